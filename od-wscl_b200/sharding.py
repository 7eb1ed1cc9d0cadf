"""Multi-GPU host logic of the hot path: images shard across ranks, the rank-local contrastive bank
stays local (modeling/roi_heads/weak_head/loss.py:276-347 only sees the rank's own targets/proposals), and the one
exchange per step is the gradient sum (tools/train_net.py:50-55): DistributedDataParallel's bucket all-reduce, with the
big fully-connected weight gradients optionally summed across ranks by the weight-gradient GEMM itself (PeerGradSum).
One process per GPU; `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) is the plumbing."""
import os

import torch
import torch.distributed as dist


def env_world():
    """(world_size, rank, local_rank) as torchrun exports them."""
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def images_per_gpu(ims_per_batch, world):
    """data/build.py:150-155: the global batch must divide evenly over the ranks."""
    if ims_per_batch % world != 0:
        raise ValueError("SOLVER.IMS_PER_BATCH (%d) must be divisible by the number of GPUs (%d) used."
                         % (ims_per_batch, world))
    return ims_per_batch // world


def shard_image_ids(ims_per_batch, world, rank):
    """Global image indices of one step that rank `rank` owns (contiguous block, as the reference's
    DistributedSampler + BatchSampler hand them out per iteration)."""
    per = images_per_gpu(ims_per_batch, world)
    return list(range(rank * per, (rank + 1) * per))


def rank_seed(base, rank, images_per_rank):
    """synth_batch seeds image i of a rank with seed+i: give every rank a disjoint seed range so the global batch
    of a step is the same set of images for any world size."""
    return base + rank * images_per_rank


PEER_PUSH_ALL_MAX = 2         # largest world whose fc gradients are pushed to every replica from the GEMM epilogue


def plan(world):
    """(peer gradient sum on?, SMs left to NCCL during the backward) for a world size, from the round-2 measurements
    (profiles/r02_scaling.md; ms/step at configs[1], same-box 1-GPU step 16.8-17.3):
        2 GPUs   peer sum (push to all) + margin 8: 17.6-17.9     all-reduce + margin 16: 18.2-18.5
        4 GPUs   peer sum (push to all) + margin 8: 18.6          all-reduce + margin 16: 18.3
        8 GPUs   peer sum (reduce-scatter) + margin 8: 19.0       all-reduce + margin 16: 18.5-18.7     push to all: 20.8-21.3
    ODWSCL_PEER_SUM=0|1 and ODWSCL_SM_MARGIN override."""
    env = os.environ.get("ODWSCL_PEER_SUM", "auto")
    peer = env == "1" or (env == "auto" and world == 2)
    margin = int(os.environ.get("ODWSCL_SM_MARGIN", "8" if peer else "16"))
    return peer, margin


def configure_nccl():
    """Call before init_process_group: NCCL may use at most as many CTAs as the SMs the persistent conv / fc kernels leave
    free during the backward, so an all-reduce kernel always finds room beside them and never holds a persistent grid's
    last CTAs back (see wrap_ddp).  An explicit NCCL_MAX_CTAS wins."""
    _, margin = plan(int(os.environ.get("WORLD_SIZE", "1")))
    if margin > 0:
        os.environ.setdefault("NCCL_MAX_CTAS", str(margin))


class _BackwardBegins(torch.autograd.Function):
    """Identity on a loss tensor whose backward runs FIRST in the backward pass and switches the SM margin on: the NCCL
    gradient all-reduce only runs beside the backward kernels, so the forward keeps every SM."""

    @staticmethod
    def forward(ctx, x, margin):
        ctx.margin = margin
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        from . import capi
        capi.set_sm_margin(ctx.margin)
        return g, None


class MarginedDDP(torch.nn.parallel.DistributedDataParallel):
    """DDP whose forward runs with all SMs (margin 0) and whose backward leaves `sm_margin` SMs to NCCL."""

    def __init__(self, *a, sm_margin=0, **k):
        super().__init__(*a, **k)
        self.sm_margin = sm_margin

    def forward(self, *a, **k):
        from . import capi
        capi.set_sm_margin(0)
        if getattr(self, "peer", None) is not None and not getattr(self.peer, "hooked", False) and torch.is_grad_enabled():
            raise RuntimeError("wrap_ddp(): call sharding.hook_optimizer(model, optimizer) before training -- the fc weight "
                               "gradients are summed across ranks by hooks around optimizer.step()")
        out = super().forward(*a, **k)
        if self.sm_margin > 0 and torch.is_grad_enabled() and isinstance(out, tuple) and isinstance(out[0], dict):
            losses = {n: (_BackwardBegins.apply(v, self.sm_margin) if torch.is_tensor(v) and v.requires_grad else v)
                      for n, v in out[0].items()}
            out = (losses,) + tuple(out[1:])
        return out


class _PeerTarget:
    """One weight's slice of the symmetric gradient buffer (fc._LinearFn.backward talks to this)."""

    def __init__(self, owner, param, view, mc_ptr, peer_ptrs):
        import weakref
        self.owner, self.param, self.view, self.mc_ptr, self.peer_ptrs = owner, weakref.ref(param), view, mc_ptr, peer_ptrs
        rows = view.shape[0]
        self.rows_per_owner = ((rows + owner.world - 1) // owner.world + 31) // 32 * 32
        r0 = min(rows, owner.rank * self.rows_per_owner)
        r1 = min(rows, r0 + self.rows_per_owner) if owner.rank < owner.world - 1 else rows
        self.own = (r0, r1)                                    # the rows this rank sums and later broadcasts

    def sum_product(self, dz, x, pending):
        from . import capi
        a2, b2 = pending if pending is not None else (None, None)
        if self.owner.scatter:
            capi.fc_gemm_peer_sum(dz, x, self.view, 0, self.owner.scale, a_mn=True, b_mn=True, A2=a2, B2=b2,
                                  peer_ptrs=self.peer_ptrs, rows_per_owner=self.rows_per_owner)
        else:
            capi.fc_gemm_peer_sum(dz, x, self.view, self.mc_ptr, self.owner.scale, a_mn=True, b_mn=True, A2=a2, B2=b2)

    def gather(self):
        """reduce-scatter form: this rank's rows hold the complete sum -- replicate them to every rank."""
        from . import capi
        r0, r1 = self.own
        if r1 > r0:
            capi.peer_broadcast(self.view[r0:r1], self.mc_ptr + 4 * r0 * self.view.shape[1])

    def add(self, local_grad):
        from . import capi
        capi.peer_add(local_grad.contiguous(), self.mc_ptr, self.owner.scale)

    def grad(self):
        """What backward returns for the weight: the buffer slice once per step (autograd adopts it as .grad without a copy:
        a fresh view object, so nothing else references it), None when .grad already is that memory."""
        p = self.param()
        g = None if p is None else p.grad
        if (g is not None and g.data_ptr() == self.view.data_ptr()) or self.owner.handed.get(id(self)):
            return None
        self.owner.handed[id(self)] = True
        return self.view.view_as(self.view)


class PeerGradSum:
    """Cross-rank gradient sum of the big fully-connected weights WITHOUT an all-reduce: the gradients live in one
    symmetric-memory buffer (same offset on every rank, one NVSwitch multicast address), the weight-gradient GEMM adds its
    tiles to all replicas from its epilogue (csrc/fc_gemm.cu kPeerSum, multimem.red), and two rank barriers per step --
    before the optimizer reads, after it has zeroed -- are the only synchronisation.  fc6 + fc7 + Sim_Net are 536 of the
    611 MB a step all-reduces; what is left goes through DistributedDataParallel as before.
    Gradients of these weights are complete only after `before_step()` (the optimizer hooks call it), and they are
    cleared by `after_step()`, not by `optimizer.zero_grad()`: a backward whose gradients are discarded without an
    optimizer step must be followed by `after_step()` on every rank.

    Two forms.  world <= PEER_PUSH_ALL_MAX: every tile goes to EVERY replica (multimem.red, one NVLink operation replicated
    by the switch) and nothing else is needed -- measured at 2 GPUs: 17.6 ms/step against 18.2-18.5 with the all-reduce.
    A rank then takes in (world - 1) gradients per step, which at 8 GPUs (3.75 GB) saturates its NVLink ingress (21 ms/step
    measured).  Larger worlds (only with ODWSCL_PEER_SUM=1: plan() keeps them on the all-reduce, which measured 0.3 ms
    faster) reduce-scatter from the epilogue (each 32-row block of the gradient is added to its owner rank only: red.add
    over NVLink to the peer-mapped replica) and all-gather the owned rows by multicast store in before_step():
    (world - 1) / world of the gradient in and out per rank, the gather (0.6 ms at 8 GPUs) exposed."""

    def __init__(self, params, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        mode = os.environ.get("ODWSCL_PEER_MODE", "auto")      # all | scatter | auto
        self.scatter = mode == "scatter" or (mode == "auto" and self.world > PEER_PUSH_ALL_MAX)
        if self.world > 8:
            raise RuntimeError("peer gradient sum: one NVSwitch domain of at most 8 ranks")
        self.scale = 1.0 / self.world                          # DistributedDataParallel averages
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]
        self.buf = symm_mem.empty(sum(sizes), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        mc = int(self.hdl.multicast_ptr)
        if mc == 0:
            raise RuntimeError("symmetric memory without a multicast address (no NVSwitch / NVLS on this box)")
        self.handed, self.targets, off = {}, [], 0
        for p, n in zip(params, sizes):
            if p.shape[1] % 4:
                raise RuntimeError("peer gradient sum: in_features must be a multiple of 4")
            p._odw_peer = _PeerTarget(self, p, self.buf[off:off + p.numel()].view_as(p), mc + 4 * off,
                                      [int(b) + 4 * off for b in self.hdl.buffer_ptrs])
            self.targets.append(p._odw_peer)
            off += n
        self.buf.zero_()
        self.hdl.barrier(channel=0)

    def before_step(self):
        """Every rank's weight-gradient GEMMs of this step have finished (stream-ordered on each rank, then the barrier)."""
        self.hdl.barrier(channel=0)
        if self.scatter:
            for t in self.targets:
                t.gather()
            self.hdl.barrier(channel=0)

    def after_step(self):
        """The optimizer has consumed the sums: zero the replica, and let nobody start the next backward before every
        replica is zero."""
        self.buf.zero_()
        self.handed.clear()
        self.hdl.barrier(channel=1)

    def hook_optimizer(self, optimizer):
        optimizer.register_step_pre_hook(lambda *_: self.before_step())
        optimizer.register_step_post_hook(lambda *_: self.after_step())
        self.hooked = True


PEER_MIN_NUMEL = 4 << 20      # weights at least this large leave DDP's buckets (fc6 102.8 M, fc7 / Sim_Net 16.8 M each)


def _peer_candidates(model):
    """(name, parameter) of the 2-D weights fc.linear consumes directly (modules that route through modeling/fc.py)."""
    out = []
    for name, p in model.named_parameters():
        if p.requires_grad and p.dim() == 2 and p.numel() >= PEER_MIN_NUMEL:
            out.append((name, p))
    return out


def wrap_ddp(model, device=None):
    """DDP exactly as the reference wraps it (tools/train_net.py:50-55: broadcast_buffers=False), with
    static_graph instead of find_unused_parameters -- every parameter receives a gradient on this path."""
    ids = [device.index] if device is not None and device.type == "cuda" else None
    if ids is not None:
        from . import capi
        # SMs left to the NCCL all-reduce kernels.  The persistent conv / fc kernels otherwise occupy every SM (1 CTA per SM,
        # ~200 KB of shared memory each, so an NCCL CTA cannot co-reside); NCCL then grabs SMs between two of our launches
        # and the NEXT persistent grid no longer fits: its last CTAs -- and the split-K slices spinning on them -- wait for
        # the 411 MB fc6 bucket to finish.  Measured at 2 GPUs: steps of 37 / 67 / 131 ms among 17.8 ms ones with margin 0,
        # a flat 18.2-19.1 ms with margin 8 (profiles/r02_scaling.md).
        use_peer, margin = plan(dist.get_world_size())
    else:
        use_peer, margin = False, 0
    # 611 MB of fp32 gradients per step: large buckets (NVSwitch collectives are latency-, not link-bound) and gradients
    # stored as views of the buckets
    bucket_mb = int(os.environ.get("ODWSCL_BUCKET_MB", "128"))
    peer = None
    if use_peer and dist.get_world_size() > 1:
        cand = _peer_candidates(model)
        try:
            peer = PeerGradSum([p for _, p in cand], device)
            torch.nn.parallel.DistributedDataParallel._set_params_and_buffers_to_ignore_for_model(model, [n for n, _ in cand])
            for _, p in cand:                                   # DDP's initial rank-0 broadcast skips ignored parameters
                dist.broadcast(p.data, src=0)
        except Exception as e:                                  # no NVLS: everything stays on the bucket all-reduce
            if dist.get_rank() == 0:
                print("[od-wscl_b200] peer gradient sum unavailable (%s); all gradients use the DDP all-reduce" % (e,), flush=True)
            for _, p in cand:
                if hasattr(p, "_odw_peer"):
                    del p._odw_peer
            model._ddp_params_and_buffers_to_ignore = []
            peer = None
    ddp = MarginedDDP(model, device_ids=ids, broadcast_buffers=False, static_graph=True, bucket_cap_mb=bucket_mb,
                      gradient_as_bucket_view=True, sm_margin=margin)
    ddp.peer = peer
    return ddp


def hook_optimizer(ddp_model, optimizer):
    """Must be called once with the optimizer that steps a wrap_ddp() model: the peer-summed gradients are synchronised by
    hooks around optimizer.step()."""
    peer = getattr(ddp_model, "peer", None)
    if peer is not None:
        peer.hook_optimizer(optimizer)


def max_over_ranks(value, device):
    """Bench timing rule: the slowest rank's device time is the step time."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ranks(value, device):
    """Every rank's value (diagnosis: which part of a multi-GPU step time is the slowest GPU's clock)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return [float(value)]
    t = torch.tensor([float(value)], device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def proposals_per_step(world, images_per_rank, proposals_per_image):
    """Units all ranks process in one step (weak scaling: per-rank work is fixed)."""
    return world * images_per_rank * proposals_per_image
