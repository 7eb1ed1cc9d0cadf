"""Multi-GPU host logic of the hot path: images shard across ranks, the rank-local contrastive bank
stays local (modeling/roi_heads/weak_head/loss.py:276-347 only sees the rank's own targets/proposals), and the one
exchange per step is the gradient all-reduce (tools/train_net.py:50-55).  One process per GPU; `torch.distributed`
(NCCL on the GPU box, gloo in the CPU tests) is the plumbing."""
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
        out = super().forward(*a, **k)
        if self.sm_margin > 0 and torch.is_grad_enabled() and isinstance(out, tuple) and isinstance(out[0], dict):
            losses = {n: (_BackwardBegins.apply(v, self.sm_margin) if torch.is_tensor(v) and v.requires_grad else v)
                      for n, v in out[0].items()}
            out = (losses,) + tuple(out[1:])
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
        margin = int(os.environ.get("ODWSCL_SM_MARGIN", "8"))
    else:
        margin = 0
    # 611 MB of fp32 gradients per step: large buckets (NVSwitch collectives are latency-, not link-bound) and gradients
    # stored as views of the buckets
    bucket_mb = int(os.environ.get("ODWSCL_BUCKET_MB", "128"))
    return MarginedDDP(model, device_ids=ids, broadcast_buffers=False, static_graph=True, bucket_cap_mb=bucket_mb,
                       gradient_as_bucket_view=True, sm_margin=margin)


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
