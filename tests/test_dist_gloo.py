"""N>1 host logic on CPU: world_size-2 gloo processes.  The path shards by image with no collective
in the forward math; the one exchange is the gradient all-reduce (mean), checked here against a single-process
run over the union of the shards."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _CpuHead(torch.nn.Module):
    """CPU stand-in with MISTPredictor's parameter set and outputs (the product module runs only on the sm_100a fc
    kernel and refuses CPU tensors); what is under test here is the sharding / all-reduce host logic."""

    def __init__(self, in_channels, nc=21):
        super().__init__()
        for n, o in (("cls_score", nc), ("det_score", nc), ("ref1", nc), ("bbox_pred1", 4 * nc), ("ref2", nc),
                     ("bbox_pred2", 4 * nc), ("ref3", nc), ("bbox_pred3", 4 * nc)):
            setattr(self, n, torch.nn.Linear(in_channels, o))

    def forward(self, x, proposals):
        return (self.cls_score(x), self.det_score(x), [self.ref1(x), self.ref2(x), self.ref3(x)],
                [self.bbox_pred1(x), self.bbox_pred2(x), self.bbox_pred3(x)])


def _make_head():
    torch.manual_seed(0)
    m = _CpuHead(64)
    for p in m.parameters():                      # the reference init (std 1e-3) makes every gradient tiny
        torch.nn.init.normal_(p, std=0.05)
    return m


def _shard_inputs(image_ids, n=48):
    from odwscl_b200.synth import synth_batch
    xs, props = [], []
    for i in image_ids:
        _, _, boxes, _ = synth_batch(1, n, 320, 256, seed=1234 + i)
        g = torch.Generator().manual_seed(77 + i)
        xs.append(torch.randn(n, 64, generator=g))
        props.append(boxes[0])
    return torch.cat(xs), props


def _loss(model, x, props):
    cls, det, refs, bbs = model(x, props)
    per_img = []
    for c, d in zip(cls.split([len(p) for p in props]), det.split([len(p) for p in props])):
        s = (torch.softmax(c, 1) * torch.softmax(d, 0)).sum(0).clamp(1e-6, 1 - 1e-6)     # loss.py:349-354 (MIL)
        per_img.append(-(torch.log(1 - s)).mean())
    return torch.stack(per_img).mean() + sum(r.square().mean() for r in refs) + sum(b.square().mean() for b in bbs)


def _worker(rank, world, port, ims_per_batch, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    from odwscl_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.env_world() == (world, rank, rank)
        ids = sharding.shard_image_ids(ims_per_batch, world, rank)
        model = _make_head()
        ddp = sharding.wrap_ddp(model)
        x, props = _shard_inputs(ids)
        _loss(ddp, x, props).backward()
        grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        slow = sharding.max_over_ranks(10.0 + rank, torch.device("cpu"))
        if rank == 0:
            torch.save({"grads": grads, "ids": ids, "slow": slow}, out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from odwscl_b200 import sharding
    world, ims = 2, 4
    out = str(tmp_path / "rank0.pt")
    mp.start_processes(_worker, args=(world, _free_port(), ims, out), nprocs=world, join=True, start_method="spawn")
    got = torch.load(out)
    assert got["ids"] == [0, 1] and got["slow"] == 11.0
    # single process: mean over ranks of the per-rank losses == DDP's averaged gradient
    model = _make_head()
    total = 0
    for r in range(world):
        x, props = _shard_inputs(sharding.shard_image_ids(ims, world, r))
        total = total + _loss(model, x, props) / world
    total.backward()
    for k, p in model.named_parameters():
        torch.testing.assert_close(got["grads"][k], p.grad, rtol=1e-5, atol=1e-7)


def test_sharding_rules():
    from odwscl_b200 import sharding
    assert sharding.images_per_gpu(8, 4) == 2
    with pytest.raises(ValueError):
        sharding.images_per_gpu(8, 3)                      # data/build.py:150-155
    ids = [sharding.shard_image_ids(8, 4, r) for r in range(4)]
    assert sorted(sum(ids, [])) == list(range(8))
    assert sharding.rank_seed(1234, 3, 2) == 1240
    assert sharding.proposals_per_step(8, 2, 2000) == 32000


def test_multi_gpu_plan(monkeypatch):
    """Which gradient exchange a world size gets (measured choice, profiles/r02_scaling.md) and the NCCL CTA cap that goes
    with the SM margin; the peer-sum targets split a weight's rows into 32-row-aligned owner blocks that cover it."""
    from odwscl_b200 import sharding
    for k in ("ODWSCL_PEER_SUM", "ODWSCL_SM_MARGIN", "NCCL_MAX_CTAS"):
        monkeypatch.delenv(k, raising=False)
    assert sharding.plan(2) == (True, 8)
    assert sharding.plan(4) == (False, 16) and sharding.plan(8) == (False, 16)
    monkeypatch.setenv("ODWSCL_PEER_SUM", "1")
    assert sharding.plan(8) == (True, 8)
    monkeypatch.setenv("ODWSCL_SM_MARGIN", "4")
    assert sharding.plan(8) == (True, 4)
    monkeypatch.delenv("ODWSCL_PEER_SUM")
    monkeypatch.delenv("ODWSCL_SM_MARGIN")
    monkeypatch.setenv("WORLD_SIZE", "8")
    sharding.configure_nccl()
    import os
    assert os.environ["NCCL_MAX_CTAS"] == "16"
    monkeypatch.setenv("NCCL_MAX_CTAS", "24")              # an explicit setting wins
    sharding.configure_nccl()
    assert os.environ["NCCL_MAX_CTAS"] == "24"

    class _Owner:
        def __init__(self, world, rank):
            self.world, self.rank, self.scatter, self.scale = world, rank, True, 1.0 / world
    import torch
    for rows in (4096, 300, 31, 128):
        for world in (2, 3, 8):
            p = torch.nn.Parameter(torch.zeros(rows, 8))
            spans = [sharding._PeerTarget(_Owner(world, r), p, p.data, 0, [0] * world) for r in range(world)]
            rpo = spans[0].rows_per_owner
            assert rpo % 32 == 0 and rpo * world >= rows
            covered = [r for t in spans for r in range(*t.own)]
            assert covered == list(range(rows))             # disjoint, in order, complete
            for t in spans:                                  # the kernel's rule: owner = min(row0 // rpo, world - 1)
                for r in range(*t.own):
                    assert min((r // 32 * 32) // rpo, world - 1) == t.owner.rank


def test_peer_target_hands_the_gradient_view_once_per_step():
    """fc._LinearFn.backward returns `_PeerTarget.grad()` for a peer-summed weight: the buffer slice the first time in a
    step (autograd adopts it as .grad), None for a second gradient of the same weight in that step and whenever .grad
    already IS the buffer (otherwise autograd would add the buffer to itself); after_step() re-arms it."""
    import torch
    from odwscl_b200 import sharding

    class _Owner:
        world, rank, scatter, scale = 2, 0, False, 0.5

        def __init__(self):
            self.handed = {}
    owner = _Owner()
    buf = torch.zeros(64 * 8)
    p = torch.nn.Parameter(torch.zeros(64, 8))
    t = sharding._PeerTarget(owner, p, buf.view(64, 8), 0, [0, 0])
    g1 = t.grad()
    assert g1 is not None and g1.data_ptr() == buf.data_ptr() and g1 is not t.view     # a fresh view object of the buffer
    assert t.grad() is None                                                            # second product of the step
    p.grad = g1
    owner.handed.clear()                                                               # what after_step() does
    assert t.grad() is None                                                            # .grad already is the buffer
    p.grad = None                                                                      # zero_grad(set_to_none=True)
    g2 = t.grad()
    assert g2 is not None and g2.data_ptr() == buf.data_ptr()
