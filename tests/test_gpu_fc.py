"""GPU parity of the fully-connected block (csrc/fc_gemm.cu, modeling/fc.py) against fp64 torch restatements of
nn.Linear forward / backward (modeling/backbone/vgg16.py:122-130, sim_net.py:25-26, roi_weak_predictors.py:158-165).
TF32 products of operands that carry <= 10 mantissa bits are EXACT, so on such inputs every layout / schedule variant
must reproduce the fp32 result to accumulation-order error; generic inputs are held to the TF32 bound, and the
3xTF32 strict mode to fp32-class accuracy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from odwscl_b200 import capi as c
    c.lib()
    return c


def _q(t, s=8):
    return (t * s).round() / s


# (M, N, K): whole rounds + split-K tail, few-tile split-K path, ragged edges in every dimension, tiny, the fc7 shape
SHAPES = [(300, 357, 4096), (128, 128, 32), (513, 300, 416), (164, 4096, 1024), (4000, 128, 4096), (8000, 4096, 512),
          (1, 21, 64), (2500, 5000, 96), (37, 1000, 2500)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_fc_gemm_layouts_exact(capi, M, N, K, a_mn, b_mn):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = _q(torch.randn(M, K, generator=g))
    B = _q(torch.randn(N, K, generator=g))
    ref = (A.double() @ B.double().T).float()
    Ad = (A.T.contiguous() if a_mn else A).cuda()
    Bd = (B.T.contiguous() if b_mn else B).cuda()
    got = capi.fc_gemm(Ad, Bd, a_mn=a_mn, b_mn=b_mn).cpu()
    assert got.shape == (M, N)
    torch.testing.assert_close(got, ref, rtol=1e-6, atol=2e-5 * K ** 0.5)


def test_fc_gemm_epilogues(capi):
    g = torch.Generator().manual_seed(11)
    M, N, K = 700, 520, 320
    A, B = _q(torch.randn(M, K, generator=g)), _q(torch.randn(N, K, generator=g))
    bias = _q(torch.randn(N, generator=g))
    ref = (A.double() @ B.double().T + bias.double()).float()
    Ad, Bd, bd = A.cuda(), B.cuda(), bias.cuda()
    tol = dict(rtol=1e-6, atol=2e-5 * K ** 0.5)
    torch.testing.assert_close(capi.fc_gemm(Ad, Bd, bias=bd).cpu(), ref, **tol)
    torch.testing.assert_close(capi.fc_gemm(Ad, Bd, bias=bd, relu=True).cpu(), ref.clamp(min=0), **tol)
    # accumulate (beta = 1) into an existing tensor, incl. a column slice with its own pitch
    C0 = _q(torch.randn(M, N + 12, generator=g)).cuda()
    out = C0[:, 4:4 + N]
    before = out.clone()
    capi.fc_gemm(Ad, Bd, out=out, accumulate=True)
    torch.testing.assert_close(out.cpu(), ((A.double() @ B.double().T) + before.cpu().double()).float(), **tol)
    assert torch.equal(C0[:, :4].cpu(), C0.cpu()[:, :4]) and float((C0[:, 4 + N:] - C0[:, 4 + N:]).abs().sum()) == 0.0
    # derivative mask of the layer below
    y_prev = torch.randn(M, N, generator=g).clamp(min=0).cuda()
    got = capi.fc_gemm(Ad, Bd, mask_src=y_prev, mask_scale=2.0).cpu()
    exp = torch.where(y_prev.cpu() > 0, (A.double() @ B.double().T).float() * 2.0, torch.zeros(()))
    torch.testing.assert_close(got, exp, **tol)
    # TF32 rounding of the output: exactly cvt.rna of the unrounded result
    raw = capi.fc_gemm(Ad, Bd, bias=bd)
    rounded = capi.fc_gemm(Ad, Bd, bias=bd, round_tf32=True)
    assert torch.equal(rounded, capi.round_tf32_(raw.contiguous().clone()))
    # Dropout: zero where relu is zero, survivors scaled by 1/(1-p), keep rate 1-p, seed-reproducible
    for p in (0.5, 0.3):
        y = capi.fc_gemm(Ad, Bd, bias=bd, relu=True, dropout_p=p, seed=77).cpu()
        pos = ref > 0
        assert float(y[~pos].abs().sum()) == 0.0
        kept = y > 0
        torch.testing.assert_close(y[kept], ref[kept] * (1.0 / (1.0 - p)), rtol=1e-5, atol=1e-5 * K ** 0.5)
        rate = float(kept.sum()) / float(pos.sum())
        assert abs(rate - (1.0 - p)) < 1e-2, (p, rate)
        assert torch.equal(capi.fc_gemm(Ad, Bd, bias=bd, relu=True, dropout_p=p, seed=77).cpu(), y)
        assert not torch.equal(capi.fc_gemm(Ad, Bd, bias=bd, relu=True, dropout_p=p, seed=78).cpu() > 0, kept)


def test_fc_gemm_generic_and_strict(capi):
    """Generic fp32 inputs: single pass within the TF32 bound; 3xTF32 split at fp32-class accuracy."""
    from odwscl_b200.modeling import fc
    g = torch.Generator().manual_seed(5)
    M, N, K = 1000, 640, 25088
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.01
    ref = A.double() @ B.double().T
    got = capi.fc_gemm(A.cuda(), B.cuda()).cpu().double()
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 2e-3 * scale            # operands truncated to 10 mantissa bits
    strict = fc._gemm(A.cuda(), B.cuda(), True).cpu().double()
    err32 = float(((A @ B.T).double() - ref).abs().max())
    # K = 25088 terms accumulated with truncation in TMEM: ~1e-5 of the output scale (an MKL sgemm: ~4e-7)
    assert float((strict - ref).abs().max()) <= 2e-5 * scale, (float((strict - ref).abs().max()), err32)


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("strict", [False, True])
def test_linear_autograd_vs_torch(capi, act, strict):
    """fc.linear forward + backward (dX, dW, db) against torch's nn.functional.linear in fp64 on exactly
    representable inputs; with Dropout the mask is read off the product's own output."""
    from odwscl_b200.modeling import fc
    g = torch.Generator().manual_seed(act)
    M, K, N = 333, 192, 136
    x = _q(torch.randn(M, K, generator=g), 4).cuda().requires_grad_(True)
    w = _q(torch.randn(N, K, generator=g), 4).cuda().requires_grad_(True)
    b = _q(torch.randn(N, generator=g), 4).cuda().requires_grad_(True)
    gy = _q(torch.randn(M, N, generator=g), 4).cuda()
    p = 0.5 if act == 2 else 0.0
    y = fc.linear(x, w, b, act=act, p=p, seed=5, strict=strict)
    y.backward(gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    z = torch.nn.functional.linear(xd, wd, bd)
    if act:
        keep = (y.detach() > 0).double() * (1.0 / (1.0 - p))
        yr = torch.relu(z) * (keep if act == 2 else 1.0)      # relu: zero gradient AT zero (clamp passes it)
        if act == 2:                      # units with z > 0 that were dropped are zero in y: consistent by construction
            assert float((y.detach().double() - yr).abs().max()) <= 1e-4
    else:
        yr = z
    torch.testing.assert_close(y.detach().double(), yr.detach(), rtol=1e-6, atol=1e-4)
    yr.backward(gy.double())
    for got, ref, nm in ((x.grad, xd.grad, "dx"), (w.grad, wd.grad, "dw"), (b.grad, bd.grad, "db")):
        torch.testing.assert_close(got.double(), ref, rtol=1e-6, atol=2e-4, msg=lambda m, nm=nm: nm + ": " + m)


def test_linear_two_call_weight_gradient_fold(capi):
    """The "main" + "small" calls of one layer (fc6 twice per step): one weight / bias gradient equal to the sum."""
    from odwscl_b200.modeling import fc
    g = torch.Generator().manual_seed(9)
    K, N = 256, 160
    w = _q(torch.randn(N, K, generator=g), 4).cuda().requires_grad_(True)
    b = _q(torch.randn(N, generator=g), 4).cuda().requires_grad_(True)
    x1 = _q(torch.randn(600, K, generator=g), 4).cuda().requires_grad_(True)
    x2 = _q(torch.randn(70, K, generator=g), 4).cuda().requires_grad_(True)
    stash = {}
    y1 = fc.linear(x1, w, b, act=1, stash=stash, role="main")
    y2 = fc.linear(x2, w, b, act=1, stash=stash, role="small")
    (y1.sum() * 0.5 + (y2 * y2).sum() * 0.25).backward()
    wd, bd = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    x1d, x2d = x1.detach().double().requires_grad_(True), x2.detach().double().requires_grad_(True)
    r1 = torch.relu(torch.nn.functional.linear(x1d, wd, bd))
    r2 = torch.relu(torch.nn.functional.linear(x2d, wd, bd))
    (r1.sum() * 0.5 + (r2 * r2).sum() * 0.25).backward()
    torch.testing.assert_close(w.grad.double(), wd.grad, rtol=1e-5, atol=1e-2)
    torch.testing.assert_close(b.grad.double(), bd.grad, rtol=1e-5, atol=1e-2)
    torch.testing.assert_close(x1.grad.double(), x1d.grad, rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(x2.grad.double(), x2d.grad, rtol=1e-5, atol=1e-2)


def test_colsum(capi):
    g = torch.Generator().manual_seed(2)
    x = _q(torch.randn(5000, 357, generator=g)).cuda()
    torch.testing.assert_close(capi.colsum(x).cpu(), x.cpu().double().sum(0).float(), rtol=1e-6, atol=1e-3)
    out = torch.ones(357, device="cuda")
    capi.colsum(x[:, :357], out=out, accumulate=True)
    torch.testing.assert_close(out.cpu(), x.cpu().double().sum(0).float() + 1, rtol=1e-6, atol=1e-3)


def test_dropblock_segmented_vs_per_call(capi):
    """One segmented launch == the reference's one-DropBlock-call-per-(image, class) loop (loss.py:299): every
    segment renormalised by its own numel / sum; rows past the last offset are zero-filled."""
    from oracle import oracle as orc
    g = torch.Generator().manual_seed(4)
    seg = [0, 13, 13, 40, 77]                       # an empty segment in the middle
    R, pad = seg[-1], 9
    x = torch.randn(R + pad, 24, 7, 7, generator=g)
    cen = (torch.rand(R + pad, 7, 7, generator=g) < 0.3).float()
    off = torch.tensor(seg, dtype=torch.int32).cuda()
    y, sc = capi.dropblock_seg(x.cuda(), cen.cuda(), 1, off, len(seg) - 1)
    for p in range(len(seg) - 1):
        a, b = seg[p], seg[p + 1]
        if b > a:
            np.testing.assert_allclose(y[a:b].cpu().numpy(), orc.dropblock(x[a:b], cen[a:b], 1).numpy(), rtol=1e-6, atol=1e-7)
    assert float(y[R:].abs().sum()) == 0.0
    gy, _ = capi.dropblock_seg(x.cuda(), cen.cuda(), 1, off, len(seg) - 1, sc)
    assert torch.equal(gy, y)


def test_model_gradients_fused_activation_backward(capi):
    """Folding each ReLU/Dropout derivative into the consumer's dgrad epilogue (fc.FUSE_ACT_BWD) gives the same
    parameter gradients as the separate elementwise passes, Dropout active, on a full train step."""
    from oracle import oracle as orc
    from odwscl_b200.config import cfg
    from odwscl_b200.modeling import build_detection_model, fc
    from odwscl_b200.structures import BoxList
    model = build_detection_model(cfg)
    model.load_state_dict(orc.synth_state_dict(21, seed=0), strict=True)
    model.cuda().train()
    images, boxes, labels = orc.synth_batch(2, 150, 400, 320, 21, seed=31)
    props = [BoxList(b.cuda(), (400, 320), "xyxy") for b in boxes]
    targets = []
    for lab in labels:
        t = BoxList(torch.zeros((len(lab), 4)), (400, 320), "xyxy")
        t.add_field("labels", torch.as_tensor(lab))
        targets.append(t)
    fe = model.roi_heads.feature_extractor

    def run(fuse):
        fc.FUSE_ACT_BWD = fuse
        fc._seed_state["ctr"] = 0                       # same Philox keys in both runs
        g = torch.Generator().manual_seed(3)
        sampler = lambda n, h, w, gamma, dev: (torch.rand(n, h, w, generator=g) < gamma).float().to(dev)
        fe.dropblock.centre_sampler = sampler
        fe.sim_drop.centre_sampler = sampler
        gn = torch.Generator().manual_seed(4)
        fe.noise_sampler = lambda shape, dev: torch.randn(tuple(shape), generator=gn).to(dev)
        model.zero_grad(set_to_none=True)
        losses, _ = model(images.cuda(), targets, props)
        sum(losses.values()).backward()
        return ({k: float(v) for k, v in losses.items()},
                {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    try:
        l1, g1 = run(True)
        l0, g0 = run(False)
    finally:
        fc.FUSE_ACT_BWD = True
    assert l1 == l0
    assert set(g1) == set(g0)
    for k in g0:
        # (det_score.bias: its true gradient is 0 -- a softmax over proposals is shift-invariant -- what is left is
        # summation-order noise ~1e-10, hence the absolute floor)
        torch.testing.assert_close(g1[k], g0[k], rtol=2e-4, atol=2e-4 * float(g0[k].abs().max()) + 1e-8,
                                   msg=lambda m, k=k: k + ": " + m)


def test_fc_gemm_second_operand_pair(capi):
    """C = A B^T + A2 B2^T in one accumulation (K-concatenated operand pairs), every layout the block uses."""
    g = torch.Generator().manual_seed(8)
    M, N, K1, K2 = 300, 520, 1000, 77
    A, B = _q(torch.randn(M, K1, generator=g)), _q(torch.randn(N, K1, generator=g))
    A2, B2 = _q(torch.randn(M, K2, generator=g)), _q(torch.randn(N, K2, generator=g))
    ref = (A.double() @ B.double().T + A2.double() @ B2.double().T).float()
    for a_mn, b_mn in ((True, True), (False, False), (False, True)):
        f = lambda t, mn: (t.T.contiguous() if mn else t).cuda()
        got = capi.fc_gemm(f(A, a_mn), f(B, b_mn), a_mn=a_mn, b_mn=b_mn, A2=f(A2, a_mn), B2=f(B2, b_mn)).cpu()
        torch.testing.assert_close(got, ref, rtol=1e-6, atol=2e-5 * (K1 + K2) ** 0.5)


@pytest.mark.gpu
def test_peer_gradient_sum_two_gpus():
    """The fused weight-gradient GEMM + cross-rank sum (csrc/fc_gemm.cu kPeerSum, sharding.PeerGradSum) on 2 GPUs of one
    NVSwitch box: kernel result == all-reduced local products (bit for bit), one model step's gradients == plain DDP's
    (both forms: push to all, reduce-scatter + gather).  Needs 2 GPUs; the single-GPU round-end box skips it
    (profiles/r02_peer_sum_check_n2.txt holds the 2-GPU run of this round)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", os.path.join(root, "scripts", "peer_sum_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PEER_SUM_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
