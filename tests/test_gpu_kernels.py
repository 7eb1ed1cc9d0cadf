"""GPU parity tests proper: the sm_100a kernels, called through the C ABI (odwscl_b200.capi ->
libodwscl_sm100.so), against the CPU oracle (oracle/) and the reference-generated golden vectors
(tests/golden/).  Bit-exact for index / selection work; stated tolerances for floating point."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from odwscl_b200 import capi as c
    c.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return c


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ------------------------------------------------------------------------------- ROIPool
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_roi_pool_golden(capi, golden, tag):
    G = golden("roi_pool.npz")
    ph, pw = [int(x) for x in G[tag + "_pooled"]]
    out, arg = capi.roi_pool_forward(cu(G[tag + "_feat"]), cu(G[tag + "_rois"]), 0.125, ph, pw)
    assert np.array_equal(out.cpu().numpy(), G[tag + "_out"])
    assert np.array_equal(arg.cpu().numpy(), G[tag + "_argmax"])
    B, C, H, W = G[tag + "_feat"].shape
    gi = capi.roi_pool_backward(cu(G[tag + "_grad_out"]), cu(G[tag + "_rois"]), arg, ph, pw, B, C, H, W)
    np.testing.assert_allclose(gi.cpu().numpy(), G[tag + "_grad_in"], rtol=1e-5, atol=1e-5)


def _rand_rois(g, B, H, W, n, scale=0.125, wild=True):
    iw, ih = W / scale, H / scale
    x1 = torch.rand(n, generator=g) * iw * (1.2 if wild else 0.9) - (0.1 * iw if wild else 0)
    y1 = torch.rand(n, generator=g) * ih * (1.2 if wild else 0.9) - (0.1 * ih if wild else 0)
    w = torch.rand(n, generator=g) * iw * 0.8 - (0.05 * iw if wild else 0)
    h = torch.rand(n, generator=g) * ih * 0.8 - (0.05 * ih if wild else 0)
    b = torch.randint(0, B, (n,), generator=g).float()
    return torch.stack([b, x1, y1, x1 + w, y1 + h], 1)


@pytest.mark.parametrize("B,C,H,W,R,quant", [(2, 8, 19, 27, 64, False), (1, 128, 38, 50, 97, True),
                                             (2, 132, 20, 31, 50, False), (1, 512, 12, 17, 33, True)])
def test_roi_pool_fast_path_vs_oracle(capi, B, C, H, W, R, quant):
    """7x7, C % 4 == 0 -> channels-last warp-per-bin-row kernel; bit-exact incl. ties / empty bins."""
    g = torch.Generator().manual_seed(B * 1000 + C)
    feat = torch.randn(B, C, H, W, generator=g)
    if quant:
        feat = (feat * 2).round() / 2
    rois = _rand_rois(g, B, H, W, R)
    eo, ea = orc.roi_pool_forward(feat.numpy(), rois.numpy(), 0.125, 7, 7)
    out, arg = capi.roi_pool_forward(feat.cuda(), rois.cuda(), 0.125, 7, 7)
    assert np.array_equal(out.cpu().numpy(), eo)
    assert np.array_equal(arg.cpu().numpy(), ea)
    go = torch.randn(out.shape, generator=g)
    gi = capi.roi_pool_backward(go.cuda(), rois.cuda(), arg, 7, 7, B, C, H, W)
    egi = orc.roi_pool_backward(go.numpy(), ea, rois.numpy(), B, C, H, W)
    np.testing.assert_allclose(gi.cpu().numpy(), egi, rtol=1e-4, atol=1e-4)


def test_roi_pool_empty_and_errors(capi):
    feat = torch.randn(1, 8, 5, 5).cuda()
    out, arg = capi.roi_pool_forward(feat, torch.zeros((0, 5)).cuda(), 0.125, 7, 7)
    assert out.shape == (0, 8, 7, 7) and arg.shape == (0, 8, 7, 7)
    gi = capi.roi_pool_backward(out, torch.zeros((0, 5)).cuda(), arg, 7, 7, 1, 8, 5, 5)
    assert float(gi.abs().sum()) == 0.0
    from odwscl_b200 import _C
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):      # csrc/ROIPool.h:23
        _C.roi_pool_forward(feat.cpu(), torch.zeros((1, 5)), 0.125, 7, 7)


def test_roi_pool_full_size_properties(capi):
    """BASELINE configs[1] shape: 2 x 512 x 76 x 128 map, 2000 rois / image.  Size-independent
    properties: out == feat[argmax]; argmax inside the roi's clipped extent; backward conserves
    mass; agreement with torchvision's kernel (same Caffe2 lineage) on the GPU."""
    import torchvision
    B, C, H, W, N = 2, 512, 76, 128, 2000
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(B, C, H, W, generator=g).cuda()
    rois = torch.cat([torch.cat([torch.full((N, 1), float(b)), orc.synth_boxes(N, 1000, 600, g)], 1)
                      for b in range(B)]).cuda()
    out, arg = capi.roi_pool_forward(feat, rois, 0.125, 7, 7)
    tv_out, tv_arg = torch.ops.torchvision.roi_pool(feat, rois, 0.125, 7, 7)
    assert torch.equal(out, tv_out)
    assert torch.equal(arg, tv_arg.int())
    bidx = rois[:, 0].long()
    planes = feat.view(B, C, H * W)[bidx]                       # [R,C,HW]
    valid = arg >= 0
    gathered = torch.gather(planes, 2, arg.clamp(min=0).long().view(B * N, C, 49)).view_as(out)
    assert torch.equal(gathered[valid], out[valid])
    assert float(out[~valid].abs().sum()) == 0.0
    go = torch.randn(out.shape, generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")
    gi = capi.roi_pool_backward(go, rois, arg, 7, 7, B, C, H, W)
    tot_in, tot_out = float(gi.double().sum()), float((go * valid).double().sum())
    assert abs(tot_in - tot_out) <= 1e-3 * max(1.0, abs(tot_out))
    # linearity of the backward in grad_out
    gi2 = capi.roi_pool_backward(2.0 * go, rois, arg, 7, 7, B, C, H, W)
    torch.testing.assert_close(gi2, 2.0 * gi, rtol=1e-4, atol=1e-3)
    # the reference's own scatter-add arithmetic (torchvision kernel, same lineage) at full size
    tv_gi = torch.ops.torchvision._roi_pool_backward(go, rois, tv_arg, 0.125, 7, 7, B, C, H, W)
    torch.testing.assert_close(gi, tv_gi, rtol=1e-4, atol=1e-3)
    # channels-last entry points (what the conv stack feeds): identical forward, same backward
    feat_cl = feat.contiguous(memory_format=torch.channels_last)
    out_cl, arg_cl = capi.roi_pool_forward(feat_cl, rois, 0.125, 7, 7)
    assert torch.equal(out_cl, out) and torch.equal(arg_cl, arg)
    gi_cl = capi.roi_pool_backward(go, rois, arg, 7, 7, B, C, H, W, channels_last=True)
    assert gi_cl.shape == gi.shape
    torch.testing.assert_close(gi_cl.contiguous(), gi, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("B,C,H,W,R", [(2, 6, 152, 200, 300), (1, 5, 160, 250, 64), (3, 16, 30, 40, 1500)])
def test_roi_pool_bwd_plane_variants(capi, B, C, H, W, R):
    """Backward plane kernel with 2 / 1 channels per CTA (large maps, odd C) and several roi chunks per image."""
    g = torch.Generator().manual_seed(C * 7 + R)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = _rand_rois(g, B, H, W, R)
    out, arg = capi.roi_pool_forward(feat.cuda(), rois.cuda(), 0.125, 7, 7)
    eo, ea = orc.roi_pool_forward(feat.numpy(), rois.numpy(), 0.125, 7, 7)
    assert np.array_equal(out.cpu().numpy(), eo) and np.array_equal(arg.cpu().numpy(), ea)
    go = torch.randint(-8, 9, out.shape, generator=g).float() / 4        # exactly representable sums: bit-exact
    gi = capi.roi_pool_backward(go.cuda(), rois.cuda(), arg, 7, 7, B, C, H, W)
    egi = orc.roi_pool_backward(go.numpy(), ea, rois.numpy(), B, C, H, W)
    assert np.array_equal(gi.cpu().numpy(), egi)


# ------------------------------------------------------------------------------- ROIAlign
@pytest.mark.parametrize("tag", ["a", "b"])
def test_roi_align_golden(capi, golden, tag):
    G = golden("roi_align.npz")
    sr = int(G[tag + "_sr"])
    out = capi.roi_align_forward(cu(G[tag + "_feat"]), cu(G[tag + "_rois"]), 0.125, 7, 7, sr)
    np.testing.assert_allclose(out.cpu().numpy(), G[tag + "_out"], rtol=1e-5, atol=1e-5)
    B, C, H, W = G[tag + "_feat"].shape
    gi = capi.roi_align_backward(cu(G[tag + "_grad_out"]), cu(G[tag + "_rois"]), 0.125, 7, 7, B, C, H, W, sr)
    np.testing.assert_allclose(gi.cpu().numpy(), G[tag + "_grad_in"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_roi_align_golden_channels_last(capi, golden, tag):
    """The same reference-generated vectors through the channels-last kernels (NHWC forward, plane-centric backward)."""
    G = golden("roi_align.npz")
    sr = int(G[tag + "_sr"])
    feat = cu(G[tag + "_feat"])
    B, C, H, W = feat.shape
    if C % 4:
        pytest.skip("channels-last path needs C % 4 == 0")
    out = capi.roi_align_forward(feat.contiguous(memory_format=torch.channels_last), cu(G[tag + "_rois"]), 0.125, 7, 7, sr)
    np.testing.assert_allclose(out.cpu().numpy(), G[tag + "_out"], rtol=1e-5, atol=1e-5)
    gi = capi.roi_align_backward(cu(G[tag + "_grad_out"]), cu(G[tag + "_rois"]), 0.125, 7, 7, B, C, H, W, sr, channels_last=True)
    np.testing.assert_allclose(gi.contiguous().cpu().numpy(), G[tag + "_grad_in"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,C,H,W,R,sr", [(2, 64, 38, 50, 300, 0), (1, 128, 76, 128, 120, 2), (2, 8, 20, 31, 64, 0)])
def test_roi_align_channels_last_vs_oracle(capi, B, C, H, W, R, sr):
    """Larger random cases (adaptive sampling grids on big rois, rois partly outside the map, sampling_ratio 0 and 2)
    against the CPU oracle (csrc/cuda/ROIAlign_cuda.cu:64-122,177-254 restated in oracle/odwscl_oracle.c), both layouts."""
    g = torch.Generator().manual_seed(R + sr)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = _rand_rois(g, B, H, W, R)
    go = torch.randn(R, C, 7, 7, generator=g)
    exp = orc.roi_align_forward(feat.numpy(), rois.numpy(), 0.125, 7, 7, sr)
    exp_gi = orc.roi_align_backward(go.numpy(), rois.numpy(), 0.125, B, C, H, W, sr)
    scale = float(np.abs(exp_gi).max())
    for cl in (True, False):
        f = feat.cuda().contiguous(memory_format=torch.channels_last) if cl else feat.cuda()
        out = capi.roi_align_forward(f, rois.cuda(), 0.125, 7, 7, sr)
        # 16 products per bin summed in fp32: nvcc contracts them into FMAs, gcc -O2 does not, so a bin whose samples nearly
        # cancel differs by a few ulp of the LARGEST term (|feat| <= ~5), not of the result: atol 5e-5
        np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=1e-5, atol=5e-5)
        gi = capi.roi_align_backward(go.cuda(), rois.cuda(), 0.125, 7, 7, B, C, H, W, sr, channels_last=cl)
        np.testing.assert_allclose(gi.contiguous().cpu().numpy(), exp_gi, rtol=1e-4, atol=2e-5 * scale + 1e-6)


# ------------------------------------------------------------------------------- IoU / NMS
def test_iou_nms_golden(capi, golden):
    G = golden("boxes.npz")
    P, Q, S = cu(G["P"]), cu(G["Q"]), cu(G["scores"])
    assert np.array_equal(capi.box_iou(P, Q, True).cpu().numpy(), G["iou"])
    assert np.array_equal(capi.nms(P, S, 0.3).cpu().numpy(), G["tv_nms_full"])
    cl = torch.from_numpy(G["cluster"]).cuda()
    for t in range(3):
        thr = float(G["easy_nms_thr_%d" % t])
        keep = cl[capi.nms(P[cl], S[cl], thr)]
        assert np.array_equal(keep.cpu().numpy(), G["easy_nms_%d" % t])
    # legacy `_C.nms`: CUDA rule is '>' (csrc/cuda/nms.cu:60); compare with the oracle's ge=False
    Su = cu(G["scores_u"])
    for thr in (0.1, 0.3, 0.7):
        assert np.array_equal(capi.nms_legacy(P, Su, thr).cpu().numpy(), orc.nms_legacy(G["P"], G["scores_u"], thr, ge=False))


@pytest.mark.parametrize("n,ties", [(1, False), (31, True), (700, True), (2000, False), (4000, True), (8192, True),
                                    (8193, True), (20000, True)])       # > 8192: the three-launch large path
def test_nms_random_vs_oracle(capi, n, ties):
    g = torch.Generator().manual_seed(n)
    P = orc.synth_boxes(n, 1000, 600, g)
    if n > 40:
        P[n // 2: n // 2 + 10] = P[:10]                       # exact duplicates
    s = torch.rand(n, generator=g)
    if ties:
        s = (s * 16).round() / 16
    for thr in (0.1, 0.5):
        got = capi.nms(P.cuda(), s.cuda(), thr).cpu().numpy()
        assert np.array_equal(got, orc.nms_tv(P.numpy(), s.numpy(), thr))
    assert np.array_equal(capi.box_iou(P.cuda(), P[:7].cuda(), False).cpu().numpy(),
                          orc.box_iou(P.numpy(), P[:7].numpy(), False))


def test_nms_large_legacy_and_per_class(capi):
    """Beyond the single-CTA limit: `_C.nms` semantics (ascending output) and the all-class test-time filter."""
    g = torch.Generator().manual_seed(5)
    n = 9000
    P = orc.synth_boxes(n, 1000, 600, g)
    s = (torch.rand(n, generator=g) * 64).round() / 64
    got = capi.nms_legacy(P.cuda(), s.cuda(), 0.3).cpu().numpy()
    assert np.array_equal(got, orc.nms_legacy(P.numpy(), s.numpy(), 0.3, False))
    C = 3
    boxes = torch.cat([P, P + 3.0, P.flip(0)], dim=1).contiguous()                  # [n, C*4]
    scores = torch.stack([torch.rand(n, generator=g), s, (torch.rand(n, generator=g) * 8).round() / 8], dim=1).contiguous()
    keep, cnt = capi.nms_per_class(boxes.cuda(), scores.cuda(), 0.25, 0.4)
    keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    assert cnt[0] == 0
    for j in range(1, C):
        cand = np.nonzero(scores[:, j].numpy() > np.float32(0.25))[0]
        exp = cand[orc.nms_tv(boxes[:, 4 * j:4 * j + 4].numpy()[cand], scores[:, j].numpy()[cand], 0.4)]
        assert np.array_equal(keep[j, :cnt[j]], exp), j


def test_nms_empty(capi):
    assert capi.nms(torch.zeros((0, 4)).cuda(), torch.zeros((0,)).cuda(), 0.5).numel() == 0


# ------------------------------------------------------------------------------- SupCon
@pytest.fixture(params=["tiles", "tensor"])
def supcon_path(request, monkeypatch):
    """Both SupCon implementations: the fused FFMA tile kernels (csrc/supcon.cu) and the tcgen05 3xTF32 path
    (csrc/supcon_tc.cu), which the host picks for banks of >= SUPCON_TC_MIN_ROWS rows."""
    from odwscl_b200.modeling import sim_head
    monkeypatch.setattr(sim_head, "SUPCON_TC_MIN_ROWS", 0 if request.param == "tensor" else 1 << 30)
    return request.param


@pytest.mark.parametrize("tag", ["a", "b"])
def test_supcon_golden(capi, golden, tag, supcon_path):
    """fp32 loss within 1e-4 rel (north_star tolerance), gradient within 2e-4 rel."""
    from odwscl_b200.modeling.sim_head import supcon_bank_loss
    G = golden("supcon.npz")
    f = cu(G[tag + "_feats"]).requires_grad_(True)
    M = f.shape[0]
    src = torch.arange(M, dtype=torch.int32, device="cuda")
    lab = cu(G[tag + "_labels"]).to(torch.int32)
    loss = supcon_bank_loss(f, f.new_zeros((1, 128)), src, lab, cu(G[tag + "_w"]),
                            torch.full((1,), M, dtype=torch.int32, device="cuda"), M + 37, 0.2)
    loss.backward()
    ref = float(G[tag + "_loss"])
    assert abs(float(loss) - ref) <= 1e-4 * abs(ref)
    np.testing.assert_allclose(f.grad.cpu().numpy(), G[tag + "_grad"], rtol=2e-3, atol=2e-8)


def test_supcon_module_duplicates(capi, supcon_path):
    """bank rows that repeat a source row (the [m] fallback, loss.py:338) accumulate their grads."""
    from odwscl_b200.modeling.sim_head import supcon_bank_loss
    g = torch.Generator().manual_seed(3)
    Fm = torch.nn.functional.normalize(torch.randn(50, 128, generator=g), dim=1)
    E = torch.nn.functional.normalize(torch.randn(20, 128, generator=g), dim=1)
    src = torch.randint(0, 70, (90,), generator=g)
    lab = torch.randint(0, 3, (90,), generator=g)
    w = torch.rand(90, generator=g)
    Fc, Ec = Fm.clone().requires_grad_(True), E.clone().requires_grad_(True)
    bank = torch.cat([Fc, Ec])[src]
    ref = orc.supcon_v2(bank, lab.float(), w, 0.2)
    ref.backward()
    Fg, Eg = Fm.cuda().requires_grad_(True), E.cuda().requires_grad_(True)
    loss = supcon_bank_loss(Fg, Eg, src.int().cuda(), lab.int().cuda(), w.cuda(),
                            torch.full((1,), 90, dtype=torch.int32, device="cuda"), 128, 0.2)
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    np.testing.assert_allclose(Fg.grad.cpu().numpy(), Fc.grad.numpy(), rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(Eg.grad.cpu().numpy(), Ec.grad.numpy(), rtol=2e-3, atol=1e-7)


@pytest.mark.parametrize("M,Mcap", [(3000, 3072), (2051, 2051), (1, 4)])
def test_supcon_large_bank_vs_oracle(capi, M, Mcap, supcon_path):
    """Bank sizes of 8 images per rank (configs[2]); M not a multiple of the tile sizes, padded bound > M, repeated source
    rows, rows from both F and E; a one-row bank (no other rows: the reference's 0/0) must give the same non-finite loss."""
    from odwscl_b200.modeling.sim_head import supcon_bank_loss
    g = torch.Generator().manual_seed(M)
    nF, nE = max(1, M // 2), max(1, M // 4)
    Fm = torch.nn.functional.normalize(torch.randn(nF, 128, generator=g), dim=1)
    E = torch.nn.functional.normalize(torch.randn(nE, 128, generator=g), dim=1)
    src = torch.randint(0, nF + nE, (M,), generator=g)
    lab = torch.randint(0, 6, (M,), generator=g)
    w = torch.rand(M, generator=g)
    Fc, Ec = Fm.clone().requires_grad_(True), E.clone().requires_grad_(True)
    ref = orc.supcon_v2(torch.cat([Fc, Ec])[src], lab.float(), w, 0.2)
    pad = torch.zeros(Mcap - M, dtype=torch.int32)
    Fg, Eg = Fm.cuda().requires_grad_(True), E.cuda().requires_grad_(True)
    loss = supcon_bank_loss(Fg, Eg, torch.cat([src.int(), pad]).cuda(), torch.cat([lab.int(), pad]).cuda(),
                            torch.cat([w, pad.float()]).cuda(), torch.full((1,), M, dtype=torch.int32, device="cuda"), Mcap, 0.2)
    if M == 1:
        assert not np.isfinite(float(ref)) and not np.isfinite(float(loss))
        return
    ref.backward()
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    scale = float(Fc.grad.abs().max())
    np.testing.assert_allclose(Fg.grad.cpu().numpy(), Fc.grad.numpy(), rtol=2e-3, atol=2e-4 * scale)
    np.testing.assert_allclose(Eg.grad.cpu().numpy(), Ec.grad.numpy(), rtol=2e-3, atol=2e-4 * scale)


# ------------------------------------------------------------------------------- DropBlock / sim
@pytest.mark.parametrize("block", [1, 3])
def test_dropblock_vs_oracle(capi, block):
    g = torch.Generator().manual_seed(block)
    x = torch.randn(37, 12, 7, 7, generator=g)
    cen = (torch.rand(37, 7, 7, generator=g) < 0.3 / block ** 2).float()
    y, sc = capi.dropblock(x.cuda(), cen.cuda(), block)
    np.testing.assert_allclose(y.cpu().numpy(), orc.dropblock(x, cen, block).numpy(), rtol=1e-6, atol=1e-7)
    gy, _ = capi.dropblock(x.cuda(), cen.cuda(), block, sc)          # backward re-uses the stored scale
    assert torch.equal(gy, y)


@pytest.mark.parametrize("n", [333, 2000])
def test_sim_nxn(capi, n):
    """tcgen05 3xTF32 similarity vs the fp32 matmul of loss.py:319 (fp64 as the arbiter)."""
    g = torch.Generator().manual_seed(0)
    Fm = torch.nn.functional.normalize(torch.randn(n, 128, generator=g) + 1.5, dim=1)
    got = capi.sim_nxn(Fm.cuda()).cpu()
    ref64 = (Fm.double() @ Fm.double().T)
    err = (got.double() - ref64).abs().max().item()
    err32 = ((Fm @ Fm.T).double() - ref64).abs().max().item()
    # fp32-class accuracy: the tensor core accumulates with truncation, measured ~4x an MKL sgemm's error
    assert err <= max(8 * err32, 4e-6), (err, err32)
    torch.testing.assert_close(got, Fm @ Fm.T, rtol=1e-5, atol=4e-6)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (200, 72, 64), (513, 300, 416), (4000, 4096, 1024)])
def test_gemm_nt_tf32(capi, M, N, K):
    """tcgen05 TF32 GEMM.  Inputs carry <= 10 mantissa bits, so TF32 products are exact and the
    result must match an fp32 matmul to fp32 accumulation error; then generic inputs at TF32 tolerance."""
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 8).round() / 8
    B = (torch.randn(N, K, generator=g) * 8).round() / 8
    got = capi.gemm_nt_tf32(A.cuda(), B.cuda()).cpu()
    ref = (A.double() @ B.double().T).float()
    torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-5 * K ** 0.5)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    got = capi.gemm_nt_tf32(A.cuda(), B.cuda()).cpu()
    ref = (A.double() @ B.double().T).float()
    assert float((got - ref).abs().max()) <= 8e-3 * K ** 0.5          # TF32: operands truncated to 10 mantissa bits (2^-10 relative each)


@pytest.mark.parametrize("B,C,H,W,R,S", [(2, 512, 76, 128, 600, 37), (1, 6, 152, 200, 90, 5), (2, 5, 20, 31, 64, 0)])
def test_roi_pool_bwd_multi_source(capi, B, C, H, W, R, S):
    """Several consumers of the pooled features (two dense gradients + gathered rows) scattered in one launch equal
    the single-source backward of their sum."""
    g = torch.Generator().manual_seed(R + S)
    feat = torch.randn(B, C, H, W, generator=g).cuda()
    rois = _rand_rois(g, B, H, W, R).cuda()
    out, arg = capi.roi_pool_forward(feat, rois, 0.125, 7, 7)
    q = lambda t: (t * 4).round() / 4                                   # exactly representable sums: bit-exact
    g1, g2 = q(torch.randn(out.shape, generator=g)).cuda(), q(torch.randn(out.shape, generator=g)).cuda()
    srows = torch.randint(0, R, (S,), generator=g).cuda() if S else None
    sgrad = q(torch.randn((S, C, 7, 7), generator=g)).cuda() if S else None
    got = capi.roi_pool_backward_multi(g1, g2, srows, sgrad, rois, arg, B, C, H, W)
    tot = g1 + g2
    if S:
        tot.index_add_(0, srows, sgrad)
    exp = capi.roi_pool_backward(tot, rois, arg, 7, 7, B, C, H, W)
    assert got is not None and torch.equal(got.contiguous(), exp)
    only1 = capi.roi_pool_backward_multi(g1, None, None, None, rois, arg, B, C, H, W)
    assert torch.equal(only1.contiguous(), capi.roi_pool_backward(g1, rois, arg, 7, 7, B, C, H, W))
    # second source through a per-(roi, bin) multiplier (the fused DropBlock backward)
    cen = (torch.rand(R, 7, 7, generator=g) < 0.05).float().cuda()
    _, sc = capi.dropblock(g2, cen, 3)
    sc[1] = 2.0                                                          # power of two: products stay exact
    bm = capi.dropblock_mask(cen, 3, sc)
    masked = capi.roi_pool_backward_multi(g1, g2, srows, sgrad, rois, arg, B, C, H, W, mask2=bm)
    tot2 = g1 + g2 * bm.view(R, 1, 7, 7)
    if S:
        tot2.index_add_(0, srows, sgrad)
    assert torch.equal(masked.contiguous(), capi.roi_pool_backward(tot2, rois, arg, 7, 7, B, C, H, W))


def test_relu_dropout_fused(capi):
    """ReLU + Dropout(p) in one in-place pass: zero where x <= 0, kept units scaled by 1/(1-p), keep rate 1-p, different
    masks for different seeds; backward = gy * scale on the surviving units (no mask tensor)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4096, 512, generator=g).cuda()
    for p in (0.0, 0.5, 0.3):
        y = capi.relu_dropout_(x.clone(), p, 1234)
        pos = x > 0
        assert float(y[~pos].abs().sum()) == 0.0
        kept = y > 0
        torch.testing.assert_close(y[kept], x[kept] * (1.0 / (1.0 - p)), rtol=1e-6, atol=0)
        rate = float(kept.sum()) / float(pos.sum())
        assert abs(rate - (1.0 - p)) < 5e-3, (p, rate)
        gy = torch.randn(x.shape, generator=g).cuda()
        gx = capi.relu_dropout_backward(y, gy, p)
        torch.testing.assert_close(gx, torch.where(kept, gy * (1.0 / (1.0 - p)), torch.zeros_like(gy)), rtol=1e-6, atol=0)
    y1, y2 = capi.relu_dropout_(x.clone(), 0.5, 1), capi.relu_dropout_(x.clone(), 0.5, 2)
    assert not torch.equal(y1 > 0, y2 > 0)
    assert torch.equal(capi.relu_dropout_(x.clone(), 0.5, 1), y1)          # same seed, same mask


def test_conv_weight_xform(capi):
    g = torch.Generator().manual_seed(5)
    w = torch.randn(96, 64, 3, 3, generator=g).cuda()
    wk, wd = capi.conv_weight_xform(w, want_fwd=True, want_dgrad=True, round_tf32=False)
    assert torch.equal(wk, w.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(wd, w.flip(2, 3).permute(1, 2, 3, 0).contiguous())
    wk_r, wd_r = capi.conv_weight_xform(w, want_fwd=True, want_dgrad=True, round_tf32=True)
    assert torch.equal(wk_r, capi.round_tf32_(wk.clone())) and torch.equal(wd_r, capi.round_tf32_(wd.clone()))


def test_roi_pool_forward_with_fused_dropblock_copy(capi):
    """ROIPool forward that also writes the DropBlock-augmented copy (rows [R, 2R) of the [2R,C,7,7] batch buffer) ==
    the plain forward followed by the separate DropBlock pass, bit for bit; the stored scale / mask match too."""
    g = torch.Generator().manual_seed(12)
    B, C, H, W, R = 2, 64, 38, 50, 333
    feat = torch.randn(B, C, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    rois = _rand_rois(g, B, H, W, R).cuda()
    cen = (torch.rand(R, 7, 7, generator=g) < 0.3 / 9).float().cuda()
    buf = torch.empty((2 * R, C, 7, 7), device="cuda")
    arg, sc, bm = capi.roi_pool_forward_aug(feat, rois, 0.125, cen, 3, buf)
    out, arg0 = capi.roi_pool_forward(feat, rois, 0.125, 7, 7)
    aug, sc0 = capi.dropblock(out, cen, 3)
    assert torch.equal(buf[:R], out) and torch.equal(arg, arg0)
    assert torch.equal(sc, sc0) and torch.equal(bm, capi.dropblock_mask(cen, 3, sc0))
    assert torch.equal(buf[R:], aug)


def test_aug_positives_builder(capi):
    """One kernel == gather + per-segment DropBlock + `noise * x + x` + concatenation (loss.py:296-305), forward and
    backward, with an injected noise tensor (bit-exact vs torch); with the built-in Philox noise: N(0,1) statistics,
    same draws in forward and backward, zero padding rows."""
    g = torch.Generator().manual_seed(21)
    R, C, Kc = 300, 16, 40
    seg = [0, 9, 9, 30]                                   # K = 30 real rows, 10 padding rows, an empty segment
    K, P = seg[-1], len(seg) - 1
    pooled = torch.randn(R, C, 7, 7, generator=g).cuda()
    rows = torch.randint(0, R, (Kc,), generator=g).cuda()
    cen = (torch.rand(Kc, 7, 7, generator=g) < 0.3).float().cuda()
    noise = torch.randn(Kc, C, 7, 7, generator=g).cuda()
    off = torch.tensor(seg, dtype=torch.int32).cuda()
    out, sc = capi.aug_positives(pooled, rows, off, P, cen, 1, noise=noise)
    X = pooled[rows]
    drop_ref, _ = capi.dropblock_seg(X, cen, 1, off, P)
    assert torch.equal(out[:K], drop_ref[:K]) and float(out[K:Kc].abs().sum()) == 0.0
    assert torch.equal(out[Kc:Kc + K], (noise * X + X)[:K]) and float(out[Kc + K:].abs().sum()) == 0.0
    # backward vs autograd of the same expression
    gy = torch.randn(2 * Kc, C, 7, 7, generator=g).cuda()
    gx, _ = capi.aug_positives(gy, rows, off, P, cen, 1, scale_seg=sc, noise=noise, backward=True)
    Xr = X.clone().requires_grad_(True)
    bm = (1 - cen) * 1.0
    scale = torch.zeros(Kc, device="cuda")
    for p in range(P):
        if seg[p + 1] > seg[p]:
            scale[seg[p]:seg[p + 1]] = sc[p, 1]
    ref = torch.cat([Xr * (bm * scale[:, None, None])[:, None], noise * Xr + Xr])
    valid = torch.zeros(2 * Kc, device="cuda"); valid[:K] = 1; valid[Kc:Kc + K] = 1
    (ref * gy * valid[:, None, None, None]).sum().backward()
    torch.testing.assert_close(gx, Xr.grad, rtol=1e-6, atol=1e-6)
    # built-in noise
    o1, sc1 = capi.aug_positives(pooled, rows, off, P, cen, 1, seed=123)
    o2, _ = capi.aug_positives(pooled, rows, off, P, cen, 1, seed=124)
    assert torch.equal(o1[:Kc], out[:Kc]) and not torch.equal(o1[Kc:], o2[Kc:])
    eps = (o1[Kc:Kc + K] / X[:K] - 1).flatten()
    assert abs(float(eps.mean())) < 0.02 and abs(float(eps.std()) - 1.0) < 0.02
    gx1, _ = capi.aug_positives(gy, rows, off, P, cen, 1, scale_seg=sc1, seed=123, backward=True)
    exp = gy[:K] * (bm * scale[:, None, None])[:K, None] + gy[Kc:Kc + K] * (o1[Kc:Kc + K] / X[:K])
    torch.testing.assert_close(gx1[:K], exp, rtol=2e-4, atol=2e-4)
