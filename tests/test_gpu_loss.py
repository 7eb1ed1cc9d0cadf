"""GPU parity of the contrastive head (object discovery + SupCon + od_layer + MIL/refine losses)
against the reference-generated golden vectors and the oracle, and of the full model at
BASELINE.json configs[0]."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as orc
from tests.helpers import case_tensors, grid_sim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


class CpuStandInExtractor:
    """The tiny stand-in of tests/golden/roi_reg_loss.npz.  It computes on the CPU exactly as the
    fixture generator did, so the embeddings fed to the GPU kernels are bit-identical to the
    reference run; what is under test is everything the product does with them."""

    def __init__(self, t, rng):
        self.t, self.rng = t, rng

    def drop_pool(self, x):
        x = x.cpu()
        return orc.dropblock(x, self.rng.dropblock_centres(x.shape[0], 1), 1)

    def noise_pool(self, x):
        x = x.cpu()
        return self.rng.noise(x.shape) * x + x

    def forward_neck(self, x):
        return torch.relu(F.linear(x.reshape(x.shape[0], -1), self.t["fe_w"], self.t["fe_b"]))


def product_loss_case(G, tag):
    from odwscl_b200.config import cfg
    from odwscl_b200.modeling.loss import RoIRegLossComputation
    from odwscl_b200.structures import BoxList
    t = case_tensors(G, tag)          # CPU leaf tensors
    dev = {k: t[k].detach().cuda().requires_grad_(True) for k in
           ("cls", "det", "simf", "pooled", "ref0", "ref1", "ref2", "bb0", "bb1", "bb2")}
    rng = orc.StochasticSource(t["seed"])
    fe = CpuStandInExtractor(t, rng)
    model_sim = lambda h: grid_sim(h, t["ms_w"], t["ms_b"]).cuda()
    props = [BoxList(b.cuda(), (500, 375), "xyxy") for b in t["boxes"]]
    targets = []
    for lab in t["labels"]:
        tg = BoxList(torch.zeros((len(lab), 4)), (500, 375), "xyxy")
        tg.add_field("labels", torch.as_tensor(lab))
        targets.append(tg)
    ev = RoIRegLossComputation(cfg)
    ev.batch_aug = False              # per-(image, class) drop/noise order, to replay the reference RNG
    losses, accs = ev([dev["cls"]], [dev["det"]], [dev["ref0"], dev["ref1"], dev["ref2"]],
                      [dev["bb0"], dev["bb1"], dev["bb2"]], dev["simf"], dev["pooled"], fe, model_sim, props, targets)
    return t, dev, ev, losses, accs


@pytest.mark.parametrize("tag", ["a", "b"])
def test_roi_reg_loss_golden(golden, tag):
    """Index selection bit-exact (discovered instances, pseudo labels); losses <= 1e-4 rel
    (north_star); gradients within 2e-3 rel of the reference's autograd."""
    G = golden("roi_reg_loss.npz")
    t, dev, ev, losses, accs = product_loss_case(G, tag)
    st = ev.last_state
    pair_img, pair_cls = st.pair_img.cpu().numpy(), st.pair_cls.cpu().numpy()
    inst, cnt = st.inst.cpu().numpy(), st.inst_cnt.cpu().numpy()
    seen = set()
    for p in range(st.P):
        for i in range(3):
            key = "%s_inst_%d_%d_%d" % (tag, pair_img[p], i, pair_cls[p])
            seen.add(key)
            assert np.array_equal(inst[p, i, :cnt[p, i]], G[key]), key
    assert seen == {k for k in G if k.startswith(tag + "_inst_")}
    from odwscl_b200 import capi
    pl, lw, rt = capi.od_layer(st, 0.5)
    off = 0
    for b, n in enumerate(t["sizes"]):
        for i in range(3):
            assert np.array_equal(pl[i, off:off + n].cpu().numpy(), G["%s_pl_%d_%d" % (tag, b, i)])
            np.testing.assert_allclose(lw[i, off:off + n].cpu().numpy(), G["%s_lw_%d_%d" % (tag, b, i)], rtol=1e-5)
            np.testing.assert_allclose(rt[i, off:off + n].cpu().numpy(), G["%s_rt_%d_%d" % (tag, b, i)],
                                       rtol=1e-5, atol=1e-6)
        off += n
    for k, v in losses.items():
        ref = float(G["%s_loss_%s" % (tag, k)])
        assert abs(float(v) - ref) <= 1e-4 * abs(ref) + 1e-9, (k, float(v), ref)
    for k, v in accs.items():
        assert abs(float(v) - float(G["%s_acc_%s" % (tag, k)])) <= 1e-6, k
    sum(losses.values()).backward()
    for k in ("cls", "det", "simf", "pooled", "ref0", "ref1", "ref2", "bb0", "bb1", "bb2"):
        np.testing.assert_allclose(dev[k].grad.cpu().numpy(), G["%s_g_%s" % (tag, k)], rtol=2e-3, atol=2e-8,
                                   err_msg=k)


@pytest.mark.parametrize("sizes,pos", [
    ([1500, 1100], [[2, 9, 17], [5, 9]]),
    ([8192], [[3, 11]]),                                        # the largest proposal list the C ABI accepts (smem carve)
    ([40] * 30, [[c % 20 for c in range(b, b + 5)] for b in range(30)]),   # 150 pairs (> 128), many small images
])
def test_discovery_stagewise_vs_oracle_random(sizes, pos):
    """Larger random cases on the exact similarity grid (N = 1500 / 1100 with 3 + 2 classes; one image at the
    Ncap = 8192 limit; 150 (image, class) pairs): every discovered set, bank row order and weight order equals the
    oracle's, bit for bit."""
    from odwscl_b200 import capi
    from oracle.gen_golden import grid_features
    g = torch.Generator().manual_seed(21)
    C = 21
    pos = [sorted(set(p)) for p in pos]
    R = sum(sizes)
    boxes = [orc.synth_boxes(n, 1000, 600, g) for n in sizes]
    mk = lambda *s: torch.randn(*s, generator=g) * 2.0
    final = torch.softmax(mk(R, C), 1) * torch.cat([torch.softmax(d, 0) for d in mk(R, C).split(sizes)])
    refl = [mk(R, C) for _ in range(3)]
    simf = grid_features(R, g)
    emb = {}

    def embed_aug(b, c, I, kind):
        e = grid_features(len(I), torch.Generator().manual_seed(1000 * b + 10 * c + (kind == "drop")))
        emb[(b, c, kind)] = e
        return e
    bank, Wt, inst, idx, tr = orc.discover(boxes, final.split(sizes), [r.split(sizes) for r in refl], simf.split(sizes),
                                           pos, embed_aug, 0.5, 0.1, C)
    ofeat, olab = [], []
    for c, rows in enumerate(bank):
        if rows:
            ofeat.append(torch.cat(rows)); olab += [c] * ofeat[-1].shape[0]
    ofeat, ow = torch.cat(ofeat), torch.cat([w.view(-1) for w in Wt])

    pair_img = [b for b in range(len(sizes)) for _ in pos[b]]
    pair_cls = [c for b in range(len(sizes)) for c in pos[b]]
    P = len(pair_img)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32).cuda()
    img_off = [0]
    for n_ in sizes:
        img_off.append(img_off[-1] + n_)
    scores = (final.cuda().contiguous(), torch.softmax(refl[0], 1).cuda().contiguous(),
              torch.softmax(refl[1], 1).cuda().contiguous())
    st = capi.discover_phase_a(torch.cat(boxes).cuda(), i32(img_off), scores, i32(pair_img), i32(pair_cls),
                               max(sizes), 0.5)
    offA = st.offA.cpu().numpy()
    K = int(offA[P])
    rowsA = st.rowsA[:K].cpu().numpy()
    for p in range(P):
        exp = tr["phaseA_idx"][(pair_img[p], pair_cls[p])] + img_off[pair_img[p]]
        assert np.array_equal(rowsA[offA[p]:offA[p + 1]], exp)
    E = torch.cat([emb[(pair_img[p], pair_cls[p], "drop")] for p in range(P)] +
                  [emb[(pair_img[p], pair_cls[p], "noise")] for p in range(P)]).cuda()
    capi.discover_phase_b(st, simf.cuda(), E, 0.1)
    capi.bank_assemble(st, C - 1, 3 * K + 3 * sum(sizes[b] for b in pair_img))
    M = int(st.M[0])
    instd, cnt = st.inst.cpu().numpy(), st.inst_cnt.cpu().numpy()
    tau = st.tau.cpu().numpy()
    for p in range(P):
        for i in range(3):
            key = (pair_img[p], i, pair_cls[p])
            assert tau[p, i] == np.float32(tr["tau"][key]), key
            assert np.array_equal(instd[p, i, :cnt[p, i]], inst[key[0]][i][key[2]]), key
    assert M == ofeat.shape[0]
    V = torch.cat([simf, E.cpu()])
    assert torch.equal(V[st.row_src[:M].cpu().long()], ofeat)
    assert st.row_lab[:M].cpu().tolist() == olab
    np.testing.assert_allclose(st.row_w[:M].cpu().numpy(), ow.numpy(), rtol=2e-6)


def test_model_cfg1_golden(golden):
    """BASELINE.json configs[0] (1 x 600x600, 256 proposals, VGG16 random init): the product
    GeneralizedRCNN on the GPU (strict fp32) against the reference's losses.  ROI features within
    1e-4 rel; losses within 1e-3 rel (13 conv layers of cuDNN-vs-MKL rounding in between);
    loss_sim is compared on its own."""
    from odwscl_b200.config import cfg
    from odwscl_b200.modeling import build_detection_model
    from odwscl_b200.structures import BoxList
    G = golden("model_cfg1.npz")
    model = build_detection_model(cfg)
    model.load_state_dict(orc.synth_state_dict(21, seed=0), strict=True)
    model.cuda().train()
    for m in model.modules():
        if hasattr(m, "strict_fp32"):
            m.strict_fp32 = True                  # 3-pass TF32 split: fp32-accurate convolutions / fc GEMMs for the 1e-4 gate
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    images, boxes, labels = orc.synth_batch(1, 256, 600, 600, 21, seed=1234)
    rng = orc.StochasticSource(4242)
    fe = model.roi_heads.feature_extractor
    sampler = lambda n, h, w, gamma, dev: (torch.rand(n, h, w, generator=rng.g) < gamma).float().to(dev)
    fe.dropblock.centre_sampler = sampler
    fe.sim_drop.centre_sampler = sampler
    fe.noise_sampler = lambda shape, dev: rng.noise(tuple(shape)).to(dev)
    model.roi_heads.loss_evaluator.batch_aug = False
    with torch.no_grad():
        feat = model.backbone(images.cuda())[0]
    ref_s = torch.from_numpy(G["feat_sample"])
    got_s = feat[0, ::37, ::5, ::7].cpu()
    assert float((got_s - ref_s).abs().max()) <= 1e-4 * float(ref_s.abs().max())
    props = [BoxList(b.cuda(), (600, 600), "xyxy") for b in boxes]
    targets = []
    for lab in labels:
        t = BoxList(torch.zeros((len(lab), 4)), (600, 600), "xyxy")
        t.add_field("labels", torch.as_tensor(lab))
        targets.append(t)
    caught = {}
    h = model.roi_heads.model_sim.register_forward_hook(
        lambda m, i, o: (caught.setdefault("simf", o.detach()), None)[1])
    losses, accs = model(images.cuda(), targets, props)
    h.remove()
    # Sim_Net embeddings of all proposals vs the reference's (unit rows -> absolute tolerance)
    assert float((caught["simf"].cpu() - torch.from_numpy(G["simf"])).abs().max()) <= 1e-4
    # discovered pseudo-GT sets vs the reference's.  `Sim[m] >= tau` (loss.py:324) is decided on the last ulp
    # of an fp32 dot product and all cosines sit in a 0.1-wide band at random init, so a set
    # may legitimately differ between MKL and the GPU; a difference is accepted only when the product's own
    # similarity row has an element within 2e-6 of tau for that (branch, class) -- then the dependent branch
    # losses are compared loosely.  Otherwise everything is held to 1e-3.
    st = model.roi_heads.loss_evaluator.last_state
    inst, cnt = st.inst.cpu().numpy(), st.inst_cnt.cpu().numpy()
    tau, amax = st.tau.cpu().numpy(), st.amax.cpu().numpy()
    pair_cls = st.pair_cls.cpu().numpy()
    Fg = caught["simf"]
    flipped = []
    for p in range(st.P):
        for i in range(3):
            got = inst[p, i, :cnt[p, i]]
            exp = G["inst_0_%d_%d" % (i, pair_cls[p])]
            if not np.array_equal(got, exp):
                row = (Fg @ Fg[int(amax[p, i])]).cpu().numpy()
                margin = float(np.abs(row - tau[p, i]).min())
                # multi-class images: the other class's top proposal m_n is kept or dropped on whether its
                # self-similarity rounds to > 1.0 (loss.py:327, SURVEY App. A QUIRK) -- also an ulp decision
                for q in range(st.P):
                    if q != p:
                        mn = int(amax[q, i])
                        margin = min(margin, abs(float(Fg[mn] @ Fg[mn]) - 1.0))
                flipped.append((i, int(pair_cls[p]), margin, len(got), len(exp)))
                assert margin <= 2e-6, ("unjustified selection difference", flipped[-1])
    print("selection flips (branch, class, margin, n_got, n_ref):", flipped)
    loose = {"loss_ref_cls%d" % i for i, *_ in flipped} | {"loss_ref_reg%d" % i for i, *_ in flipped}
    if flipped:
        loose.add("loss_sim")
    for k, v in losses.items():
        ref = float(G["loss_" + k])
        tol = 5e-2 if k in loose else 1e-3
        assert abs(float(v) - ref) <= tol * abs(ref), (k, float(v), ref, flipped)
    sum(losses.values()).backward()
    gsum = sum(float(p.grad.abs().sum()) for p in model.parameters() if p.grad is not None)
    assert np.isfinite(gsum) and gsum > 0


def test_speculative_k_matches_synced():
    """The step without a host sync (augmented-positives batch padded to a bound on K, masked on the device) computes
    the same losses and gradients as the path that reads K back; an exceeded bound raises `overflow` instead."""
    from odwscl_b200.config import cfg
    from odwscl_b200.modeling import build_detection_model
    from odwscl_b200.structures import BoxList
    model = build_detection_model(cfg)
    model.load_state_dict(orc.synth_state_dict(21, seed=0), strict=True)
    model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    images, boxes, labels = orc.synth_batch(2, 200, 400, 320, 21, seed=77)
    props = [BoxList(b.cuda(), (400, 320), "xyxy") for b in boxes]
    targets = []
    for lab in labels:
        t = BoxList(torch.zeros((len(lab), 4)), (400, 320), "xyxy")
        t.add_field("labels", torch.as_tensor(lab))
        targets.append(t)
    fe, ev = model.roi_heads.feature_extractor, model.roi_heads.loss_evaluator

    def run():
        g = torch.Generator().manual_seed(99)      # CPU generator: rand(n)[:k] == rand(k), so padded batches see the same masks
        sampler = lambda n, h, w, gamma, dev: (torch.rand(n, h, w, generator=g) < gamma).float().to(dev)
        fe.dropblock.centre_sampler = sampler
        fe.sim_drop.centre_sampler = sampler
        gn = torch.Generator().manual_seed(7)
        fe.noise_sampler = lambda shape, dev: torch.randn(tuple(shape), generator=gn).to(dev)
        model.zero_grad(set_to_none=True)
        losses, _ = model(images.cuda(), targets, props)
        sum(losses.values()).backward()
        grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        return {k: float(v) for k, v in losses.items()}, grads

    ev.speculative_k = False
    ref_l, ref_g = run()
    K = int(ev.last_state.offA[-1])
    assert K > 0
    ev.speculative_k = True
    run()                                          # first call: reads K back once and sets the bound
    assert ev._k_cap is not None and ev._k_cap >= K
    got_l, got_g = run()                           # no host sync in here
    assert float(ev.overflow) == 0.0
    for k in ref_l:
        assert abs(got_l[k] - ref_l[k]) <= 1e-5 * max(abs(ref_l[k]), 1e-3), (k, got_l[k], ref_l[k])
    # the padded batch changes the row count of the fc weight-gradient contractions, hence their split-K partition and
    # summation order (the tensor core adds into TMEM with truncation); scatter-adds use fp32 atomics: order noise,
    # scaled by the tensor
    # Sim_Net's parameters only receive the SupCon gradient (|g| ~ 1e-8 here): the padded batch runs the small fc GEMMs
    # with another split-K partition, the augmented embeddings move by ~1e-4 (single-pass TF32, truncating TMEM
    # accumulation) and SupCon at T = 0.2 over cosines packed in [0.9, 1] amplifies that to a few % of max|g|
    for k in ref_g:
        a = 0.1 if "model_sim" in k else 1e-3
        torch.testing.assert_close(got_g[k], ref_g[k], rtol=2e-4, atol=a * float(ref_g[k].abs().max()) + 1e-9)
    # the same step with SupCon on its tensor-core path (production: banks of >= 3072 rows; forced here): the bank bound
    # then comes from the row counts read back so far, and the loss / gradients agree with the tile kernels
    from odwscl_b200.modeling import sim_head
    M = int(ev.last_state.M[1])
    assert M == int(ev.last_state.M[0]) and ev._poll_m_cap() is not None and ev._m_cap >= M
    old_thr, sim_head.SUPCON_TC_MIN_ROWS = sim_head.SUPCON_TC_MIN_ROWS, 64
    try:
        tc_l, tc_g = run()
        assert float(ev.overflow) == 0.0
        assert abs(tc_l["loss_sim"] - got_l["loss_sim"]) <= 1e-5 * abs(got_l["loss_sim"])
        for k in ref_g:
            a = 0.1 if "model_sim" in k else 1e-3
            torch.testing.assert_close(tc_g[k], got_g[k], rtol=2e-4, atol=a * float(got_g[k].abs().max()) + 1e-9)
        ev._m_cap, ev._m_event = 64, None          # bank bound too small: flagged, finite
        bad_l, _ = run()
        assert float(ev.overflow) == 1.0 and int(ev.last_state.M[0]) == 64 and int(ev.last_state.M[1]) == M
        assert all(np.isfinite(v) for v in bad_l.values())
        torch.cuda.synchronize()
        assert ev._poll_m_cap() >= M               # ... and the read-back raises the bound for the redo
    finally:
        sim_head.SUPCON_TC_MIN_ROWS = old_thr
    ev._k_cap = 1                                  # bound too small: flagged, finite, nothing out of range
    ev._k_event = None
    bad_l, _ = run()
    assert float(ev.overflow) == 1.0
    assert all(np.isfinite(v) for v in bad_l.values())


@pytest.mark.parametrize("name,B,N,W,H,C", [
    ("voc12_max_scale", 1, 300, 1600, 1200, 21),      # cfg 4: 152x200 map -> ROIPool-backward plane kernel with 1 channel per CTA
    ("coco_shape", 1, 4000, 640, 480, 81),            # cfg 5: 4000 proposals, 81 classes
    ("batch4_mixed", 4, 500, 500, 375, 21),           # several images per GPU (cfg 3 shape, scaled down), odd tile counts
])
def test_full_step_other_configs(name, B, N, W, H, C):
    """One full train step (forward, all eight losses, backward) at the shapes of the other BASELINE configs:
    finite losses, finite non-zero gradients on every trainable parameter, and the step without a host sync agrees."""
    from odwscl_b200.config import get_cfg_defaults
    from odwscl_b200.modeling import build_detection_model
    from odwscl_b200.structures import BoxList
    cfg = get_cfg_defaults()
    cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES = C
    torch.manual_seed(0)
    model = build_detection_model(cfg).cuda().train()
    images, boxes, labels = orc.synth_batch(B, N, W, H, C, seed=4321)
    props = [BoxList(b.cuda(), (W, H), "xyxy") for b in boxes]
    targets = []
    for lab in labels:
        t = BoxList(torch.zeros((len(lab), 4)), (W, H), "xyxy")
        t.add_field("labels", torch.as_tensor(lab))
        targets.append(t)
    ev = model.roi_heads.loss_evaluator
    ev.speculative_k = True
    clean_spec_passes = 0
    for it in range(6):                               # passes after the first run the speculative (sync-free) path
        speculative = ev._poll_k_cap() is not None
        model.zero_grad(set_to_none=True)
        losses, accs = model(images.cuda(), targets, props)
        assert set(losses) == {"loss_img", "loss_ref_cls0", "loss_ref_reg0", "loss_ref_cls1", "loss_ref_reg1",
                               "loss_ref_cls2", "loss_ref_reg2", "loss_sim"}
        total = sum(losses.values())
        total.backward()
        assert np.isfinite(float(total)), (name, {k: float(v) for k, v in losses.items()})
        if float(ev.overflow) != 0.0:                 # K (volatile at random init: it follows the DropBlock / Dropout
            torch.cuda.synchronize()                  # draws) outgrew the bound: the step is flagged, the bound is raised
            continue                                  # from the true K that was read back, and the step is redone
        clean_spec_passes += int(speculative)
        for k, p in model.named_parameters():
            if p.requires_grad:
                assert p.grad is not None and bool(torch.isfinite(p.grad).all()), (name, k)
        gsum = sum(float(p.grad.abs().sum()) for p in model.parameters() if p.grad is not None)
        assert gsum > 0
        if clean_spec_passes >= 1:
            break
    assert clean_spec_passes >= 1, "no speculative pass completed without overflow"


@pytest.mark.parametrize("sizes,C,agn", [([300, 211], 21, False), ([500], 81, False), ([64, 64, 64], 21, True)])
def test_head_loss_kernels_vs_torch(sizes, C, agn):
    """csrc/head_loss.cu (scores, seven losses, four accuracies, closed-form logits gradient) against an fp64 torch
    restatement of loss.py:234-259,349-406 with autograd."""
    from odwscl_b200 import capi
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(sum(sizes) + C)
    B, R = len(sizes), sum(sizes)
    Q = 8 if agn else 4 * C
    W = 5 * C + 3 * Q
    logits = (torch.randn(R, W, generator=g) * 1.5).cuda()
    off = [0]
    for n in sizes:
        off.append(off[-1] + n)
    img_off = torch.tensor(off, dtype=torch.int32).cuda()
    labels = torch.zeros(B, C)
    for b in range(B):
        labels[b, torch.randperm(C - 1, generator=g)[: 1 + b % 3] + 1] = 1.0
    pl = torch.randint(0, C, (3, R), generator=g)
    pl[:, ::3] = 0                                     # a share of background rows
    lw = torch.rand(3, R, generator=g)
    rt = torch.randn(3, R, 4, generator=g) * 1.5       # both smooth-L1 regimes
    hs = capi.head_scores(logits, C, Q, img_off, B)
    out, grad = capi.head_loss(hs, labels.cuda(), pl.cuda(), lw.cuda(), rt.cuda(), agn, 1e-8)
    # ---- fp64 restatement
    z = logits.double().cpu().requires_grad_(True)
    cls, det = z[:, :C], z[:, C:2 * C]
    refs = [z[:, 2 * C + i * (C + Q): 3 * C + i * (C + Q)] for i in range(3)]
    bbs = [z[:, 3 * C + i * (C + Q): 3 * C + i * (C + Q) + Q] for i in range(3)]
    final = F.softmax(cls, 1) * torch.cat([F.softmax(d, 0) for d in det.split(sizes)])
    torch.testing.assert_close(hs.final_score.cpu().double(), final.detach(), rtol=2e-5, atol=1e-9)
    torch.testing.assert_close(hs.sm1.cpu().double(), F.softmax(refs[0], 1).detach(), rtol=2e-5, atol=1e-9)
    torch.testing.assert_close(hs.sm2.cpu().double(), F.softmax(refs[1], 1).detach(), rtol=2e-5, atol=1e-9)
    L = [0.0] * 7
    acc = [0.0] * 4
    ar4 = torch.arange(4)

    def topk_acc(score, lab):
        k = max(int(lab.sum()), 1)
        return float(lab[score.topk(k)[1]].mean()) if lab.numel() else 0.0
    labd = labels.double()
    for b, (f, n) in enumerate(zip(final.split(sizes), sizes)):
        img = torch.clamp(f.sum(0), 1e-8, 1 - 1e-8)
        L[0] = L[0] + F.binary_cross_entropy(img, labd[b])
        kk = max(int(labels[b].sum()), 1)
        acc[0] += float(labd[b][img.topk(kk)[1]].sum() / kk)
        for i in range(3):
            lm = 3 if i == 0 else 1
            zi = refs[i][off[b]:off[b + 1]]
            pli, lwi = pl[i, off[b]:off[b + 1]], lw[i, off[b]:off[b + 1]].double()
            L[1 + 2 * i] = L[1 + 2 * i] + lm * (F.cross_entropy(zi, pli, reduction="none") * lwi).mean()
            fg = (pli > 0).double()
            idx = (ar4[None] + 4).expand(n, 4) if agn else 4 * pli[:, None] + ar4[None]
            d = (bbs[i][off[b]:off[b + 1]].gather(1, idx) - rt[i, off[b]:off[b + 1]].double()).abs()
            sl = torch.where(d < 1, 0.5 * d * d, d - 0.5)
            L[2 + 2 * i] = L[2 + 2 * i] + lm * ((sl * (lwi * fg)[:, None]).sum(1)).mean()
            rs = zi.sum(0)[1:]
            acc[1 + i] += float(labd[b][1:][rs.topk(kk)[1]].sum() / kk)
    total = sum(L) / B
    total.backward()
    got = out.cpu().double()
    for k in range(7):
        assert abs(float(got[k]) - float(L[k]) / B) <= 2e-5 * abs(float(L[k]) / B) + 1e-9, (k, float(got[k]), float(L[k]) / B)
    for k in range(4):
        assert abs(float(got[7 + k]) - acc[k] / B) <= 1e-6, (k, float(got[7 + k]), acc[k] / B)
    torch.testing.assert_close(grad[:, :W].cpu().double(), z.grad, rtol=1e-4, atol=1e-7 * float(z.grad.abs().max()) + 1e-12)
    # upstream scaling: block k scaled by g[k]
    up = torch.tensor([2.0, 0.5, 3.0, 1.0, 0.0, -1.0, 4.0]).cuda()
    g2 = capi.head_grad_scale_(grad.clone(), C, Q, up)
    blk = lambda k: (slice(0, 2 * C) if k == 0 else
                     (lambda i, r: slice(2 * C + i * (C + Q) + (C if r else 0), 2 * C + i * (C + Q) + (C + Q if r else C)))((k - 1) // 2, (k - 1) % 2))
    for k in range(7):
        torch.testing.assert_close(g2[:, blk(k)], grad[:, blk(k)] * up[k], rtol=1e-6, atol=0)
