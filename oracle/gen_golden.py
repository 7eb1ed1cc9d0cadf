"""Generate tests/golden/*.npz from the REFERENCE ITSELF (imported from /root/reference through
oracle/ref_shims.py) and from torchvision (the reference's third-party ROIPool lineage / NMS).

Run in the build container only:  python -m oracle.gen_golden
TEST INFRASTRUCTURE ONLY.  The fixtures are committed; the GPU box never runs this script.

Fixture design notes
  * Similarity features in the contrastive fixtures live on an exact grid (multiples of 1/16,
    |x| <= 1/2, few non-zeros) so every fp32 dot product / row sum is exact regardless of
    summation order: the discrete `Sim >= tau` and `Sim > 1.0` decisions (loss.py:324-330) are
    then well defined across MKL / cuBLAS / hand-written kernels and can be compared BIT-EXACTLY.
  * Stochastic layers are replayed through oracle.StochasticSource (Dropout frozen; DropBlock
    centres and noise drawn from a seeded generator that the reference consumes through a
    module-level `torch` proxy -- the reference code itself is not modified).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from oracle import ref_shims              # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


class TorchProxy:
    """`torch` as seen by a reference module, with rand/normal replayed from a StochasticSource."""

    def __init__(self, rng):
        self._rng = rng

    def __getattr__(self, n):
        return getattr(torch, n)

    def rand(self, *shape, **k):
        return torch.rand(*shape, generator=self._rng.g)

    def normal(self, mean, std, size=None, device=None, **k):
        return self._rng.noise(tuple(size))


def rois_cases(g, B, H, W, scale, n):
    """random rois in image pixels incl. malformed / out-of-map / tiny ones."""
    iw, ih = W / scale, H / scale
    x1 = torch.rand(n, generator=g) * iw * 1.2 - 0.1 * iw
    y1 = torch.rand(n, generator=g) * ih * 1.2 - 0.1 * ih
    w = torch.rand(n, generator=g) * iw * 0.8 - 0.05 * iw
    h = torch.rand(n, generator=g) * ih * 0.8 - 0.05 * ih
    b = torch.randint(0, B, (n,), generator=g).float()
    r = torch.stack([b, x1, y1, x1 + w, y1 + h], 1)
    r[0, 1:] = torch.tensor([0.0, 0.0, iw - 1, ih - 1])          # whole map
    r[1, 1:] = torch.tensor([3.0, 3.0, 3.0, 3.0])                # 1x1
    r[2, 1:] = torch.tensor([iw + 50, ih + 50, iw + 90, ih + 90])  # fully outside
    r[3, 1:] = torch.tensor([-90.0, -90.0, -50.0, -50.0])        # fully outside (negative)
    r[4, 1:] = torch.tensor([40.0, 40.0, 10.0, 10.0])            # inverted
    r[5, 1:] = torch.tensor([4.0, 4.0, 4.0 + 0.5 / scale, 4.0 + 2.5 / scale])   # .5 rounding
    return r


def gen_roi_pool():
    g = torch.Generator().manual_seed(7)
    out = {}
    for tag, (B, C, H, W, ph, pw, quant) in {
        "a": (2, 6, 19, 27, 7, 7, False), "b": (1, 4, 12, 9, 3, 5, True), "c": (3, 5, 38, 50, 7, 7, True)}.items():
        feat = torch.randn(B, C, H, W, generator=g)
        if quant:   # plateaus / ties
            feat = (feat * 2).round() / 2
        rois = rois_cases(g, B, H, W, 0.125, 40)
        o, a = torch.ops.torchvision.roi_pool(feat, rois, 0.125, ph, pw)
        go = torch.randn(o.shape, generator=g)
        gi = torch.ops.torchvision._roi_pool_backward(go, rois, a, 0.125, ph, pw, B, C, H, W)
        for k, v in dict(feat=feat, rois=rois, out=o, argmax=a.int(), grad_out=go, grad_in=gi,
                         pooled=np.array([ph, pw])).items():
            out[tag + "_" + k] = v.numpy() if torch.is_tensor(v) else v
    save("roi_pool.npz", **out)


def gen_roi_align(refc):
    g = torch.Generator().manual_seed(8)
    out = {}
    for tag, (B, C, H, W, sr) in {"a": (2, 4, 19, 27, 0), "b": (1, 3, 12, 9, 2)}.items():
        feat = torch.randn(B, C, H, W, generator=g)
        rois = rois_cases(g, B, H, W, 0.125, 24)
        o = refc.roi_align_forward(feat, rois, 0.125, 7, 7, sr)            # reference CPU source
        go = torch.randn(o.shape, generator=g)
        gi = torch.ops.torchvision._roi_align_backward(go, rois, 0.125, 7, 7, B, C, H, W, sr, False)
        out.update({tag + "_feat": feat.numpy(), tag + "_rois": rois.numpy(), tag + "_out": o.numpy(),
                    tag + "_grad_out": go.numpy(), tag + "_grad_in": gi.numpy(), tag + "_sr": np.array(sr)})
    save("roi_align.npz", **out)


def gen_boxes(refc):
    from wetectron.structures.bounding_box import BoxList
    from wetectron.structures.boxlist_ops import boxlist_iou
    from wetectron.utils.utils import cal_iou, easy_nms
    g = torch.Generator().manual_seed(9)
    P = orc.synth_boxes(300, 500, 375, g)
    P = torch.cat([P, P[:20], P[5:15] + torch.tensor([1.0, 0, 1.0, 0])])      # exact duplicates, near dups
    Q = orc.synth_boxes(17, 500, 375, g)
    bl, ql = BoxList(P, (500, 375), "xyxy"), BoxList(Q, (500, 375), "xyxy")
    iou = boxlist_iou(bl, ql)
    scores = (torch.rand(P.shape[0], generator=g) * 8).round() / 8               # heavy ties
    scores_u = torch.rand(P.shape[0], generator=g)          # tie-free: the legacy sort is not stable
    out = dict(P=P.numpy(), Q=Q.numpy(), iou=iou.numpy(), scores=scores.numpy(), scores_u=scores_u.numpy())
    for t, m in enumerate([0, 17, 123]):
        idx, _ = cal_iou(bl, torch.tensor(m), 0.5)
        out["cal_iou_%d" % t] = idx.numpy(); out["cal_iou_m_%d" % t] = np.array(m)
    cluster = torch.nonzero(torch.rand(P.shape[0], generator=g) < 0.7).view(-1)
    for t, thr in enumerate([0.1, 0.3, 0.7]):
        out["easy_nms_%d" % t] = easy_nms(bl, cluster, scores, nms_iou=thr).numpy()
        out["easy_nms_thr_%d" % t] = np.array(thr, np.float32)
        out["legacy_nms_%d" % t] = refc.nms(P, scores_u, thr).numpy()            # cpu/nms_cpu.cpp
    out["cluster"] = cluster.numpy()
    import torchvision
    out["tv_nms_full"] = torchvision.ops.nms(P, scores, 0.3).numpy()
    save("boxes.npz", **out)


def grid_features(n, g, d=128):
    """rows on the exact grid: 16-18 non-zeros of +-1/4 -> |row|^2 in {1, 1.0625, 1.125}."""
    F_ = torch.zeros(n, d)
    proto = torch.randint(0, 6, (n,), generator=g)
    base = [torch.randperm(d, generator=torch.Generator().manual_seed(100 + k))[:18] for k in range(6)]
    for r in range(n):
        k = 16 + int(torch.randint(0, 3, (1,), generator=g))
        pos = base[int(proto[r])][:k].clone()
        nflip = int(torch.randint(0, 7, (1,), generator=g))
        repl = torch.randint(0, d, (nflip,), generator=g)
        pos[torch.randperm(k, generator=g)[:nflip]] = repl
        pos = torch.unique(pos)
        F_[r, pos] = 0.25
        neg = torch.rand(pos.numel(), generator=g) < 0.15
        F_[r, pos[neg]] = -0.25
    return F_


def gen_supcon():
    from wetectron.modeling.roi_heads.sim_head.sim_loss import SupConLossV2
    g = torch.Generator().manual_seed(10)
    out = {}
    for tag, sizes in {"a": [0, 40, 0, 57, 23], "b": [300, 0, 211]}.items():
        bank = [torch.nn.functional.normalize(torch.randn(s, 128, generator=g) + 2.0 * (i % 3), dim=1)
                .requires_grad_(True) if s else torch.zeros((0,)) for i, s in enumerate(sizes)]
        w = torch.rand(sum(sizes), generator=g) * 0.01
        loss = SupConLossV2(0.2)(bank, w, "cpu")
        loss.backward()
        feats = torch.cat([b for b in bank if b.numel()])
        labels = torch.cat([torch.full((s,), float(i)) for i, s in enumerate(sizes) if s])
        grads = torch.cat([b.grad for b in bank if b.numel()])
        out.update({tag + "_feats": feats.detach().numpy(), tag + "_labels": labels.numpy(), tag + "_w": w.numpy(),
                    tag + "_loss": loss.detach().numpy(), tag + "_grad": grads.numpy()})
    save("supcon.npz", **out)


class TinyExtractor(torch.nn.Module):
    """Stand-in with the three methods loss.py:298-304 calls, built from reference parts."""

    def __init__(self, c, hid, rngproxy):
        super().__init__()
        from wetectron.modeling.dropblock.drop_block import DropBlock2D
        self.fc = torch.nn.Linear(c * 49, hid)
        self.sim_drop = DropBlock2D(block_size=1, drop_prob=0.3)
        self._t = rngproxy

    def forward_neck(self, x):
        return torch.relu(self.fc(x.view(x.shape[0], -1)))

    def drop_pool(self, x):
        return self.sim_drop(x)

    def noise_pool(self, x):      # vgg16.py:177-180, verbatim formula
        noise = self._t.normal(0, 1 ** 2, size=x.shape, device=x.device)
        return noise * x + x


class GridSim(torch.nn.Module):
    """Sim head whose outputs are snapped to the exact 1/16 grid (straight-through)."""

    def __init__(self, hid):
        super().__init__()
        self.fc = torch.nn.Linear(hid, 128)

    def forward(self, x):
        y = torch.nn.functional.normalize(self.fc(x), dim=1) * 4.0
        q = torch.clamp(torch.round(y * 4) / 16, -0.5, 0.5)
        return y / 4.0 + (q - y / 4.0).detach()


def gen_roi_reg_loss():
    import wetectron.modeling.dropblock.drop_block as db_mod
    from wetectron.config import cfg
    from wetectron.modeling.roi_heads.weak_head import loss as loss_mod
    from wetectron.modeling.roi_heads.weak_head import pseudo_label_generator as plg
    from wetectron.structures.bounding_box import BoxList
    g = torch.Generator().manual_seed(11)
    Cc, hid, C = 4, 32, 21
    out = {}
    for tag, (sizes, labels) in {"a": ([220, 180], [[8, 13], [4, 13]]), "b": ([150], [[6]])}.items():
        rng = orc.StochasticSource(500 + len(sizes))
        proxy = TorchProxy(rng)
        db_mod.torch = proxy
        R = sum(sizes)
        boxes = [orc.synth_boxes(n, 500, 375, g) for n in sizes]
        props = [BoxList(b, (500, 375), "xyxy") for b in boxes]
        targets = []
        for lab in labels:
            t = BoxList(torch.tensor([[10.0, 10, 100, 100]] * len(lab)), (500, 375), "xyxy")
            t.add_field("labels", torch.tensor(lab))
            targets.append(t)
        mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).requires_grad_(True)
        cls, det = mk(R, C, sc=2.0), mk(R, C, sc=2.0)
        refs = [mk(R, C, sc=2.0) for _ in range(3)]
        bbs = [mk(R, 4 * C, sc=0.5) for _ in range(3)]
        simf = grid_features(R, g).requires_grad_(True)
        pooled = mk(R, Cc, 7, 7)
        torch.manual_seed(77)
        fe, ms = TinyExtractor(Cc, hid, proxy), GridSim(hid)
        fe.train(); ms.train()
        ev = loss_mod.RoIRegLossComputation(cfg)
        captured = []
        orig = plg.od_layer.__call__

        def spy(self, proposals, source_score, labels_, device, pgt_instance, return_targets=False):
            r = orig(self, proposals, source_score, labels_, device, pgt_instance, return_targets)
            captured.append(([p.clone() for p in pgt_instance], [x.clone() for x in r]))
            return r
        plg.od_layer.__call__ = spy
        try:
            losses, acc = ev([cls], [det], refs, bbs, simf, pooled, fe, ms, props, targets)
        finally:
            plg.od_layer.__call__ = orig
            db_mod.torch = torch
        total = sum(losses.values())
        total.backward()
        o = {"sizes": np.array(sizes), "cls": cls, "det": det, "simf": simf, "pooled": pooled,
             "fe_w": fe.fc.weight, "fe_b": fe.fc.bias, "ms_w": ms.fc.weight, "ms_b": ms.fc.bias,
             "g_cls": cls.grad, "g_det": det.grad, "g_simf": simf.grad, "g_pooled": pooled.grad,
             "rng_seed": np.array(500 + len(sizes))}
        for i in range(3):
            o["ref%d" % i] = refs[i]; o["bb%d" % i] = bbs[i]
            o["g_ref%d" % i] = refs[i].grad; o["g_bb%d" % i] = bbs[i].grad
        for b in range(len(sizes)):
            o["boxes%d" % b] = boxes[b]; o["labels%d" % b] = np.array(labels[b])
        for k, v in losses.items():
            o["loss_" + k] = v.detach()
        for k, v in acc.items():
            o["acc_" + k] = torch.as_tensor(v)
        k = 0
        for b in range(len(sizes)):
            for i in range(3):
                inst, (pl, lw, rt) = captured[k]; k += 1
                for c in range(C - 1):
                    if inst[c].numel():
                        o["inst_%d_%d_%d" % (b, i, c)] = inst[c]
                o["pl_%d_%d" % (b, i)] = pl; o["lw_%d_%d" % (b, i)] = lw; o["rt_%d_%d" % (b, i)] = rt
        for kk, v in o.items():
            out[tag + "_" + kk] = v.detach().numpy() if torch.is_tensor(v) else v
    save("roi_reg_loss.npz", **out)


def setup_cfg():
    """configs/voc/voc07_contra_db_b8_lr0.01_mcg.yaml + the README's `nms 0.1 lmda 0.03 iou 0.5 temp 0.2`."""
    from wetectron.config import cfg
    cfg.merge_from_file(os.path.join(ref_shims.REF_ROOT, "configs/voc/voc07_contra_db_b8_lr0.01_mcg.yaml"))
    cfg.merge_from_list(["nms", 0.1, "lmda", 0.03, "iou", 0.5, "temp", 0.2, "MODEL.DEVICE", "cpu"])
    return cfg


def build_reference_model(num_classes=21):
    from wetectron.config import cfg
    from wetectron.modeling.detector import build_detection_model
    model = build_detection_model(cfg)
    return model, cfg


def gen_model_cfg1():
    """BASELINE.json configs[0]: 1 synthetic 600x600 image, 256 proposals, VGG16 random init,
    CPU forward + losses through the reference's GeneralizedRCNN."""
    import wetectron.modeling.backbone.vgg16 as vgg_mod
    import wetectron.modeling.dropblock.drop_block as db_mod
    from wetectron.structures.bounding_box import BoxList
    from wetectron.structures.image_list import to_image_list
    model, cfg = build_reference_model()
    sd = orc.synth_state_dict(21, seed=0)
    missing = model.load_state_dict(sd, strict=True)
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0                                  # freeze Dropout
    images, boxes, labels = orc.synth_batch(1, 256, 600, 600, 21, seed=1234)
    rng = orc.StochasticSource(4242)
    proxy = TorchProxy(rng)
    db_mod.torch = proxy; vgg_mod.torch = proxy
    try:
        props = [BoxList(b, (600, 600), "xyxy") for b in boxes]
        targets = []
        for lab in labels:
            t = BoxList(torch.tensor([[10.0, 10, 100, 100]] * len(lab)), (600, 600), "xyxy")
            t.add_field("labels", torch.as_tensor(lab))
            targets.append(t)
        il = to_image_list([im[:, :600, :600] for im in images], 32)
        feats = {}
        h = model.backbone.register_forward_hook(lambda m, i, o: feats.__setitem__("feat", o[0].detach()))
        # first Sim_Net call of the step = sim_feature of all proposals (weak_head.py:110)
        h2 = model.roi_heads.model_sim.register_forward_hook(
            lambda m, i, o: (feats.setdefault("simf", o.detach().clone()), None)[1])
        h3 = model.roi_heads.predictor.register_forward_hook(
            lambda m, i, o: feats.__setitem__("pred", [o[0].detach(), o[1].detach()] + [r.detach() for r in o[2]]))
        from wetectron.modeling.roi_heads.weak_head import pseudo_label_generator as plg
        captured = []
        orig = plg.od_layer.__call__

        def spy(self, proposals, source_score, labels_, device, pgt_instance, return_targets=False):
            r = orig(self, proposals, source_score, labels_, device, pgt_instance, return_targets)
            captured.append(([p.clone() for p in pgt_instance], [x.clone() for x in r]))
            return r
        plg.od_layer.__call__ = spy
        try:
            losses, acc = model(il, targets, props)
        finally:
            plg.od_layer.__call__ = orig
        h.remove(); h2.remove(); h3.remove()
    finally:
        db_mod.torch = torch; vgg_mod.torch = torch
    o = {"loss_" + k: v.detach().numpy() for k, v in losses.items()}
    # discovered pseudo-GT sets / pseudo labels per refinement branch (1 image) and the head outputs they
    # were mined from: lets the GPU test tell a genuine mismatch from an ulp-level `Sim >= tau` flip
    for i in range(3):
        inst, (pl, lw, rt) = captured[i]
        for c in range(20):
            if inst[c].numel():
                o["inst_0_%d_%d" % (i, c)] = inst[c].numpy()
        o["pl_0_%d" % i] = pl.numpy(); o["lw_0_%d" % i] = lw.numpy()
    o["simf"] = feats["simf"].numpy()
    for k, v in zip(("cls", "det", "ref0", "ref1", "ref2"), feats["pred"]):
        o["head_" + k] = v.numpy()
    o.update({"acc_" + k: np.asarray(float(v)) for k, v in acc.items()})
    f = feats["feat"]
    o["feat_shape"] = np.array(f.shape); o["feat_sample"] = f[0, ::37, ::5, ::7].numpy()
    o["feat_absmean"] = f.abs().mean().numpy()
    save("model_cfg1.npz", **o)


def main():
    ref_shims.install()
    setup_cfg()
    from oracle import build_ref
    refc = build_ref.load()
    which = sys.argv[1:] or ["roi_pool", "roi_align", "boxes", "supcon", "roi_reg_loss", "model_cfg1"]
    if "roi_pool" in which: gen_roi_pool()
    if "roi_align" in which: gen_roi_align(refc)
    if "boxes" in which: gen_boxes(refc)
    if "supcon" in which: gen_supcon()
    if "roi_reg_loss" in which: gen_roi_reg_loss()
    if "model_cfg1" in which: gen_model_cfg1()


if __name__ == "__main__":
    main()
