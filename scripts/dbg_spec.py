"""Debug (GPU box): per-parameter gradient difference between the synced and the speculative-K step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as orc
from odwscl_b200.config import cfg
from odwscl_b200.modeling import build_detection_model, fc
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
    t = BoxList(torch.zeros((len(lab), 4)), (400, 320), "xyxy"); t.add_field("labels", torch.as_tensor(lab)); targets.append(t)
fe, ev = model.roi_heads.feature_extractor, model.roi_heads.loss_evaluator
def run():
    g = torch.Generator().manual_seed(99)
    sampler = lambda n, h, w, gamma, dev: (torch.rand(n, h, w, generator=g) < gamma).float().to(dev)
    fe.dropblock.centre_sampler = sampler; fe.sim_drop.centre_sampler = sampler
    gn = torch.Generator().manual_seed(7)
    fe.noise_sampler = lambda shape, dev: torch.randn(tuple(shape), generator=gn).to(dev)
    model.zero_grad(set_to_none=True)
    losses, _ = model(images.cuda(), targets, props)
    sum(losses.values()).backward()
    return {k: float(v) for k, v in losses.items()}, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}, ev.last_state.E.clone()
for tag, flags in (("default", {}), ("nofuse", {"fuse": False}), ("nofold", {"fold": False})):
    fc.FUSE_ACT_BWD = flags.get("fuse", True)
    fe.merge_fc6_wgrad = flags.get("fold", True)
    ev.speculative_k = False; ev._k_cap = None; ev._k_event = None
    l0, g0, E0 = run()
    K = int(ev.last_state.offA[-1])
    ev.speculative_k = True
    run()
    l1, g1, E1 = run()
    print("==", tag, "K", K, "cap", ev._k_cap, "overflow", float(ev.overflow), "E diff", float((E0[:2*K] - E1[:2*K]).abs().max()))
    for k in g0:
        d = float((g0[k] - g1[k]).abs().max()); m = float(g0[k].abs().max())
        if d > 1e-4 * m:
            print("   %-50s max|g| %.3e  max diff %.3e  rel %.2e" % (k, m, d, d / max(m, 1e-30)))
