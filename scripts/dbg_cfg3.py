"""Debug (GPU box): where do non-finite values first appear at a given batch / image size?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odwscl_b200.config import get_cfg_defaults
from odwscl_b200.modeling import build_detection_model
from odwscl_b200.structures import BoxList
from odwscl_b200.synth import synth_batch
B, N, W, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
torch.manual_seed(0)
model = build_detection_model(get_cfg_defaults()).cuda().train()
images, rois, boxes, labels = synth_batch(B, N, W, H, 21, seed=1234)
props = [BoxList(b.cuda(), (W, H), "xyxy") for b in boxes]
fin = lambda t: bool(torch.isfinite(t).all())
with torch.no_grad():
    feat = model.backbone(images.cuda())[0]
    print("feat", tuple(feat.shape), fin(feat), float(feat.abs().max()))
    fe = model.roi_heads.feature_extractor
    clean, aug, pooled = fe.forward_clean_and_aug([feat], props)
    print("pooled", fin(pooled), float(pooled.abs().max()), "clean", fin(clean), float(clean.abs().max()), "aug", fin(aug), float(aug.abs().max()))
    simf = model.roi_heads.model_sim(clean)
    print("simf", fin(simf))
    cls, det, refs, bbs = model.roi_heads.predictor(aug, props)
    print("cls", fin(cls), "det", fin(det), [fin(r) for r in refs], [fin(b) for b in bbs])
torch.cuda.synchronize()
print("ok")
# ---- full train-mode forward with the loss (the bench's call)
targets = []
for lab in labels:
    t = BoxList(torch.zeros((len(lab), 4)), (W, H), "xyxy"); t.add_field("labels", torch.as_tensor(lab)); targets.append(t)
print("labels", labels)
losses, _ = model(images.cuda(), targets, props)
torch.cuda.synchronize()
print({k: float(v) for k, v in losses.items()})
sum(losses.values()).backward()
torch.cuda.synchronize()
print("backward ok")
