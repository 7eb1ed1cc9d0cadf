"""Per-step device time of the resident bench step (diagnostic): python scripts/step_trace.py [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from odwscl_b200 import capi
from odwscl_b200.config import cfg
from odwscl_b200.modeling import build_detection_model
from odwscl_b200.structures import BoxList
from odwscl_b200.synth import synth_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_detection_model(cfg).to(dev).train()
opt = bench.make_optimizer(model)
ev = model.roi_heads.loss_evaluator; ev.speculative_k = True
images, rois, boxes, labels = synth_batch(2, 2000, 1000, 600, 21, seed=1234)
targets = []
for lab in labels:
    t = BoxList(torch.zeros((len(lab), 4)), (1000, 600), "xyxy"); t.add_field("labels", torch.as_tensor(lab)); targets.append(t)
images_d = images.to(dev); props = [BoxList(b.to(dev), (1000, 600), "xyxy") for b in boxes]
def step():
    losses, _ = model(images_d, targets, props)
    total = sum(losses.values())
    opt.zero_grad(set_to_none=True)
    total.backward()
    opt.found_inf = ev.overflow
    opt.step()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
torch.cuda.synchronize()
evs[0].record()
for i in range(n):
    step(); evs[i + 1].record()
torch.cuda.synchronize()
ms = [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(n)]
print("ms/step:", ms)
print("mem MB: allocated %.0f reserved %.0f" % (torch.cuda.memory_allocated() / 2**20, torch.cuda.memory_reserved() / 2**20),
      "num_alloc_retries", torch.cuda.memory_stats().get("num_alloc_retries"), "cap", ev._k_cap)
