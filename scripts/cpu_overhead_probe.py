"""How much of the step is host launch overhead?  Same model / same launch sequence on a tiny workload (GPU time
negligible): ms/step here ~= CPU enqueue cost of one step.  usage (GPU box): python scripts/cpu_overhead_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odwscl_b200 import capi
from odwscl_b200.config import cfg
from odwscl_b200.modeling import build_detection_model
from odwscl_b200.structures import BoxList
from odwscl_b200.synth import synth_batch
import bench

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_detection_model(cfg).to(dev).train()
opt = bench.make_optimizer(model)
out = {}
for tag, (W, H, N) in {"tiny": (320, 256, 64), "bench": (1000, 600, 2000)}.items():
    images, rois, boxes, labels = synth_batch(2, N, W, H, 21, seed=1234)
    targets = []
    for lab in labels:
        t = BoxList(torch.zeros((len(lab), 4)), (W, H), "xyxy"); t.add_field("labels", torch.as_tensor(lab)); targets.append(t)
    images_d = images.to(dev); props = [BoxList(b.to(dev), (W, H), "xyxy") for b in boxes]
    def step():
        losses, _ = model(images_d, targets, props)
        total = sum(losses.values())
        opt.zero_grad(set_to_none=True)
        total.backward()
        opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    l0 = capi.launch_count
    t0 = time.perf_counter()
    for _ in range(10): step()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    out[tag] = {"cpu_enqueue_ms_per_step": t_cpu * 100, "wall_ms_per_step": t_all * 100, "capi_launches_per_step": (capi.launch_count - l0) / 10}
print(json.dumps(out))
