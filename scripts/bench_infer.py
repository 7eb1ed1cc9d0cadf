"""Test-time throughput of the path: eval-mode GeneralizedRCNN (conv stack -> ROIPool -> fc6/fc7 -> MIST heads ->
decode + all-class NMS + detections cap) on the bench workload.  python scripts/bench_infer.py [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odwscl_b200 import capi
from odwscl_b200.config import cfg
from odwscl_b200.modeling import build_detection_model
from odwscl_b200.structures import BoxList
from odwscl_b200.synth import synth_batch
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_detection_model(cfg).to(dev).eval()
images, rois, boxes, _ = synth_batch(2, 2000, 1000, 600, 21, seed=1234, pin=True)
images_d = images.to(dev); props = [BoxList(b.to(dev), (1000, 600), "xyxy") for b in boxes]
with torch.no_grad():
    for _ in range(5):
        res = model(images_d, None, props)
    torch.cuda.synchronize()
    l0 = capi.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        res = model(images_d, None, props)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(json.dumps({"metric": "test-time proposals/sec (2000 ROIs/img, 1000x600, single scale, detections out)", "value": 4000 / (ms * 1e-3),
                  "unit": "proposals/s", "ms_per_step": ms, "n_gpus": 1, "steps": steps, "detections_per_image": [len(r) for r in res],
                  "gpu_launches_per_step": (capi.launch_count - l0) / steps}))
