"""Launch the ROIPool forward/backward a few times at the bench shape (for ncu captures), channels-last like the model."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odwscl_b200 import capi                      # noqa: E402
from odwscl_b200.synth import synth_batch         # noqa: E402

_, rois, _, _ = synth_batch(2, 2000, 1000, 600, seed=1234)
feat = torch.randn(2, 512, 76, 128, device="cuda").contiguous(memory_format=torch.channels_last)
rois = rois.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tf, tb = [], []
for it in range(6):
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out, arg = capi.roi_pool_forward(feat, rois, 0.125, 7, 7)
    e[1].record()
    g = capi.roi_pool_backward(out, rois, arg, 7, 7, 2, 512, 76, 128, channels_last=True)
    e[2].record()
    torch.cuda.synchronize()
    tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
print("fwd ms", [round(t, 4) for t in tf], "bwd ms", [round(t, 4) for t in tb])
print("done", float(out.sum()), float(g.sum()))
# the floor of a "stage the roi region once per (roi, slab)" design: the same bytes, each cell once, a plain max
tp = []
for it in range(6):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pr = capi.probe_roi_stream(feat, rois, 0.125); b.record()
    torch.cuda.synchronize()
    tp.append(a.elapsed_time(b))
import numpy as np
bx = rois[:, 1:].cpu().numpy()
q = lambda v: np.floor(v * 0.125 + 0.5)
cells = float(((q(bx[:, 2]) - q(bx[:, 0]) + 1).clip(1) * (q(bx[:, 3]) - q(bx[:, 1]) + 1).clip(1)).sum())
gb = cells * 512 * 4 / 1e9
print("stream probe ms", [round(t, 4) for t in tp], "bytes once per roi: %.2f GB -> %.2f TB/s" % (gb, gb / (min(tp) * 1e-3) / 1e3))
