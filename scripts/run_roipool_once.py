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
