"""Launch the ROIPool forward/backward a few times at the bench shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odwscl_b200 import capi                      # noqa: E402
from odwscl_b200.synth import synth_batch         # noqa: E402

_, rois, _, _ = synth_batch(2, 2000, 1000, 600, seed=1234)
feat = torch.randn(2, 512, 76, 128, device="cuda")
rois = rois.cuda()
for _ in range(3):
    out, arg = capi.roi_pool_forward(feat, rois, 0.125, 7, 7)
    g = capi.roi_pool_backward(out, rois, arg, 7, 7, 2, 512, 76, 128)
torch.cuda.synchronize()
print("done", float(out.sum()), float(g.sum()))
