"""Launch the WGRAD kernel a few times at the conv4_2 / conv5_x bench shape (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odwscl_b200 import capi                      # noqa: E402
x = torch.randn(2, 76, 128, 512, device="cuda")
dz = torch.randn(2, 76, 128, 512, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(6):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dw, db = capi.conv3x3_wgrad_nhwc(x, dz, dilation=2); b.record()
    torch.cuda.synchronize(); ts.append(round(a.elapsed_time(b), 4))
print("wgrad+bias_grad ms", ts, "TF/s", round(2 * 2 * 76 * 128 * 512 * 512 * 9 / (min(ts) * 1e-3) / 1e12, 1))
