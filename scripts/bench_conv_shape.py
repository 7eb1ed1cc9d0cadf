"""Time the conv kernel at an arbitrary shape: python scripts/bench_conv_shape.py B H W Cin Cout dil [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odwscl_b200 import capi
B, H, W, Cin, Cout, dil = [int(v) for v in sys.argv[1:7]]
x = torch.randn(B, H, W, Cin, device="cuda"); w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05; b = torch.randn(Cout, device="cuda")
wk = w.permute(0, 2, 3, 1).contiguous(); y = torch.empty(B, H, W, Cout, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(cold):
    ts = []
    for _ in range(12):
        if cold: flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); capi.conv3x3_nhwc(x, wk, b, dilation=dil, flags=capi.CONV_RELU, out=y); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]
fl = 2.0 * B * H * W * Cout * 9 * Cin
for cold in (True, False):
    t = run(cold)
    print(json.dumps({"shape": [B, H, W, Cin, Cout, dil], "cold_l2": cold, "ms": round(t, 4), "tflops": round(fl / t / 1e9, 1)}))
