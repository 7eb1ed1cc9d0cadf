"""Per-layer timing of the tcgen05 conv kernel vs cuDNN (TF32, channels_last) at the bench shape.
usage (GPU box): python scripts/bench_conv.py [B]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from odwscl_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
layers = [("conv1_2", 608, 1024, 64, 64, 1), ("conv2_1", 304, 512, 64, 128, 1), ("conv2_2", 304, 512, 128, 128, 1),
          ("conv3_1", 152, 256, 128, 256, 1), ("conv3_2", 152, 256, 256, 256, 1), ("conv4_1", 76, 128, 256, 512, 1),
          ("conv4_2", 76, 128, 512, 512, 1), ("conv5_1", 76, 128, 512, 512, 2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


for name, H, W, Cin, Cout, dil in layers:
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    b = torch.randn(Cout, device="cuda")
    wk = w.permute(0, 2, 3, 1).contiguous()
    y = torch.empty(B, H, W, Cout, device="cuda")
    xcl = x.permute(0, 3, 1, 2)          # NCHW view with channels_last strides
    wcl = w.contiguous(memory_format=torch.channels_last)
    t_ours = timeit(lambda: capi.conv3x3_nhwc(x, wk, b, dilation=dil, flags=capi.CONV_RELU, out=y))
    t_cudnn = timeit(lambda: F.relu_(F.conv2d(xcl, wcl, b, padding=dil, dilation=dil)))
    fl = 2.0 * B * H * W * Cout * 9 * Cin
    print(json.dumps({"layer": name, "B": B, "HxW": [H, W], "Cin": Cin, "Cout": Cout, "gflop": fl / 1e9,
                      "ours_ms": round(t_ours, 4), "ours_tflops": round(fl / t_ours / 1e9, 1),
                      "cudnn_ms(conv+relu)": round(t_cudnn, 4), "cudnn_tflops": round(fl / t_cudnn / 1e9, 1)}))
