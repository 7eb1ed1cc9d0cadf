"""Per-shape timing of the fc block's GEMMs (csrc/fc_gemm.cu) next to cuBLAS TF32 (torch.matmul) on the same shapes.
GPU box only.  Usage: python scripts/bench_fc.py [R] [Kc]   (R proposals per rank step, Kc padded positives)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from odwscl_b200 import capi

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
KC = int(sys.argv[2]) if len(sys.argv) > 2 else 384
ONLY = set(sys.argv[3].split(",")) if len(sys.argv) > 3 else None
ITERS = int(os.environ.get("FC_ITERS", "5"))
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=ITERS):
    fn(); fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def case(name, M, N, K, a_mn, b_mn):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    out = torch.empty((M, (N + 3) // 4 * 4), device=dev)[:, :N]
    ours = timeit(lambda: capi.fc_gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out))
    At = A.t() if a_mn else A
    Bt = B if b_mn else B.t()
    ref = timeit(lambda: torch.matmul(At, Bt))
    fl = 2.0 * M * N * K
    print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "a_mn": a_mn, "b_mn": b_mn, "ms": round(ours, 4),
                      "tflops": round(fl / ours / 1e9, 1), "cublas_ms": round(ref, 4), "cublas_tflops": round(fl / ref / 1e9, 1)}),
          flush=True)
    return ours, ref


tot_o = tot_r = 0.0
for name, M, N, K, a, b in [
    ("fc6.fwd", 2 * R, 4096, 25088, 0, 0), ("fc7.fwd", 2 * R, 4096, 4096, 0, 0), ("sim0.fwd", R, 4096, 4096, 0, 0),
    ("sim2.fwd", R, 128, 4096, 0, 0), ("pred.fwd", R, 357, 4096, 0, 0),
    ("fc6s.fwd", 2 * KC, 4096, 25088, 0, 0), ("fc7s.fwd", 2 * KC, 4096, 4096, 0, 0), ("sim0s.fwd", 2 * KC, 4096, 4096, 0, 0),
    ("sim2s.fwd", 2 * KC, 128, 4096, 0, 0),
    ("pred.dgrad", R, 4096, 357, 0, 1), ("sim2.dgrad", R, 4096, 128, 0, 1), ("sim0.dgrad", R, 4096, 4096, 0, 1),
    ("fc7.dgrad", 2 * R, 4096, 4096, 0, 1), ("fc6.dgrad", 2 * R, 25088, 4096, 0, 1),
    ("sim2s.dgrad", 2 * KC, 4096, 128, 0, 1), ("sim0s.dgrad", 2 * KC, 4096, 4096, 0, 1), ("fc7s.dgrad", 2 * KC, 4096, 4096, 0, 1),
    ("fc6s.dgrad", 2 * KC, 25088, 4096, 0, 1),
    ("pred.wgrad", 357, 4096, R, 1, 1), ("sim2.wgrad", 128, 4096, R, 1, 1), ("sim0.wgrad", 4096, 4096, R, 1, 1),
    ("fc7.wgrad", 4096, 4096, 2 * R, 1, 1), ("fc6.wgrad", 4096, 25088, 2 * R, 1, 1),
    ("sim2s.wgrad", 128, 4096, 2 * KC, 1, 1), ("sim0s.wgrad", 4096, 4096, 2 * KC, 1, 1), ("fc7s.wgrad", 4096, 4096, 2 * KC, 1, 1),
    ("fc6s.wgrad", 4096, 25088, 2 * KC, 1, 1),
]:
    if ONLY is not None and name not in ONLY:
        continue
    o, r = case(name, M, N, K, bool(a), bool(b))
    tot_o += o; tot_r += r
print(json.dumps({"total_ms_ours": round(tot_o, 3), "total_ms_cublas": round(tot_r, 3)}))
