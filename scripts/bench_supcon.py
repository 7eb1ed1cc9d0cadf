"""SupCon forward + backward alone at the bank sizes of configs[1] (M ~ 1.1 k) and configs[2] (M = 4-6 k): the fused FFMA
tile kernels (csrc/supcon.cu) against the tensor-core path (csrc/supcon_tc.cu).  CUDA events, L2 flushed between launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odwscl_b200 import capi                      # noqa: E402
from odwscl_b200.modeling import sim_head         # noqa: E402

import argparse                                   # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, nargs="*", default=[1100, 2048, 4096, 6000])
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--paths", nargs="*", default=["tiles", "tensor"])       # for ncu captures: --M 6000 --iters 3 --paths tensor
args = ap.parse_args()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for M in args.M:
    g = torch.Generator(device="cuda").manual_seed(M)
    Fm = torch.nn.functional.normalize(torch.randn(M, 128, device="cuda", generator=g), dim=1).requires_grad_(True)
    E = torch.zeros(1, 128, device="cuda")
    src = torch.arange(M, dtype=torch.int32, device="cuda")
    lab = torch.randint(0, 20, (M,), device="cuda", generator=g).int()
    w = torch.rand(M, device="cuda", generator=g)
    Md = torch.full((1,), M, dtype=torch.int32, device="cuda")
    out = {}
    for name, thr in (("tiles", 1 << 30), ("tensor", 0)):
        if name not in args.paths:
            out[name] = (float("nan"),) * 4
            continue
        sim_head.SUPCON_TC_MIN_ROWS = thr
        ts = []
        for it in range(args.iters):
            flush.zero_()
            Fm.grad = None
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            loss = sim_head.supcon_bank_loss(Fm, E, src, lab, w, Md, M, 0.2)
            b.record()
            loss.backward()
            c.record()
            torch.cuda.synchronize()
            ts.append((a.elapsed_time(b), b.elapsed_time(c)))
        out[name] = (min(t[0] for t in ts[-4:]), min(t[1] for t in ts[-4:]), float(loss), float(Fm.grad.abs().max()))
    f = 2.0 * M * M * 128
    print("M %5d  tiles fwd %.3f bwd %.3f ms | tensor fwd %.3f bwd %.3f ms | loss %.6f / %.6f  max|grad| %.3e / %.3e  "
          "(2 M^2 128 = %.1f GFLOP per contraction)" % (M, out["tiles"][0], out["tiles"][1], out["tensor"][0], out["tensor"][1],
                                                        out["tiles"][2], out["tensor"][2], out["tiles"][3], out["tensor"][3], f / 1e9))
