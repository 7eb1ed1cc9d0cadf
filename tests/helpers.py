"""Shared test helpers: rebuild the tiny stand-in modules of tests/golden/roi_reg_loss.npz."""
import numpy as np
import torch
import torch.nn.functional as F


def grid_sim(x, w, b):
    """oracle/gen_golden.py::GridSim (outputs snapped to the exact 1/16 grid, straight-through)."""
    y = F.normalize(F.linear(x, w, b), dim=1) * 4.0
    q = torch.clamp(torch.round(y * 4) / 16, -0.5, 0.5)
    return y / 4.0 + (q - y / 4.0).detach()


def case_tensors(G, tag, device="cpu", grad=True):
    g = lambda k: torch.from_numpy(G[tag + "_" + k]).to(device)
    sizes = [int(s) for s in G[tag + "_sizes"]]
    t = dict(sizes=sizes)
    for k in ("cls", "det", "simf", "pooled", "ref0", "ref1", "ref2", "bb0", "bb1", "bb2",
              "fe_w", "fe_b", "ms_w", "ms_b"):
        t[k] = g(k).clone().requires_grad_(grad)
    t["boxes"] = [g("boxes%d" % b) for b in range(len(sizes))]
    t["labels"] = [G[tag + "_labels%d" % b] for b in range(len(sizes))]
    t["seed"] = int(G[tag + "_rng_seed"])
    return t
