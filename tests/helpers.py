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


class KeyedSource:
    """Replayable stochastic layers whose draws for the contrastive branch depend ONLY on the proposal's global row
    id: any batching of the augmented positives (per (image, class) call as in the reference, one batch, a padded
    speculative batch) sees the same DropBlock centres and the same noise for the same proposal.  Same interface
    as oracle.StochasticSource plus *_rows variants; Dropout is the identity (its Philox stream on the device has
    no CPU counterpart)."""

    def __init__(self, seed, drop_prob=0.3):
        self.seed, self.drop_prob = int(seed), drop_prob
        self.g = torch.Generator().manual_seed(self.seed)
        self._c, self._n = {}, {}

    def dropblock_centres(self, n, block):                       # the [R,7,7] DropBlock of weak_head.py:111
        gamma = self.drop_prob / (block ** 2)
        return (torch.rand(n, 7, 7, generator=self.g) < gamma).float()

    def dropout(self, x):
        return x

    def _gen(self, r, k):
        return torch.Generator().manual_seed(self.seed * 1000003 + 2 * int(r) + k)

    def dropblock_centres_rows(self, rows, block):
        gamma = self.drop_prob / (block ** 2)
        out = []
        for r in np.asarray(rows).reshape(-1).tolist():
            if r not in self._c:
                self._c[r] = (torch.rand(7, 7, generator=self._gen(r, 0)) < gamma).float()
            out.append(self._c[r])
        return torch.stack(out) if out else torch.zeros((0, 7, 7))

    def noise_rows(self, rows, shape):
        out = []
        for r in np.asarray(rows).reshape(-1).tolist():
            if r not in self._n:
                self._n[r] = torch.randn(tuple(shape[1:]), generator=self._gen(r, 1))
            out.append(self._n[r])
        return torch.stack(out) if out else torch.zeros(tuple(shape))
