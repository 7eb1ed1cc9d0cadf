"""Sim_Net (roi_heads/sim_head/sim_net.py:7-26) and SupConLossV2 (sim_head/sim_loss.py:44-80).
State-dict keys: roi_heads.model_sim.mlp.{0,2}.  The loss runs as the fused kernels of
csrc/supcon.cu (forward and backward), fed by a row-id list into [F ; E] instead of a concatenated
bank."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import capi


class _L2NormFn(Function):
    """F.normalize(z, dim=1) (sim_net.py:26) as one kernel forward, one backward (csrc/elementwise.cu)."""

    @staticmethod
    def forward(ctx, z):
        y, inv = capi.l2norm_forward(z)
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, inv = ctx.saved_tensors
        return capi.l2norm_backward(y, g.contiguous(), inv)


class Sim_Net(nn.Module):
    def __init__(self, config, in_dim):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_dim, in_dim), nn.ReLU(inplace=True), nn.Linear(in_dim, 128))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)

        self.strict_fp32 = False       # True: 3xTF32 split products (parity tests)

    def forward(self, roi_feat, in_mask_scale=None, role=None):
        """in_mask_scale: roi_feat comes from run_classifier(..., fuse_out_bwd=True) and this is its only consumer.
        role: "main" (all proposals, weak_head.py:110) / "small" (augmented positives, loss.py:301,305) -- the weight
        gradients of the two calls of a training step are produced as one tensor per layer (fc._LinearFn)."""
        from . import fc
        fuse = fc.FUSE_ACT_BWD and torch.is_grad_enabled()
        if role == "main":
            self._stash = {"mlp0": {}, "mlp2": {}}
        st = getattr(self, "_stash", None) if role is not None else None
        h = fc.linear(roi_feat, self.mlp[0].weight, self.mlp[0].bias, act=fc.ACT_RELU, round_out=True,
                      strict=self.strict_fp32, in_mask_scale=in_mask_scale, act_bwd_fused=fuse,
                      stash=st["mlp0"] if st is not None else None, role=role if st is not None else None)
        return _L2NormFn.apply(fc.linear(h, self.mlp[2].weight, self.mlp[2].bias, strict=self.strict_fp32,
                                         in_mask_scale=1.0 if fuse else None, stash=st["mlp2"] if st is not None else None,
                                         role=role if st is not None else None))


# Banks of at least this many rows (bound Mcap) take the tensor-core path (csrc/supcon_tc.cu: both M x M contractions
# as 3xTF32 tcgen05 GEMMs); smaller ones the fused FFMA tile kernels (csrc/supcon.cu: 3 launches, nothing in HBM).
SUPCON_TC_MIN_ROWS = int(os.environ.get("ODWSCL_SUPCON_TC_MIN", "3072"))   # measured crossover: profiles/r02_supcon.txt


class _SupConBankFn(Function):
    """loss = mean_r( -log(pos_r / all_r) * w_r ) over the bank rows V[row_src], V = [Fm ; E]."""

    @staticmethod
    def forward(ctx, Fm, E, row_src, row_lab, row_w, M_dev, Mcap, inv_temp, tc):
        Fm, E = Fm.contiguous(), E.contiguous()
        ctx.tc = (Mcap >= SUPCON_TC_MIN_ROWS) if tc is None else bool(tc)
        if ctx.tc:
            loss, stats, ws = capi.supcon_tc_forward(Fm, E, row_src, row_lab, row_w, M_dev, Mcap, inv_temp)
            ctx.ws = ws                                        # S and the split bank rows: consumed by the one backward
        else:
            loss, stats = capi.supcon_forward(Fm, E, row_src, row_lab, row_w, M_dev, Mcap, inv_temp)
        ctx.save_for_backward(Fm, E, row_src, row_lab, row_w, M_dev, stats)
        ctx.Mcap, ctx.inv_temp = Mcap, inv_temp
        return loss.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        Fm, E, row_src, row_lab, row_w, M_dev, stats = ctx.saved_tensors
        g = g.contiguous().view(1).float()
        if ctx.tc:
            ws, ctx.ws = ctx.ws, None                          # H overwrites S: a second backward would need a new forward
            if ws is None:
                raise RuntimeError("SupCon (tensor-core path): backward called twice on the same graph")
            dF, dE = capi.supcon_tc_backward(Fm, E, row_src, row_lab, row_w, M_dev, ctx.Mcap, ctx.inv_temp, stats, g, ws)
        else:
            dF, dE = capi.supcon_backward(Fm, E, row_src, row_lab, row_w, M_dev, ctx.Mcap, ctx.inv_temp, stats, g)
        return dF, dE, None, None, None, None, None, None, None


def supcon_bank_loss(Fm, E, row_src, row_lab, row_w, M_dev, Mcap, temperature, tc=None):
    """tc: None = by size (Mcap >= SUPCON_TC_MIN_ROWS: right when Mcap is the row count or a tight bound on it); the
    tensor-core path works on all Mcap x Mcap entries, so a caller with a loose bound decides itself (weak_head/loss.py)."""
    return _SupConBankFn.apply(Fm, E, row_src, row_lab, row_w, M_dev, Mcap, 1.0 / temperature, tc)


class SupConLossV2(nn.Module):
    """Drop-in for sim_loss.py:44-80: forward(overlaps_enc: list of [n_c,128] per class, score_col, device)."""

    def __init__(self, temperature=0.2):
        super().__init__()
        self.temperature = temperature

    def forward(self, overlaps_enc, score_col, device=None):
        feats, labs = [], []
        for i, e in enumerate(overlaps_enc):
            if e.shape[0] != 0:
                feats.append(e)
                labs.append(torch.full((e.shape[0],), i, dtype=torch.int32, device=e.device))
        feats, labs = torch.cat(feats), torch.cat(labs)
        M = feats.shape[0]
        src = torch.arange(M, dtype=torch.int32, device=feats.device)
        M_dev = torch.full((1,), M, dtype=torch.int32, device=feats.device)
        E = feats.new_zeros((1, 128))
        return supcon_bank_loss(feats, E, src, labs, score_col.detach().float().contiguous(), M_dev, M,
                                self.temperature)
