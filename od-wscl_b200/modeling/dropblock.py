"""DropBlock2D (modeling/dropblock/drop_block.py:7-71) on the device: the centre mask is sampled
with the device generator (the reference samples on the CPU and copies, :42-45) and the block
mask / renormalisation / multiply run as one fused pass (csrc/dropblock.cu), forward and backward."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import capi


class _DropBlockFn(Function):
    @staticmethod
    def forward(ctx, x, centres, block, n_valid=None):
        y, scale_io = capi.dropblock(x, centres, block, n_valid=n_valid)
        ctx.save_for_backward(centres, scale_io)
        ctx.block, ctx.n_valid = block, n_valid
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        centres, scale_io = ctx.saved_tensors
        gx, _ = capi.dropblock(gy.contiguous(), centres, ctx.block, scale_io, n_valid=ctx.n_valid)
        return gx, None, None, None


class _DropBlockSegFn(Function):
    """DropBlock over P row segments with one renormalisation per segment (csrc/dropblock.cu, segmented variant)."""

    @staticmethod
    def forward(ctx, x, centres, block, seg_off, P):
        y, scale_seg = capi.dropblock_seg(x, centres, block, seg_off, P)
        ctx.save_for_backward(centres, seg_off, scale_seg)
        ctx.block, ctx.P = block, P
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        centres, seg_off, scale_seg = ctx.saved_tensors
        gx, _ = capi.dropblock_seg(gy.contiguous(), centres, ctx.block, seg_off, ctx.P, scale_seg)
        return gx, None, None, None, None


class DropBlock2D(nn.Module):
    def __init__(self, drop_prob, block_size):
        super().__init__()
        self.drop_prob = drop_prob
        self.block_size = block_size
        self.centre_sampler = None     # test hook: callable(n, h, w, gamma, device) -> float mask

    def forward(self, x, n_valid=None, seg_off=None):
        """`n_valid` (int32 device tensor [1], optional): x is a padded batch whose first n_valid rows are real.
        `seg_off` (int32 device tensor [P+1], optional): x is a batch of P row segments, each renormalised on its own
        (= P separate calls of the reference module); rows past seg_off[P] are padding."""
        assert x.dim() == 4, "Expected input with 4 dimensions (bsize, channels, height, width)"
        if not self.training or self.drop_prob == 0.0:
            return x
        gamma = self.drop_prob / (self.block_size ** 2)                 # :69-70
        n, _, h, w = x.shape
        if n == 0:
            return x
        if self.centre_sampler is not None:
            centres = self.centre_sampler(n, h, w, gamma, x.device)
        else:
            centres = (torch.rand(n, h, w, device=x.device) < gamma).float()   # :42
        if seg_off is not None:
            return _DropBlockSegFn.apply(x.contiguous(), centres.contiguous(), self.block_size, seg_off,
                                         seg_off.numel() - 1)
        return _DropBlockFn.apply(x.contiguous(), centres.contiguous(), self.block_size, n_valid)
