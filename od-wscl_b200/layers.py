"""Mirror of ``wetectron.layers`` for the hot path: ROIPool / ROIAlign autograd wrappers
(layers/roi_pool.py:11-57, layers/roi_align.py:11-60), nms (layers/nms.py:6), smooth_l1_loss
(layers/smooth_l1_loss.py:4-16).  Same class names, ctor arguments and repr."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import _C, capi


class _ROIPool(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale):
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.input_shape = input.size()
        ctx.channels_last = capi._is_nhwc(input)       # conv stack output: grad goes back in the same layout
        output, argmax = _C.roi_pool_forward(input, roi, spatial_scale, ctx.output_size[0], ctx.output_size[1])
        ctx.save_for_backward(roi, argmax)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, argmax = ctx.saved_tensors
        bs, ch, h, w = ctx.input_shape
        if ctx.channels_last:
            grad_input = capi.roi_pool_backward(grad_output, rois, argmax, ctx.output_size[0], ctx.output_size[1],
                                                bs, ch, h, w, channels_last=True)
        else:
            grad_input = _C.roi_pool_backward(grad_output, None, rois, argmax, ctx.spatial_scale, ctx.output_size[0],
                                              ctx.output_size[1], bs, ch, h, w)
        return grad_input, None, None, None


roi_pool = _ROIPool.apply


class _PoolAugFn(Function):
    """ROIPool + DropBlock of the pooled features into ONE [2R,C,7,7] buffer (rows [0,R) clean, [R,2R) augmented), so
    fc6/fc7 run once over both (weak_head.py:107-113: weights read once, one WGRAD, no gradient accumulation pass) and
    the backward scatters every consumer's gradient in one ROIPool-backward launch: the clean half, the DropBlock
    backward of the augmented half, and the rows the contrastive branch gathered (stashed by _GatherRowsFn)."""

    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, centres, block, stash):
        ph, pw = _pair(output_size)
        R, C = roi.shape[0], input.shape[1]
        buf = torch.empty((2 * R, C, ph, pw), dtype=torch.float32, device=input.device)
        bmask = None
        if (ph, pw) == (7, 7) and capi._is_nhwc(input) and C % 4 == 0 and spatial_scale > 0:
            # one pass: the pooling kernel writes the DropBlock-augmented copy from the values it has staged
            argmax, scale_io, bmask = capi.roi_pool_forward_aug(input, roi, spatial_scale, centres, block, buf)
        else:
            _, argmax = capi.roi_pool_forward(input, roi, spatial_scale, ph, pw, out=buf[:R])
            _, scale_io = capi.dropblock(buf[:R], centres, block, out=buf[R:])
        ctx.bmask = bmask
        ctx.save_for_backward(roi, argmax, centres, scale_io)
        ctx.block, ctx.stash, ctx.input_shape, ctx.R = block, stash, input.size(), R
        ctx.channels_last = capi._is_nhwc(input)
        ctx.out_hw = (ph, pw)
        return buf

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        rois, argmax, centres, scale_io = ctx.saved_tensors
        R = ctx.R
        bs, ch, h, w = ctx.input_shape
        g = g.contiguous()
        srows, sgrad = ctx.stash.pop("rows", None), ctx.stash.pop("grads", None)
        gin = None
        if ctx.channels_last and ctx.out_hw == (7, 7):
            # the DropBlock backward of the augmented half (g[R:] * block_mask * scale) happens inside the scatter
            bmask = ctx.bmask if ctx.bmask is not None else capi.dropblock_mask(centres, ctx.block, scale_io)
            gin = capi.roi_pool_backward_multi(g[:R], g[R:], srows, sgrad, rois, argmax, bs, ch, h, w, mask2=bmask)
        if gin is None:                                                          # maps too large for the plane kernel
            g_aug, _ = capi.dropblock(g[R:], centres, ctx.block, scale_io)
            tot = g[:R] + g_aug
            if srows is not None and srows.numel() > 0:
                tot.index_add_(0, srows, sgrad)
            gin = capi.roi_pool_backward(tot, rois, argmax, ctx.out_hw[0], ctx.out_hw[1], bs, ch, h, w,
                                         channels_last=ctx.channels_last)
        return gin, None, None, None, None, None, None


class _GatherRowsFn(Function):
    """rows of the clean half of a _PoolAugFn buffer; the gradient is handed to that node through `stash` instead of
    being expanded to a dense zero-filled [R,C,7,7] tensor (the graph edge to `buf` keeps the execution order)."""

    @staticmethod
    def forward(ctx, buf, rows, R, stash):
        ctx.stash = stash
        ctx.save_for_backward(rows)
        return buf[:R].index_select(0, rows)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (rows,) = ctx.saved_tensors
        if "rows" in ctx.stash:                                                  # a second gather of the same step
            ctx.stash["rows"] = torch.cat([ctx.stash["rows"], rows])
            ctx.stash["grads"] = torch.cat([ctx.stash["grads"], g.contiguous()])
        else:
            ctx.stash["rows"], ctx.stash["grads"] = rows, g.contiguous()
        return None, None, None, None


class _AugPositivesFn(Function):
    """The DropBlock and the noise views of the Phase-A positives (loss.py:296-305) straight from the clean half of a
    _PoolAugFn buffer: one kernel instead of gather + DropBlock + randn + mul + add + cat; the gradient w.r.t. the gathered
    rows is handed to the ROIPool backward through `stash` (as _GatherRowsFn does)."""

    @staticmethod
    def forward(ctx, buf, rows, R, stash, seg_off, P, centres, block, noise, seed):
        out, scale_seg = capi.aug_positives(buf[:R], rows, seg_off, P, centres, block, noise=noise, seed=seed)
        ctx.save_for_backward(rows, seg_off, centres, scale_seg, noise if noise is not None else rows.new_zeros(0))
        ctx.stash, ctx.P, ctx.block, ctx.seed, ctx.has_noise = stash, P, block, seed, noise is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        rows, seg_off, centres, scale_seg, noise = ctx.saved_tensors
        gx, _ = capi.aug_positives(g.contiguous(), rows, seg_off, ctx.P, centres, ctx.block, scale_seg=scale_seg,
                                   noise=noise if ctx.has_noise else None, seed=ctx.seed, backward=True)
        if "rows" in ctx.stash:
            ctx.stash["rows"] = torch.cat([ctx.stash["rows"], rows])
            ctx.stash["grads"] = torch.cat([ctx.stash["grads"], gx])
        else:
            ctx.stash["rows"], ctx.stash["grads"] = rows, gx
        return (None,) * 10


aug_positives = _AugPositivesFn.apply


class _SplitRowsFn(Function):
    """x[:R], x[R:] as two outputs whose backward is ONE concatenation (instead of two zero-padded slices + an add)."""

    @staticmethod
    def forward(ctx, x, R):
        ctx.R, ctx.shape = R, x.shape
        return x[:R], x[R:]

    @staticmethod
    @once_differentiable
    def backward(ctx, g1, g2):
        out = torch.empty(ctx.shape, dtype=g1.dtype if g1 is not None else g2.dtype,
                          device=g1.device if g1 is not None else g2.device)
        if g1 is None:
            out[:ctx.R].zero_()
        else:
            out[:ctx.R].copy_(g1)
        if g2 is None:
            out[ctx.R:].zero_()
        else:
            out[ctx.R:].copy_(g2)
        return out, None


pool_and_augment = _PoolAugFn.apply
gather_rows = _GatherRowsFn.apply
split_rows = _SplitRowsFn.apply


class ROIPool(nn.Module):
    def __init__(self, output_size, spatial_scale):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale

    def forward(self, input, rois):
        return roi_pool(input.float(), rois.float(), self.output_size, self.spatial_scale)   # amp.float_function

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s)" % (self.__class__.__name__, self.output_size, self.spatial_scale)


class _ROIAlign(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, sampling_ratio):
        ctx.save_for_backward(roi)
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.sampling_ratio = sampling_ratio
        ctx.input_shape = input.size()
        ctx.channels_last = capi._is_nhwc(input)      # conv stack output: the channels-last kernels, grad in the same layout
        return _C.roi_align_forward(input, roi, spatial_scale, ctx.output_size[0], ctx.output_size[1], sampling_ratio)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        bs, ch, h, w = ctx.input_shape
        if ctx.channels_last:
            grad_input = capi.roi_align_backward(grad_output, rois, ctx.spatial_scale, ctx.output_size[0], ctx.output_size[1],
                                                 bs, ch, h, w, ctx.sampling_ratio, channels_last=True)
        else:
            grad_input = _C.roi_align_backward(grad_output, rois, ctx.spatial_scale, ctx.output_size[0],
                                               ctx.output_size[1], bs, ch, h, w, ctx.sampling_ratio)
        return grad_input, None, None, None, None


roi_align = _ROIAlign.apply


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        return roi_align(input.float(), rois.float(), self.output_size, self.spatial_scale, self.sampling_ratio)

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s, sampling_ratio=%s)" % (
            self.__class__.__name__, self.output_size, self.spatial_scale, self.sampling_ratio)


nms = _C.nms


def smooth_l1_loss(input, target, beta=1. / 9, reduction=True):
    """layers/smooth_l1_loss.py:4-16."""
    n = torch.abs(input - target)
    loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction:
        return loss.sum()
    return loss
