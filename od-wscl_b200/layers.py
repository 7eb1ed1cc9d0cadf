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
        return _C.roi_align_forward(input, roi, spatial_scale, ctx.output_size[0], ctx.output_size[1], sampling_ratio)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        bs, ch, h, w = ctx.input_shape
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
