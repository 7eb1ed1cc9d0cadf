"""Mirror of the reference's pybind module ``wetectron._C`` (csrc/vision.cpp:9-24): same 14 names,
same argument order.  The five functions on the hot path run the sm_100a kernels through the C
ABI; the nine RetinaNet / DCN exports (out of scope) exist so that
``wetectron/layers/__init__.py:16-20`` imports, and raise when called."""
import torch

from . import capi


def roi_pool_forward(input, rois, spatial_scale, pooled_height, pooled_width):
    """csrc/ROIPool.h:11-24 -> (output [R,C,ph,pw] fp32, argmax [R,C,ph,pw] int32)."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")          # csrc/ROIPool.h:23
    return capi.roi_pool_forward(input, rois, spatial_scale, pooled_height, pooled_width)


def roi_pool_backward(grad, input, rois, argmax, spatial_scale, pooled_height, pooled_width,
                      batch_size, channels, height, width):
    """csrc/ROIPool.h:26-45 -> grad_input [B,C,H,W]."""
    if not grad.is_cuda:
        raise RuntimeError("Not implemented on the CPU")          # csrc/ROIPool.h:44
    return capi.roi_pool_backward(grad, rois, argmax, pooled_height, pooled_width, batch_size, channels,
                                  height, width)


def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    """csrc/ROIAlign.h:11-26."""
    return capi.roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio)


def roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height,
                       width, sampling_ratio):
    """csrc/ROIAlign.h:28-45."""
    return capi.roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels,
                                   height, width, sampling_ratio)


def nms(dets, scores, threshold):
    """csrc/nms.h:10-28: legacy NMS (+1 convention), kept indices ascending; empty guard :17-18."""
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.long, device=dets.device)
    return capi.nms_legacy(dets, scores, threshold)


def _out_of_scope(name):
    def fn(*a, **k):
        raise RuntimeError("wetectron._C.%s is outside the proposal-feature hot path and is not provided "
                           "by odwscl_b200 (SURVEY.md 2.2 K8-K10)" % name)
    fn.__name__ = name
    return fn


for _n in ("sigmoid_focalloss_forward", "sigmoid_focalloss_backward", "deform_conv_forward",
           "deform_conv_backward_input", "deform_conv_backward_parameters", "modulated_deform_conv_forward",
           "modulated_deform_conv_backward", "deform_psroi_pooling_forward", "deform_psroi_pooling_backward"):
    globals()[_n] = _out_of_scope(_n)
