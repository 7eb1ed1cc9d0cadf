"""ctypes binding of libodwscl_sm100.so (include/odwscl.h).

PyTorch is plumbing here: it owns device memory and the current stream; every call passes raw
device pointers + sizes + the stream handle through the C ABI.  There is NO fallback: if the
library is missing or a call fails, a RuntimeError is raised (the reference raises through
AT_ASSERTM / THCudaCheck, csrc/cuda/ROIPool_cuda.cu:115-116,151).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_size_t, c_void_p

import time

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libodwscl_sm100.so")
_lib = None
launch_count = 0          # kernels-launching C-ABI calls made so far (bench.py reports it)

# kernels enqueued per entry point (for the `gpu_launches` bench key)
_LAUNCHES = {
    "odwscl_roi_pool_fwd_f32": 2, "odwscl_roi_pool_bwd_f32": 1, "odwscl_roi_pool_fwd_nhwc_f32": 1,
    "odwscl_roi_pool_bwd_nhwc_f32": 1, "odwscl_roi_pool_bwd_nhwc_multi_f32": 1, "odwscl_roi_pool_fwd_nhwc_aug_f32": 1,
    "odwscl_dropblock_prepare_f32": 3, "odwscl_roi_align_fwd_f32": 1,
    "odwscl_roi_align_bwd_f32": 1, "odwscl_roi_align_fwd_nhwc_f32": 1, "odwscl_roi_align_bwd_nhwc_f32": 1, "odwscl_box_iou_f32": 1, "odwscl_nms_f32": 1, "odwscl_nms_legacy_f32": 1,
    "odwscl_discover_phase_a_f32": 2, "odwscl_discover_phase_b_f32": 2, "odwscl_bank_assemble": 1,
    "odwscl_supcon_fwd_f32": 2, "odwscl_supcon_bwd_f32": 1, "odwscl_supcon_tc_fwd_f32": 4, "odwscl_supcon_tc_bwd_f32": 4, "odwscl_od_layer_f32": 1,
    "odwscl_dropblock_f32": 3, "odwscl_dropblock_rows_f32": 3, "odwscl_dropblock_seg_f32": 2, "odwscl_dropblock_mask_f32": 1, "odwscl_sim_nxn_f32": 2, "odwscl_gemm_nt_tf32": 1,
    "odwscl_conv3x3_nhwc_tf32": 1, "odwscl_conv3x3_wgrad_nhwc_tf32": 2, "odwscl_conv3x3_c3_f32": 1, "odwscl_maxpool2x2_nhwc_f32": 1,
    "odwscl_maxpool2x2_nhwc_bwd_f32": 1, "odwscl_split_tf32": 1,
    "odwscl_relu_dropout_fwd_f32": 1, "odwscl_relu_dropout_bwd_f32": 1, "odwscl_conv_weight_xform_f32": 1,
    "odwscl_fc_gemm_tf32": 1, "odwscl_fc_gemm_peer_sum_tf32": 1, "odwscl_peer_add_f32": 1, "odwscl_fc_gemm_peer_scatter_tf32": 1, "odwscl_peer_broadcast_f32": 1, "odwscl_colsum_f32": 1, "odwscl_l2norm_fwd_f32": 1, "odwscl_l2norm_bwd_f32": 1,
    "odwscl_spec_index": 1, "odwscl_aug_positives_f32": 2,
    "odwscl_head_scores_f32": 5, "odwscl_head_loss_f32": 2, "odwscl_head_grad_scale_f32": 1,
}

_P, _I, _F, _Z = c_void_p, c_int, c_float, c_size_t
_SIGS = {
    "odwscl_roi_pool_fwd_ws_bytes": (_Z, [_I] * 7),
    "odwscl_roi_pool_fwd_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _F, _I, _I, _P, _P, _P, _Z, _P]),
    "odwscl_roi_pool_bwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "odwscl_roi_pool_fwd_nhwc_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _F, _P, _P, _P]),
    "odwscl_roi_pool_bwd_nhwc_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "odwscl_roi_pool_fwd_nhwc_aug_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _F, _P, _P, _P, _P, _P]),
    "odwscl_dropblock_prepare_f32": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "odwscl_probe_roi_stream_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _F, _P, _P]),
    "odwscl_roi_pool_bwd_nhwc_multi_f32": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "odwscl_dropblock_mask_f32": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "odwscl_roi_align_fwd_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _F, _I, _I, _I, _P, _P]),
    "odwscl_roi_align_bwd_f32": (_I, [_P, _P, _I, _F, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "odwscl_roi_align_fwd_nhwc_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _F, _I, _P, _P]),
    "odwscl_roi_align_bwd_nhwc_f32": (_I, [_P, _P, _I, _F, _I, _I, _I, _I, _I, _P, _P]),
    "odwscl_box_iou_f32": (_I, [_P, _I, _P, _I, _I, _P, _P]),
    "odwscl_nms_f32": (_I, [_P, _P, _I, _F, _P, _P, _P]),
    "odwscl_nms_legacy_f32": (_I, [_P, _P, _I, _F, _P, _P, _P]),
    "odwscl_nms_per_class_f32": (_I, [_P, _P, _I, _I, _F, _F, _P, _P, _P]),
    "odwscl_nms_large_ws_bytes": (_Z, [_I]),
    "odwscl_nms_large_f32": (_I, [_P, _I, _P, _I, _I, _F, _F, _I, _P, _P, _P, _P, _Z, _P]),
    "odwscl_discover_phase_a_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _F] + [_P] * 7 + [_P]),
    "odwscl_discover_phase_b_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I] + [_P] * 8 + [_F] + [_P] * 7 + [_P]),
    "odwscl_bank_assemble": (_I, [_P, _P, _I, _I, _I, _I, _I] + [_P] * 8 + [_I] + [_P] * 4 + [_P]),
    "odwscl_supcon_fwd_f32": (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _F, _P, _P, _P]),
    "odwscl_supcon_bwd_f32": (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _F, _P, _P, _P, _P, _P]),
    "odwscl_supcon_tc_ws_bytes": (_Z, [_I]),
    "odwscl_supcon_tc_fwd_f32": (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _F, _P, _Z, _P, _P, _P]),
    "odwscl_supcon_tc_bwd_f32": (_I, [_I, _P, _P, _P, _P, _I, _F, _P, _Z, _P, _P, _P, _P, _P]),
    "odwscl_od_layer_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P, _P, _F, _P, _P, _P, _P]),
    "odwscl_dropblock_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "odwscl_dropblock_rows_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    "odwscl_dropblock_seg_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P]),
    "odwscl_sim_nxn_ws_bytes": (_Z, [_I]),
    "odwscl_sim_nxn_f32": (_I, [_P, _I, _P, _P, _Z, _P]),
    "odwscl_gemm_nt_tf32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "odwscl_conv3x3_nhwc_tf32": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P]),
    "odwscl_conv3x3_wgrad_nhwc_tf32": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "odwscl_conv3x3_c3_f32": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "odwscl_maxpool2x2_nhwc_f32": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "odwscl_maxpool2x2_nhwc_bwd_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "odwscl_split_tf32": (_I, [_P, ctypes.c_longlong, _P, _P, _P]),
    "odwscl_relu_dropout_fwd_f32": (_I, [_P, ctypes.c_longlong, _F, ctypes.c_ulonglong, _P]),
    "odwscl_relu_dropout_bwd_f32": (_I, [_P, _P, _P, ctypes.c_longlong, _F, _P]),
    "odwscl_conv_weight_xform_f32": (_I, [_P, _I, _I, _P, _P, _I, _P]),
    "odwscl_fc_gemm_tf32": (_I, [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _I, _F, _F, ctypes.c_ulonglong, _I,
                                 _P, _I, _P, _I, _I, _P]),
    "odwscl_fc_gemm_peer_sum_tf32": (_I, [_P, _I, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _I, _P, _I, _P, _I, _I, _P]),
    "odwscl_peer_add_f32": (_I, [_P, _P, ctypes.c_longlong, _F, _P]),
    "odwscl_fc_gemm_peer_scatter_tf32": (_I, [_P, _I, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P, _I, _P, _I,
                                              _I, _P]),
    "odwscl_peer_broadcast_f32": (_I, [_P, _P, ctypes.c_longlong, _P]),
    "odwscl_colsum_f32": (_I, [_P, ctypes.c_longlong, _I, _I, _P, _I, _P]),
    "odwscl_head_scores_ws_bytes": (_Z, [_I, _I]),
    "odwscl_head_scores_f32": (_I, [_P, _I, _I, _I, _I, _P, _I] + [_P] * 7 + [_P, _Z, _P]),
    "odwscl_head_loss_f32": (_I, [_P, _I, _I, _I, _I, _I, _P, _I] + [_P] * 8 + [_F, _P, _P, _P, _P]),
    "odwscl_head_grad_scale_f32": (_I, [_P, _I, ctypes.c_longlong, _I, _I, _P, _P]),
    "odwscl_l2norm_fwd_f32": (_I, [_P, _I, _I, _I, _F, _P, _P, _P]),
    "odwscl_l2norm_bwd_f32": (_I, [_P, _P, _P, _I, _I, _F, _P, _P]),
    "odwscl_spec_index": (_I, [_P, _P, _I, ctypes.c_longlong, _P, _P, _P, _P]),
    "odwscl_aug_positives_f32": (_I, [_P, _I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, ctypes.c_ulonglong, _I, _P, _P]),
    "odwscl_set_sm_margin": (_I, [_I]),
    "odwscl_version": (_I, []),
    "odwscl_strerror": (ctypes.c_char_p, [_I]),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Load the C-ABI library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libodwscl_sm100.so is missing (%s): build it with `python od-wscl_b200/csrc/build.py`; "
                "there is no CPU / eager fallback for the hot path" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def set_sm_margin(sms):
    """SMs the persistent conv / fc kernels leave to concurrent NCCL kernels (0 = none)."""
    rc = lib().odwscl_set_sm_margin(int(sms))
    if rc != 0:
        raise RuntimeError("odwscl_set_sm_margin(%d) failed" % sms)


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    return c_void_p(t.data_ptr())


# Live per-call timing (bench.py): set `profile = []` and every C-ABI call is bracketed by CUDA events recorded
# on the stream it launches on; entries are (name, (kind, amount) or None, ev0, ev1).  `amount` is the ALGORITHMIC
# work of the call: FLOP for the tensor-bound kernels, bytes for the HBM-bound ones.
profile = None
_WORK = {
    # (x, B, H, W, Cin, w, bias, Cout, ...)
    "odwscl_conv3x3_nhwc_tf32": lambda a: ("flop", 2.0 * a[1] * a[2] * a[3] * a[7] * 9 * a[4]),
    # (x, dz, B, H, W, Cin, Cout, ...)
    "odwscl_conv3x3_wgrad_nhwc_tf32": lambda a: ("flop", 2.0 * a[2] * a[3] * a[4] * a[5] * a[6] * 9),
    # (feat, B, C, H, W, rois, R, ...): map once + rois + out + int32 argmax
    "odwscl_roi_pool_fwd_nhwc_f32": lambda a: ("byte", 4.0 * a[1] * a[2] * a[3] * a[4] + 20.0 * a[6] + 8.0 * a[6] * a[2] * 49),
    "odwscl_roi_pool_fwd_f32": lambda a: ("byte", 4.0 * a[1] * a[2] * a[3] * a[4] + 20.0 * a[6] + 8.0 * a[6] * a[2] * a[8] * a[9]),
    # same contract figure for the variant that also writes the augmented copy (its extra 4*R*C*49 bytes replace a
    # separate pass and are not counted as algorithmic work of ROIPool)
    "odwscl_roi_pool_fwd_nhwc_aug_f32": lambda a: ("byte", 4.0 * a[1] * a[2] * a[3] * a[4] + 20.0 * a[6] + 8.0 * a[6] * a[2] * 49),
    # (grad, argmax, rois, R, B, C, H, W, ...): grad_out + argmax reads, zero + write of the map
    "odwscl_roi_pool_bwd_nhwc_f32": lambda a: ("byte", 8.0 * a[3] * a[5] * 49 + 8.0 * a[4] * a[5] * a[6] * a[7]),
    "odwscl_roi_pool_bwd_f32": lambda a: ("byte", 8.0 * a[3] * a[5] * a[8] * a[9] + 8.0 * a[4] * a[5] * a[6] * a[7]),
    # (g1, g2, mask2, srows, sgrad, S, argmax, rois, R, B, C, H, W, ...): same algorithmic figure as the single-source call
    # (SURVEY 8d counts ONE gradient read; the second dense source replaces a separate accumulate pass)
    "odwscl_roi_pool_bwd_nhwc_multi_f32": lambda a: ("byte", 8.0 * a[8] * a[10] * 49 + 8.0 * a[9] * a[10] * a[11] * a[12]),
    # (F, N, out, ...): N x N x 128 contraction
    "odwscl_sim_nxn_f32": lambda a: ("flop", 2.0 * a[1] * a[1] * 128),
    # (A, B, C, M, N, K, ...)
    "odwscl_gemm_nt_tf32": lambda a: ("flop", 2.0 * a[3] * a[4] * a[5]),
    # (A, lda, a_mn, B, ldb, b_mn, C, ldc, M, N, K, ...)
    "odwscl_fc_gemm_tf32": lambda a: ("flop", 2.0 * a[8] * a[9] * (a[10] + a[23])),
    # (A, lda, a_mn, B, ldb, b_mn, C, C_mc, ldc, M, N, K, scale, max_pairs, A2, lda2, B2, ldb2, K2, stream)
    "odwscl_fc_gemm_peer_sum_tf32": lambda a: ("flop", 2.0 * a[9] * a[10] * (a[11] + a[18])),
    # (A, lda, a_mn, B, ldb, b_mn, C, peer_C, n_peers, rows_per_owner, ldc, M, N, K, scale, max_pairs, A2, lda2, B2, ldb2, K2)
    "odwscl_fc_gemm_peer_scatter_tf32": lambda a: ("flop", 2.0 * a[11] * a[12] * (a[13] + a[20])),
}


trace = None               # host-side diagnosis (bench.py ODWSCL_BENCH_DEBUG=2): [(t_enter, t_exit, name)] of every C-ABI call


def _call(name, *args):
    global launch_count
    if profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if trace is not None:
        t_in = time.perf_counter()
        rc = getattr(lib(), name)(*args)
        trace.append((t_in, time.perf_counter(), name))
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed: %s (%d)" % (name, lib().odwscl_strerror(rc).decode(), rc))
    launch_count += _LAUNCHES.get(name, 0)
    if profile is not None:
        e1.record()
        w = _WORK.get(name)
        profile.append((name, w(args) if w else None, e0, e1))


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (no CPU implementation: csrc/ROIPool.h:23)" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


# ------------------------------------------------------------------------------------------
def _is_nhwc(t):
    """NCHW-shaped tensor whose memory is channels-last (what the conv stack hands to the pooler)."""
    return t.dim() == 4 and t.shape[1] > 1 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def roi_pool_forward(feat, rois, scale, ph, pw, out=None):
    """`out` (optional): a contiguous [R,C,ph,pw] fp32 buffer to pool into (e.g. the first half of a larger batch)."""
    rois = _chk(rois, torch.float32, "rois")
    B, C, H, W = feat.shape
    R = rois.shape[0]
    nhwc = _is_nhwc(feat) and ph == 7 and pw == 7 and C % 4 == 0 and feat.dtype == torch.float32 and feat.is_cuda
    if not nhwc:
        feat = _chk(feat, torch.float32, "input")
    if out is None:
        out = torch.empty((R, C, ph, pw), dtype=torch.float32, device=feat.device)
    assert out.shape == (R, C, ph, pw) and out.is_contiguous() and out.dtype == torch.float32
    arg = torch.empty((R, C, ph, pw), dtype=torch.int32, device=feat.device)
    if out.numel() == 0:
        return out, arg
    with torch.cuda.device(feat.device):
        if nhwc:            # memory is already [B,H,W,C]: no transpose pass
            _call("odwscl_roi_pool_fwd_nhwc_f32", _ptr(feat), B, C, H, W, _ptr(rois), R, float(scale), _ptr(out),
                  _ptr(arg), _stream())
        else:
            n = lib().odwscl_roi_pool_fwd_ws_bytes(B, C, H, W, R, ph, pw)
            ws = _workspace(n, feat.device)
            _call("odwscl_roi_pool_fwd_f32", _ptr(feat), B, C, H, W, _ptr(rois), R, float(scale), ph, pw,
                  _ptr(out), _ptr(arg), _ptr(ws), ws.numel(), _stream())
    return out, arg


def roi_pool_forward_aug(feat, rois, scale, centres, block, buf):
    """ROIPool (7x7) of a channels-last map into buf[:R] AND its DropBlock-augmented copy into buf[R:] in one pass.
    Returns (argmax, scale_io, bmask [R,49])."""
    rois = _chk(rois, torch.float32, "rois")
    centres = _chk(centres, torch.float32, "centres")
    B, C, H, W = feat.shape
    R = rois.shape[0]
    assert _is_nhwc(feat) and C % 4 == 0 and feat.dtype == torch.float32 and feat.is_cuda
    assert buf.shape == (2 * R, C, 7, 7) and buf.is_contiguous() and buf.dtype == torch.float32
    arg = torch.empty((R, C, 7, 7), dtype=torch.int32, device=feat.device)
    scale_io = torch.empty((2,), dtype=torch.float32, device=feat.device)
    bmask = torch.empty((R, 49), dtype=torch.float32, device=feat.device)
    if R == 0:
        return arg, scale_io, bmask
    with torch.cuda.device(feat.device):
        _call("odwscl_dropblock_prepare_f32", _ptr(centres), R, 7, 7, int(block), _ptr(scale_io), _ptr(bmask), _stream())
        _call("odwscl_roi_pool_fwd_nhwc_aug_f32", _ptr(feat), B, C, H, W, _ptr(rois), R, float(scale), _ptr(buf[:R]),
              _ptr(arg), _ptr(bmask), _ptr(buf[R:]), _stream())
    return arg, scale_io, bmask


def probe_roi_stream(feat, rois, scale):
    """Measurement aid (scripts/run_roipool_once.py): see odwscl_probe_roi_stream_f32."""
    assert _is_nhwc(feat)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    out = torch.empty((R * ((C + 127) // 128) * 7,), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _call("odwscl_probe_roi_stream_f32", _ptr(feat), B, C, H, W, _ptr(_chk(rois, torch.float32, "rois")), R, float(scale),
              _ptr(out), _stream())
    return out


def roi_pool_backward(grad, rois, argmax, ph, pw, B, C, H, W, channels_last=False):
    grad, rois = _chk(grad, torch.float32, "grad"), _chk(rois, torch.float32, "rois")
    argmax = _chk(argmax, torch.int32, "argmax")
    with torch.cuda.device(grad.device):
        if channels_last and ph == 7 and pw == 7:
            gin = torch.empty((B, H, W, C), dtype=torch.float32, device=grad.device)
            _call("odwscl_roi_pool_bwd_nhwc_f32", _ptr(grad), _ptr(argmax), _ptr(rois), rois.shape[0], B, C, H, W,
                  _ptr(gin), _stream())
            return gin.permute(0, 3, 1, 2)
        gin = torch.empty((B, C, H, W), dtype=torch.float32, device=grad.device)
        _call("odwscl_roi_pool_bwd_f32", _ptr(grad), _ptr(argmax), _ptr(rois), rois.shape[0], B, C, H, W, ph, pw,
              _ptr(gin), _stream())
    return gin


def dropblock_mask(centres, block, scale_io):
    """[R,49] per-(roi, bin) factor block_mask * scale of a DropBlock whose scale_io is known."""
    centres = _chk(centres, torch.float32, "centres")
    R, ph, pw = centres.shape
    out = torch.empty((R, ph * pw), dtype=torch.float32, device=centres.device)
    with torch.cuda.device(centres.device):
        _call("odwscl_dropblock_mask_f32", _ptr(centres), R, ph, pw, int(block), _ptr(scale_io), _ptr(out), _stream())
    return out


def roi_pool_backward_multi(grad, grad2, srows, sgrad, rois, argmax, B, C, H, W, mask2=None):
    """NHWC grad map of a pooled tensor with up to two dense consumers and one sparse (gathered-rows) consumer.
    Returns None when the map does not fit the plane-centric kernel (the caller sums the gradients itself)."""
    if H * W * 4 > 220 * 1024:
        return None
    grad, rois = _chk(grad, torch.float32, "grad"), _chk(rois, torch.float32, "rois")
    argmax = _chk(argmax, torch.int32, "argmax")
    if grad2 is not None:
        grad2 = _chk(grad2, torch.float32, "grad2")
        assert grad2.shape == grad.shape
    if mask2 is not None:
        mask2 = _chk(mask2, torch.float32, "mask2")
        assert grad2 is not None and mask2.numel() == grad.shape[0] * 49
    S = 0
    if srows is not None and srows.numel() > 0:
        srows, sgrad = _chk(srows, torch.int64, "srows"), _chk(sgrad, torch.float32, "sgrad")
        S = srows.numel()
        assert sgrad.shape[0] == S and sgrad.numel() == S * C * 49
    gin = torch.empty((B, H, W, C), dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        _call("odwscl_roi_pool_bwd_nhwc_multi_f32", _ptr(grad), _ptr(grad2), _ptr(mask2), _ptr(srows) if S else None,
              _ptr(sgrad) if S else None, S, _ptr(argmax), _ptr(rois), rois.shape[0], B, C, H, W, _ptr(gin), _stream())
    return gin.permute(0, 3, 1, 2)


def roi_align_forward(feat, rois, scale, ph, pw, sampling_ratio):
    rois = _chk(rois, torch.float32, "rois")
    B, C, H, W = feat.shape
    if _is_nhwc(feat) and (ph, pw) == (7, 7) and C % 4 == 0 and feat.dtype == torch.float32 and feat.is_cuda:
        out = torch.empty((rois.shape[0], C, 7, 7), dtype=torch.float32, device=feat.device)
        with torch.cuda.device(feat.device):       # memory is [B,H,W,C]: the channels-last kernel, no transpose
            _call("odwscl_roi_align_fwd_nhwc_f32", _ptr(feat), B, C, H, W, _ptr(rois), rois.shape[0], float(scale),
                  int(sampling_ratio), _ptr(out), _stream())
        return out
    feat = _chk(feat, torch.float32, "input")
    out = torch.empty((rois.shape[0], C, ph, pw), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _call("odwscl_roi_align_fwd_f32", _ptr(feat), B, C, H, W, _ptr(rois), rois.shape[0], float(scale), ph, pw,
              int(sampling_ratio), _ptr(out), _stream())
    return out


def roi_align_backward(grad, rois, scale, ph, pw, B, C, H, W, sampling_ratio, channels_last=False):
    grad, rois = _chk(grad, torch.float32, "grad"), _chk(rois, torch.float32, "rois")
    if channels_last and (ph, pw) == (7, 7) and C % 4 == 0 and H * W * 16 <= 220 * 1024:
        gin = torch.empty((B, H, W, C), dtype=torch.float32, device=grad.device)
        with torch.cuda.device(grad.device):
            _call("odwscl_roi_align_bwd_nhwc_f32", _ptr(grad), _ptr(rois), rois.shape[0], float(scale), B, C, H, W,
                  int(sampling_ratio), _ptr(gin), _stream())
        return gin.permute(0, 3, 1, 2)
    gin = torch.empty((B, C, H, W), dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        _call("odwscl_roi_align_bwd_f32", _ptr(grad), _ptr(rois), rois.shape[0], float(scale), ph, pw, B, C, H, W,
              int(sampling_ratio), _ptr(gin), _stream())
    return gin


def box_iou(a, b, plus_one=True):
    a, b = _chk(a, torch.float32, "boxes a").view(-1, 4), _chk(b, torch.float32, "boxes b").view(-1, 4)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _call("odwscl_box_iou_f32", _ptr(a), a.shape[0], _ptr(b), b.shape[0], int(plus_one), _ptr(out), _stream())
    return out


NMS_SINGLE_CTA_MAX = 8192      # boxes the one-CTA kernels hold in shared memory; larger inputs take the three-launch path


def _nms_large(boxes, box_stride, scores, score_stride, n, score_thr, thr, legacy, keep, cnt):
    """boxes / scores: tensors whose data_ptr() is the first box / score; strides in floats."""
    dev = boxes.device
    ws = _workspace(lib().odwscl_nms_large_ws_bytes(n), dev)
    k64 = _ptr(keep) if keep.dtype == torch.int64 else None
    k32 = _ptr(keep) if keep.dtype == torch.int32 else None
    _call("odwscl_nms_large_f32", _ptr(boxes), int(box_stride), _ptr(scores), int(score_stride), n, float(score_thr),
          float(thr), int(legacy), k64, k32, _ptr(cnt), _ptr(ws), ws.numel(), _stream())


def _nms(name, boxes, scores, thr):
    boxes, scores = _chk(boxes, torch.float32, "boxes").view(-1, 4), _chk(scores, torch.float32, "scores").view(-1)
    n = boxes.shape[0]
    keep = torch.empty((n,), dtype=torch.int64, device=boxes.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        if n > NMS_SINGLE_CTA_MAX:
            _nms_large(boxes, 4, scores, 1, n, float("-inf"), thr, name == "odwscl_nms_legacy_f32", keep, cnt)
        else:
            _call(name, _ptr(boxes), _ptr(scores), n, float(thr), _ptr(keep), _ptr(cnt), _stream())
    return keep, cnt


def nms_padded(boxes, scores, thr):
    """torchvision-semantics NMS; returns (keep [n] padded, n_keep [1]) without synchronising."""
    return _nms("odwscl_nms_f32", boxes, scores, thr)


def nms(boxes, scores, thr):
    keep, cnt = _nms("odwscl_nms_f32", boxes, scores, thr)
    return keep[: int(cnt.item())]


def nms_legacy(boxes, scores, thr):
    keep, cnt = _nms("odwscl_nms_legacy_f32", boxes, scores, thr)
    return keep[: int(cnt.item())]


def nms_per_class(boxes, scores, score_thr, nms_thr):
    """boxes [N,C*4], scores [N,C] -> (keep [C,N] int32, n_keep [C] int32); all foreground classes in one launch."""
    boxes, scores = _chk(boxes, torch.float32, "boxes"), _chk(scores, torch.float32, "scores")
    N, C = scores.shape
    assert boxes.shape == (N, C * 4)
    keep = torch.empty((C, max(N, 1)), dtype=torch.int32, device=boxes.device)
    cnt = torch.zeros((C,), dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        if N > NMS_SINGLE_CTA_MAX:      # e.g. the UNION of many test-time views: class by class through the large path
            for j in range(1, C):
                _nms_large(boxes[:, 4 * j:], 4 * C, scores[:, j:], C, N, score_thr, nms_thr, False, keep[j], cnt[j:j + 1])
        else:
            _call("odwscl_nms_per_class_f32", _ptr(boxes), _ptr(scores), N, C, float(score_thr), float(nms_thr), _ptr(keep),
                  _ptr(cnt), _stream())
    return keep, cnt


def sim_nxn(F):
    F = _chk(F, torch.float32, "F")
    assert F.shape[1] == 128
    out = torch.empty((F.shape[0], F.shape[0]), dtype=torch.float32, device=F.device)
    with torch.cuda.device(F.device):
        ws = _workspace(lib().odwscl_sim_nxn_ws_bytes(F.shape[0]), F.device)
        _call("odwscl_sim_nxn_f32", _ptr(F), F.shape[0], _ptr(out), _ptr(ws), ws.numel(), _stream())
    return out


def gemm_nt_tf32(A, B):
    """C = A @ B.T with single-pass TF32 tensor-core math (tcgen05)."""
    A, B = _chk(A, torch.float32, "A"), _chk(B, torch.float32, "B")
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K and K % 4 == 0
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        _call("odwscl_gemm_nt_tf32", _ptr(A), _ptr(B), _ptr(C), M, N, K, N, _stream())
    return C


def dropblock(x, centres, block, scale_io=None, n_valid=None, out=None):
    """y = x * block_mask * numel/sum.  Pass the returned scale_io back in for the backward.  `n_valid` (int32 device
    tensor [1]) restricts the renormalisation and the output to the first n_valid rows of a padded batch."""
    x, centres = _chk(x, torch.float32, "x"), _chk(centres, torch.float32, "centres")
    R, C, ph, pw = x.shape
    y = torch.empty_like(x) if out is None else out
    assert y.shape == x.shape and y.is_contiguous() and y.dtype == torch.float32
    reuse = scale_io is not None
    if scale_io is None:
        scale_io = torch.empty((2,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        if n_valid is None:
            _call("odwscl_dropblock_f32", _ptr(x), _ptr(centres), R, C, ph, pw, int(block), _ptr(y), _ptr(scale_io),
                  int(reuse), _stream())
        else:
            _call("odwscl_dropblock_rows_f32", _ptr(x), _ptr(centres), R, C, ph, pw, int(block), _ptr(y),
                  _ptr(scale_io), int(reuse), _ptr(_chk(n_valid, torch.int32, "n_valid")), _stream())
    return y, scale_io


def dropblock_seg(x, centres, block, seg_off, P, scale_seg=None):
    """DropBlock over a batch of P row segments (seg_off int32 [>=P+1], device), each renormalised on its own (the
    reference's per-(image, class) drop_pool calls, loss.py:299); rows >= seg_off[P] are padding (zero-filled)."""
    x, centres = _chk(x, torch.float32, "x"), _chk(centres, torch.float32, "centres")
    seg_off = _chk(seg_off, torch.int32, "seg_off")
    assert seg_off.numel() >= P + 1 and P > 0
    R, C, ph, pw = x.shape
    y = torch.empty_like(x)
    reuse = scale_seg is not None
    if scale_seg is None:
        scale_seg = torch.empty((P, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _call("odwscl_dropblock_seg_f32", _ptr(x), _ptr(centres), R, C, ph, pw, int(block), _ptr(y), _ptr(seg_off), int(P),
              _ptr(scale_seg), int(reuse), _stream())
    return y, scale_seg


# ---------------------------------------------------------------- conv stack (NHWC)
CONV_RELU, CONV_ACCUM, CONV_MASK, CONV_ROUND = 1, 2, 4, 8
SUPCON_SPLITS = 8          # ODWSCL_SUPCON_SPLITS in include/odwscl.h


def conv3x3_nhwc(x, w_krsc, bias, dilation=1, flags=0, mask_src=None, out=None):
    """x [B,H,W,Cin], w_krsc [Cout,3,3,Cin] -> y [B,H,W,Cout] (tcgen05 implicit GEMM, TF32)."""
    x, w_krsc = _chk(x, torch.float32, "x"), _chk(w_krsc, torch.float32, "w")
    B, H, W, Cin = x.shape
    Cout = w_krsc.shape[0]
    assert w_krsc.shape == (Cout, 3, 3, Cin)
    y = out if out is not None else torch.empty((B, H, W, Cout), dtype=torch.float32, device=x.device)
    assert y.is_contiguous() and y.shape == (B, H, W, Cout)
    if mask_src is not None:
        mask_src = _chk(mask_src, torch.float32, "mask_src")
        assert mask_src.shape == y.shape
    with torch.cuda.device(x.device):
        _call("odwscl_conv3x3_nhwc_tf32", _ptr(x), B, H, W, Cin, _ptr(w_krsc), _ptr(bias), Cout, int(dilation),
              int(flags), _ptr(mask_src), _ptr(y), _stream())
    return y


def conv3x3_wgrad_nhwc(x, dz, dilation=1, accumulate_into=None, want_bias=True):
    """x [B,H,W,Cin], dz [B,H,W,Cout] -> (dw_krsc [Cout,3,3,Cin], db [Cout] or None)."""
    x, dz = _chk(x, torch.float32, "x"), _chk(dz, torch.float32, "dz")
    B, H, W, Cin = x.shape
    Cout = dz.shape[3]
    assert dz.shape[:3] == x.shape[:3]
    dw = torch.empty((Cout, 3, 3, Cin), dtype=torch.float32, device=x.device)
    db = torch.empty((Cout,), dtype=torch.float32, device=x.device) if want_bias else None
    with torch.cuda.device(x.device):
        _call("odwscl_conv3x3_wgrad_nhwc_tf32", _ptr(x), _ptr(dz), B, H, W, Cin, Cout, int(dilation), _ptr(dw),
              _ptr(db), _stream())
    if accumulate_into is not None:
        accumulate_into.add_(dw)
        dw = accumulate_into
    return dw, db


def conv3x3_c3(x_nchw, w_oihw, bias, relu=True, round_tf32=False):
    x_nchw, w_oihw = _chk(x_nchw, torch.float32, "x"), _chk(w_oihw, torch.float32, "w")
    B, _, H, W = x_nchw.shape
    Cout = w_oihw.shape[0]
    y = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x_nchw.device)
    with torch.cuda.device(x_nchw.device):
        _call("odwscl_conv3x3_c3_f32", _ptr(x_nchw), B, H, W, _ptr(w_oihw), _ptr(bias), Cout,
              (CONV_RELU if relu else 0) | (CONV_ROUND if round_tf32 else 0), _ptr(y), _stream())
    return y


def maxpool2x2_nhwc(x):
    x = _chk(x, torch.float32, "x")
    B, H, W, C = x.shape
    y = torch.empty((B, H // 2, W // 2, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _call("odwscl_maxpool2x2_nhwc_f32", _ptr(x), B, H, W, C, _ptr(y), _stream())
    return y


def maxpool2x2_nhwc_bwd(x, gy, relu_mask):
    x, gy = _chk(x, torch.float32, "x"), _chk(gy, torch.float32, "gy")
    B, H, W, C = x.shape
    gx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _call("odwscl_maxpool2x2_nhwc_bwd_f32", _ptr(x), _ptr(gy), B, H, W, C, int(relu_mask), _ptr(gx), _stream())
    return gx


def round_tf32_(x):
    """in place: x <- rna_tf32(x)"""
    assert x.is_contiguous() and x.dtype == torch.float32
    with torch.cuda.device(x.device):
        _call("odwscl_split_tf32", _ptr(x), x.numel(), _ptr(x), None, _stream())
    return x


def split_tf32(x):
    x = _chk(x, torch.float32, "x")
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        _call("odwscl_split_tf32", _ptr(x), x.numel(), _ptr(hi), _ptr(lo), _stream())
    return hi, lo


def conv_weight_xform(w_oihw, want_fwd=True, want_dgrad=False, round_tf32=True):
    """[Cout,Cin,3,3] -> ([Cout,3,3,Cin] or None, tap-flipped [Cin,3,3,Cout] or None), TF32-rounded, one pass."""
    w_oihw = _chk(w_oihw, torch.float32, "w")
    Cout, Cin = w_oihw.shape[:2]
    wk = torch.empty((Cout, 3, 3, Cin), dtype=torch.float32, device=w_oihw.device) if want_fwd else None
    wd = torch.empty((Cin, 3, 3, Cout), dtype=torch.float32, device=w_oihw.device) if want_dgrad else None
    with torch.cuda.device(w_oihw.device):
        _call("odwscl_conv_weight_xform_f32", _ptr(w_oihw), Cout, Cin, _ptr(wk), _ptr(wd), int(round_tf32), _stream())
    return wk, wd


def relu_dropout_(x, p, seed):
    """in place: x <- relu(x) * Bernoulli(1-p) / (1-p)"""
    assert x.is_contiguous() and x.dtype == torch.float32 and x.numel() % 4 == 0
    with torch.cuda.device(x.device):
        _call("odwscl_relu_dropout_fwd_f32", _ptr(x), x.numel(), float(p), ctypes.c_ulonglong(seed & (2 ** 64 - 1)), _stream())
    return x


def relu_dropout_backward(y, gy, p):
    y, gy = _chk(y, torch.float32, "y"), _chk(gy, torch.float32, "gy")
    gx = torch.empty_like(gy)
    with torch.cuda.device(y.device):
        _call("odwscl_relu_dropout_bwd_f32", _ptr(y), _ptr(gy), _ptr(gx), y.numel(), float(p), _stream())
    return gx


# ---------------------------------------------------------------- fully-connected block (csrc/fc_gemm.cu)
FC_BIAS, FC_ACCUM, FC_RELU, FC_DROPOUT, FC_MASK, FC_ROUND = 1, 2, 4, 8, 16, 32
FC_MAX_PAIRS = 0           # > 0: cap on resident CTA pairs of the fc GEMMs (sharding.py sets it when NCCL shares the SMs)


def _rows2d(t, name):
    """(t', pitch): a 2-D fp32 CUDA tensor with contiguous, 16-byte aligned rows and a pitch (floats) that is a multiple
    of 4 -- what a TMA descriptor needs.  Tensors that already qualify (incl. row / column slices) are used where they
    lie; anything else is copied once into a pitch-padded buffer."""
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (the fc block has no CPU implementation)" % name)
    if t.dtype != torch.float32 or t.dim() != 2:
        raise RuntimeError("%s must be a 2-D float32 tensor" % name)
    rows, cols = t.shape
    unit = cols <= 1 or t.stride(1) == 1
    if rows <= 1:
        if unit and t.data_ptr() % 16 == 0:
            return t, (cols + 3) // 4 * 4
    elif unit and t.stride(0) % 4 == 0 and t.stride(0) >= cols and t.data_ptr() % 16 == 0:
        return t, t.stride(0)
    ld = (cols + 3) // 4 * 4
    buf = torch.empty((rows, ld), dtype=t.dtype, device=t.device)
    buf[:, :cols].copy_(t)
    return buf[:, :cols], ld


def fc_gemm(A, B, a_mn=False, b_mn=False, out=None, bias=None, relu=False, dropout_p=0.0, seed=0, mask_src=None,
            mask_scale=1.0, accumulate=False, round_tf32=False, A2=None, B2=None):
    """C[M,N] (+)= sum_k A(m,k) B(n,k) on the persistent tcgen05 CTA-pair kernel (TF32 math, fp32 accumulate).
    A: [M,K] (a_mn False) or [K,M] (a_mn True); B: [N,K] (b_mn False) or [K,N] (b_mn True); rows contiguous.
    Fused epilogue: bias, accumulate, ReLU, Dropout(p, seed), derivative mask (mask_src > 0) * mask_scale, TF32 round.
    (A2, B2): a second operand pair of the same layouts whose contraction is appended (C = A B^T + A2 B2^T)."""
    (A, lda), (B, ldb) = _rows2d(A, "A"), _rows2d(B, "B")
    K2, lda2, ldb2 = 0, 0, 0
    if A2 is not None:
        (A2, lda2), (B2, ldb2) = _rows2d(A2, "A2"), _rows2d(B2, "B2")
        K2 = A2.shape[0] if a_mn else A2.shape[1]
        assert (A2.shape[1] if a_mn else A2.shape[0]) == (A.shape[1] if a_mn else A.shape[0])
        assert (B2.shape[0] if b_mn else B2.shape[1]) == K2 and (B2.shape[1] if b_mn else B2.shape[0]) == (B.shape[1] if b_mn else B.shape[0])
    K, M = (A.shape[0], A.shape[1]) if a_mn else (A.shape[1], A.shape[0])
    Kb, N = (B.shape[0], B.shape[1]) if b_mn else (B.shape[1], B.shape[0])
    if K != Kb:
        raise RuntimeError("fc_gemm: contraction sizes differ (%d vs %d)" % (K, Kb))
    if out is None:
        if accumulate:
            raise RuntimeError("fc_gemm: accumulate needs `out`")
        ldc = (N + 3) // 4 * 4
        out = torch.empty((M, ldc), dtype=torch.float32, device=A.device)[:, :N]
    if tuple(out.shape) != (M, N) or out.dtype != torch.float32 or (N > 1 and out.stride(1) != 1) or out.data_ptr() & 15 \
            or (M > 1 and out.stride(0) < N):
        raise RuntimeError("fc_gemm: `out` must be a [%d,%d] float32 tensor with contiguous, 16-byte aligned rows" % (M, N))
    flags = 0
    if bias is not None:
        bias = _chk(bias, torch.float32, "bias")
        flags |= FC_BIAS
    if accumulate:
        flags |= FC_ACCUM
    if relu:
        flags |= FC_RELU
    if dropout_p > 0.0:
        flags |= FC_DROPOUT
    ld_mask = 0
    if mask_src is not None:
        mask_src, ld_mask = _rows2d(mask_src, "mask_src")
        assert tuple(mask_src.shape) == (M, N)
        flags |= FC_MASK
    if round_tf32:
        flags |= FC_ROUND
    if M == 0 or N == 0:
        return out
    if K == 0:
        if not accumulate:
            out.zero_()
        return out
    ldc = out.stride(0) if M > 1 else max(N, out.stride(0))
    with torch.cuda.device(A.device):
        _call("odwscl_fc_gemm_tf32", _ptr(A), int(lda), int(a_mn), _ptr(B), int(ldb), int(b_mn), _ptr(out), int(ldc), M, N,
              K, flags, _ptr(bias), _ptr(mask_src), int(ld_mask), float(mask_scale), float(dropout_p),
              ctypes.c_ulonglong(seed & (2 ** 64 - 1)), int(FC_MAX_PAIRS), _ptr(A2) if K2 else None, int(lda2),
              _ptr(B2) if K2 else None, int(ldb2), int(K2), _stream())
    return out


def fc_gemm_peer_sum(A, B, out, out_multicast_ptr, scale, a_mn=False, b_mn=False, A2=None, B2=None, peer_ptrs=None,
                     rows_per_owner=0):
    """out (this rank's replica, zeroed by the caller on every rank) and every peer's replica += scale * (A B^T [+ A2 B2^T]):
    the product leaves the epilogue as multimem.red.add through `out_multicast_ptr`, the NVSwitch multicast address of
    `out` (sharding.PeerGradSum owns the symmetric allocation and the two rank barriers around the step's use of it).
    With `peer_ptrs` (every rank's address of `out`, this rank's included) the product is reduce-scattered instead: rows
    [r * rows_per_owner, ...) are added to rank r's replica only."""
    (A, lda), (B, ldb) = _rows2d(A, "A"), _rows2d(B, "B")
    K2, lda2, ldb2 = 0, 0, 0
    if A2 is not None:
        (A2, lda2), (B2, ldb2) = _rows2d(A2, "A2"), _rows2d(B2, "B2")
        K2 = A2.shape[0] if a_mn else A2.shape[1]
    K, M = (A.shape[0], A.shape[1]) if a_mn else (A.shape[1], A.shape[0])
    Kb, N = (B.shape[0], B.shape[1]) if b_mn else (B.shape[1], B.shape[0])
    if K != Kb or K == 0:
        raise RuntimeError("fc_gemm_peer_sum: contraction sizes (%d vs %d)" % (K, Kb))
    if tuple(out.shape) != (M, N) or out.dtype != torch.float32 or not out.is_contiguous() or out.data_ptr() & 15 \
            or (peer_ptrs is None and (int(out_multicast_ptr) & 15 or not out_multicast_ptr)):
        raise RuntimeError("fc_gemm_peer_sum: `out` must be a contiguous, 16-byte aligned [%d,%d] float32 view of a "
                           "symmetric-memory buffer with a multicast address" % (M, N))
    if peer_ptrs is not None:
        tab = (c_void_p * len(peer_ptrs))(*[int(q) for q in peer_ptrs])
        with torch.cuda.device(A.device):
            _call("odwscl_fc_gemm_peer_scatter_tf32", _ptr(A), int(lda), int(a_mn), _ptr(B), int(ldb), int(b_mn), _ptr(out),
                  ctypes.cast(tab, c_void_p), len(peer_ptrs), int(rows_per_owner), int(N), M, N, K, float(scale),
                  int(FC_MAX_PAIRS), _ptr(A2) if K2 else None, int(lda2), _ptr(B2) if K2 else None, int(ldb2), int(K2),
                  _stream())
        return out
    with torch.cuda.device(A.device):
        _call("odwscl_fc_gemm_peer_sum_tf32", _ptr(A), int(lda), int(a_mn), _ptr(B), int(ldb), int(b_mn), _ptr(out),
              c_void_p(int(out_multicast_ptr)), int(N), M, N, K, float(scale), int(FC_MAX_PAIRS), _ptr(A2) if K2 else None,
              int(lda2), _ptr(B2) if K2 else None, int(ldb2), int(K2), _stream())
    return out


def peer_add(src, dst_multicast_ptr, scale=1.0):
    """every rank's replica += scale * src, through the multicast address of the replica (numel % 4 == 0)."""
    src = _chk(src, torch.float32, "src")
    with torch.cuda.device(src.device):
        _call("odwscl_peer_add_f32", _ptr(src), c_void_p(int(dst_multicast_ptr)), src.numel(), float(scale), _stream())


def peer_broadcast(src, dst_multicast_ptr):
    """every rank's replica = src (numel % 4 == 0): the all-gather half of the reduce-scattered gradient."""
    src = _chk(src, torch.float32, "src")
    with torch.cuda.device(src.device):
        _call("odwscl_peer_broadcast_f32", _ptr(src), c_void_p(int(dst_multicast_ptr)), src.numel(), _stream())


def colsum(x, out=None, accumulate=False):
    """out[c] (+)= sum_r x[r, c]."""
    x, ld = _rows2d(x, "x")
    rows, cols = x.shape
    if out is None:
        out = torch.empty((cols,), dtype=torch.float32, device=x.device)
        accumulate = False
    with torch.cuda.device(x.device):
        _call("odwscl_colsum_f32", _ptr(x), rows, cols, int(ld), _ptr(out), int(accumulate), _stream())
    return out


def l2norm_forward(z, eps=1e-12):
    z, ld = _rows2d(z, "z")
    R, D = z.shape
    y = torch.empty((R, D), dtype=torch.float32, device=z.device)
    inv = torch.empty((R,), dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        _call("odwscl_l2norm_fwd_f32", _ptr(z), int(ld), R, D, float(eps), _ptr(y), _ptr(inv), _stream())
    return y, inv


def l2norm_backward(y, g, inv, eps=1e-12):
    y, g = _chk(y, torch.float32, "y"), _chk(g, torch.float32, "g")
    R, D = y.shape
    dz = torch.empty_like(y)
    with torch.cuda.device(y.device):
        _call("odwscl_l2norm_bwd_f32", _ptr(y), _ptr(g), _ptr(inv), R, D, float(eps), _ptr(dz), _stream())
    return dz


def aug_positives(src, rows, seg_off, P, centres, block, scale_seg=None, noise=None, seed=0, backward=False):
    """Forward: src = pooled [R,C,7,7] (or [R, D]) -> out [2 Kc, C,7,7] (DropBlock views, then noise views), scale_seg [P,2].
    Backward: src = gradient of out -> gradient w.r.t. the Kc gathered rows."""
    src = _chk(src, torch.float32, "src")
    centres = _chk(centres, torch.float32, "centres")
    Kc, ph, pw = centres.shape
    cells = ph * pw
    D = src[0].numel()
    compute = scale_seg is None
    if compute:
        scale_seg = torch.empty((P, 2), dtype=torch.float32, device=src.device)
    if noise is not None:
        noise = _chk(noise, torch.float32, "noise")
        assert noise.numel() == Kc * D
    dst = torch.empty(((Kc if backward else 2 * Kc),) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
    if Kc > 0:
        with torch.cuda.device(src.device):
            _call("odwscl_aug_positives_f32", _ptr(src), D, cells, _ptr(_chk(rows, torch.int64, "rows")), Kc,
                  _ptr(_chk(seg_off, torch.int32, "seg_off")), int(P), _ptr(centres), int(block), _ptr(scale_seg), int(compute),
                  _ptr(noise), ctypes.c_ulonglong(seed & (2 ** 64 - 1)), int(backward), _ptr(dst), _stream())
    return dst, scale_seg


def spec_index(k_dev, rowsA, Kc, sel_n):
    """-> (rows int64 [Kc], sel int64 [sel_n], overflow fp32 [1]); see odwscl_spec_index."""
    dev = rowsA.device
    rows = torch.empty((Kc,), dtype=torch.int64, device=dev)
    sel = torch.empty((sel_n,), dtype=torch.int64, device=dev)
    ovf = torch.empty((1,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("odwscl_spec_index", _ptr(_chk(k_dev, torch.int32, "k")), _ptr(_chk(rowsA, torch.int32, "rowsA")), int(Kc),
              int(sel_n), _ptr(rows), _ptr(sel), _ptr(ovf), _stream())
    return rows, sel, ovf


# ---------------------------------------------------------------- MIL + refinement losses (csrc/head_loss.cu)
class HeadScores:
    """Outputs of head_scores: what object discovery reads (final_score, sm1, sm2) + the statistics the loss reuses."""
    pass


def head_scores(logits, C, Q, img_off, B):
    """logits [R, >= 5C+3Q] (row pitch = stride(0)), img_off int32 [B+1] on the device."""
    logits, ld = _rows2d(logits, "logits")
    R = logits.shape[0]
    dev = logits.device
    f32 = dict(dtype=torch.float32, device=dev)
    hs = HeadScores()
    hs.logits, hs.ld, hs.R, hs.C, hs.Q, hs.B, hs.img_off = logits, ld, R, C, Q, B, img_off
    hs.det_max, hs.det_sum = torch.empty((B, C), **f32), torch.empty((B, C), **f32)
    hs.ref_colsum = torch.empty((3, B, C), **f32)
    hs.final_score, hs.sm1, hs.sm2 = (torch.empty((R, C), **f32) for _ in range(3))
    hs.img_score = torch.empty((B, C), **f32)
    with torch.cuda.device(dev):
        ws = _workspace(lib().odwscl_head_scores_ws_bytes(B, C), dev)
        _call("odwscl_head_scores_f32", _ptr(logits), int(ld), R, C, Q, _ptr(_chk(img_off, torch.int32, "img_off")), B,
              _ptr(hs.det_max), _ptr(hs.det_sum), _ptr(hs.ref_colsum), _ptr(hs.final_score), _ptr(hs.sm1), _ptr(hs.sm2),
              _ptr(hs.img_score), _ptr(ws), ws.numel(), _stream())
    return hs


def head_loss(hs, img_labels, pl, lw, rt, cls_agnostic, eps):
    """-> (out11 [11] = 7 losses + 4 accuracies (already / B), grad_logits [R, ld])."""
    dev = hs.logits.device
    out = torch.empty((11,), dtype=torch.float32, device=dev)
    grad = torch.empty((hs.R, hs.ld), dtype=torch.float32, device=dev)
    if hs.ld > 5 * hs.C + 3 * hs.Q:
        grad[:, 5 * hs.C + 3 * hs.Q:].zero_()                   # pitch padding columns
    partial = torch.empty((max((hs.R + 7) // 8, 1) * 6,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("odwscl_head_loss_f32", _ptr(hs.logits), int(hs.ld), hs.R, hs.C, hs.Q, int(cls_agnostic), _ptr(hs.img_off),
              hs.B, _ptr(hs.det_max), _ptr(hs.det_sum), _ptr(hs.img_score), _ptr(_chk(img_labels, torch.float32, "labels")),
              _ptr(_chk(pl, torch.int64, "pl")), _ptr(_chk(lw, torch.float32, "lw")), _ptr(_chk(rt, torch.float32, "rt")),
              _ptr(hs.ref_colsum), float(eps), _ptr(grad), _ptr(partial), _ptr(out), _stream())
    return out, grad


def head_grad_scale_(grad, C, Q, upstream7):
    with torch.cuda.device(grad.device):
        _call("odwscl_head_grad_scale_f32", _ptr(grad), int(grad.stride(0)), grad.shape[0], C, Q,
              _ptr(_chk(upstream7, torch.float32, "upstream")), _stream())
    return grad


# ---------------------------------------------------------------- object discovery / SupCon
class DiscoveryState:
    """Device buffers of one object-discovery pass (sizes: P pairs, Ncap proposals per image)."""
    pass


def discover_phase_a(boxes, img_off, scores, pair_img, pair_cls, Ncap, thres):
    dev = boxes.device
    R, C = scores[0].shape
    P = pair_img.numel()
    B = img_off.numel() - 1
    i32 = dict(dtype=torch.int32, device=dev)
    st = DiscoveryState()
    st.P, st.Ncap, st.R, st.C, st.B = P, Ncap, R, C, B
    st.boxes, st.img_off, st.scores, st.pair_img, st.pair_cls = boxes, img_off, scores, pair_img, pair_cls
    st.amax = torch.empty((P, 3), **i32)
    st.member = torch.empty((P, Ncap), dtype=torch.uint8, device=dev)
    st.cntA = torch.empty((P,), **i32)
    st.offA = torch.empty((P + 1,), **i32)
    st.rowsA = torch.empty((max(P * Ncap, 1),), **i32)
    st.hardA = torch.empty((max(P * Ncap, 1),), dtype=torch.float32, device=dev)
    st.colsum = torch.empty((P,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("odwscl_discover_phase_a_f32", _ptr(boxes), _ptr(img_off), B, R, C, _ptr(scores[0]), _ptr(scores[1]),
              _ptr(scores[2]), _ptr(pair_img), _ptr(pair_cls), P, Ncap, float(thres), _ptr(st.amax),
              _ptr(st.member), _ptr(st.cntA), _ptr(st.offA), _ptr(st.rowsA), _ptr(st.hardA), _ptr(st.colsum),
              _stream())
    return st


def discover_phase_b(st, F, E, nms_thr, sim_rows_in=None):
    dev = F.device
    P, Ncap = st.P, st.Ncap
    i32 = dict(dtype=torch.int32, device=dev)
    st.inst = torch.empty((P, 3, Ncap), **i32)
    st.inst_cnt = torch.empty((P, 3), **i32)
    st.newl = torch.empty((P, 3, Ncap), **i32)
    st.new_cnt = torch.empty((P, 3), **i32)
    st.hardB = torch.empty((P, 3, Ncap), dtype=torch.float32, device=dev)
    st.tau = torch.empty((P, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("odwscl_discover_phase_b_f32", _ptr(st.boxes), _ptr(st.img_off), st.B, st.R, st.C, _ptr(st.scores[0]),
              _ptr(st.scores[1]), _ptr(st.scores[2]), _ptr(st.pair_img), _ptr(st.pair_cls), P, Ncap, _ptr(F),
              _ptr(E), _ptr(st.amax), _ptr(st.member), _ptr(st.cntA), _ptr(st.offA), _ptr(st.rowsA),
              _ptr(st.colsum), float(nms_thr), _ptr(st.inst), _ptr(st.inst_cnt), _ptr(st.newl), _ptr(st.new_cnt),
              _ptr(st.hardB), _ptr(st.tau), _ptr(sim_rows_in), _stream())
    return st


def bank_assemble(st, num_fg_classes, Mcap):
    dev = st.boxes.device
    st.Mcap = Mcap
    st.row_src = torch.empty((max(Mcap, 1),), dtype=torch.int32, device=dev)
    st.row_lab = torch.empty((max(Mcap, 1),), dtype=torch.int32, device=dev)
    st.row_w = torch.empty((max(Mcap, 1),), dtype=torch.float32, device=dev)
    st.M = torch.zeros((2,), dtype=torch.int32, device=dev)     # rows written (<= Mcap), unclamped row count
    with torch.cuda.device(dev):
        _call("odwscl_bank_assemble", _ptr(st.pair_img), _ptr(st.pair_cls), st.P, st.B, st.R, st.Ncap,
              int(num_fg_classes), _ptr(st.img_off), _ptr(st.cntA), _ptr(st.offA), _ptr(st.rowsA), _ptr(st.hardA),
              _ptr(st.newl), _ptr(st.new_cnt), _ptr(st.hardB), Mcap, _ptr(st.row_src), _ptr(st.row_lab),
              _ptr(st.row_w), _ptr(st.M), _stream())
    return st


def supcon_forward(F, E, row_src, row_lab, row_w, M_dev, Mcap, inv_temp):
    dev = F.device
    stats = torch.empty(((1 + SUPCON_SPLITS) * max(Mcap, 1), 4), dtype=torch.float32, device=dev)   # merged rows + split scratch
    loss = torch.zeros((1,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("odwscl_supcon_fwd_f32", _ptr(F), _ptr(E), F.shape[0], _ptr(row_src), _ptr(row_lab), _ptr(row_w),
              _ptr(M_dev), Mcap, float(inv_temp), _ptr(stats), _ptr(loss), _stream())
    return loss, stats


def supcon_tc_forward(F, E, row_src, row_lab, row_w, M_dev, Mcap, inv_temp):
    """SupCon forward with S = V V^T on the tensor cores (3xTF32 via fc_gemm); returns (loss, stats, ws) -- `ws` carries S to
    supcon_tc_backward."""
    dev = F.device
    stats = torch.empty(((1 + SUPCON_SPLITS) * max(Mcap, 1), 4), dtype=torch.float32, device=dev)
    loss = torch.zeros((1,), dtype=torch.float32, device=dev)
    nbytes = int(lib().odwscl_supcon_tc_ws_bytes(int(Mcap)))
    ws = torch.empty((max(nbytes, 16) // 4,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("odwscl_supcon_tc_fwd_f32", _ptr(F), _ptr(E), F.shape[0], _ptr(row_src), _ptr(row_lab), _ptr(row_w),
              _ptr(M_dev), Mcap, float(inv_temp), _ptr(ws), nbytes, _ptr(stats), _ptr(loss), _stream())
    return loss, stats, ws


def supcon_tc_backward(F, E, row_src, row_lab, row_w, M_dev, Mcap, inv_temp, stats, gscale, ws):
    dF = torch.zeros_like(F)
    dE = torch.zeros_like(E) if E is not None and E.numel() else None
    with torch.cuda.device(F.device):
        _call("odwscl_supcon_tc_bwd_f32", F.shape[0], _ptr(row_src), _ptr(row_lab), _ptr(row_w), _ptr(M_dev), Mcap,
              float(inv_temp), _ptr(ws), ws.numel() * 4, _ptr(stats), _ptr(gscale), _ptr(dF), _ptr(dE), _stream())
    return dF, dE


def supcon_backward(F, E, row_src, row_lab, row_w, M_dev, Mcap, inv_temp, stats, gscale):
    dF = torch.zeros_like(F)
    dE = torch.zeros_like(E) if E is not None and E.numel() else None
    with torch.cuda.device(F.device):
        _call("odwscl_supcon_bwd_f32", _ptr(F), _ptr(E), F.shape[0], _ptr(row_src), _ptr(row_lab), _ptr(row_w),
              _ptr(M_dev), Mcap, float(inv_temp), _ptr(stats), _ptr(gscale), _ptr(dF), _ptr(dE), _stream())
    return dF, dE


def od_layer(st, fg_thr):
    dev = st.boxes.device
    labels = torch.empty((3, st.R), dtype=torch.int64, device=dev)
    weights = torch.empty((3, st.R), dtype=torch.float32, device=dev)
    targets = torch.empty((3, st.R, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("odwscl_od_layer_f32", _ptr(st.boxes), _ptr(st.img_off), st.B, st.R, st.C, _ptr(st.scores[0]),
              _ptr(st.scores[1]), _ptr(st.scores[2]), _ptr(st.pair_img), _ptr(st.pair_cls), st.P, st.Ncap,
              _ptr(st.inst), _ptr(st.inst_cnt), float(fg_thr), _ptr(labels), _ptr(weights), _ptr(targets), _stream())
    return labels, weights, targets
