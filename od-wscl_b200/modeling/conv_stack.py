"""The VGG16-OICR conv body (modeling/backbone/vgg16.py:26-36,58-83) executed by the hand-written sm_100a
kernels of csrc/conv3x3.cu, channels-last end to end, as ONE autograd node:

  forward   conv1_1 (FFMA, NCHW image -> NHWC) then 12 tcgen05 implicit-GEMM convolutions with fused
            bias + ReLU and three NHWC max-pools; the result is returned as an NCHW-shaped view of the
            NHWC buffer (channels_last strides), which the ROIPool binding consumes without a transpose.
  backward  per trainable layer: DGRAD = the same tcgen05 kernel on the flipped / transposed weights with the
            previous activation's ReLU derivative fused into its epilogue; the max-pool backward fuses the
            same mask; WGRAD + bias gradient = the tcgen05 pixel-contraction kernel (see `wgrad`).
  Frozen layers (FREEZE_CONV_BODY_AT, vgg16.py:48-55) are neither differentiated nor kept alive.

`strict=True` (tests) evaluates every convolution as a 3-pass TF32 hi/lo split (fp32-accurate); the default
is single-pass TF32 with fp32 accumulation -- what the reference's pinned torch 1.7.1 runs on tensor-core
GPUs (cudnn.allow_tf32 defaults to True).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import capi

OVERLAP_WGRAD = True       # run each layer's WGRAD on a side stream next to the DGRAD chain

# (cout, dilation, relu, pool_after) for the 13 convolutions of vgg_cfg["VGG16-OICR"]
LAYERS = [(64, 1, True, False), (64, 1, True, True), (128, 1, True, False), (128, 1, True, True),
          (256, 1, True, False), (256, 1, True, False), (256, 1, True, True),
          (512, 1, True, False), (512, 1, True, False), (512, 1, True, False),
          (512, 2, True, False), (512, 2, True, False), (512, 2, False, False)]


def _conv(x, wk, bias, dil, flags, strict, mask_src=None, feeds_conv=True, w_rounded=False):
    """One convolution.  Single-pass mode: `x` is already TF32-rounded by its producer, the weights are rounded
    here (or by conv_weight_xform: w_rounded), and the output is rounded in the epilogue when another convolution
    consumes it (the tensor core truncates fp32 operands; rounding first is what cuDNN's TF32 kernels do on load)."""
    if not strict:
        if not w_rounded:
            capi.round_tf32_(wk)
        return capi.conv3x3_nhwc(x, wk, bias, dilation=dil, flags=flags | (capi.CONV_ROUND if feeds_conv else 0),
                                 mask_src=mask_src)
    xh, xl = capi.split_tf32(x)
    wh, wl = capi.split_tf32(wk)
    y = capi.conv3x3_nhwc(xh, wh, bias, dilation=dil)
    capi.conv3x3_nhwc(xh, wl, None, dilation=dil, flags=capi.CONV_ACCUM, out=y)
    return capi.conv3x3_nhwc(xl, wh, None, dilation=dil, flags=capi.CONV_ACCUM | flags, mask_src=mask_src, out=y)


def wgrad(x_nhwc, dz_nhwc, w_shape, dil, strict):
    """dW [Cout,Cin,3,3] and db [Cout] of one layer from its NHWC input and the NHWC gradient of its
    pre-activation output (tcgen05 WGRAD kernel; both operands are already TF32-rounded in single-pass mode)."""
    if not strict:
        dw, db = capi.conv3x3_wgrad_nhwc(x_nhwc, dz_nhwc, dilation=dil)
    else:
        xh, xl = capi.split_tf32(x_nhwc)
        zh, zl = capi.split_tf32(dz_nhwc)
        dw, _ = capi.conv3x3_wgrad_nhwc(xh, zh, dilation=dil, want_bias=False)
        capi.conv3x3_wgrad_nhwc(xh, zl, dilation=dil, accumulate_into=dw, want_bias=False)
        capi.conv3x3_wgrad_nhwc(xl, zh, dilation=dil, accumulate_into=dw, want_bias=False)
        db = dz_nhwc.sum((0, 1, 2))
    return dw.permute(0, 3, 1, 2).contiguous(), db          # [Cout,3,3,Cin] -> torch's [Cout,Cin,3,3]


_side_streams = {}
_frozen_xform = {}         # (data_ptr, version, shape) -> fprop operand of a FROZEN layer (its weights never change)


def _side_stream(device):
    """One extra stream per device: the WGRAD of a layer only depends on that layer's dZ, so it runs beside the DGRAD
    chain and fills the SMs the DGRAD's last partial wave leaves idle (304 tiles on 148 SMs = 2.05 waves)."""
    s = _side_streams.get(device.index)
    if s is None:
        s = _side_streams[device.index] = torch.cuda.Stream(device=device)
    return s


class _VGGStackFn(Function):
    @staticmethod
    def forward(ctx, x, strict, *params):
        n = len(LAYERS)
        ws, bs = params[0::2], params[1::2]
        needs_w = [ctx.needs_input_grad[2 + 2 * i] or ctx.needs_input_grad[3 + 2 * i] for i in range(n)]
        first_train = min([i for i in range(n) if needs_w[i]], default=n)
        saved = {}                  # tensors the backward needs, by name
        wd_ops = {}                 # dgrad weight operands prepared in the forward (single-pass mode)
        a = capi.conv3x3_c3(x, ws[0], bs[0], relu=LAYERS[0][2], round_tf32=not strict)
        if first_train == 0:
            saved["x"] = x
        for i in range(1, n + 1):
            # `a` is the (post-ReLU) output of layer i-1; pool it if the table says so, then feed layer i
            cout, dil, relu, pool = LAYERS[i - 1]
            if pool:
                if i - 1 >= first_train:
                    saved["a%d" % (i - 1)] = a          # pre-pool activation: pool backward + ReLU mask
                a = capi.maxpool2x2_nhwc(a)
            if i == n:
                break
            if i >= first_train:
                saved["in%d" % i] = a                   # input of trainable layer i (wgrad; mask of layer i-1)
            if strict:
                wk = ws[i].permute(0, 2, 3, 1).contiguous()
            elif not needs_w[i] and i < first_train:
                # frozen layer (FREEZE_CONV_BODY_AT): the operand layout is computed once and kept
                key = (ws[i].data_ptr(), ws[i]._version, tuple(ws[i].shape))
                wk = _frozen_xform.get(key)
                if wk is None:
                    for k_old in [k for k in _frozen_xform if k[0] == key[0]]:
                        del _frozen_xform[k_old]
                    wk, _ = capi.conv_weight_xform(ws[i], want_fwd=True, want_dgrad=False)
                    _frozen_xform[key] = wk
            else:       # both operand layouts (fprop now, tap-flipped dgrad later) TF32-rounded in one pass over W
                wk, wd_ops[i] = capi.conv_weight_xform(ws[i], want_fwd=True, want_dgrad=i > first_train)
            a = _conv(a, wk, bs[i], LAYERS[i][1], capi.CONV_RELU if LAYERS[i][2] else 0, strict, feeds_conv=i < n - 1,
                      w_rounded=not strict)
        ctx.wd_ops = wd_ops
        ctx.strict, ctx.first_train, ctx.needs_w = strict, first_train, needs_w
        ctx.names = list(saved)
        ctx.save_for_backward(*saved.values(), *[ws[i] for i in range(n)])
        return a                                        # [B,Hf,Wf,512] NHWC

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        n = len(LAYERS)
        tensors = ctx.saved_tensors
        saved = dict(zip(ctx.names, tensors[:len(ctx.names)]))
        ws = tensors[len(ctx.names):]
        grads = [None] * (2 * n)
        dz = g.contiguous()                             # gradient of layer n-1's pre-activation (no ReLU after conv5_3)
        if not ctx.strict:
            dz = capi.round_tf32_(dz.clone() if dz.data_ptr() == g.data_ptr() else dz)
        main = torch.cuda.current_stream(dz.device)
        side = _side_stream(dz.device) if OVERLAP_WGRAD else None
        for i in range(n - 1, ctx.first_train - 1, -1):
            cout, dil, relu, pool = LAYERS[i]
            if i == 0:
                xin = saved["x"]
                _, dw, db = torch.ops.aten.convolution_backward(dz.permute(0, 3, 1, 2), xin, ws[0], [cout], [1, 1], [1, 1],
                                                                [1, 1], False, [0, 0], 1, [False, True, True])
                grads[0], grads[1] = dw.contiguous(), db
                break
            xin = saved["in%d" % i]
            if ctx.needs_w[i]:
                if side is None or i == ctx.first_train:
                    grads[2 * i], grads[2 * i + 1] = wgrad(xin, dz, ws[i].shape, dil, ctx.strict)
                else:
                    side.wait_stream(main)              # dz (and everything before it) is ready
                    with torch.cuda.stream(side):
                        dw, db = wgrad(xin, dz, ws[i].shape, dil, ctx.strict)
                    for t_ in (dz, xin):
                        t_.record_stream(side)          # the caching allocator must not recycle them under the side stream
                    for t_ in (dw, db):
                        t_.record_stream(main)
                    grads[2 * i], grads[2 * i + 1] = dw, db
            if i == ctx.first_train:
                break
            # DGRAD: d(input of layer i).  The input is either layer i-1's post-ReLU output (mask fused in the
            # epilogue) or its max-pooled version (mask fused into the pool backward).
            wd = ctx.wd_ops.get(i)
            pre = wd is not None
            if not pre:
                wd = ws[i].flip(2, 3).permute(1, 2, 3, 0).contiguous()      # [Cin,3,3,Cout], taps flipped
            prev_pool = LAYERS[i - 1][3]
            if prev_pool:
                dp = _conv(dz, wd, None, dil, 0, ctx.strict, w_rounded=pre)
                dz = capi.maxpool2x2_nhwc_bwd(saved["a%d" % (i - 1)], dp, relu_mask=LAYERS[i - 1][2])
            elif LAYERS[i - 1][2]:
                dz = _conv(dz, wd, None, dil, capi.CONV_MASK, ctx.strict, mask_src=xin, w_rounded=pre)
            else:
                dz = _conv(dz, wd, None, dil, 0, ctx.strict, w_rounded=pre)
        if side is not None:
            main.wait_stream(side)                      # gradients are consumed (DDP hooks, optimizer) on the main stream
        return (None, None) + tuple(grads)


def vgg_stack(x, convs, strict=False):
    """x [B,3,H,W] NCHW image batch; convs = the 13 nn.Conv2d parameter holders.  Returns the stride-8 feature map
    as an NCHW-shaped tensor with channels-last memory."""
    if not x.is_cuda:
        raise RuntimeError("the conv stack has no CPU implementation (sm_100a kernels only)")
    params = []
    for c in convs:
        params += [c.weight, c.bias]
    y = _VGGStackFn.apply(x.float().contiguous(), bool(strict), *params)
    return y.permute(0, 3, 1, 2)
