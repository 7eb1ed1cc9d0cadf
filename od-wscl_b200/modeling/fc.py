"""The fully-connected block of the path on the hand-written tcgen05 GEMM (csrc/fc_gemm.cu): nn.Linear (+ ReLU
(+ Dropout)) forward, input gradient and weight gradient of fc6 / fc7 (modeling/backbone/vgg16.py:122-130), Sim_Net
(roi_heads/sim_head/sim_net.py:10-26) and the MIST predictor heads (roi_heads/weak_head/roi_weak_predictors.py:158-165).
The nn.Linear modules only hold the parameters (the reference's state-dict keys); no cuBLAS call is left on the path.

  forward   Y = act(X W^T + b)     one launch: bias, ReLU, Dropout (Philox, no mask tensor) and TF32 rounding fused
  backward  dZ = dY * act'(Y)      one pass (fused into the producing dgrad's epilogue when the producer is ours)
            dX = dZ W              W read where it lies as an MN-major operand
            dW = dZ^T X            both operands MN-major straight from the activations; a second (small-batch) call of
                                   the same layer is folded in with the accumulate epilogue instead of a separate
                                   gradient + add (the two fc6 calls of a step, weak_head.py:107-112 / loss.py:299-310)
            db = column sums of dZ

`strict` evaluates every product as a 3xTF32 split (hi*hi + hi*lo + lo*hi: fp32-class accuracy) for the parity tests;
the default is single-pass TF32, the arithmetic torch 1.7.1 (the reference's pin) runs nn.Linear with on tensor-core GPUs.
"""
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import capi

ACT_NONE, ACT_RELU, ACT_RELU_DROPOUT = 0, 1, 2
FUSE_ACT_BWD = True        # fold the ReLU/Dropout derivative of a layer into the dgrad epilogue of its (only) consumer
_seed_state = {"ctr": 0}


def next_dropout_seed():
    """Philox key of one Dropout call: (process seed, rank, call counter) -- no host RNG round trip, distinct per rank
    (the reference leaves DDP ranks independently seeded)."""
    _seed_state["ctr"] += 1
    rank = int(os.environ.get("RANK", "0"))
    base = torch.initial_seed() & 0xFFFFFFFFFFFF
    return (base * 0x9E3779B97F4A7C15 + rank * 0xBF58476D1CE4E5B9 + _seed_state["ctr"] * 0x94D049BB133111EB) & (2 ** 64 - 1)


STRICT_K_CHUNK = 2048      # contraction length per TMEM accumulation in strict mode (see _gemm)


def _gemm(A, B, strict, **kw):
    """fc_gemm, or its fp32-class form for the parity tests: every product as a 3xTF32 hi/lo split (hi*hi + hi*lo +
    lo*hi; the dropped lo*lo term is <= 2^-22 relative), and the contraction cut into chunks of STRICT_K_CHUNK whose
    partial results are added in the epilogue (round-to-nearest fp32) -- the tensor core adds into its TMEM accumulator
    with truncation, which over K = 25088 terms costs ~1e-5 relative (measured), too much for a 1e-4 end-to-end gate.
    The bias joins the first launch, the non-linear part of the epilogue the last."""
    if not strict:
        return capi.fc_gemm(A, B, **kw)
    a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
    Ah, Al = capi.split_tf32(A if A.is_contiguous() else A.contiguous())
    Bh, Bl = capi.split_tf32(B if B.is_contiguous() else B.contiguous())
    K = A.shape[0] if a_mn else A.shape[1]
    out, first = kw.get("out"), True
    passes = []
    for k0 in range(0, max(K, 1), STRICT_K_CHUNK):
        k1 = min(K, k0 + STRICT_K_CHUNK)
        ka = (lambda t: t[k0:k1]) if a_mn else (lambda t: t[:, k0:k1])
        kb = (lambda t: t[k0:k1]) if b_mn else (lambda t: t[:, k0:k1])
        passes += [(ka(Ah), kb(Bh)), (ka(Ah), kb(Bl)), (ka(Al), kb(Bh))]
    for i, (a, b) in enumerate(passes):
        last = i == len(passes) - 1
        extra = dict(relu=kw.get("relu", False), dropout_p=kw.get("dropout_p", 0.0), seed=kw.get("seed", 0),
                     mask_src=kw.get("mask_src"), mask_scale=kw.get("mask_scale", 1.0)) if last else {}
        out = capi.fc_gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out, bias=kw.get("bias") if first else None,
                           accumulate=kw.get("accumulate", False) if first else True, **extra)
        first = False
    return out


class _LinearFn(Function):
    """y = act(x W^T + b).  `stash` / `role` fold the weight gradients of two calls of the SAME layer into one tensor:
    the "small" call's backward (it runs first: its node is younger) only stashes (dZ, x), the "main" call's backward
    accumulates them into its own dW / db through the GEMM's accumulate epilogue.  If the order is ever the other way
    round the small call returns its own gradient, so the result never depends on the assumption."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, p, seed, round_out, strict, stash, role, in_mask_scale, act_bwd_fused, peer):
        x2 = x
        ctx.peer = peer
        y = _gemm(x2, weight, strict, bias=bias, relu=act != ACT_NONE,
                  dropout_p=p if act == ACT_RELU_DROPOUT else 0.0, seed=seed, round_tf32=round_out and not strict)
        ctx.save_for_backward(x2, weight, y if act != ACT_NONE else None)
        ctx.act, ctx.p, ctx.strict, ctx.stash, ctx.role = act, (p if act == ACT_RELU_DROPOUT else 0.0), strict, stash, role
        ctx.has_bias = bias is not None
        ctx.in_mask_scale, ctx.act_bwd_fused = in_mask_scale, act_bwd_fused
        if role == "main" and stash is not None:
            stash["has_main"] = True
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight, y = ctx.saved_tensors
        if ctx.act != ACT_NONE and not ctx.act_bwd_fused:
            dz = capi.relu_dropout_backward(y, g, ctx.p)              # dY * [y > 0] / (1 - p)
        else:
            dz = g                                                    # the consumer's dgrad epilogue already applied it
        gx = None
        if ctx.needs_input_grad[0]:
            if ctx.in_mask_scale is not None:                         # x is the producer's ReLU(+Dropout) output
                gx = _gemm(dz, weight, ctx.strict, b_mn=True, mask_src=x, mask_scale=ctx.in_mask_scale)
            else:
                gx = _gemm(dz, weight, ctx.strict, b_mn=True)
        st = ctx.stash
        need_w = ctx.needs_input_grad[1]
        if ctx.role == "small" and st is not None and st.get("has_main", False) and not st.get("main_done", False):
            st.setdefault("pending", []).append((dz, x))
            return (gx,) + (None,) * 12
        gw = gb = None
        pend = st.pop("pending", []) if (ctx.role == "main" and st is not None) else []
        if need_w:
            peer = ctx.peer
            if peer is not None and len(pend) <= 1 and not ctx.strict:
                # multi-GPU: the product is added to EVERY rank's gradient replica from the GEMM epilogue (NVSwitch multicast
                # reduction) -- this weight has no all-reduce (sharding.PeerGradSum)
                peer.sum_product(dz, x, pend[0] if pend else None)
                gw = peer.grad()
            elif len(pend) == 1 and not ctx.strict:
                # the small call's (dZ, x) ride in the same accumulation as a second operand pair: one GEMM, one tensor
                gw = capi.fc_gemm(dz, x, a_mn=True, b_mn=True, A2=pend[0][0], B2=pend[0][1])
            else:
                gw = _gemm(dz, x, ctx.strict, a_mn=True, b_mn=True)
                for dzs, xs in pend:
                    _gemm(dzs, xs, ctx.strict, a_mn=True, b_mn=True, out=gw, accumulate=True)
                if peer is not None:
                    peer.add(gw)
                    gw = peer.grad()
            if ctx.has_bias:
                gb = capi.colsum(dz)
                for dzs, _ in pend:
                    capi.colsum(dzs, out=gb, accumulate=True)
        if ctx.role == "main" and st is not None:
            st["main_done"] = True
        return (gx, gw, gb) + (None,) * 10


def linear(x, weight, bias=None, act=ACT_NONE, p=0.0, seed=0, round_out=False, strict=False, stash=None, role=None,
           in_mask_scale=None, act_bwd_fused=False):
    """act(x @ weight.T + bias) on the sm_100a fc kernel.  CUDA fp32 only: there is no CPU / library fallback.
    in_mask_scale (float): `x` is the ReLU(+Dropout) output of a layer created with act_bwd_fused=True whose ONLY
    consumer is this call -- the input gradient leaves this layer's dgrad epilogue already multiplied by that layer's
    derivative (in_mask_scale * [x > 0]), and the producer's backward skips its own elementwise pass."""
    if not x.is_cuda:
        raise RuntimeError("the fully-connected block has no CPU implementation (sm_100a kernels only)")
    if x.dim() != 2:
        x = x.reshape(x.shape[0], -1)
    if not FUSE_ACT_BWD and (in_mask_scale is not None or act_bwd_fused):
        raise RuntimeError("fused activation backward requested while fc.FUSE_ACT_BWD is off")
    return _LinearFn.apply(x.float(), weight, bias, int(act), float(p), int(seed), bool(round_out), bool(strict), stash, role,
                           None if in_mask_scale is None else float(in_mask_scale), bool(act_bwd_fused),
                           getattr(weight, "_odw_peer", None))
