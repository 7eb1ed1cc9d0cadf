"""CPU oracle for the OD-WSCL proposal-feature hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package
(``od-wscl_b200/``).  Allowed importers: ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, as the checker.

What it is: a restatement, on the CPU, of what the reference *executes* on the
hot path.  Integer / index / selection
arithmetic is plain C (``odwscl_oracle.c``) or explicit Python loops over numpy /
torch CPU scalars; dense fp32 contractions (conv, linear, matmul, softmax, CE)
call ``torch`` on the CPU -- the very library the reference calls for them
(``torch`` pinned 1.7.1 by the reference's README.md:23,32; call sites
modeling/backbone/vgg16.py:35,151,161, roi_heads/sim_head/sim_net.py:26,
roi_heads/weak_head/loss.py:319-320, roi_heads/sim_head/sim_loss.py:60).

Pinning: the reference ships no tests or golden vectors.
The oracle is pinned against outputs of the reference's own Python imported in
the build container (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``) and
against ``torchvision`` (the reference's third-party NMS / ROIPool lineage);
``tests/test_oracle_golden.py`` is the check.

All file:line citations are relative to ``/root/reference/wetectron``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborc.so")
_SRC = os.path.join(_HERE, "odwscl_oracle.c")
_lib_cache = None


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/liborc.so (git-ignored)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
             "-fvisibility=hidden", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib_cache
    if _lib_cache is None:
        _lib_cache = ctypes.CDLL(build())
        _lib_cache.orc_nms_tv_f32.restype = ctypes.c_int
        _lib_cache.orc_nms_legacy_f32.restype = ctypes.c_int
    return _lib_cache


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- #
# A3/A4  ROIPool (csrc/cuda/ROIPool_cuda.cu:16-108)
# --------------------------------------------------------------------------- #
def roi_pool_forward(feat, rois, scale: float, ph: int, pw: int):
    feat, rois = _f32(feat), _f32(rois)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.empty((R, C, ph, pw), np.float32)
    arg = np.empty((R, C, ph, pw), np.int32)
    lib().orc_roi_pool_fwd_f32(_p(feat), B, C, H, W, _p(rois), R, ctypes.c_float(scale), ph, pw,
                               _p(out), _p(arg))
    return out, arg


def roi_pool_backward(grad_out, argmax, rois, B, C, H, W):
    grad_out, rois = _f32(grad_out), _f32(rois)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    R, _, ph, pw = grad_out.shape
    gin = np.empty((B, C, H, W), np.float32)
    lib().orc_roi_pool_bwd_f32(_p(grad_out), _p(argmax), _p(rois), R, B, C, H, W, ph, pw, _p(gin))
    return gin


# A5  ROIAlign (csrc/cuda/ROIAlign_cuda.cu:64-122,177-254)
def roi_align_forward(feat, rois, scale: float, ph: int, pw: int, sampling_ratio: int):
    feat, rois = _f32(feat), _f32(rois)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.empty((R, C, ph, pw), np.float32)
    lib().orc_roi_align_fwd_f32(_p(feat), B, C, H, W, _p(rois), R, ctypes.c_float(scale), ph, pw,
                                sampling_ratio, _p(out))
    return out


def roi_align_backward(grad_out, rois, scale, B, C, H, W, sampling_ratio):
    grad_out, rois = _f32(grad_out), _f32(rois)
    R, _, ph, pw = grad_out.shape
    gin = np.empty((B, C, H, W), np.float32)
    lib().orc_roi_align_bwd_f32(_p(grad_out), _p(rois), R, ctypes.c_float(scale), ph, pw, B, C, H, W,
                                sampling_ratio, _p(gin))
    return gin


# --------------------------------------------------------------------------- #
# A9/A10  IoU and NMS
# --------------------------------------------------------------------------- #
def box_iou(a, b, plus_one: bool = True):
    """structures/boxlist_ops.py:127-160 (plus_one) / torchvision IoU (not)."""
    a, b = _f32(a).reshape(-1, 4), _f32(b).reshape(-1, 4)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().orc_box_iou_f32(_p(a), a.shape[0], _p(b), b.shape[0], int(plus_one), _p(out))
    return out


def nms_tv(boxes, scores, thr: float):
    """torchvision.ops.nms semantics (structures/boxlist_ops.py:57)."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    keep = np.empty((boxes.shape[0],), np.int64)
    n = lib().orc_nms_tv_f32(_p(boxes), _p(scores), boxes.shape[0], ctypes.c_float(thr), _p(keep))
    return keep[:n].copy()


def nms_legacy(boxes, scores, thr: float, ge: bool):
    """`_C.nms` (csrc/nms.h:10-28): ge=True -> cpu/nms_cpu.cpp, ge=False -> cuda/nms.cu."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    keep = np.empty((boxes.shape[0],), np.int64)
    n = lib().orc_nms_legacy_f32(_p(boxes), _p(scores), boxes.shape[0], ctypes.c_float(thr), int(ge),
                                 _p(keep))
    return keep[:n].copy()


def cal_iou(boxes, m: int, thres: float):
    """utils/utils.py:23-26: ascending ids j with IoU+1(P[j], P[m]) >= thres."""
    boxes = _f32(boxes)
    row = box_iou(boxes, boxes[m:m + 1], True)[:, 0]
    return np.nonzero(row >= np.float32(thres))[0].astype(np.int64)


def easy_nms(boxes, cluster, score_col, thr: float):
    """utils/utils.py:29-33: torchvision NMS on the `cluster` subset, re-mapped."""
    cluster = np.asarray(cluster, np.int64)
    if cluster.size == 0:
        return cluster
    keep = nms_tv(_f32(boxes)[cluster], _f32(score_col)[cluster], thr)
    return cluster[keep]


# --------------------------------------------------------------------------- #
# A11  object discovery (roi_heads/weak_head/loss.py:271-345, Appendix A)
# --------------------------------------------------------------------------- #
def _first_argmax(col: torch.Tensor) -> int:
    return int(torch.argmax(col).item())


def discover(boxes: Sequence[torch.Tensor], final_score: Sequence[torch.Tensor],
             ref_logits: Sequence[Sequence[torch.Tensor]], sim_feature: Sequence[torch.Tensor],
             pos_classes: Sequence[Sequence[int]], embed_aug: Callable, thres: float, nms: float,
             num_classes: int):
    """Phase A + Phase B of the contrastive section.

    boxes[b] [N,4]; final_score[b] [N,C]; ref_logits[i][b] [N,C] raw logits of
    branch i; sim_feature[b] [N,128] (may require grad); pos_classes[b] ascending
    0-based class ids; embed_aug(b, c, I, kind) -> [k,128] embeddings of the
    'drop' / 'noise' augmented positives (loss.py:298-305).
    Returns bank (list per class of row tensors), Wt, inst[b][i][c] (np int64),
    idx[b][c] and a trace dict used by stage-wise parity tests.
    """
    B = len(boxes)
    nc = num_classes - 1
    idx = [[np.zeros((0,), np.int64) for _ in range(nc)] for _ in range(B)]
    bank: List[List[torch.Tensor]] = [[] for _ in range(nc)]
    Wt: List[torch.Tensor] = []
    trace = {"phaseA_idx": {}, "tau": {}, "close": {}, "sim_rows": {}, "new": {}, "argmax": {}}

    def src(b, i):      # loss.py:283,313
        return final_score[b] if i == 0 else F.softmax(ref_logits[i - 1][b], dim=1)

    # ---- Phase A: loss.py:281-307
    for b in range(B):
        P = boxes[b].detach().numpy()
        for i in range(3):
            pscore = src(b, i)[:, 1:].detach()
            for c in pos_classes[b]:
                m = _first_argmax(pscore[:, c])
                trace["argmax"][(b, i, c)] = m
                nb = cal_iou(P, m, thres)
                idx[b][c] = np.unique(np.concatenate([idx[b][c], nb]))
        for c in pos_classes[b]:
            I = idx[b][c]
            trace["phaseA_idx"][(b, c)] = I.copy()
            It = torch.from_numpy(I)
            S = final_score[b]
            h = S[It, c + 1] / S[:, c + 1].sum()                       # loss.py:294
            bank[c].append(sim_feature[b][It]); Wt.append(h)
            bank[c].append(embed_aug(b, c, It, "drop")); Wt.append(h)  # loss.py:298-301
            bank[c].append(embed_aug(b, c, It, "noise")); Wt.append(h) # loss.py:303-305
    coll = [torch.cat(rows).detach().clone() if rows else None for rows in bank]   # loss.py:307

    # ---- Phase B: loss.py:311-345
    inst = [[[np.zeros((0,), np.int64) for _ in range(nc)] for _ in range(3)] for _ in range(B)]
    for b in range(B):
        P = boxes[b].detach().numpy()
        Fb = sim_feature[b]
        sim_mat = torch.mm(Fb.detach(), Fb.detach().T)                  # loss.py:319
        pos = list(pos_classes[b])
        for i in range(3):
            pscore = src(b, i)[:, 1:].detach()
            for c in pos:
                m = _first_argmax(pscore[:, c])
                tau = torch.mm(Fb.detach()[m].view(1, -1), coll[c].T).mean()    # loss.py:320
                close = torch.ge(sim_mat[m], tau)                       # loss.py:324/330
                if len(pos) > 1:
                    for n_c in pos:
                        if n_c == c:
                            continue
                        mn = _first_argmax(pscore[:, n_c])
                        close = torch.ge(close, sim_mat[mn])            # loss.py:327 (bool vs float quirk)
                close_i = close.nonzero(as_tuple=False).view(-1).numpy().astype(np.int64)
                trace["tau"][(b, i, c)] = float(tau)
                trace["sim_rows"][(b, i, c)] = sim_mat[m].numpy().copy()
                trace["close"][(b, i, c)] = close_i.copy()
                kept = easy_nms(P, close_i, pscore[:, c].numpy(), nms)  # loss.py:332
                if kept.size == 0:                                      # loss.py:333
                    kept = np.array([m], np.int64)
                inst[b][i][c] = np.concatenate([inst[b][i][c], kept])   # loss.py:334
                new = np.setdiff1d(kept, idx[b][c])                     # loss.py:336-337 (sorted)
                if new.size == 0:                                       # loss.py:338
                    new = np.array([m], np.int64)
                trace["new"][(b, i, c)] = new.copy()
                nt = torch.from_numpy(new)
                bank[c].append(Fb[nt])                                  # loss.py:340
                idx[b][c] = np.unique(np.concatenate([idx[b][c], new])) # loss.py:341
                S = final_score[b]
                Wt.append((S[nt, c + 1] / S[:, c + 1].sum()).view(-1))  # loss.py:343-345
    return bank, Wt, inst, idx, trace


# --------------------------------------------------------------------------- #
# A12  SupConLossV2 (roi_heads/sim_head/sim_loss.py:49-80)
# --------------------------------------------------------------------------- #
def supcon_v2(features: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor,
              temperature: float) -> torch.Tensor:
    sim = torch.div(torch.matmul(features, features.T), temperature)
    row_max, _ = torch.max(sim, dim=1, keepdim=True)
    sim = sim - row_max.detach()
    logits_mask = torch.ones_like(sim)
    logits_mask.fill_diagonal_(0)
    exp_sim = torch.exp(sim)
    label_mask = torch.eq(labels.view(-1, 1), labels.view(-1, 1).T).float()
    mask = logits_mask * label_mask
    log_prob = torch.log((exp_sim * mask).sum(1) / (exp_sim * logits_mask).sum(1))
    return (-log_prob * weights.detach()).mean()


def supcon_from_bank(bank, Wt, temperature: float):
    feats, labels = [], []
    for c, rows in enumerate(bank):            # class-major concat, sim_loss.py:55-58
        if rows:
            f = torch.cat(rows)
            if f.shape[0]:
                feats.append(f)
                labels.append(torch.full((f.shape[0],), float(c)))
    feats = torch.cat(feats)
    labels = torch.cat(labels)
    w = torch.cat([x.view(-1) for x in Wt]).detach()   # execution order (misaligned on purpose)
    return supcon_v2(feats, labels, w, temperature), feats, labels, w


# --------------------------------------------------------------------------- #
# A13  od_layer (roi_heads/weak_head/pseudo_label_generator.py:135-197)
# --------------------------------------------------------------------------- #
def box_encode(gt: np.ndarray, prop: np.ndarray, weights=(10.0, 10.0, 5.0, 5.0)) -> torch.Tensor:
    """modeling/box_coder.py:22-50 (+1 widths)."""
    g, p = torch.as_tensor(gt, dtype=torch.float32), torch.as_tensor(prop, dtype=torch.float32)
    ew = p[:, 2] - p[:, 0] + 1; eh = p[:, 3] - p[:, 1] + 1
    ex = p[:, 0] + 0.5 * ew; ey = p[:, 1] + 0.5 * eh
    gw = g[:, 2] - g[:, 0] + 1; gh = g[:, 3] - g[:, 1] + 1
    gx = g[:, 0] + 0.5 * gw; gy = g[:, 1] + 0.5 * gh
    wx, wy, ww, wh = weights
    return torch.stack((wx * (gx - ex) / ew, wy * (gy - ey) / eh,
                        ww * torch.log(gw / ew), wh * torch.log(gh / eh)), dim=1)


def od_layer(boxes: torch.Tensor, source_score: torch.Tensor, pos: Sequence[int],
             inst_i: Sequence[np.ndarray], fg_thresh: float = 0.5):
    P = boxes.detach().numpy()
    prob = source_score[:, 1:].detach().clone()
    gtb, gtc, gts = [], [], []
    for c in pos:
        col = prob[:, c]
        m = _first_argmax(col)
        sb = inst_i[c]
        if sb.size == 0:
            gtb.append(P[m:m + 1]); gtc.append(np.array([c + 1])); gts.append(col[m:m + 1].numpy().copy())
        else:
            gtb.append(P[sb]); gtc.append(np.full(sb.shape, c + 1)); gts.append(col[torch.from_numpy(sb)].numpy().copy())
        prob[m].fill_(0)                                   # :165 zero the whole row
    N = P.shape[0]
    if not gtb:
        return torch.zeros(N, dtype=torch.long), torch.zeros(N), torch.zeros(N, 4)
    gtb = np.concatenate(gtb).astype(np.float32); gtc = np.concatenate(gtc); gts = np.concatenate(gts)
    ov = box_iou(P, gtb, True)
    mx = ov.max(axis=1); am = ov.argmax(axis=1)            # numpy first-max, :176-177
    labels = torch.from_numpy(gtc[am].astype(np.int64))
    weights = torch.from_numpy(gts[am].astype(np.float32))
    labels[torch.from_numpy(mx <= np.float32(fg_thresh))] = 0     # :183 (le)
    targets = box_encode(gtb[am], P)
    return labels, weights, targets


# --------------------------------------------------------------------------- #
# A15  DropBlock2D with an injected centre mask (modeling/dropblock/drop_block.py:29-66)
# --------------------------------------------------------------------------- #
def dropblock(x: torch.Tensor, centre_mask: torch.Tensor, block_size: int) -> torch.Tensor:
    bm = F.max_pool2d(centre_mask[:, None].float(), kernel_size=block_size, stride=1,
                      padding=block_size // 2)
    bm = 1 - bm.squeeze(1)
    out = x * bm[:, None]
    return out * bm.numel() / bm.sum()


def smooth_l1(x, t, beta=1.0):
    """layers/smooth_l1_loss.py:4-16, reduction=False."""
    n = torch.abs(x - t)
    return torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)


# --------------------------------------------------------------------------- #
# RoIRegLossComputation.__call__ (roi_heads/weak_head/loss.py:233-411)
# --------------------------------------------------------------------------- #
def roi_reg_loss(cls_logit, det_logit, ref_logits, bbox_preds, sim_feature, boxes, labels_per_img,
                 embed_aug, *, thres=0.5, nms=0.1, lmda=0.03, temp=0.2, epsilon=1e-8, return_trace=False):
    """cls_logit/det_logit [R,C]; ref_logits/bbox_preds 3x[R,C]/[R,4C]; sim_feature [R,128];
    boxes: list of [N_b,4]; labels_per_img: list of int arrays (1-based class ids)."""
    sizes = [b.shape[0] for b in boxes]
    C = cls_logit.shape[1]
    cls = F.softmax(cls_logit, dim=1)
    det = torch.cat([F.softmax(d, dim=0) for d in det_logit.split(sizes)])
    final = cls * det
    final_l = final.split(sizes)
    ref_l = [r.split(sizes) for r in ref_logits]
    box_l = [r.split(sizes) for r in bbox_preds]
    simf = sim_feature.split(sizes)
    pos = []
    img_labels = []
    for lab in labels_per_img:
        v = torch.zeros(C); v[torch.as_tensor(np.unique(lab)).long()] = 1; v[0] = 0
        img_labels.append(v)
        pos.append([int(x) for x in v[1:].eq(1).nonzero(as_tuple=False)[:, 0]])
    bank, Wt, inst, idx, trace = discover(boxes, final_l, ref_l, simf, pos, embed_aug, thres, nms, C)
    loss_sim, feats, flabels, w = supcon_from_bank(bank, Wt, temp)
    losses = {"loss_img": 0.0}
    for i in range(3):
        losses["loss_ref_cls%d" % i] = 0.0
        losses["loss_ref_reg%d" % i] = 0.0
    losses["loss_sim"] = lmda * loss_sim
    pseudo = {}
    for b in range(len(boxes)):
        img_score = torch.clamp(final_l[b].sum(0), min=epsilon, max=1 - epsilon)
        losses["loss_img"] = losses["loss_img"] + F.binary_cross_entropy(img_score, img_labels[b].clamp(0, 1))
        for i in range(3):
            source = final_l[b] if i == 0 else F.softmax(ref_l[i - 1][b], dim=1)
            pl, lw, rt = od_layer(boxes[b], source, pos[b], inst[b][i])
            pseudo[(b, i)] = (pl, lw, rt)
            lm = 3 if i == 0 else 1
            losses["loss_ref_cls%d" % i] = losses["loss_ref_cls%d" % i] + lm * torch.mean(
                F.cross_entropy(ref_l[i][b], pl, reduction="none") * lw)
            pi = torch.nonzero(pl > 0, as_tuple=False).squeeze(1)
            mp = 4 * pl[pi][:, None] + torch.tensor([0, 1, 2, 3])
            reg = lm * torch.sum(smooth_l1(box_l[i][b][pi[:, None], mp], rt[pi]) * lw[pi, None])
            losses["loss_ref_reg%d" % i] = losses["loss_ref_reg%d" % i] + reg / pl.numel()
    for k in losses:
        if "sim" not in k:
            losses[k] = losses[k] / len(boxes)
    if return_trace:
        return losses, dict(inst=inst, idx=idx, trace=trace, bank_feats=feats, bank_labels=flabels,
                            bank_w=w, pseudo=pseudo, final=final_l)
    return losses


# --------------------------------------------------------------------------- #
# A1  VGG16-OICR conv stack (modeling/backbone/vgg16.py:26-36,58-93)
# --------------------------------------------------------------------------- #
VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "I", "512-D", "512-D", "512-D"]
CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]   # nn.Sequential indices (state-dict keys)


def vgg16_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix="backbone.body.features.") -> torch.Tensor:
    li = 0
    convs = [v for v in VGG_CFG if v not in ("M", "I")]
    k = 0
    for v in VGG_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2); li += 1
        elif v == "I":
            li += 1
        else:
            dil = 2 if isinstance(v, str) else 1
            w, b = sd[prefix + "%d.weight" % li], sd[prefix + "%d.bias" % li]
            x = F.conv2d(x, w, b, padding=dil, dilation=dil)
            k += 1
            if k < len(convs):          # the last ReLU is dropped (vgg16.py:82-83)
                x = F.relu(x)
            li += 2
    return x


class _RoiPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois, scale, ph, pw):
        out, arg = roi_pool_forward(feat.detach().numpy(), rois.numpy(), scale, ph, pw)
        ctx.save_for_backward(rois)
        ctx.arg = arg
        ctx.shape = feat.shape
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        B, C, H, W = ctx.shape
        gi = roi_pool_backward(g.contiguous().numpy(), ctx.arg, rois.numpy(), B, C, H, W)
        return torch.from_numpy(gi), None, None, None, None


def roi_pool_autograd(feat, rois, scale=0.125, ph=7, pw=7):
    return _RoiPoolFn.apply(feat, rois, scale, ph, pw)


def model_forward(sd: Dict[str, torch.Tensor], images: torch.Tensor, boxes: List[torch.Tensor],
                  labels_per_img, rng, *, thres=0.5, nms=0.1, lmda=0.03, temp=0.2, return_trace=False):
    """GeneralizedRCNN.forward in train mode (modeling/detector/generalized_rcnn.py:57-97 ->
    roi_heads/weak_head/weak_head.py:101-122).  `sd` is a reference-keyed state dict;
    `rng` supplies the stochastic layers (see StochasticSource below)."""
    feat = vgg16_forward(images, sd)
    rois = torch.cat([torch.cat([torch.full((b.shape[0], 1), float(i)), b], dim=1)
                      for i, b in enumerate(boxes)])                      # poolers.py:85-96
    pooled = roi_pool_autograd(feat, rois)                                 # [R,512,7,7]
    p = "roi_heads.feature_extractor.classifier."

    def neck(x):                                                           # vgg16.py:159-162
        x = x.reshape(x.shape[0], -1)
        x = rng.dropout(F.relu(F.linear(x, sd[p + "1.weight"], sd[p + "1.bias"])))
        x = rng.dropout(F.relu(F.linear(x, sd[p + "4.weight"], sd[p + "4.bias"])))
        return x

    def sim_net(x):                                                        # sim_net.py:25-26
        q = "roi_heads.model_sim.mlp."
        h = F.relu(F.linear(x, sd[q + "0.weight"], sd[q + "0.bias"]))
        return F.normalize(F.linear(h, sd[q + "2.weight"], sd[q + "2.bias"]), dim=1)

    clean = neck(pooled)
    simf = sim_net(clean)
    aug = neck(dropblock(pooled, rng.dropblock_centres(pooled.shape[0], 3), 3))   # weak_head.py:111-112
    q = "roi_heads.predictor."
    lin = lambda n: F.linear(aug, sd[q + n + ".weight"], sd[q + n + ".bias"])
    cls, det = lin("cls_score"), lin("det_score")
    refs = [lin("ref1"), lin("ref2"), lin("ref3")]
    bbs = [lin("bbox_pred1"), lin("bbox_pred2"), lin("bbox_pred3")]
    sizes = [b.shape[0] for b in boxes]
    pooled_l = pooled.split(sizes)
    img_off = [0]
    for n in sizes:
        img_off.append(img_off[-1] + n)
    keyed = hasattr(rng, "noise_rows")        # row-keyed source: draws depend on the proposal's global row id only
    emb_log = {}

    def embed_aug(b, c, I, kind):
        x = pooled_l[b][I]
        if kind == "drop":                                                 # vgg16.py:173-175
            cen = rng.dropblock_centres_rows(I + img_off[b], 1) if keyed else rng.dropblock_centres(x.shape[0], 1)
            x = dropblock(x, cen, 1)
        else:                                                              # vgg16.py:177-180
            nz = rng.noise_rows(I + img_off[b], x.shape) if keyed else rng.noise(x.shape)
            x = nz * x + x
        e = sim_net(neck(x))
        emb_log[(b, c, kind)] = e.detach()
        return e

    out = roi_reg_loss(cls, det, refs, bbs, simf, boxes, labels_per_img, embed_aug,
                       thres=thres, nms=nms, lmda=lmda, temp=temp, return_trace=return_trace)
    if return_trace:
        out[1].update(feat=feat.detach(), pooled=pooled.detach(), simf=simf.detach(), cls=cls.detach(), det=det.detach(),
                      refs=[r.detach() for r in refs], bbs=[r.detach() for r in bbs], emb=emb_log,
                      clean=clean.detach(), aug=aug.detach())
    return out


class StochasticSource:
    """Replayable stand-in for the reference's stochastic layers.

    mode 'seeded': DropBlock centres ~ Bernoulli(p/bs^2) (drop_block.py:42,69-70),
    noise ~ N(0,1) (vgg16.py:178), Dropout(0.5) (vgg16.py:125,128) drawn from one
    seeded CPU generator.  mode 'off': Dropout is identity; DropBlock/noise still
    seeded (used for golden vectors, which freeze Dropout -- SURVEY.md 7.3 item 7)."""

    def __init__(self, seed: int, dropout: bool = False, drop_prob: float = 0.3):
        self.g = torch.Generator().manual_seed(seed)
        self.use_dropout = dropout
        self.drop_prob = drop_prob

    def dropblock_centres(self, n, block):
        gamma = self.drop_prob / (block ** 2)
        return (torch.rand(n, 7, 7, generator=self.g) < gamma).float()

    def noise(self, shape):
        return torch.randn(shape, generator=self.g)

    def dropout(self, x):
        if not self.use_dropout:
            return x
        keep = (torch.rand(x.shape, generator=self.g) >= 0.5).float()
        return x * keep * 2.0


# --------------------------------------------------------------------------- #
# Synthetic workload generator -- shared by tests and bench
# --------------------------------------------------------------------------- #
def synth_boxes(n: int, W: int, H: int, gen: torch.Generator) -> torch.Tensor:
    """MCG-style integer boxes, min side 20, clipped, de-duplicated (data/datasets/voc.py:108-111)."""
    out = torch.zeros((0, 4))
    while out.shape[0] < n:
        k = 2 * n
        x1 = torch.rand(k, generator=gen) * (W - 40)
        y1 = torch.rand(k, generator=gen) * (H - 40)
        w = 20 + torch.rand(k, generator=gen) * (W - 21 - x1)
        h = 20 + torch.rand(k, generator=gen) * (H - 21 - y1)
        b = torch.stack([x1, y1, x1 + w, y1 + h], 1).round()
        b[:, 0::2].clamp_(0, W - 1); b[:, 1::2].clamp_(0, H - 1)
        ok = ((b[:, 2] - b[:, 0]) >= 20) & ((b[:, 3] - b[:, 1]) >= 20)
        out = torch.unique(torch.cat([out, b[ok]]), dim=0)
        out = out[torch.randperm(out.shape[0], generator=gen)]
    return out[:n].contiguous()


def synth_state_dict(num_classes: int = 21, seed: int = 0, conv_c=(64, 128, 256, 512, 512),
                     fc_dim: int = 4096, pool_c: int = 512) -> Dict[str, torch.Tensor]:
    """Random-init weights under the reference's state-dict keys.  Each tensor is
    drawn from its own generator keyed by the parameter name, so any module order gives the same
    values.  std: conv kaiming fan_out (vgg16.py:38-42); fc 0.01 (vgg16.py:142-146); predictor
    0.001 (roi_weak_predictors.py:133-137); Sim_Net kaiming fan_out (sim_net.py:18-22)."""
    import zlib
    sd = {}

    def draw(name, shape, std):
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
        sd[name] = torch.randn(shape, generator=g) * std

    chans = [conv_c[0]] * 2 + [conv_c[1]] * 2 + [conv_c[2]] * 3 + [conv_c[3]] * 3 + [conv_c[4]] * 3
    cin = 3
    for li, co in zip(CONV_IDX, chans):
        draw("backbone.body.features.%d.weight" % li, (co, cin, 3, 3), (2.0 / (co * 9)) ** 0.5)
        sd["backbone.body.features.%d.bias" % li] = torch.zeros(co)
        cin = co
    p = "roi_heads.feature_extractor.classifier."
    draw(p + "1.weight", (fc_dim, pool_c * 49), 0.01); sd[p + "1.bias"] = torch.zeros(fc_dim)
    draw(p + "4.weight", (fc_dim, fc_dim), 0.01); sd[p + "4.bias"] = torch.zeros(fc_dim)
    q = "roi_heads.predictor."
    for n, o in (("cls_score", num_classes), ("det_score", num_classes), ("ref1", num_classes),
                 ("ref2", num_classes), ("ref3", num_classes), ("bbox_pred1", 4 * num_classes),
                 ("bbox_pred2", 4 * num_classes), ("bbox_pred3", 4 * num_classes)):
        draw(q + n + ".weight", (o, fc_dim), 0.001); sd[q + n + ".bias"] = torch.zeros(o)
    s = "roi_heads.model_sim.mlp."
    draw(s + "0.weight", (fc_dim, fc_dim), (2.0 / fc_dim) ** 0.5); sd[s + "0.bias"] = torch.zeros(fc_dim)
    draw(s + "2.weight", (128, fc_dim), (2.0 / 128) ** 0.5); sd[s + "2.bias"] = torch.zeros(128)
    return sd


def synth_batch(B: int, N: int, W: int, H: int, num_classes: int = 21, seed: int = 1234):
    """images [B,3,Hp,Wp] (padded to /32, structures/image_list.py:58-65), boxes, labels."""
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    images = torch.zeros(B, 3, Hp, Wp)
    boxes, labels = [], []
    for i in range(B):
        g = torch.Generator().manual_seed(seed + i)
        images[i, :, :H, :W] = torch.randn(3, H, W, generator=g) * 50
        boxes.append(synth_boxes(N, W, H, g))
        labels.append(torch.randperm(num_classes - 1, generator=g)[:2].numpy() + 1)
    return images, boxes, labels
