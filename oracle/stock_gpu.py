"""Stock-PyTorch GPU restatement of the reference's train step -- the stand-in for "the reference's own GPU build".

TEST / BENCH INFRASTRUCTURE ONLY (bench.py --impl stock-gpu); never imported by the product package.

The reference's CUDA extension does not compile against current PyTorch (`#include <THC/THC.h>`, csrc/cuda/
ROIPool_cuda.cu:5) and /root/reference does not exist on the GPU box, so the reference itself cannot be timed on a
B200.  This module restates what its GPU build executes, op for op, with the SAME libraries it calls: cuDNN through
F.conv2d (backbone/vgg16.py:26-36), torchvision.ops.roi_pool (the CUDA kernel of the same Caffe2 lineage and launch
geometry as csrc/cuda/ROIPool_cuda.cu:16-108), cuBLAS through F.linear (vgg16.py:122-130, sim_net.py:25-26,
roi_weak_predictors.py:158-165), torchvision.ops.nms (structures/boxlist_ops.py:57), and the reference's Python loop
structure WITH its host synchronisations: DropBlock masks sampled on the CPU and copied (dropblock/drop_block.py:42-45),
the two triple loops of weak_head/loss.py:281-345 (nonzero / unique / data-dependent indexing / a full N x N torch.mm per
inner iteration), od_layer's .cpu().numpy() round trip (pseudo_label_generator.py:176-177), SupConLossV2's dense M x M
temporaries (sim_loss.py:60-80).  TF32 on (torch 1.7.1's default, the reference's pin).  Same synthetic inputs, same
state-dict keys as the oracle."""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc


def _iou_plus1(a, b):
    """structures/boxlist_ops.py:127-160 (legacy +1 convention)."""
    area = lambda x: (x[:, 2] - x[:, 0] + 1) * (x[:, 3] - x[:, 1] + 1)
    lt = torch.max(a[:, None, :2], b[:, :2])
    rb = torch.min(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (area(a)[:, None] + area(b) - inter)


def _dropblock(x, block, p=0.3):
    gamma = p / (block ** 2)
    mask = (torch.rand(x.shape[0], *x.shape[2:]) < gamma).float().to(x.device)        # CPU sample + H2D, :42-45
    bm = 1 - F.max_pool2d(mask[:, None], kernel_size=block, stride=1, padding=block // 2).squeeze(1)
    out = x * bm[:, None]
    return out * bm.numel() / bm.sum()


def _supcon(feats, labels, w, temp):
    sim = torch.div(torch.matmul(feats, feats.T), temp)
    sim = sim - sim.max(dim=1, keepdim=True)[0].detach()
    lmask = torch.ones_like(sim)
    lmask.fill_diagonal_(0)
    e = torch.exp(sim)
    mask = lmask * torch.eq(labels.view(-1, 1), labels.view(-1, 1).T).float()
    return (-torch.log((e * mask).sum(1) / (e * lmask).sum(1)) * w.detach()).mean()


def _od_layer(P, source, pos, inst_i, dev):
    prob = source[:, 1:].clone()
    gtb, gtc, gts = [], [], []
    for c in pos:
        col = prob[:, c]
        m = torch.argmax(col)
        sb = inst_i[c]
        if sb.nelement() == 0:
            gtb.append(P[m].view(1, -1)); gtc.append(torch.full((1,), c + 1, device=dev)); gts.append(col[m].view(1))
        else:
            gtb.append(P[sb]); gtc.append(torch.full((sb.numel(),), c + 1, device=dev)); gts.append(col[sb])
        prob[m].fill_(0)
    gtb, gtc, gts = torch.cat(gtb), torch.cat(gtc), torch.cat(gts)
    ov = _iou_plus1(P, gtb)
    ov_h = ov.cpu().numpy()                                                   # :176-177 host round trip
    mx = torch.tensor(ov_h.max(axis=1), device=dev)
    am = torch.tensor(ov_h.argmax(axis=1), device=dev)
    labels = gtc[am].clone()
    weights = gts[am]
    labels[mx.le(0.5).nonzero(as_tuple=False)[:, 0]] = 0
    g, p = gtb[am], P
    ew = p[:, 2] - p[:, 0] + 1; eh = p[:, 3] - p[:, 1] + 1
    ex = p[:, 0] + 0.5 * ew; ey = p[:, 1] + 0.5 * eh
    gw = g[:, 2] - g[:, 0] + 1; gh = g[:, 3] - g[:, 1] + 1
    gx = g[:, 0] + 0.5 * gw; gy = g[:, 1] + 0.5 * gh
    tg = torch.stack((10 * (gx - ex) / ew, 10 * (gy - ey) / eh, 5 * torch.log(gw / ew), 5 * torch.log(gh / eh)), 1)
    return labels, weights, tg


def train_step(sd, images, boxes, labels_per_img, *, thres=0.5, nms=0.1, lmda=0.03, temp=0.2, eps=1e-8):
    """One forward + loss of the reference on whatever device `sd` / `images` live on; returns the loss dict."""
    import torchvision
    dev = images.device
    feat = orc.vgg16_forward(images, sd)                                       # cuDNN
    rois = torch.cat([torch.cat([torch.full((b.shape[0], 1), float(i), device=dev), b], 1) for i, b in enumerate(boxes)])
    pooled = torchvision.ops.roi_pool(feat, rois, (7, 7), 0.125)
    p, q, pr = "roi_heads.feature_extractor.classifier.", "roi_heads.model_sim.mlp.", "roi_heads.predictor."

    def neck(x):
        x = x.reshape(x.shape[0], -1)
        x = F.dropout(F.relu(F.linear(x, sd[p + "1.weight"], sd[p + "1.bias"])), 0.5, True)
        return F.dropout(F.relu(F.linear(x, sd[p + "4.weight"], sd[p + "4.bias"])), 0.5, True)

    def sim_net(x):
        h = F.relu(F.linear(x, sd[q + "0.weight"], sd[q + "0.bias"]))
        return F.normalize(F.linear(h, sd[q + "2.weight"], sd[q + "2.bias"]), dim=1)
    clean = neck(pooled)
    simf = sim_net(clean)
    aug = neck(_dropblock(pooled, 3))
    lin = lambda n: F.linear(aug, sd[pr + n + ".weight"], sd[pr + n + ".bias"])
    cls_l, det_l = lin("cls_score"), lin("det_score")
    refs = [lin("ref1"), lin("ref2"), lin("ref3")]
    bbs = [lin("bbox_pred1"), lin("bbox_pred2"), lin("bbox_pred3")]
    sizes = [b.shape[0] for b in boxes]
    C = cls_l.shape[1]
    final = (F.softmax(cls_l, 1) * torch.cat([F.softmax(d, 0) for d in det_l.split(sizes)])).split(sizes)
    ref_l = [r.split(sizes) for r in refs]
    box_l = [r.split(sizes) for r in bbs]
    simf_l, pooled_l = simf.split(sizes), pooled.split(sizes)
    B, nc = len(boxes), C - 1
    img_labels, pos = [], []
    for lab in labels_per_img:
        v = torch.zeros(C, device=dev); v[torch.as_tensor(np.unique(lab), device=dev).long()] = 1; v[0] = 0
        img_labels.append(v)
        pos.append([int(x) for x in v[1:].eq(1).nonzero(as_tuple=False)[:, 0]])          # host sync, loss.py:270
    src = lambda b, i: final[b] if i == 0 else F.softmax(ref_l[i - 1][b], dim=1)
    idx = [[torch.zeros(0, dtype=torch.long, device=dev) for _ in range(nc)] for _ in range(B)]
    bank = [[] for _ in range(nc)]
    Wt = []
    for b in range(B):                                                          # Phase A, loss.py:281-307
        P = boxes[b]
        for i in range(3):
            ps = src(b, i)[:, 1:].detach()
            for c in pos[b]:
                m = torch.argmax(ps[:, c])
                ov = _iou_plus1(P, P[m].view(1, 4))
                nb = torch.nonzero(torch.ge(ov, thres).max(dim=1)[0]).view(-1)
                _ = _iou_plus1(P, P[m].view(1, 4))[nb]                          # computed twice, utils.py:24-25
                idx[b][c] = torch.unique(torch.cat([idx[b][c], nb]))
        for c in pos[b]:
            I = idx[b][c]
            S = final[b]
            h = S[I, c + 1] / S[:, c + 1].sum()
            bank[c].append(simf_l[b][I]); Wt.append(h)
            x = pooled_l[b][I]
            bank[c].append(sim_net(neck(_dropblock(x, 1)))); Wt.append(h)
            bank[c].append(sim_net(neck(torch.randn(x.shape, device=dev) * x + x))); Wt.append(h)
    coll = [torch.cat(r).detach().clone() if r else None for r in bank]
    inst = [[[torch.zeros(0, dtype=torch.long, device=dev) for _ in range(nc)] for _ in range(3)] for _ in range(B)]
    for b in range(B):                                                          # Phase B, loss.py:311-345
        P, Fb = boxes[b], simf_l[b]
        for i in range(3):
            ps = src(b, i)[:, 1:].detach()
            for c in pos[b]:
                m = torch.argmax(ps[:, c])
                sim_mat = torch.mm(Fb.detach(), Fb.detach().T)                  # the full N x N per inner iteration, :319
                tau = torch.mm(Fb.detach()[m].view(1, -1), coll[c].T).mean()
                close = torch.ge(sim_mat[m], tau)
                if len(pos[b]) > 1:
                    for n_c in pos[b]:
                        if n_c != c:
                            close = torch.ge(close, sim_mat[torch.argmax(ps[:, n_c])])
                close_i = close.nonzero(as_tuple=False).view(-1)
                keep = torchvision.ops.nms(P[close_i], ps[close_i, c], nms)
                kept = close_i[keep]
                if kept.nelement() == 0:
                    kept = m.view(1)
                inst[b][i][c] = torch.cat([inst[b][i][c], kept])
                comb, cnt = torch.cat([kept, idx[b][c]]).unique(return_counts=True)     # loss.py:336-337
                inter = comb[cnt > 1]
                comb2, cnt2 = torch.cat([kept, inter]).unique(return_counts=True)
                new = comb2[cnt2 == 1]
                if new.nelement() == 0:
                    new = m.view(1)
                bank[c].append(Fb[new])
                idx[b][c] = torch.unique(torch.cat([idx[b][c], new]))
                S = final[b]
                Wt.append((S[new, c + 1] / S[:, c + 1].sum()).view(-1))
    feats, flab = [], []
    for c, rows in enumerate(bank):
        if rows:
            f = torch.cat(rows)
            if f.shape[0]:
                feats.append(f); flab.append(torch.full((f.shape[0],), float(c), device=dev))
    losses = {"loss_sim": lmda * _supcon(torch.cat(feats), torch.cat(flab), torch.cat([x.view(-1) for x in Wt]).detach(), temp)}
    losses["loss_img"] = 0.0
    for i in range(3):
        losses["loss_ref_cls%d" % i] = 0.0
        losses["loss_ref_reg%d" % i] = 0.0
    for b in range(B):
        img_score = torch.clamp(final[b].sum(0), min=eps, max=1 - eps)
        losses["loss_img"] = losses["loss_img"] + F.binary_cross_entropy(img_score, img_labels[b].clamp(0, 1))
        for i in range(3):
            pl, lw, rt = _od_layer(boxes[b], src(b, i).detach(), pos[b], inst[b][i], dev)
            lm = 3 if i == 0 else 1
            losses["loss_ref_cls%d" % i] = losses["loss_ref_cls%d" % i] + lm * torch.mean(
                F.cross_entropy(ref_l[i][b], pl, reduction="none") * lw)
            pi = torch.nonzero(pl > 0, as_tuple=False).squeeze(1)
            mp = 4 * pl[pi][:, None] + torch.tensor([0, 1, 2, 3], device=dev)
            d = torch.abs(box_l[i][b][pi[:, None], mp] - rt[pi])
            reg = lm * torch.sum(torch.where(d < 1, 0.5 * d ** 2, d - 0.5) * lw[pi, None])
            losses["loss_ref_reg%d" % i] = losses["loss_ref_reg%d" % i] + reg / pl.numel()
    for k in losses:
        if "sim" not in k:
            losses[k] = losses[k] / B
    return losses
