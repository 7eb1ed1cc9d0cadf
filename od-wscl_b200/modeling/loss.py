"""RoIRegLossComputation (roi_heads/weak_head/loss.py:172-411), same call signature and loss /
accuracy keys, re-designed for the device:

  reference                                           here
  ------------------------------------------------    -------------------------------------------
  2 triple-nested Python loops, ~20-40 tiny kernels   discover phase A / phase B kernels, one CTA
  and >15 host syncs per (image, branch, class)       per (image, class); ONE host sync per step
  N x N torch.mm per inner iteration (loss.py:319)    only the similarity rows the rule reads
  torchvision NMS with a host sweep                   sort + sweep in shared memory
  SupConLossV2: GEMM + ~12 elementwise M x M passes   fused tile kernels, fwd and bwd
  od_layer via .cpu().numpy() (plg.py:176-177)        one kernel per (image, branch)
  2 small fc6/fc7/Sim_Net passes per (image, class)   one batched pass over all augmented positives

The single host synchronisation reads K = number of Phase-A positives (needed to size the
augmented-positives GEMM batch)."""
import torch
import torch.nn.functional as F

from .. import capi
from ..config import cfg as global_cfg
from . import registry
from . import sim_head
from .sim_head import SupConLossV2, supcon_bank_loss


def _gather(pooled, rows):
    """pooled[rows]; a pooled tensor produced by forward_clean_and_aug routes the gradient through its own gather."""
    g = getattr(pooled, "_odw_gather", None)
    return g(rows) if g is not None else pooled.index_select(0, rows)


class _PinnedStaging:
    """A small ring of page-locked host buffers for the per-step bookkeeping tensors (pair lists, offsets, multi-hot
    labels).  `tensor.pin_memory()` every step goes through cudaHostAlloc whenever the caching host allocator has no free
    block -- a driver call that can stall the enqueueing thread for tens of milliseconds; the ring allocates four slots
    once and guards each slot with an event recorded after its asynchronous copy."""

    def __init__(self, slots=4):
        self.slots = [None] * slots
        self.i = 0

    def to_device(self, host, dev):
        """host: 1-D CPU tensor -> device copy (asynchronous, from pinned memory)."""
        n = host.numel()
        k = self.i
        self.i = (self.i + 1) % len(self.slots)
        slot = self.slots[k]
        if slot is None or slot[0].numel() < n or slot[0].dtype != host.dtype:
            slot = [torch.empty((max(n, 256),), dtype=host.dtype).pin_memory(), None]
            self.slots[k] = slot
        if slot[1] is not None:
            slot[1].synchronize()                      # the copy that last read this slot (4 steps ago) is long done
        slot[0][:n].copy_(host)
        out = slot[0][:n].to(dev, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        return out


def _host_labels(target):
    lab = target.get_field("labels_host") if target.has_field("labels_host") else target.get_field("labels")
    if isinstance(lab, torch.Tensor):
        lab = lab.cpu()
    return sorted(set(int(x) for x in lab))


class _HeadLossFn(torch.autograd.Function):
    """(loss_img, loss_ref_cls0, loss_ref_reg0, ..., loss_ref_reg2, acc_img, acc_ref0..2) of loss.py:349-406 from the
    heads' logits buffer (csrc/head_loss.cu).  The forward already writes d(sum of losses)/d(logits) in closed form; the
    backward only scales the column block of every loss by its upstream gradient."""

    @staticmethod
    def forward(ctx, logits, hs, img_labels, pl, lw, rt, cls_agnostic, eps):
        out, grad = capi.head_loss(hs, img_labels, pl, lw, rt, cls_agnostic, eps)
        ctx.save_for_backward(grad)
        ctx.C, ctx.Q, ctx.width = hs.C, hs.Q, logits.shape[1]
        res = out.unbind(0)                 # 11 scalars: no select / scatter nodes in the autograd graph
        ctx.mark_non_differentiable(*res[7:])
        return res

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        (grad,) = ctx.saved_tensors
        zero = None
        up = []
        for g in gs[:7]:
            if g is None:
                zero = grad.new_zeros(()) if zero is None else zero
                g = zero
            up.append(g.reshape(()).float())
        capi.head_grad_scale_(grad, ctx.C, ctx.Q, torch.stack(up))
        return (grad[:, :ctx.width],) + (None,) * 7


@registry.ROI_WEAK_LOSS.register("RoIRegLoss")
class RoIRegLossComputation(object):
    def __init__(self, cfg):
        self.refine_p = cfg.MODEL.ROI_WEAK_HEAD.OICR_P
        self.contra = cfg.SOLVER.CONTRA
        if not (self.contra and self.refine_p == 0):
            raise NotImplementedError("only the shipped configuration (CONTRA: True, OICR_P: 0.0 -> od_layer) "
                                      "is on the hot path (SURVEY 2.1 #7)")
        self.cls_agnostic_bbox_reg = cfg.MODEL.CLS_AGNOSTIC_BBOX_REG
        self.nms = cfg.nms
        self.sim_lmda = cfg.lmda
        self.p_thres = cfg.thres
        self.p_iou = cfg.iou            # read but unused, as in the reference (loss.py:197-198)
        self.temp = cfg.temp
        if cfg.loss != "supconv2":
            raise NotImplementedError("only cfg.loss == 'supconv2' works in the reference (SURVEY App. D)")
        self.sim_loss = SupConLossV2(self.temp)
        self.fg_thresh = cfg.MODEL.ROI_HEADS.FG_IOU_THRESHOLD
        self.batch_aug = True           # False: per-(image,class) drop/noise calls in reference order (RNG replay)
        self.last_state = None          # DiscoveryState of the last call (tests / diagnostics)
        # Speculative sizing of the augmented-positives batch: K (the number of Phase-A positives) is the ONLY
        # quantity the step reads back from the device.  With speculative_k the batch is sized from the largest K
        # seen so far (padded, masked on the device) and no host synchronisation happens inside the step;
        # `overflow` [1] fp32 is 1.0 when the true K exceeded the bound -- the caller must then discard the step
        # (bench.py hands it to the fused optimizer as `found_inf`, which skips the update, and redoes the step).
        self.speculative_k = False
        self.k_margin = 2.0
        self.k_granule = 256
        self.overflow = None
        self._k_cap = None
        self._m_cap, self._m_host, self._m_event = None, None, None
        self._k_host = None
        self._k_event = None
        self._stage_i32 = _PinnedStaging()
        self._stage_f32 = _PinnedStaging()

    def __call__(self, class_score, det_score, ref_scores, ref_bbox_preds, sim_feature, clean_pooled_feats,
                 feature_extractor, model_sim, proposals, targets, epsilon=1e-8):
        sizes = [len(p) for p in proposals]
        B, R = len(sizes), sum(sizes)
        # the eight heads' logits as ONE [R, 5C+3Q] buffer (what MISTPredictor's single GEMM leaves); pieces handed in
        # separately (tests, foreign predictors) are concatenated once
        C = class_score[0].shape[1]
        Q = ref_bbox_preds[0].shape[1]
        logits = getattr(class_score[0], "_odw_logits", None) if len(class_score) == 1 else None
        if logits is None or logits.shape[1] != 5 * C + 3 * Q:
            logits = torch.cat([torch.cat(class_score, dim=0), torch.cat(det_score, dim=0), ref_scores[0],
                                ref_bbox_preds[0], ref_scores[1], ref_bbox_preds[1], ref_scores[2], ref_bbox_preds[2]], dim=1)
        dev = logits.device
        # ---- host-side bookkeeping from the (host) image labels: pairs, offsets, multi-hot labels
        pos = [[c - 1 for c in _host_labels(t) if c > 0] for t in targets]               # loss.py:270
        pair_img = [b for b in range(B) for _ in pos[b]]
        pair_cls = [c for b in range(B) for c in pos[b]]
        P = len(pair_img)
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        meta = self._stage_i32.to_device(torch.tensor(pair_img + pair_cls + offs, dtype=torch.int32), dev)
        pair_img_d, pair_cls_d, img_off_d = meta[:P], meta[P:2 * P], meta[2 * P:]
        img_labels = torch.zeros((B, C), dtype=torch.float32)
        for b in range(B):
            for c in pos[b]:
                img_labels[b, c + 1] = 1.0
        img_labels_d = self._stage_f32.to_device(img_labels.view(-1), dev).view(B, C)
        boxes = torch.cat([p.bbox for p in proposals], dim=0).float().contiguous()
        Ncap = max(sizes)
        # ---- loss.py:234-259: class softmax x per-image proposal softmax, supervisors of the refinement branches
        hs = capi.head_scores(logits.detach(), C, Q, img_off_d, B)

        # ---- contrastive object discovery (loss.py:271-347)
        scores = (hs.final_score, hs.sm1, hs.sm2)
        Fm = sim_feature.contiguous()
        st = capi.discover_phase_a(boxes, img_off_d, scores, pair_img_d, pair_cls_d, Ncap, self.p_thres)
        spec = self.speculative_k and self.batch_aug and P > 0 and self._poll_k_cap() is not None
        if spec:
            E, K = self._augmented_positives_speculative(st, P, Ncap, clean_pooled_feats, feature_extractor, model_sim)
        else:
            E, K = self._augmented_positives_synced(st, P, clean_pooled_feats, feature_extractor, model_sim)
        capi.discover_phase_b(st, Fm.detach(), E.detach(), self.nms)
        Mcap = 3 * K + 3 * sum(sizes[b] for b in pair_img)      # what the rule can produce at most: usually >> M
        # Large banks (8 images per rank) go to the tensor-core SupCon, whose cost is Mcap^2: it gets a bound from the row
        # counts read back so far (non-blocking, like K); a bank that outgrows it raises `overflow` and the step is redone.
        m_cap = self._poll_m_cap() if spec else None
        use_tc = m_cap is not None and min(m_cap, Mcap) >= sim_head.SUPCON_TC_MIN_ROWS
        if use_tc:
            Mcap = min(m_cap, Mcap)
        capi.bank_assemble(st, C - 1, Mcap)
        if spec:
            self._record_m(st.M[1:2])
            if use_tc:
                self.overflow = torch.maximum(self.overflow, (st.M[1:2] > Mcap).to(self.overflow.dtype))
        st.E = E.detach()                       # [2K,128] augmented-positive embeddings (drop rows, then noise rows)
        self.last_state = st
        loss_sim = self.sim_lmda * supcon_bank_loss(Fm, E, st.row_src, st.row_lab, st.row_w, st.M, Mcap,
                                                    self.temp, tc=use_tc)              # loss.py:347
        # ---- pseudo labels (loss.py:364-368 -> od_layer) and the MIL + refinement losses (loss.py:349-406)
        pl, lw, rt = capi.od_layer(st, self.fg_thresh)
        out = _HeadLossFn.apply(logits, hs, img_labels_d, pl, lw, rt, bool(self.cls_agnostic_bbox_reg), float(epsilon))
        losses = {"loss_img": out[0]}
        for i in range(3):
            losses["loss_ref_cls%d" % i] = out[1 + 2 * i]
            losses["loss_ref_reg%d" % i] = out[2 + 2 * i]
        losses["loss_sim"] = loss_sim
        accs = {"acc_img": out[7], "acc_ref0": out[8], "acc_ref1": out[9], "acc_ref2": out[10]}
        return losses, accs

    # ------------------------------------------------------------------ augmented positives (loss.py:299-310)
    def _poll_k_cap(self):
        """Bound for the next speculative step from the K values read back so far (non-blocking)."""
        if self._k_event is not None and self._k_event.query():
            k = int(self._k_host[0])
            cap = self._cap_for(k)
            self._k_cap = cap if self._k_cap is None else max(self._k_cap, cap)
            self._k_event = None
        return self._k_cap

    def _poll_m_cap(self):
        """Bound on the SupCon bank rows from the counts read back so far (margin 1.25, grid of 256)."""
        if self._m_event is not None and self._m_event.query():
            cap = (int(int(self._m_host[0]) * 1.25) + 255) // 256 * 256
            self._m_cap = cap if self._m_cap is None else max(self._m_cap, cap)
            self._m_event = None
        return self._m_cap

    def _record_m(self, mdev):
        if self._m_host is None:
            self._m_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        if self._m_event is None:                      # one readback in flight at a time
            self._m_host.copy_(mdev, non_blocking=True)
            self._m_event = torch.cuda.Event()
            self._m_event.record()

    def _cap_for(self, k):
        """Bound for a batch of k positives: margin, then a coarse grid so the padded shapes (GEMM heuristics, allocator
        blocks) change rarely."""
        g = self.k_granule
        return (int(k * self.k_margin) + 64 + g - 1) // g * g

    def _record_k(self, kdev):
        if self._k_host is None:
            self._k_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        if self._k_event is None:                      # one readback in flight at a time
            self._k_host.copy_(kdev, non_blocking=True)
            self._k_event = torch.cuda.Event()
            self._k_event.record()

    @staticmethod
    def _batched_views(clean_pooled_feats, rows, seg_off, P, feature_extractor):
        """[DropBlock views ; noise views] of the positives `rows`, every (image, class) segment renormalised on its own
        (loss.py:299): one fused kernel when the pooled tensor comes from forward_clean_and_aug, else op by op."""
        fused = getattr(clean_pooled_feats, "_odw_aug_positives", None)
        if fused is not None and rows.numel() > 0:
            return fused(rows, seg_off, P)
        X = _gather(clean_pooled_feats, rows)
        feature_extractor._aug_rows = rows                       # test hook: row-keyed replay of the stochastic layers
        return torch.cat([feature_extractor.drop_pool(X, seg_off=seg_off), feature_extractor.noise_pool(X)], dim=0)

    @staticmethod
    def _embed(aug, feature_extractor, model_sim):
        """Sim_Net(neck(aug)) (loss.py:300-301,304-305).  Sim_Net is the only consumer of this neck call, so fc7's
        ReLU/Dropout derivative rides in Sim_Net's dgrad epilogue (fc.linear in_mask_scale) when the extractor supports it."""
        scale = getattr(feature_extractor, "out_act_scale", None)
        from . import fc
        if scale is not None:                                   # the product's extractor + Sim_Net
            role = "small" if getattr(model_sim, "_stash", None) is not None else None
            if fc.FUSE_ACT_BWD and torch.is_grad_enabled():
                return model_sim(feature_extractor.forward_neck(aug, fuse_out_bwd=True), in_mask_scale=scale(), role=role)
            return model_sim(feature_extractor.forward_neck(aug), role=role)
        return model_sim(feature_extractor.forward_neck(aug))

    def _augmented_positives_synced(self, st, P, clean_pooled_feats, feature_extractor, model_sim):
        offA_h = st.offA.cpu()                                   # the one host sync of the step
        K = int(offA_h[P]) if P > 0 else 0
        if self.speculative_k:
            cap = self._cap_for(K)
            self._k_cap = cap if self._k_cap is None else max(self._k_cap, cap)
            self.overflow = torch.zeros((1,), dtype=torch.float32, device=st.offA.device)
        rows = st.rowsA[:K].long()
        if self.batch_aug:
            aug = self._batched_views(clean_pooled_feats, rows, st.offA, P, feature_extractor)
        else:
            drops, noises = [], []
            for p in range(P):
                rp = rows[int(offA_h[p]):int(offA_h[p + 1])]
                Xp = _gather(clean_pooled_feats, rp)
                feature_extractor._aug_rows = rp
                drops.append(feature_extractor.drop_pool(Xp))
                noises.append(feature_extractor.noise_pool(Xp))
            aug = torch.cat(drops + noises, dim=0)
        E = self._embed(aug, feature_extractor, model_sim).contiguous()                  # [2K,128]
        return E, K

    def _augmented_positives_speculative(self, st, P, Ncap, clean_pooled_feats, feature_extractor, model_sim):
        """Same arithmetic over a batch padded to the bound Kc: rows past the device-resident K are masked (DropBlock
        renormalises each (image, class) segment of the first K rows on its own), and the [2K,128] layout the discovery kernels address (drop rows,
        then noise rows, stride K) is rebuilt with a gather -- nothing is read back."""
        Kc = min(self._k_cap, P * Ncap)
        kdev = st.offA[P:P + 1]                                   # int32 [1], device
        # K <= P*Ncap always: every address the discovery kernels form in the [2K] layout is covered by `sel`
        rows, sel, self.overflow = capi.spec_index(kdev, st.rowsA, Kc, 2 * P * Ncap)
        aug = self._batched_views(clean_pooled_feats, rows, st.offA, P, feature_extractor)
        Epad = self._embed(aug, feature_extractor, model_sim)     # [2Kc,128]
        E = Epad.index_select(0, sel).contiguous()
        self._record_k(kdev)
        return E, Kc


def make_roi_weak_loss_evaluator(cfg):
    return registry.ROI_WEAK_LOSS[cfg.MODEL.ROI_WEAK_HEAD.LOSS](cfg)
