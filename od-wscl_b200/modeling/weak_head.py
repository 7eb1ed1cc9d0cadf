"""ROIWeakRegHead (roi_heads/weak_head/weak_head.py:72-157): feature extract -> Sim_Net ->
DropBlock aug -> MISTPredictor -> RoIRegLoss.  Attribute names (feature_extractor, predictor,
model_sim, loss_evaluator) and the return tuple are the reference's."""
import torch
from torch import nn

from . import registry
from .loss import make_roi_weak_loss_evaluator
from .sim_head import Sim_Net
from . import predictors, vgg16  # noqa: F401  (registers the builders)


class ROIWeakRegHead(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        self.feature_extractor = registry.ROI_BOX_FEATURE_EXTRACTORS[cfg.MODEL.ROI_BOX_HEAD.FEATURE_EXTRACTOR](cfg, in_channels)
        self.predictor = registry.ROI_WEAK_PREDICTOR[cfg.MODEL.ROI_WEAK_HEAD.PREDICTOR](cfg, self.feature_extractor.out_channels)
        self.loss_evaluator = make_roi_weak_loss_evaluator(cfg)
        self.HEUR = cfg.MODEL.ROI_WEAK_HEAD.REGRESS_HEUR
        self.DB_METHOD = cfg.DB.METHOD
        self.model_sim = Sim_Net(cfg, self.feature_extractor.out_channels)
        from .postprocess import PostProcessor
        self.strong_post_processor = PostProcessor(                     # box_head/inference.py:260-283
            score_thresh=cfg.MODEL.ROI_HEADS.SCORE_THRESH, nms=cfg.MODEL.ROI_HEADS.NMS,
            detections_per_img=cfg.MODEL.ROI_HEADS.DETECTIONS_PER_IMG, weights=cfg.MODEL.ROI_HEADS.BBOX_REG_WEIGHTS,
            cls_agnostic_bbox_reg=cfg.MODEL.CLS_AGNOSTIC_BBOX_REG, bbox_aug_enabled=cfg.TEST.BBOX_AUG.ENABLED)

    def go_through_cdb(self, features, proposals, model_cdb):              # weak_head.py:87-99
        if not self.training or self.DB_METHOD == "none":
            return features
        if self.DB_METHOD == "dropblock":
            return self.feature_extractor.forward_dropblock(features, proposals)
        raise ValueError("DB.METHOD %r is outside the hot path (every shipped config uses 'dropblock')" % self.DB_METHOD)

    def forward(self, features, proposals, targets=None, model_cdb=None, iteration=None):
        fe = self.feature_extractor
        self.model_sim._stash = None             # per-step state of the two-call weight-gradient fold (fc._LinearFn)
        if self.training and self.DB_METHOD == "dropblock" and getattr(fe, "can_fuse_clean_aug", lambda: False)():
            # :107 and :111-112 as ONE fc6/fc7 batch over [clean; DropBlock-augmented] pooled features
            # each half of the fc7 output has exactly one consumer, whose dgrad epilogue applies fc7's ReLU/Dropout mask
            from . import fc
            fuse = fc.FUSE_ACT_BWD
            clean_roi_feats, aug_roi_feats, clean_pooled_feats = fe.forward_clean_and_aug(features, proposals,
                                                                                          fuse_out_bwd=fuse)
            s7 = fe.out_act_scale() if fuse else None
            sim_feature = self.model_sim(clean_roi_feats, in_mask_scale=s7, role="main")              # :110
            cls_score, det_score, ref_scores, ref_bbox_preds = self.predictor(aug_roi_feats, proposals,
                                                                              in_mask_scale=s7)       # :113
            loss_img, accuracy_img = self.loss_evaluator([cls_score], [det_score], ref_scores, ref_bbox_preds,
                                                         sim_feature, clean_pooled_feats, fe, self.model_sim,
                                                         proposals, targets)                          # :120
            return aug_roi_feats, proposals, loss_img, accuracy_img
        clean_roi_feats, clean_pooled_feats = self.feature_extractor.forward(features, proposals)     # :107
        if not self.training:
            cls_score, det_score, ref_scores, ref_bbox_preds = self.predictor(clean_roi_feats, proposals)
            # testing_forward, HEUR "AVG" (weak_head.py:124-135): mean of the refinement heads, decode, per-class NMS
            if self.HEUR != "AVG":
                raise NotImplementedError("REGRESS_HEUR %r: every shipped config tests with 'AVG'" % self.HEUR)
            final_score = torch.mean(torch.stack(ref_scores), dim=0)
            final_regression = torch.mean(torch.stack(ref_bbox_preds), dim=0)
            result = self.strong_post_processor((final_score, final_regression), proposals, softmax_on=False)
            return clean_roi_feats, result, {}, {}
        sim_feature = self.model_sim(clean_roi_feats)                                                 # :110
        aug_pooled_feats = self.go_through_cdb(clean_pooled_feats, proposals, model_cdb)              # :111
        aug_roi_feats = self.feature_extractor.forward_neck(aug_pooled_feats)                         # :112
        cls_score, det_score, ref_scores, ref_bbox_preds = self.predictor(aug_roi_feats, proposals)   # :113
        loss_img, accuracy_img = self.loss_evaluator([cls_score], [det_score], ref_scores, ref_bbox_preds,
                                                     sim_feature, clean_pooled_feats, self.feature_extractor,
                                                     self.model_sim, proposals, targets)              # :120
        return aug_roi_feats, proposals, loss_img, accuracy_img


def build_roi_weak_head(cfg, in_channels):
    if not cfg.MODEL.ROI_WEAK_HEAD.REGRESS_ON:
        raise NotImplementedError("ROIWeakHead (no regression) is not selected by any shipped config")
    return ROIWeakRegHead(cfg, in_channels)
