"""MISTPredictor (roi_heads/weak_head/roi_weak_predictors.py:112-187): 8 linear heads.  State-dict
keys roi_heads.predictor.{cls_score,det_score,ref1..3,bbox_pred1..3}.  The eight GEMMs over the same
[R,4096] input are issued as ONE launch of the tcgen05 fc kernel (csrc/fc_gemm.cu) against the row-concatenated weights (views into the
per-head parameters are rebuilt each call, so autograd and the state dict are unchanged)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import registry


@registry.ROI_WEAK_PREDICTOR.register("MISTPredictor")
class MISTPredictor(nn.Module):
    def __init__(self, config, in_channels):
        super().__init__()
        nc = config.MODEL.ROI_BOX_HEAD.NUM_CLASSES
        nreg = 2 if config.MODEL.CLS_AGNOSTIC_BBOX_REG else nc
        self.cls_score = nn.Linear(in_channels, nc)
        self.det_score = nn.Linear(in_channels, nc)
        self.ref1 = nn.Linear(in_channels, nc)
        self.bbox_pred1 = nn.Linear(in_channels, nreg * 4)
        self.ref2 = nn.Linear(in_channels, nc)
        self.bbox_pred2 = nn.Linear(in_channels, nreg * 4)
        self.ref3 = nn.Linear(in_channels, nc)
        self.bbox_pred3 = nn.Linear(in_channels, nreg * 4)
        self.strict_fp32 = False       # True: 3xTF32 split products (parity tests)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0, std=0.001)
                nn.init.constant_(m.bias, 0)

    def forward(self, x, proposals, in_mask_scale=None):
        """in_mask_scale: x comes from run_classifier(..., fuse_out_bwd=True) and this is its only consumer."""
        assert x.dim() == 2
        heads = [self.cls_score, self.det_score, self.ref1, self.bbox_pred1, self.ref2, self.bbox_pred2,
                 self.ref3, self.bbox_pred3]
        W = torch.cat([h.weight for h in heads], 0)
        b = torch.cat([h.bias for h in heads], 0)
        from . import fc
        y = fc.linear(x, W, b, strict=self.strict_fp32, in_mask_scale=in_mask_scale)
        cls_logit, det_logit, ref1, bb1, ref2, bb2, ref3, bb3 = y.split([h.out_features for h in heads], dim=1)
        if self.training:
            cls_logit._odw_logits = y      # the loss consumes the eight heads as one buffer (csrc/head_loss.cu)
        if not self.training:
            cls_logit = F.softmax(cls_logit, dim=1)
            det_logit = torch.cat([F.softmax(d, dim=0) for d in det_logit.split([len(p) for p in proposals])], 0)
            ref1, ref2, ref3 = F.softmax(ref1, dim=1), F.softmax(ref2, dim=1), F.softmax(ref3, dim=1)
        return cls_logit, det_logit, [ref1, ref2, ref3], [bb1, bb2, bb3]
