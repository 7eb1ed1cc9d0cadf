"""GeneralizedRCNN (modeling/detector/generalized_rcnn.py:23-97): backbone -> pre-computed rois ->
roi_heads.  forward(images, targets, rois, model_cdb=None, iteration=None) -> (losses, accuracy) in
train mode.  Attribute names `backbone`, `roi_heads` and every state-dict key match the reference."""
from torch import nn

from ..structures import to_image_list
from . import registry
from .weak_head import build_roi_weak_head
from . import vgg16  # noqa: F401


class GeneralizedRCNN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.backbone = registry.BACKBONES[cfg.MODEL.BACKBONE.CONV_BODY](cfg)
        self.roi_heads = build_roi_weak_head(cfg, self.backbone.out_channels)

    def forward(self, images, targets=None, rois=None, model_cdb=None, iteration=None):
        if self.training and targets is None:
            raise ValueError("In training mode, targets should be passed")
        images = to_image_list(images)
        features = self.backbone(images.tensors)                                       # :73
        proposals = rois                                                               # :75-78, no RPN
        x, result, detector_losses, accuracy = self.roi_heads(features, proposals, targets, model_cdb, iteration)
        if self.training:
            losses = {}
            losses.update(detector_losses)
            return losses, accuracy
        return result


def build_detection_model(cfg):
    return GeneralizedRCNN(cfg)
