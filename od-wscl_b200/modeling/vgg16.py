"""VGG16-OICR conv body and the VGG16.roi_head feature extractor (modeling/backbone/vgg16.py).
Module tree and state-dict keys are the reference's: backbone.body.features.{0..28},
roi_heads.feature_extractor.classifier.{1,4}."""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import registry
from .conv_stack import vgg_stack
from .dropblock import DropBlock2D
from .poolers import Pooler


class Identity(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x):
        return x


vgg_cfg = {
    "VGG16-OICR": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "I", "512-D", "512-D", "512-D"],
}


def make_layers(cfg_list):
    """vgg16.py:58-83: 13 convs, 3 max-pools, pool4 removed ('I'), conv5 dilation 2, last ReLU dropped."""
    layers, in_ch = [], 3
    for v in cfg_list:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        elif v == "I":
            layers.append(Identity())
        elif isinstance(v, str) and "-D" in v:
            c = int(v.split("-")[0])
            layers += [nn.Conv2d(in_ch, c, kernel_size=3, padding=2, dilation=2), nn.ReLU(inplace=True)]
            in_ch = c
        else:
            layers += [nn.Conv2d(in_ch, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            in_ch = v
    return nn.Sequential(*layers[:-1])


class VGG_Base(nn.Module):
    def __init__(self, features, cfg, init_weights=True):
        super().__init__()
        self.features = features
        if init_weights:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                    nn.init.constant_(m.bias, 0)
        freeze_at = cfg.MODEL.BACKBONE.FREEZE_CONV_BODY_AT           # vgg16.py:48-55
        if freeze_at >= 0:
            for layer in range([5, 10, 17, 23, 29][freeze_at - 1]):
                for p in self.features[layer].parameters():
                    p.requires_grad = False

        self.strict_fp32 = False     # True: 3-pass TF32 hi/lo split in every convolution (parity tests)

    def forward(self, x):
        # self.features only holds the parameters (state-dict keys backbone.body.features.{0,2,...,28});
        # the arithmetic runs in the tcgen05 conv stack (csrc/conv3x3.cu), channels-last
        convs = [m for m in self.features if isinstance(m, nn.Conv2d)]
        return [vgg_stack(x, convs, strict=self.strict_fp32)]


@registry.BACKBONES.register("VGG16-OICR")
def add_conv_body(cfg, dim_in=3):
    body = VGG_Base(make_layers(vgg_cfg[cfg.MODEL.BACKBONE.CONV_BODY]), cfg)
    model = nn.Sequential(OrderedDict([("body", body)]))
    model.out_channels = 512
    return model


@registry.ROI_BOX_FEATURE_EXTRACTORS.register("VGG16.roi_head")
class VGG16FC67ROIFeatureExtractor(nn.Module):
    def __init__(self, config, in_channels, init_weights=True):
        super().__init__()
        assert in_channels == 512
        res = config.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        self.pooler = Pooler(output_size=(res, res), scales=config.MODEL.ROI_BOX_HEAD.POOLER_SCALES,
                             sampling_ratio=config.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO)
        self.classifier = nn.Sequential(Identity(), nn.Linear(512 * 7 * 7, 4096), nn.ReLU(inplace=True),
                                        nn.Dropout(), nn.Linear(4096, 4096), nn.ReLU(inplace=True), nn.Dropout())
        self.out_channels = 4096
        if config.DB.METHOD == "dropblock":
            self.dropblock = DropBlock2D(block_size=3, drop_prob=0.3)
        self.sim_drop = DropBlock2D(block_size=1, drop_prob=0.3)
        self.noise_sampler = None      # test hook: callable(shape, device) -> N(0,1) tensor
        self.strict_fp32 = False       # True: every fc product as a 3xTF32 split (parity tests)
        self.merge_fc6_wgrad = True    # train: the weight gradients of the two fc6 / fc7 calls of a step leave as one tensor each
        self._fc6_stash = None
        self.fuse_clean_aug = True     # train: ROIPool + DropBlock into one [2R,...] batch, fc6/fc7 once
        if init_weights:
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    nn.init.normal_(m.weight, 0, 0.01)
                    nn.init.constant_(m.bias, 0)

    def run_classifier(self, x, role=None, fuse_out_bwd=False):
        """self.classifier(x) (vgg16.py:122-130): Linear + ReLU + Dropout twice, each as ONE launch of the tcgen05 fc
        kernel (csrc/fc_gemm.cu) with bias, ReLU, Philox Dropout and TF32 rounding fused into its epilogue.  Same
        parameters, same state-dict keys.  `role` ("main" / "small") marks the two fc6 calls of a training step so their
        weight gradients leave as one tensor (fc._LinearFn).  fc6's ReLU/Dropout derivative is applied by fc7's dgrad
        epilogue; with `fuse_out_bwd` the caller promises that every consumer of the result does the same for fc7
        (fc.linear(..., in_mask_scale=self.out_act_scale())), so no separate elementwise backward pass is left."""
        from . import fc
        c = self.classifier
        stash = self._fc6_stash if (role is not None and self.merge_fc6_wgrad) else None     # {"fc6": {}, "fc7": {}}
        fuse = fc.FUSE_ACT_BWD and torch.is_grad_enabled()
        p6 = float(c[3].p) if self.training else 0.0
        p7 = float(c[6].p) if self.training else 0.0
        x = fc.linear(x, c[1].weight, c[1].bias, act=fc.ACT_RELU_DROPOUT if p6 > 0.0 else fc.ACT_RELU, p=p6,
                      seed=fc.next_dropout_seed() if p6 > 0.0 else 0, round_out=True, strict=self.strict_fp32,
                      stash=stash["fc6"] if stash is not None else None, role=role if stash is not None else None,
                      act_bwd_fused=fuse)
        x = fc.linear(x, c[4].weight, c[4].bias, act=fc.ACT_RELU_DROPOUT if p7 > 0.0 else fc.ACT_RELU, p=p7,
                      seed=fc.next_dropout_seed() if p7 > 0.0 else 0, round_out=True, strict=self.strict_fp32,
                      stash=stash["fc7"] if stash is not None else None, role=role if stash is not None else None,
                      in_mask_scale=1.0 / (1.0 - p6) if fuse else None, act_bwd_fused=fuse and fuse_out_bwd)
        return x

    def out_act_scale(self):
        """1 / (1 - p) of the last Dropout: the scale a consumer passes as in_mask_scale (see run_classifier)."""
        return 1.0 / (1.0 - (float(self.classifier[6].p) if self.training else 0.0))

    def forward(self, x, proposals):                     # vgg16.py:148-153
        self._fc6_stash = None
        pooled_feat = self.pooler(x, proposals)
        x = self.run_classifier(pooled_feat.view(pooled_feat.shape[0], -1))
        return x, pooled_feat

    def can_fuse_clean_aug(self):
        from ..layers import ROIPool
        return (self.fuse_clean_aug and self.training and hasattr(self, "dropblock") and self.dropblock.drop_prob > 0
                and isinstance(self.pooler.poolers[0], ROIPool))

    def forward_clean_and_aug(self, x, proposals, fuse_out_bwd=False):
        """weak_head.py:107 + :111-112 in one pass: returns (clean_roi_feats, aug_roi_feats, clean_pooled_feats).
        The same arithmetic as forward() followed by forward_neck(forward_dropblock(pooled)); the two fc6/fc7 passes
        run as one batch, and clean_pooled_feats carries `_odw_gather(rows)` for the contrastive branch (its gradient
        joins the single ROIPool backward instead of a dense zero-filled tensor)."""
        from ..layers import gather_rows, pool_and_augment, split_rows
        rois = self.pooler.convert_to_roi_format(proposals).float()
        pool, db = self.pooler.poolers[0], self.dropblock
        R = rois.shape[0]
        ph, pw = pool.output_size if isinstance(pool.output_size, (tuple, list)) else (pool.output_size, pool.output_size)
        gamma = db.drop_prob / (db.block_size ** 2)                         # drop_block.py:69-70
        if db.centre_sampler is not None:
            centres = db.centre_sampler(R, ph, pw, gamma, rois.device)
        else:
            centres = (torch.rand(R, ph, pw, device=rois.device) < gamma).float()
        stash = {}
        buf = pool_and_augment(x[0].float(), rois, (ph, pw), pool.spatial_scale, centres.contiguous(), db.block_size, stash)
        # one stash per layer and step: the main call here, the small (augmented positives) calls later
        self._fc6_stash = {"fc6": {}, "fc7": {}} if self.merge_fc6_wgrad else None
        feats = self.run_classifier(buf.view(2 * R, -1), role="main" if self._fc6_stash is not None else None,
                                    fuse_out_bwd=fuse_out_bwd)
        clean, aug = split_rows(feats, R)
        pooled = buf.detach()[:R]
        pooled._odw_gather = lambda rows: gather_rows(buf, rows, R, stash)
        pooled._odw_aug_positives = lambda rows, seg_off, P: self._aug_positives(buf, rows, R, stash, seg_off, P)
        return clean, aug, pooled

    def _aug_positives(self, buf, rows, R, stash, seg_off, P):
        """[drop_pool(x); noise_pool(x)] of the rows `rows` of the clean pooled features (vgg16.py:173-180) in one kernel;
        the test hooks (centre_sampler / noise_sampler) are honoured."""
        from ..layers import aug_positives
        from . import fc
        db = self.sim_drop
        n = rows.numel()
        gamma = db.drop_prob / (db.block_size ** 2)
        self._aug_rows = rows
        if db.centre_sampler is not None:
            centres = db.centre_sampler(n, 7, 7, gamma, rows.device)
        else:
            centres = (torch.rand(n, 7, 7, device=rows.device) < gamma).float()
        noise = None
        if self.noise_sampler is not None:
            noise = self.noise_sampler((n,) + tuple(buf.shape[1:]), rows.device).contiguous()
        return aug_positives(buf, rows, R, stash, seg_off, P, centres.contiguous(), db.block_size, noise,
                             fc.next_dropout_seed())

    def forward_pooler(self, x, proposals):
        return self.pooler(x, proposals)

    def forward_neck(self, x, fuse_out_bwd=False):       # vgg16.py:159-162
        return self.run_classifier(x.view(x.shape[0], -1), role="small" if self._fc6_stash is not None else None,
                                   fuse_out_bwd=fuse_out_bwd)

    def forward_dropblock(self, pooled_feats, proposals):  # vgg16.py:165-167
        return self.dropblock(pooled_feats)

    def drop_pool(self, pooled_feats, n_valid=None, seg_off=None):     # vgg16.py:173-175
        """`seg_off`: the batch holds several (image, class) groups of positives; each is renormalised on its own, as
        the reference's one-call-per-group loop does (loss.py:299)."""
        if seg_off is not None:
            return self.sim_drop(pooled_feats, seg_off=seg_off)
        return self.sim_drop(pooled_feats, n_valid) if n_valid is not None else self.sim_drop(pooled_feats)

    def noise_pool(self, pooled_feats):                  # vgg16.py:177-180
        if self.noise_sampler is not None:
            noise = self.noise_sampler(pooled_feats.shape, pooled_feats.device)
        else:
            noise = torch.randn(pooled_feats.shape, device=pooled_feats.device)
        return noise * pooled_feats + pooled_feats
