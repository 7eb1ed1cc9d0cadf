"""Config for the hot path: the subset of ``wetectron.config.defaults`` the path reads, with the
values of configs/voc/voc07_contra_db_b8_lr0.01_mcg.yaml merged in (the reference's modules read a
global ``cfg``; so do ours: poolers.py:7,66-75, pseudo_label_generator.py:7,18, loss.py:16).
Lower-case contrastive knobs follow config/defaults.py:540-551 and the README example
(`nms 0.1 lmda 0.03 iou 0.5 temp 0.2`)."""
import copy


class CfgNode(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            old = node.get(parts[-1])
            if old is not None and not isinstance(v, type(old)) and not isinstance(old, (tuple, list)):
                v = type(old)(v)
            node[parts[-1]] = v

    def freeze(self):
        pass


_DEFAULTS = {
    "MODEL": {
        "DEVICE": "cuda",
        "CLS_AGNOSTIC_BBOX_REG": False,
        "BACKBONE": {"CONV_BODY": "VGG16-OICR", "FREEZE_CONV_BODY_AT": 2},
        "ROI_HEADS": {"FG_IOU_THRESHOLD": 0.5, "BBOX_REG_WEIGHTS": (10.0, 10.0, 5.0, 5.0),
                      # test time (configs/voc/voc07_contra_db_b8_lr0.01_mcg.yaml:8-10, config/defaults.py:233)
                      "SCORE_THRESH": 0.0, "NMS": 0.4, "DETECTIONS_PER_IMG": 100},
        "ROI_BOX_HEAD": {"FEATURE_EXTRACTOR": "VGG16.roi_head", "POOLER_METHOD": "ROIPool",
                         "POOLER_RESOLUTION": 7, "POOLER_SCALES": (0.125,), "POOLER_SAMPLING_RATIO": 0,
                         "NUM_CLASSES": 21},
        "ROI_WEAK_HEAD": {"PREDICTOR": "MISTPredictor", "LOSS": "RoIRegLoss", "OICR_P": 0.0,
                          "REGRESS_ON": True, "PARTIAL_LABELS": "none", "ROI_LOSS_REFINE": False,
                          "REGRESS_HEUR": "AVG"},
    },
    "SOLVER": {"CONTRA": True, "MAX_ITER": 30000, "BASE_LR": 0.01, "MOMENTUM": 0.9, "WEIGHT_DECAY": 0.0001,
               "BIAS_LR_FACTOR": 2, "WEIGHT_DECAY_BIAS": 0},
    "DB": {"METHOD": "dropblock"},
    # the multi-scale TTA loop (engine/bbox_aug.py) is outside this package: a single-scale eval filters directly
    "TEST": {"BBOX_AUG": {"ENABLED": False}},
    "DATALOADER": {"SIZE_DIVISIBILITY": 32},
    "nms": 0.1, "lmda": 0.03, "iou": 0.5, "temp": 0.2, "thres": 0.5, "loss": "supconv2", "pos_update": 0.0,
    "OUTPUT_DIR": ".",
}

cfg = CfgNode(_DEFAULTS)


def get_cfg_defaults():
    return CfgNode(_DEFAULTS)
