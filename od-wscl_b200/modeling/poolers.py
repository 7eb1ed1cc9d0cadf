"""Mirror of modeling/poolers.py:46-128 (single-level fast path :108-109).  Reads the global cfg for
the pooling method exactly as the reference does (:66-75)."""
import torch
from torch import nn

from ..config import cfg
from ..layers import ROIAlign, ROIPool


class Pooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio):
        super().__init__()
        poolers = []
        for scale in scales:
            if cfg.MODEL.ROI_BOX_HEAD.POOLER_METHOD == "ROIPool":
                poolers.append(ROIPool(output_size, spatial_scale=scale))
            elif cfg.MODEL.ROI_BOX_HEAD.POOLER_METHOD == "ROIAlign":
                poolers.append(ROIAlign(output_size, spatial_scale=scale, sampling_ratio=sampling_ratio))
            else:
                raise ValueError("please use valid pooler function")
        if len(poolers) != 1:
            raise NotImplementedError("multi-level (FPN) pooling is outside the VGG16 hot path")
        self.poolers = nn.ModuleList(poolers)
        self.output_size = output_size

    def convert_to_roi_format(self, boxes):
        """poolers.py:85-96: [R,5] = (image index as float, x1, y1, x2, y2)."""
        if isinstance(boxes, torch.Tensor):           # already [R,5] (pre-built by the data path)
            return boxes
        concat = torch.cat([b.bbox for b in boxes], dim=0)
        ids = torch.cat([torch.full((len(b), 1), i, dtype=concat.dtype, device=concat.device)
                         for i, b in enumerate(boxes)], dim=0)
        return torch.cat([ids, concat], dim=1)

    def forward(self, x, boxes):
        rois = self.convert_to_roi_format(boxes)
        return self.poolers[0](x[0], rois)
