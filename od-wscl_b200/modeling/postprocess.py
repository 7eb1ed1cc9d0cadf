"""Test-time post-processing: PostProcessor of roi_heads/box_head/inference.py:13-258 as the weak head
uses it (weak_head.py:124-145, HEUR "AVG": mean of the three refinement heads' softmaxed scores and box deltas,
softmax_on=False) -- BoxCoder.decode (modeling/box_coder.py:52-95), clip_to_image(remove_empty=False), then
filter_results: per-class score threshold + NMS + the detections_per_img cap.  The per-class loop of 20-80 torchvision NMS
calls (each with its own host sweep) is ONE launch of odwscl_nms_per_class_f32; the cap keeps the reference's kthvalue
rule (ties at the threshold are all kept) without leaving the device."""
import math

import torch

from .. import capi
from ..structures import BoxList

BBOX_XFORM_CLIP = math.log(1000.0 / 16)


def decode_boxes(rel_codes, boxes, weights=(10.0, 10.0, 5.0, 5.0)):
    """BoxCoder.decode (modeling/box_coder.py:52-95): deltas [N, C*4] applied to reference boxes [N,4] -> [N, C*4].
    Written over a [N,C,4] view; every element goes through the reference's operations in the reference's order
    (+1 widths, centre + 0.5 w, delta / weight, clamp of dw / dh at log(1000/16), exp, the "- 1" on x2 / y2)."""
    ref = boxes.to(rel_codes.dtype)
    n = rel_codes.shape[0]
    d = rel_codes.view(n, -1, 4)
    size = ref[:, 2:4] - ref[:, 0:2] + 1                        # [N,2] (w, h)
    centre = ref[:, 0:2] + 0.5 * size                           # [N,2]
    w_xy = torch.as_tensor(weights[0:2], dtype=d.dtype, device=d.device)
    w_wh = torch.as_tensor(weights[2:4], dtype=d.dtype, device=d.device)
    shift = d[:, :, 0:2] / w_xy                                 # dx, dy
    grow = torch.clamp(d[:, :, 2:4] / w_wh, max=BBOX_XFORM_CLIP)  # dw, dh
    new_centre = shift * size[:, None, :] + centre[:, None, :]
    new_size = torch.exp(grow) * size[:, None, :]
    out = torch.empty_like(d)
    out[:, :, 0:2] = new_centre - 0.5 * new_size
    out[:, :, 2:4] = new_centre + 0.5 * new_size - 1
    return out.view(n, -1)


def clip_boxes(boxes_nc4, width, height):
    """BoxList.clip_to_image(remove_empty=False) on the [N*C,4] view (TO_REMOVE = 1)."""
    b = boxes_nc4.reshape(-1, 4).clone()
    b[:, 0].clamp_(min=0, max=width - 1)
    b[:, 1].clamp_(min=0, max=height - 1)
    b[:, 2].clamp_(min=0, max=width - 1)
    b[:, 3].clamp_(min=0, max=height - 1)
    return b.reshape(boxes_nc4.shape)


class PostProcessor(torch.nn.Module):
    def __init__(self, score_thresh=0.05, nms=0.5, detections_per_img=100, weights=(10.0, 10.0, 5.0, 5.0),
                 cls_agnostic_bbox_reg=False, bbox_aug_enabled=False):
        super().__init__()
        self.score_thresh, self.nms, self.detections_per_img = score_thresh, nms, detections_per_img
        self.weights = weights
        self.cls_agnostic_bbox_reg = cls_agnostic_bbox_reg
        self.bbox_aug_enabled = bbox_aug_enabled

    def forward(self, x, boxes, softmax_on=True):
        class_logits, box_regression = x
        class_prob = torch.softmax(class_logits, -1) if softmax_on else class_logits
        sizes = [len(b) for b in boxes]
        concat = torch.cat([b.bbox for b in boxes], dim=0)
        if self.cls_agnostic_bbox_reg:
            box_regression = box_regression[:, -4:]
        proposals = decode_boxes(box_regression.view(sum(sizes), -1), concat, self.weights)
        if self.cls_agnostic_bbox_reg:
            proposals = proposals.repeat(1, class_prob.shape[1])
        C = class_prob.shape[1]
        results = []
        for prob, bx, ref in zip(class_prob.split(sizes), proposals.split(sizes), boxes):
            w, h = ref.size
            bx = clip_boxes(bx, w, h)
            if self.bbox_aug_enabled:                       # inference.py:86: TTA filters after averaging
                r = BoxList(bx.reshape(-1, 4), ref.size, "xyxy")
                r.add_field("scores", prob.reshape(-1))
            else:
                r = self.filter_results(bx, prob, ref.size)
            results.append(r)
        return results

    def filter_results(self, boxes, scores, image_size):
        """boxes [N,C*4], scores [N,C] -> BoxList with `scores` and `labels` (class-major, descending score in a class)."""
        N, C = scores.shape
        dev = scores.device
        keep, cnt = capi.nms_per_class(boxes, scores, self.score_thresh, self.nms)
        ar = torch.arange(keep.shape[1], device=dev)[None]
        valid = ar < cnt[:, None]                                            # [C,N]
        cls = torch.arange(C, device=dev)[:, None].expand_as(keep)[valid]     # class-major order = cat_boxlist order
        idx = keep[valid].long()
        sc = scores[idx, cls]
        bx = boxes.view(N, C, 4)[idx, cls]
        n_det = sc.numel()                                                   # one sync, as len(result) in the reference
        if n_det > self.detections_per_img > 0:
            thr, _ = torch.kthvalue(sc, n_det - self.detections_per_img + 1)
            sel = (sc >= thr).nonzero(as_tuple=False).squeeze(1)
            bx, sc, cls = bx[sel], sc[sel], cls[sel]
        out = BoxList(bx, image_size, "xyxy")
        out.add_field("scores", sc)
        out.add_field("labels", cls)
        return out
