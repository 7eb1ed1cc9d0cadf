"""Test-time augmentation loop (engine/bbox_aug.py:11-141; the second half of SURVEY 8f row N4): the detector is run on
every (scale, flip) view of the images, the per-proposal class scores and decoded boxes of the views are brought back to
the frame of the first view and averaged ("AVG") or concatenated ("UNION"), and ONE filter_results pass (all-class NMS +
detections cap, modeling/postprocess.py) produces the detections.

The views are handed in already transformed (resized / flipped / normalised image tensors and their proposals -- data.py
has the box-side transforms; image resampling is the data loader's job): one dict per view with
    images  [B,3,H,W] tensor on the device (or an ImageList)
    rois    list of BoxList, proposals in that view's frame
    hflip   True when the view is mirrored (its detections are mirrored back, BoxList.transpose(FLIP_LEFT_RIGHT))
The first view is the identity transform (bbox_aug.py:26-30).  The model must be in eval mode with
`roi_heads.strong_post_processor.bbox_aug_enabled = True`, so that each call returns the un-filtered [N*C] boxes and scores
(box_head/inference.py:86)."""
import torch

from . import data
from .structures import BoxList


def to_first_frame(boxlist, hflip, first_size):
    """Undo the view's mirror, then scale the boxes to the first view's image size (bbox_aug.py:16-24,118-119)."""
    b = boxlist.bbox
    if hflip:
        b = data.hflip_boxes(b, boxlist.size[0])
    if tuple(boxlist.size) != tuple(first_size):
        b = data.resize_boxes(b, boxlist.size, first_size)
    out = BoxList(b, first_size, "xyxy")
    out.add_field("scores", boxlist.get_field("scores"))
    return out


def merge_views(boxlists_t, heur="AVG"):
    """boxlists_t: the same image under every view, already in the first view's frame (bbox_aug.py:55-66)."""
    if heur == "UNION":
        bbox = torch.cat([b.bbox for b in boxlists_t])
        scores = torch.cat([b.get_field("scores") for b in boxlists_t])
    elif heur == "AVG":
        bbox = torch.mean(torch.stack([b.bbox for b in boxlists_t]), dim=0)
        scores = torch.mean(torch.stack([b.get_field("scores") for b in boxlists_t]), dim=0)
    else:
        raise ValueError("please use proper BBOX_AUG.HEUR ")
    out = BoxList(bbox, boxlists_t[0].size, "xyxy")
    out.add_field("scores", scores)
    return out


@torch.no_grad()
def im_detect_bbox_aug(model, views, num_classes, heur="AVG"):
    """Run every view, merge per image, filter once.  Returns one BoxList of detections per image."""
    pp = model.roi_heads.strong_post_processor
    if not pp.bbox_aug_enabled:
        raise RuntimeError("set model.roi_heads.strong_post_processor.bbox_aug_enabled = True for the TTA loop")
    per_image = None
    for view in views:
        boxlists = model(view["images"], None, view["rois"])
        if per_image is None:
            per_image = [[] for _ in boxlists]
        for i, bl in enumerate(boxlists):
            first_size = per_image[i][0].size if per_image[i] else bl.size
            per_image[i].append(to_first_frame(bl, bool(view.get("hflip", False)) , first_size))
    results = []
    for boxlists_t in per_image:
        merged = merge_views(boxlists_t, heur)
        n = merged.bbox.shape[0] // num_classes
        results.append(pp.filter_results(merged.bbox.reshape(n, num_classes * 4),
                                         merged.get_field("scores").reshape(n, num_classes), merged.size))
    return results
