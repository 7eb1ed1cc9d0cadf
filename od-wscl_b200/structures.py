"""Mirror of the parts of ``wetectron.structures`` the hot path touches: BoxList
(structures/bounding_box.py:13-260), boxlist_iou / boxlist_nms_index (structures/boxlist_ops.py:
38-61,127-160), ImageList / to_image_list (structures/image_list.py:33-75).  Masks / keypoints are
out of scope."""
import torch

from . import capi


class BoxList(object):
    """xyxy (or xywh) boxes of one image + named fields; legacy +1 pixel convention."""

    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox should have shape [n, 4], got {}".format(tuple(bbox.shape)))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox, self.size, self.mode = bbox, image_size, mode
        self.extra_fields = {}

    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def convert(self, mode):
        if mode == self.mode:
            return self
        b = self.bbox
        if mode == "xyxy":      # from xywh (bounding_box.py:86-96)
            nb = torch.stack((b[:, 0], b[:, 1], b[:, 0] + (b[:, 2] - 1).clamp(min=0), b[:, 1] + (b[:, 3] - 1).clamp(min=0)), 1)
        else:
            nb = torch.stack((b[:, 0], b[:, 1], b[:, 2] - b[:, 0] + 1, b[:, 3] - b[:, 1] + 1), 1)
        out = BoxList(nb, self.size, mode)
        out.extra_fields = dict(self.extra_fields)
        return out

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)      # bounding_box.py:231-241
        return b[:, 2] * b[:, 3]

    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def __repr__(self):
        return "BoxList(num_boxes=%d, image_width=%s, image_height=%s, mode=%s)" % (
            len(self), self.size[0], self.size[1], self.mode)


def boxlist_iou(boxlist1, boxlist2):
    """structures/boxlist_ops.py:127-160 (+1 convention), one kernel instead of ~8."""
    if boxlist1.size != boxlist2.size:
        raise RuntimeError("boxlists should have same image size, got {}, {}".format(boxlist1, boxlist2))
    return capi.box_iou(boxlist1.convert("xyxy").bbox, boxlist2.convert("xyxy").bbox, plus_one=True)


def boxlist_nms_index(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    """structures/boxlist_ops.py:38-61: torchvision-semantics NMS, returns (boxlist, keep)."""
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = capi.nms(boxlist.bbox, boxlist.get_field(score_field), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    return boxlist[keep].convert(mode), keep


def cat_boxlist(bboxes):
    size, mode = bboxes[0].size, bboxes[0].mode
    out = BoxList(torch.cat([b.bbox for b in bboxes], 0), size, mode)
    for f in bboxes[0].fields():
        out.add_field(f, torch.cat([b.get_field(f) for b in bboxes], 0))
    return out


class ImageList(object):
    def __init__(self, tensors, image_sizes):
        self.tensors, self.image_sizes = tensors, image_sizes

    def to(self, *args, **kwargs):
        return ImageList(self.tensors.to(*args, **kwargs), self.image_sizes)


def to_image_list(tensors, size_divisible=0):
    """structures/image_list.py:33-75: zero-pad a list of [C,H,W] to a common /size_divisible shape."""
    if isinstance(tensors, ImageList):
        return tensors
    if isinstance(tensors, torch.Tensor) and tensors.dim() == 4 and size_divisible == 0:
        return ImageList(tensors, [t.shape[-2:] for t in tensors])
    if isinstance(tensors, torch.Tensor):
        tensors = list(tensors) if tensors.dim() == 4 else [tensors]
    max_size = [max(s) for s in zip(*[t.shape for t in tensors])]
    if size_divisible > 0:
        d = size_divisible
        max_size[1] = (max_size[1] + d - 1) // d * d
        max_size[2] = (max_size[2] + d - 1) // d * d
    batched = tensors[0].new_zeros((len(tensors),) + tuple(max_size))
    for img, pad in zip(tensors, batched):
        pad[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
    return ImageList(batched, [t.shape[-2:] for t in tensors])
