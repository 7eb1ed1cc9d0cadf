"""Input side of the path: what sits between the proposal files / decoded images and
GeneralizedRCNN.forward.  Host-side restatements of the reference's functions (numpy / torch-CPU, as in the reference)
plus the one B200-specific piece: a pinned-memory, copy-stream prefetcher so the H2D copy of batch i+1 runs under the
compute of batch i.

  unique_boxes            wetectron/utils/boxes... via data/datasets/voc.py:100 (hash rows, keep first occurrence)
  clip_to_image           structures/bounding_box.py:207-218 (TO_REMOVE = 1, remove_empty)
  remove_small_boxes      structures/boxlist_ops.py:64-77 (xywh sides >= min_size)
  filter_proposals        data/datasets/voc.py:100-111 = unique -> clip(remove_empty) -> remove_small(20)
  resize_boxes / hflip    structures/bounding_box.py:93-128,130-166 (ratio scaling; flip with TO_REMOVE = 1)
  normalize_image         data/transforms/transforms.py:122-132 (to BGR 0-255, subtract PIXEL_MEAN, divide by PIXEL_STD)
  to_image_list           structures/image_list.py:33-75 (zero-pad to the batch max, rounded up to SIZE_DIVISIBILITY)
"""
import math

import numpy as np
import torch

PIXEL_MEAN = (102.9801, 115.9465, 122.7717)       # config/defaults.py INPUT.PIXEL_MEAN (BGR)
PIXEL_STD = (1.0, 1.0, 1.0)


def unique_boxes(boxes, scale=1.0):
    """Indices of the first occurrence of every distinct (rounded) box, ascending (utils hash trick of the reference)."""
    v = np.array([1, 1e3, 1e6, 1e9])
    hashes = np.round(np.asarray(boxes, dtype=np.float64) * scale).dot(v)
    _, index = np.unique(hashes, return_index=True)
    return np.sort(index)


def clip_to_image(boxes, width, height, remove_empty=True):
    b = boxes.clone()
    b[:, 0].clamp_(min=0, max=width - 1)
    b[:, 1].clamp_(min=0, max=height - 1)
    b[:, 2].clamp_(min=0, max=width - 1)
    b[:, 3].clamp_(min=0, max=height - 1)
    if remove_empty:
        keep = (b[:, 3] > b[:, 1]) & (b[:, 2] > b[:, 0])
        b = b[keep]
    return b


def remove_small_boxes(boxes, min_size):
    ws = boxes[:, 2] - boxes[:, 0] + 1
    hs = boxes[:, 3] - boxes[:, 1] + 1
    keep = ((ws >= min_size) & (hs >= min_size)).nonzero().squeeze(1)
    return boxes[keep]


def filter_proposals(rois, width, height, min_size=20):
    """numpy [n,4] proposals of one image -> float32 tensor as the reference's dataset hands them to the transforms."""
    rois = np.asarray(rois)
    rois = rois[unique_boxes(rois), :]
    b = torch.tensor(rois.astype(np.float64)).float()          # BoxList stores float32
    b = clip_to_image(b, width, height, remove_empty=True)
    return remove_small_boxes(b, min_size)


def resize_boxes(boxes, old_size, new_size):
    """sizes are (width, height)"""
    rw, rh = new_size[0] / old_size[0], new_size[1] / old_size[1]
    if rw == rh:
        return boxes * rw
    x1, y1, x2, y2 = boxes.split(1, dim=-1)
    return torch.cat((x1 * rw, y1 * rh, x2 * rw, y2 * rh), dim=-1)


def hflip_boxes(boxes, width):
    x1, y1, x2, y2 = boxes.split(1, dim=-1)
    return torch.cat((width - x2 - 1, y1, width - x1 - 1, y2), dim=-1)


def normalize_image(img_rgb01, to_bgr255=True, mean=PIXEL_MEAN, std=PIXEL_STD):
    """[3,H,W] RGB in [0,1] (ToTensor) -> the network input."""
    x = img_rgb01[[2, 1, 0]] * 255 if to_bgr255 else img_rgb01
    mean = torch.as_tensor(mean, dtype=x.dtype)[:, None, None]
    std = torch.as_tensor(std, dtype=x.dtype)[:, None, None]
    return (x - mean) / std


def to_image_list(tensors, size_divisible=32, pin=False):
    """list of [3,h,w] -> ([B,3,H,W] zero-padded, [(h,w)...]); H, W = batch max rounded up to size_divisible."""
    max_size = [max(s) for s in zip(*[img.shape for img in tensors])]
    if size_divisible > 0:
        max_size[1] = int(math.ceil(max_size[1] / size_divisible) * size_divisible)
        max_size[2] = int(math.ceil(max_size[2] / size_divisible) * size_divisible)
    batched = torch.zeros((len(tensors), *max_size), dtype=tensors[0].dtype, pin_memory=pin)
    for img, pad in zip(tensors, batched):
        pad[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
    return batched, [tuple(im.shape[-2:]) for im in tensors]


class HostPrefetcher:
    """Double-buffered H2D staging: `next()` returns the device copies of the batch handed to the previous `feed()` and
    the copy of the following batch is already running on a dedicated stream.  Tensors must be pinned."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.pending = None

    def feed(self, *host_tensors):
        with torch.cuda.stream(self.stream):
            dev = [t.to(self.device, non_blocking=True) for t in host_tensors]
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self.pending = (dev, ev)

    def next(self):
        dev, ev = self.pending
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)
        self.pending = None
        return dev


# ------------------------------------------------------------------------------------------------------------------
# N3 remainder: proposal file -> per-image proposals, scale selection, image resize, one training example
def resize_size(image_size, min_size, max_size, pick=None):
    """Resize.get_size (data/transforms/transforms.py:42-64): target (height, width) for an image of (width, height).
    min_size: int or sequence (multi-scale: one is drawn per image, `pick` = the drawn value for replay)."""
    import random
    w, h = image_size
    if not isinstance(min_size, (list, tuple)):
        min_size = (min_size,)
    size = pick if pick is not None else random.choice(min_size)
    if max_size is not None:
        mn, mx = float(min((w, h))), float(max((w, h)))
        if mx / mn * size > max_size:
            size = int(round(max_size * mn / mx))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        return (int(size * h / w), size)
    return (size, int(size * w / h))


def resize_image(img, size_hw):
    """torchvision.transforms.functional.resize on a PIL image as the reference calls it (transforms.py:68): bilinear."""
    from PIL import Image
    return img.resize((size_hw[1], size_hw[0]), Image.BILINEAR)


def image_to_tensor(img):
    """torchvision ToTensor (transforms.py:113-115): PIL RGB uint8 -> float32 [3,H,W] in [0,1]."""
    a = np.asarray(img, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None].repeat(3, axis=2)
    return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))).float().div(255)


class ProposalFile:
    """The pickled proposal file of the reference (`{'boxes': [ndarray per image], 'indexes' | 'ids': [image id ...],
    'scores': ...}`; data/datasets/voc.py:62-66,87-111): per-image lookup + the unique / clip / min-size filter."""

    def __init__(self, path_or_dict):
        if isinstance(path_or_dict, dict):
            self.data = path_or_dict
        else:
            import pickle
            with open(path_or_dict, "rb") as f:
                self.data = pickle.load(f, encoding="latin1")
        self.id_field = "indexes" if "indexes" in self.data else "ids"          # compat fix, voc.py:93
        self._pos = {int(i): k for k, i in enumerate(self.data[self.id_field])}

    def __len__(self):
        return len(self._pos)

    def rois(self, image_id, width, height, min_size=20):
        """float32 [n,4] xyxy proposals of one image as the dataset hands them to the transforms (voc.py:94-111)."""
        return filter_proposals(self.data["boxes"][self._pos[int(image_id)]], width, height, min_size)


def prepare_example(img, rois, min_size, max_size, flip, pick=None, target_boxes=None):
    """One training example through build_transforms(is_train=True) (data/transforms/build.py): Resize -> horizontal flip
    (the caller draws `flip`) -> ToTensor -> Normalize(to_bgr255).  img: PIL RGB; rois / target_boxes: float32 [n,4] in
    the original image.  Returns (image [3,h,w] float32, rois, target_boxes, (w, h))."""
    from PIL import Image
    w0, h0 = img.size
    size_hw = resize_size((w0, h0), min_size, max_size, pick)
    img = resize_image(img, size_hw)
    w1, h1 = img.size
    rois = resize_boxes(rois, (w0, h0), (w1, h1))
    if target_boxes is not None:
        target_boxes = resize_boxes(target_boxes, (w0, h0), (w1, h1))
    if flip:
        img = img.transpose(Image.FLIP_LEFT_RIGHT)
        rois = hflip_boxes(rois, w1)
        if target_boxes is not None:
            target_boxes = hflip_boxes(target_boxes, w1)
    return normalize_image(image_to_tensor(img)), rois, target_boxes, (w1, h1)
