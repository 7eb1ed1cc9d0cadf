"""Synthetic workload of SURVEY.md 8(d) / BASELINE.json `configs`: images `randn * 50` padded to a
multiple of 32, MCG-style integer boxes (min side 20, clipped, de-duplicated: data/datasets/voc.py:
108-111), two labelled classes per image.  Host tensors (pinned on request) -- the bench copies
them to the device inside its end-to-end timed region."""
import torch


def synth_boxes(n, W, H, gen):
    out = torch.zeros((0, 4))
    while out.shape[0] < n:
        k = 2 * n
        x1 = torch.rand(k, generator=gen) * (W - 40)
        y1 = torch.rand(k, generator=gen) * (H - 40)
        w = 20 + torch.rand(k, generator=gen) * (W - 21 - x1)
        h = 20 + torch.rand(k, generator=gen) * (H - 21 - y1)
        b = torch.stack([x1, y1, x1 + w, y1 + h], 1).round()
        b[:, 0::2].clamp_(0, W - 1)
        b[:, 1::2].clamp_(0, H - 1)
        ok = ((b[:, 2] - b[:, 0]) >= 20) & ((b[:, 3] - b[:, 1]) >= 20)
        out = torch.unique(torch.cat([out, b[ok]]), dim=0)
        out = out[torch.randperm(out.shape[0], generator=gen)]
    return out[:n].contiguous()


def synth_batch(B, N, W, H, num_classes=21, seed=1234, pin=False, labels_per_image=2):
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    images = torch.zeros(B, 3, Hp, Wp)
    boxes, labels = [], []
    for i in range(B):
        g = torch.Generator().manual_seed(seed + i)
        images[i, :, :H, :W] = torch.randn(3, H, W, generator=g) * 50
        boxes.append(synth_boxes(N, W, H, g))
        labels.append((torch.randperm(num_classes - 1, generator=g)[:labels_per_image] + 1).tolist())
    rois = torch.cat([torch.cat([torch.full((b.shape[0], 1), float(i)), b], 1) for i, b in enumerate(boxes)])
    if pin:
        images, rois = images.pin_memory(), rois.pin_memory()
    return images, rois, boxes, labels
