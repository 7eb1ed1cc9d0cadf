"""Golden vectors for the input side from the REFERENCE ITSELF: data/datasets/coco.py unique_boxes,
BoxList.clip_to_image / resize / transpose, boxlist_ops.remove_small_boxes, transforms.Normalize, to_image_list.
Run in the build container only:  python -m oracle.gen_golden_data   (TEST INFRASTRUCTURE ONLY)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims              # noqa: E402


def main():
    ref_shims.install()
    from wetectron.data.datasets.coco import unique_boxes
    from wetectron.data.transforms.transforms import Normalize
    from wetectron.structures.bounding_box import FLIP_LEFT_RIGHT, BoxList
    from wetectron.structures.boxlist_ops import remove_small_boxes
    from wetectron.structures.image_list import to_image_list
    rs = np.random.RandomState(7)
    W, H = 500, 375
    n = 400
    x1 = rs.uniform(-30, W, n); y1 = rs.uniform(-30, H, n)
    raw = np.stack([x1, y1, x1 + rs.uniform(1, 300, n), y1 + rs.uniform(1, 250, n)], 1).round()
    raw[50:80] = raw[10:40]                         # duplicates
    raw[100:110, 2] = raw[100:110, 0] + 5           # small boxes
    raw = raw.astype(np.float32)
    keep = unique_boxes(raw)
    rois = raw[keep, :]
    bl = BoxList(torch.tensor(rois.astype(np.float64)), (W, H), mode="xyxy")
    bl = bl.clip_to_image(remove_empty=True)
    bl = remove_small_boxes(boxlist=bl, min_size=20)
    out = {"raw": raw, "W": W, "H": H, "unique_keep": keep, "filtered": bl.bbox.numpy()}
    for tag, size in (("eq", (1000, 750)), ("neq", (864, 600))):
        out["resized_" + tag] = bl.resize(size).bbox.numpy()
        out["size_" + tag] = np.array(size)
    out["flipped"] = bl.transpose(FLIP_LEFT_RIGHT).bbox.numpy()
    g = torch.Generator().manual_seed(3)
    img = torch.rand(3, 37, 53, generator=g)
    out["img"] = img.numpy()
    out["normalized"] = Normalize(mean=[102.9801, 115.9465, 122.7717], std=[1.0, 1.0, 1.0], to_bgr255=True)(img)[0].numpy()
    imgs = [torch.rand(3, h, w, generator=g) for h, w in ((37, 53), (64, 40), (50, 70))]
    il = to_image_list(imgs, 32)
    for i, im in enumerate(imgs):
        out["il_in%d" % i] = im.numpy()
    out["il_tensors"] = il.tensors.numpy()
    out["il_sizes"] = np.array([list(s) for s in il.image_sizes])
    path = os.path.join(ROOT, "tests", "golden", "data_side.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})




def gen_postprocess():
    """N4: PostProcessor.filter_results and BoxCoder.decode of the reference (CPU)."""
    ref_shims.install()
    from wetectron.modeling.box_coder import BoxCoder
    from wetectron.modeling.roi_heads.box_head.inference import PostProcessor
    g = torch.Generator().manual_seed(11)
    out = {}
    for tag, (N, C, W, H, cap) in {"a": (300, 21, 500, 375, 100), "b": (900, 21, 800, 600, 100), "c": (64, 81, 640, 480, 40)}.items():
        x1 = torch.rand(N, generator=g) * (W - 60); y1 = torch.rand(N, generator=g) * (H - 60)
        ref_boxes = torch.stack([x1, y1, x1 + 20 + torch.rand(N, generator=g) * (W - 21 - x1),
                                 y1 + 20 + torch.rand(N, generator=g) * (H - 21 - y1)], 1).round()
        reg = torch.randn(N, C * 4, generator=g) * 0.5
        scores = torch.softmax(torch.randn(N, C, generator=g) * 2, 1)
        scores[::7] = scores[3]                                   # exact score ties across proposals
        coder = BoxCoder(weights=(10., 10., 5., 5.))
        dec = coder.decode(reg, ref_boxes)
        dec_saved = dec.clone()                                   # prepare_boxlist + clip_to_image clamp `dec` in place
        pp = PostProcessor(score_thresh=0.0 if tag != "c" else 0.02, nms=0.4, detections_per_img=cap, box_coder=coder)
        bl = pp.prepare_boxlist(dec, scores, (W, H)).clip_to_image(remove_empty=False)
        clipped = bl.bbox.reshape(N, C * 4).clone()
        res = pp.filter_results(bl, C)
        out.update({tag + "_ref_boxes": ref_boxes.numpy(), tag + "_reg": reg.numpy(), tag + "_scores": scores.numpy(),
                    tag + "_decoded": dec_saved.numpy(), tag + "_clipped": clipped.numpy(), tag + "_size": np.array([W, H]),
                    tag + "_cap": cap, tag + "_thr": pp.score_thresh,
                    tag + "_out_boxes": res.bbox.numpy(), tag + "_out_scores": res.get_field("scores").numpy(),
                    tag + "_out_labels": res.get_field("labels").numpy()})
    path = os.path.join(ROOT, "tests", "golden", "postprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items() if "out" in k})


def gen_tta():
    """N4: the merge of engine/bbox_aug.py (transpose back, resize to the first view, mean) with the reference's BoxList."""
    ref_shims.install()
    from wetectron.structures.bounding_box import FLIP_LEFT_RIGHT, BoxList
    g = torch.Generator().manual_seed(23)
    n = 120
    views = [((500, 375), False), ((500, 375), True), ((640, 480), False), ((1000, 750), True), ((864, 600), False)]
    out, lists = {}, []
    for v, (size, flip) in enumerate(views):
        W, H = size
        x1 = torch.rand(n, generator=g) * (W - 40); y1 = torch.rand(n, generator=g) * (H - 40)
        b = torch.stack([x1, y1, x1 + 10 + torch.rand(n, generator=g) * 25, y1 + 10 + torch.rand(n, generator=g) * 25], 1)
        sc = torch.rand(n, generator=g)
        bl = BoxList(b.clone(), size, "xyxy"); bl.add_field("scores", sc)
        out["v%d_boxes" % v], out["v%d_scores" % v], out["v%d_size" % v], out["v%d_flip" % v] = b.numpy(), sc.numpy(), np.array(size), flip
        if flip:
            bl = bl.transpose(FLIP_LEFT_RIGHT)
        lists.append(bl if v == 0 else bl.resize(lists[0].size))
    out["avg_boxes"] = torch.mean(torch.stack([b.bbox for b in lists]), dim=0).numpy()
    out["avg_scores"] = torch.mean(torch.stack([b.get_field("scores") for b in lists]), dim=0).numpy()
    out["union_boxes"] = torch.cat([b.bbox for b in lists]).numpy()
    out["n_views"] = len(views)
    path = os.path.join(ROOT, "tests", "golden", "tta_merge.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    if "--tta" in sys.argv:
        gen_tta()
    elif "--postprocess" in sys.argv:
        gen_postprocess()
    else:
        main()
