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


def gen_input_pipeline():
    """N3 remainder: Resize.get_size, the whole train transform chain and the proposal lookup of PascalVOCDataset, from
    the reference's own classes (data/transforms/transforms.py, data/datasets/voc.py:87-111)."""
    ref_shims.install()
    from PIL import Image
    from wetectron.data.transforms import transforms as T
    from wetectron.data.datasets.coco import unique_boxes
    from wetectron.structures.bounding_box import BoxList
    from wetectron.structures.boxlist_ops import remove_small_boxes
    out = {}
    # (1) get_size over shapes x scales (random.choice replaced by a 1-element tuple per call)
    cases = []
    for (w, h) in ((500, 375), (375, 500), (333, 500), (500, 500), (1000, 240), (640, 480)):
        for s in (480, 576, 688, 864, 1200, 600):
            for mx in (2000, 1000):
                oh, ow = T.Resize(s, mx).get_size((w, h))
                cases.append([w, h, s, mx, oh, ow])
    out["get_size"] = np.array(cases)
    # (2) one example through Resize -> RandomHorizontalFlip(prob 1 / 0) -> ToTensor -> Normalize
    rs = np.random.RandomState(5)
    arr = rs.randint(0, 256, size=(75, 100, 3), dtype=np.uint8)
    out["img_u8"] = arr
    raw = np.stack([rs.uniform(0, 70, 60), rs.uniform(0, 45, 60), rs.uniform(0, 70, 60) + 29, rs.uniform(0, 45, 60) + 29], 1).round().astype(np.float32)
    raw[7] = raw[3]
    out["raw_rois"] = raw
    ids = [11, 42, 7]
    prop = {"boxes": [raw[::2], raw, raw[5:25]], "indexes": ids}
    for k, i in enumerate(ids):
        out["prop_boxes_%d" % k] = prop["boxes"][k]
    out["prop_ids"] = np.array(ids)
    # dataset lookup + filter (voc.py:87-111) for image id 42 on a 100x75 image
    rois = prop["boxes"][prop["indexes"].index(42)]
    rois = rois[unique_boxes(rois), :]
    bl = BoxList(torch.tensor(rois.astype(np.float64)), (100, 75), mode="xyxy").clip_to_image(remove_empty=True)
    bl = remove_small_boxes(boxlist=bl, min_size=20)
    out["lookup_42"] = bl.bbox.numpy()
    for tag, flip in (("flip", 1.0), ("noflip", 0.0)):
        tf = T.Compose([T.Resize(120, 400), T.RandomHorizontalFlip(flip), T.ToTensor(),
                        T.Normalize(mean=[102.9801, 115.9465, 122.7717], std=[1.0, 1.0, 1.0], to_bgr255=True)])
        img = Image.fromarray(arr, "RGB")
        tgt = BoxList(torch.tensor([[10., 12., 60., 50.]]), img.size, mode="xyxy")
        im, tg, ro = tf(img, tgt, BoxList(bl.bbox.clone(), img.size, mode="xyxy"))
        out["ex_%s_img" % tag], out["ex_%s_rois" % tag], out["ex_%s_tgt" % tag] = im.numpy(), ro.bbox.numpy(), tg.bbox.numpy()
        out["ex_%s_size" % tag] = np.array(ro.size)
    path = os.path.join(ROOT, "tests", "golden", "input_pipeline.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    if "--tta" in sys.argv:
        gen_tta()
    elif "--postprocess" in sys.argv:
        gen_postprocess()
    elif "--input" in sys.argv:
        gen_input_pipeline()
    else:
        main()
