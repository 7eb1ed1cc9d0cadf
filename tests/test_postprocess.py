"""Test-time post-processing against the reference's own PostProcessor / BoxCoder
(oracle/gen_golden_data.py --postprocess -> tests/golden/postprocess.npz)."""
import numpy as np
import pytest
import torch

from odwscl_b200.modeling import postprocess as pp


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_decode_and_clip_match_reference_cpu(golden, tag):
    G = golden("postprocess.npz")
    dec = pp.decode_boxes(torch.from_numpy(G[tag + "_reg"]), torch.from_numpy(G[tag + "_ref_boxes"]))
    assert np.array_equal(dec.numpy(), G[tag + "_decoded"])                      # same torch-CPU ops, same order
    W, H = (int(v) for v in G[tag + "_size"])
    assert np.array_equal(pp.clip_boxes(dec, W, H).numpy(), G[tag + "_clipped"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_filter_results_matches_reference(golden, tag):
    """Per-class threshold + NMS(0.4) + detections cap, all classes in one launch: same detections, same order,
    bit-exact (index selection on identical fp32 inputs; exact score ties included)."""
    G = golden("postprocess.npz")
    W, H = (int(v) for v in G[tag + "_size"])
    proc = pp.PostProcessor(score_thresh=float(G[tag + "_thr"]), nms=0.4, detections_per_img=int(G[tag + "_cap"]))
    res = proc.filter_results(torch.from_numpy(G[tag + "_clipped"]).cuda(), torch.from_numpy(G[tag + "_scores"]).cuda(), (W, H))
    assert np.array_equal(res.get_field("labels").cpu().numpy(), G[tag + "_out_labels"])
    assert np.array_equal(res.get_field("scores").cpu().numpy(), G[tag + "_out_scores"])
    assert np.array_equal(res.bbox.cpu().numpy(), G[tag + "_out_boxes"])


@pytest.mark.gpu
def test_postprocessor_end_to_end_and_edges(golden):
    from odwscl_b200.structures import BoxList
    G = golden("postprocess.npz")
    W, H = (int(v) for v in G["a_size"])
    proc = pp.PostProcessor(score_thresh=0.0, nms=0.4, detections_per_img=100)
    props = [BoxList(torch.from_numpy(G["a_ref_boxes"]).cuda(), (W, H), "xyxy")]
    out = proc((torch.from_numpy(G["a_scores"]).cuda(), torch.from_numpy(G["a_reg"]).cuda()), props, softmax_on=False)[0]
    # decode runs on the GPU here (expf differs from the CPU's in the last ulp): same detections up to 1e-4 px
    assert np.array_equal(out.get_field("labels").cpu().numpy(), G["a_out_labels"])
    np.testing.assert_allclose(out.bbox.cpu().numpy(), G["a_out_boxes"], rtol=0, atol=1e-3)
    # nothing above the threshold -> empty result; a single proposal -> itself per class
    none = proc.filter_results(torch.zeros(5, 84).cuda(), torch.zeros(5, 21).cuda(), (W, H))
    assert len(none) == 0
    one = proc.filter_results(torch.tensor([[1.0, 2.0, 30.0, 40.0] * 3]).cuda(), torch.tensor([[0.2, 0.5, 0.3]]).cuda(), (W, H))
    assert one.get_field("labels").tolist() == [1, 2] and one.get_field("scores").tolist() == [0.5, pytest.approx(0.3)]


@pytest.mark.gpu
def test_model_eval_returns_detections():
    """GeneralizedRCNN in eval mode (generalized_rcnn.py:93-97): one BoxList of detections per image with `scores` and
    `labels`, at most DETECTIONS_PER_IMG (+ ties), labels in 1..C-1, boxes inside the image."""
    from odwscl_b200.config import cfg
    from odwscl_b200.modeling import build_detection_model
    from odwscl_b200.structures import BoxList
    from odwscl_b200.synth import synth_batch
    torch.manual_seed(0)
    model = build_detection_model(cfg).cuda().eval()
    images, _, boxes, _ = synth_batch(2, 300, 400, 320, seed=5)
    props = [BoxList(b.cuda(), (400, 320), "xyxy") for b in boxes]
    with torch.no_grad():
        res = model(images.cuda(), None, props)
    assert len(res) == 2
    for r in res:
        lab, sc = r.get_field("labels"), r.get_field("scores")
        assert 0 < len(r) <= 100 + 20 and lab.numel() == sc.numel() == len(r)
        assert int(lab.min()) >= 1 and int(lab.max()) <= 20
        b = r.bbox
        assert float(b[:, 0].min()) >= 0 and float(b[:, 2].max()) <= 399 and float(b[:, 3].max()) <= 319
        # descending scores inside every class (NMS keep order), classes ascending
        assert bool((lab[1:] >= lab[:-1]).all())
        same = lab[1:] == lab[:-1]
        assert bool((sc[1:][same] <= sc[:-1][same]).all())


def test_tta_merge_matches_reference(golden):
    """engine/bbox_aug.py merge: mirror back, resize to the first view's frame, AVG / UNION -- bit-exact against the
    reference's BoxList.transpose / resize / torch.mean (oracle/gen_golden_data.py --tta)."""
    from odwscl_b200 import bbox_aug
    from odwscl_b200.structures import BoxList
    G = golden("tta_merge.npz")
    lists = []
    for v in range(int(G["n_views"])):
        bl = BoxList(torch.from_numpy(G["v%d_boxes" % v]), tuple(int(x) for x in G["v%d_size" % v]), "xyxy")
        bl.add_field("scores", torch.from_numpy(G["v%d_scores" % v]))
        first = lists[0].size if lists else bl.size
        lists.append(bbox_aug.to_first_frame(bl, bool(G["v%d_flip" % v]), first))
    avg = bbox_aug.merge_views(lists, "AVG")
    assert np.array_equal(avg.bbox.numpy(), G["avg_boxes"]) and np.array_equal(avg.get_field("scores").numpy(), G["avg_scores"])
    assert np.array_equal(bbox_aug.merge_views(lists, "UNION").bbox.numpy(), G["union_boxes"])
    with pytest.raises(ValueError):
        bbox_aug.merge_views(lists, "MAX")


@pytest.mark.gpu
def test_tta_loop_end_to_end():
    """Three views (identity, mirrored, rescaled) through the eval-mode detector, merged and filtered once."""
    from odwscl_b200 import bbox_aug, data
    from odwscl_b200.config import cfg
    from odwscl_b200.modeling import build_detection_model
    from odwscl_b200.structures import BoxList
    from odwscl_b200.synth import synth_batch
    torch.manual_seed(0)
    model = build_detection_model(cfg).cuda().eval()
    model.roi_heads.strong_post_processor.bbox_aug_enabled = True
    images, _, boxes, _ = synth_batch(1, 200, 400, 320, seed=9)
    W, H = 400, 320
    img = images[:, :, :H, :W]
    views = [{"images": images.cuda(), "rois": [BoxList(boxes[0].cuda(), (W, H), "xyxy")], "hflip": False},
             {"images": images.flip(3).cuda() if images.shape[3] == W else torch.nn.functional.pad(img.flip(3), (0, images.shape[3] - W)).cuda(),
              "rois": [BoxList(data.hflip_boxes(boxes[0], W).cuda(), (W, H), "xyxy")], "hflip": True}]
    big = torch.nn.functional.interpolate(img, size=(480, 600), mode="bilinear", align_corners=False)
    big_p, _ = data.to_image_list([big[0]], 32)
    views.append({"images": big_p.cuda(), "rois": [BoxList(data.resize_boxes(boxes[0], (W, H), (600, 480)).cuda(), (600, 480), "xyxy")],
                  "hflip": False})
    res = bbox_aug.im_detect_bbox_aug(model, views, 21, "AVG")
    assert len(res) == 1 and 0 < len(res[0]) <= 120
    r = res[0]
    assert tuple(r.size) == (W, H) and int(r.get_field("labels").min()) >= 1
    assert bool(torch.isfinite(r.bbox).all()) and bool(torch.isfinite(r.get_field("scores")).all())
