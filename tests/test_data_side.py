"""Input side: host functions of od-wscl_b200/data.py against vectors produced by the reference's own
functions (oracle/gen_golden_data.py -> tests/golden/data_side.npz).  Bit-exact: integer / index work and fp32
elementwise arithmetic in the same order."""
import numpy as np
import pytest
import torch

from odwscl_b200 import data


def test_filter_resize_flip_match_reference(golden):
    G = golden("data_side.npz")
    W, H = int(G["W"]), int(G["H"])
    assert np.array_equal(data.unique_boxes(G["raw"]), G["unique_keep"])
    b = data.filter_proposals(G["raw"], W, H, min_size=20)
    assert b.dtype == torch.float32 and np.array_equal(b.numpy(), G["filtered"])
    for tag in ("eq", "neq"):
        size = tuple(int(x) for x in G["size_" + tag])
        assert np.array_equal(data.resize_boxes(b, (W, H), size).numpy(), G["resized_" + tag])
    assert np.array_equal(data.hflip_boxes(b, W).numpy(), G["flipped"])


def test_normalize_and_image_list_match_reference(golden):
    G = golden("data_side.npz")
    assert np.array_equal(data.normalize_image(torch.from_numpy(G["img"])).numpy(), G["normalized"])
    imgs = [torch.from_numpy(G["il_in%d" % i]) for i in range(3)]
    batched, sizes = data.to_image_list(imgs, 32)
    assert np.array_equal(batched.numpy(), G["il_tensors"])
    assert [list(s) for s in sizes] == G["il_sizes"].tolist()


def test_filter_edge_cases():
    empty = data.filter_proposals(np.zeros((0, 4), dtype=np.float32), 100, 80)
    assert empty.shape == (0, 4)
    one = data.filter_proposals(np.array([[-5, -5, 300, 300]], dtype=np.float32), 100, 80)
    assert one.tolist() == [[0.0, 0.0, 99.0, 79.0]]
    tiny = data.filter_proposals(np.array([[10, 10, 28, 40]], dtype=np.float32), 100, 80)       # 19 wide: dropped
    assert tiny.shape == (0, 4)


@pytest.mark.gpu
def test_host_prefetcher_round_trip():
    dev = torch.device("cuda", 0)
    pf = data.HostPrefetcher(dev)
    a = [torch.randn(3, 64, 64).pin_memory() for _ in range(4)]
    pf.feed(a[0])
    for i in range(4):
        (d,) = pf.next()
        if i + 1 < 4:
            pf.feed(a[i + 1])
        assert torch.equal(d.cpu(), a[i])


def test_scale_selection_matches_reference(golden):
    """Resize.get_size (transforms.py:42-64) over portrait / landscape / square / panoramic shapes x the multi-scale set."""
    G = golden("input_pipeline.npz")
    for w, h, s, mx, oh, ow in G["get_size"].tolist():
        assert data.resize_size((w, h), s, mx) == (oh, ow), (w, h, s, mx)
    assert data.resize_size((500, 375), (480, 576, 688), 2000, pick=576) == (576, 768)


def test_proposal_file_and_example_pipeline_match_reference(golden, tmp_path):
    """Pickled proposal file -> per-image lookup + filter (voc.py:87-111), then Resize -> flip -> ToTensor -> Normalize
    (build_transforms) on a PIL image: bit-exact against the reference's own dataset / transform classes."""
    import pickle
    from PIL import Image
    G = golden("input_pipeline.npz")
    prop = {"boxes": [G["prop_boxes_%d" % k] for k in range(3)], "indexes": [int(i) for i in G["prop_ids"]]}
    path = tmp_path / "proposals.pkl"
    with open(path, "wb") as f:
        pickle.dump(prop, f)
    pf = data.ProposalFile(str(path))
    assert len(pf) == 3
    rois = pf.rois(42, 100, 75)
    assert np.array_equal(rois.numpy(), G["lookup_42"])
    legacy = data.ProposalFile({"boxes": prop["boxes"], "ids": prop["indexes"]})          # the 'ids' spelling
    assert np.array_equal(legacy.rois(42, 100, 75).numpy(), G["lookup_42"])
    img = Image.fromarray(G["img_u8"], "RGB")
    tgt = torch.tensor([[10., 12., 60., 50.]])
    for tag, flip in (("flip", True), ("noflip", False)):
        im, ro, tg, size = data.prepare_example(img, rois, 120, 400, flip, target_boxes=tgt)
        assert list(size) == G["ex_%s_size" % tag].tolist()
        assert np.array_equal(im.numpy(), G["ex_%s_img" % tag])
        assert np.array_equal(ro.numpy(), G["ex_%s_rois" % tag])
        assert np.array_equal(tg.numpy(), G["ex_%s_tgt" % tag])
