"""Pins the CPU oracle (oracle/) against golden vectors produced by the reference itself
(oracle/gen_golden.py) -- SURVEY.md section 8c.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.helpers import case_tensors, grid_sim


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_roi_pool_bit_exact(golden, tag):
    G = golden("roi_pool.npz")
    ph, pw = [int(x) for x in G[tag + "_pooled"]]
    out, arg = orc.roi_pool_forward(G[tag + "_feat"], G[tag + "_rois"], 0.125, ph, pw)
    assert np.array_equal(out, G[tag + "_out"])
    assert np.array_equal(arg, G[tag + "_argmax"])
    B, C, H, W = G[tag + "_feat"].shape
    gi = orc.roi_pool_backward(G[tag + "_grad_out"], arg, G[tag + "_rois"], B, C, H, W)
    np.testing.assert_allclose(gi, G[tag + "_grad_in"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_roi_align(golden, tag):
    G = golden("roi_align.npz")
    sr = int(G[tag + "_sr"])
    out = orc.roi_align_forward(G[tag + "_feat"], G[tag + "_rois"], 0.125, 7, 7, sr)
    np.testing.assert_allclose(out, G[tag + "_out"], rtol=1e-5, atol=1e-5)   # fp32 sum order differs
    B, C, H, W = G[tag + "_feat"].shape
    gi = orc.roi_align_backward(G[tag + "_grad_out"], G[tag + "_rois"], 0.125, B, C, H, W, sr)
    np.testing.assert_allclose(gi, G[tag + "_grad_in"], rtol=1e-4, atol=1e-5)


def test_iou_nms_bit_exact(golden):
    G = golden("boxes.npz")
    assert np.array_equal(orc.box_iou(G["P"], G["Q"], True), G["iou"])
    for t in range(3):
        assert np.array_equal(orc.cal_iou(G["P"], int(G["cal_iou_m_%d" % t]), 0.5), G["cal_iou_%d" % t])
        thr = float(G["easy_nms_thr_%d" % t])
        assert np.array_equal(orc.easy_nms(G["P"], G["cluster"], G["scores"], thr), G["easy_nms_%d" % t])
        assert np.array_equal(orc.nms_legacy(G["P"], G["scores_u"], thr, ge=True), G["legacy_nms_%d" % t])
    assert np.array_equal(orc.nms_tv(G["P"], G["scores"], 0.3), G["tv_nms_full"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_supcon(golden, tag):
    G = golden("supcon.npz")
    f = torch.from_numpy(G[tag + "_feats"]).requires_grad_(True)
    loss = orc.supcon_v2(f, torch.from_numpy(G[tag + "_labels"]), torch.from_numpy(G[tag + "_w"]), 0.2)
    loss.backward()
    assert abs(float(loss) - float(G[tag + "_loss"])) <= 1e-6 * abs(float(G[tag + "_loss"]))
    np.testing.assert_allclose(f.grad.numpy(), G[tag + "_grad"], rtol=1e-4, atol=1e-9)


def oracle_loss_case(G, tag):
    t = case_tensors(G, tag)
    rng = orc.StochasticSource(t["seed"])
    pooled_l = t["pooled"].split(t["sizes"])

    def embed_aug(b, c, I, kind):
        x = pooled_l[b][I]
        if kind == "drop":
            x = orc.dropblock(x, rng.dropblock_centres(x.shape[0], 1), 1)
        else:
            x = rng.noise(x.shape) * x + x
        h = torch.relu(torch.nn.functional.linear(x.reshape(x.shape[0], -1), t["fe_w"], t["fe_b"]))
        return grid_sim(h, t["ms_w"], t["ms_b"])

    losses, tr = orc.roi_reg_loss(t["cls"], t["det"], [t["ref0"], t["ref1"], t["ref2"]],
                                  [t["bb0"], t["bb1"], t["bb2"]], t["simf"], t["boxes"], t["labels"],
                                  embed_aug, thres=0.5, nms=0.1, lmda=0.03, temp=0.2, return_trace=True)
    return t, losses, tr


@pytest.mark.parametrize("tag", ["a", "b"])
def test_roi_reg_loss_matches_reference(golden, tag):
    """Object discovery index sets bit-exact, all 8 losses and input gradients within 1e-5 rel."""
    G = golden("roi_reg_loss.npz")
    t, losses, tr = oracle_loss_case(G, tag)
    for b in range(len(t["sizes"])):
        for i in range(3):
            for c in range(20):
                key = "%s_inst_%d_%d_%d" % (tag, b, i, c)
                got = tr["inst"][b][i][c]
                if key in G:
                    assert np.array_equal(got, G[key]), key
                else:
                    assert got.size == 0, key
            pl, lw, rt = tr["pseudo"][(b, i)]
            assert np.array_equal(pl.numpy(), G["%s_pl_%d_%d" % (tag, b, i)])
            np.testing.assert_allclose(lw.numpy(), G["%s_lw_%d_%d" % (tag, b, i)], rtol=1e-6)
            np.testing.assert_allclose(rt.numpy(), G["%s_rt_%d_%d" % (tag, b, i)], rtol=1e-5, atol=1e-6)
    for k, v in losses.items():
        ref = float(G["%s_loss_%s" % (tag, k)])
        assert abs(float(v) - ref) <= 1e-5 * abs(ref) + 1e-9, (k, float(v), ref)
    sum(losses.values()).backward()
    for k in ("cls", "det", "simf", "pooled", "ref0", "ref1", "ref2", "bb0", "bb1", "bb2"):
        np.testing.assert_allclose(t[k].grad.numpy(), G["%s_g_%s" % (tag, k)], rtol=2e-4, atol=1e-8, err_msg=k)


def test_model_cfg1_matches_reference(golden):
    """BASELINE.json configs[0]: 1x600x600, 256 proposals, VGG16 random init, full train-mode
    forward through the reference GeneralizedRCNN vs the oracle's model_forward."""
    G = golden("model_cfg1.npz")
    sd = orc.synth_state_dict(21, seed=0)
    images, boxes, labels = orc.synth_batch(1, 256, 600, 600, 21, seed=1234)
    with torch.no_grad():
        feat = orc.vgg16_forward(images, sd)
    assert list(feat.shape) == list(G["feat_shape"])
    np.testing.assert_allclose(feat[0, ::37, ::5, ::7].numpy(), G["feat_sample"], rtol=1e-4, atol=1e-4)
    losses = orc.model_forward(sd, images, boxes, labels, orc.StochasticSource(4242))
    for k, v in losses.items():
        ref = float(G["loss_" + k])
        assert abs(float(v) - ref) <= 1e-4 * abs(ref), (k, float(v), ref)
