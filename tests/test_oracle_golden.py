"""CPU: the oracle restatement against the committed golden vectors produced by the unmodified reference
(tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from oracle import postproc_oracle as po
from _util import VARIANTS, load_head_golden, rel_to_max, GOLDEN


@pytest.mark.parametrize("name,simple,inverse", VARIANTS)
def test_head_oracle_matches_reference_golden(name, simple, inverse):
    d, cms, fm, tn = load_head_golden(name)
    cf = ho.prepare_class_features(cms)
    assert rel_to_max(cf, d["class_features"]) < 1e-6
    loc, score, corners = ho.head_forward(cf, fm, tn, simple, inverse)
    # the reference inverts with LU (torch.inverse), the oracle in closed form: allow a few fp32 ulps of the range
    assert rel_to_max(score, d["score"]) < 5e-6
    assert rel_to_max(loc, d["loc"]) < 5e-5
    assert rel_to_max(corners, d["corners"]) < 5e-6


def test_head_oracle_emulated_fp16_within_parity_bar():
    """The fp16-operand plan of the CUDA path (emulated on the CPU) stays inside the 1e-3 bar on the golden inputs."""
    d, cms, fm, tn = load_head_golden("affine_inverse")
    cf = ho.prepare_class_features(cms)
    loc, score, corners = ho.head_forward(cf, fm, tn, False, True, emulate=True)
    assert rel_to_max(score, d["score"]) < 1e-3
    assert rel_to_max(loc, d["loc"]) < 2e-3      # pessimistic plan (no centring / hi-lo rows), see DESIGN.md
    assert rel_to_max(corners, d["corners"]) < 1e-3


def test_decode_pyramid_oracle_matches_reference_golden():
    z = np.load(GOLDEN + "/decode_pyramid.npz")
    L = len(z["img_sizes"])
    res = po.decode_pyramid([z["loc_%d" % l] for l in range(L)], [z["cls_%d" % l] for l in range(L)],
                            [tuple(s) for s in z["img_sizes"]], [tuple(s) for s in z["fm_sizes"]],
                            list(z["class_ids"]), float(z["score_thr"]), float(z["iou_thr"]),
                            target_size=tuple(z["target"]), corners_pyr=[z["corners_%d" % l] for l in range(L)])
    assert res["boxes"].shape == z["boxes"].shape
    np.testing.assert_array_equal(res["labels"], z["labels"])
    np.testing.assert_array_equal(res["scores"], z["scores"])           # bit-exact selection and order
    np.testing.assert_allclose(res["boxes"], z["boxes"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(res["default_boxes"], z["default_boxes"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(res["transform_corners"], z["transform_corners"], rtol=0, atol=1e-3)


def test_chunked_nms_oracle_matches_reference_golden():
    z = np.load(GOLDEN + "/nms_chunked.npz")
    n = int(z["n"])
    g = torch.Generator().manual_seed(int(z["seed"]))
    ctr = torch.rand(n, 2, generator=g) * 600
    wh = torch.rand(n, 2, generator=g) * 120 + 20
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=1)
    scores = (torch.rand(n, generator=g) * 64).round() / 64
    assert abs(float(boxes.double().sum()) - float(z["box_checksum"])) < 1e-6
    keep = po.chunked_nms(boxes.numpy(), scores.numpy(), 0.3)
    np.testing.assert_array_equal(keep, z["keep"])                       # bit-exact kept set and order


def test_greedy_nms_edge_cases():
    assert po.greedy_nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.3).shape == (0,)
    b = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [5, 5, 5, 5]], np.float32)
    s = np.array([0.5, 0.5, 0.1, 0.9], np.float32)
    # identical boxes with tied scores: the lower index wins; the degenerate box (area 0) never suppresses
    np.testing.assert_array_equal(po.greedy_nms(b, s, 0.3), [3, 0, 2])


@pytest.mark.parametrize("name,simple", [("affine", False), ("simple", True)])
def test_singular_transforms_follow_reference_failure_handling(name, simple):
    """head.py:123-146: a chunk of <= 65535 matrices containing an exactly singular one is regularised as a whole
    (+1e-5 on the diagonal of the 3x3) - against outputs of the reference's own aligner on CPU."""
    from _util import singular_params
    gold = np.load(os.path.join(GOLDEN, "theta_singular.npz"))
    P = 4 if simple else 6

    def theta_rows(p):
        return torch.stack(ho.theta_from_params(p, simple, True), dim=-1).reshape(-1, 6)

    # one chunk: every matrix regularised; the singular ones come out ~1e5 large (fp32 LU vs fp64 closed form: 1e-2)
    sing = [tuple(r) for r in gold["small_{}_singular".format(name)].tolist()]
    p = singular_params(21, 2, P, 5, 7, sing)
    th = theta_rows(p)
    ref = torch.from_numpy(gold["small_{}_theta".format(name)]).reshape(-1, 6)
    assert bool(torch.isfinite(th).all())
    srows = [(n * 5 + y) * 7 + x for (n, y, x) in sing]
    mask = torch.ones(th.shape[0], dtype=torch.bool)
    mask[srows] = False
    assert ((th[mask] - ref[mask]).abs() <= 2e-5 * ref[mask].abs().clamp_min(1.0)).all()
    assert ((th[~mask] - ref[~mask]).abs() <= 1e-2 * ref[~mask].abs().max(dim=1, keepdim=True).values).all()
    # the regularisation is visible on the regular matrices: without it the difference is ~1e-5, well above 2e-6
    plain = torch.stack([x.reshape(-1) for x in _plain_inverse(p, simple)], dim=1)
    assert ((plain[mask] - ref[mask]).abs().max() > 5e-6)

    # two chunks (65535 + 349), singular matrices only in the second: the first chunk is NOT regularised
    NB, H, W = 2, 182, 181
    sing2 = [tuple(r) for r in gold["chunks_{}_singular".format(name)].tolist()]
    p2 = singular_params(22, NB, P, H, W, sing2)
    th2 = theta_rows(p2)
    head, tail = torch.from_numpy(gold["chunks_{}_head".format(name)]), torch.from_numpy(gold["chunks_{}_tail".format(name)])
    plain2 = torch.stack([x.reshape(-1) for x in _plain_inverse(p2, simple)], dim=1)
    assert torch.equal(th2[:65535], plain2[:65535])                       # untouched chunk = plain closed form
    assert ((th2[:256] - head).abs() <= 3e-6 * head.abs().clamp_min(1.0)).all()
    t_or = th2[65535 - 128:]
    srows2 = [(n * H + y) * W + x - (65535 - 128) for (n, y, x) in sing2]
    m2 = torch.ones(t_or.shape[0], dtype=torch.bool)
    m2[srows2] = False
    assert ((t_or[m2] - tail[m2]).abs() <= 2e-5 * tail[m2].abs().clamp_min(1.0)).all()
    assert bool(torch.isfinite(t_or).all())
    assert abs(float(th2[:65535].double().sum()) - float(gold["chunks_{}_sum0".format(name)])) < 1e-3 * 65535 * 1e-3


def _plain_inverse(p, simple):
    if simple:
        z = torch.zeros_like(p[:, 0])
        a, b, tx, c, d, ty = p[:, 0], z, p[:, 1], z, p[:, 2], p[:, 3]
    else:
        a, b, tx, c, d, ty = (p[:, i] for i in range(6))
    det = a * d - b * c
    ia, ib, ic, id_ = d / det, -b / det, -c / det, a / det
    return ia, ib, -(ia * tx + ib * ty), ic, id_, -(ic * tx + id_ * ty)


def _voc_arrays(data):
    """voc_eval.py:28-31: predictions are resized to the ground-truth image size first (BoxList.resize,
    bounding_box.py:138-163: one fp32 multiply when both ratios agree, per axis otherwise)."""
    pb, pl, ps, gb, gl, gd = [], [], [], [], [], []
    for (b, l, s, psize, gt, gtl, gtd, gsize) in data:
        rw, rh = float(gsize[0]) / psize[0], float(gsize[1]) / psize[1]
        scaled = b * rw if rw == rh else b * torch.tensor([rw, rh, rw, rh], dtype=torch.float32)
        pb.append(scaled.numpy()); pl.append(l.numpy()); ps.append(s.numpy())
        gb.append(gt.numpy()); gl.append(gtl.numpy()); gd.append(gtd.numpy())
    return pb, pl, ps, gb, gl, gd


@pytest.mark.parametrize("tag,seed,quant", [("distinct", 91, 0), ("ties", 92, 16)])
def test_voc_oracle_matches_reference_golden(tag, seed, quant):
    from oracle import voc_oracle as vo
    from _util import voc_inputs
    gold = np.load(os.path.join(GOLDEN, "voc_eval.npz"))
    arrays = _voc_arrays(voc_inputs(seed, quant=quant))
    for thr in (0.5, 0.3):
        for m07 in (False, True):
            r = vo.eval_detection_voc(*arrays, iou_thresh=thr, use_07_metric=m07)
            key = "{}_thr{}_{}".format(tag, int(thr * 10), "07" if m07 else "area")
            np.testing.assert_array_equal(r["n_pos"], gold[key + "_n_pos"])
            np.testing.assert_allclose(r["ap_per_class"], gold[key + "_ap_per_class"], rtol=0, atol=1e-12, equal_nan=True)
            np.testing.assert_allclose(r["recall_per_class"], gold[key + "_recall_per_class"], rtol=0, atol=1e-12, equal_nan=True)
            np.testing.assert_allclose([r["map"], r["map_weighted"], r["recall"], r["ap_joint_classes"]], gold[key + "_scalars"],
                                       rtol=0, atol=1e-12)
            if not m07 and thr == 0.5:
                np.testing.assert_allclose(r["prec"][4], gold[key + "_prec4"], rtol=0, atol=1e-15, equal_nan=True)
                np.testing.assert_allclose(r["rec"][4], gold[key + "_rec4"], rtol=0, atol=1e-15)
