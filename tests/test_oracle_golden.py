"""CPU: the oracle restatement against the committed golden vectors produced by the unmodified reference
(tests/golden/make_golden.py)."""
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
