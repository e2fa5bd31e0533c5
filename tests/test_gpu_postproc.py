"""GPU: decode + NMS kernels against the reference golden vectors and the CPU oracle (bit-exact selection)."""
import numpy as np
import pytest
import torch

from _util import GOLDEN
from oracle import postproc_oracle as po

pytestmark = pytest.mark.gpu


def _coder(nms_across=False):
    from os2d_b200.box_coder import Os2dBoxCoder, BoxGridGenerator
    from os2d_b200.structures import FeatureMapSize
    gen = BoxGridGenerator(FeatureMapSize(w=240, h=240), FeatureMapSize(w=16, h=16))
    return Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, gen, lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)),
                        do_nms_across_classes=nms_across)


def test_decode_pyramid_matches_reference_golden():
    from os2d_b200.structures import FeatureMapSize
    from os2d_b200.box_coder import make_resize_transform
    z = np.load(GOLDEN + "/decode_pyramid.npz")
    L = len(z["img_sizes"])
    tgt = FeatureMapSize(w=int(z["target"][0]), h=int(z["target"][1]))
    res = _coder().decode_pyramid([torch.from_numpy(z["loc_%d" % l]).cuda() for l in range(L)],
                                  [torch.from_numpy(z["cls_%d" % l]).cuda() for l in range(L)],
                                  [FeatureMapSize(w=int(w), h=int(h)) for (w, h) in z["img_sizes"]],
                                  [int(c) for c in z["class_ids"]], nms_score_threshold=float(z["score_thr"]),
                                  nms_iou_threshold=float(z["iou_thr"]),
                                  inverse_box_transforms=[make_resize_transform(tgt) for _ in range(L)],
                                  transform_corners_pyramid=[torch.from_numpy(z["corners_%d" % l]).cuda() for l in range(L)])
    assert len(res) == z["boxes"].shape[0]
    np.testing.assert_array_equal(res.get_field("labels").cpu().numpy(), z["labels"])
    np.testing.assert_array_equal(res.get_field("scores").cpu().numpy(), z["scores"])     # bit-exact selection + order
    np.testing.assert_allclose(res.bbox_xyxy.cpu().numpy(), z["boxes"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(res.get_field("default_boxes").bbox_xyxy.cpu().numpy(), z["default_boxes"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(res.get_field("transform_corners").cpu().numpy(), z["transform_corners"], rtol=0, atol=1e-3)
    assert res.image_size == tgt


def _rand_boxes(n, seed, quant=0):
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(n, 2, generator=g) * 600
    wh = torch.rand(n, 2, generator=g) * 120 + 20
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=1)
    scores = torch.rand(n, generator=g)
    if quant:
        scores = (scores * quant).round() / quant
    return boxes, scores


def test_chunked_nms_matches_reference_golden():
    """> 10000 candidates: the chunk-of-10000 / iterate-to-fixpoint semantics of bounding_box.py:344-387."""
    from os2d_b200.box_coder import nms
    from os2d_b200.structures import BoxList, FeatureMapSize
    z = np.load(GOLDEN + "/nms_chunked.npz")
    boxes, scores = _rand_boxes(int(z["n"]), int(z["seed"]), 64)
    assert abs(float(boxes.double().sum()) - float(z["box_checksum"])) < 1e-6
    bl = BoxList(boxes.cuda(), FeatureMapSize(w=800, h=800))
    bl.add_field("scores", scores.cuda())
    keep = nms(bl, 0.3)
    np.testing.assert_array_equal(keep.cpu().numpy(), z["keep"])


@pytest.mark.parametrize("n,quant", [(1, 0), (2, 0), (777, 0), (6400, 0), (5000, 8), (10000, 0)])
def test_single_chunk_nms_matches_oracle(n, quant):
    from os2d_b200.box_coder import nms
    from os2d_b200.structures import BoxList, FeatureMapSize
    boxes, scores = _rand_boxes(n, 1000 + n, quant)
    bl = BoxList(boxes.cuda(), FeatureMapSize(w=800, h=800))
    bl.add_field("scores", scores.cuda())
    keep = nms(bl, 0.3).cpu().numpy()
    np.testing.assert_array_equal(keep, po.chunked_nms(boxes.numpy(), scores.numpy(), 0.3))


def test_per_label_nms_and_empty_inputs():
    from os2d_b200.box_coder import nms
    from os2d_b200.structures import BoxList, FeatureMapSize
    boxes, scores = _rand_boxes(3000, 5)
    labels = torch.randint(0, 7, (3000,), generator=torch.Generator().manual_seed(1))
    bl = BoxList(boxes.cuda(), FeatureMapSize(w=800, h=800))
    bl.add_field("scores", scores.cuda())
    bl.add_field("labels", labels.cuda())
    keep = nms(bl, 0.3, do_separate_per_label=True).cpu().numpy()
    ref = []
    for l in sorted(set(labels.tolist())):
        ids = np.nonzero(labels.numpy() == l)[0]
        ref.append(ids[po.chunked_nms(boxes.numpy()[ids], scores.numpy()[ids], 0.3)])
    np.testing.assert_array_equal(keep, np.concatenate(ref))
    empty = BoxList(torch.zeros(0, 4).cuda(), FeatureMapSize(w=8, h=8))
    empty.add_field("scores", torch.zeros(0).cuda())
    assert nms(empty, 0.3).numel() == 0


@pytest.mark.parametrize("across", [False, True])
def test_decode_pyramid_matches_oracle_random(across):
    """Random loc / scores, two levels, duplicated class ids, score == threshold ties, all-invalid class."""
    from os2d_b200.structures import FeatureMapSize
    from os2d_b200.box_coder import make_resize_transform
    g = torch.Generator().manual_seed(12)
    img_sizes = [(320, 256), (640, 512)]
    fm_sizes = [(20, 16), (40, 32)]
    C = 5
    class_ids = [3, 9, 3, 4, 11]
    loc_pyr = [torch.randn(C, 4, w * h, generator=g) * 1.5 for (w, h) in fm_sizes]
    cls_pyr = [torch.rand(C, w * h, generator=g) for (w, h) in fm_sizes]
    cls_pyr[0][1, :50] = 0.5              # exactly the threshold: must be dropped (strict >)
    cls_pyr[0][4] = 0.1
    cls_pyr[1][4] = 0.2                   # class 11 has no candidate
    cor_pyr = [torch.randn(C, 8, w * h, generator=g) * 100 for (w, h) in fm_sizes]
    tgt = (1280, 1024)
    res = _coder(across).decode_pyramid([t.cuda() for t in loc_pyr], [t.cuda() for t in cls_pyr],
                                        [FeatureMapSize(w=w, h=h) for (w, h) in img_sizes], class_ids,
                                        nms_score_threshold=0.5, nms_iou_threshold=0.3,
                                        inverse_box_transforms=[make_resize_transform(FeatureMapSize(w=tgt[0], h=tgt[1]))] * 2,
                                        transform_corners_pyramid=[t.cuda() for t in cor_pyr])
    ref = po.decode_pyramid([t.numpy() for t in loc_pyr], [t.numpy() for t in cls_pyr], img_sizes, fm_sizes, class_ids,
                            0.5, 0.3, target_size=tgt, corners_pyr=[t.numpy() for t in cor_pyr], nms_across_classes=across)
    np.testing.assert_array_equal(res.get_field("labels").cpu().numpy(), ref["labels"])
    np.testing.assert_array_equal(res.get_field("scores").cpu().numpy(), ref["scores"])
    np.testing.assert_allclose(res.bbox_xyxy.cpu().numpy(), ref["boxes"], rtol=1e-5, atol=1e-2)
    np.testing.assert_allclose(res.get_field("transform_corners").cpu().numpy(), ref["transform_corners"], rtol=1e-6, atol=1e-3)
    assert 11 not in res.get_field("labels").tolist()


def _same_detections(a, b):
    assert len(a) == len(b)
    assert torch.equal(a.get_field("labels"), b.get_field("labels"))
    assert torch.equal(a.get_field("scores"), b.get_field("scores"))
    assert torch.equal(a.bbox_xyxy, b.bbox_xyxy)
    assert torch.equal(a.get_field("default_boxes").bbox_xyxy, b.get_field("default_boxes").bbox_xyxy)
    if a.has_field("transform_corners"):
        assert torch.equal(a.get_field("transform_corners"), b.get_field("transform_corners"))


@pytest.mark.parametrize("case", ["cfg2_all_anchors", "seven_levels_chunked", "merged_views_chunked", "quantised_ties"])
def test_fused_decode_nms_equals_staged_path(case):
    """The two-launch fused path (csrc/detect.cu) against the staged round-1 path (decode kernel + torch ordering + segment
    NMS kernel, itself pinned by the goldens above) at sizes the CPU oracle cannot reach: every field bit for bit."""
    from os2d_b200.structures import FeatureMapSize
    from os2d_b200.box_coder import make_resize_transform
    g = torch.Generator().manual_seed(321)
    if case == "cfg2_all_anchors":           # BASELINE configs[1]: 100 classes x 6400 anchors, every anchor a candidate
        sides, C, thr, ids, quant = [80], 100, float("-inf"), list(range(100)), 0
    elif case == "seven_levels_chunked":     # configs[3] pyramid: 52 740 anchors per class, > 10000 candidates per label
        sides, C, thr, ids, quant = [40, 50, 64, 80, 96, 112, 128], 6, 0.55, [5, 1, 9, 2, 7, 3], 0
    elif case == "merged_views_chunked":     # duplicated class ids: views merge into one label before NMS
        sides, C, thr, ids, quant = [64, 80], 6, 0.3, [4, 4, 8, 4, 8, 2], 0
    else:                                    # tied scores: stable order by candidate position
        sides, C, thr, ids, quant = [50, 64], 4, 0.4, [0, 1, 2, 3], 32
    loc_pyr = [(torch.randn(C, 4, s * s, generator=g) * 1.2).cuda() for s in sides]
    cls_pyr = [torch.rand(C, s * s, generator=g) for s in sides]
    if quant:
        cls_pyr = [(t * quant).round() / quant for t in cls_pyr]
    cls_pyr = [t.cuda() for t in cls_pyr]
    cor_pyr = [(torch.randn(C, 8, s * s, generator=g) * 100).cuda() for s in sides]
    sizes = [FeatureMapSize(w=16 * s, h=16 * s) for s in sides]
    inv = [make_resize_transform(FeatureMapSize(w=1280, h=1280)) for _ in sides]
    coder = _coder()
    kw = dict(nms_score_threshold=thr, nms_iou_threshold=0.3, inverse_box_transforms=inv, transform_corners_pyramid=cor_pyr)
    fused = coder.decode_pyramid(loc_pyr, cls_pyr, sizes, ids, **kw)
    staged = coder.decode_pyramid_staged(loc_pyr, cls_pyr, sizes, ids, **kw)
    assert len(fused) > 0
    _same_detections(fused, staged)
    again = coder.decode_pyramid(loc_pyr, cls_pyr, sizes, ids, **kw)     # workspace / counter reuse
    _same_detections(fused, again)


def test_fused_decode_nms_no_candidates_and_no_corners():
    from os2d_b200.structures import FeatureMapSize
    g = torch.Generator().manual_seed(4)
    loc = [torch.randn(3, 4, 400, generator=g).cuda()]
    cls = [torch.rand(3, 400, generator=g).cuda()]
    sizes = [FeatureMapSize(w=320, h=320)]
    none = _coder().decode_pyramid(loc, cls, sizes, [0, 1, 2], nms_score_threshold=2.0)
    assert len(none) == 0 and none.get_field("labels").numel() == 0
    some = _coder().decode_pyramid(loc, cls, sizes, [0, 1, 2], nms_score_threshold=0.5)
    _same_detections(some, _coder().decode_pyramid_staged(loc, cls, sizes, [0, 1, 2], nms_score_threshold=0.5))
    assert not some.has_field("transform_corners")


def test_decode_pyramid_async_handles_in_flight():
    """decode_pyramid_async: several calls submitted before any result is read (the pipelined use) give the synchronous results."""
    from os2d_b200.structures import FeatureMapSize
    g = torch.Generator().manual_seed(9)
    coder = _coder()
    sizes = [FeatureMapSize(w=640, h=480)]
    calls = []
    for k in range(4):
        loc = [torch.randn(5, 4, 40 * 30, generator=g).cuda()]
        cls = [torch.rand(5, 40 * 30, generator=g).cuda()]
        calls.append((loc, cls))
    sync = [coder.decode_pyramid(loc, cls, sizes, [2, 0, 1, 2, 4], nms_score_threshold=0.3) for loc, cls in calls]
    pend = [coder.decode_pyramid_async(loc, cls, sizes, [2, 0, 1, 2, 4], nms_score_threshold=0.3) for loc, cls in calls]
    for p, ref in zip(reversed(pend), reversed(sync)):                   # any order of completion
        got = p.result()
        assert got is p.result()                                         # idempotent
        _same_detections(got, ref)
