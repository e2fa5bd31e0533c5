"""GPU: BASELINE.json-sized workloads checked through size-independent properties (the oracle would take minutes at
these sizes): a class subset against the oracle, class-chunking invariance, plane independence across the batch."""
import pytest
import torch

from _util import rel_to_max, TOL
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


def _setup(C, B, S, simple, inverse, seed):
    from os2d_b200 import head as bh
    from os2d_b200.structures import FeatureMapSize
    g = torch.Generator().manual_seed(seed)
    cms = (torch.randn(C, 1024, 15, 15, generator=g) * 0.5 + 0.2).relu()
    fm = (torch.randn(B, 1024, S, S, generator=g) * 0.5 + 0.2).relu()
    P = 4 if simple else 6
    tn = ho.random_transform_net(P, seed=1, spread=0.005)
    hc = bh.build_os2d_head_creator(simple, True, inverse, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    return hc, cms, fm, tn


def test_config2_1280px_100_classes_v2():
    """configs[1]: 1280 px (80x80 map), 100 classes, V2 head."""
    hc, cms, fm, tn = _setup(100, 1, 80, False, True, 0)
    with torch.no_grad():
        head = hc.create_os2d_head([cms[i:i + 1].cuda() for i in range(100)])
        loc, rec, _, corners = head(fm.cuda())
        # class chunking (bounded workspace) must not change a single bit
        head.max_planes_per_call = 32
        loc2, rec2, _, corners2 = head(fm.cuda())
    torch.cuda.synchronize()
    assert loc.shape == (1, 100, 4, 80, 80) and rec.shape == (1, 100, 1, 80, 80) and corners.shape == (1, 100, 8, 80, 80)
    assert torch.equal(loc, loc2) and torch.equal(rec, rec2) and torch.equal(corners, corners2)
    assert bool(torch.isfinite(loc).all()) and float(rec.min()) >= -1 and float(rec.max()) <= 1
    sub = [0, 57, 99]
    cf = ho.prepare_class_features([cms[i:i + 1] for i in sub])
    oloc, osc, ocor = ho.head_forward(cf, fm, tn, False, True, class_chunk=3)
    assert rel_to_max(rec[:, sub].cpu(), osc) < TOL
    assert rel_to_max(loc[:, sub].cpu(), oloc) < TOL
    assert rel_to_max(corners[:, sub].cpu(), ocor) < TOL


def test_config5_batch8_960px_and_v1_head():
    """configs[4] shape (8 x 960 px -> 60x60 maps) with the V1 simplified-affine head of configs[3]; fewer classes."""
    hc, cms, fm, tn = _setup(12, 8, 60, True, False, 3)
    with torch.no_grad():
        head = hc.create_os2d_head([cms[i:i + 1].cuda() for i in range(12)])
        loc, rec, _, corners = head(fm.cuda())
        l1, r1, _, c1 = head(fm[5:6].cuda())
    torch.cuda.synchronize()
    assert torch.equal(l1[0], loc[5]) and torch.equal(r1[0], rec[5]) and torch.equal(c1[0], corners[5])
    cf = ho.prepare_class_features([cms[i:i + 1] for i in (3, 11)])
    oloc, osc, ocor = ho.head_forward(cf, fm[2:3], tn, True, False, class_chunk=2)
    assert rel_to_max(rec[2:3, [3, 11]].cpu(), osc) < TOL
    assert rel_to_max(loc[2:3, [3, 11]].cpu(), oloc) < TOL
    assert rel_to_max(corners[2:3, [3, 11]].cpu(), ocor) < TOL


@pytest.mark.parametrize("side", [96, 112, 128])
def test_config4_large_pyramid_levels_v1_head(side):
    """configs[3]: the three largest levels of the 7-scale pyramid (1536 / 1792 / 2048 px -> 96^2 / 112^2 / 128^2 maps, up to
    16 384 locations) with the V1 head (simplified affine P = 4, no inverse): a class subset against the oracle, class-chunk
    invariance bit for bit, non-square variant of the level."""
    C = 6
    hc, cms, fm, tn = _setup(C, 1, side, True, False, 40 + side)
    with torch.no_grad():
        head = hc.create_os2d_head([cms[i:i + 1].cuda() for i in range(C)])
        loc, rec, _, corners = head(fm.cuda())
        head.max_planes_per_call = 4
        head._cmax_cache.clear()
        loc2, rec2, _, corners2 = head(fm.cuda())
        # non-square level of the same scale (1280 x 960 image at this pyramid scale)
        hh = side * 3 // 4
        l3, r3, _, c3 = head(fm[:, :, :hh].contiguous().cuda())
    torch.cuda.synchronize()
    assert loc.shape == (1, C, 4, side, side)
    assert torch.equal(loc, loc2) and torch.equal(rec, rec2) and torch.equal(corners, corners2)
    sub = [1, 4]
    cf = ho.prepare_class_features([cms[i:i + 1] for i in sub])
    oloc, osc, ocor = ho.head_forward(cf, fm, tn, True, False, class_chunk=2)
    assert rel_to_max(rec[:, sub].cpu(), osc) < TOL
    assert rel_to_max(loc[:, sub].cpu(), oloc) < TOL
    assert rel_to_max(corners[:, sub].cpu(), ocor) < TOL
    oloc, osc, ocor = ho.head_forward(cf[:1], fm[:, :, :hh].contiguous(), tn, True, False, class_chunk=1)
    assert rel_to_max(r3[:, sub[:1]].cpu(), osc) < TOL
    assert rel_to_max(l3[:, sub[:1]].cpu(), oloc) < TOL
    assert rel_to_max(c3[:, sub[:1]].cpu(), ocor) < TOL


def test_config4_seven_level_run_head_to_detections():
    """configs[3] end to end on the device: 7-level pyramid of feature maps (40..128) -> V1 head per level -> decode_pyramid
    with the reference defaults (score threshold -inf, IoU 0.3: 52 740 candidates per class, chunked NMS to the fixpoint).
    The fused two-launch post-processing equals the staged path bit for bit on the head's real outputs."""
    from os2d_b200.box_coder import Os2dBoxCoder, make_resize_transform
    from os2d_b200.structures import FeatureMapSize
    sides = [40, 50, 64, 80, 96, 112, 128]
    C = 4
    hc, cms, _, tn = _setup(C, 1, 8, True, False, 77)
    g = torch.Generator().manual_seed(78)
    coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level,
                         lambda sz: FeatureMapSize(w=-(-sz.w // 16), h=-(-sz.h // 16)))
    tgt = FeatureMapSize(w=1280, h=1280)
    loc_p, cls_p, cor_p = [], [], []
    with torch.no_grad():
        head = hc.create_os2d_head([cms[i:i + 1].cuda() for i in range(C)])
        for s in sides:
            fm = (torch.randn(1, 1024, s, s, generator=g) * 0.5 + 0.2).relu().cuda()
            loc, rec, _, corners = head(fm)
            loc_p.append(loc[0].view(C, 4, s * s))
            cls_p.append(rec[0].view(C, s * s))
            cor_p.append(corners[0].view(C, 8, s * s))
        kw = dict(nms_score_threshold=float("-inf"), nms_iou_threshold=0.3,
                  inverse_box_transforms=[make_resize_transform(tgt)] * 7, transform_corners_pyramid=cor_p)
        sizes = [FeatureMapSize(w=16 * s, h=16 * s) for s in sides]
        dets = coder.decode_pyramid(loc_p, cls_p, sizes, [3, 0, 2, 1], **kw)
        staged = coder.decode_pyramid_staged(loc_p, cls_p, sizes, [3, 0, 2, 1], **kw)
    assert len(dets) > 0 and dets.image_size == tgt
    for f in ("scores", "labels", "transform_corners"):
        assert torch.equal(dets.get_field(f), staged.get_field(f))
    assert torch.equal(dets.bbox_xyxy, staged.bbox_xyxy)
    assert torch.equal(dets.get_field("default_boxes").bbox_xyxy, staged.get_field("default_boxes").bbox_xyxy)


def test_pyramid_decode_seven_levels_chunked():
    """configs[3] post-processing shape: 7 pyramid levels (52 740 anchors per class => chunked NMS).  Decoded boxes are
    compared with the oracle within tolerance; the NMS itself is checked on identical boxes (bit-exact) so that an ulp of
    expf cannot flip a decision."""
    import numpy as np
    from os2d_b200.box_coder import Os2dBoxCoder, BoxGridGenerator, nms, make_resize_transform
    from os2d_b200.structures import FeatureMapSize, BoxList
    from oracle import postproc_oracle as po
    sides = [40, 50, 64, 80, 96, 112, 128]
    g = torch.Generator().manual_seed(4)
    C = 2
    loc_pyr = [torch.randn(C, 4, s * s, generator=g) for s in sides]
    cls_pyr = [torch.rand(C, s * s, generator=g) for s in sides]
    gen = BoxGridGenerator(FeatureMapSize(w=240, h=240), FeatureMapSize(w=16, h=16))
    coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, gen, lambda sz: FeatureMapSize(w=-(-sz.w // 16), h=-(-sz.h // 16)))
    tgt = FeatureMapSize(w=1280, h=1280)
    dets = coder.decode_pyramid([t.cuda() for t in loc_pyr], [t.cuda() for t in cls_pyr],
                                [FeatureMapSize(w=16 * s, h=16 * s) for s in sides], [0, 1], nms_score_threshold=float("-inf"),
                                nms_iou_threshold=0.3, inverse_box_transforms=[make_resize_transform(tgt)] * 7)
    assert len(dets) > 0 and dets.image_size == tgt
    sc = dets.get_field("scores")
    for lab in (0, 1):
        s_l = sc[dets.get_field("labels") == lab]
        assert bool((s_l[:-1] >= s_l[1:]).all())             # per label: score-descending
    # NMS of the GPU-decoded candidates, recomputed by the oracle on the very same boxes
    for lab in (0, 1):
        boxes, scores = [], []
        for s, loc, cls in zip(sides, loc_pyr, cls_pyr):
            anc = po.anchors_xyxy(s, s)
            b = po.decode_boxes(np.ascontiguousarray(loc[lab].numpy().T), anc)
            b[:, 0::2] = np.clip(b[:, 0::2], 0, 16 * s)
            b[:, 1::2] = np.clip(b[:, 1::2], 0, 16 * s)
            ok = ~((b[:, 3] <= b[:, 1]) | (b[:, 2] <= b[:, 0]))
            boxes.append((b[ok] * np.float32(1280.0 / (16 * s))).astype(np.float32))
            scores.append(cls[lab].numpy()[ok])
        boxes, scores = np.concatenate(boxes), np.concatenate(scores)
        assert boxes.shape[0] > 30000                          # chunked path (> 3 chunks of 10000)
        bl = BoxList(torch.from_numpy(boxes).cuda(), tgt)
        bl.add_field("scores", torch.from_numpy(scores).cuda())
        keep = nms(bl, 0.3).cpu().numpy()
        np.testing.assert_array_equal(keep, po.chunked_nms(boxes, scores, 0.3))
        # and the end-to-end detections of this label agree with the oracle up to decisions an ulp could flip
        n_dets = int((dets.get_field("labels") == lab).sum())
        assert abs(n_dets - keep.shape[0]) <= max(3, keep.shape[0] // 100)


@pytest.mark.parametrize("corr_sms", [16, 24])
def test_stage_concurrent_corr_conv1_is_bit_identical(corr_sms):
    """K1 running NEXT TO conv1 (os2d_correlate_conv1_concurrent: K1 on `corr_sms` SMs releases each finished plane, conv1 on the
    other SMs acquires it before its first TMA load) must not change a bit relative to the two kernels one after the other."""
    hc, cms, fm, tn = _setup(30, 1, 80, False, True, 11)
    with torch.no_grad():
        head = hc.create_os2d_head([cms[i:i + 1].cuda() for i in range(30)])
        head.concurrent_corr_sms = 0
        loc, rec, _, corners = head(fm.cuda())
        head.concurrent_corr_sms = corr_sms
        for _ in range(3):                      # repeated: flag reset and stream joins between calls
            loc2, rec2, _, corners2 = head(fm.cuda())
            assert torch.equal(loc, loc2) and torch.equal(rec, rec2) and torch.equal(corners, corners2)
        head.max_planes_per_call = 12           # chunked: several concurrent launches per call
        head._cmax_cache.clear()
        loc3, rec3, _, corners3 = head(fm.cuda())
    torch.cuda.synchronize()
    assert torch.equal(loc, loc3) and torch.equal(rec, rec3) and torch.equal(corners, corners3)
