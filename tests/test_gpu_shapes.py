"""GPU: shape sweep of the head (edge geometry: maps smaller than a tile, one strip, odd sizes, single pixel rows of
tiles, single class, D = 64) against the CPU oracle.  Every case runs under a timeout: a wrong barrier protocol hangs."""
import pytest
import torch

from _util import rel_to_max, TOL
from oracle import head_oracle as ho

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]

SHAPES = [
    # (B, C, D, H, W, simple, inverse)
    (1, 1, 64, 2, 2, False, True),        # smallest legal map (the reference divides by W-1, H-1)
    (1, 2, 64, 3, 17, True, False),       # 3 strips (odd), 3 rows
    (2, 1, 128, 9, 8, False, False),      # exactly one strip
    (1, 3, 64, 33, 7, False, True),       # taller than one row tile, narrower than a strip
    (1, 2, 1024, 16, 16, True, True),
    (1, 1, 64, 65, 31, False, True),      # 3 row tiles, 4 strips (last one 7 wide)
    (3, 2, 64, 12, 41, False, True),      # 6 strips
    (1, 5, 192, 25, 25, True, False),
]


@pytest.mark.parametrize("B,C,D,H,W,simple,inverse", SHAPES)
def test_head_shape_sweep(B, C, D, H, W, simple, inverse):
    from os2d_b200 import head as bh
    from os2d_b200.structures import FeatureMapSize
    g = torch.Generator().manual_seed(H * 100 + W)
    sizes = [(15, 15), (7, 22), (18, 5), (1, 9), (30, 30)]
    cms = [(torch.randn(1, D, *sizes[i % len(sizes)], generator=g) * 0.5 + 0.2).relu() for i in range(C)]
    fm = (torch.randn(B, D, H, W, generator=g) * 0.5 + 0.2).relu()
    P = 4 if simple else 6
    tn = ho.random_transform_net(P, seed=3, spread=0.004)
    hc = bh.build_os2d_head_creator(simple, True, inverse, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head([c.cuda() for c in cms])
        loc, rec, _, corners = head(fm.cuda())
    torch.cuda.synchronize()
    cf = ho.prepare_class_features(cms)
    oloc, osc, ocor = ho.head_forward(cf, fm, tn, simple, inverse)
    assert rel_to_max(head.class_feature_maps.cpu(), cf) < 1e-5
    assert rel_to_max(rec.cpu(), osc) < TOL
    assert rel_to_max(corners.cpu(), ocor) < TOL
    # loc is ~0 where the regressed transform is the identity: compare against the anchor-relative scale as well
    assert float((loc.cpu() - oloc).abs().max()) < TOL * max(float(oloc.abs().max()), 1.0)


def test_bad_arguments_fail_loudly():
    from os2d_b200 import head as bh
    from os2d_b200 import _cabi
    from os2d_b200.structures import FeatureMapSize
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head([torch.rand(1, 96, 15, 15).cuda()])          # D = 96 is not a multiple of 64
        with pytest.raises(_cabi.Os2dB200Error):
            head(torch.rand(1, 96, 8, 8).cuda())
        head = hc.create_os2d_head([torch.rand(1, 64, 15, 15).cuda()])
        with pytest.raises(ValueError):
            head(torch.rand(1, 64, 1, 8).cuda())                                 # 1-pixel-high map: reference yields NaN
        with pytest.raises(AssertionError):
            head(torch.rand(1, 128, 8, 8).cuda())                                # feature dimensionality mismatch


def test_workspace_is_bounded_in_bytes():
    """the per-call workspace (z / raw / h1 / h2 / params, ~1.47 KB per plane and location) is bounded in BYTES:
    many class views on a large pyramid level are processed in class chunks, bit-identically to one big chunk."""
    import torch
    from os2d_b200 import head as bh
    from os2d_b200.structures import FeatureMapSize
    from oracle import head_oracle as ho
    from _util import synth_inputs
    tn = ho.random_transform_net(6, seed=2, spread=0.005)
    cms, fm = synth_inputs(5, 2, 40, 36, [(15, 15)] * 9)
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head([c.cuda() for c in cms])
        ref = head(fm.cuda())
        N, B, P = 40 * 36, 2, 6
        per_plane = N * (30 * 16 + 225 * 2 + 256 + 256 + 4 * P)
        assert head._classes_per_chunk(B, N, P, fm.cuda().device) == 9 or head._classes_per_chunk(B, N, P, fm.cuda().device) >= 9
        head.workspace_bytes = 7 * per_plane                       # room for 7 planes = 3 classes of a 2-image batch
        head._cmax_cache.clear()
        assert head._classes_per_chunk(B, N, P, fm.cuda().device) == 3
        out = head(fm.cuda())
        head.workspace_bytes = 1                                   # degenerate budget: one class at a time, never zero
        head._cmax_cache.clear()
        assert head._classes_per_chunk(B, N, P, fm.cuda().device) == 1
        out1 = head(fm.cuda())
    for a, b, c in zip(ref, out, out1):
        assert torch.equal(a, b) and torch.equal(a, c)
    # what the default budget does for the case of the advice: 1000 views on a 150x150 level
    free, _ = torch.cuda.mem_get_info()
    big = head._classes_per_chunk.__func__(type("H", (), {"max_planes_per_call": 4096, "workspace_bytes": None, "_cmax_cache": {}})(),
                                           1, 150 * 150, 6, fm.cuda().device)
    assert big * 150 * 150 * 1466 <= free
