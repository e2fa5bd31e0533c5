"""GPU: the stand-alone public methods of the drop-in classes that take / return the large tensors of the reference
(TransformationNet.forward, Os2dAlignment.forward / prepare_transform_parameters_for_grid_sampler,
Os2dHead.resample_of_correlation_map_fast/_simple) against the CPU oracle."""
import pytest
import torch

from _util import rel_to_max, synth_inputs, TOL
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


def _setup(simple, inverse, B=2, H=14, W=19):
    from os2d_b200 import head as bh
    from os2d_b200.structures import FeatureMapSize
    P = 4 if simple else 6
    tn = ho.random_transform_net(P, seed=9, spread=0.005)
    cms, fm = synth_inputs(31, B, H, W, [(15, 15), (12, 18), (19, 11)], D=64)
    hc = bh.build_os2d_head_creator(simple, True, inverse, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    cf = ho.prepare_class_features(cms)
    corr = ho.correlate(cf, ho.l2_normalize(fm, 1e-5))          # [B*C,225,H,W] fp32, what the reference materialises
    return hc, tn, cms, fm, cf, corr, P


@pytest.mark.parametrize("simple,inverse", [(False, True), (True, False)])
def test_transformation_net_and_alignment_forward(simple, inverse):
    hc, tn, cms, fm, cf, corr, P = _setup(simple, inverse)
    p_ref, _ = ho.transform_net(corr, tn)
    with torch.no_grad():
        p = hc.aligner.parameter_regressor(corr.cuda())
        grids = hc.aligner(corr.cuda())
        theta = hc.aligner.prepare_transform_parameters_for_grid_sampler(p)
    assert p.shape == p_ref.shape
    assert rel_to_max(p.cpu(), p_ref) < TOL
    a, b, tx, c, d, ty = ho.theta_from_params(p_ref, simple, inverse)
    th_ref = torch.stack([a, b, tx, c, d, ty], dim=-1).reshape(-1, 2, 3)
    assert theta.shape == th_ref.shape and rel_to_max(theta.cpu(), th_ref) < TOL
    lin = torch.linspace(-1, 1, 15)
    gx = a[..., None, None] * lin[None, None, None, None, :] + b[..., None, None] * lin[None, None, None, :, None] + tx[..., None, None]
    gy = c[..., None, None] * lin[None, None, None, None, :] + d[..., None, None] * lin[None, None, None, :, None] + ty[..., None, None]
    g_ref = torch.stack([gx, gy], dim=-1)                        # [NB,H,W,15,15,2]
    assert grids.shape == g_ref.shape and rel_to_max(grids.cpu(), g_ref) < TOL


@pytest.mark.parametrize("simple,inverse", [(False, True), (True, False)])
def test_resample_of_correlation_map_with_explicit_grid(simple, inverse):
    from os2d_b200.head import Os2dHead
    hc, tn, cms, fm, cf, corr, P = _setup(simple, inverse)
    B, C = fm.shape[0], len(cms)
    H, W = fm.shape[-2:]
    p_ref, _ = ho.transform_net(corr, tn)
    theta = ho.theta_from_params(p_ref, simple, inverse)
    score_ref = ho.resample_and_pool(corr, theta).view(B, C, 1, H, W)
    # the unit-coordinate grid the reference head feeds to the resampler (head.py:371-384)
    a, b, tx, c, d, ty = theta
    lin = torch.linspace(-1, 1, 15)
    gx = a[..., None, None] * lin[None, None, None, None, :] + b[..., None, None] * lin[None, None, None, :, None] + tx[..., None, None]
    gy = c[..., None, None] * lin[None, None, None, None, :] + d[..., None, None] * lin[None, None, None, :, None] + ty[..., None, None]
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, W, 1, 1)
    ys = torch.arange(H, dtype=torch.float32).view(1, H, 1, 1, 1)
    ux = ((gx * 7.5 + xs + 0.5) / (W - 1) * 2 - 1).clamp(-1, 1)
    uy = ((gy * 7.5 + ys + 0.5) / (H - 1) * 2 - 1).clamp(-1, 1)
    grid = torch.stack([ux, uy], dim=-1).view(B, C, H, W, 15, 15, 2)
    with torch.no_grad():
        head = hc.create_os2d_head([m.cuda() for m in cms])
        out = Os2dHead.resample_of_correlation_map_fast(corr.view(B, C, 225, H, W).cuda(), grid.cuda(), head.class_pool_mask)
        out2 = Os2dHead.resample_of_correlation_map_simple(corr.view(B, C, 225, H, W).cuda(), grid.cuda(), head.class_pool_mask)
    assert out.shape == (B, C, 1, H, W)
    assert rel_to_max(out.cpu(), score_ref) < 1e-5               # fp32 in, fp32 math: far below the 1e-3 bar
    assert torch.equal(out, out2)
