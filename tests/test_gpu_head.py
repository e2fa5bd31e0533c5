"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes) via the drop-in classes, against the
committed reference golden vectors and against the CPU oracle on seeded inputs."""
import pytest
import torch

from _util import VARIANTS, load_head_golden, rel_to_max, synth_inputs, TOL
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


def _creator(simple, inverse, tn):
    from os2d_b200 import head as bh
    from os2d_b200.structures import FeatureMapSize
    hc = bh.build_os2d_head_creator(simple, True, inverse, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    return hc


def _run(hc, cms, fm):
    with torch.no_grad():
        head = hc.create_os2d_head([c.cuda() for c in cms])
        loc, rec, rec_detached, corners = head(fm.cuda())
    torch.cuda.synchronize()
    return head, loc, rec, rec_detached, corners


@pytest.mark.parametrize("name,simple,inverse", VARIANTS)
def test_head_matches_reference_golden(name, simple, inverse):
    """Golden vectors were produced by the unmodified reference (tests/golden/make_golden.py), D = 64."""
    d, cms, fm, tn = load_head_golden(name)
    hc = _creator(simple, inverse, tn)
    head, loc, rec, rec_detached, corners = _run(hc, cms, fm)
    assert loc.shape == d["loc"].shape and rec.shape == d["score"].shape and corners.shape == d["corners"].shape
    assert loc.dtype == torch.float32 and loc.is_cuda
    assert rec_detached is rec                       # head.py:400-402 under no-grad
    assert rel_to_max(head.class_feature_maps.cpu(), d["class_features"]) < 1e-5
    assert rel_to_max(rec.cpu(), d["score"]) < TOL
    assert rel_to_max(loc.cpu(), d["loc"]) < TOL
    assert rel_to_max(corners.cpu(), d["corners"]) < TOL


@pytest.mark.parametrize("simple,inverse", [(False, True), (True, False)])
@pytest.mark.parametrize("B,H,W", [(1, 32, 32), (2, 37, 45)])
def test_head_matches_oracle_d1024(simple, inverse, B, H, W):
    """ResNet-C4 sized features (D = 1024), multi-tile geometry (H > 32, widths not multiple of 16), ragged class maps."""
    P = 4 if simple else 6
    tn = ho.random_transform_net(P, seed=21, spread=0.005)
    cms, fm = synth_inputs(40 + H, B, H, W, [(15, 15), (12, 18), (19, 11), (15, 15), (8, 25)])
    hc = _creator(simple, inverse, tn)
    head, loc, rec, _, corners = _run(hc, cms, fm)
    cf = ho.prepare_class_features(cms)
    oloc, osc, ocor = ho.head_forward(cf, fm, tn, simple, inverse)
    assert rel_to_max(rec.cpu(), osc) < TOL
    assert rel_to_max(loc.cpu(), oloc) < TOL
    assert rel_to_max(corners.cpu(), ocor) < TOL
    # elementwise form of the bar: |a-b| <= TOL*|b| + 1e-4*max|b| ... checked on the score map
    ref = osc
    assert bool(((rec.cpu() - ref).abs() <= TOL * ref.abs() + 1e-4 * ref.abs().max()).all())


def test_identity_transform_gives_zero_loc():
    """Default TransformNet init (head.py:631-642) regresses the identity: loc == 0 and corners == anchor corners."""
    from os2d_b200 import head as bh
    from os2d_b200.structures import FeatureMapSize
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.eval()
    cms, fm = synth_inputs(3, 1, 20, 24, [(15, 15), (10, 20)])
    head, loc, rec, _, corners = _run(hc, cms, fm)
    assert float(loc.abs().max()) < 1e-4
    x = (torch.arange(24, device="cuda") + 0.5) * 16
    assert torch.allclose(corners[0, 0, 0], (x - 120).view(1, 24).expand(20, 24), atol=1e-3)
    assert float(rec.min()) >= -1.0 and float(rec.max()) <= 1.0


def test_class_batching_and_image_batching_invariance():
    """One C-class head == C single-class heads; B = 2 == two B = 1 calls (bit-exact: every (image, class) plane is
    computed independently).  This is what justifies class-axis sharding over GPUs."""
    tn = ho.random_transform_net(6, seed=5, spread=0.005)
    cms, fm = synth_inputs(9, 2, 21, 19, [(15, 15), (12, 18), (19, 11)])
    hc = _creator(False, True, tn)
    head, loc, rec, _, corners = _run(hc, cms, fm)
    for c in range(3):
        _, l1, r1, _, c1 = _run(hc, cms[c:c + 1], fm)
        assert torch.equal(l1[:, 0], loc[:, c]) and torch.equal(r1[:, 0], rec[:, c]) and torch.equal(c1[:, 0], corners[:, c])
    for b in range(2):
        with torch.no_grad():
            l1, r1, _, c1 = head(fm[b:b + 1].cuda())
        assert torch.equal(l1[0], loc[b]) and torch.equal(r1[0], rec[b]) and torch.equal(c1[0], corners[b])


def test_outputs_are_fresh_tensors():
    """evaluate.py keeps the outputs of successive calls in lists (evaluate.py:351-357): no buffer may be reused."""
    tn = ho.random_transform_net(6, seed=5, spread=0.005)
    cms, fm = synth_inputs(10, 1, 16, 16, [(15, 15)])
    hc = _creator(False, True, tn)
    head, loc_a, rec_a, _, cor_a = _run(hc, cms, fm)
    keep = loc_a.clone()
    with torch.no_grad():
        loc_b, rec_b, _, cor_b = head((fm * 0.5 + 0.1).cuda())
    torch.cuda.synchronize()
    assert loc_b.data_ptr() != loc_a.data_ptr() and torch.equal(loc_a, keep)


def test_grad_mode_and_cpu_inputs_raise():
    tn = ho.random_transform_net(6, seed=5, spread=0.005)
    cms, fm = synth_inputs(11, 1, 8, 8, [(15, 15)])
    hc = _creator(False, True, tn)
    with torch.no_grad():
        head = hc.create_os2d_head([c.cuda() for c in cms])
    with pytest.raises(RuntimeError):
        head(fm.cuda())                      # grad enabled + trainable aligner parameters
    with torch.no_grad():
        with pytest.raises(RuntimeError):
            head(fm)                         # CPU tensor


def test_model_forward_api():
    """Os2dModel.forward(feature_maps=..., class_head=...) -> (loc [B,C,4,N], cls [B,C,N], cls_detached, fm size, corners [B,C,8,N])."""
    from os2d_b200.model import Os2dModel
    from os2d_b200.structures import FeatureMapSize
    torch.manual_seed(0)
    net = Os2dModel(is_cuda=True, backbone_arch="resnet50", use_inverse_geom_model=True, simplify_affine=False)
    img = torch.randn(1, 3, 200, 264, device="cuda")
    cls_imgs = [torch.randn(3, 240, 240, device="cuda"), torch.randn(3, 200, 288, device="cuda")]
    loc, cls, cls_d, fm_size, corners = net(images=img, class_images=cls_imgs)
    assert fm_size == FeatureMapSize(w=17, h=13) == net.get_feature_map_size(FeatureMapSize(w=264, h=200))
    N = 17 * 13
    assert loc.shape == (1, 2, 4, N) and cls.shape == (1, 2, N) and corners.shape == (1, 2, 8, N)
    assert bool(torch.isfinite(cls).all()) and float(loc.abs().max()) < 1e-4     # identity init


def test_one_cta_correlation_variant_matches():
    """The 1-CTA correlation main loop (OS2D_B200_CORR_1CTA=1, A/B switch of csrc/corr.cu) gives the same result as the
    default 2-CTA (cta_group::2) variant; run in a subprocess because the switch is read once per process."""
    import os
    import subprocess
    import sys
    from _util import ROOT
    code = r"""
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from _util import synth_inputs
from oracle import head_oracle as ho
from os2d_b200 import head as bh
from os2d_b200.structures import FeatureMapSize
tn = ho.random_transform_net(6, seed=5, spread=0.005)
cms, fm = synth_inputs(9, 1, 21, 19, [(15, 15), (12, 18), (19, 11)])
hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False); hc.eval()
with torch.no_grad():
    loc, rec, _, cor = hc.create_os2d_head([c.cuda() for c in cms])(fm.cuda())
torch.save([loc.cpu(), rec.cpu(), cor.cpu()], sys.argv[1])
""" % (ROOT, ROOT)
    outs = []
    for variant in ("2cta", "1cta"):
        env = dict(os.environ)
        env.pop("OS2D_B200_CORR_1CTA", None)
        if variant == "1cta":
            env["OS2D_B200_CORR_1CTA"] = "1"
        path = "/tmp/os2d_b200_corr_%s.pt" % variant
        r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(torch.load(path))
    for a, b in zip(*outs):
        assert torch.equal(a, b)        # same MMA order per accumulator element => bit-identical


def test_submit_with_async_resample_equals_forward():
    """Os2dHead.submit (K3 on a side stream, overlapping the next call's tensor kernels) == forward, bit for bit, also with
    several submits in flight, chunked workspaces and caller-provided output views."""
    from os2d_b200 import head as bh
    from os2d_b200.structures import FeatureMapSize
    tn = ho.random_transform_net(6, seed=3, spread=0.005)
    cms, fm = synth_inputs(21, 2, 33, 29, [(15, 15), (12, 18), (19, 11), (15, 15), (10, 20), (15, 15), (9, 9)])
    _, fm2 = synth_inputs(22, 2, 33, 29, [(15, 15)])
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head([c.cuda() for c in cms])
        refs = [head(f.cuda()) for f in (fm, fm2)]
        pend = [head.submit(f.cuda()) for f in (fm, fm2, fm, fm2)]          # four calls in flight
        for i, (out, ev) in enumerate(pend):
            torch.cuda.current_stream().wait_event(ev)
            for a, b in zip(out, refs[i % 2]):
                assert torch.equal(a, b)
        head.max_planes_per_call = 6                                        # several chunks, each with its own K3 launch
        head._cmax_cache.clear()
        out, ev = head.submit(fm.cuda())
        ev.synchronize()
        for a, b in zip(out, refs[0]):
            assert torch.equal(a, b)
        B, C, N = 2, 7, 33 * 29
        buf = torch.zeros(B, C, 13, N, device="cuda")
        none, ev = head.submit(fm2.cuda(), out_views=(buf[:, :, 0:1], buf[:, :, 1:5], buf[:, :, 5:13]))
        assert none is None
        ev.synchronize()
        assert torch.equal(buf[:, :, 1:5].reshape(B, C, 4, 33, 29), refs[1][0])
        assert torch.equal(buf[:, :, 0:1].reshape(B, C, 1, 33, 29), refs[1][1])
