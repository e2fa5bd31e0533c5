"""CPU: host-side logic of the drop-in boundary (no kernels run)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from os2d_b200 import head as bh
from os2d_b200.structures import FeatureMapSize, BoxList, cat_boxlist
from os2d_b200.box_coder import Os2dBoxCoder
from oracle import head_oracle as ho
from oracle import postproc_oracle as po

REFERENCE_HEAD_KEYS = sorted(
    "aligner.parameter_regressor." + k for k in
    ["conv.0.weight", "conv.0.bias", "conv.1.weight", "conv.1.bias", "conv.1.running_mean", "conv.1.running_var",
     "conv.1.num_batches_tracked", "conv.3.weight", "conv.3.bias", "conv.4.weight", "conv.4.bias", "conv.4.running_mean",
     "conv.4.running_var", "conv.4.num_batches_tracked", "linear.weight", "linear.bias"])


def _creator(simple=False, inverse=True):
    return bh.build_os2d_head_creator(simple, False, inverse, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))


def test_state_dict_keys_match_reference():
    hc = _creator()
    assert sorted(hc.state_dict().keys()) == REFERENCE_HEAD_KEYS     # 16 tensors
    sd = hc.state_dict()
    assert tuple(sd["aligner.parameter_regressor.conv.0.weight"].shape) == (128, 225, 7, 7)
    assert tuple(sd["aligner.parameter_regressor.conv.3.weight"].shape) == (64, 128, 5, 5)
    assert tuple(sd["aligner.parameter_regressor.linear.weight"].shape) == (6, 64, 5, 5)
    assert tuple(_creator(simple=True).state_dict()["aligner.parameter_regressor.linear.weight"].shape) == (4, 64, 5, 5)


def test_identity_init_and_box_grid():
    hc = _creator()
    lin = hc.aligner.parameter_regressor.linear
    assert float(lin.weight.detach().abs().max()) == 0.0
    assert lin.bias.tolist() == [1, 0, 0, 0, 1, 0]
    assert _creator(simple=True).aligner.parameter_regressor.linear.bias.tolist() == [1, 0, 1, 0]
    g = hc.box_grid_generator_image_level
    assert (g.box_size.w, g.box_size.h, g.box_stride.w, g.box_stride.h) == (240, 240, 16, 16)   # head.py:216-238
    f = hc.box_grid_generator_feature_map_level
    assert (f.box_size.w, f.box_stride.w) == (15, 1)
    boxes = g.create_strided_boxes_columnfirst(FeatureMapSize(w=5, h=3))
    np.testing.assert_array_equal(boxes.numpy(), po.anchors_xyxy(5, 3))


def _dequant_layer(blob, ks, rows, ci):
    """shared-memory image [sc][dy][dx][kg][128][8] -> [128, ci_pad, ks, ks] fp32"""
    nsub = blob.shape[0]
    return blob.float().permute(4, 0, 3, 5, 1, 2).reshape(128, nsub * 16, ks, ks)[:rows, :ci]


@pytest.mark.parametrize("P", [6, 4])
def test_packed_transform_net_reproduces_the_convolutions(P):
    """The packed fp16 operands + fp32 alpha/beta reproduce conv+BN(eval) of the reference network (emulated in fp64
    on the dequantised blobs), including the DC side channels of layer 1 and the hi/lo rows of layers 2 and 3."""
    tn = ho.random_transform_net(P, seed=2, spread=0.01)
    pw = bh.pack_transform_net(tn, P, "cpu")
    g = torch.Generator().manual_seed(0)
    # ---- layer 1 ----
    z = torch.rand(1, 225, 9, 8, generator=g) + 0.5
    z = z / z.pow(2).sum(1, keepdim=True).sqrt()
    m = z.mean(1, keepdim=True)
    x = torch.zeros(1, 240, 9, 8, dtype=torch.float64)
    x[:, :225] = (z - m) * bh.SCALE_Z
    x[:, 225] = x[:, 226] = (m * bh.SCALE_MEAN)[:, 0]
    w = _dequant_layer(pw["w1"], 7, 128, 240).double()
    acc = F.conv2d(x, w, None, padding=3)
    got = F.relu(acc * pw["alpha1"].double().view(1, -1, 1, 1) + pw["beta1"].double().view(1, -1, 1, 1))
    a1, b1 = ho.fold_bn(tn["conv.0.weight"], tn["conv.0.bias"], tn["conv.1.weight"], tn["conv.1.bias"],
                        tn["conv.1.running_mean"], tn["conv.1.running_var"])
    ref = F.relu(F.conv2d(z.double(), tn["conv.0.weight"].double(), None, padding=3) * a1.double().view(1, -1, 1, 1)
                 + b1.double().view(1, -1, 1, 1))
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-4
    # ---- layer 2: hi + lo / 2048 ----
    h1 = torch.rand(1, 128, 7, 9, generator=g).double()
    w = _dequant_layer(pw["w2"], 5, 128, 128).double()
    acc = F.conv2d(h1, w, None, padding=2)
    rh = pw["row_hi2"]                                  # hi row of channel c; the residual row is rh + 16
    comb = acc[:, rh] + acc[:, rh + 16] / bh.LO_SCALE
    assert torch.equal(pw["alpha2"][rh], pw["alpha2"][rh + 16])
    got = F.relu(comb * pw["alpha2"][rh].double().view(1, -1, 1, 1) + pw["beta2"][rh].double().view(1, -1, 1, 1))
    a2, b2 = ho.fold_bn(tn["conv.3.weight"], tn["conv.3.bias"], tn["conv.4.weight"], tn["conv.4.bias"],
                        tn["conv.4.running_mean"], tn["conv.4.running_var"])
    ref = F.relu(F.conv2d(h1, tn["conv.3.weight"].double(), None, padding=2) * a2.double().view(1, -1, 1, 1)
                 + b2.double().view(1, -1, 1, 1))
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-6
    # ---- layer 3 ----
    h2 = torch.rand(1, 64, 6, 6, generator=g).double()
    # scatter-form blob [16 chunk8 (hi 0..7, lo 8..15)][NPAD][8]: rows (dy*5+dx)*P + co
    blob = pw["w3"].double()
    npad = blob.shape[1]
    assert npad == (25 * P + 15) // 16 * 16
    wrows = (blob[:8] + blob[8:]).permute(1, 0, 2).reshape(npad, 64)          # hi + lo, [n, ci]
    assert float(wrows[25 * P:].abs().max()) == 0.0
    w = wrows[:25 * P].view(5, 5, P, 64).permute(2, 3, 0, 1)                   # [P,64,5,5]
    got = F.conv2d(h2, w, None, padding=2) * pw["alpha3"][0].double() + pw["beta3"][:P].double().view(1, -1, 1, 1)
    ref = F.conv2d(h2, tn["linear.weight"].double(), tn["linear.bias"].double(), padding=2)
    assert float((got - ref).abs().max()) < 1e-6


def test_packed_weights_cache_invalidates_on_change():
    hc = _creator()
    hc.eval()
    reg = hc.aligner.parameter_regressor
    a = reg.packed_weights()
    assert reg.packed_weights() is a
    with torch.no_grad():
        reg.linear.bias.add_(1.0)
    b = reg.packed_weights()
    assert b is not a and float(b["beta3"][0]) == 2.0


def test_train_mode_batchnorm_is_rejected():
    hc = _creator()
    hc.train()
    with pytest.raises(RuntimeError):
        hc.aligner.parameter_regressor.packed_weights()
    hc.aligner.parameter_regressor.freeze_bn()
    hc.aligner.parameter_regressor.packed_weights()


def test_cpu_tensors_fail_loudly():
    hc = _creator()
    with pytest.raises(RuntimeError):
        hc.create_os2d_head([torch.rand(1, 64, 15, 15)])
    bc = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level, lambda s: FeatureMapSize(w=2, h=2))
    with pytest.raises(RuntimeError):
        bc.decode_pyramid([torch.zeros(1, 4, 4)], [torch.zeros(1, 4)], [FeatureMapSize(w=32, h=32)], [0])


def test_feature_map_size_and_boxlist():
    s = FeatureMapSize(w=3, h=4)
    assert s == FeatureMapSize(img=torch.zeros(1, 2, 4, 3)) and hash(s) == hash(FeatureMapSize(w=3, h=4))
    with pytest.raises(AttributeError):
        s.w = 5
    b = BoxList(torch.tensor([[0., 0, 10, 20], [5, 5, 5, 9]]), FeatureMapSize(w=100, h=50))
    b.add_field("scores", torch.tensor([0.1, 0.9]))
    assert b.get_mask_empty_boxes().tolist() == [False, True]
    r = b.resize(FeatureMapSize(w=200, h=100))
    assert r.bbox_xyxy[0].tolist() == [0, 0, 20, 40] and r.get_field("scores") is b.get_field("scores")
    r2 = b.resize(FeatureMapSize(w=200, h=50))
    assert r2.bbox_xyxy[0].tolist() == [0, 0, 20, 20]
    c = cat_boxlist([b, b])
    assert len(c) == 4 and c.get_field("scores").tolist() == pytest.approx([0.1, 0.9, 0.1, 0.9])
    assert len(b[torch.tensor([True, False])]) == 1
    cb = BoxList(torch.tensor([[50., 25, 20, 10]]), FeatureMapSize(w=100, h=50), mode="cx_cy_w_h")
    assert cb.bbox_xyxy[0].tolist() == [40, 20, 60, 30]


def test_build_loc_targets_matches_oracle_encoding():
    theta = [torch.full((1, 2, 2), v) for v in (1.1, 0.05, 0.02, -0.03, 0.9, -0.01)]
    loc, _ = ho.boxes_and_corners(theta, 2, 2)
    lin = torch.linspace(-1, 1, 15)
    a, b, tx, c, d, ty = [t[0, 0, 0].item() for t in theta]
    gx = a * lin[None, :] + b * lin[:, None] + tx
    gy = c * lin[None, :] + d * lin[:, None] + ty
    X, Y = gx * 120 + 8, gy * 120 + 8
    cls_box = BoxList(torch.tensor([[X.min(), Y.min(), X.max(), Y.max()]]), FeatureMapSize(w=2, h=2))
    anchor = BoxList(torch.tensor([[8 - 120., 8 - 120, 8 + 120, 8 + 120]]), FeatureMapSize(w=2, h=2))
    enc = Os2dBoxCoder.build_loc_targets(cls_box, anchor)
    assert torch.allclose(enc[0], loc[0, :, 0, 0], atol=1e-5)


def test_inverse_transform_probe_accepts_resizes_and_refuses_flips():
    """decode_pyramid applies the inverse box transform as a per-axis rescale; a transform that does anything else
    (flip, crop) must be refused instead of silently mis-placing boxes."""
    import pytest
    from os2d_b200.box_coder import _probe_transform_target, make_resize_transform
    from os2d_b200.structures import BoxList, FeatureMapSize
    img = FeatureMapSize(w=320, h=240)
    assert _probe_transform_target(make_resize_transform(FeatureMapSize(w=640, h=480)), img) == FeatureMapSize(w=640, h=480)
    # a plain callable that resizes with different ratios per axis (BoxList.resize's second branch)
    t = _probe_transform_target(lambda b: b.resize(FeatureMapSize(w=160, h=480)), img)
    assert (t.w, t.h) == (160, 480)

    def hflip(boxes):
        b = boxes.bbox_xyxy.clone()
        w = boxes.image_size.w
        out = BoxList(torch.stack([w - b[:, 2], b[:, 1], w - b[:, 0], b[:, 3]], dim=1), boxes.image_size)
        return out
    with pytest.raises(NotImplementedError):
        _probe_transform_target(hflip, img)


def test_packed_feature_maps_mirror_the_tensor_interface():
    """head.PackedFeatureMaps stands in for the [B,D,H,W] tensor the reference passes around: shape / size / device / batch slicing."""
    import pytest
    from os2d_b200.head import PackedFeatureMaps
    packed = torch.zeros(3, 5 * 7, 64, dtype=torch.float16)
    fm = PackedFeatureMaps(packed, 5, 7)
    assert tuple(fm.shape) == (3, 64, 5, 7) and fm.size(1) == 64 and fm.size() == fm.shape and fm.requires_grad is False
    one = fm[1]
    assert tuple(one.shape) == (1, 64, 5, 7) and tuple(fm[1:3].shape) == (2, 64, 5, 7)
    with pytest.raises(AssertionError):
        PackedFeatureMaps(packed.float(), 5, 7)            # the operand is fp16 by construction
    with pytest.raises(AssertionError):
        PackedFeatureMaps(packed, 5, 8)                    # H * W must match the packed rows
