"""GPU: the channels-last backbone tail that writes the head's fp16 operand directly (SURVEY.md section 8f row 3:
os2d/modeling/feature_extractor.py:23-72 + head.py:339 fused, os2d_pack_image_features_nhwc) against the standard path
(fp32 NCHW feature map -> os2d_pack_image_features inside Os2dHead.forward)."""
import pytest
import torch

from _util import rel_to_max
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


def test_nhwc_pack_kernel_matches_torch():
    import ctypes
    from os2d_b200 import _cabi
    lib = _cabi.load()
    g = torch.Generator().manual_seed(2)
    for dtype, rows, D in ((torch.float32, 333, 1024), (torch.float16, 70, 1024), (torch.float32, 9, 256)):
        a = torch.randn(rows, D, generator=g).to(dtype).cuda()
        b = torch.randn(rows, D, generator=g).to(dtype).cuda()
        out = torch.empty(rows, D, dtype=torch.float16, device="cuda")
        for with_b, relu in ((True, 1), (False, 0)):
            rc = lib.os2d_pack_image_features_nhwc(_cabi.ptr(a), _cabi.ptr(b) if with_b else None, 1 if dtype == torch.float16 else 0,
                                                   relu, rows, D, _cabi.ptr(out), _cabi.stream_ptr())
            _cabi.check(rc, "os2d_pack_image_features_nhwc")
            x = a.float() + (b.float() if with_b else 0)
            if relu:
                x = x.relu()
            ref = 32.0 * x / (x.norm(dim=1, keepdim=True) + 1e-5)
            assert float((out.float() - ref).abs().max()) <= 2e-3 * float(ref.abs().max())     # one fp16 ulp of the largest value
    assert lib.os2d_pack_image_features_nhwc(_cabi.ptr(a), None, 0, 0, 9, 250, _cabi.ptr(out), _cabi.stream_ptr()) != 0


def test_packed_backbone_tail_equals_standard_path():
    from os2d_b200.model import Os2dModel
    from os2d_b200.head import PackedFeatureMaps
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    net = Os2dModel(is_cuda=True, backbone_arch="resnet50", merge_branch_parameters=True, use_inverse_geom_model=True,
                    simplify_affine=False)
    net.os2d_head_creator.aligner.parameter_regressor.load_state_dict(dict(ho.random_transform_net(6, seed=11, spread=0.005)), strict=False)
    net.eval()
    g = torch.Generator().manual_seed(4)
    images = torch.randn(2, 3, 160, 208, generator=g).cuda()
    class_images = [torch.randn(3, 64, 80, generator=g).cuda(), torch.randn(3, 96, 48, generator=g).cuda()]
    with torch.no_grad():
        loc, cls, _, size, corners = net(images, class_images)
        packed = net.net_feature_maps.forward_packed(images)
        assert isinstance(packed, PackedFeatureMaps) and packed.shape == (2, 1024, 10, 13) and packed.packed.dtype == torch.float16
        head = net.os2d_head_creator.create_os2d_head(net.net_label_features(class_images))
        l2, c2, _, s2, k2 = net(class_head=head, feature_maps=packed)
        one = net(class_head=head, feature_maps=packed[1])                 # batch slicing of the packed operand
        net.use_packed_feature_maps = True
        l3, c3, _, s3, k3 = net(images, class_images)
    assert size == s2 == s3
    for a, b in ((c2, cls), (l2, loc), (k2, corners)):
        assert rel_to_max(a, b) < 1e-4          # same arithmetic; only cuDNN's NHWC kernels and the norm's summation order differ
    assert torch.equal(l3, l2) and torch.equal(c3, c2) and torch.equal(k3, k2)
    assert torch.equal(one[1][0], c2[1])
