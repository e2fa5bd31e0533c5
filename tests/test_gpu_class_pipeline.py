"""GPU: class-side pipeline: ragged one-launch packing of differently sized class feature maps
and the size-batched class branch of the backbone, against the per-class path and the oracle."""
import pytest
import torch

from _util import rel_to_max
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


def _maps(seed, sizes, D=64):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(1, D, h, w, generator=g) * 0.5 + 0.2).relu() for (h, w) in sizes]


def test_ragged_pack_equals_per_class_pack_and_oracle():
    from os2d_b200 import head as bh
    sizes = [(15, 15), (12, 18), (19, 11), (15, 15), (7, 30), (12, 18), (1, 1), (2, 40)]
    cms = _maps(5, sizes)
    dev_maps = [m.cuda() for m in cms]
    cf, pk = bh._prepare_class_operands(dev_maps)                      # ragged: one launch
    for i, m in enumerate(dev_maps):                                   # uniform entry point, class by class
        cf1, pk1 = bh._prepare_class_operands([m])
        assert torch.equal(cf[i:i + 1], cf1) and torch.equal(pk[i:i + 1], pk1)
    ref = ho.prepare_class_features(cms)
    assert rel_to_max(cf.cpu(), ref) < 1e-6
    # fp16 operand = fp16(32 * normalised value), row k = tx*15 + ty, rows 225.. zero
    expect = (ref * 32.0).permute(0, 3, 2, 1).reshape(len(cms), 225, -1)
    assert torch.equal(pk[:, 225:], torch.zeros_like(pk[:, 225:]))
    assert (pk[:, :225].float().cpu() - expect).abs().max() <= 2.0 ** -11 * expect.abs().max() * 1.01
    # a sliced contiguous batch takes the uniform path and gives the same bits as the pointer path
    batch = torch.cat([dev_maps[0], dev_maps[3]], dim=0)
    cf_b, pk_b = bh._prepare_class_operands([batch[0:1], batch[1:2]])
    assert torch.equal(cf_b[0], cf[0]) and torch.equal(cf_b[1], cf[3]) and torch.equal(pk_b[1], pk[3])


def test_size_batched_class_branch_matches_per_image_loop():
    from os2d_b200.model import Os2dModel
    torch.manual_seed(0)
    net = Os2dModel(is_cuda=True, backbone_arch="resnet50", use_inverse_geom_model=True, simplify_affine=False)
    net.eval()
    g = torch.Generator().manual_seed(3)
    shapes = [(240, 240), (192, 304), (240, 240), (320, 176), (192, 304), (240, 240)]
    images = [torch.randn(3, h, w, generator=g).cuda() for (h, w) in shapes]
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False     # isolate the effect of batching from TF32 algorithm choices
    try:
        _check_batched_branch(net, images, shapes, g)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


def _check_batched_branch(net, images, shapes, g):
    with torch.no_grad():
        batched = net.net_label_features(images)
        loop = [net.net_label_features.net_class_features(im.unsqueeze(0)) for im in images]
        for a, b, (h, w) in zip(batched, loop, shapes):
            assert a.shape == b.shape == (1, 1024, -(-h // 16), -(-w // 16))
            assert rel_to_max(a, b) < 5e-4      # same math, cuDNN may choose another algorithm for another batch size
        # the whole class side -> head operand, and the head on top of it
        head_b = net.os2d_head_creator.create_os2d_head(batched)
        head_l = net.os2d_head_creator.create_os2d_head(loop)
        assert rel_to_max(head_b.class_feature_maps, head_l.class_feature_maps) < 5e-4
        fm = net.net_feature_maps(torch.randn(1, 3, 384, 512, generator=g).cuda())
        lb, sb, _, cb = head_b(fm)
        ll, sl, _, cl = head_l(fm)
        assert rel_to_max(sb, sl) < 1e-3 and rel_to_max(lb, ll) < 1e-3 and rel_to_max(cb, cl) < 1e-3
