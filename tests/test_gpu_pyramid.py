"""GPU: image pyramid levels (os2d_b200.pyramid) are bit-identical to PIL Image.resize(BILINEAR)
+ torchvision ToTensor + Normalize (the reference's data path) and to the integer oracle."""
import numpy as np
import pytest
import torch

from oracle import resize_oracle as ro

pytestmark = pytest.mark.gpu
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _image(w, h, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    img[: h // 4] = rng.integers(0, 2, size=(h // 4, w, 3), dtype=np.uint8) * 255
    return img


@pytest.mark.parametrize("src,dst", [((64, 48), (32, 24)), ((97, 61), (41, 77)), ((33, 47), (160, 47)), ((200, 150), (37, 29)),
                                     ((20, 30), (20, 30)), ((685, 507), (1280, 947)), ((640, 480), (224, 168))])
def test_level_equals_oracle_and_pillow(src, dst):
    from os2d_b200.pyramid import resize_normalize
    img = _image(src[0], src[1], src[0] + dst[1])
    out, out_u8 = resize_normalize(torch.from_numpy(img).cuda(), dst[0], dst[1], MEAN, STD, return_bytes=True)
    ref_u8 = ro.resize_bilinear_u8(img, dst[0], dst[1])
    assert np.array_equal(out_u8.cpu().numpy(), ref_u8)
    assert np.array_equal(out.cpu().numpy(), ro.to_tensor_normalize(ref_u8, MEAN, STD))
    PIL_Image = pytest.importorskip("PIL.Image")
    pil = np.asarray(PIL_Image.fromarray(img, "RGB").resize(dst, PIL_Image.BILINEAR))
    assert np.array_equal(out_u8.cpu().numpy(), pil)
    t = torch.from_numpy(pil.copy()).permute(2, 0, 1).to(torch.float32).div(255)
    t = (t - torch.tensor(MEAN).view(3, 1, 1)) / torch.tensor(STD).view(3, 1, 1)
    assert torch.equal(out.cpu(), t)


def test_image_pyramid_levels_and_inverse_transforms():
    from os2d_b200.pyramid import image_pyramid
    from os2d_b200.structures import BoxList, FeatureMapSize
    img = _image(320, 240, 5)
    scales = (0.5, 0.8, 1, 1.6)
    levels, sizes, inverse = image_pyramid(img, scales, MEAN, STD)
    assert [(s.w, s.h) for s in sizes] == ro.pyramid_sizes(320, 240, scales)
    for lvl, size in zip(levels, sizes):
        assert lvl.shape == (3, size.h, size.w) and lvl.is_cuda
        ref = ro.to_tensor_normalize(ro.resize_bilinear_u8(img, size.w, size.h), MEAN, STD)
        assert np.array_equal(lvl.cpu().numpy(), ref)
    boxes = BoxList(torch.tensor([[10.0, 20.0, 50.0, 60.0]]), sizes[0])
    back = inverse[0](boxes)
    assert back.image_size == FeatureMapSize(w=320, h=240) and torch.equal(back.bbox_xyxy, torch.tensor([[20.0, 40.0, 100.0, 120.0]]))
    with pytest.raises(ValueError):
        image_pyramid(np.zeros((4, 4), dtype=np.uint8))
