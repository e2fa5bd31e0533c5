"""CPU: the integer restatement of Pillow's bilinear resize (oracle/resize_oracle.py) against the installed Pillow, bit for
bit, and ToTensor + Normalize against torchvision-style tensor arithmetic."""
import numpy as np
import pytest
import torch

from oracle import resize_oracle as ro

PIL = pytest.importorskip("PIL")
from PIL import Image   # noqa: E402

CASES = [((64, 48), (32, 24)), ((64, 48), (96, 72)), ((97, 61), (41, 77)), ((50, 50), (50, 20)), ((33, 47), (160, 47)),
         ((200, 150), (37, 29)), ((31, 17), (7, 5)), ((20, 30), (20, 30)), ((123, 77), (124, 76)), ((640, 480), (224, 168))]


@pytest.mark.parametrize("src,dst", CASES)
def test_resize_oracle_equals_pillow(src, dst):
    rng = np.random.default_rng(src[0] * 1000 + dst[0])
    img = rng.integers(0, 256, size=(src[1], src[0], 3), dtype=np.uint8)
    img[: src[1] // 3] = rng.integers(0, 2, size=(src[1] // 3, src[0], 3), dtype=np.uint8) * 255      # saturated stripes
    ref = np.asarray(Image.fromarray(img, "RGB").resize(dst, Image.BILINEAR))
    out = ro.resize_bilinear_u8(img, dst[0], dst[1])
    assert out.shape == ref.shape and np.array_equal(out, ref)


def test_to_tensor_normalize_is_fp32_exact():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(9, 11, 3), dtype=np.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    t = torch.from_numpy(img).permute(2, 0, 1).to(torch.float32).div(255)           # torchvision ToTensor
    t = (t - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1)      # torchvision Normalize
    assert np.array_equal(ro.to_tensor_normalize(img, mean, std), t.numpy())
