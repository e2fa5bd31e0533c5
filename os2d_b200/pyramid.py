"""Image pyramid on the device: replacement for the evaluation branch of
``DataloaderOneShotDetection._transform_image_to_pyramid`` (os2d/data/dataloader.py:272-343, no augmentation):

    for every pyramid scale s: PIL ``img.resize((int(w*s), int(h*s)), Image.BILINEAR)`` -> ToTensor -> Normalize

The uint8 image is uploaded once; every level is two kernel launches (csrc/pyramid.cu) whose output equals the
PIL + torchvision result bit for bit (tests/test_gpu_pyramid.py), already on the GPU in the layout the backbone takes.
The coefficient tables of Pillow's resampler are tiny and are computed here in double precision exactly like
``precompute_coeffs`` / ``normalize_coeffs_8bpc`` of Pillow's Resample.c (bilinear filter, support 1).
"""
import ctypes
import functools

import numpy as np
import torch

from . import _cabi
from .structures import FeatureMapSize

PRECISION_BITS = 32 - 8 - 2


@functools.lru_cache(maxsize=256)
def resize_coefficients(in_size, out_size):
    """(bounds int32 [out,2] = (first tap, tap count), coeffs int32 [out,ksize] with 22 fractional bits, ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C cast: truncation
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    n = xmax - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    a = np.abs((x + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(a < 1.0, 1.0 - a, 0.0)
    w = np.where(np.arange(ksize)[None, :] < n[:, None], w, 0.0)
    ww = np.add.accumulate(w, axis=1)[:, -1]                                  # left-to-right sum, like the C loop
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    coeffs = (w * float(1 << PRECISION_BITS) + 0.5).astype(np.int64).astype(np.int32)     # weights are >= 0
    coeffs = np.where(np.arange(ksize)[None, :] < n[:, None], coeffs, 0).astype(np.int32)
    bounds = np.stack([xmin, n], axis=1).astype(np.int32)
    return bounds, np.ascontiguousarray(coeffs), ksize


def _as_u8_hwc(image, device):
    """PIL image / ndarray / tensor [H,W,3] uint8 -> contiguous CUDA uint8 tensor."""
    if isinstance(image, torch.Tensor):
        t = image
    else:
        arr = np.asarray(image)
        if arr.ndim != 3 or arr.shape[2] != 3:
            raise ValueError("expected an RGB image [H,W,3], got shape {}".format(arr.shape))
        t = torch.from_numpy(np.array(arr, copy=True))     # PIL hands out read-only buffers
    if t.dtype != torch.uint8 or t.dim() != 3 or t.size(2) != 3:
        raise ValueError("expected a uint8 image [H,W,3]")
    return t.to(device).contiguous()


@_cabi.on_device_of
def resize_normalize(img_u8, out_w, out_h, mean, std, return_bytes=False):
    """One pyramid level: CUDA uint8 [H,W,3] -> fp32 [3,out_h,out_w] (and optionally the resized bytes [out_h,out_w,3])."""
    if img_u8.device.type != "cuda":
        raise RuntimeError("os2d_b200 requires CUDA tensors (no CPU path)")
    lib = _cabi.load()
    H, W = img_u8.shape[:2]
    dev = img_u8.device
    xb, xc, xk = resize_coefficients(W, out_w)
    yb, yc, yk = resize_coefficients(H, out_h)
    xb_d, xc_d = torch.from_numpy(xb).to(dev), torch.from_numpy(xc).to(dev)
    yb_d, yc_d = torch.from_numpy(yb).to(dev), torch.from_numpy(yc).to(dev)
    tmp = torch.empty(H, out_w, 3, dtype=torch.uint8, device=dev)
    out = torch.empty(3, out_h, out_w, dtype=torch.float32, device=dev)
    out_u8 = torch.empty(out_h, out_w, 3, dtype=torch.uint8, device=dev) if return_bytes else None
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    _cabi.check(lib.os2d_resize_level(_cabi.ptr(img_u8), H, W, out_h, out_w, _cabi.ptr(xb_d), _cabi.ptr(xc_d), xk,
                                      _cabi.ptr(yb_d), _cabi.ptr(yc_d), yk, ctypes.cast(m, ctypes.c_void_p),
                                      ctypes.cast(s, ctypes.c_void_p), _cabi.ptr(tmp), _cabi.ptr(out), _cabi.ptr(out_u8),
                                      _cabi.stream_ptr()), "os2d_resize_level")
    for t in (xb_d, xc_d, yb_d, yc_d, tmp):
        t.record_stream(torch.cuda.current_stream())
    return (out, out_u8) if return_bytes else out


def image_pyramid(image, pyramid_scales=(1,), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), device="cuda"):
    """Evaluation-time image pyramid (dataloader.py:318-343).  Returns (list of fp32 [3,h_l,w_l] CUDA tensors,
    list of FeatureMapSize per level, list of box inverse transforms: callables BoxList -> BoxList resized back to the
    original image, as the TransformList entries of transforms.py:77)."""
    img = _as_u8_hwc(image, torch.device(device))
    H, W = img.shape[:2]
    image_size = FeatureMapSize(w=W, h=H)
    levels, sizes, inverse = [], [], []
    for s in pyramid_scales:
        size = FeatureMapSize(w=int(W * s), h=int(H * s))                    # dataloader.py:322
        levels.append(resize_normalize(img, size.w, size.h, mean, std))
        sizes.append(size)
        inverse.append(lambda boxes, _t=image_size: boxes.resize(_t))
    return levels, sizes, inverse
