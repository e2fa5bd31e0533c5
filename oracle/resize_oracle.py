"""CPU restatement (numpy, integer arithmetic) of the image-pyramid resize of the reference's evaluation data path:
``img.resize((w, h), Image.BILINEAR)`` (os2d/structures/transforms.py:72, called per pyramid level from
os2d/data/dataloader.py:322-334) followed by ``ToTensor`` + ``Normalize`` (dataloader.py:336-341).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  The arithmetic lives in a third-party dependency that is not vendored
in the reference: Pillow (``Image.resize`` -> ``ImagingResample``, src/libImaging/Resample.c; the reference's INSTALL.md
pins no version, this image has Pillow 12.2.0).  Restated here from its published algorithm and pinned bit for bit
against the installed Pillow by tests/test_oracle_resize.py (random images, up- and down-scaling, odd sizes):
  * per output coordinate: centre = (i + 0.5) * scale, support = max(scale, 1) (bilinear filter, support 1), taps
    xmin = int(centre - support + 0.5) .. xmax = int(centre + support + 0.5) clipped to the image, triangle weights
    1 - |x + xmin - centre + 0.5| / max(scale, 1) normalised to sum 1 in double precision, then rounded to fixed point
    with 22 fractional bits (PRECISION_BITS = 32 - 8 - 2);
  * horizontal pass over all rows into a uint8 image, then vertical pass; each output byte is
    clip8((2^21 + sum_k pixel_k * coeff_k) >> 22);
  * ToTensor: byte / 255 in fp32; Normalize: (x - mean) / std in fp32.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bilinear_coeffs(in_size, out_size):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter over the whole axis.
    Returns (bounds [out,2] int32 = (first tap, number of taps), coeffs [out,ksize] int32 fixed point, ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = np.zeros(n, dtype=np.float64)
        for x in range(n):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = w.sum() if n else 0.0
        # Pillow accumulates ww in a loop (left to right); numpy's pairwise sum equals it for these short vectors only
        # by luck, so do it the same way
        ww = 0.0
        for x in range(n):
            ww += w[x]
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            coeffs[xx, x] = int(w[x] * (1 << PRECISION_BITS) - 0.5) if w[x] < 0 else int(w[x] * (1 << PRECISION_BITS) + 0.5)
        bounds[xx] = (xmin, n)
    return bounds, coeffs, ksize


def _pass(img, bounds, coeffs, axis):
    """One separable pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H,W,C] image."""
    src = img.astype(np.int64)
    out_n = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = out_n
    out = np.zeros(shape, dtype=np.uint8)
    for o in range(out_n):
        first, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(shape[:axis] + shape[axis + 1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for k in range(n):
            acc = acc + np.take(src, first + k, axis=axis) * int(coeffs[o, k])
        val = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        if axis == 0:
            out[o] = val
        else:
            out[:, o] = val
    return out


def resize_bilinear_u8(img, out_w, out_h):
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C] like PIL.Image.resize((out_w, out_h), Image.BILINEAR): horizontal pass first
    (skipped when the width does not change), then vertical pass (skipped when the height does not change)."""
    H, W = img.shape[:2]
    cur = img
    if out_w != W:
        b, c, _ = bilinear_coeffs(W, out_w)
        cur = _pass(cur, b, c, axis=1)
    if out_h != H:
        b, c, _ = bilinear_coeffs(H, out_h)
        cur = _pass(cur, b, c, axis=0)
    return cur


def to_tensor_normalize(img_u8, mean, std):
    """torchvision ToTensor + Normalize in fp32: [H,W,3] uint8 -> [3,H,W] float32, (byte / 255 - mean) / std."""
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    return (x - m) / s


def pyramid_sizes(w, h, scales):
    """dataloader.py:322: FeatureMapSize(w=int(w * s), h=int(h * s)) per pyramid scale."""
    return [(int(w * s), int(h * s)) for s in scales]
