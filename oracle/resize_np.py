"""ORACLE — test infrastructure only (see oracle/__init__.py).

numpy restatement of Pillow's bilinear `Image.resize` on 8-bit images — what `transforms.Resize((height, width))` does to the
24 reprojected memory panoramas of a segment before they condition the next clip (dataset/CameraTrajDataset.py:597-600 through
`dataset.transform(numpy_to_image(m))`, unified_loop_consistency.py:422).  Pillow's algorithm (src/libImaging/Resample.c:
`precompute_coeffs`, `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`): a separable
triangle filter whose support grows with the down-scaling factor, coefficients normalised in double precision and rounded
to 22-bit fixed point, a horizontal pass into an 8-bit intermediate, then a vertical pass; every pass accumulates integers
from 1 << 21 and clips (sum >> 22) to [0, 255].

PARITY PINNED by the library itself: Pillow is installed in this image; tests/test_resize_host.py compares this restatement
with `PIL.Image.resize` bit for bit, tests/test_gpu_resize.py the CUDA kernels with both.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter over the whole axis (box = full image):
    returns (bounds int32 [out, 2] = (first input index, tap count), kk int32 [out, ksize], ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            w = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - w if w < 1.0 else 0.0
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << PRECISION_BITS)) if k[x] < 0 else int(0.5 + k[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img: np.ndarray, axis: int, out_size: int) -> np.ndarray:
    """One 8-bit resampling pass along `axis` of a [H, W, C] uint8 image."""
    bounds, kk, ksize = coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for i in range(out_size):
        lo, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(n):
            acc += src[lo + t] * int(kk[i, t])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """[H, W, C] uint8 -> [out_h, out_w, C] uint8 exactly as PIL.Image.resize((out_w, out_h), BILINEAR): horizontal pass first
    (skipped when the width does not change), then the vertical one."""
    x = img
    if x.shape[1] != out_w:
        x = _pass(x, 1, out_w)
    if x.shape[0] != out_h:
        x = _pass(x, 0, out_h)
    return np.ascontiguousarray(x)
