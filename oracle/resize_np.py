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


def _bilinear(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _bicubic(x: float) -> float:
    """Resample.c bicubic_filter (a = -0.5, support 2) — `Image.resize(..., BICUBIC)`, the resize of
    third_party/vggt/vggt/utils/load_fn.py:166."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0)}


def coeffs(in_size: int, out_size: int, filt: str = "bilinear"):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear / bicubic filter over the whole axis (box = full
    image): returns (bounds int32 [out, 2] = (first input index, tap count), kk int32 [out, ksize], ksize)."""
    fn, base_support = FILTERS[filt]
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = base_support * filterscale
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
            w = fn((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << PRECISION_BITS)) if k[x] < 0 else int(0.5 + k[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img: np.ndarray, axis: int, out_size: int, filt: str = "bilinear") -> np.ndarray:
    """One 8-bit resampling pass along `axis` of a [H, W, C] uint8 image."""
    bounds, kk, ksize = coeffs(img.shape[axis], out_size, filt)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for i in range(out_size):
        lo, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(n):
            acc += src[lo + t] * int(kk[i, t])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_u8(img: np.ndarray, out_h: int, out_w: int, filt: str = "bilinear") -> np.ndarray:
    """[H, W, C] uint8 -> [out_h, out_w, C] uint8 exactly as PIL.Image.resize((out_w, out_h), BILINEAR | BICUBIC): horizontal
    pass first (skipped when the width does not change), then the vertical one."""
    x = img
    if x.shape[1] != out_w:
        x = _pass(x, 1, out_w, filt)
    if x.shape[0] != out_h:
        x = _pass(x, 0, out_h, filt)
    return np.ascontiguousarray(x)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    return resize_u8(img, out_h, out_w, "bilinear")


def vggt_preprocess(frames_u8: np.ndarray) -> np.ndarray:
    """third_party/vggt/vggt/utils/load_fn.py:135-170 (`load_and_preprocess_images`, mode "crop") on RGB frames uint8
    [N, H, W, 3] of one size: width -> 518, height -> round(H * 518 / W / 14) * 14 by PIL BICUBIC, ToTensor, centre crop of the
    height to 518 -> float32 [N, 3, h, 518] in [0, 1].  (The reference reaches it through a lossless PNG round trip,
    unified_loop_consistency.py:339-348.)"""
    N, H, W, _ = frames_u8.shape
    new_w = 518
    new_h = round(H * (new_w / W) / 14) * 14
    out = np.stack([resize_u8(f, new_h, new_w, "bicubic") for f in frames_u8]).astype(np.float32) / 255.0
    out = out.transpose(0, 3, 1, 2)
    if new_h > 518:
        y0 = (new_h - 518) // 2
        out = out[:, :, y0:y0 + 518]
    return np.ascontiguousarray(out)


def load_and_preprocess_images_np(image_path_list, mode: str = "crop") -> np.ndarray:
    """third_party/vggt/vggt/utils/load_fn.py:95-230 restated with the numpy BICUBIC above: decode (alpha composited on
    white), resize, ToTensor, centre crop ("crop") / white padding to 518 x 518 ("pad"), white padding of differently shaped
    images to the largest -> float32 [N, 3, H, W]."""
    from PIL import Image

    if len(image_path_list) == 0:
        raise ValueError("At least 1 image is required")
    if mode not in ["crop", "pad"]:
        raise ValueError("Mode must be either 'crop' or 'pad'")
    target = 518
    images = []
    for path in image_path_list:
        img = Image.open(path)
        if img.mode == "RGBA":
            img = Image.alpha_composite(Image.new("RGBA", img.size, (255, 255, 255, 255)), img)
        a = np.asarray(img.convert("RGB"))
        H, W = a.shape[:2]
        if mode == "pad":
            if W >= H:
                nw, nh = target, round(H * (target / W) / 14) * 14
            else:
                nh, nw = target, round(W * (target / H) / 14) * 14
        else:
            nw, nh = target, round(H * (target / W) / 14) * 14
        t = resize_u8(a, nh, nw, "bicubic").astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
        if mode == "crop" and nh > target:
            y0 = (nh - target) // 2
            t = t[:, y0:y0 + target]
        if mode == "pad":
            hp, wp = target - t.shape[1], target - t.shape[2]
            if hp > 0 or wp > 0:
                t = np.pad(t, ((0, 0), (hp // 2, hp - hp // 2), (wp // 2, wp - wp // 2)), constant_values=1.0)
        images.append(t)
    shapes = {(t.shape[1], t.shape[2]) for t in images}
    if len(shapes) > 1:
        mh, mw = max(s[0] for s in shapes), max(s[1] for s in shapes)
        images = [np.pad(t, ((0, 0), ((mh - t.shape[1]) // 2, mh - t.shape[1] - (mh - t.shape[1]) // 2),
                             ((mw - t.shape[2]) // 2, mw - t.shape[2] - (mw - t.shape[2]) // 2)), constant_values=1.0) for t in images]
    return np.ascontiguousarray(np.stack(images))
