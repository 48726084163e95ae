"""Image pre-processing of the CLIP conditioning branch (once per clip, torch ops on the image's device).

Mirrors evoworld/pipeline/pipeline_evoworld.py:255-291 (`_encode_image`: x*2-1 -> anti-aliased resize to 224x224 ->
(x+1)/2 -> CLIP mean/std normalisation) and :746-850 (`_resize_with_antialiasing` = separable Gaussian blur with reflect
padding, sigma = max((factor-1)/2, 0.001), window = odd(int(max(4 sigma, 3))), then bicubic `interpolate` with
align_corners=True).  Pinned against the identical function in evoworld/trainer/trainer_utils.py:68-179, which imports in
the build container (tests/golden/make_pipeline_golden.py -> tests/golden/pipeline_golden.npz).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import torch
import torch.nn.functional as F

# transformers' OPENAI_CLIP_MEAN / OPENAI_CLIP_STD: what CLIPImageProcessor(do_normalize=True) applies
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
CLIP_SIZE = (224, 224)


def blur_window(size: int, factor: float) -> Tuple[int, float]:
    """(odd window length, sigma) of the anti-aliasing Gaussian for a down-scaling factor in/out."""
    sigma = max((factor - 1.0) / 2.0, 0.001)
    k = int(max(2.0 * 2 * sigma, 3))
    return (k + 1 if k % 2 == 0 else k), sigma


def gaussian_taps(k: int, sigma: torch.Tensor) -> torch.Tensor:
    """Normalised Gaussian taps centred on k // 2, evaluated in sigma's dtype (a 0-d tensor)."""
    x = torch.arange(k, device=sigma.device, dtype=sigma.dtype) - k // 2
    if k % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma.pow(2.0)))
    return g / g.sum(-1, keepdim=True)


def _blur_axis(x: torch.Tensor, taps: torch.Tensor, axis: int) -> torch.Tensor:
    """Depth-wise 1-D correlation along H (axis -2) or W (axis -1) with reflect padding."""
    b, c, h, w = x.shape
    k = taps.numel()
    front, rear = (k - 1) // 2, (k - 1) - (k - 1) // 2
    if axis == -1:
        x = F.pad(x, (front, rear, 0, 0), mode="reflect")
        kern = taps.to(x.dtype).reshape(1, 1, 1, k).expand(c, 1, 1, k)
    else:
        x = F.pad(x, (0, 0, front, rear), mode="reflect")
        kern = taps.to(x.dtype).reshape(1, 1, k, 1).expand(c, 1, k, 1)
    return F.conv2d(x, kern.contiguous(), groups=c)


def resize_with_antialiasing(image: torch.Tensor, size: Sequence[int], interpolation: str = "bicubic",
                             align_corners: bool = True) -> torch.Tensor:
    """image [B,C,H,W] float -> [B,C,size[0],size[1]]: Gaussian pre-filter (W first, then H) + bicubic resample."""
    if image.dim() != 4:
        raise ValueError(f"resize_with_antialiasing expects [B,C,H,W], got {tuple(image.shape)}")
    h, w = image.shape[-2:]
    (ky, sy), (kx, sx) = blur_window(h, h / size[0]), blur_window(w, w / size[1])
    sig = torch.tensor([sy, sx], dtype=image.dtype, device=image.device)
    out = _blur_axis(image, gaussian_taps(kx, sig[1]), -1)
    out = _blur_axis(out, gaussian_taps(ky, sig[0]), -2)
    return F.interpolate(out, size=tuple(size), mode=interpolation, align_corners=align_corners)


def clip_preprocess(image01: torch.Tensor, feature_extractor=None) -> torch.Tensor:
    """image in [0,1] [B,3,H,W] -> CLIP pixel_values [B,3,224,224] (pipeline_evoworld.py:270-286).  With a
    transformers CLIPImageProcessor the normalisation is delegated to it exactly as the reference does; without one the
    CLIP mean/std above are applied."""
    x = image01 * 2.0 - 1.0
    x = resize_with_antialiasing(x, CLIP_SIZE)
    x = (x + 1.0) / 2.0
    if feature_extractor is not None:
        pv = feature_extractor(images=x, do_normalize=True, do_center_crop=False, do_resize=False, do_rescale=False,
                               return_tensors="pt").pixel_values
        return pv.to(x.device)
    mean = torch.tensor(CLIP_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


# ---------------------------------------------------------------------------------------------
# Pillow-exact bilinear resize of 8-bit images (dataset/CameraTrajDataset.py:597-600: transforms.Resize on PIL images)
# ---------------------------------------------------------------------------------------------
_PIL_PRECISION_BITS = 32 - 8 - 2
_pil_tables: dict = {}


def pil_resize_tables(in_size: int, out_size: int, filt: str = "bilinear"):
    """Pillow's bilinear / bicubic coefficient tables for one axis (libImaging/Resample.c precompute_coeffs +
    normalize_coeffs_8bpc, box = the whole axis): (bounds int32 [out, 2] = (first input index, tap count), k int32 [out, ksize],
    ksize).  Double precision with Pillow's operation order (the weights of an output sample are summed tap by tap)."""
    import numpy as np

    if filt not in ("bilinear", "bicubic"):
        raise ValueError(f"pil_resize_tables: filter {filt!r}")
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = (1.0 if filt == "bilinear" else 2.0) * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = 0.0 + (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5), 0.0).astype(np.int64)
    xmax = np.minimum(np.trunc(center + support + 0.5), float(in_size)).astype(np.int64) - xmin
    k = np.zeros((out_size, ksize), dtype=np.float64)
    ww = np.zeros(out_size, dtype=np.float64)
    for t in range(ksize):
        w = np.abs((t + xmin - center + 0.5) * ss)
        if filt == "bilinear":
            w = np.where(w < 1.0, 1.0 - w, 0.0)
        else:  # Resample.c bicubic_filter, a = -0.5
            a = -0.5
            w = np.where(w < 1.0, ((a + 2.0) * w - (a + 3.0)) * w * w + 1, np.where(w < 2.0, (((w - 5) * w + 8) * w - 4) * a, 0.0))
        w = np.where(t < xmax, w, 0.0)
        k[:, t] = w
        ww = ww + w
    k = np.where(ww[:, None] != 0.0, k / np.where(ww == 0.0, 1.0, ww)[:, None], k)
    kk = np.trunc(np.where(k < 0, -0.5, 0.5) + k * float(1 << _PIL_PRECISION_BITS)).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, kk, ksize


def resize_pil_u8(images: torch.Tensor, height: int, width: int, filt: str = "bilinear") -> torch.Tensor:
    """uint8 [N, H, W, 3] (or [H, W, 3]) on a CUDA device -> uint8 [N, height, width, 3], bit-identical to
    `PIL.Image.resize((width, height), BILINEAR | BICUBIC)` of every image — i.e. to `transforms.Resize((height, width))` on the PIL copy
    (evw_resize_pil_u8).  Tables are cached per (size pair, device)."""
    from . import _lib

    _lib.require_cuda(images, "images")
    single = images.dim() == 3
    x = images[None] if single else images
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
        raise ValueError(f"resize_pil_u8 expects uint8 [N,H,W,3], got {images.dtype} {tuple(images.shape)}")
    x = x.contiguous()
    N, H, W, _ = x.shape
    key = (H, W, height, width, str(x.device), filt)
    if key not in _pil_tables:
        bx, kx, ksx = pil_resize_tables(W, width, filt)
        by, ky, ksy = pil_resize_tables(H, height, filt)
        to = lambda a: torch.from_numpy(a).contiguous().to(x.device)
        _pil_tables[key] = (to(bx), to(kx), ksx, to(by), to(ky), ksy)
    bx, kx, ksx, by, ky, ksy = _pil_tables[key]
    out = torch.empty((N, height, width, 3), dtype=torch.uint8, device=x.device)
    tmp = torch.empty((N, H, width, 3), dtype=torch.uint8, device=x.device) if (width != W and height != H) else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().evw_resize_pil_u8(_lib.ptr(x), _lib.ptr(tmp), _lib.ptr(out), N, H, W, height, width, _lib.ptr(bx),
                                                _lib.ptr(kx), ksx, _lib.ptr(by), _lib.ptr(ky), ksy, _lib.stream_ptr(x.device)),
                   "evw_resize_pil_u8")
    return out[0] if single else out


def _to_tensor_lut(device) -> torch.Tensor:
    """ToTensor's byte / 255 as a table computed on the host: a CUDA division by a scalar multiplies by 1 / 255 instead."""
    key = ("to_tensor", str(device))
    if key not in _pil_tables:
        _pil_tables[key] = (torch.arange(256, dtype=torch.float32) / 255).to(device)
    return _pil_tables[key]


def _vggt_target_size(H: int, W: int, mode: str):
    """load_fn.py:150-163: (new_height, new_width) of the BICUBIC resize for a W x H image."""
    target = 518
    if mode == "pad":
        if W >= H:
            return round(H * (target / W) / 14) * 14, target
        return target, round(W * (target / H) / 14) * 14
    return round(H * (target / W) / 14) * 14, target


def vggt_preprocess_u8(frames: torch.Tensor, mode: str = "crop") -> torch.Tensor:
    """`load_and_preprocess_images(paths, mode)` (third_party/vggt/vggt/utils/load_fn.py:135-193) for RGB frames uint8
    [N, H, W, 3] of one size already on the device — the reference writes them to PNG files and reads them back
    (unified_loop_consistency.py:339-348; PNG is lossless): Pillow BICUBIC resize (bit-exact: evw_resize_pil_u8 with the
    bicubic tables) to width 518 ("crop": height round(H * 518 / W / 14) * 14, centre-cropped to 518 when taller) or to the
    larger side 518 ("pad": then white padding to 518 x 518), ToTensor -> float32 [N, 3, h, w] in [0, 1]."""
    if mode not in ("crop", "pad"):
        raise ValueError("Mode must be either 'crop' or 'pad'")
    N, H, W, _ = frames.shape
    new_h, new_w = _vggt_target_size(H, W, mode)
    x = _to_tensor_lut(frames.device)[resize_pil_u8(frames, new_h, new_w, "bicubic").permute(0, 3, 1, 2).to(torch.int32)]
    if mode == "crop" and new_h > 518:
        y0 = (new_h - 518) // 2
        x = x[:, :, y0:y0 + 518]
    if mode == "pad":
        hp, wp = 518 - x.shape[2], 518 - x.shape[3]
        if hp > 0 or wp > 0:
            x = torch.nn.functional.pad(x, (wp // 2, wp - wp // 2, hp // 2, hp - hp // 2), mode="constant", value=1.0)
    return x.contiguous()


def decode_rgb(path) -> "np.ndarray":
    """load_fn.py:138-149: open an image file, composite an alpha channel onto white, convert to RGB -> uint8 [H, W, 3] (host)."""
    import numpy as np
    from PIL import Image

    img = Image.open(path)
    if img.mode == "RGBA":
        background = Image.new("RGBA", img.size, (255, 255, 255, 255))
        img = Image.alpha_composite(background, img)
    return np.asarray(img.convert("RGB"))


def load_and_preprocess_images(image_path_list, mode: str = "crop", device=None) -> torch.Tensor:
    """Drop-in for `vggt.utils.load_fn.load_and_preprocess_images` (third_party/vggt/vggt/utils/load_fn.py:95-230; call site
    unified_loop_consistency.py:348): the files are decoded on the host (PIL), everything after — Pillow-exact BICUBIC
    resize, ToTensor, centre crop / white padding, padding of differently shaped images to the largest — runs on the device.
    Returns float32 [N, 3, H, W] ON THE DEVICE (the caller's `.to(device)` is then a no-op).  Same errors as the reference."""
    if len(image_path_list) == 0:
        raise ValueError("At least 1 image is required")
    if mode not in ("crop", "pad"):
        raise ValueError("Mode must be either 'crop' or 'pad'")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None or dev.type != "cuda":
        raise RuntimeError("evoworld_b200 load_and_preprocess_images needs a CUDA device (no CPU fallback)")
    decoded = [decode_rgb(p) for p in image_path_list]
    out = [None] * len(decoded)
    by_size: dict = {}
    for i, a in enumerate(decoded):
        by_size.setdefault(a.shape[:2], []).append(i)
    import numpy as np

    for idx in by_size.values():          # one resize call per source size (the loop's frames all share one)
        batch = torch.from_numpy(np.stack([decoded[i] for i in idx])).to(dev, non_blocking=True)
        res = vggt_preprocess_u8(batch, mode)
        for j, i in enumerate(idx):
            out[i] = res[j]
    shapes = {(t.shape[1], t.shape[2]) for t in out}
    if len(shapes) > 1:
        print(f"Warning: Found images with different shapes: {shapes}")
        mh, mw = max(s[0] for s in shapes), max(s[1] for s in shapes)
        padded = []
        for t in out:
            hp, wp = mh - t.shape[1], mw - t.shape[2]
            if hp > 0 or wp > 0:
                t = torch.nn.functional.pad(t, (wp // 2, wp - wp // 2, hp // 2, hp - hp // 2), mode="constant", value=1.0)
            padded.append(t)
        out = padded
    return torch.stack(out)
