"""evw_resize_pil_u8 (csrc/resize.cu) against Pillow itself: bit-exact for down- and up-scaling, odd sizes, the panorama size."""
import numpy as np
import pytest
import torch

from evoworld_b200.image_ops import resize_pil_u8
from oracle import resize_np as R

pytestmark = pytest.mark.gpu
PIL = pytest.importorskip("PIL")


def _pil(img, h, w):
    from PIL import Image

    return np.asarray(Image.fromarray(img).resize((w, h), Image.BILINEAR))


@pytest.mark.parametrize("N,H,W,h,w", [(2, 100, 200, 58, 102), (1, 37, 53, 11, 20), (3, 40, 64, 40, 32), (1, 31, 17, 64, 40),
                                       (2, 50, 50, 50, 50), (1, 9, 300, 3, 7), (2, 33, 35, 33, 70)])
def test_resize_equals_pillow(N, H, W, h, w, cuda_device, built_lib):
    rng = np.random.default_rng(N * 7 + H)
    imgs = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    imgs[:, : H // 3] = 255
    imgs[:, -(H // 4):] = 0
    got = resize_pil_u8(torch.from_numpy(imgs).to(cuda_device), h, w).cpu().numpy()
    for i in range(N):
        assert np.array_equal(got[i], _pil(imgs[i], h, w)), f"image {i}"
        assert np.array_equal(got[i], R.resize_bilinear_u8(imgs[i], h, w))


def test_memory_panoramas(cuda_device, built_lib):
    """The reference's case: 24 reprojected panoramas 1000 x 2000 -> 576 x 1024 (unified_loop_consistency.py:422)."""
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, (24, 1000, 2000, 3), dtype=np.uint8)
    imgs[:, 400:600, 500:900] = 0  # holes of the splat
    x = torch.from_numpy(imgs).to(cuda_device)
    got = resize_pil_u8(x, 576, 1024)
    assert got.shape == (24, 576, 1024, 3)
    g = got.cpu().numpy()
    for i in (0, 11, 23):
        assert np.array_equal(g[i], _pil(imgs[i], 576, 1024))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        resize_pil_u8(x, 576, 1024)
    e1.record()
    torch.cuda.synchronize()
    print(f"24 panoramas 1000x2000 -> 576x1024: {e0.elapsed_time(e1) / 5 * 1e3:.0f} us")
    single = resize_pil_u8(x[3], 576, 1024)
    assert torch.equal(single, got[3])
    with pytest.raises(ValueError):
        resize_pil_u8(x.float(), 576, 1024)
    with pytest.raises(RuntimeError):
        resize_pil_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), 4, 4)


@pytest.mark.parametrize("N,H,W,h,w", [(2, 384, 512, 392, 518), (1, 37, 53, 11, 20), (1, 31, 17, 64, 40), (2, 9, 300, 3, 7)])
def test_bicubic_resize_equals_pillow(N, H, W, h, w, cuda_device, built_lib):
    """The same kernels with Pillow's BICUBIC tables (negative coefficients): the resize in front of VGGT (load_fn.py:166)."""
    from PIL import Image

    rng = np.random.default_rng(N * 11 + H)
    imgs = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    imgs[:, : H // 3] = 255
    imgs[:, -(H // 4):] = 0
    got = resize_pil_u8(torch.from_numpy(imgs).to(cuda_device), h, w, "bicubic").cpu().numpy()
    for i in range(N):
        assert np.array_equal(got[i], np.asarray(Image.fromarray(imgs[i]).resize((w, h), Image.BICUBIC))), f"image {i}"


def test_vggt_preprocess(cuda_device, built_lib):
    from evoworld_b200.image_ops import vggt_preprocess_u8

    frames = np.random.default_rng(9).integers(0, 256, (5, 384, 512, 3), dtype=np.uint8)
    got = vggt_preprocess_u8(torch.from_numpy(frames).to(cuda_device))
    assert got.shape == (5, 3, 392, 518) and got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy(), R.vggt_preprocess(frames))


@pytest.mark.parametrize("names,mode", [(("a", "b"), "crop"), (("a", "rgba"), "crop"), (("a", "tall", "wide"), "crop"),
                                        (("a", "tall", "wide"), "pad"), (("tall",), "crop")])
def test_load_and_preprocess_images(names, mode, cuda_device, built_lib, tmp_path, capsys):
    """The drop-in loader (files decoded on the host, resize / ToTensor / crop / pad on the device) against the numpy
    restatement of the reference loader (pinned against the reference itself in tests/test_resize_host.py), bit for bit."""
    from PIL import Image

    from evoworld_b200.image_ops import load_and_preprocess_images

    rng = np.random.default_rng(11)
    paths = []
    for n in names:
        H, W, ch = {"a": (384, 512, 3), "b": (384, 512, 3), "rgba": (384, 512, 4), "tall": (640, 400, 3), "wide": (200, 640, 3)}[n]
        paths.append(str(tmp_path / f"{n}.png"))
        Image.fromarray(rng.integers(0, 256, (H, W, ch), dtype=np.uint8), "RGBA" if ch == 4 else "RGB").save(paths[-1])
    got = load_and_preprocess_images(paths, mode=mode)
    want = R.load_and_preprocess_images_np(paths, mode=mode)
    assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)
