"""The numpy restatement of Pillow's 8-bit bilinear resize (oracle/resize_np.py) against Pillow itself, bit for bit — the
oracle of the memory-panorama resize (dataset/CameraTrajDataset.py:597-600) is pinned by the library the reference calls."""
import numpy as np
import pytest

from oracle import resize_np as R

PIL = pytest.importorskip("PIL")


@pytest.mark.parametrize("H,W,h,w", [(100, 200, 58, 102), (37, 53, 11, 20), (40, 64, 40, 32), (31, 17, 64, 40), (50, 50, 50, 50),
                                     (125, 250, 72, 128), (9, 300, 3, 7)])
def test_restatement_equals_pillow(H, W, h, w):
    from PIL import Image

    rng = np.random.default_rng(H * 1000 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[: H // 3] = 255  # saturated and zero regions exercise the clipping
    img[-(H // 4):] = 0
    want = np.asarray(Image.fromarray(img).resize((w, h), Image.BILINEAR))
    got = R.resize_bilinear_u8(img, h, w)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_torchvision_resize_is_this_function():
    """transforms.Resize((h, w)) on a PIL image — the reference's call (CameraTrajDataset.py:597) — is PIL's bilinear resize."""
    tv = pytest.importorskip("torchvision")
    from PIL import Image
    from torchvision import transforms

    img = np.random.default_rng(0).integers(0, 256, (60, 120, 3), dtype=np.uint8)
    want = np.asarray(transforms.Resize((36, 64))(Image.fromarray(img)))
    assert np.array_equal(R.resize_bilinear_u8(img, 36, 64), want)


def test_coefficient_tables():
    b, k, ks = R.coeffs(2000, 1024)
    assert ks == 5 and b.shape == (1024, 2) and k.shape == (1024, 5)
    assert (k.sum(1) - (1 << R.PRECISION_BITS)).__abs__().max() <= 3  # rounded coefficients sum to 1.0 within a few ulps
    b, k, ks = R.coeffs(10, 40)  # up-scaling: two-tap interpolation, support 1
    assert ks == 3 and int(b[:, 1].max()) <= 3


@pytest.mark.parametrize("n_in,n_out", [(2000, 1024), (1000, 576), (37, 11), (17, 64), (300, 7), (50, 50), (64, 32), (1024, 2000)])
def test_product_tables_equal_the_restatement(n_in, n_out):
    """evoworld_b200/image_ops.py::pil_resize_tables (vectorised, what the kernel consumes) against the scalar restatement."""
    from evoworld_b200.image_ops import pil_resize_tables

    b0, k0, s0 = R.coeffs(n_in, n_out)
    b1, k1, s1 = pil_resize_tables(n_in, n_out)
    assert s0 == s1 and np.array_equal(b0, b1) and np.array_equal(k0, k1)
