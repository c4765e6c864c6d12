"""CPU check of the resizing tables (host functions of the C-ABI library, no GPU needed): the integer passes the device
kernels run are emulated here with numpy from the very tables pp_resample_* return and compared with the libraries the
reference calls -- PIL.Image.resize (BICUBIC default and NEAREST) and cv2.resize (INTER_LINEAR).  The kernels themselves
are compared with the same libraries on the GPU (tests/test_gpu_ops.py::test_resize_kernels)."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")
Image = pytest.importorskip("PIL.Image")


def _tables():
    import __graft_entry__ as g
    g.build()
    from propainter_b200 import ops
    return ops.resample_tables


def emu_bicubic(img, size, resample_tables):
    H, W, _ = img.shape
    Wo, Ho = size
    x = img.astype(np.int64)
    if Wo != W:
        b, kk = (t.numpy() for t in resample_tables("bicubic", W, Wo))
        out = np.empty((H, Wo, 3), np.int64)
        for xx in range(Wo):
            lo, n = b[xx]
            out[:, xx] = ((1 << 21) + (x[:, lo:lo + n] * kk[xx, :n, None].astype(np.int64)).sum(1)) >> 22
        x = np.clip(out, 0, 255)
    if Ho != H:
        b, kk = (t.numpy() for t in resample_tables("bicubic", H, Ho))
        out = np.empty((Ho, x.shape[1], 3), np.int64)
        for yy in range(Ho):
            lo, n = b[yy]
            out[yy] = ((1 << 21) + (x[lo:lo + n] * kk[yy, :n, None, None].astype(np.int64)).sum(0)) >> 22
        x = np.clip(out, 0, 255)
    return x.astype(np.uint8)


def emu_linear_cv(img, size, resample_tables):
    H, W, _ = img.shape
    Wo, Ho = size
    xo, xa = (t.numpy().astype(np.int64) for t in resample_tables("linear_cv", W, Wo, None, True))
    yo, ya = (t.numpy().astype(np.int64) for t in resample_tables("linear_cv", H, Ho, None, False))
    s = img.astype(np.int64)
    x1 = np.minimum(xo + 1, W - 1)
    hrow = s[:, xo] * xa[None, :, 0, None] + s[:, x1] * xa[None, :, 1, None]          # [H, Wo, 3] ints
    y0, y1 = np.clip(yo, 0, H - 1), np.clip(yo + 1, 0, H - 1)
    S0, S1 = hrow[y0], hrow[y1]
    b0, b1 = ya[:, 0, None, None], ya[:, 1, None, None]
    return ((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


@pytest.mark.parametrize("H,W,size", [(243, 437, (432, 240)), (240, 432, (436, 246)), (100, 150, (72, 48)), (64, 64, (64, 40)), (37, 53, (160, 96))])
def test_bicubic_tables_reproduce_pil(H, W, size):
    rt = _tables()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img, mode="RGB").resize(size))
    assert np.array_equal(emu_bicubic(img, size, rt), ref)


@pytest.mark.parametrize("H,W,size", [(243, 437, (432, 240)), (100, 150, (72, 48)), (37, 53, (160, 96)), (720, 1283, (1280, 720))])
def test_nearest_tables_reproduce_pil(H, W, size):
    rt = _tables()
    rng = np.random.default_rng(1)
    m = (rng.integers(0, 2, (H, W), dtype=np.uint8) * 255)
    ref = np.array(Image.fromarray(m, mode="L").resize(size, Image.NEAREST))
    (ix,), (iy,) = rt("nearest", W, size[0]), rt("nearest", H, size[1])
    assert np.array_equal(m[iy.numpy()][:, ix.numpy()], ref)


@pytest.mark.parametrize("H,W,size", [(240, 432, (437, 243)), (240, 432, (432, 241)), (64, 96, (101, 77)), (120, 200, (150, 90))])
def test_linear_tables_reproduce_cv2(H, W, size):
    rt = _tables()
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = cv2.resize(img, size)
    got = emu_linear_cv(img, size, rt)
    d = np.abs(got.astype(int) - ref.astype(int))
    # OpenCV builds may route 8-bit INTER_LINEAR through vendor code paths; the portable fixed-point path is bit-exact
    assert d.max() <= 1 and (d > 0).mean() < 1e-3, (d.max(), (d > 0).mean())
