"""Metrics of propainter_b200/evaluate.py (CPU): PSNR and the skimage-style SSIM (uniform 65x65 windows, sample covariance)
against a direct numpy evaluation of the published definition (core/metrics.py:20-47 calls skimage's compare_ssim)."""
import numpy as np
import torch

from propainter_b200.evaluate import epe, fid_from_activations, psnr_frames, ssim_frames


def ssim_numpy(a, b, win=65, R=255.0):
    a, b = a.astype(np.float64), b.astype(np.float64)
    H, W, C = a.shape
    n = win * win
    vals = []
    for c in range(C):
        x, y = a[..., c], b[..., c]
        acc = []
        for i in range(0, H - win + 1, 7):                 # subsample window positions: the mean of S is what is compared
            for j in range(0, W - win + 1, 7):
                px, py = x[i:i + win, j:j + win].ravel(), y[i:i + win, j:j + win].ravel()
                ux, uy = px.mean(), py.mean()
                vx, vy = px.var(ddof=1), py.var(ddof=1)
                vxy = ((px - ux) * (py - uy)).sum() / (n - 1)
                c1, c2 = (0.01 * R) ** 2, (0.03 * R) ** 2
                acc.append(((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2)))
        vals.append(acc)
    return np.array(vals)


def test_psnr_and_ssim_definitions():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (2, 100, 121, 3), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-20, 21, a.shape), 0, 255).astype(np.uint8)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    ps = psnr_frames(ta, tb)
    for t in range(2):
        mse = ((a[t].astype(np.float64) - b[t].astype(np.float64)) ** 2).mean()
        assert abs(ps[t].item() - 20 * np.log10(255 / np.sqrt(mse))) < 1e-9
    assert psnr_frames(ta, ta)[0].item() == float("inf") and abs(ssim_frames(ta, ta)[0].item() - 1.0) < 1e-12
    # full-resolution map of S from the torch implementation, sampled at the same window positions as the numpy loop
    import torch.nn.functional as F
    x = ta[0].permute(2, 0, 1).double()[:, None]
    y = tb[0].permute(2, 0, 1).double()[:, None]
    n, cn = 65 * 65, 65 * 65 / (65 * 65 - 1.0)
    box = lambda z: F.avg_pool2d(z, 65, stride=1)
    ux, uy = box(x), box(y)
    vx, vy, vxy = cn * (box(x * x) - ux * ux), cn * (box(y * y) - uy * uy), cn * (box(x * y) - ux * uy)
    c1, c2 = 6.5025, 58.5225
    S = (((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2)))[:, 0, ::7, ::7]
    ref = ssim_numpy(a[0], b[0])
    assert np.abs(S.reshape(3, -1).numpy() - ref).max() < 1e-9
    assert abs(ssim_frames(ta[:1], tb[:1])[0].item() - float((((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))).flatten(1).mean(1).mean())) < 1e-12


def test_epe_and_fid():
    f = torch.zeros(1, 2, 4, 4)
    g = f.clone()
    g[:, 0] = 3
    g[:, 1] = 4
    assert abs(epe(f, g) - 5.0) < 1e-6
    rng = np.random.default_rng(1)
    a = rng.normal(size=(200, 8))
    assert abs(fid_from_activations(a, a)) < 1e-6 and fid_from_activations(a, a + 1.0) > 7.0
