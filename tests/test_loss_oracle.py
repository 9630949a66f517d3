"""Pins oracle/loss_oracle.py (L1 + SSIM with the fused-ssim conventions) against torch autograd of the same formula written with
conv2d, and against central finite differences."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import loss_oracle as lo  # noqa: E402


def _images(h=37, w=45, seed=0):
    rng = np.random.default_rng(seed)
    y = rng.uniform(0, 1, (h, w, 3))
    x = np.clip(y + 0.1 * rng.normal(size=(h, w, 3)), 0, 1.2)
    return x, y


def _torch_loss(x, y, l1w, sw):
    win = torch.tensor(lo.gaussian_window(), dtype=torch.float64)
    k = (win[:, None] * win[None, :])[None, None].repeat(3, 1, 1, 1)
    X, Y = (t.permute(2, 0, 1)[None] for t in (x, y))
    conv = lambda t: torch.nn.functional.conv2d(t, k, padding=5, groups=3)  # noqa: E731
    mu1, mu2 = conv(X), conv(Y)
    s1, s2, s12 = conv(X * X) - mu1 * mu1, conv(Y * Y) - mu2 * mu2, conv(X * Y) - mu1 * mu2
    m = ((2 * mu1 * mu2 + lo.C1) * (2 * s12 + lo.C2)) / ((mu1 * mu1 + mu2 * mu2 + lo.C1) * (s1 + s2 + lo.C2))
    ssim = m[:, :, 5:-5, 5:-5].mean()
    return l1w * (x - y).abs().mean() + sw * (1 - ssim), ssim


def test_loss_and_gradient_match_torch_autograd():
    x, y = _images()
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    loss_t, ssim_t = _torch_loss(tx, torch.tensor(y, dtype=torch.float64), 0.8, 0.2)
    loss_t.backward()
    loss, l1, ssim, grad = lo.loss_and_gradient(x, y, 0.8, 0.2)
    assert abs(loss - float(loss_t)) <= 1e-12 and abs(ssim - float(ssim_t)) <= 1e-12
    assert np.allclose(grad, tx.grad.numpy(), rtol=1e-9, atol=1e-14)


def test_gradient_matches_finite_differences():
    x, y = _images(23, 19, seed=3)
    _, _, _, grad = lo.loss_and_gradient(x, y, 0.8, 0.2)
    rng = np.random.default_rng(1)
    for _ in range(12):
        i, j, c = rng.integers(0, 23), rng.integers(0, 19), rng.integers(0, 3)
        e = np.zeros_like(x)
        e[i, j, c] = 1e-6
        fd = (lo.loss_and_gradient(x + e, y, 0.8, 0.2)[0] - lo.loss_and_gradient(x - e, y, 0.8, 0.2)[0]) / 2e-6
        assert abs(fd - grad[i, j, c]) <= 1e-6 * max(1.0, abs(fd)) + 5e-9


def test_identical_images_give_ssim_one_and_float32_agrees():
    x, y = _images()
    assert abs(lo.loss_and_gradient(y, y)[2] - 1.0) <= 1e-12
    a = lo.loss_and_gradient(x, y, dtype=np.float32)
    b = lo.loss_and_gradient(x, y)
    assert abs(a[0] - b[0]) <= 1e-6 and np.abs(a[3] - b[3]).max() <= 1e-7
