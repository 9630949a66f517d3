"""CPU restatement (numpy float64 / float32) of the image loss of the training step -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module; the product (3dgrut_b200/) never does.

    loss = lambda_l1 * mean|x - y|  +  lambda_ssim * (1 - SSIM(x, y))                      threedgrut/trainer.py:698-739

  * l1_loss: threedgrut/model/losses.py:20-21.
  * ssim: threedgrut/model/losses.py:30-33 calls fused_ssim(img1, img2, padding="valid") of the third-party package
    fused-ssim @ git 1272e21a282342e89537159e4bad508b19b34157 (requirements_extra.txt:2), which is NOT under /root/reference.
    Its published algorithm (Wang et al. 2004 as used by 3DGS): per channel, 11x11 Gaussian window (sigma 1.5, separable, normalised),
    zero padding for the convolutions, C1 = 0.01^2, C2 = 0.03^2,
        mu1 = w*x, mu2 = w*y, s1 = w*x^2 - mu1^2, s2 = w*y^2 - mu2^2, s12 = w*xy - mu1 mu2,
        map = (2 mu1 mu2 + C1)(2 s12 + C2) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2)),
    padding="valid" crops 5 pixels on every side before the mean.
Parity is UNPINNED against the package itself (absent, no network); the restatement is pinned against torch autograd of the same
formula written with conv2d (tests/test_loss_oracle.py) and against finite differences."""
import numpy as np

C1, C2 = 0.01 ** 2, 0.03 ** 2


def gaussian_window(size=11, sigma=1.5, dtype=np.float64):
    x = np.arange(size, dtype=np.float64) - size // 2
    g = np.exp(-(x ** 2) / (2.0 * sigma ** 2))
    return (g / g.sum()).astype(dtype)


def _conv(img, w):
    """separable 'same' correlation with zero padding over the two leading axes of [H,W,C]"""
    from scipy.ndimage import correlate1d

    out = correlate1d(img, w, axis=0, mode="constant", cval=0.0)
    return correlate1d(out, w, axis=1, mode="constant", cval=0.0)


def ssim_map(x, y, dtype=np.float64):
    x, y = np.asarray(x, dtype), np.asarray(y, dtype)
    w = gaussian_window(dtype=dtype)
    mu1, mu2 = _conv(x, w), _conv(y, w)
    s1 = _conv(x * x, w) - mu1 * mu1
    s2 = _conv(y * y, w) - mu2 * mu2
    s12 = _conv(x * y, w) - mu1 * mu2
    a, b = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    c, d = mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
    return a * b / (c * d), (mu1, mu2, s1, s2, s12, a, b, c, d)


def loss_and_gradient(x, y, lambda_l1=0.8, lambda_ssim=0.2, dtype=np.float64):
    """x (prediction), y (target): [H,W,3].  Returns (loss, l1, ssim, dloss/dx [H,W,3])."""
    x, y = np.asarray(x, dtype), np.asarray(y, dtype)
    H, W, C = x.shape
    m, (mu1, mu2, s1, s2, s12, a, b, c, d) = ssim_map(x, y, dtype)
    valid = np.zeros((H, W, 1), dtype)
    if H > 10 and W > 10:
        valid[5:-5, 5:-5] = 1.0
    count = max(float(valid.sum()) * C, 1.0)
    ssim = float((m * valid).sum() / count)
    l1 = float(np.abs(x - y).mean())
    loss = lambda_l1 * l1 + lambda_ssim * (1.0 - ssim)
    # d(1 - ssim)/dmap = -valid / count; chain through mu1, s1, s12 (functions of x through the window)
    g = -lambda_ssim * valid / count
    dm_dmu1 = (2 * mu2 * b) / (c * d) - (2 * mu1 * a * b) / (c * c * d) - 2 * mu1 * (-(a * b) / (c * d * d)) - mu2 * (2 * a) / (c * d)
    # s1 = w*x^2 - mu1^2 and s12 = w*xy - mu1 mu2 carry mu1: the terms with -2 mu1 d/ds1 and -mu2 d/ds12 are folded in above
    dm_ds1 = -(a * b) / (c * d * d)
    dm_ds12 = (2 * a) / (c * d)
    w = gaussian_window(dtype=dtype)
    grad = _conv(g * dm_dmu1, w) + 2 * x * _conv(g * dm_ds1, w) + y * _conv(g * dm_ds12, w)
    grad = grad + lambda_l1 * np.sign(x - y) / x.size
    return loss, l1, ssim, grad
