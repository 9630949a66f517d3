"""GPU parity of the image loss kernels (3dgrut_b200/csrc/gut_loss.cu) against oracle/loss_oracle.py (float64).
Tolerance: loss terms 1e-6 absolute; gradient |diff| <= 2e-6 * max|grad| + 1e-10 (121-tap fp32 convolutions)."""
import numpy as np
import pytest

from oracle import loss_oracle as lo

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("size", [(64, 64), (75, 53), (37, 130)])
@pytest.mark.parametrize("weights", [(0.8, 0.2), (0.0, 1.0), (1.0, 0.0)])
def test_image_loss_matches_oracle(size, weights):
    import losses

    h, w = size
    rng = np.random.default_rng(h * 1000 + w)
    y = rng.uniform(0, 1, (h, w, 3)).astype(np.float32)
    x = np.clip(y + 0.1 * rng.normal(size=(h, w, 3)), 0, 1.2).astype(np.float32)
    pred = np.concatenate([x, rng.uniform(0, 1, (h, w, 1)).astype(np.float32)], -1)
    dev = torch.device("cuda", 0)
    loss, l1, ssim, d = losses.image_loss(torch.from_numpy(pred).to(dev), torch.from_numpy(y).to(dev), *weights)
    ref_loss, ref_l1, ref_ssim, ref_grad = lo.loss_and_gradient(x.astype(np.float64), y.astype(np.float64), *weights)
    assert abs(float(l1) - ref_l1) <= 1e-6 and abs(float(ssim) - ref_ssim) <= 1e-6 and abs(float(loss) - ref_loss) <= 1e-6
    d = d.cpu().numpy()
    assert np.all(d[..., 3] == 0)
    err = np.abs(d[..., :3] - ref_grad).max()
    print(f"[parity] loss {size} {weights}: max |grad diff| {err:.3e} (max |grad| {np.abs(ref_grad).max():.3e})")
    assert err <= 2e-6 * np.abs(ref_grad).max() + 1e-10


def test_image_loss_rejects_cpu_tensors():
    import losses

    with pytest.raises(RuntimeError):
        losses.image_loss(torch.zeros(16, 16, 4), torch.zeros(16, 16, 3))
