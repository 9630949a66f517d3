"""Pins oracle/adam_oracle.py against the reference's own dependency: torch autograd through the reference's activations
(threedgrut/model/model.py:102-118, utils/misc.py:46-50) + torch.optim.Adam on the CPU, and against the selective rule restated
from threedgrut/optimizers/optimizers.cu:66-80."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import adam_oracle as ao  # noqa: E402


def _state(n=257, seed=0):
    rng = np.random.default_rng(seed)
    params = {
        "positions": rng.normal(size=(n, 3)), "density": rng.normal(size=(n, 1)), "rotation": rng.normal(size=(n, 4)),
        "scale": rng.normal(-3, 0.7, size=(n, 3)), "features_albedo": rng.normal(size=(n, 3)), "features_specular": 0.1 * rng.normal(size=(n, 45)),
    }
    params = {k: v.astype(np.float32) for k, v in params.items()}
    dp = rng.normal(size=(n, 12)).astype(np.float32)
    ds = rng.normal(size=(n, 48)).astype(np.float32)
    return params, dp, ds


LRS = dict(positions=1.6e-4, density=0.05, rotation=1e-3, scale=5e-3, features_albedo=2.5e-3, features_specular=1.25e-4)


def _torch_reference(params, grads_seq, steps):
    """Three steps of torch.optim.Adam over autograd of the activation chain, with the given per-step (dp, ds)."""
    leaves = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
    opt = torch.optim.Adam([{"params": [leaves[k]], "lr": LRS[k]} for k in ao.GROUPS], lr=0.0, eps=1e-15)
    for dp, ds in grads_seq[:steps]:
        opt.zero_grad()
        act = torch.cat([leaves["positions"], torch.sigmoid(leaves["density"]), torch.nn.functional.normalize(leaves["rotation"]),
                         torch.exp(leaves["scale"]), torch.zeros_like(leaves["density"])], 1)
        feat = torch.cat([leaves["features_albedo"], leaves["features_specular"]], 1)
        ((act * torch.tensor(dp)).sum() + (feat * torch.tensor(ds)).sum()).backward()
        opt.step()
    return {k: v.detach().numpy() for k, v in leaves.items()}


def test_chain_rule_matches_autograd():
    params, dp, ds = _state()
    leaves = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
    act = torch.cat([leaves["positions"], torch.sigmoid(leaves["density"]), torch.nn.functional.normalize(leaves["rotation"]),
                     torch.exp(leaves["scale"]), torch.zeros_like(leaves["density"])], 1)
    feat = torch.cat([leaves["features_albedo"], leaves["features_specular"]], 1)
    ((act * torch.tensor(dp)).sum() + (feat * torch.tensor(ds)).sum()).backward()
    got = ao.raw_gradients(params, dp, ds)
    for k in ao.GROUPS:
        assert np.allclose(got[k], leaves[k].grad.numpy(), rtol=2e-6, atol=1e-7), k


def test_three_adam_steps_match_torch_optim():
    params, _, _ = _state()
    rng = np.random.default_rng(5)
    seq = [(rng.normal(size=(257, 12)).astype(np.float32), rng.normal(size=(257, 48)).astype(np.float32)) for _ in range(3)]
    want = _torch_reference(params, seq, 3)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v = {k: np.zeros_like(vv) for k, vv in params.items()}
    for t, (dp, ds) in enumerate(seq, 1):
        p, m, v = ao.gaussian_adam_step(p, m, v, LRS, dp, ds, eps=1e-15, step=t)
    for k in ao.GROUPS:
        assert np.allclose(p[k], want[k], rtol=1e-5, atol=1e-6), k


def test_selective_rule_and_mask():
    rng = np.random.default_rng(2)
    p, g = rng.normal(size=(50, 3)).astype(np.float32), rng.normal(size=(50, 3)).astype(np.float32)
    m, v = rng.normal(size=(50, 3)).astype(np.float32) * 0.1, rng.uniform(0, 1, size=(50, 3)).astype(np.float32)
    vis = rng.uniform(size=50) > 0.4
    pn, mn, vn = ao.adam_update(p, g, m, v, 0.01, 0.9, 0.999, 1e-8, selective=True, visibility=vis)
    # restated literally from optimizers.cu:66-80 (float arithmetic: `1.0f - b2` is 0.00100004673, not 0.001)
    b1, b2, lr, eps = np.float32(0.9), np.float32(0.999), np.float32(0.01), np.float32(1e-8)
    m_ref = b1 * m + (np.float32(1) - b1) * g
    v_ref = b2 * v + (np.float32(1) - b2) * g * g
    p_ref = p - lr * m_ref / (np.sqrt(v_ref) + eps)
    assert np.allclose(pn[vis], p_ref[vis], rtol=1e-6) and np.allclose(mn[vis], m_ref[vis], rtol=1e-6) and np.allclose(vn[vis], v_ref[vis], rtol=1e-6)
    assert np.array_equal(pn[~vis], p[~vis]) and np.array_equal(mn[~vis], m[~vis]) and np.array_equal(vn[~vis], v[~vis])
