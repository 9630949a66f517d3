"""CPU restatement (numpy, float32) of the optimizer step of the Gaussian parameters -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module; the product (3dgrut_b200/) never does.

Follows
  * selective Adam: threedgrut/optimizers/optimizers.cu:49-83 (kernel), threedgrut/optimizers/__init__.py:86-124 (step):
        m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p += -lr m / (sqrt(v) + eps)      on rows with visibility, no bias correction
  * Adam: torch.optim.Adam as the reference constructs it (threedgrut/model/model.py:807-810: lr per group, eps, default betas,
    no weight decay, no amsgrad):  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
  * activation chain rule the reference leaves to autograd (threedgrut/model/model.py:102-118, utils/misc.py:46-50):
        density = sigmoid(raw), scale = exp(raw), rotation = torch.nn.functional.normalize(raw) (eps 1e-12),
        features = cat(features_albedo [N,3], features_specular [N,45])                      (model.py:94-96)
Pinned by tests/test_adam_oracle.py against torch autograd + torch.optim.Adam on the CPU (the reference's own dependency)."""
import numpy as np

GROUPS = ("positions", "density", "rotation", "scale", "features_albedo", "features_specular")
WIDTHS = (3, 1, 4, 3, 3, 45)

f32 = np.float32


def raw_gradients(params, d_particles, d_sph):
    """params: dict of raw (pre-activation) arrays; d_particles [N,12], d_sph [N,48] w.r.t. the activated values.
    Returns the dict of gradients w.r.t. the raw parameters (float32)."""
    dp = np.asarray(d_particles, f32)
    ds = np.asarray(d_sph, f32)
    raw_d = np.asarray(params["density"], f32)
    raw_s = np.asarray(params["scale"], f32)
    raw_r = np.asarray(params["rotation"], f32)
    s = (f32(1) / (f32(1) + np.exp(-raw_d))).astype(f32)
    norm = np.maximum(np.sqrt((raw_r * raw_r).sum(1, keepdims=True, dtype=f32)), f32(1e-12)).astype(f32)
    q = (raw_r / norm).astype(f32)
    gq = dp[:, 4:8]
    dot = (q * gq).sum(1, keepdims=True, dtype=f32)
    return {
        "positions": dp[:, 0:3].copy(),
        "density": (dp[:, 3:4] * s * (f32(1) - s)).astype(f32),
        "rotation": ((gq - q * dot) / norm).astype(f32),
        "scale": (dp[:, 8:11] * np.exp(raw_s)).astype(f32),
        "features_albedo": ds[:, 0:3].copy(),
        "features_specular": ds[:, 3:48].copy(),
    }


def adam_update(p, g, m, v, lr, b1=0.9, b2=0.999, eps=1e-15, step=1, selective=False, visibility=None):
    """One step on one tensor; returns (p, m, v) as new float32 arrays.  visibility: [N] truthy flags (selective mode)."""
    p, g, m, v = (np.asarray(a, f32) for a in (p, g, m, v))
    b1, b2, lr, eps = f32(b1), f32(b2), f32(lr), f32(eps)
    m_new = (b1 * m + (f32(1) - b1) * g).astype(f32)
    v_new = (b2 * v + (f32(1) - b2) * g * g).astype(f32)
    if selective:
        bc1, bc2s = f32(1), f32(1)
    else:
        bc1, bc2s = f32(1.0 - float(b1) ** step), f32(np.sqrt(1.0 - float(b2) ** step))
    p_new = (p - (lr / bc1) * m_new / (np.sqrt(v_new) / bc2s + eps)).astype(f32)
    if selective and visibility is not None:
        keep = np.asarray(visibility).reshape(-1).astype(bool)
        shape = (-1,) + (1,) * (p.ndim - 1)
        k = keep.reshape(shape)
        p_new, m_new, v_new = np.where(k, p_new, p), np.where(k, m_new, m), np.where(k, v_new, v)
    return p_new.astype(f32), m_new.astype(f32), v_new.astype(f32)


def gaussian_adam_step(params, moments_m, moments_v, lrs, d_particles, d_sph, b1=0.9, b2=0.999, eps=1e-15, step=1, selective=False,
                       visibility=None):
    """The fused step: chain rule + Adam on all six groups.  Returns new (params, m, v) dicts."""
    grads = raw_gradients(params, d_particles, d_sph)
    out_p, out_m, out_v = {}, {}, {}
    for name in GROUPS:
        out_p[name], out_m[name], out_v[name] = adam_update(params[name], grads[name], moments_m[name], moments_v[name], lrs[name], b1, b2, eps,
                                                            step, selective, visibility)
    return out_p, out_m, out_v
