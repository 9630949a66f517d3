"""CPU checks of the brute-force 3DGRT oracle (grt_oracle_* in oracle/gut_oracle.c):
its adjoint against torch autograd (fp64, same ordered hit lists, including the reference's quirk that the backward
re-trace excludes the last processed hit) and basic invariants of the proxies."""
import numpy as np
import pytest

import scenes
from helpers import rel_l2
from oracle import gut_oracle as go

torch = pytest.importorskip("torch")


def test_proxy_scale_matches_closed_form():
    cfg = go.grt_config()
    sc = scenes.scene_c1(n=50)
    kscl, bb = go.grt_proxies(cfg, sc.particles, clamping=True)
    dns = sc.particles[:, 3].astype(np.float64)
    minr = np.minimum(0.0113 / dns, 0.97)
    r = (np.log(minr) / (-4.5 / 81.0)) ** 0.25  # kernelScale, degree 4 (particlePrimitives.cu:27-51)
    assert np.allclose(kscl, sc.particles[:, 8:11] * r[:, None], rtol=2e-5)
    assert np.all(bb[:3] <= sc.particles[:, 0:3].min(0)) and np.all(bb[3:] >= sc.particles[:, 0:3].max(0))


def _candidates(sc, kscl, o, d, tmin, tmax):
    """numpy restatement of the candidate rule, used only to order hits for the autograd model"""
    P = sc.particles.astype(np.float64)
    out = []
    for i in range(P.shape[0]):
        r, x, y, z = P[i, 4:8]
        Rt = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y)],
                       [2 * (x * y - r * z), 1 - 2 * (x * x + z * z), 2 * (y * z + r * x)],
                       [2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)]])
        oi = (Rt @ (o - P[i, 0:3])) / kscl[i]
        di = (Rt @ d) / kscl[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            t0, t1 = (-1 - oi) / di, (1 - oi) / di
        tin, tout = max(tmin, np.minimum(t0, t1).max()), min(tmax, np.maximum(t0, t1).min())
        if not tin <= tout:
            continue
        t = -(oi @ di) / (di @ di)
        if t > tmin and t < tmax:
            out.append((t, i))
    return sorted(out)


def test_grt_oracle_backward_equals_autograd():
    sc = scenes.scene_c1(n=40, seed=5, width=20, height=16)
    cfg = go.grt_config()
    c2w = np.asarray(sc.camera(2, 7), np.float32)
    ro, rd = sc.rays()
    rgb, alpha, dist, hits, vis = go.grt_trace(cfg, sc.particles, sc.sph, 3, ro[0], rd[0], c2w)
    assert hits.max() >= 3
    rng = np.random.default_rng(0)
    d_rgb = rng.normal(size=rgb.shape).astype(np.float32)
    d_alpha = rng.normal(size=alpha.shape).astype(np.float32)
    d_dist = (0.2 * rng.normal(size=alpha.shape)).astype(np.float32)
    dp, ds = go.grt_trace_bwd(cfg, sc.particles, sc.sph, 3, ro[0], rd[0], c2w, rgb, alpha, dist, d_rgb, d_alpha, d_dist)

    from test_oracle_autograd import _rot_rows, _sh

    f64 = torch.float64
    Pt = torch.tensor(sc.particles, dtype=f64)
    pos, dns, quat, scl = (Pt[:, 0:3].clone().requires_grad_(True), Pt[:, 3].clone().requires_grad_(True),
                           Pt[:, 4:8].clone().requires_grad_(True), Pt[:, 8:11].clone().requires_grad_(True))
    sph = torch.tensor(sc.sph, dtype=f64).reshape(-1, 16, 3).clone().requires_grad_(True)
    kscl, bb = go.grt_proxies(cfg, sc.particles)
    R, t = c2w[:3, :3].astype(np.float64), c2w[:3, 3].astype(np.float64)
    loss = torch.zeros((), dtype=f64)
    ro2, rd2 = ro[0].reshape(-1, 3).astype(np.float64), rd[0].reshape(-1, 3).astype(np.float64)
    last = dist.reshape(-1, 2)[:, 1]
    for k in range(ro2.shape[0]):
        o, d = R @ ro2[k] + t, R @ rd2[k]
        with np.errstate(divide="ignore"):
            t0s, t1s = (bb[:3] - o) / d, (bb[3:] - o) / d
        tmin, tmax = max(0.0, np.minimum(t0s, t1s).max()), np.maximum(t0s, t1s).min()
        if not tmin <= tmax:
            continue
        ot, dt = torch.tensor(o, dtype=f64), torch.tensor(d, dtype=f64)
        T = torch.ones((), dtype=f64)
        C, D = torch.zeros(3, dtype=f64), torch.zeros((), dtype=f64)
        for (ts, i) in _candidates(sc, kscl.astype(np.float64), o, d, max(0.0, tmin - 1e-9), tmax + 1e-9):
            if float(T) <= 1e-3:
                break
            Rr = _rot_rows(quat[i])
            gro = (Rr @ (ot - pos[i])) / scl[i]
            grdu = (Rr @ dt) / scl[i]
            grd = grdu / grdu.norm()
            gray = torch.linalg.cross(grd, gro).pow(2).sum()
            gres = torch.exp(-0.0555555555556 * gray * gray)
            a = torch.clamp(gres * dns[i], max=0.99)
            if not (float(gres) > 0.0113 and float(a) > 1.0 / 255.0):
                continue
            tt = (scl[i] * grd * (-(grd * gro).sum())).norm()
            col = torch.clamp(_sh(sph[i], dt), min=0.0)
            # the reference's backward re-trace stops strictly before the last processed hit (referenceBwdOptix.cu:115,125):
            # that hit's own parameters receive no gradient, earlier hits still see it through the transmittance
            if ts >= float(last[k]) * (1.0 - 1e-5):
                tt, col, a = tt.detach(), col.detach(), a.detach()
            w = a * T
            C = C + w * col
            D = D + w * tt
            T = T * (1 - a)
        loss = loss + (C * torch.tensor(d_rgb.reshape(-1, 3)[k], dtype=f64)).sum() + (1 - T) * float(d_alpha.reshape(-1)[k]) \
            + D * float(d_dist.reshape(-1)[k])
    loss.backward()
    assert rel_l2(pos.grad.numpy(), dp[:, 0:3]) < 2e-3
    assert rel_l2(dns.grad.numpy(), dp[:, 3]) < 2e-3
    assert rel_l2(quat.grad.numpy(), dp[:, 4:8]) < 2e-3
    assert rel_l2(scl.grad.numpy(), dp[:, 8:11]) < 2e-3
    assert rel_l2(sph.grad.numpy().reshape(-1, 48), ds) < 2e-3
