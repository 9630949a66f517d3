"""Independent check of the oracle's hand-derived adjoints (G7 + G8) against torch autograd in fp64.

The forward compositing is re-stated in differentiable torch (same sorted lists, same accept tests) and
d(loss)/d(pos, density, quat, scale, SH) from autograd must agree with oracle.render_backward.  The reference's
hand adjoint does not mask the min(0.99, .) clamp, so the scene keeps alpha below it."""
import numpy as np
import pytest

import scenes
from helpers import rel_l2
from oracle import gut_oracle as go

torch = pytest.importorskip("torch")

C0, C1 = 0.28209479177387814, 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277,
      -0.5900435899266435]


def _sh(c, d):
    x, y, z = d[0], d[1], d[2]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    b = [C0, -C1 * y, C1 * z, -C1 * x, C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy),
         C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy), C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
         C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy), C3[6] * x * (xx - 3 * yy)]
    return sum(b[k] * c[k] for k in range(16)) + 0.5


def _rot_rows(q):
    r, x, y, z = q[0], q[1], q[2], q[3]
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y)]),
        torch.stack([2 * (x * y - r * z), 1 - 2 * (x * x + z * z), 2 * (y * z + r * x)]),
        torch.stack([2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)])])


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_backward_equals_autograd(seed):
    sc = scenes.scene_c1(n=90, seed=seed + 20, width=48, height=40)
    cfg = go.default_config()
    pose = scenes.pose7_from_c2w(sc.camera(seed, 5))
    cam = go.make_camera(sc.width, sc.height, sc.fx, sc.fy, sc.cx, sc.cy, pose)
    ro, rd = sc.rays()
    pr, bn, rgba, dist, hits = go.forward_all(cfg, cam, ro, rd, sc.particles, sc.sph, 3)
    assert bn.sorted_values.size > 200
    rng = np.random.default_rng(seed)
    d_rgba = rng.normal(size=rgba.shape).astype(np.float32)
    d_dist = (0.2 * rng.normal(size=dist.shape)).astype(np.float32)
    dp, ds = go.render_backward(cfg, cam, ro, rd, sc.particles, sc.sph, 3, pr, bn, rgba, dist, d_rgba, d_dist)

    f64 = torch.float64
    P = torch.tensor(sc.particles, dtype=f64)
    pos = P[:, 0:3].clone().requires_grad_(True)
    dns = P[:, 3].clone().requires_grad_(True)
    quat = P[:, 4:8].clone().requires_grad_(True)
    scl = P[:, 8:11].clone().requires_grad_(True)
    sph = torch.tensor(sc.sph, dtype=f64).reshape(-1, 16, 3).clone().requires_grad_(True)
    _, inv, campos = go.sensor_matrices(cam)
    inv_t = torch.tensor(inv, dtype=f64)  # [4 cols, 3]
    campos_t = torch.tensor(campos, dtype=f64)
    W, H = sc.width, sc.height
    gx = (W + 15) // 16
    ro_t = torch.tensor(ro.reshape(-1, 3), dtype=f64)
    rd_t = torch.tensor(rd.reshape(-1, 3), dtype=f64)
    o_w = ro_t @ inv_t[:3] + inv_t[3]
    d_w = rd_t @ inv_t[:3]
    loss = torch.zeros((), dtype=f64)
    g_rgba = torch.tensor(d_rgba.reshape(-1, 4), dtype=f64)
    g_dist = torch.tensor(d_dist.reshape(-1), dtype=f64)
    visible = pr.tiles_count > 0
    rgb_p = {}
    for i in np.nonzero(visible)[0]:
        v = pos[i] - campos_t
        rgb_p[i] = torch.clamp(_sh(sph[i], v / v.norm()), min=0.0)
    for tile in range(bn.ranges.shape[0]):
        b, e = bn.ranges[tile]
        tx, ty = tile % gx, tile // gx
        ys, xs = np.meshgrid(np.arange(ty * 16, min(H, ty * 16 + 16)), np.arange(tx * 16, min(W, tx * 16 + 16)), indexing="ij")
        pix = torch.tensor((ys * W + xs).reshape(-1))
        o, d = o_w[pix], d_w[pix]
        T = torch.ones(len(pix), dtype=f64)
        alive = torch.ones(len(pix), dtype=torch.bool)
        C = torch.zeros((len(pix), 3), dtype=f64)
        D = torch.zeros(len(pix), dtype=f64)
        for k in range(b, e):
            i = int(bn.sorted_values[k])
            R = _rot_rows(quat[i])
            gro = ((o - pos[i]) @ R.T) / scl[i]
            grdu = (d @ R.T) / scl[i]
            grd = grdu / grdu.norm(dim=1, keepdim=True)
            gray = torch.linalg.cross(grd, gro).pow(2).sum(1)
            gres = torch.exp(-0.5 * gray)
            alpha = torch.clamp(gres * dns[i], max=0.99)
            t = (scl[i] * grd * (-(grd * gro).sum(1, keepdim=True))).norm(dim=1)
            acc = alive & (gres > 0.0113) & (alpha > 1.0 / 255.0) & (t > 0)
            w = torch.where(acc, alpha * T, torch.zeros_like(T))
            C = C + w[:, None] * rgb_p[i]
            D = D + w * t
            T = torch.where(acc, T * (1 - alpha), T)
            alive = alive & ~(T < 1e-4)
        loss = loss + (C * g_rgba[pix, :3]).sum() + ((1 - T) * g_rgba[pix, 3]).sum() + (D * g_dist[pix]).sum()
    loss.backward()
    assert rel_l2(pos.grad.numpy(), dp[:, 0:3]) < 2e-4
    assert rel_l2(dns.grad.numpy(), dp[:, 3]) < 2e-4
    assert rel_l2(quat.grad.numpy(), dp[:, 4:8]) < 2e-4
    assert rel_l2(scl.grad.numpy(), dp[:, 8:11]) < 2e-4
    assert rel_l2(sph.grad.numpy().reshape(-1, 48), ds) < 2e-4
