"""The algebra G7 / G8 rely on since round 2 (3dgrut_b200/csrc/gut_render.cu, section comment "G7 backward"), checked in float64 numpy
against the chain the reference's hand adjoint writes (threedgut_tracer/include/3dgut/kernels/cuda/models/gaussianParticles.cuh:684-747):

  * grduGrd in closed form, without and with the distance gradient's extra terms,
  * d scale and d quat as contractions of W = grduGrd (x) d (+ groGrd (x) (o - o_f)) and of G = groGrd, applied once per particle.

No GPU, no library: this pins the derivation itself; the kernels are pinned against the oracle in tests/test_gut_parity_gpu.py."""
import numpy as np


def _rot(q):
    """quaternionWXYZToMatrix rows as the kernels stage them (r_i . d = (R d)_i)"""
    r, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y)],
                     [2 * (x * y - r * z), 1 - 2 * (x * x + z * z), 2 * (y * z + r * x)],
                     [2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)]])


def _quat_contraction(q, m):
    """matmul_bw_quat (gaussianParticles.cuh:719-747): gradient of the rows above w.r.t. (r, x, y, z) for dL/dR = m"""
    r, x, y, z = q
    return np.array([
        2 * (z * (m[0, 1] - m[1, 0]) + y * (m[2, 0] - m[0, 2]) + x * (m[1, 2] - m[2, 1])),
        2 * (y * (m[0, 1] + m[1, 0]) + z * (m[0, 2] + m[2, 0]) + r * (m[1, 2] - m[2, 1])) - 4 * x * (m[1, 1] + m[2, 2]),
        2 * (x * (m[0, 1] + m[1, 0]) + r * (m[2, 0] - m[0, 2]) + z * (m[1, 2] + m[2, 1])) - 4 * y * (m[0, 0] + m[2, 2]),
        2 * (r * (m[0, 1] - m[1, 0]) + x * (m[0, 2] + m[2, 0]) + y * (m[1, 2] + m[2, 1])) - 4 * z * (m[0, 0] + m[1, 1])])


def _pair(rng, q, s, mu, o, d, depth):
    """One (pixel, particle) pair: the reference's chain -> (groGrd, grduGrd, per-pair d scale, per-pair dL/dR), plus our closed form."""
    R = _rot(q)
    go = (R @ (o - mu)) / s
    u = (R @ d) / s
    l = u @ u
    il = 1 / np.sqrt(l)
    gd = u * il
    cc = np.cross(gd, go)
    gray_g = rng.normal()                      # dL/d|grd x gro|^2, whatever the response / blending adjoint made of it
    k = 2 * cc * gray_g
    gd_g = np.cross(go, k)                     # grdGrd
    go_g = np.cross(k, gd)                     # groGrd
    pd = -(gd @ go)
    ex = np.zeros(3)
    closed_extra = np.zeros(3)
    sd = 0.0
    if depth:
        dd = gd * pd
        h = s * dd
        hs = rng.normal() / np.sqrt(h @ h)     # (weight / gdist) * Dgrad
        hg = h * hs
        sd = hg @ (s * gd)
        gd_g = gd_g + s * hg * pd - go * sd
        ex = dd * hg                           # gsclRayHitGrd
        closed_extra = s * hg
    # reference chain: normalize adjoint
    ug = il * gd_g - il ** 3 * u * (gd_g @ u)
    # ours
    P = np.cross(k, gd)
    if depth:
        ug_closed = il * (pd * (P + closed_extra - 2 * sd * gd) - sd * go)
        go_g = go_g - gd * sd
    else:
        ug_closed = (pd * il) * P
    rdg = ug / s                               # rayDirRGrd
    prg = go_g / s                             # gposcrGrd
    d_scale = ex - u * rdg - go * prg
    dR = np.outer(rdg, d) + np.outer(prg, o - mu)
    d_pos = -(R.T @ prg)
    return go_g, ug, ug_closed, d_scale, dR, d_pos, ex


def test_closed_form_normalize_adjoint_matches_the_chain():
    rng = np.random.default_rng(3)
    for depth in (False, True):
        for _ in range(200):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            s = np.exp(rng.normal(size=3))
            mu, o = rng.normal(size=3), 4 * rng.normal(size=3)
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            _, ug, ug_closed, _, _, _, _ = _pair(rng, q, s, mu, o, d, depth)
            assert np.allclose(ug, ug_closed, rtol=1e-9, atol=1e-11 * (1 + np.abs(ug).max())), (depth, ug, ug_closed)


def test_w_and_g_sums_give_the_per_pair_gradients():
    """Sum over many pixels of one particle: per-pair (d pos, d scale, d quat) == the maps G8 applies to G, W and the depth extras."""
    rng = np.random.default_rng(5)
    for depth in (False, True):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        s = np.exp(0.5 * rng.normal(size=3))
        mu = rng.normal(size=3)
        o_f = 4 * rng.normal(size=3)
        R = _rot(q)
        ref_pos, ref_scale, ref_quat = np.zeros(3), np.zeros(3), np.zeros(4)
        G, W, EX = np.zeros(3), np.zeros((3, 3)), np.zeros(3)
        for i in range(50):
            o = o_f + (0.3 * rng.normal(size=3) if i % 2 else 0.0)   # half the pixels off the frame origin (GENERAL tiles)
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            go_g, ug, _, d_scale, dR, d_pos, ex = _pair(rng, q, s, mu, o, d, depth)
            ref_pos += d_pos; ref_scale += d_scale; ref_quat += _quat_contraction(q, dR)
            # what G7 accumulates for this pair
            G += go_g
            W += np.outer(ug, d) + np.outer(go_g, o - o_f)
            EX += ex
        # what G8 applies once
        pc = o_f - mu
        go_f = (R @ pc) / s
        pr = G / s
        pos = -(R.T @ pr)
        scale = EX - np.einsum("ik,ik->i", R, W) / (s * s) - go_f * pr
        quat = _quat_contraction(q, W / s[:, None] + np.outer(pr, pc))
        assert np.allclose(pos, ref_pos, rtol=1e-9, atol=1e-10)
        assert np.allclose(scale, ref_scale, rtol=1e-9, atol=1e-9), (depth, scale, ref_scale)
        assert np.allclose(quat, ref_quat, rtol=1e-9, atol=1e-9), (depth, quat, ref_quat)


def test_closed_form_is_at_least_as_accurate_as_the_chain_in_float32():
    """The claim of DESIGN.md section 4 (item 4): in float32 the closed form grduGrd = (pd / |grdu|) groGrd does not lose accuracy against
    the reference's three-step chain -- which forms grd |gro|^2-sized terms and cancels them -- measured against float64 on geometries with
    |gro| in the hundreds (small Gaussians seen from a few units away), the regime of the C2 / C3 scenes."""
    rng = np.random.default_rng(11)
    f32 = np.float32
    err_chain, err_closed = [], []
    for _ in range(400):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        s = np.exp(rng.normal(size=3) * 0.5) * 0.01                      # scales around 0.01
        mu = rng.normal(size=3) * 0.5
        o = mu + 4.0 * (lambda v: v / np.linalg.norm(v))(rng.normal(size=3))   # 4 units away: |gro| ~ 400
        aim = mu + s.max() * rng.normal(size=3)                          # the ray passes within ~1 sigma of the centre
        d = (aim - o) / np.linalg.norm(aim - o)
        gray_g = rng.normal()

        def run(dt):
            R = _rot(q).astype(dt)
            go = (R @ (o - mu).astype(dt)) / s.astype(dt)
            u = (R @ d.astype(dt)) / s.astype(dt)
            il = dt(1) / np.sqrt(u @ u)
            gd = u * il
            cc = np.cross(gd, go)
            k = dt(2) * cc * dt(gray_g)
            gd_g = np.cross(go, k)
            P = np.cross(k, gd)
            chain = il * gd_g - il * il * il * u * (gd_g @ u)
            closed = (-(gd @ go) * il) * P
            return chain, closed

        truth, truth2 = run(np.float64)
        assert np.allclose(truth, truth2, rtol=1e-7, atol=1e-9 * np.abs(truth).max())
        chain, closed = run(f32)
        scale = np.abs(truth).max()
        err_chain.append(np.abs(chain - truth).max() / scale)
        err_closed.append(np.abs(closed - truth).max() / scale)
    err_chain, err_closed = np.array(err_chain), np.array(err_closed)
    print(f"[algebra] float32 relative error vs float64: chain median {np.median(err_chain):.2e} p99 {np.quantile(err_chain, 0.99):.2e}; "
          f"closed form median {np.median(err_closed):.2e} p99 {np.quantile(err_closed, 0.99):.2e}")
    assert np.median(err_closed) <= np.median(err_chain) * 1.05
    assert np.quantile(err_closed, 0.99) <= np.quantile(err_chain, 0.99) * 1.05
