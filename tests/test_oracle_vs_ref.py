"""Pins the C restatement (oracle/gut_oracle.c) against the REFERENCE's own hand-written CUDA math compiled for
the host from /root/reference (oracle/_ref/libgut_ref.so, oracle/ref_gut.cpp).  Only runs where the reference is
mounted (the build container); the committed golden vectors (tests/golden) carry the same pins elsewhere."""
import numpy as np
import pytest

import scenes
from oracle import gut_oracle as go
from oracle import gut_ref as gr

pytestmark = pytest.mark.skipif(not gr.available(), reason="needs /root/reference to build oracle/_ref")


def _cam(sc, i, n=6):
    pose = scenes.pose7_from_c2w(sc.camera(i, n))
    return go.make_camera(sc.width, sc.height, sc.fx, sc.fy, sc.cx, sc.cy, pose), pose


@pytest.mark.parametrize("cam_index", range(6))
def test_projection_and_keys_bit_identical(cam_index):
    sc = scenes.scene_c1(bands=True)
    cfg = go.default_config()
    cam, pose = _cam(sc, cam_index)
    pr = go.project(cfg, cam, sc.particles, sc.sph, 3)
    rf = gr.project(sc.particles, sc.sph, 3, sc.width, sc.height, [sc.fx, sc.fy], [sc.cx, sc.cy], pose, pose)
    assert pr.tiles_count.sum() > 1000
    assert np.array_equal(pr.tiles_count, rf["tiles_count"])
    assert np.array_equal(pr.depth.view(np.uint32), rf["depth"].view(np.uint32))
    for k in ("proj_pos", "conic_opacity", "extent"):
        assert np.array_equal(getattr(pr, k), rf[k]), k
    vis = pr.tiles_count > 0
    assert np.array_equal(pr.rgb[vis], rf["rgb"][vis])
    # visibility: identical wherever the reference's value is defined (it reads an uninitialised covariance when the
    # projection was rejected, gutProjector.cuh:245-275)
    differs = pr.visibility != rf["visibility"]
    assert not np.any(differs & (pr.visibility == 1))
    bn = go.bin_tiles(cfg, cam, pr)
    keys, vals = gr.expand(sc.width, sc.height, rf["tiles_count"], rf["proj_pos"], rf["conic_opacity"], rf["extent"], rf["depth"])
    assert np.array_equal(keys, bn.unsorted_keys) and np.array_equal(vals, bn.unsorted_values)
    order = np.argsort(keys, kind="stable")  # tile bits are the high bits: a full 64-bit stable sort == the masked one
    assert np.array_equal(keys[order], bn.sorted_keys) and np.array_equal(vals[order], bn.sorted_values)


FISHEYE = (0.05, -0.01, 0.002, -0.0003, 0.33)  # k1..k4, max angle (rad): some particles fall outside the valid cone


@pytest.mark.parametrize("cam_index", range(4))
def test_fisheye_projection_bit_identical(cam_index):
    """OpenCV fisheye model (cameraProjections.cuh:120-146) through the whole projection + key expansion: bit-identical with the
    reference's own code compiled for the host (both sides call the same libm atan2f)."""
    sc = scenes.scene_c1(bands=True)
    cfg = go.default_config()
    pose = scenes.pose7_from_c2w(sc.camera(cam_index, 4))
    f = 1.2 * sc.width  # fisheye focal: pixels per radian (the image spans about +-0.4 rad, the valid cone 0.33)
    cam = go.make_camera(sc.width, sc.height, f, f, sc.cx, sc.cy, pose, fisheye=FISHEYE)
    pr = go.project(cfg, cam, sc.particles, sc.sph, 3)
    gr.set_camera_model(FISHEYE)
    try:
        rf = gr.project(sc.particles, sc.sph, 3, sc.width, sc.height, [f, f], [sc.cx, sc.cy], pose, pose)
        keys, vals = gr.expand(sc.width, sc.height, rf["tiles_count"], rf["proj_pos"], rf["conic_opacity"], rf["extent"], rf["depth"])
    finally:
        gr.set_camera_model(None)
    assert pr.tiles_count.sum() > 500
    assert (pr.tiles_count == 0).sum() > 0  # the max-angle / resolution rejections are exercised
    assert np.array_equal(pr.tiles_count, rf["tiles_count"])
    assert np.array_equal(pr.depth.view(np.uint32), rf["depth"].view(np.uint32))
    for k in ("proj_pos", "conic_opacity", "extent"):
        assert np.array_equal(getattr(pr, k), rf[k]), k
    bn = go.bin_tiles(cfg, cam, pr)
    assert np.array_equal(keys, bn.unsorted_keys) and np.array_equal(vals, bn.unsorted_values)
    # and it differs from the pinhole projection of the same scene (the model switch is really taken)
    pin = go.project(cfg, go.make_camera(sc.width, sc.height, f, f, sc.cx, sc.cy, pose), sc.particles, sc.sph, 3)
    assert not np.array_equal(pin.proj_pos, pr.proj_pos)


def _ftheta(width, height, reference_poly):
    """equidistant-like f-theta camera: backward polynomial theta = a1 r + a3 r^3, forward its low-order inverse"""
    f = 1.2 * width
    a1, a3 = 1.0 / f, 0.04 / f ** 3
    return dict(reference_poly=reference_poly, bw=[0.0, a1, 0.0, a3, 0.0, 0.0], fw=[0.0, f, 0.0, -0.04 * f, 0.0, 0.0], cde=[1.0, 0.001, -0.002],
                max_angle=0.36, principal=(width / 2.0 - 0.5, height / 2.0 - 0.5))


@pytest.mark.parametrize("reference_poly", [0, 1])
@pytest.mark.parametrize("cam_index", range(3))
def test_ftheta_projection_bit_identical(cam_index, reference_poly):
    """f-theta model (cameraProjections.cuh:148-198), both reference polynomials (Newton inversion of the backward polynomial /
    direct forward polynomial), through projection + key expansion: bit-identical with the reference's code compiled for the host."""
    sc = scenes.scene_c1(bands=True)
    cfg = go.default_config()
    pose = scenes.pose7_from_c2w(sc.camera(cam_index, 3))
    ft = _ftheta(sc.width, sc.height, reference_poly)
    cam = go.make_camera(sc.width, sc.height, 1.0, 1.0, 0.0, 0.0, pose, ftheta=ft)
    pr = go.project(cfg, cam, sc.particles, sc.sph, 3)
    gr.set_ftheta(ft)
    try:
        rf = gr.project(sc.particles, sc.sph, 3, sc.width, sc.height, [1.0, 1.0], list(ft["principal"]), pose, pose)
        keys, vals = gr.expand(sc.width, sc.height, rf["tiles_count"], rf["proj_pos"], rf["conic_opacity"], rf["extent"], rf["depth"])
    finally:
        gr.set_camera_model(None)
    assert pr.tiles_count.sum() > 500 and (pr.tiles_count == 0).sum() > 0
    assert np.array_equal(pr.tiles_count, rf["tiles_count"])
    for k in ("proj_pos", "conic_opacity", "extent"):
        assert np.array_equal(getattr(pr, k), rf[k]), k
    bn = go.bin_tiles(cfg, cam, pr)
    assert np.array_equal(keys, bn.unsorted_keys) and np.array_equal(vals, bn.unsorted_values)


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
@pytest.mark.parametrize("model", ["pinhole", "fisheye"])
def test_rolling_shutter_projection_bit_identical(kind, model):
    """projectPointWithShutter (cameraProjections.cuh:218-257) with a sensor that moves and rotates during the exposure: 5 iterations of
    pose(time of the projected row / column) -> projection, for the four readout directions."""
    sc = scenes.scene_c1(bands=True)
    cfg = go.default_config()
    p0 = scenes.pose7_from_c2w(sc.camera(1, 40))
    p1 = scenes.pose7_from_c2w(sc.camera(2, 40))  # 9 degrees further along the orbit
    fe = FISHEYE if model == "fisheye" else None
    f = 1.2 * sc.width if model == "fisheye" else sc.fx
    cam = go.make_camera(sc.width, sc.height, f, f, sc.cx, sc.cy, p0, p1, fisheye=fe, rolling_shutter=kind)
    pr = go.project(cfg, cam, sc.particles, sc.sph, 3)
    gr.set_camera_model(fe)
    gr.set_rolling_shutter(kind)
    try:
        rf = gr.project(sc.particles, sc.sph, 3, sc.width, sc.height, [f, f], [sc.cx, sc.cy], p0, p1)
    finally:
        gr.set_rolling_shutter(0)
        gr.set_camera_model(None)
    assert pr.tiles_count.sum() > 500
    assert np.array_equal(pr.tiles_count, rf["tiles_count"])
    assert np.array_equal(pr.depth.view(np.uint32), rf["depth"].view(np.uint32))
    for k in ("proj_pos", "conic_opacity", "extent"):
        assert np.array_equal(getattr(pr, k), rf[k]), k
    # the shutter really matters: the global-shutter projection of the same poses differs
    glob = go.project(cfg, go.make_camera(sc.width, sc.height, f, f, sc.cx, sc.cy, p0, p1, fisheye=fe), sc.particles, sc.sph, 3)
    assert not np.array_equal(glob.proj_pos, pr.proj_pos)


def test_sensor_pose_maths_identical():
    sc = scenes.scene_c1()
    for i in range(8):
        cam, pose = _cam(sc, i, 8)
        a = go.sensor_matrices(cam)
        b = gr.sensor_matrices(pose, pose)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def _random_hit_case(rng):
    pos = rng.normal(size=3) * 0.3
    scl = np.exp(rng.normal(np.log(0.2), 0.5, 3))
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    p = np.concatenate([pos, [rng.uniform(0.02, 1.0)], q, scl, [0]]).astype(np.float32)
    ro = np.array([0, 0, -3], np.float32) + rng.normal(size=3).astype(np.float32) * 0.1
    rd = pos + rng.normal(size=3) * 0.25 - ro
    rd = (rd / np.linalg.norm(rd)).astype(np.float32)
    return p, ro, rd


@pytest.mark.parametrize("degree", [2, 4])
def test_single_hit_forward_and_adjoint_match_reference(degree):
    rng = np.random.default_rng(degree)
    cfg = go.default_config()
    cfg.kernel_degree = degree
    accepted, worst = 0, 0.0
    for _ in range(1500):
        p, ro, rd = _random_hit_case(rng)
        rgb = rng.uniform(0, 1, 3).astype(np.float32)
        T, C0, D = float(rng.uniform(0.05, 1)), rng.uniform(0, 0.5, 3).astype(np.float32), float(rng.uniform(0, 2))
        acc_ref, T_ref, _, D_ref = gr.hit_fwd(degree, ro, rd, p, rgb, T, C0, D)
        acc, alpha, t = go.hit_forward(cfg, ro, rd, p)
        assert acc == acc_ref
        if acc:
            w = np.float32(alpha) * np.float32(T)
            assert abs(np.float32(T) * (np.float32(1) - np.float32(alpha)) - T_ref) <= 1e-6
            assert abs(np.float32(D) + np.float32(t) * w - D_ref) <= 1e-5 * max(1.0, abs(D_ref))
        Tint, Cint, Dint = T * rng.uniform(0.001, 0.9), C0 + rng.uniform(0.1, 1, 3).astype(np.float32), D + rng.uniform(0.1, 3)
        Tg, Cg, Dg = float(rng.normal()), rng.normal(size=3).astype(np.float32), float(rng.normal())
        g_ref, rg_ref, Tb_ref, _, _ = gr.hit_bwd(degree, ro, rd, p, rgb, 1e-4, Tint, T, Tg, Cint, C0, Cg, Dint, D, Dg)
        acc2, g, rg, Tb, _, _ = go.hit_backward(cfg, ro, rd, p, rgb, Tint, T, Tg, Cint, C0, Cg, Dint, D, Dg)
        assert acc2 == acc
        if acc:
            accepted += 1
            worst = max(worst, float(np.abs(g_ref[:11] - g).max() / (np.abs(g_ref[:11]).max() + 1e-12)), float(np.abs(rg_ref - rg).max()))
            assert abs(Tb - Tb_ref) <= 1e-6
    assert accepted > 500
    assert worst <= 5e-5


def test_sph_eval_and_coefficient_adjoint_match_reference():
    rng = np.random.default_rng(5)
    for deg in range(4):
        for _ in range(50):
            c = rng.normal(size=48).astype(np.float32)
            d = rng.normal(size=3)
            d = (d / np.linalg.norm(d)).astype(np.float32)
            assert np.array_equal(go.sph_eval(deg, c, d), gr.sph(deg, c, d, clamped=False))
