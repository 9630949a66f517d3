#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the REFERENCE's own hand-written CUDA math compiled for the host
(oracle/_ref/libgut_ref.so <- /root/reference sources, see oracle/ref_gut.cpp).  Run in the build container:

    python tests/golden/make_golden.py

The fixtures are what pins oracle/gut_oracle.c on machines where /root/reference does not exist (the GPU box)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "3dgrut_b200")):
    sys.path.insert(0, p)

import scenes  # noqa: E402
from oracle import gut_ref as gr  # noqa: E402


def projection():
    sc = scenes.scene_c1(n=300, seed=3, width=96, height=64)
    out = dict(particles=sc.particles, sph=sc.sph, width=sc.width, height=sc.height, fx=sc.fx, fy=sc.fy, cx=sc.cx, cy=sc.cy)
    for i in range(3):
        pose = scenes.pose7_from_c2w(sc.camera(i, 3))
        rf = gr.project(sc.particles, sc.sph, 3, sc.width, sc.height, [sc.fx, sc.fy], [sc.cx, sc.cy], pose, pose)
        keys, vals = gr.expand(sc.width, sc.height, rf["tiles_count"], rf["proj_pos"], rf["conic_opacity"], rf["extent"], rf["depth"])
        view, inv, pos = gr.sensor_matrices(pose, pose)
        out.update({f"pose{i}": pose, f"view{i}": view, f"inv{i}": inv, f"campos{i}": pos, f"keys{i}": keys, f"vals{i}": vals})
        out.update({f"{k}{i}": v for k, v in rf.items()})
    np.savez_compressed(os.path.join(HERE, "gut_projection_ref.npz"), **out)


FISHEYE = (0.05, -0.01, 0.002, -0.0003, 0.33)  # k1..k4, max angle


def projection_fisheye():
    """Same pins through the OpenCV fisheye model (cameraProjections.cuh:120-146).  atan2f comes from the libm of the machine that
    runs this script; the fixture therefore pins the oracle bit for bit only where libm agrees (glibc 2.3x: correctly rounded in
    practice) -- the test accepts a <= 1e-3 fraction of differing tile counts and compares float fields with a 1e-6 tolerance."""
    sc = scenes.scene_c1(n=300, seed=3, width=96, height=64)
    f = 1.2 * sc.width
    out = dict(particles=sc.particles, sph=sc.sph, width=sc.width, height=sc.height, fx=f, fy=f, cx=sc.cx, cy=sc.cy, fisheye=np.asarray(FISHEYE, np.float32))
    gr.set_camera_model(FISHEYE)
    try:
        for i in range(3):
            pose = scenes.pose7_from_c2w(sc.camera(i, 3))
            rf = gr.project(sc.particles, sc.sph, 3, sc.width, sc.height, [f, f], [sc.cx, sc.cy], pose, pose)
            keys, vals = gr.expand(sc.width, sc.height, rf["tiles_count"], rf["proj_pos"], rf["conic_opacity"], rf["extent"], rf["depth"])
            out.update({f"pose{i}": pose, f"keys{i}": keys, f"vals{i}": vals})
            out.update({f"{k}{i}": v for k, v in rf.items() if k != "visibility"})  # undefined for rejected particles in the reference
    finally:
        gr.set_camera_model(None)
    np.savez_compressed(os.path.join(HERE, "gut_projection_fisheye_ref.npz"), **out)


def projection_ftheta():
    """f-theta model (cameraProjections.cuh:148-198), backward polynomial as the reference (Newton inversion)."""
    sc = scenes.scene_c1(n=300, seed=3, width=96, height=64)
    f = 1.2 * sc.width
    a1, a3 = 1.0 / f, 0.04 / f ** 3
    ft = dict(reference_poly=0, bw=[0.0, a1, 0.0, a3, 0.0, 0.0], fw=[0.0, f, 0.0, -0.04 * f, 0.0, 0.0], cde=[1.0, 0.001, -0.002], max_angle=0.36,
              principal=(sc.width / 2.0 - 0.5, sc.height / 2.0 - 0.5))
    out = dict(particles=sc.particles, sph=sc.sph, width=sc.width, height=sc.height, bw=np.asarray(ft["bw"], np.float32),
               fw=np.asarray(ft["fw"], np.float32), cde=np.asarray(ft["cde"], np.float32), max_angle=np.float32(ft["max_angle"]),
               principal=np.asarray(ft["principal"], np.float32))
    gr.set_ftheta(ft)
    try:
        for i in range(3):
            pose = scenes.pose7_from_c2w(sc.camera(i, 3))
            rf = gr.project(sc.particles, sc.sph, 3, sc.width, sc.height, [1.0, 1.0], list(ft["principal"]), pose, pose)
            out.update({f"pose{i}": pose})
            out.update({f"{k}{i}": v for k, v in rf.items() if k != "visibility"})
    finally:
        gr.set_camera_model(None)
    np.savez_compressed(os.path.join(HERE, "gut_projection_ftheta_ref.npz"), **out)


def hits():
    rng = np.random.default_rng(2024)
    rows = []
    for degree in (2, 4):
        for _ in range(300):
            pos = rng.normal(size=3) * 0.3
            scl = np.exp(rng.normal(np.log(0.2), 0.5, 3))
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            p = np.concatenate([pos, [rng.uniform(0.02, 1.0)], q, scl, [0]]).astype(np.float32)
            ro = np.array([0, 0, -3], np.float32) + rng.normal(size=3).astype(np.float32) * 0.1
            rd = pos + rng.normal(size=3) * 0.25 - ro
            rd = (rd / np.linalg.norm(rd)).astype(np.float32)
            rgb = rng.uniform(0, 1, 3).astype(np.float32)
            T, C0, D = np.float32(rng.uniform(0.05, 1)), rng.uniform(0, 0.5, 3).astype(np.float32), np.float32(rng.uniform(0, 2))
            Tint, Cint, Dint = np.float32(T * rng.uniform(0.001, 0.9)), (C0 + rng.uniform(0.1, 1, 3)).astype(np.float32), np.float32(D + rng.uniform(0.1, 3))
            Tg, Cg, Dg = np.float32(rng.normal()), rng.normal(size=3).astype(np.float32), np.float32(rng.normal())
            acc, T1, C1, D1 = gr.hit_fwd(degree, ro, rd, p, rgb, float(T), C0, float(D))
            g, rg, Tb, Cb, Db = gr.hit_bwd(degree, ro, rd, p, rgb, 1e-4, float(Tint), float(T), float(Tg), Cint, C0, Cg, float(Dint), float(D), float(Dg))
            rows.append(np.concatenate([[degree], p, ro, rd, rgb, [T], C0, [D], [Tint], Cint, [Dint], [Tg], Cg, [Dg],
                                        [acc, T1], C1, [D1], g, rg, [Tb], Cb, [Db]]).astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "gut_hits_ref.npz"), rows=np.stack(rows))


def sph():
    rng = np.random.default_rng(9)
    c = rng.normal(size=(64, 48)).astype(np.float32)
    d = rng.normal(size=(64, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    out = np.stack([np.stack([gr.sph(deg, c[i], d[i], clamped=False) for i in range(64)]) for deg in range(4)])
    np.savez_compressed(os.path.join(HERE, "gut_sph_ref.npz"), coeffs=c, dirs=d, rgb=out)


if __name__ == "__main__":
    assert gr.available(), "oracle/_ref could not be built (needs /root/reference)"
    projection()
    projection_fisheye()
    projection_ftheta()
    hits()
    sph()
    print("golden fixtures written to", HERE)
