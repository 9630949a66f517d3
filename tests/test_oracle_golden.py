"""Oracle vs the committed golden vectors (generated from the reference's own code by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import gut_oracle as go

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_projection_keys_and_pose_maths_match_reference_fixture():
    z = np.load(os.path.join(G, "gut_projection_ref.npz"))
    cfg = go.default_config()
    for i in range(3):
        cam = go.make_camera(int(z["width"]), int(z["height"]), float(z["fx"]), float(z["fy"]), float(z["cx"]), float(z["cy"]), z[f"pose{i}"])
        view, inv, pos = go.sensor_matrices(cam)
        assert np.array_equal(view, z[f"view{i}"]) and np.array_equal(inv, z[f"inv{i}"]) and np.array_equal(pos, z[f"campos{i}"])
        pr = go.project(cfg, cam, z["particles"], z["sph"], 3)
        assert np.array_equal(pr.tiles_count, z[f"tiles_count{i}"])
        assert np.array_equal(pr.depth.view(np.uint32), z[f"depth{i}"].view(np.uint32))
        assert np.array_equal(pr.proj_pos, z[f"proj_pos{i}"]) and np.array_equal(pr.conic_opacity, z[f"conic_opacity{i}"])
        assert np.array_equal(pr.extent, z[f"extent{i}"])
        vis = pr.tiles_count > 0
        assert vis.sum() > 50 and np.array_equal(pr.rgb[vis], z[f"rgb{i}"][vis])
        bn = go.bin_tiles(cfg, cam, pr)
        assert np.array_equal(bn.unsorted_keys, z[f"keys{i}"]) and np.array_equal(bn.unsorted_values, z[f"vals{i}"])
        order = np.argsort(z[f"keys{i}"], kind="stable")
        assert np.array_equal(bn.sorted_keys, z[f"keys{i}"][order]) and np.array_equal(bn.sorted_values, z[f"vals{i}"][order])
        # ranges partition the sorted stream by tile
        tiles = (bn.sorted_keys >> np.uint64(32)).astype(np.int64)
        for t in np.unique(tiles):
            b, e = bn.ranges[t]
            assert np.all(tiles[b:e] == t) and (e - b) == np.sum(tiles == t)


def test_fisheye_projection_matches_reference_fixture():
    """OpenCV fisheye model: the fixture was produced by the reference's own projection code on the build machine's libm (atan2f);
    on the same libm the oracle is bit-identical (tests/test_oracle_vs_ref.py), elsewhere a last-bit atan2f difference may flip a
    borderline tile count."""
    z = np.load(os.path.join(G, "gut_projection_fisheye_ref.npz"))
    cfg = go.default_config()
    for i in range(3):
        cam = go.make_camera(int(z["width"]), int(z["height"]), float(z["fx"]), float(z["fy"]), float(z["cx"]), float(z["cy"]), z[f"pose{i}"],
                             fisheye=tuple(float(v) for v in z["fisheye"]))
        pr = go.project(cfg, cam, z["particles"], z["sph"], 3)
        same = pr.tiles_count == z[f"tiles_count{i}"]
        assert same.mean() >= 0.999 and (z[f"tiles_count{i}"] == 0).sum() > 0 and z[f"tiles_count{i}"].sum() > 100
        assert np.array_equal(pr.depth.view(np.uint32), z[f"depth{i}"].view(np.uint32))
        vis = same & (pr.tiles_count > 0)
        for k in ("proj_pos", "conic_opacity", "extent"):
            assert np.allclose(getattr(pr, k)[vis], z[f"{k}{i}"][vis], rtol=1e-5, atol=1e-5), k
        if same.all():
            bn = go.bin_tiles(cfg, cam, pr)
            assert np.array_equal(bn.unsorted_keys, z[f"keys{i}"]) and np.array_equal(bn.unsorted_values, z[f"vals{i}"])


def test_ftheta_projection_matches_reference_fixture():
    """f-theta model, same libm caveat as the fisheye fixture."""
    z = np.load(os.path.join(G, "gut_projection_ftheta_ref.npz"))
    cfg = go.default_config()
    ft = dict(reference_poly=0, bw=z["bw"], fw=z["fw"], cde=z["cde"], max_angle=float(z["max_angle"]), principal=tuple(float(v) for v in z["principal"]))
    for i in range(3):
        cam = go.make_camera(int(z["width"]), int(z["height"]), 1.0, 1.0, 0.0, 0.0, z[f"pose{i}"], ftheta=ft)
        pr = go.project(cfg, cam, z["particles"], z["sph"], 3)
        same = pr.tiles_count == z[f"tiles_count{i}"]
        assert same.mean() >= 0.999 and (z[f"tiles_count{i}"] == 0).sum() > 0 and z[f"tiles_count{i}"].sum() > 100
        vis = same & (pr.tiles_count > 0)
        for k in ("proj_pos", "conic_opacity", "extent"):
            assert np.allclose(getattr(pr, k)[vis], z[f"{k}{i}"][vis], rtol=1e-5, atol=1e-5), k


def test_single_hit_forward_and_adjoint_match_reference_fixture():
    rows = np.load(os.path.join(G, "gut_hits_ref.npz"))["rows"]
    cfg = go.default_config()
    accepted = 0
    for r in rows:
        it = iter(r)
        take = lambda k: np.array([next(it) for _ in range(k)], np.float32)  # noqa: E731
        degree = int(next(it))
        p, ro, rd, rgb = take(12), take(3), take(3), take(3)
        T, C0, D, Tint, Cint, Dint, Tg, Cg, Dg = take(1)[0], take(3), take(1)[0], take(1)[0], take(3), take(1)[0], take(1)[0], take(3), take(1)[0]
        acc_ref, T1 = int(next(it)), np.float32(next(it))
        C1, D1, g_ref, rg_ref, Tb_ref = take(3), take(1)[0], take(12), take(3), take(1)[0]
        cfg.kernel_degree = degree
        acc, alpha, t = go.hit_forward(cfg, ro, rd, p)
        assert acc == acc_ref
        acc2, g, rg, Tb, _, _ = go.hit_backward(cfg, ro, rd, p, rgb, float(Tint), float(T), float(Tg), Cint, C0, Cg, float(Dint), float(D), float(Dg))
        assert acc2 == acc_ref
        if acc:
            accepted += 1
            w = np.float32(alpha) * T
            assert abs(T * (np.float32(1) - np.float32(alpha)) - T1) <= 1e-6
            assert abs(D + np.float32(t) * w - D1) <= 1e-5 * max(1.0, abs(float(D1)))
            assert np.abs(g_ref[:11] - g).max() <= 5e-5 * (np.abs(g_ref[:11]).max() + 1e-12)
            assert np.abs(rg_ref - rg).max() <= 1e-6
            assert abs(Tb - Tb_ref) <= 1e-6
    assert accepted > 300


def test_sph_matches_reference_fixture():
    z = np.load(os.path.join(G, "gut_sph_ref.npz"))
    for deg in range(4):
        for i in range(z["coeffs"].shape[0]):
            assert np.array_equal(go.sph_eval(deg, z["coeffs"][i], z["dirs"][i]), z["rgb"][deg, i])


def test_higher_msb():
    lib = go.lib()
    for n, want in [(64, 7), (2500, 12), (4056, 12), (4096, 13), (1, 1), (255, 8), (256, 9)]:
        assert lib.gut_oracle_higher_msb(n) == want, n
