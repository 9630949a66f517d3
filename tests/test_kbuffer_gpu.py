"""GPU parity of the sorted (k-buffer) 3DGUT variant (3dgrut_b200/csrc/gut_render_kbuffer.cu) against the oracle's k-buffer forward and
backward (oracle/gut_oracle.c: gut_oracle_render_forward_kbuffer / _backward_kbuffer).  Same tolerances as the unsorted path."""
import numpy as np
import pytest

import scenes
from helpers import image_error_report, oracle_camera, rel_l2, tracer_pose
from oracle import gut_oracle as go

pytestmark = pytest.mark.gpu  # first hardware run: round 2 (profiles/r02_a_gate_pytest.log), K = 4 / 16, kernel degrees 2 / 4
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("k,degree", [(4, 2), (16, 2), (16, 4)])
def test_kbuffer_forward_and_gradients(k, degree):
    import b200_native as nat

    sc = scenes.scene_c1()
    sc.particles[:, 8:11] *= 2.0  # overlapping Gaussians: the per-ray hit order differs from the depth order of the lists
    c2w = sc.camera(3, 10)
    pose = tracer_pose(c2w)
    cfg = go.default_config()
    cfg.kernel_degree = degree
    cam, _ = oracle_camera(sc, c2w, pose)
    ro, rd = sc.rays()
    pr = go.project(cfg, cam, sc.particles, sc.sph, 3)
    bn = go.bin_tiles(cfg, cam, pr)
    rgba_ref, dist_ref, hits_ref = go.render_forward_kbuffer(cfg, cam, k, ro, rd, sc.particles, pr, bn)
    base = go.render_forward(cfg, cam, ro, rd, sc.particles, pr, bn)[0]
    assert np.abs(base - rgba_ref).max() > 1e-3  # the sorted variant really differs on this scene
    rng = np.random.default_rng(k)
    d_rgba = rng.normal(size=rgba_ref.shape).astype(np.float32)
    d_dist = (0.1 * rng.normal(size=dist_ref.shape)).astype(np.float32)
    dp_ref, ds_ref = go.render_backward_kbuffer(cfg, cam, k, ro, rd, sc.particles, sc.sph, 3, pr, bn, rgba_ref, dist_ref, d_rgba, d_dist)

    ncfg = nat.default_config()
    ncfg.kernel_degree = degree
    ncfg.k_buffer_size = k
    ctx = nat.Context(ncfg, 0)
    ncam = nat.Camera()
    ncam.width, ncam.height = sc.width, sc.height
    ncam.principal[:] = [sc.cx, sc.cy]
    ncam.focal[:] = [sc.fx, sc.fy]
    ncam.pose_start[:] = [float(v) for v in pose]
    ncam.pose_end[:] = [float(v) for v in pose]
    n, hw = sc.n, sc.width * sc.height
    rgba, dist, hits, vis = (np.zeros((hw, 4), np.float32), np.zeros(hw, np.float32), np.zeros(hw, np.float32), np.zeros(n, np.float32))
    p = lambda a: a.ctypes.data  # noqa: E731
    ro_c, rd_c = np.ascontiguousarray(ro), np.ascontiguousarray(rd)
    ctx.forward_host(ncam, n, p(sc.particles), p(sc.sph), 3, p(ro_c), p(rd_c), p(rgba), p(dist), p(hits), p(vis))
    mean_e, max_e, bad = image_error_report(f"kbuffer K={k} deg={degree} rgba", rgba.reshape(rgba_ref.shape), rgba_ref)
    assert mean_e <= 1e-5 and max_e <= 2e-2 and bad <= max(3, int(2e-4 * hw))
    assert float(np.mean(hits.reshape(hits_ref.shape) == hits_ref)) >= 0.999
    dp, ds = np.zeros((n, 12), np.float32), np.zeros((n, 48), np.float32)
    ctx.backward_host(ncam, n, p(sc.particles), p(sc.sph), 3, p(ro_c), p(rd_c), p(rgba), p(d_rgba), p(dist), p(d_dist), p(dp), p(ds))
    errs = dict(dp=rel_l2(dp, dp_ref), ds=rel_l2(ds, ds_ref))
    print(f"[parity] kbuffer K={k} deg={degree} gradient rel-L2:", {a: f"{b:.2e}" for a, b in errs.items()})
    assert max(errs.values()) <= 1e-3
    ctx.close()
