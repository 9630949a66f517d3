"""Parity against the REFERENCE'S OWN 3DGUT kernels running on the GPU: oracle/_ref/libgut_ref_cuda.so is threedgut_tracer/src/gutRenderer.cu
(projectOnTiles, CUB scan, expandTileProjections, 44-bit CUB radix sort, tile ranges, render, renderBackward with the hand-written adjoint,
projectBackward, and the host orchestration around them) compiled UNMODIFIED for sm_100a in the build container, with only the slangc output
replaced by a hand translation (oracle/ref_cuda/threedgutSlang.cuh).  This is the pin the round-1 oracle lacked: the CPU oracle and the
product are both compared with what the reference's kernels compute on identical tensors.

The reference binary is built the way its setup script builds it (-use_fast_math -O3: FMA contraction, approximate div / sqrt / exp), ours
keeps the projection stage IEEE, so integers are compared as "equal except for a counted borderline set":
  tile counts equal on >= 99.9 % of the particles, and wherever they are equal for ALL particles of a tile list the sorted (key, value)
  stream is bit-identical; RGBA / dist mean |diff| <= 1e-5, outliers as in the other parity tests; gradients rel-L2 <= 3e-3 against the
  reference (two fast-math evaluations of a discontinuous accept test; the oracle-vs-reference figure is printed beside ours).
First run on a B200 (profiles/r02_f_reference_kernels_on_gpu.md): C1 tile counts equal on 100 %, all 64 tile lists identical in order; C2
tile counts equal on 99.9983-99.9993 % of 300k particles, 2478-2482 of 2500 tile lists identical in order (the reference's -use_fast_math
build contracts the depth FMA: 3.4 % of the depth keys differ in the last bit, which reorders neighbours with near-equal depth), images
35-56 of 640 000 pixels off by more than 1e-4, gradients 2e-5 .. 6e-4 -- except d_quat of C2 camera 41 at 2.22e-3, the frame and tensor
whose fp32-vs-fp64 yardstick is 2.24e-3 (one borderline anisotropic particle, tests/test_gut_headline_parity_gpu.py): hence 3e-3 here.
Skipped when the library was not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

import scenes
from helpers import image_error_report, oracle_frame, rel_l2, tracer_pose

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _have():
    from oracle import gut_ref_cuda

    return gut_ref_cuda.available()


def _run_reference(sc, pose, d_rgba, d_dist):
    from oracle import gut_ref_cuda as grc

    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    ro, rd = sc.rays()
    particles, sph, tro, trd = t(sc.particles), t(sc.sph), t(ro), t(rd)
    rr = grc.ReferenceRaster()
    s = torch.cuda.current_stream(dev).cuda_stream
    rgba, dist, hits, vis = rr.trace(torch, s, 0, sc.sph_degree, particles, sph, sc.width, sc.height, sc.fx, sc.fy, sc.cx, sc.cy, pose, tro, trd)
    torch.cuda.synchronize()
    tiles = ((sc.width + 15) // 16) * ((sc.height + 15) // 16)
    dbg = {k: rr.debug(k, sc.n, tiles) for k in ("tiles_count", "sorted_keys", "sorted_values", "ranges", "depth", "rgb")}
    dp, ds = rr.trace_bwd(torch, s, 0, sc.sph_degree, particles, sph, sc.width, sc.height, sc.fx, sc.fy, sc.cx, sc.cy, pose, tro, trd, rgba,
                          t(d_rgba), dist, t(d_dist))
    torch.cuda.synchronize()
    out = dict(rgba=rgba.cpu().numpy(), dist=dist.cpu().numpy(), hits=hits.cpu().numpy(), vis=vis.cpu().numpy().view(np.int32), dp=dp.cpu().numpy(),
               ds=ds.cpu().numpy(), **dbg)
    rr.close()
    return out


def _run_ours(sc, pose, d_rgba, d_dist):
    import b200_native as nat

    ctx = nat.Context(nat.default_config(), 0)
    cam = nat.Camera()
    cam.width, cam.height = sc.width, sc.height
    cam.principal[:] = [sc.cx, sc.cy]
    cam.focal[:] = [sc.fx, sc.fy]
    cam.pose_start[:] = [float(v) for v in pose]
    cam.pose_end[:] = [float(v) for v in pose]
    n, hw = sc.n, sc.width * sc.height
    rgba, dist, hits, vis = (np.zeros((hw, 4), np.float32), np.zeros(hw, np.float32), np.zeros(hw, np.float32), np.zeros(n, np.float32))
    p = lambda a: a.ctypes.data  # noqa: E731
    ro, rd = sc.rays()
    ro, rd = np.ascontiguousarray(ro), np.ascontiguousarray(rd)
    ctx.forward_host(cam, n, p(sc.particles), p(sc.sph), sc.sph_degree, p(ro), p(rd), p(rgba), p(dist), p(hits), p(vis))
    dbg = dict(tiles_count=ctx.debug_copy(nat.DBG_TILES_COUNT), sorted_keys=ctx.debug_copy(nat.DBG_SORTED_KEYS),
               sorted_values=ctx.debug_copy(nat.DBG_SORTED_VALUES), ranges=ctx.debug_copy(nat.DBG_TILE_RANGES), depth=ctx.debug_copy(nat.DBG_DEPTH))
    dp, ds = np.zeros((n, 12), np.float32), np.zeros((n, 48), np.float32)
    d_rgba, d_dist = np.ascontiguousarray(d_rgba), np.ascontiguousarray(d_dist)
    ctx.backward_host(cam, n, p(sc.particles), p(sc.sph), sc.sph_degree, p(ro), p(rd), p(rgba), p(d_rgba), p(dist), p(d_dist), p(dp), p(ds))
    ctx.close()
    return dict(rgba=rgba.reshape(sc.height, sc.width, 4), dist=dist.reshape(sc.height, sc.width, 1), hits=hits.reshape(sc.height, sc.width, 1),
                vis=vis.view(np.int32), dp=dp, ds=ds, **dbg)


def _compare(label, sc, cam_index, n_cams, with_oracle):
    c2w = sc.camera(cam_index, n_cams)
    pose = tracer_pose(c2w)
    rng = np.random.default_rng(cam_index)
    d_rgba = rng.normal(size=(sc.height, sc.width, 4)).astype(np.float32)
    d_dist = (0.1 * rng.normal(size=(sc.height, sc.width, 1))).astype(np.float32)
    ref = _run_reference(sc, pose, d_rgba, d_dist)
    ours = _run_ours(sc, pose, d_rgba, d_dist)
    arms = {"ours": ours}
    if with_oracle:
        o = oracle_frame(sc, c2w, seed=cam_index, pose=pose)
        arms["oracle"] = dict(rgba=o["rgba"], dist=o["dist"], hits=o["hits"], dp=o["dp"], ds=o["ds"], tiles_count=o["pr"].tiles_count,
                              sorted_keys=o["bn"].sorted_keys, sorted_values=o["bn"].sorted_values, ranges=o["bn"].ranges, depth=o["pr"].depth)
        assert np.array_equal(o["d_rgba"], d_rgba) and np.array_equal(o["d_dist"], d_dist)
    P = sc.width * sc.height
    cols = dict(pos=slice(0, 3), dns=slice(3, 4), quat=slice(4, 8), scl=slice(8, 11))
    for name, a in arms.items():
        tc_same = float(np.mean(a["tiles_count"] == ref["tiles_count"]))
        depth_same = float(np.mean(np.asarray(a["depth"]).view(np.uint32) == ref["depth"].view(np.uint32)))
        print(f"[ref-gpu] {label} cam{cam_index} {name}: tile counts equal on {tc_same * 100:.4f} % of {sc.n} particles "
              f"(I {int(np.asarray(a['tiles_count'], np.int64).sum())} vs reference {int(ref['tiles_count'].astype(np.int64).sum())}), depth bits equal on {depth_same * 100:.4f} %")
        assert tc_same >= 0.999
        if tc_same == 1.0 and depth_same == 1.0:
            assert np.array_equal(a["sorted_keys"], ref["sorted_keys"]) and np.array_equal(a["sorted_values"], ref["sorted_values"])
            assert np.array_equal(a["ranges"], ref["ranges"])
            print(f"[ref-gpu] {label} cam{cam_index} {name}: sorted (key, value) stream and tile ranges BIT-IDENTICAL to the reference's CUB 44-bit sort")
        else:
            # per tile: the particle SETS must agree except where a tile count differed
            same_tiles = 0
            T = ref["ranges"].shape[0]
            for tix in range(T):
                ra, rb = ref["ranges"][tix]
                oa, ob = a["ranges"][tix]
                if (rb - ra) == (ob - oa) and np.array_equal(a["sorted_values"][oa:ob], ref["sorted_values"][ra:rb]):
                    same_tiles += 1
            print(f"[ref-gpu] {label} cam{cam_index} {name}: {same_tiles}/{T} tile lists identical (order included)")
            assert same_tiles >= 0.97 * T
        mean_e, max_e, bad = image_error_report(f"{label} cam{cam_index} {name} vs reference-gpu rgba", a["rgba"].reshape(ref["rgba"].shape), ref["rgba"])
        assert mean_e <= 1e-5 and max_e <= 5e-2 and bad <= max(3, int(4e-4 * P))
        dscale = max(1.0, float(np.abs(ref["dist"][ref["dist"] < 1e5]).max()))
        mean_e, _, bad = image_error_report(f"{label} cam{cam_index} {name} vs reference-gpu dist", a["dist"].reshape(ref["dist"].shape), ref["dist"],
                                            atol=1e-4 * dscale)
        assert mean_e <= 1e-5 * dscale and bad <= max(3, int(4e-4 * P))
        same_hits = float(np.mean(a["hits"].reshape(ref["hits"].shape) == ref["hits"]))
        errs = {k: rel_l2(a["dp"][:, v], ref["dp"][:, v]) for k, v in cols.items()}
        errs["sph"] = rel_l2(a["ds"], ref["ds"])
        print(f"[ref-gpu] {label} cam{cam_index} {name}: hit counts equal on {same_hits * 100:.4f} % of pixels; gradient rel-L2 vs reference-gpu:",
              {k: f"{v:.2e}" for k, v in errs.items()})
        assert same_hits >= 0.999
        for k, v in errs.items():
            assert v <= 3e-3, (name, k, v)


@pytest.mark.skipif(not _have(), reason="oracle/_ref/libgut_ref_cuda.so not built (needs /root/reference at build time)")
def test_c1_reference_kernels_vs_oracle_and_ours():
    """C1 (1k Gaussians, 128x128): CPU oracle AND product against the reference's kernels."""
    _compare("c1", scenes.scene_c1(), 1, 8, with_oracle=True)


@pytest.mark.skipif(not _have(), reason="oracle/_ref/libgut_ref_cuda.so not built (needs /root/reference at build time)")
def test_c1_dc_only_reference_kernels():
    _compare("c1-dc", scenes.scene_c1(bands=False), 5, 8, with_oracle=True)


@pytest.mark.skipif(not _have(), reason="oracle/_ref/libgut_ref_cuda.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("cam_index", [3, 41])
def test_c2_reference_kernels_vs_oracle_and_ours(cam_index):
    """BASELINE configs[1] at full scale (300k Gaussians, 800x800): product and oracle against the reference's kernels."""
    _compare("c2", scenes.scene_c2(), cam_index, 100, with_oracle=True)


@pytest.mark.skipif(not _have(), reason="oracle/_ref/libgut_ref_cuda.so not built (needs /root/reference at build time)")
def test_c3_like_reference_kernels_vs_ours():
    sc = scenes.scene_c3(n=400_000)
    _compare("c3-400k", sc, 2, 16, with_oracle=False)
