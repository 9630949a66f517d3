"""World-size-2 gloo test (CPU) of the view-parallel plumbing: disjoint cameras per rank, one bucketed all-reduce,
result == the serial sum of the per-view gradients (computed with the CPU oracle)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

import scenes  # noqa: E402
from helpers import oracle_frame  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import view_parallel as vp

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.scene_c1(n=120, width=48, height=48)
    (view,) = vp.views_for_rank(step=3, rank=rank, world=world, num_views=10)
    ref = oracle_frame(sc, sc.camera(view, 10), seed=view)
    dp, ds = torch.from_numpy(ref["dp"]), torch.from_numpy(ref["ds"])
    grads = [dp[:, 0:3].contiguous(), dp[:, 4:8].contiguous(), dp[:, 8:11].contiguous(), dp[:, 3:4].contiguous(), ds]
    params = [torch.full((4,), float(rank))]
    vp.broadcast_parameters(params)
    assert float(params[0][0]) == 0.0
    red = vp.allreduce_gradients(grads)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), view=view, **{f"g{i}": r.numpy() for i, r in enumerate(red)})
    dist.destroy_process_group()


def test_view_parallel_allreduce_matches_serial_sum(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    views = [int(o["view"]) for o in outs]
    assert len(set(views)) == world  # disjoint cameras
    sc = scenes.scene_c1(n=120, width=48, height=48)
    tot = None
    for v in views:
        ref = oracle_frame(sc, sc.camera(v, 10), seed=v)
        dp, ds = ref["dp"], ref["ds"]
        parts = [dp[:, 0:3], dp[:, 4:8], dp[:, 8:11], dp[:, 3:4], ds]
        tot = parts if tot is None else [a + b for a, b in zip(tot, parts)]
    for r in range(world):
        for i, t in enumerate(tot):
            assert np.allclose(outs[r][f"g{i}"], t, rtol=1e-6, atol=1e-7)


def test_views_for_rank_partition():
    import view_parallel as vp

    for world in (1, 2, 4, 8):
        for step in range(5):
            seen = [v for r in range(world) for v in vp.views_for_rank(step, r, world, 1000, views_per_rank=3)]
            assert len(seen) == len(set(seen)) == world * 3


class _OracleRaster:
    """CPU stand-in for SplatRaster.sph_grad_from_views built on the oracle's SH evaluation (the product's version is a CUDA kernel;
    this one only lets the gloo test run the exchange's collectives and check the identity d_sph = sum_v basis(dir_v) x g_v)."""

    def sph_grad_from_views(self, deg, particle_density, positions, g_all, out=None):
        from oracle import gut_oracle as go

        pos = particle_density[:, 0:3].numpy()
        n = pos.shape[0]
        res = np.zeros((n, 16, 3), np.float64)
        eye = np.eye(16, dtype=np.float32)
        for v, cam in enumerate(np.asarray(positions, np.float32)):
            d = pos - cam[None]
            d = d / np.linalg.norm(d, axis=1, keepdims=True)
            for i in range(n):
                g = g_all[v, i, 0:3].numpy().astype(np.float64)
                if not g.any():
                    continue
                # basis_j = SH evaluation of the unit coefficient e_j (the oracle adds 0.5 to the radiance: subtract it)
                basis = np.array([go.sph_eval(deg, np.repeat(eye[j][:, None], 3, 1), d[i])[0] - 0.5 for j in range(16)])
                res[i] += basis[:, None] * g[None, :]
        out.copy_(torch.from_numpy(res.reshape(n, 48).astype(np.float32)))
        return out


def _compact_worker(rank, world, port, out_dir):
    import view_parallel as vp

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.scene_c1(n=60, width=32, height=32)
    views = [vp.views_for_rank(step=1, rank=r, world=world, num_views=10)[0] for r in range(world)]
    ref = oracle_frame(sc, sc.camera(views[rank], 10), seed=views[rank])
    ex = vp.CompactGradientExchange(_OracleRaster(), sc.n, torch.device("cpu"))
    d_particles, g = ex.out()
    d_particles.copy_(torch.from_numpy(ref["dp"]))
    # a view's SH gradient row is basis x g and basis_0 is the constant 0.2820948: g = row[0] / basis_0
    g[:, 0:3] = torch.from_numpy(ref["ds"][:, 0:3] / 0.28209479177387814)
    g[:, 3] = 0
    positions = np.stack([np.asarray(sc.camera(v, 10), np.float32)[:3, 3] for v in views])
    dp, ds = ex.exchange(sc.sph_degree, torch.from_numpy(sc.particles), positions)
    np.savez(os.path.join(out_dir, f"compact{rank}.npz"), views=np.array(views), dp=dp.numpy(), ds=ds.numpy())
    dist.destroy_process_group()


def test_compact_exchange_matches_serial_sum(tmp_path):
    """all-reduce [N,12] + all-gather [N,4] + rebuild == the serial sum of the per-view [N,12] / [N,48] oracle gradients."""
    world = 2
    mp.spawn(_compact_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"compact{r}.npz") for r in range(world)]
    views = [int(v) for v in outs[0]["views"]]
    sc = scenes.scene_c1(n=60, width=32, height=32)
    refs = [oracle_frame(sc, sc.camera(v, 10), seed=v) for v in views]
    dp_sum, ds_sum = sum(r["dp"] for r in refs), sum(r["ds"] for r in refs)
    assert np.abs(ds_sum).max() > 0
    for o in outs:
        assert np.allclose(o["dp"], dp_sum, rtol=1e-6, atol=1e-7)
        assert np.allclose(o["ds"], ds_sum, rtol=2e-5, atol=1e-6 * float(np.abs(ds_sum).max()))
    assert np.array_equal(outs[0]["ds"], outs[1]["ds"])  # replicas stay bit-identical


def _accumulate_worker(rank, world, port, out_dir):
    import view_parallel as vp

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.scene_c1(n=60, width=32, height=32)
    V = 2
    mine = vp.views_for_rank(step=1, rank=rank, world=world, num_views=10, views_per_rank=V)
    ex = vp.CompactGradientExchange(_OracleRaster(), sc.n, torch.device("cpu"), views_per_rank=V)
    for slot, view in enumerate(mine):
        ref = oracle_frame(sc, sc.camera(view, 10), seed=view)
        d_particles, g = ex.out(slot)
        d_particles.copy_(torch.from_numpy(ref["dp"]))
        g[:, 0:3] = torch.from_numpy(ref["ds"][:, 0:3] / 0.28209479177387814)
        g[:, 3] = 0
        ex.submit(slot)  # accumulates, starts this view's all-gather without waiting
    # slot-major, rank-minor: the order the gathered [V, world, N, 4] buffer is laid out in
    order = [vp.views_for_rank(1, r, world, 10, V)[j] for j in range(V) for r in range(world)]
    positions = np.stack([np.asarray(sc.camera(v, 10), np.float32)[:3, 3] for v in order])
    dp, ds = ex.finish(sc.sph_degree, torch.from_numpy(sc.particles), positions)
    np.savez(os.path.join(out_dir, f"acc{rank}.npz"), views=np.array(order), dp=dp.numpy(), ds=ds.numpy())
    dist.destroy_process_group()


def test_accumulated_exchange_over_two_views_per_rank(tmp_path):
    """views_per_rank = 2: per-view asynchronous all-gathers + ONE all-reduce of the accumulated [N,12] + rebuild over all 4 views
    == the serial sum of the four per-view oracle gradients, bit-identical on both ranks."""
    world = 2
    mp.spawn(_accumulate_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"acc{r}.npz") for r in range(world)]
    views = [int(v) for v in outs[0]["views"]]
    assert len(set(views)) == 4
    sc = scenes.scene_c1(n=60, width=32, height=32)
    refs = [oracle_frame(sc, sc.camera(v, 10), seed=v) for v in views]
    dp_sum, ds_sum = sum(r["dp"] for r in refs), sum(r["ds"] for r in refs)
    for o in outs:
        assert np.allclose(o["dp"], dp_sum, rtol=1e-5, atol=1e-6)
        assert np.allclose(o["ds"], ds_sum, rtol=2e-5, atol=1e-6 * float(np.abs(ds_sum).max()))
    assert np.array_equal(outs[0]["ds"], outs[1]["ds"]) and np.array_equal(outs[0]["dp"], outs[1]["dp"])
