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
