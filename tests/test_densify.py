"""CPU tests of the replica-consistent densification (3dgrut_b200/densify.py) against a literal numpy restatement of the reference's
GS strategy (threedgrut/strategy/gs.py) and, under gloo with two ranks, of the property it exists for: replicas that saw different
views take identical decisions and stay bit-identical.  The reference's own strategy code cannot run here (it allocates on "cuda"),
so this row is parity-UNPINNED: the numpy restatement below is ours too (file:line cited)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

import densify  # noqa: E402


def _params(n=400, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {
        "positions": torch.randn((n, 3), generator=g), "density": torch.randn((n, 1), generator=g) * 2,
        "rotation": torch.randn((n, 4), generator=g), "scale": torch.randn((n, 3), generator=g) * 0.7 - 3.0,
        "features_albedo": torch.randn((n, 3), generator=g), "features_specular": torch.randn((n, 45), generator=g) * 0.1,
    }


def _moments(params, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.rand(v.shape, generator=g) for k, v in params.items()}


def test_gradient_buffer_clone_split_prune_reset_follow_the_reference_rules():
    params = _params()
    m, v = _moments(params, 1), _moments(params, 2)
    conf = densify.DensifyConfig(clone_grad_threshold=0.05, split_grad_threshold=0.05, relative_size_threshold=0.05, seed=3)
    d = densify.GSDensifier(params, [m, v], conf)
    g = torch.Generator().manual_seed(9)
    grads = torch.randn((400, 3), generator=g) * 0.05
    grads[::3] = 0  # particles that were not visible
    cam = torch.tensor([0.5, -1.0, 4.0])
    pos0 = params["positions"].clone()
    d.update_gradient_buffer(grads, cam)
    # gs.py:127-137 restated
    mask = (grads != 0).any(1).numpy()
    want = np.zeros((400, 1), np.float32)
    dist_ = np.linalg.norm(pos0.numpy()[mask] - cam.numpy(), axis=1, keepdims=True)
    want[mask] = np.linalg.norm(grads.numpy()[mask] * dist_, axis=1, keepdims=True) / 2
    assert np.allclose(d.grad_norm_accum.numpy(), want, rtol=1e-6) and np.array_equal(d.grad_norm_denom.numpy()[:, 0], mask.astype(np.int32))

    scale0, scene_extent = torch.exp(params["scale"]).numpy().copy(), 1.0
    gn = np.nan_to_num(want[:, 0] / mask.astype(np.float32), nan=0.0)
    clone_mask = (gn >= 0.05) & (scale0.max(1) <= 0.05 * scene_extent)
    split_mask = (gn >= 0.05) & (scale0.max(1) > 0.05 * scene_extent)
    assert clone_mask.sum() > 5 and split_mask.sum() > 5
    old = {k: t.clone() for k, t in params.items()}
    old_m = m["positions"].clone()
    d.densify(scene_extent)
    n_new = 400 + clone_mask.sum() - split_mask.sum() + 2 * split_mask.sum()
    assert d.n == n_new and all(t.shape[0] == n_new for t in params.values()) and m["scale"].shape[0] == n_new
    keep = ~split_mask
    n_keep = keep.sum()
    # survivors keep their order, values and moments (gs.py:186,192); clones are exact copies with zero moments (gs.py:215-222)
    assert torch.equal(params["positions"][:n_keep], old["positions"][torch.from_numpy(keep)])
    assert torch.equal(m["positions"][:n_keep], old_m[torch.from_numpy(keep)])
    assert torch.equal(params["features_specular"][n_keep:n_keep + clone_mask.sum()], old["features_specular"][torch.from_numpy(clone_mask)])
    assert float(m["positions"][n_keep:].abs().max()) == 0.0
    # the two children of a split: scale / (0.8 * 2) in activated space (gs.py:180-183), positions = parent + R (sigma * eps) (gs.py:166-171)
    kids_scale = torch.exp(params["scale"][n_keep + clone_mask.sum():]).numpy()
    parents = np.tile(scale0[split_mask], (2, 1))
    assert np.allclose(kids_scale, parents / 1.6, rtol=1e-5)
    kids_pos = params["positions"][n_keep + clone_mask.sum():].numpy()
    off = kids_pos - np.tile(old["positions"].numpy()[split_mask], (2, 1))
    assert np.all(np.linalg.norm(off, axis=1) <= 6.0 * np.linalg.norm(parents, axis=1))  # within 6 sigma
    assert float(d.grad_norm_accum.abs().max()) == 0.0 and d.grad_norm_accum.shape[0] == n_new  # buffers reset (gs.py:285-297)

    dens = torch.sigmoid(params["density"]).squeeze(1)
    pruned = d.prune_opacity()  # gs.py:268-283
    assert pruned == int((dens < conf.prune_density_threshold).sum()) and d.n == n_new - pruned
    d.reset_density()  # gs.py:315-328: density <= new_max_density, its moments cleared
    assert float(torch.sigmoid(params["density"]).max()) <= conf.new_max_density + 1e-6
    assert float(m["density"].abs().max()) == 0.0 and float(v["density"].abs().max()) == 0.0 and float(m["scale"].abs().max()) > 0.0


def test_check_step_condition_matches_reference():
    assert densify.check_step_condition(600, 500, 15000, 300) and not densify.check_step_condition(500, 500, 15000, 100)
    assert not densify.check_step_condition(700, 500, 15000, 300) and not densify.check_step_condition(15000, 500, 15000, 300)
    assert densify.check_step_condition(3000, 0, -1, 3000) and not densify.check_step_condition(100, -1, -1, 50)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = _params()
    moments = [_moments(params, 1), _moments(params, 2)]
    d = densify.GSDensifier(params, moments, densify.DensifyConfig(clone_grad_threshold=0.05, split_grad_threshold=0.05,
                                                                   relative_size_threshold=0.05, seed=3))
    g = torch.Generator().manual_seed(100 + rank)  # every rank saw a different view: different gradients, different sensor
    for _ in range(3):
        grads = torch.randn((400, 3), generator=g) * 0.05
        grads[torch.rand(400, generator=g) < 0.3] = 0
        d.update_gradient_buffer(grads, torch.randn(3, generator=g) * 3)
    d.densify(1.0)
    d.prune_opacity()
    np.savez(os.path.join(out_dir, f"dens{rank}.npz"), **{k: v.numpy() for k, v in params.items()}, m=moments[0]["scale"].numpy())
    dist.destroy_process_group()


def test_replicas_stay_bit_identical_through_densification(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(tmp_path / f"dens{r}.npz") for r in range(world))
    assert a["positions"].shape[0] != 400  # something happened
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k


def test_mcmc_relocation_formula_matches_the_reference_kernel_loops():
    """compute_relocation vs gaussian_mcmc.cu:36-70 restated literally (double loops)."""
    import math

    rng = np.random.default_rng(0)
    M, n_max = 64, 51
    binoms = torch.tensor([[math.comb(n, k) if k <= n else 0 for k in range(n_max)] for n in range(n_max)], dtype=torch.float32)
    o = torch.from_numpy(rng.uniform(0.01, 0.95, M).astype(np.float32))
    s = torch.from_numpy(rng.uniform(0.01, 0.3, (M, 3)).astype(np.float32))
    ratios = torch.from_numpy(rng.integers(1, 9, M).astype(np.int32))
    new_o, new_s = densify.compute_relocation(o, s, ratios, binoms)
    for idx in range(M):
        n = int(ratios[idx])
        no = 1.0 - (1.0 - float(o[idx])) ** (1.0 / n)
        denom = 0.0
        for i in range(1, n + 1):
            for k in range(i):
                denom += float(binoms[i - 1, k]) * ((-1.0) ** k / math.sqrt(k + 1)) * no ** (k + 1)
        assert abs(float(new_o[idx]) - no) <= 1e-6
        assert np.allclose(new_s[idx].numpy(), float(o[idx]) / denom * s[idx].numpy(), rtol=2e-4)


def _mcmc_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = _params()
    params["density"][::7] = -8.0  # some dead Gaussians (opacity 3e-4)
    moments = [_moments(params, 1), _moments(params, 2)]
    d = densify.MCMCDensifier(params, moments, densify.MCMCConfig(seed=5))
    dead_before = int((torch.sigmoid(params["density"]) <= 0.005).sum())
    moved = d.relocate()
    added = d.add()
    d.perturb(1.6e-4)
    np.savez(os.path.join(out_dir, f"mcmc{rank}.npz"), moved=moved, added=added, dead_before=dead_before,
             dead_after=int((torch.sigmoid(params["density"]) <= 0.005).sum()), **{k: v.numpy() for k, v in params.items()},
             m=moments[0]["positions"].numpy())
    dist.destroy_process_group()


def test_mcmc_relocate_add_perturb_and_replica_consistency(tmp_path):
    world = 2
    mp.spawn(_mcmc_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(tmp_path / f"mcmc{r}.npz") for r in range(world))
    assert int(a["moved"]) == int(a["dead_before"]) > 0 and int(a["dead_after"]) == 0  # every dead Gaussian moved onto a live one
    assert int(a["added"]) == 20 and a["positions"].shape[0] == 420                      # +5 %
    assert np.all(a["m"][400:] == 0)                                                      # new Gaussians start with empty moments
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k                                              # replicas bit-identical
