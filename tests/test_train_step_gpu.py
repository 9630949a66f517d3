"""System test of the path as training uses it: activations -> 3DGUT forward -> L1 gradient -> compact backward -> (single-rank) exchange
-> fused Adam.  Fits perturbed Gaussians to images rendered from the unperturbed ones; a sign or scale error anywhere in the chain
(renderer adjoint, SH adjoint, compact rebuild, activation chain rule, Adam) makes the loss go up or stall."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_short_fit_reduces_the_loss():
    import train_step
    from threedgut_tracer.tracer import ShutterType, fromOpenCVPinholeCameraModelParameters

    dev = torch.device("cuda", 0)
    sc = scenes.scene_c1(n=600, width=96, height=96)
    W, H = sc.width, sc.height
    sensor = fromOpenCVPinholeCameraModelParameters(np.array([W, H]), ShutterType.GLOBAL, np.array([sc.cx, sc.cy], np.float32),
                                                    np.array([sc.fx, sc.fy], np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32),
                                                    np.zeros(4, np.float32))
    ro, rd = sc.rays()
    rays_o, rays_d = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    P = torch.from_numpy(sc.particles).to(dev)
    S = torch.from_numpy(sc.sph).to(dev)

    def raw_from(particles, sph):
        dns = particles[:, 3:4].clamp(1e-4, 1 - 1e-4)
        return {"positions": particles[:, 0:3].clone(), "density": torch.log(dns / (1 - dns)), "rotation": particles[:, 4:8].clone(),
                "scale": torch.log(particles[:, 8:11]), "features_albedo": sph[:, 0:3].clone(), "features_specular": sph[:, 3:48].clone()}

    lrs = dict(positions=2e-3, density=0.05, rotation=1e-3, scale=5e-3, features_albedo=1e-2, features_specular=5e-4)
    truth = train_step.GaussianTrainStep(raw_from(P, S), lrs)
    views = [scenes.pose7_from_c2w(sc.camera(i, 6)) for i in range(6)]
    targets = [truth.render(rays_o, rays_d, sensor, p)[0][..., :3].clone() for p in views]
    gen = torch.Generator(device=dev).manual_seed(0)
    P2, S2 = P.clone(), S.clone()
    P2[:, 0:3] += 0.02 * torch.randn((sc.n, 3), device=dev, generator=gen)
    P2[:, 8:11] *= torch.exp(0.2 * torch.randn((sc.n, 3), device=dev, generator=gen))
    S2[:, 0:3] += 0.5 * torch.randn((sc.n, 3), device=dev, generator=gen)
    fit = train_step.GaussianTrainStep(raw_from(P2, S2), lrs)

    def mean_loss():
        return float(np.mean([float((fit.render(rays_o, rays_d, sensor, p)[0][..., :3] - t).abs().mean()) for p, t in zip(views, targets)]))

    before = mean_loss()
    for it in range(90):
        fit.step(rays_o, rays_d, sensor, views[it % 6], targets[it % 6])
    after = mean_loss()
    print(f"[train-step] mean L1 over 6 views: {before:.5f} -> {after:.5f} after 90 steps")
    assert np.isfinite(after) and after < 0.6 * before
    assert fit.optimizer.steps == 90
