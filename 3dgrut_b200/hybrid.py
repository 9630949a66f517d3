"""3DGRUT hybrid rendering (BASELINE.json config 5): primary camera rays through the 3DGUT rasteriser, arbitrary secondary
rays (here: mirror reflections off a plane) through the 3DGRT tracer, both on the same Gaussians.

The reference ships no training-path implementation of the hybrid (SURVEY.md note N1): `conf.render.method` selects one
tracer and `MixtureOfGaussians.trace` exposes the other for arbitrary rays (threedgrut/model/model.py:918-930).  This
module is that composition: two calls through the two drop-in tracers, gradients of both flow into the same parameters."""
from __future__ import annotations

import torch


def mirror_rays(rays_ori: torch.Tensor, rays_dir: torch.Tensor, T_to_world: torch.Tensor, plane_point, plane_normal):
    """World-space reflection of camera rays off a plane; returns (origins, directions, hit mask), shapes [1,H,W,3] / [1,H,W,1]."""
    R, t = T_to_world[0, :3, :3], T_to_world[0, :3, 3]
    o = rays_ori @ R.T + t
    d = rays_dir @ R.T
    n = torch.as_tensor(plane_normal, dtype=o.dtype, device=o.device)
    n = n / n.norm()
    p0 = torch.as_tensor(plane_point, dtype=o.dtype, device=o.device)
    denom = (d * n).sum(-1, keepdim=True)
    s = ((p0 - o) * n).sum(-1, keepdim=True) / torch.where(denom.abs() > 1e-8, denom, torch.full_like(denom, 1e-8))
    hit = (s > 0) & (denom < 0)
    origin = o + s * d
    refl = d - 2.0 * denom * n
    return torch.where(hit, origin, o), torch.where(hit, refl, d), hit


def render_hybrid(gut_tracer, grt_tracer, gaussians, gpu_batch, plane_point=(0.0, 0.0, -1.2), plane_normal=(0.0, 0.0, 1.0),
                  reflectivity: float = 0.3, train: bool = False, frame_id: int = 0):
    """Primary pass (3DGUT) + one secondary bounce (3DGRT); returns the primary dict extended with `pred_secondary` and
    `pred_features_hybrid` = primary + reflectivity * (1 - primary opacity) * secondary radiance on reflecting pixels."""
    primary = gut_tracer.render(gaussians, gpu_batch, train=train, frame_id=frame_id)
    sec_o, sec_d, hit = mirror_rays(gpu_batch.rays_ori, gpu_batch.rays_dir, gpu_batch.T_to_world.to(gpu_batch.rays_ori.device),
                                    plane_point, plane_normal)

    class _SecondaryBatch:
        rays_ori = sec_o.contiguous()
        rays_dir = sec_d.contiguous()
        T_to_world = torch.eye(4, device=sec_o.device, dtype=sec_o.dtype)[None]  # rays are already in world space

    grt_tracer.build_acc(gaussians, rebuild=True)
    secondary = grt_tracer.render(gaussians, _SecondaryBatch, train=train, frame_id=frame_id)
    weight = reflectivity * (1.0 - primary["pred_opacity"]) * hit.to(primary["pred_opacity"].dtype)
    out = dict(primary)
    out["pred_secondary"] = secondary["pred_features"]
    out["pred_features_hybrid"] = primary["pred_features"] + weight * secondary["pred_features"]
    return out
