"""Host-side mirror of the reference's 3DGRT tracer surface, backed by our LBVH ray tracer (csrc/grt.cu).

Same names, argument meaning and tensor contracts as the reference:
  Tracer / Tracer._Autograd / build_acc / render        threedgrt_tracer/tracer.py:50-255
  OptixTracer{trace, trace_bwd, build_bvh}              threedgrt_tracer/bindings.cpp:32-38, include/3dgrt/optixTracer.h:128-177
The class keeps the name OptixTracer for drop-in compatibility; there is no OptiX here (B200 has no RT cores).
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np
import torch

import b200_native as native


def _cfg_get(conf, path, default):
    cur = conf
    for key in path.split("."):
        if cur is None:
            return default
        cur = cur.get(key, None) if isinstance(cur, dict) else getattr(cur, key, None)
    return default if cur is None else cur


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


class OptixTracer:
    """Python twin of lib3dgrt_cc.OptixTracer (bindings.cpp:32-38); constructor arguments as optixTracer.h:128-141."""

    def __init__(self, path=None, cuda_path=None, pipeline="reference", backward_pipeline="referenceBwd", primitive="instances",
                 particle_kernel_degree=4, particle_kernel_min_response=0.0113, particle_kernel_max_alpha=0.99,
                 particle_kernel_density_clamping=True, particle_radiance_sph_degree=3, enable_normals=False, enable_hitcounts=True):
        if not torch.cuda.is_available():
            raise RuntimeError("threedgrt_tracer (B200): CUDA device required; there is no CPU path")
        if pipeline not in ("reference",) or backward_pipeline not in ("referenceBwd",):
            raise NotImplementedError("only the default reference / referenceBwd pipelines are built")
        if primitive != "instances":
            raise NotImplementedError("only the default `instances` proxy is built (configs/render/3dgrt.yaml:11)")
        if enable_normals:
            raise NotImplementedError("normals output is not built")
        if int(particle_radiance_sph_degree) != 3:
            raise NotImplementedError("this build stores 16 SH coefficients per particle")
        cfg = native.grt_default_config()
        cfg.kernel_degree = int(particle_kernel_degree)
        cfg.min_response = float(particle_kernel_min_response)
        cfg.max_alpha = float(particle_kernel_max_alpha)
        cfg.density_clamping = int(bool(particle_kernel_density_clamping))
        self._cfg = cfg
        self._ctx = {}

    def _context(self, device) -> native.GrtContext:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._ctx:
            self._ctx[idx] = native.GrtContext(self._cfg, idx)
        return self._ctx[idx]

    def build_bvh(self, mog_pos, mog_rot, mog_scl, mog_dns, rebuild=True, allow_update=False):
        """optixTracer.cpp:616-890"""
        dev = mog_pos.device
        pos, rot, scl, dns = (t.detach().contiguous().float() for t in (mog_pos, mog_rot, mog_scl, mog_dns))
        self._keep = (pos, rot, scl, dns)  # keep alive until the stream has consumed them
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._context(dev).build_bvh(stream, int(pos.shape[0]), _ptr(pos), _ptr(rot), _ptr(scl), _ptr(dns), rebuild, allow_update)

    @staticmethod
    def _r2w(ray_to_world) -> np.ndarray:
        m = ray_to_world.detach().cpu().numpy().reshape(-1, 4, 4)[0][:3, :4]  # first pose, 3 rows (optixTracer.cpp:931)
        return np.ascontiguousarray(m, dtype=np.float32)

    def trace(self, frame_id, ray_to_world, ray_ori, ray_dir, particle_density, particle_features, render_opts, sph_degree, min_transmittance):
        """optixTracer.cpp:893-960 -> (feat [B,H,W,3], alpha [B,H,W,1], hit [B,H,W,2], normals [B,H,W,3], hits [B,H,W,1], vis [N,1])"""
        dev = ray_ori.device
        b, h, w = (int(v) for v in ray_ori.shape[:3])
        n = int(particle_density.shape[0])
        particle_density, particle_features = particle_density.contiguous(), particle_features.contiguous()
        ray_ori, ray_dir = ray_ori.contiguous(), ray_dir.contiguous()
        opts = dict(dtype=torch.float32, device=dev)
        feat, alpha = torch.empty((b, h, w, 3), **opts), torch.empty((b, h, w, 1), **opts)
        hit, hits = torch.empty((b, h, w, 2), **opts), torch.empty((b, h, w, 1), **opts)
        nrm = torch.zeros((b, h, w, 3), **opts)
        vis = torch.empty((max(n, 1), 1), **opts)
        r2w = self._r2w(ray_to_world)
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._context(dev).trace(stream, n, _ptr(particle_density), _ptr(particle_features), int(sph_degree), float(min_transmittance), b, h, w,
                                 _ptr(ray_ori), _ptr(ray_dir), r2w.ctypes.data, _ptr(feat), _ptr(alpha), _ptr(hit), _ptr(hits), _ptr(vis))
        return feat, alpha, hit, nrm, hits, vis[:n]

    def set_replay(self, enable, device):
        """Measurement-free switch (no reference twin): record the forward's hit lists for the backward (grtb200_set_replay)."""
        if getattr(self, "_replay", None) != bool(enable):
            self._context(device).set_replay(bool(enable))
            self._replay = bool(enable)

    def trace_counters(self, ray_to_world, ray_ori, ray_dir, particle_density, particle_features, sph_degree, min_transmittance):
        """Measurement helper (no reference twin): work counters of one forward trace -- rays, k-nearest queries, node visits, box / proxy
        tests, candidate and accepted hits (grtb200_debug_trace_counters); writes no image."""
        dev = ray_ori.device
        b, h, w = (int(v) for v in ray_ori.shape[:3])
        n = int(particle_density.shape[0])
        particle_density, particle_features = particle_density.contiguous(), particle_features.contiguous()
        ray_ori, ray_dir = ray_ori.contiguous(), ray_dir.contiguous()
        vis = torch.empty((max(n, 1), 1), dtype=torch.float32, device=dev)
        r2w = self._r2w(ray_to_world)
        stream = torch.cuda.current_stream(dev).cuda_stream
        return self._context(dev).trace_counters(stream, n, _ptr(particle_density), _ptr(particle_features), int(sph_degree), float(min_transmittance),
                                                 b, h, w, _ptr(ray_ori), _ptr(ray_dir), r2w.ctypes.data, _ptr(vis))

    def trace_bwd(self, frame_id, ray_to_world, ray_ori, ray_dir, ray_features, ray_density, ray_hit_distance, ray_normals, particle_density,
                  particle_features, ray_features_grd, ray_density_grd, ray_hit_distance_grd, ray_normals_grd, render_opts, sph_degree,
                  min_transmittance):
        """optixTracer.cpp:962-1031 -> (dDensity [N,12], dFeatures [N,48])"""
        dev = ray_ori.device
        b, h, w = (int(v) for v in ray_ori.shape[:3])
        n = int(particle_density.shape[0])
        particle_density, particle_features = particle_density.contiguous(), particle_features.contiguous()
        ray_ori, ray_dir = ray_ori.contiguous(), ray_dir.contiguous()
        rf, rd_, rh = ray_features.contiguous(), ray_density.contiguous(), ray_hit_distance.contiguous()
        g_f, g_a = ray_features_grd.contiguous().float(), ray_density_grd.contiguous().float()
        g_d = ray_hit_distance_grd.contiguous().float()
        if g_d.shape[-1] != 1:
            g_d = g_d[..., 0:1].contiguous()
        d_density = torch.empty((max(n, 1), 12), dtype=torch.float32, device=dev)
        d_features = torch.empty((max(n, 1), 48), dtype=torch.float32, device=dev)
        r2w = self._r2w(ray_to_world)
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._context(dev).trace_bwd(stream, n, _ptr(particle_density), _ptr(particle_features), int(sph_degree), float(min_transmittance), b, h, w,
                                     _ptr(ray_ori), _ptr(ray_dir), r2w.ctypes.data, _ptr(rf), _ptr(rd_), _ptr(rh), _ptr(g_f), _ptr(g_a),
                                     _ptr(g_d), _ptr(d_density), _ptr(d_features))
        return d_density[:n], d_features[:n]

    def native_context(self, device=None) -> native.GrtContext:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        return self._context(dev)


class Tracer:
    class _Autograd(torch.autograd.Function):
        """threedgrt_tracer/tracer.py:51-164"""

        @staticmethod
        def forward(ctx, tracer_wrapper, frame_id, ray_to_world, ray_ori, ray_dir, mog_pos, mog_rot, mog_scl, mog_dns, mog_sph, render_opts,
                    sph_degree, min_transmittance):
            particle_density = torch.concat([mog_pos, mog_dns, mog_rot, mog_scl, torch.zeros_like(mog_dns)], dim=1)
            ray_features, ray_density, ray_hit_distance, ray_normals, hits_count, mog_visibility = tracer_wrapper.trace(
                frame_id, ray_to_world, ray_ori, ray_dir, particle_density, mog_sph, render_opts, sph_degree, min_transmittance)
            ctx.save_for_backward(ray_to_world, ray_ori, ray_dir, ray_features, ray_density, ray_hit_distance, ray_normals, particle_density, mog_sph)
            ctx.frame_id, ctx.render_opts, ctx.sph_degree = frame_id, render_opts, sph_degree
            ctx.min_transmittance, ctx.tracer_wrapper = min_transmittance, tracer_wrapper
            return ray_features, ray_density, ray_hit_distance[:, :, :, 0:1], ray_normals, hits_count, mog_visibility

        @staticmethod
        def backward(ctx, ray_features_grd, ray_density_grd, ray_hit_distance_grd, ray_normals_grd, ray_hits_count_grd_UNUSED,
                     mog_visibility_grd_UNUSED):
            (ray_to_world, ray_ori, ray_dir, ray_features, ray_density, ray_hit_distance, ray_normals, particle_density, mog_sph) = ctx.saved_tensors
            particle_density_grd, mog_sph_grd = ctx.tracer_wrapper.trace_bwd(
                ctx.frame_id, ray_to_world, ray_ori, ray_dir, ray_features, ray_density, ray_hit_distance, ray_normals, particle_density, mog_sph,
                ray_features_grd, ray_density_grd, ray_hit_distance_grd, ray_normals_grd, ctx.render_opts, ctx.sph_degree, ctx.min_transmittance)
            mog_pos_grd, mog_dns_grd, mog_rot_grd, mog_scl_grd, _ = torch.split(particle_density_grd, [3, 1, 4, 3, 1], dim=1)
            return (None, None, None, None, None, mog_pos_grd, mog_rot_grd, mog_scl_grd, mog_dns_grd, mog_sph_grd, None, None, None)

    class RenderOpts(IntEnum):
        NONE = 0
        DEFAULT = NONE

    def __init__(self, conf):
        self.device = "cuda"
        self.conf = conf
        self.num_update_bvh = 0
        torch.zeros(1, device=self.device)
        g = lambda k, d: _cfg_get(conf, "render." + k, d)  # noqa: E731
        self.tracer_wrapper = OptixTracer(
            None, None, g("pipeline_type", "reference"), g("backward_pipeline_type", "referenceBwd"), g("primitive_type", "instances"),
            g("particle_kernel_degree", 4), g("particle_kernel_min_response", 0.0113), g("particle_kernel_max_alpha", 0.99),
            g("particle_kernel_density_clamping", True), g("particle_radiance_sph_degree", 3), g("enable_normals", False),
            g("enable_hitcounts", True))
        self._min_transmittance = float(g("min_transmittance", 0.001))
        self._clamping = bool(g("particle_kernel_density_clamping", True))
        self._max_updates = int(g("max_consecutive_bvh_update", 1))
        self._timings_on = bool(g("enable_kernel_timings", False))
        self.timings = {}

    def build_acc(self, gaussians, rebuild=True):
        """tracer.py:198-216"""
        allow_bvh_update = (self._max_updates > 1) and not self._clamping
        rebuild_bvh = rebuild or self._clamping or self.num_update_bvh >= self._max_updates
        self.tracer_wrapper.build_bvh(
            gaussians.positions.view(-1, 3).contiguous(), gaussians.rotation_activation(gaussians.rotation).view(-1, 4).contiguous(),
            gaussians.scale_activation(gaussians.scale).view(-1, 3).contiguous(),
            gaussians.density_activation(gaussians.density).view(-1, 1).contiguous(), rebuild_bvh, allow_bvh_update)
        self.num_update_bvh = 0 if rebuild_bvh else self.num_update_bvh + 1

    def render(self, gaussians, gpu_batch, train=False, frame_id=0):
        """tracer.py:218-255"""
        start = end = None
        if self._timings_on:
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
        # hit lists for the backward's replay are only worth recording when a backward can follow
        self.tracer_wrapper.set_replay(bool(train) or torch.is_grad_enabled(), gpu_batch.rays_ori.device)
        pred_features, pred_opacity, pred_dist, pred_normals, hits_count, mog_visibility = Tracer._Autograd.apply(
            self.tracer_wrapper, frame_id, gpu_batch.T_to_world.contiguous(), gpu_batch.rays_ori.contiguous(), gpu_batch.rays_dir.contiguous(),
            gaussians.positions.contiguous(), gaussians.get_rotation().contiguous(), gaussians.get_scale().contiguous(),
            gaussians.get_density().contiguous(), gaussians.get_features().contiguous(), Tracer.RenderOpts.DEFAULT,
            gaussians.n_active_features, self._min_transmittance)
        frame_ms = 0.0
        if self._timings_on:
            end.record()
            end.synchronize()
            frame_ms = start.elapsed_time(end)
            self.timings["forward_render"] = frame_ms
        return {
            "pred_features": pred_features,
            "pred_opacity": pred_opacity,
            "pred_dist": pred_dist,
            "pred_normals": torch.nn.functional.normalize(pred_normals, dim=3),
            "hits_count": hits_count,
            "frame_time_ms": frame_ms,
            "mog_visibility": mog_visibility,
        }
