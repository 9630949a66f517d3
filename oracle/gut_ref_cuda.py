"""ctypes front-end of oracle/_ref/libgut_ref_cuda.so: the REFERENCE's own 3DGUT renderer (threedgut_tracer/src/gutRenderer.cu and
everything it includes: projectOnTiles, CUB scan, expandTileProjections, the 44-bit CUB radix sort, tile ranges, render,
renderBackward with the hand-written adjoint, projectBackward) compiled unmodified for sm_100a from /root/reference, with the slangc
output replaced by oracle/ref_cuda/threedgutSlang.cuh (see that file and oracle/ref_cuda/ref_cuda_driver.cu).

TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/ (parity pin on the GPU) and by bench.py's `reference_gpu` block (the same-box GPU
denominator).  The library is built in the build container (`make -C oracle refcuda`) and travels to the GPU box as a prebuilt file;
nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libgut_ref_cuda.so")
_LIB = None

STAGES = ("project", "prepare_expand", "expand", "sort", "render", "forward", "render_backward", "project_backward", "backward")


def available() -> bool:
    if os.path.exists(_SO):
        return True
    if os.path.isdir("/root/reference/threedgut_tracer"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "refcuda"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            return False
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libgut_ref_cuda.so not built (make -C oracle refcuda, needs /root/reference)")
        _LIB = C.CDLL(_SO)
        _LIB.refcuda_create.restype = C.c_void_p
        _LIB.refcuda_last_error.restype = C.c_char_p
        _LIB.refcuda_debug_copy.restype = C.c_int64
    return _LIB


def _fp(a):
    return np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))


class ReferenceRaster:
    """The reference's SplatRaster.trace / trace_bwd over torch CUDA tensors (pinhole, global shutter, default render config)."""

    def __init__(self):
        self.h = C.c_void_p(lib().refcuda_create())
        if not self.h:
            raise RuntimeError("refcuda_create failed (no CUDA device)")

    def close(self):
        if self.h:
            lib().refcuda_destroy(self.h)
            self.h = None

    def set_timing(self, on: bool):
        lib().refcuda_set_timing(self.h, C.c_int(int(on)))

    def stage_times(self):
        ms = (C.c_float * 9)()
        lib().refcuda_stage_times(self.h, ms)
        return dict(zip(STAGES, [float(v) for v in ms]))

    def _cam(self, fx, fy, cx, cy, pose):
        pose = np.asarray(pose, np.float32).reshape(7)
        return _fp([fx, fy]), _fp([cx, cy]), _fp(pose), _fp(pose)

    def trace(self, torch, stream, frame, sph_degree, particles, sph, width, height, fx, fy, cx, cy, pose, rays_o, rays_d):
        n = particles.shape[0]
        dev = particles.device
        rgba = torch.empty((height, width, 4), device=dev)
        dist = torch.empty((height, width, 1), device=dev)
        hits = torch.empty((height, width, 1), device=dev)
        vis = torch.empty((n, 1), device=dev)
        f, p, p0, p1 = self._cam(fx, fy, cx, cy, pose)
        rc = lib().refcuda_forward(self.h, C.c_void_p(stream), C.c_uint32(frame), C.c_int(sph_degree), C.c_int64(n), C.c_void_p(particles.data_ptr()),
                                   C.c_void_p(sph.data_ptr()), C.c_int(width), C.c_int(height), f, p, p0, p1, C.c_void_p(rays_o.data_ptr()),
                                   C.c_void_p(rays_d.data_ptr()), C.c_void_p(rgba.data_ptr()), C.c_void_p(dist.data_ptr()),
                                   C.c_void_p(hits.data_ptr()), C.c_void_p(vis.data_ptr()))
        if rc:
            raise RuntimeError("refcuda_forward: " + lib().refcuda_last_error(self.h).decode())
        return rgba, dist, hits, vis

    def trace_bwd(self, torch, stream, frame, sph_degree, particles, sph, width, height, fx, fy, cx, cy, pose, rays_o, rays_d, rgba, d_rgba, dist,
                  d_dist, out=None):
        n = particles.shape[0]
        dev = particles.device
        dp, ds = out if out is not None else (torch.empty((n, 12), device=dev), torch.empty((n, 48), device=dev))
        f, p, p0, p1 = self._cam(fx, fy, cx, cy, pose)
        rc = lib().refcuda_backward(self.h, C.c_void_p(stream), C.c_uint32(frame), C.c_int(sph_degree), C.c_int64(n), C.c_void_p(particles.data_ptr()),
                                    C.c_void_p(sph.data_ptr()), C.c_int(width), C.c_int(height), f, p, p0, p1, C.c_void_p(rays_o.data_ptr()),
                                    C.c_void_p(rays_d.data_ptr()), C.c_void_p(rgba.data_ptr()), C.c_void_p(d_rgba.data_ptr()),
                                    C.c_void_p(dist.data_ptr()), C.c_void_p(d_dist.data_ptr()), C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr()))
        if rc:
            raise RuntimeError("refcuda_backward: " + lib().refcuda_last_error(self.h).decode())
        return dp, ds

    def debug(self, what: str, n: int, tiles: int):
        idx = {"tiles_count": (0, np.uint32), "sorted_keys": (1, np.uint64), "sorted_values": (2, np.uint32), "ranges": (3, np.uint32),
               "depth": (4, np.float32), "rgb": (5, np.float32)}[what]
        nbytes = lib().refcuda_debug_copy(self.h, C.c_int(idx[0]), None, C.c_int64(n), C.c_int64(tiles))
        if nbytes < 0:
            raise RuntimeError("refcuda_debug_copy: no forward context")
        out = np.empty(nbytes // np.dtype(idx[1]).itemsize, idx[1])
        if nbytes:
            got = lib().refcuda_debug_copy(self.h, C.c_int(idx[0]), out.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int64(tiles))
            if got != nbytes:
                raise RuntimeError("refcuda_debug_copy failed")
        if what == "ranges":
            out = out.reshape(-1, 2)
        if what == "rgb":
            out = out.reshape(-1, 3)
        return out
