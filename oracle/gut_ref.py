"""ctypes front-end of oracle/_ref/libgut_ref.so: the reference's own hand-written CUDA math compiled
for the host from /root/reference (oracle/ref_gut.cpp).  TEST INFRASTRUCTURE ONLY; exists only where
/root/reference is mounted (the build container) -- tests skip when it is absent."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libgut_ref.so")
_LIB = None


def available() -> bool:
    if os.path.exists(_SO):
        return True
    if os.path.isdir("/root/reference/threedgut_tracer"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            return False
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libgut_ref.so not built (needs /root/reference)")
        _LIB = C.CDLL(_SO)
        _LIB.ref_hit_fwd.restype = C.c_int
    return _LIB


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sensor_matrices(p0, p1):
    p0, p1 = _f(p0), _f(p1)
    view, inv, pos = np.zeros(12, np.float32), np.zeros(12, np.float32), np.zeros(3, np.float32)
    lib().ref_sensor_matrices(_p(p0), _p(p1), _p(view), _p(inv), _p(pos))
    return view.reshape(4, 3), inv.reshape(4, 3), pos


def set_camera_model(fisheye=None):
    """Camera model of the following project() / expand() calls: None = OpenCV pinhole, (k1,k2,k3,k4,max_angle) = OpenCV fisheye."""
    if fisheye is None:
        lib().ref_set_camera_model(C.c_int(0), None)
    else:
        f = _f(np.asarray(fisheye, np.float32).reshape(5))
        lib().ref_set_camera_model(C.c_int(1), _p(f))


def set_rolling_shutter(kind=0):
    """0 = global shutter, 1..4 = rolling top-to-bottom / left-to-right / bottom-to-top / right-to-left (the oracle's numbering)."""
    lib().ref_set_shutter(C.c_int(4 if kind == 0 else kind - 1))


def set_ftheta(ftheta):
    """f-theta model for the following project() calls (dict as in gut_oracle.make_camera); set_camera_model(None) resets."""
    v = _f(np.concatenate([ftheta["bw"], ftheta["fw"], ftheta["cde"], [ftheta["max_angle"]]]).astype(np.float32))
    lib().ref_set_ftheta(C.c_int(int(ftheta["reference_poly"])), _p(v))


def project(particles, sph, degree, width, height, focal, pp, p0, p1):
    particles, sph, focal, pp, p0, p1 = map(_f, (particles, sph, focal, pp, p0, p1))
    n = particles.shape[0]
    out = dict(tiles_count=np.zeros(n, np.uint32), proj_pos=np.zeros((n, 2), np.float32),
               conic_opacity=np.zeros((n, 4), np.float32), extent=np.zeros((n, 2), np.float32),
               depth=np.zeros(n, np.float32), rgb=np.zeros((n, 3), np.float32), visibility=np.zeros(n, np.int32))
    lib().ref_project(C.c_int64(n), _p(particles), _p(sph), C.c_int(degree), C.c_int(width), C.c_int(height), _p(focal), _p(pp),
                      _p(p0), _p(p1), _p(out["tiles_count"], C.c_uint32), _p(out["proj_pos"]), _p(out["conic_opacity"]),
                      _p(out["extent"]), _p(out["depth"]), _p(out["rgb"]), _p(out["visibility"], C.c_int))
    return out


def expand(width, height, tiles_count, proj_pos, conic_opacity, extent, depth):
    n = tiles_count.shape[0]
    offs = np.cumsum(tiles_count.astype(np.uint64)).astype(np.uint32)
    total = int(offs[-1]) if n else 0
    keys, vals = np.zeros(max(total, 1), np.uint64), np.zeros(max(total, 1), np.uint32)
    lib().ref_expand(C.c_int64(n), C.c_int(width), C.c_int(height), _p(offs, C.c_uint32), _p(_f(proj_pos)), _p(_f(conic_opacity)),
                     _p(_f(extent)), _p(_f(depth)), _p(keys, C.c_uint64), _p(vals, C.c_uint32))
    return keys[:total], vals[:total]


def sph(degree, coeffs, direction, clamped=False):
    out = np.zeros(3, np.float32)
    lib().ref_sph(C.c_int(degree), _p(_f(coeffs)), _p(_f(direction)), C.c_int(int(clamped)), _p(out))
    return out


def sph_bwd(degree, direction, rgb_grad, unclamped):
    g = np.zeros(48, np.float32)
    lib().ref_sph_bwd(C.c_int(degree), _p(_f(direction)), _p(_f(rgb_grad)), _p(_f(unclamped)), _p(g))
    return g


def hit_fwd(degree, ro, rd, particle, rgb, T, Cacc, D):
    T_, D_ = C.c_float(T), C.c_float(D)
    Cc = _f(Cacc).copy()
    acc = lib().ref_hit_fwd(C.c_int(degree), _p(_f(ro)), _p(_f(rd)), _p(_f(particle)), _p(_f(rgb)), C.byref(T_), _p(Cc), C.byref(D_))
    return acc, T_.value, Cc, D_.value


def hit_bwd(degree, ro, rd, particle, rgb, min_t, Tint, T, Tgrad, Cint, Cacc, Cgrad, Dint, D, Dgrad):
    T_, D_ = C.c_float(T), C.c_float(D)
    Cc = _f(Cacc).copy()
    grad, rg = np.zeros(12, np.float32), np.zeros(3, np.float32)
    lib().ref_hit_bwd(C.c_int(degree), _p(_f(ro)), _p(_f(rd)), _p(_f(particle)), _p(_f(rgb)), C.c_float(min_t), C.c_float(Tint),
                      C.byref(T_), C.c_float(Tgrad), _p(_f(Cint)), _p(Cc), _p(_f(Cgrad)), C.c_float(Dint), C.byref(D_), C.c_float(Dgrad),
                      _p(grad), _p(rg))
    return grad, rg, T_.value, Cc, D_.value
