"""debug helper: one reference forward+backward on a small scene with intermediate buffers printed"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3dgrut_b200")]
import numpy as np, torch
import scenes
from oracle import gut_ref_cuda as grc
sc = scenes.scene_c1()
dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ro, rd = sc.rays()
particles, sph, tro, trd = t(sc.particles), t(sc.sph), t(ro), t(rd)
pose = scenes.pose7_from_c2w(sc.camera(1, 8))
print("pose", pose, "fx", sc.fx, sc.fy, sc.cx, sc.cy, sc.width, sc.height)
out = np.zeros(20, np.float32)
p7 = np.asarray(pose, np.float32)
grc.lib().refcuda_debug_camera(C.c_int(sc.width), C.c_int(sc.height), grc._fp([sc.fx, sc.fy]), grc._fp([sc.cx, sc.cy]), grc._fp(p7), grc._fp(p7), out.ctypes.data_as(C.c_void_p))
print("ref camera: view cols", out[:12], "pos", out[12:15], "params", out[15:])
import b200_native as nat
cam = nat.Camera(); cam.width, cam.height = sc.width, sc.height
cam.principal[:] = [sc.cx, sc.cy]; cam.focal[:] = [sc.fx, sc.fy]
cam.pose_start[:] = [float(v) for v in pose]; cam.pose_end[:] = [float(v) for v in pose]
print("our camera pos", nat.camera_position(cam))
rr = grc.ReferenceRaster()
s = torch.cuda.current_stream(dev).cuda_stream
rgba, dist, hits, vis = rr.trace(torch, s, 0, sc.sph_degree, particles, sph, sc.width, sc.height, sc.fx, sc.fy, sc.cx, sc.cy, pose, tro, trd)
torch.cuda.synchronize()
print("forward ok", float(rgba.sum()), float(hits.sum()), "vis", int(vis.view(torch.int32).sum()))
tc = rr.debug("tiles_count", sc.n, 64); dp = rr.debug("depth", sc.n, 64)
print("tiles_count sum", int(tc.sum()), "nonzero", int((tc > 0).sum()), "depth min/max", float(dp.min()), float(dp.max()))
if tc.sum() > 0:
    d_rgba = torch.randn_like(rgba); d_dist = torch.zeros_like(dist)
    dpp, ds = rr.trace_bwd(torch, s, 0, sc.sph_degree, particles, sph, sc.width, sc.height, sc.fx, sc.fy, sc.cx, sc.cy, pose, tro, trd, rgba, d_rgba, dist, d_dist)
    torch.cuda.synchronize()
    print("backward ok", float(dpp.abs().sum()), float(ds.abs().sum()))
