"""debug helper: one reference forward+backward on a small scene (run under compute-sanitizer on the GPU box)"""
import os, sys
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
rr = grc.ReferenceRaster()
s = torch.cuda.current_stream(dev).cuda_stream
print("forward...", flush=True)
rgba, dist, hits, vis = rr.trace(torch, s, 0, sc.sph_degree, particles, sph, sc.width, sc.height, sc.fx, sc.fy, sc.cx, sc.cy, pose, tro, trd)
torch.cuda.synchronize()
print("forward ok", float(rgba.sum()), float(hits.sum()), flush=True)
d_rgba = torch.randn_like(rgba); d_dist = torch.zeros_like(dist)
dp, ds = rr.trace_bwd(torch, s, 0, sc.sph_degree, particles, sph, sc.width, sc.height, sc.fx, sc.fy, sc.cx, sc.cy, pose, tro, trd, rgba, d_rgba, dist, d_dist)
torch.cuda.synchronize()
print("backward ok", float(dp.abs().sum()), float(ds.abs().sum()), flush=True)
