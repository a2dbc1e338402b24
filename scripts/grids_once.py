"""Timing probe of nvf_build_grids on the synthetic vox10 cloud (CUDA events around the call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, numpy as np, torch
from nvfpcc_b200 import grids, synth

ap = argparse.ArgumentParser()
ap.add_argument("--resolution", type=int, default=1024)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
pts = synth.sphere_shell_points(a.resolution)
origins = synth.leaf_origins(pts)
p = torch.from_numpy(pts).cuda(); o = torch.from_numpy(origins).cuda()
for _ in range(2):
    r = grids.build_grids(p, o, want_dist32=True, max_cells=origins.shape[0])
torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = grids.build_grids(p, o, want_dist32=True, check=False, max_cells=origins.shape[0]); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
n = origins.shape[0]
ms = float(np.median(ts))
print("blocks %d points %d: %.3f ms  (%.1f Mvox/s, output %.1f GB/s)" % (n, len(pts), ms, n * 32768 / ms / 1e3, n * 32768 * 13 / ms / 1e6))
print("occupied", int(r["gt"].sum().item()), "max dist", float(r["dist"].max().item()), "status", int(r["status"].item()))
