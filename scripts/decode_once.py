"""One decode of the synthetic vox10 cloud (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--chanstr", default="8,16,8,8")
ap.add_argument("--resolution", type=int, default=1024)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
torch.cuda.set_device(0)
pts, origins = bench.make_cloud(a.resolution)
dw = bench.DecodeWorkload(a, 0, 1, pts, origins)
for i in range(a.reps):
    dw.step(i, False)
torch.cuda.synchronize()
print("decoded", dw.n_all, "blocks ->", dw.points, "points")
