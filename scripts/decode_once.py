"""One decode of the synthetic cloud bracketed by cudaProfilerStart/Stop (use ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--chanstr", default="8,16,8,8")
ap.add_argument("--resolution", type=int, default=1024)
a = ap.parse_args()
torch.cuda.set_device(0)
pts, origins = bench.make_cloud(a.resolution)
dw = bench.DecodeWorkload(a.chanstr, 0, 1, pts, origins, 0.65)
for i in range(2):
    dw.step(i, False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
dw.step(2, False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("decoded", dw.n_all, "blocks")
