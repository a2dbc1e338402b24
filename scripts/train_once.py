"""A few weight-loop train steps on the synthetic vox10 blocks (for ncu captures):
the last step is bracketed by cudaProfilerStart/Stop (use ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--chanstr", default="8,16,8,8")
ap.add_argument("--resolution", type=int, default=1024)
ap.add_argument("--train-blocks", type=int, default=32)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()
torch.cuda.set_device(0)
pts, origins = bench.make_cloud(a.resolution)
tw = bench.TrainWorkload(a, 0, 1, pts, origins)
for i in range(a.warm):
    tw.step(i, False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tw.step(a.warm, False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("train step done, loss", float(tw.last_loss))
