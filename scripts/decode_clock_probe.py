"""Diagnostic: per-step kernel time of the fused decode against NVML clocks/power over a long run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--chanstr", default="8,16,8,8")
ap.add_argument("--resolution", type=int, default=1024)
ap.add_argument("--reps", type=int, default=60)
ap.add_argument("--flush", type=int, default=1)
a = ap.parse_args()
torch.cuda.set_device(0)
pts, origins = bench.make_cloud(a.resolution)
dw = bench.DecodeWorkload(a, 0, 1, pts, origins)
flush = torch.zeros(128 * 1024 * 1024, device="cuda")
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
for i in range(2):
    dw.step(i, False)
torch.cuda.synchronize()
dw.kernel_events = []
rows = []
for i in range(a.reps):
    if a.flush:
        bench.flush_l2(flush)
    dw.step(i, False)
    torch.cuda.synchronize()
    rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                 pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
ms = [x.elapsed_time(y) for x, y in dw.kernel_events]
for i in range(0, a.reps, 4):
    print("step %3d kernel %.2f ms  sm %d MHz  %.0f W  reasons 0x%x" % (i, ms[i], rows[i][0], rows[i][1], rows[i][2]))
