"""Timeline of the host-fed weight loop (bench.TrainWorkload.e2e_loop): GPU events of 4 consecutive steps (kernels
collapsed to first / last of each graph replay, memcpys listed) and the CPU time of each phase of the loop."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

torch.cuda.set_device(0)
args = argparse.Namespace(chanstr="8,16,8,8", resolution=1024, train_blocks=0, no_graph=False)
pts, origins = bench.make_cloud(1024)
tw = bench.TrainWorkload(args, 0, 1, pts, origins)
tw.e2e_loop(0, 5)
torch.cuda.synchronize()
t0 = time.perf_counter(); tw.e2e_loop(5, 50); torch.cuda.synchronize(); t1 = time.perf_counter()
print("e2e loop: %.1f us per step" % ((t1 - t0) / 50 * 1e6))
for i in range(5):
    tw.step(i, False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    tw.step(i, False)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("device-resident loop: %.1f us per step" % ((t1 - t0) / 50 * 1e6))
# CPU phases
from nvfpcc_b200 import trainer
f = tw.feeder
ph = {"take": 0.0, "step": 0.0, "release": 0.0, "submit": 0.0, "stats": 0.0}
f.submit(tw.batch_idx(0))
for i in range(50):
    a = time.perf_counter(); (gt, dst), slot = f.take(); b = time.perf_counter()
    st = tw.ws.step(tw.emb_batches[i % len(tw.emb_batches)], gt, dst, q=1, n_pts=tw.npts[i % len(tw.emb_batches)]); c = time.perf_counter()
    f.release(slot); d = time.perf_counter()
    f.submit(tw.batch_idx(i + 1)); e = time.perf_counter()
    f.read_stats(st); g = time.perf_counter()
    for k, v in zip(ph, (b - a, c - b, d - c, e - d, g - e)):
        ph[k] += v
f.drain(); torch.cuda.synchronize()
print("CPU us per step:", {k: round(v / 50 * 1e6, 1) for k, v in ph.items()})
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tw.e2e_loop(100, 3); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tw.e2e_loop(110, 4); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    n = e.name
    if "Memcpy" in n or "latent_fwd" in n or "adam_tick" in n or "elementwise" in n or "copy" in n.lower():
        print("%9.1f %8.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, n[:90]))
