"""Concurrent timeline of ONE graph-replayed weight-loop step (CUPTI through torch.profiler; nsys is not in the image).

    python scripts/timeline.py [--chanstr 8,16,8,8] [--batch 16] [--out gpurun_out/timeline.txt]

ncu serialises kernels, so it cannot show what overlaps what; this prints every kernel of the last profiled replay
with its stream, start offset and duration, plus the step's wall time and the busy time per stream."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chanstr", default="8,16,8,8")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--out", default="gpurun_out/timeline.txt")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    args = argparse.Namespace(chanstr=a.chanstr, resolution=1024, train_blocks=128, no_graph=False)
    bench.HP["batch"] = a.batch
    pts, origins = bench.make_cloud(1024)
    tw = bench.TrainWorkload(args, 0, 1, pts, origins)
    for i in range(5):
        tw.step(i, False)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tw.step(5, False)           # CUPTI warm-up replay
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tw.step(6, False)           # exactly one replay in this profile
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    last = evs
    t0 = last[0].time_range.start
    lines = []
    busy = {}
    for e in last:
        st = getattr(e, "stream", None)
        if st is None:
            st = getattr(e, "device_resource_id", -1)
        d = e.time_range.end - e.time_range.start
        busy[st] = busy.get(st, 0.0) + d
        lines.append("%9.1f %8.1f  s%-3s %s" % (e.time_range.start - t0, d, st, e.name[:110]))
    wall = max(e.time_range.end for e in last) - t0
    hdr = ["step wall %.1f us, %d kernels; busy per stream: %s" % (
        wall, len(last), ", ".join("s%s %.1f" % (k, v) for k, v in sorted(busy.items(), key=lambda kv: -kv[1]))),
        "%9s %8s  %-4s %s" % ("start_us", "dur_us", "strm", "kernel")]
    txt = "\n".join(hdr + lines)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        f.write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
