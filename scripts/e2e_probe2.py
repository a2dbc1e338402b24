import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
torch.cuda.set_device(0)
args = argparse.Namespace(chanstr="8,16,8,8", resolution=1024, train_blocks=0, no_graph=False)
pts, origins = bench.make_cloud(1024)
tw = bench.TrainWorkload(args, 0, 1, pts, origins)
tw.e2e_loop(0, 5); torch.cuda.synchronize()
f = tw.feeder
ph = {}
def T(name, t0):
    ph[name] = ph.get(name, 0.0) + time.perf_counter() - t0
def submit(idx):
    s = f._head & 1
    t = time.perf_counter()
    if f._head >= 2: f.h2d_done[s].synchronize()
    T("sync_h2d_done", t); t = time.perf_counter()
    torch.index_select(f.gt_host, 0, idx, out=f.pin[s][0]); T("gather_gt", t); t = time.perf_counter()
    torch.index_select(f.dist_host, 0, idx, out=f.pin[s][1]); T("gather_dist", t); t = time.perf_counter()
    with torch.cuda.stream(f.copy_stream):
        if f._head >= 2: f.copy_stream.wait_event(f.consumed[s])
        f.dev[s][0].copy_(f.pin[s][0], non_blocking=True); T("enqueue_gt", t); t = time.perf_counter()
        f.dev[s][1].copy_(f.pin[s][1], non_blocking=True); T("enqueue_dist", t); t = time.perf_counter()
        f.h2d_done[s].record(f.copy_stream); f.ready[s].record(f.copy_stream)
    f._head += 1
    T("events", t)
submit(tw.batch_idx(0))
for i in range(50):
    (gt, dst), slot = f.take()
    st = tw.ws.step(tw.emb_batches[i % len(tw.emb_batches)], gt, dst, q=1, n_pts=tw.npts[i % len(tw.emb_batches)])
    f.release(slot)
    submit(tw.batch_idx(i + 1))
    f.read_stats(st)
f.drain(); torch.cuda.synchronize()
print({k: round(v / 50 * 1e6, 1) for k, v in ph.items()})
print("threads", torch.get_num_threads(), "gt_host pinned", tw.gt_host.is_pinned(), tw.gt_host.dtype, tw.gt_host.shape)
