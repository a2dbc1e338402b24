"""Experiment: one 16-block decoder fwd+loss+bwd chain on one stream vs two 8-block chains on two streams
(software pipelining of the latency-bound small layers of one half against the FMA-bound big layers of the other)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfpcc_b200 import _lib, network, synth
import bench

torch.cuda.set_device(0)
b = _lib.cuda_binding()
network.set_seed(synth.synthetic_seed())
net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8").cuda()
w = {k: v.detach().contiguous() for k, v in net.reconstructor.effective_weights(2).items()}
desc = b.desc(3, (8, 16, 8, 8))
N = 16
g = torch.Generator(device="cuda").manual_seed(0)
lat = torch.round(torch.randn(N, 3, 2, 2, 2, device="cuda", generator=g) * 3)
gt = (torch.rand(N, 1, 32, 32, 32, device="cuda", generator=g) < 0.03).float()
dist = torch.rand(N, 1, 32, 32, 32, device="cuda", generator=g) * 4 * (1 - gt)


def chain(sl):
    out, c1, c0, ws, keep = b.train_forward(desc, w, lat[sl])
    sums, seeds = b.loss_seeds(out, c1, c0, gt[sl], dist[sl])
    gl, grads = b.train_backward(desc, w, lat[sl], ws, seeds[0], seeds[1], seeds[2], True, True)
    return grads


def run(parts, streams):
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    for sl, s in zip(parts, streams):
        with torch.cuda.stream(s):
            chain(sl)
    for s in streams:
        cur.wait_stream(s)


def timeit(parts, streams, graph=True, reps=30):
    for _ in range(3):
        run(parts, streams)
    torch.cuda.synchronize()
    if graph:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run(parts, streams)
        fn = gr.replay
    else:
        fn = lambda: run(parts, streams)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


s1, s2, s3, s4 = (torch.cuda.Stream() for _ in range(4))
print("1 x 16 blocks, 1 stream : %.3f ms" % timeit([slice(0, 16)], [s1]))
print("2 x  8 blocks, 2 streams: %.3f ms" % timeit([slice(0, 8), slice(8, 16)], [s1, s2]))
print("2 x  8 blocks, 1 stream : %.3f ms" % timeit([slice(0, 8), slice(8, 16)], [s1, s1]))
print("4 x  4 blocks, 4 streams: %.3f ms" % timeit([slice(0, 4), slice(4, 8), slice(8, 12), slice(12, 16)], [s1, s2, s3, s4]))
print("1 x 32 blocks-equivalent (2 x 16 on 2 streams): %.3f ms" % timeit([slice(0, 16), slice(0, 16)], [s1, s2]))
