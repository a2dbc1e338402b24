"""Multi-rank smoke of the rows built around the path (run under torchrun on >= 2 GPUs):
trainer.fit with rank-sharded leaves + all-reduced weight gradients, codec.encode / decode with rank-sharded
reconstruction.  Checks: weights stay identical on all ranks, the loss goes down, the gathered cloud equals the
single-process reconstruction of the same pack, bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from nvfpcc_b200 import codec, grids, network, synth, trainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pts = synth.sphere_shell_points(256)
origins = synth.leaf_origins(pts)[:63]      # 32 + 31 leaves: rank 1 meets a new batch size (15) at a step where rank 0 replays
g = grids.build_grids(pts, origins, want_dist32=True)
network.set_seed(synth.synthetic_seed()); torch.manual_seed(0)
net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8").cuda()
emb, hist = trainer.fit(net, g["gt"], g["dist32"], epochs=3, batchsize=16, lr=1e-3, lmbda=1.0, w1=10.0, w2=57.0, phase_change=10)
flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
ref = flat.clone(); dist.broadcast(ref, 0)
assert torch.equal(flat, ref), "weights diverged between ranks"
e0 = emb.detach().clone(); dist.broadcast(e0, 0)
assert torch.equal(emb.detach(), e0), "gathered embeddings differ between ranks"
assert all(np.isfinite(h["loss"]) for h in hist) and hist[-1]["bce"] < hist[0]["bce"], hist
state = codec.quantize_state({k: v.detach().cpu() for k, v in net.state_dict().items()}, 16)
net.load_state_dict(state, strict=False)
enc = codec.encode(net, emb.detach(), origins, 0.5, weights_state=state)
network.set_seed(synth.synthetic_seed())
dec = codec.decode(enc["total_pack"], 3, "8,16,8,8", 0.5)
if rank == 0:
    lat = net.get_latent_code(emb.detach().cuda())["quantized_latent"]
    single = net.decode_points(lat, torch.from_numpy(origins.astype(np.int32)).cuda(), 0.5, return_host=True)["coords"].numpy()
    assert np.array_equal(enc["points"], single) and np.array_equal(dec, single), "sharded reconstruction differs"
    print("peer all-reduce: %s" % ("on" if getattr(trainer, "_last_opt", None) is not None and trainer._last_opt._symm is not None else "off"))
    print("dist check ok: world %d, bce %.1f -> %.1f, %d points, latent stream %d B" % (
        world, hist[0]["bce"], hist[-1]["bce"], single.shape[0], len(enc["total_pack"]["latent_pack"]["latent_byte_stream"])))
else:
    assert dec is None and enc["points"] is None
dist.barrier()
dist.destroy_process_group()
