"""Phase timing of the multi-rank decode e2e path (torchrun): decode, size exchange, point-to-point gather, D2H."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from nvfpcc_b200 import dist as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pts, origins = bench.make_cloud(1024)
dw = bench.DecodeWorkload("8,16,8,8", rank, world, pts, origins, 0.65)
def sync():
    torch.cuda.synchronize()
    return time.perf_counter()
for it in range(5):
    dist.barrier(); t0 = sync()
    r = dw.net.decode_points(dw.lat_host, dw.org_host, dw.thh, return_host=False)
    t1 = sync()
    c, n = D.gather_points(r["coords"], r["counts"])
    t2 = sync()
    if c is not None:
        ch = c.cpu()
    t3 = sync()
    if rank == 0:
        print("iter %d: decode %.2f ms, gather %.2f ms, d2h %.2f ms (%d points)" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, ch.shape[0]))
dist.barrier()
dist.destroy_process_group()
