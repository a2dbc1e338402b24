"""Prints the epoch history of trainer.fit on a small synthetic cloud (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvfpcc_b200 import grids, network, synth, trainer
pts = synth.sphere_shell_points(256)
origins = synth.leaf_origins(pts)[:40]
g = grids.build_grids(pts, origins, want_dist32=True)
for lmbda, ep in ((200.0, 12), (1.0, 12)):
    network.set_seed(synth.synthetic_seed()); torch.manual_seed(0)
    net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8").cuda()
    emb, hist = trainer.fit(net, g["gt"], g["dist32"], epochs=ep, batchsize=16, lr=1e-3, lmbda=lmbda, w1=10.0, w2=57.0, phase_change=6)
    for h in hist:
        print(lmbda, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in h.items()})
