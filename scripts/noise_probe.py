"""q = 1 loss statistics of the fused step (in-kernel Philox noise) vs the autograd path (torch.rand): same state,
same batch, lr = 0, so the only difference between steps is the noise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nvfpcc_b200 import network, synth, trainer
from oracle.gen_golden import fixture_inputs
g = np.load("tests/golden/nvf_A.npz")
gt, dist = torch.from_numpy(g["tr_gt"]).float().cuda(), torch.from_numpy(g["tr_dist"]).float().cuda()
for state in ("fixture", "init"):
    for fused in (False, True):
        network.set_seed(synth.synthetic_seed())
        net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8")
        if state == "fixture":
            net.load_state_dict(fixture_inputs("A")["sd"])
        net = net.cuda()
        opt = trainer.FusedAdam(net.parameters(), lr=0.0)
        ws = trainer.WeightStep(net, opt, 2, 849338.0, 200.0, 10.0, 57.0, use_graph=False, fused=fused)
        emb = torch.ones(2, 3, 2, 2, 2).cuda()
        torch.manual_seed(0)
        L = torch.stack([ws.step(emb, gt, dist, q=1).clone() for _ in range(24)]).cpu()
        print(state, "fused" if fused else "autograd", "loss mean %.1f std %.1f | bce mean %.1f std %.1f | b_latent %.4f" % (
            L[:, 0].mean(), L[:, 0].std(), L[:, 1].mean(), L[:, 1].std(), L[:, 4].mean()))
