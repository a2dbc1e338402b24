"""GPU test of the epoch driver trainer.fit (train(), NVFPCC.py:103-296, on the sync-free steps): the grids
built on the GPU feed the training loop, the loss goes down, checkpoints are the reference's files, and the
trained network + embeddings go through encode -> pack -> decode."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fit_trains_checkpoints_and_round_trips(gpu, tmp_path):
    from nvfpcc_b200 import codec, grids, network, synth, trainer
    pts = synth.sphere_shell_points(256)
    origins = synth.leaf_origins(pts)[:40]
    g = grids.build_grids(pts, origins, want_dist32=True)
    gt, dist = g["gt"], g["dist32"]
    network.set_seed(synth.synthetic_seed())
    torch.manual_seed(0)
    net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8").cuda()
    logs = []
    emb, hist = trainer.fit(net, gt, dist, epochs=4, batchsize=16, lr=1e-3, lmbda=200.0, w1=10.0, w2=57.0, wemb=5.0,
                            phase_change=3, checkpoint_dir=str(tmp_path / "train"), save_every=2, log=logs.append)
    assert len(hist) == 4 and logs == hist
    assert [h["q"] for h in hist] == [1, 1, 1, 2]
    assert all(np.isfinite(h["loss"]) and np.isfinite(h["emb_loss"]) for h in hist)
    assert hist[2]["bce"] < 0.5 * hist[0]["bce"]                              # it learns (noise phase, q = 1)
    # q = 2 rounds the still tiny kernel deltas (|d| < 1/32 after 9 steps) back to zero: the reference's behaviour
    assert hist[3]["bce"] > hist[2]["bce"]
    assert abs(hist[0]["n_pts"] * 3 - float(gt.sum()) * (3 * 16 / 40)) < 0.35 * float(gt.sum())   # ~16 leaves per batch
    assert emb.shape == (40, 3, 2, 2, 2) and emb.requires_grad and not torch.equal(emb, torch.ones_like(emb))
    # checkpoints: NVFPCC.py:284-291
    files = sorted(os.listdir(tmp_path / "train"))
    assert files == ["0000.ckpt", "0000_emb.ckpt", "0002.ckpt", "0002_emb.ckpt"]
    sd = torch.load(tmp_path / "train" / "0002.ckpt", map_location="cpu")
    assert list(sd.keys()) == list(net.state_dict().keys())
    e2 = torch.load(tmp_path / "train" / "0002_emb.ckpt", map_location="cpu")
    assert e2.shape == (40, 3, 2, 2, 2)
    # resume from the checkpoint (load_weights / load_emb of the reference CLI)
    network.set_seed(synth.synthetic_seed())
    net2 = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8")
    net2.load_state_dict(sd)
    net2 = net2.cuda()
    emb3, hist3 = trainer.fit(net2, gt, dist, epochs=4, start_epoch=3, batchsize=16, lr=1e-3, lmbda=200.0, w1=10.0,
                              w2=57.0, phase_change=3, emb=e2.detach())
    assert len(hist3) == 1 and hist3[0]["q"] == 2 and np.isfinite(hist3[0]["loss"])
    # trained model -> quantise (manipulate_weights.py) -> encode -> decode, bit-exact
    state = codec.quantize_state({k: v.detach().cpu() for k, v in net.state_dict().items()}, 16)
    net.load_state_dict(state, strict=False)
    enc = codec.encode(net, emb.detach(), origins, 0.5, weights_state=state)
    network.set_seed(synth.synthetic_seed())
    dec = codec.decode(enc["total_pack"], 3, "8,16,8,8", 0.5)
    assert np.array_equal(dec, enc["points"])


def test_gather_batch_equals_indexing(gpu):
    from nvfpcc_b200 import trainer
    g = torch.Generator().manual_seed(3)
    n_all, n = 57, 16
    emb = torch.randn(n_all, 3, 2, 2, 2, generator=g).cuda()
    gt = (torch.rand(n_all, 1, 32, 32, 32, generator=g) < 0.05).float().cuda()
    dist = torch.rand(n_all, 1, 32, 32, 32, generator=g).cuda()
    idx = torch.randint(0, n_all, (n,), generator=g).cuda()
    e, a, d = torch.zeros(n, 3, 2, 2, 2).cuda(), torch.zeros(n, 1, 32, 32, 32).cuda(), torch.zeros(n, 1, 32, 32, 32).cuda()
    trainer._gather_batch(emb, gt, dist, idx, e, a, d)
    assert torch.equal(e, emb[idx]) and torch.equal(a, gt[idx]) and torch.equal(d, dist[idx])
