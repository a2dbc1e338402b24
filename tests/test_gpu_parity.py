"""GPU parity tests proper: the CUDA library, called through the C ABI, against
the oracle and the golden vectors (same assertions as the emulator tier)."""
import numpy as np
import pytest
import torch

from oracle import nvf_oracle as O
from oracle.gen_golden import fixture_inputs
from tests.helpers import check_decode_against_oracle, check_train_against_oracle, eff_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from nvfpcc_b200 import _lib
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return _lib.cuda_binding()


def test_decode_fused_A(gpu, golden_A):
    fx = fixture_inputs("A")
    lat = torch.cat([fx["latents"], fx["latents"].flip(0) * 0.5 + 1, fx["latents"][:1] * -1.0], 0).round()
    check_decode_against_oracle(gpu, fx, lat, thh=0.5, golden=golden_A, dev="cuda")


def test_decode_layerwise_B(gpu, golden_B):
    fx = fixture_inputs("B")
    check_decode_against_oracle(gpu, fx, fx["latents"], thh=0.5, golden=golden_B, dev="cuda")


def test_decode_many_blocks_deterministic_and_batch_invariant(gpu):
    """rc_enc.ply == rc_dec.ply (README.md:63): the same latent must give bit-identical
    probabilities whatever the batch composition / CTA it lands on, run to run."""
    fx = fixture_inputs("A")
    desc = gpu.desc(fx["ch"], fx["channels"])
    w = eff_weights(fx["sd"], 2, "cuda")
    g = torch.Generator().manual_seed(3)
    lat = torch.round(torch.randn(333, 3, 2, 2, 2, generator=g) * 3).cuda()
    org = (torch.arange(333 * 3, dtype=torch.int32).reshape(333, 3) * 32).cuda()
    a = gpu.decode(desc, w, lat, org, 0.5, want_prob=True)
    b = gpu.decode(desc, w, lat, org, 0.5, want_prob=True)
    assert torch.equal(a["prob"], b["prob"]) and torch.equal(a["coords"], b["coords"])
    perm = torch.randperm(333, generator=g).cuda()
    c = gpu.decode(desc, w, lat[perm], org[perm], 0.5, want_prob=True)
    assert torch.equal(c["prob"], a["prob"][perm])
    one = gpu.decode(desc, w, lat[7:8], org[7:8], 0.5, want_prob=True)
    assert torch.equal(one["prob"][0], a["prob"][7])
    # encode->decode round trip property at scale: points of batch == concatenation of per-block points
    off = torch.cumsum(a["counts"].long(), 0)
    assert int(off[-1]) == a["coords"].shape[0]
    s = int(off[6])
    assert torch.equal(one["coords"], a["coords"][s:s + int(a["counts"][7])])


def test_decode_empty_and_cap_overflow(gpu):
    fx = fixture_inputs("A")
    desc = gpu.desc(fx["ch"], fx["channels"])
    w = eff_weights(fx["sd"], 2, "cuda")
    r = gpu.decode(desc, w, torch.zeros(0, 3, 2, 2, 2).cuda(), torch.zeros(0, 3, dtype=torch.int32).cuda(), 0.5)
    assert r["coords"].shape == (0, 3) and int(r["total"]) == 0
    lat = fx["latents"][:1].cuda()
    full = gpu.decode(desc, w, lat, None, 0.3)
    small = gpu.decode(desc, w, lat, None, 0.3, cap=5)
    assert int(full["total"]) > 5 and torch.equal(full["coords"], small["coords"])


def test_fused_A_equals_layerwise_A(gpu):
    """The fused decode kernel and the layer-wise kernels are independent implementations."""
    fx = fixture_inputs("A")
    desc = gpu.desc(fx["ch"], fx["channels"])
    w = eff_weights(fx["sd"], 2, "cuda")
    lat = torch.round(torch.randn(5, 3, 2, 2, 2, generator=torch.Generator().manual_seed(11)) * 3).cuda()
    a = gpu.decode(desc, w, lat, None, 0.5, want_prob=True)
    out, cls1, cls0, ws, keep = gpu.train_forward(desc, w, lat)
    np.testing.assert_allclose(a["prob"].cpu().numpy(), out.cpu().numpy(), rtol=0, atol=2e-6)


def test_train_forward_backward_A(gpu, golden_A):
    check_train_against_oracle(gpu, fixture_inputs("A"), golden_A, dev="cuda")


def test_train_forward_backward_B(gpu, golden_A):
    fx = fixture_inputs("B")
    check_train_against_oracle(gpu, fx, golden_A, dev="cuda", blocks=[1])


def test_backward_flags_skip_work(gpu, golden_A):
    fx = fixture_inputs("A")
    desc = gpu.desc(fx["ch"], fx["channels"])
    w = eff_weights(fx["sd"], 2, "cuda")
    lat = torch.round(torch.randn(2, 3, 2, 2, 2, generator=torch.Generator().manual_seed(1)) * 3).cuda()
    out, cls1, cls0, ws, keep = gpu.train_forward(desc, w, lat)
    g = [torch.randn_like(t) * 1e-3 for t in (out, cls1, cls0)]
    gl_full, gw_full = gpu.train_backward(desc, w, lat, ws, g[0], g[1], g[2], True, True)
    gl_only, none = gpu.train_backward(desc, w, lat, ws, g[0], g[1], g[2], False, True)
    assert none == {} and torch.equal(gl_full, gl_only)
    nol, gw_only = gpu.train_backward(desc, w, lat, ws, g[0], g[1], g[2], True, False)
    assert nol is None
    for k in gw_full:
        assert torch.equal(gw_full[k], gw_only[k]), k


def _make_step(graph):
    from nvfpcc_b200 import network, synth, trainer
    network.set_seed(synth.synthetic_seed())
    net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8").cuda()
    net.entropy_coder.noise_scale = 0.0
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
    return net, trainer.WeightStep(net, opt, batch=2, n_total=849338.0, lmbda=200.0, w1=10.0, w2=57.0, use_graph=graph)


def test_graphed_weight_step_equals_eager(gpu, golden_A):
    """The CUDA-graph replay of the weight-loop step (NVFPCC.py:149-223) must walk the same
    trajectory as the eager step: q=2 and zero latent noise make both deterministic."""
    gt = torch.from_numpy(golden_A["tr_gt"]).float().cuda()
    dist = torch.from_numpy(golden_A["tr_dist"]).float().cuda()
    emb = torch.ones(2, 3, 2, 2, 2).cuda()
    stats = {}
    params = {}
    for graph in (False, True):
        net, ws = _make_step(graph)
        hist = []
        for i in range(4):
            hist.append(ws.step(emb, gt if i % 2 == 0 else gt.flip(0), dist if i % 2 == 0 else dist.flip(0), q=2).clone())
        stats[graph] = torch.stack(hist).cpu()
        params[graph] = [p.detach().cpu().clone() for p in net.parameters()]
        if graph:
            assert ws.launches_per_step > 20
    assert torch.isfinite(stats[True]).all()
    np.testing.assert_allclose(stats[True].numpy(), stats[False].numpy(), rtol=1e-5, atol=1e-6)
    for a, b in zip(params[True], params[False]):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-6)
    # the loss of the 3rd visit of a batch is below its 1st visit (Adam is actually stepping)
    assert stats[True][2, 0] < stats[True][0, 0]


def test_weight_step_matches_oracle_step(gpu, golden_A):
    """One eager WeightStep == one oracle step (loss terms and post-Adam weights)."""
    from nvfpcc_b200 import network, synth
    gt = torch.from_numpy(golden_A["tr_gt"]).float()
    dist = torch.from_numpy(golden_A["tr_dist"]).float()
    net, ws = _make_step(False)
    sd0 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    emb = torch.ones(2, 3, 2, 2, 2)
    st = ws.step(emb.cuda(), gt.cuda(), dist.cuda(), q=2).cpu()
    # oracle: same state, zero latent noise (x + (0.5 - 0.5)), autograd + Adam
    sd = {k: v.clone() for k, v in sd0.items()}
    params = {k: v.requires_grad_(True) for k, v in sd.items() if not k.endswith(("_init", "pedestal"))}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    res = O.net_forward(emb, sd, "train", 2, latent_noise=torch.full((2, 3, 2, 2, 2), 0.5))
    L = O.train_loss(res, gt, dist, gt.sum(), 849338.0, 200.0, 10.0, 57.0)
    L["loss"].backward()
    opt.step()
    assert abs(st[0].item() - L["loss"].item()) <= 1e-4 * abs(L["loss"].item())
    new = net.state_dict()
    for k, v in params.items():
        got = new[k].detach().cpu()
        # Adam's first step moves every weight by ~lr * sign(grad): compare the update direction where it is defined
        d_got, d_ref = got - sd0[k], v.detach() - sd0[k]
        big = d_ref.abs() > 5e-4
        flips = (torch.sign(d_got[big]) != torch.sign(d_ref[big])).float().mean().item() if big.any() else 0.0
        assert flips <= 2e-3, (k, flips)
        np.testing.assert_allclose(got.numpy(), v.detach().numpy(), rtol=0, atol=2.1e-3)


@pytest.mark.parametrize("q", [0, 1, 2])
def test_fused_param_prep_matches_torch_path(gpu, q):
    """nvf_param_prep (+ backward) against the torch ops it replaces (utils/network.py:606-620 weight
    preparation, gdn_3d.py:143-150 reparametrisation, :4777-4778 net_bits) incl. autograd."""
    from nvfpcc_b200 import network, ops, synth, _lib
    fx = fixture_inputs("A")
    network.set_seed(synth.synthetic_seed())
    nets = []
    for _ in range(2):
        network.seed_ptr = 0
        n = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8")
        n.load_state_dict(fx["sd"])          # perturbed state: non-trivial kernels, sigma, mu, beta, gamma
        nets.append(n.cuda().reconstructor)
    ref, fus = nets
    with torch.no_grad():                    # push some IGDN entries below their bounds (LowerBound branches)
        for r in (ref, fus):
            r.activation.gamma[0, 1] = 1e-7
            r.activation.gamma[2, 3] = -0.5
            r.activation.beta[1] = 1e-4
    g = torch.Generator(device="cuda").manual_seed(5)
    nq = sum(p.numel() for p in ref.get_q_params())
    noise = torch.rand(nq, device="cuda", generator=g) if q == 1 else None
    # torch path with the same noise
    w_ref = {}
    off = 0
    for name, key in (("up0", "up0"), ("conv0", "conv0"), ("up1", "up1"), ("conv1", "conv1"), ("up2", "up2"),
                      ("conv2", "conv2"), ("conv2_cls", "cls2")):
        m = getattr(ref, name)
        k = m.kernel
        if q == 1:
            k = k + (noise[off:off + k.numel()].view_as(k) - 0.5) * (1 / 16)
        elif q == 2:
            k = network.bypass_round16(k)
        off += m.kernel.numel()
        w_ref[key + "_w"], w_ref[key + "_b"] = k + m.kernel_init, m.b + m.b_init
    w_ref["cls1_w"], w_ref["cls1_b"] = ref.conv1_cls.effective()
    w_ref["cls0_w"], w_ref["cls0_b"] = ref.conv0_cls.effective()
    w_ref["igdn_beta"], w_ref["igdn_gamma"] = ref.activation.effective()
    bits_ref = ref.net_bits()
    outs = ops._ParamPrepFn.apply(3, (8, 16, 8, 8), q, noise, float(fus.activation.beta_bound),
                                  float(fus.activation.gamma_bound), float(fus.activation.reparam_pedestal),
                                  *[fus.raw_tensors()[k] for k in ops.RAW_FIELDS])
    w_fus, bits_fus = dict(zip(_lib.WEIGHT_FIELDS, outs[:-1])), outs[-1]
    for k in _lib.WEIGHT_FIELDS:
        np.testing.assert_allclose(w_fus[k].detach().cpu().numpy(), w_ref[k].detach().cpu().numpy(), rtol=1e-6, atol=1e-7,
                                   err_msg=k)
    np.testing.assert_allclose(bits_fus.detach().cpu().numpy(), bits_ref.detach().cpu().numpy(), rtol=2e-5)
    # backward with random cotangents
    cot = {k: torch.randn(w_ref[k].shape, device="cuda", generator=g) for k in _lib.WEIGHT_FIELDS}
    cb = torch.rand(7, device="cuda", generator=g) + 0.1
    for w, bits in ((w_ref, bits_ref), (w_fus, bits_fus)):
        (sum((w[k] * cot[k]).sum() for k in _lib.WEIGHT_FIELDS) + (bits * cb).sum()).backward()
    for (kn, pr), (_, pf) in zip(ref.named_parameters(), fus.named_parameters()):
        assert pr.grad is not None and pf.grad is not None, kn
        a, b = pf.grad.cpu().numpy(), pr.grad.cpu().numpy()
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=1e-5 * max(1.0, float(np.abs(b).max())), err_msg=kn)
