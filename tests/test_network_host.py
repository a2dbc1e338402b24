"""Host-side mirror of the reference interface (nvfpcc_b200.network): state_dict
compatibility and the torch-side parameter transforms, checked on CPU against the
oracle (which is pinned to the reference by the golden vectors)."""
import numpy as np
import pytest
import torch

from nvfpcc_b200 import network as N
from nvfpcc_b200 import synth
from oracle import nvf_oracle as O
from oracle.gen_golden import fixture_inputs


def build(ch, chanstr):
    N.set_seed(synth.synthetic_seed())
    return N.Net(None, "Gaussian", ch=ch, channel_str=chanstr)


@pytest.mark.parametrize("chanstr,n_seed", [("8,16,8,8", 52127), ("16,32,16,16", 200623)])
def test_constructor_state_equals_reference_layout(chanstr, n_seed):
    net = build(3, chanstr)
    assert N.seed_ptr == n_seed
    ref = O.make_state(3, [int(c) for c in chanstr.split(",")], synth.synthetic_seed())
    sd = net.state_dict()
    ref = {k: v for k, v in ref.items() if not k.startswith("_")}
    assert list(sd.keys()) == list(ref.keys())          # same keys, same registration order
    for k in sd:
        assert sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]), k
    assert sum(p.numel() for p in net.parameters()) == (52219 if chanstr == "8,16,8,8" else 200915)


def test_loads_reference_and_quantised_checkpoints():
    fx = fixture_inputs("A")
    net = build(3, "8,16,8,8")
    net.load_state_dict(fx["sd"], strict=True)
    q = O.quantize_state_16(fx["sd"])                    # manipulate_weights.py output: fewer keys
    missing = net.load_state_dict(q, strict=False)
    assert not missing.unexpected_keys
    assert all(("_init" in k) or ("_cls" in k and "conv2" not in k) for k in missing.missing_keys)


def test_effective_weights_and_rates_match_oracle():
    fx = fixture_inputs("A")
    net = build(3, "8,16,8,8")
    net.load_state_dict(fx["sd"])
    wo = O.effective_decoder_weights(fx["sd"], 2)
    w = net.reconstructor.effective_weights(2)
    ren = {"cls2": "conv2_cls", "cls1": "conv1_cls", "cls0": "conv0_cls"}
    for k, v in w.items():
        base, suf = k.rsplit("_", 1)
        ko = ren.get(base, base) + "_" + suf
        assert torch.equal(v, wo[ko]), k
    assert torch.allclose(net.reconstructor.net_bits(), O.net_bits(fx["sd"]))
    torch.manual_seed(3)
    lat = net.latent_gen(fx["emb"])
    assert torch.equal(lat, O.latent_gen(fx["emb"], fx["sd"]))
    torch.manual_seed(5)
    r, bits = net.entropy_coder(lat, "train")
    torch.manual_seed(5)
    noise = torch.rand_like(lat)
    r2, bits2 = O.entropy_coder(lat, fx["sd"], "train", noise)
    assert torch.equal(r, r2) and torch.allclose(bits, bits2)
    code = net.get_latent_code(fx["emb"])
    assert torch.equal(code["quantized_latent"], torch.round(lat))


def test_q1_noise_and_ste_gradients():
    net = build(3, "8,16,8,8")
    torch.manual_seed(0)
    w1, _ = net.reconstructor.up2.effective(1)
    d = (w1 - net.reconstructor.up2.kernel_init).detach()
    assert d.abs().max() <= 1 / 32 + 1e-7 and d.abs().max() > 1 / 64
    k = net.reconstructor.conv2.kernel
    w2, _ = net.reconstructor.conv2.effective(2)
    w2.sum().backward()
    assert torch.equal(k.grad, torch.ones_like(k))       # straight-through (utils/network.py:43-50)
