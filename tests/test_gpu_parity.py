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
