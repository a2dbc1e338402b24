"""In-process entropy coders (csrc/nvf_entropy.cpp via nvfpcc_b200.entropy) against known answers
produced by the reference's own coders (oracle/gen_golden_entropy.py) and, when the compiled
reference helper is present (oracle/_ref, build container and GPU box), against it directly.
Host code only: no GPU needed."""
import os
import subprocess as sp

import numpy as np
import pytest
import torch

from nvfpcc_b200 import _lib, entropy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "module_arithmeticcoding")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "entropy_small.npz"))


def test_arithmetic_streams_equal_the_reference_helper(golden):
    for name in golden["arith_cases"]:
        sym, mu, sg = (golden["arith_%s_%s" % (name, k)] for k in ("sym", "mu", "sigma"))
        ref = golden["arith_%s_stream" % name].tobytes()
        got = entropy.arithmetic_encode(sym, mu, sg)
        assert got == ref, name
        assert np.array_equal(entropy.arithmetic_decode(ref, mu, sg), sym), name


@pytest.mark.skipif(not os.path.isfile(REF_EXE), reason="oracle/_ref not built")
def test_arithmetic_coder_against_compiled_reference_random():
    rng = np.random.default_rng(123)
    for trial in range(6):
        ch = int(rng.integers(1, 5))
        nb = int(rng.integers(1, 400))
        mus = rng.normal(0, 3, ch).astype(np.float32)
        sig = (np.abs(rng.normal(1.5, 2, ch)) + 0.02).astype(np.float32)
        lat = np.clip(np.rint(rng.normal(mus[None, :, None], 1.5 * sig[None, :, None], size=(nb, ch, 8))), -512, 511)
        mu = (np.broadcast_to(mus[None, :, None], lat.shape).reshape(-1) + np.float32(512)).astype(np.float32)
        sg = np.broadcast_to(sig[None, :, None], lat.shape).reshape(-1).astype(np.float32)
        sym = (lat.reshape(-1) + 512).astype(np.int16)
        length = np.array([sym.size], dtype=np.int64)
        ref = sp.run([REF_EXE, "e", "1", "1"], input=length.tobytes() + sym.tobytes() + mu.tobytes() + sg.tobytes(),
                     stdout=sp.PIPE, check=True).stdout
        assert entropy.arithmetic_encode(sym, mu, sg) == ref
        back = sp.run([REF_EXE, "d", "1", "1"], input=length.tobytes() + mu.tobytes() + sg.tobytes() + ref,
                      stdout=sp.PIPE, check=True).stdout
        assert np.array_equal(np.frombuffer(back, dtype=np.int16), entropy.arithmetic_decode(ref, mu, sg))


def test_arithmetic_errors():
    mu, sg = np.full(4, 512, np.float32), np.full(4, 2, np.float32)
    with pytest.raises(_lib.NvfError):
        entropy.arithmetic_encode(np.array([0, 1, 1025, 3], np.int16), mu, sg)      # symbol out of range
    with pytest.raises(ValueError):
        entropy.arithmetic_encode(np.zeros(3, np.int16), mu, sg)
    # a stream that these models cannot have produced is reported, not asserted on
    # (code value 0 lies below the first symbol's interval when the model has mass below symbol 0)
    with pytest.raises(_lib.NvfError, match="corrupt"):
        entropy.arithmetic_decode(bytes(16), np.full(4, -3.0, np.float32), np.full(4, 2.0, np.float32))


def test_latent_pack_round_trip_and_layout():
    g = torch.Generator().manual_seed(0)
    lat = torch.round(torch.randn(37, 3, 2, 2, 2, generator=g) * 3)
    sigma = torch.tensor([1.5, 0.7, 3.0]).reshape(1, 3, 1, 1, 1)
    mu = torch.tensor([0.2, -1.0, 0.0]).reshape(1, 3, 1, 1, 1)
    pack = entropy.encode_latents(lat, sigma, mu)
    assert set(pack) == {"shape", "latent_byte_stream", "sigma", "mu", "length"}       # NVFPCC.py:476-482
    assert isinstance(pack["latent_byte_stream"], bytes) and pack["length"].dtype == np.int64
    assert tuple(pack["shape"]) == (37, 3, 2, 2, 2) and int(pack["length"][0]) == 37 * 24
    back = entropy.decode_latents(pack)
    assert back.dtype == np.float32 and np.array_equal(back, lat.numpy())
    with pytest.raises(ValueError):
        entropy.encode_latents(lat + 0.25, sigma, mu)


def test_huffman_codebook_and_stream_equal_the_reference_module(golden):
    from oracle.gen_golden_entropy import quantised_kernels
    pool = quantised_kernels()
    eles = np.concatenate([p.reshape(-1) for p in pool])
    pdf, bins = entropy.get_pdf(eles)
    assert np.array_equal(pdf, golden["huff_pdf"]) and np.array_equal(bins, golden["huff_bins"])
    codebook, inv = entropy.get_huffman_codebook(pdf, bins)
    assert list(inv.keys()) == [str(w) for w in golden["huff_words"]]                 # same words, same order
    assert [int(v) for v in inv.values()] == golden["huff_symbols"].tolist()
    assert abs(entropy.est_rate(pdf, bins, codebook) - float(golden["huff_rate"])) < 1e-12
    stream, shapes = entropy.entropy_encode(pool, codebook)
    assert stream == golden["huff_stream"].tobytes()
    dec = entropy.entropy_decode(stream, inv, len(eles), shapes)
    assert all(d.dtype == np.float32 and np.array_equal(d, p) for d, p in zip(dec, pool))
    with pytest.raises(_lib.NvfError, match="corrupt"):
        entropy.entropy_decode(stream[: len(stream) // 2], inv, len(eles), shapes)


def test_weight_pack_layout_and_state_round_trip():
    from nvfpcc_b200 import network, synth
    network.set_seed(synth.synthetic_seed())
    net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8")
    sd = synth.random_kernel_deltas({k: v.clone() for k, v in net.state_dict().items()}, quantize=True)
    pack = entropy.enc_dec_from_state(sd)
    assert list(pack) == ["bit_stream", "inv_codebook", "element_length", "shape_list", "as_is_pool",
                          "keys_quantize", "keys_code_as_is"]                          # util_code_quantized_weights.py:199-207
    assert pack["element_length"] == sum(int(np.prod(s)) for s in pack["shape_list"])
    nd = entropy.state_from_weight_pack(pack)
    for k in entropy.keys_quantize + entropy.keys_code_as_is:
        assert torch.equal(nd[k], sd[k].float()), k
    sd_bad = dict(sd)
    sd_bad["reconstructor.up0.kernel"] = sd["reconstructor.up0.kernel"] + 0.01
    with pytest.raises(ValueError):
        entropy.enc_dec_from_state(sd_bad)


def test_single_symbol_alphabet():
    pool = [np.zeros((2, 3), np.float32)]
    pdf, bins = entropy.get_pdf(pool[0].reshape(-1))
    codebook, inv = entropy.get_huffman_codebook(pdf, bins)
    stream, shapes = entropy.entropy_encode(pool, codebook)
    assert stream == b"" and inv == {"": 0}
    assert np.array_equal(entropy.entropy_decode(stream, inv, 6, shapes)[0], pool[0])


REF_ROOT = os.environ.get("NVF_REFERENCE_ROOT", "/root/reference")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF_ROOT, "util_code_quantized_weights.py")),
                    reason="reference checkout absent")
def test_reference_module_decodes_our_weight_pack():
    """The pack written here is readable by the reference's own entropy_decode (util_code_quantized_weights.py:130-148)."""
    import sys
    from oracle.gen_golden_entropy import bitstream_stub
    from nvfpcc_b200 import network, synth
    sys.modules.setdefault("bitstream", bitstream_stub())
    sys.path.insert(0, REF_ROOT)
    try:
        import util_code_quantized_weights as U
    finally:
        sys.path.remove(REF_ROOT)
    network.set_seed(synth.synthetic_seed())
    net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8")
    sd = synth.random_kernel_deltas({k: v.clone() for k, v in net.state_dict().items()}, quantize=True)
    pack = entropy.enc_dec_from_state(sd)
    dec = U.entropy_decode(pack["bit_stream"], pack["inv_codebook"], pack["element_length"], pack["shape_list"])
    for k, v in zip(pack["keys_quantize"], dec):
        assert np.array_equal(v / 16, sd[k].numpy()), k
    assert U.keys_quantize == entropy.keys_quantize and U.keys_code_as_is == entropy.keys_code_as_is and U.qp == entropy.qp


def test_quantize_state_and_ply_writer(tmp_path):
    from nvfpcc_b200 import codec, grids, network, synth
    network.set_seed(synth.synthetic_seed())
    net = network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8")
    sd = synth.random_kernel_deltas({k: v.clone() for k, v in net.state_dict().items()}, quantize=False)
    q = codec.quantize_state(sd, 16)
    assert not any(("conv1_cls" in k or "conv0_cls" in k) for k in q)                 # manipulate_weights.py drops the aux heads
    for k in entropy.keys_quantize:
        assert torch.equal(q[k], torch.round(sd[k] * 16) / 16)
    assert torch.equal(q["reconstructor.conv2.b"], sd["reconstructor.conv2.b"])
    entropy.enc_dec_from_state(q)                                                       # discrete -> codable
    pts = np.array([[1, 2, 3], [1023, 0, 77]], dtype=np.int32)
    codec.write_ply_ascii(str(tmp_path / "a.ply"), pts)
    assert np.array_equal(grids.read_ply_xyz(str(tmp_path / "a.ply")), pts.astype(np.float64))


# ----------------------------------------------------------------------------- property tests
from hypothesis import given, settings, strategies as st


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(0, 300), st.floats(0.003, 40.0), st.floats(-30.0, 30.0))
def test_arithmetic_round_trip_property(seed, n, sigma_scale, mu_shift):
    """Any symbol sequence in range round-trips under any per-symbol Gaussian parameters; when the compiled
    reference helper is present the stream is also byte-identical to its output."""
    rng = np.random.default_rng(seed)
    mu = (512 + mu_shift + rng.normal(0, 5, n)).astype(np.float32)
    sg = (np.abs(rng.normal(0, sigma_scale, n)) + 1e-3).astype(np.float32)
    sym = np.clip(np.rint(rng.normal(mu, np.minimum(sg * 2, 200))), 0, 1023).astype(np.int16)
    stream = entropy.arithmetic_encode(sym, mu, sg)
    assert np.array_equal(entropy.arithmetic_decode(stream, mu, sg), sym)
    if os.path.isfile(REF_EXE):
        length = np.array([n], dtype=np.int64)
        ref = sp.run([REF_EXE, "e", "1", "1"], input=length.tobytes() + sym.tobytes() + mu.tobytes() + sg.tobytes(),
                     stdout=sp.PIPE).stdout
        assert stream == ref


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 400), st.floats(0.2, 12.0))
def test_huffman_round_trip_property(seed, n, spread):
    rng = np.random.default_rng(seed)
    pool = [np.rint(rng.laplace(0, spread, size=(n,))).astype(np.float32), np.rint(rng.normal(0, spread, size=(3, 5))).astype(np.float32)]
    eles = np.concatenate([p.reshape(-1) for p in pool])
    pdf, bins = entropy.get_pdf(eles)
    codebook, inv = entropy.get_huffman_codebook(pdf, bins)
    assert len(set(inv.keys())) == len(inv) == len(pdf)
    words = sorted(inv.keys())
    assert not any(b.startswith(a) for a, b in zip(words, words[1:]))          # prefix free
    stream, shapes = entropy.entropy_encode(pool, codebook)
    assert len(stream) == (sum(len(codebook[int(v)]) for v in eles) + 7) // 8
    dec = entropy.entropy_decode(stream, inv, len(eles), shapes)
    assert all(np.array_equal(a, b) for a, b in zip(pool, dec))
