"""GPU tests of the batched encode / decode drivers (nvfpcc_b200.codec): pack layout, rc_enc == rc_dec
bit-exactly (README.md:63), agreement with the oracle, and that the reference's own helper decodes the
latent stream of a pack written here."""
import os
import pickle
import subprocess as sp

import numpy as np
import pytest
import torch

from oracle import nvf_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "module_arithmeticcoding")
THH = 0.64


def make_case(n_blocks=300, chanstr="8,16,8,8"):
    from nvfpcc_b200 import codec, network, synth
    network.set_seed(synth.synthetic_seed())
    net = network.Net(None, "Gaussian", ch=3, channel_str=chanstr)
    sd = net.state_dict()
    sd.update(synth.random_kernel_deltas(sd, seed=1, sigma=0.05, quantize=True))
    g = torch.Generator().manual_seed(5)
    sd["entropy_coder.sigma"] = torch.tensor([2.5, 3.0, 1.5]).reshape(1, 3, 1, 1, 1)
    sd["entropy_coder.mu"] = torch.tensor([0.3, -0.2, 0.0]).reshape(1, 3, 1, 1, 1)
    net.load_state_dict(sd)
    net = net.cuda()
    pts = synth.sphere_shell_points(1024)
    origins = synth.leaf_origins(pts)[:n_blocks]
    emb = torch.randn(n_blocks, 3, 2, 2, 2, generator=g) * 4
    with torch.no_grad():                       # calibrate to a realistic occupancy
        lat = net.get_latent_code(emb[:32].cuda())["quantized_latent"]
        p = net.reconstruct(lat, 2)
        logit = torch.log(p) - torch.log1p(-p)
        net.reconstructor.conv2_cls.b += float(np.log(THH / (1 - THH))) - torch.quantile(logit.flatten()[::7], 0.979)
    state = codec.quantize_state({k: v.detach().cpu() for k, v in net.state_dict().items()}, 16)
    return net, emb, origins, state


def test_encode_decode_round_trip_bit_exact(gpu, tmp_path):
    from nvfpcc_b200 import codec, network, synth
    net, emb, origins, state = make_case()
    pack_fn, enc_ply, dec_ply = (str(tmp_path / n) for n in ("pack.pk", "rc_enc.ply", "rc_dec.ply"))
    enc = codec.encode(net, emb, origins, THH, weights_state=state, pack_fn=pack_fn, ply_fn=enc_ply, chunk=128)
    tp = enc["total_pack"]
    # pack layout (NVFPCC.py:486-493, 476-482; util_code_quantized_weights.py:199-207)
    assert list(tp) == ["net_weight_pack", "origins", "latent_pack"]
    assert tp["origins"].dtype == np.int16 and tp["origins"].shape == (300, 3)
    assert set(tp["latent_pack"]) == {"shape", "latent_byte_stream", "sigma", "mu", "length"}
    assert enc["latent_bits"] == 8 * len(tp["latent_pack"]["latent_byte_stream"]) > 0
    assert 0.005 < enc["points"].shape[0] / (300 * 32768) < 0.06
    # decoder side: fresh network from the seed table + the pack file only, different chunking
    network.set_seed(synth.synthetic_seed())
    dec = codec.decode(pack_fn, 3, "8,16,8,8", THH, ply_fn=dec_ply, chunk=77)
    assert dec.dtype == np.int32 and np.array_equal(dec, enc["points"])
    assert open(enc_ply, "rb").read() == open(dec_ply, "rb").read()
    # the latents the decoder saw are the encoder's rounded latents
    from nvfpcc_b200 import entropy
    with torch.no_grad():
        q = net.get_latent_code(emb.cuda())["quantized_latent"].cpu().numpy()
    assert np.array_equal(entropy.decode_latents(tp["latent_pack"]), q)


def test_decoded_cloud_matches_oracle(gpu):
    from nvfpcc_b200 import codec, entropy, network, synth
    net, emb, origins, state = make_case(n_blocks=24)
    enc = codec.encode(net, emb, origins, THH, weights_state=state)
    network.set_seed(synth.synthetic_seed())
    pts = codec.decode(enc["total_pack"], 3, "8,16,8,8", THH)
    # oracle: decoded state on top of the constructor state, decoded latents, dense threshold
    network.set_seed(synth.synthetic_seed())
    sd = {k: v.clone() for k, v in network.Net(None, "Gaussian", ch=3, channel_str="8,16,8,8").state_dict().items()}
    sd.update(entropy.state_from_weight_pack(enc["total_pack"]["net_weight_pack"]))
    lat = torch.from_numpy(entropy.decode_latents(enc["total_pack"]["latent_pack"]))
    with torch.no_grad():
        ref = O.reconstruct(lat, sd, q=2)
    pts_ref, _ = O.threshold_points(ref, enc["total_pack"]["origins"].astype(np.int32), THH)
    key = lambda a: set(map(tuple, a.tolist()))
    diff = key(pts) ^ key(pts_ref)
    assert len(diff) <= 1e-4 * 24 * 32768
    # every disagreeing voxel sits within 1e-4 of the threshold
    org = enc["total_pack"]["origins"].astype(np.int64)
    for v in diff:
        b = int(np.nonzero((org == (np.array(v) // 32 * 32)).all(1))[0][0])
        i, j, k = (np.array(v) - org[b]).tolist()
        assert abs(float(ref[b, 0, i, j, k]) - THH) <= 1e-4


@pytest.mark.skipif(not os.path.isfile(REF_EXE), reason="oracle/_ref not built")
def test_reference_helper_decodes_our_latent_stream(gpu):
    from nvfpcc_b200 import codec
    net, emb, origins, state = make_case(n_blocks=64)
    lp = codec.encode(net, emb, origins, THH, weights_state=state)["total_pack"]["latent_pack"]
    s = tuple(lp["shape"])
    flat_sigma = np.broadcast_to(lp["sigma"].numpy().astype(np.float32), s).reshape(-1)
    flat_mu = np.broadcast_to(lp["mu"].numpy().astype(np.float32), s).reshape(-1) + np.float32(512)
    cmd = lp["length"].tobytes() + flat_mu.tobytes() + flat_sigma.tobytes() + lp["latent_byte_stream"]     # NVFPCC.py:598
    out = sp.run([REF_EXE, "d", "1", "1"], input=cmd, stdout=sp.PIPE, check=True).stdout
    with torch.no_grad():
        q = net.get_latent_code(emb.cuda())["quantized_latent"].cpu().numpy()
    assert np.array_equal(np.frombuffer(out, dtype=np.int16).astype(np.float32).reshape(s) - 512, q)


def test_pack_is_plain_pickle_and_errors(gpu, tmp_path):
    from nvfpcc_b200 import codec
    net, emb, origins, state = make_case(n_blocks=8)
    enc = codec.encode(net, emb, origins, THH, weights_state=state, pack_fn=str(tmp_path / "p.pk"))
    tp = pickle.load(open(tmp_path / "p.pk", "rb"))
    assert tp["latent_pack"]["latent_byte_stream"] == enc["total_pack"]["latent_pack"]["latent_byte_stream"]
    with pytest.raises(ValueError):
        codec.encode(net, emb[:4], origins, THH, weights_state=state)
    bad = dict(state)
    bad["reconstructor.conv2.kernel"] = state["reconstructor.conv2.kernel"] + 0.013
    with pytest.raises(ValueError):
        codec.encode(net, emb, origins, THH, weights_state=bad)


def test_zero_leaves_round_trip(gpu):
    from nvfpcc_b200 import codec, network, synth
    net, emb, origins, state = make_case(n_blocks=8)
    enc = codec.encode(net, emb[:0], origins[:0], THH, weights_state=state)
    assert enc["points"].shape == (0, 3) and enc["total_pack"]["origins"].shape == (0, 3)
    network.set_seed(synth.synthetic_seed())
    dec = codec.decode(enc["total_pack"], 3, "8,16,8,8", THH)
    assert dec.shape == (0, 3)
