"""Batched encode / decode drivers around the NVF decoder path.

Drop-in for the bodies of encode() and decode() of the reference (NVFPCC.py:395-554, 557-652)
with the frozen artefacts kept as they are: the `pack.pk` pickle
    {'net_weight_pack': ..., 'origins': int16 (N,3), 'latent_pack': ...}          (NVFPCC.py:486-493)
and the `rc_enc.ply` / `rc_dec.ply` ASCII clouds, which must be identical (README.md:63).

What changes is how the work is done:
* every leaf block goes through ONE batched `nvf_decode` call per chunk (fused decoder +
  threshold + ordered coordinate emission on the GPU) instead of a per-block Python loop with
  MinkowskiEngine dense->sparse conversions and two host syncs per block (NVFPCC.py:505-539,
  625-638);
* each block is paired with ITS origin (the reference indexes `origins[i]` with the batch
  index, NVFPCC.py:534-536, which is only right for --batchsize 1);
* the number of blocks comes from the pack (`origins.shape[0]`), not from `--N 917`
  (NVFPCC.py:613,741);
* the latent / weight bit streams are produced in process (nvfpcc_b200.entropy) - same bytes;
* with torch.distributed initialised, blocks are split into contiguous ranges per rank and the
  coordinates gathered in rank order (nvfpcc_b200.dist), which reproduces the 1-GPU point order.

There is no CPU fallback: the reconstruction needs a CUDA device and the built library.
"""
from __future__ import annotations

import pickle
from typing import Dict, Optional

import numpy as np
import torch

from . import dist as nvdist
from . import entropy, network

DEFAULT_CHUNK = 4096     # leaf blocks per nvf_decode call (workspace ~ 4 KB mask + scratch per block)


def quantize_state(d: Dict[str, torch.Tensor], iqp: int = 16) -> Dict[str, torch.Tensor]:
    """manipulate_weights.py:19-50: round the 7 decoder kernels to multiples of 1/iqp, keep the listed
    side tensors, drop everything else (aux heads)."""
    bypass = ['latent_gen.h_analysis_2.kernel', 'latent_gen.h_analysis_2.b', 'latent_gen.h_analysis_2.kernel_init',
              'latent_gen.h_analysis_2.b_init', 'latent_gen.gdn_2.beta', 'latent_gen.gdn_2.gamma',
              'latent_gen.gdn_2.pedestal']
    handled = set(entropy.keys_quantize)
    keys = ['entropy_coder.sigma', 'entropy_coder.mu', 'reconstructor.activation.beta',
            'reconstructor.activation.gamma', 'reconstructor.activation.pedestal']
    for layer in ('up0', 'conv0', 'up1', 'conv1', 'up2', 'conv2', 'conv2_cls'):
        keys += ['reconstructor.%s.kernel' % layer, 'reconstructor.%s.b' % layer]
    keys += ['reconstructor.likelihood_model.sigma', 'reconstructor.likelihood_model.mu']
    nd = {}
    with torch.no_grad():
        for k in bypass:
            nd[k] = d[k].clone()
        for k in keys:
            nd[k] = torch.round(d[k] * iqp) / iqp if k in handled else d[k].clone()
    return nd


def _reconstruct(net: network.Net, latents: torch.Tensor, origins: np.ndarray, thh: float, chunk: int,
                 gather: bool = True) -> Optional[np.ndarray]:
    """All leaf blocks -> (K,3) int32 points in block order then row-major (i,j,k) order.  Under
    torch.distributed each rank decodes its contiguous block range; rank 0 returns the whole cloud."""
    n = int(latents.shape[0])
    rank, ws = nvdist.world()
    lo, hi = nvdist.block_range(n, rank, ws)
    dev = next(net.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("nvfpcc_b200.codec needs the network on a CUDA device (no CPU fallback)")
    org = torch.as_tensor(np.asarray(origins).astype(np.int32)).to(dev)
    lat = latents.to(dev, torch.float32)
    coords, counts = [], []
    for s in range(lo, hi, chunk):
        e = min(hi, s + chunk)
        r = net.decode_points(lat[s:e], org[s:e], thh, q=2, return_host=False)
        coords.append(r["coords"])
        counts.append(r["counts"])
    c = torch.cat(coords, 0) if coords else torch.zeros((0, 3), dtype=torch.int32, device=dev)
    k = torch.cat(counts, 0) if counts else torch.zeros((0,), dtype=torch.int32, device=dev)
    if ws > 1 and gather:
        c, k = nvdist.gather_points(c, k, dst=0)
        if c is None:
            return None
    return c.cpu().numpy()


def encode(net: network.Net, emb: torch.Tensor, origins, thh: float, weights_state: Optional[Dict] = None,
           pack_fn: Optional[str] = None, ply_fn: Optional[str] = None, chunk: int = DEFAULT_CHUNK) -> Dict:
    """encode() of the reference (NVFPCC.py:395-554) for an already loaded `net` and `emb`.

    weights_state: the (1/16-quantised) checkpoint to put into the pack; defaults to net.state_dict().
    Returns {'total_pack', 'points' (encoder-side reconstruction, rank 0), 'latent_bits', 'net_bits'}.
    """
    origins = np.asarray(origins)
    if origins.shape[0] != emb.shape[0]:
        raise ValueError("one embedding per leaf origin is required (%d vs %d)" % (emb.shape[0], origins.shape[0]))
    sd = weights_state if weights_state is not None else net.state_dict()
    net_weight_pack = entropy.enc_dec_from_state(sd)                                   # 1. network parameters
    np_origins = np.array(origins, dtype=np.int16)                                     # 2. cube origins
    if not np.array_equal(np_origins, origins):
        raise ValueError("leaf origins do not fit int16 (pack layout, NVFPCC.py:439)")
    dev = next(net.parameters()).device
    with torch.no_grad():                                                              # 3. embeddings
        info = net.get_latent_code(emb.to(dev))
    latent_pack = entropy.encode_latents(info['quantized_latent'], info['sigma'].detach().cpu(),
                                         info['mu'].detach().cpu())
    total_pack = {'net_weight_pack': net_weight_pack, 'origins': np_origins, 'latent_pack': latent_pack}
    rank, _ = nvdist.world()
    if pack_fn and rank == 0:
        with open(pack_fn, 'wb') as f:
            pickle.dump(total_pack, f)
    pts = _reconstruct(net, info['quantized_latent'], np_origins, thh, chunk)
    if ply_fn and pts is not None:
        write_ply_ascii(ply_fn, pts)
    return {'total_pack': total_pack, 'points': pts,
            'latent_bits': len(latent_pack['latent_byte_stream']) * 8,
            'net_bits': len(net_weight_pack['bit_stream']) * 8}


def decode(total_pack, ch: int, channel_str: str, thh: float, qp: float = 16, device=None,
           ply_fn: Optional[str] = None, chunk: int = DEFAULT_CHUNK, net: Optional[network.Net] = None):
    """decode() of the reference (NVFPCC.py:557-652): pack (dict or file name) -> (K,3) int32 points on
    rank 0.  `network.set_seed(...)` must hold the same seed table as at encode time (SEED3.npy)."""
    if isinstance(total_pack, (str, bytes)):
        with open(total_pack, 'rb') as f:
            total_pack = pickle.load(f)
    if not torch.cuda.is_available():
        raise RuntimeError("nvfpcc_b200.codec needs a CUDA device (no CPU fallback)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if net is None:
        network.seed_ptr = 0
        net = network.Net(None, "Gaussian", ch=ch, channel_str=channel_str)
    nd = entropy.state_from_weight_pack(total_pack['net_weight_pack'], qp)             # 1. network weights
    net.load_state_dict(nd, strict=False)
    net = net.to(dev)
    latents = torch.from_numpy(entropy.decode_latents(total_pack['latent_pack']))      # 2. latent codes
    np_origins = total_pack['origins']                                                 # 3. cube origins
    if latents.shape[0] != np_origins.shape[0]:
        raise ValueError("pack holds %d latents for %d origins" % (latents.shape[0], np_origins.shape[0]))
    pts = _reconstruct(net, latents, np_origins, thh, chunk)
    if ply_fn and pts is not None:
        write_ply_ascii(ply_fn, pts)
    return pts


def write_ply_ascii(fn: str, points: np.ndarray) -> None:
    """ASCII PLY with double x,y,z vertices - the layout o3d.io.write_point_cloud(..., write_ascii=True)
    gives rc_enc.ply / rc_dec.ply (NVFPCC.py:548-554, 642-650; coordinates are rounded integers)."""
    p = np.round(np.asarray(points)).astype(np.int64)
    with open(fn, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment Created by nvfpcc_b200\nelement vertex %d\n"
                "property double x\nproperty double y\nproperty double z\nend_header\n" % p.shape[0])
        np.savetxt(f, p, fmt="%d")
