"""Entropy coding of the NVF bitstream, in process.

Two wire formats are frozen by the reference and reproduced here byte for byte:

* the latent code - `latent_pack['latent_byte_stream']`, which encode()/decode()
  (NVFPCC.py:446-477, 588-607) obtain by piping int16 symbols and tiled float32 mu / sigma
  through the `./module_arithmeticcoding e|d 1 1` helper.  `arithmetic_encode` /
  `arithmetic_decode` call the same coder inside libnvf_b200.so (csrc/nvf_entropy.cpp): no
  subprocess, no tiled arrays on a pipe, one cumulative table per channel.
* the weight code - `net_weight_pack` of util_code_quantized_weights.py (Huffman code of the
  1/16-quantised kernels + float32 side tensors).  The functions below keep that module's names,
  arguments and return values (`enc_dec_from_file`, `entropy_encode`, `entropy_decode`,
  `get_pdf`, `get_huffman_codebook`, ...), so `import nvfpcc_b200.entropy as entropy_module`
  replaces `import util_code_quantized_weights as entropy_module` (NVFPCC.py:22); the bit walk
  (a Python loop over a '0'/'1' string in the reference, `bitstream` package) runs in C++.

Host code: works without a GPU, but needs the built library (no pure-Python fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _lib

EXPORTS = ("nvf_arith_encode_bound", "nvf_arith_encode_host", "nvf_arith_decode_host", "nvf_huffman_encode_host",
           "nvf_huffman_decode_host")
LATENT_OFFSET = 512        # NVFPCC.py:447 / :593
FL_LEVELS = (1, 1)         # argv of the helper as called by the reference (NVFPCC.py:464,469,600)

qp = 16                    # util_code_quantized_weights.py:9
keys_quantize = [          # util_code_quantized_weights.py:10-18
    'reconstructor.up0.kernel', 'reconstructor.conv0.kernel', 'reconstructor.up1.kernel',
    'reconstructor.conv1.kernel', 'reconstructor.up2.kernel', 'reconstructor.conv2.kernel',
    'reconstructor.conv2_cls.kernel',
]
keys_code_as_is = [        # util_code_quantized_weights.py:20-35
    'entropy_coder.sigma', 'entropy_coder.mu', 'reconstructor.activation.beta', 'reconstructor.activation.gamma',
    'reconstructor.activation.pedestal', 'reconstructor.up0.b', 'reconstructor.conv0.b', 'reconstructor.up1.b',
    'reconstructor.conv1.b', 'reconstructor.up2.b', 'reconstructor.conv2.b', 'reconstructor.conv2_cls.b',
    'reconstructor.likelihood_model.sigma', 'reconstructor.likelihood_model.mu',
]

_bound = None


def _lib_handle():
    global _bound
    if _bound is None:
        b = _lib.cuda_binding()          # raises when libnvf_b200.so is not built
        L = b.lib
        for name in EXPORTS:
            if not hasattr(L, name):
                raise _lib.NvfError("%s does not export %s" % (b.path, name))
        vp, sz = C.c_void_p, C.c_size_t
        L.nvf_arith_encode_bound.argtypes = [C.c_int64, C.POINTER(sz)]
        L.nvf_arith_encode_host.argtypes = [vp, vp, vp, C.c_int64, C.c_int, C.c_int, vp, sz, C.POINTER(sz)]
        L.nvf_arith_decode_host.argtypes = [vp, sz, vp, vp, C.c_int64, C.c_int, C.c_int, vp]
        L.nvf_huffman_encode_host.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int32, vp, sz, C.POINTER(sz)]
        L.nvf_huffman_decode_host.argtypes = [vp, sz, vp, vp, vp, C.c_int32, C.c_int64, vp]
        _bound = b
    return _bound


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


# ----------------------------------------------------------------------------- latent code
def arithmetic_encode(symbols: np.ndarray, mu: np.ndarray, sigma: np.ndarray, levels: Tuple[int, int] = FL_LEVELS) -> bytes:
    """symbols int16 in [0,1024], mu / sigma float32, all flat and of equal length -> byte stream
    (what `sp.run([EXE_ARITH,'e','1','1'], input=length+coeff+mu+sigma).stdout` returns, NVFPCC.py:461-465)."""
    b = _lib_handle()
    sym = np.ascontiguousarray(symbols, dtype=np.int16).reshape(-1)
    mu = np.ascontiguousarray(mu, dtype=np.float32).reshape(-1)
    sg = np.ascontiguousarray(sigma, dtype=np.float32).reshape(-1)
    if not (sym.size == mu.size == sg.size):
        raise ValueError("symbols, mu and sigma must have the same number of elements")
    cap = C.c_size_t(0)
    b.check(b.lib.nvf_arith_encode_bound(sym.size, C.byref(cap)), "nvf_arith_encode_bound")
    out = np.empty(cap.value, dtype=np.uint8)
    n_out = C.c_size_t(0)
    b.check(b.lib.nvf_arith_encode_host(_p(sym), _p(mu), _p(sg), sym.size, levels[0], levels[1], _p(out), out.size,
                                        C.byref(n_out)), "nvf_arith_encode_host")
    return out[: n_out.value].tobytes()


def arithmetic_decode(stream: bytes, mu: np.ndarray, sigma: np.ndarray, levels: Tuple[int, int] = FL_LEVELS) -> np.ndarray:
    """Inverse of arithmetic_encode: int16 symbols, one per (mu, sigma) entry (NVFPCC.py:598-601)."""
    b = _lib_handle()
    mu = np.ascontiguousarray(mu, dtype=np.float32).reshape(-1)
    sg = np.ascontiguousarray(sigma, dtype=np.float32).reshape(-1)
    if mu.size != sg.size:
        raise ValueError("mu and sigma must have the same number of elements")
    buf = np.frombuffer(bytes(stream), dtype=np.uint8)
    out = np.empty(mu.size, dtype=np.int16)
    b.check(b.lib.nvf_arith_decode_host(_p(buf) if buf.size else None, buf.size, _p(mu), _p(sg), mu.size, levels[0],
                                        levels[1], _p(out)), "nvf_arith_decode_host")
    return out


def encode_latents(quantized_latent, sigma, mu) -> Dict:
    """`arithmetic_enc` of encode() (NVFPCC.py:446-483): rounded latents (N,ch,2,2,2) and the
    (1,ch,1,1,1) entropy-model tensors -> `latent_pack` with the reference's keys."""
    import torch

    s = quantized_latent.shape
    np_tensor = quantized_latent.detach().cpu().numpy()
    d = np_tensor.astype(np.int16)
    if not np.sum(np.abs(d - np_tensor)) < 1e-6:
        raise ValueError("latents are not integers")
    flat_coeff = d.reshape(-1) + LATENT_OFFSET
    flat_sigma = np.broadcast_to(sigma.detach().cpu().numpy().astype(np.float32), s).reshape(-1)
    flat_mu = np.broadcast_to(mu.detach().cpu().numpy().astype(np.float32), s).reshape(-1) + np.float32(LATENT_OFFSET)
    stream = arithmetic_encode(flat_coeff, flat_mu, flat_sigma)
    back = arithmetic_decode(stream, flat_mu, flat_sigma)
    if not np.array_equal(back, flat_coeff):                         # the reference's own round-trip assert, :467-474
        raise _lib.NvfError("latent code does not round-trip")
    return {'shape': s, 'latent_byte_stream': stream, 'sigma': sigma, 'mu': mu,
            'length': np.array([flat_coeff.shape[0]], dtype=np.int64)}


def decode_latents(latent_pack: Dict) -> np.ndarray:
    """Step 2 of decode() (NVFPCC.py:584-606): `latent_pack` -> float32 latents of shape `shape`."""
    s = tuple(latent_pack['shape'])
    sigma, mu = latent_pack['sigma'], latent_pack['mu']
    flat_sigma = np.broadcast_to(sigma.detach().cpu().numpy().astype(np.float32), s).reshape(-1)
    flat_mu = np.broadcast_to(mu.detach().cpu().numpy().astype(np.float32), s).reshape(-1) + np.float32(LATENT_OFFSET)
    n = int(np.asarray(latent_pack['length']).reshape(-1)[0])
    if n != flat_mu.size:
        raise ValueError("latent_pack length %d does not match its shape %s" % (n, s))
    sym = arithmetic_decode(latent_pack['latent_byte_stream'], flat_mu, flat_sigma)
    return (sym - LATENT_OFFSET).astype(np.float32).reshape(s)


# ----------------------------------------------------------------------------- weight code
def read_elements_from_state(ws) -> Tuple[np.ndarray, List[np.ndarray], List[np.ndarray]]:
    """util_code_quantized_weights.py:37-52 for an already loaded state dict."""
    pool = [ws[k].detach().cpu().numpy() * qp for k in keys_quantize]
    as_is_pool = [ws[k].detach().cpu().numpy() for k in keys_code_as_is]
    eles = np.concatenate([i.reshape(-1) for i in pool])
    if not np.abs(np.sum(np.round(eles) - eles)) < 1e-3:
        print("Warning: the loaded elements are not discrete!")
        raise ValueError('The loaded elements are not discrete.')
    return eles, pool, as_is_pool


def read_elements_from_file(fn):
    import torch
    return read_elements_from_state(torch.load(fn, map_location=torch.device('cpu')))


def get_pdf(eles):
    """util_code_quantized_weights.py:54-66."""
    lower_bound = int(np.min(eles))
    upper_bound = int(np.max(eles))
    bins = np.linspace(lower_bound, upper_bound + 1, (upper_bound - lower_bound) + 2).astype(int)
    freq = np.histogram(eles, bins)[0].astype(np.float64)
    pdf = freq / np.sum(freq)
    nz = np.where(pdf != 0)
    return pdf[nz], bins[nz]


def get_huffman_codebook(pdf, bins):
    """util_code_quantized_weights.py:74-100: repeatedly merge the two least probable nodes
    (stable sort by probability, merged node appended last), first child = bit 0.
    Returns (codebook: symbol -> bool array, inv_codebook: '0101' -> symbol)."""
    nodes = [(p, None, k) for p, k in zip(pdf, bins)]           # (probability, children, symbol)
    while len(nodes) > 1:
        pool = sorted(nodes, key=lambda x: x[0])
        a, b = pool[0], pool[1]
        nodes = pool[2:] + [(a[0] + b[0], (a, b), None)]
    codebook, inv_codebook = {}, {}
    stack = [(nodes[0], [])]
    while stack:                                                 # pre-order, child 0 before child 1
        node, word = stack.pop()
        if node[1] is None:
            codebook[node[2]] = np.array(word).astype(bool)
            inv_codebook[''.join(str(c) for c in word)] = node[2]
            continue
        stack.append((node[1][1], word + [1]))
        stack.append((node[1][0], word + [0]))
    return codebook, inv_codebook


def est_rate(pdf, bins, codebook):
    """util_code_quantized_weights.py:98-106."""
    return float(sum(len(codebook[s]) * p for p, s in zip(pdf, bins)))


def _code_arrays(words: Sequence[Tuple[int, str]]):
    syms = np.array([int(s) for s, _ in words], dtype=np.int32)
    lens = np.array([len(w) for _, w in words], dtype=np.uint8)
    if lens.size and int(lens.max()) > 64:
        raise ValueError("Huffman codewords longer than 64 bits are not supported")
    bits = np.array([int(w, 2) if w else 0 for _, w in words], dtype=np.uint64)
    return syms, lens, bits


def entropy_encode(tensor_list, codebook):
    """util_code_quantized_weights.py:108-128: concatenated codewords, MSB-first, zero padded."""
    b = _lib_handle()
    shape_list = [t.shape for t in tensor_list]
    flat = np.concatenate([np.asarray(t).reshape(-1) for t in tensor_list]) if tensor_list else np.zeros(0)
    sym = np.rint(flat).astype(np.int32)
    if not np.all(np.abs(sym - flat) < 1e-3):
        raise ValueError("elements are not integers")
    words = [(int(s), ''.join('1' if c else '0' for c in w)) for s, w in codebook.items()]
    cs, cl, cb = _code_arrays(words)
    out = np.empty(int(sym.size) * 8 + 8, dtype=np.uint8)
    n_out = C.c_size_t(0)
    b.check(b.lib.nvf_huffman_encode_host(_p(sym), sym.size, _p(cs), _p(cl), _p(cb), cs.size, _p(out), out.size,
                                          C.byref(n_out)), "nvf_huffman_encode_host")
    return out[: n_out.value].tobytes(), shape_list


def entropy_decode(byte_str, inv_codebook, nsymbol, shape_list):
    """util_code_quantized_weights.py:130-148: -> list of float32 tensors of shapes `shape_list`."""
    b = _lib_handle()
    cs, cl, cb = _code_arrays([(int(s), w) for w, s in inv_codebook.items()])
    buf = np.frombuffer(bytes(byte_str), dtype=np.uint8)
    out = np.empty(int(nsymbol), dtype=np.int32)
    b.check(b.lib.nvf_huffman_decode_host(_p(buf) if buf.size else None, buf.size, _p(cs), _p(cl), _p(cb), cs.size,
                                          int(nsymbol), _p(out)), "nvf_huffman_decode_host")
    tensors, at = [], 0
    for s in shape_list:
        n_sym = int(np.prod(s))
        tensors.append(out[at:at + n_sym].astype(np.float32).reshape(s))
        at += n_sym
    return tensors


def est_fp_bit_consumption(tensor_list):
    """util_code_quantized_weights.py:150-156."""
    total = 0
    for t in tensor_list:
        assert t.dtype == np.float32
        total += int(np.prod(t.shape))
    return total * 32


def enc_dec_from_state(ws, verbose: bool = False) -> Dict:
    """enc_dec_from_file (util_code_quantized_weights.py:186-207) for a loaded state dict."""
    eles, pool, as_is_pool = read_elements_from_state(ws)
    pdf, bins = get_pdf(eles)
    codebook, inv_codebook = get_huffman_codebook(pdf, bins)
    bit_stream, shape_list = entropy_encode(pool, codebook)
    if verbose:
        print('Estimated E(l): ', est_rate(pdf, bins, codebook))
        print('Bit-stream length in bytes: ', len(bit_stream))
        n_bits_as_is = est_fp_bit_consumption(as_is_pool)
        print('Extra bits: ', n_bits_as_is)
        print('Total bits: ', n_bits_as_is + len(bit_stream) * 8)
    dec_pool = entropy_decode(bit_stream, inv_codebook, len(eles), shape_list)
    for a, b in zip(pool, dec_pool):
        assert np.sum(np.abs(a - b)) < 1e-6
    return {'bit_stream': bit_stream, 'inv_codebook': inv_codebook, 'element_length': len(eles),
            'shape_list': shape_list, 'as_is_pool': as_is_pool, 'keys_quantize': keys_quantize,
            'keys_code_as_is': keys_code_as_is}


def enc_dec_from_file(filename):
    import torch
    return enc_dec_from_state(torch.load(filename, map_location=torch.device('cpu')), verbose=True)


def state_from_weight_pack(net_weight_pack: Dict, qp_: float = qp):
    """Step 1 of decode() (NVFPCC.py:571-583): weight pack -> partial state dict (float32 tensors)."""
    import torch

    dec_pool = entropy_decode(net_weight_pack['bit_stream'], net_weight_pack['inv_codebook'],
                              net_weight_pack['element_length'], net_weight_pack['shape_list'])
    nd = {}
    for k, v in zip(net_weight_pack['keys_quantize'], dec_pool):
        nd[k] = torch.from_numpy(v).float() / qp_
    for k, v in zip(net_weight_pack['keys_code_as_is'], net_weight_pack['as_is_pool']):
        nd[k] = torch.from_numpy(np.asarray(v)).float()
    return nd
