"""Generate tests/golden/entropy_small.npz from the reference's own entropy coders (run in the
build container only: needs /root/reference and `make -C oracle`).

    python oracle/gen_golden_entropy.py

TEST INFRASTRUCTURE ONLY.
* Latent code: the reference helper compiled unmodified into oracle/_ref/module_arithmeticcoding
  (oracle/Makefile) encodes seeded symbol sequences exactly as encode() drives it
  (NVFPCC.py:446-465); the byte streams are the known answers for csrc/nvf_entropy.cpp.
* Weight code: the reference module util_code_quantized_weights.py is imported unmodified with a
  stub `bitstream` package (absent here; MSB-first packing = numpy.packbits) and
  produces the Huffman codebook + stream of a seeded set of 1/16-quantised kernels.
"""
from __future__ import annotations

import os
import subprocess as sp
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF_ROOT = os.environ.get("NVF_REFERENCE_ROOT", "/root/reference")
EXE = os.path.join(HERE, "_ref", "module_arithmeticcoding")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def arith_cases():
    """(name, symbols int16, mu f32, sigma f32): tiled per-channel models like encode() builds them,
    plus edge cases (empty, extreme symbols, tiny sigma, every element its own model)."""
    rng = np.random.default_rng(5)
    cases = []
    for name, nb, ch in (("ch3", 200, 3), ("ch4", 64, 4)):
        mus = rng.normal(0, 2, ch).astype(np.float32)
        sig = (np.abs(rng.normal(2, 1.5, ch)) + 0.05).astype(np.float32)
        lat = np.clip(np.rint(rng.normal(mus[None, :, None], sig[None, :, None], size=(nb, ch, 8))), -512, 511)
        mu = np.broadcast_to(mus[None, :, None], lat.shape).reshape(-1).astype(np.float32) + np.float32(512)
        sg = np.broadcast_to(sig[None, :, None], lat.shape).reshape(-1).astype(np.float32)
        cases.append((name, (lat.reshape(-1) + 512).astype(np.int16), mu, sg))
    cases.append(("empty", np.zeros(0, np.int16), np.zeros(0, np.float32), np.zeros(0, np.float32)))
    ext = np.array([0, 1023, 512, 1, 1022, 512, 300, 700], dtype=np.int16)
    cases.append(("extremes", ext, np.full(8, 512.25, np.float32), np.full(8, 3.0, np.float32)))
    cases.append(("tiny_sigma", np.full(40, 515, np.int16), np.full(40, 515.1, np.float32), np.full(40, 1e-3, np.float32)))
    n = 500
    mu = (512 + rng.normal(0, 4, n)).astype(np.float32)
    sg = (np.abs(rng.normal(1, 2, n)) + 0.01).astype(np.float32)
    sym = np.clip(np.rint(rng.normal(mu, sg)), 0, 1023).astype(np.int16)
    cases.append(("per_element", sym, mu, sg))
    return cases


def ref_arith(sym, mu, sg):
    length = np.array([sym.size], dtype=np.int64)
    r = sp.run([EXE, "e", "1", "1"], input=length.tobytes() + sym.tobytes() + mu.tobytes() + sg.tobytes(),
               stdout=sp.PIPE, check=True)
    d = sp.run([EXE, "d", "1", "1"], input=length.tobytes() + mu.tobytes() + sg.tobytes() + r.stdout,
               stdout=sp.PIPE, check=True)
    assert np.array_equal(np.frombuffer(d.stdout, dtype=np.int16), sym)
    return r.stdout


def bitstream_stub():
    class BitStream:
        def __init__(self, data=None):
            self.bits = np.zeros(0, dtype=bool)
            if data is not None:
                self.bits = np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8)).astype(bool)

        def write(self, arr, _type):
            self.bits = np.concatenate([self.bits, np.asarray(arr, dtype=bool)])

        def __len__(self):
            return int(self.bits.size)

        def read(self, _type, n):
            out, self.bits = np.packbits(self.bits[: 8 * n]).tobytes(), self.bits[8 * n:]
            return out

        def __str__(self):
            return ''.join('1' if b else '0' for b in self.bits)

    m = types.ModuleType("bitstream")
    m.BitStream = BitStream
    return m


def quantised_kernels(seed=3):
    """1/16-quantised kernels with a skewed, realistic distribution (values times qp=16)."""
    rng = np.random.default_rng(seed)
    shapes = [(3, 8, 5, 5, 5), (8, 16, 5, 5, 5), (16, 8, 5, 5, 5), (8, 8, 4, 4, 4), (1, 8, 3, 3, 3)]
    return [np.rint(rng.laplace(0, 1.3, size=s)).astype(np.float32) for s in shapes]


def main():
    out = {}
    names = []
    for name, sym, mu, sg in arith_cases():
        stream = ref_arith(sym, mu, sg)
        names.append(name)
        out["arith_%s_sym" % name], out["arith_%s_mu" % name], out["arith_%s_sigma" % name] = sym, mu, sg
        out["arith_%s_stream" % name] = np.frombuffer(stream, dtype=np.uint8)
        print("arith %-12s %6d symbols -> %5d bytes" % (name, sym.size, len(stream)))
    out["arith_cases"] = np.array(names)

    sys.modules["bitstream"] = bitstream_stub()
    sys.path.insert(0, REF_ROOT)
    import util_code_quantized_weights as U    # noqa: E402
    pool = quantised_kernels()
    eles = np.concatenate([p.reshape(-1) for p in pool])
    pdf, bins = U.get_pdf(eles)
    codebook, inv = U.get_huffman_codebook(pdf, bins)
    stream, shape_list = U.entropy_encode(pool, codebook)
    dec = U.entropy_decode(stream, inv, len(eles), shape_list)
    assert all(np.array_equal(a, b) for a, b in zip(pool, dec))
    words = list(inv.keys())
    out["huff_words"] = np.array(words)
    out["huff_symbols"] = np.array([int(inv[w]) for w in words], dtype=np.int64)
    out["huff_pdf"], out["huff_bins"] = pdf, bins
    out["huff_stream"] = np.frombuffer(stream, dtype=np.uint8)
    out["huff_rate"] = np.array(U.est_rate(pdf, bins, codebook))
    print("huffman: %d symbols, %d codewords, %d bytes, E(l)=%.4f" % (len(eles), len(words), len(stream), out["huff_rate"]))
    np.savez_compressed(os.path.join(GOLDEN, "entropy_small.npz"), **out)


if __name__ == "__main__":
    main()
