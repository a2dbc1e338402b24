"""Import the UNMODIFIED reference modules from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY, and only usable in the build container (the GPU box
has no /root/reference).  Used by oracle/gen_golden.py to produce the golden
fixtures under tests/golden/ and by tests/test_oracle_vs_reference.py (skipped
when the reference checkout is absent).

Recipe: stub the un-installed third-party modules that
the reference imports but the live classes never use, chdir to a directory
holding a synthetic SEED3.npy / SEED4_Gaussian.npy (utils/network.py:20-21 load
them from the CWD at import time), then import utils.network / utils.loss /
gdn_3d from the reference checkout.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np

REF_ROOT = os.environ.get("NVF_REFERENCE_ROOT", "/root/reference")
SEED_LEN = 300_000


def synthetic_seed(n: int = SEED_LEN) -> np.ndarray:
    """The synthetic stand-in for SEED3.npy."""
    return np.random.default_rng(0).random(n)


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "network.py"))


_cached = None


def load():
    """Returns (network, loss, gdn_3d) reference modules."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REF_ROOT)
    for name in ("open3d", "MinkowskiEngine", "bitstream", "IPython"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    tmp = tempfile.mkdtemp(prefix="nvf_ref_seed_")
    np.save(os.path.join(tmp, "SEED3.npy"), synthetic_seed())
    np.save(os.path.join(tmp, "SEED4_Gaussian.npy"), np.zeros(4))
    cwd = os.getcwd()
    sys.path.insert(0, REF_ROOT)
    try:
        os.chdir(tmp)
        import utils.network as network  # noqa: E402
        import utils.loss as loss  # noqa: E402
        import gdn_3d  # noqa: E402
    finally:
        os.chdir(cwd)
        sys.path.remove(REF_ROOT)
    _cached = (network, loss, gdn_3d)
    return _cached


def build_net(ch: int, channels):
    """Constructs the reference `Net` (NVFPCC.py:32-45) from the reference's
    own live classes.  NVFPCC.py itself is not imported (it needs CUDA and
    prints at import); the 14-line class body is reproduced structurally here
    by composing the reference modules in the same order."""
    import torch.nn as nn

    network, _, _ = load()
    network.seed_ptr = 0

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.latent_gen = network.SingleLayerLatentGen(in_channels=ch, out_channels=ch)
            self.entropy_coder = network.QuantGaussianLikelihood(in_channels=ch)
            self.reconstructor = network.CompDecoder(None, "Gaussian", useIGDN=True, in_channels=ch,
                                                     channels=tuple(int(c) for c in channels))

        def forward(self, emb, mode, q):
            latent = self.latent_gen(emb)
            latent_rounded, latent_likelihood = self.entropy_coder(latent, mode)
            out, out_cls_list, net_bits = self.reconstructor(latent_rounded, q)
            return out, out_cls_list, net_bits, latent_likelihood

        def reconstruct(self, latent, q):
            return self.reconstructor(latent, q)[0]

    return Net()
