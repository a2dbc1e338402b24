"""Test-only: builds (g++) and loads the CPU emulator of the NVF kernels.

The emulator compiles the same kernel bodies / host orchestration as the CUDA
library and runs them sequentially on CPU tensors.  Used by the non-GPU tests
to validate kernel logic; never imported by the product package.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "nvf_emu.cpp")
LIB = os.path.join(HERE, "libnvf_emu.so")
CSRC = os.path.join(ROOT, "nvfpcc_b200", "csrc")
_binding = None


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [SRC, os.path.join(ROOT, "include", "nvf_b200.h")] + [
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build():
    if _stale():
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB, SRC])
    return LIB


def binding():
    global _binding
    if _binding is None:
        from nvfpcc_b200._lib import Binding
        _binding = Binding(build())
    return _binding
