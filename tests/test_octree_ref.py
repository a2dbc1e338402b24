"""synth.leaf_origins (the block order every sharded path relies on) against the reference's own octree
builder, compiled unmodified into oracle/_ref/get_octree (oracle/Makefile).  Skipped when the binary is absent."""
import os
import subprocess as sp

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "get_octree")


@pytest.mark.skipif(not os.path.isfile(EXE), reason="oracle/_ref not built")
def test_leaf_origins_match_get_octree(tmp_path):
    from nvfpcc_b200 import synth
    # a small shell inside the reference's fixed 1024^3 root (get_octree.cpp:783), colours as its parser expects
    pts = synth.sphere_shell_points(1024, radius=90.0)
    ply = tmp_path / "cloud.ply"
    with open(ply, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % len(pts))
        np.savetxt(f, np.concatenate([pts, np.zeros_like(pts)], 1), fmt="%d")
    out1, out2 = tmp_path / "origins.txt", tmp_path / "subtree.txt"
    sp.run([EXE, str(ply), str(out1), str(out2)], check=True, stdout=sp.DEVNULL, stderr=sp.DEVNULL, timeout=600)
    ref = np.loadtxt(out1, delimiter=",", ndmin=2).astype(np.int64)
    got = synth.leaf_origins(pts).astype(np.int64)
    assert ref.shape == got.shape
    assert np.array_equal(ref, got)          # same leaves, same (DFS / Morton, z-major) order
