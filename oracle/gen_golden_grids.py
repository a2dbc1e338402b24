"""Generate tests/golden/grids_small.npz by EXECUTING the unmodified reference script
util_get_grids.py (run in the build container only: needs /root/reference).

    python oracle/gen_golden_grids.py

TEST INFRASTRUCTURE ONLY.  open3d is not installed, so the script runs against a stub
`open3d` whose `io.read_point_cloud` hands back the fixture cloud and whose
`geometry.KDTreeFlann.search_knn_vector_3d(p, 1)` answers by exhaustive search.  Everything
else - cube template, axis order, origin handling, sqrt, dtype, `dist == 0`, the file names
it writes - is the reference's own code.  The fixture stores the squared distances as uint16
(the generator asserts that sqrt(d2) reproduces the reference's float64 `dist` bit for bit).
"""
from __future__ import annotations

import os
import runpy
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF_ROOT = os.environ.get("NVF_REFERENCE_ROOT", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def fixture_cloud():
    """A piece of a voxelised sphere shell plus a few isolated voxels, and six leaf origins
    chosen so that nearest points lie in the own leaf, in neighbouring leaves that are NOT
    listed, and (single-voxel leaf) up to a leaf diagonal away."""
    c, R = np.array([70, 75, 80]), 58.0
    ax = np.arange(0, 160)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    d = np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2)
    shell = np.stack(np.nonzero(np.abs(d - R) < 0.5), 1)
    extra = np.array([[200, 40, 40], [231, 63, 63], [224, 32, 95], [5, 5, 200]])
    pts = np.unique(np.concatenate([shell, extra], 0), axis=0).astype(np.int32)
    cells = np.unique(pts // 32, axis=0)
    rng = np.random.default_rng(7)
    pick = rng.choice(len(cells), size=4, replace=False)
    origins = np.concatenate([cells[pick] * 32, np.array([[192, 32, 32], [0, 0, 192]])], 0).astype(np.int32)
    return pts, origins


def run_reference(points: np.ndarray, origins: np.ndarray):
    refs = points.astype(np.float64)

    class _Pcd:
        pass

    class _KDTreeFlann:
        def __init__(self, pcd):
            self.refs = np.asarray(pcd.points)

        def search_knn_vector_3d(self, p, k):
            assert k == 1
            d2 = np.sum(np.square(self.refs - np.asarray(p, dtype=np.float64)), -1)
            i = int(np.argmin(d2))
            return 1, [i], [float(d2[i])]

    def read_point_cloud(_fn):
        pcd = _Pcd()
        pcd.points = refs
        return pcd

    o3d = types.ModuleType("open3d")
    o3d.io = types.SimpleNamespace(read_point_cloud=read_point_cloud)
    o3d.geometry = types.SimpleNamespace(KDTreeFlann=_KDTreeFlann)
    saved = sys.modules.get("open3d")
    sys.modules["open3d"] = o3d
    tmp = tempfile.mkdtemp(prefix="nvf_ref_grids_")
    cwd, argv = os.getcwd(), sys.argv
    try:
        os.chdir(tmp)
        np.savetxt("fixture_l5_origins.txt", origins, delimiter=",", fmt="%d")
        sys.argv = ["util_get_grids.py", "fixture.ply", "5"]
        runpy.run_path(os.path.join(REF_ROOT, "util_get_grids.py"), run_name="__main__")
        gt = np.load("fixture_l5_gt_grid.npy")
        dist = np.load("fixture_l5_dist.npy")
        org = np.load("fixture_l5_origins.npy")
    finally:
        os.chdir(cwd)
        sys.argv = argv
        if saved is None:
            del sys.modules["open3d"]
        else:
            sys.modules["open3d"] = saved
    return gt, dist, org


def main():
    pts, origins = fixture_cloud()
    print("fixture: %d points, %d leaves" % (len(pts), len(origins)))
    gt, dist, org = run_reference(pts, origins)
    assert gt.dtype == np.uint8 and dist.dtype == np.float64 and gt.shape == dist.shape == (len(origins), 1, 32, 32, 32)
    d2 = np.rint(dist ** 2).astype(np.int64)
    assert d2.max() < 65535 and np.array_equal(np.sqrt(d2.astype(np.float64)), dist)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "grids_small.npz"), points=pts, origins=origins,
                        origins_saved=org, origins_saved_dtype=str(org.dtype), gt=np.packbits(gt.reshape(-1)),
                        d2=d2.astype(np.uint16).reshape(len(origins), -1), dist_dtype=str(dist.dtype),
                        gt_dtype=str(gt.dtype))
    print("wrote grids_small.npz: occupied %d, max dist %.3f" % (int(gt.sum()), float(dist.max())))


if __name__ == "__main__":
    main()
