"""GPU parity of nvf_build_grids (include/nvf_prep_b200.h) against the oracle and against the
fixture written by the unmodified reference script: distances are float64 and must be bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import grids_oracle as GO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(points, origins, **kw):
    from nvfpcc_b200 import grids
    r = grids.build_grids(points, origins, want_gt=True, want_dist64=True, want_dist32=True, want_d2=True, **kw)
    return {k: v.cpu().numpy() for k, v in r.items()}


def assert_exact(r, gt_ref, dist_ref):
    assert r["dist"].dtype == np.float64 and r["gt"].dtype == np.uint8
    assert np.array_equal(r["dist"], dist_ref)
    assert np.array_equal(r["gt"], gt_ref)
    assert np.array_equal(r["dist32"], dist_ref.astype(np.float32))       # utils/dataloader.py:171 .float()
    assert np.array_equal(r["d2"].astype(np.int64).reshape(dist_ref.shape), np.rint(dist_ref ** 2).astype(np.int64))


def test_matches_reference_script_fixture():
    g = np.load(os.path.join(ROOT, "tests", "golden", "grids_small.npz"))
    n = g["origins"].shape[0]
    dist_ref = np.sqrt(g["d2"].astype(np.float64)).reshape(n, 1, 32, 32, 32)
    gt_ref = np.unpackbits(g["gt"])[: n * 32768].reshape(n, 1, 32, 32, 32)
    r = run(g["points"], g["origins"], max_cells=256)
    assert_exact(r, gt_ref, dist_ref)


def test_sphere_shell_all_leaves_match_oracle():
    from nvfpcc_b200 import synth
    pts = synth.sphere_shell_points(256)
    origins = synth.leaf_origins(pts)
    gt_ref, dist_ref = GO.build_grids(pts, origins, workers=-1)
    r = run(pts, origins)
    assert_exact(r, gt_ref, dist_ref)
    assert int(r["gt"].sum()) == pts.shape[0]                              # every point lands in exactly one leaf
    r2 = run(pts, origins)
    assert all(np.array_equal(r[k], r2[k]) for k in r)                     # deterministic


def test_unaligned_and_negative_origins():
    rng = np.random.default_rng(3)
    pts = np.unique(rng.integers(-70, 120, size=(4000, 3)), axis=0).astype(np.int32)
    origins = np.array([[-37, 5, 11], [0, 0, 0], [13, -64, 50], [77, 81, -3], [-70, -70, -70]], dtype=np.int32)
    # every leaf must hold a point for the exact-radius guarantee: add one inside each
    pts = np.unique(np.concatenate([pts, origins + rng.integers(0, 32, size=origins.shape)], 0), axis=0).astype(np.int32)
    gt_ref, dist_ref = GO.build_grids(pts, origins)
    r = run(pts, origins, max_cells=1024)
    assert_exact(r, gt_ref, dist_ref)


def test_single_point_leaf_reaches_the_full_diagonal():
    # one voxel in the far corner of its leaf; the only other points are farther than the leaf diagonal
    pts = np.array([[31, 31, 31], [200, 0, 0], [0, 200, 0]], dtype=np.int32)
    origins = np.array([[0, 0, 0]], dtype=np.int32)
    gt_ref, dist_ref = GO.build_grids(pts, origins)
    r = run(pts, origins)
    assert_exact(r, gt_ref, dist_ref)
    assert r["dist"][0, 0, 0, 0, 0] == np.sqrt(3 * 31 * 31)
    # nearer points in neighbouring cells on every side win over the own point
    pts2 = np.concatenate([pts, [[-1, 0, 0], [0, -2, 5], [3, 3, -3], [32, 31, 31], [16, 40, 16]]]).astype(np.int32)
    gt_ref, dist_ref = GO.build_grids(pts2, origins)
    assert_exact(run(pts2, origins), gt_ref, dist_ref)


def test_dense_and_sparse_random_leaves():
    rng = np.random.default_rng(11)
    dense = rng.integers(0, 64, size=(60000, 3))
    sparse = rng.integers(64, 192, size=(60, 3))
    pts = np.unique(np.concatenate([dense, sparse], 0), axis=0).astype(np.int32)
    origins = (np.unique(pts // 32, axis=0) * 32).astype(np.int32)
    gt_ref, dist_ref = GO.build_grids(pts, origins)
    assert_exact(run(pts, origins), gt_ref, dist_ref)


def test_status_flags_raise():
    from nvfpcc_b200 import _lib, grids
    pts = np.array([[5, 5, 5]], dtype=np.int32)
    with pytest.raises(_lib.NvfError, match="no cloud point"):
        grids.build_grids(pts, np.array([[320, 320, 320]], dtype=np.int32))          # empty leaf, nothing in reach
    many = (np.arange(300)[:, None] * np.array([[32, 0, 0]])).astype(np.int32)
    with pytest.raises(_lib.NvfError, match="max_cells"):
        grids.build_grids(many, many[:2], max_cells=64)
    r = grids.build_grids(pts, np.array([[320, 320, 320]], dtype=np.int32), check=False, want_d2=True)
    assert int(r["status"].item()) == grids.STATUS_NOT_FOUND and bool(torch.isinf(r["dist"]).all())


def test_cli_writes_the_reference_files(tmp_path, monkeypatch):
    from nvfpcc_b200 import grids, synth
    pts = synth.sphere_shell_points(128)
    origins = synth.leaf_origins(pts)
    monkeypatch.chdir(tmp_path)
    with open("cloud.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nend_header\n" % len(pts))
        np.savetxt(f, pts, fmt="%d")
    np.savetxt("cloud_l5_origins.txt", origins, delimiter=",", fmt="%d")
    assert grids.main(["grids", "sub/dir/../../cloud.ply".replace("sub/dir/../../", ""), "5"]) == 0
    gt, dist, org = np.load("cloud_l5_gt_grid.npy"), np.load("cloud_l5_dist.npy"), np.load("cloud_l5_origins.npy")
    assert gt.dtype == np.uint8 and dist.dtype == np.float64 and org.dtype == np.float64
    gt_ref, dist_ref = GO.build_grids(pts, origins)
    assert np.array_equal(gt, gt_ref) and np.array_equal(dist, dist_ref) and np.array_equal(org, origins)


def test_full_vox10_cloud_properties_and_sampled_oracle():
    """BASELINE size (849 338 points, 1247 leaves): size-independent properties of a distance field on every leaf,
    exact comparison with the oracle on a random sample of leaves."""
    from nvfpcc_b200 import grids, synth
    pts = synth.sphere_shell_points(1024)
    origins = synth.leaf_origins(pts)
    r = grids.build_grids(pts, origins, want_d2=True, max_cells=origins.shape[0])
    gt, dist = r["gt"], r["dist"]
    assert int(gt.sum().item()) == pts.shape[0]                        # every cloud point is one occupied voxel
    assert bool(((dist == 0) == (gt == 1)).all())
    assert float(dist.max().item()) <= 31 * np.sqrt(3) + 1e-9          # non-empty leaves bound the distance
    d = dist[:, 0]
    for ax in (1, 2, 3):                                               # 1-Lipschitz along every axis
        assert float((d.narrow(ax, 1, 31) - d.narrow(ax, 0, 31)).abs().max().item()) <= 1.0 + 1e-12
    d2 = r["d2"].cpu().numpy().astype(np.int64)
    assert np.array_equal(np.sqrt(d2.astype(np.float64)).reshape(dist.shape), dist.cpu().numpy())
    # the occupied voxels, mapped back through the origins, are exactly the cloud
    idx = torch.nonzero(gt[:, 0]).cpu().numpy()
    back = origins[idx[:, 0]].astype(np.int64) + idx[:, 1:]
    key = lambda a: np.sort(a[:, 0] * (1 << 40) + a[:, 1] * (1 << 20) + a[:, 2])
    assert np.array_equal(key(back), key(pts.astype(np.int64)))
    sel = np.random.default_rng(0).choice(origins.shape[0], size=24, replace=False)
    gt_ref, dist_ref = GO.build_grids(pts, origins[sel], workers=-1)
    assert np.array_equal(dist.cpu().numpy()[sel], dist_ref) and np.array_equal(gt.cpu().numpy()[sel], gt_ref)


def test_empty_inputs():
    from nvfpcc_b200 import _lib, grids
    pts = np.array([[1, 2, 3]], dtype=np.int32)
    r = grids.build_grids(pts, np.zeros((0, 3), np.int32))
    assert r["gt"].shape == (0, 1, 32, 32, 32) and r["dist"].shape == (0, 1, 32, 32, 32)
    r = grids.build_grids(np.zeros((0, 3), np.int32), np.zeros((1, 3), np.int32), check=False)
    assert int(r["status"].item()) == grids.STATUS_NOT_FOUND
    with pytest.raises(ValueError):
        grids.build_grids(np.zeros((4, 2), np.int32), np.zeros((1, 3), np.int32))
    with pytest.raises(ValueError):
        grids.build_grids(np.array([[0.5, 1, 2]]), np.zeros((1, 3), np.int32))


def test_host_results_equal_device_float64():
    """build_grids_host (pinned staging buffers) returns exactly the float64 distances and uint8 grid of the device path."""
    from nvfpcc_b200 import grids, synth
    pts = synth.sphere_shell_points(256)
    origins = synth.leaf_origins(pts)[:40]
    dev = grids.build_grids(pts, origins, want_gt=True, want_dist64=True)
    host = grids.build_grids_host(pts, origins)
    assert host["dist"].dtype == np.float64 and host["gt"].dtype == np.uint8
    assert np.array_equal(host["dist"], dev["dist"].cpu().numpy())
    assert np.array_equal(host["gt"], dev["gt"].cpu().numpy())
