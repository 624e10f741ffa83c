"""The only real point clouds of the reference tree (ndt_omp/data/251370668.pcd, 251371071.pcd; committed as
tests/golden/real_lidar_pair.npz by tests/golden/make_real_clouds.py, which also stored what the oracle computed on them).
69 k points each, 7 % exact duplicates: the tie rule (d2, index) decides real neighbour lists here.
CPU: the .pcd reader and the oracle against the fixture. GPU: the CUDA library against the oracle and the fixture, and
the reference test's own bar (gicp_test.cpp:148-166: forward / backward registration within 5 cm and 1 degree)."""
import importlib
import os

import numpy as np
import pytest

from helpers import pose_err
from oracle_binding import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
pcd = importlib.import_module("go-rio_b200.pcd")


@pytest.fixture(scope="module")
def real():
    d = np.load(os.path.join(HERE, "golden", "real_lidar_pair.npz"))
    tgt = np.zeros((d["target_xyz"].shape[0], 4), np.float32); tgt[:, :3] = d["target_xyz"]
    src = np.zeros((d["source_xyz"].shape[0], 4), np.float32); src[:, :3] = d["source_xyz"]
    return d, src, tgt


def test_pcd_reader_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    cloud = rng.normal(size=(257, 4)).astype(np.float32)
    for binary in (True, False):
        path = str(tmp_path / ("b.pcd" if binary else "a.pcd"))
        pcd.write_pcd(path, cloud, binary=binary)
        back = pcd.read_pcd(path)
        assert back.dtype.names == ("x", "y", "z", "intensity") and back.shape[0] == 257
        got = np.stack([back[n] for n in back.dtype.names], axis=1)
        assert np.array_equal(got, cloud) if binary else np.allclose(got, cloud, rtol=0, atol=0)
    bad = cloud.copy(); bad[5, 1] = np.nan
    path = str(tmp_path / "n.pcd")
    pcd.write_pcd(path, bad)
    assert pcd.xyz_label(pcd.read_pcd(path)).shape == (256, 4)  # non-finite points are dropped, label column is zero


def test_fixture_is_the_reference_data(real):
    d, src, tgt = real
    assert tgt.shape[0] == 69088 and src.shape[0] == 69792  # the WIDTH of the two .pcd headers
    assert len(np.unique(tgt[:, :3], axis=0)) == 64057      # exact duplicates are part of the data
    path = "/root/reference/ndt_omp/data/251370668.pcd"
    if os.path.exists(path):  # (in the build container only)
        assert np.array_equal(pcd.xyz_label(pcd.read_pcd(path))[:, :3], tgt[:, :3])


def test_oracle_on_real_clouds_matches_the_fixture(real):
    d, src, tgt = real
    o = Oracle(search=1, threads=os.cpu_count() or 1)
    o.set_params(max_correspondence_distance=2.0, maha_fp64=1)
    o.set_input_target(tgt); o.set_input_source(src)
    e, H, b = o.linearize(np.eye(4))
    assert abs(e - float(d["err_I"])) / e < 1e-12 and np.array_equal(o.get_correspondences()[0], d["corr_I"])
    assert np.array_equal(o.get_neighbors(1)[d["knn_rows"]], d["knn_target"])
    # the brute-force search (the definition) agrees with the kd-tree on the sampled rows, duplicates included
    xyz = tgt[:, :3]
    for r in d["knn_rows"][::50]:
        dd = xyz - xyz[r]
        d2 = (dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]
        assert np.array_equal(np.lexsort((np.arange(xyz.shape[0]), d2))[:20], d["knn_target"][list(d["knn_rows"]).index(r)])
    r = o.align()
    assert np.abs(r["T64"] - d["T64_default"]).max() < 1e-12 and r["iterations"] == int(d["iterations_default"])


@pytest.mark.gpu
def test_gpu_matches_oracle_on_real_clouds(gorio, real):
    d, src, tgt = real
    g = gorio.FastAPDGICP(0)
    g.set_params(max_correspondence_distance=2.0, maha_fp64=1)
    g.set_input_target(tgt); g.set_input_source(src)
    e, H, b = g.linearize(np.eye(4))
    assert np.array_equal(g.get_correspondences()[0], d["corr_I"])  # bit-exact correspondences, 7 % duplicate points
    assert np.array_equal(g.get_neighbors(1)[d["knn_rows"]], d["knn_target"])
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert abs(e - float(d["err_I"])) / e < 1e-10 and rel(H, d["H_I"]) < 1e-10 and rel(b, d["b_I"]) < 1e-9
    for host_loop in (1, 0):  # 69 k source points: the host-driven loop either way (the device loop serves <= 32 768)
        g.set_params(host_loop=host_loop)
        r = g.align()
        dt, dr = pose_err(r["T64"], d["T64_default"])
        assert dt < 1e-6 and dr < 1e-6 and r["iterations"] == int(d["iterations_default"]) and r["converged"]
    # every neighbour list against the oracle, live (thread-per-point and warp-per-point search kernels alike)
    o = Oracle(search=1, threads=os.cpu_count() or 1)
    o.set_params(max_correspondence_distance=2.0, maha_fp64=1)
    o.set_input_target(tgt); o.set_input_source(src)
    o.linearize(np.eye(4))
    assert np.array_equal(g.get_neighbors(1), o.get_neighbors(1)) and np.array_equal(g.get_neighbors(0), o.get_neighbors(0))
    g.close()


@pytest.mark.gpu
def test_reference_alignment_bar_on_real_clouds(gorio, real):
    """gicp_test.cpp:148-166 (the reference's only alignment test, on its own pair of real clouds): forward and backward
    registration with the deployed parameters, each within 5 cm / 1 degree of the pose found on the CPU, both converged,
    and inverse to one another within the same bar"""
    d, src, tgt = real
    g = gorio.FastAPDGICP(0)
    g.set_params(max_correspondence_distance=2.0, transformation_epsilon=0.1)
    g.set_input_target(tgt); g.set_input_source(src)
    fwd = g.align()
    g.swap_source_and_target()
    bwd = g.align()
    t_tol, r_tol = 0.05, np.pi / 180.0
    for got, want in ((fwd, d["T64_deployed"]), (bwd, d["T64_deployed_backward"])):
        dt, dr = pose_err(got["T64"], want)
        assert got["converged"] and dt < t_tol and dr < r_tol
        assert dt < 1e-5 and dr < 1e-5  # (and in fact equal to the oracle's pose within the fp32-storage tolerance)
    dt, dr = pose_err(fwd["T64"] @ bwd["T64"], np.eye(4))
    assert dt < t_tol and dr < r_tol
    g.close()


@pytest.mark.gpu
def test_vgicp_alignment_bar_on_real_clouds(gorio, real):
    """gicp_test.cpp:141-166 runs the same forward / backward bar on FastVGICP: on the reference tree's real lidar pair the
    voxelised registration lands within 5 cm / 1 degree of the pose FastAPDGICP's CPU restatement found, in both
    directions, and the two directions are inverse to one another within the same bar"""
    d, src, tgt = real
    g = gorio.FastAPDGICP(0)
    o = Oracle(search=1, threads=0)
    for r in (g, o):
        r.set_params(variant=2, voxel_resolution=1.0, voxel_search=2)  # the reference's defaults: 1 m voxels, DIRECT1
        r.set_input_target(tgt); r.set_input_source(src)
    fwd, fwd_o = g.align(), o.align()
    coords, counts, _, _ = g.vgicp_voxels()
    assert counts.sum() == tgt.shape[0] and coords.shape[0] > 1000
    assert np.array_equal(coords, o.vgicp_voxels()[0]) and np.array_equal(counts, o.vgicp_voxels()[1])
    assert np.array_equal(g.vgicp_correspondences()[0], o.vgicp_correspondences()[0])
    dt, dr = pose_err(fwd["T64"], fwd_o["T64"])
    assert dt < 1e-6 and dr < 1e-6 and fwd["iterations"] == fwd_o["iterations"]
    g.swap_source_and_target()
    bwd = g.align()
    t_tol, r_tol = 0.05, np.pi / 180.0
    for got, want in ((fwd, d["T64_deployed"]), (bwd, d["T64_deployed_backward"])):
        dt, dr = pose_err(got["T64"], want)
        assert got["converged"] and dt < t_tol and dr < r_tol, (dt, dr)
    dt, dr = pose_err(fwd["T64"] @ bwd["T64"], np.eye(4))
    assert dt < t_tol and dr < r_tol
    g.close()
