"""Pins the C++ oracle against the independent NumPy/SciPy restatement
(tests/numpy_restatement.py). The reference has no golden vectors for this path
(SURVEY.md §4), so this cross-check plus the property tests are what pin it."""
import numpy as np
import pytest

import numpy_restatement as nr
from oracle_binding import Oracle

REGS = {"NONE": 0, "MIN_EIG": 1, "NORMALIZED_MIN_EIG": 2, "PLANE": 3, "FROBENIUS": 4}


@pytest.fixture(scope="module")
def pair(synth):
    return synth.scan_pair(1001, 600)


def _rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def test_knn_matches_ckdtree(pair):
    src, tgt, _ = pair
    o = Oracle(search=1)
    o.set_input_source(src); o.set_input_target(tgt)
    o.get_target_covariances()
    nb = o.get_neighbors(1)
    d, idx = nr.knn_sets(tgt[:, :3], 20)
    # compare where the fp64 gap between the 20th and 21st neighbour is clear of fp32 rounding
    clear = (d[:, 20] - d[:, 19]) > 1e-4 * d[:, 20]
    assert clear.mean() > 0.95
    same = np.array([set(a) == set(b) for a, b in zip(nb[clear], idx[clear, :20])])
    assert same.all()
    assert (nb[:, 0] == np.arange(tgt.shape[0])).all()  # self is the nearest (d2 = 0)


def test_brute_kd_nanoflann_agree(pair):
    src, tgt, _ = pair
    outs = []
    for search, ref in ((0, False), (1, False), (2, True)):
        try:
            o = Oracle(search=search, ref=ref)
        except OSError:
            pytest.skip("oracle/_ref not built")
        o.set_input_source(src); o.set_input_target(tgt)
        o.get_target_covariances()
        outs.append(o.get_neighbors(1))
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("reg", list(REGS))
def test_covariances(pair, reg):
    src, tgt, _ = pair
    o = Oracle(search=1)
    o.set_params(regularization=REGS[reg])
    o.set_input_source(src); o.set_input_target(tgt)
    C = o.get_target_covariances()[:, :3, :3]
    nb = o.get_neighbors(1)
    Cn = nr.covariances(tgt, nb, reg)
    if reg in ("PLANE", "MIN_EIG", "NORMALIZED_MIN_EIG"):
        # the smallest singular vector is ill-conditioned when sigma2 ~ sigma3: compare well-separated points
        S = np.linalg.svd(nr.covariances(tgt, nb, "NONE"), compute_uv=False)
        ok = (S[:, 1] - S[:, 2]) / S[:, 0] > 1e-3
        assert ok.mean() > 0.9
        assert np.abs(C[ok] - Cn[ok]).max() < 1e-9
    else:
        assert _rel(C, Cn) < 1e-12
    assert np.abs(o.get_target_covariances()[:, 3, :]).max() == 0.0


@pytest.mark.parametrize("variant", [0, 1], ids=["apdgicp", "fastgicp"])
def test_update_correspondences_and_linearize(pair, synth, variant):
    """variant 1: the oracle's FastGICP switch (fast_gicp_impl.hpp:157 no noise term, :205 unit weights) against NumPy"""
    gicp = variant == 1
    src, tgt, Tgt = pair
    T = Tgt @ synth.make_pose([0.05, -0.03, 0.01], np.deg2rad([0.1, -0.2, 0.4]))
    o = Oracle(search=1)
    o.set_params(max_correspondence_distance=2.0, variant=variant)
    o.set_input_source(src); o.set_input_target(tgt)
    err, H, b = o.linearize(T)
    corr, sq = o.get_correspondences()
    # correspondences: fp32 query against a brute-force fp32 scan
    q = nr.transform_f32(T, src[:, :3])
    d2 = ((q[:, None, :] - tgt[None, :, :3]) ** 2)
    d2 = ((d2[:, :, 0] + d2[:, :, 1]).astype(np.float32) + d2[:, :, 2]).astype(np.float32)
    nn = d2.argmin(axis=1)
    dmin = d2[np.arange(src.shape[0]), nn]
    expect = np.where(dmin.astype(np.float64) < 4.0, nn, -1)
    assert np.array_equal(corr, expect)
    assert np.array_equal(sq, dmin)
    assert (corr >= 0).sum() > 50
    # Mahalanobis, H, b, err against numpy
    cs = o.get_source_covariances()[:, :3, :3]
    ct = o.get_target_covariances()[:, :3, :3]
    M = nr.mahalanobis(T, src, tgt, cs, ct, corr, gicp=gicp)
    Mo = o.get_mahalanobis()[:, :3, :3]
    assert _rel(Mo, M) < 1e-9
    e2, H2, b2 = nr.linearize(T, src, tgt, cs, corr, M, gicp=gicp)
    assert abs(err - e2) / e2 < 1e-10
    assert _rel(H, H2) < 1e-10
    assert _rel(b, b2) < 1e-9
    # compute_error at a trial pose uses the stale correspondences / Mahalanobis (:310-346)
    T2 = synth.make_pose([0.01, 0.0, 0.0], [0, 0, 0.001]) @ T
    e3 = o.compute_error(T2)
    e4, _, _ = nr.linearize(T2, src, tgt, cs, corr, M, gicp=gicp)
    assert abs(e3 - e4) / e4 < 1e-10


def test_lm_step_matches_numpy(pair):
    """One LM outer iteration re-derived with numpy.linalg.solve (lsq_registration_impl.hpp:127-170)."""
    src, tgt, _ = pair
    o = Oracle(search=1)
    o.set_params(max_correspondence_distance=2.0, max_iterations=1)
    o.set_input_source(src); o.set_input_target(tgt)
    r = o.align()
    tr = o.lm_trace()
    assert tr.shape[0] >= 1 and r["iterations"] == 0
    y0, H, b = o.linearize(np.eye(4))
    assert abs(y0 - tr[0, 2]) / y0 < 1e-12
    lam = 1e-9 * np.abs(np.diag(H)).max()
    assert abs(lam - tr[0, 5]) / lam < 1e-12
    d = np.linalg.solve(H + lam * np.eye(6), -b)
    assert abs(np.linalg.norm(d) - tr[0, 6]) / np.linalg.norm(d) < 1e-8
    if tr[0, 7] == 1.0:  # accepted
        X = np.eye(4); X[:3, :3] = nr.so3_exp(d[:3]); X[:3, 3] = d[3:]
        assert np.abs(X - r["T64"]).max() < 1e-9
        assert np.abs(X.astype(np.float32) - r["T"]).max() == 0 or np.abs(X - r["T"]).max() < 1e-6


# ---- FastVGICP ----
VGICP = dict(variant=2)


@pytest.mark.parametrize("method,mode", [("DIRECT1", 0), ("DIRECT7", 0), ("DIRECT27", 1), ("DIRECT7", 2)])
def test_vgicp_voxelmap_correspondences_and_sums(pair, method, mode):
    """the oracle's FastVGICP (voxel map, voxel correspondences, H / b / err, stale compute_error) against the NumPy
    restatement with a dict as the voxel map"""
    src, tgt, Tgt = pair
    search = {"DIRECT27": 0, "DIRECT7": 1, "DIRECT1": 2}[method]
    res = 1.5
    o = Oracle(search=1)
    o.set_params(**VGICP, voxel_resolution=res, voxel_search=search, voxel_mode=mode)
    o.set_input_source(src); o.set_input_target(tgt)
    cs, ct = o.get_source_covariances()[:, :3, :3], o.get_target_covariances()[:, :3, :3]
    T = Tgt @ nr_pose(0.05, 0.01)
    err, H, b = o.linearize(T)
    vox = nr.vgicp_voxelmap(tgt, ct, res, multiplicative=(mode == 2))
    coords, counts, means, covs = o.vgicp_voxels()
    assert len(vox) == coords.shape[0] and counts.sum() == tgt.shape[0]
    for c, n, m, C in zip(coords, counts, means, covs):
        v = vox[tuple(int(x) for x in c)]
        assert v[0] == n and _rel(m, v[1][:3]) < (1e-9 if mode == 2 else 1e-12) and _rel(C, v[2][:3, :3]) < 1e-9  # (multiplicative: sums of inverses of 1e-3-conditioned matrices)
    # ascending (z, y, x)
    key = (coords[:, 2].astype(np.int64) * 4096 + coords[:, 1]) * 4096 + coords[:, 0]
    assert (np.diff(key) > 0).all()
    e_n, H_n, b_n, corr_n = nr.vgicp_linearize(T, src, cs, vox, res, method)
    vc, maha = o.vgicp_correspondences()
    got = [(i, tuple(int(x) for x in coords[v])) for i in range(vc.shape[0]) for v in vc[i] if v >= 0]
    assert got == corr_n and len(got) > 50
    assert abs(err - e_n) / e_n < 1e-10 and _rel(H, H_n) < 1e-10 and _rel(b, b_n) < 1e-9
    # compute_error at a trial pose: stale correspondences and Mahalanobis matrices
    T2 = T @ nr_pose(0.02, 0.004)
    e2 = o.compute_error(T2)
    e2_n = nr.vgicp_linearize(T2, src, cs, vox, res, method, T_corr=T)[0]
    assert abs(e2 - e2_n) / e2_n < 1e-10


def nr_pose(t, r):
    T = np.eye(4)
    T[:3, :3] = nr.so3_exp(np.array([r, -0.5 * r, 2 * r]))
    T[:3, 3] = [t, -t, 0.3 * t]
    return T


def test_vgicp_align_recovers_the_motion(pair):
    """the whole FastVGICP registration in the oracle (2 m voxels, DIRECT27): converges, and lands within the noise of the
    600-point scans of the true motion"""
    src, tgt, Tgt = pair
    o = Oracle(search=1)
    o.set_params(variant=2, voxel_resolution=2.0, voxel_search=0)
    o.set_input_source(src); o.set_input_target(tgt)
    r = o.align()
    assert r["converged"] and 1 < r["iterations"] < 40
    d = np.linalg.inv(Tgt) @ r["T64"]
    assert np.abs(d[:3, 3]).max() < 0.25 and np.abs(d[:3, :3] - np.eye(3)).max() < 0.02
