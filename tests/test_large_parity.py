"""Oracle parity at a size larger than L2: the seed-4000 tiled 2 M-point pair (256 MB of per-point state against the
126 MB L2) against tests/golden/tiled_2m_oracle.npz, which tests/golden/make_large_fixture.py wrote from the CPU oracle.
This is where the GPU runs its large-cloud kernels — the radix-sort grid build, the thread-per-point kNN kernel, the
one-lane search kernel, the bulk-copy linearize ring — which the small parity cases never reach."""
import importlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def big(synth):
    d = np.load(os.path.join(HERE, "golden", "tiled_2m_oracle.npz"))
    src, tgt, T = synth.tiled_cloud_pair(int(d["seed"]), int(d["n"]))
    return d, src, tgt, T


def _checksum(c):
    c = c.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    i = np.arange(c.shape[0], dtype=np.uint64)
    return int(np.bitwise_xor.reduce((c + np.uint64(1)) * (i * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019))))


def test_fixture_describes_the_generated_clouds(big):
    d, src, tgt, T = big
    assert src.shape == (2_000_000, 4) and np.array_equal(src[:4], d["src_head"]) and np.array_equal(tgt[:4], d["tgt_head"])
    assert np.array_equal(T, d["T"])


@pytest.mark.gpu
def test_large_cloud_matches_the_oracle_fixture(gorio, big):
    d, src, tgt, T = big
    g = gorio.FastAPDGICP(0)
    g.set_params(max_correspondence_distance=2.0, maha_fp64=1)
    g.set_input_target(tgt)
    g.set_input_source(src)
    e, H, b = g.linearize(T)
    rel = lambda a, bb: float(np.abs(a - bb).max() / np.abs(bb).max())
    assert abs(e - float(d["err"])) / e < 1e-10 and rel(H, d["H"]) < 1e-10 and rel(b, d["b"]) < 1e-9
    c, sq = g.get_correspondences()
    rows = d["knn_rows"]
    assert int((c >= 0).sum()) == int(d["n_matched"]) and _checksum(c) == int(d["corr_checksum"])  # all 2 M correspondences, bit-exact
    assert np.array_equal(c[rows], d["corr_rows"]) and np.array_equal(sq[rows][c[rows] >= 0], d["sqd_rows"][d["corr_rows"] >= 0])
    assert abs(g.compute_error(d["T_trial"]) - float(d["err_trial_stale"])) / float(d["err_trial_stale"]) < 1e-10
    assert np.array_equal(g.get_neighbors(1)[rows], d["knn_target"]) and np.array_equal(g.get_neighbors(0)[rows], d["knn_source"])
    cov = g.get_source_covariances()[rows]
    assert np.abs(cov - d["cov_source_rows"]).max() < 1e-9
    # a second, warm-started pass at the trial pose and back: still the oracle's sums at T
    g.linearize(d["T_trial"])
    e2, H2, _ = g.linearize(T)
    assert abs(e2 - float(d["err"])) / e2 < 1e-10 and rel(H2, d["H"]) < 1e-10
    # a chain of millimetre motions and back (matches kept by their stored bounds, corr.cu): the same correspondences,
    # distances and sums as the first, cold pass
    Tk = T.copy()
    for k in range(4):
        Tk[:3, 3] += np.array([0.002, -0.001, 0.0015])
        g.linearize(Tk)
    e3, H3, b3 = g.linearize(T)
    c3, sq3 = g.get_correspondences()
    assert e3 == e2 and np.array_equal(H3, H2)
    assert np.array_equal(c3, c) and np.array_equal(sq3[c3 >= 0], sq[c >= 0])
    g.close()


@pytest.mark.gpu
def test_large_cloud_sharded_matches_the_oracle_fixture(gorio, big):
    """the same sums from 4 in-process ranks (chunk table, in-kernel exchange), against the oracle's 28 doubles"""
    d, src, tgt, T = big
    grp = gorio.Group([0, 0, 0, 0], max_correspondence_distance=2.0, maha_fp64=1)
    grp.set_input_target(tgt)
    grp.set_input_source(src)
    e, H, b = grp.linearize(T)
    rel = lambda a, bb: float(np.abs(a - bb).max() / np.abs(bb).max())
    assert abs(e - float(d["err"])) / e < 1e-10 and rel(H, d["H"]) < 1e-10 and rel(b, d["b"]) < 1e-9
    got = np.full(src.shape[0], -1, np.int32)
    for r in grp.ranks:
        c_r, _ = r.get_correspondences()
        got[c_r >= 0] = c_r[c_r >= 0]
    assert _checksum(got) == int(d["corr_checksum"])
    grp.close()
