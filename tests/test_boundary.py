"""The parts of the drop-in boundary that real callers reach besides align(): the search method behind
pcl::Registration::getSearchMethodTarget() (apd_nearest_k / apd_source_nearest), and the multi-device batch context with
its one shared queue of pairs (apd_align_batch(..., n_devices))."""
import importlib

import numpy as np
import pytest

from helpers import DEPLOYED


def _brute_knn(cloud, queries, k):
    """FLANN L2_Simple's fp32 arithmetic ((dx*dx + dy*dy) + dz*dz), ordered by (d2, index)"""
    idx = np.empty((queries.shape[0], k), np.int32)
    d2o = np.empty((queries.shape[0], k), np.float32)
    c = cloud[:, :3].astype(np.float32)
    for i, q in enumerate(queries[:, :3].astype(np.float32)):
        d = c - q
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        order = np.lexsort((np.arange(c.shape[0]), d2))[:k]
        idx[i], d2o[i] = order, d2[order]
    return idx, d2o


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 5, 20, 32])
def test_nearest_k_of_arbitrary_queries(gorio, synth, k):
    src, tgt, T = synth.submap_pair(2004, n_source=1500, n_frames=6, n_per_frame=1500)
    g = gorio.FastAPDGICP(0)
    g.set_input_target(tgt)
    g.set_input_source(src)
    rng = np.random.default_rng(k)
    # queries inside the cloud, on its points, and far outside its bounding box
    q = np.concatenate([src[:300, :3], tgt[:100, :3], tgt[:50, :3] + rng.normal(0, 3.0, (50, 3)).astype(np.float32),
                        np.array([[500.0, -400.0, 90.0], [-300.0, 0.0, 0.0]], np.float32)])
    for which, cloud in ((1, tgt), (0, src)):
        gi, gd = g.nearest_k(q, k, which=which)
        bi, bd = _brute_knn(cloud, q, k)
        assert np.array_equal(gi, bi) and np.array_equal(gd, bd)
    g.close()


@pytest.mark.gpu
def test_nearest_k_on_a_tiny_cloud(gorio):
    tiny = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 2, 0, 1]], np.float32)
    g = gorio.FastAPDGICP(0)
    g.set_input_target(tiny)
    gi, gd = g.nearest_k(np.array([[0.1, 0.1, 0.0]], np.float32), 5)
    assert list(gi[0]) == [0, 1, 2, -1, -1] and np.isinf(gd[0, 3:]).all()
    g.close()


@pytest.mark.gpu
def test_source_nearest_is_the_fitness_pass_per_point(gorio, synth):
    src, tgt, T = synth.submap_pair(2005, n_source=1200, n_frames=5, n_per_frame=1500)
    g = gorio.FastAPDGICP(0)
    g.set_params(**DEPLOYED)
    g.set_input_target(tgt)
    g.set_input_source(src)
    r = g.align()
    idx, d2, xyz = g.source_nearest()
    Tf = r["T"].astype(np.float32)
    x, y, z = src[:, 0], src[:, 1], src[:, 2]
    exp = np.stack([((Tf[i, 0] * x + Tf[i, 1] * y) + Tf[i, 2] * z) + Tf[i, 3] for i in range(3)], axis=1).astype(np.float32)
    assert np.array_equal(xyz, exp)  # Eigen's Isometry3f * point, fp32, one rounding per operation
    bi, bd = _brute_knn(tgt, xyz, 1)
    assert np.array_equal(idx, bi[:, 0]) and np.array_equal(d2, bd[:, 0])
    score, n_in, n_inl = g.fitness()
    assert abs(score - d2.astype(np.float64).mean()) / score < 1e-12 and n_in == src.shape[0] and n_inl == int((d2 < 0.25).sum())
    # an explicit pose
    idx2, d22, xyz2 = g.source_nearest(np.eye(4))
    assert np.array_equal(xyz2, src[:, :3]) and np.array_equal(idx2, _brute_knn(tgt, src, 1)[0][:, 0])
    g.close()


@pytest.mark.gpu
def test_multi_device_batch_shares_one_queue(gorio, synth):
    """two pools (here both on device 0 — the box has one GPU; on a node they are different GPUs) drain ONE queue of
    heterogeneous pairs: same results as a single pool, every pair done once, both pools took part"""
    pairs = []
    for i in range(24):  # pairs of very different cost
        s, t, _ = synth.submap_pair(2100 + i, n_source=300 + 150 * (i % 5), n_frames=3 + (i % 4), n_per_frame=800)
        pairs.append((s, t, None))
    one = gorio.Batch(0, n_workers=4, **DEPLOYED)
    ref = one.align(pairs, with_fitness=True)
    one.close()
    multi = gorio.Batch([0, 0], n_workers=4, **DEPLOYED)
    assert multi.devices == [0, 0]
    got = multi.align(pairs, with_fitness=True)
    took = multi.device_pairs()
    assert sum(took) == len(pairs) and min(took) > 0
    for a, b in zip(ref, got):
        assert a["status"] == b["status"] == 0 and np.array_equal(a["T"], b["T"]) and a["fitness"] == b["fitness"]
        assert a["converged"] == b["converged"] and a["iterations"] == b["iterations"]
    # device-resident clouds name one device's memory: refused on a multi-device context
    with pytest.raises(gorio.ApdError):
        multi.align([((1, 10), (1, 10), None)])
    multi.close()


@pytest.mark.gpu
def test_batch_takes_the_pcl_layout(gorio, synth):
    """the pool through stride-48 pageable pcl::PointXYZINormal clouds (the drop-in class's real input) gives the packed result"""
    pairs = [synth.submap_pair(2200 + i, n_source=600, n_frames=4, n_per_frame=900)[:2] for i in range(6)]
    b = gorio.Batch(0, n_workers=3, **DEPLOYED)
    ref = b.align([(s, t, None) for s, t in pairs])
    prep = b.prepare([(synth.to_pcl_xyzinormal(s), synth.to_pcl_xyzinormal(t), None) for s, t in pairs], layout=(48, 0, 16))
    got = b.align(prep)
    for a, c in zip(ref, got):
        assert a["status"] == c["status"] == 0 and np.array_equal(a["T"], c["T"]) and a["fitness"] == c["fitness"]
    b.close()
