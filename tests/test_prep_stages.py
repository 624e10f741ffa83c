"""The stages on either side of the registration that run on the same grid (SURVEY.md 8f "next" rows): the radius
searches of the preprocessing nodelet (RadiusOutlierRemoval, DBSCAN's neighbour queries) and the submap assembly +
pcl::VoxelGrid downsample of the scan-to-map branch — the CUDA library against the CPU restatement
(oracle/apd_prep_oracle.cpp), bit for bit. CPU: the restatement against an independent NumPy computation."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_binding import oracle_lib


def o_radius(cloud, radius):
    lib = oracle_lib()
    n = cloud.shape[0]
    c = np.ascontiguousarray(cloud, np.float32)
    counts = np.empty(n, np.int32)
    lib.apdo_radius_search(c.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_double(radius), counts.ctypes.data_as(C.c_void_p), None, None)
    offsets = np.zeros(n + 1, np.int64)
    offsets[1:] = np.cumsum(counts)
    idx = np.empty(int(offsets[-1]), np.int32)
    lib.apdo_radius_search(c.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_double(radius), None, offsets.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p))
    return counts, offsets, idx


def o_voxel(cloud, leaf):
    lib = oracle_lib()
    c = np.ascontiguousarray(cloud, np.float32)
    out = np.empty((max(1, c.shape[0]), 4), np.float32)
    m = C.c_int32()
    rc = lib.apdo_voxel_grid(c.ctypes.data_as(C.c_void_p), C.c_int32(c.shape[0]), C.c_float(leaf), out.ctypes.data_as(C.c_void_p), C.byref(m))
    return out[: m.value].copy(), rc


def o_submap(clouds, poses):
    lib = oracle_lib()
    arrs = [np.ascontiguousarray(c, np.float32) for c in clouds]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    ns = (C.c_int32 * len(arrs))(*[a.shape[0] for a in arrs])
    P = np.ascontiguousarray(np.stack([np.asarray(T, np.float64).T for T in poses])).reshape(-1)
    out = np.empty((sum(a.shape[0] for a in arrs), 4), np.float32)
    m = C.c_int32()
    lib.apdo_submap_assemble(ptrs, ns, P.ctypes.data_as(C.c_void_p), C.c_int32(len(arrs)), out.ctypes.data_as(C.c_void_p), C.byref(m))
    return out[: m.value]


def test_radius_oracle_against_numpy(synth):
    src, _, _ = synth.scan_pair(1200, 900)
    counts, offsets, idx = o_radius(src, 2.0)
    x = src[:, :3].astype(np.float32)
    for i in (0, 17, 400, src.shape[0] - 1):
        d = x[i] - x
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        want = np.flatnonzero(d2 < np.float32(2.0) * np.float32(2.0))
        assert counts[i] == want.size and np.array_equal(idx[offsets[i]:offsets[i + 1]], want) and i in want


def test_voxel_oracle_against_numpy(synth):
    _, tgt, _ = synth.submap_pair(2400, n_source=200, n_frames=3, n_per_frame=700)
    leaf = 0.5
    out, rc = o_voxel(tgt, leaf)
    assert rc == 0
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(tgt[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(axis=0)
    dims = ijk.max(axis=0) + 1
    key = ijk[:, 0] + ijk[:, 1] * dims[0] + ijk[:, 2] * dims[0] * dims[1]
    uniq, inverse = np.unique(key, return_inverse=True)
    assert out.shape[0] == uniq.size
    mean = np.stack([np.bincount(inverse, weights=tgt[:, a].astype(np.float64)) / np.bincount(inverse) for a in range(3)], axis=1)
    assert np.abs(out[:, :3] - mean).max() < 1e-4  # (float sums against double sums)
    any_label = np.bincount(inverse, weights=(tgt[:, 3] > 0)) > 0
    assert np.array_equal(out[:, 3] > 0, any_label) and set(np.unique(out[:, 3])) <= {0.0, 1.0}
    # a leaf far too small for the extent: PCL passes the cloud through
    wide = tgt.copy(); wide[0, 0] += 1e6
    out2, rc2 = o_voxel(wide, 0.001)
    assert rc2 == 1 and np.array_equal(out2, wide)


def _dbscan_cases():
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_dbscan_reference as m

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dbscan_reference.npz"))
    return m, d


def test_dbscan_restatement_equals_the_reference_code(synth):
    """tests/golden/dbscan_reference.npz was written by the REFERENCE's own DBSCAN headers (DBSCAN_simple.h / DBSCAN_kdtree.h
    compiled in place, oracle/ref_dbscan.cpp): the NumPy transcription gives the same labels and cluster counts on every
    case — 1 to 154 clusters, both parameter sets; and, where oracle/_ref holds that build, so does a live run of it"""
    import numpy_restatement as nr

    m, d = _dbscan_cases()
    lib_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libapd_ref_dbscan.so")
    lib = C.CDLL(lib_path) if os.path.exists(lib_path) else None
    for name, spec, params in m.CASES:
        cloud = m.cloud_of(spec)
        assert abs(cloud[:, :3].astype(np.float64).sum() - float(d[name + "_cloud_sum"])) < 1e-6  # the cloud the fixture was made on
        eps, mp, lo, hi = params
        labels, nc = nr.dbscan_labels(cloud, eps, int(mp), int(lo), int(hi))
        assert nc == int(d[name + "_n_clusters"]) and np.array_equal(labels, d[name + "_labels"]), name
        if lib is not None:
            live, nc_live = m.reference_labels(lib, cloud, params)
            assert nc_live == nc and np.array_equal(live, labels)


def test_dbscan_restatement_is_sane(synth):
    """the NumPy transcription of DBSCANKdtreeCluster + the nodelet's ranking: labels are 1 .. n_clusters, clusters nearer
    to the sensor get the smaller label, members of a cluster are mutually reachable at the expansion radius"""
    import numpy_restatement as nr

    _, cloud, _ = synth.submap_pair(2500, n_source=500, n_frames=3, n_per_frame=1200)
    labels, nc = nr.dbscan_labels(cloud)
    assert nc >= 5 and set(np.unique(labels)) == set(range(nc + 1))
    rng = [np.linalg.norm(cloud[labels == k, :3].mean(axis=0)) for k in range(1, nc + 1)]
    assert all(a <= b + 1e-4 for a, b in zip(rng, rng[1:]))
    assert min((labels == k).sum() for k in range(1, nc + 1)) >= 20


@pytest.mark.gpu
@pytest.mark.parametrize("params", [dict(), dict(eps=1.2, core_min_pts=6, min_cluster=10, max_cluster=300)])
def test_dbscan_labels_match_the_restatement(gorio, synth, params):
    """the cluster labels the registration reads in normal_x (preprocessing_nodelet_ntu.cpp:520-567): GPU radius searches with
    the reference's range-dependent radii + the reference's growth loop, against the literal NumPy transcription"""
    import numpy_restatement as nr

    src, tgt, _ = synth.submap_pair(2501, n_source=2500, n_frames=6, n_per_frame=1500)
    g = gorio.FastAPDGICP(0)
    g.set_input_target(tgt); g.set_input_source(src)
    kw = dict(eps=0.9, min_pts=10, min_cluster=20, max_cluster=25000)
    kw.update({{"core_min_pts": "min_pts"}.get(k, k): v for k, v in params.items()})
    for which, cloud in ((0, src), (1, tgt)):
        want, nc_want = nr.dbscan_labels(cloud, **kw)
        got, nc = g.dbscan_labels(which=which, **params)
        assert nc == nc_want and np.array_equal(got, want)
    assert nc_want >= 5
    g.close()


@pytest.mark.gpu
def test_dbscan_labels_match_the_reference_code(gorio, synth):
    """apd_dbscan_labels against the output of the reference's own DBSCAN headers (tests/golden/dbscan_reference.npz)"""
    m, d = _dbscan_cases()
    g = gorio.FastAPDGICP(0)
    for name, spec, params in m.CASES:
        cloud = m.cloud_of(spec)
        g.set_input_source(cloud)
        eps, mp, lo, hi = params
        got, nc = g.dbscan_labels(eps=eps, core_min_pts=int(mp), min_cluster=int(lo), max_cluster=int(hi), which=0)
        assert nc == int(d[name + "_n_clusters"]) and np.array_equal(got, d[name + "_labels"]), name
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("radius", [0.9, 2.0])
def test_radius_search_matches_the_oracle(gorio, synth, radius):
    """RadiusOutlierRemoval (radius 2, min 2 neighbours) and DBSCAN (eps 0.9) parameters of the preprocessing nodelet"""
    src, tgt, _ = synth.submap_pair(2401, n_source=1500, n_frames=4, n_per_frame=1500)
    g = gorio.FastAPDGICP(0)
    g.set_input_target(tgt); g.set_input_source(src)
    for which, cloud in ((0, src), (1, tgt)):
        oc, oo, oi = o_radius(cloud, radius)
        gc, go, gi = g.radius_search(radius, which=which, lists=True)
        assert np.array_equal(gc, oc) and np.array_equal(go, oo) and np.array_equal(g.radius_search(radius, which=which), oc)
        for i in range(cloud.shape[0]):  # the same SETS (rows are unordered on the GPU)
            assert np.array_equal(np.sort(gi[go[i]:go[i + 1]]), oi[oo[i]:oo[i + 1]])
        keep_g, keep_o = gc > 2, oc > 2  # pcl::RadiusOutlierRemoval: kept when k (itself included) > min_neighbors
        assert np.array_equal(keep_g, keep_o)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("leaf", [0.1, 0.5, 2.0])
def test_voxel_downsample_matches_the_oracle(gorio, synth, leaf):
    _, tgt, _ = synth.submap_pair(2402, n_source=200, n_frames=5, n_per_frame=1500)
    tgt = tgt.copy(); tgt[5, 1] = np.nan; tgt[77, 0] = np.inf  # non-finite points are skipped
    g = gorio.FastAPDGICP(0)
    got = g.voxel_downsample(tgt, leaf)
    want, rc = o_voxel(tgt, leaf)
    assert rc == 0 and got.shape == want.shape and np.array_equal(got, want)  # same voxels, same order, same float sums
    # the 48-byte PCL layout goes through the same staging
    assert np.array_equal(g.voxel_downsample(synth.to_pcl_xyzinormal(np.nan_to_num(tgt, posinf=0.0)), leaf), o_voxel(np.nan_to_num(tgt, posinf=0.0), leaf)[0])
    g.close()


@pytest.mark.gpu
def test_submap_assembly_matches_the_oracle_and_registers(gorio, synth):
    """scan_matching_odometry_nodelet.cpp:602-618 end to end: keyframes -> submap (on the device) -> target of a registration"""
    frames = list(synth.drive_frames(5003, 6, 1200))
    clouds = [np.ascontiguousarray(c) for _, c, _ in frames[:5]]
    gts = [gt for _, _, gt in frames]
    poses = [np.linalg.inv(gts[i]) @ gts[4] for i in range(5)]  # odom_i^-1 * odom_last, as the nodelet composes it
    poses = [np.linalg.inv(P) for P in poses]  # (points of keyframe i expressed in the last keyframe's frame)
    g = gorio.FastAPDGICP(0)
    g.set_params(max_correspondence_distance=2.0, transformation_epsilon=0.1)
    raw = g.submap_assemble(clouds, poses, 0.0)
    assert np.array_equal(raw, o_submap(clouds, poses))  # double-precision transform, cast to float: bit-exact
    sub = g.submap_assemble(clouds, poses, 0.1, set_as_target=True)
    want, rc = o_voxel(o_submap(clouds, poses), 0.1)
    assert rc == 0 and np.array_equal(sub, want) and g.n_target == want.shape[0]
    # the submap is the target now (it never left the device): same registration as with the cloud set from the host
    src = np.ascontiguousarray(frames[5][1])
    g.set_input_source(src)
    r1 = g.align()
    g2 = gorio.FastAPDGICP(0)
    g2.set_params(max_correspondence_distance=2.0, transformation_epsilon=0.1)
    g2.set_input_target(want); g2.set_input_source(src)
    r2 = g2.align()
    assert np.array_equal(r1["T64"], r2["T64"]) and r1["converged"] == r2["converged"]
    g.close(); g2.close()
