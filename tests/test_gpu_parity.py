"""GPU parity tests: the CUDA library (through its C-ABI) against the CPU oracle
on identical seeded inputs. Bars: bit-exact neighbour indices and
correspondences; covariances <= 1e-9 abs where the normal is well conditioned;
H / b / err within 1e-10 relative (fp64 Mahalanobis storage) or 1e-6 (fp32
storage; north_star's bar is 1e-5); pose within 1e-6 m / 1e-6 rad; equal
convergence flag and iteration counts."""
import ctypes

import numpy as np
import pytest

from helpers import DEPLOYED, moved_copy, pose_err, rel
from oracle_binding import Oracle

pytestmark = pytest.mark.gpu

REGS = {"NONE": 0, "MIN_EIG": 1, "NORMALIZED_MIN_EIG": 2, "PLANE": 3, "FROBENIUS": 4}


def make(gorio, src, tgt, **kw):
    g = gorio.FastAPDGICP(0)
    o = Oracle(search=1)
    g.set_params(**kw)
    o.set_params(**kw)
    for r in (g, o):
        r.set_input_target(tgt)
        r.set_input_source(src)
    return g, o


@pytest.fixture(scope="module")
def c1(synth):
    return synth.scan_pair(1000, 1000)


@pytest.fixture(scope="module")
def c2(synth):
    return synth.submap_pair(2000)


@pytest.fixture(scope="module")
def c2_small(synth):
    return synth.submap_pair(2001, n_source=800, n_frames=6, n_per_frame=1500)


# ---------------------------------------------------------------- kNN ----
@pytest.mark.parametrize("k", [1, 5, 20, 32, 50])
def test_knn_indices_bit_exact(gorio, c1, k):
    src, tgt, _ = c1
    g, o = make(gorio, src, tgt, k_correspondences=k)
    o.get_source_covariances(); o.get_target_covariances()
    for which in (0, 1):
        assert np.array_equal(g.get_neighbors(which), o.get_neighbors(which))


def test_knn_indices_bit_exact_submap(gorio, c2):
    src, tgt, _ = c2
    g, o = make(gorio, src, tgt)
    o.get_target_covariances()
    assert np.array_equal(g.get_neighbors(1), o.get_neighbors(1))


def test_knn_kernels_match(gorio, c2_small, monkeypatch):
    """APD_KNN_MODE selects the warp-per-point or the thread-per-point kNN kernel (auto: by cloud size and k):
    same neighbours, same covariances"""
    src, tgt, _ = c2_small
    hs = []
    for mode in ("thread", "warp"):
        monkeypatch.setenv("APD_KNN_MODE", mode)
        hs.append(gorio.FastAPDGICP(0))
    monkeypatch.delenv("APD_KNN_MODE")
    for r in hs:
        r.set_input_target(tgt); r.set_input_source(src)
    for r in hs[1:]:
        for which in (0, 1):
            assert np.array_equal(hs[0].get_neighbors(which), r.get_neighbors(which))
        assert np.array_equal(hs[0].get_target_covariances(), r.get_target_covariances())


@pytest.mark.parametrize("cpp", ["0.5", "2", "16"])
def test_knn_any_cell_size(gorio, c2_small, monkeypatch, cpp):
    """the grid resolution only changes how many shells a search walks, never the result"""
    src, tgt, _ = c2_small
    monkeypatch.setenv("APD_CELLS_PER_POINT", cpp)
    g, o = make(gorio, src, tgt, regularization=0)
    o.get_target_covariances(); o.get_source_covariances()
    for which in (0, 1):
        assert np.array_equal(g.get_neighbors(which), o.get_neighbors(which))
    assert np.array_equal(g.get_target_covariances(), o.get_target_covariances())


@pytest.mark.parametrize("mode", ["thread", "warp"])
def test_raw_covariance_bit_exact(gorio, c2_small, monkeypatch, mode):
    """NONE regularisation: both kNN kernels sum in the oracle's order with one rounding per operation"""
    src, tgt, _ = c2_small
    monkeypatch.setenv("APD_KNN_MODE", mode)
    g, o = make(gorio, src, tgt, regularization=0)
    assert np.array_equal(g.get_target_covariances(), o.get_target_covariances())
    assert np.array_equal(g.get_source_covariances(), o.get_source_covariances())


def test_knn_clustered_and_degenerate_geometry(gorio):
    """dense blob + far outliers + a coplanar sheet + a collinear run: ring expansion and tie handling"""
    rng = np.random.default_rng(7)
    blob = rng.normal(0, 0.05, (3000, 3))
    far = rng.uniform(-400, 400, (40, 3))
    sheet = np.stack([rng.uniform(0, 10, 500), rng.uniform(0, 10, 500), np.full(500, 3.0)], axis=1)
    line = np.stack([np.arange(100) * 0.25, np.zeros(100), np.zeros(100)], axis=1) + 20.0  # exact distance ties
    xyz = np.concatenate([blob, far, sheet, line]).astype(np.float32)
    xyz = np.unique(xyz, axis=0)
    rng.shuffle(xyz)
    cloud = np.concatenate([xyz, np.zeros((xyz.shape[0], 1), np.float32)], axis=1)
    g, o = make(gorio, cloud[:500], cloud)
    o.get_target_covariances()
    assert np.array_equal(g.get_neighbors(1), o.get_neighbors(1))


# --------------------------------------------------------- covariances ----
@pytest.mark.parametrize("reg", list(REGS))
def test_covariances(gorio, c1, reg):
    src, tgt, _ = c1
    g, o = make(gorio, src, tgt, regularization=REGS[reg])
    Cg, Co = g.get_target_covariances(), o.get_target_covariances()
    assert np.abs(Cg[:, 3, :]).max() == 0 and np.abs(Cg[:, :, 3]).max() == 0
    if reg in ("PLANE", "MIN_EIG", "NORMALIZED_MIN_EIG"):
        o2 = Oracle(search=1); o2.set_params(regularization=0); o2.set_input_target(tgt); o2.set_input_source(src)
        S = np.linalg.svd(o2.get_target_covariances()[:, :3, :3], compute_uv=False)
        ok = (S[:, 1] - S[:, 2]) / S[:, 0] > 1e-6
        assert ok.mean() > 0.99
        assert np.abs(Cg[ok] - Co[ok]).max() < 1e-9
    else:
        assert rel(Cg, Co) < 1e-12
    # symmetric storage
    assert np.abs(Cg - Cg.transpose(0, 2, 1)).max() == 0


def test_set_get_covariances_roundtrip(gorio, c1):
    src, tgt, _ = c1
    g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    rng = np.random.default_rng(3)
    A = rng.normal(size=(src.shape[0], 3, 3))
    C = np.zeros((src.shape[0], 4, 4))
    C[:, :3, :3] = A @ A.transpose(0, 2, 1) + 0.05 * np.eye(3)
    for r in (g, o):
        r.set_source_covariances(C)
    assert np.abs(g.get_source_covariances() - C).max() < 1e-15
    T = np.eye(4)
    eg, Hg, bg = g.linearize(T)
    eo, Ho, bo = o.linearize(T)
    assert abs(eg - eo) / eo < 1e-10 and rel(Hg, Ho) < 1e-10 and rel(bg, bo) < 1e-9


# ----------------------------------------------- correspondences, H, b ----
POSES = [
    ([0, 0, 0], [0, 0, 0]),
    ([0.3, -0.2, 0.05], [0.2, -0.3, 1.5]),
    ([-1.5, 2.0, 0.3], [1.0, 2.0, -8.0]),
    ([40.0, -30.0, 5.0], [0.0, 0.0, 90.0]),  # mostly outside the target's bounding box
]


@pytest.mark.parametrize("fp64", [1, 0])
@pytest.mark.parametrize("pose", range(len(POSES)))
@pytest.mark.parametrize("thr", [2.0, None])
def test_linearize_parity(gorio, synth, c1, pose, thr, fp64):
    src, tgt, _ = c1
    kw = dict(maha_fp64=fp64)
    if thr is not None:
        kw["max_correspondence_distance"] = thr
    g, o = make(gorio, src, tgt, **kw)
    t, rpy = POSES[pose]
    T = synth.make_pose(t, np.deg2rad(rpy))
    eg, Hg, bg = g.linearize(T)
    eo, Ho, bo = o.linearize(T)
    cg, sg = g.get_correspondences()
    co, so = o.get_correspondences()
    assert np.array_equal(cg, co)                      # bit-exact, including -1
    assert np.array_equal(sg[co >= 0], so[co >= 0])    # fp32 d2 bit-exact where matched
    if thr is None:
        assert (co >= 0).all()
    if (co >= 0).sum() == 0:
        assert eg == 0.0 and np.abs(Hg).max() == 0.0
        return
    tol = 1e-10 if fp64 else 1e-6
    assert rel(g.get_mahalanobis(), o.get_mahalanobis()) < (1e-9 if fp64 else 2e-7)
    assert abs(eg - eo) / abs(eo) < tol
    assert rel(Hg, Ho) < tol
    assert rel(bg, bo) < (tol * 10)
    assert np.array_equal(Hg, Hg.T)
    # compute_error at trial poses keeps the stale correspondences / Mahalanobis
    for d in ([0.02, 0, 0], [0, -0.05, 0.01]):
        T2 = synth.make_pose(d, [0, 0, 0.002]) @ T
        assert abs(g.compute_error(T2) - o.compute_error(T2)) / abs(o.compute_error(T2)) < tol
    # err-only linearize (H == nullptr, :278-280)
    assert abs(g.linearize(T, want_hb=False) - eo) / abs(eo) < tol


def test_linearize_parity_submap(gorio, synth, c2):
    src, tgt, Tgt = c2
    g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    for T in (np.eye(4), Tgt):
        eg, Hg, bg = g.linearize(T)
        eo, Ho, bo = o.linearize(T)
        assert np.array_equal(g.get_correspondences()[0], o.get_correspondences()[0])
        assert abs(eg - eo) / eo < 1e-10 and rel(Hg, Ho) < 1e-10 and rel(bg, bo) < 1e-9


@pytest.mark.parametrize("mode", ["lane", "wide"])
def test_update_correspondences_variants(gorio, synth, c2, monkeypatch, mode):
    """the two search variants of update_correspondences (8 lanes per query: small sources; one lane per query: large
    sources) against the oracle: bit-exact correspondences and squared distances, Mahalanobis within 1e-10, at poses
    near and far from the answer"""
    src, tgt, Tgt = c2
    monkeypatch.setenv("APD_CORR_MODE", mode)
    for thr in (2.0, None):
        kw = dict(maha_fp64=1, host_loop=1)
        if thr is not None:
            kw["max_correspondence_distance"] = thr
        g, o = make(gorio, src, tgt, **kw)
        # consecutive passes on one handle warm-start from the previous one (seeded searches, provably-unmatched points
        # skipped): small LM-like steps, a return to an earlier pose, and a jump far away must all stay exact
        steps = [Tgt @ synth.make_pose([0.02 * i, -0.01 * i, 0.005], [0.0, 0.001 * i, -0.002 * i]) for i in range(1, 5)]
        for T in [np.eye(4), Tgt] + steps + [Tgt, Tgt @ synth.make_pose([3.0, -2.0, 0.5], [0.02, -0.01, 0.3]), steps[0]]:
            g.update_correspondences(T); o.update_correspondences(T)
            cg, sg = g.get_correspondences()
            co, so = o.get_correspondences()
            assert np.array_equal(cg, co)
            assert np.array_equal(sg[cg >= 0], so[co >= 0])  # (sq_distances_ of rejected points is write-only scratch, :180)
            assert rel(g.get_mahalanobis(), o.get_mahalanobis()) < 1e-10


def test_kept_matches_equal_a_cold_search(gorio, synth, c2, monkeypatch):
    """the single-lane search keeps a match without searching when the bound it stored for every OTHER target point
    proves it (corr.cu): a chain of small motions — bounds carried from pass to pass, never refreshed for most points —
    must give, at every pose, exactly what a fresh handle's cold search and the oracle give"""
    src, tgt, Tgt = c2
    monkeypatch.setenv("APD_CORR_MODE", "lane")
    g, o = make(gorio, src, tgt, max_correspondence_distance=2.0, maha_fp64=1, host_loop=1)
    T = Tgt.copy()
    rng = np.random.default_rng(5)
    for i in range(14):
        step = synth.make_pose(rng.normal(0, 0.004, 3), rng.normal(0, 0.0004, 3)) if i % 5 != 4 else synth.make_pose([0.3, -0.2, 0.05], [0.0, 0.01, 0.02])
        T = T @ step
        g.update_correspondences(T); o.update_correspondences(T)
        cg, sg = g.get_correspondences()
        co, so = o.get_correspondences()
        assert np.array_equal(cg, co), i
        assert np.array_equal(sg[cg >= 0], so[co >= 0]), i
        if i in (3, 13):
            cold = gorio.FastAPDGICP(0)
            cold.set_params(max_correspondence_distance=2.0, maha_fp64=1, host_loop=1)
            cold.set_input_target(tgt); cold.set_input_source(src)
            cold.update_correspondences(T)
            cc, sc = cold.get_correspondences()
            assert np.array_equal(cg, cc) and np.array_equal(sg[cg >= 0], sc[cc >= 0])
            assert rel(g.get_mahalanobis(), cold.get_mahalanobis()) == 0.0
            cold.close()


def test_cluster_label_weight(gorio, synth, c1):
    """cl_weight = 1/N when source.normal_x == target.normal_x (:271-273)"""
    src, tgt, _ = c1
    s2 = src.copy(); s2[:, 3] = -5.0  # no label ever matches
    g, o = make(gorio, s2, tgt, **DEPLOYED, maha_fp64=1)
    g2, o2 = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    e_none, e_lab = g.linearize(np.eye(4), want_hb=False), g2.linearize(np.eye(4), want_hb=False)
    assert e_lab > e_none
    assert abs(e_none - o.linearize(np.eye(4), want_hb=False)) / e_none < 1e-10
    assert abs(e_lab - o2.linearize(np.eye(4), want_hb=False)) / e_lab < 1e-10


def test_linearize_is_bit_deterministic(gorio, c2_small):
    src, tgt, _ = c2_small
    outs = []
    for _ in range(3):
        g = gorio.FastAPDGICP(0)
        g.set_params(**DEPLOYED)
        g.set_input_target(tgt); g.set_input_source(src)
        e, H, b = g.linearize(np.eye(4))
        outs.append((e, H.copy(), b.copy(), g.align()["T64"]))
    for e, H, b, T in outs[1:]:
        assert e == outs[0][0] and np.array_equal(H, outs[0][1]) and np.array_equal(b, outs[0][2]) and np.array_equal(T, outs[0][3])


# ----------------------------------------------------------------- align ----
def _check_align(g, o, guess=None, pose_tol=1e-6):
    rg, ro = g.align(guess), o.align(guess)
    dt, dr = pose_err(rg["T64"], ro["T64"])
    assert dt < pose_tol and dr < pose_tol, (dt, dr)
    assert rg["converged"] == ro["converged"] and rg["iterations"] == ro["iterations"]
    tg, to = g.lm_trace(), o.lm_trace()
    assert tg.shape == to.shape
    assert np.array_equal(tg[:, [0, 1, 7]], to[:, [0, 1, 7]])  # same LM path (outer, inner, accept)
    assert np.abs(rg["T"].astype(np.float64) - rg["T64"]).max() < 1e-5
    assert rel(rg["H"], ro["H"]) < 1e-8
    return rg, ro


@pytest.mark.parametrize("params", [DEPLOYED, dict(max_correspondence_distance=2.0), dict()])
def test_align_parity_scan_pair(gorio, synth, params):
    for seed in (1000, 1001, 1002):
        src, tgt, _ = synth.scan_pair(seed, 1000)
        g, o = make(gorio, src, tgt, **params, maha_fp64=1)
        _check_align(g, o)


def test_align_parity_submap(gorio, c2):
    src, tgt, _ = c2
    g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    _check_align(g, o)
    g, o = make(gorio, src, tgt, max_correspondence_distance=2.0, maha_fp64=1)
    _check_align(g, o)


def test_align_fp32_mahalanobis_within_north_star_tolerance(gorio, synth, c2):
    src, tgt, _ = c2
    g, o = make(gorio, src, tgt, max_correspondence_distance=2.0, maha_fp64=0)
    rg, ro = g.align(), o.align()
    dt, dr = pose_err(rg["T64"], ro["T64"])
    assert dt < 1e-6 and dr < 1e-6
    assert rel(rg["H"], ro["H"]) < 1e-5


@pytest.mark.parametrize("optimizer", [1, 0])
def test_host_loop_matches_oracle_and_device_loop(gorio, c2_small, optimizer):
    """apd_align has two drivers of the same kernels' arithmetic: the device-resident loop (lm.cu, default for
    sources <= 32768 points) and the host loop (host_loop=1; large / sharded clouds). Both must match the oracle,
    and each other far below the pose tolerance (they differ only in summation order and libm vs CUDA sin/cos)."""
    src, tgt, _ = c2_small
    kw = dict(max_correspondence_distance=2.0, optimizer=optimizer, max_iterations=8 if optimizer == 0 else 64, maha_fp64=1)
    gh, o = make(gorio, src, tgt, **kw, host_loop=1)
    rh, _ = _check_align(gh, o)
    gd, o = make(gorio, src, tgt, **kw, host_loop=0)
    rd, _ = _check_align(gd, o)
    assert gd.kernel_ms()["lm"][1] == 1 and gh.kernel_ms()["lm"][1] == 0
    dt, dr = pose_err(rh["T64"], rd["T64"])
    assert dt < 1e-7 and dr < 1e-7, (dt, dr)
    assert np.array_equal(gh.get_correspondences()[0], gd.get_correspondences()[0])
    assert rel(gh.get_mahalanobis(), gd.get_mahalanobis()) < 1e-12
    # compute_error after a device-loop align reads the correspondences the loop left behind
    assert abs(gd.compute_error(rd["T64"]) - gh.compute_error(rd["T64"])) / gh.compute_error(rd["T64"]) < 1e-9


@pytest.mark.parametrize("cluster", [1, 2, 4, 16])
def test_device_loop_cluster_sizes(gorio, c2_small, monkeypatch, cluster):
    """the number of CTAs per registration only changes the summation order"""
    src, tgt, _ = c2_small
    g8, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)  # default: 8
    r4, _ = _check_align(g8, o)
    monkeypatch.setenv("APD_LM_CLUSTER", str(cluster))
    g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    r, _ = _check_align(g, o)
    dt, dr = pose_err(r["T64"], r4["T64"])
    assert dt < 1e-7 and dr < 1e-7, (dt, dr)


# ------------------------------------------------- FastGICP behind the same kernels ----
GICP = dict(variant=1)  # APD_VARIANT_GICP: reference fast_gicp_impl.hpp (registrations.cpp:28-37 "FAST_GICP")


@pytest.mark.parametrize("host_loop", [0, 1])
@pytest.mark.parametrize("thr", [2.0, None])
def test_gicp_variant_align_parity(gorio, synth, c1, c2_small, host_loop, thr):
    """FastGICP = the combined covariance without the radar noise term (fast_gicp_impl.hpp:157) and unit weights
    (:205); everything else is the APDGICP path. Same bars as APDGICP against the oracle's FastGICP."""
    kw = dict(**GICP, maha_fp64=1, host_loop=host_loop)
    if thr is not None:
        kw["max_correspondence_distance"] = thr
    for src, tgt, _ in (c1, c2_small):
        g, o = make(gorio, src, tgt, **kw)
        rg, ro = _check_align(g, o)
        assert np.array_equal(g.get_correspondences()[0], o.get_correspondences()[0])
        assert rel(g.get_mahalanobis(), o.get_mahalanobis()) < 1e-9
        # and it is a different cost from APDGICP's
        ga, _ = make(gorio, src, tgt, **{**kw, "variant": 0})
        assert not np.array_equal(ga.align()["T64"], rg["T64"])


def test_gicp_variant_linearize_and_switching(gorio, synth, c1):
    src, tgt, _ = c1
    g, o = make(gorio, src, tgt, **GICP, max_correspondence_distance=2.0, maha_fp64=1)
    t, rpy = POSES[1]
    T = synth.make_pose(t, np.deg2rad(rpy))
    eg, Hg, bg = g.linearize(T)
    eo, Ho, bo = o.linearize(T)
    assert rel(Hg, Ho) < 1e-10 and rel(bg, bo) < 1e-9 and abs(eg - eo) / eo < 1e-10
    assert abs(g.compute_error(T) - o.compute_error(T)) / eo < 1e-10
    # the radar noise parameters are not read
    g.set_params(dist_var=5.0, azimuth_var=3.0, elevation_var=4.0)
    assert g.linearize(T)[0] == eg
    # switching the variant on a live handle: the weights and the stored Mahalanobis matrices follow
    for r in (g, o):
        r.set_params(variant=0, dist_var=0.86, azimuth_var=0.5, elevation_var=1.0)
    ea, eoa = g.linearize(T)[0], o.linearize(T)[0]
    assert abs(ea - eoa) / eoa < 1e-10 and ea != eg
    for r in (g, o):
        r.set_params(variant=1)
    assert g.linearize(T)[0] == eg
    with pytest.raises(gorio.ApdError):
        g.set_params(variant=7)


def test_gicp_variant_in_a_pool(gorio, synth, monkeypatch):
    monkeypatch.setenv("APD_LM_CLUSTER", "4")
    monkeypatch.setenv("APD_LM_MINB", "2")
    pairs = [synth.submap_pair(3100 + i, n_source=600, n_frames=4, n_per_frame=1000)[:2] + (None,) for i in range(4)]
    b = gorio.Batch(0, n_workers=4, max_correspondence_distance=2.0, transformation_epsilon=0.1, variant=1)
    res = b.align(b.prepare(pairs))
    b.close()
    for (s, t, _), r in zip(pairs, res):
        g = gorio.FastAPDGICP(0)
        g.set_params(max_correspondence_distance=2.0, transformation_epsilon=0.1, variant=1)
        g.set_input_target(t); g.set_input_source(s)
        assert r["status"] == 0 and np.array_equal(r["T"], g.align()["T"])


# ------------------------------------------------- FastVGICP (vgicp.cu) ----
VOXEL_SEARCH = {"DIRECT27": 0, "DIRECT7": 1, "DIRECT1": 2}


@pytest.mark.parametrize("method,mode,res", [("DIRECT1", 0, 1.0), ("DIRECT7", 0, 1.5), ("DIRECT27", 1, 2.0), ("DIRECT7", 2, 1.5)])
def test_vgicp_voxelmap_correspondences_and_sums(gorio, synth, c1, c2_small, method, mode, res):
    """fast_gicp::FastVGICP (reference impl/fast_vgicp_impl.hpp, fast_vgicp_voxel.hpp) against the oracle: the voxel map
    (coordinates and counts exact), the voxel correspondence table (exact), Mahalanobis / H / b / err / stale compute_error"""
    kw = dict(variant=2, voxel_resolution=res, voxel_search=VOXEL_SEARCH[method], voxel_mode=mode)
    for src, tgt, Tgt in (c1, c2_small):
        g, o = make(gorio, src, tgt, **kw)
        T = Tgt @ synth.make_pose([0.05, -0.03, 0.01], [0.0, 0.002, 0.01])
        eg, Hg, bg = g.linearize(T)
        eo, Ho, bo = o.linearize(T)
        cg, ng, mg, Cg = g.vgicp_voxels()
        co, no, mo, Co = o.vgicp_voxels()
        assert np.array_equal(cg, co) and np.array_equal(ng, no) and ng.sum() == tgt.shape[0]
        assert rel(mg, mo) < (1e-9 if mode == 2 else 1e-13) and rel(Cg, Co) < (1e-8 if mode == 2 else 1e-12)
        vg, Mg = g.vgicp_correspondences()
        vo, Mo = o.vgicp_correspondences()
        assert np.array_equal(vg, vo) and (vg >= 0).sum() > 50
        assert rel(Mg, Mo) < 1e-9
        assert abs(eg - eo) / eo < 1e-10 and rel(Hg, Ho) < 1e-10 and rel(bg, bo) < 1e-9
        T2 = T @ synth.make_pose([0.02, 0.01, -0.01], [0.001, 0.0, -0.004])
        assert abs(g.compute_error(T2) - o.compute_error(T2)) / eo < 1e-10
        # a second pass at another pose: the table is rewritten, the voxel map is kept
        e2g, e2o = g.linearize(T2, want_hb=False), o.linearize(T2, want_hb=False)
        assert abs(e2g - e2o) / e2o < 1e-10 and np.array_equal(g.vgicp_correspondences()[0], o.vgicp_correspondences()[0])
        g.close()


@pytest.mark.parametrize("optimizer", [1, 0])
def test_vgicp_align_parity(gorio, synth, c1, c2_small, optimizer):
    """the whole FastVGICP registration (voxel map rebuilt per align, LM or GN on the host): same pose, converged flag,
    iteration count and LM trace as the oracle"""
    for (src, tgt, _), kw in ((c1, dict(voxel_resolution=2.0, voxel_search=0)), (c2_small, dict(voxel_resolution=1.0, voxel_search=1))):
        g, o = make(gorio, src, tgt, variant=2, optimizer=optimizer, transformation_epsilon=0.01, **kw)
        rg, ro = _check_align(g, o)
        assert np.array_equal(g.vgicp_correspondences()[0], o.vgicp_correspondences()[0])
        tg, to = g.lm_trace(), o.lm_trace()
        assert tg.shape == to.shape and np.array_equal(tg[:, [0, 1, 7]], to[:, [0, 1, 7]])
        # swapping rebuilds the voxel map over the new target (fast_vgicp_impl.hpp:46-54)
        g.swap_source_and_target(); o.swap_source_and_target()
        _check_align(g, o)
        assert np.array_equal(g.vgicp_voxels()[0], o.vgicp_voxels()[0])
        g.close()


def test_vgicp_parameters_and_switching(gorio, synth, c1):
    src, tgt, Tgt = c1
    g, o = make(gorio, src, tgt, variant=2, voxel_resolution=1.5, voxel_search=1)
    e0 = g.linearize(Tgt, want_hb=False)
    for r in (g, o):
        r.set_params(voxel_resolution=3.0)  # a new resolution rebuilds the map
    e1 = g.linearize(Tgt, want_hb=False)
    # (the oracle keeps its map until the target changes, as the reference does; a fresh oracle gives the comparison)
    o2 = make(gorio, src, tgt, variant=2, voxel_resolution=3.0, voxel_search=1)[1]
    assert e1 != e0 and abs(e1 - o2.linearize(Tgt, want_hb=False)) / e1 < 1e-10
    for bad in (dict(voxel_resolution=0.0), dict(voxel_search=3), dict(voxel_mode=5)):
        with pytest.raises(gorio.ApdError):
            g.set_params(**bad)
    # a resolution whose voxel grid over the target's extent would not fit 32-bit keys is refused, not wrapped
    g.set_params(voxel_resolution=1e-8)
    with pytest.raises(gorio.ApdError):
        g.linearize(Tgt, want_hb=False)
    # back to APDGICP on the live handle
    g.set_params(variant=0, voxel_resolution=1.0, voxel_search=2, voxel_mode=0)
    oa = make(gorio, src, tgt)[1]
    assert abs(g.linearize(Tgt, want_hb=False) - oa.linearize(Tgt, want_hb=False)) < 1e-6 * abs(e0)  # (fp32 Mahalanobis storage)
    g.close()


def test_vgicp_in_a_pool(gorio, synth):
    """a pool whose handles run FastVGICP (host-driven loop inside the workers) gives what lone handles give"""
    pairs = [synth.submap_pair(3200 + i, n_source=600, n_frames=4, n_per_frame=1000)[:2] + (None,) for i in range(5)]
    kw = dict(variant=2, voxel_resolution=1.5, voxel_search=1, transformation_epsilon=0.01)
    b = gorio.Batch(0, n_workers=3, **kw)
    res = b.align(b.prepare(pairs))
    b.close()
    for (s, t, _), r in zip(pairs, res):
        g = gorio.FastAPDGICP(0)
        g.set_params(**kw)
        g.set_input_target(t); g.set_input_source(s)
        ra = g.align()
        assert r["status"] == 0 and np.array_equal(r["T"], ra["T"]) and r["iterations"] == ra["iterations"]
        g.close()


def _handle(gorio, monkeypatch, lazy, src, tgt, **kw):
    monkeypatch.setenv("APD_LAZY_TARGET_COV", lazy)
    g = gorio.FastAPDGICP(0)
    g.set_params(**kw)
    g.set_input_target(tgt)
    g.set_input_source(src)
    return g


@pytest.mark.parametrize("reg", ["PLANE", "MIN_EIG", "FROBENIUS"])
def test_target_covariances_on_demand_are_the_eager_ones(gorio, c2, monkeypatch, reg):
    """The device loop computes the covariance of a target point the first time a source point matches it
    (APD_LAZY_TARGET_COV, default for scan-vs-submap sizes) instead of all 60 k up front: same kNN search, same
    arithmetic -> the SAME BITS in every output, and one kNN launch less."""
    src, tgt, _ = c2
    kw = dict(**DEPLOYED, maha_fp64=1, regularization=REGS[reg])
    ge = _handle(gorio, monkeypatch, "0", src, tgt, **kw)
    gl = _handle(gorio, monkeypatch, "1", src, tgt, **kw)
    re_, rl = ge.align(), gl.align()
    assert np.array_equal(re_["T64"], rl["T64"]) and np.array_equal(re_["H"], rl["H"])
    assert re_["iterations"] == rl["iterations"] and re_["converged"] == rl["converged"]
    assert np.array_equal(ge.lm_trace(), gl.lm_trace())
    assert np.array_equal(ge.get_correspondences()[0], gl.get_correspondences()[0])
    assert np.array_equal(ge.get_mahalanobis(), gl.get_mahalanobis())
    # the on-demand handle took the fused kernel: grids, source covariances and the loop in ONE launch; the eager one ran the
    # separate kernels (host-sized grids) — and still every output above has the same bits
    assert gl.kernel_ms()["knn_cov"][1] == 0 and gl.kernel_ms()["grid"][1] == 0 and gl.kernel_ms()["lm"][1] == 1
    # asking for all of them completes the cloud; the ones the loop made are not distinguishable
    assert np.array_equal(ge.get_target_covariances(), gl.get_target_covariances())
    if reg == "PLANE":
        o = Oracle(search=1)
        o.set_params(**kw)
        o.set_input_target(tgt)
        o.set_input_source(src)
        _check_align(gl, o)


def test_target_covariances_on_demand_persist_across_sources(gorio, synth, c2_small, monkeypatch):
    """a keyframe target serves many scans: what one align computed the next one finds; a parameter change in between
    does not touch covariances that exist (reference :152-154: target_covs_ is only computed when empty) and the rest
    of the cloud is completed with the parameters the first align saw"""
    src, tgt, _ = c2_small
    kw = dict(max_correspondence_distance=2.0, maha_fp64=1)
    ge = _handle(gorio, monkeypatch, "0", src, tgt, **kw)
    gl = _handle(gorio, monkeypatch, "1", src, tgt, **kw)
    o = Oracle(search=1)
    o.set_params(**kw)
    o.set_input_target(tgt)
    o.set_input_source(src)
    T = np.eye(4)
    T[:3, 3] = [0.3, -0.2, 0.05]
    sources = [src, src[::2].copy(), moved_copy(synth, src, T), src[1::3].copy()]
    for i, s in enumerate(sources):
        if i == 2:
            for r in (ge, gl, o):
                r.set_params(**kw, k_correspondences=10, regularization=REGS["MIN_EIG"])
        for r in (ge, gl, o):
            r.set_input_source(s)
        re_, rl = ge.align(), gl.align()
        assert np.array_equal(re_["T64"], rl["T64"]), i
        assert np.array_equal(ge.get_mahalanobis(), gl.get_mahalanobis()), i
        ro = o.align()
        dt, dr = pose_err(rl["T64"], ro["T64"])
        assert dt < 1e-6 and dr < 1e-6 and rl["iterations"] == ro["iterations"], (i, dt, dr)
    assert np.array_equal(ge.get_target_covariances(), gl.get_target_covariances())
    # swap: the partly covered target becomes the source and is completed
    gl2 = _handle(gorio, monkeypatch, "1", src, tgt, **kw)
    ge2 = _handle(gorio, monkeypatch, "0", src, tgt, **kw)
    for g in (gl2, ge2):
        g.align()
        g.swap_source_and_target()
    assert np.array_equal(ge2.align()["T64"], gl2.align()["T64"])


def test_loop_kernel_register_builds_agree(gorio, c2_small, monkeypatch):
    """lm_kernel exists in two register allocations (128 registers: lone handles; 64 registers, two CTAs per SM: pool
    workers). Same source, same arithmetic: same LM path, poses equal far below the tolerance."""
    src, tgt, _ = c2_small
    res = []
    for minb in ("1", "2"):
        monkeypatch.setenv("APD_LM_MINB", minb)
        g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
        r, _ = _check_align(g, o)
        res.append((r, g.lm_trace()))
    assert np.array_equal(res[0][1][:, [0, 1, 7]], res[1][1][:, [0, 1, 7]])
    dt, dr = pose_err(res[0][0]["T64"], res[1][0]["T64"])
    assert dt < 1e-10 and dr < 1e-10, (dt, dr)


def test_device_loop_tiny_and_ragged_sources(gorio, synth, c2_small):
    """fewer source points than CTAs x lanes, and counts that do not divide by the cluster size"""
    src, tgt, _ = c2_small
    for n in (21, 63, 257, 799):
        s = src[:n].copy()
        g, o = make(gorio, s, tgt, max_correspondence_distance=2.0, maha_fp64=1)
        _check_align(g, o)


def test_lm_failure_is_reported(gorio, c2_small):
    """lm_max_iterations exhausted -> converged stays false (lsq_registration_impl.hpp:71-74), same as the oracle"""
    src, tgt, _ = c2_small
    for host_loop in (0, 1):
        g, o = make(gorio, src, tgt, max_correspondence_distance=2.0, lm_max_iterations=1, lm_init_lambda_factor=1e-30,
                    max_iterations=3, maha_fp64=1, host_loop=host_loop)
        _check_align(g, o)


def test_align_gauss_newton(gorio, c2_small):
    src, tgt, _ = c2_small
    g, o = make(gorio, src, tgt, max_correspondence_distance=2.0, optimizer=0, max_iterations=8, maha_fp64=1)
    _check_align(g, o)


def test_align_with_guess_and_aligned_output(gorio, synth, c2_small):
    src, tgt, Tgt = c2_small
    g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    guess = (Tgt @ synth.make_pose([0.1, 0.05, 0.0], [0, 0, 0.01])).astype(np.float32)
    _check_align(g, o, guess)
    rg = g.align(guess, want_aligned=True)
    ro = o.align(guess, want_aligned=True)
    assert np.array_equal(rg["aligned"], ro["aligned"]) or np.abs(rg["aligned"] - ro["aligned"]).max() < 1e-4


def test_recovers_known_transform(gorio, synth, c2_small):
    """reference gicp_test.cpp:148-166 bar: < 0.05 m, < 1 deg, converged"""
    _, tgt, _ = c2_small
    D = synth.make_pose([0.3, -0.2, 0.05], np.deg2rad([0.3, -0.4, 2.0]))
    src = moved_copy(synth, tgt, D)[::3].copy()
    g = gorio.FastAPDGICP(0)
    g.set_params(max_correspondence_distance=2.0)
    g.set_input_target(tgt); g.set_input_source(src)
    r = g.align()
    dt, dr = pose_err(r["T64"], D)
    assert r["converged"] and dt < 0.05 and dr < np.deg2rad(1.0)
    score, n_in, n_inl = g.fitness()
    assert score < 1e-3 and n_in == src.shape[0] and n_inl == src.shape[0]
    # identity on identical clouds: zero motion, ~zero error
    g.set_input_source(tgt[::3].copy())
    r = g.align()
    dt, dr = pose_err(r["T64"], np.eye(4))
    assert r["converged"] and dt < 1e-6 and dr < 1e-6


def test_swap_source_and_target(gorio, synth, c2_small):
    """gicp_test.cpp:157-200: swap then align gives the inverse transform, covariances travel with the clouds"""
    src, tgt, _ = c2_small
    g, o = make(gorio, src, tgt, max_correspondence_distance=2.0, maha_fp64=1)
    _check_align(g, o)
    cs = g.get_source_covariances()
    g.swap_source_and_target(); o.swap_source_and_target()
    assert np.array_equal(g.get_target_covariances(), cs)
    with pytest.raises(gorio.ApdError):
        g.compute_error(np.eye(4))  # correspondences_ were cleared by the swap (:96)
    _check_align(g, o)
    # swap, then set a new source / a new target
    g.swap_source_and_target(); o.swap_source_and_target()
    s2 = src[::2].copy()
    g.set_input_source(s2); o.set_input_source(s2)
    _check_align(g, o)


def test_clear_and_cache_key(gorio, c1):
    src, tgt, _ = c1
    g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    g.set_input_target(tgt, key=42)
    g.align()
    l0 = g.launch_count()
    g.set_input_target(tgt.copy(), key=42)  # same key, same content: early-out (:128), nothing is rebuilt
    assert g.launch_count() == l0
    _check_align(g, o)
    # a stale key: the same number now names other points — it must not keep the old cloud
    t2 = tgt[::-1].copy()[: tgt.shape[0] - 7]
    g.set_input_target(t2, key=42); o.set_input_target(t2)
    _check_align(g, o)
    t3 = t2.copy(); t3[:, 0] += 0.25  # same size, same address pattern, other content
    g.set_input_target(t3, key=42); o.set_input_target(t3)
    _check_align(g, o)
    g.set_input_target(tgt, key=43); o.set_input_target(tgt)
    _check_align(g, o)
    g.clear_source()
    with pytest.raises(gorio.ApdError) as e:
        g.align()
    assert e.value.code == 1
    g.set_input_source(src)
    _check_align(g, o)


def test_fused_kernel_equals_the_separate_kernels(gorio, synth, c2_small, monkeypatch):
    """APD_FUSED=0 builds grids and source covariances with the GPU-wide kernels (boxes reduced on the device, grids sized on
    the host); the default does all of it inside the loop kernel's launch (device-sized grids: grid_desc.hpp is one function
    for both). Same bits everywhere, 1 launch against 11; keyframe reuse (target kept) and odd sizes included."""
    src, tgt, _ = c2_small
    monkeypatch.setenv("APD_LAZY_TARGET_COV", "1")
    kw = dict(**DEPLOYED, maha_fp64=1)
    monkeypatch.setenv("APD_FUSED", "0")
    gs = gorio.FastAPDGICP(0); gs.set_params(**kw)
    monkeypatch.setenv("APD_FUSED", "1")
    gf = gorio.FastAPDGICP(0); gf.set_params(**kw)
    for cut_s, cut_t in ((0, 0), (7, 13), (301, 999)):
        s_, t_ = src[cut_s:].copy(), tgt[cut_t:].copy()
        for g in (gs, gf):
            g.clear_target(); g.clear_source()
            g.set_input_target(t_); g.set_input_source(s_)
        l0 = gf.launch_count()
        rs, rf = gs.align(), gf.align()
        assert gf.launch_count() - l0 == 1 and gs.kernel_ms()["grid"][1] > 0
        assert np.array_equal(rs["T64"], rf["T64"]) and np.array_equal(rs["H"], rf["H"]) and rs["iterations"] == rf["iterations"]
        assert np.array_equal(gs.lm_trace(), gf.lm_trace())
        assert np.array_equal(gs.get_correspondences()[0], gf.get_correspondences()[0])
        assert np.array_equal(gs.get_source_covariances(), gf.get_source_covariances())
        assert np.array_equal(gs.get_neighbors(0), gf.get_neighbors(0)) and np.array_equal(gs.get_neighbors(1), gf.get_neighbors(1))
        # the next scan against the SAME target: the fused kernel prepares the source only (the target's grid and the
        # covariances it has met so far stay)
        s2 = s_[::2].copy()
        for g in (gs, gf):
            g.set_input_source(s2)
        l0 = gf.launch_count()
        rs, rf = gs.align(), gf.align()
        assert gf.launch_count() - l0 == 1 and np.array_equal(rs["T64"], rf["T64"]) and np.array_equal(gs.lm_trace(), gf.lm_trace())
        assert np.array_equal(gs.get_target_covariances(), gf.get_target_covariances())
        assert gf.fitness() == gs.fitness()


def test_streamed_frames_with_recycled_addresses(gorio, synth):
    """a sequence that copies each frame and lets it go reuses host addresses; keys derived from addresses then repeat with
    different content. The library must register the frames it was given (it silently kept the previous one)."""
    frames = [synth.scan_pair(1100 + i, 600) for i in range(4)]
    g = gorio.FastAPDGICP(0)
    g.set_params(**DEPLOYED, maha_fp64=1)
    o = Oracle(search=1)
    o.set_params(**DEPLOYED, maha_fp64=1)
    buf_s, buf_t = np.empty((600, 4), np.float32), np.empty((600, 4), np.float32)
    for s, t, _ in frames:
        n_s, n_t = s.shape[0], t.shape[0]
        assert n_s <= 600 and n_t <= 600
        buf_s[:n_s], buf_t[:n_t] = s, t  # the SAME host buffers every frame: address-derived keys collide by construction
        g.set_input_target(buf_t[:n_t], key=buf_t.ctypes.data); o.set_input_target(t)
        g.set_input_source(buf_s[:n_s], key=buf_s.ctypes.data); o.set_input_source(s)
        _check_align(g, o)


def test_promotion_of_the_source_to_target_reuses_its_covariances(gorio, synth, c2_small):
    """scan_matching_odometry_nodelet.cpp:587-588: the registered frame becomes the keyframe. Same cache key as the
    source -> grid and covariances are copied, not rebuilt; results are those of a fresh handle."""
    src, tgt, _ = c2_small
    nxt = src[::2].copy()
    g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    g.set_input_source(src, key=7)
    _check_align(g, o)
    k0 = g.kernel_ms()["knn_cov"][1]
    g.set_input_target(src, key=7); o.set_input_target(src)  # promotion
    assert g.kernel_ms()["knn_cov"][1] == k0
    assert np.array_equal(g.get_target_covariances(), g.get_source_covariances())
    g.set_input_source(nxt, key=8); o.set_input_source(nxt)
    _check_align(g, o)
    k1 = g.kernel_ms()["knn_cov"][1]
    fresh, _ = make(gorio, nxt, src, **DEPLOYED, maha_fp64=1)
    assert np.array_equal(fresh.align()["T64"], g.align()["T64"])
    # only the new source needed covariances — and with the target's all valid the loop kernel computed them itself (fused
    # prologue, source only): no separate kNN launch at all
    assert k1 - k0 == 0 and fresh.kernel_ms()["knn_cov"][1] == 4


def test_pcl_xyzinormal_layout(gorio, synth, c1):
    src, tgt, _ = c1
    g = gorio.FastAPDGICP(0)
    g.set_params(**DEPLOYED, maha_fp64=1)
    g.set_input_target(synth.to_pcl_xyzinormal(tgt)); g.set_input_source(synth.to_pcl_xyzinormal(src))
    g2, _ = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    assert np.array_equal(g.align()["T64"], g2.align()["T64"])
    # odd point counts (the staging loop handles two points per turn) and a layout that takes the general gather
    s3, t3 = src[:777].copy(), tgt[:999].copy()
    g.set_input_target(synth.to_pcl_xyzinormal(t3)); g.set_input_source(synth.to_pcl_xyzinormal(s3))
    g2.set_input_target(t3); g2.set_input_source(s3)
    r3 = g2.align()["T64"]
    assert np.array_equal(g.align()["T64"], r3)
    wide = np.zeros((999, 6), np.float32); wide[:, 1:5] = t3  # stride 24, xyz at 4, label at 16
    g._call("set_target", wide.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(999), ctypes.c_int32(24), ctypes.c_int32(4),
            ctypes.c_int32(16), ctypes.c_uint64(0))
    g.n_target = 999
    assert np.array_equal(g.align()["T64"], r3)


def test_error_codes(gorio, c1):
    src, tgt, _ = c1
    g = gorio.FastAPDGICP(0)
    g.set_input_target(tgt); g.set_input_source(src[:10])  # fewer than k points
    with pytest.raises(gorio.ApdError) as e:
        g.align()
    assert e.value.code == 3
    g.set_params(k_correspondences=129)
    g.set_input_source(src)
    with pytest.raises(gorio.ApdError) as e:
        g.align()
    assert e.value.code == 4
    g2 = gorio.FastAPDGICP(0)
    with pytest.raises(gorio.ApdError):
        g2.align()  # no clouds
    g2.set_input_source(np.zeros((0, 4), np.float32))
    g2.set_input_target(tgt)
    with pytest.raises(gorio.ApdError):
        g2.align()  # empty source


# --------------------------------------------------------------- fitness ----
def test_fitness_parity(gorio, synth, c2_small):
    """pcl getFitnessScore(max_range) of the final pose, and — with an explicit pose and a finite max_range —
    InformationMatrixCalculator::calc_fitness_score(cloud1 = target, cloud2 = source, relpose, max_range)
    (information_matrix_calculator.cpp:55-86): the same mean of 1-NN squared distances within max_range."""
    src, tgt, Tgt = c2_small
    g, o = make(gorio, src, tgt, **DEPLOYED, maha_fp64=1)
    for T in (None, Tgt.astype(np.float32)):
        if T is None:
            g.align(); o.align()
        for max_range in (np.finfo(np.float64).max, 1.0):
            sg, ng, ig = g.fitness(T, max_range)
            so, no, io = o.fitness(T, max_range)
            assert (ng, ig) == (no, io)
            assert abs(sg - so) / so < 1e-12
    s, n, _ = g.fitness(Tgt.astype(np.float32), 0.0)
    assert n == 0 and s == np.finfo(np.float64).max  # DBL_MAX when nothing is in range


# ------------------------------------------------------- batch and sizes ----
def test_align_batch_matches_sequential(gorio, synth, monkeypatch):
    monkeypatch.setenv("APD_LM_CLUSTER", "4")  # pool workers default to 4 CTAs per registration, lone handles to 8: same summation order for the bit comparison
    monkeypatch.setenv("APD_LM_MINB", "2")     # ... and to the 64-register build of the loop kernel, lone handles to the 128-register one
    pairs = []
    for seed in range(3000, 3006):
        s, t, _ = synth.submap_pair(seed, n_source=600, n_frames=4, n_per_frame=1000)
        pairs.append((s, t, None))
    p = gorio.ApdParams()
    gorio.load().apd_default_params(ctypes.byref(p))
    p.max_correspondence_distance = 2.0
    p.transformation_epsilon = 0.1
    res = gorio.align_batch(pairs, p, n_streams=3)
    for (s, t, _), r in zip(pairs, res):
        g = gorio.FastAPDGICP(0)
        g.set_params(max_correspondence_distance=2.0, transformation_epsilon=0.1)
        g.set_input_target(t); g.set_input_source(s)
        ra = g.align()
        assert r["status"] == 0
        assert np.array_equal(r["T"], ra["T"]) and r["converged"] == ra["converged"] and r["iterations"] == ra["iterations"]
        assert abs(r["fitness"] - g.fitness()[0]) < 1e-12
        assert r["n_inliers"] == g.fitness(None, np.finfo(np.float64).max, 0.25)[2]
    # a persistent context gives the same answers, call after call, for host and for HBM-resident clouds
    import torch
    b = gorio.Batch(0, n_workers=4, max_correspondence_distance=2.0, transformation_epsilon=0.1)
    dev = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for s, t, _ in pairs for x in (s, t)]
    dpairs = [((dev[2 * i].data_ptr(), pairs[i][0].shape[0]), (dev[2 * i + 1].data_ptr(), pairs[i][1].shape[0]), None) for i in range(len(pairs))]
    # ... and for host clouds in page-locked memory, which the pool copies from directly (no staging pass)
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for s, t, _ in pairs for x in (s, t)]
    ppairs = [(pin[2 * i].numpy(), pin[2 * i + 1].numpy(), None) for i in range(len(pairs))]
    for prepared in (b.prepare(pairs), b.prepare(dpairs), b.prepare(ppairs)):
        for _ in range(2):
            res2 = b.align(prepared)
            for r, r2 in zip(res, res2):
                assert r2["status"] == 0 and np.array_equal(r["T"], r2["T"]) and r["fitness"] == r2["fitness"]
                assert (r["converged"], r["iterations"], r["n_inliers"]) == (r2["converged"], r2["iterations"], r2["n_inliers"])
    # one launch of the loop kernel per registration, or per two (a ready registration waits up to 200 us for a partner)
    assert 3 * len(pairs) <= b.kernel_ms()["lm"][1] <= 6 * len(pairs)
    b.close()


def test_pool_two_registrations_per_launch(gorio, synth, monkeypatch):
    """a device runs at most 128 grids at a time, so a pool launches two ready registrations in one grid; who shares a launch
    with whom changes nothing: bit-identical results with one per launch, and fewer launches than registrations"""
    pairs = []
    for seed in range(3100, 3108):
        s, t, _ = synth.submap_pair(seed, n_source=600, n_frames=4, n_per_frame=1000)
        pairs.append((s, t, None))
    out = {}
    for hold in ("0", "50000"):
        monkeypatch.setenv("APD_PAIR_HOLD_US", hold)
        b = gorio.Batch(0, n_workers=4, max_correspondence_distance=2.0, transformation_epsilon=0.1)
        prepared = b.prepare(pairs)
        res = [b.align(prepared) for _ in range(2)][-1]
        out[hold] = (res, b.kernel_ms()["lm"][1])
        b.close()
    one, two = out["0"], out["50000"]
    assert one[1] == 2 * len(pairs) and two[1] < 2 * len(pairs)
    for r1, r2 in zip(one[0], two[0]):
        assert r1["status"] == 0 and r2["status"] == 0
        assert np.array_equal(r1["T"], r2["T"]) and r1["fitness"] == r2["fitness"]
        assert (r1["converged"], r1["iterations"], r1["n_inliers"]) == (r2["converged"], r2["iterations"], r2["n_inliers"])


def test_pool_any_submap_grid_resolution(gorio, synth, monkeypatch):
    """a submap whose covariances are computed on demand gets a coarser grid (2 cells per point instead of 4: fewer, longer
    candidate lists); the resolution decides how the exact searches walk, never what they find — bit-identical poses,
    fitness and iteration counts at 1 / 2 / 4 cells per point, and equal to the eager path"""
    pairs = []
    for seed in range(3150, 3154):
        s, t, _ = synth.submap_pair(seed, n_source=600, n_frames=10, n_per_frame=1000)  # 10 000 points: above the scans' rule
        pairs.append((s, t, None))
    out = {}
    for key, env in (("1", {"APD_CELLS_PER_POINT_MID_LAZY": "1"}), ("2", {}), ("4", {"APD_CELLS_PER_POINT_MID_LAZY": "4"}),
                     ("eager", {"APD_LAZY_TARGET_COV": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        b = gorio.Batch(0, n_workers=4, max_correspondence_distance=2.0, transformation_epsilon=0.1)
        out[key] = b.align(b.prepare(pairs))
        b.close()
        for k in env:
            monkeypatch.delenv(k)
    for key in ("1", "4", "eager"):
        for r1, r2 in zip(out["2"], out[key]):
            assert r1["status"] == 0 and r2["status"] == 0
            assert np.array_equal(r1["T"], r2["T"]) and r1["fitness"] == r2["fitness"]
            assert (r1["converged"], r1["iterations"], r1["n_inliers"]) == (r2["converged"], r2["iterations"], r2["n_inliers"])


def test_pool_registrations_longer_than_the_stall_check(gorio, synth, monkeypatch):
    """256 eager registrations in flight take ~30 ms each — longer than the 20 ms after which a worker asks the stream whether
    its kernel is still alive. Two ways that check reported a healthy pair as "finished without publishing its result"
    (one pair in a thousand of bench.py's eager block): the kernel finishing between the worker's look at the result and
    the stream query, and — device clouds, whose boxes a kernel reduces first — a handle asking the stream of the PARTNER
    that had launched its previous registration. Every pair of every pass must succeed, with the same pose each pass,
    from HBM-resident and from host clouds"""
    import torch
    monkeypatch.setenv("APD_LAZY_TARGET_COV", "0")
    scenes = [synth.submap_pair(3300 + i)[:2] for i in range(4)]
    dev = [(torch.from_numpy(np.ascontiguousarray(s)).cuda(), torch.from_numpy(np.ascontiguousarray(t)).cuda()) for s, t in scenes]
    host_pairs = [(scenes[i % 4][0], scenes[i % 4][1], None) for i in range(256)]
    dev_pairs = [((dev[i % 4][0].data_ptr(), scenes[i % 4][0].shape[0]), (dev[i % 4][1].data_ptr(), scenes[i % 4][1].shape[0]), None) for i in range(256)]
    b = gorio.Batch(0, n_workers=256, max_correspondence_distance=2.0, transformation_epsilon=0.1)
    first = None
    for pairs in (dev_pairs, host_pairs):
        prepared = b.prepare(pairs)
        rep = b.repeat(prepared, 6)
        b.align(rep, with_fitness=False, parse=False)
        v = b.results_view(rep["res"])
        assert not np.any(v["status"] != 0), np.nonzero(v["status"] != 0)[0][:10]
        T = v["T"].reshape(6, 256, 16)
        for j in range(1, 6):
            assert np.array_equal(T[0], T[j])
        for i in range(4, 256):
            assert np.array_equal(T[0][i], T[0][i % 4])
        first = T[0].copy() if first is None else first
        assert np.array_equal(first, T[0])
    b.close()


def test_pool_reports_errors_per_pair(gorio, synth):
    """a pair that cannot be registered (fewer source points than k: APD_ERR_TOO_FEW, where the reference reads
    uninitialised memory) fails alone; the pool carries on with the others, call after call"""
    good = synth.submap_pair(3200, n_source=600, n_frames=4, n_per_frame=1000)
    bad_src = good[0][:7].copy()
    pairs = [(good[0], good[1], None), (bad_src, good[1], None), (good[0], good[1], None), (good[0][:300].copy(), good[1], None)]
    b = gorio.Batch(0, n_workers=3, max_correspondence_distance=2.0, transformation_epsilon=0.1)
    prepared = b.prepare(pairs)
    for _ in range(2):
        res = b.align(prepared)
        assert [r["status"] for r in res] == [0, 3, 0, 0]
        assert np.array_equal(res[0]["T"], res[2]["T"]) and res[0]["converged"]
    b.close()


def test_large_cloud_properties(gorio, synth):
    """2 M points (size-independent properties; the oracle would take minutes here):
    identical clouds -> zero cost; moved copy -> align recovers the motion;
    H from a subset + H from the rest == H from all (linearity of the reduction)."""
    src, tgt, Tgt = synth.tiled_cloud_pair(4000, 2_000_000)
    g = gorio.FastAPDGICP(0)
    g.set_params(max_correspondence_distance=2.0)
    g.set_input_target(tgt); g.set_input_source(tgt)
    e, H, b = g.linearize(np.eye(4))
    assert e == 0.0 and np.abs(b).max() == 0.0
    c, sq = g.get_correspondences()
    assert np.array_equal(c, np.arange(tgt.shape[0])) and sq.max() == 0.0
    g.set_input_source(src)
    e_all, H_all, b_all = g.linearize(Tgt)
    cs = g.get_source_covariances()
    half = src.shape[0] // 2
    parts = []
    for sl in (slice(0, half), slice(half, None)):
        g.set_input_source(src[sl].copy())
        g.set_source_covariances(cs[sl])  # a point's covariance depends on its cloud: keep the full cloud's
        parts.append(g.linearize(Tgt))
    # cl_weight = 1/N differs between the runs; compare H and b (unweighted, :289-290)
    assert rel(parts[0][1] + parts[1][1], H_all) < 1e-9
    assert rel(parts[0][2] + parts[1][2], b_all) < 1e-7
    # an exact moved copy of a subset is registered back onto the cloud
    sub = moved_copy(synth, tgt[::10].copy(), Tgt)
    g.set_input_source(sub)
    r = g.align()
    dt, dr = pose_err(r["T64"], Tgt)
    assert r["converged"] and dt < 5e-3 and dr < 1e-4, (r["converged"], r["iterations"], dt, dr)
