"""Known-answer tests against the committed fixture tests/golden/apdgicp_c1_small.npz (made by
tests/golden/make_golden.py from the CPU oracle after cross-checking it with the NumPy restatement; the
reference itself holds no fixture for FastAPDGICP and cannot be built here — SURVEY.md §0.2, §8c).

CPU (-m "not gpu"): the oracle still reproduces the fixture (guards the checker against drift).
GPU (-m gpu): the CUDA library, through the C-ABI, against the fixture with the bars of SURVEY.md §8c."""
import os

import numpy as np
import pytest

from helpers import pose_err, rel
from oracle_binding import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "apdgicp_c1_small.npz")
GOLD_GICP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fastgicp_c1_small.npz")  # variant = 1
DEPLOYED = dict(max_correspondence_distance=2.0, transformation_epsilon=0.1)
ALIGNS = {"lm_deployed": DEPLOYED, "lm_default": dict(max_correspondence_distance=2.0),
          "gn": dict(max_correspondence_distance=2.0, optimizer=0, max_iterations=6)}


@pytest.fixture(scope="module", params=[0, 1], ids=["apdgicp", "fastgicp"])
def gold(request):
    """the fixture of one variant; gold["variant"] is passed on to set_params"""
    g = dict(np.load(GOLD_GICP if request.param else GOLD))
    g["variant"] = request.param
    return g


def sym6(c4):
    return np.stack([c4[:, 0, 0], c4[:, 0, 1], c4[:, 0, 2], c4[:, 1, 1], c4[:, 1, 2], c4[:, 2, 2]], axis=1)


def trial(synth, T):
    return T @ synth.make_pose([0.01, 0.0, -0.01], [0.0, 0.001, 0.0])


def check(r, gold, synth, tol_cov, tol_lin, tol_b, exact_trace=True):
    """r: a Registration (oracle or CUDA) with the fixture's clouds set and DEPLOYED params"""
    S = np.abs(gold["cov_target"]).max()
    assert np.abs(sym6(r.get_target_covariances()) - gold["cov_target"]).max() <= tol_cov * max(1.0, S)
    assert np.abs(sym6(r.get_source_covariances()) - gold["cov_source"]).max() <= tol_cov * max(1.0, S)
    assert np.array_equal(r.get_neighbors(1), gold["nb_target"]) and np.array_equal(r.get_neighbors(0), gold["nb_source"])
    for name, T in (("I", np.eye(4)), ("P2", gold["pose2"])):
        err, H, b = r.linearize(T)
        c, sq = r.get_correspondences()
        assert np.array_equal(c, gold[f"corr_{name}"])
        assert np.array_equal(sq[c >= 0], gold[f"sqd_{name}"][c >= 0])
        assert (c >= 0).sum() > 100 and (c < 0).sum() > 0  # the fixture exercises both branches of :183
        assert rel(sym6(r.get_mahalanobis()), gold[f"maha_{name}"]) < tol_lin
        assert rel(H, gold[f"H_{name}"]) < tol_lin and rel(b, gold[f"b_{name}"]) < tol_b
        assert abs(err - gold[f"err_{name}"]) / gold[f"err_{name}"] < tol_lin
        e2 = r.compute_error(trial(synth, T))
        assert abs(e2 - gold[f"err_trial_{name}"]) / gold[f"err_trial_{name}"] < tol_lin


def check_align(r, gold, name):
    res = r.align()
    dt, dr = pose_err(res["T64"], gold[f"{name}_T64"])
    assert dt < 1e-6 and dr < 1e-6  # north_star: 1e-6 m / 1e-6 rad
    assert [int(res["converged"]), res["iterations"]] == list(gold[f"{name}_flags"])
    tr = r.lm_trace()
    assert tr.shape == gold[f"{name}_trace"].shape and np.array_equal(tr[:, [0, 1, 7]], gold[f"{name}_trace"][:, [0, 1, 7]])
    assert rel(res["H"], gold[f"{name}_H"]) < 1e-8
    s, n_in, n_inl = r.fitness()
    assert abs(s - gold[f"{name}_fitness"][0]) / gold[f"{name}_fitness"][0] < 1e-9
    assert (n_in, n_inl) == (int(gold[f"{name}_fitness"][1]), int(gold[f"{name}_fitness"][2]))


def test_oracle_reproduces_the_fixture(gold, synth):
    for search in (0, 1):  # brute force (what made the fixture) and the oracle's kd-tree
        o = Oracle(search=search)
        o.set_params(maha_fp64=1, variant=gold["variant"], **DEPLOYED)
        o.set_input_target(gold["target"]); o.set_input_source(gold["source"])
        check(o, gold, synth, 1e-13, 1e-12, 1e-11)
    for name, kw in ALIGNS.items():
        o = Oracle(search=1)
        o.set_params(maha_fp64=1, variant=gold["variant"], **kw)
        o.set_input_target(gold["target"]); o.set_input_source(gold["source"])
        check_align(o, gold, name)


def test_fixture_is_a_registration_problem(gold):
    """sanity of the fixture itself: 400-point scans with the radar noise model (0.5 / 1 deg at up to 100 m) are a
    noise-dominated problem, so only coarse agreement with the generating motion is expected"""
    dt, dr = pose_err(gold["lm_default_T64"], gold["T_true"])
    assert dt < 1.0 and dr < np.deg2rad(3.0)
    assert gold["lm_deployed_flags"][0] == 1
    assert np.array_equal(gold["lm_default_T64"][3], [0, 0, 0, 1])


@pytest.mark.gpu
@pytest.mark.parametrize("fp64", [1, 0])
def test_cuda_matches_the_fixture(gorio, gold, synth, fp64):
    g = gorio.FastAPDGICP(0)
    g.set_params(maha_fp64=fp64, variant=gold["variant"], **DEPLOYED)
    g.set_input_target(gold["target"]); g.set_input_source(gold["source"])
    # fp64 Mahalanobis storage: only summation order differs (1e-10); fp32 storage (default): 1e-6 (north_star: 1e-5 on H)
    check(g, gold, synth, 1e-9, 1e-10 if fp64 else 1e-6, 1e-9 if fp64 else 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(ALIGNS))
@pytest.mark.parametrize("host_loop", [0, 1])
def test_cuda_align_matches_the_fixture(gorio, gold, name, host_loop):
    g = gorio.FastAPDGICP(0)
    g.set_params(maha_fp64=1, host_loop=host_loop, variant=gold["variant"], **ALIGNS[name])
    g.set_input_target(gold["target"]); g.set_input_source(gold["source"])
    check_align(g, gold, name)
