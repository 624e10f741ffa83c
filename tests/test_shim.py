"""The C++ drop-in class (go-rio_b200/include/fast_gicp/gicp/fast_apdgicp.hpp):
compiles against a stub of the pcl::Registration interface (no PCL/Eigen in this
image), and — on a GPU — gives the same result as the C-ABI driven from ctypes
when used the way 4DRadarSLAM's factory and nodelets use it."""
import json
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_DIR = os.path.join(REPO, "go-rio_b200", "shim_test")


def test_shim_compiles_against_the_registration_interface(gorio):
    if not os.path.exists(gorio.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    subprocess.check_call(["bash", os.path.join(SHIM_DIR, "build.sh")], env={**os.environ, "MAKEFLAGS": ""})
    assert os.path.exists(os.path.join(SHIM_DIR, "test_shim"))


def test_shim_keeps_the_reference_surface():
    """every public/protected member function of the reference class is declared by the shim"""
    text = open(os.path.join(REPO, "go-rio_b200", "include", "fast_gicp", "gicp", "fast_apdgicp.hpp")).read()
    text += open(os.path.join(REPO, "go-rio_b200", "include", "fast_gicp", "gicp", "lsq_registration.hpp")).read()
    for name in ["setNumThreads", "setCorrespondenceRandomness", "setRegularizationMethod", "setAzimuthVar", "setElevationVar",
                 "setDistVar", "swapSourceAndTarget", "clearSource", "clearTarget", "setInputSource", "setSourceCovariances",
                 "setInputTarget", "setTargetCovariances", "getSourceCovariances", "getTargetCovariances", "computeTransformation",
                 "update_correspondences", "linearize", "compute_error", "setRotationEpsilon", "setInitialLambdaFactor",
                 "setDebugPrint", "getFinalHessian", "evaluateCost", "setSearchMethodTarget"]:
        assert name in text, name


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["apdgicp", "gicp", "vgicp"])
def test_shim_matches_the_c_abi(gorio, synth, tmp_path, variant):
    exe = os.path.join(SHIM_DIR, "test_shim")
    if not os.path.exists(exe):
        subprocess.check_call(["bash", os.path.join(SHIM_DIR, "build.sh")])
    src, tgt, _ = synth.submap_pair(2002, n_source=1000, n_frames=5, n_per_frame=1500)
    fs, ft = str(tmp_path / "s.f32"), str(tmp_path / "t.f32")
    src.tofile(fs); tgt.tofile(ft)
    out = subprocess.run([exe, fs, str(src.shape[0]), ft, str(tgt.shape[0]), variant], capture_output=True, text=True, check=True)
    r = json.loads(out.stdout.strip().splitlines()[-1])
    g = gorio.FastAPDGICP(0)
    if variant == "vgicp":  # fast_gicp::FastVGICP as the shim test's factory sets it up (registrations.cpp:64-72)
        g.set_params(transformation_epsilon=0.1, variant=2, voxel_resolution=2.0, voxel_search=1)
    else:
        g.set_params(max_correspondence_distance=2.0, transformation_epsilon=0.1, variant=1 if variant == "gicp" else 0)
    g.set_input_target(synth.to_pcl_xyzinormal(tgt)); g.set_input_source(synth.to_pcl_xyzinormal(src))
    ra = g.align(want_aligned=True)
    assert bool(r["converged"]) == ra["converged"]
    assert np.array_equal(np.array(r["T"], dtype=np.float32).reshape(4, 4), ra["T"])
    assert abs(r["fitness_gpu"] - g.fitness()[0]) < 1e-12
    assert np.allclose(r["aligned0"], ra["aligned"][0], atol=0)
    assert r["n_aligned"] == src.shape[0] and r["n_cov"] == tgt.shape[0]
    assert abs(r["cost"] - g.linearize(ra["T"].astype(np.float64), want_hb=False)) / r["cost"] < 1e-12
    assert "swapped: converged=" in out.stderr
    # --- the base class's search method is the GPU grid (apd_search.hpp) ---
    # pcl::Registration::align -> initCompute built no CPU kd-tree (the stub counts what real PCL would build per target)
    assert r["tree_builds"] == 0 and r["tree_builds_end"] == 0
    # getFitnessScore() through the base pointer (PCL's per-point loop; its own transform rounds differently in the last bit)
    assert abs(r["fitness_pcl"] - r["fitness_gpu"]) / r["fitness_gpu"] < 1e-6
    # the nodelet's inlier loop over the aligned cloud (scan_matching_odometry_nodelet.cpp:680-688)
    assert r["nodelet_inliers"] == r["inliers"] == g.fitness()[2]
    # 2 n point-at-a-time queries were served by ONE pass over the source, none by a round trip of its own
    assert r["search_passes"] == 1 and r["search_hits"] == 2 * src.shape[0] and r["search_singles"] == 0
    assert r["knn_equal"] == 1  # an arbitrary query, k = 5: same indices and distances as brute force
    # --- keyframe promotion keeps grid + covariances on the device: only the new scan's covariances are computed — and with
    # every target covariance valid the loop kernel computes them itself (fused prologue): no kNN launch at all
    if variant == "vgicp":  # (host-driven loop: covariances of both clouds up front; the promoted keyframe still brings its own along)
        assert r["promoted_same_pose"] == 1 and r["knn_launches_promoted"] < r["knn_launches_first"]
        return
    assert r["knn_launches_first"] == 4 and r["knn_launches_promoted"] == 0 and r["promoted_same_pose"] == 1
