#!/usr/bin/env python
"""Generates tests/golden/apdgicp_c1_small.npz — known-answer vectors for the FastAPDGICP path — and
tests/golden/fastgicp_c1_small.npz — the same for the FastGICP variant (apd_params.variant = 1) on the same clouds.

The reference holds no golden vector, test or fixture for FastAPDGICP and cannot be built in
this image (no Eigen / PCL / FLANN), so these vectors come from the CPU oracle (oracle/, a line-by-line
restatement of fast_apdgicp_impl.hpp / lsq_registration_impl.hpp), AFTER this script has checked the oracle
against the independent NumPy/SciPy restatement (tests/numpy_restatement.py) on the same inputs. The input
clouds are stored in the file, so the fixture does not depend on the synthetic generator staying unchanged.

    python tests/golden/make_golden.py        # rewrites both .npz (run from the repo root)
    python tests/golden/make_golden.py gicp   # only the FastGICP one
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy_restatement as nr  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

OUT = os.path.join(HERE, "apdgicp_c1_small.npz")
OUT_GICP = os.path.join(HERE, "fastgicp_c1_small.npz")
DEPLOYED = dict(max_correspondence_distance=2.0, transformation_epsilon=0.1)


def sym6(c4):  # n x 4 x 4 -> n x 6 (xx xy xz yy yz zz)
    return np.stack([c4[:, 0, 0], c4[:, 0, 1], c4[:, 0, 2], c4[:, 1, 1], c4[:, 1, 2], c4[:, 2, 2]], axis=1)


def main(variant=0, path=OUT):
    gicp = variant == 1
    synth = importlib.import_module("go-rio_b200.synth")
    src, tgt, T_true = synth.scan_pair(4242, 400)
    src, tgt = np.ascontiguousarray(src), np.ascontiguousarray(tgt)
    pose2 = T_true @ synth.make_pose([0.05, -0.03, 0.01], [0.002, -0.001, 0.004])
    out = dict(source=src, target=tgt, T_true=T_true, pose2=pose2)

    def fresh(**kw):
        o = Oracle(search=0)  # brute force: the definitional search
        o.set_params(maha_fp64=1, variant=variant, **kw)
        o.set_input_target(tgt)
        o.set_input_source(src)
        return o

    o = fresh(**DEPLOYED)
    cov_t, cov_s = o.get_target_covariances(), o.get_source_covariances()
    out["nb_target"], out["nb_source"] = o.get_neighbors(1), o.get_neighbors(0)
    out["cov_target"], out["cov_source"] = sym6(cov_t), sym6(cov_s)
    # pin the oracle against the NumPy restatement before trusting it
    S = np.linalg.svd(nr.covariances(tgt, out["nb_target"], "NONE"), compute_uv=False)
    gap_ok = (S[:, 1] - S[:, 2]) / S[:, 0] > 1e-3  # the plane normal is ill-conditioned when sigma2 ~ sigma3
    assert np.abs(cov_t[gap_ok, :3, :3] - nr.covariances(tgt, out["nb_target"], "PLANE")[gap_ok]).max() < 1e-9
    d, idx = nr.knn_sets(tgt[:, :3], 20)
    clear = (d[:, 20] - d[:, 19]) > 1e-4 * d[:, 20]
    assert all(set(a) == set(b) for a, b in zip(out["nb_target"][clear], idx[clear, :20]))
    for name, T in (("I", np.eye(4)), ("P2", pose2)):
        err, H, b = o.linearize(T)
        c, sq = o.get_correspondences()
        Mn = nr.mahalanobis(T, src, tgt, cov_s[:, :3, :3], cov_t[:, :3, :3], c, gicp=gicp)
        en, Hn, bn = nr.linearize(T, src, tgt, cov_s[:, :3, :3], c, Mn, gicp=gicp)
        assert abs(err - en) / en < 1e-10 and np.abs(H - Hn).max() / np.abs(Hn).max() < 1e-10 and np.abs(b - bn).max() / np.abs(bn).max() < 1e-9
        out[f"corr_{name}"], out[f"sqd_{name}"] = c, sq
        out[f"maha_{name}"] = sym6(o.get_mahalanobis())
        out[f"H_{name}"], out[f"b_{name}"], out[f"err_{name}"] = H, b, np.float64(err)
        out[f"err_trial_{name}"] = np.float64(o.compute_error(T @ synth.make_pose([0.01, 0.0, -0.01], [0.0, 0.001, 0.0])))
    for name, kw in (("lm_deployed", DEPLOYED), ("lm_default", dict(max_correspondence_distance=2.0)),
                     ("gn", dict(max_correspondence_distance=2.0, optimizer=0, max_iterations=6))):
        o2 = fresh(**kw)
        r = o2.align()
        out[f"{name}_T64"], out[f"{name}_H"] = r["T64"], r["H"]
        out[f"{name}_flags"] = np.array([int(r["converged"]), r["iterations"]], np.int32)
        out[f"{name}_trace"] = o2.lm_trace()
        s, n_in, n_inl = o2.fitness()
        out[f"{name}_fitness"] = np.array([s, n_in, n_inl], np.float64)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "LM deployed:", out["lm_deployed_flags"], "GN:", out["gn_flags"])


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] != "gicp":
        main(0, OUT)
    main(1, OUT_GICP)
