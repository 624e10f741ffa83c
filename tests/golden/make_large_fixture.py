#!/usr/bin/env python
"""Writes tests/golden/tiled_2m_oracle.npz: what the CPU oracle computes on the seed-4000 tiled 2 M-point pair (the C4
cloud at a tenth of its size: 256 MB of per-point state, larger than the 126 MB L2, so the GPU's streaming kernels and
its thread-per-point kNN kernel run as they do at 20 M) — H, b, err of one linearisation at the ground-truth pose (28
doubles), err at a trial pose, the matched-point count, a checksum of all correspondences and 1 000 sampled kNN rows of the
target. The clouds themselves are regenerated from the seed by the test (synth.tiled_cloud_pair is deterministic), so the
fixture is a few hundred KB. Takes a few minutes of CPU (all cores); run in the build container:

    python tests/golden/make_large_fixture.py
"""
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from oracle_binding import Oracle  # noqa: E402

synth = importlib.import_module("go-rio_b200.synth")
N = 2_000_000


def corr_checksum(c):
    """order-dependent 64-bit checksum of an int32 correspondence array"""
    c = c.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    i = np.arange(c.shape[0], dtype=np.uint64)
    return int(np.bitwise_xor.reduce((c + np.uint64(1)) * (i * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019))))


def main():
    t0 = time.time()
    src, tgt, T = synth.tiled_cloud_pair(4000, N)
    o = Oracle(search=1, threads=os.cpu_count() or 1)
    o.set_params(max_correspondence_distance=2.0, maha_fp64=1)
    o.set_input_target(tgt)
    o.set_input_source(src)
    e, H, b = o.linearize(T)
    print("linearize", time.time() - t0)
    c, sq = o.get_correspondences()
    rows = np.linspace(0, N - 1, 1000).astype(np.int64)
    T2 = T @ synth.make_pose([0.03, -0.02, 0.01], [0.001, -0.002, 0.003])
    out = {"n": np.int64(N), "seed": np.int64(4000), "T": T, "T_trial": T2, "err": np.float64(e), "H": H, "b": b,
           "err_trial_stale": np.float64(o.compute_error(T2)), "n_matched": np.int64((c >= 0).sum()), "corr_checksum": np.uint64(corr_checksum(c)),
           "corr_rows": c[rows], "sqd_rows": sq[rows], "knn_rows": rows, "knn_target": o.get_neighbors(1)[rows], "knn_source": o.get_neighbors(0)[rows],
           "src_head": src[:4], "tgt_head": tgt[:4]}
    cs = o.get_source_covariances()[rows]
    out["cov_source_rows"] = cs
    np.savez_compressed(os.path.join(HERE, "tiled_2m_oracle.npz"), **out)
    print("done", time.time() - t0, {k: getattr(v, "shape", v) for k, v in out.items() if k not in ("H", "b", "T", "T_trial")})


if __name__ == "__main__":
    main()
