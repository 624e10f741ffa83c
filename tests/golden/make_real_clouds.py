#!/usr/bin/env python
"""Writes tests/golden/real_lidar_pair.npz: the two real clouds of the reference tree (ndt_omp/data/251370668.pcd,
251371071.pcd — the only real point clouds in /root/reference; 69 088 and 69 792 points, 7 % of them exact duplicates,
which makes the (d2, index) tie rule visible) together with what the CPU oracle computes on them: the registration of
one onto the other (default and deployed parameters), 1 000 sampled kNN rows, the correspondences and H / b / err at the
identity. Run in the build container (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_real_clouds.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from oracle_binding import Oracle  # noqa: E402

pcd = importlib.import_module("go-rio_b200.pcd")
DATA = "/root/reference/ndt_omp/data"


def main():
    tgt = pcd.xyz_label(pcd.read_pcd(os.path.join(DATA, "251370668.pcd")))
    src = pcd.xyz_label(pcd.read_pcd(os.path.join(DATA, "251371071.pcd")))
    out = {"target_xyz": tgt[:, :3].copy(), "source_xyz": src[:, :3].copy()}
    o = Oracle(search=1, threads=os.cpu_count() or 1)
    o.set_params(max_correspondence_distance=2.0, maha_fp64=1)
    o.set_input_target(tgt)
    o.set_input_source(src)
    e, H, b = o.linearize(np.eye(4))  # (computes the covariances, hence the neighbour lists)
    rows = np.linspace(0, tgt.shape[0] - 1, 1000).astype(np.int64)
    out["knn_rows"] = rows
    out["knn_target"] = o.get_neighbors(1)[rows]
    out["err_I"], out["H_I"], out["b_I"] = np.float64(e), H, b
    c, _ = o.get_correspondences()
    out["corr_I"] = c
    r = o.align()
    out["T64_default"], out["iterations_default"], out["converged_default"] = r["T64"], np.int32(r["iterations"]), np.bool_(r["converged"])
    o.set_params(transformation_epsilon=0.1)
    r = o.align()
    out["T64_deployed"], out["iterations_deployed"] = r["T64"], np.int32(r["iterations"])
    o.swap_source_and_target()
    r = o.align()
    out["T64_deployed_backward"] = r["T64"]
    np.savez_compressed(os.path.join(HERE, "real_lidar_pair.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
