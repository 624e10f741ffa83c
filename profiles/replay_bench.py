#!/usr/bin/env python
"""Config C5: odometry replay (go-rio_b200/replay.py, the call pattern of ScanMatchingOdometryNodelet::matching) over a
synthetic drive: 10 k frames of ~1 k radar points through ONE registration handle (the frames are sequential: the guess of
frame i is the result of frame i-1), next to the CPU restatement on a shorter prefix of the same drive.

    python profiles/replay_bench.py [--frames 10000] [--cpu-frames 500]
"""
import argparse
import importlib
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


NOISE = 0.2  # scale of the synthetic sensor noise: at 1 the open-loop odometry is a biased random walk (synth.drive_frames)


def run(n_frames=10000, cpu_frames=500, points=1000, device=0):
    """returns the replay block (a dict); bench.py puts a shorter run of it into its JSON line"""
    import argparse as _a
    args = _a.Namespace(frames=n_frames, cpu_frames=cpu_frames, points=points)
    import numpy as np

    gorio = importlib.import_module("go-rio_b200")
    synth = importlib.import_module("go-rio_b200.synth")
    replay = importlib.import_module("go-rio_b200.replay")
    from oracle_binding import ORACLE_REF_SO, Oracle

    frames = list(synth.drive_frames(5000, args.frames, args.points, noise=NOISE))
    kw = dict(max_correspondence_distance=2.0, transformation_epsilon=0.1)
    g = gorio.FastAPDGICP(device)
    g.set_params(**kw)
    replay.replay(g, frames[:50])  # warm-up (allocations, module load)
    g = gorio.FastAPDGICP(device)
    g.set_params(**kw)
    l0 = g.launch_count()
    t = time.perf_counter()
    rg = replay.replay(g, frames)
    t_gpu = time.perf_counter() - t
    cores = len(os.sched_getaffinity(0))
    ref = os.path.exists(ORACLE_REF_SO)
    o = Oracle(search=2 if ref else 1, threads=cores, ref=ref)
    o.set_params(**kw)
    sub = frames[: args.cpu_frames]
    t = time.perf_counter()
    ro = replay.replay(o, sub)
    t_cpu = time.perf_counter() - t
    g2 = gorio.FastAPDGICP(device)
    g2.set_params(**kw)
    rg2 = replay.replay(g2, sub)
    return {
        "workload": f"C5 odometry replay, {args.frames} frames x {args.points} points, scan-to-scan (ScanMatchingOdometryNodelet::matching + KeyframeUpdater), "
                    f"deployed parameters, sensor noise x {NOISE}",
        "gpu_frames_per_s": args.frames / t_gpu, "gpu_ms_per_frame": 1e3 * t_gpu / args.frames,
        "gpu_launches_per_frame": (g.launch_count() - l0) / args.frames,
        "keyframes": rg["n_keyframes"], "not_converged": rg["n_not_converged"], "mean_lm_iterations": float(np.mean(rg["iterations"])),
        "path_m": rg["path_m"], "final_drift_m": rg["final_drift_m"], "drift_fraction_of_path": rg["final_drift_m"] / max(rg["path_m"], 1e-9),
        "cpu_frames_per_s": len(sub) / t_cpu, "cpu_threads": cores, "cpu_frames": len(sub),
        "cpu_kind": "oracle/_ref (reference loop structure + nanoflann)" if ref else "oracle port",
        "same_keyframes_on_cpu_prefix": rg2["n_keyframes"] == ro["n_keyframes"],
        "max_pose_difference_on_cpu_prefix_m": float(np.abs(rg2["poses"][:, :3, 3] - ro["poses"][:, :3, 3]).max()),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--cpu-frames", type=int, default=500)
    ap.add_argument("--points", type=int, default=1000)
    args = ap.parse_args()
    print(json.dumps(run(args.frames, args.cpu_frames, args.points)))


if __name__ == "__main__":
    main()
