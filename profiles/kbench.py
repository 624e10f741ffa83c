#!/usr/bin/env python
"""Kernel-level micro-benchmark / ncu target (not the bench contract; see bench.py).

    python profiles/kbench.py [--mode c2|c1|big] [--n N] [--reps R]

Runs the {clear; set; align} protocol on ONE handle with CUDA-event profiling on and
prints the average device milliseconds per kernel class and per registration.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="c2", choices=["c1", "c2", "big"])
    ap.add_argument("--n", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--fp64", type=int, default=0)
    args = ap.parse_args()
    import torch

    gorio = importlib.import_module("go-rio_b200")
    synth = importlib.import_module("go-rio_b200.synth")
    dev = torch.device("cuda", 0)
    if args.mode == "c1":
        src, tgt, T = synth.scan_pair(1000, 1000)
    elif args.mode == "c2":
        src, tgt, T = synth.submap_pair(2000)
    else:
        src, tgt, T = synth.tiled_cloud_pair(4000, args.n)
    ds, dt = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev)
    g = gorio.FastAPDGICP(0)
    kw = dict(max_correspondence_distance=2.0, maha_fp64=args.fp64)
    if args.mode != "big":
        kw["transformation_epsilon"] = 0.1
    g.set_params(**kw)

    def one():
        g.clear_target(); g.clear_source()
        g.set_input_target_device(dt.data_ptr(), tgt.shape[0])
        g.set_input_source_device(ds.data_ptr(), src.shape[0])
        if args.mode == "big":
            g.linearize(T)
            g.compute_error(T)
            return None
        return g.align()

    for _ in range(3):
        r = one()
    g.set_profiling(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.reps):
        r = one()
    ev1.record()
    torch.cuda.synchronize()
    k = g.kernel_ms()
    out = {"mode": args.mode, "n_source": int(src.shape[0]), "n_target": int(tgt.shape[0]), "reps": args.reps,
           "wall_ms_per_rep": ev0.elapsed_time(ev1) / args.reps,
           "kernel_ms_per_rep": {c: round(ms / args.reps, 5) for c, (ms, cnt) in k.items()},
           "launches_per_rep": {c: cnt / args.reps for c, (ms, cnt) in k.items()}}
    if r is not None:
        out["iterations"] = r["iterations"]
        out["converged"] = r["converged"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
