#!/usr/bin/env python
"""Where a pooled registration spends its host wall time (APD_BATCH_TRACE=1): one mode, one pool, a few steps.

    python profiles/batch_trace.py --mode dev|host [--streams 64] [--pairs 256] [--steps 10]
"""
import argparse
import importlib
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ["APD_BATCH_TRACE"] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="dev", choices=["dev", "host"])
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch

    gorio = importlib.import_module("go-rio_b200")
    synth = importlib.import_module("go-rio_b200.synth")
    dev = torch.device("cuda", 0)
    host_pairs = bench.make_pairs(synth, 0, args.pairs)
    keep, cache, pairs = [], {}, []
    for s, t in host_pairs:
        if args.mode == "host":
            pairs.append((s, t, None))
            continue
        key = (s.ctypes.data, t.ctypes.data)
        if key not in cache:
            ds, dt = torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev)
            keep += [ds, dt]
            cache[key] = ((ds.data_ptr(), s.shape[0]), (dt.data_ptr(), t.shape[0]), None)
        pairs.append(cache[key])
    batch = gorio.Batch(0, n_workers=args.streams, **bench.DEPLOYED)
    prep = batch.prepare(pairs)
    for _ in range(3):
        batch.align(prep, with_fitness=False, parse=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        batch.align(prep, with_fitness=False, parse=False)
    dt_s = time.perf_counter() - t0
    print(f"mode={args.mode} streams={args.streams} pairs={args.pairs}: {args.pairs * args.steps / dt_s:.0f} registrations/s "
          f"({1e3 * dt_s / args.steps:.2f} ms per step)", flush=True)
    batch.close()


if __name__ == "__main__":
    main()
