#!/usr/bin/env python
"""What bounds the batch pool: the launch rate of the host process, or the GPU? (profiling aid, not the bench contract)

    python profiles/pool_probe.py [--pairs 2048] [--streams 128] [--steps 5]

Prints the driver's empty-kernel launch ceiling for a few thread counts, then runs the C2 pool and reports, measured
under full load: registrations/s, launches/s, the loop kernel's mean duration and CTA-slot occupancy (from the kernels'
own %globaltimer stamps), and the host milliseconds per phase."""
import argparse
import ctypes
import importlib
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2048)
    ap.add_argument("--distinct", type=int, default=64)
    ap.add_argument("--streams", type=int, default=192)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-launch-rate", action="store_true")
    args = ap.parse_args()
    synth = importlib.import_module("go-rio_b200.synth")
    pairs = bench.make_pairs(synth, 0, 1, args.pairs, args.distinct)
    import torch
    gorio = importlib.import_module("go-rio_b200")
    lib = gorio.load()
    out = {}
    if not args.no_launch_rate:
        rate = (ctypes.c_double * 2)()
        out["launch_rate_per_s"] = {}
        for th in (1, 2, 4, 8):
            lib.apd_debug_launch_rate(0, 128, th, 40000, rate)
            out["launch_rate_per_s"][f"{th}_threads"] = [round(rate[0]), round(rate[1])]
    dev = torch.device("cuda", 0)
    cache, dev_pairs, keep = {}, [], []
    for s, t in pairs:
        if s.ctypes.data not in cache:
            ds, dt = torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev)
            keep += [ds, dt]
            cache[s.ctypes.data] = ((ds.data_ptr(), s.shape[0]), (dt.data_ptr(), t.shape[0]), None)
        dev_pairs.append(cache[s.ctypes.data])
    b = gorio.Batch(0, n_workers=args.streams, **bench.DEPLOYED)
    prep = b.prepare(dev_pairs)
    for _ in range(2):
        b.align(prep, with_fitness=False, parse=False)
    b.load_stats(reset=True)
    l0 = b.launch_count()
    torch.cuda.synchronize()
    c0 = time.process_time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.align(prep, with_fitness=False, parse=False)
    wall = time.perf_counter() - t0
    cpu = time.process_time() - c0
    st = b.load_stats()
    n = args.steps * args.pairs
    ctas = int(os.environ.get("APD_LM_CLUSTER", "2"))
    out.update({"streams": args.streams, "cluster": ctas, "threads": os.environ.get("APD_BATCH_THREADS", "default"),
                "registrations_per_s": round(n / wall), "launches_per_s": round((b.launch_count() - l0) / wall),
                "launches_per_registration": (b.launch_count() - l0) / n,
                "lm_kernel_ms_mean_under_load": st["lm_kernel_ms"] / max(1, st["registrations"]),
                "cta_slot_occupancy": st["lm_kernel_ms"] * ctas / (296.0 * wall * 1e3),
                "host_cpu_ms_per_registration": 1e3 * cpu / n,
                "host_ms_per_registration": {k: v / max(1, st["registrations"]) for k, v in st.items() if k.startswith("host_")},
                "lm_phase_ms_per_registration": {k: round(v / max(1, st["registrations"]), 4) for k, v in st["lm_phase_ms"].items() if v}})
    print(json.dumps(out))
    b.close()


if __name__ == "__main__":
    main()
