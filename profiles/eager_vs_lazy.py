#!/usr/bin/env python
"""Do the pool's on-demand and eager target covariances give the same poses on the bench's own pairs? (diagnostic)

    python profiles/eager_vs_lazy.py [--pairs 256]
"""
import argparse
import importlib
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--stress", type=int, default=0, help="K > 0: only the eager pool, 256 in flight, K passes back to back as bench.py runs it")
    ap.add_argument("--device", action="store_true", help="clouds resident on the device, as bench.py's value leg")
    args = ap.parse_args()
    synth = importlib.import_module("go-rio_b200.synth")
    pairs = bench.make_pairs(synth, 0, 1, args.pairs, 0)
    gorio = importlib.import_module("go-rio_b200")
    items = [(s, t, None) for s, t in pairs]
    if args.device:
        import torch
        keep = [(torch.from_numpy(s).cuda(), torch.from_numpy(t).cuda()) for s, t in pairs]
        items = [((ds.data_ptr(), ds.shape[0]), (dt.data_ptr(), dt.shape[0]), None) for ds, dt in keep]
    if args.stress:
        os.environ["APD_LAZY_TARGET_COV"] = "0"
        b = gorio.Batch(0, n_workers=256, **bench.DEPLOYED)
        prep = b.prepare(items)
        for rnd in range(3):
            rep = b.repeat(prep, args.stress)
            b.align(rep, with_fitness=False, parse=False)
            v = b.results_view(rep["res"])
            bad = np.nonzero(v["status"] != 0)[0]
            print("round", rnd, "failed pairs:", [(int(i), int(i) % len(items), int(v["status"][i])) for i in bad][:10], "of", rep["n"])
        b.close()
        return
    out = {}
    for name, env in (("lazy", {}), ("eager", {"APD_LAZY_TARGET_COV": "0"}), ("lazy4", {"APD_CELLS_PER_POINT_MID_LAZY": "4"})):
        os.environ.update(env)
        b = gorio.Batch(0, n_workers=64, **bench.DEPLOYED)
        for k in env:
            del os.environ[k]
        prep = b.prepare(items)
        out[name] = [b.align(prep, with_fitness=True) for _ in range(2)]
        b.close()
    for name in ("eager", "lazy4"):
        diff = []
        for i, (a, c) in enumerate(zip(out["lazy"][1], out[name][1])):
            if not np.array_equal(a["T"], c["T"]) or a["iterations"] != c["iterations"]:
                diff.append((i, float(np.abs(np.asarray(a["T"], dtype=np.float64) - np.asarray(c["T"], dtype=np.float64)).max()), a["iterations"], c["iterations"],
                             a["converged"], c["converged"], a["fitness"], c["fitness"]))
        print(name, "pairs that differ from lazy:", len(diff), "of", len(pairs))
        for d in diff[:12]:
            print("   ", d)
    rep = sum(not np.array_equal(a["T"], c["T"]) for a, c in zip(out["lazy"][0], out["lazy"][1]))
    rep_e = sum(not np.array_equal(a["T"], c["T"]) for a, c in zip(out["eager"][0], out["eager"][1]))
    print("run-to-run differences: lazy", rep, "eager", rep_e)


if __name__ == "__main__":
    main()
