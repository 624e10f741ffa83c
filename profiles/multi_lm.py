#!/usr/bin/env python
"""ncu target: ONE launch of the loop kernel over N pairs (N x cluster CTAs), i.e. the kernel under the co-residency a
pool runs it with (ncu profiles kernels one at a time, so a pool's own launches are always seen alone).

    python profiles/multi_lm.py [--jobs 74] [--repeat 2]
"""
import argparse
import ctypes
import importlib
import os
import sys

os.environ.setdefault("APD_LAZY_TARGET_COV", "1")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=74)
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    synth = importlib.import_module("go-rio_b200.synth")
    pairs = bench.make_pairs(synth, 0, 1, args.jobs, 0)
    import torch
    gorio = importlib.import_module("go-rio_b200")
    lib = gorio.load()
    dev = torch.device("cuda", 0)
    hs, keep = [], []
    for s, t in pairs:
        ds, dt = torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev)
        keep += [ds, dt]
        g = gorio.FastAPDGICP(0)
        g.set_params(**bench.DEPLOYED)
        g.set_input_target_device(dt.data_ptr(), t.shape[0])
        g.set_input_source_device(ds.data_ptr(), s.shape[0])
        hs.append(g)
    arr = (ctypes.c_void_p * len(hs))(*[g._h for g in hs])
    lib.apd_debug_multi_align.restype = ctypes.c_int
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rc = lib.apd_debug_multi_align(arr, ctypes.c_int32(len(hs)), ctypes.c_int32(1))  # warm-up
    assert rc == 0, rc
    torch.cuda.synchronize()
    e0.record()
    rc = lib.apd_debug_multi_align(arr, ctypes.c_int32(len(hs)), ctypes.c_int32(args.repeat))
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, rc
    print({"jobs": args.jobs, "repeat": args.repeat, "ms_per_launch_incl_host": e0.elapsed_time(e1) / args.repeat})


if __name__ == "__main__":
    main()
