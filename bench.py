#!/usr/bin/env python
"""bench.py — FastAPDGICP registrations/s (scan-to-submap) and linearize GB/s.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic scan/submap
pairs: for every pair {clearTarget; clearSource; setInputTarget; setInputSource;
align} (the reference benchmark protocol, fast_apdgicp/src/align.cpp:57-83) with
the deployed parameters (4DRadarSLAM/launch/ntu_loop2.launch:88-99).

  value     registrations/s, clouds already resident in HBM (apd_set_*_device)
  e2e       the same through the C-ABI with HOST buffers (apd_set_source/target
            stage + copy the clouds, apd_align copies the pose back)
  roofline  the linearize kernel on a cloud larger than L2 (CUDA events on the
            handle's stream, live in this run) against the measured HBM peak
  cpu_baseline  the reference-structure CPU restatement (oracle/_ref: OpenMP +
            the reference tree's nanoflann; else the oracle port) on a bounded
            sample of the same pairs, all host threads

N > 1: the pairs are sharded across ranks (one process per GPU, no data-path
collective — pairs are independent, reference loop_detector.cpp:222-236), weak
scaling; value = all pairs / max-over-ranks time.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# one hardware work queue per worker stream of the batch pool (read when the CUDA context is created, i.e. before torch
# touches the GPU; libapdgicp.so sets the same default when it is loaded — see apdgicp.cu: apd_default_connections)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

DEPLOYED = dict(max_correspondence_distance=2.0, transformation_epsilon=0.1)
METRIC = "APDGICP registrations/sec (scan-to-submap)"
UNIT = "registrations/s"
WORKLOAD = ("C2 scan-to-submap: 2000-pt radar scan vs 60000-pt keyframe submap, k=20, deployed params "
            "(max_corr_dist 2.0, trans_eps 0.1, LM, PLANE)")  # the same string in both arms
BYTES_PER_POINT_LINEARIZE = 64  # src 16 + corr 4 + tgt 16 + maha 24 + geo 4
BYTES_PER_POINT_KNNCOV = 68     # read point 16, write cov 48 + geo 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=512, help="scan/submap pairs per step and per GPU (config C3: 4096 pairs over 8 GPUs)")
    ap.add_argument("--streams", type=int, default=64, help="registrations in flight per GPU: handles (CUDA stream each) of the batch context, driven by a few host threads")
    ap.add_argument("--roofline-points", type=int, default=20_000_000)
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--roofline-reps", type=int, default=10)
    ap.add_argument("--roofline-only", action="store_true", help="profiling aid: skip the registrations/s part")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-pairs", type=int, default=8, help="pairs per step of the reference arm")
    ap.add_argument("--no-fused", action="store_true", help="c4: ncclAllReduce after the reduction kernels instead of the peer-memory exchange inside them")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2 (default): scan-to-submap pairs, sharded by pair; c4: one large cloud, source-sharded with an NCCL all-reduce")
    return ap.parse_args()


def make_pairs(synth, rank, n_pairs, distinct=8):
    """C2-shaped pairs: 2000-point scan vs 60k-point submap. `distinct` scenes are
    generated (numpy generation costs ~1 s each) and cycled to n_pairs. Weak scaling
    means the SAME work per GPU: every rank registers the same scenes (seeds 2000..),
    starting at a different one. (With per-rank seeds the ranks' scenes need 33 to 65
    outer iterations per 8 scenes, and the max-over-ranks time measures the unluckiest
    draw: 8 x B200 read 80 % of 8 x the 1-GPU rate for that reason alone.)"""
    base = []
    for i in range(min(distinct, n_pairs)):
        s, t, _ = synth.submap_pair(2000 + i)
        base.append((np.ascontiguousarray(s), np.ascontiguousarray(t)))
    return [base[(i + rank) % len(base)] for i in range(n_pairs)]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------ reference arm ----
def load_cpu_impl():
    from oracle_binding import ORACLE_REF_SO, oracle_lib  # tests/oracle_binding.py (checker side)
    if os.path.exists(ORACLE_REF_SO):
        return oracle_lib(ref=True), "port", 2, "oracle/_ref (reference loop structure + the reference tree's nanoflann kd-tree, OpenMP)"
    return oracle_lib(ref=False), "port", 1, "oracle port (own exact kd-tree, OpenMP)"


def cpu_registrations(lib, search, pairs, reps):
    """Runs the CPU restatement on `pairs` x reps with all host threads; returns (seconds, n)."""
    h = ctypes.c_void_p()
    lib.apdo_create(ctypes.byref(h))
    gorio = importlib.import_module("go-rio_b200")
    p = gorio.ApdParams()
    lib.apdo_get_params(h, ctypes.byref(p))
    p.max_correspondence_distance = DEPLOYED["max_correspondence_distance"]
    p.transformation_epsilon = DEPLOYED["transformation_epsilon"]
    lib.apdo_set_params(h, ctypes.byref(p))
    lib.apdo_set_search(h, ctypes.c_int(search))
    # setNumThreads(0) = all cores (registrations.cpp:41); torchrun exports OMP_NUM_THREADS=1, so ask for the cores explicitly
    lib.apdo_set_num_threads(h, ctypes.c_int(host_cores()))
    ms = (ctypes.c_double * reps)()
    total = 0.0
    for s, t in pairs:
        rc = lib.apdo_bench_align(h, s.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(s.shape[0]), t.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_int32(t.shape[0]), ctypes.c_int32(16), ctypes.c_int32(0), ctypes.c_int32(12), None,
                                  ctypes.c_int32(reps), ms)
        assert rc == 0
        total += sum(ms) / 1e3
    lib.apdo_destroy(h)
    return total, len(pairs) * reps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth = importlib.import_module("go-rio_b200.synth")
    lib, kind, search, what = load_cpu_impl()
    cores = host_cores()
    pairs = make_pairs(synth, 0, args.ref_pairs, distinct=min(8, args.ref_pairs))
    for _ in range(args.warmup):
        cpu_registrations(lib, search, pairs[:1], 1)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        t, n = cpu_registrations(lib, search, pairs, 1)
        t_total += t
        n_total += n
    value = n_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "pairs_per_step": len(pairs), "protocol": "clearTarget;clearSource;setInputTarget;setInputSource;align"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{len(pairs)} pairs per step x {args.steps} steps; {what}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------- clocks ----
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._p = None

    def __enter__(self):
        # ONE nvidia-smi process in loop mode (the recipe's clocks line): starting a process per sample re-initialises
        # NVML every 50 ms and measurably slows kernel launches of the 32-64 worker threads being timed
        try:
            self._p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                        "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._p = None
        return self

    def __exit__(self, *a):
        if self._p is None:
            return
        self._p.terminate()
        try:
            out, _ = self._p.communicate(timeout=6)
        except Exception:
            self._p.kill()
            out = ""
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) >= 7:
                self.samples.append(f)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = max(float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples if len(s) > 3 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------ our arm ----
LM_RESULT_HEAD_BYTES = 1984  # apdgicp.cu kLmHeadBytes: pose, final hessian, flags and the first LM trace rows, per registration


def run_c4(args):
    """Config C4: `--roofline-points` source points vs as many target points, the source split contiguously over the
    ranks, the target replicated; one step = one LM iteration's device work (linearize = update_correspondences +
    H/b/err reduction + all-reduce of 28 doubles, then one compute_error + all-reduce of 1 double)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gorio = importlib.import_module("go-rio_b200")
    synth = importlib.import_module("go-rio_b200.synth")
    sharding = importlib.import_module("go-rio_b200.sharding")
    n = args.roofline_points
    src, tgt, T = synth.tiled_cloud_pair(4000, n)
    dsrc, dtgt = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev)
    g = gorio.FastAPDGICP(local_rank)
    g.set_params(max_correspondence_distance=2.0)
    if world > 1:
        sharding.init_comm(g, gorio.load(), rank, world, n, dist, dev, fused=not args.no_fused)
    # every rank sets the same full clouds; the library slices the work (covariances all-gathered, H/b/err all-reduced)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_setup = time.perf_counter()
    g.set_input_target_device(dtgt.data_ptr(), n)
    g.set_input_source_device(dsrc.data_ptr(), n)
    g.linearize(T)  # grids + covariances (+ all-gather) + the first linearisation
    torch.cuda.synchronize()
    t_setup = torch.tensor([time.perf_counter() - t_setup], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_setup, op=dist.ReduceOp.MAX)
    for _ in range(max(1, args.warmup)):
        g.linearize(T)
        g.compute_error(T)
    l0 = g.launch_count()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        err, H, bb = g.linearize(T)
        err2 = g.compute_error(T)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = g.launch_count() - l0
    g.set_profiling(True)
    g.linearize(T)
    g.compute_error(T)
    k = g.kernel_ms()
    if rank == 0:
        print(json.dumps({
            "metric": "APDGICP LM-iteration throughput on one large cloud (source points/s)", "value": n * args.steps / (ms / 1e3),
            "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4: {n} source vs {n} target points over {world} GPU(s): covariances computed by slices and "
                                   "all-gathered, source sliced for update_correspondences/linearize/compute_error, all-reduce of 28 "
                                   "doubles per linearize and 1 per compute_error "
                                   + ("by ncclAllReduce" if (args.no_fused or world == 1) else "inside the reduction kernels over NVLink peer memory"),
                       "l2": "inputs larger than L2"},
            "setup_ms": 1e3 * float(t_setup.item()), "setup": "grid builds + kNN covariances of both clouds (+ all-gather) + first linearize, wall clock, max over ranks",
            "gpu_launches": int(launches), "err": err, "err_trial": err2,
            "kernels_rank0_ms": {c: v[0] for c, v in k.items()},
        }), flush=True)
    if world > 1:
        g.comm_destroy()
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE line (the JSON): libraries that chat on fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c4":
        return run_c4(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the registration path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gorio = importlib.import_module("go-rio_b200")
    synth = importlib.import_module("go-rio_b200.synth")
    dev = torch.device("cuda", local_rank)

    line = None
    host_pairs = []
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    if args.roofline_only:
        line = {"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "note": "roofline-only profiling run"}
    else:
        host_pairs = make_pairs(synth, rank, args.pairs)
        # e2e inputs live in page-locked host memory (the contract's "pinned host memory"): the pool copies packed
        # float4 clouds straight out of it; pageable clouds would be staged through the handles' own pinned buffers
        pinned = {}
        def pin(a):
            if a.ctypes.data not in pinned:
                t = torch.from_numpy(a).pin_memory()
                pinned[a.ctypes.data] = (t, t.numpy())
            return pinned[a.ctypes.data][1]
        host_pairs = [(pin(s), pin(t)) for s, t in host_pairs]
        # HBM-resident copies of the clouds for `value`
        dev_tensors, dev_pairs = [], []
        cache = {}
        for s, t in host_pairs:
            key = (s.ctypes.data, t.ctypes.data)
            if key not in cache:
                ds, dt = torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev)
                dev_tensors += [ds, dt]
                cache[key] = (ds.data_ptr(), s.shape[0], dt.data_ptr(), t.shape[0])
            dev_pairs.append(cache[key])
        # the batch context of the C-ABI (apd_batch_*): `--streams` workers, one handle / CUDA stream / host thread each
        batch = gorio.Batch(local_rank, n_workers=args.streams, **DEPLOYED)
        prep_dev = batch.prepare([((ds, ns), (dt, nt), None) for ds, ns, dt, nt in dev_pairs])
        prep_host = batch.prepare([(s, t, None) for s, t in host_pairs])
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(prepared, steps, warmup):
            for _ in range(warmup):
                batch.align(prepared, with_fitness=False, parse=False)
            ms_total = 0.0
            l0 = batch.launch_count()
            cpu0 = time.process_time()
            for _ in range(steps):
                flush.zero_()  # L2 flush between timed iterations (untimed)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                batch.align(prepared, with_fitness=False, parse=False)  # returns when every pair's result is on the host
                e1.record()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                ms_total += float(ms.item())
            launches = batch.launch_count() - l0
            cpu_ms_per_pair = 1e3 * (time.process_time() - cpu0) / (steps * prepared["n"])  # all threads of this process
            return ms_total, launches, batch.align(prepared, with_fitness=False), cpu_ms_per_pair

        with ClockSampler(local_rank) as clocks:
            ms_dev, launches, results, cpu_dev = timed(prep_dev, args.steps, args.warmup)
            ms_e2e, _, results_h, cpu_e2e = timed(prep_host, args.steps, args.warmup)
        total_pairs = args.pairs * world
        value = total_pairs * args.steps / (ms_dev / 1e3)
        e2e_value = total_pairs * args.steps / (ms_e2e / 1e3)
        same = all(np.array_equal(a["T"], b["T"]) for a, b in zip(results, results_h))
        ok = all(r["status"] == 0 for r in results + results_h)
        h2d = sum(s.nbytes + t.nbytes for s, t in host_pairs)
        d2h = args.pairs * LM_RESULT_HEAD_BYTES

        # per-kernel-class device time of one step (profiling events on the worker streams; separate, untimed pass)
        batch.set_profiling(True)
        batch.align(prep_dev, with_fitness=False, parse=False)
        kms = {k: [ms, cnt] for k, (ms, cnt) in batch.kernel_ms().items()}
        batch.set_profiling(False)
        tot = sum(v[0] for v in kms.values()) or 1.0
        kernels = {k: {"ms_per_step": round(v[0], 4), "launches_per_step": v[1], "share": round(v[0] / tot, 4)} for k, v in kms.items()}

        if rank == 0:
            n_tgt_pts = sum(t.shape[0] for _, t in host_pairs[:1])
            knn_ms = kms.get("knn_cov", [0.0, 0])[0]
            knn_pts = sum(s.shape[0] + t.shape[0] for s, t in host_pairs)
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "pairs_per_step_per_gpu": args.pairs, "streams_per_gpu": args.streams,
                           "api": "apd_batch_align_device (value) / apd_batch_align (e2e)", "optimizer_loop": "device-resident (lm.cu), one launch per registration",
                           "protocol": "clearTarget;clearSource;setInputTarget;setInputSource;align",
                           "l2": "flushed (256 MiB write) between timed steps", "sharding": "pairs across ranks, no collective; every rank registers the same 8 scenes (equal work per GPU)",
                           "target_points": n_tgt_pts},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "same_result_as_device_resident": bool(same), "all_pairs_ok": bool(ok),
                        "api": "apd_batch_align (host AoS clouds in, poses out)"},
                "gpu_launches": int(launches),
                "host_cpu_ms_per_registration": {"device_resident": round(cpu_dev, 4), "e2e": round(cpu_e2e, 4), "host_cores": host_cores(),
                                                 "note": "process CPU time of rank 0 (all worker threads, incl. the L2 flush / barrier between steps) per registration"},
                "kernels": kernels,
                "knn_cov_roofline": {"bound": "hbm", "achieved": (BYTES_PER_POINT_KNNCOV * knn_pts / 1e9) / (knn_ms / 1e3) if knn_ms > 0 else None,
                                     "peak": peak, "unit": "GB/s", "note": "search-bound (L2-resident candidates), reported for the step's dominant kernel"},
                "clocks": clocks.summary(),
            }

    # ---- roofline of the linearize kernel on a cloud larger than L2 (rank 0, N = 1 only) ----
    if rank == 0 and world == 1 and not args.no_roofline:
        if not args.roofline_only:
            batch.close()
        dev_tensors = None
        torch.cuda.empty_cache()
        n = args.roofline_points
        src, tgt, Tgt = synth.tiled_cloud_pair(4000, n)
        ds, dt = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev)
        g = gorio.FastAPDGICP(local_rank)
        g.set_params(max_correspondence_distance=2.0)
        g.set_input_target_device(dt.data_ptr(), n)
        g.set_input_source_device(ds.data_ptr(), n)
        T = Tgt.copy()
        g.set_profiling(True)
        g.linearize(T)  # builds grids, covariances, correspondences
        k0 = g.kernel_ms()
        g.set_profiling(False)
        for _ in range(3):
            g.linearize(T)
            g.compute_error(T)
        g.set_profiling(True)
        reps = max(1, args.roofline_reps)
        for _ in range(reps):
            g.linearize(T)
            g.compute_error(T)
        k = g.kernel_ms()
        lin_ms = k["linearize"][0] / reps
        err_ms = k["error"][0] / reps
        corr_ms = k["corr"][0] / reps
        achieved = BYTES_PER_POINT_LINEARIZE * n / 1e9 / (lin_ms / 1e3)
        n_valid = int((g.get_correspondences()[0] >= 0).sum())
        traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of the same kernel at the same size
        try:
            cap = json.load(open(os.path.join(REPO, "profiles", "r01_ncu_linearize.json")))
            if cap.get("points") == n:
                kk = cap["linearize_kernel<fp32 maha, H+b+err>"]
                traffic = kk["dram_bytes_read"] + kk["dram_bytes_write"]
        except Exception:
            pass
        line["roofline"] = {
            "bound": "hbm", "kernel": "linearize_kernel<fp32 maha, H+b+err>", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/r01_ncu_linearize.json (ncu capture, bytes per launch)" if traffic else None,
            "algorithmic_bytes": BYTES_PER_POINT_LINEARIZE * n, "peak_source": peak_src, "points": n, "matched_points": n_valid,
            "bytes_per_point": BYTES_PER_POINT_LINEARIZE, "ms_per_launch": lin_ms,
            "compute_error": {"ms_per_launch": err_ms, "achieved": BYTES_PER_POINT_LINEARIZE * n / 1e9 / (err_ms / 1e3)},
            "update_correspondences": {"ms_per_launch": corr_ms, "achieved": 148 * n / 1e9 / (corr_ms / 1e3), "bytes_per_point": 148},
            "grid_build": {"ms_per_cloud": k0["grid"][0] / 2, "launches_per_cloud": k0["grid"][1] // 2},
            "knn_covariance": {"ms_per_cloud": k0["knn_cov"][0] / 2, "achieved": BYTES_PER_POINT_KNNCOV * n / 1e9 / (k0["knn_cov"][0] / 2 / 1e3),
                               "bytes_per_point": BYTES_PER_POINT_KNNCOV, "mqueries_per_s": n / 1e6 / (k0["knn_cov"][0] / 2 / 1e3)},
            "timing": "CUDA events on the handle's stream around each launch; working set 1.28 GB > L2",
        }
        g.close()

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.roofline_only:
        lib, kind, search, what = load_cpu_impl()
        lib.apdo_max_threads.restype = ctypes.c_int
        sample = host_pairs[:8]
        seen, uniq = set(), []
        for s, t in sample:
            if s.ctypes.data not in seen:
                seen.add(s.ctypes.data)
                uniq.append((s, t))
        t1, n1 = cpu_registrations(lib, search, uniq[:1], 1)
        reps = max(3, min(200, int(round(12.0 / max(t1 * len(uniq), 1e-3)))))  # ~12 s of CPU work
        t_cpu, n_cpu = cpu_registrations(lib, search, uniq, reps)
        line["cpu_baseline"] = {"value": n_cpu / t_cpu, "unit": UNIT, "cores": host_cores(), "kind": kind,
                                "sample": f"{len(uniq)} of the step's pairs x {reps} repetitions ({t_cpu:.1f} s wall); {what}"}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
