#!/usr/bin/env python
"""bench.py — FastAPDGICP registrations/s (scan-to-submap) and linearize GB/s.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic scan/submap
pairs: for every pair {clearTarget; clearSource; setInputTarget; setInputSource;
align} (the reference benchmark protocol, fast_apdgicp/src/align.cpp:57-83) with
the deployed parameters (4DRadarSLAM/launch/ntu_loop2.launch:88-99).

The batch is config C3 of BASELINE.json: 4096 scan/submap pairs, EVERY ONE A DIFFERENT
SCENE (C2-shaped: 2000-point scan vs 60 000-point keyframe submap, seeds 3000 + i),
sharded across the N ranks by pair index (pair i -> rank i mod N, no data-path
collective: pairs are independent, reference loop_detector.cpp:222-236). The total
work is fixed, so this is strong scaling; at N = 1 the one GPU registers all 4096.

  value     registrations/s, clouds already resident in HBM (apd_batch_align_device)
  e2e       the same through the C-ABI with HOST buffers in the layout the drop-in
            class receives: pageable 48-byte pcl::PointXYZINormal clouds
            (apd_batch_align stages, copies, registers, brings the poses back)
  e2e_packed  the same with packed float4 clouds in page-locked memory (no staging)
  eager     device-resident rate with ALL target covariances computed up front, as
            the reference does (the default computes the ~4 100 a registration meets)
  parity_vs_cpu  max |dT| against the CPU restatement on the first pairs of the batch
  c4        config C4, one 20 M x 20 M registration source-sharded over the N ranks:
            ms per LM iteration (linearize + compute_error, all-reduce inside the
            reduction kernels over NVLink peer memory), err against a committed constant
  roofline  (N = 1) the linearize kernel on that cloud (> L2), CUDA events on the
            handle's stream, against the measured HBM peak; + the loop kernel's share
  replay    (N = 1) config C5: frames/s of the scan-to-scan odometry front end over a synthetic drive, next to
            the CPU restatement on a prefix of the same drive (same keyframes, same poses)
  cpu_baseline  (N = 1) the reference-structure CPU restatement (oracle/_ref: OpenMP +
            the reference tree's nanoflann; else the oracle port) on a bounded sample
"""
import argparse
import ctypes
import importlib
import json
import multiprocessing
import os
import subprocess
import sys
import time

import numpy as np

# one hardware work queue per worker stream of the batch pool: read by the driver when the CUDA context is created, i.e.
# before torch touches the GPU (the library itself no longer changes the environment when it is loaded)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

DEPLOYED = dict(max_correspondence_distance=2.0, transformation_epsilon=0.1)
METRIC = "APDGICP registrations/sec (scan-to-submap)"
UNIT = "registrations/s"
WORKLOAD = ("C3 batch of C2 scan-to-submap pairs: 2000-pt radar scan vs 60000-pt keyframe submap, every pair a different scene "
            "(seeds 3000+i), k=20, deployed params (max_corr_dist 2.0, trans_eps 0.1, LM, PLANE)")  # the same string in both arms
BYTES_PER_POINT_LINEARIZE = 64  # src 16 + corr 4 + tgt 16 + maha 24 + geo 4
BYTES_PER_POINT_CORR = 148
SCENE_SEED0 = 3000              # C3 pair i is scene 3000 + i
# err of one linearisation of the seed-4000 tiled 20 M-point pair at its ground-truth pose (sum over 20 M points): the
# value every sharding (N = 1, 2, 4, 8; NCCL or in-kernel exchange) must reproduce to 12 digits (summation order differs)
C4_ERR_20M = 4856.67264058955


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=4096, help="scan/submap pairs per step over ALL GPUs (config C3: 4096)")
    ap.add_argument("--distinct", type=int, default=0, help="distinct scenes among the pairs (0: all; fewer are cycled — profiling aid)")
    ap.add_argument("--streams", type=int, default=256, help="registrations in flight per GPU: handles (CUDA stream each) of the batch context, driven by a few host threads")
    ap.add_argument("--roofline-points", type=int, default=20_000_000)
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-replay", action="store_true")
    ap.add_argument("--replay-frames", type=int, default=2000, help="frames of the C5 odometry replay block (N = 1; profiles/replay_bench.py runs all 10 000)")
    ap.add_argument("--roofline-reps", type=int, default=10)
    ap.add_argument("--roofline-only", action="store_true", help="profiling aid: skip the registrations/s part")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-pairs", type=int, default=8, help="pairs per step of the reference arm")
    ap.add_argument("--no-fused", action="store_true", help="c4: ncclAllReduce after the reduction kernels instead of the peer-memory exchange inside them")
    ap.add_argument("--workload", default="c3", choices=["c3", "c4"],
                    help="c3 (default): the batch of scan-to-submap pairs (+ the c4 block); c4: only the large source-sharded cloud")
    return ap.parse_args()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _scene(seed):
    synth = importlib.import_module("go-rio_b200.synth")
    s, t, _ = synth.submap_pair(seed)
    return np.ascontiguousarray(s), np.ascontiguousarray(t)


def pair_indices(rank, world, total):
    """config C3's split: pair i -> rank i mod N (go-rio_b200/sharding.py: shard_pairs)"""
    return importlib.import_module("go-rio_b200.sharding").shard_pairs(total, rank, world)


def make_pairs(synth, rank, world, total, distinct=0, procs=None):
    """This rank's share of the `total` C2-shaped pairs, pair i being scene SCENE_SEED0 + (i mod distinct): (source, target)
    float32 [n,4] arrays. Scenes are generated by a pool of forked workers (0.1 s of numpy each) BEFORE the process
    touches CUDA."""
    distinct = total if distinct <= 0 else min(distinct, total)
    mine = pair_indices(rank, world, total)
    need = sorted({i % distinct for i in mine})
    seeds = [SCENE_SEED0 + j for j in need]
    if procs is None:
        procs = max(1, min(32, host_cores() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    if procs > 1 and len(seeds) > 4:
        with multiprocessing.get_context("fork").Pool(procs) as pool:
            scenes = pool.map(_scene, seeds, chunksize=max(1, len(seeds) // (8 * procs)))
    else:
        scenes = []
        for sd in seeds:
            s, t, _ = synth.submap_pair(sd)
            scenes.append((np.ascontiguousarray(s), np.ascontiguousarray(t)))
    by = dict(zip(need, scenes))
    return [by[i % distinct] for i in mine]


# ------------------------------------------------------------ reference arm ----
REF_CODE_SO = os.path.join(REPO, "oracle", "_ref", "libapd_ref_apdgicp_nf.so")


def load_cpu_impl():
    """The CPU arm. Preferred: the REFERENCE'S OWN CODE (fast_apdgicp / lsq_registration headers compiled where they lie under
    /root/reference against the stand-ins of oracle/ref_stubs, with the kd-tree the reference tree vendors; oracle/Makefile
    -> oracle/_ref, which travels to the GPU box) — kind "reference". Else the oracle port."""
    from oracle_binding import ORACLE_REF_SO, oracle_lib  # tests/oracle_binding.py (checker side)
    if os.path.exists(REF_CODE_SO):
        lib = ctypes.CDLL(REF_CODE_SO)
        lib.aref_create.restype = ctypes.c_void_p
        return lib, "reference", -1, ("oracle/_ref/libapd_ref_apdgicp_nf.so: the reference's own FastAPDGICP / LsqRegistration headers compiled in place "
                                     "(-O3, OpenMP) against stand-ins for Eigen / PCL, nearest neighbours through the reference tree's nanoflann kd-tree")
    if os.path.exists(ORACLE_REF_SO):
        return oracle_lib(ref=True), "port", 2, "oracle/_ref (reference loop structure + the reference tree's nanoflann kd-tree, OpenMP)"
    return oracle_lib(ref=False), "port", 1, "oracle port (own exact kd-tree, OpenMP)"


def cpu_handle(lib, search):
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
    return h


def ref_code_registrations(lib, pairs, reps, want_poses=False):
    """the reference's own code on `pairs` x reps, all host threads (setNumThreads as registrations.cpp:41 asks: all cores)"""
    h = ctypes.c_void_p(lib.aref_create())
    C = ctypes
    lib.aref_set_params(h, C.c_int(20), C.c_int(3), C.c_double(DEPLOYED["max_correspondence_distance"]), C.c_double(0.86), C.c_double(0.5), C.c_double(1.0),
                        C.c_int(64), C.c_int(0), C.c_double(2e-3), C.c_double(DEPLOYED["transformation_epsilon"]), C.c_int(10), C.c_double(1e-9),
                        C.c_int(host_cores()))
    ms = (C.c_double * reps)()
    total, poses = 0.0, []
    for s, t in pairs:
        s32, t32 = np.ascontiguousarray(s, np.float32), np.ascontiguousarray(t, np.float32)
        T = np.zeros(16, np.float32)
        conv, it = C.c_int(), C.c_int()
        rc = lib.aref_bench_align(h, s32.ctypes.data_as(C.c_void_p), C.c_int(s32.shape[0]), t32.ctypes.data_as(C.c_void_p), C.c_int(t32.shape[0]),
                                  C.c_int(reps), ms, T.ctypes.data_as(C.c_void_p), C.byref(conv), C.byref(it))
        assert rc == 0
        total += sum(ms) / 1e3
        poses.append((T.reshape(4, 4).T.copy(), bool(conv.value), it.value))
    lib.aref_destroy(h)
    return (total, len(pairs) * reps, poses) if want_poses else (total, len(pairs) * reps)


def cpu_registrations(lib, search, pairs, reps):
    """Runs the CPU arm on `pairs` x reps with all host threads; returns (seconds, n)."""
    if search < 0:
        return ref_code_registrations(lib, pairs, reps)
    h = cpu_handle(lib, search)
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


def cpu_poses(pairs):
    """final float poses of the CPU restatement (the checker) for `pairs`, and whether each converged"""
    from oracle_binding import ORACLE_REF_SO, Oracle
    ref = os.path.exists(ORACLE_REF_SO)
    o = Oracle(search=2 if ref else 1, threads=host_cores(), ref=ref)
    o.set_params(**DEPLOYED)
    out = []
    for s, t in pairs:
        o.clear_target(); o.clear_source()
        o.set_input_target(t); o.set_input_source(s)
        r = o.align()
        out.append((r["T"], r["converged"], r["iterations"]))
    o.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth = importlib.import_module("go-rio_b200.synth")
    lib, kind, search, what = load_cpu_impl()
    cores = host_cores()
    pairs = make_pairs(synth, 0, 1, args.ref_pairs)  # the first pairs of the batch our arm registers
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
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "pairs_per_step": len(pairs), "sample": f"the first {len(pairs)} pairs of the batch per step (a bounded sample: the CPU path "
                   "needs ~20 ms per pair on all cores)", "protocol": "clearTarget;clearSource;setInputTarget;setInputSource;align"},
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
        # NVML every 50 ms and measurably slows kernel launches of the worker threads being timed
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


def c4_block(args, torch, dist, gorio, synth, rank, local_rank, world, dev, steps, with_roofline, peak, peak_src):
    """Config C4 on the N ranks of this run: `--roofline-points` source points vs as many target points; every rank holds
    both full clouds, the library cuts the cell-sorted source into interleaved chunks (covariances computed by chunks and
    all-gathered; update_correspondences / linearize / compute_error over the rank's chunks in ONE launch each; the 28 / 1
    sums exchanged inside the reduction kernels over NVLink peer memory). One step = one LM iteration's device work:
    linearize (= update_correspondences + H/b/err) + one compute_error. At N = 1 the same handle gives the roofline block."""
    sharding = importlib.import_module("go-rio_b200.sharding")
    n = args.roofline_points
    src, tgt, T = synth.tiled_cloud_pair(4000, n)
    dsrc, dtgt = torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev)
    del src, tgt
    g = gorio.FastAPDGICP(local_rank)
    g.set_params(max_correspondence_distance=2.0)
    if world > 1:
        sharding.init_comm(g, gorio.load(), rank, world, n, dist, dev, fused=not args.no_fused)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_setup = time.perf_counter()
    g.set_input_target_device(dtgt.data_ptr(), n)
    g.set_input_source_device(dsrc.data_ptr(), n)
    g.set_profiling(True)
    g.linearize(T)  # grids + covariances (+ all-gather) + the first linearisation
    torch.cuda.synchronize()
    k0 = g.kernel_ms()
    g.set_profiling(False)
    t_setup = torch.tensor([time.perf_counter() - t_setup], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_setup, op=dist.ReduceOp.MAX)
    # The pose moves between the steps as it does between the last iterations of an LM run (millimetres, also at the far
    # edge of the 2 km wide tiled area): each update_correspondences pass then starts from the previous pass's matches, as
    # it does inside align().
    walk = [T @ synth.make_pose([0.004 * a, -0.003 * b, 0.002 * c], [0.0, 1e-6 * b, 2e-6 * a])
            for a, b, c in ((1, 0, 1), (1, 1, 0), (0, 1, 1), (0, 0, 0))]
    for i in range(4):
        g.linearize(walk[i % 4])
        g.compute_error(walk[i % 4])
    l0 = g.launch_count()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        g.linearize(walk[i % 4])
        g.compute_error(walk[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = g.launch_count() - l0
    reps = max(1, args.roofline_reps)

    def profiled(poses):
        g.set_profiling(True)
        for i in range(reps):
            g.linearize(poses[i % len(poses)])
            g.compute_error(poses[i % len(poses)])
        kk = g.kernel_ms()
        g.set_profiling(False)
        return kk

    k = profiled(walk)                                   # the timed steps' poses (moves of ~5 mm)
    far = [T @ synth.make_pose([0.10, -0.06, 0.02], [0.0, 1e-5, 4e-5]), T]  # (the tiled area is 2 km wide: 4e-5 rad = 8 cm at its far edge)
    k_far = profiled(far)                                # early-iteration moves (~12-20 cm)
    k_same = profiled([T])                               # the same pose again
    err, H, bb = g.linearize(T)
    err2 = g.compute_error(T)
    lin_ms, err_ms, corr_ms = k["linearize"][0] / reps, k["error"][0] / reps, k["corr"][0] / reps
    corr_passes = {"cold_first_pass": k0["corr"][0], "pose_moved_12cm": k_far["corr"][0] / reps, "pose_moved_5mm": corr_ms,
                   "same_pose": k_same["corr"][0] / reps}
    n_local = -(-n // world)  # source points this rank serves
    block = {
        "workload": f"C4: {n} source vs {n} target points, source-sharded over {world} GPU(s) in interleaved chunks; covariances by chunks + "
                    "all-gather; per LM iteration ONE update_correspondences pass (search + Mahalanobis kernels), ONE linearize and ONE "
                    "compute_error launch per rank; all-reduce of 28 / 1 doubles "
                    + ("by ncclAllReduce" if (args.no_fused and world > 1) else ("inside the reduction kernels over NVLink peer memory" if world > 1 else "not needed (1 GPU)")),
        "ms_per_step": ms / steps, "steps": steps, "points_per_s": n * steps / (ms / 1e3), "scaling": "strong",
        "launches_per_step": launches / steps,
        "kernels_rank0_ms": {"update_correspondences": corr_ms, "linearize": lin_ms, "compute_error": err_ms},
        "update_correspondences_ms_by_motion": corr_passes,
        "poses": "the pose moves ~5 mm between steps (the last iterations of an LM run); err is checked at the ground-truth pose after the timed steps",
        "setup_ms": 1e3 * float(t_setup.item()),
        "setup": "grid builds + kNN covariances of both clouds (+ all-gather) + first linearize, wall clock, max over ranks",
        "err": err, "err_trial": err2,
        "err_expected": C4_ERR_20M if n == 20_000_000 else None,
        "err_equal_across_N": (abs(err - C4_ERR_20M) / C4_ERR_20M < 1e-12) if n == 20_000_000 else None,
        "timing": "CUDA events around the steps, max over ranks (inputs 1.28 GB per pass > L2)",
    }
    # ---- the whole registration through align() (the C++ host loop of large / sharded clouds: per outer iteration one
    # linearize, per LM trial one compute_error, the 6x6 solve on the host of every rank), from the identity guess ----
    try:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l1 = g.launch_count()
        t_al = time.perf_counter()
        res = g.align(np.eye(4))
        torch.cuda.synchronize()
        t_al = torch.tensor([time.perf_counter() - t_al], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_al, op=dist.ReduceOp.MAX)
        trials = int(g.lm_trace().shape[0])
        E = res["T64"] @ np.linalg.inv(T)
        block["align"] = {
            "guess": "identity (the clouds are ~0.5 m apart)", "converged": bool(res["converged"]), "outer_iterations": int(res["iterations"]),
            "lm_trials": trials, "ms": 1e3 * float(t_al.item()), "ms_per_outer_iteration": 1e3 * float(t_al.item()) / max(1, int(res["iterations"])),
            "launches": int(g.launch_count() - l1),
            "pose_error_vs_ground_truth_m_rad": [float(np.linalg.norm(E[:3, 3])), float(np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)))],
            "timing": "wall clock around apd_align (covariances and grids already built), max over ranks",
        }
    except Exception as e:
        block["align"] = {"error": str(e)}
    roof = None
    if with_roofline:
        achieved = BYTES_PER_POINT_LINEARIZE * n_local / 1e9 / (lin_ms / 1e3)
        n_valid = int((g.get_correspondences()[0] >= 0).sum())
        traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of the same kernel at the same size
        try:
            cap = json.load(open(os.path.join(REPO, "profiles", "r02_ncu_linearize.json")))
            if cap.get("points") == n:
                kk = cap["linearize_kernel<fp32 maha, H+b+err>"]
                traffic = kk["dram_bytes_read"] + kk["dram_bytes_write"]
        except Exception:
            pass
        roof = {
            "bound": "hbm", "kernel": "linearize_kernel<fp32 maha, H+b+err>", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/r02_ncu_linearize.json (ncu capture, bytes per launch)" if traffic else None,
            "algorithmic_bytes": BYTES_PER_POINT_LINEARIZE * n, "peak_source": peak_src, "points": n, "matched_points": n_valid,
            "bytes_per_point": BYTES_PER_POINT_LINEARIZE, "ms_per_launch": lin_ms,
            "compute_error": {"ms_per_launch": err_ms, "achieved": BYTES_PER_POINT_LINEARIZE * n / 1e9 / (err_ms / 1e3),
                              "frac": BYTES_PER_POINT_LINEARIZE * n / 1e9 / (err_ms / 1e3) / peak},
            "update_correspondences": {"ms_per_pass": corr_ms, "launches_per_pass": 2, "achieved": BYTES_PER_POINT_CORR * n / 1e9 / (corr_ms / 1e3),
                                       "frac": BYTES_PER_POINT_CORR * n / 1e9 / (corr_ms / 1e3) / peak, "bytes_per_point": BYTES_PER_POINT_CORR,
                                       "ms_by_motion_since_previous_pass": corr_passes,
                                       "note": "search kernel (fp32, index work; a match is kept without a search when the bound stored for every other target point proves it) "
                                               "+ Mahalanobis kernel (fp64, coalesced); the quoted pass follows a pass ~5 mm away, as inside an LM run"},
            "grid_build": {"ms_per_cloud": k0["grid"][0] / 2, "launches_per_cloud": k0["grid"][1] // 2},
            "knn_covariance": {"ms_per_cloud": k0["knn_cov"][0] / 2, "mqueries_per_s": n / 1e6 / (k0["knn_cov"][0] / 2 / 1e3)},
            "timing": "CUDA events on the handle's stream around each launch; working set 1.28 GB > L2",
        }
    if world > 1:
        g.comm_destroy()
    g.close()
    del dsrc, dtgt
    torch.cuda.empty_cache()
    return block, roof


def main():
    args = parse()
    # stdout carries exactly ONE line (the JSON): libraries that chat on fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    synth = importlib.import_module("go-rio_b200.synth")
    only_c4 = args.workload == "c4"
    host_pairs = []
    t_gen = time.perf_counter()
    if not args.roofline_only and not only_c4:
        host_pairs = make_pairs(synth, rank, world, args.pairs, args.distinct)  # (forks: before CUDA is touched)
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the registration path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gorio = importlib.import_module("go-rio_b200")
    gorio_mod = gorio

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    line = {"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world}
    if args.roofline_only or only_c4:
        line["note"] = "profiling run: only the large-cloud (C4 / roofline) part"
    else:
        n_mine = len(host_pairs)
        uniq = {}
        for s, t in host_pairs:
            uniq.setdefault(s.ctypes.data, (s, t))
        uniq = list(uniq.values())
        # --- e2e_packed inputs: packed float4 clouds in ONE page-locked buffer (the pool copies them without staging) ---
        total_f = sum(s.size + t.size for s, t in uniq)
        pinned = torch.empty(total_f, dtype=torch.float32).pin_memory()
        pinned_np = pinned.numpy()
        pin_of, off = {}, 0
        for s, t in uniq:
            ps = pinned_np[off:off + s.size].reshape(s.shape); off += s.size
            pt = pinned_np[off:off + t.size].reshape(t.shape); off += t.size
            ps[...] = s
            pt[...] = t
            pin_of[s.ctypes.data] = (ps, pt)
        pinned_pairs = [pin_of[s.ctypes.data] for s, _ in host_pairs]
        # --- e2e inputs: pageable 48-byte pcl::PointXYZINormal clouds, what the drop-in class hands over ---
        pcl_of = {s.ctypes.data: (synth.to_pcl_xyzinormal(s), synth.to_pcl_xyzinormal(t)) for s, t in uniq}
        pcl_pairs = [pcl_of[s.ctypes.data] for s, _ in host_pairs]
        # --- HBM-resident copies for `value` ---
        dev_all = pinned.to(dev, non_blocking=True)
        torch.cuda.synchronize()
        base_ptr = dev_all.data_ptr()
        host_base = pinned_np.ctypes.data
        dev_pairs = [((base_ptr + (ps.ctypes.data - host_base), ps.shape[0]), (base_ptr + (pt.ctypes.data - host_base), pt.shape[0]), None)
                     for ps, pt in pinned_pairs]
        # the batch context of the C-ABI (apd_batch_*): `--streams` registrations in flight, a few host threads
        batch = gorio.Batch(local_rank, n_workers=args.streams, **DEPLOYED)
        prep_dev = batch.prepare(dev_pairs)
        prep_packed = batch.prepare([(s, t, None) for s, t in pinned_pairs])
        prep_pcl = batch.prepare([(s, t, None) for s, t in pcl_pairs], layout=(48, 0, 16))
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(b, prepared, steps, warmup):
            """W untimed passes over the batch, then EXACTLY `steps` passes inside ONE timed region (barrier + synchronize
            on both sides, CUDA events, max over ranks). The passes go through the pool's queue back to back — the pool is
            not drained between steps, as a loop-closure service would run it; every pass copies / registers every pair
            again (nothing is cached across passes: each pair is its own clearTarget/clearSource/set/align). Inputs per
            pass (>= 512 MB per GPU) exceed the 126 MB L2, which is also flushed before the timed region."""
            if warmup > 0:
                b.align(b.repeat(prepared, warmup), with_fitness=False, parse=False)
            rep = b.repeat(prepared, steps)
            l0 = b.launch_count()
            flush.zero_()
            barrier()
            cpu0 = time.process_time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.align(rep, with_fitness=False, parse=False)  # returns when every pair's result is on the host
            e1.record()
            torch.cuda.synchronize()
            cpu_ms_per_pair = 1e3 * (time.process_time() - cpu0) / max(1, rep["n"])  # all threads of this process
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            launches = b.launch_count() - l0
            # every pass must have produced the same results
            v = b.results_view(rep["res"]).reshape(steps, prepared["n"])
            stable = bool(all(np.array_equal(v["T"][0], v["T"][j]) and np.array_equal(v["status"][0], v["status"][j]) for j in range(1, steps)))
            last = (gorio_mod.ApdResult * prepared["n"]).from_buffer(rep["res"], (steps - 1) * prepared["n"] * ctypes.sizeof(gorio_mod.ApdResult))
            return float(ms.item()), launches, gorio_mod._results(last), cpu_ms_per_pair, stable

        with ClockSampler(local_rank) as clocks:
            batch.load_stats(reset=True)
            ms_dev, launches, results, cpu_dev, st0 = timed(batch, prep_dev, args.steps, args.warmup)
            load = batch.load_stats(reset=True)  # the loop kernels' own %globaltimer stamps (warm-up + timed passes)
            ms_pcl, _, results_pcl, cpu_pcl, st1 = timed(batch, prep_pcl, args.steps, args.warmup)
            ms_packed, _, results_packed, cpu_packed, st2 = timed(batch, prep_packed, args.steps, args.warmup)
        value = args.pairs * args.steps / (ms_dev / 1e3)
        e2e_value = args.pairs * args.steps / (ms_pcl / 1e3)
        packed_value = args.pairs * args.steps / (ms_packed / 1e3)
        same = st0 and st1 and st2 and all(np.array_equal(a["T"], b["T"]) and np.array_equal(a["T"], c["T"]) for a, b, c in zip(results, results_pcl, results_packed))
        every = results + results_pcl + results_packed
        counts = torch.tensor([sum(r["status"] == 0 for r in every), sum(bool(r["converged"]) for r in every), len(every),
                               sum(r["iterations"] for r in results)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(counts)
        n_ok, n_conv, n_all, n_iter = (int(x) for x in counts.tolist())
        h2d_pcl = sum(s.nbytes + t.nbytes for s, t in pcl_pairs)
        h2d_packed = sum(s.nbytes + t.nbytes for s, t in pinned_pairs)
        d2h = n_mine * LM_RESULT_HEAD_BYTES

        # per-kernel-class device time of one step (profiling events on the worker streams; separate, untimed pass)
        batch.set_profiling(True)
        batch.align(prep_dev, with_fitness=False, parse=False)
        kms = {k: [ms, cnt] for k, (ms, cnt) in batch.kernel_ms().items()}
        batch.set_profiling(False)
        tot = sum(v[0] for v in kms.values()) or 1.0
        kernels = {k: {"ms_per_step": round(v[0], 4), "launches_per_step": v[1], "share": round(v[0] / tot, 4)} for k, v in kms.items()}
        batch.close()

        # --- the same with every target covariance computed up front, as the reference does (calculate_covariances over
        # all 60 000 target points per registration): separates the algorithmic part of the speed-up from the hardware part
        eager = None
        if not args.no_eager:
            os.environ["APD_LAZY_TARGET_COV"] = "0"
            b2 = gorio.Batch(local_rank, n_workers=args.streams, **DEPLOYED)
            del os.environ["APD_LAZY_TARGET_COV"]
            sub = batch_sub = b2.prepare(dev_pairs[: max(1, min(n_mine, 1024 // world))])
            ms_eager, _, res_eager, _, _ = timed(b2, sub, max(2, args.steps // 4), 1)
            eager_diff = [(i, float(np.abs(a["T"].astype(np.float64) - b["T"].astype(np.float64)).max()), a["iterations"], b["iterations"], a["status"], b["status"])
                          for i, (a, b) in enumerate(zip(results, res_eager)) if not np.array_equal(a["T"], b["T"])]
            eager_same = not eager_diff
            if eager_diff:
                print("eager != on-demand:", len(eager_diff), "of", len(res_eager), eager_diff[:8], file=sys.stderr)
            eager = {"value": world * batch_sub["n"] * max(2, args.steps // 4) / (ms_eager / 1e3), "unit": UNIT, "pairs_per_step": world * batch_sub["n"],
                     "same_poses_as_on_demand": bool(eager_same), "pairs_compared": len(res_eager), "pairs_that_differ": len(eager_diff),
                     "note": "APD_LAZY_TARGET_COV=0: all 60 000 target covariances per registration (the reference's work); the default "
                             "computes the ones the registration meets (~4 100), bit-identical outputs"}
            b2.close()

        # --- parity against the CPU restatement on the first pairs of the batch (rank 0; the same pairs the reference arm times) ---
        parity = None
        if rank == 0:
            try:
                k = min(8, n_mine)
                cpu = cpu_poses(host_pairs[:k])
                dT = max(float(np.abs(results_pcl[i]["T"].astype(np.float64) - cpu[i][0].astype(np.float64)).max()) for i in range(k))
                ref_code = None
                if os.path.exists(REF_CODE_SO):  # ... and against the reference's own code (oracle/_ref) on the same pairs
                    rl = ctypes.CDLL(REF_CODE_SO)
                    rl.aref_create.restype = ctypes.c_void_p
                    _, _, rp = ref_code_registrations(rl, host_pairs[:k], 1, want_poses=True)
                    ref_code = {"max_abs_dT": max(float(np.abs(results_pcl[i]["T"].astype(np.float64) - rp[i][0].astype(np.float64)).max()) for i in range(k)),
                                "same_converged_and_iterations": all(bool(results_pcl[i]["converged"]) == rp[i][1] and results_pcl[i]["iterations"] == rp[i][2] for i in range(k)),
                                "what": "oracle/_ref/libapd_ref_apdgicp_nf.so: the reference's own headers compiled in place (stand-in Eigen / PCL, nanoflann kd-tree)"}
                parity = {"pairs": k, "max_abs_dT": dT, "vs_reference_code": ref_code,
                          "same_converged_and_iterations": all(bool(results_pcl[i]["converged"]) == bool(cpu[i][1]) and results_pcl[i]["iterations"] == cpu[i][2] for i in range(k)),
                          "against": "the CPU restatement (oracle), the first pairs of the batch — the pairs the reference arm times",
                          "bar": "float poses; fp32 Mahalanobis storage: within north_star's 1e-5 (the fp64 bar, 1e-6 m / 1e-6 rad, is held in tests/)"}
            except Exception as e:  # the checker is optional here (tests/ hold the parity bar)
                parity = {"error": str(e)}

        line.update({
            "value": value, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": args.pairs, "pairs_per_step_per_gpu": n_mine,
                       "distinct_scenes": args.pairs if args.distinct <= 0 else min(args.distinct, args.pairs),
                       "streams_per_gpu": args.streams, "api": "apd_batch_align_device (value) / apd_batch_align (e2e)",
                       "optimizer_loop": "device-resident (lm.cu), one launch per registration or per two (a device runs at most 128 grids at a time: two ready registrations share a launch)",
                       "mahalanobis_storage": "fp32 (6 x 4 B per point; arithmetic fp64; within north_star's 1e-5, tests/test_gpu_parity.py)",
                       "protocol": "clearTarget;clearSource;setInputTarget;setInputSource;align",
                       "l2": "inputs larger than L2 (>= 512 MB of clouds per GPU per step against 126 MB); flushed (256 MiB write) before the timed region",
                       "timed_region": "the K steps go through the pool's queue back to back inside ONE region bracketed by barrier + synchronize (no drain between steps); every step copies / registers every pair again",
                       "sharding": "pair i -> rank i mod N, no collective; every pair is a different scene (heterogeneous iteration counts)",
                       "scene_generation_s": round(t_gen, 1)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_pcl, "d2h_bytes_per_step": d2h,
                    "layout": "pageable 48-byte pcl::PointXYZINormal AoS (staged to pinned float4 by the pool, then copied)",
                    "same_result_as_device_resident": bool(same), "all_pairs_ok": bool(n_ok == n_all), "pairs_converged": n_conv, "pairs_run": n_all,
                    "api": "apd_batch_align (host AoS clouds in, poses out)",
                    "host_cpu_ms_per_pair": round(cpu_pcl, 4), "host_cores": host_cores(),
                    "host_staging_ceiling": round(host_cores() / max(cpu_pcl, 1e-6) * 1e3, 0),
                    "host_note": "reading 2.9 MB of pageable 48-byte records per pair is host-core work (go-rio_b200/csrc/host_stage.hpp): the node's "
                                 "cores stage host_staging_ceiling pairs/s whatever the number of GPUs; e2e_packed is the same run from packed pinned clouds"},
            "e2e_packed": {"value": packed_value, "unit": UNIT, "h2d_bytes_per_step": h2d_packed, "d2h_bytes_per_step": d2h,
                           "layout": "packed float4 {x,y,z,label} in page-locked memory (no staging pass)"},
            "eager": eager,
            "parity_vs_cpu": parity,
            "mean_outer_iterations": n_iter / max(1, args.pairs),
            "gpu_launches": int(launches),
            "host_cpu_ms_per_registration": {"device_resident": round(cpu_dev, 4), "e2e": round(cpu_pcl, 4), "e2e_packed": round(cpu_packed, 4),
                                             "host_cores": host_cores(),
                                             "note": "process CPU time of rank 0 (all worker threads) inside the timed region, per registration"},
            "kernels": kernels,
            "clocks": clocks.summary(),
        })
        # the step's dominant kernel: the device-resident loop. It is latency / issue bound (0.04 % of DRAM peak), so its
        # live figure is how full it keeps the GPU; the ncu figures of the same kernel are in profiles/
        lm_ms = kms.get("lm", [0.0, 0])[0]
        ctas = 2 if "APD_LM_CLUSTER" not in os.environ else int(os.environ["APD_LM_CLUSTER"])  # (a pool's default cluster size)
        under_load_ms = load["lm_kernel_ms"] / max(1.0, load["registrations"])  # mean duration of a loop kernel while the pool is full
        line["loop_kernel"] = {"kernel": "lm_kernel<fp32 maha, 2 CTAs/SM>", "share_of_step_kernel_time": round(lm_ms / tot, 4),
                               "ms_per_registration_under_load": under_load_ms, "ms_per_registration_profiled_pass": lm_ms / max(1, n_mine),
                               "ctas_per_registration": ctas,
                               "cta_slot_occupancy": round(under_load_ms * ctas * n_mine / (296.0 * (ms_dev / args.steps)), 4),
                               "note": "CTA-slot occupancy = mean loop-kernel duration under load (the kernels' own time stamps during the timed passes) x CTAs "
                                       "x registrations per step / (148 SMs x 2 slots x step time); the profiled pass runs with timing events and less "
                                       "concurrency; issue-slot %, L2 hit rate, stalls: profiles/r02_ncu_lm_kernel.txt"}
        del dev_all, flush
        torch.cuda.empty_cache()

    # ---- config C4 on the ranks of this run; at N = 1 also the roofline of the linearize kernel ----
    if not args.no_c4 and not (args.no_roofline and world == 1 and not only_c4):
        block, roof = c4_block(args, torch, dist, gorio, synth, rank, local_rank, world, dev, max(5, args.steps), world == 1, peak, peak_src)
        line["c4"] = block
        if roof is not None:
            line["roofline"] = roof
        if only_c4 or args.roofline_only:
            line.update({"metric": "APDGICP LM-iteration throughput on one large cloud (source points/s)", "value": block["points_per_s"], "unit": "points/s",
                         "ms_per_step": block["ms_per_step"], "steps": block["steps"], "scaling": "strong", "higher_is_better": True})

    # ---- config C5: the scan-to-scan odometry front end over a synthetic drive, one handle, frames in sequence (N = 1) ----
    if rank == 0 and world == 1 and not args.no_replay and host_pairs:
        try:
            sys.path.insert(0, os.path.join(REPO, "profiles"))
            import replay_bench
            line["replay"] = replay_bench.run(args.replay_frames, min(300, args.replay_frames), 1000, local_rank)
        except Exception as e:
            line["replay"] = {"error": str(e)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline and host_pairs:
        lib, kind, search, what = load_cpu_impl()
        sample = host_pairs[:8]
        t1, n1 = cpu_registrations(lib, search, sample[:1], 1)
        reps = max(1, min(50, int(round(12.0 / max(t1 * len(sample), 1e-3)))))  # ~12 s of CPU work
        t_cpu, n_cpu = cpu_registrations(lib, search, sample, reps)
        line["cpu_baseline"] = {"value": n_cpu / t_cpu, "unit": UNIT, "cores": host_cores(), "kind": kind,
                                "sample": f"the first {len(sample)} pairs of the batch x {reps} repetitions ({t_cpu:.1f} s wall); {what}"}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
