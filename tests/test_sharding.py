"""N > 1 paths. CPU: the partition arithmetic and, with world_size-2 gloo, the
source-sharded linearize identity (sum over shards of H, b == H, b of the whole
cloud) driven through the oracle. GPU (needs >= 2 devices, else skipped): the same
through the CUDA library with its NCCL all-reduce, launched with torchrun."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sharding = importlib.import_module("go-rio_b200.sharding")


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1000, 20_000_001):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_pairs_cover_all():
    for n in (0, 5, 4096):
        for world in (1, 2, 8):
            got = sorted(i for r in range(world) for i in sharding.shard_pairs(n, r, world))
            assert got == list(range(n))


def test_bench_shards_distinct_scenes_by_pair_index(synth, monkeypatch):
    """config C3 as bench.py runs it: 4096 pairs, pair i = scene 3000 + i, pair i -> rank i mod N; every scene is generated
    once, by the rank that registers it"""
    import bench
    calls = []

    def fake_pair(seed, **kw):
        calls.append(seed)
        return np.full((2, 4), seed, np.float32), np.full((3, 4), seed, np.float32), None

    monkeypatch.setattr(synth, "submap_pair", fake_pair)
    world, total = 8, 64
    seen = []
    for rank in range(world):
        pairs = bench.make_pairs(synth, rank, world, total, procs=1)
        seeds = [int(s[0, 0]) for s, _ in pairs]
        assert seeds == [3000 + i for i in range(rank, total, world)]
        seen += seeds
    assert sorted(seen) == [3000 + i for i in range(total)] and sorted(calls) == sorted(seen)
    # fewer distinct scenes (a profiling aid) are cycled
    pairs = bench.make_pairs(synth, 0, 1, 16, distinct=4, procs=1)
    assert [int(s[0, 0]) for s, _ in pairs] == [3000 + i % 4 for i in range(16)]


_WORKER = r'''
import importlib, json, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["APD_REPO"]); sys.path.insert(0, os.path.join(os.environ["APD_REPO"], "tests"))
from oracle_binding import Oracle
sharding = importlib.import_module("go-rio_b200.sharding")
synth = importlib.import_module("go-rio_b200.synth")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
src, tgt, T = synth.submap_pair(2003, n_source=900, n_frames=4, n_per_frame=1000)
src[:, 3] = -1.0                      # no label matches: cl_weight (1/N, N = shard size in the oracle) drops out
full = Oracle(search=1); full.set_params(max_correspondence_distance=2.0)
full.set_input_target(tgt); full.set_input_source(src)
e_full, H_full, b_full = full.linearize(T)
cs = full.get_source_covariances()
b, e = sharding.shard_range(src.shape[0], rank, world)
part = Oracle(search=1); part.set_params(max_correspondence_distance=2.0)
part.set_input_target(tgt); part.set_input_source(src[b:e].copy()); part.set_source_covariances(cs[b:e])
e_p, H_p, b_p = part.linearize(T)
buf = torch.tensor(np.concatenate([H_p.reshape(-1), b_p, [e_p]]))
dist.all_reduce(buf)                  # the 28-value exchange of the sharded path (here 36+6+1 for simplicity)
got = buf.numpy()
ok = (np.abs(got[:36].reshape(6, 6) - H_full).max() / np.abs(H_full).max() < 1e-12 and
      np.abs(got[36:42] - b_full).max() / np.abs(b_full).max() < 1e-10 and abs(got[42] - e_full) / e_full < 1e-12)
if rank == 0:
    print(json.dumps({"ok": bool(ok), "world": world}))
dist.destroy_process_group()
'''


def _torchrun(script_text, tmp_path, nproc, extra_env=None, timeout=600):
    script = tmp_path / "worker.py"
    script.write_text(script_text)
    env = {**os.environ, "APD_REPO": REPO, **(extra_env or {})}
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script)], capture_output=True, text=True, env=env, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_source_sharded_linearize_gloo(tmp_path):
    r = _torchrun(_WORKER, tmp_path, 2)
    assert r["ok"] and r["world"] == 2


_GPU_WORKER = r'''
import ctypes, importlib, json, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["APD_REPO"])
gorio = importlib.import_module("go-rio_b200")
sharding = importlib.import_module("go-rio_b200.sharding")
synth = importlib.import_module("go-rio_b200.synth")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
src, tgt, T = synth.tiled_cloud_pair(4001, 300_000)
kw = dict(max_correspondence_distance=2.0, maha_fp64=1)
full = gorio.FastAPDGICP(local); full.set_params(**kw)
full.set_input_target(tgt); full.set_input_source(src)
e_full, H_full, b_full = full.linearize(T)
part = gorio.FastAPDGICP(local); part.set_params(**kw)
sharding.init_comm(part, gorio.load(), rank, world, src.shape[0], dist, torch.device("cuda", local), fused=os.environ["APD_FUSED"] == "1")
part.set_input_target(tgt); part.set_input_source(src)   # the same full clouds on every rank
e_p, H_p, b_p = part.linearize(T)           # covariance slices all-gathered, H/b/err all-reduced inside the library
cov_ok = bool(np.array_equal(part.get_target_covariances(), full.get_target_covariances()) and
              np.array_equal(part.get_source_covariances(), full.get_source_covariances()))
c_full, _ = full.get_correspondences(); c_part, _ = part.get_correspondences()
mine = c_part >= 0
counts = torch.tensor([int(mine.sum()), int((c_full >= 0).sum())], device="cuda")
dist.all_reduce(counts[:1])
corr_ok = bool(np.array_equal(c_part[mine], c_full[mine]) and int(counts[0]) == int(counts[1]))
e2_full, e2_p = full.compute_error(T), part.compute_error(T)
r_full = full.align()
r_part = part.align()                        # every rank runs the same LM decisions on the reduced values
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
res = {"H": rel(H_p, H_full), "b": rel(b_p, b_full), "err": abs(e_p - e_full) / e_full, "err2": abs(e2_p - e2_full) / e2_full,
       "pose": float(np.abs(r_part["T64"] - r_full["T64"]).max()), "iters": [r_part["iterations"], r_full["iterations"]],
       "conv": [r_part["converged"], r_full["converged"]], "cov_ok": cov_ok, "corr_ok": corr_ok}
poses = [None] * world
dist.all_gather_object(poses, r_part["T64"].tolist())
res["same_pose_on_all_ranks"] = all(p == poses[0] for p in poses)
part.comm_destroy()
if rank == 0:
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.gpu
@pytest.mark.parametrize("fused", ["0", "1"])
def test_sharded_registration_two_gpus(tmp_path, fused):
    """fused = 0: ncclAllReduce after the reduction kernels; 1: the exchange over NVLink peer memory inside them"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(_GPU_WORKER, tmp_path, 2, extra_env={"APD_FUSED": fused})
    assert r["H"] < 1e-11 and r["b"] < 1e-9 and r["err"] < 1e-11 and r["err2"] < 1e-11, r
    assert r["pose"] < 1e-9 and r["iters"][0] == r["iters"][1] and r["conv"][0] == r["conv"][1], r
    assert r["same_pose_on_all_ranks"] and r["cov_ok"] and r["corr_ok"], r


# ---- the sharded arithmetic on ONE GPU: the ranks are handles of this process (apd_group_*), exchanging through the
# same PeerMailbox code as the multi-process path (plain pointers instead of cudaIpc mappings) --------------------------
@pytest.fixture(scope="module")
def tiled300k(synth):
    return synth.tiled_cloud_pair(4001, 300_000)


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_sharded_group_on_one_gpu(gorio, tiled300k, nranks):
    """2-4 ranks on device 0 against the unsharded handle: the chunk table of the one-launch kernels, the in-kernel
    exchange of the 28 / 1 sums, the peer-to-peer all-gather of the covariance chunks, the redundant LM loop"""
    src, tgt, T = tiled300k
    kw = dict(max_correspondence_distance=2.0, maha_fp64=1)
    full = gorio.FastAPDGICP(0)
    full.set_params(**kw)
    full.set_input_target(tgt)
    full.set_input_source(src)
    e_full, H_full, b_full = full.linearize(T)
    grp = gorio.Group([0] * nranks, **kw)
    grp.set_input_target(tgt)
    grp.set_input_source(src)
    e_p, H_p, b_p = grp.linearize(T)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert rel(H_p, H_full) < 1e-11 and rel(b_p, b_full) < 1e-9 and abs(e_p - e_full) / e_full < 1e-11
    c_full, _ = full.get_correspondences()
    owned = np.zeros(src.shape[0], bool)
    for r in grp.ranks:  # every rank holds ALL covariances (gathered) and its own chunks of the correspondences
        assert np.array_equal(r.get_target_covariances(), full.get_target_covariances())
        assert np.array_equal(r.get_source_covariances(), full.get_source_covariances())
        c_r, _ = r.get_correspondences()
        mine = c_r >= 0
        assert np.array_equal(c_r[mine], c_full[mine]) and not (owned & mine).any()
        owned |= mine
    assert np.array_equal(owned, c_full >= 0)
    assert abs(grp.compute_error(T) - full.compute_error(T)) / e_full < 1e-11
    # a second pass (warm-started searches) and a different pose
    T2 = T @ synth_pose()
    e2_p, H2_p, _ = grp.linearize(T2)
    e2_f, H2_f, _ = full.linearize(T2)
    assert abs(e2_p - e2_f) / e2_f < 1e-11 and rel(H2_p, H2_f) < 1e-11
    r_full, r_part = full.align(), grp.align()
    assert np.abs(r_part["T64"] - r_full["T64"]).max() < 1e-9
    assert r_part["iterations"] == r_full["iterations"] and r_part["converged"] == r_full["converged"]
    traces = [r.lm_trace() for r in grp.ranks]
    assert all(np.array_equal(t, traces[0]) for t in traces)  # identical LM decisions from bit-identical totals on every rank
    grp.close()
    full.close()


def synth_pose():
    synth = importlib.import_module("go-rio_b200.synth")
    return synth.make_pose([0.03, -0.02, 0.01], [0.001, -0.002, 0.003])


@pytest.mark.gpu
def test_sharded_group_against_the_oracle(gorio, tiled300k):
    """the sharded sums at 300 k points against the CPU oracle (H, b, err of one linearisation; bit-exact correspondences)"""
    from oracle_binding import Oracle
    src, tgt, T = tiled300k
    kw = dict(max_correspondence_distance=2.0, maha_fp64=1)
    o = Oracle(search=1)
    o.set_params(**kw)
    o.set_input_target(tgt)
    o.set_input_source(src)
    e_o, H_o, b_o = o.linearize(T)
    c_o, _ = o.get_correspondences()
    grp = gorio.Group([0, 0], **kw)
    grp.set_input_target(tgt)
    grp.set_input_source(src)
    e_p, H_p, b_p = grp.linearize(T)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert abs(e_p - e_o) / e_o < 1e-10 and rel(H_p, H_o) < 1e-10 and rel(b_p, b_o) < 1e-9
    got = np.full(src.shape[0], -1, np.int32)
    for r in grp.ranks:
        c_r, _ = r.get_correspondences()
        got[c_r >= 0] = c_r[c_r >= 0]
    assert np.array_equal(got, c_o)
    grp.close()


@pytest.mark.gpu
def test_group_small_clouds_and_odd_sizes(gorio, synth):
    """chunks shorter than a tile, empty chunks (n < ranks x 4 x 256) and a rank count that does not divide anything"""
    for case, nranks in ((0, 2), (1, 3), (2, 4)):
        src, tgt, T = (synth.scan_pair(1001, 1000), synth.scan_pair(1002, 777), synth.submap_pair(2001, n_source=5000))[case]
        kw = dict(max_correspondence_distance=2.0, maha_fp64=1, host_loop=1)
        full = gorio.FastAPDGICP(0)
        full.set_params(**kw)
        full.set_input_target(tgt)
        full.set_input_source(src)
        grp = gorio.Group([0] * nranks, **kw)
        grp.set_input_target(tgt)
        grp.set_input_source(src)
        e_f, H_f, b_f = full.linearize(np.eye(4))
        e_p, H_p, b_p = grp.linearize(np.eye(4))
        assert abs(e_p - e_f) / e_f < 1e-11 and np.abs(H_p - H_f).max() / np.abs(H_f).max() < 1e-11
        r_f, r_p = full.align(), grp.align()
        assert np.abs(r_p["T64"] - r_f["T64"]).max() < 1e-9 and r_p["iterations"] == r_f["iterations"]
        grp.close()
        full.close()
