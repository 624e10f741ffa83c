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


def test_bench_gives_every_rank_the_same_work(synth, monkeypatch):
    """weak scaling = equal work per GPU: every rank of bench.py registers the same scenes, starting at a different one"""
    import bench
    calls = []

    def fake_pair(seed, **kw):
        calls.append(seed)
        return np.full((2, 4), seed, np.float32), np.full((3, 4), seed, np.float32), None

    monkeypatch.setattr(synth, "submap_pair", fake_pair)
    per_rank = []
    for rank in range(8):
        pairs = bench.make_pairs(synth, rank, 512)
        per_rank.append(sorted(int(s[0, 0]) for s, _ in pairs))
        assert int(pairs[0][0][0, 0]) == 2000 + rank  # rotated start
    assert all(p == per_rank[0] for p in per_rank) and set(per_rank[0]) == set(range(2000, 2008))
    assert sorted(set(calls)) == list(range(2000, 2008))


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
