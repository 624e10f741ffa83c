timeout 900 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --workload c4 --steps 10 --warmup 2 2>/dev/null | tail -1 > gpurun_out/c4_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c4 --gpus 2 --steps 10 --warmup 2 2>gpurun_out/c4_n2.err | tail -1 > gpurun_out/c4_n2.json
tail -3 gpurun_out/c4_n2.err
python - <<'PY'
import json
for f in ("c4_n1","c4_n2"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["n_gpus"], round(d["ms_per_step"],3), round(d["setup_ms"],1), d["err"], d["kernels_rank0_ms"])
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
