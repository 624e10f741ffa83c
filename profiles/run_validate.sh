timeout 600 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -15
for f in "" "--no-fused"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c4 --gpus 2 --steps 20 --warmup 3 $f 2>gpurun_out/c4_n2.err | tail -1 > gpurun_out/c4_n2$f.json
python - <<PY
import json
d=json.load(open("gpurun_out/c4_n2$f.json")); print("$f", d["n_gpus"], round(d["ms_per_step"],3), round(d["setup_ms"],1), d["err"], d["gpu_launches"], d["kernels_rank0_ms"])
PY
done
