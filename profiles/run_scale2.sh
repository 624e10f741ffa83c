# 2 x B200: the N > 1 GPU tests, the C2 pool on both ranks, the sharded C4 step
timeout 600 python -m pytest tests -m gpu -x -q -k "shard or nccl or comm or two or rank" 2>&1 | tail -4
for n in 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>gpurun_out/c2_n$n.err | tail -1 > gpurun_out/c2_n${n}.json
python - <<PY
import json
d=json.load(open("gpurun_out/c2_n${n}.json")); print("c2", $n, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["host_cpu_ms_per_registration"], d["clocks"])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 bench.py --workload c4 --gpus $n --steps 10 --warmup 3 2>gpurun_out/c4_n$n.err | tail -1 > gpurun_out/c4_n${n}.json
cut -c1-700 gpurun_out/c4_n${n}.json
done
nproc
