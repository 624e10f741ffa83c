# 8 x B200: the C2 pool on every rank (weak scaling, pairs sharded across ranks, no collective).
# (The sharded C4 step at N = 8 was taken with the same build earlier in the round: 0.70 ms per step, see r01_scaling.md.)
n=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 6 --warmup 3 --no-cpu-baseline --no-roofline 2>gpurun_out/c2_n$n.err | tail -1 > gpurun_out/c2_n${n}.json
python - <<PY
import json
d=json.load(open("gpurun_out/c2_n${n}.json")); print("c2", $n, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["host_cpu_ms_per_registration"], d["clocks"])
PY
nproc
