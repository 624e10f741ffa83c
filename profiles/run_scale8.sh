for n in 8; do
APD_POLL_WAIT_US=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/c2_n$n.err | tail -1 > gpurun_out/c2_n${n}_poll.json
python - <<PY
import json
d=json.load(open("gpurun_out/c2_n${n}_poll.json")); print("c2 poll", $n, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["host_cpu_ms_per_registration"])
PY
done
nvidia-smi topo -m | head -14
lscpu | grep -i "numa\|socket\|model name" | head
