# pass 44 (8 GPUs): the bench line as the driver launches it at N = 8 (final build: two registrations per launch, non-temporal staging)
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_final2.json 2> gpurun_out/r02_bench_n8_final2.err; echo rc=$? wall=${SECONDS}s
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n8_final2.json').read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["e2e_packed"]["value"]), d["e2e"]["host_cpu_ms_per_pair"], d["c4"]["ms_per_step"], d["c4"]["align"]["ms"], d["c4"]["err_equal_across_N"], d["loop_kernel"]["cta_slot_occupancy"])
P
nproc
