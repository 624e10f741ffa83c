# pass 49 (2 GPUs): the 2-GPU tests and the bench line as the driver launches it (final build: submap grid at 2 cells per point under on-demand covariances)
timeout 600 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -3
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_final2.json 2> gpurun_out/r02_bench_n2_final2.err; echo rc=$? wall=${SECONDS}s
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_final2.json').read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["e2e_packed"]["value"]), d["e2e"]["host_cpu_ms_per_pair"], d["c4"]["ms_per_step"], d["c4"]["align"]["ms"], d["c4"]["err_equal_across_N"], d["loop_kernel"]["cta_slot_occupancy"])
P
