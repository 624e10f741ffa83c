# pass 35 (1 GPU): the bench line with two registrations per launch
SECONDS=0
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench35.json 2> gpurun_out/r02_bench35.err; echo bench rc=$? wall=${SECONDS}s
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench35.json').read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["host_cpu_ms_per_pair"], d["e2e_packed"]["value"], d["eager"]["value"], d["gpu_launches"], d["loop_kernel"], d["parity_vs_cpu"]["max_abs_dT"], d["e2e"]["all_pairs_ok"], d["e2e"]["same_result_as_device_resident"])
P
tail -3 gpurun_out/r02_bench35.err
