# pass 36 (1 GPU): end to end from pageable PCL clouds — non-temporal stores in the staging pass, one / two registrations per launch
B="python bench.py --steps 10 --warmup 3 --no-roofline --no-c4 --no-eager --no-replay --no-cpu-baseline"
run() { echo "== $*" >> gpurun_out/r02_probe36.txt; env "$@" timeout 400 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['e2e']['value']), d['e2e']['host_cpu_ms_per_pair'], round(d['e2e_packed']['value']), d['loop_kernel']['cta_slot_occupancy'])" >> gpurun_out/r02_probe36.txt; }
: > gpurun_out/r02_probe36.txt
run APD_STAGE_NT=0 APD_PAIR_HOLD_US=200
run APD_STAGE_NT=1 APD_PAIR_HOLD_US=200
run APD_STAGE_NT=0 APD_PAIR_HOLD_US=0
run APD_STAGE_NT=1 APD_PAIR_HOLD_US=0
run APD_STAGE_NT=1 APD_PAIR_HOLD_US=50
cat gpurun_out/r02_probe36.txt
