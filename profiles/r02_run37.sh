# pass 37 (1 GPU): end to end from pageable PCL clouds — registrations in flight and host threads
B="python bench.py --steps 10 --warmup 3 --no-roofline --no-c4 --no-eager --no-replay --no-cpu-baseline"
run() { echo "== $*" >> gpurun_out/r02_probe37.txt; env "$@" timeout 400 $B $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['e2e']['value']), d['e2e']['host_cpu_ms_per_pair'], round(d['e2e_packed']['value']), d['loop_kernel']['cta_slot_occupancy'])" >> gpurun_out/r02_probe37.txt; }
: > gpurun_out/r02_probe37.txt
EXTRA="--streams 256" run APD_NOP=1
EXTRA="--streams 384" run APD_NOP=1
EXTRA="--streams 256" run APD_BATCH_THREADS=24
EXTRA="--streams 256" run APD_BATCH_THREADS=12
cat gpurun_out/r02_probe37.txt; nproc; lscpu | grep -i "model name\|numa\|socket\|thread"
