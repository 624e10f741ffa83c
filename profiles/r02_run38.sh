# pass 38 (1 GPU): where the loop kernel's time goes under the pool's load, by phase (a -DAPD_LM_PHASE_TIMING build; CTA 0's clock)
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
APD_LIB=$PWD/go-rio_b200/libapdgicp_phase.so timeout 300 $P 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['registrations_per_s'], d['lm_kernel_ms_mean_under_load'], json.dumps(d.get('lm_phase_ms_per_registration')))" | tee gpurun_out/r02_probe38.txt
