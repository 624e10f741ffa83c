set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "on_demand or align_parity or swap or promotion" 2>&1 | tail -15
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline"
APD_LAZY_TARGET_COV=0 timeout 300 $B > gpurun_out/lazy0.json 2> gpurun_out/lazy0.err; echo rc=$?
timeout 300 $B > gpurun_out/lazy_auto.json 2> gpurun_out/lazy_auto.err; echo rc=$?
cat gpurun_out/lazy0.json gpurun_out/lazy_auto.json
timeout 300 python profiles/kbench.py --mode c2 --reps 50 > gpurun_out/kbench_c2_lazy.log 2>&1; tail -20 gpurun_out/kbench_c2_lazy.log
APD_LAZY_TARGET_COV=0 timeout 300 python profiles/kbench.py --mode c2 --reps 50 > gpurun_out/kbench_c2_lazy0.log 2>&1; tail -20 gpurun_out/kbench_c2_lazy0.log
