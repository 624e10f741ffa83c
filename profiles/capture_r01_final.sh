# Round-1 final evidence: plain bench (the numbers), then the ncu launch list of a short bench run, then ncu --set full
# of the device-resident optimizer loop (C2 shapes, target covariances on demand as in the pool) and of the linearize /
# compute_error kernels (20 M points). Every ncu pass runs only after the same command has exited 0 without ncu.
# Numbers printed under ncu are not used.
set -x
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2>/dev/null; echo ref rc=$?
SHORT="python bench.py --steps 1 --warmup 1 --pairs 8 --streams 2 --no-cpu-baseline --roofline-points 20000000 --roofline-reps 2"
timeout 600 $SHORT > gpurun_out/short_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01_final.csv $SHORT > gpurun_out/ncu_launch_final.log 2>&1
echo launches rc=$?
export APD_LAZY_TARGET_COV=1 APD_LM_CLUSTER=4
K2="python profiles/kbench.py --mode c2 --reps 2"
timeout 300 $K2 > gpurun_out/kbench_c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 3 -c 1 -o gpurun_out/prof_lm_r01 -f $K2 > gpurun_out/ncu_lm.log 2>&1
echo lm rc=$?
unset APD_LAZY_TARGET_COV APD_LM_CLUSTER
timeout 300 python profiles/kbench.py --mode c1 --reps 50 > gpurun_out/kbench_c1.log 2>&1
timeout 300 python profiles/kbench.py --mode c2 --reps 50 > gpurun_out/kbench_c2.log 2>&1
timeout 300 python profiles/replay_bench.py > gpurun_out/replay_c5.json 2> gpurun_out/replay_c5.err
cat gpurun_out/kbench_c1.log gpurun_out/kbench_c2.log gpurun_out/replay_c5.json
cat gpurun_out/bench_r01_final.json gpurun_out/bench_r01_reference.json
