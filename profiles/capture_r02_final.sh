# End-of-round capture (run under gpurun, one GPU). Each ncu pass runs only after the same command has exited 0 without
# ncu; numbers printed by a run under ncu are never used as bench values.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2>/dev/null; echo ref rc=$?
# launch list of a (short) bench command: the kernels' shares of a step
SHORT="python bench.py --steps 1 --warmup 1 --pairs 64 --streams 16 --no-cpu-baseline --no-eager --no-replay --roofline-reps 2"
timeout 600 $SHORT > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv $SHORT > /dev/null 2>&1
# the loop kernel under the pool's residency, and the large-cloud kernels
export APD_LM_CLUSTER=2 APD_LM_MINB=2
timeout 300 python profiles/multi_lm.py --jobs 148 --repeat 2 > /dev/null && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 1 -c 1 -o gpurun_out/prof_lm_final_r02 -f python profiles/multi_lm.py --jobs 148 --repeat 1 > gpurun_out/ncu_lm_final.log 2>&1
unset APD_LM_CLUSTER APD_LM_MINB
C4="python bench.py --workload c4 --steps 3 --roofline-reps 3 --no-cpu-baseline"
timeout 600 $C4 > /dev/null 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"corr_search|maha_kernel|linearize_kernel" -s 12 -c 4 -o gpurun_out/prof_c4_final_r02 -f $C4 > gpurun_out/ncu_c4_final.log 2>&1
cut -c1-1500 gpurun_out/bench_r02_final.json
