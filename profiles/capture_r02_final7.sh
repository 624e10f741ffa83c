# End-of-round evidence for the final build (submap grid at 2 cells per point under on-demand covariances): two more probes
# (one CTA per registration; scan grid at 0.5 cells per point), the launch list of a short bench command, and ncu --set full
# of the loop kernel under the pool's residency. Each ncu pass runs only after the same command has exited 0 without ncu.
set -x
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe48.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-120 >> gpurun_out/r02_probe48.txt; }
: > gpurun_out/r02_probe48.txt
run APD_NOP=1
run APD_CELLS_PER_POINT_SMALL=0.5
run APD_NOP=1
run APD_CELLS_PER_POINT_SMALL=0.5
echo "== APD_LM_CLUSTER=1 --streams 384" >> gpurun_out/r02_probe48.txt
APD_LM_CLUSTER=1 timeout 300 $P --streams 384 2>&1 | cut -c1-120 >> gpurun_out/r02_probe48.txt
cat gpurun_out/r02_probe48.txt
SHORT="python bench.py --steps 1 --warmup 1 --pairs 64 --streams 16 --no-cpu-baseline --no-eager --no-replay --roofline-reps 2"
timeout 600 $SHORT > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final7.csv $SHORT > /dev/null 2>&1
export APD_LM_CLUSTER=2 APD_LM_MINB=2
timeout 300 python profiles/multi_lm.py --jobs 148 --repeat 2 > /dev/null && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 1 -c 1 -o gpurun_out/prof_lm_final7_r02 -f python profiles/multi_lm.py --jobs 148 --repeat 1 > gpurun_out/ncu_lm_final7.log 2>&1
ls -la gpurun_out/
