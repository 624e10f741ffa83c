# End-of-round capture, fourth part (final build: two registrations per launch, 32-bit stand-in sort in the kNN flush,
# non-temporal staging): the whole suite, smoke, both bench arms as the driver runs them, the launch list of a short bench
# command, and ncu of the loop kernel under the pool's residency. Each ncu pass runs only after the same command has exited
# 0 without ncu; numbers printed by a run under ncu are never used as bench values.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
SECONDS=0
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_reference4.json 2>/dev/null; echo ref rc=$? wall=${SECONDS}s
SECONDS=0
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final4.json 2> gpurun_out/bench_r02_final4.err; echo bench rc=$? wall=${SECONDS}s
SHORT="python bench.py --steps 1 --warmup 1 --pairs 64 --streams 16 --no-cpu-baseline --no-eager --no-replay --roofline-reps 2"
timeout 600 $SHORT > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final4.csv $SHORT > /dev/null 2>&1
export APD_LM_CLUSTER=2 APD_LM_MINB=2
timeout 300 python profiles/multi_lm.py --jobs 148 --repeat 2 > /dev/null && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 1 -c 1 -o gpurun_out/prof_lm_final4_r02 -f python profiles/multi_lm.py --jobs 148 --repeat 1 > gpurun_out/ncu_lm_final4.log 2>&1
unset APD_LM_CLUSTER APD_LM_MINB
cut -c1-400 gpurun_out/bench_r02_final4.json
