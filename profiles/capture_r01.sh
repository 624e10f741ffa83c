set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
PROF="python bench.py --steps 1 --warmup 1 --pairs 4 --streams 1 --no-cpu-baseline --roofline-points 20000000 --roofline-reps 2"
timeout 600 $PROF > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01.csv $PROF > gpurun_out/ncu_launch.log 2>&1
echo launch-list rc=$?
timeout 600 $PROF > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:linearize_kernel -c 6 -o gpurun_out/prof_linearize_r01 $PROF > gpurun_out/ncu_full.log 2>&1
echo full rc=$?
ls -la gpurun_out
