# End-of-round capture, second part (final build): the whole suite, smoke, the bench line and the reference arm as the driver
# runs them, and ncu of the COLD search kernel (the first update_correspondences pass of the C4 block: no previous pass to
# start from). Each ncu pass runs only after the same command has exited 0 without ncu.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
/usr/bin/time -v timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo bench rc=$?
grep -E "Elapsed|Maximum resident" gpurun_out/bench_r02_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2>/dev/null; echo ref rc=$?
C4="python bench.py --workload c4 --steps 3 --roofline-reps 3 --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"corr_search" -c 1 -o gpurun_out/prof_c4_cold_r02 -f $C4 > gpurun_out/ncu_c4_cold.log 2>&1
cut -c1-400 gpurun_out/bench_r02_final.json
