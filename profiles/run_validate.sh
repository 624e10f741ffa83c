set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_head.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/pytest_head.log
for s in 8 16 32; do
timeout 600 python bench.py --steps 20 --warmup 3 --streams $s --no-roofline --no-cpu-baseline > gpurun_out/bench_s$s.json 2> gpurun_out/bench_s$s.err; echo bench rc=$?
tail -3 gpurun_out/bench_s$s.err
done
timeout 600 python bench.py --steps 20 --warmup 3 --streams 16 --pairs 128 --no-roofline --no-cpu-baseline > gpurun_out/bench_p128.json 2> gpurun_out/bench_p128.err
