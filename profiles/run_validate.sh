timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_head.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_head.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_d.json 2> gpurun_out/bench_r01_d.err; echo bench rc=$?
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo ref rc=$?
