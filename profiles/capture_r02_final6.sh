# End-of-round validation of HEAD (submap grid at 2 cells per point when its covariances are on demand): the whole GPU suite, smoke, and both
# bench arms as the driver runs them. No ncu in this pass.
set -x
SECONDS=0
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6; echo suite wall=${SECONDS}s
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
SECONDS=0
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_reference5.json 2>/dev/null; echo ref rc=$? wall=${SECONDS}s
SECONDS=0
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final6.json 2> gpurun_out/bench_r02_final6.err; echo bench rc=$? wall=${SECONDS}s
cut -c1-600 gpurun_out/bench_r02_final6.json
