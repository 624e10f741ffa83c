# pass 32 (1 GPU): the whole -m gpu suite, then both bench arms as the driver runs them (the CPU arm is now the reference's own code)
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_tests32.txt; echo tests rc=$? wall=${SECONDS}s; cat gpurun_out/r02_tests32.txt
SECONDS=0
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference32.json 2> gpurun_out/r02_bench_reference32.err; echo ref rc=$? wall=${SECONDS}s
cut -c1-600 gpurun_out/r02_bench_reference32.json
SECONDS=0
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench32.json 2> gpurun_out/r02_bench32.err; echo bench rc=$? wall=${SECONDS}s
cut -c1-300 gpurun_out/r02_bench32.json; tail -3 gpurun_out/r02_bench32.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
