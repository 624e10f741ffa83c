# pass 39 (1 GPU): the searches ask for the cache lines of their segments / row bounds ahead of the loads (prefetch.global.L2 / L1) — A/B
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe39.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-260 >> gpurun_out/r02_probe39.txt; }
: > gpurun_out/r02_probe39.txt
run APD_LIB=$PWD/go-rio_b200/libapdgicp_NOPF.so
run APD_NOP=1
run APD_LIB=$PWD/go-rio_b200/libapdgicp_L1.so
run APD_LIB=$PWD/go-rio_b200/libapdgicp_NOPF.so
run APD_NOP=1
cat gpurun_out/r02_probe39.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
