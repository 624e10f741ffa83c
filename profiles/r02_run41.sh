# pass 41 (1 GPU): kNN — where the own row holds k points it is scanned first and its k-th distance prunes the other eight rows of the cube
# 64-bit sort when two candidates agree in those bits — parity tests, then the pool
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_code.py tests/test_large_parity.py -m gpu -x -q 2>&1 | tail -3
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe41.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-260 >> gpurun_out/r02_probe41.txt; }
: > gpurun_out/r02_probe41.txt
run APD_NOP=1
run APD_NOP=1
cat gpurun_out/r02_probe41.txt
