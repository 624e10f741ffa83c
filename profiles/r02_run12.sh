# round 2, twelfth GPU pass: tests; pool with clusters of 2 / 192 in flight / unordered target grid; the scan's grid resolution
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_tests12.txt
cat gpurun_out/r02_tests12.txt
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe12.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-330 >> gpurun_out/r02_probe12.txt; }
: > gpurun_out/r02_probe12.txt
run APD_NOP=1
run APD_CELLS_PER_POINT_SMALL=2
run APD_CELLS_PER_POINT_SMALL=1
run APD_CELLS_PER_POINT_SMALL=0.5
run APD_CELLS_PER_POINT_SMALL=1 APD_CELLS_PER_POINT=3
cat gpurun_out/r02_probe12.txt
