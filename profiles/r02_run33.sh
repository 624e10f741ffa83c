# pass 33 (1 GPU): two pooled registrations per launch (a device runs at most 128 grids at a time: 128 of the 148 cluster
# slots were filled) — pool tests, then the pool with one / two per launch and 192 / 256 / 320 in flight
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pool or batch" 2>&1 | tail -4
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe33.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-330 >> gpurun_out/r02_probe33.txt; }
: > gpurun_out/r02_probe33.txt
run APD_PAIR_HOLD_US=0
run APD_PAIR_HOLD_US=200
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3 --streams 256"
run APD_PAIR_HOLD_US=200
run APD_PAIR_HOLD_US=1000
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3 --streams 320"
run APD_PAIR_HOLD_US=200
cat gpurun_out/r02_probe33.txt
