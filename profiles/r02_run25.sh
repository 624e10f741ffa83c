# pass 25 (1 GPU): grid build of the fused kernel with four points per thread in flight (A/B on one box: previous build first)
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or pool or lazy or demand" 2>&1 | tail -3
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
for i in 1 2; do
  APD_LIB=$PWD/go-rio_b200/libapdgicp_prev.so timeout 300 $P 2>&1 | cut -c1-200
  timeout 300 $P 2>&1 | cut -c1-200
done
