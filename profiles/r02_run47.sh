# pass 47 (1 GPU): the prologue's cluster-wide scan of the cell counters four counters per lane and turn (-DAPD_SCAN_VEC=1),
# with eight points per thread in flight in the rank / scatter passes, and with the evict-first hints; fused / pool tests
# on the scan build first
APD_LIB=$PWD/go-rio_b200/_exp_scanvec.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_code.py -m gpu -x -q 2>&1 | tail -3
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe47.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-120 >> gpurun_out/r02_probe47.txt; }
: > gpurun_out/r02_probe47.txt
run APD_NOP=1
for v in scanvec scanvec_u8 scanvec_hints; do run APD_LIB=$PWD/go-rio_b200/_exp_$v.so; done
run APD_NOP=1
run APD_LIB=$PWD/go-rio_b200/_exp_scanvec.so
cat gpurun_out/r02_probe47.txt
