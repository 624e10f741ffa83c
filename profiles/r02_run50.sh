# pass 50 (1 GPU): the loop kernel with 448 / 384 threads per CTA x 2 CTAs per SM (72 / 80 registers instead of 64: fewer
# spills, 28 / 24 warps per SM instead of 32) on the C3 probe
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe50.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-120 >> gpurun_out/r02_probe50.txt; }
: > gpurun_out/r02_probe50.txt
run APD_NOP=1
run APD_LIB=$PWD/go-rio_b200/_exp_t448.so
run APD_LIB=$PWD/go-rio_b200/_exp_t384.so
run APD_NOP=1
cat gpurun_out/r02_probe50.txt
