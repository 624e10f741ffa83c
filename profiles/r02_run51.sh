# pass 51 (1 GPU): the loop kernel with 576 / 640 threads per CTA x 2 CTAs per SM (56 / 48 registers instead of 64: more
# spills, 36 / 40 warps per SM instead of 32) on the C3 probe
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe51.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-120 >> gpurun_out/r02_probe51.txt; }
: > gpurun_out/r02_probe51.txt
run APD_NOP=1
run APD_LIB=$PWD/go-rio_b200/_exp_t576.so
run APD_LIB=$PWD/go-rio_b200/_exp_t640.so
run APD_NOP=1
cat gpurun_out/r02_probe51.txt
