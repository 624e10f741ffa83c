# pass 42 (1 GPU): lanes per 1-NN query in the loop kernel (build-time APD_LM_G = 2 / 4 / 8 / 16) now that two registrations share a launch
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe42.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-260 >> gpurun_out/r02_probe42.txt; }
: > gpurun_out/r02_probe42.txt
run APD_NOP=1
for G in 2 4 16; do run APD_LIB=$PWD/go-rio_b200/libapdgicp_G$G.so; done
run APD_NOP=1
cat gpurun_out/r02_probe42.txt
