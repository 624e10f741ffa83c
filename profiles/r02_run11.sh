# round 2, eleventh GPU pass: fewer CTAs per registration, more registrations resident
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe11.txt; env "${@:2}" timeout 300 $P --streams $1 2>&1 | cut -c1-330 >> gpurun_out/r02_probe11.txt; }
: > gpurun_out/r02_probe11.txt
T=$PWD/go-rio_b200/libapdgicp_t1024.so
run 128 APD_LIB=$T APD_LM_MINB=1 APD_LM_CLUSTER=1
run 192 APD_LIB=$T APD_LM_MINB=1 APD_LM_CLUSTER=1
run 256 APD_LIB=$T APD_LM_MINB=1 APD_LM_CLUSTER=1
run 128 APD_LIB=$T APD_LM_MINB=1 APD_LM_CLUSTER=2
run 192 APD_LIB=$T APD_LM_MINB=1 APD_LM_CLUSTER=2
run 192 APD_LIB=$T APD_LM_MINB=1 APD_LM_CLUSTER=2 APD_CELLS_PER_POINT=3
run 320 APD_LM_CLUSTER=1
run 384 APD_LM_CLUSTER=1
run 192 APD_LM_CLUSTER=2
run 256 APD_LM_CLUSTER=2
cat gpurun_out/r02_probe11.txt
