# round 2, fifth GPU pass: where the loop kernel's time goes UNDER LOAD (phase clock build), in-flight x cluster sweep of
# the fused path, all GPU tests incl. the new large / real-data / boundary ones
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_tests5.txt
cat gpurun_out/r02_tests5.txt
: > gpurun_out/r02_probe5.txt
P="python profiles/pool_probe.py --no-launch-rate"
APD_LIB=$PWD/go-rio_b200/libapdgicp_diag.so timeout 300 $P --streams 128 >> gpurun_out/r02_probe5.txt 2>&1
APD_LIB=$PWD/go-rio_b200/libapdgicp_diag.so timeout 300 $P --streams 8 --pairs 256 >> gpurun_out/r02_probe5.txt 2>&1
for S in 32 48 64 96 128; do
  APD_LM_CLUSTER=8 timeout 300 $P --streams $S >> gpurun_out/r02_probe5.txt 2>&1
done
for S in 48 64 96 128; do
  APD_LM_CLUSTER=4 timeout 300 $P --streams $S >> gpurun_out/r02_probe5.txt 2>&1
done
APD_LM_CLUSTER=4 APD_LM_MINB=1 timeout 300 $P --streams 64 >> gpurun_out/r02_probe5.txt 2>&1
cut -c1-1100 gpurun_out/r02_probe5.txt
