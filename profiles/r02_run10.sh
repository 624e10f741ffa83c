# round 2, tenth GPU pass: the loop kernel's shape against the L2 working set (resident registrations = CTA slots /
# cluster): 512 threads x 2 per SM (pool default) vs 1024 / 768 threads x 1 per SM, clusters of 2 / 4 / 8; grid resolution
P="python profiles/pool_probe.py --no-launch-rate --streams 128"
run() { echo "== $*" >> gpurun_out/r02_probe10.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-330 >> gpurun_out/r02_probe10.txt; }
: > gpurun_out/r02_probe10.txt
run APD_LM_CLUSTER=4
run APD_LM_CLUSTER=4 APD_CELLS_PER_POINT=2
run APD_LM_CLUSTER=4 APD_CELLS_PER_POINT=3
run APD_LM_CLUSTER=4 APD_CELLS_PER_POINT=6
for C in 2 4 8; do
  run APD_LIB=$PWD/go-rio_b200/libapdgicp_t1024.so APD_LM_MINB=1 APD_LM_CLUSTER=$C
done
for C in 2 4; do
  run APD_LIB=$PWD/go-rio_b200/libapdgicp_t768.so APD_LM_MINB=1 APD_LM_CLUSTER=$C
done
run APD_LIB=$PWD/go-rio_b200/libapdgicp_t1024.so APD_LM_MINB=1 APD_LM_CLUSTER=4 APD_CELLS_PER_POINT=2
cat gpurun_out/r02_probe10.txt
