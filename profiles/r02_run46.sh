# pass 46 (1 GPU): grid resolution of the submap alone (APD_CELLS_PER_POINT_MID; the scan stays at 1 cell per point) on the
# C3 probe, and evict-first hints on the prologue's read-once / write-once traffic (-DAPD_STREAM_HINTS=1)
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe46.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-120 >> gpurun_out/r02_probe46.txt; }
: > gpurun_out/r02_probe46.txt
run APD_NOP=1
run APD_LIB=$PWD/go-rio_b200/_exp_hints.so
for c in 1 1.5 2 2.5 3; do run APD_CELLS_PER_POINT_MID=$c; done
run APD_CELLS_PER_POINT_MID=2 APD_LIB=$PWD/go-rio_b200/_exp_hints.so
run APD_CELLS_PER_POINT_MID=2 APD_CELLS_PER_POINT_SMALL=0.5
run APD_NOP=1
cat gpurun_out/r02_probe46.txt
