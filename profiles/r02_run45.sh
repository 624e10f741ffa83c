# pass 45 (1 GPU): build-time knobs of the loop kernel's searches on the C3 probe — kNN flush insert limit 6 / 10 / 16,
# thin shells below radius 2 / 3 / 4, lanes per 1-NN query 4 / 8 / 16 — and the grid resolution of the target (2 / 3 / 4 / 6 cells per point)
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
run() { echo "== $*" >> gpurun_out/r02_probe45.txt; env "$@" timeout 300 $P 2>&1 | cut -c1-120 >> gpurun_out/r02_probe45.txt; }
: > gpurun_out/r02_probe45.txt
run APD_NOP=1
for v in im6 im16 thin2 thin4 g4 g16; do run APD_LIB=$PWD/go-rio_b200/_exp_$v.so; done
run APD_NOP=1
for c in 2 3 6; do run APD_CELLS_PER_POINT=$c; done
cat gpurun_out/r02_probe45.txt
