# round 2, first GPU pass: the split update_correspondences + one-launch sharded kernels + in-process group
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_tests1.txt
for L in 1 2 4 8; do
  echo "lanes $L" >> gpurun_out/r02_corr_lanes.txt
  APD_CORR_LANES=$L timeout 600 python profiles/kbench.py --mode big --n 20000000 --reps 10 >> gpurun_out/r02_corr_lanes.txt 2>&1
done
cat gpurun_out/r02_tests1.txt gpurun_out/r02_corr_lanes.txt
