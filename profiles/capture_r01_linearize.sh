# ncu captures of the linearize / compute_error kernels on a 20 M-point cloud (working set > L2):
# launch list first, then --set full of the H+b+err and the err-only kernels.
set -x
CMD="python bench.py --roofline-only --roofline-points 20000000 --roofline-reps 3"
timeout 600 $CMD > gpurun_out/lin_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_roofline.csv $CMD > gpurun_out/ncu_lin_launch.log 2>&1
echo rc=$?
timeout 600 $CMD > gpurun_out/lin_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:linearize_kernel -s 6 -c 4 -o gpurun_out/prof_linearize20m_r01 $CMD > gpurun_out/ncu_lin_full.log 2>&1
echo rc=$?
tail -2 gpurun_out/lin_plain.log
