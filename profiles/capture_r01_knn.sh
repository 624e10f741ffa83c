# ncu --set full captures of the kNN+covariance kernel (C2 shapes) and of update_correspondences (4 M points)
set -x
CMD1="python profiles/kbench.py --mode c2 --reps 2"
timeout 300 $CMD1 > gpurun_out/knn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_cov_kernel -s 6 -c 2 -o gpurun_out/prof_knn_r01 $CMD1 > gpurun_out/ncu_knn.log 2>&1
echo knn rc=$?
CMD2="python profiles/kbench.py --mode big --n 4000000 --reps 1"
timeout 300 $CMD2 > gpurun_out/corr_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:update_corr_kernel -s 3 -c 1 -o gpurun_out/prof_corr_r01 $CMD2 > gpurun_out/ncu_corr.log 2>&1
echo corr rc=$?
