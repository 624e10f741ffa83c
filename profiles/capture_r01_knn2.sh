# ncu --set full of the (buffered-merge) warp-per-point kNN kernel on the C2 shapes
set -x
CMD1="python profiles/kbench.py --mode c2 --reps 2"
timeout 300 $CMD1 > gpurun_out/knn2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_cov_kernel -s 7 -c 1 -o gpurun_out/prof_knn2_r01 $CMD1 > gpurun_out/ncu_knn2.log 2>&1
echo rc=$?
