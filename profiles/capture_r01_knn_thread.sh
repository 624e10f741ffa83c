# ncu --set full of the thread-per-point kNN kernel: C2 shapes and 4 M points
set -x
CMD1="python profiles/kbench.py --mode c2 --reps 2"
timeout 300 $CMD1 > gpurun_out/knnt_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_cov_thread_kernel -s 6 -c 2 -o gpurun_out/prof_knnt_c2_r01 $CMD1 > gpurun_out/ncu_knnt.log 2>&1
echo rc=$?
CMD2="python profiles/kbench.py --mode big --n 4000000 --reps 1"
timeout 300 $CMD2 > gpurun_out/knnt_big_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_cov_thread_kernel -s 6 -c 1 -o gpurun_out/prof_knnt_big_r01 $CMD2 > gpurun_out/ncu_knnt_big.log 2>&1
echo rc=$?
