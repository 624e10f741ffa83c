# round 2, fourth GPU pass: the fused registration kernel (one launch per registration) against the separate kernels,
# the new bench line end to end, and an ncu --set full of the loop kernel as a pool worker runs it
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_tests4.txt
cat gpurun_out/r02_tests4.txt
: > gpurun_out/r02_probe4.txt
for F in 0 1; do
  APD_FUSED=$F timeout 300 python profiles/pool_probe.py --streams 128 --no-launch-rate >> gpurun_out/r02_probe4.txt 2>&1
  APD_FUSED=$F APD_LM_CLUSTER=2 timeout 300 python profiles/pool_probe.py --streams 128 --no-launch-rate >> gpurun_out/r02_probe4.txt 2>&1
done
APD_FUSED=1 APD_LM_CLUSTER=8 timeout 300 python profiles/pool_probe.py --streams 128 --no-launch-rate >> gpurun_out/r02_probe4.txt 2>&1
cat gpurun_out/r02_probe4.txt
export APD_LAZY_TARGET_COV=1 APD_LM_CLUSTER=4 APD_LM_MINB=2 APD_FUSED=0
K2="python profiles/kbench.py --mode c2 --reps 2"
timeout 200 $K2 > gpurun_out/kbench_c2_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 3 -c 1 -o gpurun_out/prof_lm_r02a -f $K2 > gpurun_out/ncu_lm_r02a.log 2>&1
echo lm rc=$?
unset APD_LAZY_TARGET_COV APD_LM_CLUSTER APD_LM_MINB APD_FUSED
timeout 900 python bench.py --steps 5 --warmup 3 --pairs 1024 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
echo bench rc=$?; tail -3 gpurun_out/r02_bench4.err
cat gpurun_out/r02_bench4.json
