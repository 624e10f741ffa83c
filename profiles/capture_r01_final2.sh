# End-of-round capture after the 64-register pool build of the loop kernel: GPU tests, the bench line, the reference arm,
# ncu --set full of lm_kernel as a pool worker runs it (APD_LAZY_TARGET_COV=1 APD_LM_CLUSTER=4 APD_LM_MINB=2).
set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2>/dev/null; echo ref rc=$?
export APD_LAZY_TARGET_COV=1 APD_LM_CLUSTER=4 APD_LM_MINB=2
K2="python profiles/kbench.py --mode c2 --reps 2"
timeout 200 $K2 > gpurun_out/kbench_c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 3 -c 1 -o gpurun_out/prof_lm64_r01 -f $K2 > gpurun_out/ncu_lm64.log 2>&1
echo lm rc=$?
cut -c1-1200 gpurun_out/bench_r01_final.json
