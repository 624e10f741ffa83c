# round 2, eighth GPU pass: search kernel with compacted row loop; ncu --set full of the loop kernel UNDER LOAD (one launch
# over 74 pairs = 296 CTAs, 2 per SM) and alone (4 CTAs), fused path
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_large_parity.py tests/test_real_data.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload c4 --steps 5 --roofline-reps 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['c4']['kernels_rank0_ms']), d['c4']['ms_per_step'], d['c4']['err_equal_across_N'], d['roofline']['frac'])"
timeout 300 python profiles/multi_lm.py --jobs 74 --repeat 3
timeout 300 python profiles/multi_lm.py --jobs 1 --repeat 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 1 -c 1 -o gpurun_out/prof_lm_load_r02 -f python profiles/multi_lm.py --jobs 74 --repeat 1 > gpurun_out/ncu_lm_load.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 1 -c 1 -o gpurun_out/prof_lm_alone_r02 -f python profiles/multi_lm.py --jobs 1 --repeat 1 > gpurun_out/ncu_lm_alone.log 2>&1; echo rc=$?
