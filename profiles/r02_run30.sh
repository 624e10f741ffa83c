# pass 30 (1 GPU): branch-free tracking of the second-nearest bound (fmax / fmin / select per candidate) — measured 3 % SLOWER than the branchy form (cold 2.45 vs 2.37 ms, 12 cm 1.93 vs 1.85): reverted
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_large_parity.py tests/test_sharding.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload c4 --steps 8 --roofline-reps 8 --no-cpu-baseline > gpurun_out/r02_c4_30.json 2> gpurun_out/r02_c4_30.err; echo c4 rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_c4_30.json'))
print(json.dumps(d['c4']['kernels_rank0_ms']), d['c4']['ms_per_step'], d['c4']['update_correspondences_ms_by_motion'], d['c4']['err_equal_across_N'], d['c4']['align']['ms'])
"
