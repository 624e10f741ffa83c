# pass 27 (1 GPU): reduction kernels publish their totals into pinned host memory (no copy, no stream wait); whole suite, C4 block
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload c4 --steps 8 --roofline-reps 8 --no-cpu-baseline > gpurun_out/r02_c4_27.json 2> gpurun_out/r02_c4_27.err; echo c4 rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_c4_27.json'))
print(json.dumps(d['c4']['kernels_rank0_ms']), d['c4']['ms_per_step'], d['c4']['err_equal_across_N'], d['c4']['align']['ms'], d['c4']['align']['outer_iterations'])
print(d['roofline']['frac'], d['roofline']['compute_error']['frac'])
"
