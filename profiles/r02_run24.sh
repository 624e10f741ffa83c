# pass 24 (1 GPU): searching queries packed into the block's first warps after the kept-match test; two-level final sum of
# the reduction kernels; FastVGICP in a pool. Whole suite, then the C4 block.
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --workload c4 --steps 8 --roofline-reps 8 --no-cpu-baseline > gpurun_out/r02_c4_24.json 2> gpurun_out/r02_c4_24.err; echo c4 rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_c4_24.json'))
print(json.dumps(d['c4']['kernels_rank0_ms']), d['c4']['ms_per_step'], d['c4']['update_correspondences_ms_by_motion'], d['c4']['err_equal_across_N'], d['c4']['align'])
print(d['roofline']['frac'], d['roofline']['compute_error']['frac'], d['roofline']['traffic'])
"
