# pass 19: the full GPU suite twice (group ranks drain + meet after their collectives), then the C4 block by motion
for i in 1 2; do timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; done
timeout 600 python bench.py --workload c4 --steps 8 --roofline-reps 8 --no-cpu-baseline > gpurun_out/r02_c4_19.json 2> gpurun_out/r02_c4_19.err; echo c4 rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_c4_19.json'))
print(json.dumps(d['c4']['kernels_rank0_ms']), d['c4']['ms_per_step'], d['c4']['update_correspondences_ms_by_motion'], d['c4']['err_equal_across_N'])
"
