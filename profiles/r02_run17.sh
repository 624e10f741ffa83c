# round 2, pass 17 (1 GPU): matches kept by their stored bounds in the single-lane search (tests + C4 block by motion),
# bench steps back to back through the pool's queue
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_tests17.txt
cat gpurun_out/r02_tests17.txt
timeout 900 python bench.py --steps 5 --warmup 3 --pairs 2048 > gpurun_out/r02_bench17.json 2> gpurun_out/r02_bench17.err; echo bench rc=$?; tail -3 gpurun_out/r02_bench17.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench17.json'))
print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, 'e2e', d['e2e']['value'], 'packed', d['e2e_packed']['value'], 'eager', d['eager']['value'], d['host_cpu_ms_per_registration'])
print(d['e2e']['same_result_as_device_resident'], d['parity_vs_cpu']); print(d['loop_kernel']); print(json.dumps(d['c4']['kernels_rank0_ms']), d['c4']['ms_per_step'], d['c4']['update_correspondences_ms_by_motion'], d['c4']['err_equal_across_N']); print(d['roofline']['frac'], d['roofline']['update_correspondences']['frac'], d['cpu_baseline'])
"
