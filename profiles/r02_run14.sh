# round 2, pass 14 (1 GPU): tests in isolation and in full (lazy kernel loading vs in-process ranks), neighbour points staged
# for the covariance passes, pool buffers reserved up front, bench end to end with a short batch
timeout 600 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()"
timeout 300 python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3 2>&1 | cut -c1-330
timeout 900 python bench.py --steps 5 --warmup 3 --pairs 1024 > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err; echo bench rc=$?; tail -2 gpurun_out/r02_bench14.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench14.json'))
print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, 'e2e', d['e2e']['value'], 'packed', d['e2e_packed']['value'], 'eager', d['eager']['value'], d['host_cpu_ms_per_registration'])
print(d['parity_vs_cpu']); print(d['loop_kernel']); print(json.dumps(d['c4']['kernels_rank0_ms']), d['c4']['ms_per_step']); print(d['roofline']['frac'], d['cpu_baseline']); print(json.dumps(d.get('replay'))[:900])
"
