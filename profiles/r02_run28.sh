# pass 28 (2 GPUs): the multi-process sharded tests; bench.py --gpus 2 end to end
timeout 900 python -m pytest tests/test_sharding.py tests/test_large_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2c.json 2> gpurun_out/r02_bench_n2c.err
echo bench rc=$?; tail -3 gpurun_out/r02_bench_n2c.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2c.json'))
print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e']['value'], d['e2e_packed']['value'], d['eager']['value'] if d.get('eager') else None, d['parity_vs_cpu'], d['host_cpu_ms_per_registration'])
print(json.dumps(d['c4'])[:1500])
"
