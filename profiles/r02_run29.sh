# pass 29 (8 GPUs, final build): bench.py --gpus 8 as the driver launches it (C3 strong scaling over 4096 distinct scenes + the C4 block)
nproc
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err
echo bench rc=$?; tail -3 gpurun_out/r02_bench_n8_final.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n8_final.json'))
print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e']['value'], d['e2e_packed']['value'], d['eager']['value'] if d.get('eager') else None, d['parity_vs_cpu'], d['host_cpu_ms_per_registration'], d['clocks'])
print(json.dumps(d['c4'])[:2500])
"
