# round 2, 2-GPU pass: the multi-process sharded path (NCCL all-gather + cudaIpc mailboxes) and bench.py --gpus 2 end to end
timeout 900 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --pairs 1024 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo bench rc=$?; tail -3 gpurun_out/r02_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2.json'))
print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e']['value'], d['e2e_packed']['value'], d['eager']['value'] if d.get('eager') else None, d['parity_vs_cpu'])
print(json.dumps(d['c4'])[:900])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-400
