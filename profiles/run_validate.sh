timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_head.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_head.log
run() { timeout 600 python bench.py --steps 10 --warmup 3 --streams $1 --pairs $2 --no-roofline --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']})"; }
echo "spin s16 p128"; run 16 128
echo "block s16 p128"; APD_BLOCKING_SYNC=1 run 16 128
echo "block s32 p256"; APD_BLOCKING_SYNC=1 run 32 256
echo "block s48 p384"; APD_BLOCKING_SYNC=1 run 48 384
