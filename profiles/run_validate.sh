run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --streams $1 --no-roofline --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; }
echo "16 spin"; run 16
echo "12 spin"; run 12
echo "16 block"; APD_BLOCKING_SYNC=1 run 16
echo "24 block"; APD_BLOCKING_SYNC=1 run 24
echo "32 block"; APD_BLOCKING_SYNC=1 run 32
