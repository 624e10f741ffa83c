# round 2, second GPU pass: lane-parallel 1-NN candidate scan, insertion flush in the kNN, dynamic work distribution in
# the loop kernel; in-flight / cluster sweep of the pool
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_tests2.txt
cat gpurun_out/r02_tests2.txt
run() { echo "== streams=$1 cluster=$2 pairs=$3" >> gpurun_out/r02_sweep2.txt; APD_LM_CLUSTER=$2 timeout 600 python bench.py --steps 10 --warmup 3 --pairs $3 --streams $1 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['host_cpu_ms_per_registration'], {k:(v['ms_per_step'],v['launches_per_step']) for k,v in d['kernels'].items() if v['ms_per_step']})" >> gpurun_out/r02_sweep2.txt 2>&1; }
run 64 4 512
run 64 4 2048
run 128 4 2048
run 128 2 2048
run 96 4 2048
run 128 8 2048
cat gpurun_out/r02_sweep2.txt
timeout 600 python bench.py --steps 5 --warmup 3 --roofline-only 2>/dev/null | tail -1 > gpurun_out/r02_roofline2.json
cat gpurun_out/r02_roofline2.json
