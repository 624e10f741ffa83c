# round 2, third GPU pass: what bounds the pool (launch rate vs GPU), the new bench line end to end
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_tests3.txt
cat gpurun_out/r02_tests3.txt
: > gpurun_out/r02_probe3.txt
timeout 300 python profiles/pool_probe.py --streams 128 >> gpurun_out/r02_probe3.txt 2>&1
APD_LM_CLUSTER=2 timeout 300 python profiles/pool_probe.py --streams 128 --no-launch-rate >> gpurun_out/r02_probe3.txt 2>&1
APD_BATCH_THREADS=2 timeout 300 python profiles/pool_probe.py --streams 128 --no-launch-rate >> gpurun_out/r02_probe3.txt 2>&1
APD_BATCH_THREADS=4 timeout 300 python profiles/pool_probe.py --streams 128 --no-launch-rate >> gpurun_out/r02_probe3.txt 2>&1
APD_BATCH_THREADS=16 timeout 300 python profiles/pool_probe.py --streams 128 --no-launch-rate >> gpurun_out/r02_probe3.txt 2>&1
APD_POLL_WAIT_US=10 timeout 300 python profiles/pool_probe.py --streams 128 --no-launch-rate >> gpurun_out/r02_probe3.txt 2>&1
cat gpurun_out/r02_probe3.txt
/usr/bin/time -v timeout 900 python bench.py --steps 5 --warmup 3 --pairs 1024 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
tail -3 gpurun_out/r02_bench3.err | head -2; grep -E "Elapsed|Maximum resident" gpurun_out/r02_bench3.err
cat gpurun_out/r02_bench3.json
