timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python profiles/replay_bench.py --frames 10000 --cpu-frames 500 > gpurun_out/replay_c5.json 2> gpurun_out/replay_c5.err; echo rc=$?
cat gpurun_out/replay_c5.json; tail -3 gpurun_out/replay_c5.err
