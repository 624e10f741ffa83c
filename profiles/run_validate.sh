timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_head.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/pytest_head.log
