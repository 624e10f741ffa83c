# pass 18: is the 1-GPU multi-rank group sensitive to the streams' hardware queues? + the sharding tests alone, 3 times
timeout 600 python profiles/group_alias_probe.py 2>&1 | tail -30
for i in 1 2 3; do timeout 300 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -2; done
