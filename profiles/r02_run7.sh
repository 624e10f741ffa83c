# round 2, seventh GPU pass: tests (incl. radius / voxel / submap), the single-lane search at 20 M points, the pool
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_tests7.txt
cat gpurun_out/r02_tests7.txt
timeout 600 python bench.py --workload c4 --steps 5 --roofline-reps 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['c4']['kernels_rank0_ms']), d['c4']['ms_per_step'], d['c4']['err_equal_across_N'], d['roofline']['frac'], json.dumps(d['roofline']['update_correspondences']))" > gpurun_out/r02_c4_7.txt 2>&1
cat gpurun_out/r02_c4_7.txt
timeout 300 python profiles/kbench.py --mode big --n 20000000 --reps 5 > gpurun_out/r02_kbench_big7.txt 2>&1; cat gpurun_out/r02_kbench_big7.txt
timeout 300 python profiles/pool_probe.py --no-launch-rate --streams 128 2>&1 | cut -c1-330
