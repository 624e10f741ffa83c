timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 300 python profiles/kbench.py --mode c1 --reps 50
timeout 300 python profiles/kbench.py --mode c2 --reps 50
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']}); print(d['roofline']['update_correspondences'], d['roofline']['frac'])"
