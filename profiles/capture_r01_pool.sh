timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for t in 4 8 16; do for s in 32 64; do APD_BATCH_THREADS=$t timeout 200 python profiles/batch_trace.py --mode dev --streams $s --pairs 1024 --steps 4 2>&1 | grep -E "mode=|trace" | sed "s/^/threads=$t /"; done; done
for t in 8 16; do APD_BATCH_THREADS=$t timeout 200 python profiles/batch_trace.py --mode host --streams 64 --pairs 1024 --steps 4 2>&1 | grep -E "mode=|trace" | sed "s/^/threads=$t /"; done
APD_LAZY_SEED=0 timeout 200 python profiles/batch_trace.py --mode dev --streams 64 --pairs 1024 --steps 4 2>&1 | grep -E "mode=|trace" | sed "s/^/noseed /"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --streams 64 2>/dev/null | cut -c1-2000
