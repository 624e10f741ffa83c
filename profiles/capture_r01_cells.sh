for c in 0.5 1 2 4 8; do APD_CELLS_PER_POINT=$c timeout 200 python profiles/batch_trace.py --mode dev --streams 64 --pairs 1024 --steps 4 2>&1 | grep -E "mode=" | sed "s/^/cells_per_point=$c /"; done
for c in 1 2 4; do APD_CELLS_PER_POINT=$c timeout 200 python profiles/kbench.py --mode c2 --reps 30 2>&1 | tail -1 | cut -c1-300 | sed "s/^/cpp=$c /"; done
