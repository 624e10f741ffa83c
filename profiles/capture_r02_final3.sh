# End-of-round capture, third part: the bench line as the driver runs it (wall clock of the whole command alongside)
SECONDS=0
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo bench rc=$? wall=${SECONDS}s
cut -c1-300 gpurun_out/bench_r02_final.json
