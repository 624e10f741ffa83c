#!/usr/bin/env bash
# compute-sanitizer evidence: memcheck + racecheck + synccheck of (a) the smoke registration, (b) one
# pooled batch through the fused registration kernel with target covariances on demand (the path the bench runs), (c) an
# eager lone registration through the separate kernels and the host-driven loop. Run on a B200 (gpurun). racecheck sees
# shared memory only; the on-demand covariance publication is global memory (release / acquire flag, see lm.cu).
set -u
OUT=gpurun_out/sanitizer_r02
mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
gorio = importlib.import_module("go-rio_b200"); synth = importlib.import_module("go-rio_b200.synth")
case = sys.argv[1]
kw = dict(max_correspondence_distance=2.0, transformation_epsilon=0.1)
if case == "smoke":
    import __graft_entry__ as ge
    ge.smoke()
elif case == "pool":
    pairs = [synth.submap_pair(2300 + i, n_source=500, n_frames=4, n_per_frame=800)[:2] for i in range(6)]
    b = gorio.Batch(0, n_workers=3, **kw)
    r = b.align([(s, t, None) for s, t in pairs], with_fitness=True)
    assert all(x["status"] == 0 for x in r), r
    print("pool ok", [x["iterations"] for x in r]); b.close()
elif case == "eager":
    s, t, _ = synth.submap_pair(2310, n_source=500, n_frames=4, n_per_frame=800)
    for host_loop in (0, 1):
        g = gorio.FastAPDGICP(0); g.set_params(**kw, host_loop=host_loop)
        g.set_input_target(t); g.set_input_source(s); r = g.align(); g.fitness()
        print("eager ok", host_loop, r["iterations"]); g.close()
elif case == "group":
    s, t, T = synth.tiled_cloud_pair(4001, 40000)
    grp = gorio.Group([0, 0], max_correspondence_distance=2.0)
    grp.set_input_target(t); grp.set_input_source(s); print("group ok", grp.linearize(T)[0]); grp.close()
PY
run() {
  timeout 900 compute-sanitizer --tool $1 --print-limit 20 python /tmp/san_case.py $2 > $OUT/$1_$2.log 2>&1
  echo "$1 $2 rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$1_$2.log | tail -1) | $(grep -E ' ok' $OUT/$1_$2.log | tail -1)"
}
for case in smoke pool eager group; do run memcheck $case; done
for case in pool eager; do run racecheck $case; done
for case in pool group; do run synccheck $case; done
