"""Does an in-process group of ranks on ONE device depend on which hardware queues its streams land on? Creates n dummy
handles (a CUDA stream each) before a 2-rank / 4-rank group and times one sharded linearize (a rank's reduction kernel
waits inside the kernel for its peers' kernels: two rank streams behind one hardware queue would serialise them)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gorio = importlib.import_module("go-rio_b200")
synth = importlib.import_module("go-rio_b200.synth")
src, tgt, T = synth.scan_pair(1001, 1000)
kw = dict(max_correspondence_distance=2.0, maha_fp64=1, host_loop=1)
dummies = []
for n in [int(a) for a in sys.argv[1:]] or [0, 3, 7, 15, 27, 28, 29, 30, 31, 32, 33, 47, 63]:
    while len(dummies) < n:
        dummies.append(gorio.FastAPDGICP(0))
    for ranks in (2, 4):
        grp = gorio.Group([0] * ranks, **kw)
        grp.set_input_target(tgt); grp.set_input_source(src)
        t0 = time.perf_counter()
        e, H, b = grp.linearize(np.eye(4))
        t1 = time.perf_counter()
        e2, _, _ = grp.linearize(np.eye(4))
        t2 = time.perf_counter()
        print(f"dummies {n:3d} ranks {ranks}: err {e:.6f} / {e2:.6f}  first {1e3 * (t1 - t0):8.1f} ms  second {1e3 * (t2 - t1):8.1f} ms", flush=True)
        grp.close()
