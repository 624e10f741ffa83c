"""go-rio_b200 — B200-native FastAPDGICP registration (Go-RIO hot path).

The product is the CUDA library go-rio_b200/csrc -> libapdgicp.so behind the C-ABI
of include/apdgicp.h, plus the C++ shim header that re-creates the reference class
on top of it. This Python package is the harness side only: it loads the library
(ctypes) for tests and bench.py. There is no CPU fallback: ``load()`` raises if the
library is missing, and every compute call fails without an sm_100 device.

(The directory name has a hyphen; import it with
``importlib.import_module("go-rio_b200")``.)
"""
import ctypes
import os

from . import _binding
from ._binding import (  # noqa: F401
    ApdError, ApdPair, ApdParams, ApdResult, Registration,
    OPT_GN, OPT_LM, REG_FROBENIUS, REG_MIN_EIG, REG_NONE, REG_NORMALIZED_MIN_EIG, REG_PLANE,
)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libapdgicp.so")

_lib = None


def load():
    """Loads libapdgicp.so (built in-tree by __graft_entry__.build() / csrc/build.sh)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the registration path)"
            )
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    return _lib


def FastAPDGICP(device=0):
    """A registration object on `device` (mirror of fast_gicp::FastAPDGICP)."""
    return Registration(load(), "apd_", device)


def align_batch(pairs, params=None, device=0, n_streams=4, with_fitness=True):
    """apd_align_batch over a list of (source[n,4] f32, target[m,4] f32, guess 4x4 or None)."""
    import numpy as np

    lib = load()
    if params is None:
        params = ApdParams()
        lib.apd_default_params(ctypes.byref(params))
    n = len(pairs)
    arr = (ApdPair * n)()
    keep = []
    for i, (s, t, g) in enumerate(pairs):
        s = np.ascontiguousarray(s, dtype=np.float32)
        t = np.ascontiguousarray(t, dtype=np.float32)
        keep += [s, t]
        arr[i].source = s.ctypes.data
        arr[i].n_source = s.shape[0]
        arr[i].target = t.ctypes.data
        arr[i].n_target = t.shape[0]
        if g is not None:
            gg = np.ascontiguousarray(np.asarray(g, dtype=np.float32).T).reshape(-1)
            keep.append(gg)
            arr[i].guess = gg.ctypes.data
        else:
            arr[i].guess = None
    res = (ApdResult * n)()
    lib.apd_align_batch.restype = ctypes.c_int
    rc = lib.apd_align_batch(ctypes.c_int(device), ctypes.byref(params), arr, ctypes.c_int32(n), ctypes.c_int32(16),
                             ctypes.c_int32(0), ctypes.c_int32(12), ctypes.c_int32(n_streams),
                             ctypes.c_int32(1 if with_fitness else 0), res)
    if rc != 0:
        raise ApdError(rc, "apd_align_batch")
    out = []
    for r in res:
        out.append(dict(T=np.array(r.T, dtype=np.float32).reshape(4, 4).T.copy(), fitness=r.fitness,
                        converged=bool(r.converged), iterations=r.iterations, status=r.status, n_inliers=r.n_inliers))
    return out
