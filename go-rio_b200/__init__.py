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

# see csrc/apdgicp.cu (apd_default_connections): one hardware work queue per worker stream; must precede the CUDA context
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.environ.get("APD_LIB") or os.path.join(PKG_DIR, "libapdgicp.so")  # APD_LIB: experiment builds

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


def _pair_array(pairs, keep, raw=False):
    import numpy as np

    n = len(pairs)
    arr = (ApdPair * n)()
    for i, (s, t, g) in enumerate(pairs):
        if isinstance(s, tuple):  # (device pointer, count): cloud already on the GPU as float4 {x,y,z,label}
            arr[i].source, arr[i].n_source = s
            arr[i].target, arr[i].n_target = t
        elif raw:  # structured AoS (e.g. the 48-byte pcl::PointXYZINormal): passed as it lies
            keep += [s, t]
            arr[i].source, arr[i].n_source = s.ctypes.data, s.shape[0]
            arr[i].target, arr[i].n_target = t.ctypes.data, t.shape[0]
        else:
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
    return arr


def _results(res):
    import numpy as np

    return [dict(T=np.array(r.T, dtype=np.float32).reshape(4, 4).T.copy(), fitness=r.fitness, converged=bool(r.converged),
                 iterations=r.iterations, status=r.status, n_inliers=r.n_inliers) for r in res]


class Batch:
    """apd_batch_*: a persistent pool of `n_workers` handles / streams / host threads for independent pairs."""

    def __init__(self, device=0, n_workers=8, params=None, **kw):
        """device: one device index, or a list of them (apd_batch_create_multi: n_workers registrations in flight per
        device, one shared queue of pairs)"""
        self._lib = load()
        self._b = ctypes.c_void_p()
        self.devices = list(device) if isinstance(device, (list, tuple)) else [device]
        arr = (ctypes.c_int32 * len(self.devices))(*self.devices)
        self._lib.apd_batch_create_multi.restype = ctypes.c_int
        rc = self._lib.apd_batch_create_multi(arr, ctypes.c_int32(len(self.devices)), ctypes.c_int32(n_workers), ctypes.byref(self._b))
        if rc != 0:
            raise ApdError(rc, "apd_batch_create")
        if params is None:
            params = ApdParams()
            self._lib.apd_default_params(ctypes.byref(params))
        for k, v in kw.items():
            setattr(params, k, v)
        rc = self._lib.apd_batch_set_params(self._b, ctypes.byref(params))
        if rc != 0:
            raise ApdError(rc, "apd_batch_set_params")

    def prepare(self, pairs, layout=(16, 0, 12)):
        """list of (source, target, guess) -> a reusable C array; source/target are [n,4] f32 arrays
        (host clouds), structured PCL arrays with layout=(stride, xyz_off, label_off), or (device_ptr, n) tuples
        (clouds resident in HBM)."""
        keep = []
        structured = len(pairs) > 0 and not isinstance(pairs[0][0], tuple) and pairs[0][0].dtype.names is not None
        arr = _pair_array(pairs, keep, raw=structured)
        device = len(pairs) > 0 and isinstance(pairs[0][0], tuple)
        return dict(arr=arr, keep=keep, n=len(pairs), device=device, res=(ApdResult * len(pairs))(), layout=layout)

    def repeat(self, prepared, k):
        """the prepared pair list k times over, as ONE list (k passes over a batch without draining the pool in between)"""
        n = prepared["n"]
        arr = (ApdPair * (n * k))()
        size = ctypes.sizeof(ApdPair) * n
        for j in range(k):
            ctypes.memmove(ctypes.addressof(arr) + j * size, prepared["arr"], size)
        return dict(arr=arr, keep=[prepared], n=n * k, device=prepared["device"], res=(ApdResult * (n * k))(), layout=prepared.get("layout", (16, 0, 12)))

    @staticmethod
    def results_view(res):
        """the C result array as a NumPy structured array (no copy)"""
        import numpy as np

        dt = np.dtype([("T", np.float32, (16,)), ("fitness", np.float64), ("converged", np.int32), ("iterations", np.int32),
                       ("status", np.int32), ("n_inliers", np.int32)])
        assert dt.itemsize == ctypes.sizeof(ApdResult)
        return np.frombuffer(res, dtype=dt)

    def device_pairs(self):
        out = (ctypes.c_int64 * len(self.devices))()
        self._lib.apd_batch_device_pairs(self._b, out, ctypes.c_int32(len(self.devices)))
        return list(out)

    def align(self, prepared, with_fitness=True, parse=True):
        if not isinstance(prepared, dict):
            prepared = self.prepare(prepared)
        n, arr, res = prepared["n"], prepared["arr"], prepared["res"]
        wf = ctypes.c_int32(1 if with_fitness else 0)
        if prepared["device"]:
            rc = self._lib.apd_batch_align_device(self._b, arr, ctypes.c_int32(n), wf, res)
        else:
            stride, xo, lo = prepared.get("layout", (16, 0, 12))
            rc = self._lib.apd_batch_align(self._b, arr, ctypes.c_int32(n), ctypes.c_int32(stride), ctypes.c_int32(xo), ctypes.c_int32(lo), wf, res)
        if rc != 0:
            raise ApdError(rc, "apd_batch_align")
        return _results(res) if parse else res

    def launch_count(self):
        self._lib.apd_batch_launch_count.restype = ctypes.c_int64
        return self._lib.apd_batch_launch_count(self._b)

    def load_stats(self, reset=True):
        """see apd_batch_get_load_stats"""
        st = (ctypes.c_double * 17)()
        self._lib.apd_batch_get_load_stats(self._b, st, ctypes.c_int32(17), ctypes.c_int32(1 if reset else 0))
        names = ["registrations", "lm_kernel_ms", "host_set_ms", "host_bbox_wait_ms", "host_enqueue_prep_ms", "host_enqueue_loop_ms", "host_result_wait_ms"]
        out = dict(zip(names, st))
        phases = ["boxes_grids", "source_cov", "nn_search", "lazy_knn", "lazy_cov", "mahalanobis", "sums_solve", "lm_trials", "fitness", "spare"]
        out["lm_phase_ms"] = {k: st[7 + i] for i, k in enumerate(phases)}  # (zeros unless the library was built with -DAPD_LM_PHASE_TIMING)
        return out

    def set_profiling(self, on):
        self._lib.apd_batch_set_profiling(self._b, ctypes.c_int32(1 if on else 0))

    def kernel_ms(self):
        n = len(_binding.KERNEL_CLASSES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        self._lib.apd_batch_get_kernel_ms(self._b, ms, cnt)
        return {k: (ms[i], cnt[i]) for i, k in enumerate(_binding.KERNEL_CLASSES)}

    def close(self):
        if self._b:
            self._lib.apd_batch_destroy(self._b)
            self._b = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """apd_group_*: the handles of ONE process as the ranks of one sharded registration (devices[i] = the device of rank i;
    several ranks may share a device). The calls fan out to all ranks from the library's own threads."""

    def __init__(self, devices, **params):
        import numpy as np

        self._np = np
        self._lib = load()
        self.ranks = [Registration(self._lib, "apd_", d) for d in devices]
        if params:
            for r in self.ranks:
                r.set_params(**params)
        arr = (ctypes.c_void_p * len(self.ranks))(*[r._h for r in self.ranks])
        self._g = ctypes.c_void_p()
        self._lib.apd_group_create.restype = ctypes.c_int
        rc = self._lib.apd_group_create(arr, ctypes.c_int32(len(self.ranks)), ctypes.byref(self._g))
        if rc != 0:
            raise ApdError(rc, "apd_group_create")
        self._keep = []

    def _check(self, rc, what):
        if rc != 0:
            raise ApdError(rc, what + ": " + "; ".join(r.last_error() for r in self.ranks))

    def _set(self, name, cloud):
        a, n, stride, xo, lo = Registration._layout(cloud)
        self._keep = self._keep[-3:] + [a]
        fn = getattr(self._lib, name)
        fn.restype = ctypes.c_int
        self._check(fn(self._g, a.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(n), ctypes.c_int32(stride), ctypes.c_int32(xo),
                       ctypes.c_int32(lo)), name)
        return n

    def set_input_source(self, cloud):
        n = self._set("apd_group_set_source", cloud)
        for r in self.ranks:
            r.n_source = n

    def set_input_target(self, cloud):
        n = self._set("apd_group_set_target", cloud)
        for r in self.ranks:
            r.n_target = n

    def _set_device(self, name, dptrs, n):
        arr = (ctypes.c_void_p * len(self.ranks))(*dptrs)
        fn = getattr(self._lib, name)
        fn.restype = ctypes.c_int
        self._check(fn(self._g, arr, ctypes.c_int32(n)), name)

    def set_input_source_device(self, dptrs, n):
        self._set_device("apd_group_set_source_device", dptrs, n)
        for r in self.ranks:
            r.n_source = n

    def set_input_target_device(self, dptrs, n):
        self._set_device("apd_group_set_target_device", dptrs, n)
        for r in self.ranks:
            r.n_target = n

    def linearize(self, T, want_hb=True):
        np = self._np
        t = _binding._colmajor(T, np.float64)
        H, b, err = np.empty(36, np.float64), np.empty(6, np.float64), ctypes.c_double()
        self._lib.apd_group_linearize.restype = ctypes.c_int
        self._check(self._lib.apd_group_linearize(self._g, t.ctypes.data_as(ctypes.c_void_p), H.ctypes.data_as(ctypes.c_void_p) if want_hb else None,
                                                  b.ctypes.data_as(ctypes.c_void_p) if want_hb else None, ctypes.byref(err)), "apd_group_linearize")
        return (err.value, H.reshape(6, 6).T.copy(), b) if want_hb else err.value

    def compute_error(self, T):
        t = _binding._colmajor(T, self._np.float64)
        err = ctypes.c_double()
        self._lib.apd_group_compute_error.restype = ctypes.c_int
        self._check(self._lib.apd_group_compute_error(self._g, t.ctypes.data_as(ctypes.c_void_p), ctypes.byref(err)), "apd_group_compute_error")
        return err.value

    def align(self, guess=None):
        np = self._np
        g = None if guess is None else _binding._colmajor(guess, np.float32)
        T, T64, H = np.empty(16, np.float32), np.empty(16, np.float64), np.empty(36, np.float64)
        conv, it = ctypes.c_int32(), ctypes.c_int32()
        self._lib.apd_group_align.restype = ctypes.c_int
        self._check(self._lib.apd_group_align(self._g, g.ctypes.data_as(ctypes.c_void_p) if g is not None else None, T.ctypes.data_as(ctypes.c_void_p),
                                              T64.ctypes.data_as(ctypes.c_void_p), H.ctypes.data_as(ctypes.c_void_p), ctypes.byref(conv),
                                              ctypes.byref(it)), "apd_group_align")
        return dict(T=T.reshape(4, 4).T.copy(), T64=T64.reshape(4, 4).T.copy(), H=H.reshape(6, 6).T.copy(), converged=bool(conv.value),
                    iterations=it.value)

    def launch_count(self):
        return sum(r.launch_count() for r in self.ranks)

    def close(self):
        if self._g:
            self._lib.apd_group_destroy(self._g)
            self._g = ctypes.c_void_p()
        for r in self.ranks:
            r.close()
        self.ranks = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def align_batch(pairs, params=None, device=0, n_streams=4, with_fitness=True):
    """apd_align_batch over a list of (source[n,4] f32, target[m,4] f32, guess 4x4 or None)."""
    lib = load()
    if params is None:
        params = ApdParams()
        lib.apd_default_params(ctypes.byref(params))
    n = len(pairs)
    keep = []
    arr = _pair_array(pairs, keep)
    res = (ApdResult * n)()
    lib.apd_align_batch.restype = ctypes.c_int
    rc = lib.apd_align_batch(ctypes.c_int(device), ctypes.byref(params), arr, ctypes.c_int32(n), ctypes.c_int32(16),
                             ctypes.c_int32(0), ctypes.c_int32(12), ctypes.c_int32(n_streams),
                             ctypes.c_int32(1 if with_fitness else 0), res)
    if rc != 0:
        raise ApdError(rc, "apd_align_batch")
    return _results(res)
