"""ctypes binding of the C-ABI declared in include/apdgicp.h.

The wrapper class is generic over (library, symbol prefix) because the CPU oracle
exports the same ABI under the prefix ``apdo_`` (tests/oracle_binding.py); this
module itself never touches the oracle.

Python is only the test/bench harness here — the host side of the product is the
C++ shim in go-rio_b200/include/fast_gicp/gicp/fast_apdgicp.hpp.
"""
import ctypes as C

import numpy as np

APD_OK, APD_ERR_INVALID, APD_ERR_CUDA, APD_ERR_TOO_FEW, APD_ERR_UNSUPPORTED, APD_ERR_COMM = range(6)
REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
OPT_GN, OPT_LM = 0, 1
KERNEL_CLASSES = ["grid", "knn_cov", "corr", "linearize", "error", "fitness", "lm"]


class ApdParams(C.Structure):
    _fields_ = [
        ("k_correspondences", C.c_int32),
        ("regularization", C.c_int32),
        ("max_correspondence_distance", C.c_double),
        ("dist_var", C.c_double),
        ("azimuth_var", C.c_double),
        ("elevation_var", C.c_double),
        ("max_iterations", C.c_int32),
        ("optimizer", C.c_int32),
        ("rotation_epsilon", C.c_double),
        ("transformation_epsilon", C.c_double),
        ("lm_max_iterations", C.c_int32),
        ("lm_debug_print", C.c_int32),
        ("lm_init_lambda_factor", C.c_double),
        ("maha_fp64", C.c_int32),
        ("host_loop", C.c_int32),
        ("variant", C.c_int32),
        ("voxel_search", C.c_int32),
        ("voxel_resolution", C.c_double),
        ("voxel_mode", C.c_int32),
        ("reserved_", C.c_int32),
    ]


class ApdPair(C.Structure):
    _fields_ = [
        ("source", C.c_void_p),
        ("n_source", C.c_int32),
        ("target", C.c_void_p),
        ("n_target", C.c_int32),
        ("guess", C.c_void_p),
    ]


class ApdResult(C.Structure):
    _fields_ = [
        ("T", C.c_float * 16),
        ("fitness", C.c_double),
        ("converged", C.c_int32),
        ("iterations", C.c_int32),
        ("status", C.c_int32),
        ("n_inliers", C.c_int32),
    ]


class ApdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"apd status {code}: {msg}")
        self.code = code


# symbols every implementation of the ABI exports (without prefix)
CORE_SYMBOLS = [
    "default_params", "create", "destroy", "last_error", "set_params", "get_params",
    "set_source", "set_target", "swap_source_and_target", "clear_source", "clear_target",
    "set_source_covariances", "set_target_covariances", "get_source_covariances",
    "get_target_covariances", "get_neighbors", "align", "linearize", "compute_error",
    "update_correspondences", "get_correspondences", "get_mahalanobis", "fitness", "get_lm_trace",
]
# symbols only the CUDA library exports
PRODUCT_SYMBOLS = CORE_SYMBOLS + [
    "abi_version", "set_source_device", "set_target_device", "align_batch", "batch_create", "batch_destroy",
    "batch_set_params", "batch_align", "batch_align_device", "batch_launch_count", "batch_set_profiling",
    "batch_get_kernel_ms", "comm_unique_id",
    "comm_init", "comm_peer_handle", "comm_peer_attach", "comm_destroy", "group_create", "group_destroy", "group_size",
    "group_set_params", "group_set_source", "group_set_target", "group_set_source_device", "group_set_target_device", "group_align",
    "group_linearize", "group_compute_error", "nearest_k", "source_nearest", "batch_create_multi", "batch_device_pairs", "batch_get_load_stats", "debug_launch_rate", "debug_multi_align", "radius_search", "voxel_downsample", "submap_assemble",
    "align_batch_multi", "stream", "launch_count", "set_profiling", "get_kernel_ms",
]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _colmajor(m, dtype):
    """4x4 (or 6x6) numpy matrix -> flat column-major buffer (Eigen layout)."""
    return np.ascontiguousarray(np.asarray(m, dtype=dtype).T).reshape(-1)


class Registration:
    """Mirror of fast_gicp::FastAPDGICP over the C-ABI (method names follow the
    reference class, fast_apdgicp.hpp:47-88, in snake_case)."""

    def __init__(self, lib, prefix="apd_", device=0):
        self._lib = lib
        self._p = prefix
        self._h = C.c_void_p()
        if prefix == "apd_":
            self._call("create", C.c_int(device), C.byref(self._h), handle=False)
        else:
            self._call("create", C.byref(self._h), handle=False)
        self._keep = []
        self.n_source = 0
        self.n_target = 0

    # -- plumbing -------------------------------------------------------
    def _fn(self, name):
        return getattr(self._lib, self._p + name)

    def _call(self, name, *args, handle=True):
        fn = self._fn(name)
        fn.restype = C.c_int
        rc = fn(self._h, *args) if handle else fn(*args)
        if rc != APD_OK:
            raise ApdError(rc, self.last_error() if self._h else name)
        return rc

    def last_error(self):
        fn = self._fn("last_error")
        fn.restype = C.c_char_p
        s = fn(self._h)
        return s.decode() if s else ""

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters -----------------------------------------------------
    def get_params(self):
        p = ApdParams()
        self._call("get_params", C.byref(p))
        return p

    def set_params(self, **kw):
        p = self.get_params()
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        self._call("set_params", C.byref(p))
        return p

    # -- clouds ---------------------------------------------------------
    @staticmethod
    def _layout(cloud):
        """Accepts float32 [n,4] {x,y,z,label} or a structured PointXYZINormal array."""
        if cloud.dtype.names:  # structured (PCL AoS)
            a = np.ascontiguousarray(cloud)
            return a, a.shape[0], a.dtype.itemsize, a.dtype.fields["x"][1], a.dtype.fields["normal_x"][1]
        a = _f32(cloud)
        assert a.ndim == 2 and a.shape[1] in (3, 4)
        return a, a.shape[0], a.shape[1] * 4, 0, (12 if a.shape[1] == 4 else -1)

    def set_input_source(self, cloud, key=0):
        a, n, stride, xo, lo = self._layout(cloud)
        self._call("set_source", a.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_int32(stride), C.c_int32(xo),
                   C.c_int32(lo), C.c_uint64(key))
        self.n_source = n

    def set_input_target(self, cloud, key=0):
        a, n, stride, xo, lo = self._layout(cloud)
        self._call("set_target", a.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_int32(stride), C.c_int32(xo),
                   C.c_int32(lo), C.c_uint64(key))
        self.n_target = n

    def set_input_source_device(self, dptr, n):
        self._call("set_source_device", C.c_void_p(dptr), C.c_int32(n))
        self.n_source = n

    def set_input_target_device(self, dptr, n):
        self._call("set_target_device", C.c_void_p(dptr), C.c_int32(n))
        self.n_target = n

    def swap_source_and_target(self):
        self._call("swap_source_and_target")
        self.n_source, self.n_target = self.n_target, self.n_source

    def clear_source(self):
        self._call("clear_source")
        self.n_source = 0

    def clear_target(self):
        self._call("clear_target")
        self.n_target = 0

    def _get_covs(self, name, n):
        out = np.empty((n, 16), dtype=np.float64)
        self._call(name, out.ctypes.data_as(C.c_void_p), C.c_int32(n))
        return out.reshape(n, 4, 4).transpose(0, 2, 1).copy()  # column-major -> [n,4,4]

    def get_source_covariances(self):
        return self._get_covs("get_source_covariances", self.n_source)

    def get_target_covariances(self):
        return self._get_covs("get_target_covariances", self.n_target)

    def _set_covs(self, name, covs):
        a = np.ascontiguousarray(np.asarray(covs, dtype=np.float64).transpose(0, 2, 1)).reshape(-1, 16)
        self._call(name, a.ctypes.data_as(C.c_void_p), C.c_int32(a.shape[0]))

    def set_source_covariances(self, covs):
        self._set_covs("set_source_covariances", covs)

    def set_target_covariances(self, covs):
        self._set_covs("set_target_covariances", covs)

    def get_neighbors(self, which):
        n = self.n_source if which == 0 else self.n_target
        k = self.get_params().k_correspondences
        out = np.empty((n, k), dtype=np.int32)
        self._call("get_neighbors", C.c_int32(which), out.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_int32(k))
        return out

    # -- hot path -------------------------------------------------------
    def align(self, guess=None, want_aligned=False):
        """Returns dict(T float32 4x4, T64, H 6x6, converged, iterations[, aligned])."""
        g = None if guess is None else _colmajor(guess, np.float32)
        T = np.empty(16, np.float32)
        T64 = np.empty(16, np.float64)
        H = np.empty(36, np.float64)
        conv, it = C.c_int32(), C.c_int32()
        aligned = np.empty((self.n_source, 3), np.float32) if want_aligned else None
        self._call(
            "align",
            g.ctypes.data_as(C.c_void_p) if g is not None else None,
            T.ctypes.data_as(C.c_void_p), T64.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p),
            C.byref(conv), C.byref(it),
            aligned.ctypes.data_as(C.c_void_p) if aligned is not None else None,
        )
        out = dict(T=T.reshape(4, 4).T.copy(), T64=T64.reshape(4, 4).T.copy(), H=H.reshape(6, 6).T.copy(),
                   converged=bool(conv.value), iterations=it.value)
        if want_aligned:
            out["aligned"] = aligned
        return out

    def linearize(self, T, want_hb=True):
        t = _colmajor(T, np.float64)
        H = np.empty(36, np.float64)
        b = np.empty(6, np.float64)
        err = C.c_double()
        self._call("linearize", t.ctypes.data_as(C.c_void_p),
                   H.ctypes.data_as(C.c_void_p) if want_hb else None,
                   b.ctypes.data_as(C.c_void_p) if want_hb else None, C.byref(err))
        if want_hb:
            return err.value, H.reshape(6, 6).T.copy(), b
        return err.value

    def compute_error(self, T):
        t = _colmajor(T, np.float64)
        err = C.c_double()
        self._call("compute_error", t.ctypes.data_as(C.c_void_p), C.byref(err))
        return err.value

    def update_correspondences(self, T):
        t = _colmajor(T, np.float64)
        self._call("update_correspondences", t.ctypes.data_as(C.c_void_p))

    def get_correspondences(self):
        idx = np.empty(self.n_source, np.int32)
        sq = np.empty(self.n_source, np.float32)
        self._call("get_correspondences", idx.ctypes.data_as(C.c_void_p), sq.ctypes.data_as(C.c_void_p),
                   C.c_int32(self.n_source))
        return idx, sq

    def get_mahalanobis(self):
        return self._get_covs("get_mahalanobis", self.n_source)

    # -- FastVGICP parity hooks -------------------------------------------
    N_OFFSETS = {0: 27, 1: 7, 2: 1}  # APD_VOXEL_DIRECT27 / 7 / 1

    def vgicp_voxels(self):
        """the Gaussian voxel map of the target: (coords int32[n,3], counts int32[n], means f64[n,3], covs f64[n,3,3])"""
        n = C.c_int32()
        self._call("vgicp_get_voxels", C.byref(n), None, None, None, None, C.c_int32(0))
        m = n.value
        coords, counts = np.zeros((m, 3), np.int32), np.zeros(m, np.int32)
        means, covs = np.zeros((m, 3), np.float64), np.zeros((m, 3, 3), np.float64)
        self._call("vgicp_get_voxels", C.byref(n), coords.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                   means.ctypes.data_as(C.c_void_p), covs.ctypes.data_as(C.c_void_p), C.c_int32(m))
        return coords, counts, means, covs

    def vgicp_correspondences(self):
        """(voxel int32[n_source, n_offsets] (-1: none), maha f64[n_source, n_offsets, 3, 3]) of the last linearize"""
        no = self.N_OFFSETS[self.get_params().voxel_search]
        vox = np.zeros((self.n_source, no), np.int32)
        maha = np.zeros((self.n_source, no, 3, 3), np.float64)
        self._call("vgicp_get_correspondences", vox.ctypes.data_as(C.c_void_p), maha.ctypes.data_as(C.c_void_p), C.c_int32(self.n_source), C.c_int32(no))
        return vox, maha

    def fitness(self, T=None, max_range=np.finfo(np.float64).max, inlier_sq_thr=0.25):
        t = None if T is None else _colmajor(T, np.float32)
        score, nr, ni = C.c_double(), C.c_int32(), C.c_int32()
        self._call("fitness", t.ctypes.data_as(C.c_void_p) if t is not None else None, C.c_double(max_range),
                   C.byref(score), C.byref(nr), C.c_double(inlier_sq_thr), C.byref(ni))
        return score.value, nr.value, ni.value

    def nearest_k(self, queries, k, which=1):
        """exact k nearest neighbours of arbitrary query points [n,3+] in the target (which=1) or source (0) cloud:
        (indices [n,k] int32, squared distances [n,k] float32), ascending by (d2, index)"""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        n = q.shape[0]
        idx = np.empty((n, k), np.int32)
        d2 = np.empty((n, k), np.float32)
        self._call("nearest_k", C.c_int32(which), q.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_int32(q.shape[1] * 4), C.c_int32(k),
                   idx.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p))
        return idx, d2

    def source_nearest(self, T=None):
        """nearest target point of every source point under pose T (None: the final transformation):
        (index [n], squared distance [n], transformed xyz [n,3])"""
        t = None if T is None else _colmajor(T, np.float32)
        n = self.n_source
        idx, d2, xyz = np.empty(n, np.int32), np.empty(n, np.float32), np.empty((n, 3), np.float32)
        self._call("source_nearest", t.ctypes.data_as(C.c_void_p) if t is not None else None, idx.ctypes.data_as(C.c_void_p),
                   d2.ctypes.data_as(C.c_void_p), xyz.ctypes.data_as(C.c_void_p), C.c_int32(n))
        return idx, d2, xyz

    def radius_search(self, radius, which=1, lists=False):
        """neighbours within `radius` of every point of the cloud (pcl radiusSearch semantics: d2 < r^2, the point itself
        included): counts [n]; with lists=True also (offsets [n+1], indices) in CSR form"""
        n = self.n_source if which == 0 else self.n_target
        counts = np.empty(n, np.int32)
        if not lists:
            self._call("radius_search", C.c_int32(which), C.c_double(radius), counts.ctypes.data_as(C.c_void_p), None, None, C.c_int64(0), C.c_int32(n))
            return counts
        offsets = np.empty(n + 1, np.int64)
        self._call("radius_search", C.c_int32(which), C.c_double(radius), counts.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), None,
                   C.c_int64(0), C.c_int32(n))
        indices = np.empty(int(offsets[-1]), np.int32)
        self._call("radius_search", C.c_int32(which), C.c_double(radius), None, offsets.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p),
                   C.c_int64(indices.shape[0]), C.c_int32(n))
        return counts, offsets, indices

    def dbscan_labels(self, eps=0.9, core_min_pts=10, min_cluster=20, max_cluster=25000, which=0):
        """(labels float32[n], n_clusters): DBSCANKdtreeCluster + the preprocessing nodelet's ranking of the clusters"""
        n = self.n_source if which == 0 else self.n_target
        labels = np.zeros(n, np.float32)
        nc = C.c_int32()
        self._call("dbscan_labels", C.c_int32(which), C.c_double(eps), C.c_int32(core_min_pts), C.c_int32(min_cluster), C.c_int32(max_cluster),
                   labels.ctypes.data_as(C.c_void_p), C.byref(nc), C.c_int32(n))
        return labels, nc.value

    def voxel_downsample(self, cloud, leaf):
        """pcl::VoxelGrid with a cubic leaf: float32 [m,4] {x,y,z,label}"""
        a, n, stride, xo, lo = self._layout(cloud)
        out = np.empty((max(n, 1), 4), np.float32)
        m = C.c_int32()
        self._call("voxel_downsample", a.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_int32(stride), C.c_int32(xo), C.c_int32(lo), C.c_double(leaf),
                   out.ctypes.data_as(C.c_void_p), C.c_int32(out.shape[0]), C.byref(m))
        return out[: m.value].copy()

    def submap_assemble(self, clouds, poses, leaf, set_as_target=False):
        """scan_matching_odometry_nodelet.cpp:602-618: clouds[c] moved by poses[c] (4x4), concatenated, voxel-downsampled"""

        class Ref(C.Structure):
            _fields_ = [("pts", C.c_void_p), ("n", C.c_int32)]

        arrs = [_f32(c) for c in clouds]
        refs = (Ref * len(arrs))()
        for i, a in enumerate(arrs):
            refs[i].pts, refs[i].n = a.ctypes.data, a.shape[0]
        P = np.ascontiguousarray(np.stack([np.asarray(T, np.float64).T for T in poses])).reshape(-1)
        total = sum(a.shape[0] for a in arrs)
        out = np.empty((max(total, 1), 4), np.float32)
        m = C.c_int32()
        self._call("submap_assemble", refs, P.ctypes.data_as(C.c_void_p), C.c_int32(len(arrs)), C.c_int32(16), C.c_int32(0), C.c_int32(12),
                   C.c_double(leaf), C.c_int32(1 if set_as_target else 0), out.ctypes.data_as(C.c_void_p), C.c_int32(out.shape[0]), C.byref(m))
        if set_as_target:
            self.n_target = m.value
        return out[: m.value].copy()

    def lm_trace(self, max_rows=1024):
        rows = np.zeros((max_rows, 8), np.float64)
        n = C.c_int32()
        self._call("get_lm_trace", rows.ctypes.data_as(C.c_void_p), C.c_int32(max_rows), C.byref(n))
        return rows[: n.value].copy()

    # -- product-only instrumentation ------------------------------------
    def stream(self):
        fn = self._fn("stream")
        fn.restype = C.c_void_p
        return fn(self._h)

    def launch_count(self):
        fn = self._fn("launch_count")
        fn.restype = C.c_int64
        return fn(self._h)

    def set_profiling(self, on):
        self._call("set_profiling", C.c_int32(1 if on else 0))

    def kernel_ms(self):
        ms = (C.c_double * len(KERNEL_CLASSES))()
        cnt = (C.c_int64 * len(KERNEL_CLASSES))()
        self._call("get_kernel_ms", ms, cnt)
        return {k: (ms[i], cnt[i]) for i, k in enumerate(KERNEL_CLASSES)}

    def comm_init(self, id128, rank, nranks, n_source_total):
        buf = (C.c_char * 128).from_buffer_copy(bytes(id128))
        self._call("comm_init", buf, C.c_int32(rank), C.c_int32(nranks), C.c_int64(n_source_total))

    def comm_peer_handle(self):
        buf = (C.c_char * 64)()
        self._call("comm_peer_handle", buf)
        return bytes(buf)

    def comm_peer_attach(self, handles):
        blob = b"".join(handles)
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._call("comm_peer_attach", buf)

    def comm_destroy(self):
        self._call("comm_destroy")
