#!/usr/bin/env python
"""Writes tests/golden/gicp_reference.npz from the REFERENCE's own FastGICP / FastVGICP code
(oracle/_ref/libapd_ref_gicp.so = fast_gicp/gicp/{fast_gicp,fast_vgicp,fast_vgicp_voxel}.hpp + impl/ compiled in place
against the stand-ins of oracle/ref_stubs; `make -C oracle`). Run in the container that holds /root/reference:
    python tests/golden/make_gicp_reference.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_apdgicp_reference as mk  # noqa: E402  (clouds, poses)

REPO = os.path.dirname(os.path.dirname(HERE))
FLT_MAX = mk.FLT_MAX
DEFAULTS = dict(k=20, regularization=3, max_corr=FLT_MAX, max_iterations=64, gauss_newton=0, rot_eps=2e-3, trans_eps=5e-4, threads=1,
                voxel_resolution=1.0, voxel_search=2, voxel_mode=0)
VGICP_CASES = [("d1_r10_add", dict(voxel_search=2, voxel_resolution=1.0, voxel_mode=0)), ("d7_r15_add", dict(voxel_search=1, voxel_resolution=1.5, voxel_mode=0)),
               ("d27_r20_addw", dict(voxel_search=0, voxel_resolution=2.0, voxel_mode=1)), ("d7_r15_mul", dict(voxel_search=1, voxel_resolution=1.5, voxel_mode=2))]


class Ref:
    """ctypes face of oracle/ref_gicp.cpp; variant 1 FastGICP, 2 FastVGICP"""

    def __init__(self, lib, variant, **kw):
        self.lib, self.variant = lib, variant
        lib.gref_create.restype = C.c_void_p
        lib.gref_linearize.restype = C.c_double
        lib.gref_compute_error.restype = C.c_double
        self.h = C.c_void_p(lib.gref_create(C.c_int(variant)))
        p = dict(DEFAULTS); p.update(kw)
        lib.gref_set_params(self.h, C.c_int(p["k"]), C.c_int(p["regularization"]), C.c_double(p["max_corr"]), C.c_int(p["max_iterations"]),
                            C.c_int(p["gauss_newton"]), C.c_double(p["rot_eps"]), C.c_double(p["trans_eps"]), C.c_int(p["threads"]),
                            C.c_double(p["voxel_resolution"]), C.c_int(p["voxel_search"]), C.c_int(p["voxel_mode"]))

    def set_clouds(self, src, tgt):
        self.src, self.tgt = np.ascontiguousarray(src, np.float32), np.ascontiguousarray(tgt, np.float32)
        self.lib.gref_set_clouds(self.h, self.src.ctypes.data_as(C.c_void_p), C.c_int(self.src.shape[0]), self.tgt.ctypes.data_as(C.c_void_p),
                                 C.c_int(self.tgt.shape[0]))

    def linearize(self, T):
        t = np.ascontiguousarray(np.asarray(T, np.float64).T).reshape(-1)
        H, b = np.zeros((6, 6)), np.zeros(6)
        e = self.lib.gref_linearize(self.h, t.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
        return e, H, b

    def compute_error(self, T):
        t = np.ascontiguousarray(np.asarray(T, np.float64).T).reshape(-1)
        return self.lib.gref_compute_error(self.h, t.ctypes.data_as(C.c_void_p))

    def correspondences(self):
        n = self.src.shape[0]
        idx, sqd, maha = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 3, 3))
        self.lib.gref_get_correspondences(self.h, idx.ctypes.data_as(C.c_void_p), sqd.ctypes.data_as(C.c_void_p), maha.ctypes.data_as(C.c_void_p))
        return idx, sqd, maha

    def voxels(self):
        m = self.lib.gref_get_voxels(self.h, None, None, None, None)
        coords, counts, means, covs = np.zeros((m, 3), np.int32), np.zeros(m, np.int32), np.zeros((m, 3)), np.zeros((m, 3, 3))
        self.lib.gref_get_voxels(self.h, coords.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), means.ctypes.data_as(C.c_void_p),
                                 covs.ctypes.data_as(C.c_void_p))
        return coords, counts, means, covs

    def voxel_correspondences(self):
        m = self.lib.gref_get_voxel_correspondences(self.h, None, None, None)
        si, vi, maha = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros((m, 3, 3))
        self.lib.gref_get_voxel_correspondences(self.h, si.ctypes.data_as(C.c_void_p), vi.ctypes.data_as(C.c_void_p), maha.ctypes.data_as(C.c_void_p))
        return si, vi, maha

    def align(self, guess=None):
        g = None if guess is None else np.ascontiguousarray(np.asarray(guess, np.float32).T).reshape(-1)
        T, H = np.zeros(16, np.float32), np.zeros((6, 6))
        conv, it = C.c_int(), C.c_int()
        self.lib.gref_align(self.h, g.ctypes.data_as(C.c_void_p) if g is not None else None, T.ctypes.data_as(C.c_void_p), C.byref(conv), C.byref(it),
                            H.ctypes.data_as(C.c_void_p))
        return T.reshape(4, 4).T.copy(), bool(conv.value), it.value, H

    def swap(self):
        self.lib.gref_swap(self.h)
        self.src, self.tgt = self.tgt, self.src


def load_lib():
    return C.CDLL(os.path.join(REPO, "oracle", "_ref", "libapd_ref_gicp.so"))


def main():
    lib = load_lib()
    out = {}
    for cname in mk.CLOUDS:
        src, tgt, Tgt = mk.clouds(cname)
        P = mk.poses(Tgt)
        # ---- FastGICP ----
        for thr_name, thr in (("thr2", 2.0), ("nothr", FLT_MAX)):
            r = Ref(lib, 1, max_corr=thr)
            r.set_clouds(src, tgt)
            for pi, T in enumerate(P):
                key = f"gicp_{cname}_{thr_name}_p{pi}"
                e, H, b = r.linearize(T)
                idx, sqd, maha = r.correspondences()
                out[key + "_err"], out[key + "_H"], out[key + "_b"] = np.float64(e), H, b
                out[key + "_corr"], out[key + "_sqd"], out[key + "_maha"] = idx, sqd, maha
                out[key + "_err_trial_stale"] = np.float64(r.compute_error(P[(pi + 1) % len(P)]))
        for aname, kw in (("lm_default", {}), ("lm_deployed", dict(max_corr=2.0, trans_eps=0.1)), ("gn_thr2", dict(max_corr=2.0, gauss_newton=1))):
            r = Ref(lib, 1, **kw)
            r.set_clouds(src, tgt)
            T, conv, it, H = r.align()
            key = f"gicp_{cname}_align_{aname}"
            out[key + "_T"], out[key + "_converged"], out[key + "_iterations"], out[key + "_H"] = T, np.bool_(conv), np.int32(it), H
            r.swap()
            T2, conv2, it2, _ = r.align()
            out[key + "_swapped_T"], out[key + "_swapped_converged"], out[key + "_swapped_iterations"] = T2, np.bool_(conv2), np.int32(it2)
            print("gicp", cname, aname, conv, it, "| swapped", conv2, it2)
        # ---- FastVGICP ----
        for vname, vkw in VGICP_CASES:
            r = Ref(lib, 2, **vkw)
            r.set_clouds(src, tgt)
            for pi, T in enumerate(P[1:3]):
                key = f"vgicp_{cname}_{vname}_p{pi}"
                e, H, b = r.linearize(T)
                out[key + "_err"], out[key + "_H"], out[key + "_b"] = np.float64(e), H, b
                si, vi, maha = r.voxel_correspondences()
                out[key + "_vsrc"], out[key + "_vvox"], out[key + "_vmaha"] = si, vi, maha
                out[key + "_err_trial_stale"] = np.float64(r.compute_error(P[3]))
            coords, counts, means, covs = r.voxels()
            key = f"vgicp_{cname}_{vname}"
            out[key + "_coords"], out[key + "_counts"], out[key + "_means"], out[key + "_covs"] = coords, counts, means, covs
            ra = Ref(lib, 2, trans_eps=0.01, **vkw)
            ra.set_clouds(src, tgt)
            T, conv, it, H = ra.align()
            out[key + "_align_T"], out[key + "_align_converged"], out[key + "_align_iterations"], out[key + "_align_H"] = T, np.bool_(conv), np.int32(it), H
            print("vgicp", cname, vname, "voxels", coords.shape[0], "corr", si.shape[0], "align", conv, it)
    np.savez_compressed(os.path.join(HERE, "gicp_reference.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
