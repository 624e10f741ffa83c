"""Loads the CPU oracle (oracle/, TEST INFRASTRUCTURE) behind the same Python
wrapper the CUDA library uses, so parity tests run identical call sequences."""
import ctypes
import importlib
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libapd_oracle.so")
ORACLE_REF_SO = os.path.join(ORACLE_DIR, "_ref", "libapd_oracle_ref.so")

gorio = importlib.import_module("go-rio_b200")
_libs = {}


def build_oracle():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("apd_oracle.cpp", "apd_oracle_capi.cpp", "apd_prep_oracle.cpp", "apd_vgicp_oracle.cpp", "apd_oracle.hpp", "apd_math.hpp")]
    stale = (not os.path.exists(ORACLE_SO)) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], env={**os.environ, "MAKEFLAGS": ""})


def oracle_lib(ref=False):
    path = ORACLE_REF_SO if ref else ORACLE_SO
    if path not in _libs:
        if not ref:
            build_oracle()
        _libs[path] = ctypes.CDLL(path)
    return _libs[path]


def Oracle(search=1, threads=1, ref=False):
    """search: 0 brute force, 1 exact kd-tree, 2 nanoflann (ref build only)."""
    lib = oracle_lib(ref)
    o = gorio.Registration(lib, "apdo_")
    assert lib.apdo_set_search(o._h, ctypes.c_int(search)) == 0
    lib.apdo_set_num_threads(o._h, ctypes.c_int(threads))
    return o
