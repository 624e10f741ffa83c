"""The serial part of the optimizer (go-rio_b200/csrc/host_math.hpp: pivoted 6x6 LDL^T, so3_exp + pose assembly, pose
composition, is_converged — reference lsq_registration_impl.hpp:83-173, so3.hpp:59-78) is one header that the host LM
loop and the device-resident loop kernel both compile. Here it is compiled with g++ and pinned on the CPU against
NumPy and against the independent restatement in tests/numpy_restatement.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import numpy_restatement as nr

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hm") / "libhm.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                           "-I" + os.path.join(REPO, "go-rio_b200", "csrc"), os.path.join(REPO, "tests", "host_math_capi.cpp"), "-o", out])
    return ctypes.CDLL(out)


def _p(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.c_void_p)


def test_ldlt_solve6_matches_numpy(hm):
    rng = np.random.default_rng(7)
    for trial in range(200):
        J = rng.normal(size=(40, 6)) * rng.uniform(1e-3, 1e3, size=6)  # badly scaled columns, as H's rotation / translation blocks are
        H = J.T @ J
        if trial % 3 == 0:
            H += 1e-9 * np.abs(np.diag(H)).max() * np.eye(6)  # the LM damping of the first trial (lsq :131-138)
        b = rng.normal(size=6)
        x = np.empty(6)
        assert hm.hm_ldlt_solve6(_p(H), _p(b), x.ctypes.data_as(ctypes.c_void_p)) == 1
        ref = np.linalg.solve(H, b)
        assert np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max() * max(1.0, np.linalg.cond(H) * 1e-7)
        assert np.abs(H @ x - b).max() <= 1e-9 * (np.abs(H).max() * np.abs(x).max() + np.abs(b).max())
    # an indefinite but regular matrix (LDL^T, not Cholesky) and a singular one
    H = np.diag([4.0, -3.0, 2.0, 1.0, -5.0, 6.0]); H[0, 1] = H[1, 0] = 0.5
    b = np.arange(1.0, 7.0); x = np.empty(6)
    assert hm.hm_ldlt_solve6(_p(H), _p(b), x.ctypes.data_as(ctypes.c_void_p)) == 1
    assert np.allclose(x, np.linalg.solve(H, b), rtol=1e-12, atol=0)
    assert hm.hm_ldlt_solve6(_p(np.zeros((6, 6))), _p(b), x.ctypes.data_as(ctypes.c_void_p)) == 0


def test_delta_from_twist_is_so3_exp_plus_translation(hm):
    rng = np.random.default_rng(8)
    twists = [rng.normal(size=6) * s for s in (1e-9, 1e-6, 1e-5, 1e-3, 0.1, 1.0, 3.0)] + [np.zeros(6)]
    for d in twists:
        out = np.empty(16)
        hm.hm_delta_from_twist(_p(d), out.ctypes.data_as(ctypes.c_void_p))
        P = out.reshape(4, 4)
        assert np.abs(P[:3, :3] - nr.so3_exp(d[:3])).max() < 1e-15  # the same formulas (Taylor branch below theta^2 = 1e-10)
        assert np.array_equal(P[:3, 3], d[3:]) and np.array_equal(P[3], [0, 0, 0, 1])
        assert np.abs(P[:3, :3] @ P[:3, :3].T - np.eye(3)).max() < 1e-14
        # against the closed form of the exponential map
        th = np.linalg.norm(d[:3])
        if th > 1e-4:
            k = d[:3] / th
            K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            assert np.abs(P[:3, :3] - (np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K)).max() < 1e-14


def test_compose_and_is_converged(hm):
    rng = np.random.default_rng(9)
    for _ in range(50):
        A, B = np.eye(4), np.eye(4)
        A[:3, :3], B[:3, :3] = nr.so3_exp(rng.normal(size=3)), nr.so3_exp(rng.normal(size=3))
        A[:3, 3], B[:3, 3] = rng.normal(size=3) * 10, rng.normal(size=3) * 10
        C = np.empty(16)
        hm.hm_compose(_p(A), _p(B), C.ctypes.data_as(ctypes.c_void_p))
        assert np.abs(C.reshape(4, 4) - A @ B).max() < 1e-13
    hm.hm_is_converged.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
    # lsq_registration_impl.hpp:83-92 with the deployed epsilons (rotation 2e-3, translation 0.1)
    for d, expect in (([1e-4, 0, 0, 0.05, 0, 0], 1), ([1e-4, 0, 0, 0.11, 0, 0], 0), ([0, 0, 3e-3, 0, 0, 0], 0), ([0, 0, 1.9e-3, 0, 0.09, 0], 1)):
        P = np.empty(16)
        hm.hm_delta_from_twist(_p(np.array(d, dtype=np.float64)), P.ctypes.data_as(ctypes.c_void_p))
        assert hm.hm_is_converged(_p(P), 2e-3, 0.1) == expect, d


def test_unpack_upper_and_guess_layout(hm):
    u = np.arange(1.0, 22.0)
    H = np.empty(36)
    hm.hm_unpack_upper(_p(u), H.ctypes.data_as(ctypes.c_void_p))
    H = H.reshape(6, 6)
    assert np.array_equal(H, H.T) and np.array_equal(H[np.triu_indices(6)], u)
    g = np.arange(16, dtype=np.float32)  # column-major 4x4, as Eigen::Matrix4f::data() hands it over
    P = np.empty(16)
    hm.hm_from_colmajor_f32(g.ctypes.data_as(ctypes.c_void_p), P.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(P.reshape(4, 4), g.reshape(4, 4).T.astype(np.float64))
