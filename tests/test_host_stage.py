"""Host-side preparation of a cloud (go-rio_b200/csrc/host_stage.hpp, included by the library): staging of the caller's
AoS points into packed {x,y,z,label} with the bounding box found on the way, and the sizing of the uniform grid.
Compiled with g++ and checked on the CPU: every layout gives the same bytes as a NumPy gather, NaNs do not poison the
box, and the grid obeys the limits the kernels rely on."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hs(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hs") / "libhs.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O3", "-fPIC", "-shared", "-I" + os.path.join(REPO, "go-rio_b200", "csrc"),
                           os.path.join(REPO, "tests", "host_stage_capi.cpp"), "-o", out])
    lib = ctypes.CDLL(out)
    lib.hs_size_grid.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    return lib


def aligned(n_floats):
    raw = np.zeros(n_floats + 8, np.float32)
    off = (-raw.ctypes.data % 16) // 4
    return raw[off:off + n_floats]


def stage(hs, buf, n, stride, xyz_off, label_off):
    dst = aligned(4 * max(n, 1))
    bbox = np.zeros(6, np.float32)
    hs.hs_stage_cloud(buf.ctypes.data_as(ctypes.c_void_p), n, stride, xyz_off, label_off, dst.ctypes.data_as(ctypes.c_void_p),
                      bbox.ctypes.data_as(ctypes.c_void_p))
    return dst[:4 * n].reshape(n, 4), bbox


@pytest.mark.parametrize("n", [0, 1, 2, 3, 777, 1000, 60001])
def test_every_layout_stages_the_same_points(hs, synth, n):
    rng = np.random.default_rng(n)
    pts = np.ascontiguousarray((rng.normal(size=(n, 4)) * [30, 30, 3, 1]).astype(np.float32))
    pts[:, 3] = rng.integers(0, 9, size=n)
    want_box = np.concatenate([pts[:, :3].min(axis=0), pts[:, :3].max(axis=0)]) if n else None
    # packed float4
    d, b = stage(hs, pts, n, 16, 0, 12)
    assert np.array_equal(d, pts)
    if n:
        assert np.array_equal(b, want_box)
        b2 = np.zeros(6, np.float32)
        hs.hs_bounds_of_packed(pts.ctypes.data_as(ctypes.c_void_p), n, b2.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(b2, want_box)
    # pcl::PointXYZINormal (48 bytes: x y z 1 | normal_x normal_y normal_z 0 | intensity curvature pad pad)
    if n:
        pcl = synth.to_pcl_xyzinormal(pts)
        raw = np.frombuffer(pcl.tobytes(), dtype=np.uint8).copy()
        assert pcl.dtype.itemsize == 48
        d, b = stage(hs, raw, n, 48, pcl.dtype.fields["x"][1], pcl.dtype.fields["normal_x"][1])
        assert np.array_equal(d, pts) and np.array_equal(b, want_box)
    # a layout without a fast path: stride 24, xyz at 4, label at 16; and one without a label
    wide = np.zeros((n, 6), np.float32)
    wide[:, 1:5] = pts
    d, b = stage(hs, wide, n, 24, 4, 16)
    assert np.array_equal(d, pts)
    d, b = stage(hs, wide, n, 24, 4, -1)
    assert np.array_equal(d[:, :3], pts[:, :3]) and not d[:, 3].any()
    if n:
        assert np.array_equal(b, want_box)


def test_nan_points_do_not_poison_the_box(hs):
    pts = np.array([[1, 2, 3, 0], [np.nan, 5, 6, 0], [-4, np.nan, 9, 0], [7, 8, np.nan, 0], [0, 0, 0, 0]], np.float32)
    for n in (5, 4):  # both the 4-chain body and its tail
        _, b = stage(hs, pts, n, 16, 0, 12)
        b2 = np.zeros(6, np.float32)
        hs.hs_bounds_of_packed(pts.ctypes.data_as(ctypes.c_void_p), n, b2.ctypes.data_as(ctypes.c_void_p))
        want = np.concatenate([np.nanmin(pts[:n, :3], axis=0), np.nanmax(pts[:n, :3], axis=0)])
        assert np.array_equal(b, want) and np.array_equal(b2, want)


@pytest.mark.parametrize("n,ext", [(1000, (120, 170, 14)), (2000, (100, 100, 0.0)), (60000, (140, 200, 16)), (20_000_000, (2000, 2000, 25)),
                                   (50, (1e-4, 1e-4, 1e-4)), (5000, (1e5, 3, 3)), (1, (0, 0, 0))])
@pytest.mark.parametrize("cpp", [0.5, 4.0, 8.0])
def test_grid_sizing_obeys_the_kernels_limits(hs, n, ext, cpp):
    lo = np.array([-3.25, 10.5, -1.75], np.float32)
    bbox = np.concatenate([lo, lo + np.array(ext, np.float32)]).astype(np.float32)
    out = np.zeros(8)
    ncells = hs.hs_size_grid(bbox.ctypes.data_as(ctypes.c_void_p), n, cpp, out.ctypes.data_as(ctypes.c_void_p))
    ox, oy, oz, inv_cell, cell, nx, ny, nz = out
    assert (ox, oy, oz) == tuple(float(v) for v in lo)
    assert 1 <= nx <= 2049 and 1 <= ny <= 2049 and 1 <= nz <= 2049 and ncells == nx * ny * nz <= (1 << 28) + (1 << 24)
    assert cell > 0 and abs(cell * inv_cell - 1) < 1e-6
    # the max corner maps inside the grid with the kernels' own fp32 expression (common.cuh: cell_coord)
    for mx, o, dim in zip(bbox[3:], lo, (nx, ny, nz)):
        c = np.floor((np.float32(mx) - np.float32(o)) * np.float32(inv_cell))
        assert 0 <= c <= dim - 1
    # about cells_per_point cells per point where the extent allows it
    if n >= 1000 and min(ext) > 1 and max(ext) < 1e4:
        assert 0.2 * cpp * n <= ncells <= 2.5 * max(64.0, cpp * n)
