"""Minimal reader / writer of binary and ascii .pcd files (PCL's point cloud format, v0.7) for the harness: the
reference's own sample clouds (ndt_omp/data/*.pcd) and dumps of real radar frames are .pcd files. Fields of type F/I/U,
any SIZE, COUNT 1; `DATA binary` and `DATA ascii` (not binary_compressed)."""
import numpy as np

_TYPES = {("F", 4): "<f4", ("F", 8): "<f8", ("I", 1): "<i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8",
          ("U", 1): "<u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8"}


def read_pcd(path):
    """returns a structured numpy array with the file's fields"""
    with open(path, "rb") as f:
        raw = f.read()
    header, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, *vals = line.split()
        header[key.upper()] = vals
        if key.upper() == "DATA":
            break
    fields, sizes, types = header["FIELDS"], [int(v) for v in header["SIZE"]], header["TYPE"]
    counts = [int(v) for v in header.get("COUNT", ["1"] * len(fields))]
    if any(c != 1 for c in counts):
        raise ValueError("COUNT != 1 is not supported")
    n = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header["HEIGHT"][0])
    dtype = np.dtype([(nm, _TYPES[(t, s)]) for nm, t, s in zip(fields, types, sizes)])
    kind = header["DATA"][0].lower()
    if kind == "binary":
        return np.frombuffer(raw, dtype=dtype, count=n, offset=pos).copy()
    if kind == "ascii":
        rows = np.loadtxt(raw[pos:].decode("ascii").splitlines(), ndmin=2)
        out = np.zeros(rows.shape[0], dtype=dtype)
        for j, nm in enumerate(fields):
            out[nm] = rows[:, j]
        return out
    raise ValueError(f"DATA {kind} is not supported")


def write_pcd(path, cloud, binary=True):
    """cloud: structured array (fields written as they are) or float32 [n,3|4] (x y z [intensity])"""
    if cloud.dtype.names is None:
        names = ["x", "y", "z", "intensity"][: cloud.shape[1]]
        st = np.zeros(cloud.shape[0], dtype=[(nm, "<f4") for nm in names])
        for j, nm in enumerate(names):
            st[nm] = cloud[:, j]
        cloud = st
    names = cloud.dtype.names
    kinds = {"f": "F", "i": "I", "u": "U"}
    hdr = ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7", "FIELDS " + " ".join(names),
           "SIZE " + " ".join(str(cloud.dtype[nm].itemsize) for nm in names), "TYPE " + " ".join(kinds[cloud.dtype[nm].kind] for nm in names),
           "COUNT " + " ".join("1" for _ in names), f"WIDTH {cloud.shape[0]}", "HEIGHT 1", "VIEWPOINT 0 0 0 1 0 0 0", f"POINTS {cloud.shape[0]}",
           "DATA " + ("binary" if binary else "ascii")]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        if binary:
            f.write(np.ascontiguousarray(cloud).tobytes())
        else:
            for row in cloud:
                f.write((" ".join(repr(float(v)) for v in row) + "\n").encode("ascii"))


def xyz_label(cloud, label_field=None):
    """structured pcd array -> float32 [n,4] {x,y,z,label} with non-finite points dropped (label 0 unless a field is named)"""
    out = np.zeros((cloud.shape[0], 4), np.float32)
    out[:, 0], out[:, 1], out[:, 2] = cloud["x"], cloud["y"], cloud["z"]
    if label_field is not None:
        out[:, 3] = cloud[label_field]
    return np.ascontiguousarray(out[np.isfinite(out[:, :3]).all(axis=1)])
