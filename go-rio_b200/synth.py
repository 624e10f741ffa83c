"""Deterministic synthetic 4D-radar-shaped clouds for tests and bench.py.

Follows SURVEY.md §8(d): sensor at the origin, 120 deg x 30 deg field of view,
range 1-100 m with density ~ r^0.9, a scene of ground / vertical planes / blobs /
clutter, per-point sensor noise drawn from the reference's own noise model
(sigma_range = r*0.86/400, sigma_az = 0.5 deg, sigma_el = 1.0 deg,
reference fast_apdgicp.hpp:116-118) and a cluster label per point that mimics the
DBSCAN rank the reference's preprocessing writes into ``normal_x``
(4DRadarSLAM/apps/preprocessing_nodelet_ntu.cpp:558-567).

Clouds are float32 arrays [n, 4] = {x, y, z, label}; ``to_pcl_xyzinormal`` packs
them into the 48-byte pcl::PointXYZINormal layout the reference's callers use.
"""
import numpy as np

FOV_AZ = np.deg2rad(60.0)
FOV_EL = np.deg2rad(15.0)
R_MIN, R_MAX = 1.0, 100.0
GROUND_Z = -1.5

# pcl::PointXYZINormal [PCL 1.10 point_types.hpp]: 48 bytes
PCL_XYZINORMAL = np.dtype(
    {
        "names": ["x", "y", "z", "intensity", "normal_x", "normal_y", "normal_z", "curvature"],
        "formats": ["<f4"] * 8,
        "offsets": [0, 4, 8, 12, 16, 20, 24, 32],
        "itemsize": 48,
    }
)


def to_pcl_xyzinormal(cloud):
    out = np.zeros(cloud.shape[0], dtype=PCL_XYZINORMAL)
    out["x"], out["y"], out["z"] = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    out["intensity"] = 1.0
    out["normal_x"] = cloud[:, 3]
    return out


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def rot_xyz(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def make_pose(t, rpy):
    T = np.eye(4)
    T[:3, :3] = rot_xyz(*rpy)
    T[:3, 3] = t
    return T


def random_motion(rng, scale=1.0, rot_scale=1.0):
    """Relative motion of SURVEY.md §8(d): t_xy ~ U(-.5,.5) m, t_z ~ U(-.05,.05),
    yaw ~ U(-3,3) deg, roll/pitch ~ U(-.5,.5) deg."""
    t = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.05, 0.05)]) * scale
    rpy = np.deg2rad([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-3, 3)]) * scale * rot_scale
    return make_pose(t, rpy)


class Scene:
    """Static world: ground plane z = -1.5, vertical wall rectangles, blobs."""

    def __init__(self, seed, extent=(0.0, 110.0, -90.0, 90.0), n_walls=8, n_blobs=24):
        rng = _rng(seed)
        x0, x1, y0, y1 = extent
        self.extent = extent
        self.walls = []  # (origin xyz, unit direction xy, length, height)
        for _ in range(n_walls):
            o = np.array([rng.uniform(x0 + 8, x1), rng.uniform(y0, y1), GROUND_Z])
            ang = rng.uniform(0, np.pi)
            self.walls.append((o, np.array([np.cos(ang), np.sin(ang), 0.0]), rng.uniform(15, 50), rng.uniform(3, 8)))
        self.blobs = np.stack(
            [rng.uniform(x0 + 5, x1 * 0.8, n_blobs), rng.uniform(y0 * 0.6, y1 * 0.6, n_blobs), rng.uniform(-1.2, 1.0, n_blobs)], axis=1
        )


def _sample_range(rng, n, lo=R_MIN, hi=R_MAX):
    # density ~ r^0.9  ->  CDF ~ r^1.9
    u = rng.random(n)
    return (u * (hi**1.9 - lo**1.9) + lo**1.9) ** (1 / 1.9)


def _in_fov(p):
    r = np.linalg.norm(p, axis=1)
    az = np.arctan2(p[:, 1], p[:, 0])
    el = np.arcsin(np.clip(p[:, 2] / np.maximum(r, 1e-9), -1, 1))
    return (r >= R_MIN) & (r <= R_MAX) & (np.abs(az) <= FOV_AZ) & (np.abs(el) <= FOV_EL)


def _to_sensor(T_ws, pw):
    R, t = T_ws[:3, :3], T_ws[:3, 3]
    return (pw - t) @ R  # R^T (pw - t)


def _draw(sampler, n, rng):
    """Rejection-sample n sensor-frame points with `sampler(m) -> [m,3]`, FOV-filtered."""
    out = []
    have = 0
    tries = 0
    while have < n and tries < 64:
        p = sampler(max(256, int((n - have) * 3)))
        p = p[_in_fov(p)]
        out.append(p)
        have += p.shape[0]
        tries += 1
    p = np.concatenate(out, axis=0) if out else np.zeros((0, 3))
    if p.shape[0] < n:  # object never in view: fall back to clutter-like samples
        extra = n - p.shape[0]
        r = _sample_range(rng, extra)
        az = rng.uniform(-FOV_AZ, FOV_AZ, extra)
        el = rng.uniform(-FOV_EL, FOV_EL, extra)
        q = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], axis=1)
        p = np.concatenate([p, q], axis=0)
    return p[:n]


def radar_scan(scene, T_ws, n, seed, noise=True):
    """One n-point scan of `scene` from sensor pose T_ws (sensor -> world).
    Returns float32 [n,4] in the SENSOR frame."""
    rng = _rng(seed)
    n_ground = int(round(0.5 * n))
    n_wall = int(round(0.3 * n))
    n_blob = int(round(0.1 * n))
    n_clutter = n - n_ground - n_wall - n_blob
    R, t = T_ws[:3, :3], T_ws[:3, 3]
    parts, obj = [], []

    def ground(m):
        r = _sample_range(rng, m, 6.0, R_MAX)
        az = rng.uniform(-FOV_AZ, FOV_AZ, m)
        x, y = r * np.cos(az), r * np.sin(az)
        z = (GROUND_Z - t[2] - R[2, 0] * x - R[2, 1] * y) / R[2, 2]
        return np.stack([x, y, z], axis=1)

    parts.append(_draw(ground, n_ground, rng))
    obj.append(np.full(n_ground, 1, dtype=np.int64))

    nw = len(scene.walls)
    per = [n_wall // nw + (1 if i < n_wall % nw else 0) for i in range(nw)]
    for wi, (o, d, length, height) in enumerate(scene.walls):
        if per[wi] == 0:
            continue

        def wall(m, o=o, d=d, length=length, height=height):
            pw = o + np.outer(rng.uniform(0, length, m), d) + np.outer(rng.uniform(0, height, m), [0, 0, 1.0])
            return _to_sensor(T_ws, pw)

        parts.append(_draw(wall, per[wi], rng))
        obj.append(np.full(per[wi], 2 + wi, dtype=np.int64))

    nb = scene.blobs.shape[0]
    which = rng.integers(0, nb, n_blob)
    pw = scene.blobs[which] + rng.normal(0, 0.3, (n_blob, 3))
    pb = _to_sensor(T_ws, pw)
    bad = ~_in_fov(pb)
    if bad.any():  # blobs out of view: pull them to a visible clutter position, keep their id
        m = int(bad.sum())
        r = _sample_range(rng, m)
        az = rng.uniform(-FOV_AZ, FOV_AZ, m)
        el = rng.uniform(-FOV_EL, FOV_EL, m)
        pb[bad] = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], axis=1)
    parts.append(pb)
    obj.append(2 + nw + which)

    r = _sample_range(rng, n_clutter)
    az = rng.uniform(-FOV_AZ, FOV_AZ, n_clutter)
    el = rng.uniform(-FOV_EL, FOV_EL, n_clutter)
    parts.append(np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], axis=1))
    obj.append(np.zeros(n_clutter, dtype=np.int64))

    p = np.concatenate(parts, axis=0)
    obj = np.concatenate(obj)

    if noise:
        rr = np.linalg.norm(p, axis=1)
        az = np.arctan2(p[:, 1], p[:, 0])
        el = np.arcsin(np.clip(p[:, 2] / rr, -1, 1))
        rr = rr + rng.normal(0, 1, n) * rr * 0.86 / 400
        az = az + rng.normal(0, np.deg2rad(0.5), n)
        el = el + rng.normal(0, np.deg2rad(1.0), n)
        p = np.stack([rr * np.cos(el) * np.cos(az), rr * np.cos(el) * np.sin(az), rr * np.sin(el)], axis=1)

    # label = object id ranked by centroid distance, 0 for clutter
    label = np.zeros(n, dtype=np.float32)
    ids = np.unique(obj[obj > 0])
    if ids.size:
        cd = np.array([np.linalg.norm(p[obj == i].mean(axis=0)) for i in ids])
        for rank, i in enumerate(ids[np.argsort(cd, kind="stable")]):
            label[obj == i] = float(rank + 1)

    perm = rng.permutation(n)
    cloud = np.concatenate([p[perm].astype(np.float32), label[perm, None]], axis=1).astype(np.float32)
    return _dedup(cloud, rng)


def _dedup(cloud, rng):
    """No duplicate xyz (exact float ties make the kNN tie rule visible)."""
    for _ in range(8):
        _, first = np.unique(cloud[:, :3], axis=0, return_index=True)
        if first.size == cloud.shape[0]:
            return cloud
        dup = np.ones(cloud.shape[0], dtype=bool)
        dup[first] = False
        cloud[dup, :3] += rng.normal(0, 1e-3, (int(dup.sum()), 3)).astype(np.float32)
    return cloud


def scan_pair(seed, n=1000):
    """Config C1: two n-point scans of one scene. Returns (source, target, T_gt)
    with T_gt mapping source-frame points into the target frame."""
    rng = _rng(seed)
    scene = Scene(seed * 7919 + 1)
    A = np.eye(4)
    delta = random_motion(rng)
    B = A @ delta
    target = radar_scan(scene, A, n, seed * 3 + 1)
    m = int(np.ceil(n / 0.8))
    src_full = radar_scan(scene, B, m, seed * 3 + 2)
    keep = np.sort(_rng(seed * 3 + 3).permutation(m)[:n])  # 20 % dropout
    return src_full[keep].copy(), target, delta


def submap_pair(seed, n_source=2000, n_frames=30, n_per_frame=2000, path_len=15.0):
    """Config C2: an n_source scan against a keyframe submap (union of n_frames
    scans along a path, expressed in the frame of the LAST pose)."""
    rng = _rng(seed)
    scene = Scene(seed * 7919 + 1, extent=(0.0, 110.0 + path_len, -90.0, 90.0))
    poses = []
    for f in range(n_frames):
        s = f / max(1, n_frames - 1)
        yaw = np.deg2rad(4.0) * np.sin(2.0 * s)
        poses.append(make_pose([path_len * s, 0.4 * np.sin(3.0 * s), 0.0], [0.0, 0.0, yaw]))
    ref = poses[-1]
    ref_inv = np.linalg.inv(ref)
    chunks = []
    for f, P in enumerate(poses):
        sc = radar_scan(scene, P, n_per_frame, seed * 1000 + f)
        M = ref_inv @ P
        xyz = sc[:, :3].astype(np.float64) @ M[:3, :3].T + M[:3, 3]
        chunks.append(np.concatenate([xyz.astype(np.float32), sc[:, 3:4]], axis=1))
    target = _dedup(np.concatenate(chunks, axis=0).astype(np.float32), rng)
    delta = random_motion(rng)
    B = ref @ delta
    m = int(np.ceil(n_source / 0.8))
    src_full = radar_scan(scene, B, m, seed * 1000 + 999)
    keep = np.sort(_rng(seed * 3 + 3).permutation(m)[:n_source])
    return src_full[keep].copy(), target, delta


def tiled_cloud_pair(seed, n, base_n=78125, pitch=125.0):
    """Config C4: two n-point clouds (n up to 20 M) — a submap-density base cloud
    tiled over a square area, each tile with its own centimetre-level jitter so no
    two points coincide; the source is the same world re-observed with independent
    5 cm noise and moved by a small rigid motion. Returns (source, target, T_gt)."""
    rng = _rng(seed)
    base_n = min(base_n, n)
    _, base, _ = submap_pair(seed, n_source=64, n_frames=max(2, base_n // 2000), n_per_frame=min(2000, base_n))
    if base.shape[0] < base_n:
        reps = int(np.ceil(base_n / base.shape[0]))
        base = np.concatenate([base] * reps, axis=0)
    base = base[:base_n]
    tiles = int(np.ceil(n / base_n))
    side = int(np.ceil(np.sqrt(tiles)))
    tgt = np.empty((tiles * base_n, 4), dtype=np.float32)
    src = np.empty((tiles * base_n, 4), dtype=np.float32)
    # keep the rotation small enough that the far edge of the tiled area moves by < ~0.5 m
    # (a scan-sized yaw would displace points a kilometre away by tens of metres)
    delta = random_motion(rng, rot_scale=min(1.0, 10.0 / (side * pitch)))
    Dinv = np.linalg.inv(delta)
    for ti in range(tiles):
        off = np.array([(ti % side) * pitch, (ti // side) * pitch, 0.0])
        w = base[:, :3].astype(np.float64) + off
        lab = base[:, 3] + np.float32(ti * 64)
        a = w + rng.normal(0, 0.01, w.shape)
        b = (w + rng.normal(0, 0.05, w.shape)) @ Dinv[:3, :3].T + Dinv[:3, 3]
        sl = slice(ti * base_n, (ti + 1) * base_n)
        tgt[sl, :3], tgt[sl, 3] = a, lab
        src[sl, :3], src[sl, 3] = b, lab
    return src[:n].copy(), tgt[:n].copy(), delta
