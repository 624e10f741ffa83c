"""Deterministic synthetic 4D-radar-shaped clouds for tests and bench.py.

Follows SURVEY.md §8(d). A scene is a fixed set of radar scatterers in the world
(50 % on the ground plane z = -1.5 m, 30 % on vertical wall rectangles, 10 % in
0.3 m Gaussian blobs — poles, vehicles —, 10 % uniform clutter), laid out so that a
sensor sees them with a 120 deg x 30 deg field of view, ranges 1-100 m and a range
density ~ r^0.9. A scan observes the scatterers in view from a pose with random
dropout and per-point sensor noise drawn from the reference's own noise model
(sigma_range = r*0.86/400, sigma_az = 0.5 deg, sigma_el = 1.0 deg, reference
fast_apdgicp.hpp:116-118); consecutive scans therefore see mostly the same physical
reflectors, as a radar does. Each point carries a cluster label that mimics the
DBSCAN rank the reference's preprocessing writes into ``normal_x``
(4DRadarSLAM/apps/preprocessing_nodelet_ntu.cpp:558-567): object id ranked by
centroid distance, 0 for clutter.

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
        "offsets": [0, 4, 8, 32, 16, 20, 24, 36],  # PCL_ADD_POINT4D | PCL_ADD_NORMAL4D | {intensity, curvature, pad, pad}
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


def _to_world(T_ws, ps):
    return ps @ T_ws[:3, :3].T + T_ws[:3, 3]


def _polar_box(rng, m, lo=R_MIN):
    r = _sample_range(rng, m, lo)
    az = rng.uniform(-FOV_AZ, FOV_AZ, m)
    el = rng.uniform(-FOV_EL, FOV_EL, m)
    return np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], axis=1)


def make_walls(rng, extent, n_walls):
    x0, x1, y0, y1 = extent
    walls = []
    for _ in range(n_walls):
        o = np.array([rng.uniform(x0 + 8, x1), rng.uniform(y0, y1), GROUND_Z])
        ang = rng.uniform(0, np.pi)
        walls.append((o, np.array([np.cos(ang), np.sin(ang), 0.0]), rng.uniform(15, 50), rng.uniform(3, 8)))
    return walls


def scatterers_seen_from(rng, T_ws, m, walls, blobs, obj_base=0):
    """m world scatterers laid out as a sensor at pose T_ws sees the scene: (points [m,3], object id [m]).
    Object ids: 1 ground, 2.. walls, then blobs, 0 clutter (offset by obj_base for all but clutter)."""
    n_ground, n_wall, n_blob = int(round(0.5 * m)), int(round(0.3 * m)), int(round(0.1 * m))
    n_clutter = m - n_ground - n_wall - n_blob
    R, t = T_ws[:3, :3], T_ws[:3, 3]
    pts, obj = [], []
    # ground: horizontal range / azimuth in the sensor frame, z from the plane equation
    r = _sample_range(rng, n_ground, 6.0)
    az = rng.uniform(-FOV_AZ, FOV_AZ, n_ground)
    x, y = r * np.cos(az), r * np.sin(az)
    z = (GROUND_Z - t[2] - R[2, 0] * x - R[2, 1] * y) / R[2, 2]
    pts.append(_to_world(T_ws, np.stack([x, y, z], axis=1)))
    obj.append(np.full(n_ground, obj_base + 1))
    # walls: uniform on the rectangles, kept if in view (rejection), spread over the walls
    nw = max(1, len(walls))
    got = 0
    for wi, (o, d, length, height) in enumerate(walls):
        want = n_wall // nw + (1 if wi < n_wall % nw else 0)
        if want == 0:
            continue
        pw = o + np.outer(rng.uniform(0, length, 4 * want), d) + np.outer(rng.uniform(0, height, 4 * want), [0, 0, 1.0])
        pw = pw[_in_fov(_to_sensor(T_ws, pw))][:want]
        pts.append(pw)
        obj.append(np.full(pw.shape[0], obj_base + 2 + wi))
        got += pw.shape[0]
    # blobs
    if len(blobs):
        which = rng.integers(0, len(blobs), n_blob)
        pw = blobs[which] + rng.normal(0, 0.3, (n_blob, 3))
        keep = _in_fov(_to_sensor(T_ws, pw))
        pts.append(pw[keep])
        obj.append(obj_base + 2 + len(walls) + which[keep])
        got += int(keep.sum())
    else:
        n_blob = 0
    # clutter (also tops up what the walls / blobs could not place in view)
    n_fill = n_clutter + (n_wall + n_blob - got)
    pts.append(_to_world(T_ws, _polar_box(rng, n_fill)))
    obj.append(np.zeros(n_fill, dtype=np.int64))
    return np.concatenate(pts, axis=0), np.concatenate(obj).astype(np.int64)


class Scene:
    """A fixed set of world scatterers with their object ids."""

    def __init__(self, points, obj):
        self.points = points
        self.obj = obj

    @staticmethod
    def around(seed, poses, n_per_pose, extent=(0.0, 110.0, -90.0, 90.0), n_walls=8, n_blobs=24):
        """scatterers laid out as seen from each of `poses` (n_per_pose each), one shared wall / blob geometry"""
        rng = _rng(seed)
        walls = make_walls(rng, extent, n_walls)
        x0, x1, y0, y1 = extent
        blobs = np.stack([rng.uniform(x0 + 5, x1 * 0.8, n_blobs), rng.uniform(y0 * 0.6, y1 * 0.6, n_blobs), rng.uniform(-1.2, 1.0, n_blobs)], axis=1)
        P, O = [], []
        for T in poses:
            p, o = scatterers_seen_from(rng, T, n_per_pose, walls, blobs)
            P.append(p)
            O.append(o)
        return Scene(np.concatenate(P, axis=0), np.concatenate(O))


def radar_scan(scene, T_ws, n, seed, noise=True):
    """One scan of `scene` from sensor pose T_ws (sensor -> world): the scatterers in view, randomly
    thinned to at most n, with sensor noise (noise: True = the spec'd sigmas, a float scales them, False / 0 = none).
    Returns float32 [<= n, 4] in the SENSOR frame."""
    rng = _rng(seed)
    ps = _to_sensor(T_ws, scene.points)
    vis = np.flatnonzero(_in_fov(ps))
    if vis.size > n:
        vis = np.sort(rng.choice(vis, n, replace=False))
    p, obj = ps[vis], scene.obj[vis]
    m = p.shape[0]
    if noise:
        rr = np.linalg.norm(p, axis=1)
        az = np.arctan2(p[:, 1], p[:, 0])
        el = np.arcsin(np.clip(p[:, 2] / rr, -1, 1))
        ns = 1.0 if noise is True else float(noise)
        rr = rr + rng.normal(0, 1, m) * rr * 0.86 / 400 * ns
        az = az + rng.normal(0, np.deg2rad(0.5), m) * ns
        el = el + rng.normal(0, np.deg2rad(1.0), m) * ns
        p = np.stack([rr * np.cos(el) * np.cos(az), rr * np.cos(el) * np.sin(az), rr * np.sin(el)], axis=1)
    # label = object id ranked by centroid distance, 0 for clutter
    label = np.zeros(m, dtype=np.float32)
    ids = np.unique(obj[obj > 0])
    if ids.size:
        cd = np.array([np.linalg.norm(p[obj == i].mean(axis=0)) for i in ids])
        for rank, i in enumerate(ids[np.argsort(cd, kind="stable")]):
            label[obj == i] = float(rank + 1)
    perm = rng.permutation(m)
    cloud = np.concatenate([p[perm].astype(np.float32), label[perm, None]], axis=1).astype(np.float32)
    return _dedup(cloud, rng)


def _duplicate_rows(xyz):
    """mask of the rows whose xyz equals that of an EARLIER row (what np.unique(axis=0, return_index=True) leaves out),
    found through a 1-D sort of a 64-bit hash of the coordinate bits — the row-wise unique costs 30x more"""
    bits = np.ascontiguousarray(xyz + np.float32(0.0)).view(np.uint32).astype(np.uint64)  # (+0.0: -0.0 and 0.0 are one value)
    h = (bits[:, 0] * np.uint64(0x9E3779B97F4A7C15)) ^ (bits[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F)) ^ (bits[:, 2] * np.uint64(0x165667B19E3779F9))
    order = np.argsort(h, kind="stable")
    same = h[order][1:] == h[order][:-1]
    dup = np.zeros(xyz.shape[0], dtype=bool)
    if not same.any():
        return dup
    # rows that share a hash with a neighbour: settle those few exactly
    cand = np.unique(np.concatenate([order[1:][same], order[:-1][same]]))
    _, first = np.unique(xyz[cand], axis=0, return_index=True)
    dup[cand] = True
    dup[cand[first]] = False
    return dup


def _dedup(cloud, rng):
    """No duplicate xyz (exact float ties make the kNN tie rule visible)."""
    for _ in range(8):
        dup = _duplicate_rows(cloud[:, :3])
        if not dup.any():
            return cloud
        cloud[dup, :3] += rng.normal(0, 1e-3, (int(dup.sum()), 3)).astype(np.float32)
    return cloud


def scan_pair(seed, n=1000):
    """Config C1: two n-point scans of one scene (same reflectors, independent noise, 20 % dropout).
    Returns (source, target, T_gt) with T_gt mapping source-frame points into the target frame."""
    rng = _rng(seed)
    A = np.eye(4)
    delta = random_motion(rng)
    B = A @ delta
    scene = Scene.around(seed * 7919 + 1, [A], int(np.ceil(n / 0.8 * 1.08)))
    target = radar_scan(scene, A, n, seed * 3 + 1)
    source = radar_scan(scene, B, n, seed * 3 + 2)
    return source, target, delta


def submap_pair(seed, n_source=2000, n_frames=30, n_per_frame=2000, path_len=15.0):
    """Config C2: an n_source scan against a keyframe submap (union of n_frames scans taken along a
    path, expressed in the frame of the LAST pose)."""
    rng = _rng(seed)
    poses = []
    for f in range(n_frames):
        s = f / max(1, n_frames - 1)
        yaw = np.deg2rad(4.0) * np.sin(2.0 * s)
        poses.append(make_pose([path_len * s, 0.4 * np.sin(3.0 * s), 0.0], [0.0, 0.0, yaw]))
    ref = poses[-1]
    ref_inv = np.linalg.inv(ref)
    anchor = poses[:: max(1, n_frames // 4)] + [ref]
    per = int(np.ceil(max(n_per_frame, n_source) / 0.8 * 1.1 / len(anchor) * 1.6))
    scene = Scene.around(seed * 7919 + 1, anchor, per, extent=(0.0, 110.0 + path_len, -90.0, 90.0))
    chunks = []
    for f, P in enumerate(poses):
        sc = radar_scan(scene, P, n_per_frame, seed * 1000 + f)
        M = ref_inv @ P
        xyz = sc[:, :3].astype(np.float64) @ M[:3, :3].T + M[:3, 3]
        chunks.append(np.concatenate([xyz.astype(np.float32), sc[:, 3:4]], axis=1))
    target = _dedup(np.concatenate(chunks, axis=0).astype(np.float32), rng)
    delta = random_motion(rng)
    B = ref @ delta
    source = radar_scan(scene, B, n_source, seed * 1000 + 999)
    return source, target, delta


def tiled_cloud_pair(seed, n, base_n=78125, pitch=125.0):
    """Config C4: two n-point clouds (n up to 20 M) — a submap-density base cloud tiled over a square
    area, each tile with its own centimetre-level jitter so no two points coincide; the source is the
    same world re-observed with independent 5 cm noise and moved by a small rigid motion.
    Returns (source, target, T_gt)."""
    rng = _rng(seed)
    base_n = min(base_n, n)
    _, base, _ = submap_pair(seed, n_source=64, n_frames=max(2, base_n // 2000), n_per_frame=min(2000, base_n))
    if base.shape[0] < base_n:
        reps = int(np.ceil(base_n / base.shape[0]))
        base = np.concatenate([base + np.float32(0.013 * r) for r in range(reps)], axis=0)
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


# --------------------------------------------------------------------------
# Config C5: a long drive for the odometry replay
# --------------------------------------------------------------------------
class CorridorScene:
    """An unbounded street-like world for long trajectories: per 100 m chunk a fixed set of
    scatterers (ground, walls on both sides, pole/vehicle blobs, clutter) generated from the chunk
    index, so any pose sees a deterministic neighbourhood."""

    def __init__(self, seed, per_chunk=1800):
        self.seed = seed
        self.per_chunk = per_chunk
        self._cache = {}

    def _chunk(self, ci):
        if ci not in self._cache:
            rng = _rng(self.seed * 100003 + (ci + 50000))
            x0 = ci * 100.0
            walls = []
            for j in range(6):
                side = 1.0 if j % 2 == 0 else -1.0
                o = np.array([x0 + rng.uniform(0, 100), side * rng.uniform(8, 30), GROUND_Z])
                ang = rng.normal(0.0, 0.25)
                walls.append((o, np.array([np.cos(ang), np.sin(ang), 0.0]), rng.uniform(15, 45), rng.uniform(3, 8)))
            blobs = np.stack([x0 + rng.uniform(0, 100, 8), rng.uniform(-25, 25, 8), rng.uniform(-1.2, 1.0, 8)], axis=1)
            m = self.per_chunk
            ng, nw, nb = int(0.5 * m), int(0.3 * m), int(0.1 * m)
            P = [np.stack([x0 + rng.uniform(0, 100, ng), rng.uniform(-60, 60, ng), np.full(ng, GROUND_Z)], axis=1)]
            O = [np.full(ng, ci * 100 + 1)]
            for wi, (o, d, length, height) in enumerate(walls):
                k = nw // len(walls)
                P.append(o + np.outer(rng.uniform(0, length, k), d) + np.outer(rng.uniform(0, height, k), [0, 0, 1.0]))
                O.append(np.full(k, ci * 100 + 2 + wi))
            which = rng.integers(0, len(blobs), nb)
            P.append(blobs[which] + rng.normal(0, 0.3, (nb, 3)))
            O.append(ci * 100 + 10 + which)
            nc = m - sum(p.shape[0] for p in P)
            P.append(np.stack([x0 + rng.uniform(0, 100, nc), rng.uniform(-60, 60, nc), rng.uniform(-1.5, 12, nc)], axis=1))
            O.append(np.zeros(nc, dtype=np.int64))
            self._cache[ci] = (np.concatenate(P, axis=0), np.concatenate(O).astype(np.int64))
        return self._cache[ci]

    def local(self, T_ws):
        """the scatterers of the chunks within ~100 m ahead of / behind the pose, as a Scene"""
        c0 = int(np.floor(T_ws[0, 3] / 100.0))
        P, O = zip(*[self._chunk(ci) for ci in range(c0 - 1, c0 + 3)])
        return Scene(np.concatenate(P, axis=0), np.concatenate(O))


def drive_trajectory(n_frames, seed, hz=10.0):
    """Smooth 6-DoF trajectory, mostly forward, speed 2-6 m/s (the reference caps the ego velocity at
    12-15 m/s, launch/ntu_loop2.launch:84, but bridges fast motion with its Doppler ego-velocity guess,
    which is outside this path: the replay runs with guess = previous transform only), gentle yaw /
    roll / pitch oscillation. Returns poses [n,4,4]."""
    rng = _rng(seed)
    ph = rng.uniform(0, 2 * np.pi, 5)
    poses = np.zeros((n_frames, 4, 4))
    x = y = 0.0
    for i in range(n_frames):
        t = i / hz
        v = 4.0 + 2.0 * np.sin(0.05 * t + ph[0])
        yaw = 0.25 * np.sin(0.03 * t + ph[1])
        x += v * np.cos(yaw) / hz
        y += v * np.sin(yaw) / hz
        z = 0.05 * np.sin(0.2 * t + ph[2])
        poses[i] = make_pose([x, y, z], [0.01 * np.sin(0.3 * t + ph[3]), 0.01 * np.sin(0.25 * t + ph[4]), yaw])
    return poses


def drive_frames(seed, n_frames, n_points=1000, noise=True):
    """generator of (frame index, cloud [<= n_points,4] float32 in the sensor frame, ground-truth pose).
    noise: scale of the sensor noise (True = 1 = the spec'd sigmas). With the full noise a pair of 1000-point scans
    registers a 0.8 m step to +-15 % with a +12 % forward bias (measured with the CPU restatement; without noise: 1.006
    +- 0.03) — open-loop scan-to-scan odometry is then a biased random walk, which the reference closes with its Doppler
    ego-velocity guess and the IMU, both outside this path. The replay benchmark therefore runs at a reduced scale."""
    scene = CorridorScene(seed)
    poses = drive_trajectory(n_frames, seed)
    for i in range(n_frames):
        yield i, radar_scan(scene.local(poses[i]), poses[i], n_points, seed * 1000003 + i, noise=noise), poses[i]
