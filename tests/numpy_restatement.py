"""Independent NumPy/SciPy restatement of the FastAPDGICP math, written from the
reference source (fast_apdgicp_impl.hpp:160-411, lsq_registration_impl.hpp) without
looking at oracle/: np.linalg.svd instead of the oracle's Jacobi solver,
np.linalg.inv instead of adjugates, scipy cKDTree (fp64) instead of the fp32
kd-tree. It pins the oracle's arithmetic; it is TEST INFRASTRUCTURE.
"""
import numpy as np
from scipy.spatial import cKDTree


def knn_sets(cloud_xyz, k):
    tree = cKDTree(cloud_xyz.astype(np.float64))
    d, idx = tree.query(cloud_xyz.astype(np.float64), k=k + 1)
    return d, idx


def covariances(cloud, nbr_idx, reg="PLANE"):
    """fast_apdgicp_impl.hpp:366-405. nbr_idx [n,k] neighbour indices (incl. self)."""
    P = cloud[:, :3].astype(np.float64)[nbr_idx]            # [n,k,3]
    k = nbr_idx.shape[1]
    C = np.einsum("nki,nkj->nij", P - P.mean(axis=1, keepdims=True), P - P.mean(axis=1, keepdims=True)) / k
    if reg == "NONE":
        return C
    if reg == "FROBENIUS":
        Ci = np.linalg.inv(C + 1e-3 * np.eye(3))
        nrm = np.sqrt((Ci ** 2).sum(axis=(1, 2)))[:, None, None]
        return np.linalg.inv(Ci / nrm)
    U, S, Vt = np.linalg.svd(C)
    if reg == "PLANE":
        vals = np.tile(np.array([1.0, 1.0, 1e-3]), (C.shape[0], 1))
    elif reg == "MIN_EIG":
        vals = np.maximum(S, 1e-3)
    elif reg == "NORMALIZED_MIN_EIG":
        vals = np.maximum(S / S[:, :1], 1e-3)
    else:
        raise ValueError(reg)
    return np.einsum("nij,nj,nkj->nik", U, vals, Vt.transpose(0, 2, 1))


def noise_cov(pt32, dist_var, az_var, el_var):
    """fast_apdgicp_impl.hpp:193-210 for float32 points pt32 [n,3]."""
    p = pt32.astype(np.float64)
    dist = np.sqrt((p ** 2).sum(axis=1))
    s = np.stack([dist * dist_var / 400, dist * np.sin(az_var / 180 * np.pi), dist * np.sin(el_var / 180 * np.pi)], axis=1)
    rho = np.sqrt(pt32[:, 0] * pt32[:, 0] + pt32[:, 1] * pt32[:, 1]).astype(np.float32)
    el = np.arctan2(rho.astype(np.float64), p[:, 2]).astype(np.float32).astype(np.float64)
    az = np.arctan2(p[:, 1], p[:, 0]).astype(np.float32).astype(np.float64)
    ce, se, ca, sa = np.cos(el), np.sin(el), np.cos(az), np.sin(az)
    n = p.shape[0]
    Ry = np.zeros((n, 3, 3)); Rz = np.zeros((n, 3, 3))
    Ry[:, 0, 0] = ce; Ry[:, 0, 2] = se; Ry[:, 1, 1] = 1; Ry[:, 2, 0] = -se; Ry[:, 2, 2] = ce
    Rz[:, 0, 0] = ca; Rz[:, 0, 1] = -sa; Rz[:, 1, 0] = sa; Rz[:, 1, 1] = ca; Rz[:, 2, 2] = 1
    A = (Rz @ Ry) * s[:, None, :]
    return A @ A.transpose(0, 2, 1)


def transform_f32(T, xyz32):
    Tf = T.astype(np.float32)
    x, y, z = xyz32[:, 0], xyz32[:, 1], xyz32[:, 2]
    out = np.empty_like(xyz32, dtype=np.float32)
    for r in range(3):
        out[:, r] = ((Tf[r, 0] * x + Tf[r, 1] * y).astype(np.float32) + Tf[r, 2] * z).astype(np.float32) + Tf[r, 3]
    return out


def mahalanobis(T, src, tgt, cov_src, cov_tgt, corr, dist_var=0.86, az_var=0.5, el_var=1.0, gicp=False):
    """fast_apdgicp_impl.hpp:193-218 given the correspondences; gicp: fast_gicp_impl.hpp:157-161 (no noise term)."""
    pt = transform_f32(T, src[:, :3])
    cov_r = np.zeros((src.shape[0], 3, 3)) if gicp else noise_cov(pt, dist_var, az_var, el_var)
    R = T[:3, :3]
    valid = corr >= 0
    M = np.zeros((src.shape[0], 3, 3))
    cB = cov_tgt[np.where(valid, corr, 0)]
    RCR = (cB + cov_r) + R @ (cov_src + cov_r) @ R.T
    M[valid] = np.linalg.inv(RCR[valid])
    return M


def linearize(T, src, tgt, cov_src, corr, M, gicp=False):
    """fast_apdgicp_impl.hpp:247-304 given correspondences and Mahalanobis; gicp: fast_gicp_impl.hpp:168-237 (unit weights)."""
    n = src.shape[0]
    valid = corr >= 0
    a = src[:, :3].astype(np.float64)
    b = tgt[np.where(valid, corr, 0), :3].astype(np.float64)
    tA = a @ T[:3, :3].T + T[:3, 3]
    e = b - tA
    sv = np.linalg.svd(cov_src, compute_uv=False)
    geo = sv[:, 2] / sv[:, 0]
    cl = np.where(tgt[np.where(valid, corr, 0), 3] == src[:, 3], 1.0 / n, 0.0)
    q = np.einsum("ni,nij,nj->n", e, M, e)
    err = (q if gicp else (1.0 + geo + cl) * q)[valid].sum()
    J = np.zeros((n, 3, 6))
    J[:, 0, 1] = -tA[:, 2]; J[:, 0, 2] = tA[:, 1]
    J[:, 1, 0] = tA[:, 2];  J[:, 1, 2] = -tA[:, 0]
    J[:, 2, 0] = -tA[:, 1]; J[:, 2, 1] = tA[:, 0]
    J[:, 0, 3] = J[:, 1, 4] = J[:, 2, 5] = -1.0
    H = np.einsum("nia,nij,njb->nab", J, M, J)[valid].sum(axis=0)
    bb = np.einsum("nia,nij,nj->na", J, M, e)[valid].sum(axis=0)
    return err, H, bb


def so3_exp(w):
    th2 = w @ w
    if th2 < 1e-10:
        im = 0.5 - th2 / 48 + th2 * th2 / 3840
        re = 1 - th2 / 8 + th2 * th2 / 384
    else:
        th = np.sqrt(th2)
        im = np.sin(th / 2) / th
        re = np.cos(th / 2)
    qw, qx, qy, qz = re, im * w[0], im * w[1], im * w[2]
    return np.array([
        [1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
        [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
        [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)],
    ])


# ---- FastVGICP (fast_vgicp_impl.hpp, fast_vgicp_voxel.hpp), written from the reference source with a Python dict as
# the voxel map and np.linalg.inv on the 4x4 matrices the reference inverts ----
def vgicp_offsets(method):
    if method == "DIRECT1":
        return [(0, 0, 0)]
    if method == "DIRECT7":
        return [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    return [(i - 1, j - 1, k - 1) for i in range(3) for j in range(3) for k in range(3)]


def vgicp_voxelmap(tgt, cov_tgt, resolution, multiplicative=False):
    """create_voxelmap (fast_vgicp_voxel.hpp:131-158): {coord: [n, mean4, cov4x4]}"""
    vox = {}
    P = np.concatenate([tgt[:, :3].astype(np.float64), np.ones((tgt.shape[0], 1))], axis=1)
    coords = np.floor(P[:, :3] / resolution - 0.5).astype(np.int64)
    for i in range(tgt.shape[0]):
        C4 = np.zeros((4, 4)); C4[:3, :3] = cov_tgt[i]
        v = vox.setdefault(tuple(coords[i]), [0, np.zeros(4), np.zeros((4, 4))])
        v[0] += 1
        if multiplicative:
            Ci = C4.copy(); Ci[3, 3] = 1.0
            Ci = np.linalg.inv(Ci)
            v[2] += Ci
            v[1] += Ci @ P[i]
        else:
            v[1] += P[i]
            v[2] += C4
    for v in vox.values():
        if multiplicative:
            v[2][3, 3] = 1.0
            v[1][3] = 1.0
            v[2] = np.linalg.inv(v[2])
            v[1] = v[2] @ v[1]
        else:
            v[1] = v[1] / v[0]
            v[2] = v[2] / v[0]
    return vox


def vgicp_linearize(T, src, cov_src, vox, resolution, method="DIRECT1", T_corr=None):
    """update_correspondences at T_corr (default T) + the sums at T (fast_vgicp_impl.hpp:74-181).
    Returns err, H, b, and the correspondence list [(i, coord)]."""
    Tc = T if T_corr is None else T_corr
    A = np.concatenate([src[:, :3].astype(np.float64), np.ones((src.shape[0], 1))], axis=1)
    tc = A @ Tc.T
    base = np.floor(tc[:, :3] / resolution - 0.5).astype(np.int64)
    err, H, b, corr = 0.0, np.zeros((6, 6)), np.zeros(6), []
    for i in range(src.shape[0]):
        for off in vgicp_offsets(method):
            c = tuple(base[i] + np.array(off))
            if c not in vox:
                continue
            n, mean, cov = vox[c]
            CA = np.zeros((4, 4)); CA[:3, :3] = cov_src[i]
            RCR = cov + Tc @ CA @ Tc.T
            RCR[3, 3] = 1.0
            M = np.linalg.inv(RCR)
            M[3, 3] = 0.0
            tA = T @ A[i]
            e = mean - tA
            w = np.sqrt(n)
            err += w * e @ M @ e
            J = np.zeros((4, 6))
            J[:3, :3] = np.array([[0, -tA[2], tA[1]], [tA[2], 0, -tA[0]], [-tA[1], tA[0], 0]])
            J[:3, 3:] = -np.eye(3)
            H += w * J.T @ M @ J
            b += w * J.T @ M @ e
            corr.append((i, c))
    return err, H, b, corr


# ---- DBSCANKdtreeCluster::extract + the preprocessing nodelet's cluster labels (4DRadarSLAM/include/dbscan/DBSCAN_simple.h:27-104,
# apps/preprocessing_nodelet_ntu.cpp:520-567): a literal transcription, brute-force neighbour sets, small clouds only ----
def dbscan_radius_lists(cloud, eps):
    x = cloud[:, :3].astype(np.float32)
    s = (x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1]) + x[:, 2] * x[:, 2]        # float expression (:35-37)
    nf = np.sqrt(s)                                                        # std::sqrt(float)
    rad_seed = np.abs(nf.astype(np.float64) - 1.0) / 50.0 + eps            # double norm; std::abs(norm - 1)/50 + eps_ (:38)
    rad_exp = ((nf - np.float32(1)) / np.float32(100)).astype(np.float64) + eps  # (float - 1)/100 + eps_ (:60-62)
    out = []
    for rad in (rad_seed, rad_exp):
        r2 = (rad * rad).astype(np.float32)                                # [ext] pcl KdTreeFLANN: static_cast<float>(radius * radius)
        lists = []
        for i in range(x.shape[0]):
            d = x[i] - x
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            lists.append(np.flatnonzero(d2 < r2[i]))                       # FLANN RadiusResultSet: dist < radius
        out.append(lists)
    return out


def dbscan_labels(cloud, eps=0.9, min_pts=10, min_cluster=20, max_cluster=25000):
    n = cloud.shape[0]
    seed_nb, exp_nb = dbscan_radius_lists(cloud, eps)
    UN, ING, DONE = 0, 1, 2
    is_noise = [False] * n
    types = [UN] * n
    clusters = []
    for i in range(n):
        if types[i] == DONE:
            continue
        nn = seed_nb[i]
        if len(nn) < min_pts:
            is_noise[i] = True
            continue
        queue = [i]
        types[i] = DONE
        for j in nn:
            if j != i:
                queue.append(int(j))
                types[j] = ING
        sq = 1
        while sq < len(queue):
            ci = queue[sq]
            if is_noise[ci] or types[ci] == DONE:
                types[ci] = DONE
                sq += 1
                continue
            nn = exp_nb[ci]
            if len(nn) >= min_pts:
                for j in nn:
                    if types[j] == UN:
                        queue.append(int(j))
                        types[j] = ING
            types[ci] = DONE
            sq += 1
        if min_cluster <= len(queue) <= max_cluster:
            clusters.append(sorted(set(queue)))
    x = cloud[:, :3].astype(np.float32)
    dist = []
    for c in clusters:
        sx = sy = sz = np.float32(0)
        for idx in c:
            sx = np.float32(sx + x[idx, 0]); sy = np.float32(sy + x[idx, 1]); sz = np.float32(sz + x[idx, 2])
        m = np.float32(len(c))
        cx, cy, cz = np.float32(sx / m), np.float32(sy / m), np.float32(sz / m)
        dist.append(float(np.sqrt(np.float64(cx) ** 2 + np.float64(cy) ** 2 + np.float64(cz) ** 2)))
    labels = np.zeros(n, np.float32)
    for rank, k in enumerate(np.argsort(np.array(dist), kind="stable")):
        labels[clusters[k]] = rank + 1
    return labels, len(clusters)
