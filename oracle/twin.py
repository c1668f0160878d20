"""Independent NumPy/SciPy twin of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE.  A second, separately written statement of SURVEY.md
Appendix A built on library routines (scipy.spatial.cKDTree, numpy.linalg.svd /
eigh) — used only by tests/ to cross-validate the C++ oracle, because the
reference pins nothing for this path and Open3D is not available
(PARITY UNPINNED, see oracle.cpp).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def voxel_downsample(pts, voxel, origin=None):
    """A.1: -> centroids, keys, counts in lexicographic key order."""
    pts = np.asarray(pts, dtype=np.float64)
    if origin is None:
        origin = pts.min(axis=0) - 0.5 * voxel
    keys = np.floor((pts - np.asarray(origin, dtype=np.float64)) / voxel).astype(np.int32)
    uk, inv, cnt = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    sums = np.zeros((len(uk), 3))
    np.add.at(sums, inv, pts)  # unbuffered: accumulates in input order
    return sums / cnt[:, None], uk, cnt.astype(np.int32)


def nn_search(src, tgt, radius):
    """A.3/A.4: strict d2 < r2, 1-NN."""
    src = np.asarray(src, dtype=np.float64)
    tgt = np.asarray(tgt, dtype=np.float64)
    tree = cKDTree(tgt)
    _, j = tree.query(src, k=1)
    diff = src - tgt[j]
    d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
    ok = d2 < radius * radius
    return np.where(ok, j, -1).astype(np.int32), np.where(ok, d2, 0.0)


def umeyama(P, Q):
    """A.5: rigid T minimising sum |R p + t - q|^2 (Eigen::umeyama, no scaling)."""
    P = np.asarray(P, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    mp, mq = P.mean(axis=0), Q.mean(axis=0)
    S = (Q - mq).T @ (P - mp) / len(P)
    U, _, Vt = np.linalg.svd(S)
    s = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        s[2] = -1
    R = U @ np.diag(s) @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = mq - R @ mp
    return T


def vec6_to_mat4(x):
    a, b, g = x[:3]
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = x[3:]
    return T


def point_to_plane_step(P, Q, N):
    """A.6 on already-paired rows."""
    P, Q, N = (np.asarray(a, dtype=np.float64) for a in (P, Q, N))
    r = np.einsum("ij,ij->i", P - Q, N)
    J = np.concatenate([np.cross(P, N), N], axis=1)
    x = np.linalg.solve(J.T @ J, -(J.T @ r))
    return vec6_to_mat4(x)


def transform(pts, T):
    pts = np.asarray(pts, dtype=np.float64)
    return pts @ T[:3, :3].T + T[:3, 3]


def estimate_normals(pts, radius, max_nn):
    """A.2 with a library eigen-solver; sign is arbitrary."""
    pts = np.asarray(pts, dtype=np.float64)
    tree = cKDTree(pts)
    d, j = tree.query(pts, k=min(max_nn, len(pts)))
    if d.ndim == 1:
        d, j = d[:, None], j[:, None]
    out = np.zeros_like(pts)
    for i in range(len(pts)):
        nb = j[i][d[i] ** 2 < radius * radius]
        if len(nb) < 3:
            out[i] = (0, 0, 1)
            continue
        q = pts[nb]
        C = (q.T @ q) / len(nb) - np.outer(q.mean(0), q.mean(0))
        w, v = np.linalg.eigh(C)
        out[i] = v[:, 0]
    return out


def registration_icp(src, tgt, max_corr, init=None, max_iteration=30, rel_fitness=1e-6, rel_rmse=1e-6,
                     mode="p2p", tgt_normals=None):
    """A.3 loop. -> T, fitness, rmse, iterations, corr."""
    src = np.asarray(src, dtype=np.float64)
    tgt = np.asarray(tgt, dtype=np.float64)
    T = np.eye(4) if init is None else np.array(init, dtype=np.float64)
    pcd = transform(src, T)

    def ev(p):
        idx, d2 = nn_search(p, tgt, max_corr)
        m = idx >= 0
        n = int(m.sum())
        if n == 0:
            return idx, 0.0, 0.0
        return idx, n / len(p), float(np.sqrt(d2[m].sum() / n))

    idx, fit, rmse = ev(pcd)
    it = 0
    while it < max_iteration:
        m = idx >= 0
        if m.any():
            if mode == "p2p":
                upd = umeyama(pcd[m], tgt[idx[m]])
            else:
                upd = point_to_plane_step(pcd[m], tgt[idx[m]], np.asarray(tgt_normals)[idx[m]])
        else:
            upd = np.eye(4)
        T = upd @ T
        pcd = transform(pcd, upd)
        pf, pr = fit, rmse
        idx, fit, rmse = ev(pcd)
        it += 1
        if abs(pf - fit) < rel_fitness and abs(pr - rmse) < rel_rmse:
            break
    return T, fit, rmse, it, idx


def statistical_outlier_mask(pts, nb_neighbors, std_ratio):
    pts = np.asarray(pts, dtype=np.float64)
    tree = cKDTree(pts)
    d, _ = tree.query(pts, k=min(nb_neighbors, len(pts)))
    avg = d.mean(axis=1)
    ok = avg > 0
    mu = avg[ok].mean()
    sd = np.sqrt(((avg[ok] - mu) ** 2).sum() / (ok.sum() - 1))
    return ok & (avg < mu + std_ratio * sd)


def radius_outlier_mask(pts, nb_points, radius):
    pts = np.asarray(pts, dtype=np.float64)
    tree = cKDTree(pts)
    cnt = tree.query_ball_point(pts, r=radius, return_length=True)
    return cnt > nb_points
