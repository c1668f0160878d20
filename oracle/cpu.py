"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
PARITY UNPINNED: see the header of oracle.cpp (no Open3D, no reference golden
vectors exist for this path).

Every function takes/returns float64 NumPy arrays, the dtype Open3D computes in
(Vector3dVector widens float32 input exactly).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++, OpenMP)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_voxel_downsample.restype = C.c_int64
        _lib.orc_nn_search.restype = C.c_int64
        _lib.orc_nn_search_brute.restype = C.c_int64
        _lib.orc_stencil_candidates.restype = C.c_int64
        _lib.orc_statistical_outlier_mask.restype = C.c_int64
        _lib.orc_radius_outlier_mask.restype = C.c_int64
        _lib.orc_registration_icp.restype = C.c_int
    return _lib


def _pts(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError(f"expected [N,3] points, got {a.shape}")
    return a


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(n)))


def min_max_bound(pts):
    pts = _pts(pts)
    mn = np.zeros(3)
    mx = np.zeros(3)
    lib().orc_min_max_bound(_ptr(pts, C.c_double), C.c_int64(len(pts)), _ptr(mn, C.c_double), _ptr(mx, C.c_double))
    return mn, mx


def voxel_keys(pts, voxel: float, origin=None) -> np.ndarray:
    pts = _pts(pts)
    keys = np.zeros((len(pts), 3), dtype=np.int32)
    org = None if origin is None else np.ascontiguousarray(origin, dtype=np.float64)
    lib().orc_voxel_keys(_ptr(pts, C.c_double), C.c_int64(len(pts)), C.c_double(voxel), _ptr(org, C.c_double),
                         _ptr(keys, C.c_int32))
    return keys


def voxel_downsample(pts, voxel: float, origin=None):
    """-> (centroids [M,3] f64, keys [M,3] i32, counts [M] i32), canonical (lexicographic key) order."""
    pts = _pts(pts)
    n = len(pts)
    out = np.zeros((n, 3))
    keys = np.zeros((n, 3), dtype=np.int32)
    cnt = np.zeros(n, dtype=np.int32)
    org = None if origin is None else np.ascontiguousarray(origin, dtype=np.float64)
    m = lib().orc_voxel_downsample(_ptr(pts, C.c_double), C.c_int64(n), C.c_double(voxel), _ptr(org, C.c_double),
                                   _ptr(out, C.c_double), _ptr(keys, C.c_int32), _ptr(cnt, C.c_int32))
    return out[:m].copy(), keys[:m].copy(), cnt[:m].copy()


def transform(pts, T) -> np.ndarray:
    pts = _pts(pts)
    T = np.ascontiguousarray(T, dtype=np.float64)
    out = np.empty_like(pts)
    lib().orc_transform(_ptr(pts, C.c_double), C.c_int64(len(pts)), _ptr(T, C.c_double), _ptr(out, C.c_double))
    return out


def nn_search(src, tgt, radius: float, brute: bool = False):
    """-> (idx [Ns] i32 (-1 = none), d2 [Ns] f64)."""
    src, tgt = _pts(src), _pts(tgt)
    idx = np.zeros(len(src), dtype=np.int32)
    d2 = np.zeros(len(src))
    fn = lib().orc_nn_search_brute if brute else lib().orc_nn_search
    fn(_ptr(src, C.c_double), C.c_int64(len(src)), _ptr(tgt, C.c_double), C.c_int64(len(tgt)), C.c_double(radius),
       _ptr(idx, C.c_int32), _ptr(d2, C.c_double))
    return idx, d2


def knn(query, tgt, k: int, radius: float = 0.0):
    """kNN (radius<=0) or hybrid search. -> (idx [N,k], d2 [N,k], cnt [N])."""
    query, tgt = _pts(query), _pts(tgt)
    idx = np.zeros((len(query), k), dtype=np.int32)
    d2 = np.zeros((len(query), k))
    cnt = np.zeros(len(query), dtype=np.int32)
    lib().orc_knn(_ptr(query, C.c_double), C.c_int64(len(query)), _ptr(tgt, C.c_double), C.c_int64(len(tgt)),
                  C.c_int(k), C.c_double(radius), _ptr(idx, C.c_int32), _ptr(d2, C.c_double), _ptr(cnt, C.c_int32))
    return idx, d2, cnt


def stencil_candidates(query, tgt, cell: float, origin) -> int:
    query, tgt = _pts(query), _pts(tgt)
    org = np.ascontiguousarray(origin, dtype=np.float64)
    return int(lib().orc_stencil_candidates(_ptr(query, C.c_double), C.c_int64(len(query)), _ptr(tgt, C.c_double),
                                            C.c_int64(len(tgt)), C.c_double(cell), _ptr(org, C.c_double)))


def umeyama(src, tgt, corr_src, corr_tgt) -> np.ndarray:
    src, tgt = _pts(src), _pts(tgt)
    cs = np.ascontiguousarray(corr_src, dtype=np.int32)
    ct = np.ascontiguousarray(corr_tgt, dtype=np.int32)
    T = np.zeros((4, 4))
    lib().orc_umeyama(_ptr(src, C.c_double), _ptr(tgt, C.c_double), _ptr(cs, C.c_int32), _ptr(ct, C.c_int32),
                      C.c_int64(len(cs)), _ptr(T, C.c_double))
    return T


def point_to_plane_step(src, tgt, tgt_normals, corr_src, corr_tgt) -> np.ndarray:
    src, tgt, nrm = _pts(src), _pts(tgt), _pts(tgt_normals)
    cs = np.ascontiguousarray(corr_src, dtype=np.int32)
    ct = np.ascontiguousarray(corr_tgt, dtype=np.int32)
    T = np.zeros((4, 4))
    lib().orc_point_to_plane_step(_ptr(src, C.c_double), _ptr(tgt, C.c_double), _ptr(nrm, C.c_double),
                                  _ptr(cs, C.c_int32), _ptr(ct, C.c_int32), C.c_int64(len(cs)), _ptr(T, C.c_double))
    return T


def svd3(M):
    M = np.ascontiguousarray(M, dtype=np.float64).reshape(3, 3)
    U, S, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
    lib().orc_svd3(_ptr(M, C.c_double), _ptr(U, C.c_double), _ptr(S, C.c_double), _ptr(V, C.c_double))
    return U, S, V


def eig3_smallest(Cm) -> np.ndarray:
    Cm = np.ascontiguousarray(Cm, dtype=np.float64).reshape(3, 3)
    v = np.zeros(3)
    lib().orc_eig3_smallest(_ptr(Cm, C.c_double), _ptr(v, C.c_double))
    return v


def estimate_normals(pts, radius: float, max_nn: int) -> np.ndarray:
    pts = _pts(pts)
    out = np.zeros_like(pts)
    lib().orc_estimate_normals(_ptr(pts, C.c_double), C.c_int64(len(pts)), C.c_double(radius), C.c_int(max_nn),
                               _ptr(out, C.c_double))
    return out


class ICPResult:
    __slots__ = ("transformation", "fitness", "inlier_rmse", "iterations", "correspondence")

    def __init__(self, T, fitness, rmse, iters, corr):
        self.transformation = T
        self.fitness = fitness
        self.inlier_rmse = rmse
        self.iterations = iters
        self.correspondence = corr


def registration_icp(src, tgt, max_corr: float, init=None, mode: str = "p2p", tgt_normals=None,
                     max_iteration: int = 30, relative_fitness: float = 1e-6,
                     relative_rmse: float = 1e-6) -> ICPResult:
    src, tgt = _pts(src), _pts(tgt)
    init = np.eye(4) if init is None else np.ascontiguousarray(init, dtype=np.float64)
    nrm = None if tgt_normals is None else _pts(tgt_normals)
    T = np.zeros((4, 4))
    fit, rmse, iters = C.c_double(0), C.c_double(0), C.c_int32(0)
    corr = np.zeros(len(src), dtype=np.int32)
    rc = lib().orc_registration_icp(
        _ptr(src, C.c_double), C.c_int64(len(src)), _ptr(tgt, C.c_double), C.c_int64(len(tgt)),
        _ptr(nrm, C.c_double), C.c_double(max_corr), _ptr(init, C.c_double),
        C.c_int({"p2p": 0, "p2l": 1}[mode]), C.c_int(max_iteration), C.c_double(relative_fitness),
        C.c_double(relative_rmse), _ptr(T, C.c_double), C.byref(fit), C.byref(rmse), C.byref(iters),
        _ptr(corr, C.c_int32))
    if rc != 0:
        raise RuntimeError("point-to-plane ICP requires target normals")
    return ICPResult(T, fit.value, rmse.value, iters.value, corr)


def statistical_outlier_mask(pts, nb_neighbors: int, std_ratio: float) -> np.ndarray:
    pts = _pts(pts)
    keep = np.zeros(len(pts), dtype=np.uint8)
    lib().orc_statistical_outlier_mask(_ptr(pts, C.c_double), C.c_int64(len(pts)), C.c_int(nb_neighbors),
                                       C.c_double(std_ratio), _ptr(keep, C.c_uint8))
    return keep.astype(bool)


def radius_outlier_mask(pts, nb_points: int, radius: float) -> np.ndarray:
    pts = _pts(pts)
    keep = np.zeros(len(pts), dtype=np.uint8)
    lib().orc_radius_outlier_mask(_ptr(pts, C.c_double), C.c_int64(len(pts)), C.c_int(nb_points),
                                  C.c_double(radius), _ptr(keep, C.c_uint8))
    return keep.astype(bool)


class ResidentMapCPU:
    """The resident-map CPU leg (oracle.cpp: ResidentMap): the map stays a voxel hash of (sum, n)
    accumulators that doubles as the search grid, nothing is re-voxelised or re-indexed per scan.
    The C++ twin of oracle/pipeline.py:ResidentSequence (same results, bit for bit) and the
    like-for-like CPU baseline of b2_slam_scan (icp_processor.py:43-119 with a resident index)."""

    def __init__(self, map_voxel, map_origin, ds_voxel=0.02, max_corr=0.05, max_iteration=30,
                 relative_fitness=1e-6, relative_rmse=1e-6):
        L = lib()
        L.orc_resident_create.restype = C.c_void_p
        L.orc_resident_size.restype = C.c_int64
        L.orc_resident_process.restype = C.c_int
        org = np.ascontiguousarray(map_origin, dtype=np.float64)
        self._h = C.c_void_p(L.orc_resident_create(C.c_double(map_voxel), _ptr(org, C.c_double), C.c_double(ds_voxel),
                                                   C.c_double(max_corr), C.c_int(max_iteration),
                                                   C.c_double(relative_fitness), C.c_double(relative_rmse)))
        self.last_result = None

    def close(self):
        if self._h:
            lib().orc_resident_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self) -> int:
        return int(lib().orc_resident_size(self._h))

    @property
    def pose(self) -> np.ndarray:
        T = np.zeros((4, 4))
        lib().orc_resident_get_pose(self._h, _ptr(T, C.c_double))
        return T

    @pose.setter
    def pose(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64)
        lib().orc_resident_set_pose(self._h, _ptr(T, C.c_double))

    def fuse(self, points, T=None):
        p = np.ascontiguousarray(points, dtype=np.float32)
        T = np.eye(4) if T is None else np.ascontiguousarray(T, dtype=np.float64)
        lib().orc_resident_fuse(self._h, _ptr(p, C.c_float), C.c_int64(len(p)), _ptr(T, C.c_double))

    def process(self, points):
        p = np.ascontiguousarray(points, dtype=np.float32)
        T = np.zeros((4, 4))
        fit, rmse, iters = C.c_double(0), C.c_double(0), C.c_int32(0)
        rc = lib().orc_resident_process(self._h, _ptr(p, C.c_float), C.c_int64(len(p)), _ptr(T, C.c_double),
                                        C.byref(fit), C.byref(rmse), C.byref(iters))
        if rc == 1:
            return None
        self.last_result = ICPResult(T, fit.value, rmse.value, iters.value, None)
        return self.last_result

    def map_points(self):
        n = self.size()
        keys = np.zeros((n, 3), dtype=np.int32)
        cent = np.zeros((n, 3), dtype=np.float32)
        cnt = np.zeros(n, dtype=np.int32)
        lib().orc_resident_export(self._h, _ptr(keys, C.c_int32), _ptr(cent, C.c_float), _ptr(cnt, C.c_int32))
        return cent, keys, cnt
