"""ctypes binding of libb2slam.so (include/b2slam.h) + thin torch-tensor plumbing.

There is no CPU fallback: importing works anywhere (so the C ABI can be
inspected on a CPU box), but creating an :class:`Engine` without the built
library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb2slam.so")

B2_OK = 0
P2P, P2L = 0, 1

# every symbol include/b2slam.h declares
EXPORTS = [
    "b2_version", "b2_create", "b2_destroy", "b2_last_error", "b2_set_stream", "b2_synchronize",
    "b2_launch_count", "b2_pack_xyz", "b2_unpack_xyz", "b2_project_pixels", "b2_voxel_downsample",
    "b2_estimate_normals", "b2_nn_search", "b2_icp", "b2_transform", "b2_map_init", "b2_map_reset",
    "b2_map_fuse", "b2_map_size", "b2_map_export", "b2_map_icp", "b2_map_icp_corr", "b2_slam_scan", "b2_slam_set_pose",
    "b2_slam_get_pose", "b2_icp_batch", "b2_slam_scan_dev", "b2_map_stencil_candidates",
    "b2_radius_outlier_mask", "b2_statistical_outlier_mask", "b2_select_points", "b2_stencil_candidates",
    "b2_set_icp_driver", "b2_ingest_pointcloud2", "b2_slam_results", "b2_icp_batch_p2l",
    "b2_estimate_normals_batch", "b2_map_export_records", "b2_map_merge_records", "b2_set_scan_path", "b2_scan_downsample", "b2_slam_scan_records",
    "b2_set_dense_driver",
]


class IcpParams(C.Structure):
    _fields_ = [("max_correspondence_distance", C.c_double), ("max_iteration", C.c_int), ("mode", C.c_int),
                ("relative_fitness", C.c_double), ("relative_rmse", C.c_double)]


class IcpResult(C.Structure):
    _fields_ = [("transformation", C.c_double * 16), ("fitness", C.c_double), ("inlier_rmse", C.c_double),
                ("iterations", C.c_int), ("n_correspondences", C.c_int)]

    def matrix(self) -> np.ndarray:
        return np.array(self.transformation[:], dtype=np.float64).reshape(4, 4)


RESULT_DTYPE = np.dtype([("transformation", np.float64, (4, 4)), ("fitness", np.float64),
                         ("inlier_rmse", np.float64), ("iterations", np.int32), ("n_correspondences", np.int32)])
assert RESULT_DTYPE.itemsize == C.sizeof(IcpResult)

_lib = None


def load():
    """dlopen libb2slam.so and declare the prototypes of include/b2slam.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_longlong, C.c_double
    pd = C.POINTER(C.c_double)
    lib.b2_version.restype = C.c_char_p
    lib.b2_last_error.restype = C.c_char_p
    lib.b2_last_error.argtypes = [vp]
    lib.b2_set_icp_driver.argtypes = [vp, i32]
    lib.b2_set_scan_path.argtypes = [vp, i32]
    lib.b2_set_dense_driver.argtypes = [vp, i32]
    lib.b2_launch_count.restype = C.c_ulonglong
    lib.b2_launch_count.argtypes = [vp]
    protos = {
        "b2_create": [i32, C.POINTER(vp)],
        "b2_destroy": [vp],
        "b2_set_stream": [vp, vp],
        "b2_synchronize": [vp],
        "b2_pack_xyz": [vp, vp, i32, i32, vp],
        "b2_unpack_xyz": [vp, vp, i32, vp, i32],
        "b2_project_pixels": [vp, vp, i32, f64, f64, f64, f64, i32, vp],
        "b2_ingest_pointcloud2": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, pd, i32, vp, C.POINTER(C.c_int)],
        "b2_voxel_downsample": [vp, vp, i32, f64, pd, pd, vp, vp, C.POINTER(i32)],
        "b2_scan_downsample": [vp, vp, i32, f64, vp, C.POINTER(i32)],
        "b2_estimate_normals": [vp, vp, i32, f64, i32, vp],
        "b2_nn_search": [vp, vp, i32, vp, i32, pd, f64, vp, vp],
        "b2_icp": [vp, vp, i32, vp, i32, vp, pd, C.POINTER(IcpParams), C.POINTER(IcpResult), vp],
        "b2_transform": [vp, vp, i32, pd, vp],
        "b2_map_init": [vp, f64, pd, f64, i64],
        "b2_map_reset": [vp],
        "b2_map_fuse": [vp, vp, i32, pd],
        "b2_map_size": [vp, C.POINTER(i64)],
        "b2_map_export": [vp, vp, vp, i64, C.POINTER(i64)],
        "b2_map_export_records": [vp, vp, vp, i64, C.POINTER(i64)],
        "b2_map_merge_records": [vp, vp, vp, i64],
        "b2_map_icp": [vp, vp, i32, pd, C.POINTER(IcpParams), C.POINTER(IcpResult)],
        "b2_map_icp_corr": [vp, vp, i32, pd, C.POINTER(IcpParams), C.POINTER(IcpResult), vp, vp],
        "b2_slam_scan": [vp, vp, i32, f64, C.POINTER(IcpParams), C.POINTER(IcpResult)],
        "b2_slam_scan_dev": [vp, vp, i32, f64, C.POINTER(IcpParams), C.POINTER(IcpResult)],
        "b2_slam_scan_records": [vp, vp, i32, f64, f64, f64, f64, i32, f64, C.POINTER(IcpParams), C.POINTER(IcpResult)],
        "b2_slam_results": [vp, i32, C.POINTER(IcpResult)],
        "b2_map_stencil_candidates": [vp, vp, i32, pd, C.POINTER(i64)],
        "b2_radius_outlier_mask": [vp, vp, i32, i32, f64, vp, C.POINTER(i32)],
        "b2_statistical_outlier_mask": [vp, vp, i32, i32, f64, f64, vp, C.POINTER(i32)],
        "b2_select_points": [vp, vp, i32, vp, vp, C.POINTER(i32)],
        "b2_stencil_candidates": [vp, vp, i32, vp, i32, pd, f64, C.POINTER(i64)],
        "b2_slam_set_pose": [vp, pd],
        "b2_slam_get_pose": [vp, pd],
        "b2_icp_batch": [vp, vp, vp, i32, vp, vp, i32, i32, f64, vp, C.POINTER(IcpParams), vp],
        "b2_icp_batch_p2l": [vp, vp, vp, i32, vp, vp, i32, i32, f64, f64, i32, vp, C.POINTER(IcpParams), vp],
        "b2_estimate_normals_batch": [vp, vp, vp, i32, i32, f64, i32, vp],
    }
    for name, args in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    _lib = lib
    return lib


class B2Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b2slam error {code}: {msg}")
        self.code = code


def _dptr(T):
    """float64[16] host pointer (or NULL) for a 4x4 array-like."""
    if T is None:
        return None, None
    a = np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(16))
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def make_params(max_corr: float, max_iter: int = 30, mode: int = P2P, rel_fitness: float = 1e-6,
                rel_rmse: float = 1e-6) -> IcpParams:
    return IcpParams(float(max_corr), int(max_iter), int(mode), float(rel_fitness), float(rel_rmse))


class Engine:
    """One handle of the C ABI bound to one CUDA device; torch tensors are only used as device buffers."""

    def __init__(self, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("b2slam needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = load()
        self.device = torch.device("cuda", device)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        rc = self.lib.b2_create(int(device), C.byref(h))
        if rc != B2_OK or not h:
            raise B2Error(rc, "b2_create failed (no usable CUDA device?)")
        self.h = h
        # All kernels run on one torch-visible side stream: tensors handed out by this class are
        # allocated under it (so the caching allocator tracks their use) and CUDA events recorded
        # on it time exactly the library's work.
        self.stream = torch.cuda.Stream(device=self.device)
        self._check(self.lib.b2_set_stream(self.h, C.c_void_p(self.stream.cuda_stream)))

    # -- plumbing -----------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.b2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != B2_OK:
            raise B2Error(rc, self.lib.b2_last_error(self.h).decode())

    def empty(self, shape, dtype):
        with self.torch.cuda.stream(self.stream):
            return self.torch.empty(shape, dtype=dtype, device=self.device)

    def empty_points(self, n: int):
        return self.empty((max(int(n), 0), 4), self.torch.float32)

    def _pts(self, t):
        torch = self.torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2
                and t.shape[1] == 4 and t.is_contiguous()):
            raise ValueError("expected a contiguous CUDA float32 [n,4] tensor")
        # inputs may have been produced on the caller's stream
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        return t

    def synchronize(self):
        self._check(self.lib.b2_synchronize(self.h))

    def set_icp_driver(self, driver) -> None:
        """'auto' | 'group' | 'sweep' | 'sweep4' (or 0..3): the kernel that runs ICP steps on pair / slab grids."""
        code = {"auto": 0, "group": 1, "sweep": 2, "sweep4": 3}.get(driver, driver)
        self._check(self.lib.b2_set_icp_driver(self.h, int(code)))

    def set_scan_path(self, path) -> None:
        """'hash' (default: sort-free scratch-hash passes) | 'sorted' (radix-sort pipeline) for slam_scan*."""
        code = {"hash": 0, "sorted": 1}.get(path, path)
        self._check(self.lib.b2_set_scan_path(self.h, int(code)))

    def set_dense_driver(self, driver) -> None:
        """'persistent' (default: one launch per registration against the brick-major map) | 'chained' (one launch
        per ICP step)."""
        code = {"persistent": 0, "chained": 1}.get(driver, driver)
        self._check(self.lib.b2_set_dense_driver(self.h, int(code)))

    def launch_count(self) -> int:
        return int(self.lib.b2_launch_count(self.h))

    # -- layout -------------------------------------------------------------
    def pack(self, xyz):
        """float32 [n,3] (numpy / CPU tensor / CUDA tensor) -> CUDA float32 [n,4]."""
        torch = self.torch
        if isinstance(xyz, torch.Tensor) and xyz.is_cuda:
            src = xyz.to(torch.float32).contiguous()
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            n = src.shape[0]
            out = self.empty_points(n)
            self._check(self.lib.b2_pack_xyz(self.h, src.data_ptr(), n, 0, out.data_ptr()))
            return out
        a = np.ascontiguousarray(xyz.numpy() if isinstance(xyz, torch.Tensor) else xyz, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] != 3:
            raise ValueError(f"expected [n,3] points, got {a.shape}")
        out = self.empty_points(len(a))
        self._check(self.lib.b2_pack_xyz(self.h, a.ctypes.data, len(a), 1, out.data_ptr()))
        self.synchronize()
        return out

    def unpack(self, pts) -> np.ndarray:
        pts = self._pts(pts)
        out = np.empty((pts.shape[0], 3), dtype=np.float32)
        self._check(self.lib.b2_unpack_xyz(self.h, pts.data_ptr(), pts.shape[0], out.ctypes.data, 1))
        return out

    def project_pixels(self, records: np.ndarray, fx, fy, cx, cy, numpy2: bool = False):
        rec = np.ascontiguousarray(records, dtype=np.float32)
        out = self.empty_points(len(rec))
        self._check(self.lib.b2_project_pixels(self.h, rec.ctypes.data, len(rec), fx, fy, cx, cy, int(numpy2),
                                               out.data_ptr()))
        self.synchronize()
        return out

    def ingest_pointcloud2(self, data, n_points: int, point_step: int = 12, offsets=(0, 4, 8),
                           is_bigendian: bool = False, skip_nans: bool = True, intrinsics=None,
                           numpy2: bool = False):
        """PointCloud2 payload (bytes / bytearray / numpy uint8 on the host, or a CUDA uint8 tensor) ->
        CUDA float32 [m,4] rows of the records without NaN, in order; with ``intrinsics=(fx, fy, cx, cy)``
        the pixel->metric projection is applied in the same pass (b2_ingest_pointcloud2)."""
        torch = self.torch
        on_host = not (isinstance(data, torch.Tensor) and data.is_cuda)
        if on_host:
            if isinstance(data, torch.Tensor):
                data = data.contiguous().numpy()
            if isinstance(data, np.ndarray):
                buf = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
            else:
                buf = np.frombuffer(data, dtype=np.uint8)
            nbytes, ptr = buf.size, (buf.ctypes.data if buf.size else None)
        else:
            buf = data.contiguous().view(torch.uint8).reshape(-1)
            nbytes, ptr = buf.numel(), buf.data_ptr()
        if nbytes < int(n_points) * int(point_step):
            raise ValueError(f"payload of {nbytes} bytes is shorter than {n_points} x {point_step}")
        out = self.empty_points(int(n_points))
        k = (C.c_double * 4)(*[float(v) for v in intrinsics]) if intrinsics is not None else None
        m = C.c_int(0)
        self._check(self.lib.b2_ingest_pointcloud2(self.h, ptr, int(on_host), int(n_points), int(point_step),
                                                   int(offsets[0]), int(offsets[1]), int(offsets[2]),
                                                   int(is_bigendian), int(skip_nans), k, int(numpy2),
                                                   out.data_ptr(), C.byref(m)))
        if on_host:
            self.synchronize()  # the staging copy reads the caller's buffer
        return out[:m.value]

    # -- primitives ---------------------------------------------------------
    def voxel_downsample(self, pts, voxel: float, origin=None, T=None, want_keys: bool = True):
        """-> (centroids [m,4] (w = count bits), keys int32 [m,3] or None)."""
        torch = self.torch
        pts = self._pts(pts)
        n = pts.shape[0]
        out = self.empty_points(n)
        keys = self.empty((n, 3), torch.int32) if want_keys else None
        _o, op = (None, None) if origin is None else _dptr3(origin)
        _t, tp = _dptr(T)
        m = C.c_int(0)
        self._check(self.lib.b2_voxel_downsample(self.h, pts.data_ptr(), n, float(voxel), op, tp, out.data_ptr(),
                                                 keys.data_ptr() if want_keys else None, C.byref(m)))
        return out[:m.value], (keys[:m.value] if want_keys else None)

    def scan_downsample(self, pts, voxel: float):
        """Sort-free voxel down-sample of one scan (Open3D origin) -> centroids [m,4] (w = count bits) in the
        order of each voxel's first input point."""
        pts = self._pts(pts)
        n = pts.shape[0]
        out = self.empty_points(n)
        m = C.c_int(0)
        self._check(self.lib.b2_scan_downsample(self.h, pts.data_ptr(), n, float(voxel), out.data_ptr(), C.byref(m)))
        return out[:m.value]

    def estimate_normals(self, pts, radius: float, max_nn: int):
        pts = self._pts(pts)
        out = self.empty_points(pts.shape[0])
        self._check(self.lib.b2_estimate_normals(self.h, pts.data_ptr(), pts.shape[0], float(radius), int(max_nn),
                                                 out.data_ptr()))
        return out

    def nn_search(self, src, tgt, radius: float, T=None):
        torch = self.torch
        src, tgt = self._pts(src), self._pts(tgt)
        idx = self.empty(src.shape[0], torch.int32)
        d2 = self.empty(src.shape[0], torch.float64)
        _t, tp = _dptr(T)
        self._check(self.lib.b2_nn_search(self.h, src.data_ptr(), src.shape[0], tgt.data_ptr(), tgt.shape[0], tp,
                                          float(radius), idx.data_ptr(), d2.data_ptr()))
        return idx, d2

    def icp(self, src, tgt, params: IcpParams, init=None, tgt_normals=None, want_corr: bool = False):
        torch = self.torch
        src, tgt = self._pts(src), self._pts(tgt)
        corr = self.empty(src.shape[0], torch.int32) if want_corr else None
        _i, ip = _dptr(init)
        res = IcpResult()
        self._check(self.lib.b2_icp(self.h, src.data_ptr(), src.shape[0], tgt.data_ptr(), tgt.shape[0],
                                    self._pts(tgt_normals).data_ptr() if tgt_normals is not None else None, ip,
                                    C.byref(params), C.byref(res), corr.data_ptr() if want_corr else None))
        return res, corr

    def transform(self, pts, T, out=None):
        pts = self._pts(pts)
        out = self.empty_points(pts.shape[0]) if out is None else out
        _t, tp = _dptr(T)
        self._check(self.lib.b2_transform(self.h, pts.data_ptr(), pts.shape[0], tp, out.data_ptr()))
        return out

    # -- map maintenance ------------------------------------------------------
    def radius_outlier_mask(self, pts, nb_points: int, radius: float):
        """-> (keep uint8 [n] CUDA tensor, number kept)."""
        pts = self._pts(pts)
        keep = self.empty(pts.shape[0], self.torch.uint8)
        m = C.c_int(0)
        self._check(self.lib.b2_radius_outlier_mask(self.h, pts.data_ptr(), pts.shape[0], int(nb_points),
                                                    float(radius), keep.data_ptr(), C.byref(m)))
        return keep, m.value

    def statistical_outlier_mask(self, pts, nb_neighbors: int, std_ratio: float, cell: float = 0.0):
        pts = self._pts(pts)
        keep = self.empty(pts.shape[0], self.torch.uint8)
        m = C.c_int(0)
        self._check(self.lib.b2_statistical_outlier_mask(self.h, pts.data_ptr(), pts.shape[0], int(nb_neighbors),
                                                         float(std_ratio), float(cell), keep.data_ptr(), C.byref(m)))
        return keep, m.value

    def select_points(self, pts, keep):
        pts = self._pts(pts)
        out = self.empty_points(pts.shape[0])
        m = C.c_int(0)
        self._check(self.lib.b2_select_points(self.h, pts.data_ptr(), pts.shape[0], keep.data_ptr(), out.data_ptr(),
                                              C.byref(m)))
        return out[:m.value]

    # -- resident map ---------------------------------------------------------
    def map_init(self, voxel: float, origin, max_corr: float, capacity_cells: int):
        _o, op = _dptr3(origin)
        self._check(self.lib.b2_map_init(self.h, float(voxel), op, float(max_corr), int(capacity_cells)))

    def map_reset(self):
        self._check(self.lib.b2_map_reset(self.h))

    def map_fuse(self, pts, T=None):
        pts = self._pts(pts)
        _t, tp = _dptr(T)
        self._check(self.lib.b2_map_fuse(self.h, pts.data_ptr(), pts.shape[0], tp))

    def map_size(self) -> int:
        n = C.c_longlong(0)
        self._check(self.lib.b2_map_size(self.h, C.byref(n)))
        return int(n.value)

    def map_export(self, want_keys: bool = True):
        torch = self.torch
        n = self.map_size()
        out = self.empty_points(n)
        keys = self.empty((n, 3), torch.int32) if want_keys else None
        m = C.c_longlong(0)
        if n:
            self._check(self.lib.b2_map_export(self.h, out.data_ptr(), keys.data_ptr() if want_keys else None, n,
                                               C.byref(m)))
        return out[:m.value], (keys[:m.value] if want_keys else None)

    def map_export_records(self):
        """The map as mergeable records in canonical key order -> (keys int32 [m,3], sums float64 [m,4] =
        (sum x, sum y, sum z, count)) — what ranks exchange when partial maps are gathered."""
        torch = self.torch
        n = self.map_size()
        keys = self.empty((n, 3), torch.int32)
        sums = self.empty((n, 4), torch.float64)
        m = C.c_longlong(0)
        if n:
            self._check(self.lib.b2_map_export_records(self.h, keys.data_ptr(), sums.data_ptr(), n, C.byref(m)))
        return keys[:m.value], sums[:m.value]

    def map_merge_records(self, keys, sums):
        """Add exported voxel records (unique keys) to the resident map, weighted by their counts."""
        torch = self.torch
        if not (keys.is_cuda and sums.is_cuda and keys.dtype == torch.int32 and sums.dtype == torch.float64
                and keys.is_contiguous() and sums.is_contiguous() and keys.shape[0] == sums.shape[0]):
            raise ValueError("expected contiguous CUDA int32 [n,3] keys and float64 [n,4] sums")
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self._check(self.lib.b2_map_merge_records(self.h, keys.data_ptr(), sums.data_ptr(), keys.shape[0]))

    def map_icp(self, src, params: IcpParams, init=None) -> IcpResult:
        src = self._pts(src)
        _i, ip = _dptr(init)
        res = IcpResult()
        self._check(self.lib.b2_map_icp(self.h, src.data_ptr(), src.shape[0], ip, C.byref(params), C.byref(res)))
        return res

    def map_icp_corr(self, src, params: IcpParams, init=None):
        """-> (result, d2 float64 [ns], match float32 [ns,4]): the registration plus the correspondence set of the
        last evaluated transform (squared distances and matched voxel centroids, w = 1 where matched)."""
        torch = self.torch
        src = self._pts(src)
        _i, ip = _dptr(init)
        d2 = self.empty(src.shape[0], torch.float64)
        match = self.empty((src.shape[0], 4), torch.float32)
        res = IcpResult()
        self._check(self.lib.b2_map_icp_corr(self.h, src.data_ptr(), src.shape[0], ip, C.byref(params), C.byref(res),
                                             d2.data_ptr(), match.data_ptr()))
        return res, d2, match

    def slam_scan(self, xyz_host: np.ndarray, ds_voxel: float, params: IcpParams, want_result: bool = True):
        """xyz_host: float32 [n,3] C-contiguous host array, pinned (uploaded in place) or pageable (staged
        through the library's pinned ring).  With
        ``want_result=False`` the call only enqueues work (a PINNED array must then stay alive until the next
        synchronising call); fetch the poses later with ``slam_results``."""
        if xyz_host.dtype != np.float32 or xyz_host.ndim != 2 or xyz_host.shape[1] != 3 or \
                not xyz_host.flags["C_CONTIGUOUS"]:
            raise ValueError("expected a C-contiguous float32 [n,3] host array")
        res = IcpResult() if want_result else None
        self._check(self.lib.b2_slam_scan(self.h, xyz_host.ctypes.data, xyz_host.shape[0], float(ds_voxel),
                                          C.byref(params), C.byref(res) if want_result else None))
        return res

    def slam_scan_records(self, records_host: np.ndarray, fx, fy, cx, cy, numpy2: bool, ds_voxel: float,
                          params: IcpParams, want_result: bool = True):
        """The scan step fed with the wire-format records (x_px, y_px, depth) float32 [n,3] of one depth image:
        the projection (generic_point_cloud_processor.py:93-108) runs on the device as part of the upload."""
        if records_host.dtype != np.float32 or records_host.ndim != 2 or records_host.shape[1] != 3 or \
                not records_host.flags["C_CONTIGUOUS"]:
            raise ValueError("expected a C-contiguous float32 [n,3] host array of (x_px, y_px, depth) records")
        res = IcpResult() if want_result else None
        self._check(self.lib.b2_slam_scan_records(self.h, records_host.ctypes.data, records_host.shape[0], float(fx),
                                                  float(fy), float(cx), float(cy), int(bool(numpy2)), float(ds_voxel),
                                                  C.byref(params), C.byref(res) if want_result else None))
        return res

    def slam_scan_dev(self, scan4, ds_voxel: float, params: IcpParams, want_result: bool = False):
        """Same step for a scan already resident on the device (float32 [n,4]); asynchronous unless
        a result is requested."""
        scan4 = self._pts(scan4)
        res = IcpResult() if want_result else None
        self._check(self.lib.b2_slam_scan_dev(self.h, scan4.data_ptr(), scan4.shape[0], float(ds_voxel),
                                              C.byref(params), C.byref(res) if want_result else None))
        return res

    def slam_results(self, count: int):
        """Results of the last ``count`` scans submitted with ``want_result=False``, oldest first (one stream
        synchronisation for the whole batch)."""
        arr = (IcpResult * int(count))()
        self._check(self.lib.b2_slam_results(self.h, int(count), arr))
        return list(arr)

    def stencil_candidates(self, src, T=None, tgt=None, cell: float = 0.0) -> int:
        """Points under the 27-cell stencils of the (transformed) source: against the resident map, or
        against `tgt` gridded with edge `cell`."""
        src = self._pts(src)
        _t, tp = _dptr(T)
        n = C.c_longlong(0)
        if tgt is None:
            self._check(self.lib.b2_map_stencil_candidates(self.h, src.data_ptr(), src.shape[0], tp, C.byref(n)))
        else:
            tgt = self._pts(tgt)
            self._check(self.lib.b2_stencil_candidates(self.h, src.data_ptr(), src.shape[0], tgt.data_ptr(),
                                                       tgt.shape[0], tp, float(cell), C.byref(n)))
        return int(n.value)

    def set_pose(self, T):
        _t, tp = _dptr(T)
        self._check(self.lib.b2_slam_set_pose(self.h, tp))

    def get_pose(self) -> np.ndarray:
        out = np.zeros(16, dtype=np.float64)
        self._check(self.lib.b2_slam_get_pose(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.reshape(4, 4)

    # -- batch ----------------------------------------------------------------
    def icp_batch(self, src, src_offsets, tgt, tgt_offsets, params: IcpParams, ds_voxel: float = 0.0, init=None,
                  normal_radius: float = 0.0, normal_max_nn: int = 20, sync: bool = True):
        """src/tgt: CUDA float32 [sum,4]; offsets: CUDA int32 [n_pairs+1].
        -> CUDA uint8 tensor viewable as RESULT_DTYPE records (stays on the device for NCCL gathers).
        ``params.mode == P2L``: the normals of every (down-sampled) target are estimated first with
        (``normal_radius``, ``normal_max_nn``).

        The C call only enqueues the batch on the engine's stream.  ``sync=True`` (default) waits for it and
        raises device-side failures here; with ``sync=False`` the caller's current stream is ordered after the
        engine's stream (so ``.cpu()`` or a collective on it sees finished records) and failures surface at the
        next ``synchronize()``."""
        torch = self.torch
        src, tgt = self._pts(src), self._pts(tgt)
        n_pairs = src_offsets.numel() - 1
        assert src_offsets.dtype == torch.int32 and tgt_offsets.dtype == torch.int32
        res = self.empty((n_pairs, C.sizeof(IcpResult)), torch.uint8)
        ip = init.data_ptr() if init is not None else None
        if params.mode == P2L:
            self._check(self.lib.b2_icp_batch_p2l(self.h, src.data_ptr(), src_offsets.data_ptr(), src.shape[0],
                                                  tgt.data_ptr(), tgt_offsets.data_ptr(), tgt.shape[0], n_pairs,
                                                  float(ds_voxel), float(normal_radius), int(normal_max_nn), ip,
                                                  C.byref(params), res.data_ptr()))
        else:
            self._check(self.lib.b2_icp_batch(self.h, src.data_ptr(), src_offsets.data_ptr(), src.shape[0],
                                              tgt.data_ptr(), tgt_offsets.data_ptr(), tgt.shape[0], n_pairs,
                                              float(ds_voxel), ip, C.byref(params), res.data_ptr()))
        if sync:
            self.synchronize()
        else:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return res

    def estimate_normals_batch(self, pts, offsets, radius: float, max_nn: int):
        """Normals of a batch of clouds (concatenated rows, CUDA int32 [n_clouds+1] offsets)."""
        pts = self._pts(pts)
        assert offsets.dtype == self.torch.int32
        out = self.empty_points(pts.shape[0])
        self._check(self.lib.b2_estimate_normals_batch(self.h, pts.data_ptr(), offsets.data_ptr(), pts.shape[0],
                                                       offsets.numel() - 1, float(radius), int(max_nn), out.data_ptr()))
        return out


def _dptr3(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(3))
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def results_to_numpy(res_tensor) -> np.ndarray:
    """uint8 [n, sizeof(b2_icp_result)] tensor -> structured numpy array."""
    return res_tensor.cpu().numpy().view(RESULT_DTYPE).reshape(-1)
