"""Developer script (library built with -DB2_TAIL_TIMING): launch-to-launch period of the dense ICP step kernel inside
the graph-replayed scan step, from the kernel's own globaltimer stamps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lidar_slam_b200 as b2
import bench

eng = b2.Engine(0)
dev = torch.device("cuda", 0)
scene = bench.build_scene(0)
bench.preload_map(eng, scene, dev)
scans, poses = bench.make_scans(scene, 12, dev, seed=1)
prm = b2.make_params(0.05, 30)
d4 = [eng.pack(s) for s in scans]
eng.set_pose(poses[0])
mis = np.zeros(64, dtype=np.uint32)
eng.lib.b2_debug_dense_full_path.argtypes = [C.c_void_p]
for s in d4:
    eng.slam_scan_dev(s, 0.02, prm)
    eng.synchronize()
    eng.lib.b2_debug_dense_full_path(mis.ctypes.data)
    print("queries on the full path per launch:", mis[:34].tolist())
eng.synchronize()
res = eng.slam_results(1)[0]
buf = np.zeros(64 * 148 * 16, dtype=np.uint64)
eng.lib.b2_debug_dense_stamps.argtypes = [C.c_void_p]
assert eng.lib.b2_debug_dense_stamps(buf.ctypes.data) == 0
ts = buf.reshape(64, 148, 16).astype(np.int64)
n = res.iterations + 2
names = ["entry", "waited", "folded", "solved", "searched", "exit", "w0:arrived", "w0:spec", "w0:Tseen", "w0:cell", "w0:reloaded", "w0:finished", "last:spec", "last:reloaded", "last:arrived"]
print("last scan: iterations", res.iterations)
first = ts[:n, :, 0].min(axis=1)
last_exit = ts[:n, :, 5].max(axis=1)
print("launch-to-launch period (first entry), ns:", np.diff(first)[:n - 1].tolist())
print("gap last exit -> next first entry, ns:", (first[1:n] - last_exit[:n - 1]).tolist())
persist = bool(int(os.environ.get("B2_DENSE_PERSIST", "0")))
if persist:  # one launch: no per-step entry stamps; refer every trip to its first "arrived" stamp, ignore stale slots
    for jj in (5, 10):
        ref = ts[jj, :, 6].min()
        def stat(k):
            v = ts[jj, :, k] - ref
            v = v[(v >= 0) & (v < 100000)]
            return f"{int(v.mean())}/{int(v.max())}" if len(v) else "-"
        print(f"trip {jj} (ns after the first warp-0 arrival): " + "  ".join(f"{names[k]} {stat(k)}" for k in range(2, 15)))
    arr = np.array([ts[jj, :, 6].min() for jj in range(2, n - 1)])
    print("trip-to-trip period (first arrival), ns:", np.diff(arr).tolist())
    sys.exit(0)
for jj in (5, 10):
    t0 = first[jj]
    print(f"launch {jj}: " + "  ".join(f"{names[k]} {int((ts[jj, :, k] - t0).mean())}/{int((ts[jj, :, k] - t0).max())}" for k in range(15)))
for jj in (1, 2, 10):
  t0 = first[jj]
  print("launch", jj, "slowest by last:spec", [(int(b), int(ts[jj, b, 12] - t0)) for b in np.argsort(-ts[jj, :, 12])[:8]], "median", int(np.median(ts[jj, :, 12] - t0)))
jj = 10
t0 = first[jj]
order = np.argsort(-(ts[jj, :, 4]))[:6]
print("slowest CTAs of launch 10 (by 'searched'):")
for b in order:
    print(f"  cta {int(b):3d}: " + " ".join(f"{names[k]}={int(ts[jj, b, k] - t0)}" for k in range(15)))
print("fastest:")
for b in np.argsort(ts[jj, :, 4])[:3]:
    print(f"  cta {int(b):3d}: " + " ".join(f"{names[k]}={int(ts[jj, b, k] - t0)}" for k in range(15)))
