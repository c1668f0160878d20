#!/usr/bin/env python
"""bench.py — scans/s of ICP + fuse (49k-pt scan -> 2 M-voxel resident map), BASELINE.json config 2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: a run of
`--scans-per-step` consecutive 256x192 scans (49,152 points each) along a smooth corridor
trajectory, each voxel-down-sampled (2 cm), registered by point-to-point ICP (<= 30 iterations,
1e-6/1e-6 criteria, 5 cm gate) against the GPU-resident 5 cm voxel map (~2 M voxels) seeded with
the previous pose, then fused into that map.  Scans are sequentially dependent, so on N > 1 GPUs
every rank replays its own trajectory against its own replica of the map (weak scaling, no
data-path collective — SURVEY.md §8e); the secondary `batched_pairs` object reports the sharded
independent-pair workload (config 5 shape) with its NCCL result gather.

value  : device-timed scans/s, scan batches already resident in HBM (CUDA events on the
         library's stream, max over ranks).
e2e    : the same through the reference-facing plugin call
         (B200ICPProcessor.construct_global_map) with HOST (pinned) scans: H2D copy of every
         scan and D2H read-back of every pose inside the timed region.
--impl reference : the reference's own CPU sequence (re-voxelise + re-index the whole map per
         scan, icp_processor.py:43-119) restated on the CPU oracle — Open3D itself is not
         installable here — on a bounded sample, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAP_VOXEL = 0.05
DS_VOXEL = 0.02
MAX_CORR = 0.05
MAX_ITER = 30
MAP_ORIGIN = (-50.0, -50.0, -50.0)
CORRIDOR_LEN = 210.0          # ~2.1 M voxels at 5 cm
# voxel budget of the resident map (B2_ERR_CAPACITY beyond 70 %): 2^23 -> a pool of 131,072 bricks of 8^3 voxels
# (16-byte centroid cell + 32-byte fp64 accumulator per voxel: 3.1 GiB of the 180 GB HBM; ~27 k bricks are used)
MAP_CAPACITY = int(os.environ.get("B2_BENCH_CAP", 1 << 23))
SCAN_POINTS = 49152
METRIC = "scans/s ICP+fuse (49k-pt scan->2M-pt map)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans-per-step", type=int, default=64)
    ap.add_argument("--pairs", type=int, default=512, help="independent pairs per rank for the batched_pairs leg")
    ap.add_argument("--cpu-scans", type=int, default=10, help="scans timed by the CPU baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------- workload
def build_scene(rank: int):
    from lidar_slam_b200 import synth

    return synth.make_corridor_scene(1000 + rank, length=CORRIDOR_LEN)


def preload_map(engine, scene, device):
    """Fill the resident map with ~2 M voxels by fusing dense surface samples of the corridor."""
    import torch
    from lidar_slam_b200 import synth

    engine.map_init(MAP_VOXEL, MAP_ORIGIN, MAX_CORR, MAP_CAPACITY)
    for x0 in range(0, int(CORRIDOR_LEN), 10):
        pts = synth.sample_scene_surface(scene, 0.02, (float(x0), float(x0 + 10)), seed=x0, device=device)
        p4 = torch.zeros((pts.shape[0], 4), dtype=torch.float32, device=device)
        p4[:, :3] = pts
        engine.map_fuse(p4)
    engine.synchronize()
    return engine.map_size()


def make_scans(scene, n_scans: int, device, seed: int, start_x: float = 20.0):
    """Consecutive scans along a smooth trajectory (<= ~2.5 cm / 1 deg per frame) + their true poses."""
    from lidar_slam_b200 import synth

    poses = synth.corridor_trajectory(scene, n_scans, start_x=start_x, step=0.02, seed=seed)
    scans = [synth.render_scan(scene, p, seed=seed * 100003 + k, device=device).cpu().numpy()
             for k, p in enumerate(poses)]
    assert all(len(s) == SCAN_POINTS for s in scans), "every pixel of the synthetic scans must be valid"
    return scans, poses


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import lidar_slam_b200 as b2
    from lidar_slam_b200 import batch as b2batch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = b2.Engine(local_rank)
    scene = build_scene(rank)
    map_voxels = preload_map(eng, scene, device)
    S = args.scans_per_step
    # consecutive steps replay DIFFERENT stretches of the trajectory (up to 8 segments of S scans =
    # 400 MB of device-resident / 300 MB of pinned inputs, > the 126 MB L2), so no step finds its
    # scans or its part of the map warm in L2 from the step before
    n_seg = max(1, min(8, args.warmup + args.steps))
    scans, poses = make_scans(scene, S * n_seg, device, seed=rank + 1)
    prm = b2.make_params(MAX_CORR, MAX_ITER)

    # device-resident copies (for `value`) and pinned host copies (for `e2e`)
    dev_scans = [eng.pack(s) for s in scans]
    pinned = [torch.from_numpy(s).pin_memory() for s in scans]
    pinned_np = [p.numpy() for p in pinned]
    h2d_per_step = int(sum(p.numel() * 4 for p in pinned[:S]))
    d2h_per_step = S * 152
    step_no = [0, 0]

    def step_device():
        """One S-scan stretch with inputs resident in HBM: down-sample -> ICP (the pose chain stays on
        the device) -> fuse, enqueued asynchronously, no host read-back inside."""
        b = (step_no[0] % n_seg) * S
        step_no[0] += 1
        eng.set_pose(poses[b])
        for k in range(b, b + S):
            eng.slam_scan_dev(dev_scans[k], DS_VOXEL, prm)

    def step_e2e(per_scan_sync=False):
        """The same stretch through the C-ABI call the plugin makes (b2_slam_scan): HOST scans in (pinned),
        pose + fitness + rmse of EVERY scan read back to the host inside the step — by default once per step
        (b2_slam_results, what plugin.ProcessLoop does for a batch of pending scans); per_scan_sync=True waits
        for each scan's result before submitting the next (what a caller with one scan at a time sees)."""
        b = (step_no[1] % n_seg) * S
        step_no[1] += 1
        eng.set_pose(poses[b])
        if per_scan_sync:
            out = [eng.slam_scan(pinned_np[k], DS_VOXEL, prm) for k in range(b, b + S)]
        else:
            for k in range(b, b + S):
                eng.slam_scan(pinned_np[k], DS_VOXEL, prm, want_result=False)
            out = eng.slam_results(S)
        return out, poses[b + S - 1]

    launches0 = eng.launch_count()
    # ---- device-timed value
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = eng.launch_count()
    with ClockSampler(local_rank) as clk:
        with torch.cuda.stream(eng.stream):
            ev0.record()
            for _ in range(args.steps):
                step_device()
            ev1.record()
        eng.synchronize()
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    gpu_launches = eng.launch_count() - launches_before
    ms_per_step = ms_total / args.steps
    value = world * S * args.steps / (ms_total / 1e3)

    # ---- e2e through the plugin-facing call with host buffers
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(eng.stream):
        ev0.record()
        last, last_pose = None, None
        for _ in range(args.steps):
            last, last_pose = step_e2e()
        ev1.record()
    eng.synchronize()
    barrier()
    e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))
    e2e_value = world * S * args.steps / (e2e_ms / 1e3)
    # the same with a host wait after every scan (secondary figure)
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 2)):
        step_e2e(per_scan_sync=True)
    eng.synchronize()
    barrier()
    e2e_sync_value = world * S * max(1, args.steps // 2) / max_over_ranks(time.perf_counter() - t0)
    fit = float(np.mean([r.fitness for r in last[1:]]))
    iters = float(np.mean([r.iterations for r in last[1:]]))
    # pose accuracy of the replay against the synthetic ground truth (sanity, not a metric)
    err = float(np.linalg.norm(last[-1].matrix()[:3, 3] - last_pose[:3, 3]))

    # ---- roofline of the dominant kernel (ICP iteration) measured live
    roof = roofline_icp(eng, dev_scans, poses, prm, torch)

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "scans/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config[1]: scan->map ICP + voxel fusion, 49,152-pt scan into a ~2M-voxel resident "
                               "map, 5 cm voxels", "scans_per_step": S, "distinct_scans": S * n_seg, "scan_points": SCAN_POINTS,
                   "map_voxels": int(map_voxels), "map_voxel_m": MAP_VOXEL, "ds_voxel_m": DS_VOXEL,
                   "max_corr_m": MAX_CORR, "max_iteration": MAX_ITER, "mean_icp_iterations": round(iters, 2),
                   "mean_fitness": round(fit, 4), "final_pose_error_m": round(err, 4),
                   "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                   "map_hash_slots": MAP_CAPACITY,
                   "l2": "inputs larger than L2: %d MB of distinct device-resident scans cycled over the steps; map tables 2 GiB" % (S * n_seg * SCAN_POINTS * 16 // 1000000)},
        "e2e": {"value": round(e2e_value, 2), "unit": "scans/s", "readback": "every scan's 152-byte result, "
                "collected once per step (b2_slam_results)", "per_scan_wait_value": round(e2e_sync_value, 2),
                "h2d_bytes_per_step": h2d_per_step,
                "d2h_bytes_per_step": d2h_per_step},
        "gpu_launches": int(gpu_launches),
        "clocks": clk.summary(),
        "roofline": roof,
    }

    if not args.no_batched:
        line["batched_pairs"] = run_batched(eng, args, rank, world, device, barrier, max_over_ranks, b2, b2batch)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, scene_seed=1000, n_scans=args.cpu_scans)
    if rank == 0:
        print(json.dumps(line), flush=True)
    del launches0
    if world > 1:
        dist.destroy_process_group()


def roofline_icp(eng, dev_scans, poses, prm, torch):
    """Average duration of one ICP-iteration launch against the resident map, timed with CUDA events on
    the library stream, versus its algorithmic bytes: per query 16 B (source
    point) + 27 x 16 B (cell records of the stencil) + 16 B per candidate point; + 256 B per CTA."""
    import json as _json

    sd, _ = eng.voxel_downsample(dev_scans[1], DS_VOXEL, want_keys=False)
    ns = sd.shape[0]
    # candidates under the stencil: measured by the library's own search-only pass (counts returned in d2 slot)
    n_launch = 31
    reps = 20
    one = __import__("lidar_slam_b200").make_params(MAX_CORR, n_launch - 1, rel_fitness=0.0, rel_rmse=0.0)
    init = poses[1]
    for _ in range(3):
        eng.map_icp(sd, one, init=init)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.stream):
        ev0.record()
        for _ in range(reps):
            eng.map_icp(sd, one, init=init)
        ev1.record()
    eng.synchronize()
    us_per_launch = ev0.elapsed_time(ev1) * 1e3 / (reps * n_launch)
    cand = eng.stencil_candidates(sd, init)
    bytes_per_launch = 16 * ns + 27 * 16 * ns + 16 * cand + 256 * 296
    achieved = bytes_per_launch / (us_per_launch * 1e-6) / 1e9
    peaks = {}
    try:
        peaks = _json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return {"bound": "hbm", "kernel": "icp_iter_kernel<p2p,map>", "achieved": round(achieved, 1), "peak": peak,
            "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
            "unit": "GB/s", "frac": round(achieved / peak, 4),
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (cold caches between
            # replays): profiles/r01c_icp_iter_full_raw.csv
            "traffic": 2500352,
            "us_per_launch": round(us_per_launch, 3), "queries": int(ns), "candidates": int(cand),
            "algorithmic_bytes_per_launch": int(bytes_per_launch),
            "note": "time per launch includes the gaps between the 31 chained launches of a registration; DRAM traffic "
                    "(2.5 MB cold) is far below the algorithmic bytes because the 27-cell reads of neighbouring "
                    "queries overlap in L1/L2: the kernel is bound by dependent-access latency, not bandwidth"}


def run_batched(eng, args, rank, world, device, barrier, max_over_ranks, b2, b2batch):
    """Config-5-shaped leg: independent pairs sharded in contiguous blocks over the ranks, one NCCL
    all-gather of the 152-byte result records at the end."""
    import torch
    from lidar_slam_b200 import synth

    P = args.pairs
    n_total = P * world
    b, e = b2batch.shard_range(n_total, rank, world)
    srcs, tgts = [], []
    distinct = min(P, 16)  # synthesise a few distinct pairs per rank and tile them (inputs only)
    for i in range(distinct):
        s, t, _ = synth.make_pair(b + i, device=device)
        srcs.append(s)
        tgts.append(t)
    reps = -(-P // distinct)
    src = torch.cat((srcs * reps)[:P]).to(device)
    tgt = torch.cat((tgts * reps)[:P]).to(device)
    so = torch.arange(0, P + 1, dtype=torch.int32, device=device) * SCAN_POINTS
    S4, T4 = eng.pack(src), eng.pack(tgt)
    prm = b2.make_params(MAX_CORR, MAX_ITER)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def go():
        res = eng.icp_batch(S4, so, T4, so, prm, ds_voxel=DS_VOXEL)
        with torch.cuda.stream(eng.stream):
            return b2batch.gather_results(res, n_total)

    go()
    barrier()
    with torch.cuda.stream(eng.stream):
        ev0.record()
        allres = go()
        ev1.record()
    eng.synchronize()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    rec = b2batch.records_view(allres)
    out = {"metric": "batched pair-ICP/s", "value": round(n_total / (ms / 1e3), 2), "unit": "pairs/s",
           "pairs_total": n_total, "pairs_per_rank": P, "ms": round(ms, 3), "gathered_records": int(len(rec)),
           "mean_fitness": round(float(rec["fitness"].mean()), 4),
           "collective": "all_gather of 152 B/pair result records" if world > 1 else "none (1 rank)"}
    if rank == 0:
        # roofline of the batched ICP step: (time with 30 steps - time with 0 steps) / 30 per launch, against
        # the algorithmic bytes of one launch over all pairs of this rank (census on the distinct pairs)
        prm0 = b2.make_params(MAX_CORR, 0)
        eng.icp_batch(S4, so, T4, so, prm0, ds_voxel=DS_VOXEL)
        with torch.cuda.stream(eng.stream):
            ev0.record()
            eng.icp_batch(S4, so, T4, so, prm0, ds_voxel=DS_VOXEL)
            ev1.record()
        eng.synchronize()
        ms0 = ev0.elapsed_time(ev1)
        with torch.cuda.stream(eng.stream):
            ev0.record()
            eng.icp_batch(S4, so, T4, so, prm, ds_voxel=DS_VOXEL)
            ev1.record()
        eng.synchronize()
        us_launch = (ev0.elapsed_time(ev1) - ms0) * 1e3 / MAX_ITER
        q = c = 0
        for i in range(distinct):
            sd, _ = eng.voxel_downsample(eng.pack(srcs[i]), DS_VOXEL, want_keys=False)
            td, _ = eng.voxel_downsample(eng.pack(tgts[i]), DS_VOXEL, want_keys=False)
            q += sd.shape[0]
            c += eng.stencil_candidates(sd, None, tgt=td, cell=MAX_CORR)
        scale = P / distinct
        bytes_launch = int(scale * (16 * q + 27 * 16 * q + 16 * c))
        peak = 6545.3
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        ach = bytes_launch / (us_launch * 1e-6) / 1e9
        # DRAM bytes per pair and launch from the ncu capture of this kernel (profiles/r01c_sweep_raw.csv:
        # dram__bytes_read.sum + dram__bytes_write.sum = 352.3 + 3.7 MB for a 148-pair launch)
        traffic = int(356.0e6 / 148 * P)
        out["roofline"] = {"bound": "hbm", "kernel": "icp_sweep_kernel<p2p,pair> (gridDim.y = pairs)",
                           "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                           "traffic": traffic, "us_per_launch": round(us_launch, 1), "queries": int(scale * q),
                           "candidates": int(scale * c), "algorithmic_bytes_per_launch": bytes_launch,
                           "note": "algorithmic bytes = SURVEY 8(d) model over the FULL 27-cell stencil; the kernel "
                                   "prunes the walk (about 1/3 of those candidates are read) and L1/L2 serve the reuse, "
                                   "so the figure can exceed the HBM peak: the kernel is bound by instruction issue "
                                   "(ncu: 55 % issue-active, 20/32 threads per instruction), DRAM traffic is ~0.7 TB/s"}
    return out


# ----------------------------------------------------------------------------- CPU legs
def cpu_reference_run(n_scans: int, scene_seed: int, map_points_target: int = 2_000_000):
    """The reference-faithful CPU sequence on the oracle: raw-point map (~2 M points), and for every
    scan: re-voxelise scan and WHOLE map at 2 cm, index the map, ICP (<= 30 it.), transform, prepend."""
    import torch
    from lidar_slam_b200 import synth
    from oracle import cpu as orc
    from oracle import pipeline

    # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun exports 1)
    try:
        host_threads = len(os.sched_getaffinity(0))
    except AttributeError:
        host_threads = os.cpu_count() or 1
    orc.set_num_threads(host_threads)
    scene = synth.make_corridor_scene(scene_seed, length=CORRIDOR_LEN)
    # raw map of ~2 M points around the trajectory start (the reference's map is raw points, not voxels)
    span = 28.0
    pts = synth.sample_scene_surface(scene, 0.013, (20.0 - 4.0, 20.0 - 4.0 + span), seed=1).numpy()
    if len(pts) > map_points_target:
        pts = pts[np.random.default_rng(0).choice(len(pts), map_points_target, replace=False)]
    scans, poses = make_scans(scene, n_scans + 1, "cpu", seed=1)
    seq = pipeline.ReferenceSequence(max_iteration=MAX_ITER, maintenance="off")
    seq.global_map = pts.astype(np.float64)
    seq.point_clouds_in_map = 2
    seq.previous_transformation = [poses[0]]
    seq.process(scans[0])  # warm-up scan (page-in, thread pool)
    t0 = time.perf_counter()
    for s in scans[1:]:
        seq.process(s)
    dt = time.perf_counter() - t0
    del torch
    return n_scans / dt, orc.num_threads(), len(pts), float(seq.last_result.fitness)


def cpu_baseline(args, scene_seed: int, n_scans: int):
    v, cores, n_map, fit = cpu_reference_run(n_scans, scene_seed)
    return {"value": round(v, 4), "unit": "scans/s", "cores": cores, "kind": "port",
            "sample": f"{n_scans} consecutive 49,152-pt scans against a {n_map}-pt raw map, reference-faithful "
                      f"sequence (re-voxelise + re-index whole map per scan), oracle C++/OpenMP; last fitness {fit:.3f}",
            "note": "Open3D itself is not installable here; CPU baseline = restated oracle"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(1, args.cpu_scans)
    vals = []
    cores = n_map = 0
    for _ in range(max(1, min(args.steps, 3))):
        v, cores, n_map, _ = cpu_reference_run(n, 1000)
        vals.append(v)
    v = float(np.median(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "scans/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * n / v, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config[1]: scan->map ICP + voxel fusion, 49,152-pt scan into a ~2M-pt map "
                               "(reference-faithful CPU sequence)", "scan_points": SCAN_POINTS, "map_points": n_map,
                   "max_iteration": MAX_ITER},
        "cpu_baseline": {"value": round(v, 4), "unit": "scans/s", "cores": cores, "kind": "port",
                         "sample": f"{n} scans per step against a {n_map}-pt raw map"},
        "e2e": {"value": round(v, 4), "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference arm = oracle port of the reference's Open3D sequence (Open3D not installable offline)",
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
