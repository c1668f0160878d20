"""Profiling driver: one batched pair-ICP call (config-5 shape), nothing else.  Run under ncu."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lidar_slam_b200 as b2  # noqa: E402
from lidar_slam_b200 import batch, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=148)
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--points", type=int, default=49152)
ap.add_argument("--p2l", action="store_true", help="point-to-plane batch (normals of every target estimated first)")
args = ap.parse_args()
dev = torch.device("cuda:0")
eng = b2.Engine(0)
srcs, tgts = [], []
for i in range(min(args.pairs, 8)):
    s, t, _ = synth.make_pair(i, device=dev)
    srcs.append(s)
    tgts.append(t)
reps = -(-args.pairs // len(srcs))
src = torch.cat((srcs * reps)[: args.pairs]).to(dev)
tgt = torch.cat((tgts * reps)[: args.pairs]).to(dev)
so = torch.arange(0, args.pairs + 1, dtype=torch.int32, device=dev) * args.points
S4, T4 = eng.pack(src), eng.pack(tgt)
prm = b2.make_params(0.05, args.iters, mode=b2.P2L if args.p2l else b2.P2P)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    with torch.cuda.stream(eng.stream):
        ev0.record()
        res = eng.icp_batch(S4, so, T4, so, prm, ds_voxel=0.02, normal_radius=0.04, normal_max_nn=20)
        ev1.record()
    eng.synchronize()
print("ms", ev0.elapsed_time(ev1), "fitness", float(batch.records_view(res)["fitness"].mean()))
