"""Synthetic iPhone-LiDAR-shaped scans (SURVEY.md §8d "Synthetic scan").

The reference's sensor front-end delivers a 256x192 depth image as records
``[x_px, y_px, depth_m]`` (lidar_ios_app/lidar/ARDepthViewController.swift:98-156)
which the plugin base class back-projects with the rescaled intrinsics of
src/slam/config/lidar_config.yaml (generic_point_cloud_processor.py:82-87,93-108).
This module ray-casts axis-aligned box scenes through that pinhole model.  It
is written against torch so the same code renders on the CPU (seeded tests) and
on the GPU (bench-sized workloads); it is data plumbing, not part of the hot path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

# src/slam/config/lidar_config.yaml:1-9
WIDTH, HEIGHT = 256, 192
REF_WIDTH, REF_HEIGHT = 1920, 1440
_FX = _FY = 1342.9849
_CX, _CY = 965.6142, 717.95435


def intrinsics():
    """fx, fy, cx, cy after the rescale at generic_point_cloud_processor.py:82-87."""
    sx, sy = WIDTH / REF_WIDTH, HEIGHT / REF_HEIGHT
    return _FX * sx, _FY * sy, _CX * sx, _CY * sy


@dataclass
class Scene:
    """A hollow axis-aligned room [0,lx]x[0,ly]x[0,lz] (z up) holding solid boxes."""
    size: tuple
    boxes: np.ndarray  # [K,6] = (xmin,ymin,zmin,xmax,ymax,zmax)


def make_room_scene(seed: int = 0, size=(4.0, 3.0, 2.6), n_boxes: int = 6) -> Scene:
    rng = np.random.default_rng(seed)
    lx, ly, lz = size
    boxes = []
    for _ in range(n_boxes):
        w = rng.uniform(0.3, 0.9, size=3)
        w[2] = rng.uniform(0.3, 1.4)
        # furniture hugs the far walls so the room centre stays free for the camera
        if rng.random() < 0.5:
            x0 = rng.uniform(0.0, lx - w[0])
            y0 = ly - w[1] - rng.uniform(0.0, 0.1)
        else:
            x0 = lx - w[0] - rng.uniform(0.0, 0.1)
            y0 = rng.uniform(0.0, ly - w[1])
        boxes.append([x0, y0, 0.0, x0 + w[0], y0 + w[1], w[2]])
    return Scene(size=(lx, ly, lz), boxes=np.asarray(boxes, dtype=np.float64).reshape(-1, 6))


def make_corridor_scene(seed: int = 0, length: float = 60.0, width: float = 3.0, height: float = 2.6,
                        boxes_per_m: float = 0.8) -> Scene:
    """Long corridor along +x with pillars/crates on the y=width wall and floor clutter."""
    rng = np.random.default_rng(seed)
    n = max(1, int(length * boxes_per_m))
    xs = np.sort(rng.uniform(0.2, length - 1.0, size=n))
    boxes = []
    for x0 in xs:
        w = rng.uniform(0.2, 0.8, size=3)
        w[2] = rng.uniform(0.3, height)
        y0 = width - w[1]
        boxes.append([x0, y0, 0.0, x0 + w[0], width, w[2]])
    # pilasters on the viewed wall and low clutter on the floor: without them a corridor leaves the
    # along-axis translation unobservable and point-to-point ICP slides (the classic degeneracy)
    x = 0.3
    while x < length - 0.5:
        d = rng.uniform(0.06, 0.16)
        boxes.append([x, width - d, 0.0, x + rng.uniform(0.08, 0.2), width, height])
        if rng.random() < 0.6:
            s = rng.uniform(0.15, 0.4, size=3)
            yc = rng.uniform(1.6, width - 0.5)
            boxes.append([x + 0.3, yc, 0.0, x + 0.3 + s[0], yc + s[1], s[2]])
        x += rng.uniform(0.45, 0.9)
    return Scene(size=(length, width, height), boxes=np.asarray(boxes, dtype=np.float64).reshape(-1, 6))


def look_at_pose(eye, yaw: float, pitch: float = 0.0, roll: float = 0.0) -> np.ndarray:
    """Camera->world 4x4.  Camera axes: x right, y down, z forward.  World z is up.
    yaw rotates the forward axis in the world xy-plane (0 = +x), pitch>0 looks up."""
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    fwd = np.array([cy * cp, sy * cp, sp])
    right = np.array([sy, -cy, 0.0])
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], axis=1)
    cr, sr = math.cos(roll), math.sin(roll)
    Rr = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1.0]])
    T = np.eye(4)
    T[:3, :3] = R @ Rr
    T[:3, 3] = np.asarray(eye, dtype=np.float64)
    return T


def _ray_dirs(device, dtype):
    fx, fy, cx, cy = intrinsics()
    xs = torch.arange(WIDTH, device=device, dtype=dtype)
    ys = torch.arange(HEIGHT, device=device, dtype=dtype)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    d = torch.stack([(xx - cx) / fx, (yy - cy) / fy, torch.ones_like(xx)], dim=-1)
    return d.reshape(-1, 3), xx.reshape(-1), yy.reshape(-1)


def render_depth(scene: Scene, pose_wc: np.ndarray, noise_sigma: float = 0.003, seed: int = 0,
                 device="cpu", cull_radius: float = 8.0) -> torch.Tensor:
    """Depth image [HEIGHT*WIDTH] float64 (z along the optical axis, metres); inf where nothing is hit."""
    dt = torch.float64
    dev = torch.device(device)
    dirs_c, _, _ = _ray_dirs(dev, dt)
    T = torch.as_tensor(pose_wc, dtype=dt, device=dev)
    o = T[:3, 3]
    d = dirs_c @ T[:3, :3].T  # world directions, un-normalised: depth == ray parameter
    inv = 1.0 / torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
    lo = torch.zeros(3, dtype=dt, device=dev)
    hi = torch.as_tensor(scene.size, dtype=dt, device=dev)
    # room: the camera is inside, the wall hit is where the ray leaves the box
    t1 = (lo - o) * inv
    t2 = (hi - o) * inv
    depth = torch.maximum(t1, t2).min(dim=1).values
    boxes = scene.boxes
    if len(boxes):
        c = 0.5 * (boxes[:, :3] + boxes[:, 3:])
        near = np.linalg.norm(c[:, :2] - np.asarray(pose_wc)[:2, 3], axis=1) < cull_radius
        b = torch.as_tensor(boxes[near], dtype=dt, device=dev)
        for k in range(0, b.shape[0], 64):
            bb = b[k:k + 64]
            ta = (bb[None, :, :3] - o) * inv[:, None, :]
            tb = (bb[None, :, 3:] - o) * inv[:, None, :]
            tn = torch.minimum(ta, tb).max(dim=2).values
            tf = torch.maximum(ta, tb).min(dim=2).values
            hit = (tn <= tf) & (tn > 0)
            tn = torch.where(hit, tn, torch.full_like(tn, float("inf")))
            depth = torch.minimum(depth, tn.min(dim=1).values)
    if noise_sigma > 0:
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed))
        depth = depth + (torch.randn(depth.shape, generator=g, dtype=dt) * noise_sigma).to(dev)
    return depth


def scan_records(depth: torch.Tensor, min_depth: float = 0.3, max_depth: float = 5.0,
                 keep_prob: float = 1.0, seed: int = 0) -> torch.Tensor:
    """Wire-format records [N,3] float32 = (x_px, y_px, depth_m) for valid pixels
    (optionally Bernoulli-dropped like ARDepthViewController.swift:112)."""
    dev = depth.device
    _, xx, yy = _ray_dirs(dev, torch.float32)
    valid = (depth >= min_depth) & (depth <= max_depth)
    if keep_prob < 1.0:
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed) + 7919)
        valid &= (torch.rand(depth.shape, generator=g) < keep_prob).to(dev)
    rec = torch.stack([xx, yy, depth.to(torch.float32)], dim=1)
    return rec[valid].contiguous()


def project_records(rec: torch.Tensor, mode: str = "f64") -> torch.Tensor:
    """Vectorised statement of ProcessPointClouds.project_pixel_to_3d
    (generic_point_cloud_processor.py:93-108): float32 [N,3] metric points.
    mode 'f64': NumPy-1 scalar promotion (float32 record, Python-float intrinsics ->
    float64 arithmetic, one rounding on store); mode 'f32': NumPy-2 (NEP 50) weak
    scalars -> float32 arithmetic throughout."""
    fx, fy, cx, cy = intrinsics()
    if mode == "f64":
        r = rec.to(torch.float64)
        x = (r[:, 0] - cx) * r[:, 2] / fx
        y = (r[:, 1] - cy) * r[:, 2] / fy
        return torch.stack([x, y, r[:, 2]], dim=1).to(torch.float32)
    if mode == "f32":
        r = rec.to(torch.float32)
        f = lambda v: torch.tensor(v, dtype=torch.float32, device=rec.device)  # noqa: E731
        x = (r[:, 0] - f(cx)) * r[:, 2] / f(fx)
        y = (r[:, 1] - f(cy)) * r[:, 2] / f(fy)
        return torch.stack([x, y, r[:, 2]], dim=1)
    raise ValueError(mode)


def render_scan(scene: Scene, pose_wc: np.ndarray, seed: int = 0, noise_sigma: float = 0.003,
                device="cpu", keep_prob: float = 1.0, max_depth: float = 5.0) -> torch.Tensor:
    """Plugin-boundary scan: float32 [N,3] metric points in the camera frame."""
    depth = render_depth(scene, pose_wc, noise_sigma=noise_sigma, seed=seed, device=device)
    rec = scan_records(depth, keep_prob=keep_prob, seed=seed, max_depth=max_depth)
    return project_records(rec)


def perturb_pose(pose_wc: np.ndarray, seed: int, trans: float = 0.04, rot_deg: float = 1.5) -> np.ndarray:
    """A nearby camera pose: |t| = trans metres, rotation rot_deg about a random axis."""
    rng = np.random.default_rng(seed)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = math.radians(rot_deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)
    t = rng.normal(size=3)
    t *= trans / np.linalg.norm(t)
    D = np.eye(4)
    D[:3, :3] = R
    D[:3, 3] = t
    return pose_wc @ D


def make_pair(seed: int = 0, device="cpu"):
    """Config-1-shaped pair: (source scan, target scan, T_true source->target frame)."""
    scene = make_room_scene(seed)
    rng = np.random.default_rng(seed + 1000)
    eye = [rng.uniform(0.8, 1.4), rng.uniform(0.6, 1.0), rng.uniform(1.1, 1.5)]
    pose_t = look_at_pose(eye, yaw=rng.uniform(0.55, 0.95), pitch=rng.uniform(-0.25, -0.05),
                          roll=rng.uniform(-0.05, 0.05))
    pose_s = perturb_pose(pose_t, seed + 2000, trans=rng.uniform(0.03, 0.05), rot_deg=rng.uniform(1.0, 2.0))
    tgt = render_scan(scene, pose_t, seed=2 * seed, device=device)
    src = render_scan(scene, pose_s, seed=2 * seed + 1, device=device)
    T_true = np.linalg.inv(pose_t) @ pose_s
    return src, tgt, T_true


def corridor_trajectory(scene: Scene, n_frames: int, start_x: float = 1.5, step: float = 0.025,
                        seed: int = 0):
    """Smooth camera path along a corridor, facing the cluttered wall obliquely.
    Inter-frame motion stays inside the reference's 5 cm correspondence gate
    (<= ~3 cm / 1 degree per frame)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * math.pi, size=4)
    poses = []
    for k in range(n_frames):
        x = start_x + step * k
        y = 1.0 + 0.15 * math.sin(0.05 * k + ph[0])
        z = 1.3 + 0.05 * math.sin(0.07 * k + ph[1])
        yaw = math.pi / 2 + 0.25 * math.sin(0.03 * k + ph[2])
        pitch = -0.12 + 0.05 * math.sin(0.04 * k + ph[3])
        poses.append(look_at_pose([x, y, z], yaw=yaw, pitch=pitch))
    return poses


def sample_scene_surface(scene: Scene, spacing: float, x_range, seed: int = 0, noise_sigma: float = 0.003,
                         device="cpu") -> torch.Tensor:
    """Dense jittered samples of the room walls / floor / ceiling and the visible box
    faces inside x_range: bulk content for pre-loading a large resident map."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    x0, x1 = x_range
    lx, ly, lz = scene.size
    parts = []

    def rect(fixed_axis, fixed_val, a_rng, b_rng):
        na = max(1, int((a_rng[1] - a_rng[0]) / spacing))
        nb = max(1, int((b_rng[1] - b_rng[0]) / spacing))
        a = torch.linspace(a_rng[0], a_rng[1], na, dtype=torch.float64)
        b = torch.linspace(b_rng[0], b_rng[1], nb, dtype=torch.float64)
        aa, bb = torch.meshgrid(a, b, indexing="ij")
        aa = aa.reshape(-1) + (torch.rand(aa.numel(), generator=g, dtype=torch.float64) - 0.5) * spacing
        bb = bb.reshape(-1) + (torch.rand(bb.numel(), generator=g, dtype=torch.float64) - 0.5) * spacing
        ff = torch.full_like(aa, fixed_val) + torch.randn(aa.numel(), generator=g, dtype=torch.float64) * noise_sigma
        cols = [None, None, None]
        cols[fixed_axis] = ff
        rest = [i for i in range(3) if i != fixed_axis]
        cols[rest[0]], cols[rest[1]] = aa, bb
        parts.append(torch.stack(cols, dim=1))

    rect(2, 0.0, (x0, x1), (0.0, ly))   # floor
    rect(2, lz, (x0, x1), (0.0, ly))    # ceiling
    rect(1, 0.0, (x0, x1), (0.0, lz))   # wall y=0
    rect(1, ly, (x0, x1), (0.0, lz))    # wall y=ly
    for bx in scene.boxes:
        if bx[3] < x0 or bx[0] > x1:
            continue
        rect(1, bx[1], (bx[0], bx[3]), (bx[2], bx[5]))  # face towards the corridor
        rect(0, bx[0], (bx[1], bx[4]), (bx[2], bx[5]))
        rect(0, bx[3], (bx[1], bx[4]), (bx[2], bx[5]))
        if bx[5] < lz - 1e-6:
            rect(2, bx[5], (bx[0], bx[3]), (bx[1], bx[4]))
    return torch.cat(parts, dim=0).to(torch.float32).to(device)
