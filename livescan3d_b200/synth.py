"""Deterministic synthetic Kinect-v2-shaped inputs for the hot path.

Pure numpy, no GPU, no oracle: this only manufactures INPUTS (depth maps, colours, intrinsics, poses) in the
packed layouts the reference's P/Invoke boundary uses (KinectServer.cs:404-500 packs them; NativeUtils reads
them at depthprocessing.cpp:715-727,1646-1650):

  depth_maps   : S tightly packed little-endian u16 images, sensor i at byte offset sum_{j<i} 2*w_j*h_j
  depth_colors : S packed RGB u8 triples at sum_{j<i} 3*w_j*h_j
  intr         : 7 floats per sensor  cx, cy, fx, fy, r2, r4, r6      (depthprocessing.h:90-98)
  wt           : 12 floats per sensor t[3] then R row-major            (depthprocessing.h:50-63)

The scene (metres, world frame): floor y=-0.8, walls at |x|,|z|=2.5, a 0.4 m sphere at the origin and a 0.3 m
cube at (0.6,-0.5,0.2).  Sensor s of S sits on a ring of radius 1.75 m, height 0.3 m, yaw 360*s/S, looking at
the origin; its (R_w, t_w) satisfies p_world = R_w (p_cam + t_w), the order createVertices applies them
(depthprocessing.cpp:157-160).
"""
from __future__ import annotations

import numpy as np

KINECT_W, KINECT_H = 512, 424
_FX, _FY, _CX, _CY = 365.456, 365.456, 254.878, 205.395
_R2, _R4, _R6 = 0.0905474, -0.26819, 0.0950862

# the two real poses the reference ships (LiveScanClient/calibration.txt, bin/calibration.txt): t then R rows
FIXTURE_POSES = (
    np.array([0.0193386, 0.272806, -1.73244,
              -0.958759, -0.0636071, 0.277012,
              0.100432, 0.835942, 0.539551,
              -0.265885, 0.54512, -0.795078], dtype=np.float32),
    np.array([-0.755848, 0.191075, -1.71209,
              -0.988817, -0.0238505, 0.147212,
              0.0449593, 0.893526, 0.446755,
              -0.142193, 0.448378, -0.882462], dtype=np.float32),
)


def intrinsics(w: int = KINECT_W, h: int = KINECT_H) -> np.ndarray:
    s = w / float(KINECT_W)
    return np.array([_CX * s, _CY * s, _FX * s, _FY * s, _R2, _R4, _R6], dtype=np.float32)


def ring_pose(s: int, S: int, radius: float = 1.75, height: float = 0.3):
    """Camera-to-world rotation (columns = camera x,y,z axes in world) and camera centre, float64."""
    yaw = 2.0 * np.pi * s / S
    C = np.array([radius * np.sin(yaw), height, radius * np.cos(yaw)])
    f = -C / np.linalg.norm(C)                      # forward: look at the origin
    up0 = np.array([0.0, 1.0, 0.0])
    r = np.cross(up0, f); r /= np.linalg.norm(r)    # camera +X (sensor's left), right-handed (r, u, f)
    u = np.cross(f, r)
    R_cw = np.stack([r, u, f], axis=1)
    return R_cw, C


def pose_params(R_cw: np.ndarray, C: np.ndarray) -> np.ndarray:
    """12-float wtransform block: t_w = R_cw^T C, then R_w = R_cw row-major."""
    t_w = R_cw.T @ C
    return np.concatenate([t_w, R_cw.reshape(-1)]).astype(np.float32)


def _raycast(orig: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    """Smallest positive ray parameter to the scene, per ray (dirs are NOT normalised; camera z == 1)."""
    ox, oy, oz = orig
    dx, dy, dz = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    best = np.full(dirs.shape[0], np.inf)

    def plane(o, d, c):
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (c - o) / d
        t[~np.isfinite(t) | (t <= 1e-9)] = np.inf
        return t

    for t in (plane(oy, dy, -0.8), plane(ox, dx, 2.5), plane(ox, dx, -2.5), plane(oz, dz, 2.5), plane(oz, dz, -2.5)):
        best = np.minimum(best, t)
    # sphere r=0.4 at the origin
    a = dx * dx + dy * dy + dz * dz
    b = 2.0 * (ox * dx + oy * dy + oz * dz)
    c = ox * ox + oy * oy + oz * oz - 0.4 * 0.4
    disc = b * b - 4 * a * c
    ok = disc >= 0
    t = np.full_like(best, np.inf)
    t[ok] = (-b[ok] - np.sqrt(disc[ok])) / (2 * a[ok])
    t[t <= 1e-9] = np.inf
    best = np.minimum(best, t)
    # axis-aligned cube, side 0.3, centred (0.6,-0.5,0.2): slab test
    lo = np.array([0.6, -0.5, 0.2]) - 0.15
    hi = np.array([0.6, -0.5, 0.2]) + 0.15
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo[None, :] - orig[None, :]) / dirs
        t2 = (hi[None, :] - orig[None, :]) / dirs
    tn = np.nanmax(np.minimum(t1, t2), axis=1)
    tf = np.nanmin(np.maximum(t1, t2), axis=1)
    hit = (tn <= tf) & (tn > 1e-9)
    t = np.where(hit, tn, np.inf)
    return np.minimum(best, t)


def _hash32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x45D9F3B)) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x45D9F3B)) & np.uint64(0xFFFFFFFF)
    x = x ^ (x >> np.uint64(16))
    return x.astype(np.uint32)


def sensor_frame(s: int, S: int, w: int = KINECT_W, h: int = KINECT_H, seed_base: int = 1000, wt: np.ndarray | None = None):
    """One sensor's (depth u16 [h,w], rgb u8 [h,w,3], intr f32[7], wt f32[12])."""
    intr = intrinsics(w, h)
    if wt is None:
        R_cw, C = ring_pose(s, S)
        wt = pose_params(R_cw, C)
    else:
        wt = np.asarray(wt, dtype=np.float32)
        R_cw = wt[3:].astype(np.float64).reshape(3, 3)
        C = R_cw @ wt[:3].astype(np.float64)
    cx, cy, fx, fy = [float(v) for v in intr[:4]]
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d_cam = np.stack([(xs - cx) / fx, (cy - ys) / fy, np.ones_like(xs)], axis=-1).reshape(-1, 3)
    d_world = d_cam @ R_cw.T
    z = _raycast(C, d_world)                         # camera-space Z in metres (ray dir has z == 1)
    rng = np.random.RandomState(seed_base + s)
    mm = z * 1000.0 + rng.normal(0.0, 1.5, size=z.shape)
    fly = rng.random_sample(z.shape) < 0.005
    mm = mm + fly * rng.uniform(0.0, 300.0, size=z.shape)
    drop = rng.random_sample(z.shape) < 0.02
    mm = np.where(np.isfinite(mm), mm, 0.0)
    mm = np.where((mm < 500.0) | (mm > 8000.0) | drop, 0.0, mm)
    depth = np.rint(mm).astype(np.uint16).reshape(h, w)
    hv = _hash32(np.arange(w * h, dtype=np.uint32) ^ np.uint32(s))
    rgb = np.stack([hv & 0xFF, (hv >> 8) & 0xFF, (hv >> 16) & 0xFF], axis=-1).astype(np.uint8).reshape(h, w, 3)
    return depth, rgb, intr, wt


def make_frame(S: int, w: int = KINECT_W, h: int = KINECT_H, seed_base: int = 1000, poses=None, ring: int | None = None):
    """Packed multi-sensor frame in the P/Invoke layout.  `ring` = number of ring slots (defaults to S)."""
    depths, colors, intrs, wts = [], [], [], []
    for s in range(S):
        d, c, i, t = sensor_frame(s, ring or S, w, h, seed_base, None if poses is None else poses[s])
        depths.append(d.reshape(-1)); colors.append(c.reshape(-1)); intrs.append(i); wts.append(t)
    return {
        "n_maps": S,
        "depth_maps": np.ascontiguousarray(np.concatenate(depths)).view(np.uint8),
        "depth_colors": np.ascontiguousarray(np.concatenate(colors)),
        "widths": np.full(S, w, dtype=np.int32),
        "heights": np.full(S, h, dtype=np.int32),
        "intr": np.concatenate(intrs).astype(np.float32),
        "wt": np.concatenate(wts).astype(np.float32),
    }


DEFAULT_BOUNDS = np.array([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], dtype=np.float32)   # minX,minY,minZ,maxX,maxY,maxZ
SERVER_BOUNDS = np.array([-5, -5, -5, 5, 5, 5], dtype=np.float32)                # KinectSettings.cs:54-60
CLIENT_BOUNDS = np.array([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5], dtype=np.float32)    # liveScanClient.cpp:78-83


def rigid_offset(deg: float = 1.5, axis=(0.3, 1.0, 0.2), trans_mm=(8.0, -5.0, 6.0)):
    """The known ICP perturbation: rotation matrix (float64) and translation in metres."""
    a = np.asarray(axis, dtype=np.float64); a /= np.linalg.norm(a)
    th = np.deg2rad(deg)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    Rm = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    return Rm, np.asarray(trans_mm, dtype=np.float64) / 1000.0


def perturb(points: np.ndarray, deg: float = 1.5, axis=(0.3, 1.0, 0.2), trans_mm=(8.0, -5.0, 6.0)) -> np.ndarray:
    Rm, tv = rigid_offset(deg, axis, trans_mm)
    return (points.astype(np.float64) @ Rm.T + tv).astype(np.float32)
