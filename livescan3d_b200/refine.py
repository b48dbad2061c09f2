"""The calibration-refinement driver around ICP: LiveScanServer's refineWorker loop (MainWindowForm.cs:349-405) on top of the
library's ICP — the caller of the hot path on the server side, so that a whole "Refine calibration" click
can be reproduced and measured, not just one ICP call.

  for refineIter in range(nNumRefineIters):            # KinectSettings.cs: nNumRefineIters = 2
      for i in range(n_sensors):
          verts1 = all other sensors' (already refined) clouds, concatenated in sensor order
          ICP(verts1, verts2 = cloud i, ..., Rs[i], Ts[i], nNumICPIterations)      # cloud i moves in place; R, t accumulate

Two drivers with identical results: `refine_poses` goes through the reference's own export with host arrays (what the C#
code does), `refine_poses_device` keeps every cloud in HBM and only concatenates on the device.
"""
from __future__ import annotations

import numpy as np


def refine_poses(clouds, n_refine_iters: int = 2, n_icp_iters: int = 10):
    """clouds: list of [n_i, 3] float32 arrays.  Returns (refined clouds, Rs [S,3,3], Ts [S,3])."""
    from . import api
    clouds = [np.array(c, dtype=np.float32, order="C").reshape(-1, 3) for c in clouds]
    S = len(clouds)
    Rs = [np.eye(3, dtype=np.float32) for _ in range(S)]
    Ts = [np.zeros(3, dtype=np.float32) for _ in range(S)]
    for _ in range(int(n_refine_iters)):
        for i in range(S):
            others = [clouds[j] for j in range(S) if j != i]
            if not others or sum(len(o) for o in others) == 0 or len(clouds[i]) == 0:
                continue                                     # the reference would throw on an empty target (nanoflann.h:904)
            verts1 = np.ascontiguousarray(np.concatenate(others))
            clouds[i], Rs[i], Ts[i] = api.icp(verts1, clouds[i], Rs[i], Ts[i], n_icp_iters)
    return clouds, np.stack(Rs), np.stack(Ts)


def refine_poses_device(clouds, n_refine_iters: int = 2, n_icp_iters: int = 10):
    """clouds: list of CUDA float32 tensors [n_i, 3] (moved in place).  Returns (Rs [S,3,3], Ts [S,3]) as numpy.  Everything is
    enqueued on the current stream; the host only waits once per refine iteration, to hand each sensor's accumulated pose back
    as the next call's R, t (ICP() takes them by value)."""
    import torch
    from .device import IcpSolver
    S = len(clouds)
    n = [int(c.shape[0]) for c in clouds]
    solver = IcpSolver(max(1, sum(n) - (min(n) if n else 0)), max(1, max(n) if n else 1))
    Rs = np.stack([np.eye(3, dtype=np.float32) for _ in range(S)])
    Ts = np.zeros((S, 3), dtype=np.float32)
    poses = torch.zeros((S, 12), dtype=torch.float32, device=clouds[0].device)
    poses[:, 0] = poses[:, 4] = poses[:, 8] = 1.0
    keep_alive = []
    for it in range(int(n_refine_iters)):
        if it > 0:
            torch.cuda.current_stream().synchronize()
            rt = poses.cpu().numpy()
            Rs, Ts = rt[:, :9].reshape(S, 3, 3).copy(), rt[:, 9:].copy()
            keep_alive.clear()
        for i in range(S):
            others = [clouds[j] for j in range(S) if j != i and n[j] > 0]
            if not others or n[i] == 0:
                continue                                     # the reference would throw on an empty target (nanoflann.h:904)
            verts1 = torch.cat(others).contiguous()           # device-side concatenation in sensor order (plumbing)
            keep_alive.append(verts1)                         # the stream still reads it after this Python scope moves on
            solver.set_target(verts1)
            solver.set_source(clouds[i], 0, None, Rs[i], Ts[i])
            solver.run(n_icp_iters)
            poses[i].copy_(solver.Rt)
    torch.cuda.current_stream().synchronize()
    out = poses.cpu().numpy()
    solver.close()
    return out[:, :9].reshape(S, 3, 3).copy(), out[:, 9:].copy()


def update_calibration(world_R, world_t, camera_R, camera_t, Rs, Ts):
    """What refineWorker does with the ICP result (MainWindowForm.cs:377-405), in float32 like the C# code:
    worldTransforms[i].t += Ts[i] * R_i (row vector times matrix); cameraPoses[i].t += Ts[i]; R_i <- Rs[i]^T * R_i (both)."""
    world_R = np.array(world_R, dtype=np.float32).reshape(-1, 3, 3)
    world_t = np.array(world_t, dtype=np.float32).reshape(-1, 3)
    camera_R = np.array(camera_R, dtype=np.float32).reshape(-1, 3, 3)
    camera_t = np.array(camera_t, dtype=np.float32).reshape(-1, 3)
    for i in range(len(world_R)):
        R, T = np.asarray(Rs[i], np.float32).reshape(9), np.asarray(Ts[i], np.float32)
        tempT = np.zeros(3, np.float32)
        tempR = np.zeros((3, 3), np.float32)
        for j in range(3):
            for k in range(3):
                tempT[j] = np.float32(tempT[j] + np.float32(T[k] * world_R[i][k, j]))
            world_t[i][j] = np.float32(world_t[i][j] + tempT[j])
            camera_t[i][j] = np.float32(camera_t[i][j] + T[j])
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    tempR[j, k] = np.float32(tempR[j, k] + np.float32(R[l * 3 + j] * world_R[i][l, k]))
        world_R[i] = tempR
        camera_R[i] = tempR
    return world_R, world_t, camera_R, camera_t
