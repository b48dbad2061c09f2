"""Tuning aid (GPU): per-query work distribution of the ICP nearest-neighbour search on the bench pair."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import api, native  # noqa: E402
from livescan3d_b200.device import IcpSolver  # noqa: E402

frame, pair = bench.make_inputs(0)
A, B = bench.icp_clouds(pair, api.generate_vertices_from_depth_map)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
s = IcpSolver(len(A), len(B))
dbg = torch.zeros(3 * len(B), dtype=torch.int32, device="cuda")
lib = native.load()
lib.ls3d_icp_set_debug(s.h, C.c_void_p(dbg.data_ptr()))
s.set_target(dA)
s.set_source(dB)
for it in range(3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); s.match(); ev1.record()
    torch.cuda.synchronize()
    d = dbg.cpu().numpy().reshape(-1, 3).astype(np.int64)
    idx, d2 = s.nn()
    dist = np.sqrt(d2.cpu().numpy())
    print(f"iter {it}: match {ev0.elapsed_time(ev1) * 1000:.0f} us  block-wide (heavy) share of the queries {np.mean(d[:, 2] > 0):.3f}")
    for name, col in (("steps", 0), ("scanned", 1)):
        v = d[:, col]
        print(f"   {name}: mean {v.mean():.1f} p50 {np.percentile(v, 50):.0f} p90 {np.percentile(v, 90):.0f} p99 {np.percentile(v, 99):.0f} max {v.max()}  sum {v.sum() / 1e6:.1f}M")
    print("   nn dist [m]: p50 %.4f p90 %.4f p99 %.4f max %.3f" % tuple(np.percentile(dist, [50, 90, 99, 100])))
    far = dist > 0.05
    print(f"   far (>5cm) fraction {far.mean():.3f}: steps mean {d[far, 0].mean():.1f} scanned mean {d[far, 1].mean():.1f} | near: steps {d[~far, 0].mean():.1f} scanned {d[~far, 1].mean():.1f}")
    w = d[: len(d) // 32 * 32].reshape(-1, 32, 3)
    print(f"   per-warp max: steps mean {w[:, :, 0].max(1).mean():.1f}, scanned mean {w[:, :, 1].max(1).mean():.1f}; lane-sum/warp-max utilisation steps {w[:, :, 0].sum() / (32 * w[:, :, 0].max(1).sum() + 1):.2f} scanned {w[:, :, 1].sum() / (32 * w[:, :, 1].max(1).sum() + 1):.2f}")
    s.reduce()
