"""ICP-only profiling driver (GPU) for `ncu`: the bench pair, `iters` iterations as plain stream launches.
  python scripts/prof_icp.py [iters] [n_points_scale]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import api  # noqa: E402
from livescan3d_b200.device import IcpSolver  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
frame, pair = bench.make_inputs(0)
A, B = bench.icp_clouds(pair, api.generate_vertices_from_depth_map)
dev = torch.device("cuda", 0)
dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
s = IcpSolver(len(A), len(B))
s.set_target(dA)
s.set_source(dB)
for _ in range(iters):
    s.match(); s.reduce()
s.finish()
R, t, st = s.pose()
print("icp:", st.tolist(), t.tolist())
