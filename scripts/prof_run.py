"""Short profiling driver (GPU): the bench workloads with few launches, for `ncu --set full`.

  python scripts/prof_run.py [frames] [icp_iters]

Runs `frames` 8-sensor frames on the organized path, one on the voxel-hash path, and one ICP call of `icp_iters`
iterations as plain stream-ordered launches (no graph), all on the bench inputs (bench.make_inputs(0))."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import api  # noqa: E402
from livescan3d_b200.device import FramePipeline, IcpSolver  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3

frame, pair = bench.make_inputs(0)
dev = torch.device("cuda", 0)
d_depth = torch.from_numpy(frame["depth_maps"]).to(dev)
d_colors = torch.from_numpy(frame["depth_colors"]).to(dev)
fp = FramePipeline(frame["widths"], frame["heights"])
fp.set_params(frame["intr"], frame["wt"], bench.FRAME_BOUNDS, bench.FILTER_K, bench.FILTER_MAXDIST)
for _ in range(frames):
    fp.run(d_depth, d_colors)
torch.cuda.synchronize()
print("organized:", fp.counts.cpu().tolist())
fp.set_filter_mode(1)
fp.run(d_depth, d_colors)
torch.cuda.synchronize()
print("voxel hash:", fp.counts.cpu().tolist())

A, B = bench.icp_clouds(pair, api.generate_vertices_from_depth_map)
dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
s = IcpSolver(len(A), len(B))
s.set_target(dA)
s.set_source(dB)
for _ in range(iters):
    s.match(); s.stats(); s.sums()
s.finish()
R, t, st = s.pose()
print("icp:", st.tolist(), t.tolist())
