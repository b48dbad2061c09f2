"""Match vs reduce time per iteration of the bench ICP pair (GPU, CUDA events around each stage; plain stream launches)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import api  # noqa: E402
from livescan3d_b200.device import IcpSolver  # noqa: E402

frame, pair = bench.make_inputs(0)
A, B = bench.icp_clouds(pair, api.generate_vertices_from_depth_map)
dA, dB0 = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
dB = dB0.clone()
s = IcpSolver(len(A), len(B))
flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
its = 10
m, r = np.zeros(its), np.zeros(its)
for rep in range(6):
    dB.copy_(dB0)
    flush.zero_()
    s.set_target(dA)
    s.set_source(dB)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(its)]
    for it in range(its):
        ev[it][0].record(); s.match(); ev[it][1].record(); s.reduce(); ev[it][2].record()
    s.finish()
    torch.cuda.synchronize()
    if rep:
        m += [ev[i][0].elapsed_time(ev[i][1]) * 1000 for i in range(its)]
        r += [ev[i][1].elapsed_time(ev[i][2]) * 1000 for i in range(its)]
print("match us :", " ".join(f"{v / 5:.0f}" for v in m))
print("reduce us:", " ".join(f"{v / 5:.0f}" for v in r))
