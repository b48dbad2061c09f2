"""Per-iteration match time of the bench ICP pair (GPU), for A/B builds: LS3D_B200_LIB=<lib> python scripts/icp_time.py"""
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
acc = np.zeros(its)
whole = []
for rep in range(6):
    dB.copy_(dB0)
    flush.zero_()
    s.set_target(dA)
    s.set_source(dB)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(its + 1)]
    for it in range(its):
        ev[it].record(); s.match(); s.reduce()
    ev[its].record(); s.finish()
    torch.cuda.synchronize()
    if rep:
        acc += [ev[i].elapsed_time(ev[i + 1]) * 1000 for i in range(its)]
    dB.copy_(dB0)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.set_target(dA); s.set_source(dB); s.run(its); e1.record()
    torch.cuda.synchronize()
    if rep:
        whole.append(e0.elapsed_time(e1) * 1000)
R, t, st = s.pose()
print(os.environ.get("LS3D_B200_LIB", "default").split("/")[-1], "iter us:", " ".join(f"{v / 5:.0f}" for v in acc), "| graph call us:", f"{np.mean(whole):.0f}", "status", st.tolist(), "t", t.tolist())
