"""Stage times of ONE ICP call of the global refine schedule (target = the other 7 clouds of an 8-ring, source = sensor 0), GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import api, synth  # noqa: E402
from livescan3d_b200.device import IcpSolver  # noqa: E402

xyz = lambda v: np.stack([v["X"], v["Y"], v["Z"]], axis=1).astype(np.float32)
ring = synth.make_frame(8, synth.KINECT_W, synth.KINECT_H)
clouds = []
for i in range(8):
    c = xyz(api.generate_vertices_from_depth_map(ring, synth.SERVER_BOUNDS, i))
    if i:
        c = synth.perturb(c, deg=0.3 + 0.1 * i, trans_mm=(2.0 * i, -3.0, 1.0 * i))
    clouds.append(np.ascontiguousarray(c))
dA = torch.from_numpy(np.concatenate(clouds[1:])).cuda()
dB0 = torch.from_numpy(clouds[0]).cuda()
dB = dB0.clone()
s = IcpSolver(len(dA), len(dB))
its = 10
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = np.zeros((its, 3)); tgt = src = 0.0
reps = 4
for rep in range(reps + 1):
    dB.copy_(dB0)
    e = [ev() for _ in range(3)]
    e[0].record(); s.set_target(dA); e[1].record(); s.set_source(dB); e[2].record()
    marks = []
    for it in range(its):
        m = [ev() for _ in range(4)]
        m[0].record(); s.match(); m[1].record(); s.reduce(); m[2].record(); m[3].record()
        marks.append(m)
    s.finish()
    torch.cuda.synchronize()
    if rep:
        tgt += e[0].elapsed_time(e[1]); src += e[1].elapsed_time(e[2])
        acc += [[m[k].elapsed_time(m[k + 1]) for k in range(3)] for m in marks]
acc /= reps
print(f"n1={len(dA)} n2={len(dB)} set_target {1000 * tgt / reps:.0f} us, set_source {1000 * src / reps:.0f} us")
print("match us:", " ".join(f"{1000 * v:.0f}" for v in acc[:, 0]))
print("stats us:", " ".join(f"{1000 * v:.0f}" for v in acc[:, 1]), " sums us:", " ".join(f"{1000 * v:.0f}" for v in acc[:, 2]))
R, t, st = s.pose()
print("status", st.tolist())

# host-side cost of one call as the refine driver makes it (wall clock, stream synchronised at the end of each stage)
import time
for rep in range(3):
    dB.copy_(dB0); torch.cuda.synchronize()
    t = [time.perf_counter()]
    v1 = torch.cat([torch.from_numpy(c).cuda() for c in clouds[1:]]).contiguous(); torch.cuda.synchronize(); t.append(time.perf_counter())
    s.set_target(v1); t.append(time.perf_counter()); torch.cuda.synchronize(); t.append(time.perf_counter())
    s.set_source(dB); t.append(time.perf_counter()); torch.cuda.synchronize(); t.append(time.perf_counter())
    s.run(its); t.append(time.perf_counter()); torch.cuda.synchronize(); t.append(time.perf_counter())
    print("wall ms: cat+upload %.2f | set_target call %.2f (+sync %.2f) | set_source call %.2f (+sync %.2f) | run call %.2f (+sync %.2f)" % tuple(1000 * (t[i + 1] - t[i]) for i in range(7)))
