"""Where a global-refine ICP call spends its time (GPU): config[3], sensor 0 against the other 7 clouds."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from livescan3d_b200 import api, synth, refine
from livescan3d_b200.device import IcpSolver
xyz = lambda v: np.ascontiguousarray(np.stack([v["X"], v["Y"], v["Z"]], axis=1), dtype=np.float32)
ring = synth.make_frame(8, synth.KINECT_W, synth.KINECT_H)
clouds = []
for i in range(8):
    c = xyz(api.generate_vertices_from_depth_map(ring, synth.SERVER_BOUNDS, i))
    if i:
        c = synth.perturb(c, deg=0.3 + 0.1 * i, trans_mm=(2.0 * i, -3.0, 1.0 * i))
    clouds.append(np.ascontiguousarray(c))
dev = torch.device("cuda", 0)
dc = [torch.from_numpy(c).to(dev) for c in clouds]
v1 = torch.cat(dc[1:]).contiguous()
src0 = dc[0].clone()
s = IcpSolver(v1.shape[0], max(c.shape[0] for c in dc))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(16)]
for rep in range(4):
    src = src0.clone()
    torch.cuda.synchronize()
    ev[0].record(); s.set_target(v1); ev[1].record(); s.set_source(src); ev[2].record()
    for it in range(10):
        s.match(); ev[3 + it].record() if it < 1 else None
        s.reduce()
    ev[14].record(); s.finish(); ev[15].record()
    torch.cuda.synchronize()
print("n1", v1.shape[0], "n2", src0.shape[0])
print("set_target ms", ev[0].elapsed_time(ev[1]), "set_source", ev[1].elapsed_time(ev[2]), "first match", ev[2].elapsed_time(ev[3]), "10 iterations", ev[2].elapsed_time(ev[14]), "finish", ev[14].elapsed_time(ev[15]))
m, r = [], []
a, b, c3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
src = src0.clone(); s.set_target(v1); s.set_source(src)
for it in range(10):
    a.record(); s.match(); b.record(); s.reduce(); c3.record(); torch.cuda.synchronize()
    m.append(a.elapsed_time(b)); r.append(b.elapsed_time(c3))
print("match ms", [round(x, 3) for x in m]); print("reduce ms", [round(x, 3) for x in r])
t0 = time.perf_counter(); d2 = [c.clone() for c in dc]; torch.cuda.synchronize(); t0 = time.perf_counter()
refine.refine_poses_device(d2, 2, 10); print("refine_poses_device ms", 1000 * (time.perf_counter() - t0))
t0 = time.perf_counter(); d2 = [c.clone() for c in dc]; torch.cuda.synchronize(); t0 = time.perf_counter()
refine.refine_poses_device(d2, 2, 10); print("refine_poses_device ms (2nd)", 1000 * (time.perf_counter() - t0))
