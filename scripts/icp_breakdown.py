"""Where one ICP call spends its time (GPU, CUDA events): python scripts/icp_breakdown.py [W H]  (sensors 0/1 of the 8-ring, cull +-5 m)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from livescan3d_b200 import api, synth
from livescan3d_b200.device import IcpSolver
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 424)
xyz = lambda v: np.ascontiguousarray(np.stack([v["X"], v["Y"], v["Z"]], axis=1), dtype=np.float32)
pair = synth.make_frame(2, W, H, ring=8)
A = xyz(api.generate_vertices_from_depth_map(pair, synth.SERVER_BOUNDS, 0))
B = synth.perturb(xyz(api.generate_vertices_from_depth_map(pair, synth.SERVER_BOUNDS, 1)))
dev = torch.device("cuda", 0)
dA, dB0 = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
s = IcpSolver(len(A), len(B))
E = lambda: torch.cuda.Event(enable_timing=True)
acc = np.zeros(4); m = np.zeros(10); r = np.zeros(10)
reps = 4
for rep in range(reps + 1):
    dB = dB0.clone()
    e = [E() for _ in range(4)]; em = [[E(), E(), E()] for _ in range(10)]
    torch.cuda.synchronize()
    e[0].record(); s.set_target(dA); e[1].record(); s.set_source(dB); e[2].record()
    for it in range(10):
        em[it][0].record(); s.match(); em[it][1].record(); s.reduce(); em[it][2].record()
    s.finish(); e[3].record()
    torch.cuda.synchronize()
    if rep:
        acc += [e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[0].elapsed_time(e[3])]
        m += [em[i][0].elapsed_time(em[i][1]) for i in range(10)]; r += [em[i][1].elapsed_time(em[i][2]) for i in range(10)]
print(f"n1 {len(A)} n2 {len(B)}: set_target {acc[0] / reps:.3f} ms, set_source {acc[1] / reps:.3f}, iterations+finish {acc[2] / reps:.3f}, whole {acc[3] / reps:.3f}")
print("match ms ", " ".join(f"{v / reps:.3f}" for v in m))
print("reduce ms", " ".join(f"{v / reps:.3f}" for v in r))
dB = dB0.clone(); a, b = E(), E(); ts = []
for rep in range(5):
    dB.copy_(dB0); a.record(); s.set_target(dA); s.set_source(dB); s.run(10); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("graph call ms", " ".join(f"{t:.3f}" for t in ts))
