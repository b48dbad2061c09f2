"""Per-stage times of the sharded ICP (torchrun --nproc-per-node N): python scripts/icp_sharded_breakdown.py [W H]"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from livescan3d_b200 import api, synth
from livescan3d_b200 import dist as ldist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 424)
xyz = lambda v: np.ascontiguousarray(np.stack([v["X"], v["Y"], v["Z"]], axis=1), dtype=np.float32)
pair = synth.make_frame(2, W, H, ring=8)
A = xyz(api.generate_vertices_from_depth_map(pair, synth.SERVER_BOUNDS, 0))
B = synth.perturb(xyz(api.generate_vertices_from_depth_map(pair, synth.SERVER_BOUNDS, 1)))
dA, dB0 = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
si = ldist.ShardedIcp(len(A), len(B))
s = si.solver
b, e = ldist.slice_ranges(len(B), world)[rank]
E = lambda: torch.cuda.Event(enable_timing=True)
acc = np.zeros(3); m = np.zeros(10); r = np.zeros(10); reps = 4
for rep in range(reps + 1):
    dB = dB0.clone()
    ev = [E() for _ in range(4)]; em = [[E(), E(), E()] for _ in range(10)]
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev[0].record(); s.set_target(dA); ev[1].record(); s.set_source(dB, b, e); ev[2].record()
    for it in range(10):
        em[it][0].record(); s.match(); em[it][1].record(); s.reduce(); em[it][2].record()
    s.finish(); ev[3].record()
    torch.cuda.synchronize()
    if rep:
        acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[0].elapsed_time(ev[3])]
        m += [em[i][0].elapsed_time(em[i][1]) for i in range(10)]; r += [em[i][1].elapsed_time(em[i][2]) for i in range(10)]
print(f"rank {rank}/{world} n1 {len(A)} slice {e - b}: set_target {acc[0] / reps:.3f} set_source+sync {acc[1] / reps:.3f} whole {acc[2] / reps:.3f} | match " + " ".join(f"{v / reps:.3f}" for v in m[:5]) + " | reduce " + " ".join(f"{v / reps:.3f}" for v in r[:5]), flush=True)
dB = dB0.clone(); a, bb = E(), E(); ts = []
for rep in range(6):
    dB.copy_(dB0); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a.record(); si.run(dA, dB, 10); bb.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(bb))
print(f"rank {rank} graph call ms " + " ".join(f"{t:.3f}" for t in ts), flush=True)
si.close(); dist.barrier(); dist.destroy_process_group()
