"""2-rank debug: per-call comparison of the sharded refine schedule against the single-GPU one (torchrun --nproc-per-node 2)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from livescan3d_b200 import api, synth
from livescan3d_b200 import dist as ldist
from livescan3d_b200.device import IcpSolver
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
xyz = lambda v: np.ascontiguousarray(np.stack([v["X"], v["Y"], v["Z"]], axis=1), dtype=np.float32)
ring = synth.make_frame(8, synth.KINECT_W, synth.KINECT_H)
clouds = []
for i in range(8):
    c = xyz(api.generate_vertices_from_depth_map(ring, synth.SERVER_BOUNDS, i))
    if i: c = synth.perturb(c, deg=0.3 + 0.1 * i, trans_mm=(2.0 * i, -3.0, 1.0 * i))
    clouds.append(np.ascontiguousarray(c))
dc = [torch.from_numpy(c).to(dev) for c in clouds]
n = [len(c) for c in clouds]
one = IcpSolver(sum(n) - min(n), max(n))
si = ldist.ShardedIcp(sum(n) - min(n), max(n))
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for i in range(8):
    v1 = torch.cat([dc[j] for j in range(8) if j != i]).contiguous()
    a = dc[i].clone(); b = dc[i].clone()
    one.set_target(v1); one.set_source(a); one.run(iters); R1, t1, s1 = one.pose()
    si.run(v1, b, iters); R2, t2, s2 = si.pose()
    same = np.array_equal(R1, R2) and np.array_equal(t1, t2) and bool(torch.equal(a, b))
    tr1 = torch.empty(0)
    print(f"rank {rank} call {i}: n1 {v1.shape[0]} n2 {a.shape[0]} same {same} status one {s1.tolist()} sharded {s2.tolist()} dR {float(np.max(np.abs(R1 - R2))):.3e}", flush=True)
    dc[i] = a
dist.barrier()
dist.destroy_process_group()
