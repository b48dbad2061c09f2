"""Phase timeline of k_icp_reduce's block 0 (needs a -DLS3D_RED_TIMING=1 build: scripts/ab_build.sh redt -DLS3D_RED_TIMING=1)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LS3D_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "livescan3d_b200/csrc/build/ab/libls3d_redt.so"))
import bench
from livescan3d_b200 import api, native
from livescan3d_b200.device import IcpSolver
frame, pair = bench.make_inputs(0)
A, B = bench.icp_clouds(pair, api.generate_vertices_from_depth_map)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
s = IcpSolver(len(A), len(B)); s.set_target(dA); s.set_source(dB)
lib = C.CDLL(os.environ["LS3D_B200_LIB"])
names = ["start", "pdl wait", "sched+sync0", "pass1+store", "barrier1", "fold1", "pass2+store", "barrier2", "fold2", "pass3+store", "barrier3", "fold3", "solve"]
acc = np.zeros(12)
for it in range(6):
    s.match(); s.reduce(); torch.cuda.synchronize()
    t = (C.c_ulonglong * 16)()
    assert lib.ls3d_debug_red_timing(t) == 0
    t = np.array(t[:13], dtype=np.float64)
    if it:
        acc += np.diff(t)
for n, v in zip(names[1:], acc / 5):
    print(f"{n:14s} {v / 1000:7.2f} us")
print("total", acc.sum() / 5000)
