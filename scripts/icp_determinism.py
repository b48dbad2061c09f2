"""Bit-level repeatability of ICP on one GPU: the same small pair through fresh and reused contexts, many times."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import icp_pair, small_frame, synth
from livescan3d_b200.device import IcpSolver
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 96)
A, B = icp_pair(small_frame(S=2, w=w, h=h), synth.SERVER_BOUNDS)
dev = torch.device("cuda", 0)
dA = torch.from_numpy(A).to(dev)
ref = None
bad = 0
s = None
for rep in range(n_rep):
    if rep % 3 == 0:
        if s: s.close()
        s = IcpSolver(len(A), len(B))
    dB = torch.from_numpy(B).to(dev)
    s.set_target(dA); s.set_source(dB); s.run(5)
    R, t, st = s.pose()
    key = (R.tobytes(), t.tobytes(), dB.cpu().numpy().tobytes(), tuple(st.tolist()))
    if ref is None:
        ref = key
    elif key != ref:
        bad += 1
        print("rep", rep, "differs: status", st.tolist(), "dR", float(np.max(np.abs(R - np.frombuffer(ref[0], np.float32).reshape(3, 3)))))
print("n", len(A), len(B), "reps", n_rep, "mismatches", bad)
