"""Per-iteration differences between the GPU and oracle ICP traces (GPU box): bench pair and the small test pair."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import icp_pair, small_frame, synth, orc
from livescan3d_b200 import api
for name, (A, B) in (("small 160x120", icp_pair(small_frame(S=2, w=160, h=120), synth.DEFAULT_BOUNDS)), ("full 512x424", icp_pair(synth.make_frame(2, ring=8), synth.SERVER_BOUNDS))):
    _, wR, wt, wtr = orc.orc_icp(A, B, max_iter=10)
    _, gR, gt, gtr = api.icp_trace(A, B, max_iter=10)
    print(name, "n", len(A), len(B), "dR", float(np.max(np.abs(gR.astype(np.float64) - wR))), "dt", float(np.max(np.abs(gt.astype(np.float64) - wt))))
    for k, (g, w) in enumerate(zip(gtr, wtr)):
        print(f"  it {k}: matched {w['n_matched']} d {g['n_matched'] - w['n_matched']:+d}  accepted {w['n_accepted']} d {g['n_accepted'] - w['n_accepted']:+d}  sigma rel {abs(g['sigma'] - w['sigma']) / w['sigma']:.2e}")
