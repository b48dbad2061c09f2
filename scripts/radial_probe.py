"""Profiling driver (GPU): one radial correction of the bench frame through the host export."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from livescan3d_b200 import api
frame, _ = bench.make_inputs(0)
for _ in range(2):
    d, c = api.radial_correction(frame)
print("ok", int((d != frame["depth_maps"]).sum()))
