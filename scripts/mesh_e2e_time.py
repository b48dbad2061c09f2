"""Wall time of generateMeshFromDepthMaps (8-sensor bench frame, page-locked inputs) through the C ABI; checks counts.  For tuning the host
schedule: LS3D_E2E_CHUNKS=4 LS3D_E2E_COPY_BLOCKS=16 python scripts/mesh_e2e_time.py [steps]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import native  # noqa: E402
from livescan3d_b200.native import Mesh  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
frame, _ = bench.make_inputs(0)
lib = native.load()
S = bench.S
h_depth = torch.from_numpy(frame["depth_maps"]).pin_memory()
h_colors = torch.from_numpy(frame["depth_colors"]).pin_memory()
arrs = [np.ascontiguousarray(frame[k], t) for k, t in (("widths", np.int32), ("heights", np.int32), ("intr", np.float32), ("wt", np.float32))]
p = lambda a: C.c_void_p(a.ctypes.data)
b = [float(x) for x in bench.FRAME_BOUNDS]


def call():
    mesh = Mesh()
    lib.generateMeshFromDepthMaps(S, C.c_void_p(h_depth.data_ptr()), C.c_void_p(h_colors.data_ptr()), *[p(a) for a in arrs], C.byref(mesh), 0, *b, 0)
    r = (mesh.nVertices, mesh.nTriangles)
    lib.deleteMesh(C.byref(mesh))
    return r


for _ in range(5):
    first = call()
assert first[0] > 0 and first[1] > 0, native.last_error()
ts = []
for _ in range(steps):
    t0 = time.perf_counter(); r = call(); ts.append(time.perf_counter() - t0)
    assert r == first
ts = np.array(ts) * 1e3
print(" ".join(f"{k[9:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("LS3D_E2E_")), first, f"mean {ts.mean():.4f} ms median {np.median(ts):.4f} min {ts.min():.4f}")
