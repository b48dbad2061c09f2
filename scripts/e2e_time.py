"""Wall time of ls3d_frame_pipeline on the bench frame through the C ABI (pinned host buffers), checked against the device-resident
run; for tuning the host pipeline:  LS3D_E2E_DEPTH_CHUNKS=2 LS3D_E2E_MERGE_CHUNKS=8 python scripts/e2e_time.py [steps]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import native  # noqa: E402
from livescan3d_b200.device import FramePipeline  # noqa: E402
from livescan3d_b200.native import Mesh  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
frame, _ = bench.make_inputs(0)
lib = native.load()
S = bench.S
h_depth = torch.from_numpy(frame["depth_maps"]).pin_memory()
h_colors = torch.from_numpy(frame["depth_colors"]).pin_memory()
w_arr = np.ascontiguousarray(frame["widths"], np.int32)
h_arr = np.ascontiguousarray(frame["heights"], np.int32)
ip = np.ascontiguousarray(frame["intr"], np.float32)
wt = np.ascontiguousarray(frame["wt"], np.float32)
pm = np.zeros(S, np.int32)
p = lambda a: C.c_void_p(a.ctypes.data)
b = [float(x) for x in bench.FRAME_BOUNDS]

# device-resident result to compare against
fp = FramePipeline(frame["widths"], frame["heights"])
fp.set_params(frame["intr"], frame["wt"], bench.FRAME_BOUNDS, bench.FILTER_K, bench.FILTER_MAXDIST)
fp.run(torch.from_numpy(frame["depth_maps"]).cuda(), torch.from_numpy(frame["depth_colors"]).cuda())
n_ref = int(fp.counts.cpu()[0])
ref = fp.vertices()[:n_ref].cpu().numpy().reshape(-1)


def call(check=False):
    mesh = Mesh()
    n = lib.ls3d_frame_pipeline(S, C.c_void_p(h_depth.data_ptr()), C.c_void_p(h_colors.data_ptr()), p(w_arr), p(h_arr), p(ip), p(wt), C.byref(mesh),
                                *b, bench.FILTER_K, bench.FILTER_MAXDIST, p(pm))
    if check:
        assert n == n_ref, (n, n_ref, native.last_error())
        got = np.ctypeslib.as_array(C.cast(mesh.vertices, C.POINTER(C.c_uint8)), shape=(16 * n,)).copy()
        if ref is not None:
            assert np.array_equal(got, ref), "host pipeline differs from the device-resident run"
        assert int(pm.sum()) == n
    lib.deleteMesh(C.byref(mesh))
    return n


if os.environ.get("E2E_RAW"):
    # raw PCIe floor: the frame's input bytes up, the result's bytes down (copy engine, CUDA events)
    dd, dc = torch.empty_like(h_depth, device="cuda"), torch.empty_like(h_colors, device="cuda")
    out_d = torch.empty(16 * n_ref, dtype=torch.uint8, device="cuda")
    out_h = torch.empty(16 * n_ref, dtype=torch.uint8).pin_memory()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    up, down = [], []
    for _ in range(20):
        e0.record(); dd.copy_(h_depth, non_blocking=True); dc.copy_(h_colors, non_blocking=True); e1.record()
        out_h.copy_(out_d, non_blocking=True); e2.record(); torch.cuda.synchronize()
        up.append(e0.elapsed_time(e1)); down.append(e1.elapsed_time(e2))
    print(f"raw H2D {h_depth.numel() + h_colors.numel()} B: {np.median(up):.4f} ms ({(h_depth.numel() + h_colors.numel()) / np.median(up) / 1e6:.1f} GB/s)   "
          f"raw D2H {16 * n_ref} B: {np.median(down):.4f} ms ({16 * n_ref / np.median(down) / 1e6:.1f} GB/s)")

for _ in range(5):
    call(check=True)
ts = []
for _ in range(steps):
    t0 = time.perf_counter()
    call()
    ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(" ".join(f"{k[9:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("LS3D_E2E_")) + f" n={n_ref} "
      f"mean {ts.mean():.4f} ms  median {np.median(ts):.4f}  min {ts.min():.4f}  -> {1000.0 / ts.mean():.0f} clouds/s (checked={'bytes' if ref is not None else 'count'})")
