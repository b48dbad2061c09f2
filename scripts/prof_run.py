"""Short profiling driver (GPU): the bench workloads with few launches, for `ncu --set full`.

  python scripts/prof_run.py [frames] [icp_iters]

Runs `frames` 8-sensor frames on the organized path, one on the voxel-hash path, and one ICP call of `icp_iters`
iterations as plain stream-ordered launches (no graph), all on the bench inputs (bench.make_inputs(0))."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import api  # noqa: E402
from livescan3d_b200.device import FramePipeline, IcpSolver  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3

frame, pair = bench.make_inputs(0)
dev = torch.device("cuda", 0)
d_depth = torch.from_numpy(frame["depth_maps"]).to(dev)
d_colors = torch.from_numpy(frame["depth_colors"]).to(dev)
fp = FramePipeline(frame["widths"], frame["heights"])
fp.set_params(frame["intr"], frame["wt"], bench.FRAME_BOUNDS, bench.FILTER_K, bench.FILTER_MAXDIST)
for _ in range(frames):
    fp.run(d_depth, d_colors)
torch.cuda.synchronize()
print("organized:", fp.counts.cpu().tolist())
fp.set_filter_mode(1)
fp.run(d_depth, d_colors)
torch.cuda.synchronize()
print("voxel hash:", fp.counts.cpu().tolist())

# widened rows: unfiltered mesh with triangles, radial correction (device, on a copy), flying-pixel filter
import ctypes as C  # noqa: E402
from livescan3d_b200 import native  # noqa: E402
lib = native.load()
fp.set_filter_mode(0)
fp.set_params(frame["intr"], frame["wt"], bench.FRAME_BOUNDS, 0, 0.0)
fp.enable_triangles(True)
fp.run(d_depth, d_colors)
torch.cuda.synchronize()
print("mesh:", fp.counts.cpu().tolist())
dd, dc = d_depth.clone(), d_colors.clone()
pp = lambda a: a.ctypes.data_as(C.c_void_p)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
assert lib.ls3d_radial_correction_device(bench.S, C.c_void_p(dd.data_ptr()), C.c_void_p(dc.data_ptr()), pp(frame["widths"]), pp(frame["heights"]), pp(frame["intr"]), st) > 0
one = d_depth[: 2 * bench.W_PX * bench.H_PX]
fo = torch.empty_like(one)
assert lib.ls3d_filter_flying_pixels_device(C.c_void_p(one.data_ptr()), C.c_void_p(fo.data_ptr()), bench.W_PX, bench.H_PX, 1, 10.0, st) > 0
torch.cuda.synchronize()

# N4: the mesh just produced, re-packed on the device (PLY body, transfer frame chunking)
mc = fp.counts.cpu().numpy()
nv_m, nt_m = int(mc[0]), int(mc[4])
ply_out = torch.empty(15 * nv_m + 13 * nt_m + 16, dtype=torch.uint8, device=dev)
pv, pt = C.c_void_p(fp.vertices().data_ptr()), C.c_void_p(fp.triangles().data_ptr())
assert lib.ls3d_pack_ply_body_device(pv, nv_m, pt, nt_m, C.c_void_p(ply_out.data_ptr()), st) > 0
import numpy as np  # noqa: E402
cvs, cts = np.zeros(4096, np.int32), np.zeros(4096, np.int32)
nv_out, body = C.c_int(0), C.c_void_p(0)
print("transfer chunks:", lib.ls3d_transfer_chunks_device(pv, nv_m, pt, nt_m, pp(cvs), pp(cts), 4096, C.byref(nv_out), C.byref(body), st))
torch.cuda.synchronize()

# the host entry point with page-locked inputs (chunked schedule, colours pulled by the merge kernel, copy-out kernels);
# LS3D_E2E_GRAPH=0 makes it plain stream launches
from livescan3d_b200.native import Mesh  # noqa: E402
h_depth = torch.from_numpy(frame["depth_maps"]).pin_memory()
h_colors = torch.from_numpy(frame["depth_colors"]).pin_memory()
w_arr, h_arr = np.ascontiguousarray(frame["widths"], np.int32), np.ascontiguousarray(frame["heights"], np.int32)
ipar, wtr = np.ascontiguousarray(frame["intr"], np.float32), np.ascontiguousarray(frame["wt"], np.float32)
for _ in range(2):
    mesh = Mesh()
    n = lib.ls3d_frame_pipeline(bench.S, C.c_void_p(h_depth.data_ptr()), C.c_void_p(h_colors.data_ptr()), pp(w_arr), pp(h_arr), pp(ipar), pp(wtr), C.byref(mesh),
                                *[float(x) for x in bench.FRAME_BOUNDS], bench.FILTER_K, bench.FILTER_MAXDIST, None)
    lib.deleteMesh(C.byref(mesh))
print("host pipeline:", n)

A, B = bench.icp_clouds(pair, api.generate_vertices_from_depth_map)
dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
s = IcpSolver(len(A), len(B))
s.set_target(dA)
s.set_source(dB)
for _ in range(iters):
    s.match(); s.reduce()
s.finish()
R, t, st = s.pose()
print("icp:", st.tolist(), t.tolist())
