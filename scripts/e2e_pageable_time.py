"""End-to-end frame time through ls3d_frame_pipeline with pinned vs pageable caller buffers (GPU box), and parity of the results."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from livescan3d_b200 import native
from livescan3d_b200.native import Mesh
lib = native.load()
frame, _ = bench.make_inputs(0)
S = bench.S
p = lambda a: C.c_void_p(a.ctypes.data)
w, h = np.ascontiguousarray(frame["widths"], np.int32), np.ascontiguousarray(frame["heights"], np.int32)
ip, wt = np.ascontiguousarray(frame["intr"], np.float32), np.ascontiguousarray(frame["wt"], np.float32)
b = [float(x) for x in bench.FRAME_BOUNDS]
hd, hc = torch.from_numpy(frame["depth_maps"]).pin_memory(), torch.from_numpy(frame["depth_colors"]).pin_memory()
gd, gc = np.array(frame["depth_maps"], copy=True), np.array(frame["depth_colors"], copy=True)
def call(dp, cp, keep=False):
    mesh = Mesh(); pm = np.zeros(S, np.int32)
    n = lib.ls3d_frame_pipeline(S, dp, cp, p(w), p(h), p(ip), p(wt), C.byref(mesh), *b, bench.FILTER_K, bench.FILTER_MAXDIST, p(pm))
    out = C.string_at(mesh.vertices, n * 16) if keep and n > 0 else None
    lib.deleteMesh(C.byref(mesh))
    assert n > 0, native.last_error()
    return out
ref = call(C.c_void_p(hd.data_ptr()), C.c_void_p(hc.data_ptr()), True)
for name, dp, cp in (("pinned", C.c_void_p(hd.data_ptr()), C.c_void_p(hc.data_ptr())), ("pageable", p(gd), p(gc))):
    assert call(dp, cp, True) == ref, name
    # alternate two different pageable images so a stale staging copy would show
    if name == "pageable":
        gd2 = gd.copy(); gd2.view(np.uint16)[1000:200000] //= 2
        r2 = call(p(gd2), p(gc), True); assert r2 != ref
        assert call(p(gd), p(gc), True) == ref
    for _ in range(20): call(dp, cp)
    ts = []
    for _ in range(300):
        t0 = time.perf_counter(); call(dp, cp); ts.append(time.perf_counter() - t0)
    print(f"{name}: median {1e3 * np.median(ts):.3f} ms  p10 {1e3 * np.percentile(ts, 10):.3f}  p90 {1e3 * np.percentile(ts, 90):.3f}")
