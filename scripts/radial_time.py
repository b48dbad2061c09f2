"""Device time of the radial correction of the bench frame (GPU), for A/B builds."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from livescan3d_b200 import native
frame, _ = bench.make_inputs(0)
lib = native.load()
d0 = torch.from_numpy(frame["depth_maps"]).cuda(); c0 = torch.from_numpy(frame["depth_colors"]).cuda()
d, c = d0.clone(), c0.clone()
p = lambda a: a.ctypes.data_as(C.c_void_p)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for it in range(8):
    d.copy_(d0); c.copy_(c0)
    a.record()
    assert lib.ls3d_radial_correction_device(bench.S, C.c_void_p(d.data_ptr()), C.c_void_p(c.data_ptr()), p(frame["widths"]), p(frame["heights"]), p(frame["intr"]), st) > 0
    b.record(); torch.cuda.synchronize()
    if it >= 3: tot += a.elapsed_time(b)
print(os.environ.get("LS3D_B200_LIB", "default").split("/")[-1], f"radial {tot / 5:.3f} ms", int(d.sum()))
