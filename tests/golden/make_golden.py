"""Generates tests/golden/hotpath_small.npz from the reference's OWN sources compiled in place (oracle/_ref).

Run in the build container (needs /root/reference): `make -C oracle && python tests/golden/make_golden.py`.
The fixture travels to the GPU box, where /root/reference does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import cloud_of, icp_pair, synth, orc  # noqa: E402

S, W, H, SEED, RING = 2, 128, 96, 1000, 8
assert orc.have_ref(), "build oracle/_ref first (make -C oracle)"
fr = synth.make_frame(S, W, H, seed_base=SEED, ring=RING)
bounds = synth.DEFAULT_BOUNDS
verts, counts, triangles = orc.ref_generate_mesh(fr, bounds, with_triangles=True)
xyz, rgba = cloud_of(fr, bounds, 0)
fk = np.array([10, 10, 1, 50], dtype=np.int32)
fd = np.array([0.01, 0.1, 0.01, 0.05], dtype=np.float32)
out = dict(S=S, w=W, h=H, seed_base=SEED, ring=RING, bounds=bounds, depth_maps=fr["depth_maps"], vertices=verts, vertex_counts=counts, triangles=triangles,
           filter_k=fk, filter_maxdist=fd)
for i, (k, md) in enumerate(zip(fk, fd)):
    out[f"filter_map_{i}"] = orc.ref_filter(xyz, rgba, int(k), float(md))[2]
A, B = icp_pair(fr, bounds)
ni, nd = orc.ref_find_closest(A, B)
out["nn_index"], out["nn_d2"] = ni, nd
iters = 5
_, R, t = orc.ref_icp(A, B, max_iter=iters)
out.update(icp_iters=iters, icp_R=R, icp_t=t)
np.savez_compressed(os.path.join(HERE, "hotpath_small.npz"), **out)
print("wrote hotpath_small.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})
