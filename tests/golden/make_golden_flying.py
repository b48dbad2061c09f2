"""Generates tests/golden/flying_small.npz (N2) from the reference's OWN filterFlyingPixels, kinectCapture.cpp:132-174,
compiled in place into oracle/_ref (oracle/Makefile extracts those lines in-stream).

Run in the build container (needs /root/reference): `make -C oracle && python tests/golden/make_golden_flying.py`."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import synth, orc  # noqa: E402

assert orc.have_ref(), "build oracle/_ref first (make -C oracle)"
W, H, SEED = 96, 64, 1000
d = synth.make_frame(1, W, H, seed_base=SEED)["depth_maps"].view(np.uint16).copy()
ks = np.array([1, 2, 3, 1], dtype=np.int32)
thrs = np.array([10.0, 25.0, 5.5, 0.0], dtype=np.float32)
out = dict(w=W, h=H, seed_base=SEED, depth=d, k=ks, thr=thrs)
for i, (k, thr) in enumerate(zip(ks, thrs)):
    out[f"out_{i}"] = orc.ref_filter_flying_pixels(d, W, H, int(k), float(thr), 0)
np.savez_compressed(os.path.join(HERE, "flying_small.npz"), **out)
print("wrote flying_small.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})
