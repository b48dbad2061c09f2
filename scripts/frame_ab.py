"""Organized-path frame time of the bench frame for one library build (A/B): LS3D_B200_LIB=<lib> python scripts/frame_ab.py [steps]
prints the merged count, a checksum of the merged bytes (equal across builds = same bits), the untimed-stage frame time (median and
10th percentile) and the two stage times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200.device import FramePipeline  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
frame, _ = bench.make_inputs(0)
dev = torch.device("cuda", 0)
d_depth = torch.from_numpy(frame["depth_maps"]).to(dev)
d_colors = torch.from_numpy(frame["depth_colors"]).to(dev)
flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
fp = FramePipeline(frame["widths"], frame["heights"])
fp.set_params(frame["intr"], frame["wt"], bench.FRAME_BOUNDS, bench.FILTER_K, bench.FILTER_MAXDIST)
fp.enable_timing(False)
for _ in range(100):
    fp.run(d_depth, d_colors)
torch.cuda.synchronize()
n = int(fp.counts.cpu()[0])
chk = int(fp.vertices()[:n].to(torch.int64).sum())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
whole = []
for _ in range(steps):
    flush.zero_()
    a.record(); fp.run(d_depth, d_colors); b.record(); torch.cuda.synchronize()
    whole.append(a.elapsed_time(b))
fp.enable_timing(True)
acc = np.zeros(9)
for _ in range(50):
    flush.zero_()
    fp.run(d_depth, d_colors)
    acc += fp.stage_ms()
acc /= 50
print(os.environ.get("LS3D_B200_LIB", "default").split("/")[-1], "n", n, "chk", chk, f"frame {1000 * np.median(whole):.1f}us p10 {1000 * np.percentile(whole, 10):.1f}us |",
      f"count={1000 * acc[6]:.1f}us map={1000 * acc[0]:.1f}us")
