"""Frame-only profiling driver (GPU): `frames` 8-sensor bench frames on the organized path as plain stream-ordered launches, for
`ncu --set full --kernel-name regex:"k_organized_count|k_map_cull_compact"`.   python scripts/prof_frame.py [frames]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200.device import FramePipeline  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frame, _ = bench.make_inputs(0)
dev = torch.device("cuda", 0)
d_depth = torch.from_numpy(frame["depth_maps"]).to(dev)
d_colors = torch.from_numpy(frame["depth_colors"]).to(dev)
fp = FramePipeline(frame["widths"], frame["heights"])
fp.set_params(frame["intr"], frame["wt"], bench.FRAME_BOUNDS, bench.FILTER_K, bench.FILTER_MAXDIST)
for _ in range(frames):
    fp.run(d_depth, d_colors)
torch.cuda.synchronize()
print("organized:", fp.counts.cpu().tolist())
