"""Per-stage time of the bench frame (GPU), for A/B builds: LS3D_B200_LIB=<lib> python scripts/frame_stage_time.py [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200.device import FramePipeline  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
frame, _ = bench.make_inputs(0)
dev = torch.device("cuda", 0)
d_depth = torch.from_numpy(frame["depth_maps"]).to(dev)
d_colors = torch.from_numpy(frame["depth_colors"]).to(dev)
flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
fp = FramePipeline(frame["widths"], frame["heights"])
fp.set_params(frame["intr"], frame["wt"], bench.FRAME_BOUNDS, bench.FILTER_K, bench.FILTER_MAXDIST)
for mode, name in ((0, "auto"), (1, "voxel hash")):
    fp.set_filter_mode(mode)
    fp.enable_timing(False)
    for _ in range(100):
        fp.run(d_depth, d_colors)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    whole = []
    for _ in range(steps):
        flush.zero_()
        a.record(); fp.run(d_depth, d_colors); b.record(); torch.cuda.synchronize()
        whole.append(a.elapsed_time(b))
    fp.enable_timing(True)
    acc = np.zeros(9)
    for _ in range(steps):
        flush.zero_()
        fp.run(d_depth, d_colors)
        acc += fp.stage_ms()
    acc /= steps
    print(os.environ.get("LS3D_B200_LIB", "default"), name, "n_final", int(fp.counts.cpu()[0]), f"untimed-stage frame {1000 * np.median(whole):.1f}us |",
          " ".join(f"{n}={1000 * v:.1f}us" for n, v in zip(FramePipeline.STAGES, acc) if v > 0))
