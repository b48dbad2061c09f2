"""Stress configuration (BASELINE.json configs[4], GPU): 8 streams at 1920x1080 through the frame path, and the ICP size
sweep n in {217k, 500k, 1M, 2M}.  Prints one JSON object; device-resident, CUDA events, L2 flushed between steps.

  python scripts/stress_bench.py [steps]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from livescan3d_b200 import api, synth  # noqa: E402
from livescan3d_b200.device import FramePipeline, IcpSolver  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
hbm_peak, peak_src = bench.peaks()
flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
W, H, S = 1920, 1080, 8
t0 = time.time()
frame = synth.make_frame(S, W, H)
gen_s = time.time() - t0
px = S * W * H
d_depth = torch.from_numpy(frame["depth_maps"]).to(dev)
d_colors = torch.from_numpy(frame["depth_colors"]).to(dev)
fp = FramePipeline(frame["widths"], frame["heights"])


def timed(fn, reps=steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for it in range(reps + 3):
        flush.zero_()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if it >= 3:
            tot += a.elapsed_time(b)
    return tot / reps


out = {"workload": f"{S} x {W}x{H} u16 depth + RGB, cull +-1.5 m", "pixels": px, "synthetic_generation_s": gen_s, "hbm_peak_gbs": hbm_peak, "peak_source": peak_src}
fp.set_params(frame["intr"], frame["wt"], synth.DEFAULT_BOUNDS, 0, 0.0)
ms = timed(lambda: fp.run(d_depth, d_colors))
n = int(fp.counts.cpu()[0])
alg = 5 * px + 16 * n
out["map_cull_merge"] = {"ms": ms, "vertices": n, "alg_bytes": alg, "gbs": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / hbm_peak}
fp.enable_triangles(True)
ms = timed(lambda: fp.run(d_depth, d_colors))
out["map_cull_merge_triangles"] = {"ms": ms, "triangles": int(fp.counts.cpu()[4])}
fp.enable_triangles(False)
for k, md in ((10, 0.004), (10, 0.01)):
    fp.set_params(frame["intr"], frame["wt"], synth.DEFAULT_BOUNDS, k, md)
    for mode, name in ((0, "auto"), (1, "voxel_hash")):
        fp.set_filter_mode(mode)
        ms = timed(lambda: fp.run(d_depth, d_colors))
        c = fp.counts.cpu().numpy()
        assert c[2] == 0
        out[f"filtered_frame_k{k}_r{md}_{name}"] = {"ms": ms, "clouds_per_s": 1000.0 / ms, "merged": int(c[0])}
fp.set_filter_mode(0)
fp.close()

# ---- ICP size sweep: sensors 0 and 1 of the ring at full resolution, +-5 m (every valid pixel), strided to the target sizes
pair = synth.make_frame(2, W, H, ring=8)
xyz = lambda v: np.ascontiguousarray(np.stack([v["X"], v["Y"], v["Z"]], axis=1), dtype=np.float32)
A_full = xyz(api.generate_vertices_from_depth_map(pair, synth.SERVER_BOUNDS, 0))
B_full = synth.perturb(xyz(api.generate_vertices_from_depth_map(pair, synth.SERVER_BOUNDS, 1)))
sweep = []
for target in (217_000, 500_000, 1_000_000, 2_000_000):
    st = max(1, int(round(len(A_full) / target)))
    A, B = np.ascontiguousarray(A_full[::st]), np.ascontiguousarray(B_full[::st])
    dA, dB0 = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    dB = dB0.clone()
    s = IcpSolver(len(A), len(B))

    def call():
        s.set_target(dA); s.set_source(dB); s.run(bench.ICP_ITERS)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for it in range(steps + 2):
        dB.copy_(dB0)
        flush.zero_()
        a.record(); call(); b.record()
        torch.cuda.synchronize()
        if it >= 2:
            tot += a.elapsed_time(b)
    ms = tot / steps
    R, t, stt = s.pose()
    assert stt[0] == bench.ICP_ITERS and stt[1] == 0
    sweep.append({"n1": len(A), "n2": len(B), "ms_per_call": ms, "ms_per_iter": ms / bench.ICP_ITERS, "Mpts_iter_per_s": len(B) * bench.ICP_ITERS / ms / 1e3})
    s.close()
out["icp_sweep"] = sweep

# ---- global ICP (BASELINE.json configs[3]): the refine-calibration schedule, 2 refine iterations x 8 sensors, target = the other 7 clouds
from livescan3d_b200 import refine  # noqa: E402
ring = synth.make_frame(8, synth.KINECT_W, synth.KINECT_H)
clouds = []
for i in range(8):
    c = xyz(api.generate_vertices_from_depth_map(ring, synth.SERVER_BOUNDS, i))
    if i:
        c = synth.perturb(c, deg=0.3 + 0.1 * i, trans_mm=(2.0 * i, -3.0, 1.0 * i))
    clouds.append(np.ascontiguousarray(c))
dev_clouds0 = [torch.from_numpy(c).to(dev) for c in clouds]
# the sweep above left multi-GB buffers in torch's caching allocator; the refine driver allocates its ICP context with plain
# cudaMalloc inside the timed region, which is slow (and was what this figure measured) while that cache is full
del dA, dB, dB0
torch.cuda.empty_cache()
reps = max(3, steps // 2)
samples = []
for it in range(reps + 2):
    dc = [c.clone() for c in dev_clouds0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Rs, Ts = refine.refine_poses_device(dc, 2, bench.ICP_ITERS)
    if it >= 2:
        samples.append(time.perf_counter() - t0)
tot = float(np.median(samples)) * reps
n2_total = 2 * sum(len(c) for c in clouds)
out["global_icp_refine_8_sensors"] = {"ms_per_refine": 1000 * tot / reps, "icp_calls": 16, "iters_per_call": bench.ICP_ITERS, "n_target_per_call": int(sum(len(c) for c in clouds) - len(clouds[0])),
                                      "Mpts_iter_per_s": n2_total * bench.ICP_ITERS / (tot / reps) / 1e6, "timing": "wall clock around refine_poses_device incl. creating and destroying its ICP context (device-resident clouds, one host wait per refine iteration), median"}
if "--cpu" in sys.argv:
    from oracle import oracle_lib as orc  # noqa: E402
    v1 = np.concatenate(clouds[1:])
    t0 = time.perf_counter()
    (orc.ref_icp if orc.have_ref() else orc.orc_icp)(v1, clouds[0], max_iter=2)
    t_cpu = time.perf_counter() - t0
    out["global_icp_refine_8_sensors"]["cpu_reference_estimate_ms"] = 1000 * t_cpu / 2 * bench.ICP_ITERS * 16
    out["global_icp_refine_8_sensors"]["cpu_sample"] = f"one ICP() call of 2 iterations on sensor 0 vs the other 7 clouds ({t_cpu:.2f} s, {os.cpu_count()} cores), scaled to 16 calls x {bench.ICP_ITERS} iterations"
print(json.dumps(out))
