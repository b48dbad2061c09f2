"""Multi-GPU host side of the hot path: one process per GPU, torch.distributed for the plumbing.

Two sharded variants, both bit-identical to the single-GPU result:

  ShardedFrame  one or more sensor streams per GPU.  Every rank maps, culls and neighbour-counts its own sensors
                (no communication), the per-rank survivor counts are exchanged by peer stores inside one single-block
                kernel (ls3d_frame_publish_counts), and the final compaction kernel stores every surviving 16-byte record
                straight into EVERY rank's merged buffer at its global offset over NVLink peer stores (CUDA IPC mapped
                memory): compaction + all-gather in one kernel, no collective library call anywhere.  formMesh's order (sensor order, row-major inside a sensor,
                depthprocessing.cpp:1594-1608) is kept because ranks own contiguous sensor ranges.

  ShardedIcp    the target octree is replicated, the SOURCE points are partitioned, the dedupe slots and the reduction are
                sharded by target range.  Per iteration every rank searches its slice and atomicMin's the one-to-one keys
                (bits(d2) << 32 | ~i, icp.cpp:95-126) into the OWNER rank's slot array over NVLink; the fused reduction kernel
                then exchanges chunk partials by peer stores + flags inside the launch.  No NCCL call in the loop; the poses
                are bit-identical to the single-GPU run (canonical chunked reduction, csrc/icp.cu).

The pure planning helpers at the top (no CUDA, no library) are what the world_size-2 gloo tests in tests/ exercise.
"""
from __future__ import annotations

import numpy as np

SLOT_EMPTY = 0x7FFFFFFFFFFFFFFF


# ---------------------------------------------------------------------------------------------------------
# planning helpers (host logic; CPU-testable)
# ---------------------------------------------------------------------------------------------------------
def sensor_ranges(n_sensors: int, world: int):
    """Contiguous, balanced sensor ranges per rank: [(first, count)] * world.  Contiguity is what keeps the merged cloud
    in formMesh's sensor order when rank r writes at offset sum(counts of ranks < r)."""
    if n_sensors < 0 or world <= 0:
        raise ValueError("sensor_ranges: need n_sensors >= 0 and world > 0")
    base, extra = divmod(n_sensors, world)
    out, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((first, n))
        first += n
    return out


def slice_ranges(n: int, world: int, align: int = 32):
    """[begin, end) per rank over n items, boundaries rounded to `align` (a warp of queries never straddles two ranks)."""
    if n < 0 or world <= 0 or align <= 0:
        raise ValueError("slice_ranges: bad arguments")
    per = -(-n // world)
    per = -(-per // align) * align
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def exclusive_offsets(counts):
    """Exclusive prefix sum of the per-rank survivor counts = where each rank's records start in the merged cloud."""
    c = np.asarray(counts, dtype=np.int64).reshape(-1)
    out = np.zeros_like(c)
    if len(c) > 1:
        out[1:] = np.cumsum(c[:-1])
    return out


def pack_slot_keys(n1: int, nn_index, nn_d2, i_offset: int = 0):
    """The dedupe slot array a rank produces for its slice (numpy restatement of nn_commit in csrc/icp.cu, used by the CPU
    tests): slot[j] = min over source points i with NN j of (bits(d2) << 32 | (0xFFFFFFFF - i)), else SLOT_EMPTY.
    Smaller d2 wins; on equal d2 the LATER source index wins (icp.cpp:103).  All keys are positive as int64, so a signed
    MIN all-reduce merges per-rank arrays exactly."""
    idx = np.asarray(nn_index, dtype=np.int64).reshape(-1)
    d2 = np.ascontiguousarray(nn_d2, dtype=np.float32).reshape(-1)
    i = np.arange(len(idx), dtype=np.int64) + int(i_offset)
    keys = (d2.view(np.uint32).astype(np.int64) << 32) | (0xFFFFFFFF - i)
    slots = np.full(int(n1), SLOT_EMPTY, dtype=np.int64)
    ok = idx >= 0
    np.minimum.at(slots, idx[ok], keys[ok])
    return slots


def unpack_slot_keys(slots):
    """-> (winner source index per target point or -1, d2 of the winning match)."""
    s = np.asarray(slots, dtype=np.int64).reshape(-1)
    has = s != SLOT_EMPTY
    win = np.where(has, 0xFFFFFFFF - (s & 0xFFFFFFFF), -1).astype(np.int64)
    d2 = ((s >> 32) & 0xFFFFFFFF).astype(np.uint32).view(np.float32).copy()
    d2[~has] = 0
    return win, d2


RED_CHUNKS_MAX = 296


def reduction_chunks(n1: int) -> int:
    """How many chunks the target range is cut into for the canonical reduction (red_chunks in csrc/icp.cu): a function of n1
    alone, so the chunk partials — and every total folded from them — do not depend on how many ranks share the work."""
    return max(1, min(RED_CHUNKS_MAX, (int(n1) + 255) // 256))


def reduction_chunk_size(n1: int) -> int:
    c = reduction_chunks(n1)
    return (int(n1) + c - 1) // c


def chunk_owner_ranges(n1: int, world: int):
    """[(first chunk, end chunk)] per rank: contiguous runs of ceil(C / world) chunks (ls3d_icp_reduce)."""
    if world <= 0:
        raise ValueError("chunk_owner_ranges: world must be positive")
    c = reduction_chunks(n1)
    per = -(-c // world)
    return [(min(c, r * per), min(c, (r + 1) * per)) for r in range(world)]


def slot_owner(idx, n1: int, world: int):
    """Rank that owns the dedupe slot of target point idx (SlotMap / nn_commit in csrc/icp.cu): the owner of its chunk."""
    per = max(1, -(-reduction_chunks(n1) // world))
    return np.minimum(world - 1, (np.asarray(idx, dtype=np.int64) // max(1, reduction_chunk_size(n1))) // per)


def fold_chunk_partials(partials):
    """The fixed-order fold every rank applies to the C chunk partials (red_total in csrc/icp.cu): chunk c belongs to column
    c % 32 and row (c // 32) % 8; a cell adds its chunks in increasing c, a column its 8 cells in row order, the 32 columns are
    combined by a 5-step xor butterfly.  partials: [C] or [C, k] float64 -> scalar or [k]."""
    p = np.asarray(partials, dtype=np.float64)
    p2 = p.reshape(len(p), -1)
    cells = np.zeros((256, p2.shape[1]), dtype=np.float64)
    for c in range(len(p2)):
        cells[c % 256] = cells[c % 256] + p2[c]
    acc = np.zeros((32, p2.shape[1]), dtype=np.float64)
    for w in range(8):
        acc = acc + cells[w * 32:(w + 1) * 32]
    o = 16
    while o > 0:
        acc = acc + acc[np.arange(32) ^ o]
        o >>= 1
    return acc[0] if p.ndim > 1 else float(acc[0, 0])


def _world(group=None):
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


# ---------------------------------------------------------------------------------------------------------
# peer-mapped buffers (CUDA IPC between the ranks of one node)
# ---------------------------------------------------------------------------------------------------------
class PeerBuffer:
    """The same-sized device allocation on every rank, each mapped into every other rank's address space: ptrs[r] is
    rank r's buffer as seen from this process (NVLink loads/stores).  ptrs[rank] is the local allocation."""

    def __init__(self, nbytes: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import native
        self.lib = native.load()
        self.rank, self.world = _world(group)
        self.nbytes = int(nbytes)
        self.local = self.lib.ls3d_dev_alloc(self.nbytes)
        native.check(bool(self.local), "ls3d_dev_alloc")
        self.ptrs = [None] * self.world
        self.ptrs[self.rank] = int(self.local)
        self._opened = []
        if self.world > 1:
            h = (C.c_ubyte * 64)()
            native.check(self.lib.ls3d_ipc_export(C.c_void_p(self.local), h) == 0, "ls3d_ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(h), group=group)
            for r, hb in enumerate(handles):
                if r == self.rank:
                    continue
                buf = (C.c_ubyte * 64).from_buffer_copy(hb)
                p = self.lib.ls3d_ipc_open(buf)
                native.check(bool(p), f"ls3d_ipc_open (rank {r})")
                self.ptrs[r] = int(p)
                self._opened.append(p)

    def tensor(self, shape, typestr):
        from .device import view
        return view(self.local, shape, typestr)

    def close(self):
        import ctypes as C
        for p in self._opened:
            self.lib.ls3d_ipc_close(C.c_void_p(p))
        self._opened = []
        if self.local:
            self.lib.ls3d_dev_free(C.c_void_p(self.local))
            self.local = None


# ---------------------------------------------------------------------------------------------------------
# sensor streams sharded over GPUs, merged by peer stores
# ---------------------------------------------------------------------------------------------------------
class ShardedFrame:
    """One rig over the ranks of a node: contiguous sensor ranges per rank, merged by peer stores.  Per frame a rank enqueues
    (1) everything up to the neighbour count on its own sensors, (2) ls3d_frame_publish_counts — its survivor count goes into every
    rank's sync block and its exclusive prefix comes back, all inside one single-block kernel, (3) the compaction kernel, whose
    STG.128s land in EVERY rank's merged buffer at that prefix, (4) ls3d_frame_wait_peers.  No collective library call, no host
    round trip: the four launches are stream-ordered (and graph-capturable)."""

    def __init__(self, widths, heights, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from .device import FramePipeline, view
        self.group = group
        self.rank, self.world = _world(group)
        self.fp = FramePipeline(widths, heights)           # descriptors for the whole rig; only this rank's range is run
        self.first, self.n_own = sensor_ranges(self.fp.n_maps, self.world)[self.rank]
        self.merged = PeerBuffer(16 * self.fp.total_px, group)
        self.sync = PeerBuffer(256, group)                 # Ls3dFrameSync of every rank, mapped everywhere
        self._sync_i32 = view(self.sync.local, (64,), "<i4", self.fp.device)
        self._sync_i32.zero_()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=group)                      # every block is zero before anyone's first publish can reach it
        self._offset = view(int(self.sync.local) + 4, (1,), "<i4", self.fp.device)          # Ls3dFrameSync.offset
        self._sync_tbl = (C.c_void_p * self.world)(*[C.c_void_p(int(p)) for p in self.sync.ptrs])

    def set_params(self, intr, wt, bounds, filter_k=10, filter_max_dist=0.01):
        self.fp.set_params(intr, wt, bounds, filter_k, filter_max_dist)

    def step(self, d_depth, d_colors):
        """Enqueue one frame on the current stream.  d_depth / d_colors use the packed whole-rig layout; only this rank's
        sensor range has to hold data.  Afterwards every rank's merged buffer holds the whole merged cloud."""
        from . import native
        from .device import _stream
        lib = self.fp.lib
        if self.n_own > 0:
            self.fp.run_count(d_depth, d_colors, self.first, self.n_own)
        native.check(lib.ls3d_frame_publish_counts(self.fp.h if self.n_own > 0 else None, self.rank, self.world, self._sync_tbl, _stream()) == 0, "ls3d_frame_publish_counts")
        if self.n_own > 0:
            self.fp.merge_peers(self.merged.ptrs, self._offset, self.first, self.n_own)
        native.check(lib.ls3d_frame_wait_peers(self.rank, self.world, self._sync_tbl, _stream()) == 0, "ls3d_frame_wait_peers")

    def result(self):
        """Synchronise; -> (merged VertexC4ubV3f ndarray, per-rank counts)."""
        import torch
        from .api import VERTEX_DTYPE
        from .native import Ls3dError
        torch.cuda.current_stream().synchronize()
        blk = self._sync_i32.cpu().numpy()
        if blk[3]:
            raise Ls3dError(f"sharded frame: peer exchange error flags 0x{int(blk[3]):x}")
        counts = blk[12:12 + self.world].copy()           # cnt_val
        n = int(blk[2])                                    # total
        assert n == int(counts.sum())
        v = self.merged.tensor((self.fp.total_px, 16), "|u1")[:n].cpu().numpy().reshape(-1).view(VERTEX_DTYPE)
        return v, counts

    def close(self):
        self.sync.close()
        self.merged.close()
        self.fp.close()


# ---------------------------------------------------------------------------------------------------------
# ICP with partitioned source points
# ---------------------------------------------------------------------------------------------------------
def map_peers(local_ptr: int, group=None):
    """CUDA-IPC mapping of the same allocation on every rank: -> ([pointer of rank r's buffer valid in this process], [opened handles])."""
    import ctypes as C
    import torch.distributed as dist
    from . import native
    lib = native.load()
    rank, world = _world(group)
    ptrs = [None] * world
    ptrs[rank] = int(local_ptr)
    opened = []
    if world > 1:
        h = (C.c_ubyte * 64)()
        native.check(lib.ls3d_ipc_export(C.c_void_p(local_ptr), h) == 0, "ls3d_ipc_export")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h), group=group)
        for r, hb in enumerate(handles):
            if r == rank:
                continue
            p = lib.ls3d_ipc_open((C.c_ubyte * 64).from_buffer_copy(hb))
            native.check(bool(p), f"ls3d_ipc_open (rank {r})")
            ptrs[r] = int(p)
            opened.append(p)
    return ptrs, opened


class ShardedIcp:
    """One ICP call over the ranks of a node.  The clouds are replicated; every rank searches its slice of the source and owns a
    contiguous range of the target's dedupe slots and reduction chunks.  Per iteration there is NO collective call: the match
    kernels atomicMin their one-to-one keys straight into the owner rank's slots over NVLink, and the reduction kernel exchanges
    its chunk partials with peer stores + flags inside the launch (csrc/icp.cu, k_icp_reduce).  Every rank folds the same numbers
    in the same order, so R, t and verts2 are bit-identical to the single-GPU run on every rank.  The whole call is one CUDA
    graph per rank when it repeats."""

    def __init__(self, n1_max: int, n2_max: int, group=None):
        import ctypes as C
        from . import native
        from .device import IcpSolver
        self.group = group
        self.rank, self.world = _world(group)
        self.solver = IcpSolver(n1_max, n2_max)
        self._opened = []
        if self.world > 1:
            lib = self.solver.lib
            h = self.solver.h
            tables = []
            for getter in (lib.ls3d_icp_slots, lib.ls3d_icp_red_part, lib.ls3d_icp_red_flag):
                ptrs, opened = map_peers(int(getter(h)), group)
                self._opened += opened
                tables.append((C.c_void_p * self.world)(*[C.c_void_p(p) for p in ptrs]))
            native.check(lib.ls3d_icp_set_peers(h, self.world, self.rank, tables[0], tables[1], tables[2]) == 0, "ls3d_icp_set_peers")

    def run(self, d_verts1, d_verts2, max_iter: int, R0=None, t0=None):
        """Enqueue a whole ICP call on the current stream.  Every rank passes the full clouds (replicated); verts2 is
        transformed in place on every rank, so all ranks end with identical verts2, R, t."""
        s = self.solver
        s.set_target(d_verts1)
        n2 = d_verts2.numel() // 3
        b, e = slice_ranges(n2, self.world)[self.rank]
        s.set_source(d_verts2, b, e, R0, t0)
        s.run(int(max_iter))

    def pose(self):
        return self.solver.pose()

    def close(self):
        import ctypes as C
        if self.world > 1 and self.solver.h:
            self.solver.lib.ls3d_icp_set_peers(self.solver.h, 1, 0, None, None, None)
        for p in self._opened:
            self.solver.lib.ls3d_ipc_close(C.c_void_p(p))
        self._opened = []
        self.solver.close()


# ---------------------------------------------------------------------------------------------------------
# bench.py's "sharded" block (N > 1)
# ---------------------------------------------------------------------------------------------------------
def refine_poses_sharded(clouds, n_refine_iters: int = 2, n_icp_iters: int = 10, group=None):
    """LiveScanServer's refine schedule (MainWindowForm.cs:349-375; livescan3d_b200/refine.py) with every ICP call sharded over
    the ranks: target = the other sensors' clouds (replicated), source = cloud i, searched in slices.  clouds: CUDA tensors,
    identical on every rank, moved in place.  Returns (Rs, Ts) — the same bits on every rank and as refine_poses_device."""
    import torch
    S = len(clouds)
    n = [int(c.shape[0]) for c in clouds]
    si = ShardedIcp(max(1, sum(n) - (min(n) if n else 0)), max(1, max(n) if n else 1), group)
    try:
        Rs = np.stack([np.eye(3, dtype=np.float32) for _ in range(S)])
        Ts = np.zeros((S, 3), dtype=np.float32)
        poses = torch.zeros((S, 12), dtype=torch.float32, device=clouds[0].device)
        poses[:, 0] = poses[:, 4] = poses[:, 8] = 1.0
        keep_alive = []
        for it in range(int(n_refine_iters)):
            if it > 0:
                torch.cuda.current_stream().synchronize()
                rt = poses.cpu().numpy()
                Rs, Ts = rt[:, :9].reshape(S, 3, 3).copy(), rt[:, 9:].copy()
                keep_alive.clear()
            for i in range(S):
                others = [clouds[j] for j in range(S) if j != i and n[j] > 0]
                if not others or n[i] == 0:
                    continue
                verts1 = torch.cat(others).contiguous()
                keep_alive.append(verts1)
                si.run(verts1, clouds[i], n_icp_iters, Rs[i], Ts[i])
                poses[i].copy_(si.solver.Rt)
        torch.cuda.current_stream().synchronize()
        out = poses.cpu().numpy()
    finally:
        si.close()
    return out[:, :9].reshape(S, 3, 3).copy(), out[:, 9:].copy()


def bench_sharded(args, rank, world, dev, flush):
    """Strong-scaling variants, every one checked bit for bit against this rank's own single-GPU result (a mismatch raises, which
    makes bench.py exit non-zero): ONE rig split over the ranks (peer-store merge) and ONE ICP call with the source partitioned —
    on the bench sizes (8 x 512x424, 2 x 213 k points), on BASELINE.json configs[4] (8 x 1920x1080; 2 x 2 M points) and on
    configs[3] (the global refine schedule: 16 ICP calls against the other seven clouds).  Times are CUDA events (the refine: wall
    clock), max over ranks; the single-GPU time of the same work, measured in the same process, sits beside each."""
    import os
    import time
    import torch
    import torch.distributed as dist
    import bench
    from . import api, refine, synth
    from .device import FramePipeline, IcpSolver

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(flag: bool) -> bool:
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    steps = max(3, args.steps)
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]

    def timed(fn, prep=None, reps=steps):
        for _ in range(max(args.warmup, 3)):
            if prep:
                prep()
            fn()
        sync()
        for i in range(reps):
            if prep:
                prep()
            flush.zero_()
            ev_s[i].record()
            fn()
            ev_e[i].record()
        sync()
        return max_over_ranks(float(np.median([a.elapsed_time(b) for a, b in zip(ev_s[:reps], ev_e[:reps])])))

    def frame_case(frame, bounds, k, md, label):
        d_depth = torch.from_numpy(frame["depth_maps"]).to(dev)
        d_colors = torch.from_numpy(frame["depth_colors"]).to(dev)
        fp = FramePipeline(frame["widths"], frame["heights"])
        fp.set_params(frame["intr"], frame["wt"], bounds, k, md)
        one_ms = timed(lambda: fp.run(d_depth, d_colors))
        want, _ = fp.result()
        want = want.copy()
        fp.close()
        sf = ShardedFrame(frame["widths"], frame["heights"])
        sf.set_params(frame["intr"], frame["wt"], bounds, k, md)
        ms = timed(lambda: sf.step(d_depth, d_colors))
        got, counts = sf.result()
        same = all_ok(got.tobytes() == want.tobytes())
        sf.close()
        del d_depth, d_colors
        res = {"workload": label, "ms_per_step": ms, "clouds_per_s": 1000.0 / ms, "single_gpu_ms_per_step": one_ms, "speedup_vs_single_gpu": one_ms / ms, "scaling": "strong",
               "merged": int(counts.sum()), "per_rank_counts": [int(c) for c in counts], "bit_identical_to_single_gpu_on_every_rank": same,
               "exchange": "survivor counts and completion flags as peer stores inside two single-block kernels + peer stores of 16*n_kept bytes to every rank (NVLink, CUDA IPC); no collective call"}
        if not same:
            raise RuntimeError(f"sharded frame ({label}): the merged cloud differs from the single-GPU result on at least one rank")
        return res

    def icp_case(A, B, label):
        dA = torch.from_numpy(A).to(dev)
        dB0 = torch.from_numpy(B).to(dev)
        dB = dB0.clone()
        one = IcpSolver(len(A), len(B))

        def run_one():
            one.set_target(dA); one.set_source(dB); one.run(bench.ICP_ITERS)
        one_ms = timed(run_one, prep=lambda: dB.copy_(dB0))
        R1, t1, _ = one.pose()
        v1 = dB.clone()
        one.close()
        si = ShardedIcp(len(A), len(B))
        ms = timed(lambda: si.run(dA, dB, bench.ICP_ITERS), prep=lambda: dB.copy_(dB0))
        R, t, st = si.pose()
        same = all_ok(np.array_equal(R, R1) and np.array_equal(t, t1) and bool(torch.equal(dB, v1)) and int(st[1]) == 0)
        si.close()
        res = {"workload": label, "n1": len(A), "n2": len(B), "ms_per_step": ms, "ms_per_iter": ms / bench.ICP_ITERS, "Mpts_iter_per_s": len(B) * bench.ICP_ITERS / (ms / 1000.0) / 1e6,
               "single_gpu_ms_per_step": one_ms, "speedup_vs_single_gpu": one_ms / ms, "scaling": "strong", "bit_identical_to_single_gpu_on_every_rank": same, "status": [int(x) for x in st],
               "exchange": "dedupe keys: 64-bit atomicMin into the owner rank's slots over NVLink; partial sums: peer stores + flags inside k_icp_reduce; no collective call in the loop"}
        del dA, dB, dB0, v1
        if not same:
            raise RuntimeError(f"sharded ICP ({label}): R, t or the moved cloud differ from the single-GPU result (status {st}); they must be bit-identical")
        return res

    xyz = lambda v: np.ascontiguousarray(np.stack([v["X"], v["Y"], v["Z"]], axis=1), dtype=np.float32)
    out = {}
    frame, pair = bench.make_inputs(0)                       # every rank: the SAME rig / pair (rank 0's)
    out["frame_one_rig_over_ranks"] = frame_case(frame, bench.FRAME_BOUNDS, bench.FILTER_K, bench.FILTER_MAXDIST, "8 x 512x424, cull +-1.5 m, filter (10, 0.01)")
    A, B = bench.icp_clouds(pair, api.generate_vertices_from_depth_map)
    out["icp_source_partitioned"] = icp_case(A, B, "bench pair: sensors 0/1 of the 8-ring at 512x424, cull +-5 m")
    if os.environ.get("LS3D_BENCH_STRESS", "1") != "0":
        # ---- BASELINE.json configs[3]: the global refine (2 refine iterations x 8 sensors, target = the other 7 clouds, ~1.49 M points)
        ring = synth.make_frame(8, synth.KINECT_W, synth.KINECT_H)
        clouds = []
        for i in range(8):
            c = xyz(api.generate_vertices_from_depth_map(ring, synth.SERVER_BOUNDS, i))
            if i:
                c = synth.perturb(c, deg=0.3 + 0.1 * i, trans_mm=(2.0 * i, -3.0, 1.0 * i))
            clouds.append(np.ascontiguousarray(c))
        dev0 = [torch.from_numpy(c).to(dev) for c in clouds]

        def wall(fn, reps=3):
            ts = []
            for it in range(reps + 1):
                dc = [c.clone() for c in dev0]
                sync()
                t0 = time.perf_counter()
                r = fn(dc)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                if it:
                    ts.append(dt)
            return max_over_ranks(float(np.median(ts))) * 1000.0, r, dc
        one_ms, (R1s, T1s), dc1 = wall(lambda dc: refine.refine_poses_device(dc, 2, bench.ICP_ITERS))
        ms, (Rs, Ts), dcs = wall(lambda dc: refine_poses_sharded(dc, 2, bench.ICP_ITERS))
        same = all_ok(np.array_equal(Rs, R1s) and np.array_equal(Ts, T1s) and all(bool(torch.equal(a, b)) for a, b in zip(dc1, dcs)))
        n2_total = 2 * sum(len(c) for c in clouds)
        out["global_refine_8_sensors"] = {"workload": "BASELINE.json configs[3]: refine schedule of MainWindowForm.cs:349-375, 16 ICP calls x 10 iterations, target = the other 7 clouds",
                                          "n_target_per_call": int(sum(len(c) for c in clouds) - len(clouds[0])), "ms_per_refine": ms, "single_gpu_ms_per_refine": one_ms, "speedup_vs_single_gpu": one_ms / ms,
                                          "Mpts_iter_per_s": n2_total * bench.ICP_ITERS / (ms / 1000.0) / 1e6, "bit_identical_to_single_gpu_on_every_rank": same,
                                          "timing": "wall clock around the whole schedule (device-resident clouds, one host wait per refine iteration), median of 3, max over ranks"}
        del dev0, dc1, dcs
        if not same:
            raise RuntimeError("sharded global refine: poses or clouds differ from the single-GPU schedule")
        # ---- BASELINE.json configs[4]: 8 x 1920x1080 and the 2 M-point ICP
        W, H = 1920, 1080
        big = synth.make_frame(8, W, H)
        out["stress_frame_one_rig_over_ranks"] = frame_case(big, synth.DEFAULT_BOUNDS, 10, 0.004, "BASELINE.json configs[4]: 8 x 1920x1080, cull +-1.5 m, filter (10, 0.004)")
        del big
        pair_big = synth.make_frame(2, W, H, ring=8)
        A2 = xyz(api.generate_vertices_from_depth_map(pair_big, synth.SERVER_BOUNDS, 0))
        B2 = synth.perturb(xyz(api.generate_vertices_from_depth_map(pair_big, synth.SERVER_BOUNDS, 1)))
        torch.cuda.empty_cache()
        out["stress_icp_source_partitioned"] = icp_case(A2, B2, "BASELINE.json configs[4]: sensors 0/1 of the 8-ring at 1920x1080, cull +-5 m (2 x ~2 M points)")
    return out
