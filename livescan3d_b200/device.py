"""Device-resident drivers over Part 3 of the C ABI (include/ls3d.h): inputs and outputs stay in HBM.

torch is plumbing here — it owns the input tensors, the CUDA stream the kernels are enqueued on and (in
dist.py) the NCCL process group.  Every computation is a kernel of libls3d_b200.so; nothing in this module
computes with torch ops on the data path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import native
from .native import Ls3dError


class _DevView:
    """Expose library-owned device memory to torch without a copy (CUDA array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def view(ptr: int, shape, typestr: str, device=None) -> torch.Tensor:
    if not ptr:
        raise Ls3dError("null device pointer")
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device or torch.device("cuda", torch.cuda.current_device()))


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FramePipeline:
    """map -> world transform -> cull -> neighbour-count filter -> merge for one rig, inputs resident in HBM."""

    def __init__(self, widths, heights):
        self.lib = native.load()
        self.widths = np.ascontiguousarray(widths, dtype=np.int32)
        self.heights = np.ascontiguousarray(heights, dtype=np.int32)
        self.n_maps = len(self.widths)
        self.total_px = int((self.widths.astype(np.int64) * self.heights).sum())
        self.h = self.lib.ls3d_frame_create(self.n_maps, self.widths.ctypes.data_as(C.c_void_p), self.heights.ctypes.data_as(C.c_void_p))
        native.check(bool(self.h), "ls3d_frame_create")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.counts = view(self.lib.ls3d_frame_count_ptr(self.h), (5,), "<i4", self.device)     # n_final, n_culled, err, n_kept, n_triangles
        self.filter_on = False

    def close(self):
        if self.h:
            self.lib.ls3d_frame_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, intr, wt, bounds, filter_k=10, filter_max_dist=0.01):
        ip = np.ascontiguousarray(intr, dtype=np.float32)
        w = np.ascontiguousarray(wt, dtype=np.float32)
        assert ip.size == 7 * self.n_maps and w.size == 12 * self.n_maps
        b = [float(x) for x in bounds]
        r = self.lib.ls3d_frame_set_params(self.h, ip.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), *b, int(filter_k), float(filter_max_dist), _stream())
        native.check(r == 0, "ls3d_frame_set_params")
        self.filter_on = filter_k > 0 and filter_max_dist > 0

    def run(self, d_depth: torch.Tensor, d_colors: torch.Tensor, first_map=0, n_run=0) -> int:
        """Enqueue the whole path on the current stream (no synchronisation).  Returns kernels enqueued."""
        assert d_depth.is_cuda and d_colors.is_cuda and d_depth.is_contiguous() and d_colors.is_contiguous()
        r = self.lib.ls3d_frame_run(self.h, C.c_void_p(d_depth.data_ptr()), C.c_void_p(d_colors.data_ptr()), int(first_map), int(n_run), _stream())
        native.check(r >= 0, "ls3d_frame_run")
        return r

    def run_count(self, d_depth, d_colors, first_map=0, n_run=0) -> int:
        r = self.lib.ls3d_frame_run_count(self.h, C.c_void_p(d_depth.data_ptr()), C.c_void_p(d_colors.data_ptr()), int(first_map), int(n_run), _stream())
        native.check(r >= 0, "ls3d_frame_run_count")
        return r

    def merge_peers(self, peer_ptrs, d_offset: torch.Tensor, first_map=0, n_run=0) -> int:
        arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        r = self.lib.ls3d_frame_merge_peers(self.h, int(first_map), int(n_run), len(peer_ptrs), arr, C.c_void_p(d_offset.data_ptr()), _stream())
        native.check(r >= 0, "ls3d_frame_merge_peers")
        return r

    def enable_timing(self, on=True):
        self.lib.ls3d_frame_enable_timing(self.h, 1 if on else 0)

    STAGES = ["map_cull_compact", "hash_clear", "voxel_insert", "cell_ranges_scatter", "voxel_neighbour_count", "survivor_compact",
              "organized_neighbour_count", "whole", "triangles"]

    def stage_ms(self) -> np.ndarray:
        """Milliseconds per stage (order: FramePipeline.STAGES) of the last timed run; 0 where a stage did not run."""
        out = np.zeros(9, dtype=np.float32)
        native.check(self.lib.ls3d_frame_stage_ms(self.h, out.ctypes.data_as(C.c_void_p)) == 0, "ls3d_frame_stage_ms")
        return out

    def enable_triangles(self, on=True):
        """Also produce the depth-grid triangles on unfiltered runs (generateTrianglesGradients)."""
        self.lib.ls3d_frame_enable_triangles(self.h, 1 if on else 0)

    def triangles(self) -> torch.Tensor:
        """int32 [2 * total_px, 3] view of the triangle buffer; rows [0, counts[4]) are valid."""
        return view(self.lib.ls3d_frame_triangles(self.h), (2 * self.total_px, 3), "<i4", self.device)

    def depth_to_vertex(self) -> torch.Tensor:
        """int32 [total_px]: pixel -> index of its vertex in the culled cloud of the run, -1 = none (from the next run on)."""
        return view(self.lib.ls3d_frame_depth_to_vertex(self.h), (self.total_px,), "<i4", self.device)

    def set_filter_mode(self, mode: int):
        """0 auto, 1 voxel hash, 2 organized (pixel window)."""
        native.check(self.lib.ls3d_frame_set_filter_mode(self.h, int(mode)) == 0, "ls3d_frame_set_filter_mode")

    # ---- results (views of library memory; valid until the next run) ----
    def vertices(self) -> torch.Tensor:
        """uint8 [total_px, 16] view of the merged cloud buffer; rows [0, n_final) are valid."""
        return view(self.lib.ls3d_frame_vertices(self.h), (self.total_px, 16), "|u1", self.device)

    def culled_vertices(self) -> torch.Tensor:
        return view(self.lib.ls3d_frame_culled_vertices(self.h), (self.total_px, 16), "|u1", self.device)

    def sensor_starts(self) -> torch.Tensor:
        return view(self.lib.ls3d_frame_sensor_starts(self.h), (self.n_maps + 1,), "<i4", self.device)

    def old_to_new(self) -> torch.Tensor:
        return view(self.lib.ls3d_frame_old_to_new(self.h), (self.total_px,), "<i4", self.device)

    def keep_mask(self) -> torch.Tensor:
        """uint8 [total_px] survivor mask in pixel order (organized runs only)."""
        ptr = self.lib.ls3d_frame_keep_mask(self.h)
        native.check(bool(ptr), "ls3d_frame_keep_mask (no organized run yet)")
        return view(ptr, (self.total_px,), "|u1", self.device)

    def result(self):
        """Synchronise and fetch (vertices ndarray[VERTEX], per-sensor counts) — test/debug convenience."""
        from .api import VERTEX_DTYPE
        torch.cuda.current_stream().synchronize()
        c = self.counts.cpu().numpy()
        if c[2]:
            raise Ls3dError(f"frame pipeline device error flags 0x{int(c[2]):x}")
        n = int(c[0])
        v = self.vertices()[:n].cpu().numpy().reshape(-1).view(VERTEX_DTYPE)
        st = self.sensor_starts().cpu().numpy()
        return v, np.diff(st)


class IcpSolver:
    """Device-resident ICP: target grid built once, iterations enqueued without host round trips."""

    def __init__(self, n1_max: int, n2_max: int):
        self.lib = native.load()
        self.h = self.lib.ls3d_icp_create(int(n1_max), int(n2_max))
        native.check(bool(self.h), "ls3d_icp_create")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.n1 = self.n2 = 0
        self.Rt = view(self.lib.ls3d_icp_Rt(self.h), (12,), "<f4", self.device)
        self.status = view(self.lib.ls3d_icp_status(self.h), (4,), "<i4", self.device)

    def close(self):
        if self.h:
            self.lib.ls3d_icp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_target(self, d_verts1: torch.Tensor):
        assert d_verts1.is_cuda and d_verts1.dtype == torch.float32 and d_verts1.is_contiguous()
        self.n1 = d_verts1.numel() // 3
        self._v1 = d_verts1
        native.check(self.lib.ls3d_icp_set_target(self.h, C.c_void_p(d_verts1.data_ptr()), self.n1, _stream()) == 0, "ls3d_icp_set_target")

    def set_source(self, d_verts2: torch.Tensor, i_begin=0, i_end=None, R0=None, t0=None):
        assert d_verts2.is_cuda and d_verts2.dtype == torch.float32 and d_verts2.is_contiguous()
        self.n2 = d_verts2.numel() // 3
        self._v2 = d_verts2
        R = np.ascontiguousarray(np.eye(3) if R0 is None else R0, dtype=np.float32).reshape(9)
        t = np.ascontiguousarray(np.zeros(3) if t0 is None else t0, dtype=np.float32).reshape(3)
        ie = self.n2 if i_end is None else int(i_end)
        r = self.lib.ls3d_icp_set_source(self.h, C.c_void_p(d_verts2.data_ptr()), self.n2, int(i_begin), ie,
                                         R.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), _stream())
        native.check(r == 0, "ls3d_icp_set_source")

    def match(self):
        native.check(self.lib.ls3d_icp_match(self.h, _stream()) == 0, "ls3d_icp_match")

    def reduce(self):
        native.check(self.lib.ls3d_icp_reduce(self.h, _stream()) == 0, "ls3d_icp_reduce")

    def finish(self):
        native.check(self.lib.ls3d_icp_finish(self.h, _stream()) == 0, "ls3d_icp_finish")

    def run(self, max_iter: int):
        native.check(self.lib.ls3d_icp_run(self.h, int(max_iter), _stream()) == 0, "ls3d_icp_run")

    def slots(self) -> torch.Tensor:
        return view(self.lib.ls3d_icp_slots(self.h), (self.n1,), "<i8", self.device)

    def nn(self):
        return (view(self.lib.ls3d_icp_nn_index(self.h), (self.n2,), "<i4", self.device),
                view(self.lib.ls3d_icp_nn_dist(self.h), (self.n2,), "<f4", self.device))

    def pose(self):
        """Synchronise and return (R[3,3], t[3], status[4]) as numpy."""
        torch.cuda.current_stream().synchronize()
        rt = self.Rt.cpu().numpy()
        return rt[:9].reshape(3, 3).copy(), rt[9:].copy(), self.status.cpu().numpy()
