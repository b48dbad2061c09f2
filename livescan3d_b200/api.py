"""Host-side mirror of the reference's interface for the hot path, on numpy arrays.

Each function calls exactly one C-ABI entry of libls3d_b200.so with HOST buffers — the same call
LiveScanServer's P/Invoke would make (include/ls3d.h cites the reference declaration behind each one):

  generate_vertices_from_depth_map  -> generateVerticesFromDepthMap   (depthprocessing.h:103-105)
  generate_mesh_from_depth_maps     -> generateMeshFromDepthMaps      (depthprocessing.h:108-110)
  frame_pipeline                    -> ls3d_frame_pipeline            (createVertices + filter + formMesh)
  filter                            -> ls3d_filter                    (filter.h:64)
  icp / icp_trace                   -> ICP / ls3d_icp_trace           (icp.h:65)
  find_closest                      -> ls3d_find_closest              (icp.cpp:18-32)

There is no fallback: a missing library raises ImportError, a failing call raises Ls3dError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native
from .native import Ls3dError, Mesh, IcpTrace

VERTEX_DTYPE = np.dtype([("R", "u1"), ("G", "u1"), ("B", "u1"), ("A", "u1"), ("X", "<f4"), ("Y", "<f4"), ("Z", "<f4")])   # VertexC4ubV3f
assert VERTEX_DTYPE.itemsize == 16


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _take_mesh(lib, mesh: Mesh, what: str, triangles: bool = False):
    """Copy Mesh.vertices (and Mesh.triangles) out — what the C# side does with Marshal.Copy, KinectServer.cs:342-352,
    376-389 — then deleteMesh."""
    err = native.last_error()
    if err.startswith("note:"):                               # non-fatal: e.g. flags whose extras are outside this path
        import warnings
        warnings.warn(f"{what}: {err}")
        err = ""
    try:
        n = mesh.nVertices
        if n > 0 and not mesh.vertices:
            raise Ls3dError(f"{what}: {err or 'no vertices returned'}")
        out = np.empty(n, dtype=VERTEX_DTYPE)
        if n > 0:
            C.memmove(out.ctypes.data, mesh.vertices, n * 16)
        nt = mesh.nTriangles
        if not mesh.triangles:
            raise Ls3dError(f"{what}: Mesh.triangles must never be NULL after a call")
        tri = np.empty((nt, 3), dtype=np.int32)
        if nt > 0:
            C.memmove(tri.ctypes.data, mesh.triangles, nt * 12)
    finally:
        lib.deleteMesh(C.byref(mesh))
    if err:
        raise Ls3dError(f"{what}: {err}")
    return (out, tri) if triangles else out


def _frame_args(frame):
    return (_c(frame["depth_maps"], np.uint8), _c(frame["depth_colors"], np.uint8), _c(frame["widths"], np.int32),
            _c(frame["heights"], np.int32), _c(frame["intr"], np.float32), _c(frame["wt"], np.float32))


def generate_vertices_from_depth_map(frame: dict, bounds, depth_map_index: int) -> np.ndarray:
    """One sensor of a packed frame -> VertexC4ubV3f[n] (map, +t, R*, strict cull, row-major order)."""
    lib = native.load()
    d, c, w, h, ip, wt = _frame_args(frame)
    b = [float(x) for x in bounds]
    mesh = Mesh()
    lib.generateVerticesFromDepthMap(_ptr(d), _ptr(c), _ptr(w), _ptr(h), _ptr(ip), _ptr(wt), C.byref(mesh), *b, int(depth_map_index))
    return _take_mesh(lib, mesh, "generateVerticesFromDepthMap")


def generate_mesh_from_depth_maps(frame: dict, bounds, color_transfer: bool = False, generate_triangles: bool = False, triangles: bool = False):
    """All sensors -> one merged VertexC4ubV3f[n] in sensor order; with triangles=True -> (vertices, int32[nt,3] triangles),
    the whole Mesh generateMeshFromDepthMaps fills."""
    lib = native.load()
    d, c, w, h, ip, wt = _frame_args(frame)
    b = [float(x) for x in bounds]
    mesh = Mesh()
    lib.generateMeshFromDepthMaps(int(frame["n_maps"]), _ptr(d), _ptr(c), _ptr(w), _ptr(h), _ptr(ip), _ptr(wt), C.byref(mesh),
                                  int(bool(color_transfer)), *b, int(bool(generate_triangles)))
    return _take_mesh(lib, mesh, "generateMeshFromDepthMaps", triangles)


def radial_correction(frame: dict):
    """depthMapAndColorSetRadialCorrection on copies of the frame's packed buffers -> (depth_maps u8[], depth_colors u8[])."""
    lib = native.load()
    d = np.array(frame["depth_maps"], dtype=np.uint8, order="C")
    c = np.array(frame["depth_colors"], dtype=np.uint8, order="C")
    w, h, ip = _c(frame["widths"], np.int32), _c(frame["heights"], np.int32), _c(frame["intr"], np.float32)
    lib.depthMapAndColorSetRadialCorrection(int(frame["n_maps"]), _ptr(d), _ptr(c), _ptr(w), _ptr(h), _ptr(ip))
    err = native.last_error()
    if err:
        raise Ls3dError(f"depthMapAndColorSetRadialCorrection: {err}")
    return d, c


def filter_flying_pixels(depth_u16: np.ndarray, width: int, height: int, neighbourhood_size: int = 1, thr: float = 10.0, max_non_fitting: int = 0):
    """KinectCapture::filterFlyingPixels (kinectCapture.cpp:132-174) on a copy of one depth image -> filtered uint16[h*w]."""
    lib = native.load()
    d = np.array(depth_u16, dtype=np.uint16, order="C").reshape(-1)
    if d.size != width * height:
        raise ValueError("depth size does not match width*height")
    r = lib.ls3d_filter_flying_pixels(_ptr(d), int(width), int(height), int(neighbourhood_size), float(thr), int(max_non_fitting))
    native.check(r >= 0, "ls3d_filter_flying_pixels")
    return d


def frame_pipeline(frame: dict, bounds, filter_k: int = 10, filter_max_dist: float = 0.01):
    """map -> world transform -> cull -> per-sensor neighbour-count filter -> merge.  Returns (vertices, per_map_counts)."""
    lib = native.load()
    d, c, w, h, ip, wt = _frame_args(frame)
    b = [float(x) for x in bounds]
    mesh = Mesh()
    counts = np.zeros(int(frame["n_maps"]), dtype=np.int32)
    n = lib.ls3d_frame_pipeline(int(frame["n_maps"]), _ptr(d), _ptr(c), _ptr(w), _ptr(h), _ptr(ip), _ptr(wt), C.byref(mesh),
                                *b, int(filter_k), float(filter_max_dist), _ptr(counts))
    verts = _take_mesh(lib, mesh, "ls3d_frame_pipeline")
    if n < 0:
        raise Ls3dError(f"ls3d_frame_pipeline: {native.last_error()}")
    return verts, counts


def filter(verts: np.ndarray, colors: np.ndarray, k: int = 10, max_dist: float = 0.01):
    """filter(vertices, colors, k, maxDist) of filter.h:64 on flat arrays.

    verts: [n,3] float32, colors: [n,4] uint8 (RGB struct).  Returns (verts_kept, colors_kept, old_to_new) with
    old_to_new[i] = new index, -1 if removed (or -2 everywhere when the reference's early return applies)."""
    lib = native.load()
    v = np.array(verts, dtype=np.float32, order="C").reshape(-1, 3)
    c = np.array(colors, dtype=np.uint8, order="C").reshape(-1, 4)
    if len(c) != len(v):
        raise ValueError("verts and colors must have the same length")
    n = len(v)
    m = np.empty(n, dtype=np.int32)
    kept = lib.ls3d_filter(_ptr(v), _ptr(c), n, int(k), float(max_dist), _ptr(m))
    native.check(kept >= 0, "ls3d_filter")
    return v[:kept].copy(), c[:kept].copy(), m


def icp_trace(verts1: np.ndarray, verts2: np.ndarray, R=None, t=None, max_iter: int = 10, trace: bool = True):
    """ICP(verts1, verts2, n1, n2, R, t, maxIter) of icp.h:65.  Returns (verts2_out, R, t, trace_records|None)."""
    lib = native.load()
    v1 = _c(np.asarray(verts1).reshape(-1, 3), np.float32)
    v2 = np.array(np.asarray(verts2).reshape(-1, 3), dtype=np.float32, order="C")
    Rm = np.array(np.eye(3) if R is None else R, dtype=np.float32, order="C").reshape(9)
    tv = np.array(np.zeros(3) if t is None else t, dtype=np.float32, order="C").reshape(3)
    tr = (IcpTrace * max(int(max_iter), 1))() if trace else None
    ret = lib.ls3d_icp_trace(_ptr(v1), _ptr(v2), len(v1), len(v2), _ptr(Rm), _ptr(tv), int(max_iter), tr)
    err = native.last_error()
    if err:
        raise Ls3dError(f"ICP: {err}")
    assert ret == 1.0
    recs = None
    if trace:
        recs = [dict(n_matched=x.n_matched, n_accepted=x.n_accepted, sigma=float(x.sigma), T=np.array(x.T[:], dtype=np.float32),
                     Rk=np.array(x.Rk[:], dtype=np.float32).reshape(3, 3)) for x in tr[:max(int(max_iter), 0)]]
    return v2, Rm.reshape(3, 3), tv, recs


def icp(verts1, verts2, R=None, t=None, max_iter: int = 10):
    lib = native.load()
    v1 = _c(np.asarray(verts1).reshape(-1, 3), np.float32)
    v2 = np.array(np.asarray(verts2).reshape(-1, 3), dtype=np.float32, order="C")
    Rm = np.array(np.eye(3) if R is None else R, dtype=np.float32, order="C").reshape(9)
    tv = np.array(np.zeros(3) if t is None else t, dtype=np.float32, order="C").reshape(3)
    lib.ICP(_ptr(v1), _ptr(v2), len(v1), len(v2), _ptr(Rm), _ptr(tv), int(max_iter))
    err = native.last_error()
    if err:
        raise Ls3dError(f"ICP: {err}")
    return v2, Rm.reshape(3, 3), tv


def find_closest(verts1, verts2):
    """FindClosestPointForEach (icp.cpp:18-32): (indices uint64[n2], squared distances float32[n2])."""
    lib = native.load()
    v1 = _c(np.asarray(verts1).reshape(-1, 3), np.float32)
    v2 = _c(np.asarray(verts2).reshape(-1, 3), np.float32)
    idx = np.empty(len(v2), dtype=np.uint64)
    d2 = np.empty(len(v2), dtype=np.float32)
    r = lib.ls3d_find_closest(_ptr(v1), len(v1), _ptr(v2), len(v2), _ptr(idx), _ptr(d2))
    native.check(r == 0, "ls3d_find_closest")
    return idx, d2


def version() -> str:
    return native.load().ls3d_version().decode()
