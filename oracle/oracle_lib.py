"""ctypes access to the CPU oracle (oracle/libls3d_oracle.so) and, when built, the reference's own sources
compiled in place (oracle/_ref/libls3d_ref_{native,filter}.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, as the checker or the timed CPU baseline.  Nothing under livescan3d_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libls3d_oracle.so")
REF_NATIVE_SO = os.path.join(_HERE, "_ref", "libls3d_ref_native.so")
REF_FILTER_SO = os.path.join(_HERE, "_ref", "libls3d_ref_filter.so")

VERTEX_DTYPE = np.dtype([("R", "u1"), ("G", "u1"), ("B", "u1"), ("A", "u1"), ("X", "<f4"), ("Y", "<f4"), ("Z", "<f4")])


class OrcIcpTrace(C.Structure):
    _fields_ = [("n_matched", C.c_int), ("n_accepted", C.c_int), ("sigma", C.c_float), ("T", C.c_float * 3), ("Rk", C.c_float * 9)]


class RefMesh(C.Structure):
    _fields_ = [("nVertices", C.c_int), ("vertices", C.c_void_p), ("nTriangles", C.c_int), ("triangles", C.c_void_p)]


def build(ref: bool = True):
    """make -C oracle (the restatement always; oracle/_ref only when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE, "oracle"] + (["ref"] if ref else []), check=True, stdout=subprocess.DEVNULL)


_p = lambda a: a.ctypes.data_as(C.c_void_p)
_f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)

_orc = None
_refn = None
_reff = None


def oracle():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _orc = C.CDLL(ORACLE_SO)
        _orc.orc_icp_trace.restype = C.c_float
        _orc.orc_icp.restype = C.c_float
    return _orc


def have_ref() -> bool:
    return os.path.exists(REF_NATIVE_SO) and os.path.exists(REF_FILTER_SO)


def ref_native():
    global _refn
    if _refn is None:
        _refn = C.CDLL(REF_NATIVE_SO)
        _refn.ICP.restype = C.c_float
    return _refn


def ref_filter_lib():
    global _reff
    if _reff is None:
        _reff = C.CDLL(REF_FILTER_SO)
    return _reff


# ---------------------------------------------------------------------------------------------------------
# oracle (restatement)
# ---------------------------------------------------------------------------------------------------------
def orc_generate_mesh(frame: dict, bounds, map_index: int = -1):
    """-> (VertexC4ubV3f[n], per_map_counts)"""
    o = oracle()
    S = int(frame["n_maps"])
    w = np.ascontiguousarray(frame["widths"], dtype=np.int32)
    h = np.ascontiguousarray(frame["heights"], dtype=np.int32)
    total = int((w.astype(np.int64) * h).sum())
    out = np.zeros(total, dtype=VERTEX_DTYPE)
    counts = np.zeros(S, dtype=np.int32)
    d = np.ascontiguousarray(frame["depth_maps"], dtype=np.uint8)
    c = np.ascontiguousarray(frame["depth_colors"], dtype=np.uint8)
    ip, wt, b = _f32(frame["intr"]), _f32(frame["wt"]), _f32(bounds)
    n = o.orc_generate_mesh(S, _p(d), _p(c), _p(w), _p(h), _p(ip), _p(wt), _p(b), int(map_index), _p(out), _p(counts))
    return out[:n].copy(), counts


def orc_generate_mesh_triangles(frame: dict, bounds):
    """The (bcolor_transfer, bgenerate_triangles) = (false, false) branch of generateMeshFromDepthMaps, triangles included
    -> (VertexC4ubV3f[n], triangles int32[nt,3], per_map_vertex_counts, per_map_triangle_counts)"""
    o = oracle()
    S = int(frame["n_maps"])
    w = np.ascontiguousarray(frame["widths"], dtype=np.int32)
    h = np.ascontiguousarray(frame["heights"], dtype=np.int32)
    total = int((w.astype(np.int64) * h).sum())
    out = np.zeros(total, dtype=VERTEX_DTYPE)
    tri = np.zeros(6 * total + 3, dtype=np.int32)
    counts = np.zeros(S, dtype=np.int32)
    tcounts = np.zeros(S, dtype=np.int32)
    nt = C.c_int(0)
    d = np.ascontiguousarray(frame["depth_maps"], dtype=np.uint8)
    c = np.ascontiguousarray(frame["depth_colors"], dtype=np.uint8)
    ip, wt, b = _f32(frame["intr"]), _f32(frame["wt"]), _f32(bounds)
    n = o.orc_generate_mesh_triangles(S, _p(d), _p(c), _p(w), _p(h), _p(ip), _p(wt), _p(b), _p(out), _p(tri), C.byref(nt), _p(counts), _p(tcounts))
    return out[:n].copy(), tri[:3 * nt.value].reshape(-1, 3).copy(), counts, tcounts


def orc_radial_correction(frame: dict):
    """depthMapAndColorSetRadialCorrection on copies of the frame's buffers -> (depth_maps u8[], depth_colors u8[])"""
    o = oracle()
    d = np.array(frame["depth_maps"], dtype=np.uint8, order="C")
    c = np.array(frame["depth_colors"], dtype=np.uint8, order="C")
    w = np.ascontiguousarray(frame["widths"], dtype=np.int32)
    h = np.ascontiguousarray(frame["heights"], dtype=np.int32)
    o.orc_radial_correction(int(frame["n_maps"]), _p(d), _p(c), _p(w), _p(h), _p(_f32(frame["intr"])))
    return d, c


def orc_filter_flying_pixels(depth_u16, w: int, h: int, k: int = 1, thr: float = 10.0, max_non_fitting: int = 0):
    o = oracle()
    d = np.array(depth_u16, dtype=np.uint16, order="C").reshape(-1)
    o.orc_filter_flying_pixels(_p(d), int(w), int(h), int(k), C.c_float(thr), int(max_non_fitting))
    return d


def ref_radial_correction(frame: dict):
    """The reference's own export depthMapAndColorSetRadialCorrection (depthprocessing.cpp:1794-1815) on copies."""
    r = ref_native()
    d = np.array(frame["depth_maps"], dtype=np.uint8, order="C")
    c = np.array(frame["depth_colors"], dtype=np.uint8, order="C")
    w = np.ascontiguousarray(frame["widths"], dtype=np.int32)
    h = np.ascontiguousarray(frame["heights"], dtype=np.int32)
    ip = _f32(frame["intr"]).copy()
    r.depthMapAndColorSetRadialCorrection(int(frame["n_maps"]), _p(d), _p(c), _p(w), _p(h), _p(ip))
    return d, c


def orc_vertex_maps(depth_u16, colors, w, h, intr7, wt12, bounds):
    """createVertices side outputs for one sensor -> (n, depth_to_vertices[w*h], vertices_to_depth[n])"""
    o = oracle()
    d = np.ascontiguousarray(depth_u16, dtype=np.uint16)
    c = np.ascontiguousarray(colors, dtype=np.uint8)
    xyz = np.zeros(3 * w * h, dtype=np.float32)
    rgb = np.zeros(3 * w * h, dtype=np.uint8)
    d2v = np.zeros(w * h, dtype=np.int32)
    v2d = np.zeros(w * h, dtype=np.int32)
    ip, wt, b = _f32(intr7), _f32(wt12), _f32(bounds)
    n = o.orc_create_vertices(_p(d), _p(c), int(w), int(h), _p(ip), _p(wt), _p(b), _p(xyz), _p(rgb), _p(d2v), _p(v2d))
    return n, d2v, v2d[:n].copy()


def orc_find_closest(v1, v2, brute: bool = False):
    o = oracle()
    v1, v2 = _f32(v1).reshape(-1, 3), _f32(v2).reshape(-1, 3)
    idx = np.zeros(len(v2), dtype=np.uint64)
    d = np.zeros(len(v2), dtype=np.float32)
    fn = o.orc_find_closest_brute if brute else o.orc_find_closest
    fn(_p(v1), len(v1), _p(v2), len(v2), _p(idx), _p(d))
    return idx, d


def orc_knn_kdist(v, k: int, brute: bool = False):
    o = oracle()
    v = _f32(v).reshape(-1, 3)
    out = np.zeros(len(v), dtype=np.float32)
    (o.orc_knn_kdist_brute if brute else o.orc_knn_kdist)(_p(v), len(v), int(k), _p(out))
    return out


def orc_filter(verts, colors, k: int, max_dist: float):
    """-> (verts_kept, colors_kept, old_to_new)"""
    o = oracle()
    v = np.array(verts, dtype=np.float32, order="C").reshape(-1, 3)
    c = np.array(colors, dtype=np.uint8, order="C").reshape(-1, 4)
    m = np.zeros(len(v), dtype=np.int32)
    n = o.orc_filter(_p(v), _p(c), len(v), int(k), C.c_float(max_dist), _p(m))
    return v[:n].copy(), c[:n].copy(), m


def orc_icp(v1, v2, R=None, t=None, max_iter: int = 10):
    """-> (verts2_out, R[3,3], t[3], trace list)"""
    o = oracle()
    v1 = _f32(v1).reshape(-1, 3)
    v2 = np.array(v2, dtype=np.float32, order="C").reshape(-1, 3)
    Rm = np.array(np.eye(3) if R is None else R, dtype=np.float32, order="C").reshape(9)
    tv = np.array(np.zeros(3) if t is None else t, dtype=np.float32, order="C").reshape(3)
    tr = (OrcIcpTrace * max(max_iter, 1))()
    o.orc_icp_trace(_p(v1), _p(v2), len(v1), len(v2), _p(Rm), _p(tv), int(max_iter), tr)
    recs = [dict(n_matched=x.n_matched, n_accepted=x.n_accepted, sigma=float(x.sigma), T=np.array(x.T[:], dtype=np.float32),
                 Rk=np.array(x.Rk[:], dtype=np.float32).reshape(3, 3)) for x in tr[:max_iter]]
    return v2, Rm.reshape(3, 3), tv, recs


def orc_dedupe(indices, dists, n1: int):
    o = oracle()
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    d = _f32(dists)
    win = np.zeros(n1, dtype=np.int32)
    o.orc_dedupe(_p(idx), _p(d), len(idx), int(n1), _p(win))
    return win


# ---------------------------------------------------------------------------------------------------------
# reference (its own sources, compiled in place)
# ---------------------------------------------------------------------------------------------------------
def ref_generate_mesh(frame: dict, bounds, with_triangles: bool = False):
    """-> (vertices, per_map_counts), plus triangles int32[nt,3] as a third value when with_triangles."""
    r = ref_native()
    S = int(frame["n_maps"])
    w = np.ascontiguousarray(frame["widths"], dtype=np.int32)
    h = np.ascontiguousarray(frame["heights"], dtype=np.int32)
    d = np.array(frame["depth_maps"], dtype=np.uint8, order="C")
    c = np.array(frame["depth_colors"], dtype=np.uint8, order="C")
    ip, wt = _f32(frame["intr"]).copy(), _f32(frame["wt"]).copy()
    b = [C.c_float(float(x)) for x in bounds]
    mesh = RefMesh()
    counts = np.zeros(S, dtype=np.int32)
    r.ref_generate_mesh(S, _p(d), _p(c), _p(w), _p(h), _p(ip), _p(wt), C.byref(mesh), *b, 1 if with_triangles else 0, _p(counts))
    out = np.empty(mesh.nVertices, dtype=VERTEX_DTYPE)
    if mesh.nVertices:
        C.memmove(out.ctypes.data, mesh.vertices, mesh.nVertices * 16)
    tri = np.empty(3 * mesh.nTriangles, dtype=np.int32)
    if mesh.nTriangles:
        C.memmove(tri.ctypes.data, mesh.triangles, mesh.nTriangles * 12)
    r.deleteMesh(C.byref(mesh))
    if with_triangles:
        return out, counts, tri.reshape(-1, 3)
    return out, counts


def ref_generate_vertices_from_depth_map(frame: dict, bounds, index: int):
    """The reference's own export generateVerticesFromDepthMap (depthprocessing.cpp:1631-1657)."""
    r = ref_native()
    w = np.ascontiguousarray(frame["widths"], dtype=np.int32)
    h = np.ascontiguousarray(frame["heights"], dtype=np.int32)
    d = np.array(frame["depth_maps"], dtype=np.uint8, order="C")
    c = np.array(frame["depth_colors"], dtype=np.uint8, order="C")
    ip, wt = _f32(frame["intr"]).copy(), _f32(frame["wt"]).copy()
    b = [C.c_float(float(x)) for x in bounds]
    mesh = RefMesh()
    r.generateVerticesFromDepthMap(_p(d), _p(c), _p(w), _p(h), _p(ip), _p(wt), C.byref(mesh), *b, int(index))
    out = np.empty(mesh.nVertices, dtype=VERTEX_DTYPE)
    if mesh.nVertices:
        C.memmove(out.ctypes.data, mesh.vertices, mesh.nVertices * 16)
    r.deleteMesh(C.byref(mesh))
    return out


def ref_vertex_maps(depth_u16, colors, w, h, intr7, wt12, bounds):
    r = ref_native()
    d = np.array(depth_u16, dtype=np.uint16, order="C")
    c = np.array(colors, dtype=np.uint8, order="C")
    d2v = np.zeros(w * h, dtype=np.int32)
    v2d = np.zeros(w * h, dtype=np.int32)
    ip, wt = _f32(intr7).copy(), _f32(wt12).copy()
    b = [C.c_float(float(x)) for x in bounds]
    n = r.ref_vertex_maps(_p(d), _p(c), int(w), int(h), _p(ip), _p(wt), *b, _p(d2v), _p(v2d))
    return n, d2v, v2d[:n].copy()


def ref_find_closest(v1, v2):
    r = ref_native()
    v1 = np.array(v1, dtype=np.float32, order="C").reshape(-1, 3)
    v2 = np.array(v2, dtype=np.float32, order="C").reshape(-1, 3)
    idx = np.zeros(len(v2), dtype=np.uint64)
    d = np.zeros(len(v2), dtype=np.float32)
    r.ref_find_closest(_p(v1), len(v1), _p(v2), len(v2), _p(idx), _p(d))
    return idx, d


def ref_icp(v1, v2, R=None, t=None, max_iter: int = 10):
    """The reference's own export ICP (icp.cpp:75-177) over the mini-cv shim."""
    r = ref_native()
    v1 = np.array(v1, dtype=np.float32, order="C").reshape(-1, 3)
    v2 = np.array(v2, dtype=np.float32, order="C").reshape(-1, 3)
    Rm = np.array(np.eye(3) if R is None else R, dtype=np.float32, order="C").reshape(9)
    tv = np.array(np.zeros(3) if t is None else t, dtype=np.float32, order="C").reshape(3)
    r.ICP(_p(v1), _p(v2), len(v1), len(v2), _p(Rm), _p(tv), int(max_iter))
    return v2, Rm.reshape(3, 3), tv


def ref_filter(verts, colors, k: int, max_dist: float):
    r = ref_filter_lib()
    v = np.array(verts, dtype=np.float32, order="C").reshape(-1, 3)
    c = np.array(colors, dtype=np.uint8, order="C").reshape(-1, 4)
    m = np.zeros(len(v), dtype=np.int32)
    n = r.ref_filter(_p(v), _p(c), len(v), int(k), C.c_float(max_dist), _p(m))
    return v[:n].copy(), c[:n].copy(), m


def ref_filter_flying_pixels(depth_u16, w: int, h: int, k: int = 1, thr: float = 10.0, max_non_fitting: int = 0):
    """The reference's own KinectCapture::filterFlyingPixels (kinectCapture.cpp:132-174, compiled in place into oracle/_ref)."""
    r = ref_filter_lib()
    d = np.array(depth_u16, dtype=np.uint16, order="C").reshape(-1)
    r.ref_filter_flying_pixels(_p(d), int(w), int(h), int(k), C.c_float(thr), int(max_non_fitting))
    return d


def ref_knn_kdist(v, k: int):
    r = ref_filter_lib()
    v = np.array(v, dtype=np.float32, order="C").reshape(-1, 3)
    out = np.zeros(len(v), dtype=np.float32)
    r.ref_knn_kdist(_p(v), len(v), int(k), _p(out))
    return out
