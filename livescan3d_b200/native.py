"""ctypes binding of libls3d_b200.so (the C ABI declared in include/ls3d.h).

This is the ONLY way Python reaches the hot path: there is no Python/torch/CPU implementation behind it.
If the shared library is missing, or no B200-class CUDA device is usable, every compute entry fails loudly
(ImportError at load, Ls3dError at call time) — nothing silently falls back.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LS3D_B200_LIB") or os.path.join(_HERE, "libls3d_b200.so")     # the override is for A/B builds of the same library


class Ls3dError(RuntimeError):
    pass


class Mesh(C.Structure):
    """depthprocessing.h:42-48"""
    _fields_ = [("nVertices", C.c_int), ("vertices", C.c_void_p), ("nTriangles", C.c_int), ("triangles", C.c_void_p)]


class IcpTrace(C.Structure):
    """Ls3dIcpTrace (include/ls3d.h) == OrcIcpTrace (oracle/ls3d_oracle.cpp)"""
    _fields_ = [("n_matched", C.c_int), ("n_accepted", C.c_int), ("sigma", C.c_float), ("T", C.c_float * 3), ("Rk", C.c_float * 9)]


class ClientFrameInfo(C.Structure):
    """Ls3dClientFrameInfo (include/ls3d.h)"""
    _fields_ = [("payload_bytes", C.c_int), ("compressed", C.c_int), ("width", C.c_int), ("height", C.c_int), ("n_bodies", C.c_int), ("raw_bytes", C.c_longlong)]


class FramesInfo(C.Structure):
    """Ls3dFramesInfo (include/ls3d.h)"""
    _fields_ = [("n_maps", C.c_int), ("depth_maps", C.c_void_p), ("depth_colors", C.c_void_p), ("widths", C.c_void_p), ("heights", C.c_void_p),
                ("intr_params", C.c_void_p), ("wtransform_params", C.c_void_p)]


# every symbol include/ls3d.h declares: (restype, argtypes)
_vp, _i, _f, _ip, _fp = C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_float)
_SIG = {
    "ICP": (_f, [_vp, _vp, _i, _i, _vp, _vp, _i]),
    "generateVerticesFromDepthMap": (None, [_vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(Mesh), _f, _f, _f, _f, _f, _f, _i]),
    "generateMeshFromDepthMaps": (None, [_i, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(Mesh), _i, _f, _f, _f, _f, _f, _f, _i]),
    "depthMapAndColorSetRadialCorrection": (None, [_i, _vp, _vp, _vp, _vp, _vp]),
    "ls3d_filter_flying_pixels": (_i, [_vp, _i, _i, _i, _f, _i]),
    "ls3d_radial_correction_device": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ls3d_filter_flying_pixels_device": (_i, [_vp, _vp, _i, _i, _i, _f, _vp]),
    "createMesh": (C.POINTER(Mesh), []),
    "deleteMesh": (None, [C.POINTER(Mesh)]),
    "ls3d_filter": (_i, [_vp, _vp, _i, _i, _f, _vp]),
    "ls3d_frame_pipeline": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(Mesh), _f, _f, _f, _f, _f, _f, _i, _f, _vp]),
    "ls3d_find_closest": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "ls3d_icp_trace": (_f, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "ls3d_last_error": (C.c_char_p, []),
    "ls3d_version": (C.c_char_p, []),
    "ls3d_selftest": (_i, []),
    "ls3d_launch_count": (C.c_longlong, []),
    "ls3d_reset_launch_count": (None, []),
    "ls3d_client_frame_header": (_i, [_vp, C.c_longlong, C.POINTER(ClientFrameInfo)]),
    "ls3d_client_frame_unpack": (_i, [_vp, C.c_longlong, _vp, _vp, _vp, C.c_longlong, C.POINTER(ClientFrameInfo)]),
    "ls3d_client_frame_pack": (C.c_longlong, [_vp, _vp, _i, _i, _vp, C.c_longlong, _i, _vp, C.c_longlong]),
    "ls3d_frames_info_store": (_i, [C.c_char_p, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ls3d_frames_info_load": (_i, [C.c_char_p, C.POINTER(FramesInfo)]),
    "ls3d_frames_info_free": (None, [C.POINTER(FramesInfo)]),
    "ls3d_ply_binary_size": (C.c_longlong, [_i, _i]),
    "ls3d_write_ply_binary": (C.c_longlong, [_vp, _i, _vp, _i, _vp, C.c_longlong]),
    "ls3d_pack_ply_body_device": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "ls3d_ply_ascii_bound": (C.c_longlong, [_i, _i]),
    "ls3d_write_ply_ascii": (C.c_longlong, [_vp, _i, _vp, _i, _vp, C.c_longlong]),
    "ls3d_transfer_frame_size": (C.c_longlong, [_i, _i, _i]),
    "ls3d_set_transfer_chunk_limit": (_i, [_i]),
    "ls3d_write_transfer_frame": (C.c_longlong, [_vp, _i, _vp, _i, _vp, C.c_longlong]),
    "ls3d_transfer_chunks_device": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "ls3d_pack_transfer_body_device": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "ls3d_frame_create": (_vp, [_i, _vp, _vp]),
    "ls3d_frame_destroy": (None, [_vp]),
    "ls3d_frame_set_params": (_i, [_vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _i, _f, _vp]),
    "ls3d_frame_run": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "ls3d_frame_run_to": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ls3d_frame_run_peers": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ls3d_frame_run_count": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "ls3d_frame_merge_peers": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "ls3d_dev_alloc": (_vp, [C.c_ulonglong]),
    "ls3d_dev_free": (None, [_vp]),
    "ls3d_ipc_export": (_i, [_vp, _vp]),
    "ls3d_ipc_open": (_vp, [_vp]),
    "ls3d_ipc_close": (None, [_vp]),
    "ls3d_frame_set_filter_mode": (_i, [_vp, _i]),
    "ls3d_set_default_filter_mode": (_i, [_i]),
    "ls3d_frame_enable_timing": (None, [_vp, _i]),
    "ls3d_frame_stage_ms": (_i, [_vp, _vp]),
    "ls3d_frame_vertices": (_vp, [_vp]),
    "ls3d_frame_culled_vertices": (_vp, [_vp]),
    "ls3d_frame_count_ptr": (_vp, [_vp]),
    "ls3d_frame_sensor_starts": (_vp, [_vp]),
    "ls3d_frame_culled_starts": (_vp, [_vp]),
    "ls3d_frame_old_to_new": (_vp, [_vp]),
    "ls3d_frame_keep_mask": (_vp, [_vp]),
    "ls3d_frame_depth_to_vertex": (_vp, [_vp]),
    "ls3d_frame_enable_triangles": (None, [_vp, _i]),
    "ls3d_frame_triangles": (_vp, [_vp]),
    "ls3d_frame_triangle_starts": (_vp, [_vp]),
    "ls3d_icp_create": (_vp, [_i, _i]),
    "ls3d_icp_destroy": (None, [_vp]),
    "ls3d_icp_set_target": (_i, [_vp, _vp, _i, _vp]),
    "ls3d_icp_set_source": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ls3d_icp_match": (_i, [_vp, _vp]),
    "ls3d_frame_publish_counts": (_i, [_vp, _i, _i, _vp, _vp]),
    "ls3d_frame_wait_peers": (_i, [_i, _i, _vp, _vp]),
    "ls3d_icp_reduce": (_i, [_vp, _vp]),
    "ls3d_icp_set_peers": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "ls3d_icp_red_part": (_vp, [_vp]),
    "ls3d_icp_red_flag": (_vp, [_vp]),
    "ls3d_icp_finish": (_i, [_vp, _vp]),
    "ls3d_icp_run": (_i, [_vp, _i, _vp]),
    "ls3d_icp_set_debug": (None, [_vp, _vp]),
    "ls3d_icp_slots": (_vp, [_vp]),
    "ls3d_icp_Rt": (_vp, [_vp]),
    "ls3d_icp_nn_index": (_vp, [_vp]),
    "ls3d_icp_nn_dist": (_vp, [_vp]),
    "ls3d_icp_trace_buf": (_vp, [_vp]),
    "ls3d_icp_status": (_vp, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and type every entry point.  Raises ImportError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  livescan3d_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIG.items():
        fn = getattr(lib, name)          # AttributeError here == header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIG)


def last_error() -> str:
    return load().ls3d_last_error().decode("utf-8", "replace")


def check(ok: bool, what: str):
    if not ok:
        raise Ls3dError(f"{what}: {last_error() or 'failed'}")
