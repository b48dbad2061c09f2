"""Host mirror of the format entry points (include/ls3d.h Part 4): the client's frame blob, NativeUtils' frames
dump, binary PLY and the TransferServer mesh frame.  Thin ctypes wrappers — the work is in libls3d_b200.so (csrc/formats.cu)."""
import ctypes as C

import numpy as np

from . import native
from .native import ClientFrameInfo, FramesInfo, Ls3dError

_ptr = lambda a: C.c_void_p(a.ctypes.data)


def client_frame_pack(depth_u16: np.ndarray, colors_rgb: np.ndarray, bodies: bytes | None = None, compression_level: int = 2) -> bytes:
    """SerializeFrame (liveScanClient.cpp:185-290) for a depth image [h, w] u16 and its per-depth-pixel colours [h, w, 3] u8."""
    lib = native.load()
    d = np.ascontiguousarray(depth_u16, np.uint16)
    c = np.ascontiguousarray(colors_rgb, np.uint8)
    h, w = d.shape
    assert c.size == 3 * d.size
    b = (C.c_ubyte * len(bodies)).from_buffer_copy(bodies) if bodies else None
    nb = len(bodies) if bodies else 0
    cap = lib.ls3d_client_frame_pack(_ptr(d), _ptr(c), w, h, b, nb, int(compression_level), None, 0)
    native.check(cap >= 0, "ls3d_client_frame_pack")
    out = np.empty(cap, np.uint8)
    n = lib.ls3d_client_frame_pack(_ptr(d), _ptr(c), w, h, b, nb, int(compression_level), _ptr(out), cap)
    native.check(n >= 0, "ls3d_client_frame_pack")
    return out[:n].tobytes()


def client_frame_unpack(blob: bytes):
    """KinectSocket.ReceiveFrame (KinectSocket.cs:211-304): -> (depth [h, w] u16, colours [h, w, 3] u8, body records bytes, info)."""
    lib = native.load()
    buf = np.frombuffer(blob, np.uint8)
    info = ClientFrameInfo()
    native.check(lib.ls3d_client_frame_header(_ptr(buf), len(buf), C.byref(info)) == 0, "ls3d_client_frame_header")
    depth = np.empty((info.height, info.width), np.uint16)
    colors = np.empty((info.height, info.width, 3), np.uint8)
    bodies = np.empty(max(len(buf) * 64, 1 << 16), np.uint8)
    n = lib.ls3d_client_frame_unpack(_ptr(buf), len(buf), _ptr(depth), _ptr(colors), _ptr(bodies), len(bodies), C.byref(info))
    native.check(n >= 0, "ls3d_client_frame_unpack")
    return depth, colors, bodies[:n].tobytes(), info


def frames_info_store(filename: str, frame: dict):
    lib = native.load()
    a = [np.ascontiguousarray(frame[k], t) for k, t in (("depth_maps", np.uint8), ("depth_colors", np.uint8), ("widths", np.int32), ("heights", np.int32),
                                                        ("intr", np.float32), ("wt", np.float32))]
    native.check(lib.ls3d_frames_info_store(filename.encode(), int(frame["n_maps"]), *[_ptr(x) for x in a]) == 0, "ls3d_frames_info_store")


def frames_info_load(filename: str) -> dict:
    lib = native.load()
    fi = FramesInfo()
    native.check(lib.ls3d_frames_info_load(filename.encode(), C.byref(fi)) == 0, "ls3d_frames_info_load")
    try:
        n = fi.n_maps
        take = lambda p, cnt, t: np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(cnt,)).copy() if cnt else np.zeros(0, t)
        w, h = take(fi.widths, n, C.c_int32), take(fi.heights, n, C.c_int32)
        px = int((w.astype(np.int64) * h).sum())
        return {"n_maps": n, "widths": w, "heights": h, "depth_maps": take(fi.depth_maps, 2 * px, C.c_uint8), "depth_colors": take(fi.depth_colors, 3 * px, C.c_uint8),
                "intr": take(fi.intr_params, 7 * n, C.c_float), "wt": take(fi.wtransform_params, 12 * n, C.c_float)}
    finally:
        lib.ls3d_frames_info_free(C.byref(fi))


def write_ply_binary(vertices: np.ndarray, triangles: np.ndarray | None) -> bytes:
    """Utils.saveToPly(..., binary=true) (Utils.cs:222-293; triangles=None: the vertex-only overload :173-220) -> the file's bytes."""
    lib = native.load()
    v = np.ascontiguousarray(vertices)
    assert v.dtype.itemsize == 16
    t = None if triangles is None else np.ascontiguousarray(triangles, np.int32).reshape(-1, 3)
    nt = -1 if t is None else len(t)
    size = lib.ls3d_ply_binary_size(len(v), nt)
    out = np.empty(size, np.uint8)
    n = lib.ls3d_write_ply_binary(_ptr(v), len(v), None if t is None or not len(t) else _ptr(t), nt, _ptr(out), size)
    native.check(n == size, "ls3d_write_ply_binary")
    return out.tobytes()


def write_ply_ascii(vertices: np.ndarray, triangles: np.ndarray | None) -> bytes:
    """Utils.saveToPly(..., binary=false) (Utils.cs:204-214, :276-289) -> the file's bytes (host text codec)."""
    lib = native.load()
    v = np.ascontiguousarray(vertices)
    assert v.dtype.itemsize == 16
    t = None if triangles is None else np.ascontiguousarray(triangles, np.int32).reshape(-1, 3)
    nt = -1 if t is None else len(t)
    cap = lib.ls3d_ply_ascii_bound(len(v), nt)
    out = np.empty(cap, np.uint8)
    n = lib.ls3d_write_ply_ascii(_ptr(v) if len(v) else None, len(v), None if t is None or not len(t) else _ptr(t), nt, _ptr(out), cap)
    native.check(n >= 0, "ls3d_write_ply_ascii")
    return out[:n].tobytes()


def write_transfer_frame(vertices: np.ndarray, triangles: np.ndarray | None) -> bytes:
    """formVerticesChunks / formMeshChunks + TransferSocket.SendFrame (TransferServer.cs:179-271, TransferSocket.cs:50-105)."""
    lib = native.load()
    v = np.ascontiguousarray(vertices)
    assert v.dtype.itemsize == 16
    t = np.zeros((0, 3), np.int32) if triangles is None else np.ascontiguousarray(triangles, np.int32).reshape(-1, 3)
    args = (_ptr(v) if len(v) else None, len(v), _ptr(t) if len(t) else None, len(t))
    size = lib.ls3d_write_transfer_frame(*args, None, 0)
    native.check(size >= 0, "ls3d_write_transfer_frame")
    out = np.empty(size, np.uint8)
    n = lib.ls3d_write_transfer_frame(*args, _ptr(out), size)
    native.check(n == size, "ls3d_write_transfer_frame")
    return out.tobytes()


__all__ = ["client_frame_pack", "client_frame_unpack", "frames_info_store", "frames_info_load", "write_ply_binary", "write_ply_ascii", "write_transfer_frame", "Ls3dError"]
