"""CPU restatement of the data formats either side of the path.  TEST INFRASTRUCTURE ONLY: imported by tests/
(and bench.py's cpu_baseline leg) as the checker; nothing under livescan3d_b200/ imports it.

Each function follows the reference code it cites line by line, in plain Python / numpy:
  * client frame blob   SerializeFrame, src/LiveScanClient/liveScanClient.cpp:185-290; receiver LiveScanServer/KinectSocket.cs:211-304
  * frames dump         storeAllFramesInformation / loadAllFramesInformation, src/NativeUtils/depthprocessing.cpp:1316-1385
  * binary and ASCII PLY Utils.saveToPly, LiveScanServer/Utils.cs:173-293
  * transfer frame      formVerticesChunks / formMeshChunks, LiveScanServer/TransferServer.cs:179-271; SendFrame, TransferSocket.cs:50-105

Pinning: the frames dump is checked against the reference's own C++ functions compiled into oracle/_ref (tests/test_formats.py).
The other three are C# / Win32 code that cannot run here and the reference ships no sample files for them: PARITY UNPINNED by
fixtures — they are restated from the source text only.  zstd is the system libzstd through ctypes (the reference links zstd 1.1.3;
any zstd decoder reads any zstd frame, so only round trips are asserted on compressed blobs, never compressed bytes).
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

CHUNK_LIMIT = 65000 - 3          # TransferServer.cs:181,205


# --------------------------------------------------------------------------------------------------------- zstd
def _zstd():
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    z.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
    z.ZSTD_isError.restype = C.c_uint
    z.ZSTD_isError.argtypes = [C.c_size_t]
    return z


def zstd_compress(data: bytes, level: int) -> bytes:
    z = _zstd()
    cap = z.ZSTD_compressBound(len(data)) * 2                      # liveScanClient.cpp:272
    out = C.create_string_buffer(cap)
    n = z.ZSTD_compress(out, cap, data, len(data), level)
    assert not z.ZSTD_isError(n)
    return out.raw[:n]


def zstd_decompress(data: bytes) -> bytes:
    z = _zstd()
    want = z.ZSTD_getFrameContentSize(data, len(data))
    out = C.create_string_buffer(max(int(want), 1))
    n = z.ZSTD_decompress(out, want, data, len(data))
    assert not z.ZSTD_isError(n) and n == want
    return out.raw[:n]


# --------------------------------------------------------------------------------------------------------- client frame blob
def bodies_bytes(bodies) -> bytes:
    """bodies: list of (tracked: bool, joints: list of (type, state, x, y, z, cx, cy)) -> liveScanClient.cpp:233-264."""
    b = struct.pack("<i", len(bodies))
    for tracked, joints in bodies:
        b += struct.pack("<?i", tracked, len(joints))
        for (jt, st, x, y, z, cx, cy) in joints:
            b += struct.pack("<iifffff", jt, st, x, y, z, cx, cy)
    return b


def orc_client_frame_pack(depth_u16: np.ndarray, colors_rgb: np.ndarray, bodies=(), compression_level: int = 2) -> bytes:
    """liveScanClient.cpp:185-290 after the colour mapping (the caller hands over per-depth-pixel colours, with depth already zeroed
    where the mapping fell outside the colour image, :219-228)."""
    h, w = depth_u16.shape
    payload = np.ascontiguousarray(depth_u16, "<u2").tobytes() + np.ascontiguousarray(colors_rgb, np.uint8).tobytes() + bodies_bytes(list(bodies))
    compressed = 1 if compression_level > 0 else 0                 # :267, :629-634
    if compressed:
        payload = zstd_compress(payload, compression_level)       # :268-279
    return struct.pack("<iiii", len(payload), compressed, w, h) + payload      # :283-288


def orc_client_frame_unpack(blob: bytes):
    """KinectSocket.cs:211-304 -> (depth [h,w] u16, colours [h,w,3] u8, bodies list, raw body bytes)."""
    n_to_read, compressed, w, h = struct.unpack_from("<iiii", blob, 0)          # :225-239
    assert n_to_read > 0
    buf = blob[16:16 + n_to_read]
    if compressed == 1:
        buf = zstd_decompress(buf)                                 # :245-246
    depth = np.frombuffer(buf, "<u2", w * h, 0).reshape(h, w).copy()            # :254
    colors = np.frombuffer(buf, np.uint8, 3 * w * h, 2 * w * h).reshape(h, w, 3).copy()     # :255
    pos = start = w * h * 5                                         # :258
    (nb,) = struct.unpack_from("<i", buf, pos)
    pos += 4
    bodies = []
    for _ in range(nb):
        tracked, nj = struct.unpack_from("<?i", buf, pos)
        pos += 5
        joints = []
        for _ in range(nj):
            joints.append(struct.unpack_from("<iifffff", buf, pos))
            pos += 28
        bodies.append((tracked, joints))
    return depth, colors, bodies, buf[start:pos]


# --------------------------------------------------------------------------------------------------------- frames dump
def orc_frames_info_bytes(frame: dict) -> bytes:
    """depthprocessing.cpp:1316-1341."""
    n = int(frame["n_maps"])
    w = np.asarray(frame["widths"], "<i4")
    h = np.asarray(frame["heights"], "<i4")
    out = struct.pack("<i", n)
    if n > 0:
        out += w.tobytes() + h.tobytes()
    d = np.asarray(frame["depth_maps"], np.uint8)
    c = np.asarray(frame["depth_colors"], np.uint8)
    pd = pc = 0
    for i in range(n):
        px = int(w[i]) * int(h[i])
        out += d[pd:pd + 2 * px].tobytes() + c[pc:pc + 3 * px].tobytes()
        pd += 2 * px
        pc += 3 * px
    out += np.asarray(frame["intr"], "<f4").tobytes()[: 28 * n] + np.asarray(frame["wt"], "<f4").tobytes()[: 48 * n]
    return out


def orc_frames_info_parse(data: bytes) -> dict:
    """depthprocessing.cpp:1343-1385."""
    (n,) = struct.unpack_from("<i", data, 0)
    pos = 4
    w = np.frombuffer(data, "<i4", n, pos).copy(); pos += 4 * n
    h = np.frombuffer(data, "<i4", n, pos).copy(); pos += 4 * n
    ds, cs = [], []
    for i in range(n):
        px = int(w[i]) * int(h[i])
        ds.append(np.frombuffer(data, np.uint8, 2 * px, pos)); pos += 2 * px
        cs.append(np.frombuffer(data, np.uint8, 3 * px, pos)); pos += 3 * px
    intr = np.frombuffer(data, "<f4", 7 * n, pos).copy(); pos += 28 * n
    wt = np.frombuffer(data, "<f4", 12 * n, pos).copy(); pos += 48 * n
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
    return {"n_maps": n, "widths": w, "heights": h, "depth_maps": cat(ds), "depth_colors": cat(cs), "intr": intr, "wt": wt}


# --------------------------------------------------------------------------------------------------------- binary PLY
def orc_ply_binary(vertices: np.ndarray, triangles) -> bytes:
    """Utils.cs:222-293 with binary=true (triangles None: the overload at :173-220, fed with this cloud's xyz and colours).
    StreamWriter.WriteLine terminates with Environment.NewLine = "\\r\\n" on the Windows hosts the server runs on (:187, :236)."""
    v = np.asarray(vertices)
    head = "ply\nformat binary_little_endian 1.0" + "\r\n"
    head += "element vertex " + str(len(v)) + "\n"
    head += "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
    if triangles is not None:
        t = np.asarray(triangles, "<i4").reshape(-1, 3)
        head += "element face " + str(len(t)) + "\n"
        head += "property list uchar int vertex_index\n"
    head += "end_header\n"
    rec = np.zeros(len(v), dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")]))      # :257-266
    rec["x"], rec["y"], rec["z"], rec["r"], rec["g"], rec["b"] = v["X"], v["Y"], v["Z"], v["R"], v["G"], v["B"]
    body = rec.tobytes()
    if triangles is not None:
        f = np.zeros(len(t), dtype=np.dtype([("n", "u1"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4")]))                                  # :268-274
        f["n"], f["a"], f["b"], f["c"] = 3, t[:, 0], t[:, 1], t[:, 2]
        body += f.tobytes()
    return head.encode("ascii") + body


# --------------------------------------------------------------------------------------------------------- ASCII PLY
def net45_single_to_string(v) -> str:
    """Single.ToString(CultureInfo.InvariantCulture) on .NET Framework 4.5 (the server's target, LiveScanServer.csproj:12): the
    general format with 7 significant digits.  Restated with exact decimal arithmetic: the float's exact value is rounded half away
    from zero to 7 digits (the CLR goes through the C runtime's _ecvt); fixed notation when the decimal exponent e of the ROUNDED
    number satisfies -5 < e < 7, else d.ddddddE+XX with at least two exponent digits; trailing zeros dropped; both zeros print "0".
    No .NET runtime is available here: parity unpinned, this follows the documented behaviour."""
    import decimal
    import math
    f = float(np.float32(v))
    if math.isnan(f):
        return "NaN"
    if math.isinf(f):
        return "Infinity" if f > 0 else "-Infinity"
    if f == 0.0:
        return "0"
    d = decimal.Decimal(f)                       # exact
    sign, a = ("-" if d < 0 else ""), abs(d)
    e = a.adjusted()
    with decimal.localcontext() as ctx:
        ctx.prec = 200
        q = (a.scaleb(6 - e)).to_integral_value(rounding=decimal.ROUND_HALF_UP)      # 7-digit integer
    q = int(q)
    if q == 10 ** 7:
        q, e = 10 ** 6, e + 1
    digits = str(q).rstrip("0") or "0"
    if -5 < e < 7:
        if e >= 0:
            ip = (digits + "0" * 7)[:e + 1]
            fp = digits[e + 1:]
            return sign + ip + ("." + fp if fp else "")
        return sign + "0." + "0" * (-e - 1) + digits
    mant = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
    return sign + mant + "E" + ("-" if e < 0 else "+") + "%02d" % abs(e)


def orc_ply_ascii(vertices: np.ndarray, triangles) -> bytes:
    """Utils.cs:222-293 with binary=false (triangles None: the overload at :173-220).  WriteLine = text + "\r\n"; the header's first
    WriteLine is given a string that already ends in "\n" (:187/:236); vertex lines end in a blank (:207-211, :279-284); face
    lines are "3 " followed by the three indices with nothing between them (:285-290) — reproduced as written."""
    v = np.asarray(vertices)
    out = ["ply\nformat ascii 1.0\n" + "\r\n", "element vertex " + str(len(v)) + "\n",
           "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"]
    t = None
    if triangles is not None:
        t = np.asarray(triangles, "<i4").reshape(-1, 3)
        out.append("element face " + str(len(t)) + "\n")
        out.append("property list uchar int vertex_index\n")
    out.append("end_header\n")
    for r in v:
        s = ""
        for k in ("X", "Y", "Z"):
            s += net45_single_to_string(r[k]) + " "
        for k in ("R", "G", "B"):
            s += str(int(r[k])) + " "
        out.append(s + "\r\n")
    if t is not None:
        for a, b, c in t:
            out.append("3 " + str(int(a)) + str(int(b)) + str(int(c)) + "\r\n")
    return "".join(out).encode("ascii")


# --------------------------------------------------------------------------------------------------------- transfer frame
def orc_form_vertices_chunks(n_vertices: int):
    """TransferServer.cs:179-201."""
    vs, ts, cur = [], [], 0
    while cur < n_vertices:
        size = min(CHUNK_LIMIT, n_vertices - cur)
        vs.append(size); ts.append(0)
        cur += size
    return vs, ts


def orc_form_mesh_chunks(vertices: np.ndarray, triangles: np.ndarray, limit: int = CHUNK_LIMIT):
    """TransferServer.cs:203-271, statement by statement (including trianglesChunkStart = t at :250, which makes the first of several
    chunks report one triangle too few).  -> (new_vertices, new_triangles flat, vertices_in_chunks, triangles_in_chunks)."""
    tl = np.asarray(triangles, np.int64).reshape(-1)
    n_vertices, n_tri = len(vertices), len(tl) // 3
    chunk_index = np.full(n_vertices, -1, np.int64)
    vertices_map = np.zeros(n_vertices, np.int64)
    new_vertices = np.zeros(n_tri * 3, dtype=vertices.dtype)
    new_triangles = np.zeros(n_tri * 3, np.int32)
    tri_in, ver_in = [], []
    tri_chunk_start = cur_chunk = cur_vertex = in_chunk = 0
    for t in range(n_tri * 3):
        val = int(tl[t])
        if chunk_index[val] != cur_chunk:
            new_vertices[cur_vertex] = vertices[val]
            vertices_map[val] = in_chunk
            chunk_index[val] = cur_chunk
            cur_vertex += 1
            new_triangles[t] = in_chunk
            in_chunk += 1
        else:
            new_triangles[t] = vertices_map[val]
        if in_chunk >= limit and ((t + 1) % 3) == 0:
            cur_chunk += 1
            ver_in.append(in_chunk)
            tri_in.append((t - tri_chunk_start) // 3)
            in_chunk = 0
            tri_chunk_start = t
    if in_chunk != 0:
        ver_in.append(in_chunk)
        tri_in.append((n_tri * 3 - tri_chunk_start) // 3)
    return new_vertices[:cur_vertex], new_triangles, ver_in, tri_in


def orc_transfer_frame(vertices: np.ndarray, triangles, limit: int = CHUNK_LIMIT) -> bytes:
    """TransferServer.cs:131-156 (which chunker) + TransferSocket.cs:50-105 (the bytes)."""
    v = np.asarray(vertices)
    t = np.zeros(0, np.int32) if triangles is None else np.asarray(triangles, np.int32).reshape(-1)
    if len(t) > 0:
        v, t, vs, ts = orc_form_mesh_chunks(v, t, limit)
    else:
        vs, ts = orc_form_vertices_chunks(len(v))
    xyz = np.stack([v["X"], v["Y"], v["Z"]], axis=1).astype("<f4") if len(v) else np.zeros((0, 3), "<f4")
    rgb = np.stack([v["R"], v["G"], v["B"]], axis=1).astype(np.uint8) if len(v) else np.zeros((0, 3), np.uint8)
    return (struct.pack("<iii", len(v), len(t) // 3, len(vs)) + np.asarray(vs, "<i4").tobytes() + np.asarray(ts, "<i4").tobytes()
            + xyz.tobytes() + rgb.tobytes() + np.asarray(t, "<i4").tobytes())
