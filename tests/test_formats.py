"""SURVEY §8f N4 — the data formats either side of the path: client frame blob, frames dump, binary PLY, transfer frame.
CPU tests cover the host-side codecs (no device needed) against the restatement and, for the frames dump, against the
reference's own code in oracle/_ref; GPU tests cover the device re-packing and chunking, byte for byte."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from livescan3d_b200 import formats, native, synth  # noqa: E402
from oracle import formats_oracle as fo  # noqa: E402
from oracle import oracle_lib as orc  # noqa: E402
from common import small_frame  # noqa: E402

VERTEX = orc.VERTEX_DTYPE


def _sensor(fr, i):
    w, h = int(fr["widths"][i]), int(fr["heights"][i])
    px = w * h
    d = fr["depth_maps"].view(np.uint16)[i * px:(i + 1) * px].reshape(h, w)
    c = fr["depth_colors"][3 * i * px:3 * (i + 1) * px].reshape(h, w, 3)
    return d, c


BODIES = [(True, [(3, 2, 0.1, -0.2, 1.5, 200.5, 100.25), (7, 1, 0.4, 0.3, 1.25, 50.0, 60.0)]), (False, [])]


# ------------------------------------------------------------------------------------------------ CPU: codecs
@pytest.mark.parametrize("level", [0, 2, 5])
def test_client_frame_blob_matches_restatement(level):
    fr = small_frame(S=2, w=64, h=48)
    d, c = _sensor(fr, 1)
    bb = fo.bodies_bytes(BODIES)
    want = fo.orc_client_frame_pack(d, c, BODIES, level)
    got = formats.client_frame_pack(d, c, bb, level)
    if level == 0:
        assert got == want                                   # uncompressed blobs are byte-identical
    assert got[:4] == np.int32(len(got) - 16).tobytes() and got[4:16] == want[4:16]
    # each side reads the other's blob
    for blob in (got, want):
        gd, gc, gb, info = formats.client_frame_unpack(blob)
        od, oc, obodies, ob = fo.orc_client_frame_unpack(blob)
        assert np.array_equal(gd, d) and np.array_equal(gc, c) and gb == bb == ob
        assert np.array_equal(od, d) and np.array_equal(oc, c) and obodies[0][1][1][0] == 7
        assert (info.width, info.height, info.n_bodies, info.compressed) == (64, 48, 2, 1 if level else 0)
        assert info.raw_bytes == 5 * 64 * 48 + len(bb)


def test_client_frame_blob_without_bodies_and_bad_input():
    d = np.arange(12, dtype=np.uint16).reshape(3, 4)
    c = np.arange(36, dtype=np.uint8).reshape(3, 4, 3)
    blob = formats.client_frame_pack(d, c, None, 0)
    assert blob == fo.orc_client_frame_pack(d, c, (), 0)
    gd, gc, gb, info = formats.client_frame_unpack(blob)
    assert np.array_equal(gd, d) and np.array_equal(gc, c) and gb == b"\0\0\0\0" and info.n_bodies == 0
    lib = native.load()
    info = native.ClientFrameInfo()
    buf = np.frombuffer(blob, np.uint8)
    p = C.c_void_p(buf.ctypes.data)
    assert lib.ls3d_client_frame_header(p, 8, C.byref(info)) == -1 and b"16-byte header" in lib.ls3d_last_error()
    assert lib.ls3d_client_frame_header(p, len(buf) - 1, C.byref(info)) == -1                      # truncated payload
    end = np.zeros(16, np.uint8)                                                                   # payload size 0: "no more stored frames"
    assert lib.ls3d_client_frame_header(C.c_void_p(end.ctypes.data), 16, C.byref(info)) == -1
    bad = bytearray(formats.client_frame_pack(d, c, fo.bodies_bytes(BODIES), 0))
    bad[16 + 60 + 4 + 1:16 + 60 + 4 + 5] = np.int32(1000).tobytes()                                # a joint count the payload cannot hold
    with pytest.raises(native.Ls3dError, match="truncated body"):
        formats.client_frame_unpack(bytes(bad))
    garbage = bytes(blob[:4]) + np.int32(1).tobytes() + bytes(blob[8:])                            # flagged compressed, is not
    with pytest.raises(native.Ls3dError, match="zstd"):
        formats.client_frame_unpack(garbage)


def test_client_frame_blob_random_round_trips():
    """Property: for any image size, content, body list and compression level, pack -> unpack returns the inputs, both codecs agree on the
    uncompressed bytes, and each reads the other's blob (40 random cases, no device needed)."""
    rng = np.random.default_rng(7)
    for case in range(40):
        w, h = int(rng.integers(0, 70)), int(rng.integers(0, 50))
        d = rng.integers(0, 65536, (h, w)).astype(np.uint16)
        c = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        bodies = [(bool(rng.integers(0, 2)), [(int(rng.integers(0, 25)), int(rng.integers(0, 3)), *[float(np.float32(x)) for x in rng.normal(size=5)])
                                              for _ in range(int(rng.integers(0, 26)))]) for _ in range(int(rng.integers(0, 7)))]
        bb = fo.bodies_bytes(bodies)
        level = int(rng.choice([0, 1, 2, 9]))
        ours = formats.client_frame_pack(d, c, bb, level)
        theirs = fo.orc_client_frame_pack(d, c, bodies, level)
        if level == 0:
            assert ours == theirs, case
        for blob in (ours, theirs):
            gd, gc, gb, info = formats.client_frame_unpack(blob)
            od, oc, ob_list, ob = fo.orc_client_frame_unpack(blob)
            assert np.array_equal(gd, d) and np.array_equal(gc, c) and gb == bb == ob, case
            assert np.array_equal(od, d) and np.array_equal(oc, c) and len(ob_list) == len(bodies)
            assert (info.width, info.height, info.n_bodies) == (w, h, len(bodies))


def test_client_frames_fill_the_packed_arrays_the_path_takes():
    """KinectServer.CopyLatestFrames (KinectServer.cs:404-500): every sensor's depth / colours land back to back."""
    fr = synth.make_frame(3, 40, 30, ring=8)
    blobs = [formats.client_frame_pack(*_sensor(fr, i), None, 2 if i % 2 else 0) for i in range(3)]
    lib = native.load()
    depth = np.zeros_like(fr["depth_maps"])
    colors = np.zeros_like(fr["depth_colors"])
    px = 40 * 30
    for i, b in enumerate(blobs):
        buf = np.frombuffer(b, np.uint8)
        n = lib.ls3d_client_frame_unpack(C.c_void_p(buf.ctypes.data), len(buf), C.c_void_p(depth.ctypes.data + 2 * px * i), C.c_void_p(colors.ctypes.data + 3 * px * i), None, 0, None)
        assert n == 4, native.last_error()
    assert np.array_equal(depth, fr["depth_maps"]) and np.array_equal(colors, fr["depth_colors"])


def test_frames_dump_against_restatement_and_reference(tmp_path):
    fr = synth.make_frame(3, 50, 40, ring=8)
    fr["widths"] = np.array([50, 20, 50], np.int32)           # mixed sizes: 50x40, 20x100, 50x40 (same pixel counts)
    fr["heights"] = np.array([40, 100, 40], np.int32)
    ours = str(tmp_path / "ours.bin")
    formats.frames_info_store(ours, fr)
    data = open(ours, "rb").read()
    assert data == fo.orc_frames_info_bytes(fr)
    back = formats.frames_info_load(ours)
    for k in ("widths", "heights", "depth_maps", "depth_colors", "intr", "wt"):
        assert np.array_equal(back[k], np.asarray(fr[k]).reshape(-1)), k
    assert np.array_equal(fo.orc_frames_info_parse(data)["wt"], fr["wt"].reshape(-1))
    # empty rig, missing file, truncated file
    formats.frames_info_store(str(tmp_path / "empty.bin"), {"n_maps": 0, "depth_maps": np.zeros(0, np.uint8), "depth_colors": np.zeros(0, np.uint8),
                                                          "widths": np.zeros(0, np.int32), "heights": np.zeros(0, np.int32), "intr": np.zeros(0, np.float32), "wt": np.zeros(0, np.float32)})
    assert open(tmp_path / "empty.bin", "rb").read() == b"\0\0\0\0" and formats.frames_info_load(str(tmp_path / "empty.bin"))["n_maps"] == 0
    with pytest.raises(native.Ls3dError, match="cannot open"):
        formats.frames_info_load(str(tmp_path / "nope.bin"))
    open(tmp_path / "cut.bin", "wb").write(data[: len(data) // 2])
    with pytest.raises(native.Ls3dError, match="truncated"):
        formats.frames_info_load(str(tmp_path / "cut.bin"))
    if not orc.have_ref():
        pytest.skip("oracle/_ref not built")
    ref = orc.ref_native()
    p = lambda a: C.c_void_p(a.ctypes.data)
    theirs = str(tmp_path / "theirs.bin")
    a = [np.ascontiguousarray(fr[k]) for k in ("depth_maps", "depth_colors", "widths", "heights", "intr", "wt")]
    ref.ref_store_frames(theirs.encode(), 3, *[p(x) for x in a])                  # the reference's writer
    assert open(theirs, "rb").read() == data
    out = [np.zeros_like(x) for x in a]
    ref.ref_load_frames.restype = C.c_int
    assert ref.ref_load_frames(ours.encode(), *[p(x) for x in out]) == 3          # the reference's reader on our file
    for x, y in zip(a, out):
        assert np.array_equal(x, y)


def test_ply_and_transfer_sizes_without_a_device():
    lib = native.load()
    v = np.zeros(5, VERTEX)
    assert lib.ls3d_ply_binary_size(5, 2) == len(fo.orc_ply_binary(v, np.zeros((2, 3), np.int32)))
    assert lib.ls3d_ply_binary_size(5, -1) == len(fo.orc_ply_binary(v, None))
    assert lib.ls3d_ply_binary_size(0, 0) == len(fo.orc_ply_binary(v[:0], np.zeros((0, 3), np.int32)))
    assert lib.ls3d_transfer_frame_size(5, 2, 1) == len(fo.orc_transfer_frame(v, np.array([[0, 1, 2], [2, 3, 4]])))


def test_net45_float_text_known_answers():
    """Documented outputs of Single.ToString() on .NET Framework (general format, 7 significant digits)."""
    kat = [(0.0, "0"), (-0.0, "0"), (1.0, "1"), (-1.5, "-1.5"), (0.1, "0.1"), (1.0 / 3.0, "0.3333333"), (123456.789, "123456.8"), (1234567.0, "1234567"),
           (12345678.0, "1.234568E+07"), (1e7, "1E+07"), (9999999.0, "9999999"), (0.0001, "0.0001"), (0.00001, "1E-05"), (1.5e-5, "1.5E-05"),
           (3.4028235e38, "3.402823E+38"), (1.17549435e-38, "1.175494E-38"), (1e-45, "1.401298E-45"), (float("nan"), "NaN"),
           (float("inf"), "Infinity"), (float("-inf"), "-Infinity"), (0.5, "0.5"), (100.0, "100"), (-0.00012345678, "-0.0001234568"),
           (1234566.5, "1234567"), (1234567.5, "1234568"), (0.99999994, "0.9999999")]
    for x, want in kat:
        assert fo.net45_single_to_string(np.float32(x)) == want, (x, fo.net45_single_to_string(np.float32(x)), want)


def _awkward_floats(n, seed):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)           # every exponent, NaNs and infinities included
    f = bits.view(np.float32).copy()
    special = np.array([0.0, -0.0, 1e7, 9999999.0, 9999999.5, 99999.995, 1234566.5, 1234567.5, 8388607.5, 0.00001, 0.0000099999995, 1e-45, 3.4028235e38,
                        np.inf, -np.inf, np.nan, 0.1, 1.0 / 3.0, 16777216.0, 1048576.5, 2097151.5, 524287.25, 0.5, 0.25, 0.125], np.float32)
    f[:len(special)] = special[:min(len(special), n)]
    near = rng.normal(size=n // 2).astype(np.float32) * np.float32(2.0)                # the coordinates a cloud really has
    f[n - len(near):] = near
    return f


@pytest.mark.parametrize("n,nt", [(0, 0), (1, -1), (3, 2), (4000, 777), (60000, 30000)])
def test_ply_ascii_bytes(n, nt):
    """Host text codec: runs without a device.  Floats of every magnitude, specials, exact ties at the 8th digit."""
    v = np.zeros(n, VERTEX)
    rng = np.random.default_rng(n + 7)
    for k in "RGB":
        v[k] = rng.integers(0, 256, n)
    v["A"] = 255
    for i, k in enumerate("XYZ"):
        v[k] = _awkward_floats(n, 10 * n + i) if n else 0
    t = None if nt < 0 else rng.integers(-5, max(n, 1) + 100000, (nt, 3)).astype(np.int32)
    got = formats.write_ply_ascii(v, t)
    want = fo.orc_ply_ascii(v, t)
    assert got == want
    if nt >= 0:
        assert formats.write_ply_ascii(v, None) == fo.orc_ply_ascii(v, None)


# ------------------------------------------------------------------------------------------------ GPU: re-packing and chunking
def _random_cloud(n, seed):
    rng = np.random.default_rng(seed)
    v = np.zeros(n, VERTEX)
    for k in "RGB":
        v[k] = rng.integers(0, 256, n)
    v["A"] = 255
    for k in "XYZ":
        v[k] = rng.normal(size=n).astype(np.float32)
    return v


@pytest.mark.gpu
@pytest.mark.parametrize("n,nt", [(0, 0), (1, 0), (1, 1), (17, 5), (1000, 1999), (4099, 3)])
def test_ply_binary_bytes(n, nt):
    v = _random_cloud(n, n + nt)
    t = np.random.default_rng(nt).integers(0, max(n, 1), (nt, 3)).astype(np.int32)
    assert formats.write_ply_binary(v, t) == fo.orc_ply_binary(v, t)
    assert formats.write_ply_binary(v, None) == fo.orc_ply_binary(v, None)


@pytest.mark.gpu
def test_ply_and_transfer_frame_of_a_real_mesh():
    from livescan3d_b200 import api
    fr = small_frame(S=3, w=160, h=120)
    v, t = api.generate_mesh_from_depth_maps(fr, synth.DEFAULT_BOUNDS, triangles=True)
    assert len(t) > 1000
    assert formats.write_ply_binary(v, t) == fo.orc_ply_binary(v, t)
    lib = native.load()
    for limit in (65000 - 3, 5000, 997, 64):
        assert lib.ls3d_set_transfer_chunk_limit(limit) == 0
        try:
            assert formats.write_transfer_frame(v, t) == fo.orc_transfer_frame(v, t, limit), limit
        finally:
            lib.ls3d_set_transfer_chunk_limit(65000 - 3)
    assert formats.write_transfer_frame(v, None) == fo.orc_transfer_frame(v, None)


@pytest.mark.gpu
@pytest.mark.parametrize("limit", [3, 4, 10, 50])
def test_transfer_chunks_adversarial_index_lists(limit):
    """Random index lists (repeats inside a triangle, vertices reused across far-apart triangles, unused vertices) and tiny limits:
    chunk ends in the middle of vertex runs, chunks that close exactly at the last triangle, high-valence vertices."""
    lib = native.load()
    rng = np.random.default_rng(limit)
    assert lib.ls3d_set_transfer_chunk_limit(limit) == 0
    try:
        for n, nt in [(1, 1), (5, 40), (60, 300), (300, 200), (2000, 5000)]:
            v = _random_cloud(n, n)
            t = rng.integers(0, n, (nt, 3)).astype(np.int32)
            t[::7] = t[::7, :1]                                   # degenerate triangles (one vertex three times)
            assert formats.write_transfer_frame(v, t) == fo.orc_transfer_frame(v, t, limit), (n, nt)
        v = _random_cloud(9, 1)
        t = np.arange(9, dtype=np.int32).reshape(3, 3)            # limit 3: every triangle closes its own chunk, nothing trails
        assert formats.write_transfer_frame(v, t) == fo.orc_transfer_frame(v, t, limit)
        with pytest.raises(native.Ls3dError, match="outside"):
            formats.write_transfer_frame(v, np.array([[0, 1, 9]], np.int32))
    finally:
        lib.ls3d_set_transfer_chunk_limit(65000 - 3)


@pytest.mark.gpu
def test_transfer_frame_vertices_only_many_chunks():
    v = _random_cloud(65000 * 2 + 11, 3)
    assert formats.write_transfer_frame(v, None) == fo.orc_transfer_frame(v, None)


@pytest.mark.gpu
def test_device_resident_mesh_to_ply_and_transfer_body():
    """The frame pipeline's device buffers go straight into the packers (no host round trip of the 16-byte records)."""
    import torch
    from livescan3d_b200.device import FramePipeline, view
    fr = small_frame(S=2, w=160, h=120)
    dev = torch.device("cuda", 0)
    fp = FramePipeline(fr["widths"], fr["heights"])
    fp.set_params(fr["intr"], fr["wt"], synth.DEFAULT_BOUNDS, 0, 0.0)
    fp.enable_triangles(True)
    fp.run(torch.from_numpy(fr["depth_maps"]).to(dev), torch.from_numpy(fr["depth_colors"]).to(dev))
    torch.cuda.synchronize()
    c = fp.counts.cpu().numpy()
    n, nt = int(c[0]), int(c[4])
    lib = native.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.empty(15 * n + 13 * nt + 16, dtype=torch.uint8, device=dev)
    assert lib.ls3d_pack_ply_body_device(C.c_void_p(fp.vertices().data_ptr()), n, C.c_void_p(fp.triangles().data_ptr()), nt, C.c_void_p(out.data_ptr()), st) == 1
    v = fp.vertices()[:n].cpu().numpy().reshape(-1).view(VERTEX)
    t = fp.triangles()[:nt].cpu().numpy()
    want = fo.orc_ply_binary(v, t)
    assert out[: 15 * n + 13 * nt].cpu().numpy().tobytes() == want[len(want) - (15 * n + 13 * nt):]
    cv, ct = np.zeros(64, np.int32), np.zeros(64, np.int32)
    nv_out, body = C.c_int(0), C.c_void_p(0)
    lib.ls3d_set_transfer_chunk_limit(3000)
    try:
        chunks = lib.ls3d_transfer_chunks_device(C.c_void_p(fp.vertices().data_ptr()), n, C.c_void_p(fp.triangles().data_ptr()), nt,
                                                 C.c_void_p(cv.ctypes.data), C.c_void_p(ct.ctypes.data), 64, C.byref(nv_out), C.byref(body), st)
        assert chunks > 1, native.last_error()
        torch.cuda.synchronize()
        wantf = fo.orc_transfer_frame(v, t, 3000)
        head = 12 + 8 * chunks
        assert np.frombuffer(wantf, "<i4", 3).tolist() == [nv_out.value, nt, chunks]
        assert wantf[12:head] == cv[:chunks].tobytes() + ct[:chunks].tobytes()
        got = view(body.value, (len(wantf) - head,), "|u1", dev).cpu().numpy().tobytes()
        assert got == wantf[head:]
    finally:
        lib.ls3d_set_transfer_chunk_limit(65000 - 3)
