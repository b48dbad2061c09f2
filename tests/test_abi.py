"""The drop-in boundary on CPU: libls3d_b200.so loads without a GPU, exports every function include/ls3d.h declares
(and nothing in the header is missing from the ctypes table), keeps the reference's struct layouts, and — with no CUDA
device — fails loudly instead of computing anything on the CPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from common import ROOT

from livescan3d_b200 import native

HEADER = os.path.join(ROOT, "include", "ls3d.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(names))


def _ensure_built():
    if not os.path.exists(native.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "livescan3d_b200", "csrc"), "-j4"], check=True)


def test_header_declares_what_the_binding_types():
    fns = _header_functions()
    assert "ICP" in fns and "generateVerticesFromDepthMap" in fns and "generateMeshFromDepthMaps" in fns and "createMesh" in fns and "deleteMesh" in fns
    assert fns == native.exported_symbols(), (set(fns) ^ set(native.exported_symbols()))


def test_library_exports_every_header_symbol():
    _ensure_built()
    lib = native.load()                     # getattr on every symbol; raises if header and library disagree
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in _header_functions():
        assert name in exported, f"{name} is declared in include/ls3d.h but not exported as a C symbol"
        assert getattr(lib, name) is not None


def test_struct_layouts_match_the_reference():
    # depthprocessing.h:29-33,42-48; icp.h:15-18; utils.h:105-111 (x64)
    assert C.sizeof(native.Mesh) == 32
    assert native.Mesh.vertices.offset == 8 and native.Mesh.nTriangles.offset == 16 and native.Mesh.triangles.offset == 24
    from livescan3d_b200.api import VERTEX_DTYPE
    assert VERTEX_DTYPE.itemsize == 16 and VERTEX_DTYPE.fields["X"][1] == 4
    assert C.sizeof(native.IcpTrace) == 4 + 4 + 4 + 12 + 36


def test_mesh_lifetime_without_a_device():
    _ensure_built()
    lib = native.load()
    m = lib.createMesh()
    assert m and m.contents.nVertices == 0 and m.contents.nTriangles == 0
    lib.deleteMesh(m)                       # releases nothing, must not crash; the struct itself stays (depthprocessing.cpp:1828-1835)
    assert lib.ls3d_launch_count() >= 0


def test_no_cpu_fallback_without_a_device():
    """Without CUDA every compute entry must fail and say why — never produce numbers."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    _ensure_built()
    lib = native.load()
    v = np.zeros((8, 3), np.float32)
    c = np.zeros((8, 4), np.uint8)
    m = np.zeros(8, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.ls3d_filter(p(v), p(c), 8, 2, 0.5, p(m)) == -1
    assert "no CUDA device" in native.last_error() or "CUDA" in native.last_error()
    R = np.eye(3, dtype=np.float32).reshape(9)
    t = np.zeros(3, np.float32)
    v2 = v.copy() + 1
    assert lib.ICP(p(v), p(v2), 8, 8, p(R), p(t), 3) == 1.0                       # the reference's constant return value
    assert native.last_error() and np.array_equal(R, np.eye(3, dtype=np.float32).reshape(9)) and np.all(v2 == 1)   # untouched
    assert not lib.ls3d_frame_create(1, p(np.array([4], np.int32)), p(np.array([4], np.int32)))
    from livescan3d_b200 import api
    with pytest.raises(native.Ls3dError):
        api.filter(v, c, 2, 0.5)


def test_header_compiles_and_links_as_c(tmp_path):
    """include/ls3d.h is plain C: tests/c/abi_smoke.c is built as C99 with -pedantic -Werror, linked against the in-tree
    library like any C caller would, and run (Mesh lifetime calls + ls3d_last_error need no device)."""
    _ensure_built()
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(native.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lls3d_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout
    assert out.startswith("ok")
