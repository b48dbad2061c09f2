"""Shared helpers for the test-suite: seeded inputs and the parity predicates the north star states."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from livescan3d_b200 import synth  # noqa: E402
from oracle import oracle_lib as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def small_frame(S=2, w=128, h=96, seed_base=1000, ring=8):
    """A reduced-resolution rig (same scene, same sensor model scaled) the CPU oracle finishes in well under a second."""
    return synth.make_frame(S, w, h, seed_base=seed_base, ring=ring)


def cloud_of(frame, bounds, index):
    v, _ = orc.orc_generate_mesh(frame, bounds, index)
    xyz = np.stack([v["X"], v["Y"], v["Z"]], axis=1).astype(np.float32)
    rgba = np.stack([v["B"], v["G"], v["R"], np.zeros_like(v["R"])], axis=1).astype(np.uint8)    # RGB struct order: B,G,R,reserved
    return xyz, rgba


def xyz_of(verts):
    return np.stack([verts["X"], verts["Y"], verts["Z"]], axis=1).astype(np.float32)


def nn_parity(idx_gpu, d2_gpu, idx_ref, d2_ref):
    """North-star rule: indices identical except for equidistant ties.  Returns (n_index_mismatch, n_bad) where a
    mismatch is bad unless both report the same fp32 d2 (bit compare) or the GPU found a strictly closer point
    (nanoflann's fp32 branch bound can in principle skip a neighbour closer by a few ulp)."""
    idx_gpu = np.asarray(idx_gpu).astype(np.int64)
    idx_ref = np.asarray(idx_ref).astype(np.int64)
    mism = idx_gpu != idx_ref
    bad = mism & ~(d2_gpu.view(np.uint32) == d2_ref.view(np.uint32)) & ~(d2_gpu < d2_ref)
    return int(mism.sum()), int(bad.sum())


def rot_err(Ra, Rb):
    return float(np.max(np.abs(np.asarray(Ra, dtype=np.float64) - np.asarray(Rb, dtype=np.float64))))


def icp_pair(frame, bounds, a=0, b=1, stride=1):
    """Config-1 style pair: target = sensor a's cloud, source = sensor b's cloud with the known rigid offset."""
    A, _ = cloud_of(frame, bounds, a)
    B, _ = cloud_of(frame, bounds, b)
    B = synth.perturb(B)
    return np.ascontiguousarray(A[::stride]), np.ascontiguousarray(B[::stride])
