"""Pins the CPU oracle (oracle/ls3d_oracle.cpp, a restatement) against the reference's OWN sources compiled in
place (oracle/_ref/, built by oracle/Makefile from /root/reference) on seeded inputs, and against the golden
vectors under tests/golden/ that were generated from that build (tests/golden/make_golden.py).

The reference ships no golden vectors for this path, so these two are what "pinned" means here.
CPU only."""
import os

import numpy as np
import pytest

from common import GOLDEN, cloud_of, icp_pair, nn_parity, rot_err, small_frame, synth, orc

needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
BOUNDS = [synth.DEFAULT_BOUNDS, synth.SERVER_BOUNDS, synth.CLIENT_BOUNDS]


@needs_ref
@pytest.mark.parametrize("bi", [0, 1, 2])
@pytest.mark.parametrize("seed", [1000, 2000])
def test_vertices_bit_exact_vs_reference(bi, seed):
    fr = small_frame(S=3, seed_base=seed)
    ref_v, ref_counts = orc.ref_generate_mesh(fr, BOUNDS[bi])
    orc_v, orc_counts = orc.orc_generate_mesh(fr, BOUNDS[bi])
    assert np.array_equal(ref_counts, orc_counts)
    assert ref_v.tobytes() == orc_v.tobytes()
    # the reference's single-sensor export agrees with the corresponding slice
    one = orc.ref_generate_vertices_from_depth_map(fr, BOUNDS[bi], 1)
    s = int(ref_counts[0])
    assert one.tobytes() == ref_v[s:s + int(ref_counts[1])].tobytes()


@needs_ref
@pytest.mark.parametrize("S,w,h", [(3, 128, 96), (2, 37, 29), (1, 5, 5), (1, 4, 3), (2, 512, 424)])
def test_triangles_bit_exact_vs_reference(S, w, h):
    """generateTriangles + formMesh (depthprocessing.cpp:1659-1691,1611-1626; meshGenerator.cpp:14-181) through the reference's
    own objects vs the restatement."""
    fr = synth.make_frame(S, w, h)
    for b in BOUNDS[:2]:
        rv, rc, rt = orc.ref_generate_mesh(fr, b, with_triangles=True)
        ov, ot, oc, otc = orc.orc_generate_mesh_triangles(fr, b)
        assert rv.tobytes() == ov.tobytes() and np.array_equal(rc, oc)
        assert rt.shape == ot.shape and np.array_equal(rt, ot) and int(otc.sum()) == len(ot)


@needs_ref
@pytest.mark.parametrize("S,w,h,seed", [(3, 128, 96, 1000), (2, 37, 29, 3), (1, 5, 5, 1), (1, 3, 3, 1), (1, 2, 7, 2), (2, 512, 424, 1000)])
def test_radial_correction_bit_exact_vs_reference(S, w, h, seed):
    """depthMapAndColorSetRadialCorrection (depthprocessing.cpp:1794-1815, :191-261): the reference's own export vs the restatement."""
    fr = synth.make_frame(S, w, h, seed_base=seed)
    rd, rc = orc.ref_radial_correction(fr)
    od, oc = orc.orc_radial_correction(fr)
    assert np.array_equal(rd, od) and np.array_equal(rc, oc)
    # strong coefficients and many holes: the in-place fill cascades
    rng = np.random.default_rng(seed)
    fr2 = dict(fr)
    d = fr["depth_maps"].view(np.uint16).copy()
    d[rng.random(d.shape) < 0.3] = 0
    fr2["depth_maps"] = d.view(np.uint8)
    intr = fr["intr"].copy(); intr[4::7] = 0.8; intr[5::7] = 0.3; intr[6::7] = -0.2
    fr2["intr"] = intr
    rd, rc = orc.ref_radial_correction(fr2)
    od, oc = orc.orc_radial_correction(fr2)
    assert np.array_equal(rd, od) and np.array_equal(rc, oc)


@needs_ref
def test_fixture_poses_and_pixel_maps_vs_reference():
    fr = synth.make_frame(2, 160, 120, poses=synth.FIXTURE_POSES)
    for b in BOUNDS:
        assert orc.ref_generate_mesh(fr, b)[0].tobytes() == orc.orc_generate_mesh(fr, b)[0].tobytes()
    w, h = 160, 120
    d = fr["depth_maps"].view(np.uint16)[: w * h]
    c = fr["depth_colors"][: 3 * w * h]
    n1, a1, b1 = orc.ref_vertex_maps(d, c, w, h, fr["intr"][:7], fr["wt"][:12], synth.SERVER_BOUNDS)
    n2, a2, b2 = orc.orc_vertex_maps(d, c, w, h, fr["intr"][:7], fr["wt"][:12], synth.SERVER_BOUNDS)
    assert n1 == n2 and np.array_equal(a1, a2) and np.array_equal(b1, b2)


@needs_ref
@pytest.mark.parametrize("k,thr", [(1, 10.0), (2, 25.0), (3, 5.5), (0, 1.0), (1, 0.0), (4, 300.0)])
def test_flying_pixel_filter_bit_exact_vs_reference(k, thr):
    """N2: orc_filter_flying_pixels against the reference's own filterFlyingPixels (kinectCapture.cpp:132-174 compiled in
    place), incl. the transposed shifts (:146), the overwritten maxNonFittingNeighbours (:150), non-square and tiny images."""
    rng = np.random.default_rng(7)
    for (w, h, seed) in ((512, 424, 1000), (160, 120, 5), (37, 29, 4), (7, 5, 1), (3, 3, 2), (2, 9, 3), (9, 2, 3)):
        fr = synth.make_frame(1, w, h, seed_base=seed)
        cases = [fr["depth_maps"].view(np.uint16).copy(), rng.integers(0, 65536, w * h).astype(np.uint16),
                 (rng.integers(0, 3, w * h) * 4000).astype(np.uint16)]
        for d in cases:
            for mnf in (0, 123):
                want = orc.ref_filter_flying_pixels(d, w, h, k, thr, mnf)
                got = orc.orc_filter_flying_pixels(d, w, h, k, thr, mnf)
                assert np.array_equal(got, want), (w, h, k, thr, mnf)
        if w == 512 and k == 1 and thr == 10.0:
            assert (orc.ref_filter_flying_pixels(cases[0], w, h, k, thr) == 0).sum() > (cases[0] == 0).sum()


@needs_ref
@pytest.mark.parametrize("k,md", [(10, 0.01), (10, 0.1), (1, 0.01), (50, 0.05), (3, 0.02)])
def test_filter_bit_exact_vs_reference(k, md):
    fr = small_frame(S=1, w=160, h=120)
    xyz, rgba = cloud_of(fr, synth.DEFAULT_BOUNDS, 0)
    rv, rc, rm = orc.ref_filter(xyz, rgba, k, md)
    ov, oc, om = orc.orc_filter(xyz, rgba, k, md)
    assert np.array_equal(rm, om)
    assert rv.tobytes() == ov.tobytes() and rc.tobytes() == oc.tobytes()
    assert np.array_equal(orc.ref_knn_kdist(xyz, k).view(np.uint32), orc.orc_knn_kdist(xyz, k).view(np.uint32))


@needs_ref
@pytest.mark.parametrize("seed,w,h,bi,k,md", [(21, 37, 29, 0, 3, 0.12), (22, 5, 5, 1, 1, 0.5), (23, 97, 61, 2, 7, 0.06), (24, 128, 96, 0, 29, 0.09),
                                             (25, 64, 48, 1, 10, 0.1), (27, 31, 201, 1, 5, 0.08)])
def test_filter_seeded_sweep_vs_reference(seed, w, h, bi, k, md):
    """More rigs (odd and extreme aspect ratios, all three bounds presets, k up to 29) with partial survival: masks, vertices, colours."""
    fr = small_frame(S=2, w=w, h=h, seed_base=seed)
    for idx in (0, 1):
        xyz, rgba = cloud_of(fr, BOUNDS[bi], idx)
        assert len(xyz) > 0
        rv, rc, rm = orc.ref_filter(xyz, rgba, k, md)
        ov, oc, om = orc.orc_filter(xyz, rgba, k, md)
        assert np.array_equal(rm, om)
        assert rv.tobytes() == ov.tobytes() and rc.tobytes() == oc.tobytes()


@needs_ref
def test_filter_edge_cases_vs_reference():
    fr = small_frame(S=1, w=64, h=48)
    xyz, rgba = cloud_of(fr, synth.SERVER_BOUNDS, 0)
    for k, md in [(0, 0.01), (10, 0.0), (-1, -1.0)]:            # early return: untouched, map holds only the sentinel
        rv, rc, rm = orc.ref_filter(xyz, rgba, k, md)
        ov, oc, om = orc.orc_filter(xyz, rgba, k, md)
        assert len(rv) == len(xyz) and np.array_equal(rm, om) and np.all(rm == -2)
    few = xyz[:5]                                                # k > n: every point is removed (nanoflann.h:93)
    rv, rc, rm = orc.ref_filter(few, rgba[:5], 10, 10.0)
    ov, oc, om = orc.orc_filter(few, rgba[:5], 10, 10.0)
    assert len(rv) == 0 and len(ov) == 0 and np.array_equal(rm, om)


@needs_ref
def test_find_closest_vs_reference_and_brute_force():
    fr = small_frame(S=2, w=160, h=120)
    A, B = icp_pair(fr, synth.DEFAULT_BOUNDS)
    ri, rd = orc.ref_find_closest(A, B)
    oi, od = orc.orc_find_closest(A, B)
    assert np.array_equal(ri, oi) and np.array_equal(rd.view(np.uint32), od.view(np.uint32))
    bi, bd = orc.orc_find_closest(A, B, brute=True)
    mism, bad = nn_parity(oi, od, bi, bd)
    assert bad == 0


@needs_ref
def test_icp_vs_reference():
    fr = small_frame(S=2, w=160, h=120)
    A, B = icp_pair(fr, synth.DEFAULT_BOUNDS)
    rv, rR, rt = orc.ref_icp(A, B, max_iter=5)
    ov, oR, ot, _ = orc.orc_icp(A, B, max_iter=5)
    # same source compiled around the same mini-cv arithmetic: identical results
    assert np.array_equal(rR.view(np.uint32), oR.view(np.uint32))
    assert np.array_equal(rt.view(np.uint32), ot.view(np.uint32))
    assert rv.tobytes() == ov.tobytes()
    # accumulate-into semantics: non-identity R, non-zero t on entry (MainWindowForm.cs:330-344 passes I, 0)
    R0 = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float32)
    t0 = np.array([0.1, -0.2, 0.3], dtype=np.float32)
    rv, rR, rt = orc.ref_icp(A, B, R0, t0, max_iter=2)
    ov, oR, ot, _ = orc.orc_icp(A, B, R0, t0, max_iter=2)
    assert np.array_equal(rR.view(np.uint32), oR.view(np.uint32)) and np.array_equal(rt.view(np.uint32), ot.view(np.uint32))


@needs_ref
@pytest.mark.parametrize("seed,w,h,stride,iters", [(11, 48, 36, 1, 3), (12, 64, 48, 3, 4), (13, 40, 30, 7, 2), (14, 96, 72, 5, 6), (15, 33, 27, 2, 5),
                                                  (16, 64, 48, 29, 3), (17, 80, 60, 113, 2)])
def test_nn_and_icp_seeded_sweep_vs_reference(seed, w, h, stride, iters):
    """More rigs, sizes (down to 15 points) and iteration counts: NN indices / d2 bits and the whole ICP call, oracle vs the compiled reference."""
    fr = small_frame(S=2, w=w, h=h, seed_base=seed)
    A, B = icp_pair(fr, synth.DEFAULT_BOUNDS, stride=stride)
    assert len(A) > 0 and len(B) > 0
    ri, rd = orc.ref_find_closest(A, B)
    oi, od = orc.orc_find_closest(A, B)
    assert np.array_equal(ri, oi) and np.array_equal(rd.view(np.uint32), od.view(np.uint32))
    rv, rR, rt = orc.ref_icp(A, B, max_iter=iters)
    ov, oR, ot, _ = orc.orc_icp(A, B, max_iter=iters)
    assert np.array_equal(rR.view(np.uint32), oR.view(np.uint32)) and np.array_equal(rt.view(np.uint32), ot.view(np.uint32))
    assert rv.tobytes() == ov.tobytes()


def test_icp_recovers_known_offset():
    """Sanity of the oracle itself: ICP reduces the known 1.5 deg / (8,-5,6) mm perturbation (point-to-point ICP
    slides along the scene's planes, so 10 iterations only take part of it back)."""
    fr = small_frame(S=2, w=160, h=120)
    A, _ = cloud_of(fr, synth.DEFAULT_BOUNDS, 0)
    B = synth.perturb(A)
    ov, oR, ot, tr = orc.orc_icp(A, B, max_iter=10)
    assert np.sqrt(((ov - A) ** 2).sum(1)).mean() < 0.8 * np.sqrt(((B - A) ** 2).sum(1)).mean()
    assert tr[0]["n_matched"] > 0 and tr[0]["n_accepted"] <= tr[0]["n_matched"]


def test_oracle_vs_golden_vectors():
    """tests/golden/*.npz were written by tests/golden/make_golden.py from the compiled reference (oracle/_ref)."""
    path = os.path.join(GOLDEN, "hotpath_small.npz")
    assert os.path.exists(path), "golden fixture missing: run python tests/golden/make_golden.py in the build container"
    g = np.load(path)
    fr = synth.make_frame(int(g["S"]), int(g["w"]), int(g["h"]), seed_base=int(g["seed_base"]), ring=int(g["ring"]))
    assert fr["depth_maps"].tobytes() == g["depth_maps"].tobytes(), "synthetic generator drifted from the fixture"
    v, counts = orc.orc_generate_mesh(fr, g["bounds"])
    assert np.array_equal(counts, g["vertex_counts"]) and v.tobytes() == g["vertices"].tobytes()
    tv, tri, _, _ = orc.orc_generate_mesh_triangles(fr, g["bounds"])
    assert tv.tobytes() == g["vertices"].tobytes() and np.array_equal(tri, g["triangles"])
    xyz, rgba = cloud_of(fr, g["bounds"], 0)
    for i, (k, md) in enumerate(zip(g["filter_k"], g["filter_maxdist"])):
        _, _, m = orc.orc_filter(xyz, rgba, int(k), float(md))
        assert np.array_equal(m, g[f"filter_map_{i}"])
    A, B = icp_pair(fr, g["bounds"])
    oi, od = orc.orc_find_closest(A, B)
    assert np.array_equal(oi, g["nn_index"]) and np.array_equal(od.view(np.uint32), g["nn_d2"].view(np.uint32))
    ov, oR, ot, _ = orc.orc_icp(A, B, max_iter=int(g["icp_iters"]))
    assert rot_err(oR, g["icp_R"]) == 0.0 and np.array_equal(ot, g["icp_t"])


def test_oracle_vs_golden_flying_pixels():
    """tests/golden/flying_small.npz: outputs of the reference's own filterFlyingPixels (tests/golden/make_golden_flying.py)."""
    g = np.load(os.path.join(GOLDEN, "flying_small.npz"))
    w, h = int(g["w"]), int(g["h"])
    d = synth.make_frame(1, w, h, seed_base=int(g["seed_base"]))["depth_maps"].view(np.uint16)
    assert np.array_equal(d, g["depth"]), "synthetic generator drifted from the fixture"
    for i, (k, thr) in enumerate(zip(g["k"], g["thr"])):
        assert np.array_equal(orc.orc_filter_flying_pixels(d, w, h, int(k), float(thr), 55), g[f"out_{i}"])
