"""Seeded ICP fuzz against the oracle (the north star's pose tolerance on random small pairs), as a test.

Every seed gives two overlapping sensor clouds of a random small rig, a random stride, a random rigid offset (<= 3 deg, <= 15 mm
per axis) and 1-6 iterations.  Two things are checked per seed:

  (1) ITERATION-LEVEL parity from identical states.  For every iteration k the oracle's own source cloud after k iterations is
      handed to the GPU for ONE iteration; n_matched and n_accepted must be identical to the oracle's iteration k and sigma, T
      and Rk must agree to fp32 accumulation noise.  This is the arithmetic check: it has no exceptions.
  (2) WHOLE-CALL pose: |dR| <= 1e-5 and |dt| <= 1e-4 m after all iterations.  ICP on a small coarse cloud amplifies the ~1e-7 pose
      difference the two implementations have after an iteration (fp64 reductions + Jacobi SVD here, OpenCV-style fp32 running
      sums there): one source point that flips between two nearly equidistant targets moves the next pose by residual / n.  A seed
      outside the tolerance is therefore accepted only if (1) held for every one of its iterations AND the traces show where the
      discrete matching split (n_matched or n_accepted differ at or before the first iteration whose T / Rk differ by more than
      the fp32 noise) — i.e. a correspondence flip, not an arithmetic difference.  The number of such seeds is bounded.

`python tests/test_gpu_icp_fuzz.py [n_seeds]` prints the same statistics over more seeds (used for DESIGN.md)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import rot_err, synth, orc  # noqa: E402

pytestmark = pytest.mark.gpu

R_TOL, T_TOL = 1e-5, 1e-4            # north star: rotation entries, translation in metres
STEP_TOL = 2e-6                      # one iteration from identical inputs: T (m) and Rk entries
SEEDS = list(range(1, 49))
_misses = []


@pytest.fixture(scope="module")
def api():
    from livescan3d_b200 import api as _api
    return _api


def make_pair(seed):
    rng = np.random.default_rng(100000 + seed)
    for _ in range(20):
        w, h = int(rng.integers(40, 200)), int(rng.integers(30, 150))
        fr = synth.make_frame(2, w, h, seed_base=int(rng.integers(1, 1 << 20)), ring=int(rng.integers(6, 9)))
        clouds = []
        for i in range(2):
            v, _ = orc.orc_generate_mesh(fr, synth.SERVER_BOUNDS, i)
            clouds.append(np.ascontiguousarray(np.stack([v["X"], v["Y"], v["Z"]], axis=1), dtype=np.float32))
        st = int(rng.integers(1, 6))
        A, B = np.ascontiguousarray(clouds[0][::st]), np.ascontiguousarray(clouds[1][::st])
        deg = float(rng.uniform(0.0, 3.0))
        tr = tuple(float(x) for x in rng.uniform(-15, 15, 3))
        iters = int(rng.integers(1, 7))
        if len(A) >= 64 and len(B) >= 64:
            return A, synth.perturb(B, deg=deg, trans_mm=tr), iters
    raise RuntimeError("no usable pair")


def step_parity(api, A, B, iters, wtr):
    """(1): one GPU iteration from each of the oracle's intermediate states.  Returns a list of problems (empty = parity)."""
    bad = []
    for k in range(iters):
        state = B if k == 0 else orc.orc_icp(A, B, max_iter=k)[0]
        _, _, _, g = api.icp_trace(A, state, max_iter=1)
        g, w = g[0], wtr[k]
        if g["n_matched"] != w["n_matched"]:
            bad.append((k, "n_matched", g["n_matched"], w["n_matched"]))
        if g["n_accepted"] != w["n_accepted"]:
            bad.append((k, "n_accepted", g["n_accepted"], w["n_accepted"]))
        if not abs(g["sigma"] - w["sigma"]) <= 2e-4 * w["sigma"]:
            bad.append((k, "sigma", g["sigma"], w["sigma"]))
        dT = float(np.max(np.abs(g["T"].astype(np.float64) - w["T"])))
        dRk = float(np.max(np.abs(g["Rk"].astype(np.float64) - w["Rk"])))
        if not (dT <= STEP_TOL and dRk <= STEP_TOL):
            bad.append((k, "T/Rk", dT, dRk))
    return bad


def whole_call(api, A, B, iters):
    wv, wR, wt, wtr = orc.orc_icp(A, B, max_iter=iters)
    gv, gR, gt, gtr = api.icp_trace(A, B, max_iter=iters)
    dR, dt = rot_err(gR, wR), float(np.max(np.abs(gt.astype(np.float64) - wt)))
    split = None                     # first iteration whose update differs by more than fp32 noise
    counts = None                    # first iteration whose match / accept counts differ
    for k, (g, w) in enumerate(zip(gtr, wtr)):
        if counts is None and (g["n_matched"] != w["n_matched"] or g["n_accepted"] != w["n_accepted"]):
            counts = k
        d = max(float(np.max(np.abs(g["T"].astype(np.float64) - w["T"]))), float(np.max(np.abs(g["Rk"].astype(np.float64) - w["Rk"]))))
        if split is None and d > 10 * STEP_TOL:
            split = k
    return dR, dt, split, counts, wtr, gtr


@pytest.mark.parametrize("seed", SEEDS)
def test_icp_fuzz_seed(api, seed):
    A, B, iters = make_pair(seed)
    dR, dt, split, counts, wtr, gtr = whole_call(api, A, B, iters)
    bad = step_parity(api, A, B, iters, wtr)
    assert not bad, f"seed {seed} (n1={len(A)} n2={len(B)} iters={iters}): one-iteration parity from the oracle's states broke: {bad}"
    if dR <= R_TOL and dt <= T_TOL:
        return
    # outside the whole-call tolerance: must be a split of the discrete matching, visible in the counts no later than the update split
    _misses.append((seed, len(A), len(B), iters, dR, dt, split, counts))
    assert counts is not None and (split is None or counts <= split), \
        f"seed {seed}: pose off by dR={dR:.2e} dt={dt:.2e} m without a visible matching split (update split at {split}, counts at {counts})"
    assert dR <= 5e-4 and dt <= 5e-3, f"seed {seed}: gross disagreement dR={dR:.2e} dt={dt:.2e}"


def test_icp_fuzz_miss_rate():
    """Runs after the seeds: at most 2 of the 48 may sit outside the whole-call tolerance (each already proven to be a matching split)."""
    print("ICP fuzz: seeds outside 1e-5 / 1e-4 m:", _misses)
    assert len(_misses) <= 2, _misses


if __name__ == "__main__":
    from livescan3d_b200 import api as _api
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    out, step_bad, worst = [], 0, [0.0, 0.0]
    for seed in range(1000, 1000 + n):
        A, B, iters = make_pair(seed)
        dR, dt, split, counts, wtr, gtr = whole_call(_api, A, B, iters)
        bad = step_parity(_api, A, B, iters, wtr)
        step_bad += 1 if bad else 0
        if bad:
            print("STEP PARITY", seed, len(A), len(B), iters, bad)
        if not (dR <= R_TOL and dt <= T_TOL):
            out.append((seed, len(A), len(B), iters, dR, dt, split, counts))
            worst = [max(worst[0], dR), max(worst[1], dt)]
    print({"pairs": n, "one_iteration_parity_failures": step_bad, "outside_whole_call_tolerance": len(out), "worst_dR": worst[0], "worst_dt_m": worst[1]})
    for o in out:
        print("  miss", o)
