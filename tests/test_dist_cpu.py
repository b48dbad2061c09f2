"""The N > 1 host logic on CPU: world_size-2 `gloo` process groups drive the same planning / key-packing helpers
livescan3d_b200/dist.py uses on the GPUs, with the CPU oracle standing in for the per-rank kernels, and the result is
compared with the unsharded oracle.  (The device side of the same protocol is covered by the -m gpu tests and
bench.py's "sharded" block at N > 1.)"""
import os
import socket
import sys

import numpy as np
import pytest

from common import ROOT, cloud_of, icp_pair, small_frame, synth, orc

from livescan3d_b200 import dist as ldist


def test_sensor_and_slice_ranges():
    for S in range(0, 12):
        for world in (1, 2, 3, 4, 8):
            r = ldist.sensor_ranges(S, world)
            assert len(r) == world and sum(n for _, n in r) == S
            assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(world - 1))            # contiguous, in order
            assert max(n for _, n in r) - min(n for _, n in r) <= 1                              # balanced
    for n in (0, 1, 31, 32, 33, 1000, 212797):
        for world in (1, 2, 4, 8):
            s = ldist.slice_ranges(n, world)
            assert s[0][0] == 0 and s[-1][1] == n
            assert all(s[i][1] == s[i + 1][0] for i in range(world - 1))
            assert all(b % 32 == 0 for b, _ in s if b < n)
    assert list(ldist.exclusive_offsets([5, 0, 7, 2])) == [0, 5, 5, 12]
    assert list(ldist.exclusive_offsets([])) == []
    with pytest.raises(ValueError):
        ldist.sensor_ranges(4, 0)


def test_slot_keys_reproduce_the_one_to_one_rule():
    """pack/unpack == the reference's dedupe loop (icp.cpp:95-126) incl. 'later index wins ties', on adversarial input."""
    rng = np.random.default_rng(5)
    n1, n2 = 50, 400
    idx = rng.integers(0, n1, n2).astype(np.uint64)
    d2 = rng.choice(np.array([0.0, 1e-6, 0.25, 0.25, 3.5, 1e-30], dtype=np.float32), n2)      # many exact ties
    want = orc.orc_dedupe(idx, d2, n1)
    win, wd2 = ldist.unpack_slot_keys(ldist.pack_slot_keys(n1, idx.astype(np.int64), d2))
    assert np.array_equal(win, want)
    has = want >= 0
    assert np.array_equal(wd2[has].view(np.uint32), d2[want[has]].view(np.uint32))
    # merging per-slice arrays with MIN == packing everything at once
    a = ldist.pack_slot_keys(n1, idx[:160].astype(np.int64), d2[:160], 0)
    b = ldist.pack_slot_keys(n1, idx[160:].astype(np.int64), d2[160:], 160)
    assert np.array_equal(np.minimum(a, b), ldist.pack_slot_keys(n1, idx.astype(np.int64), d2))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        # ---- frame: sensors sharded over ranks, merged at all-gathered offsets
        S, k, md = 5, 6, 0.03
        fr = small_frame(S=S, w=96, h=72)
        first, n_own = ldist.sensor_ranges(S, world)[rank]
        parts = []
        for s in range(first, first + n_own):
            v, _ = orc.orc_generate_mesh(fr, synth.DEFAULT_BOUNDS, s)
            xyz = np.stack([v["X"], v["Y"], v["Z"]], axis=1)
            col = np.stack([v["R"], v["G"], v["B"], v["A"]], axis=1)
            _, _, m = orc.orc_filter(xyz, col, k, md)
            parts.append(v[m >= 0])
        own = np.concatenate(parts) if parts else np.zeros(0, dtype=orc.VERTEX_DTYPE)
        kept_all = torch.zeros(world, dtype=torch.int32)
        dist.all_gather_into_tensor(kept_all, torch.tensor([len(own)], dtype=torch.int32))
        offs = ldist.exclusive_offsets(kept_all.numpy())
        total = int(kept_all.sum())
        merged = torch.zeros(total * 16, dtype=torch.uint8)                      # every rank's copy of the merged cloud
        merged[offs[rank] * 16:(offs[rank] + len(own)) * 16] = torch.from_numpy(own.view(np.uint8).copy())
        dist.all_reduce(merged, op=dist.ReduceOp.MAX)                           # stands in for the peer stores (regions are disjoint)
        res["merged"] = merged.numpy().tobytes()
        res["counts"] = kept_all.numpy().tolist()

        # ---- ICP dedupe: source slices; every key goes to the OWNER rank's slots (on the GPUs: 64-bit atomicMin over NVLink peer
        # mappings; here: a MIN all-reduce after which each rank keeps only the range it owns)
        A, B = icp_pair(small_frame(S=2, w=96, h=72), synth.SERVER_BOUNDS)
        n1 = len(A)
        b, e = ldist.slice_ranges(len(B), world)[rank]
        idx, d2 = orc.orc_find_closest(A, B[b:e]) if e > b else (np.zeros(0, np.uint64), np.zeros(0, np.float32))
        slots = torch.from_numpy(ldist.pack_slot_keys(n1, idx.astype(np.int64), d2, b))
        dist.all_reduce(slots, op=dist.ReduceOp.MIN)
        S_ = ldist.reduction_chunk_size(n1)
        c0, c1 = ldist.chunk_owner_ranges(n1, world)[rank]
        own = slots.numpy()[min(n1, c0 * S_):min(n1, c1 * S_)].copy()            # the slots this rank owns
        assert np.all(ldist.slot_owner(np.arange(min(n1, c0 * S_), min(n1, c1 * S_)), n1, world) == rank)
        res["own_slots"] = (min(n1, c0 * S_), own)
        # canonical chunked reduction of the matched d2 (pass 1 of k_icp_reduce): the owner computes its chunks' partials, every rank
        # receives all C of them (peer stores on the GPUs, a SUM all-reduce of disjoint rows here) and folds them in the same order
        C_ = ldist.reduction_chunks(n1)
        table = torch.zeros(C_, 2, dtype=torch.float64)
        for c in range(c0, c1):
            win, wd2 = ldist.unpack_slot_keys(slots.numpy()[min(n1, c * S_):min(n1, (c + 1) * S_)])
            table[c, 0] = float((win >= 0).sum())
            table[c, 1] = float(wd2[win >= 0].astype(np.float64).sum())
        dist.all_reduce(table)
        res["stats"] = ldist.fold_chunk_partials(table.numpy())
        res["table"] = table.numpy().copy()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_protocols_world2_gloo():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # unsharded truth
    S, k, md = 5, 6, 0.03
    fr = small_frame(S=S, w=96, h=72)
    parts = []
    for s in range(S):
        xyz, _ = cloud_of(fr, synth.DEFAULT_BOUNDS, s)
        v, _ = orc.orc_generate_mesh(fr, synth.DEFAULT_BOUNDS, s)
        col = np.stack([v["R"], v["G"], v["B"], v["A"]], axis=1)
        _, _, m = orc.orc_filter(xyz, col, k, md)
        parts.append(v[m >= 0])
    want = np.concatenate(parts)
    for r in range(world):
        assert got[r]["merged"] == want.tobytes()                                    # every rank holds the same, correctly ordered cloud
        assert sum(got[r]["counts"]) == len(want)

    A, B = icp_pair(small_frame(S=2, w=96, h=72), synth.SERVER_BOUNDS)
    n1 = len(A)
    idx, d2 = orc.orc_find_closest(A, B)
    winner = orc.orc_dedupe(idx, d2, n1)
    has = winner >= 0
    # the owned ranges tile the target exactly, and together they hold the reference's one-to-one matches
    whole = np.full(n1, ldist.SLOT_EMPTY, dtype=np.int64)
    covered = 0
    for r in range(world):
        first, own = got[r]["own_slots"]
        whole[first:first + len(own)] = own
        covered += len(own)
    assert covered == n1
    win, wd2 = ldist.unpack_slot_keys(whole)
    assert np.array_equal(win, winner)                                               # queries are identical, so no tie can differ
    assert np.array_equal(wd2[has].view(np.uint32), d2[winner[has]].view(np.uint32))
    # the single-rank table, computed the same way, folds to the same bits on every rank (world-independence of the reduction)
    S_, C_ = ldist.reduction_chunk_size(n1), ldist.reduction_chunks(n1)
    one = np.zeros((C_, 2))
    for c in range(C_):
        w1, wd1 = ldist.unpack_slot_keys(whole[min(n1, c * S_):min(n1, (c + 1) * S_)])
        one[c] = [float((w1 >= 0).sum()), float(wd1[w1 >= 0].astype(np.float64).sum())]
    want_stats = ldist.fold_chunk_partials(one)
    for r in range(world):
        assert np.array_equal(got[r]["table"], one)
        assert np.array_equal(got[r]["stats"], want_stats) and got[r]["stats"][0] == has.sum()
    # planning helpers: owners are monotone, contiguous and cover every rank count
    for n in (1, 255, 256, 257, 5000, 212654, 2_000_000):
        for w in (1, 2, 3, 4, 8):
            rr = ldist.chunk_owner_ranges(n, w)
            assert rr[0][0] == 0 and rr[-1][1] == ldist.reduction_chunks(n) and all(rr[i][1] == rr[i + 1][0] for i in range(w - 1))
            o = ldist.slot_owner(np.arange(0, n, max(1, n // 997)), n, w)
            assert np.all(np.diff(o) >= 0) and o.min() >= 0 and o.max() < w
            S2 = ldist.reduction_chunk_size(n)
            for r, (a0, a1) in enumerate(rr):
                if a1 > a0 and a0 * S2 < n:
                    assert ldist.slot_owner([a0 * S2, min(n, a1 * S2) - 1], n, w).tolist() == [r, r]
