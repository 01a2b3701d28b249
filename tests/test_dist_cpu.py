"""CPU suite: the multi-GPU plumbing (3pre_b200/dist.py) on the gloo backend, world_size 2.

Covers the host logic of both partitionings of SURVEY.md 8e: block sharding of independent units
with a record gather, and the hypothesis-block split of one pair -- each rank evaluates its block
(here with the CPU oracle standing in for the kernels), the ranks agree on the winner with ONE
collective, and the result equals the single-process run of the reference rule."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

pd = importlib.import_module("3pre_b200.dist")


def test_split_range_is_a_partition():
    for total in (0, 1, 7, 4096, 1000001):
        for ws in (1, 2, 3, 8):
            edges = [pd.split_range(total, r, ws) for r in range(ws)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[r][1] == edges[r + 1][0] for r in range(ws - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_key_order_is_max_count_then_lowest_id():
    k = pd.pack_key
    assert k(10, 5) > k(9, 0) and k(10, 5) > k(10, 6) and k(10, 0) > k(10, 1)
    assert pd.unpack_key(k(123456, 999999)) == (123456, 999999)
    assert k(2**31 - 1, 0) < 2**63  # fits a signed 64-bit MAX reduce


def test_pick_reference_rule():
    # max count, then min ErrorSum, then lowest id; ranks without a recorded hypothesis (-1) never win
    assert pd.pick_reference([5, 7, 7, -1], [0, 10, 20, 30], [1.0, 3.0, 2.0, 0.0]) == 2
    assert pd.pick_reference([7, 7], [10, 4], [2.0, 2.0]) == 1
    assert pd.pick_reference([-1, -1], [0, 0], [0.0, 0.0]) == -1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from oracle import oracle as orc
        synth = importlib.import_module("3pre_b200.synth")
        out = {}
        # ---- independent units: each rank "solves" its block, records are gathered ---------------
        P = 11
        lo, hi = pd.split_range(P, rank, ws)
        local = torch.zeros(hi - lo, 16, dtype=torch.uint8)
        for i, p in enumerate(range(lo, hi)):
            local[i] = p  # the record of pair p is 16 bytes of value p
        allrec = pd.gather_records(local, P)
        out["gather_ok"] = bool((allrec[:, 0] == torch.arange(P, dtype=torch.uint8)).all()) and allrec.shape == (P, 16)
        # ---- one pair, hypotheses split: local winner by the oracle, one collective -----------------
        c = synth.make_correspondences(77, N=400, outlier_ratio=0.6)
        H = 600
        samples = orc.sample_sets(5, 0, H, 400, 5)
        h0, h1 = pd.split_range(H, rank, ws)
        loc = orc.ransac(c.Ya, c.Yb, samples[h0:h1], method=0, max_iteration=H + 1, adaptive=False)
        # "first" mode: max count, lowest id
        first_local = int(np.flatnonzero(loc.counts == loc.counts.max())[0])
        key = torch.tensor([pd.pack_key(int(loc.counts.max()), h0 + first_local)], dtype=torch.int64)
        pd.allreduce_max_key(key)
        out["first"] = pd.unpack_key(int(key.item()))
        # "reference" mode: (count, ErrorSum, id) of the local winner under the full rule
        w, cnt, gid, es = pd.agree_on_winner(loc.best_fit, h0 + loc.best_sample, loc.error_sum, "reference", "cpu")
        out["ref"] = (w, cnt, gid, es)
        _, cnt1, gid1, _ = pd.agree_on_winner(int(loc.counts.max()), h0 + first_local, 0.0, "first", "cpu")
        out["first2"] = (cnt1, gid1)
        if rank == 0:
            g = orc.ransac(c.Ya, c.Yb, samples, method=0, max_iteration=H + 1, adaptive=False)
            out["global"] = (int(g.counts.max()), int(np.flatnonzero(g.counts == g.counts.max())[0]), g.best_fit,
                             g.best_sample, g.error_sum)
        q.put((rank, out))
    finally:
        tdist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_two_gloo():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(ws))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    gmax, gfirst, gbest_fit, gbest_sample, ges = got[0]["global"]
    for r in range(ws):
        assert got[r]["gather_ok"]
        assert got[r]["first"] == (gmax, gfirst) and got[r]["first2"] == (gmax, gfirst)
        w, cnt, gid, es = got[r]["ref"]
        assert (cnt, gid, es) == (gbest_fit, gbest_sample, ges)
        assert w == (0 if gid < pd.split_range(600, 0, 2)[1] else 1)
