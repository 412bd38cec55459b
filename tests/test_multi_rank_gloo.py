"""The N > 1 path's host logic on CPU: two ranks over gloo.  Each rank takes its bit-balanced shard (bvg_plan_shards),
"decodes" it with the oracle (the stand-in for the CUDA path: there is no GPU here), publishes its boundary lists in the
fixed-size message of webgraph_b200/sharding.py through the one all-gather of a step, and rank r checks that what it
receives from rank r - 1 is exactly the lists its own reference chains reach back to; the per-rank (arcs, checksum) pairs
must combine to the whole graph's.  Mirrors ImmutableGraph.splitNodeIterators + the window re-read of
BVGraphNodeIterator's constructor (reference ImmutableGraph.java:379-409, BVGraph.java:1173-1183)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import graphs
from tests import oracle_binding as ob
from tests.conftest import ROOT
from webgraph_b200 import bvgraph, sharding, tools


def _rank_main(rank, world, base, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = ob.load().load(base)
        bounds = bvgraph.plan_shards(base, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        assert bounds[0] == 0 and bounds[-1] == g.n and all(a <= b for a, b in zip(bounds, bounds[1:]))
        off, succ = g.decode_range(0, g.n)  # the oracle is the checker AND the stand-in decoder of this CPU test
        reach = g.window * 3                # W x maxrefcount of the fixture
        bc = min(reach, hi - lo)
        need = sharding.any_rank(rank > 0 and bc > 0, "cpu")
        assert need
        first = hi - bc
        local_arcs = int(off[hi] - off[first])
        bcap = sharding.agree_capacity(local_arcs, "cpu")
        words = sharding.message_words(reach, bcap)
        send = torch.zeros(words, dtype=torch.int64)
        s_off, s_lists = sharding.message_views(send, reach, bcap)
        s_off[:bc + 1] = torch.from_numpy((off[first:hi + 1] - off[first]).astype(np.int64))
        s_lists[:local_arcs] = torch.from_numpy(succ[off[first]:off[hi]].astype(np.int32))
        recv = torch.zeros(world * words, dtype=torch.int64)
        sharding.exchange(send, recv)
        if rank > 0:
            msg = sharding.previous_rank_message(recv, rank, words)
            r_off, r_lists = sharding.message_views(msg, reach, bcap)
            plo, phi = bounds[rank - 1], bounds[rank]
            pbc = min(reach, phi - plo)
            pfirst = phi - pbc
            want_off = off[pfirst:phi + 1] - off[pfirst]
            assert np.array_equal(r_off[:pbc + 1].numpy(), want_off)
            assert np.array_equal(r_lists[:int(want_off[-1])].numpy(), succ[off[pfirst]:off[phi]])
        arcs, cs = g.scan_range(lo, hi)
        t = torch.tensor([arcs, cs - (1 << 64) if cs >= (1 << 63) else cs], dtype=torch.int64)
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        tot_arcs = sum(int(p[0]) for p in parts)
        x = 0
        for p in parts:
            x ^= int(p[1]) & 0xFFFFFFFFFFFFFFFF
        assert (tot_arcs, x) == g.scan_range(0, g.n)
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
            f.write("%d %d\n" % (lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shards_exchange_boundaries_over_gloo(tmp_path, world):
    off, succ, _ = graphs.copy_heavy(4000, seed=21)
    base = str(tmp_path / "g")
    tools.store_csr(base, off, succ)  # one writer thread: reference chains do cross the shard cuts
    port = 29600 + (os.getpid() % 300) + world
    mp.spawn(_rank_main, args=(world, base, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(str(tmp_path / ("ok%d" % r)))


def test_shard_plans_are_partitions_with_clean_cuts(tmp_path):
    """bvg_plan_shards / bvg_replan_shards (host only): monotone partitions of [0, n); the cuts sit where no reference crosses
    (checked against the oracle's successor lists through the compressor's own reference choices: a node whose list was
    written with a reference r copies from x - r, reference test fixture and a copy-heavy graph), and re-planning from a
    cost vector moves the cuts towards the cheaper shards."""
    from tests.conftest import CNR
    off, succ, _ = graphs.copy_heavy(20000, seed=3)
    base = str(tmp_path / "ch")
    tools.store_csr(base, off, succ)
    for b in (CNR, base):
        g = ob.load().load(b)
        for k in (1, 2, 3, 8):
            bounds = bvgraph.plan_shards(b, k)
            assert bounds[0] == 0 and bounds[-1] == g.n and all(x <= y for x, y in zip(bounds, bounds[1:]))
            for c in bounds[1:-1]:  # no chain crosses: every node of [c, c + W) decodes with nodes >= c alone
                hi = min(g.n, c + g.window)
                assert all(g.first_ancestor(x) >= c for x in range(c, hi)), (b, k, c)
        bounds = bvgraph.plan_shards(b, 4)
        again = bvgraph.replan_shards(b, bounds, [1.0, 1.0, 1.0, 5.0])
        assert again[0] == 0 and again[-1] == g.n and all(x <= y for x, y in zip(again, again[1:]))
        assert again[3] > bounds[3]  # the expensive last shard shrinks
