"""GPU parity tests: the CUDA path, called through the C ABI, against the reference's golden pair and the oracle.

Each test names the reference test it mirrors.  Bit-exact (integer work): every successor list must equal
BVGraph.successors() on the same .graph/.properties input.
"""
import os

import numpy as np
import pytest

from tests import graphs
from tests import oracle_binding as ob
from tests.conftest import CNR
from webgraph_b200 import bvgraph, tools
from webgraph_b200.bvgraph import BVGraph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cnr():
    g = BVGraph.load(CNR)
    yield g
    g.close()


# ---- BVGraphTest.testLarge (reference test/it/unimi/dsi/webgraph/BVGraphTest.java:101-119) ----

def test_cnr2000_properties(cnr):
    assert (cnr.numNodes(), cnr.numArcs(), cnr.windowSize(), cnr.maxRefCount(), cnr.minIntervalLength(), cnr.zetaK()) == \
        (325557, 3216152, 7, 3, 3, 3)
    assert cnr.randomAccess()
    assert cnr.extent()[2] == 3 and cnr.extent()[3] == 2716  # chain depth 3, max outdegree 2716


def test_cnr2000_sequential_equals_ascii(cnr, cnr_truth):
    toff, tsucc = cnr_truth
    off, succ = cnr.decodeRange(0, cnr.numNodes())
    assert np.array_equal(off, toff)
    assert np.array_equal(succ, tsucc)


def test_cnr2000_node_iterator(cnr, cnr_truth):
    toff, tsucc = cnr_truth
    it = cnr.nodeIterator()
    with pytest.raises(bvgraph.IllegalStateError):
        it.outdegree()  # BVGraph.java:1237
    x = -1
    while it.hasNext():
        x = it.nextInt()
        d = it.outdegree()
        assert d == toff[x + 1] - toff[x]
        if x % 97 == 0 or d > 500:
            assert np.array_equal(it.successorArray(), tsucc[toff[x]:toff[x + 1]])
    assert x == cnr.numNodes() - 1
    with pytest.raises(bvgraph.NoSuchElementError):
        it.nextInt()  # :1202


def test_cnr2000_random_access_all_nodes(cnr, cnr_truth):
    toff, tsucc = cnr_truth
    n = cnr.numNodes()
    perm = np.random.default_rng(3).permutation(n).astype(np.int32)
    off, succ = cnr.successorsBatch(perm)
    assert np.array_equal(np.diff(off), (toff[1:] - toff[:-1])[perm])
    # compare list by list through a gather of the truth
    idx = np.concatenate([np.arange(toff[x], toff[x + 1]) for x in perm[:20000]])
    assert np.array_equal(succ[:len(idx)], tsucc[idx])
    # checksum of everything: order-independent, covers all nodes
    assert ob.xor_checksum(toff, tsucc) == _checksum_batch(perm, off, succ)
    assert np.array_equal(cnr.outdegreeBatch(perm), np.diff(off).astype(np.int32))


def _checksum_batch(xs, off, succ):
    deg = np.diff(off)
    x64 = np.repeat(xs.astype(np.uint64), deg)
    with np.errstate(over="ignore"):
        v = x64 * np.uint64(0x9E3779B97F4A7C15) + succ.astype(np.int64).astype(np.uint64)
    return int(np.bitwise_xor.reduce(v)) if len(v) else 0


def test_cnr2000_single_node_calls(cnr, cnr_truth):
    toff, tsucc = cnr_truth
    for x in [0, 1, 2, 3, 4, 1000, 325556, 172345]:
        d = cnr.outdegree(x)
        assert d == toff[x + 1] - toff[x]
        it = cnr.successors(x)
        got = [it.nextInt() for _ in range(d + 2)]  # the reference test reads the terminating -1 (BVGraphTest.java:115)
        assert got[:d] == list(tsucc[toff[x]:toff[x + 1]]) and got[d:] == [-1, -1]
    assert list(cnr.successorArray(0)) == [1, 342, 343, 344, 345, 346, 347, 348, 349, 350, 351, 211284, 223142]
    assert list(cnr.successorArray(2)) == [211284, 223142]


def test_cnr2000_scan_checksum(cnr):
    arcs, cs = cnr.scanRange(0, cnr.numNodes())
    assert arcs == 3216152 and cs == 0xf941dd3471d172f1


def test_cnr2000_ranges_and_split_iterators(cnr, cnr_truth, oracle):
    toff, tsucc = cnr_truth
    n = cnr.numNodes()
    for lo, hi in [(1, 50), (7, 8), (1000, 1200), (n - 10, n), (n, n), (12345, 12345), (100000, 230000), (3, 4)]:
        off, succ = cnr.decodeRange(lo, hi)
        assert np.array_equal(off, toff[lo:hi + 1] - toff[lo])
        assert np.array_equal(succ, tsucc[toff[lo]:toff[hi]])
        arcs, cs = cnr.scanRange(lo, hi)
        assert arcs == toff[hi] - toff[lo]
        assert cs == ob.xor_checksum(toff[lo:hi + 1], tsucc[toff[lo]:toff[hi]], first_node=lo)
    # assertSplitIterator (WebGraphTestCase.java:66-103): every node exactly once, same lists
    for k in (1, 4, 7):
        seen = 0
        for it in cnr.splitNodeIterators(k):
            if it is None:
                continue
            while it.hasNext():
                x = it.nextInt()
                seen += 1
                if x % 1013 == 0:
                    assert np.array_equal(it.successorArray(), tsucc[toff[x]:toff[x + 1]])
        assert seen == n


def test_argument_errors(cnr):
    n = cnr.numNodes()
    with pytest.raises(ValueError):
        cnr.outdegree(n)          # BVGraph.java:860
    with pytest.raises(ValueError):
        cnr.successorArray(-1)    # :900
    with pytest.raises(ValueError):
        cnr.nodeIterator(n + 1)   # :1165
    with pytest.raises(ValueError):
        cnr.decodeRange(5, 4)
    g = BVGraph.loadOffline(CNR)  # no random access (:901), sequential still fine
    assert not g.randomAccess()
    with pytest.raises(bvgraph.UnsupportedOperationError):
        g.successorArray(3)
    with pytest.raises(bvgraph.IllegalStateError):
        g.outdegree(3)            # :869
    with pytest.raises(bvgraph.IllegalStateError):
        g.nodeIterator(5)         # :1174
    arcs, cs = g.scanRange(0, g.numNodes())
    assert arcs == 3216152 and cs == 0xf941dd3471d172f1
    g.close()


def test_sequential_graph_without_offsets_file(tmp_path, cnr_truth, monkeypatch):
    """loadSequential / loadOffline of a graph that has no .offsets (the reference never opens the file for
    offsetType <= 0, BVGraph.java:1581-1609; BVGraph.writeOffsets :2662-2676 is how it makes one): the record
    boundaries come from the .graph stream alone (bvg_boundaries.cuh) and must equal what .offsets holds."""
    import shutil
    toff, tsucc = cnr_truth
    base = str(tmp_path / "noff")
    shutil.copy(CNR + ".graph", base + ".graph")
    shutil.copy(CNR + ".properties", base + ".properties")
    with pytest.raises(IOError):
        BVGraph.load(base)  # random access needs the file
    for sub_bits in (None, 1 << 16):  # default sub-ranges; sub-ranges so short that most speculative entries are wrong
        if sub_bits:
            monkeypatch.setenv("BVG_BND_SUB_BITS", str(sub_bits))
        g = BVGraph.loadOffline(base)
        assert not g.randomAccess()
        with pytest.raises(bvgraph.UnsupportedOperationError):
            g.successorArray(3)
        assert g.scanRange(0, g.numNodes()) == (3216152, 0xf941dd3471d172f1)
        off, succ = g.decodeRange(0, g.numNodes())
        assert np.array_equal(off, toff) and np.array_equal(succ, tsucc)
        g.close()
    monkeypatch.delenv("BVG_BND_SUB_BITS")
    # non-default codings, no window, a record far longer than a sub-range; in-memory entry point with offsets = None
    off, succ, _ = graphs.copy_heavy(3000, seed=8)
    deg = np.zeros(3000, dtype=np.int64)
    deg[7] = 100000
    big_off = np.zeros(3001, dtype=np.int64)
    np.cumsum(deg, out=big_off[1:])
    big_succ = np.arange(0, 400000, 4, dtype=np.int32)
    cases = [(off, succ, dict(flags=0, window=7, maxref=3, minlen=4)), (off, succ, dict(flags=0, window=0, maxref=3, minlen=0)),
             (off, succ, dict(flags=tools.RESIDUALS_NIBBLE | tools.REFERENCES_GAMMA | tools.BLOCKS_DELTA, window=3, maxref=-1, minlen=2)),
             (big_off, big_succ, dict(flags=0, window=7, maxref=3, minlen=4))]
    monkeypatch.setenv("BVG_BND_SUB_BITS", "4096")
    for i, (o, s_, kw) in enumerate(cases):
        b = str(tmp_path / ("n%d" % i))
        tools.store_csr(b, o, s_, **kw)
        os.remove(b + ".offsets")
        g = BVGraph.loadSequential(b)
        o2, s2 = g.decodeRange(0, g.numNodes())
        assert np.array_equal(o2, o) and np.array_equal(s2, s_)
        g.close()
        g = BVGraph.fromMemory(open(b + ".graph", "rb").read(), None, len(o) - 1, len(s_), kw["window"], kw["maxref"], kw["minlen"],
                               flags=kw["flags"], offsetType=0)
        assert g.scanRange(0, g.numNodes())[0] == len(s_)
        g.close()
    # a stream cut short is an error, not a fault
    b = str(tmp_path / "n0")
    data = open(b + ".graph", "rb").read()
    with open(b + ".graph", "wb") as f:
        f.write(data[:len(data) // 2])
    with pytest.raises((IOError, bvgraph.FormatError)):
        BVGraph.loadSequential(b)


# ---- WebGraphTestCase.assertGraph (reference test/it/unimi/dsi/webgraph/WebGraphTestCase.java:158-260) ----

def assert_graph(g, off, succ):
    n = g.numNodes()
    assert n == len(off) - 1
    o2, s2 = g.decodeRange(0, n)
    assert np.array_equal(o2, off) and np.array_equal(s2, succ)
    it = g.nodeIterator()
    for x in range(n):
        assert it.hasNext() and it.nextInt() == x
        assert it.outdegree() == off[x + 1] - off[x]
        assert np.array_equal(it.successorArray(), succ[off[x]:off[x + 1]])
    assert not it.hasNext()
    if n:
        xs = np.arange(n, dtype=np.int32)
        bo, bs = g.successorsBatch(xs)
        assert np.array_equal(bo, off) and np.array_equal(bs, succ)
    # for every start s: nodeIterator(s) agrees with random access
    for s in range(0, n + 1, max(1, n // 16)):
        o3, s3 = g.decodeRange(s, n)
        assert np.array_equal(o3, off[s:] - off[s]) and np.array_equal(s3, succ[off[s]:])
    arcs, cs = g.scanRange(0, n)
    assert arcs == len(succ) and cs == ob.xor_checksum(off, succ)


@pytest.mark.parametrize("kind", ["intree", "outtree", "complete"])
def test_compression_matrix(tmp_path, kind):  # BVGraphTest.testCompression :50-99 (+ complete graphs)
    for n in range(1, 8):
        off, succ = {"intree": graphs.binary_intree, "outtree": graphs.binary_outtree, "complete": graphs.complete_graph}[kind](n)
        for w in range(3):
            for r in range(1 if w == 0 else 3):
                for i in (0, 1, 3):
                    base = str(tmp_path / "g")
                    tools.store_csr(base, off, succ, window=w, maxref=r, minlen=i, zetak=3)
                    for loader in (BVGraph.load, BVGraph.loadMapped):
                        g = loader(base)
                        assert_graph(g, off, succ)
                        g.close()


@pytest.mark.parametrize("n,p", [(5, .1), (10, .3), (100, .5), (100, .9)])
def test_erdos_renyi(tmp_path, n, p):  # ImmutableGraphTest.java:68-82
    off, succ = graphs.erdos_renyi(n, p, seed=n * 7 + int(p * 10))
    base = str(tmp_path / "er")
    tools.store_csr(base, off, succ)
    g = BVGraph.load(base)
    assert_graph(g, off, succ)
    g.close()


@pytest.mark.parametrize("flags,k", [
    (0, 1), (0, 2), (0, 5),
    (tools.OUTDEGREES_DELTA | tools.BLOCKS_DELTA | tools.RESIDUALS_DELTA | tools.REFERENCES_DELTA | tools.BLOCK_COUNT_DELTA | tools.OFFSETS_DELTA, 3),
    (tools.RESIDUALS_GAMMA | tools.REFERENCES_GAMMA | tools.BLOCK_COUNT_UNARY | tools.BLOCKS_UNARY, 3),
    (tools.RESIDUALS_NIBBLE, 3), (tools.RESIDUALS_GOLOMB | tools.BLOCKS_DELTA, 3),
])
def test_non_default_codings(tmp_path, oracle, flags, k):  # parity unpinned by the reference; checked against the oracle
    off, succ, _ = graphs.copy_heavy(1500, seed=5)
    base = str(tmp_path / "f")
    tools.store_csr(base, off, succ, flags=flags, zetak=k)
    g = BVGraph.load(base)
    assert_graph(g, off, succ)
    g.close()


def test_copy_heavy_chains_and_unbounded_refcount(tmp_path):
    off, succ, _ = graphs.copy_heavy(5000, seed=21, maxdeg=120)
    for maxref, w in [(3, 7), (-1, 7), (1, 1), (10, 16)]:
        base = str(tmp_path / ("c%d_%d" % (maxref, w)))
        st = tools.store_csr(base, off, succ, maxref=maxref, window=w)
        g = BVGraph.load(base)
        assert g.extent()[2] == st["max_ref_chain"]
        assert_graph(g, off, succ)
        g.close()


# ---- BASELINE configs C1/C2: 100 k-node synthetic power-law graph, full decode bit-exact ----

@pytest.fixture(scope="module")
def synth100k(tmp_path_factory):
    base = str(tmp_path_factory.mktemp("synth") / "pl100k")
    st, off, succ = tools.generate_store(base, 100000, 3000000, seed=0x5EED, return_csr=True, threads=4)
    return base, st, off, succ


def test_synthetic_100k_full_decode(synth100k, oracle):
    base, st, off, succ = synth100k
    g = BVGraph.load(base)
    o, s = g.decodeRange(0, g.numNodes())
    assert np.array_equal(o, off) and np.array_equal(s, succ)
    og = oracle.load(base)
    oo, os_ = og.decode_range(0, og.n)  # the oracle agrees with the generator's own lists
    assert np.array_equal(oo, off) and np.array_equal(os_, succ)
    arcs, cs = g.scanRange(0, g.numNodes())
    assert arcs == st["arcs"] and cs == st["xor_checksum"]
    g.close()


@pytest.mark.parametrize("env", [{"BVG_TILE": "1"}, {"BVG_TILE": "1", "BVG_TILE_NT": "512"}, {"BVG_TILE": "1", "BVG_LONG_D": "64", "BVG_LONG_SEG": "16", "BVG_LONG_CHUNK": "16"},
                                 {"BVG_STREAM": "1"}, {"BVG_STREAM": "1", "BVG_LONG_D": "64", "BVG_LONG_SEG": "16", "BVG_LONG_CHUNK": "16"}])
def test_alternative_scan_kernels(synth100k, cnr_truth, monkeypatch, env):
    """The two scan kernels of round 2 that are off by default (tile kernel, bvg_tile.cuh; stream-position extras kernel,
    bvg_stream.cuh): same (arcs, XOR checksum) as the truth on the reference's fixture and on the synthetic graph, whole graph,
    sub-ranges and shards with re-decoded halos."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    off, succ = cnr_truth
    g = BVGraph.load(CNR)
    assert g.scanRange(0, g.numNodes()) == (len(succ), ob.xor_checksum(off, succ))
    for lo, hi in [(0, 1), (17, 18), (1000, 200000), (325000, g.numNodes())]:
        assert g.scanRange(lo, hi) == (int(off[hi] - off[lo]), ob.xor_checksum(off[lo:hi + 1] - off[lo], succ[off[lo]:off[hi]], lo)), (lo, hi)
    g.close()
    base, st, soff, ssucc = synth100k
    g = BVGraph.load(base)
    assert g.scanRange(0, g.numNodes()) == (st["arcs"], st["xor_checksum"])
    g.close()
    bounds = bvgraph.plan_shards(base, 3)
    tot_a, tot_c = 0, 0
    for r in range(3):
        gs = BVGraph.loadShard(base, bounds[r], bounds[r + 1])
        a, c = gs.scanRange(bounds[r], bounds[r + 1])
        tot_a += a
        tot_c ^= c
        gs.close()
    assert (tot_a, tot_c) == (st["arcs"], st["xor_checksum"])


def _bfs_numpy(off, succ, source):
    n = len(off) - 1
    dist = np.full(n, -1, dtype=np.int32)
    dist[source] = 0
    frontier = np.array([source], dtype=np.int64)
    level = 0
    while len(frontier):
        starts, ends = off[frontier], off[frontier + 1]
        idx = np.concatenate([np.arange(a, b) for a, b in zip(starts, ends)]) if len(frontier) else np.zeros(0, dtype=np.int64)
        nxt = np.unique(succ[idx]) if len(idx) else np.zeros(0, dtype=np.int64)
        nxt = nxt[dist[nxt] < 0]
        level += 1
        dist[nxt] = level
        frontier = nxt.astype(np.int64)
    return dist


def test_fused_consumers_indegrees_and_bfs(cnr, cnr_truth, synth100k):
    """the counting pass of a transposition (Transform.java:977-987) done by the scan itself, and a breadth-first
    visit (ParallelBreadthFirstVisit.java:155-181) over device-side random access, against numpy on the truth."""
    off, succ = cnr_truth
    n = len(off) - 1
    assert np.array_equal(cnr.indegrees(), np.bincount(succ, minlength=n).astype(np.uint32))
    lo, hi = 1000, 200000   # a sub-range: only the arcs leaving [lo, hi)
    assert np.array_equal(cnr.indegrees(lo, hi), np.bincount(succ[off[lo]:off[hi]], minlength=n).astype(np.uint32))
    for source in (0, 12345, 325556):
        dist, levels, reached = cnr.bfs(source)
        want = _bfs_numpy(off, succ, source)
        assert np.array_equal(dist, want), source
        assert levels == int(want.max()) and reached == int((want >= 0).sum())
    base, st, soff, ssucc = synth100k
    g = BVGraph.load(base)
    assert np.array_equal(g.indegrees(), np.bincount(ssucc, minlength=g.numNodes()).astype(np.uint32))   # long records included
    dist, levels, reached = g.bfs(7)
    assert np.array_equal(dist, _bfs_numpy(soff, ssucc, 7))
    g.close()
    # shards add up (what an all-reduce of the ranks' counts gives)
    bounds = bvgraph.plan_shards(base, 3)
    tot = np.zeros(len(soff) - 1, dtype=np.uint32)
    for r in range(3):
        gs = BVGraph.loadShard(base, bounds[r], bounds[r + 1])
        tot += gs.indegrees()
        gs.close()
    assert np.array_equal(tot, np.bincount(ssucc, minlength=len(soff) - 1).astype(np.uint32))


def test_weblike_1m_full_parity(tmp_path, oracle):
    """The second benchmark workload (copy-heavy, cnr-2000's mix: ~3.6 bits/arc, ~63 % copied arcs, avgref ~1.3) at 1 M nodes:
    every list against the generator's own and against the oracle; scan checksum and sum of successors."""
    import bench
    base = str(tmp_path / "web1m")
    st, off, succ = tools.generate_store(base, 1000000, 14000000, seed=0x5EED, return_csr=True, threads=8, **bench.WEBLIKE)
    assert st["copied_arcs"] > 0.5 * st["arcs"] and st["tot_ref"] > 1.1 * st["nodes"] and st["graph_bits"] < 6 * st["arcs"]
    g = BVGraph.load(base)
    o, s = g.decodeRange(0, g.numNodes())
    assert np.array_equal(o, off) and np.array_equal(s, succ)
    assert g.scanRange(0, g.numNodes()) == (st["arcs"], st["xor_checksum"])
    assert int(s.astype(np.uint64).sum()) == st["sum_successors"]
    g.close()
    og = oracle.load(base)
    assert og.scan_range(0, og.n, threads=8) == (st["arcs"], st["xor_checksum"])
    lo, hi = 400000, 402000
    oo, os_ = og.decode_range(lo, hi)
    assert np.array_equal(os_, succ[off[lo]:off[hi]])


def test_synthetic_100k_random_access(synth100k):
    base, st, off, succ = synth100k
    g = BVGraph.load(base)
    xs = np.random.default_rng(9).integers(0, g.numNodes(), 50000).astype(np.int32)
    bo, bs = g.successorsBatch(xs)
    assert np.array_equal(np.diff(bo), (off[1:] - off[:-1])[xs])
    assert _checksum_batch(xs, bo, bs) == _checksum_batch(xs, bo, np.concatenate([succ[off[x]:off[x + 1]] for x in xs]))
    g.close()


# ---- range sharding (assertSplitIterator + BVGraph.java:1173-1183) ----

@pytest.mark.parametrize("shards", [2, 4, 8])
def test_shards_halo_redecode_and_import(synth100k, shards):
    base, st, off, succ = synth100k
    n = len(off) - 1
    cuts = [n * i // shards for i in range(shards + 1)]
    cuts[1] = min(cuts[1] + 3, cuts[2])  # an uneven cut for good measure
    tot_arcs, tot_cs = 0, 0
    prev = None
    for i in range(shards):
        g = BVGraph.loadShard(base, cuts[i], cuts[i + 1])
        lo, hi = cuts[i], cuts[i + 1]
        assert g.extent()[:2] == (lo, hi)
        o, s = g.decodeRange(lo, hi)  # halo re-decoded from the shard's own bits
        assert np.array_equal(o, off[lo:hi + 1] - off[lo]) and np.array_equal(s, succ[off[lo]:off[hi]])
        if prev is not None:  # the previous shard's exported boundary lists replace the re-decode
            import ctypes as C
            cnt, boff, blists = prev
            bvgraph._check(bvgraph.lib().bvg_halo_import(g.handle, cnt, boff.ctypes.data, blists.ctypes.data, 0))
            o2, s2 = g.decodeRange(lo, hi)
            assert np.array_equal(o2, o) and np.array_equal(s2, s)
        arcs, cs = g.scanRange(lo, hi)
        tot_arcs += arcs
        tot_cs ^= cs
        import ctypes as C
        cnt = C.c_int32()
        bvgraph._check(bvgraph.lib().bvg_boundary_count(g.handle, C.byref(cnt)))
        assert cnt.value == min(21, hi - lo)
        boff = np.zeros(cnt.value + 1, dtype=np.int64)
        cap = int(off[hi] - off[hi - cnt.value])
        blists = np.empty(max(cap, 1), dtype=np.int32)
        bvgraph._check(bvgraph.lib().bvg_boundary_export(g.handle, boff.ctypes.data, blists.ctypes.data, cap, 0))
        assert np.array_equal(blists[:cap], succ[off[hi - cnt.value]:off[hi]])
        prev = (cnt.value, boff, blists)
        g.close()
    assert tot_arcs == st["arcs"] and tot_cs == st["xor_checksum"]


def test_size_independent_properties_on_a_larger_graph(tmp_path):
    """No oracle at this size in the test budget: 1 M nodes / ~33 M arcs (long records, every chain depth) checked through
    properties that do not depend on size -- the scan of the whole equals the generator's own (arcs, checksum), equals the
    XOR / sum over uneven sub-range scans, equals the checksum recomputed on the device from the materialised CSR; rows are
    strictly increasing; random-access rows equal the CSR rows.  bench.py checks the first property at the full 1 B arcs."""
    import ctypes as C
    import torch
    base = str(tmp_path / "pl1m")
    st = tools.generate_store(base, 1_000_000, 33_000_000, seed=0xBEEF, threads=os.cpu_count() or 4)
    g = BVGraph.load(base)
    n = g.numNodes()
    assert g.scanRange(0, n) == (st["arcs"], st["xor_checksum"])
    cuts = [0, 1, 7, 1000, 333_333, 333_340, 700_001, n - 3, n]
    parts = [g.scanRange(a, b) for a, b in zip(cuts, cuts[1:])]
    x = 0
    for _, c in parts:
        x ^= c
    assert sum(a for a, _ in parts) == st["arcs"] and x == st["xor_checksum"]
    L = bvgraph.lib()
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    d_succ = torch.empty(int(st["arcs"]), dtype=torch.int32, device="cuda")
    bvgraph._check(L.bvg_decode_range(g.handle, 0, n, d_off.data_ptr(), d_succ.data_ptr(), int(st["arcs"]), 1))
    torch.cuda.synchronize()
    assert int(d_off[-1]) == st["arcs"]
    deg = d_off[1:] - d_off[:-1]
    src = torch.repeat_interleave(torch.arange(n, device="cuda", dtype=torch.int64), deg)
    # x * MIX + y mod 2^64 in int64 two's complement (torch has no uint64 arithmetic): wrap-around is what we want
    MIX = 0x9E3779B97F4A7C15 - (1 << 64)
    terms = src * MIX + (d_succ.to(torch.int64) & 0xFFFFFFFF)
    acc = terms
    while acc.numel() > 1:  # XOR-reduce by halving
        h = acc.numel() // 2
        rest = acc[2 * h:]
        acc = torch.bitwise_xor(acc[:h], acc[h:2 * h])
        if rest.numel():
            acc = torch.cat([acc, rest])
    assert (int(acc[0]) & 0xFFFFFFFFFFFFFFFF) == st["xor_checksum"]
    assert int(d_succ.to(torch.int64).sum()) == st["sum_successors"]
    # strictly increasing inside every row: the only non-increasing steps are at row starts
    drops = torch.nonzero(d_succ[1:] <= d_succ[:-1]).flatten() + 1
    starts = d_off[1:-1][deg[1:] > 0]
    assert torch.isin(drops, starts).all()
    # random access == sequential
    xs = np.random.default_rng(5).integers(0, n, 20000).astype(np.int32)
    boff, bsucc = g.successorsBatch(xs)
    h_off = d_off.cpu().numpy()
    h_succ = d_succ.cpu().numpy()
    ref = np.concatenate([h_succ[h_off[v]:h_off[v + 1]] for v in xs])
    assert np.array_equal(bsucc, ref) and np.array_equal(np.diff(boff), h_off[xs + 1] - h_off[xs])
    g.close()


def test_cursor_batches_zero_copy_and_concurrent(cnr, cnr_truth, synth100k):
    """bvg_cursor_next_batch: whole batches as views of the cursor's pinned memory; cursors of one graph drained by several
    host threads at once (each has its own stream and error word), split like ImmutableGraph.splitNodeIterators."""
    import concurrent.futures as cf
    toff, tsucc = cnr_truth
    it = cnr.nodeIterator(1234)
    seen = 1234
    while True:
        b = it.nextBatch()
        if b is None:
            break
        first, off, succ = b
        assert first == seen
        cnt = len(off) - 1
        assert np.array_equal(off - off[0], toff[first:first + cnt + 1] - toff[first])
        assert np.array_equal(succ[off[0]:off[-1]], tsucc[toff[first]:toff[first + cnt]])
        seen += cnt
    assert seen == cnr.numNodes()
    base, st, off, succ = synth100k
    g = BVGraph.load(base)
    L = bvgraph.lib()
    import ctypes as C

    def drain(lo, hi):
        cur = C.c_void_p()
        bvgraph._check(L.bvg_cursor_open(g.handle, lo, hi, C.byref(cur)))
        cn, ca, cc = C.c_int64(), C.c_int64(), C.c_uint64()
        bvgraph._check(L.bvg_cursor_drain(cur, -1, C.byref(cn), C.byref(ca), C.byref(cc)))
        L.bvg_cursor_close(cur)
        return cn.value, ca.value, cc.value
    k = 6
    n = g.numNodes()
    step = (n + k - 1) // k
    with cf.ThreadPoolExecutor(k) as ex:
        res = list(ex.map(lambda i: drain(i * step, min(n, (i + 1) * step)), range(k)))
    assert sum(r[0] for r in res) == n and sum(r[1] for r in res) == st["arcs"]
    cs = 0
    for r in res:
        cs ^= r[2]
    assert cs == st["xor_checksum"]
    g.close()


def test_cursor_drain_equals_scan(cnr, synth100k):
    """bvg_cursor_drain (the C loop a binding runs over bvg_cursor_next: double-buffered batches in pinned memory) consumes
    exactly what the consume-only scan does; mixing single steps, partial drains and a copied cursor keeps the position."""
    import ctypes as C
    L = bvgraph.lib()
    for g, want in ((cnr, (3216152, 0xf941dd3471d172f1)), (BVGraph.load(synth100k[0]), (synth100k[1]["arcs"], synth100k[1]["xor_checksum"]))):
        n = g.numNodes()
        h = C.c_void_p()
        bvgraph._check(L.bvg_cursor_open(g.handle, 0, 2 ** 31 - 1, C.byref(h)))
        nodes, arcs, cs = C.c_int64(), C.c_int64(), C.c_uint64()
        bvgraph._check(L.bvg_cursor_drain(h, -1, C.byref(nodes), C.byref(arcs), C.byref(cs)))
        assert (nodes.value, arcs.value, cs.value) == (n, want[0], want[1])
        assert L.bvg_cursor_next(h, None, None, None) == -8  # BVG_EEND, NoSuchElementException in the reference
        L.bvg_cursor_close(h)
        # from the middle, in pieces that straddle batch boundaries (65536 nodes), with a copy taken on the way
        frm = 70000 if n > 200000 else n // 3
        bvgraph._check(L.bvg_cursor_open(g.handle, frm, 2 ** 31 - 1, C.byref(h)))
        tot_n, tot_a, x = 0, 0, 0
        bvgraph._check(L.bvg_cursor_drain(h, 61000, C.byref(nodes), C.byref(arcs), C.byref(cs)))
        tot_n += nodes.value; tot_a += arcs.value; x ^= cs.value
        h2 = C.c_void_p()
        bvgraph._check(L.bvg_cursor_copy(h, 2 ** 31 - 1, C.byref(h2)))
        bvgraph._check(L.bvg_cursor_drain(h, -1, C.byref(nodes), C.byref(arcs), C.byref(cs)))
        rest = (nodes.value, arcs.value, cs.value)
        bvgraph._check(L.bvg_cursor_drain(h2, -1, C.byref(nodes), C.byref(arcs), C.byref(cs)))
        assert rest == (nodes.value, arcs.value, cs.value)
        tot_n += rest[0]; tot_a += rest[1]; x ^= rest[2]
        assert tot_n == n - frm and (tot_a, x) == g.scanRange(frm, n)
        L.bvg_cursor_close(h)
        L.bvg_cursor_close(h2)
        if g is not cnr:
            g.close()


def test_scan_memory_pipelined_pieces(synth100k):
    """bvg_scan_memory (one pass over a graph in host memory, piece p + 1 uploaded while piece p is indexed and scanned):
    every number of pieces and every sub-range gives what bvg_scan_range gives on the opened graph."""
    import ctypes as C
    base, st, off, succ = synth100k
    graph = np.fromfile(base + ".graph", dtype=np.uint8)
    offs = np.fromfile(base + ".offsets", dtype=np.uint8)
    n = len(off) - 1
    L = bvgraph.lib()
    g = BVGraph.load(base)
    a, c = C.c_int64(), C.c_uint64()
    for frm, to in ((0, n), (0, n // 3), (n // 3 + 7, n - 11), (5, 6)):
        want = g.scanRange(frm, to)
        for pieces in (1, 2, 5, 16):
            bvgraph._check(L.bvg_scan_memory(graph.ctypes.data, len(graph), offs.ctypes.data, len(offs), n, int(st["arcs"]),
                                             7, 3, 4, 3, 0, 0, frm, to, pieces, C.byref(a), C.byref(c)))
            assert (a.value, c.value) == want, (frm, to, pieces)
    assert g.scanRange(0, n) == (st["arcs"], st["xor_checksum"])
    g.close()
    assert L.bvg_scan_memory(graph.ctypes.data, len(graph), offs.ctypes.data, len(offs), n, int(st["arcs"]),
                             7, 3, 4, 3, 0, 0, 0, n + 1, 2, C.byref(a), C.byref(c)) == -1  # BVG_EINVAL


def test_halo_import_from_device_buffers_twice(synth100k):
    """What bench.py does every step at N > 1: boundary lists exported into device buffers, imported from device buffers.
    The first import sizes the halo buffers (one round trip), every later one of the same shape is a device-side copy."""
    import ctypes as C
    import torch
    base, st, off, succ = synth100k
    n = len(off) - 1
    cut = n // 2 + 5
    a = BVGraph.loadShard(base, 0, cut)
    b = BVGraph.loadShard(base, cut, n)
    L = bvgraph.lib()
    cnt = C.c_int32()
    bvgraph._check(L.bvg_boundary_count(a.handle, C.byref(cnt)))
    cap = int(off[cut] - off[cut - cnt.value])
    d_off = torch.zeros(cnt.value + 1, dtype=torch.int64, device="cuda")
    d_lists = torch.zeros(max(cap, 1), dtype=torch.int32, device="cuda")
    want = (int(st["arcs"]) - int(off[cut]), None)
    results = []
    for rep in range(3):
        d_lists.zero_()
        bvgraph._check(L.bvg_boundary_export(a.handle, d_off.data_ptr(), d_lists.data_ptr(), cap, 1))
        bvgraph._check(L.bvg_halo_import(b.handle, cnt.value, d_off.data_ptr(), d_lists.data_ptr(), 1))
        o, s = b.decodeRange(cut, n)
        assert np.array_equal(o, off[cut:n + 1] - off[cut]) and np.array_equal(s, succ[off[cut]:off[n]])
        results.append(b.scanRange(cut, n))
    torch.cuda.synchronize()
    assert np.array_equal(d_lists.cpu().numpy()[:cap], succ[off[cut - cnt.value]:off[cut]])
    assert results[0][0] == want[0] and results[0] == results[1] == results[2]
    arcs_a, cs_a = a.scanRange(0, cut)
    assert arcs_a + results[0][0] == st["arcs"] and (cs_a ^ results[0][1]) == st["xor_checksum"]
    a.close()
    b.close()


# ---- corrupt input: error code + node, never a fault (BVGraph.java:705, 1129-1131) ----

def test_corrupt_stream_reports_error(tmp_path):
    off, succ, _ = graphs.copy_heavy(3000, seed=2)
    base = str(tmp_path / "ok")
    tools.store_csr(base, off, succ)
    graph = bytearray(open(base + ".graph", "rb").read())
    offs = open(base + ".offsets", "rb").read()
    rng = np.random.default_rng(0)
    failures = 0
    for trial in range(12):
        bad = bytearray(graph)
        for _ in range(40):
            bad[int(rng.integers(0, len(bad)))] ^= int(rng.integers(1, 256))
        try:
            g = BVGraph.fromMemory(bytes(bad), offs, len(off) - 1, len(succ), 7, 3, 4)
            try:
                g.decodeRange(0, g.numNodes())
                g.successorsBatch(np.arange(g.numNodes(), dtype=np.int32))
            finally:
                g.close()
        except (bvgraph.IllegalStateError, IOError, MemoryError):
            failures += 1
    assert failures > 0  # some corruption is detected; none of it crashes or hangs the device
    g = BVGraph.fromMemory(bytes(graph), offs, len(off) - 1, len(succ), 7, 3, 4)  # the device is still healthy
    o, s = g.decodeRange(0, g.numNodes())
    assert np.array_equal(s, succ)
    g.close()
    with pytest.raises(IOError):  # truncated stream
        BVGraph.fromMemory(bytes(graph[:len(graph) // 2]), offs, len(off) - 1, len(succ), 7, 3, 4)


# ---- extremes (scaled-down BVGraphSlowTest, reference slow/it/unimi/dsi/webgraph/BVGraphSlowTest.java:30-96) ----

def test_extremes_large_ids_and_giant_lists(tmp_path):
    n = 3_000_000
    lists = {0: np.arange(0, n, 4, dtype=np.int32), 1: np.arange(1, n, 4, dtype=np.int32),
             2: np.array([n - 1], dtype=np.int32), n - 1: np.array([0, 1, n - 2], dtype=np.int32),
             n - 2: np.arange(n - 5000, n, dtype=np.int32)}
    deg = np.zeros(n, dtype=np.int64)
    for k, v in lists.items():
        deg[k] = len(v)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=off[1:])
    succ = np.concatenate([lists[k] for k in sorted(lists)])
    base = str(tmp_path / "big")
    tools.store_csr(base, off, succ, threads=4)
    g = BVGraph.load(base)
    assert g.extent()[3] == len(lists[0])
    o, s = g.decodeRange(0, 3)
    assert np.array_equal(s, succ[:off[3]])
    o, s = g.decodeRange(n - 3, n)
    assert np.array_equal(s, succ[off[n - 3]:])
    bo, bs = g.successorsBatch(np.array([n - 1, 0, n - 2, 5], dtype=np.int32))
    assert np.array_equal(bs, np.concatenate([lists[n - 1], lists[0], lists[n - 2]]))
    arcs, cs = g.scanRange(0, n)
    assert arcs == len(succ) and cs == ob.xor_checksum(off, succ)
    g.close()


# ---- the C++ mirror of the reference interface (include/bvgraph_b200.hpp) ----

def test_cpp_mirror(tmp_path):
    import shutil
    import subprocess
    from webgraph_b200 import build
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    exe = str(tmp_path / "cpp_mirror_test")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(build.cuda_library())
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp_mirror_test.cpp"),
                           "-L", libdir, "-lbvgraph_b200", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe, CNR], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["3216152", "f941dd3471d172f1", "1", "-1", "1", "3216152", "f941dd3471d172f1"]
    # EFGraph and ArcLabels of the mirror: cnr-2000 stored as an EFGraph, gamma labels j % 1000 over the fixture itself
    off, succ = ob.read_ascii_graph(CNR + ".graph-txt.gz")
    ef = str(tmp_path / "cnr-ef")
    tools.store_ef(ef, off, succ, threads=4)
    lab = str(tmp_path / "cnr-lab")
    tools.store_labels(lab, CNR, off, (np.arange(len(succ)) % 1000).astype(np.int32), tools.LABEL_GAMMA, threads=4)
    out = subprocess.run([exe, CNR, ef, lab], capture_output=True, text=True, check=True).stdout.split()
    assert out[7:] == ["3216152", "f941dd3471d172f1", "1", "1"]


@pytest.mark.gpu
def test_fused_consumer_hyperball_step(tmp_path, cnr_truth):
    """HyperBall's inner loop (HyperBall.java:875-915; its own assert block :918-929 states the property): every register of a
    node becomes the maximum over the node and its successors."""
    from webgraph_b200.bvgraph import BVGraph

    def expect(off, succ, cin):
        n = len(off) - 1
        src = np.repeat(np.arange(n), np.diff(off))
        out = cin.copy()
        np.maximum.at(out, src, cin[succ])
        return out, int(np.any(out != cin, axis=1).sum())

    rng = np.random.default_rng(7)
    off, succ = cnr_truth
    g = BVGraph.load(CNR)
    for log2m in (4, 6):
        cin = rng.integers(0, 32, (g.numNodes(), 1 << log2m), dtype=np.uint8)
        want, wmod = expect(off, succ, cin)
        got, mod = g.hyperballStep(cin, log2m)
        assert np.array_equal(got, want) and mod == wmod
        # iterating to the fixed point keeps agreeing (two more rounds)
        for _ in range(2):
            want, wmod = expect(off, succ, want)
            got, mod = g.hyperballStep(got, log2m)
            assert np.array_equal(got, want) and mod == wmod
    # a sub-range leaves the other rows alone
    cin = rng.integers(0, 32, (g.numNodes(), 16), dtype=np.uint8)
    want, _ = expect(off, succ, cin)
    got, mod = g.hyperballStep(cin, 4, 1000, 300000)
    assert np.array_equal(got[1000:300000], want[1000:300000]) and np.array_equal(got[:1000], cin[:1000]) and np.array_equal(got[300000:], cin[300000:])
    assert mod == int(np.any(want[1000:300000] != cin[1000:300000], axis=1).sum())
    with pytest.raises(ValueError):
        g.hyperballStep(cin, 5)   # shape does not match 2^log2m
    g.close()
    # a node with 60 000 successors (the block-per-node path), empty nodes, a self-loop, 512 registers
    n = 70000
    deg = np.zeros(n, dtype=np.int64)
    deg[5] = 60000
    deg[100:140] = 7
    deg[n - 1] = 2
    off2 = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=off2[1:])
    lists = [np.arange(60000, dtype=np.int32)] + [np.arange(x, x + 7, dtype=np.int32) for x in range(100, 140)] + [np.array([3, n - 1], dtype=np.int32)]
    succ2 = np.concatenate(lists)
    base = str(tmp_path / "hb")
    tools.store_csr(base, off2, succ2)
    g = BVGraph.load(base)
    for log2m in (4, 9):
        cin = rng.integers(0, 60, (n, 1 << log2m), dtype=np.uint8)
        want, wmod = expect(off2, succ2, cin)
        got, mod = g.hyperballStep(cin, log2m)
        assert np.array_equal(got, want) and mod == wmod
    g.close()
