"""The host-side compressor against the reference's bytes and through the oracle.

* re-encoding the decoded cnr-2000 with the fixture's parameters reproduces the reference's own .graph and
  .offsets byte for byte (BVGraph.store is deterministic: first strictly cheaper candidate wins, BVGraph.java:2313-2323);
* the BVGraphTest.testCompression matrix (reference test/it/unimi/dsi/webgraph/BVGraphTest.java:50-99) plus complete
  graphs round-trips through the oracle, with the same size/arc accounting identities;
* non-default codings round-trip (parity unpinned by the reference: no Java-made vector exists for them).
"""
import os

import numpy as np
import pytest

from tests import graphs
from tests import oracle_binding as ob
from tests.conftest import CNR
from webgraph_b200 import tools


def _props(path):
    out = {}
    for line in open(path):
        if "=" in line and not line.startswith("#"):
            k, v = line.strip().split("=", 1)
            out[k] = v
    return out


def test_cnr2000_reencode_is_byte_identical(tmp_path, cnr_truth):
    toff, tsucc = cnr_truth
    base = str(tmp_path / "re")
    st = tools.store_csr(base, toff, tsucc, window=7, maxref=3, minlen=3, zetak=3, threads=1)
    for ext in (".graph", ".offsets"):
        assert open(base + ext, "rb").read() == open(CNR + ext, "rb").read(), ext
    # SURVEY Appendix C statistics
    assert (st["copied_arcs"], st["intervalised_arcs"], st["residual_arcs"]) == (2130833, 361894, 723425)
    assert (st["bits_outdegrees"], st["bits_references"], st["bits_blocks"], st["bits_intervals"], st["bits_residuals"]) == \
        (1660205, 814229, 1504003, 866065, 6599402)
    assert st["max_ref_chain"] == 3 and st["max_outdegree"] == 2716
    assert st["xor_checksum"] == 0xf941dd3471d172f1


def _roundtrip(oracle, base, off, succ, **kw):
    st = tools.store_csr(base, off, succ, **kw)
    p = _props(base + ".properties")
    bits = sum(int(p[k]) for k in ("bitsforoutdegrees", "bitsforreferences", "bitsforblocks", "bitsforintervals", "bitsforresiduals"))
    assert os.path.getsize(base + ".graph") == (bits + 7) // 8               # BVGraphTest.java:66-74
    assert int(p["copiedarcs"]) + int(p["intervalisedarcs"]) + int(p["residualarcs"]) == len(succ)  # :76
    g = oracle.load(base)
    o2, s2 = g.decode_range(0, g.n)
    assert np.array_equal(o2, off) and np.array_equal(s2, succ)
    for x in range(g.n):  # random access route
        assert np.array_equal(g.successors(x, cap=max(1, int(off[x + 1] - off[x]))), succ[off[x]:off[x + 1]])
    return st


@pytest.mark.parametrize("kind", ["intree", "outtree", "complete"])
def test_compression_matrix(oracle, tmp_path, kind):
    for n in range(1, 8):
        off, succ = {"intree": graphs.binary_intree, "outtree": graphs.binary_outtree, "complete": graphs.complete_graph}[kind](n)
        for w in range(3):
            for r in range(1 if w == 0 else 3):
                for i in range(4):
                    _roundtrip(oracle, str(tmp_path / "g"), off, succ, window=w, maxref=r, minlen=i, zetak=3)


@pytest.mark.parametrize("n,p", [(5, .1), (10, .3), (100, .5), (100, .9)])
def test_erdos_renyi(oracle, tmp_path, n, p):  # ImmutableGraphTest.java:68-82
    off, succ = graphs.erdos_renyi(n, p, seed=n * 7 + int(p * 10))
    _roundtrip(oracle, str(tmp_path / "er"), off, succ)


@pytest.mark.parametrize("flags,k", [
    (0, 1), (0, 2), (0, 5),
    (tools.OUTDEGREES_DELTA | tools.BLOCKS_DELTA | tools.RESIDUALS_DELTA | tools.REFERENCES_DELTA | tools.BLOCK_COUNT_DELTA | tools.OFFSETS_DELTA, 3),
    (tools.RESIDUALS_GAMMA | tools.REFERENCES_GAMMA | tools.BLOCK_COUNT_UNARY | tools.BLOCKS_UNARY, 3),
    (tools.RESIDUALS_NIBBLE, 3), (tools.RESIDUALS_GOLOMB | tools.BLOCKS_DELTA, 3),
])
def test_non_default_codings_roundtrip(oracle, tmp_path, flags, k):
    off, succ, _ = graphs.copy_heavy(300, seed=5)
    _roundtrip(oracle, str(tmp_path / "f"), off, succ, flags=flags, zetak=k)


def test_threads_and_unbounded_chains(oracle, tmp_path):
    off, succ, _ = graphs.copy_heavy(2000, seed=11)
    st1 = _roundtrip(oracle, str(tmp_path / "a"), off, succ, threads=1)
    st4 = _roundtrip(oracle, str(tmp_path / "b"), off, succ, threads=4)
    assert st1["arcs"] == st4["arcs"] and st1["xor_checksum"] == st4["xor_checksum"]
    stu = _roundtrip(oracle, str(tmp_path / "c"), off, succ, maxref=-1)
    assert stu["max_ref_chain"] > 3  # -m -1: chains are not bounded (BVGraph.java:2689,2720)
    assert st1["max_ref_chain"] <= 3


def test_rejects_unsorted_lists(tmp_path):
    off = np.array([0, 3], dtype=np.int64)
    with pytest.raises(ValueError):
        tools.store_csr(str(tmp_path / "bad"), off, np.array([3, 2, 5], dtype=np.int32))
    with pytest.raises(ValueError):
        tools.store_csr(str(tmp_path / "bad"), off, np.array([2, 2, 5], dtype=np.int32))  # BVGraph.java:2201


def test_generator_is_deterministic_and_valid(oracle, tmp_path):
    st1, off1, succ1 = tools.generate_store(str(tmp_path / "g1"), 20000, 300000, seed=7, threads=1, return_csr=True)
    st3, off3, succ3 = tools.generate_store(str(tmp_path / "g3"), 20000, 300000, seed=7, threads=3, return_csr=True)
    assert np.array_equal(off1, off3) and np.array_equal(succ1, succ3)  # the graph does not depend on the thread count
    assert abs(st1["arcs"] - 300000) < 60000
    assert st1["xor_checksum"] == ob.xor_checksum(off1, succ1)
    for base in ("g1", "g3"):
        g = oracle.load(str(tmp_path / base))
        o, s = g.decode_range(0, g.n)
        assert np.array_equal(o, off1) and np.array_equal(s, succ1)
    assert st1["copied_arcs"] > 0 and st1["intervalised_arcs"] > 0 and st1["residual_arcs"] > 0
    assert st1["max_ref_chain"] == 3


def test_c1_graph_is_pinned(oracle, tmp_path):
    """BASELINE config C1: the 100 k-node synthetic power-law graph with BVGraph defaults, decoded by the CPU oracle, is the
    ground truth of C2 (the GPU decode of the same files, tests/test_gpu_parity.py).  The graph is this repo's own input
    (the reference ships no generator), pinned by tests/golden/synthetic/c1_100k.json so that it cannot drift unnoticed."""
    import hashlib
    import json
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "synthetic", "c1_100k.json")))
    p, e = fx["params"], fx["expect"]
    base = str(tmp_path / "c1")
    st, off, succ = tools.generate_store(base, p["n"], p["target_arcs"], seed=p["seed"], threads=p["threads"], return_csr=True)
    # ground truth first: the oracle's decode of the files == the generator's own adjacency lists, whatever they are
    g = oracle.load(base)
    o, s = g.decode_range(0, g.n)
    assert np.array_equal(o, off) and np.array_equal(s, succ)
    assert g.scan_range(0, g.n) == (int(st["arcs"]), int(st["xor_checksum"]))
    assert g.scan_range(0, g.n, threads=4) == (int(st["arcs"]), int(st["xor_checksum"]))
    # then the identity of the input.  The outdegree law goes through libm's exp/log; a host whose libm rounds one of
    # them differently makes a (valid) graph that differs in a few lists: that is reported, not failed.
    got = {k: int(st[k]) for k in ("nodes", "arcs", "graph_bits", "offsets_bits", "copied_arcs", "intervalised_arcs", "residual_arcs",
                                   "max_outdegree", "max_ref_chain", "xor_checksum", "sum_successors")}
    for ext in ("graph", "offsets"):
        got["sha256_" + ext] = hashlib.sha256(open(base + "." + ext, "rb").read()).hexdigest()
    if got != e:
        assert got["nodes"] == e["nodes"] and abs(got["arcs"] - e["arcs"]) < 0.01 * e["arcs"], "not the C1 graph at all"
        pytest.skip("generator output differs from the pinned C1 fixture on this host (libm rounding): %r" %
                    sorted(k for k in e if got[k] != e[k]))
