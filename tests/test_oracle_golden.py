"""Pins the CPU oracle on the reference's own golden pair (cnr-2000), i.e. restates
BVGraphTest.testLarge (reference test/it/unimi/dsi/webgraph/BVGraphTest.java:101-119):
the BVGraph cnr-2000.{graph,offsets,properties} must equal cnr-2000.graph-txt.gz both through the
sequential iterator and node by node through random access."""
import hashlib
import os

import numpy as np
import pytest

from tests.conftest import CNR
from tests import oracle_binding as ob

SHA = {
    ".graph": "b7d6b8bdf1218eb21edd77ce4ce091245050252972e687fc4d10a4d587f02db1",
    ".offsets": "c5268e312d8b4395518f85f6bd18f59049bb687e19b4307d45be08b3b81be3be",
    ".properties": "3fcb5ac1b1bd6505a30656726a737c7a13a3cf9e8739c9f4f18631902400dfef",
    ".graph-txt.gz": "ad05bc0dc8f826532a186a56279878eb34bbdd8ae8176e3cb0681bb571110f0a",
}


def test_fixture_integrity():
    for ext, h in SHA.items():
        with open(CNR + ext, "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == h, ext


def test_properties(oracle):
    g = oracle.load(CNR)
    assert (g.n, g.m, g.window, g.maxref, g.minlen, g.zetak, g.flags) == (325557, 3216152, 7, 3, 3, 3, 0)


def test_offsets(oracle):
    g = oracle.load(CNR)
    off = g.offsets()
    assert list(off[:6]) == [0, 85, 113, 130, 131, 151]
    assert off[-1] == 11443904 == g.graph_bytes * 8
    assert np.array_equal(g.rebuild_offsets(), off)  # BVGraph.writeOffsets equivalence


def test_sequential_equals_ascii(oracle, cnr_truth):
    toff, tsucc = cnr_truth
    g = oracle.load(CNR)
    off, succ = g.decode_range(0, g.n)
    assert np.array_equal(off, toff)
    assert np.array_equal(succ, tsucc)


def test_sequential_without_offsets(oracle, cnr_truth):
    toff, tsucc = cnr_truth
    g = oracle.load(CNR, offsets=False)
    off, succ = g.decode_range(0, g.n)
    assert np.array_equal(off, toff) and np.array_equal(succ, tsucc)
    with pytest.raises(ob.OracleError):
        g.decode_range(5, 10)  # BVGraph.java:1174
    with pytest.raises(ob.OracleError):
        g.successors(3)        # BVGraph.java:901


def test_random_access_equals_ascii(oracle, cnr_truth):
    toff, tsucc = cnr_truth
    g = oracle.load(CNR)
    rng = np.random.default_rng(1)
    nodes = np.concatenate([np.arange(0, 3000), rng.integers(0, g.n, 20000), [g.n - 1]])
    for x in nodes:
        x = int(x)
        d = g.outdegree(x)
        assert d == toff[x + 1] - toff[x]
        assert np.array_equal(g.successors(x, cap=max(d, 1)), tsucc[toff[x]:toff[x + 1]])


def test_from_every_kind_of_start(oracle, cnr_truth):
    toff, tsucc = cnr_truth
    g = oracle.load(CNR)
    for lo, hi in [(1, 50), (7, 8), (1000, 1200), (g.n - 10, g.n), (g.n, g.n), (12345, 12345)]:
        off, succ = g.decode_range(lo, hi)
        assert np.array_equal(off, toff[lo:hi + 1] - toff[lo])
        assert np.array_equal(succ, tsucc[toff[lo]:toff[hi]])


def test_known_answers_first_nodes(oracle):
    g = oracle.load(CNR)
    assert list(g.successors(0)) == [1, 342, 343, 344, 345, 346, 347, 348, 349, 350, 351, 211284, 223142]
    assert list(g.successors(1)) == [2, 3, 4, 319]
    assert list(g.successors(2)) == [211284, 223142]
    assert list(g.successors(3)) == []
    assert list(g.successors(4)) == [317]


def test_checksums(oracle, cnr_truth):
    toff, tsucc = cnr_truth
    g = oracle.load(CNR)
    arcs, cs = g.scan_range(0, g.n)
    assert arcs == 3216152
    assert int(tsucc.astype(np.int64).sum()) == 624313407838
    assert cs == ob.xor_checksum(toff, tsucc) == 0xf941dd3471d172f1
    arcs4, cs4 = g.scan_range(0, g.n, threads=4)
    assert (arcs4, cs4) == (arcs, cs)


def test_code_known_answers(oracle):
    # first bytes of cnr-2000.offsets decode (gamma) to 0,85,28,17,1,20
    vals, _ = oracle.read_codes(bytes.fromhex("81583a1241540a81"), 2, 0, 6)
    assert vals == [0, 85, 28, 17, 1, 20]
    # zeta_3 hand vectors: x=0 -> 100 ; x=1 -> 1010 ; x=6 -> 1111 ; x=7 -> 01 00000 (7 bits)
    vals, pos = oracle.read_codes(bytes([0b10010101, 0b11101000, 0b00000000]), 6, 3, 4)
    assert vals == [0, 1, 6, 7] and pos == 3 + 4 + 4 + 7
    # unary and delta: delta(0) = gamma(0) = "1"; delta(1): y=2, msb=1 -> gamma(1)="010", then "0"
    vals, pos = oracle.read_codes(bytes([0b00010000]), 5, 0, 1)
    assert vals == [3] and pos == 4
    vals, pos = oracle.read_codes(bytes([0b10100000]), 1, 0, 2)
    assert vals == [0, 1] and pos == 5
