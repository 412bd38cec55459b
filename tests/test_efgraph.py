"""EFGraph (decode half): the reference's quasi-succinct format on the device.

PARITY UNPINNED: the reference ships no EFGraph fixture (EFGraphTest.java stores its graphs at run time) and there is no JVM
here.  The writer (bvgt_store_ef, restating Accumulator / LongWordOutputBitStream), the oracle (restating LongWordBitReader /
EliasFanoSuccessorReader) and the kernels are independent restatements checked against each other, against the graphs' CSR
and against the BVGraph decode of the same graphs; the skip pointers, which enumeration never reads, are checked against the
property skipTo relies on (EFGraph.java:1160-1175)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import graphs
from tests.conftest import CNR, ROOT
from webgraph_b200 import tools

EMU_DIR = os.path.join(ROOT, "tests", "hostemu")
EMU_EF = os.path.join(EMU_DIR, "libemu_ef.so")


def skewed():
    n = 70000
    deg = np.zeros(n, dtype=np.int64)
    deg[5] = 60000
    deg[100:140] = 7
    deg[n - 1] = 2
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=off[1:])
    lists = [np.arange(0, 60000, dtype=np.int32)] + [np.arange(x, x + 7, dtype=np.int32) for x in range(100, 140)] + [np.array([3, n - 1], dtype=np.int32)]
    return off, np.concatenate(lists)


CASES = [("er10", lambda: graphs.erdos_renyi(10, .5, 1)), ("er100", lambda: graphs.erdos_renyi(100, .3, 2)), ("er100d", lambda: graphs.erdos_renyi(100, .9, 3)),
         ("complete", lambda: graphs.complete_graph(40, loops=True)), ("copy", lambda: graphs.copy_heavy(1500, seed=5)[:2]), ("skew", skewed),
         ("empty", lambda: (np.zeros(1, dtype=np.int64), np.empty(0, dtype=np.int32))),
         ("arcless", lambda: (np.zeros(51, dtype=np.int64), np.empty(0, dtype=np.int32)))]


def store(tmp_path, name, off, succ, q=8, big=False, threads=1):
    base = str(tmp_path / name)
    ub = max(len(off) - 1, int(succ.max()) + 1 if len(succ) else 0)
    bits = tools.store_ef(base, off, succ, upper_bound=ub, log2_quantum=q, big_endian=big, threads=threads)
    return base, ub, bits


def ef_geometry(d, ub, q):
    length = d + 1
    quot = ub // length
    l = 0 if quot == 0 else quot.bit_length() - 1
    ulen = length + (ub >> l)
    psize = 0 if ulen <= 1 else (ulen - 1).bit_length()
    return l, psize, (ub >> l) >> q, ulen


def getbits(words, pos, width):
    v = 0
    for i in range(width):
        p = pos + i
        v |= ((int(words[p >> 6]) >> (p & 63)) & 1) << i
    return v


@pytest.mark.parametrize("q,big", [(8, False), (0, False), (3, True)])
def test_writer_and_oracle_agree_and_skip_pointers_hold(tmp_path, oracle, q, big):
    for name, make in CASES:
        off, succ = make()
        base, ub, bits = store(tmp_path, "%s_%d" % (name, q), off, succ, q=q, big=big, threads=3)
        props = open(base + ".properties").read()
        assert "graphclass=it.unimi.dsi.webgraph.EFGraph" in props and "quantum=%d" % (1 << q) in props
        assert ("byteorder=BIG_ENDIAN" if big else "byteorder=LITTLE_ENDIAN") in props
        assert os.path.getsize(base + ".graph") == 8 * (bits // 64 + 1)   # LongWordOutputBitStream.close() always writes the buffer
        g = oracle.load_ef(base)
        assert (g.n, g.m, g.upper_bound, g.log2_quantum) == (len(off) - 1, len(succ), ub, q)
        assert g.offsets()[-1] == bits
        o, s = g.decode_range(0, g.n)
        assert np.array_equal(o, off) and np.array_equal(s, succ), name
        # skip pointers (small graphs): pointer b - 1 = number of upper bits before which exactly b * quantum zeros lie, and the
        # bit just before it is one of them (what skipTo positions on, EFGraph.java:1160-1175)
        if g.n <= 100:
            words = np.ctypeslib.as_array(g.g.contents.words, shape=(int(g.g.contents.nwords) + 2,))
            offs = g.offsets()
            for x in range(g.n):
                d = int(off[x + 1] - off[x])
                pos = int(offs[x])
                msb = 0
                while getbits(words, pos, 1) == 0:
                    pos += 1
                    msb += 1
                pos += 1 + msb
                l, psize, npointers, ulen = ef_geometry(d, ub, q)
                upper_start = pos + psize * npointers + l * (d + 1)
                assert upper_start + ulen == int(offs[x + 1]), (name, x)
                upper = [getbits(words, upper_start + i, 1) for i in range(ulen)]
                assert sum(upper) == d + 1
                zeros_before = np.concatenate([[0], np.cumsum(1 - np.array(upper))])
                for b in range(1, npointers + 1):
                    p = getbits(words, pos + (b - 1) * psize, psize)
                    assert zeros_before[p] == b << q and upper[p - 1] == 0, (name, x, b)
        g.close()


def test_writer_rejects_what_the_accumulator_rejects(tmp_path):
    off = np.array([0, 3], dtype=np.int64)
    with pytest.raises(ValueError):   # not strictly increasing (Accumulator.add, EFGraph.java:499)
        tools.store_ef(str(tmp_path / "b"), off, np.array([0, 0, 0], dtype=np.int32), upper_bound=5)
    with pytest.raises(ValueError):   # prefix sum above the upper bound (:502)
        tools.store_ef(str(tmp_path / "b"), off, np.array([0, 1, 9], dtype=np.int32), upper_bound=5)


@pytest.fixture(scope="module")
def emu_ef():
    cuda_dir = os.path.join(ROOT, "webgraph_b200", "csrc", "cuda")
    srcs = [os.path.join(EMU_DIR, "emu_ef.cpp"), os.path.join(EMU_DIR, "cuda_shim.h"), os.path.join(cuda_dir, "bvg_ef.cuh"), os.path.join(cuda_dir, "bvg_device.cuh")]
    if not os.path.exists(EMU_EF) or any(os.path.getmtime(s) > os.path.getmtime(EMU_EF) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-std=c++17", "-fPIC",
                               "-shared", "-I" + EMU_DIR, "-o", EMU_EF, srcs[0]])
    lib = C.CDLL(EMU_EF)
    lib.emu_ef_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_ulonglong)]
    return lib


def xor_checksum(off, succ):
    src = np.repeat(np.arange(len(off) - 1, dtype=np.uint64), np.diff(off))
    with np.errstate(over="ignore"):
        h = src * np.uint64(0x9E3779B97F4A7C15) + succ.astype(np.uint64)
    return int(np.bitwise_xor.reduce(h)) if len(h) else 0


@pytest.mark.parametrize("q,big", [(8, False), (2, True)])
def test_emulated_walkers_match_the_oracle(tmp_path, oracle, emu_ef, q, big):
    for name, make in CASES:
        off, succ = make()
        base, ub, bits = store(tmp_path, "%s_%d" % (name, q), off, succ, q=q, big=big)
        g = oracle.load_ef(base)
        words = np.ctypeslib.as_array(g.g.contents.words, shape=(int(g.g.contents.nwords) + 2,)).copy()
        offs = g.offsets()
        GUARD = 16
        out = np.full(len(succ) + 2 * GUARD, -7, dtype=np.int32)
        out_off = np.zeros(g.n + 1, dtype=np.int64)
        cs = C.c_ulonglong()
        rc = emu_ef.emu_ef_decode(words.ctypes.data, len(words) - 2, offs.ctypes.data, g.n, ub, q, out_off.ctypes.data, out[GUARD:].ctypes.data, len(succ), C.byref(cs))
        assert rc == 0
        assert np.all(out[:GUARD] == -7) and np.all(out[GUARD + len(succ):] == -7)
        assert np.array_equal(out_off, off) and np.array_equal(out[GUARD:GUARD + len(succ)], succ), name
        assert cs.value == xor_checksum(off, succ)
        g.close()


def test_emulated_element_parallel_writer_is_byte_identical_to_the_host_writer(tmp_path, emu_ef):
    emu_ef.emu_ef_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    emu_ef.emu_ef_compress.restype = C.c_int64
    for q in (8, 0, 3):
        for name, make in CASES:
            off, succ = make()
            if name == "skew":
                off, succ = off[:2000], succ[:off[1999]]   # keep the host loop short: the 60 000-arc node stays in
            base, ub, bits = store(tmp_path, "%s_%d" % (name, q), off, succ, q=q)
            want = np.fromfile(base + ".graph", dtype=np.uint64)
            words = np.zeros(len(want) + 2, dtype=np.uint64)
            node_bits = np.zeros(len(off), dtype=np.int64)
            got = emu_ef.emu_ef_compress(off.ctypes.data, succ.ctypes.data if len(succ) else None, len(off) - 1, ub, q,
                                         words.ctypes.data, len(words), node_bits.ctypes.data)
            assert got == bits, (name, q)
            assert np.array_equal(words[:len(want)], want), (name, q)
            assert np.all(words[len(want):] == 0)
    off = np.array([0, 3], dtype=np.int64)
    words = np.zeros(64, dtype=np.uint64)
    nb = np.zeros(2, dtype=np.int64)
    bad = np.array([0, 0, 1], dtype=np.int32)
    assert emu_ef.emu_ef_compress(off.ctypes.data, bad.ctypes.data, 1, 5, 8, words.ctypes.data, 64, nb.ctypes.data) == -1


# ---------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("q,big", [(8, False), (0, True), (4, False)])
def test_gpu_efgraph_matches_csr_oracle_and_api_protocol(tmp_path, oracle, q, big):
    from webgraph_b200 import bvgraph
    from webgraph_b200.efgraph import EFGraph
    for name, make in CASES:
        off, succ = make()
        base, ub, bits = store(tmp_path, "%s_%d" % (name, q), off, succ, q=q, big=big, threads=2)
        g = EFGraph.load(base)
        n = len(off) - 1
        assert (g.numNodes(), g.numArcs(), g.upperBound, g.quantum, g.graphBits) == (n, len(succ), ub, 1 << q, bits)
        assert g.randomAccess()
        o, s = g.decodeRange(0, n)
        assert np.array_equal(o, off) and np.array_equal(s, succ), name
        assert g.scanRange(0, n) == (len(succ), xor_checksum(off, succ))
        if n > 3:
            for frm, to in ((n // 3, n), (n // 3, 2 * n // 3 + 1), (n // 2, n // 2), (0, 7)):
                o, s = g.decodeRange(frm, to)
                assert np.array_equal(o, off[frm:to + 1] - off[frm]) and np.array_equal(s, succ[off[frm]:off[to]])
                arcs, cs = g.scanRange(frm, to)
                src = np.repeat(np.arange(frm, to, dtype=np.uint64), np.diff(off[frm:to + 1]))
                with np.errstate(over="ignore"):
                    h = src * np.uint64(0x9E3779B97F4A7C15) + succ[off[frm]:off[to]].astype(np.uint64)
                assert (arcs, cs) == (off[to] - off[frm], int(np.bitwise_xor.reduce(h)) if len(h) else 0)
        og = oracle.load_ef(base)
        step = max(1, n // 40)
        for x in range(0, n, step):
            assert g.outdegree(x) == og.outdegree(x) == off[x + 1] - off[x]
            it = g.successors(x)
            want = og.successors(x)
            assert [it.nextInt() for _ in range(len(want))] == want.tolist() and it.nextInt() == -1
        og.close()
        if 0 < n <= 1500:   # the NodeIterator protocol, every start a few nodes apart
            it = g.nodeIterator()
            seen = 0
            while it.hasNext():
                x = it.nextInt()
                assert x == seen and it.outdegree() == off[x + 1] - off[x] and np.array_equal(it.successorArray(), succ[off[x]:off[x + 1]])
                seen += 1
            assert seen == n
            it = g.nodeIterator(n // 2)
            assert it.nextInt() == n // 2 and np.array_equal(it.successorArray(), succ[off[n // 2]:off[n // 2 + 1]])
            with pytest.raises(bvgraph.NoSuchElementError):
                e = g.nodeIterator(n)
                e.nextInt()
        with pytest.raises(ValueError):
            g.outdegree(n)
        with pytest.raises(ValueError):
            g.decodeRange(0, n + 1)
        g.close()


@pytest.mark.gpu
def test_gpu_efgraph_of_cnr2000_equals_the_bvgraph(tmp_path, cnr_truth):
    """The reference's fixture graph stored as an EFGraph decodes to the same lists as its BVGraph, and both scans agree."""
    from webgraph_b200.bvgraph import BVGraph
    from webgraph_b200.efgraph import EFGraph
    off, succ = cnr_truth
    base = str(tmp_path / "cnr-ef")
    tools.store_ef(base, off, succ, threads=4)
    ef, bv = EFGraph.load(base), BVGraph.load(CNR)
    o, s = ef.decodeRange(0, ef.numNodes())
    assert np.array_equal(o, off) and np.array_equal(s, succ)
    assert ef.scanRange(0, ef.numNodes()) == bv.scanRange(0, bv.numNodes()) == (3216152, 0xf941dd3471d172f1)
    assert ef.hashCode() == bv.hashCode()
    ef.close()
    bv.close()


@pytest.mark.gpu
def test_gpu_efgraph_loader_errors(tmp_path):
    from webgraph_b200 import bvgraph
    from webgraph_b200.efgraph import EFGraph
    off, succ = graphs.erdos_renyi(60, .3, 5)
    base, ub, bits = store(tmp_path, "e", off, succ)
    good = open(base + ".properties").read()

    def with_props(text):
        with open(base + ".properties", "w") as f:
            f.write(text)

    with_props(good.replace("it.unimi.dsi.webgraph.EFGraph", "it.unimi.dsi.webgraph.BVGraph"))
    with pytest.raises(IOError):   # "cannot load a graph stored using class", EFGraph.java:716-718
        EFGraph.load(base)
    with_props(good.replace("version=0", "version=1"))
    with pytest.raises(IOError):   # :720-722
        EFGraph.load(base)
    with_props(good.replace("quantum=256", "quantum=100"))
    with pytest.raises(ValueError):   # "Illegal quantum (must be a power of 2)", :731
        EFGraph.load(base)
    with_props(good.replace("LITTLE_ENDIAN", "MIDDLE_ENDIAN"))
    with pytest.raises(ValueError):   # "Unknown byte order", :736
        EFGraph.load(base)
    with_props(good.replace("arcs=%d" % len(succ), "arcs=%d" % (len(succ) + 1)))
    with pytest.raises(bvgraph.FormatError):
        EFGraph.load(base)
    with_props(good)
    data = open(base + ".graph", "rb").read()
    with open(base + ".graph", "wb") as f:
        f.write(data[:len(data) // 2 // 8 * 8])
    with pytest.raises(IOError):   # offsets past the end of the stream
        EFGraph.load(base)
    with pytest.raises(IOError):
        EFGraph.load(str(tmp_path / "nothing"))


@pytest.mark.gpu
def test_gpu_efgraph_corrupted_streams_fail_cleanly(tmp_path):
    """Random byte flips in .graph: every call returns (an error or lists inside the node range is not promised, memory safety
    is), and the device keeps working afterwards."""
    from webgraph_b200 import bvgraph
    from webgraph_b200.efgraph import EFGraph
    off, succ = graphs.copy_heavy(3000, seed=11)[:2]
    base, ub, bits = store(tmp_path, "c", off, succ)
    data = bytearray(open(base + ".graph", "rb").read())
    rng = np.random.default_rng(3)
    outcomes = set()
    for trial in range(12):
        bad = bytearray(data)
        for p in rng.integers(0, len(bad), 1 + trial % 4):
            bad[p] ^= 1 << int(rng.integers(0, 8))
        with open(base + ".graph", "wb") as f:
            f.write(bad)
        try:
            g = EFGraph.load(base)
        except (IOError, ValueError, bvgraph.FormatError) as e:
            outcomes.add(type(e).__name__)
            continue
        try:
            g.decodeRange(0, g.numNodes())
            g.scanRange(0, g.numNodes())
            outcomes.add("decoded")
        except (IOError, bvgraph.FormatError) as e:
            outcomes.add(type(e).__name__)
        g.close()
    assert outcomes   # something happened, nothing crashed
    with open(base + ".graph", "wb") as f:
        f.write(data)
    g = EFGraph.load(base)
    o, s = g.decodeRange(0, g.numNodes())
    assert np.array_equal(o, off) and np.array_equal(s, succ)
    g.close()


@pytest.mark.gpu
def test_gpu_efgraph_store_on_the_device_is_byte_identical(tmp_path, cnr_truth):
    """EFGraph.store with the stream written by the element-parallel kernels: same .graph and .offsets bytes as the host
    writer (which restates Accumulator), and what it wrote loads and decodes."""
    from webgraph_b200.efgraph import EFGraph
    cases = [(n, m) for n, m in CASES if n not in ("empty",)] + [("cnr", lambda: cnr_truth)]
    for q in (8, 2):
        for name, make in cases:
            off, succ = make()
            base, ub, bits = store(tmp_path, "%s_%d" % (name, q), off, succ, q=q, threads=4)
            dbase = base + "-dev"
            dbits, ms = EFGraph.store(dbase, off, succ, upperBound=ub, log2Quantum=q)
            assert dbits == bits, (name, q)
            for ext in (".graph", ".offsets"):
                assert open(base + ext, "rb").read() == open(dbase + ext, "rb").read(), (name, q, ext)
            g = EFGraph.load(dbase)
            o, s = g.decodeRange(0, g.numNodes())
            assert np.array_equal(o, off) and np.array_equal(s, succ)
            g.close()
    with pytest.raises(ValueError):   # not strictly increasing
        EFGraph.store(str(tmp_path / "bad"), np.array([0, 3], dtype=np.int64), np.array([0, 0, 1], dtype=np.int32), upperBound=5)
