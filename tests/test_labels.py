"""Arc labels: the label stream of a BitStreamArcLabelledImmutableGraph.

Fixtures are written the way the reference's own test writes them (BitStreamArcLabelledGraphTest.java:131-203: labels
`x * succ + x & mask` per arc, lists of (succ + 1) * 2 elements `x * k + x & mask`), by the host tools; the oracle reads
them one node at a time (BitStreamLabelledArcIterator); the kernels' logic runs on the host (tests/hostemu) in the CPU
suite and on the device in the -m gpu suite.  PARITY UNPINNED: the reference ships no .labels file and there is no JVM
here, so writer, oracle and kernels are three independent restatements checked against each other and against the
closed-form label values."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import graphs
from tests import oracle_binding as ob
from tests.conftest import CNR, ROOT
from webgraph_b200 import tools

LABEL_MASK = (1 << 15) - 1  # BitStreamArcLabelledGraphTest.java:58 uses a 15-bit mask too
EMU_DIR = os.path.join(ROOT, "tests", "hostemu")
EMU_LABELS = os.path.join(EMU_DIR, "libemu_labels.so")


def reference_test_labels(off, succ, kind, width):
    """The label values of the reference's test for a CSR graph: (values, list_off or None)."""
    n = len(off) - 1
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    s = succ.astype(np.int64)
    mask = LABEL_MASK if kind == tools.LABEL_GAMMA else LABEL_MASK & ((1 << width) - 1)
    if kind != tools.LABEL_FIXED_LIST:
        return ((src * s + src) & mask).astype(np.int32), None
    lens = (s + 1) * 2
    lo = np.zeros(len(s) + 1, dtype=np.int64)
    np.cumsum(lens, out=lo[1:])
    arc = np.repeat(np.arange(len(s), dtype=np.int64), lens)
    k = np.arange(lo[-1], dtype=np.int64) - lo[arc]
    return ((src[arc] * k + src[arc]) & mask).astype(np.int32), lo


def write_case(tmp_path, name, off, succ, kind, width, values=None, list_off=None, threads=1):
    base = str(tmp_path / name)
    tools.store_csr(base, off, succ)
    if values is None:
        values, list_off = reference_test_labels(off, succ, kind, width)
    lbase = base + "-lab%d" % kind
    bits = tools.store_labels(lbase, os.path.basename(base), off, values, kind, width, list_off=list_off, threads=threads)
    if list_off is None:
        list_off = np.arange(len(values) + 1, dtype=np.int64)
    return base, lbase, values, list_off, bits


def skewed_graph():
    """Empty nodes next to a 60 000-successor node and a few small ones: the shape that breaks one-thread-per-node."""
    deg = np.zeros(3000, dtype=np.int64)
    deg[5] = 60000
    deg[100:140] = 7
    deg[2999] = 2
    off = np.zeros(3001, dtype=np.int64)
    np.cumsum(deg, out=off[1:])
    lists = [np.arange(0, 3 * d, 3, dtype=np.int32) % 200000 for d in deg]
    succ = np.concatenate([np.sort(l) for l in lists]).astype(np.int32)
    return off, succ


CASES = [("er10", lambda: graphs.erdos_renyi(10, .5, 1)), ("er100", lambda: graphs.erdos_renyi(100, .3, 2)),
         ("copy", lambda: graphs.copy_heavy(1500, seed=5)[:2]), ("skew", skewed_graph),
         ("empty", lambda: (np.zeros(1, dtype=np.int64), np.empty(0, dtype=np.int32))),
         ("arcless", lambda: (np.zeros(51, dtype=np.int64), np.empty(0, dtype=np.int32)))]


def cases_for(kind):
    """List labels carry (succ + 1) * 2 elements per arc (the reference's test): keep the successors small for them."""
    if kind != tools.LABEL_FIXED_LIST:
        return CASES
    return [c for c in CASES if c[0] not in ("copy", "skew")] + [("copy150", lambda: graphs.copy_heavy(150, seed=5)[:2])]


KINDS = [(tools.LABEL_GAMMA, 0), (tools.LABEL_FIXED, 15), (tools.LABEL_FIXED, 7), (tools.LABEL_FIXED, 0), (tools.LABEL_FIXED, 31),
         (tools.LABEL_FIXED_LIST, 9)]


# ---------------------------------------------------------------- CPU: writer <-> oracle, host-emulated kernels

@pytest.mark.parametrize("kind,width", KINDS)
def test_writer_and_oracle_agree_on_the_reference_tests_labels(tmp_path, oracle, kind, width):
    for name, make in cases_for(kind):
        off, succ = make()
        base, lbase, values, list_off, bits = write_case(tmp_path, "%s_%d_%d" % (name, kind, width), off, succ, kind, width)
        n = len(off) - 1
        L = oracle.load_labels(lbase, n)
        assert (L.kind, L.width) == (kind, width if kind != tools.LABEL_GAMMA else 0)
        offs = L.offsets()
        assert offs[0] == 0 and offs[-1] == bits and np.all(np.diff(offs.astype(np.int64)) >= 0)
        lo, vals = L.range(0, n, off)
        assert np.array_equal(lo, list_off) and np.array_equal(vals, values)
        if kind == tools.LABEL_FIXED:
            assert np.array_equal(np.diff(offs.astype(np.int64)), np.diff(off) * width)
        L.close()


def test_writer_is_independent_of_thread_count_and_rejects_bad_values(tmp_path):
    off, succ = graphs.erdos_renyi(300, .1, 9)
    a = write_case(tmp_path, "t1", off, succ, tools.LABEL_GAMMA, 0, threads=1)
    b = write_case(tmp_path, "t5", off, succ, tools.LABEL_GAMMA, 0, threads=5)
    for ext in (".labels", ".labeloffsets"):
        assert open(a[1] + ext, "rb").read() == open(b[1] + ext, "rb").read()
    props = open(a[1] + ".properties").read()
    assert "graphclass = it.unimi.dsi.webgraph.labelling.BitStreamArcLabelledImmutableGraph" in props
    assert "labelspec = it.unimi.dsi.webgraph.labelling.GammaCodedIntLabel(TEST)" in props and "underlyinggraph = t1" in props
    vals = np.zeros(len(succ), dtype=np.int32)
    vals[3] = 1 << 7
    with pytest.raises(ValueError):  # FixedWidthIntLabel.java:42 "Value out of range"
        tools.store_labels(str(tmp_path / "bad"), "t1", off, vals, tools.LABEL_FIXED, 7)
    vals[3] = -1
    with pytest.raises(ValueError):  # GammaCodedIntLabel.java:36 "Value cannot be negative"
        tools.store_labels(str(tmp_path / "bad"), "t1", off, vals, tools.LABEL_GAMMA)
    with pytest.raises(ValueError):  # width out of range, FixedWidthIntLabel.java:41
        tools.store_labels(str(tmp_path / "bad"), "t1", off, np.zeros(len(succ), dtype=np.int32), tools.LABEL_FIXED, 32)


@pytest.fixture(scope="module")
def emu_labels():
    cuda_dir = os.path.join(ROOT, "webgraph_b200", "csrc", "cuda")
    srcs = [os.path.join(EMU_DIR, "emu_labels.cpp"), os.path.join(EMU_DIR, "cuda_shim.h"), os.path.join(cuda_dir, "bvg_labels.cuh"),
            os.path.join(cuda_dir, "bvg_offsets.cuh"), os.path.join(cuda_dir, "bvg_device.cuh")]
    if not os.path.exists(EMU_LABELS) or any(os.path.getmtime(s) > os.path.getmtime(EMU_LABELS) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-std=c++17", "-fPIC",
                               "-shared", "-I" + EMU_DIR, "-o", EMU_LABELS, srcs[0]])
    lib = C.CDLL(EMU_LABELS)
    lib.emu_labels.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_uint64, C.c_int32, C.c_int32,
                               C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int64)]
    return lib


def run_emu_labels(emu, lbase, off, offs, kind, width, frm, to, shard=False):
    stream = np.fromfile(lbase + ".labels", dtype=np.uint8)
    stream = np.concatenate([stream, np.zeros(16, dtype=np.uint8)])
    arcs = int(off[to] - off[frm])
    GUARD = 32
    cap = 1 << 22
    lo = np.full(arcs + 1 + 2 * GUARD, -7, dtype=np.int64)
    vals = np.full(cap + 2 * GUARD, -7, dtype=np.int32)
    cs, nv = C.c_ulonglong(), C.c_int64()
    rc = emu.emu_labels(stream.ctypes.data, len(stream) - 16, offs.ctypes.data, off.ctypes.data, len(off) - 1, kind, width,
                        int(offs[frm]) // 8 if shard else 0, frm, to, lo[GUARD:].ctypes.data, vals[GUARD:].ctypes.data, cap, C.byref(cs), C.byref(nv))
    assert rc == 0, rc
    assert np.all(lo[:GUARD] == -7) and np.all(lo[GUARD + arcs + 1:] == -7) and np.all(vals[:GUARD] == -7) and np.all(vals[GUARD + nv.value:] == -7)
    return lo[GUARD:GUARD + arcs + 1], vals[GUARD:GUARD + nv.value], cs.value


@pytest.mark.parametrize("kind,width", KINDS)
def test_emulated_label_kernels_match_the_oracle(tmp_path, oracle, emu_labels, kind, width):
    for name, make in cases_for(kind):
        off, succ = make()
        base, lbase, values, list_off, _ = write_case(tmp_path, "%s_%d_%d" % (name, kind, width), off, succ, kind, width)
        n = len(off) - 1
        L = oracle.load_labels(lbase, n)
        offs = L.offsets()
        ranges = [(0, n)] + ([(n // 3, n), (n // 3, 2 * n // 3 + 1), (n // 2, n // 2)] if n > 3 else [])
        for frm, to in ranges:
            want_lo, want_vals = L.range(frm, to, off)
            for shard in (False, True):
                lo, vals, cs = run_emu_labels(emu_labels, lbase, off, offs, kind, width, frm, to, shard)
                assert np.array_equal(lo, want_lo), (name, frm, to)
                assert np.array_equal(vals, want_vals), (name, frm, to)
                assert cs == ob.label_checksum(want_lo, want_vals)
        L.close()


def test_emulated_gamma_labels_with_long_codes(tmp_path, oracle, emu_labels):
    """Gamma codes from 1 to 61 bits side by side (values up to 2^31 - 2): the speculative passes must resynchronise over
    codes longer than a 96-bit sub-range leaves room for."""
    off, succ = graphs.erdos_renyi(400, .2, 3)
    rng = np.random.default_rng(4)
    values = (rng.integers(0, 2 ** 31 - 1, len(succ)) >> rng.integers(0, 31, len(succ))).astype(np.int32)
    values[::17] = 2 ** 31 - 2
    base, lbase, values, list_off, _ = write_case(tmp_path, "long", off, succ, tools.LABEL_GAMMA, 0, values=values)
    L = oracle.load_labels(lbase, 400)
    lo, vals = L.range(0, 400, off)
    assert np.array_equal(vals, values)
    got = run_emu_labels(emu_labels, lbase, off, L.offsets(), tools.LABEL_GAMMA, 0, 0, 400)
    assert np.array_equal(got[1], values) and got[2] == ob.label_checksum(lo, vals)
    got = run_emu_labels(emu_labels, lbase, off, L.offsets(), tools.LABEL_GAMMA, 0, 123, 377, shard=True)
    assert np.array_equal(got[1], values[off[123]:off[377]])


# ---------------------------------------------------------------- GPU: the C ABI and the Python mirror

def _reference_style_checks(alg, off, succ, values, list_off, kind):
    """What BitStreamArcLabelledGraphTest.testLabels does (:205-260): sequential iterators, sequential arrays, random access."""
    n = len(off) - 1

    def expect(j):
        return values[list_off[j]:list_off[j + 1]] if kind == tools.LABEL_FIXED_LIST else int(values[j])

    def same(a, b):
        return np.array_equal(a, b) if kind == tools.LABEL_FIXED_LIST else a == b

    it = alg.nodeIterator()
    seen = 0
    while it.hasNext():
        x = it.nextInt()
        arcs = it.successors()
        d = it.outdegree()
        assert d == off[x + 1] - off[x]
        for k in range(d):
            assert arcs.nextInt() == succ[off[x] + k]
            assert same(arcs.label(), expect(off[x] + k))
        assert arcs.nextInt() == -1
        la = it.labelArray()
        assert len(la) == d and all(same(la[k] if kind == tools.LABEL_FIXED_LIST else int(la[k]), expect(off[x] + k)) for k in range(d))
        assert np.array_equal(it.successorArray(), succ[off[x]:off[x + 1]])
        seen += 1
    assert seen == n
    assert alg.randomAccess()
    for x in range(0, n, max(1, n // 50)):
        arcs = alg.successors(x)
        for k in range(alg.outdegree(x)):
            assert arcs.nextInt() == succ[off[x] + k]
            assert same(arcs.label(), expect(off[x] + k))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,width", KINDS)
def test_gpu_labels_match_the_oracle_and_the_reference_tests_protocol(tmp_path, oracle, kind, width):
    from webgraph_b200 import labelling
    for name, make in cases_for(kind):
        off, succ = make()
        base, lbase, values, list_off, bits = write_case(tmp_path, "%s_%d_%d" % (name, kind, width), off, succ, kind, width, threads=3)
        n = len(off) - 1
        alg = labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
        assert (alg.kind, alg.width, alg.labelBits) == (kind, width if kind != tools.LABEL_GAMMA else 0, bits)
        assert alg.numNodes() == n and alg.numArcs() == len(succ)
        L = oracle.load_labels(lbase, n)
        ranges = [(0, n)] + ([(n // 3, n), (n // 3, 2 * n // 3 + 1), (n // 2, n // 2)] if n > 3 else [])
        for frm, to in ranges:
            want_lo, want_vals = L.range(frm, to, off)
            lo, vals = alg.decodeLabels(frm, to)
            assert np.array_equal(lo, want_lo) and np.array_equal(vals, want_vals), (name, frm, to)
            arcs, nv, cs = alg.scanLabels(frm, to)
            assert (arcs, nv, cs) == (off[to] - off[frm], len(want_vals), ob.label_checksum(want_lo, want_vals))
        if n <= 1500:
            _reference_style_checks(alg, off, succ, values, list_off, kind)
        L.close()
        alg.close()


@pytest.mark.gpu
def test_gpu_labels_on_shards_device_buffers_and_long_gamma_codes(tmp_path, oracle):
    import torch
    from webgraph_b200 import bvgraph, labelling
    off, succ = skewed_graph()
    rng = np.random.default_rng(4)
    values = (rng.integers(0, 2 ** 31 - 1, len(succ)) >> rng.integers(0, 31, len(succ))).astype(np.int32)
    values[::17] = 2 ** 31 - 2
    base, lbase, values, list_off, _ = write_case(tmp_path, "sk", off, succ, tools.LABEL_GAMMA, 0, values=values)
    n = len(off) - 1
    # a shard of the underlying graph loads only its stretch of the label stream
    for frm, to in ((0, n), (4, 120), (40, n), (1000, 1000)):
        g = bvgraph.BVGraph.loadShard(base, frm, to)
        h = C.c_void_p()
        bvgraph._check(bvgraph.lib().bvg_labels_open(g.handle, os.fsencode(lbase), C.byref(h)))
        alg = labelling.BitStreamArcLabelledImmutableGraph(g, h)
        lo, vals = alg.decodeLabels(frm, to)
        assert np.array_equal(vals, values[off[frm]:off[to]])
        if (frm, to) == (40, n):  # node 5 holds 99 % of the labels and lies before the shard's halo
            assert alg.heldBytes < 0.2 * (len(values) * 4)
        # device buffers, filled in stream order
        arcs = int(off[to] - off[frm])
        d_vals = torch.full((max(arcs, 1),), -1, dtype=torch.int32, device="cuda")
        d_lo = torch.full((arcs + 1,), -1, dtype=torch.int64, device="cuda")
        nv = C.c_int64()
        bvgraph._check(bvgraph.lib().bvg_labels_decode_range(h, frm, to, d_lo.data_ptr(), d_vals.data_ptr(), arcs, 1, C.byref(nv)))
        torch.cuda.synchronize()
        assert nv.value == arcs and np.array_equal(d_vals.cpu().numpy()[:arcs], values[off[frm]:off[to]])
        assert np.array_equal(d_lo.cpu().numpy(), np.arange(arcs + 1))
        with pytest.raises(MemoryError):
            if arcs:
                bvgraph._check(bvgraph.lib().bvg_labels_decode_range(h, frm, to, None, d_vals.data_ptr(), arcs - 1, 1, None))
            else:
                raise MemoryError
        with pytest.raises(ValueError):  # outside the shard's extent
            alg.decodeLabels(max(0, frm - 1) if frm else 0, to + 1)
        alg.close()
    # the same from caller-owned buffers
    g = bvgraph.BVGraph.load(base)
    alg = labelling.BitStreamArcLabelledImmutableGraph.fromMemory(g, open(lbase + ".labels", "rb").read(), open(lbase + ".labeloffsets", "rb").read(),
                                                                   labelling.GAMMA)
    assert np.array_equal(alg.decodeLabels(0, n)[1], values)
    alg.close()


@pytest.mark.gpu
def test_gpu_labels_loader_errors(tmp_path):
    from webgraph_b200 import bvgraph, labelling
    off, succ = graphs.erdos_renyi(60, .3, 5)
    base, lbase, values, list_off, _ = write_case(tmp_path, "e", off, succ, tools.LABEL_GAMMA, 0)
    good = open(lbase + ".properties").read()

    def with_props(text):
        with open(lbase + ".properties", "w") as f:
            f.write(text)

    with_props(good.replace("labelspec", "labelspek"))
    with pytest.raises(IOError):  # "does not contain a label specification", :409
        labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    with_props(good.replace("underlyinggraph", "underlying"))
    with pytest.raises(IOError):  # "does not contain an underlying graph basename", :391
        labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    with_props(good.replace("GammaCodedIntLabel(TEST)", "SomeOtherLabel(TEST)"))
    with pytest.raises(bvgraph.UnsupportedOperationError):
        labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    with_props(good.replace("GammaCodedIntLabel(TEST)", "FixedWidthIntLabel(TEST,32)"))
    with pytest.raises(ValueError):  # "Width out of range", FixedWidthIntLabel.java:41
        labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    with_props(good.replace("GammaCodedIntLabel(TEST)", "FixedWidthIntLabel(TEST)"))
    with pytest.raises(bvgraph.FormatError):
        labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    # absolute underlying basenames are taken as they are, :393-395
    with_props(good.replace("underlyinggraph = e", "underlyinggraph = " + base))
    alg = labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    assert np.array_equal(alg.decodeLabels(0, 60)[1], values)
    alg.close()
    # label offsets that run past the stream: IOException at load
    with_props(good)
    data = open(lbase + ".labels", "rb").read()
    with open(lbase + ".labels", "wb") as f:
        f.write(data[:len(data) // 2])
    with pytest.raises(IOError):
        labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    # a stretch that does not hold one label per arc (offsets of another labelling): reported, not mis-assigned
    with open(lbase + ".labels", "wb") as f:
        f.write(data)
    tools.store_labels(lbase + "x", "e", off, np.zeros(len(succ), dtype=np.int32), tools.LABEL_FIXED, 3)
    os.replace(lbase + "x.labeloffsets", lbase + ".labeloffsets")
    alg = labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    with pytest.raises(bvgraph.FormatError):
        alg.decodeLabels(0, 60)
    alg.close()
    missing = str(tmp_path / "nothing")
    with pytest.raises(IOError):
        labelling.BitStreamArcLabelledImmutableGraph.load(missing)


@pytest.mark.gpu
def test_gpu_labels_on_cnr2000_all_kinds(tmp_path, oracle, cnr_truth):
    """The reference's fixture graph (325 557 nodes, 3.2 M arcs) labelled the way the reference's test labels its graphs."""
    from webgraph_b200 import bvgraph, labelling
    off, succ = cnr_truth
    for kind, width in ((tools.LABEL_GAMMA, 0), (tools.LABEL_FIXED, 13)):
        values, _ = reference_test_labels(off, succ, kind, width)
        lbase = str(tmp_path / ("cnr-lab%d" % kind))
        tools.store_labels(lbase, CNR, off, values, kind, width, threads=4)
        alg = labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
        lo, vals = alg.decodeLabels(0, alg.numNodes())
        assert np.array_equal(vals, values)
        arcs, nv, cs = alg.scanLabels(0, alg.numNodes())
        assert (arcs, nv, cs) == (len(succ), len(succ), ob.label_checksum(lo, vals))
        L = oracle.load_labels(lbase, alg.numNodes())
        for x in (0, 1, 1000, 100000, 325556):
            assert np.array_equal(alg.labelArray(x), L.node(x, int(off[x + 1] - off[x]))[1])
        L.close()
        alg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,width", [(tools.LABEL_GAMMA, 0), (tools.LABEL_FIXED, 11), (tools.LABEL_FIXED_LIST, 5)])
def test_gpu_labels_corrupted_streams_fail_cleanly(tmp_path, kind, width):
    """Random byte flips in .labels / .labeloffsets: calls return (an error, or labels that differ), nothing crashes, and the
    device keeps working afterwards."""
    from webgraph_b200 import bvgraph, labelling
    off, succ = graphs.erdos_renyi(400, .1, 21) if kind != tools.LABEL_FIXED_LIST else graphs.erdos_renyi(60, .2, 21)
    base, lbase, values, list_off, _ = write_case(tmp_path, "c", off, succ, kind, width)
    n = len(off) - 1
    rng = np.random.default_rng(5)
    outcomes = set()
    for ext in (".labels", ".labeloffsets"):
        data = bytearray(open(lbase + ext, "rb").read())
        for trial in range(8):
            bad = bytearray(data)
            for p in rng.integers(0, len(bad), 1 + trial % 3):
                bad[p] ^= 1 << int(rng.integers(0, 8))
            with open(lbase + ext, "wb") as f:
                f.write(bad)
            try:
                alg = labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
            except (IOError, ValueError, bvgraph.FormatError) as e:
                outcomes.add(type(e).__name__)
                continue
            try:
                alg.decodeLabels(0, n)
                alg.scanLabels(0, n)
                outcomes.add("decoded")
            except (IOError, MemoryError, bvgraph.FormatError) as e:
                outcomes.add(type(e).__name__)
            alg.close()
        with open(lbase + ext, "wb") as f:
            f.write(data)
    assert outcomes
    alg = labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    lo, vals = alg.decodeLabels(0, n)
    assert np.array_equal(lo, list_off) and np.array_equal(vals, values)
    alg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,width", [(tools.LABEL_FIXED, 13), (tools.LABEL_FIXED_LIST, 6), (tools.LABEL_GAMMA, 0)])
def test_gpu_labels_of_every_kind_on_shards(tmp_path, oracle, kind, width):
    """Labels opened on a shard of the underlying graph (node window with a halo, bit base inside the stream): every kind."""
    from webgraph_b200 import bvgraph, labelling
    off, succ = graphs.copy_heavy(900, seed=13, maxdeg=30, universe=120)[:2] if kind == tools.LABEL_FIXED_LIST else graphs.copy_heavy(5000, seed=13)[:2]
    base, lbase, values, list_off, _ = write_case(tmp_path, "s", off, succ, kind, width, threads=2)
    n = len(off) - 1
    L = oracle.load_labels(lbase, n)
    for frm, to in ((0, n // 3), (n // 3, 2 * n // 3), (2 * n // 3, n), (n - 1, n)):
        g = bvgraph.BVGraph.loadShard(base, frm, to)
        h = C.c_void_p()
        bvgraph._check(bvgraph.lib().bvg_labels_open(g.handle, os.fsencode(lbase), C.byref(h)))
        alg = labelling.BitStreamArcLabelledImmutableGraph(g, h)
        want_lo, want_vals = L.range(frm, to, off)
        lo, vals = alg.decodeLabels(frm, to)
        assert np.array_equal(lo, want_lo) and np.array_equal(vals, want_vals), (frm, to)
        mid = (frm + to) // 2
        lo, vals = alg.decodeLabels(mid, to)
        w2 = L.range(mid, to, off)
        assert np.array_equal(lo, w2[0]) and np.array_equal(vals, w2[1])
        assert alg.scanLabels(frm, to)[2] == ob.label_checksum(want_lo, want_vals)
        alg.close()
    L.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,width", [(tools.LABEL_FIXED, 9), (tools.LABEL_FIXED_LIST, 4)])
def test_gpu_labels_into_device_buffers(tmp_path, oracle, kind, width):
    """bvg_labels_decode_range with on_device = 1 for the fixed-width kinds (the gamma kind is covered with the shards): list
    offsets and values land in caller-owned device memory, sizes come back, too small a buffer is BVG_ENOMEM."""
    import torch
    from webgraph_b200 import bvgraph, labelling
    off, succ = graphs.erdos_renyi(70, .2, 8) if kind == tools.LABEL_FIXED_LIST else graphs.copy_heavy(2000, seed=3)[:2]
    base, lbase, values, list_off, _ = write_case(tmp_path, "d", off, succ, kind, width)
    n = len(off) - 1
    alg = labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
    L = bvgraph.lib()
    for frm, to in ((0, n), (n // 4, 3 * n // 4)):
        arcs = int(off[to] - off[frm])
        nv = C.c_int64()
        bvgraph._check(L.bvg_labels_decode_range(alg._h, frm, to, None, None, 0, 1, C.byref(nv)))
        want_lo = list_off[off[frm]:off[to] + 1] - list_off[off[frm]]
        want_vals = values[list_off[off[frm]]:list_off[off[to]]]
        assert nv.value == len(want_vals)
        d_lo = torch.full((arcs + 1,), -1, dtype=torch.int64, device="cuda")
        d_vals = torch.full((max(nv.value, 1),), -1, dtype=torch.int32, device="cuda")
        bvgraph._check(L.bvg_labels_decode_range(alg._h, frm, to, d_lo.data_ptr(), d_vals.data_ptr(), nv.value, 1, C.byref(nv)))
        torch.cuda.synchronize()
        assert np.array_equal(d_lo.cpu().numpy(), want_lo) and np.array_equal(d_vals.cpu().numpy()[:nv.value], want_vals)
        if nv.value:
            assert L.bvg_labels_decode_range(alg._h, frm, to, d_lo.data_ptr(), d_vals.data_ptr(), nv.value - 1, 1, None) == bvgraph.BVG_ENOMEM
    alg.close()
