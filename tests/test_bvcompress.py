"""BVGraph.store on the device (compress half; bvg_compress.cuh).

The device code is compared byte for byte with the host writer (webgraph_b200/csrc/tools/bvg_tools.cpp, which re-encodes the
reference's cnr-2000 fixture byte-identically: tests/test_writer_roundtrip.py) given the same compression ranges -- the host
writer's `threads` and the device's `range_nodes` cut the node range the same way when range_nodes = ceil(n / threads), as the
reference's own multi-threaded store does (BVGraph.java:2471-2550)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import graphs
from tests.conftest import CNR, ROOT
from webgraph_b200 import tools

EMU_DIR = os.path.join(ROOT, "tests", "hostemu")
EMU_BVC = os.path.join(EMU_DIR, "libemu_bvc.so")


@pytest.fixture(scope="module")
def emu_bvc():
    cuda_dir = os.path.join(ROOT, "webgraph_b200", "csrc", "cuda")
    srcs = [os.path.join(EMU_DIR, "emu_bvc.cpp"), os.path.join(EMU_DIR, "cuda_shim.h"), os.path.join(cuda_dir, "bvg_compress.cuh"), os.path.join(cuda_dir, "bvg_device.cuh")]
    if not os.path.exists(EMU_BVC) or any(os.path.getmtime(s) > os.path.getmtime(EMU_BVC) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-std=c++17", "-fPIC",
                               "-shared", "-I" + EMU_DIR, "-o", EMU_BVC, srcs[0]])
    lib = C.CDLL(EMU_BVC)
    lib.emu_bv_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64,
                                    C.c_void_p, C.c_void_p]
    lib.emu_bv_compress.restype = C.c_int64
    return lib


def skewed():
    n = 20000
    deg = np.zeros(n, dtype=np.int64)
    deg[5] = 15000
    deg[100:140] = 7
    deg[6] = 14990
    deg[n - 1] = 2
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=off[1:])
    lists = [np.arange(0, 15000, dtype=np.int32), np.arange(5, 14995, dtype=np.int32)] + [np.arange(x, x + 21, 3, dtype=np.int32) for x in range(100, 140)] + [np.array([3, n - 1], dtype=np.int32)]
    return off, np.concatenate(lists)


CASES = [("er10", lambda: graphs.erdos_renyi(10, .5, 1)), ("er100", lambda: graphs.erdos_renyi(100, .3, 2)), ("er100d", lambda: graphs.erdos_renyi(100, .9, 3)),
         ("complete", lambda: graphs.complete_graph(40, loops=True)), ("copy", lambda: graphs.copy_heavy(1500, seed=5)[:2]),
         ("copy2", lambda: graphs.copy_heavy(4000, seed=9, maxdeg=120)[:2]), ("skew", skewed),
         ("intree", lambda: graphs.binary_intree(8)), ("arcless", lambda: (np.zeros(51, dtype=np.int64), np.empty(0, dtype=np.int32)))]
PARAMS = [(7, 3, 4, 3), (1, 1, 0, 2), (16, 10, 2, 5), (7, -1, 4, 3), (0, 3, 4, 3), (3, 2, 3, 1)]   # window, maxref, minlen, zetak


def host_reference(tmp_path, name, off, succ, w, r, ml, k, threads):
    base = str(tmp_path / name)
    tools.store_csr(base, off, succ, window=w, maxref=r, minlen=ml, zetak=k, threads=threads)
    return base


@pytest.mark.parametrize("w,r,ml,k", PARAMS)
def test_emulated_device_compressor_is_byte_identical_to_the_host_writer(tmp_path, oracle, emu_bvc, w, r, ml, k):
    for name, make in CASES:
        off, succ = make()
        n = len(off) - 1
        for threads in (1, 3):
            threads = min(threads, max(n, 1))
            rn = (n + threads - 1) // threads
            base = host_reference(tmp_path, "%s_%d" % (name, threads), off, succ, w, r, ml, k, threads)
            want = np.fromfile(base + ".graph", dtype=np.uint8)
            out = np.zeros(len(want) + 64, dtype=np.uint8)
            node_bits = np.zeros(n + 1, dtype=np.int64)
            refs = np.zeros(max(n, 1), dtype=np.int8)
            bits = emu_bvc.emu_bv_compress(off.ctypes.data, succ.ctypes.data if len(succ) else None, n, w, r, ml, k, max(rn, 1),
                                           out.ctypes.data, len(out), node_bits.ctypes.data, refs.ctypes.data)
            assert bits >= 0, (name, threads)
            assert (bits + 7) // 8 == len(want), (name, threads, bits, len(want))
            assert np.array_equal(out[:len(want)], want), (name, threads)
            g = oracle.load(base)
            assert np.array_equal(node_bits.astype(np.uint64), g.offsets()), (name, threads)
            g.close()
    bad = np.array([3, 3, 4], dtype=np.int32)
    off = np.array([0, 3], dtype=np.int64)
    assert emu_bvc.emu_bv_compress(off.ctypes.data, bad.ctypes.data, 1, 7, 3, 4, 3, 1, np.zeros(64, dtype=np.uint8).ctypes.data, 64,
                                   np.zeros(2, dtype=np.int64).ctypes.data, np.zeros(1, dtype=np.int8).ctypes.data) == -1


def test_emulated_device_compressor_on_cnr2000(tmp_path, emu_bvc, cnr_truth):
    """One range = the reference's own single-threaded store: the bytes of the reference's cnr-2000.graph itself."""
    off, succ = cnr_truth
    n = len(off) - 1
    want = np.fromfile(CNR + ".graph", dtype=np.uint8)
    out = np.zeros(len(want) + 64, dtype=np.uint8)
    node_bits = np.zeros(n + 1, dtype=np.int64)
    refs = np.zeros(n, dtype=np.int8)
    bits = emu_bvc.emu_bv_compress(off.ctypes.data, succ.ctypes.data, n, 7, 3, 3, 3, n, out.ctypes.data, len(out), node_bits.ctypes.data, refs.ctypes.data)
    assert (bits + 7) // 8 == len(want)
    assert np.array_equal(out[:len(want)], want)


# ---------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("w,r,ml,k", PARAMS)
def test_gpu_compressor_is_byte_identical_and_round_trips(tmp_path, w, r, ml, k):
    from webgraph_b200.bvgraph import BVGraph
    for name, make in CASES:
        off, succ = make()
        n = len(off) - 1
        for threads in (1, 3):
            threads = min(threads, max(n, 1))
            rn = max(1, (n + threads - 1) // threads)
            base = host_reference(tmp_path, "%s_%d" % (name, threads), off, succ, w, r, ml, k, threads)
            dbase = base + "-dev"
            bits, ms = BVGraph.store(dbase, off, succ, windowSize=w, maxRefCount=r, minIntervalLength=ml, zetaK=k, rangeNodes=rn)
            for ext in (".graph", ".offsets"):
                assert open(base + ext, "rb").read() == open(dbase + ext, "rb").read(), (name, threads, ext)
        # what the device wrote with its default ranges decodes to the graph (on the device)
        dbase = str(tmp_path / ("%s-d256" % name))
        BVGraph.store(dbase, off, succ, windowSize=w, maxRefCount=r, minIntervalLength=ml, zetaK=k, rangeNodes=64)
        g = BVGraph.load(dbase)
        o, s = g.decodeRange(0, n)
        assert np.array_equal(o, off) and np.array_equal(s, succ), name
        g.close()
    with pytest.raises(ValueError):
        BVGraph.store(str(tmp_path / "bad"), np.array([0, 3], dtype=np.int64), np.array([3, 3, 4], dtype=np.int32))


@pytest.mark.gpu
def test_gpu_compressor_reproduces_the_references_cnr2000(tmp_path, cnr_truth):
    """One range = the reference's single-threaded store: the device writes the bytes of the reference's own cnr-2000.graph and
    cnr-2000.offsets."""
    from webgraph_b200.bvgraph import BVGraph
    off, succ = cnr_truth
    base = str(tmp_path / "cnr-dev")
    bits, ms = BVGraph.store(base, off, succ, windowSize=7, maxRefCount=3, minIntervalLength=3, zetaK=3, rangeNodes=len(off) - 1)
    for ext in (".graph", ".offsets"):
        assert open(base + ext, "rb").read() == open(CNR + ext, "rb").read(), ext
    # and with the default ranges it is a few per cent larger, decodes to the same lists
    bits256, _ = BVGraph.store(base + "256", off, succ, windowSize=7, maxRefCount=3, minIntervalLength=3, zetaK=3)
    assert bits <= bits256 < 1.08 * bits
    g = BVGraph.load(base + "256")
    assert g.scanRange(0, g.numNodes()) == (3216152, 0xf941dd3471d172f1)
    g.close()


@pytest.mark.gpu
def test_gpu_compress_calls_report_the_size_they_need(tmp_path):
    """Both compress entry points return BVG_ENOMEM and the size when the output buffer is too small, and succeed when called
    again with that size (what the store() wrappers fall back to when their guess is too small)."""
    from webgraph_b200 import bvgraph
    off, succ = graphs.erdos_renyi(300, .2, 4)
    n = len(off) - 1
    L = bvgraph.lib()
    nb = np.zeros(n + 1, dtype=np.int64)
    need = C.c_uint64(0)
    for call, args in ((L.bvg_bv_compress, (off.ctypes.data, succ.ctypes.data, n, 7, 3, 4, 3, 256, 0, -1)),
                       (L.bvg_ef_compress, (off.ctypes.data, succ.ctypes.data, n, 0, 8, 0, -1))):
        assert call(*args, None, 0, C.byref(need), nb.ctypes.data, None) == bvgraph.BVG_ENOMEM and need.value > 0
        small = np.zeros(need.value - 1, dtype=np.uint8)
        assert call(*args, small.ctypes.data, len(small), C.byref(need), nb.ctypes.data, None) == bvgraph.BVG_ENOMEM
        buf = np.zeros(need.value, dtype=np.uint8)
        assert call(*args, buf.ctypes.data, len(buf), C.byref(need), nb.ctypes.data, None) == bvgraph.BVG_OK
        assert nb[0] == 0 and (nb[-1] + 7) // 8 <= need.value and buf.any()
