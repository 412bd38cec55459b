"""The device-side decode steps (bvg_device.cuh: decode_extras / merge_copied and the k_random chain walk) compiled for
the HOST with ASan+UBSan (tests/hostemu) and compared with the truth.  A debugging net for out-of-bounds row writes and
arithmetic slips that costs no GPU time; the real parity tests are tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import graphs
from tests import oracle_binding as ob
from tests.conftest import CNR, ROOT
from webgraph_b200 import tools

EMU_DIR = os.path.join(ROOT, "tests", "hostemu")
EMU = os.path.join(EMU_DIR, "libemu.so")


def _find_asan():
    out = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    return out if os.path.isabs(out) and os.path.exists(out) else None


@pytest.fixture(scope="module")
def emu():
    cuda_dir = os.path.join(ROOT, "webgraph_b200", "csrc", "cuda")
    srcs = [os.path.join(EMU_DIR, "emu.cpp"), os.path.join(EMU_DIR, "emu_long.cpp"), os.path.join(EMU_DIR, "emu_offsets.cpp"),
            os.path.join(EMU_DIR, "emu_scan.cpp"), os.path.join(EMU_DIR, "emu_boundaries.cpp"), os.path.join(EMU_DIR, "emu_iterators.cpp"), os.path.join(EMU_DIR, "emu_tile.cpp"), os.path.join(EMU_DIR, "emu_stream.cpp"), os.path.join(EMU_DIR, "cuda_shim.h"), os.path.join(cuda_dir, "bvg_device.cuh"), os.path.join(cuda_dir, "bvg_long.cuh"),
            os.path.join(cuda_dir, "bvg_offsets.cuh"), os.path.join(cuda_dir, "bvg_scan.cuh"), os.path.join(cuda_dir, "bvg_boundaries.cuh"),
            os.path.join(cuda_dir, "bvg_tile.cuh"), os.path.join(cuda_dir, "bvg_stream.cuh")]
    if not os.path.exists(EMU) or any(os.path.getmtime(s) > os.path.getmtime(EMU) for s in srcs):
        # UBSan only (ASan needs LD_PRELOAD under python); bounds are enforced by guard words below
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-std=c++17", "-fPIC",
                               "-shared", "-I" + EMU_DIR, "-o", EMU] + srcs[:8])
    lib = C.CDLL(EMU)
    lib.emu_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32] + [C.c_int] * 9 + [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    lib.emu_decode_long.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32] + [C.c_int] * 9 + [C.c_void_p, C.c_void_p, C.c_int64]
    lib.emu_decode_offsets.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]
    lib.emu_boundaries.argtypes = [C.c_void_p, C.c_uint64, C.c_int64] + [C.c_int] * 9 + [C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    lib.emu_masked.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int,
                               C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_ulonglong)]
    lib.emu_scan.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emu_tile_scan.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_uint32,
                                  C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.emu_stream_scan.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emu_stream_fold.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    return lib


def run_emu(emu, oracle, base, random_mode):
    g = oracle.load(base)
    graph = np.fromfile(base + ".graph", dtype=np.uint8)
    graph = np.concatenate([graph, np.zeros(8, dtype=np.uint8)])
    offs = g.offsets()
    c = g.g.contents
    is_def = (c.outdegree_coding, c.block_coding, c.residual_coding, c.reference_coding, c.block_count_coding) == (2, 2, 6, 5, 2)
    toff, tsucc = g.decode_range(0, g.n)
    GUARD = 64
    out = np.full(len(tsucc) + 2 * GUARD, -77, dtype=np.int32)
    out_off = np.zeros(g.n + 1, dtype=np.int64)
    args = (graph.ctypes.data, len(graph) - 8, offs.ctypes.data, g.n, g.window, g.minlen, g.zetak,
            c.outdegree_coding, c.block_coding, c.residual_coding, c.reference_coding, c.block_count_coding,
            1 if is_def else 0, out_off.ctypes.data, out[GUARD:].ctypes.data, len(tsucc))
    rc = emu.emu_decode_long(*args) if random_mode == 2 else emu.emu_decode(*args, random_mode)
    assert rc == 0
    assert np.all(out[:GUARD] == -77) and np.all(out[GUARD + len(tsucc):] == -77), "row write out of bounds"
    assert np.array_equal(out_off, toff)
    assert np.array_equal(out[GUARD:GUARD + len(tsucc)], tsucc)


@pytest.mark.parametrize("random_mode", [0, 1, 2])  # 2 = every record through the split (long-record) path
def test_emulated_kernels_on_cnr2000(emu, oracle, random_mode):
    run_emu(emu, oracle, CNR, random_mode)


@pytest.mark.parametrize("n,p", [(10, .3), (100, .5), (100, .9)])
def test_emulated_kernels_on_erdos_renyi(emu, oracle, tmp_path, n, p):
    off, succ = graphs.erdos_renyi(n, p, seed=n * 7 + int(p * 10))
    base = str(tmp_path / "er")
    tools.store_csr(base, off, succ)
    run_emu(emu, oracle, base, 0)
    run_emu(emu, oracle, base, 1)
    run_emu(emu, oracle, base, 2)


@pytest.mark.parametrize("flags,k,w,r,ml", [(0, 3, 7, 3, 4), (0, 2, 1, 1, 0), (0, 5, 16, 10, 2),
                                           (tools.OUTDEGREES_DELTA | tools.BLOCKS_DELTA | tools.RESIDUALS_DELTA | tools.REFERENCES_DELTA | tools.BLOCK_COUNT_DELTA, 3, 7, -1, 4),
                                           (tools.RESIDUALS_GAMMA | tools.REFERENCES_GAMMA | tools.BLOCK_COUNT_UNARY | tools.BLOCKS_UNARY, 3, 3, 2, 3),
                                           (tools.RESIDUALS_NIBBLE, 3, 7, 3, 4), (tools.RESIDUALS_GOLOMB, 3, 7, 3, 4)])
def test_emulated_kernels_on_copy_heavy(emu, oracle, tmp_path, flags, k, w, r, ml):
    off, succ, _ = graphs.copy_heavy(1200, seed=33 + k)
    base = str(tmp_path / "ch")
    tools.store_csr(base, off, succ, flags=flags, zetak=k, window=w, maxref=r, minlen=ml)
    run_emu(emu, oracle, base, 0)
    run_emu(emu, oracle, base, 1)
    run_emu(emu, oracle, base, 2)


def _emu_offsets(emu, oracle, base):
    g = oracle.load(base)
    stream = np.fromfile(base + ".offsets", dtype=np.uint8)
    out = np.zeros(g.n + 1, dtype=np.uint64)
    passes = C.c_int(0)
    rc = emu.emu_decode_offsets(stream.ctypes.data, len(stream), g.g.contents.offset_coding, g.n, out.ctypes.data, C.byref(passes))
    assert rc == 0
    assert np.array_equal(out, g.offsets())
    return passes.value


def test_emulated_parallel_offsets_decode(emu, oracle, tmp_path):
    assert _emu_offsets(emu, oracle, CNR) >= 1  # 331 195-byte gamma stream, 96-bit sub-ranges
    off, succ, _ = graphs.copy_heavy(3000, seed=8)
    for flags in (0, tools.OFFSETS_DELTA):
        base = str(tmp_path / ("o%d" % flags))
        tools.store_csr(base, off, succ, flags=flags)
        _emu_offsets(emu, oracle, base)
    # records of very different lengths: empty nodes (1-bit gaps) next to a 100 000-successor list (long gamma gaps)
    deg = np.zeros(5000, dtype=np.int64)
    deg[7] = 100000
    deg[4000:4010] = 3
    off2 = np.zeros(5001, dtype=np.int64)
    np.cumsum(deg, out=off2[1:])
    succ2 = np.concatenate([np.arange(0, 400000, 4, dtype=np.int32)] + [np.array([1, 5, 9], dtype=np.int32)] * 10)
    base = str(tmp_path / "skew")
    tools.store_csr(base, off2, succ2)
    _emu_offsets(emu, oracle, base)


def _emu_boundaries(emu, oracle, base, sub_bits, cap=1 << 26):
    g = oracle.load(base)
    graph = np.fromfile(base + ".graph", dtype=np.uint8)
    c = g.g.contents
    is_def = (c.outdegree_coding, c.block_coding, c.residual_coding, c.reference_coding, c.block_count_coding) == (2, 2, 6, 5, 2)
    out = np.full(g.n + 1, 2**64 - 1, dtype=np.uint64)
    passes, walks = C.c_int(0), C.c_int64(0)
    rc = emu.emu_boundaries(graph.ctypes.data, len(graph), g.n, g.window, g.minlen, g.zetak, c.outdegree_coding, c.block_coding,
                            c.residual_coding, c.reference_coding, c.block_count_coding, 1 if is_def else 0, sub_bits, cap,
                            out.ctypes.data, C.byref(passes), C.byref(walks))
    assert rc == 0
    assert np.array_equal(out, g.offsets())
    nsub = max(1, -(-len(graph) * 8 // sub_bits))
    return passes.value, walks.value, nsub


@pytest.mark.parametrize("lean", [0, 1])  # 1: residual runs through the 32-bit window of the scan kernels (BVG_BND_LEAN)
def test_emulated_boundaries_from_graph_alone(emu, oracle, tmp_path, monkeypatch, lean):
    """Record boundaries found without .offsets (bvg_boundaries.cuh) == the .offsets the writer (and, for cnr-2000, the
    reference) produced; sub-ranges from far too small (every speculation wrong, passes do the work) to one."""
    monkeypatch.setenv("EMU_BND_LEAN", str(lean))
    for sub_bits in (1 << 30, 1 << 20, 1 << 17, 1 << 14):
        passes, walks, nsub = _emu_boundaries(emu, oracle, CNR, sub_bits)
        assert passes <= nsub + 1
    # a wrong chain falls onto the right one well inside a 1 Mbit sub-range: speculate, adopt, confirm
    passes, walks, nsub = _emu_boundaries(emu, oracle, CNR, 1 << 20)
    assert nsub == 11 and passes <= 4
    off, succ, _ = graphs.copy_heavy(3000, seed=8)
    for i, (flags, w, r, ml) in enumerate([(0, 7, 3, 4), (0, 0, 3, 0), (0, 1, 1, 2), (0, 16, -1, 3),
                                           (tools.OUTDEGREES_DELTA | tools.BLOCKS_DELTA | tools.RESIDUALS_DELTA | tools.REFERENCES_DELTA | tools.BLOCK_COUNT_DELTA, 7, 3, 4),
                                           (tools.RESIDUALS_GAMMA | tools.REFERENCES_GAMMA | tools.BLOCK_COUNT_UNARY | tools.BLOCKS_UNARY, 3, 2, 3),
                                           (tools.RESIDUALS_NIBBLE, 7, 3, 4)]):
        base = str(tmp_path / ("b%d" % i))
        tools.store_csr(base, off, succ, flags=flags, window=w, maxref=r, minlen=ml)
        for sub_bits in (1 << 30, 1 << 13, 1 << 10):
            _emu_boundaries(emu, oracle, base, sub_bits)
    # a record far longer than a sub-range and than the cap of unproven walks, empty nodes around it
    deg = np.zeros(5000, dtype=np.int64)
    deg[7] = 100000
    deg[4000:4010] = 3
    off2 = np.zeros(5001, dtype=np.int64)
    np.cumsum(deg, out=off2[1:])
    succ2 = np.concatenate([np.arange(0, 400000, 4, dtype=np.int32)] + [np.array([1, 5, 9], dtype=np.int32)] * 10)
    base = str(tmp_path / "skew")
    tools.store_csr(base, off2, succ2)
    for sub_bits, cap in ((1 << 12, 1 << 26), (1 << 12, 1 << 10), (1 << 16, 64)):
        _emu_boundaries(emu, oracle, base, sub_bits, cap)


def test_emulated_fold_only_stream(emu, oracle, tmp_path):
    """stream_only (copied successors folded without a merge) == the same elements pulled through next_a."""
    bases = [CNR]
    off, succ, _ = graphs.copy_heavy(2500, seed=3)
    base = str(tmp_path / "ch")
    tools.store_csr(base, off, succ)
    bases.append(base)
    for b in bases:
        g = oracle.load(b)
        graph = np.concatenate([np.fromfile(b + ".graph", dtype=np.uint8), np.zeros(8, dtype=np.uint8)])
        offs = g.offsets()
        toff, tsucc = g.decode_range(0, g.n)
        a, r = C.c_ulonglong(0), C.c_ulonglong(0)
        rc = emu.emu_stream_fold(graph.ctypes.data, len(graph) - 8, offs.ctypes.data, g.n, g.window, g.minlen, g.zetak, 1,
                                 toff.ctypes.data, tsucc.ctypes.data, C.byref(a), C.byref(r))
        assert rc == 0 and a.value == r.value and a.value != 0


def _emu_scan(emu, oracle, base):
    g = oracle.load(base)
    graph = np.concatenate([np.fromfile(base + ".graph", dtype=np.uint8), np.zeros(8, dtype=np.uint8)])
    offs = g.offsets()
    toff, tsucc = g.decode_range(0, g.n)
    rows = np.full(len(tsucc) + 1, -77, dtype=np.int32)
    out_off = np.zeros(g.n + 1, dtype=np.int64)
    parent = np.zeros(g.n + 1, dtype=np.uint8)
    res = (C.c_ulonglong * 2)()
    rc = emu.emu_scan(graph.ctypes.data, len(graph) - 8, offs.ctypes.data, g.n, g.window, g.minlen, g.zetak,
                      out_off.ctypes.data, rows.ctypes.data, parent.ctypes.data, res)
    assert rc == 0
    assert (res[0], res[1]) == g.scan_range(0, g.n)
    assert rows[len(tsucc)] == -77
    # rows of the nodes somebody copies from are materialised, sorted, exactly as the oracle decodes them
    for x in np.nonzero(parent[:g.n])[0]:
        assert np.array_equal(rows[toff[x]:toff[x + 1]], tsucc[toff[x]:toff[x + 1]]), x
    return int(parent.sum())


def test_emulated_fused_scan(emu, oracle, tmp_path):
    """k_scan_extras_lean / k_scan_merge logic (bvg_scan.cuh) record by record on the host: checksum == oracle's scan,
    parents' rows == oracle's lists."""
    assert _emu_scan(emu, oracle, CNR) > 1000
    for k, w, r, ml in [(3, 7, 3, 4), (3, 7, 3, 0), (2, 1, 1, 2), (5, 16, 10, 3), (1, 3, 2, 4)]:
        off, succ, _ = graphs.copy_heavy(2000, seed=5 + k + ml)
        base = str(tmp_path / ("s%d_%d" % (k, ml)))
        tools.store_csr(base, off, succ, zetak=k, window=w, maxref=r, minlen=ml)
        _emu_scan(emu, oracle, base)
    # ids near 2^31, gaps that need zeta codes longer than the 32-bit window, long gammas for the first interval left
    n = 300
    deg = np.zeros(n, dtype=np.int64)
    deg[5] = 10
    deg[100:110] = 3
    deg[200] = 9
    deg[250] = 2
    deg[251] = 3
    off2 = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=off2[1:])
    succ2 = np.concatenate([np.array([0, 1, 2, 3, 2 ** 30, 2 ** 30 + 1, 2 ** 30 + 2, 2 ** 30 + 3, 2 ** 30 + 4, 2 ** 31 - 2], dtype=np.int64)] +
                           [np.array([7, 2 ** 26 + i, 2 ** 31 - 5 + i // 3], dtype=np.int64) for i in range(10)] +
                           [np.array([2 ** 29 + j for j in range(8)] + [2 ** 31 - 1], dtype=np.int64)] +
                           [np.array([2 ** 28 + 5, 2 ** 28 + 900], dtype=np.int64), np.array([2 ** 30 + 1, 2 ** 30 + 2 ** 25, 2 ** 31 - 3], dtype=np.int64)]).astype(np.int32)
    for k in (3, 4):
        base = str(tmp_path / ("big%d" % k))
        tools.store_csr(base, off2, succ2, zetak=k)
        _emu_scan(emu, oracle, base)


MIX = 0x9E3779B97F4A7C15


def _fold(x, ys):
    acc = 0
    for y in ys:
        acc ^= (x * MIX + int(y)) & (2 ** 64 - 1)
    return acc


def _masked(emu, parent, blocks, extras, copied, x, variant):
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    blocks = np.ascontiguousarray(blocks, dtype=np.int32)
    extras = np.ascontiguousarray(extras, dtype=np.int32)
    out = np.full(len(parent) + len(extras) + 8, -555, dtype=np.int32)
    n, f = C.c_int32(0), C.c_ulonglong(0)
    rc = emu.emu_masked(parent.ctypes.data, len(parent), blocks.ctypes.data, len(blocks), extras.ctypes.data, len(extras), copied, x,
                        variant, out.ctypes.data, C.byref(n), C.byref(f))
    assert rc == 0
    return out[:n.value].tolist(), f.value


def test_masked_iterator_like_the_reference(emu):
    """MaskedIntIteratorTest.test (reference test/it/unimi/dsi/webgraph/MaskedIntIteratorTest.java:32-110): every
    (length, number of dropped elements) up to 19 x 19, the mask given as the alternating block lengths
    MaskedIntIterator takes (MaskedIntIterator.java:65-97: first block copied, possibly empty; an even number of blocks
    keeps the tail), here through the staged copy runs of the scan kernels -- pulled one by one, folded, folded with
    16-byte group reads -- which must agree with the filter computed directly."""
    rng = np.random.default_rng(20)
    for length in range(20):
        for zeros in range(20):
            x = np.cumsum(rng.integers(1, 1000, length)).astype(np.int32)
            keep = np.zeros(length, dtype=bool)
            keep[:max(0, length - zeros)] = True
            rng.shuffle(keep)
            expected = x[keep].tolist()
            blocks, look, curr = [], True, 0
            for i in range(length):
                if keep[i] == look:
                    curr += 1
                else:
                    blocks.append(curr)
                    look = not look
                    curr = 1
            got, _ = _masked(emu, x, blocks, [], 0, 7, 0)
            assert got == expected, (length, zeros, blocks)
            for variant in (1,):
                _, f = _masked(emu, x, blocks, [], 0, 12345, variant)
                assert f == _fold(12345, expected), (length, zeros, blocks, variant)


def test_merged_iterator_like_the_reference(emu):
    """MergedIntIteratorTest.testMerge (reference test/.../MergedIntIteratorTest.java:31-66): the ascending union of two
    ascending lists, an element present in both emitted once (MergedIntIterator.java:50-74) -- here the in-place merge
    of a record's copied successors with its extras (copied_merge); what the union loses to
    duplicates is padded with -1, as BVGraphNodeIterator leaves it (BVGraph.java:1210)."""
    rng = np.random.default_rng(21)
    for i in range(10):
        for n0, n1 in ((i, i), (i, i + 1), (i, i * 2), (i + 1, i), (i * 2, i), (i * 7, i * 5)):
            a = sorted(set(np.cumsum(rng.integers(0, 10, n0)).tolist()))   # the reference builds sets of the two lists
            b = sorted(set(np.cumsum(rng.integers(0, 10, n1)).tolist()))
            union = sorted(set(a) | set(b))
            for variant in (3,):
                got, f = _masked(emu, a, [], b, len(a), 99, variant)  # no blocks: the whole parent is copied
                assert got[:len(union)] == union and all(v == -1 for v in got[len(union):]), (a, b, variant)
                assert len(got) == len(a) + len(b)
                assert f == _fold(99, a), (a, b, variant)  # the copied successors are folded here, the extras where they are decoded
    # and both at once: a mask over the parent, extras in between
    for trial in range(200):
        dp = int(rng.integers(1, 60))
        parent = np.cumsum(rng.integers(1, 9, dp)).astype(np.int32)
        keep = rng.random(dp) < rng.random()
        blocks, look, curr = [], True, 0
        for t in range(dp):
            if keep[t] == look:
                curr += 1
            else:
                blocks.append(curr)
                look = not look
                curr = 1
        copied = parent[keep].tolist()
        pool = sorted(set(range(int(parent[-1]) + 20)) - set(copied))
        extras = sorted(rng.choice(pool, size=min(len(pool), int(rng.integers(0, 30))), replace=False).tolist())
        union = sorted(copied + extras)
        for variant in (3,):
            got, f = _masked(emu, parent, blocks, extras, len(copied), 5, variant)
            assert got == union, (trial, variant)
            assert f == _fold(5, copied)


def _emu_tile(emu, oracle, base, long_d=1024, seg=128, chunk=128, smem=56 * 1024, nt=256, lo=0, hi=None, may_refuse=False):
    g = oracle.load(base)
    hi = g.n if hi is None else hi
    graph = np.fromfile(base + ".graph", dtype=np.uint8)
    nb = len(graph)
    graph = np.concatenate([graph, np.zeros(64, dtype=np.uint8)])
    offs = g.offsets()
    out = np.zeros(2, dtype=np.uint64)
    st = np.zeros(5, dtype=np.int64)
    rc = emu.emu_tile_scan(graph.ctypes.data, nb, offs.ctypes.data, g.n, g.window, g.minlen, g.zetak, long_d, seg, chunk, smem, nt, lo, hi,
                           out.ctypes.data, st.ctypes.data)
    if may_refuse and rc == -300:  # the planner found a node whose ancestors do not fit one tile: the library keeps the general kernels
        return None
    assert rc == 0
    assert (int(out[0]), int(out[1])) == g.scan_range(lo, hi)
    return st.tolist()  # tiles, tiles with a halo, halo nodes, long records, largest tile


def test_emulated_tile_kernel(emu, oracle, tmp_path):
    """The tile kernel (bvg_tile.cuh: planner, staged stream, in-tile header parse, chain levels in shared memory, long records
    split at the sync points) phase by phase on the host, a byte buffer with guard bytes as the block's shared memory."""
    tiles, with_halo, halo_nodes, nlong, biggest = _emu_tile(emu, oracle, CNR)
    assert tiles > 100 and biggest <= 1024
    _emu_tile(emu, oracle, CNR, long_d=64, seg=16, chunk=16)                  # ~2000 records through the split path
    st = _emu_tile(emu, oracle, CNR, long_d=8, seg=3, chunk=4, smem=24 * 1024, nt=64)   # a third of all records long, small tiles
    assert st[3] > 100000 and st[1] > 0                                         # ... and tiles that need a halo
    _emu_tile(emu, oracle, CNR, lo=1000, hi=200000)                            # a sub-range: tiles cut by the range at both ends
    for k, w, r, ml in [(3, 7, 3, 4), (2, 1, 1, 0), (5, 16, -1, 2), (3, 0, 3, 4)]:  # -1: unbounded chains
        off, succ, _ = graphs.copy_heavy(3000, seed=11 + k)
        base = str(tmp_path / ("t%d_%d" % (k, w)))
        tools.store_csr(base, off, succ, zetak=k, window=w, maxref=r, minlen=ml)
        _emu_tile(emu, oracle, base, long_d=16, seg=8, chunk=8, smem=20 * 1024, nt=64, may_refuse=(r < 0))
        if r < 0:
            _emu_tile(emu, oracle, base, long_d=1024, smem=200 * 1024, nt=64, may_refuse=True)


def _emu_stream(emu, oracle, base, lo=0, hi=None):
    g = oracle.load(base)
    hi = g.n if hi is None else hi
    graph = np.fromfile(base + ".graph", dtype=np.uint8)
    nb = len(graph)
    graph = np.concatenate([graph, np.zeros(64, dtype=np.uint8)])
    offs = g.offsets()
    toff, tsucc = g.decode_range(0, g.n)
    GUARD = 64
    rows = np.full(len(tsucc) + 2 * GUARD, -77, dtype=np.int32)
    out_off = np.zeros(g.n + 1, dtype=np.int64)
    pf = np.zeros(g.n + 1, dtype=np.uint8)
    out = np.zeros(2, dtype=np.uint64)
    st = np.zeros(2, dtype=np.int64)
    rc = emu.emu_stream_scan(graph.ctypes.data, nb, offs.ctypes.data, g.n, g.window, g.minlen, g.zetak, lo, hi, out_off.ctypes.data,
                             rows[GUARD:].ctypes.data, pf.ctypes.data, out.ctypes.data, st.ctypes.data)
    assert rc == 0
    assert (int(out[0]), int(out[1])) == g.scan_range(lo, hi)
    assert np.all(rows[:GUARD] == -77) and np.all(rows[GUARD + len(tsucc):] == -77), "row write out of bounds"
    r = rows[GUARD:GUARD + len(tsucc)]
    for x in np.nonzero(pf[:g.n])[0]:
        if lo <= x < hi:
            assert np.array_equal(r[toff[x]:toff[x + 1]], tsucc[toff[x]:toff[x + 1]]), x
    return st.tolist()  # chunks, records whose intervals were merged by the fix-up step


def test_emulated_stream_position_extras(emu, oracle, tmp_path):
    """The stream-position extras kernel (bvg_stream.cuh): an entry per 2048-bit chunk, every lane walked through its chunk,
    deferred interval merges, then the copied parts: checksum == oracle's scan, stored rows == oracle's lists."""
    chunks, deferred = _emu_stream(emu, oracle, CNR)
    assert chunks > 5000 and deferred > 100
    for k, w, r, ml in [(3, 7, 3, 4), (2, 1, 1, 0), (5, 16, 10, 2), (3, 0, 3, 4)]:
        off, succ, _ = graphs.copy_heavy(3000, seed=5)
        base = str(tmp_path / ("q%d_%d" % (k, w)))
        tools.store_csr(base, off, succ, zetak=k, window=w, maxref=r, minlen=ml)
        _emu_stream(emu, oracle, base)
