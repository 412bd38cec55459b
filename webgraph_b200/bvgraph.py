"""Host-side mirror of the reference's graph API over the C ABI (include/bvgraph_b200.h).

Same names, argument meaning and error behaviour as the reference's Java classes so that parity tests read like
the reference's own tests (WebGraphTestCase.assertGraph, BVGraphTest.testLarge):

  ImmutableGraph  (reference src/it/unimi/dsi/webgraph/ImmutableGraph.java:169-772)
  BVGraph         (BVGraph.java: load :1380-1500, numNodes/numArcs, outdegree :857, successors :896,
                   successorArray ImmutableGraph.java:329, nodeIterator :1292, copy :551, randomAccess :592)
  NodeIterator    (NodeIterator.java:34-107; BVGraphNodeIterator BVGraph.java:1136-1281)
  LazyIntIterator (LazyIntIterator.java:28-44: nextInt() -> next successor or -1 forever, skip(n))

Exceptions map 1:1 from bvg_status: IllegalArgumentException -> ValueError, IllegalStateException ->
IllegalStateError, UnsupportedOperationException -> UnsupportedOperationError, NoSuchElementException ->
StopIteration/NoSuchElementError, IOException -> IOError.

All decoding happens on the GPU inside libbvgraph_b200.so; this file is ctypes glue only.  Importing it without the
built library, or calling it without a CUDA device, fails loudly (there is no CPU fallback).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

BVG_OK, BVG_EINVAL, BVG_ESTATE, BVG_EUNSUPPORTED, BVG_EIO, BVG_EFORMAT, BVG_ENOMEM, BVG_ECUDA, BVG_EEND = 0, -1, -2, -3, -4, -5, -6, -7, -8


class IllegalStateError(RuntimeError):
    pass


class UnsupportedOperationError(RuntimeError):
    pass


class NoSuchElementError(IndexError):
    pass


class CudaError(RuntimeError):
    pass


class FormatError(IOError):
    pass


_lib = None

SYMBOLS = [
    "bvg_open", "bvg_open_shard", "bvg_open_memory", "bvg_close", "bvg_info", "bvg_extent", "bvg_random_access",
    "bvg_set_stream", "bvg_device", "bvg_outdegree", "bvg_successors", "bvg_successors_batch", "bvg_outdegree_batch",
    "bvg_range_arcs", "bvg_decode_range", "bvg_scan_range", "bvg_scan_range_async", "bvg_cursor_open", "bvg_cursor_next",
    "bvg_cursor_copy", "bvg_cursor_close", "bvg_cursor_drain", "bvg_boundary_count", "bvg_boundary_export", "bvg_halo_needed",
    "bvg_halo_import", "bvg_strerror", "bvg_last_error_node", "bvg_kernel_launches", "bvg_memory_footprint",
    "bvg_open_memory_shard", "bvg_plan_shards", "bvg_scan_memory", "bvg_release_cached_memory", "bvg_profile", "bvg_profile_read",
    "bvg_scan_bits", "bvg_replan_shards", "bvg_indegrees", "bvg_bfs", "bvg_cursor_next_batch",
    "bvg_labels_underlying", "bvg_labels_open", "bvg_labels_open_memory", "bvg_labels_close", "bvg_labels_info",
    "bvg_labels_decode_range", "bvg_labels_scan_range", "bvg_hyperball_step",
    "bvg_ef_open", "bvg_ef_open_memory", "bvg_ef_close", "bvg_ef_info", "bvg_ef_outdegree", "bvg_ef_successors", "bvg_ef_range_arcs",
    "bvg_ef_decode_range", "bvg_ef_scan_range", "bvg_ef_last_error_node", "bvg_ef_compress", "bvg_bv_compress",
]


def lib():
    """Loads libbvgraph_b200.so (building it in-tree with nvcc if stale) and declares the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(_build.cuda_library())
    P, vp, i32, i64, u32, u64 = C.POINTER, C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
    L.bvg_open.argtypes = [C.c_char_p, C.c_int, P(C.c_int), C.c_int, P(vp)]
    L.bvg_open_shard.argtypes = [C.c_char_p, C.c_int, i32, i32, P(vp)]
    L.bvg_open_memory.argtypes = [vp, u64, vp, u64, i32, i64, i32, i32, i32, i32, u32, C.c_int, C.c_int, P(vp)]
    L.bvg_open_memory_shard.argtypes = [vp, u64, vp, u64, i32, i64, i32, i32, i32, i32, u32, C.c_int, C.c_int, i32, i32, P(vp)]
    L.bvg_plan_shards.argtypes = [C.c_char_p, C.c_int, P(i32)]
    L.bvg_replan_shards.argtypes = [C.c_char_p, C.c_int, P(i32), P(C.c_double), P(i32)]
    L.bvg_release_cached_memory.argtypes = [C.c_int]
    L.bvg_release_cached_memory.restype = C.c_int64
    L.bvg_scan_memory.argtypes = [vp, u64, vp, u64, i32, i64, i32, i32, i32, i32, u32, C.c_int, i32, i32, C.c_int, P(i64), P(u64)]
    L.bvg_profile.argtypes = [vp, C.c_int]
    L.bvg_profile_read.argtypes = [vp, C.c_char_p, C.c_int]
    L.bvg_close.argtypes = [vp]
    L.bvg_close.restype = None
    L.bvg_info.argtypes = [vp, P(i32), P(i64), P(i32), P(i32), P(i32), P(i32), P(u32), P(i64)]
    L.bvg_extent.argtypes = [vp, P(i32), P(i32), P(i32), P(i32)]
    L.bvg_random_access.argtypes = [vp]
    L.bvg_set_stream.argtypes = [vp, vp]
    L.bvg_device.argtypes = [vp]
    L.bvg_outdegree.argtypes = [vp, i32, P(i32)]
    L.bvg_successors.argtypes = [vp, i32, vp, i32, P(i32)]
    L.bvg_successors_batch.argtypes = [vp, vp, i64, vp, vp, i64, C.c_int]
    L.bvg_outdegree_batch.argtypes = [vp, vp, i32, i64, vp, C.c_int]
    L.bvg_range_arcs.argtypes = [vp, i32, i32, P(i64)]
    L.bvg_decode_range.argtypes = [vp, i32, i32, vp, vp, i64, C.c_int]
    L.bvg_scan_range.argtypes = [vp, i32, i32, P(i64), P(u64)]
    L.bvg_scan_range_async.argtypes = [vp, i32, i32, vp]
    L.bvg_cursor_open.argtypes = [vp, i32, i32, P(vp)]
    L.bvg_cursor_next.argtypes = [vp, P(i32), P(i32), P(P(i32))]
    L.bvg_cursor_next_batch.argtypes = [vp, P(i32), P(i32), P(P(i64)), P(P(i32))]
    L.bvg_cursor_copy.argtypes = [vp, i32, P(vp)]
    L.bvg_cursor_close.argtypes = [vp]
    L.bvg_cursor_close.restype = None
    L.bvg_cursor_drain.argtypes = [vp, C.c_int64, P(C.c_int64), P(C.c_int64), P(C.c_uint64)]
    L.bvg_boundary_count.argtypes = [vp, P(i32)]
    L.bvg_boundary_export.argtypes = [vp, vp, vp, i64, C.c_int]
    L.bvg_halo_needed.argtypes = [vp, P(i32)]
    L.bvg_halo_import.argtypes = [vp, i32, vp, vp, C.c_int]
    L.bvg_strerror.argtypes = [C.c_int]
    L.bvg_strerror.restype = C.c_char_p
    L.bvg_last_error_node.argtypes = [vp, P(i32), P(i64)]
    L.bvg_kernel_launches.argtypes = []
    L.bvg_kernel_launches.restype = i64
    L.bvg_memory_footprint.argtypes = [vp, P(i64), P(i64), P(i64)]
    L.bvg_scan_bits.argtypes = [vp, P(i64)]
    L.bvg_indegrees.argtypes = [vp, i32, i32, vp, i64, C.c_int, P(i64)]
    L.bvg_bfs.argtypes = [vp, i32, vp, C.c_int, P(i32), P(i64)]
    L.bvg_ef_open.argtypes = [C.c_char_p, C.c_int, P(vp)]
    L.bvg_ef_open_memory.argtypes = [vp, u64, vp, u64, i32, i64, i32, i32, C.c_int, C.c_int, P(vp)]
    L.bvg_ef_close.argtypes = [vp]
    L.bvg_ef_close.restype = None
    L.bvg_ef_info.argtypes = [vp, P(i32), P(i64), P(i32), P(i32), P(i64)]
    L.bvg_ef_outdegree.argtypes = [vp, i32, P(i32)]
    L.bvg_ef_successors.argtypes = [vp, i32, vp, i32, P(i32)]
    L.bvg_ef_range_arcs.argtypes = [vp, i32, i32, P(i64)]
    L.bvg_ef_decode_range.argtypes = [vp, i32, i32, vp, vp, i64, C.c_int]
    L.bvg_ef_scan_range.argtypes = [vp, i32, i32, P(i64), P(u64)]
    L.bvg_ef_last_error_node.argtypes = [vp, P(i32), P(i64)]
    L.bvg_bv_compress.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, C.c_int, C.c_int, vp, u64, P(u64), vp, P(C.c_double)]
    L.bvg_ef_compress.argtypes = [vp, vp, i32, i32, C.c_int, C.c_int, C.c_int, vp, u64, P(u64), vp, P(C.c_double)]
    L.bvg_hyperball_step.argtypes = [vp, i32, i32, C.c_int, vp, vp, C.c_int, P(i64)]
    L.bvg_labels_underlying.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.bvg_labels_open.argtypes = [vp, C.c_char_p, P(vp)]
    L.bvg_labels_open_memory.argtypes = [vp, vp, u64, vp, u64, C.c_int, C.c_int, P(vp)]
    L.bvg_labels_close.argtypes = [vp]
    L.bvg_labels_close.restype = None
    L.bvg_labels_info.argtypes = [vp, P(C.c_int), P(C.c_int), P(i64), P(i64)]
    L.bvg_labels_decode_range.argtypes = [vp, i32, i32, vp, vp, i64, C.c_int, P(i64)]
    L.bvg_labels_scan_range.argtypes = [vp, i32, i32, P(i64), P(i64), P(u64)]
    _lib = L
    return L


def _check(rc, g=None):
    if rc == BVG_OK:
        return
    msg = lib().bvg_strerror(rc).decode()
    if g is not None and rc in (BVG_ESTATE, BVG_EIO, BVG_EFORMAT):
        node, pos = C.c_int32(-1), C.c_int64(-1)
        lib().bvg_last_error_node(g, C.byref(node), C.byref(pos))
        if node.value >= 0:
            msg += " [node %d, stream position %d]" % (node.value, pos.value)
    if rc == BVG_EINVAL:
        raise ValueError(msg)                    # IllegalArgumentException
    if rc == BVG_ESTATE:
        raise IllegalStateError(msg)             # IllegalStateException
    if rc == BVG_EUNSUPPORTED:
        raise UnsupportedOperationError(msg)     # UnsupportedOperationException
    if rc == BVG_EIO:
        raise IOError(msg)                       # RuntimeException(IOException)
    if rc == BVG_EFORMAT:
        raise FormatError(msg)
    if rc == BVG_ENOMEM:
        raise MemoryError(msg)
    if rc == BVG_ECUDA:
        raise CudaError(msg)
    if rc == BVG_EEND:
        raise NoSuchElementError(msg)            # NoSuchElementException
    raise RuntimeError("bvg status %d: %s" % (rc, msg))


def immutable_graph_hash(off, succ, first_node=0):
    """ImmutableGraph.hashCode (ImmutableGraph.java:755-769) of a graph given as CSR: h = -1, then per node h = 31 h + x
    followed by its successors in REVERSE order, in 32-bit wrap-around arithmetic; returned as a Java int."""
    off = np.asarray(off, dtype=np.int64)
    n, m = len(off) - 1, int(off[-1] - off[0])
    seq = np.empty(n + m, dtype=np.uint32)
    d = np.diff(off)
    node_pos = np.arange(n, dtype=np.int64) + (off[:-1] - off[0])  # where node x sits in the sequence
    seq[node_pos] = (np.arange(n, dtype=np.int64) + first_node).astype(np.uint32)
    if m:
        # successor j of node x (0-based) goes to node_pos[x] + d[x] - j
        owner = np.repeat(np.arange(n, dtype=np.int64), d)
        j = np.arange(m, dtype=np.int64) - np.repeat(off[:-1] - off[0], d)
        seq[node_pos[owner] + d[owner] - j] = np.asarray(succ[:m] if off[0] == 0 else succ[off[0]:off[-1]], dtype=np.int64).astype(np.uint32)
    # h = (-1) 31^L + sum seq[i] 31^(L-1-i)  (mod 2^32)
    L = n + m
    pw = np.ones(L + 1, dtype=np.uint32)
    if L:
        pw[1:] = 31
        pw = np.cumprod(pw, dtype=np.uint32)  # pw[k] = 31^k mod 2^32
    h = (np.uint64(0xFFFFFFFF) * np.uint64(pw[L])) & np.uint64(0xFFFFFFFF)
    if L:
        h = (h + np.uint64(np.sum(seq.astype(np.uint64) * pw[L - 1::-1][:L].astype(np.uint64) & np.uint64(0xFFFFFFFF), dtype=np.uint64))) & np.uint64(0xFFFFFFFF)
    h = int(h)
    return h - (1 << 32) if h >= (1 << 31) else h


class LazyIntIterator:
    """LazyIntIterator over a decoded successor array (LazyIntIterators.wrap, LazyIntIterators.java:151-154)."""

    def __init__(self, arr):
        self._a = arr
        self._i = 0

    def nextInt(self):
        if self._i >= len(self._a):
            return -1  # keeps returning -1 after exhaustion (LazyIntIterator.java:35)
        v = int(self._a[self._i])
        self._i += 1
        return v

    def skip(self, n):
        k = max(0, min(n, len(self._a) - self._i))
        self._i += k
        return k

    def __iter__(self):
        return self

    def __next__(self):
        v = self.nextInt()
        if v == -1:
            raise StopIteration
        return v


class NodeIterator:
    """BVGraphNodeIterator (BVGraph.java:1136-1281) over a bvg_cursor: the device decodes batches of nodes."""

    def __init__(self, graph, handle, frm):
        self._g = graph
        self._h = handle
        self._from = frm
        self._curr = frm - 1
        self._d = 0
        self._succ = None

    def hasNext(self):
        return self._curr < min(self._upper(), self._g.numNodes()) - 1

    def _upper(self):
        return self._up

    def nextInt(self):
        node, d, ptr = C.c_int32(), C.c_int32(), C.POINTER(C.c_int32)()
        _check(lib().bvg_cursor_next(self._h, C.byref(node), C.byref(d), C.byref(ptr)), self._g._h)
        self._curr, self._d = node.value, d.value
        self._succ = np.ctypeslib.as_array(ptr, shape=(d.value,)) if d.value else np.empty(0, dtype=np.int32)
        return node.value

    def outdegree(self):
        if self._curr == self._from - 1:
            raise IllegalStateError("nextInt() has never been called")  # BVGraph.java:1237
        return self._d

    def successorArray(self, copy=True):
        """A fresh array by default.  copy=False returns a view of the cursor's pinned batch (the aliasing the reference allows,
        BVGraph.java:1228-1233): it is overwritten by a later refill and dangles once the iterator is closed."""
        if self._curr == self._from - 1:
            raise IllegalStateError("nextInt() has never been called")  # :1230
        return self._succ.copy() if copy else self._succ

    def nextBatch(self):
        """(first node, offsets, successors) of the whole batch holding the next node, zero-copy views of the cursor's pinned
        memory (bvg_cursor_next_batch): successors of node first + i are succ[off[i]:off[i + 1]].  None at the end.  Valid until
        the next call on this iterator."""
        first, count = C.c_int32(), C.c_int32()
        off, succ = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
        rc = lib().bvg_cursor_next_batch(self._h, C.byref(first), C.byref(count), C.byref(off), C.byref(succ))
        if rc == -8:  # BVG_EEND
            return None
        _check(rc, self._g._h)
        o = np.ctypeslib.as_array(off, shape=(count.value + 1,))
        sarr = np.ctypeslib.as_array(succ, shape=(max(int(o[-1]), 1),))
        self._curr = first.value + count.value - 1
        return first.value, o, sarr

    def successors(self):
        if self._curr == self._from - 1:
            raise IllegalStateError("nextInt() has never been called")  # :1222
        return LazyIntIterator(self._succ)

    def copy(self, upperBound=2**31 - 1):
        h = C.c_void_p()
        _check(lib().bvg_cursor_copy(self._h, upperBound, C.byref(h)))
        it = NodeIterator(self._g, h, self._curr + 1)
        it._up = min(upperBound, self._g.numNodes())
        return it

    def close(self):
        if self._h:
            lib().bvg_cursor_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __iter__(self):
        return self

    def __next__(self):
        if not self.hasNext():
            raise StopIteration
        return self.nextInt()


class ImmutableGraph:
    """The abstract surface (ImmutableGraph.java:254-420); BVGraph below is the only implementation here."""

    @staticmethod
    def load(basename):
        return BVGraph.load(basename)

    @staticmethod
    def loadMapped(basename):
        return BVGraph.loadMapped(basename)

    @staticmethod
    def loadOffline(basename):
        return BVGraph.loadOffline(basename)

    def successors(self, x):
        return LazyIntIterator(self.successorArray(x))

    def nodeIterator(self, frm=0):
        raise NotImplementedError

    def splitNodeIterators(self, howMany):
        """ImmutableGraph.splitNodeIterators (:379-409): ceil(n/howMany)-node ranges; trailing entries None."""
        n = self.numNodes()
        if n == 0 and howMany == 0:
            return []
        if howMany == 0:
            raise ValueError("howMany == 0")
        per = (n + howMany - 1) // howMany
        out = []
        for i in range(howMany):
            frm = i * per
            if frm >= n and not (n == 0 and i == 0):
                out.append(None)
                continue
            it = self.nodeIterator(frm).copy(min(n, frm + per)) if frm else self.nodeIterator(0).copy(min(n, per))
            out.append(it)
        return out

    def hashCode(self):
        """ImmutableGraph.hashCode (:755-769), computed from one bulk decode of the whole graph."""
        off, succ = self.decodeRange(0, self.numNodes())
        return immutable_graph_hash(off, succ)

    def equals(self, other):
        """ImmutableGraph.equals (:731-749): same node count and same successor lists through sequential iterators."""
        if self.numNodes() != other.numNodes():
            return False
        a, b = self.nodeIterator(), other.nodeIterator()
        while a.hasNext():
            a.nextInt()
            b.nextInt()
            if a.outdegree() != b.outdegree() or not np.array_equal(a.successorArray(), b.successorArray()):
                return False
        return True


class BVGraph(ImmutableGraph):
    def __init__(self, handle, basename=None):
        self._h = handle
        self._basename = basename
        n, m, w, r, ml, k = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        fl, gb = C.c_uint32(), C.c_int64()
        _check(lib().bvg_info(handle, C.byref(n), C.byref(m), C.byref(w), C.byref(r), C.byref(ml), C.byref(k), C.byref(fl), C.byref(gb)))
        self._n, self._m = n.value, m.value
        self._window, self._maxref, self._minlen, self._zetak, self._flags, self._graph_bits = w.value, r.value, ml.value, k.value, fl.value, gb.value

    # ---- loading (BVGraph.load / loadMapped / loadOffline, BVGraph.java:1380-1500) ----
    @classmethod
    def _open(cls, basename, offset_type, device=None):
        h = C.c_void_p()
        devs = (C.c_int * 1)(device) if device is not None else None
        _check(lib().bvg_open(os.fsencode(basename), offset_type, devs, 1 if device is not None else 0, C.byref(h)))
        return cls(h, str(basename))

    @classmethod
    def load(cls, basename, offsetType=1, device=None):
        return cls._open(basename, offsetType, device)

    @classmethod
    def loadMapped(cls, basename, device=None):
        return cls._open(basename, 2, device)

    @classmethod
    def loadSequential(cls, basename, device=None):
        return cls._open(basename, 0, device)

    @classmethod
    def loadOffline(cls, basename, device=None):
        return cls._open(basename, -1, device)

    @classmethod
    def loadShard(cls, basename, frm, to, device=-1):
        h = C.c_void_p()
        _check(lib().bvg_open_shard(os.fsencode(basename), device, frm, to, C.byref(h)))
        return cls(h, str(basename))

    @classmethod
    def fromMemory(cls, graph, offsets_stream, nodes, arcs, window, maxref, minlen, zetak=3, flags=0, offsetType=1, device=-1):
        g = np.frombuffer(graph, dtype=np.uint8)
        o = np.frombuffer(offsets_stream, dtype=np.uint8) if offsets_stream is not None else None  # None: sequential graph without .offsets
        h = C.c_void_p()
        _check(lib().bvg_open_memory(g.ctypes.data if len(g) else None, len(g), o.ctypes.data if o is not None else None,
                                     len(o) if o is not None else 0, nodes, arcs, window, maxref,
                                     minlen, zetak, flags, offsetType, device, C.byref(h)))
        return cls(h)

    @staticmethod
    def store(basename, off, succ, windowSize=7, maxRefCount=3, minIntervalLength=4, zetaK=3, rangeNodes=256, device=-1):
        """BVGraph.store (BVGraph.java:1679-1688, 2436-2650; defaults :454-472) with the stream produced on the device
        (bvg_bv_compress, default codings): writes <basename>.graph, .offsets (gamma-coded gaps, from the node bit positions the
        device returns) and .properties.  rangeNodes: nodes per independently compressed range (the whole graph = the
        single-threaded reference's bytes).  Returns (graph bits, milliseconds of the device kernels)."""
        from . import tools
        off = np.ascontiguousarray(off, dtype=np.int64)
        succ = np.ascontiguousarray(succ, dtype=np.int32)
        n = len(off) - 1
        need, ms = C.c_uint64(0), C.c_double(0)
        node_bits = np.zeros(n + 1, dtype=np.int64)
        L = lib()
        args = (off.ctypes.data, succ.ctypes.data if len(succ) else None, n, windowSize, maxRefCount, minIntervalLength, zetaK, rangeNodes, 0, device)
        # one call when the guess is large enough (5 bytes per arc, 2 per node), a second one with the exact size otherwise
        data = np.empty(5 * len(succ) + 2 * n + 1024, dtype=np.uint8)   # the device call fills what it reports
        rc = L.bvg_bv_compress(*args, data.ctypes.data, len(data), C.byref(need), node_bits.ctypes.data, C.byref(ms))
        if rc == BVG_ENOMEM:
            data = np.zeros(max(need.value, 1), dtype=np.uint8)
            rc = L.bvg_bv_compress(*args, data.ctypes.data, len(data), C.byref(need), node_bits.ctypes.data, C.byref(ms))
        _check(rc)
        data[:need.value].tofile(basename + ".graph")
        gaps = np.concatenate([[0], np.diff(node_bits)]).astype(np.uint64)
        codes, _ = tools.write_codes(tools.GAMMA, 0, gaps)
        with open(basename + ".offsets", "wb") as f:
            f.write(codes)
        bits = int(node_bits[-1])
        with open(basename + ".properties", "w") as f:
            f.write("#BVGraph properties\ngraphclass=it.unimi.dsi.webgraph.BVGraph\nversion=0\nnodes=%d\narcs=%d\n" % (n, len(succ)))
            f.write("windowsize=%d\nmaxrefcount=%d\nminintervallength=%d\nzetak=%d\ncompressionflags=\n"
                    % (windowSize, 2147483647 if maxRefCount < 0 else maxRefCount, minIntervalLength, zetaK))
            f.write("bitsperlink=%.3f\nbitspernode=%.3f\n" % (bits / max(len(succ), 1), bits / max(n, 1)))
        return bits, ms.value

    def close(self):
        if self._h:
            lib().bvg_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ImmutableGraph surface ----
    def numNodes(self):
        return self._n

    def numArcs(self):
        return self._m

    def randomAccess(self):
        return bool(lib().bvg_random_access(self._h))

    def hasCopiableIterators(self):
        return True

    def basename(self):
        return self._basename

    def windowSize(self):
        return self._window

    def maxRefCount(self):
        return self._maxref

    def minIntervalLength(self):
        return self._minlen

    def zetaK(self):
        return self._zetak

    def copy(self):
        return self  # the native graph is immutable and shareable (ImmutableGraph.java:157-165)

    def outdegree(self, x):
        d = C.c_int32()
        _check(lib().bvg_outdegree(self._h, x, C.byref(d)), self._h)
        return d.value

    def successorArray(self, x):
        """ImmutableGraph.successorArray (:329-333): a fresh array holding exactly the successors of x."""
        d = self.outdegree_checked(x)
        out = np.empty(max(d, 1), dtype=np.int32)
        dd = C.c_int32()
        _check(lib().bvg_successors(self._h, x, out.ctypes.data, len(out), C.byref(dd)), self._h)
        return out[:dd.value]

    def outdegree_checked(self, x):
        if not (0 <= x < self._n):
            raise ValueError("Node index out of range: %d" % x)  # BVGraph.java:900
        if not self.randomAccess():
            raise UnsupportedOperationError("Random access to successor lists is not possible with sequential or offline graphs")  # :901
        return self.outdegree(x)

    def nodeIterator(self, frm=0):
        h = C.c_void_p()
        _check(lib().bvg_cursor_open(self._h, frm, 2**31 - 1, C.byref(h)), self._h)
        it = NodeIterator(self, h, frm)
        it._up = self._n
        return it

    # ---- batched entry points (what the JNI shim would actually call) ----
    def extent(self):
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib().bvg_extent(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def rangeArcs(self, frm, to):
        a = C.c_int64()
        _check(lib().bvg_range_arcs(self._h, frm, to, C.byref(a)), self._h)
        return a.value

    def decodeRange(self, frm, to):
        """Successor lists of frm..to-1 as CSR (offsets int64[to-frm+1], successors int32[arcs]) in host memory."""
        arcs = self.rangeArcs(frm, to)
        off = np.zeros(to - frm + 1, dtype=np.int64)
        out = np.empty(max(arcs, 1), dtype=np.int32)
        _check(lib().bvg_decode_range(self._h, frm, to, off.ctypes.data, out.ctypes.data, arcs, 0), self._h)
        return off, out[:arcs]

    def scanRange(self, frm, to):
        arcs, cs = C.c_int64(), C.c_uint64()
        _check(lib().bvg_scan_range(self._h, frm, to, C.byref(arcs), C.byref(cs)), self._h)
        return arcs.value, cs.value

    def successorsBatch(self, xs):
        xs = np.ascontiguousarray(xs, dtype=np.int32)
        off = np.zeros(len(xs) + 1, dtype=np.int64)
        _check(lib().bvg_successors_batch(self._h, xs.ctypes.data, len(xs), off.ctypes.data, None, 0, 0), self._h)
        out = np.empty(max(int(off[-1]), 1), dtype=np.int32)
        _check(lib().bvg_successors_batch(self._h, xs.ctypes.data, len(xs), off.ctypes.data, out.ctypes.data, int(off[-1]), 0), self._h)
        return off, out[:off[-1]]

    def outdegreeBatch(self, xs):
        xs = np.ascontiguousarray(xs, dtype=np.int32)
        d = np.zeros(len(xs), dtype=np.int32)
        _check(lib().bvg_outdegree_batch(self._h, xs.ctypes.data, 0, len(xs), d.ctypes.data, 0), self._h)
        return d

    def setStream(self, cuda_stream):
        _check(lib().bvg_set_stream(self._h, C.c_void_p(cuda_stream)))

    def profile(self, enable):
        _check(lib().bvg_profile(self._h, 1 if enable else 0))

    def profileRead(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        _check(lib().bvg_profile_read(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def memoryFootprint(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().bvg_memory_footprint(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"stream_bytes": a.value, "offsets_bytes": b.value, "index_bytes": c.value}

    def indegrees(self, frm=None, to=None, counts=None):
        """In-degree of every node from the arcs leaving [frm, to) (the counting pass of Transform.transposeOffline, reference
        Transform.java:977-987), counted on the device by the scan itself.  Returns a uint32 array of numNodes() entries."""
        ext_from, ext_to = C.c_int32(), C.c_int32()
        _check(lib().bvg_extent(self._h, C.byref(ext_from), C.byref(ext_to), None, None))
        frm = ext_from.value if frm is None else frm
        to = ext_to.value if to is None else to
        if counts is None:
            counts = np.zeros(self._n, dtype=np.uint32)
        arcs = C.c_int64()
        _check(lib().bvg_indegrees(self._h, frm, to, counts.ctypes.data, len(counts), 0, C.byref(arcs)))
        return counts

    def bfs(self, source):
        """Distances of a breadth-first visit from `source` (-1 = unreachable), the eccentricity of the source and the number of
        nodes reached (reference algo/ParallelBreadthFirstVisit.java:155-181)."""
        dist = np.empty(self._n, dtype=np.int32)
        levels, reached = C.c_int32(), C.c_int64()
        _check(lib().bvg_bfs(self._h, source, dist.ctypes.data, 0, C.byref(levels), C.byref(reached)))
        return dist, levels.value, reached.value

    def hyperballStep(self, counters, log2m, frm=None, to=None):
        """One HyperBall iteration (HyperBall.java:875-915) on byte registers: counters uint8[numNodes, 2^log2m] ->
        (new counters, number of modified nodes); rows outside [frm, to) are returned unchanged."""
        frm, to = (0 if frm is None else frm), (self._n if to is None else to)
        cin = np.ascontiguousarray(counters, dtype=np.uint8)
        if cin.shape != (self._n, 1 << log2m):
            raise ValueError("counters must be uint8[numNodes, 2^log2m]")
        out = cin.copy()
        mod = C.c_int64()
        _check(lib().bvg_hyperball_step(self._h, frm, to, log2m, cin.ctypes.data, out.ctypes.data, 0, C.byref(mod)), self._h)
        return out, mod.value

    def scanBits(self):
        """What a scan of this graph's extent reads of the stream (bvg_scan_bits)."""
        out = (C.c_int64 * 6)()
        _check(lib().bvg_scan_bits(self._h, out))
        return {"extent_bits": out[0], "long_records": out[1], "long_residual_bits": out[2], "long_preexpanded_bits": out[3],
                "long_arcs": out[4], "schedules_built": bool(out[5])}

    @property
    def graphBits(self):
        return self._graph_bits

    @property
    def handle(self):
        return self._h


def kernel_launches():
    return int(lib().bvg_kernel_launches())


def replan_shards(basename, old_bounds, old_cost):
    """Cuts at equal shares of a measured cost (bvg_replan_shards): old_cost[j] is what shard j of old_bounds cost."""
    n = len(old_bounds) - 1
    ob = (C.c_int32 * (n + 1))(*old_bounds)
    oc = (C.c_double * n)(*[float(c) for c in old_cost])
    b = (C.c_int32 * (n + 1))()
    _check(lib().bvg_replan_shards(os.fsencode(basename), n, ob, oc, b))
    return list(b)


def plan_shards(basename, nshards):
    """Bit-balanced contiguous node ranges: bounds[0..nshards]."""
    b = (C.c_int32 * (nshards + 1))()
    _check(lib().bvg_plan_shards(os.fsencode(basename), nshards, b))
    return list(b)
