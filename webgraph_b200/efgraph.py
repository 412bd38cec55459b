"""Host-side mirror of the reference's EFGraph (src/it/unimi/dsi/webgraph/EFGraph.java) over libbvgraph_b200.so: the same
ImmutableGraph surface as bvgraph.BVGraph (load*, numNodes, numArcs, outdegree, successors, successorArray, nodeIterator), the
lists decoded on the device (include/bvgraph_b200.h, "EFGraph").  There is no CPU path."""
import ctypes as C
import os

import numpy as np

from .bvgraph import ImmutableGraph, IllegalStateError, NoSuchElementError, LazyIntIterator, _check, lib

BATCH_NODES = 1 << 18


class EFNodeIterator:
    """nodeIterator(from) of an EFGraph (the reference inherits ImmutableGraph's generic iterator, ImmutableGraph.java:254-420):
    BATCH_NODES nodes per device call."""

    def __init__(self, g, frm, upper=None):
        self._g, self._from, self._curr = g, frm, frm - 1
        self._upper = g.numNodes() if upper is None else min(upper, g.numNodes())
        self._lo = self._hi = frm
        self._off = self._succ = None

    def hasNext(self):
        return self._curr < self._upper - 1

    def nextInt(self):
        if not self.hasNext():
            raise NoSuchElementError("no more nodes")
        self._curr += 1
        if self._curr >= self._hi:
            self._lo, self._hi = self._curr, min(self._upper, self._curr + BATCH_NODES)
            self._off, self._succ = self._g.decodeRange(self._lo, self._hi)
        return self._curr

    def _row(self):
        if self._curr == self._from - 1:
            raise IllegalStateError("nextInt() has not been called")
        i = self._curr - self._lo
        return int(self._off[i]), int(self._off[i + 1])

    def outdegree(self):
        a, b = self._row()
        return b - a

    def successorArray(self):
        a, b = self._row()
        return self._succ[a:b].copy()

    def successors(self):
        a, b = self._row()
        return LazyIntIterator(self._succ[a:b])

    def copy(self, upperBound=2 ** 31 - 1):
        return EFNodeIterator(self._g, self._curr + 1, upperBound)

    def __iter__(self):
        return self

    def __next__(self):
        if not self.hasNext():
            raise StopIteration
        return self.nextInt()


class EFGraph(ImmutableGraph):
    def __init__(self, handle, basename=None):
        self._h, self._basename = handle, basename
        n, m, ub, q, bits = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32(), C.c_int64()
        _check(lib().bvg_ef_info(handle, C.byref(n), C.byref(m), C.byref(ub), C.byref(q), C.byref(bits)))
        self._n, self._m, self.upperBound, self.quantum, self.graphBits = n.value, m.value, ub.value, q.value, bits.value

    # load / loadMapped / loadOffline / loadSequential all end in loadInternal (EFGraph.java:576-640, 709-790)
    @classmethod
    def load(cls, basename, device=-1):
        h = C.c_void_p()
        _check(lib().bvg_ef_open(os.fsencode(basename), device, C.byref(h)))
        return cls(h, str(basename))

    loadMapped = loadOffline = loadSequential = load

    @staticmethod
    def store(basename, off, succ, upperBound=0, log2Quantum=8, device=-1):
        """EFGraph.store (EFGraph.java:812-888) with the stream written on the device (bvg_ef_compress): <basename>.graph
        (little-endian long words), .offsets (delta-coded gaps, written here from the node bit offsets the device returns) and
        .properties.  Returns (graph bits, milliseconds of the device kernels)."""
        from . import tools
        off = np.ascontiguousarray(off, dtype=np.int64)
        succ = np.ascontiguousarray(succ, dtype=np.int32)
        n = len(off) - 1
        need, ms = C.c_uint64(0), C.c_double(0)
        node_bits = np.zeros(n + 1, dtype=np.int64)
        L = lib()
        args = (off.ctypes.data, succ.ctypes.data if len(succ) else None, n, upperBound, log2Quantum, 0, device)
        # one call when the guess is large enough (6 bytes per arc, 16 per node), a second one with the exact size otherwise
        words = np.empty(6 * len(succ) + 16 * n + 1024, dtype=np.uint8)   # the device call fills what it reports
        rc = L.bvg_ef_compress(*args, words.ctypes.data, len(words), C.byref(need), node_bits.ctypes.data, C.byref(ms))
        if rc == -6:   # BVG_ENOMEM carries the size needed
            words = np.zeros(max(need.value, 8), dtype=np.uint8)
            rc = L.bvg_ef_compress(*args, words.ctypes.data, len(words), C.byref(need), node_bits.ctypes.data, C.byref(ms))
        _check(rc)
        words[:need.value].tofile(basename + ".graph")
        gaps = np.concatenate([[0], np.diff(node_bits)]).astype(np.uint64)
        data, _ = tools.write_codes(tools.DELTA, 0, gaps)
        with open(basename + ".offsets", "wb") as f:
            f.write(data)
        ub = upperBound if upperBound > 0 else n
        with open(basename + ".properties", "w") as f:
            f.write("#EFGraph properties\nnodes=%d\narcs=%d\n" % (n, len(succ)))
            if ub != n:
                f.write("upperbound=%d\n" % ub)
            f.write("quantum=%d\nbyteorder=LITTLE_ENDIAN\ngraphclass=it.unimi.dsi.webgraph.EFGraph\nversion=0\n" % (1 << log2Quantum))
        return int(node_bits[-1]), ms.value

    def close(self):
        if self._h:
            lib().bvg_ef_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def numNodes(self):
        return self._n

    def numArcs(self):
        return self._m

    def randomAccess(self):
        return True   # EFGraph.java:1045-1047

    def hasCopiableIterators(self):
        return True

    def basename(self):
        return self._basename

    def copy(self):
        return self

    def _checked(self, rc):
        if rc in (-4, -5):
            node, pos = C.c_int32(-1), C.c_int64(-1)
            lib().bvg_ef_last_error_node(self._h, C.byref(node), C.byref(pos))
        _check(rc)

    def outdegree(self, x):
        d = C.c_int32()
        self._checked(lib().bvg_ef_outdegree(self._h, x, C.byref(d)))
        return d.value

    def successorArray(self, x):
        d = self.outdegree(x)
        out = np.empty(max(d, 1), dtype=np.int32)
        got = C.c_int32()
        self._checked(lib().bvg_ef_successors(self._h, x, out.ctypes.data, d, C.byref(got)))
        return out[:got.value]

    def nodeIterator(self, frm=0):
        if frm < 0 or frm > self._n:
            raise ValueError("node index out of range: %d" % frm)
        return EFNodeIterator(self, frm)

    def rangeArcs(self, frm, to):
        a = C.c_int64()
        self._checked(lib().bvg_ef_range_arcs(self._h, frm, to, C.byref(a)))
        return a.value

    def decodeRange(self, frm, to):
        arcs = self.rangeArcs(frm, to)
        off = np.zeros(to - frm + 1, dtype=np.int64)
        out = np.empty(max(arcs, 1), dtype=np.int32)
        self._checked(lib().bvg_ef_decode_range(self._h, frm, to, off.ctypes.data, out.ctypes.data, arcs, 0))
        return off, out[:arcs]

    def scanRange(self, frm, to):
        arcs, cs = C.c_int64(), C.c_uint64()
        self._checked(lib().bvg_ef_scan_range(self._h, frm, to, C.byref(arcs), C.byref(cs)))
        return arcs.value, cs.value

    @property
    def handle(self):
        return self._h
