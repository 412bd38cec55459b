"""Host-side mirror of the reference's arc-labelled graphs for the label classes decoded on the GPU
(reference src/it/unimi/dsi/webgraph/labelling/BitStreamArcLabelledImmutableGraph.java, ArcLabelledNodeIterator.java,
GammaCodedIntLabel.java, FixedWidthIntLabel.java, FixedWidthIntListLabel.java).  Same method names and error behaviour;
labels come back as ints (integer labels) or numpy int32 arrays (list labels).  Everything is decoded by
libbvgraph_b200.so (include/bvgraph_b200.h, "arc labels"); there is no CPU path."""
import ctypes as C
import os

import numpy as np

from .bvgraph import BVGraph, IllegalStateError, NoSuchElementError, _check, lib

GAMMA, FIXED, FIXED_LIST = 0, 1, 2
BATCH_NODES = 1 << 18


def underlying_basename(basename):
    """The underlyinggraph property resolved against the labelled graph's directory (:391-395)."""
    buf = C.create_string_buffer(4096)
    _check(lib().bvg_labels_underlying(os.fsencode(basename), buf, len(buf)))
    return os.fsdecode(buf.value)


class LabelledArcIterator:
    """ArcLabelledNodeIterator.LabelledArcIterator: nextInt() returns the next successor (-1 at the end), label() the label
    of the arc just returned (:225-262)."""

    def __init__(self, succ, labels):
        self._succ, self._labels, self._i = succ, labels, -1

    def nextInt(self):
        if self._i + 1 >= len(self._succ):
            return -1
        self._i += 1
        return int(self._succ[self._i])

    def label(self):
        if self._i < 0:
            raise IllegalStateError("label() before nextInt()")
        v = self._labels[self._i]
        return v if isinstance(v, np.ndarray) else int(v)

    def skip(self, n):
        k = max(0, min(n, len(self._succ) - 1 - self._i))
        self._i += k
        return k


class ArcLabelledNodeIterator:
    """nodeIterator() of a labelled graph: nextInt / outdegree / successorArray / labelArray / successors
    (ArcLabelledNodeIterator.java).  Decodes BATCH_NODES nodes per device call."""

    def __init__(self, alg, frm):
        self._alg, self._from, self._curr = alg, frm, frm - 1
        self._lo = self._hi = frm
        self._off = self._succ = self._labels = None

    def hasNext(self):
        return self._curr < self._alg.numNodes() - 1

    def nextInt(self):
        if not self.hasNext():
            raise NoSuchElementError("no more nodes")
        self._curr += 1
        if self._curr >= self._hi:
            self._lo, self._hi = self._curr, min(self._alg.numNodes(), self._curr + BATCH_NODES)
            self._off, self._succ = self._alg.g.decodeRange(self._lo, self._hi)
            self._labels = self._alg.labelsOfRange(self._lo, self._hi)
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

    def labelArray(self):
        a, b = self._row()
        return self._labels[a:b]

    def successors(self):
        a, b = self._row()
        return LabelledArcIterator(self._succ[a:b], self._labels[a:b])

    def __iter__(self):
        return self

    def __next__(self):
        if not self.hasNext():
            raise StopIteration
        return self.nextInt()


class BitStreamArcLabelledImmutableGraph:
    def __init__(self, g, handle, basename=None):
        self.g, self._h, self._basename = g, handle, basename
        k, w, bits, held = C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
        _check(lib().bvg_labels_info(handle, C.byref(k), C.byref(w), C.byref(bits), C.byref(held)))
        self.kind, self.width, self.labelBits, self.heldBytes = k.value, w.value, bits.value, held.value

    # ---- loading (:385-470): the underlying graph is loaded with the same method, then the labels ----
    @classmethod
    def _load(cls, basename, graph_loader, device=None):
        g = graph_loader(underlying_basename(basename), device=device)
        h = C.c_void_p()
        try:
            _check(lib().bvg_labels_open(g.handle, os.fsencode(basename), C.byref(h)), g.handle)
        except Exception:
            g.close()
            raise
        return cls(g, h, str(basename))

    @classmethod
    def load(cls, basename, device=None):
        return cls._load(basename, BVGraph.load, device)

    @classmethod
    def loadMapped(cls, basename, device=None):
        return cls._load(basename, BVGraph.loadMapped, device)

    @classmethod
    def loadOffline(cls, basename, device=None):
        return cls._load(basename, BVGraph.loadOffline, device)

    @classmethod
    def loadSequential(cls, basename, device=None):
        return cls._load(basename, BVGraph.loadSequential, device)

    @classmethod
    def fromMemory(cls, g, labels, label_offsets, kind, width=0):
        lb = np.frombuffer(labels, dtype=np.uint8)
        ob = np.frombuffer(label_offsets, dtype=np.uint8)
        h = C.c_void_p()
        _check(lib().bvg_labels_open_memory(g.handle, lb.ctypes.data if len(lb) else None, len(lb), ob.ctypes.data, len(ob),
                                            kind, width, C.byref(h)), g.handle)
        return cls(g, h)

    def close(self, close_graph=True):
        if self._h:
            lib().bvg_labels_close(self._h)
            self._h = None
        if close_graph and self.g is not None:
            self.g.close()

    def __del__(self):
        try:
            self.close(close_graph=False)
        except Exception:
            pass

    # ---- ImmutableGraph surface, delegated (:264-300) ----
    def numNodes(self):
        return self.g.numNodes()

    def numArcs(self):
        return self.g.numArcs()

    def randomAccess(self):
        return self.g.randomAccess()

    def outdegree(self, x):
        return self.g.outdegree(x)

    def successorArray(self, x):
        return self.g.successorArray(x)

    def basename(self):
        return self._basename

    # ---- labels ----
    def decodeLabels(self, frm, to):
        """(list_off int64[arcs + 1], values int32[nvalues]) for the arcs of frm..to-1 in successor order."""
        nv = C.c_int64()
        _check(lib().bvg_labels_decode_range(self._h, frm, to, None, None, 0, 0, C.byref(nv)), self.g.handle)
        arcs = self.g.rangeArcs(frm, to)
        lo = np.zeros(arcs + 1, dtype=np.int64)
        vals = np.empty(max(nv.value, 1), dtype=np.int32)
        _check(lib().bvg_labels_decode_range(self._h, frm, to, lo.ctypes.data, vals.ctypes.data, nv.value, 0, C.byref(nv)), self.g.handle)
        return lo, vals[:nv.value]

    def labelsOfRange(self, frm, to):
        """One entry per arc: an int32 array (integer labels) or a list of int32 arrays (list labels)."""
        lo, vals = self.decodeLabels(frm, to)
        if self.kind != FIXED_LIST:
            return vals
        return [vals[lo[j]:lo[j + 1]] for j in range(len(lo) - 1)]

    def labelArray(self, x):
        return self.labelsOfRange(x, x + 1)

    def scanLabels(self, frm, to):
        arcs, nv, cs = C.c_int64(), C.c_int64(), C.c_uint64()
        _check(lib().bvg_labels_scan_range(self._h, frm, to, C.byref(arcs), C.byref(nv), C.byref(cs)), self.g.handle)
        return arcs.value, nv.value, cs.value

    def successors(self, x):
        if not self.randomAccess():
            raise IllegalStateError("random access to a graph loaded without offsets")
        return LabelledArcIterator(self.g.successorArray(x), self.labelArray(x))

    def nodeIterator(self, frm=0):
        return ArcLabelledNodeIterator(self, frm)

