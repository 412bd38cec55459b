"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")


class OrcGraph(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int64), ("window", C.c_int32), ("maxref", C.c_int32),
                ("minlen", C.c_int32), ("zetak", C.c_int32), ("flags", C.c_uint32),
                ("outdegree_coding", C.c_int), ("block_coding", C.c_int), ("residual_coding", C.c_int),
                ("reference_coding", C.c_int), ("block_count_coding", C.c_int), ("offset_coding", C.c_int),
                ("graph", C.POINTER(C.c_uint8)), ("graph_bytes", C.c_uint64),
                ("offsets", C.POINTER(C.c_uint64))]


class OrcEFGraph(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int64), ("upper_bound", C.c_int32), ("log2_quantum", C.c_int32),
                ("words", C.POINTER(C.c_uint64)), ("nwords", C.c_uint64), ("offsets", C.POINTER(C.c_uint64))]


class OrcLabels(C.Structure):
    _fields_ = [("kind", C.c_int), ("width", C.c_int), ("n", C.c_int32), ("labels", C.POINTER(C.c_uint8)),
                ("label_bytes", C.c_uint64), ("offsets", C.POINTER(C.c_uint64))]


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("bvg_oracle.c", "bvg_oracle_mt.c", "bvg_oracle.h", "Makefile")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB


class OracleError(Exception):
    def __init__(self, code):
        super().__init__("oracle error %d" % code)
        self.code = code


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        P = C.POINTER
        lib.orc_load.argtypes = [C.c_char_p, C.c_int, P(P(OrcGraph))]
        lib.orc_from_memory.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_uint32, P(P(OrcGraph))]
        lib.orc_free.argtypes = [P(OrcGraph)]
        lib.orc_free.restype = None
        lib.orc_outdegree.argtypes = [P(OrcGraph), C.c_int32, P(C.c_int32)]
        lib.orc_successors.argtypes = [P(OrcGraph), C.c_int32, C.c_void_p, C.c_int64]
        lib.orc_successors.restype = C.c_int64
        lib.orc_decode_range.argtypes = [P(OrcGraph), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
        lib.orc_decode_range.restype = C.c_int64
        lib.orc_scan_range.argtypes = [P(OrcGraph), C.c_int32, C.c_int32, P(C.c_int64), P(C.c_uint64)]
        lib.orc_scan_range_mt.argtypes = [P(OrcGraph), C.c_int32, C.c_int32, C.c_int, P(C.c_int64), P(C.c_uint64)]
        lib.orc_rebuild_offsets.argtypes = [P(OrcGraph), C.c_void_p]
        lib.orc_chain_bits.argtypes = [P(OrcGraph), C.c_int32, P(C.c_int32)]
        lib.orc_chain_bits.restype = C.c_int64
        lib.orc_chain_root.argtypes = [P(OrcGraph), C.c_int32]
        lib.orc_chain_root.restype = C.c_int32
        lib.orc_read_code.argtypes = [C.c_void_p, C.c_uint64, P(C.c_uint64), C.c_int, C.c_int]
        lib.orc_read_code.restype = C.c_uint64
        lib.orc_ef_load.argtypes = [C.c_char_p, P(P(OrcEFGraph))]
        lib.orc_ef_free.argtypes = [P(OrcEFGraph)]
        lib.orc_ef_free.restype = None
        lib.orc_ef_outdegree.argtypes = [P(OrcEFGraph), C.c_int32, P(C.c_int32)]
        lib.orc_ef_successors.argtypes = [P(OrcEFGraph), C.c_int32, C.c_void_p, C.c_int64]
        lib.orc_ef_successors.restype = C.c_int64
        lib.orc_labels_load.argtypes = [C.c_char_p, C.c_int32, P(P(OrcLabels))]
        lib.orc_labels_free.argtypes = [P(OrcLabels)]
        lib.orc_labels_free.restype = None
        lib.orc_labels_node.argtypes = [P(OrcLabels), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
        lib.orc_labels_node.restype = C.c_int64
        lib.orc_labels_range.argtypes = [P(OrcLabels), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, P(C.c_uint64)]
        lib.orc_labels_range.restype = C.c_int64

    def load_ef(self, basename):
        g = C.POINTER(OrcEFGraph)()
        rc = self.lib.orc_ef_load(basename.encode(), C.byref(g))
        if rc:
            raise OracleError(rc)
        return OracleEFGraph(self, g)

    def load_labels(self, basename, n):
        l = C.POINTER(OrcLabels)()
        rc = self.lib.orc_labels_load(basename.encode(), n, C.byref(l))
        if rc:
            raise OracleError(rc)
        return OracleLabels(self, l)

    def load(self, basename, offsets=True):
        g = C.POINTER(OrcGraph)()
        rc = self.lib.orc_load(basename.encode(), 1 if offsets else 0, C.byref(g))
        if rc:
            raise OracleError(rc)
        return OracleGraph(self, g)

    def read_codes(self, data: bytes, coding: int, k: int, count: int):
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        pos = C.c_uint64(0)
        out = []
        for _ in range(count):
            out.append(int(self.lib.orc_read_code(buf.ctypes.data, len(buf), C.byref(pos), coding, k)))
        return out, int(pos.value)


class OracleGraph:
    def __init__(self, orc, g):
        self.orc, self.g = orc, g
        c = g.contents
        self.n, self.m = c.n, c.m
        self.window, self.maxref, self.minlen, self.zetak, self.flags = c.window, c.maxref, c.minlen, c.zetak, c.flags
        self.graph_bytes = c.graph_bytes

    def close(self):
        if self.g:
            self.orc.lib.orc_free(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def offsets(self):
        c = self.g.contents
        return np.ctypeslib.as_array(c.offsets, shape=(self.n + 1,)).copy()

    def outdegree(self, x):
        d = C.c_int32()
        rc = self.orc.lib.orc_outdegree(self.g, x, C.byref(d))
        if rc:
            raise OracleError(rc)
        return d.value

    def successors(self, x, cap=None):
        if cap is None:
            cap = max(self.outdegree(x), 1)
        out = np.empty(cap, dtype=np.int32)
        d = self.orc.lib.orc_successors(self.g, x, out.ctypes.data, cap)
        if d < 0:
            raise OracleError(d)
        return out[:d]

    def decode_range(self, lo, hi):
        off = np.zeros(hi - lo + 1, dtype=np.int64)
        tot = self.orc.lib.orc_decode_range(self.g, lo, hi, off.ctypes.data, None, 0)
        if tot < 0:
            raise OracleError(tot)
        out = np.empty(max(tot, 1), dtype=np.int32)
        tot2 = self.orc.lib.orc_decode_range(self.g, lo, hi, off.ctypes.data, out.ctypes.data, tot)
        if tot2 < 0:
            raise OracleError(tot2)
        return off, out[:tot]

    def scan_range(self, lo, hi, threads=1):
        arcs, cs = C.c_int64(), C.c_uint64()
        if threads == 1:
            rc = self.orc.lib.orc_scan_range(self.g, lo, hi, C.byref(arcs), C.byref(cs))
        else:
            rc = self.orc.lib.orc_scan_range_mt(self.g, lo, hi, threads, C.byref(arcs), C.byref(cs))
        if rc:
            raise OracleError(rc)
        return arcs.value, cs.value

    def rebuild_offsets(self):
        out = np.zeros(self.n + 1, dtype=np.uint64)
        rc = self.orc.lib.orc_rebuild_offsets(self.g, out.ctypes.data)
        if rc:
            raise OracleError(rc)
        return out

    def first_ancestor(self, x):
        r = self.orc.lib.orc_chain_root(self.g, x)
        if r < 0:
            raise OracleError(r)
        return r

    def chain_bits(self, x):
        dep = C.c_int32()
        b = self.orc.lib.orc_chain_bits(self.g, x, C.byref(dep))
        if b < 0:
            raise OracleError(b)
        return b, dep.value


_ORACLE = None


class OracleEFGraph:
    """orc_efgraph: EFGraph read the way EliasFanoSuccessorReader reads it (one node at a time)."""

    def __init__(self, orc, g):
        self.orc, self.g = orc, g
        c = g.contents
        self.n, self.m, self.upper_bound, self.log2_quantum = c.n, c.m, c.upper_bound, c.log2_quantum

    def close(self):
        if self.g:
            self.orc.lib.orc_ef_free(self.g)
            self.g = None

    def offsets(self):
        return np.ctypeslib.as_array(self.g.contents.offsets, shape=(self.n + 1,)).copy()

    def outdegree(self, x):
        d = C.c_int32()
        rc = self.orc.lib.orc_ef_outdegree(self.g, x, C.byref(d))
        if rc:
            raise OracleError(rc)
        return d.value

    def successors(self, x):
        d = self.outdegree(x)
        out = np.empty(max(d, 1), dtype=np.int32)
        r = self.orc.lib.orc_ef_successors(self.g, x, out.ctypes.data, d)
        if r < 0:
            raise OracleError(r)
        return out[:d]

    def decode_range(self, lo, hi):
        rows = [self.successors(x) for x in range(lo, hi)]
        off = np.zeros(hi - lo + 1, dtype=np.int64)
        np.cumsum([len(r) for r in rows], out=off[1:])
        return off, (np.concatenate(rows) if rows else np.empty(0, dtype=np.int32)).astype(np.int32)


class OracleLabels:
    """orc_labels: the label stream read the way BitStreamLabelledArcIterator reads it (one node at a time)."""

    def __init__(self, orc, l):
        self.orc, self.l = orc, l
        c = l.contents
        self.kind, self.width, self.n = c.kind, c.width, c.n

    def close(self):
        if self.l:
            self.orc.lib.orc_labels_free(self.l)
            self.l = None

    def offsets(self):
        return np.ctypeslib.as_array(self.l.contents.offsets, shape=(self.n + 1,)).copy()

    def node(self, x, d):
        """(list_off int64[d + 1], values int32[...]) of node x, whose outdegree is d."""
        lo = np.zeros(d + 1, dtype=np.int64)
        nv = self.orc.lib.orc_labels_node(self.l, x, d, lo.ctypes.data, None, 0)
        if nv < 0:
            raise OracleError(nv)
        vals = np.empty(max(nv, 1), dtype=np.int32)
        nv2 = self.orc.lib.orc_labels_node(self.l, x, d, lo.ctypes.data, vals.ctypes.data, nv)
        if nv2 < 0:
            raise OracleError(nv2)
        return lo, vals[:nv]

    def sequential(self, frm, to, row_off, store=True):
        """Integer labels of frm..to-1 read front to back; returns (values or None, sum of the values)."""
        row_off = np.ascontiguousarray(row_off, dtype=np.int64)
        arcs = int(row_off[to] - row_off[frm])
        vals = np.empty(max(arcs, 1), dtype=np.int32) if store else None
        tot = C.c_uint64()
        rc = self.orc.lib.orc_labels_range(self.l, frm, to, row_off.ctypes.data, vals.ctypes.data if store else None, arcs, C.byref(tot))
        if rc < 0:
            raise OracleError(rc)
        return (vals[:arcs] if store else None), tot.value

    def range(self, frm, to, row_off):
        """Labels of the arcs of frm..to-1 (row_off: CSR offsets of the underlying graph), as (list_off, values)."""
        los, vals, base = [np.zeros(1, dtype=np.int64)], [], 0
        for x in range(frm, to):
            lo, v = self.node(x, int(row_off[x + 1] - row_off[x]))
            los.append(lo[1:] + base)
            vals.append(v)
            base += len(v)
        return np.concatenate(los), (np.concatenate(vals) if vals else np.empty(0, dtype=np.int32))

def label_checksum(list_off, values):
    """The fold of bvg_labels_scan_range restated with numpy (tests; include/bvgraph_b200.h)."""
    M = np.uint64(0x9E3779B97F4A7C15)
    lo = np.asarray(list_off, dtype=np.int64)
    v = np.asarray(values, dtype=np.int64).astype(np.uint64)
    arcs = len(lo) - 1
    with np.errstate(over="ignore"):
        arc_of = np.repeat(np.arange(arcs, dtype=np.int64), np.diff(lo))
        i = (np.arange(len(v), dtype=np.int64) - lo[arc_of]).astype(np.uint64)
        t = np.zeros(arcs, dtype=np.uint64)
        np.add.at(t, arc_of, (v + np.uint64(1)) * (np.uint64(2) * i + np.uint64(1)))
        t += M * np.diff(lo).astype(np.uint64)
        j = np.arange(arcs, dtype=np.uint64)
        return int(((np.uint64(2) * j + np.uint64(1)) * t).sum(dtype=np.uint64))


def load():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle(C.CDLL(build()))
    return _ORACLE


def read_ascii_graph(path):
    """ASCIIGraph format (reference ASCIIGraph.java:57-61): first line n, then one line of
    space-separated successors per node.  Returns (offsets int64[n+1], successors int32[m])."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    nl = data.index(b"\n")
    n = int(data[:nl])
    body = data[nl + 1:]
    lines = body.split(b"\n")[:n]
    counts = np.fromiter((len(l.split()) for l in lines), dtype=np.int64, count=n)
    succ = np.array(body.split(), dtype=np.int64).astype(np.int32)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    assert off[-1] == len(succ)
    return off, succ


def xor_checksum(off, succ, first_node=0):
    """XOR over arcs (x,y) of (x*0x9E3779B97F4A7C15 + y) mod 2^64."""
    n = len(off) - 1
    deg = np.diff(off)
    xs = np.repeat(np.arange(first_node, first_node + n, dtype=np.uint64), deg)
    with np.errstate(over="ignore"):
        v = xs * np.uint64(0x9E3779B97F4A7C15) + succ.astype(np.int64).astype(np.uint64)
    return int(np.bitwise_xor.reduce(v)) if len(v) else 0
