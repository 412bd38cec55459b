"""ctypes binding of the host-side tools (libbvgraph_tools.so): BVGraph compressor and the seeded
synthetic power-law generator.  See include/bvgraph_tools.h.  These PRODUCE inputs; they never decode."""
import ctypes as C
import os

import numpy as np

from . import build as _build

DELTA, GAMMA, GOLOMB, SKEWED_GOLOMB, UNARY, ZETA, NIBBLE = 1, 2, 3, 4, 5, 6, 7
# flag words, reference BVGraph.java:474-523
OUTDEGREES_DELTA = DELTA
BLOCKS_DELTA = DELTA << 4
BLOCKS_UNARY = UNARY << 4
RESIDUALS_GAMMA = GAMMA << 8
RESIDUALS_DELTA = DELTA << 8
RESIDUALS_NIBBLE = NIBBLE << 8
RESIDUALS_GOLOMB = GOLOMB << 8
REFERENCES_GAMMA = GAMMA << 12
REFERENCES_DELTA = DELTA << 12
BLOCK_COUNT_DELTA = DELTA << 16
BLOCK_COUNT_UNARY = UNARY << 16
OFFSETS_DELTA = DELTA << 20


class StoreStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "nodes", "arcs", "graph_bits", "offsets_bits", "bits_outdegrees", "bits_references", "bits_blocks",
        "bits_intervals", "bits_residuals", "copied_arcs", "intervalised_arcs", "residual_arcs", "tot_ref", "tot_dist")] + [
        ("max_outdegree", C.c_int32), ("max_ref_chain", C.c_int32), ("xor_checksum", C.c_uint64), ("sum_successors", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GenParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("target_arcs", C.c_int64), ("seed", C.c_uint64), ("zipf_s", C.c_double),
                ("p_copy", C.c_double), ("p_interval", C.c_double), ("p_local", C.c_double), ("block", C.c_int32),
                ("max_degree", C.c_int32), ("copy_run", C.c_double), ("skip_run", C.c_double), ("local_bits", C.c_int32),
                ("interval_max", C.c_int32), ("p_same_degree", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.tools_library())
        _lib.bvgt_store_csr.argtypes = [C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_uint32, C.c_int, C.POINTER(StoreStats)]
        _lib.bvgt_gen_defaults.argtypes = [C.POINTER(GenParams), C.c_int32, C.c_int64, C.c_uint64]
        _lib.bvgt_gen_defaults.restype = None
        _lib.bvgt_generate_store.argtypes = [C.c_char_p, C.POINTER(GenParams), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(StoreStats)]
        _lib.bvgt_write_codes.argtypes = [C.c_int, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        _lib.bvgt_write_codes.restype = C.c_int64
        _lib.bvgt_store_ef.argtypes = [C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        _lib.bvgt_store_labels.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    return _lib


def write_codes(coding, k, values):
    """`values` written back to back in one of dsiutils' instantaneous codes; returns (bytes, number of bits)."""
    v = np.ascontiguousarray(values, dtype=np.uint64)
    cap = 64
    while True:
        out = np.zeros(cap, dtype=np.uint8)
        nbits = lib().bvgt_write_codes(coding, k, v.ctypes.data, len(v), out.ctypes.data, cap)
        if nbits >= 0:
            return out[:(nbits + 7) // 8].tobytes(), int(nbits)
        if nbits != -1 or cap > (1 << 30):
            raise ValueError("bvgt_write_codes failed: %d" % nbits)
        cap *= 8


def store_csr(basename, off, succ, window=7, maxref=3, minlen=4, zetak=3, flags=0, threads=1):
    """BVGraph.store for a CSR graph (defaults: reference BVGraph.java:454-472)."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    succ = np.ascontiguousarray(succ, dtype=np.int32)
    st = StoreStats()
    rc = lib().bvgt_store_csr(os.fsencode(basename), len(off) - 1, off.ctypes.data, succ.ctypes.data if len(succ) else None,
                              window, maxref, minlen, zetak, flags, threads, C.byref(st))
    if rc:
        raise ValueError("bvgt_store_csr failed: %d" % rc)
    return st.as_dict()


def gen_params(n, target_arcs, seed=0x5EED, **kw):
    p = GenParams()
    lib().bvgt_gen_defaults(C.byref(p), n, target_arcs, seed)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def generate_store(basename, n, target_arcs, seed=0x5EED, window=7, maxref=3, minlen=4, zetak=3, flags=0,
                   threads=None, return_csr=False, **kw):
    """Synthetic block-local copy-model power-law graph compressed straight to <basename>.*"""
    p = gen_params(n, target_arcs, seed, **kw)
    threads = threads or (os.cpu_count() or 1)
    st = StoreStats()
    off = succ = None
    if return_csr:
        off = np.zeros(n + 1, dtype=np.int64)
        cap = int(target_arcs * 1.25) + 4 * n + 1024
        succ = np.empty(cap, dtype=np.int32)
    rc = lib().bvgt_generate_store(os.fsencode(basename), C.byref(p), window, maxref, minlen, zetak, flags, threads,
                                   off.ctypes.data if return_csr else None, succ.ctypes.data if return_csr else None,
                                   len(succ) if return_csr else 0, C.byref(st))
    if rc:
        raise ValueError("bvgt_generate_store failed: %d" % rc)
    if return_csr:
        return st.as_dict(), off, succ[:off[-1]].copy()
    return st.as_dict()


LABEL_GAMMA, LABEL_FIXED, LABEL_FIXED_LIST = 0, 1, 2


def store_labels(basename, underlying, off, values, kind, width=0, list_off=None, key="TEST", threads=1):
    """Writes <basename>.labels / .labeloffsets / .properties for the graph whose CSR row offsets are `off`
    (BitStreamArcLabelledGraphTest.java:131-203).  `values`: one int per arc (LABEL_GAMMA, LABEL_FIXED) or the
    concatenated lists addressed by `list_off` (LABEL_FIXED_LIST).  Returns the length of the label stream in bits."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    values = np.ascontiguousarray(values, dtype=np.int32)
    lo = None if list_off is None else np.ascontiguousarray(list_off, dtype=np.int64)
    bits = C.c_int64(0)
    rc = lib().bvgt_store_labels(os.fsencode(basename), os.fsencode(underlying), key.encode(), len(off) - 1, off.ctypes.data,
                                 None if lo is None else lo.ctypes.data, values.ctypes.data if len(values) else None,
                                 kind, width, threads, C.byref(bits))
    if rc:
        raise ValueError("bvgt_store_labels failed: %d" % rc)
    return int(bits.value)


def store_ef(basename, off, succ, upper_bound=0, log2_quantum=8, big_endian=False, threads=1):
    """EFGraph.store for a CSR graph (reference EFGraph.java:812-888; defaults :94, upper bound = n).  Returns the bits of the
    graph stream."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    succ = np.ascontiguousarray(succ, dtype=np.int32)
    bits = C.c_int64(0)
    rc = lib().bvgt_store_ef(os.fsencode(basename), len(off) - 1, off.ctypes.data, succ.ctypes.data if len(succ) else None,
                             upper_bound, log2_quantum, 1 if big_endian else 0, threads, C.byref(bits))
    if rc:
        raise ValueError("bvgt_store_ef failed: %d" % rc)
    return int(bits.value)
