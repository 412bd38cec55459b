// bvg_stream.cuh -- the extras of a consume-only scan decoded by stream position instead of by record.
//
// One lane per record (bvg_scan.cuh, k_scan_extras_lean) needs a length-sorted schedule to keep the 32 lanes of a warp
// busy, and it cannot split a record: records of more than a few hundred successors go through the sync points of the
// long index in kernels of their own.  Here the .graph stream is cut into chunks of STREAM_CHUNK_BITS bits and every
// lane decodes one chunk, through whatever record boundaries lie in it, exactly as BVGraphNodeIterator walks the stream
// (reference BVGraph.java:1136-1213: one record after the other, no index) -- so all lanes carry the same number of
// bits, in stream order, with no schedule and no distinction between short and long records.  What makes a chunk
// enterable is one 16-byte entry per chunk, built once at open (k_stream_entries): the first position at or after the
// chunk's first bit where the decoder's state is small -- a record start, or a position inside a residual run together
// with the node, the number of residuals left and the running successor value.
//
// The lanes of a warp are in different places of their records, so the loop is flat: a residual step (the hot path,
// BVGraph.java:939-972) is taken whenever most lanes have residuals left, and the record prologue (outdegree, reference,
// copy blocks, intervals: BVGraph.java:1048-1100) is taken by all the lanes that need one at the same time once enough of
// them wait (a warp vote), so that it is not paid on every trip.
//
// The kernel does the extras only: copied successors are merged level by level by k_scan_merge_lean / k_long_merge from
// the lists this kernel stores (stored records write their residuals right-aligned into their CSR-shaped row, intervals
// are merged in front of them when the record ends, as ScanExtras::iv_merge does); interval sections and merges of long
// records stay with k_long_extras / k_long_merge.
#pragma once
#include "bvg_device.cuh"
#include "bvg_scan.cuh"
#include "bvg_long.cuh"

namespace bvg {

#ifndef BVG_STREAM_CHUNK_BITS
#define BVG_STREAM_CHUNK_BITS 2048
#endif
constexpr uint32_t STREAM_CHUNK_BITS = BVG_STREAM_CHUNK_BITS;
constexpr uint32_t STREAM_FIRST = 0x80000000u;   // StreamEntry.rem: the next residual is the first of its run (x + nat2int)

struct StreamEntry {
    uint32_t dpos;   // entry position - chunk start (bits); the chunk is [start, next entry position)
    int32_t x;       // node the position belongs to: a record start when rem == 0, else inside its residual run
    uint32_t rem;    // residuals left in the run (| STREAM_FIRST)
    uint32_t v;      // successor value before the position (unused with STREAM_FIRST)
};

// What a lane needs to place and fold the residuals of the record it is in.
struct StreamRec {
    int32_t x;
    uint32_t d;          // outdegree
    uint32_t rem;        // residuals left
    uint32_t v;          // running successor value
    bool first;          // next residual is the first of the run
    bool fold;           // consumed by this scan (x in [from, to)), extras folded here
    bool store;          // residuals are written: tail of the row (short records) or the long record's residual buffer
    bool is_long;
    int32_t* dst;        // where residual with `rem` left goes: dst[-rem]  (dst = one past the last residual's slot)
};

// Lookup of a long record by node id (long_nodes ascending).
__device__ __forceinline__ int32_t long_find(const int32_t* __restrict__ long_nodes, int32_t nlong, int32_t x) {
    int32_t lo = 0, hi = nlong;
    while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (long_nodes[mid] < x) lo = mid + 1; else hi = mid; }
    return (lo < nlong && long_nodes[lo] == x) ? lo : -1;
}

struct StreamArgs {
    const StreamEntry* __restrict__ entries;   // nchunks + 1 (the last one: end of the stream, rem = 0)
    int64_t nchunks;                            // chunks of the graph object's stream
    int64_t first_chunk, count;                 // chunks this launch walks
    uint64_t bit0;                              // bit position (from word 0 of the stream buffer) of chunk 0
    int32_t lo, hi, from;                       // nodes [lo, hi) are decoded (lo <= from: halo nodes only as parents), [from, hi) folded
    const uint8_t* __restrict__ is_parent;      // [x - node_lo]: somebody copies from x
    const int32_t* __restrict__ long_nodes;
    int32_t nlong, long_d;
    LongIndex li;
    int32_t* long_tmp;                          // LongDst.tmp
    unsigned long long* result;
    int debug_nostore;
    // stored records with intervals whose residual run was decoded by more than one lane: their intervals are merged in front
    // of the residuals by k_stream_ivfix once every lane is done (the lane that ends the run cannot know that the others are)
    int32_t* defer_list;
    unsigned int* defer_count;
};

// ScanExtras::iv_merge for a stored short record: residuals sit right-aligned in row_extras[nout - rc .. nout), the interval
// section starts at iv_pos.  Out of line: one call per stored record with intervals, and the hot loop stays small.
template <int K>
__device__ BVG_NOINLINE void stream_iv_merge(const GraphDev& g, int32_t x, uint32_t nout, uint32_t ic, uint64_t iv_pos, int32_t* row_extras) {
    ScanExtras<K, Win> w;
    w.x = x; w.nout = (int32_t)nout; w.ic = ic; w.iv_pos = iv_pos; w.err = 0;
    int64_t total = 0;
    {
        Win c;
        c.seek(g, iv_pos);
        uint32_t prev = 0;
        for (uint32_t i = 0; i < ic; i++) {
            uint32_t left;
            if (i == 0) left = (uint32_t)(int32_t)(nat2int(c.gamma(g)) + (int64_t)x);
            else left = prev + 1u + (uint32_t)c.gamma(g);
            const uint32_t len = (uint32_t)(c.gamma(g) + (uint64_t)g.c.minlen);
            total += len;
            prev = left + len;
        }
    }
    w.rc = (int32_t)((int64_t)nout - total);   // iv_merge only needs nout - rc: where the residuals start
    w.iv_merge(g, row_extras);
}

// Everything a lane does, written against a row locator RM (RowMap) so that the host emulation can use a flat one.
template <int K, class RM>
struct StreamLane {
    Win b;
    Fold32 f;
    StreamRec r;
    uint64_t end;            // the lane stops at the first position >= end (a code boundary by construction of the entries)
    int64_t left;            // end - position of the window, kept by the residual steps (position arithmetic is 64-bit and slow)
    uint32_t folded;         // successors folded for the current record (Fold32.n)
    uint64_t iv_pos;         // stored short record in progress: position of its first interval code (for iv_merge), ~0 = none / unknown
    uint32_t ic;             // its interval count
    int32_t* row_extras;     // its row + copied
    uint32_t nout;
    unsigned long long acc;
    long long arcs;
    bool done;

    __device__ __forceinline__ void fail(const GraphDev& g, int code) {
        report(g.err, code, r.x, b.pos(g) + g.bit_base);
        done = true; r.rem = 0;
    }

    __device__ __forceinline__ bool wanted_node(const GraphDev& g, const StreamArgs& a, const RM& rm, int32_t x) const {
        return x >= a.lo && x < a.hi && rm.wanted(g, x);
    }

    // close the fold of the record in progress
    __device__ __forceinline__ void flush_fold() {
        if (folded) {
            f.n = folded; acc ^= f.finish(r.x);
            if (r.is_long) arcs += folded;   // short records count their outdegree at the prologue, long ones part by part
        }
        folded = 0;
    }

    // (re)derives the placement of node x's residuals from the index arrays; `rem` residuals of the run are left
    __device__ __forceinline__ void locate(const GraphDev& g, const StreamArgs& a, const RM& rm, int32_t x, uint32_t d, uint32_t rem, bool have_meta,
                                           int32_t l) {
        r.x = x; r.d = d; r.rem = rem;
        const bool wanted = wanted_node(g, a, rm, x);
        const bool stored = wanted && a.is_parent[x - g.node_lo] != 0 && a.debug_nostore != 1;
        const bool consumed = wanted && x >= a.from;
        r.is_long = (int32_t)d > a.long_d && a.nlong > 0;
        r.store = false; r.fold = false; r.dst = nullptr;
        f.begin(x);
        folded = 0;
        if (r.is_long) {
            if (!have_meta) l = long_find(a.long_nodes, a.nlong, x);
            if (l < 0) { r.is_long = false; }
        }
        if (r.is_long) {
            const LongMeta& m = a.li.meta[l];
            // as k_long_resid: stored long records write their residuals where k_long_extras / k_long_merge expect them and fold
            // them here only when that is already their final list; consumed ones are folded here
            if (stored) {
                r.store = true;
                int32_t* base;
                if (m.ic == 0) base = m.copied == 0 ? rm.row(g, x) : a.long_tmp + m.tmp_off + m.d;
                else base = a.long_tmp + m.tmp_off;
                r.dst = base + m.rc;
                r.fold = consumed && m.ic == 0 && m.copied == 0;
            } else r.fold = consumed;
        } else {
            r.fold = consumed;
            if (stored) { r.store = true; r.dst = rm.row(g, x) + d; }   // right-aligned: the last residual is the row's last slot
        }
    }

    __device__ __forceinline__ void init() {
        acc = 0; arcs = 0; done = true; folded = 0; iv_pos = ~0ull; ic = 0; row_extras = nullptr; nout = 0; end = 0; left = 0;
        r.x = 0; r.d = 0; r.rem = 0; r.v = 0; r.first = false; r.fold = false; r.store = false; r.is_long = false; r.dst = nullptr;
        f.begin(0);
        b.p0 = nullptr; b.idx = 0; b.lim = 0; b.w0 = b.w1 = b.q0 = b.q1 = b.q2 = 0; b.s = 0;
    }
    __device__ __forceinline__ void sync_left(const GraphDev& g) { left = (int64_t)end - (int64_t)b.pos(g); }

    // Opens the lane at the entry of chunk c (acc / arcs run on across the chunks a lane takes).
    __device__ __forceinline__ void open(const GraphDev& g, const StreamArgs& a, const RM& rm, int64_t c) {
        done = false; folded = 0; iv_pos = ~0ull; ic = 0; row_extras = nullptr; nout = 0;
        const StreamEntry e = a.entries[c], en = a.entries[c + 1];
        const uint64_t pos = a.bit0 + (uint64_t)c * STREAM_CHUNK_BITS + e.dpos;
        end = a.bit0 + (uint64_t)(c + 1) * STREAM_CHUNK_BITS + en.dpos;
        r.x = e.x; r.d = 0; r.rem = 0; r.v = e.v; r.first = false; r.fold = false; r.store = false; r.is_long = false; r.dst = nullptr;
        f.begin(e.x);
        if (pos >= end || e.x >= g.node_hi) { done = true; return; }
        b.seek(g, pos);
        left = (int64_t)(end - pos);
        const uint32_t rem = e.rem & ~STREAM_FIRST;
        if (rem) {
            const uint32_t d = (uint32_t)g.outdeg[e.x - g.node_lo];
            locate(g, a, rm, e.x, d, rem, false, -1);
            r.first = (e.rem & STREAM_FIRST) != 0;
            r.v = e.v;
            if (!r.fold && !r.store) skip_run(g);   // nobody wants these residuals: jump to the next record
        }
    }

    // Jumps over the rest of the current record (its residual run ends where the next record starts).
    __device__ __forceinline__ void skip_run(const GraphDev& g) {
        const uint64_t next = g.offsets[r.x + 1 - g.node_lo] - g.bit_base;
        r.rem = 0;
        r.x++;
        if (next >= end || r.x >= g.node_hi) { done = true; return; }
        b.seek(g, next);
        left = (int64_t)(end - next);
    }

    // One residual (BVGraph.java:954, 966).  Precondition: r.rem > 0.
    __device__ __forceinline__ void resid_step(const GraphDev& g) {
        const int k = g.c.zetak;
        uint32_t m, len;
        if (r.first) {
            r.v = (uint32_t)(int32_t)((int64_t)r.x + nat2int(zeta_any<K>(b, g, k) - 1ull));
            r.first = false;
            sync_left(g);
        } else {
            if (zeta_fast<K>(b.top(), k, m, len)) { b.skip(len); left -= (int64_t)len; }
            else { m = (uint32_t)zeta_any<K>(b, g, k); sync_left(g); }
            r.v += m;
        }
        if (r.fold) { f.add(r.v); folded++; }
        if (r.store) r.dst[-(int64_t)r.rem] = (int32_t)r.v;
        r.rem--;
    }

    // The residual run of the current record has just ended (or the record had none): finish the record.
    __device__ __forceinline__ void end_record(const GraphDev& g) {
        if (r.store && !r.is_long && iv_pos != ~0ull && ic) iv_merge_row(g);
        flush_fold();
        iv_pos = ~0ull; ic = 0;
        r.x++;
        r.store = false; r.fold = false;
        if (b.overrun()) { fail(g, E_IO); return; }
        if (left <= 0 || r.x >= g.node_hi) done = true;
    }

    __device__ __forceinline__ void iv_merge_row(const GraphDev& g) { stream_iv_merge<K>(g, r.x, nout, ic, iv_pos, row_extras); }

    // Record prologue at a record start (BVGraph.java:1048-1100): outdegree, reference, copy blocks (for the copied count),
    // intervals (folded here), and the state of the residual run.  Precondition: r.rem == 0, !done, the window at the first
    // bit of node r.x's record.
    __device__ __forceinline__ void prologue(const GraphDev& g, const StreamArgs& a, const RM& rm) {
        const Codec& c = g.c;
        const int32_t x = r.x;
        if (x >= a.hi) { done = true; return; }
        if (x < a.lo) {  // before the range: jump to its first record
            const uint64_t p = g.offsets[a.lo - g.node_lo] - g.bit_base;
            r.x = a.lo;
            if (p >= end) { done = true; return; }
            b.seek(g, p);
            left = (int64_t)(end - p);
            return;
        }
        const uint64_t d64 = b.gamma(g);
        if (d64 > 0x7fffffffull || b.overrun()) { fail(g, E_IO); return; }
        const uint32_t d = (uint32_t)d64;
        if (d == 0) {  // a one-bit record
            r.x++;
            left -= 1;
            if (left <= 0 || r.x >= g.node_hi) done = true;
            return;
        }
        int32_t l = -1;
        if ((int32_t)d > a.long_d && a.nlong > 0) l = long_find(a.long_nodes, a.nlong, x);
        if (l >= 0) {
            // long record: its header, copy blocks and intervals are in the long index; only its residual run is decoded here
            const LongMeta& m = a.li.meta[l];
            locate(g, a, rm, x, d, (uint32_t)m.rc, true, l);
            r.first = true;
            if (m.rc <= 0 || (!r.fold && !r.store)) { skip_run(g); return; }
            if (m.resid_pos >= end) { flush_fold(); r.rem = 0; done = true; return; }   // the run starts in a later chunk (its entry points there)
            b.seek(g, m.resid_pos);
            left = (int64_t)(end - m.resid_pos);
            return;
        }
        uint32_t copied = 0;
        if (c.window > 0) {
            const uint64_t ref = b.unary(g);
            if (ref > (uint64_t)c.window) { fail(g, E_STATE); return; }      // BVGraph.java:705
            if ((int64_t)ref > (int64_t)x - g.node_lo) { fail(g, E_FORMAT); return; }
            if (ref) {
                const uint64_t bc = b.gamma(g);
                int64_t total = 0, cp = 0;
                bool ok = bc <= 0x7fffffffull;
                for (uint64_t i = 0; ok && i < bc; i++) {  // :1062-1066
                    const int64_t blk = (int64_t)b.gamma(g) + (i ? 1 : 0);
                    total += blk;
                    if (!(i & 1)) cp += blk;
                    if (b.overrun()) ok = false;
                }
                const int64_t dp = g.outdeg[x - (int32_t)ref - g.node_lo];
                if (ok && !(bc & 1)) cp += dp - total;  // :1069
                if (!ok || total > dp || cp < 0 || cp > (int64_t)d) { fail(g, ok ? E_FORMAT : E_IO); return; }
                copied = (uint32_t)cp;
            }
        }
        locate(g, a, rm, x, d, 0, true, -1);
        if (!r.fold && !r.store) { skip_run(g); return; }
        if (r.fold) arcs += d;
        nout = d - copied;
        row_extras = r.store ? r.dst - nout : nullptr;
        uint32_t rc = nout;
        iv_pos = ~0ull; ic = 0;
        if (nout > 0 && c.minlen != 0) {  // interval section (:1076-1095), elements folded as they are walked
            const uint64_t ic64 = b.gamma(g);
            if (ic64 > (uint64_t)nout) { fail(g, E_IO); return; }
            ic = (uint32_t)ic64;
            iv_pos = b.pos(g);
            int64_t total = 0;
            uint32_t prev = 0;
            for (uint32_t i = 0; i < ic; i++) {
                uint32_t left;
                if (i == 0) left = (uint32_t)(int32_t)(nat2int(b.gamma(g)) + (int64_t)x);
                else left = prev + 1u + (uint32_t)b.gamma(g);
                const uint64_t len64 = b.gamma(g) + (uint64_t)c.minlen;
                total += (int64_t)len64;
                if (total > (int64_t)nout || b.overrun()) { fail(g, E_IO); return; }
                const uint32_t len = (uint32_t)len64;
                if (r.fold) { for (uint32_t j = 0; j < len; j++) f.add(left + j); folded += len; }
                prev = left + len;
            }
            rc = nout - (uint32_t)total;
        }
        r.rem = rc;
        r.first = true;
        sync_left(g);
        if (rc == 0) { end_record(g); return; }
        if (left <= 0) {  // the residual run starts in the next chunk: the record is finished by the lane that gets there
            flush_fold();
            r.rem = 0; done = true;
        }
    }

    // A deferred record (k_stream_ivfix): its interval section is found again from the record's start, then merged.
    __device__ __forceinline__ void fix_intervals(const GraphDev& g, const StreamArgs& a, const RM& rm, int32_t x) {
        r.x = x;
        r.dst = rm.row(g, x) + g.outdeg[x - g.node_lo];
        recover_intervals(g);
        if (ic) iv_merge_row(g);
    }
    __device__ __forceinline__ void recover_intervals(const GraphDev& g) {
        const Codec& c = g.c;
        Win h;
        h.seek(g, g.offsets[r.x - g.node_lo] - g.bit_base);
        const uint32_t d = (uint32_t)h.gamma(g);
        uint32_t copied = 0;
        if (c.window > 0) {
            const uint64_t ref = h.unary(g);
            if (ref) {
                const uint64_t bc = h.gamma(g);
                int64_t total = 0, cp = 0;
                for (uint64_t i = 0; i < bc && !h.overrun(); i++) {
                    const int64_t blk = (int64_t)h.gamma(g) + (i ? 1 : 0);
                    total += blk;
                    if (!(i & 1)) cp += blk;
                }
                if (!(bc & 1)) cp += (int64_t)g.outdeg[r.x - (int32_t)ref - g.node_lo] - total;
                copied = (uint32_t)(cp < 0 ? 0 : (cp > (int64_t)d ? (int64_t)d : cp));
            }
        }
        nout = d - copied;
        row_extras = r.dst - nout;
        ic = 0; iv_pos = ~0ull;
        if (nout > 0 && c.minlen != 0) {
            ic = (uint32_t)h.gamma(g);
            iv_pos = h.pos(g);
        }
    }

    // One trip of the flat loop for this lane: a residual step when it has residuals left, else a prologue.  The kernel
    // decides, per warp, which of the two everybody takes; the emulation calls this until done.
    __device__ __forceinline__ void step_resid(const GraphDev& g, const StreamArgs& a) {
        resid_step(g);
        if (r.rem == 0) {
            if (r.store && !r.is_long && iv_pos == ~0ull && g.c.minlen != 0) {
#ifdef BVG_HOST_EMULATION
                a.defer_list[(*a.defer_count)++] = r.x;
#else
                a.defer_list[atomicAdd(a.defer_count, 1u)] = r.x;
#endif
            }
            end_record(g);
        } else if (left <= 0) {   // the run goes on in the next chunk
            flush_fold();
            done = true;
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// Entries, built once at open: one thread per chunk.
// ---------------------------------------------------------------------------------------------------
template <int K>
__device__ inline void stream_entry_one(const GraphDev& g, int64_t c, int64_t nchunks, uint64_t bit0, int32_t node0, const int32_t* __restrict__ long_nodes,
                                        int32_t nlong, int32_t long_d, const LongIndex& li, StreamEntry* __restrict__ entries) {
    const int64_t n = (int64_t)g.node_hi - g.node_lo;
    const uint64_t last = g.offsets[n] - g.bit_base;
    StreamEntry e;
    e.dpos = 0; e.x = g.node_hi; e.rem = 0; e.v = 0;
    if (c >= nchunks) {  // sentinel: the end of the stream
        const uint64_t start = bit0 + (uint64_t)nchunks * STREAM_CHUNK_BITS;
        e.dpos = last >= start ? (uint32_t)(last - start) : 0u;
        entries[nchunks] = e;
        return;
    }
    const uint64_t p = bit0 + (uint64_t)c * STREAM_CHUNK_BITS;
    auto put = [&](uint64_t pos, int32_t x, uint32_t rem, uint32_t v) {
        if (pos > last) pos = last;
        e.dpos = (uint32_t)(pos - p); e.x = x; e.rem = rem; e.v = v;
        entries[c] = e;
    };
    if (p >= last) { put(last, g.node_hi, 0, 0); return; }
    // record containing p: largest i with offsets[i] <= p  (i >= node0 - node_lo since p >= bit0 = offsets[node0])
    int64_t lo = (int64_t)node0 - g.node_lo, hi = n;  // invariant: offsets[lo] <= p < offsets[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (g.offsets[mid] - g.bit_base <= p) lo = mid; else hi = mid;
    }
    const int64_t i = lo;
    const int32_t x = g.node_lo + (int32_t)i;
    const uint64_t o = g.offsets[i] - g.bit_base, onext = g.offsets[i + 1] - g.bit_base;
    if (o == p) { put(p, x, 0, 0); return; }
    // p is inside record x: where does its residual run start, how long is it
    const Codec& cd = g.c;
    Win b;
    uint64_t rpos;
    int64_t rc;
    int32_t l = -1;
    const int32_t d_idx = g.outdeg[i];
    if (d_idx > long_d && nlong > 0) l = long_find(long_nodes, nlong, x);
    if (l >= 0) {
        const LongMeta& m = li.meta[l];
        rpos = m.resid_pos; rc = m.rc;
    } else {
        b.seek(g, o);
        const uint64_t d = b.gamma(g);
        int64_t copied = 0;
        if (d != 0 && cd.window > 0) {
            const uint64_t ref = b.unary(g);
            if (ref && ref <= (uint64_t)i) {
                const uint64_t bc = b.gamma(g);
                int64_t total = 0, cp = 0;
                for (uint64_t k = 0; k < bc && !b.overrun(); k++) {
                    const int64_t blk = (int64_t)b.gamma(g) + (k ? 1 : 0);
                    total += blk;
                    if (!(k & 1)) cp += blk;
                }
                if (!(bc & 1)) cp += (int64_t)g.outdeg[i - (int64_t)ref] - total;
                copied = cp < 0 ? 0 : (cp > (int64_t)d ? (int64_t)d : cp);
            }
        }
        int64_t extra = (int64_t)d - copied;
        if (extra > 0 && cd.minlen != 0) {
            const int64_t ic = (int64_t)b.gamma(g);
            for (int64_t k = 0; k < ic && !b.overrun(); k++) {
                (void)b.gamma(g);
                extra -= (int64_t)b.gamma(g) + cd.minlen;
            }
        }
        rc = extra > 0 ? extra : 0;
        rpos = b.pos(g);
    }
    if (rc <= 0) { put(onext, x + 1, 0, 0); return; }           // no residuals: the next record start
    if (p <= rpos) { put(rpos, x, (uint32_t)rc | STREAM_FIRST, 0); return; }   // the whole run belongs to this chunk's lane
    // p is inside the run: walk to the first code boundary at or after p, from the run's start or, for a long record, from
    // the last sync point at or before p
    int64_t ord = 0;
    uint32_t v = 0;
    uint64_t start = rpos;
    if (l >= 0) {
        const LongMeta& m = li.meta[l];
        const int64_t nseg = (m.rc + li.seg - 1) / li.seg;
        int64_t sa = 0, sb = nseg;  // largest s with seg_pos[s] <= p
        while (sb - sa > 1) { const int64_t mid = (sa + sb) >> 1; if (li.seg_pos[m.seg_off + mid] <= p) sa = mid; else sb = mid; }
        ord = sa * li.seg;
        start = li.seg_pos[m.seg_off + sa];
        v = (uint32_t)li.seg_val[m.seg_off + sa];
    }
    b.seek(g, start);
    const int k = cd.zetak;
    while (ord < rc && b.pos(g) < p) {
        if (ord == 0) v = (uint32_t)(int32_t)((int64_t)x + nat2int(zeta_any<K>(b, g, k) - 1ull));
        else v += (uint32_t)zeta_any<K>(b, g, k);
        ord++;
        if (b.overrun()) break;
    }
    if (ord >= rc) { put(onext, x + 1, 0, 0); return; }
    put(b.pos(g), x, (uint32_t)(rc - ord) | (ord == 0 ? STREAM_FIRST : 0u), v);
}

#ifndef BVG_HOST_EMULATION
template <int K>
__global__ void k_stream_entries(GraphDev g, int64_t nchunks, uint64_t bit0, int32_t node0, const int32_t* __restrict__ long_nodes, int32_t nlong,
                                 int32_t long_d, LongIndex li, StreamEntry* __restrict__ entries) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= nchunks) stream_entry_one<K>(g, c, nchunks, bit0, node0, long_nodes, nlong, long_d, li, entries);
}

constexpr int STREAM_BLOCK = 128;
#ifndef BVG_STREAM_THRESH
#define BVG_STREAM_THRESH 20
#endif
#ifndef BVG_STREAM_GROUP
#define BVG_STREAM_GROUP 8      // chunks per lane: a warp owns 32 * BVG_STREAM_GROUP consecutive chunks and hands them to whichever lane is free
#endif
#ifndef BVG_STREAM_REFILL
#define BVG_STREAM_REFILL 6     // lanes that must be out of work before the warp stops to open new chunks
#endif

// A warp owns 32 * BVG_STREAM_GROUP consecutive chunks; a lane that has finished its chunk takes the next one.  A residual
// step is taken while at least BVG_STREAM_THRESH lanes have residuals left; below that, lanes out of work open new chunks
// (once enough of them wait) and lanes at a record start take their prologues together.
template <int K, class RM>
__global__ void __launch_bounds__(STREAM_BLOCK, 8) k_stream_extras(GraphDev g, StreamArgs a, RM rm) {
    const unsigned lane = threadIdx.x & 31u;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t next = warp * (32 * BVG_STREAM_GROUP);
    const int64_t last = next + 32 * BVG_STREAM_GROUP < a.count ? next + 32 * BVG_STREAM_GROUP : a.count;
    StreamLane<K, RM> L;
    L.init();
    for (;;) {
        const bool in_run = !L.done && L.r.rem > 0;
        const bool at_start = !L.done && L.r.rem == 0;
        const unsigned m_run = __ballot_sync(0xffffffffu, in_run);
        const unsigned m_start = __ballot_sync(0xffffffffu, at_start);
        const unsigned m_idle = ~(m_run | m_start);
        const bool more = next < last;
        if (m_run != 0u && (__popc(m_run) >= BVG_STREAM_THRESH || (m_start == 0u && !(more && m_idle != 0u)))) {
            if (in_run) L.step_resid(g, a);
        } else if (more && m_idle != 0u && (__popc(m_idle) >= BVG_STREAM_REFILL || m_start == 0u)) {
            const int64_t c = next + __popc(m_idle & ((1u << lane) - 1u));
            if (L.done && c < last) L.open(g, a, rm, a.first_chunk + c);
            next += __popc(m_idle);
        } else if (m_start != 0u) {
            if (at_start) L.prologue(g, a, rm);
        } else break;
        __syncwarp();
    }
    if (a.result) warp_fold(L.acc, L.arcs, a.result);
}

template <int K, class RM>
__global__ void k_stream_ivfix(GraphDev g, StreamArgs a, RM rm) {
    const unsigned int n = *a.defer_count;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        StreamLane<K, RM> L;
        L.init();
        L.fix_intervals(g, a, rm, a.defer_list[i]);
    }
}
#endif

}  // namespace bvg
