// bvg_long.cuh -- long records, split across threads.
//
// One thread per record cannot work for the heavy tail of a power-law graph: the 858 018-successor node of the 1 B-arc
// benchmark graph is ~10^6 dependent code reads on a single thread.  Codeword k+1 starts where codeword k ends, so a
// record can only be entered at positions somebody has already walked to.  The reference has the same problem and
// solves it the same way for whole records: it keeps an index of record starts (.offsets / Elias-Fano,
// BVGraph.java:1594).  For records with more than LONG_D successors this file extends that idea inside the record:
// at open, one thread walks the record once and leaves
//   * its copy blocks as prefix sums (parent position and copied count at the start of every copy block),
//   * its intervals as (left, cumulative length),
//   * a sync point (bit position, running successor value) every LONG_SEG residuals.
// Per scan the work is then data-parallel with bounded items:
//   k_long_resid   one thread per residual segment  : zeta_k codes -> successor values            (BVGraph.java:939-972)
//   k_long_extras  one thread per output chunk      : intervals U residuals by merge path          (:1103-1108, MergedIntIterator)
//   k_long_merge   one thread per output chunk      : masked parent list U extras by merge path    (:1110-1126, MaskedIntIterator)
// Files whose lists contain duplicated successors (never produced by BVGraph.store, BVGraph.java:2201) keep the
// sequential kernels' -1 fill semantics only on the short path; the merge path assumes the three parts are disjoint.
#pragma once
#include "bvg_device.cuh"
#include "bvg_scan.cuh"

namespace bvg {

#ifndef BVG_LONG_D        // overridable so that tests/hostemu can push every record through the split path
#define BVG_LONG_D 1024     // measured on the 1 B-arc benchmark graph: 2048/512/512 -> 10.98 ms per scan, 1024/128/128 -> 9.74 ms
#define BVG_LONG_SEG 128
#define BVG_LONG_CHUNK 128
#endif
constexpr int32_t LONG_D = BVG_LONG_D;          // records with more successors than this take the split path
constexpr int32_t LONG_SEG = BVG_LONG_SEG;      // residuals per sync point
constexpr int32_t LONG_CHUNK = BVG_LONG_CHUNK;  // outputs per merge-path item

struct LongMeta {
    int32_t x, d, ref, copied;
    int32_t ncb;          // copy blocks, including the implicit tail block of an even block count
    int32_t ic, ilen;     // intervals and their total length
    int32_t rc;           // residuals
    int32_t level;        // reference-chain depth
    int32_t flags;        // bit 0: somebody copies from this record (its list has to exist during a consume-only scan)
    int64_t cb_off;       // cb_cum[cb_off .. +ncb] (ncb+1 entries), cb_ppos[cb_off - l .. ] (ncb entries): see LongIndex
    int64_t iv_off;       // iv_cum[iv_off .. +ic] (ic+1 entries), iv_left (ic entries)
    int64_t seg_off;      // seg_pos / seg_val, ceil(rc / LONG_SEG) entries
    int64_t tmp_off;      // per-scan temp: residuals at [tmp_off, +rc), extras at [tmp_off + d, +d-copied)
    int64_t scan_off;     // tile scan (bvg_tile.cuh), records somebody copies from only: 3 d entries -- residuals, extras, the list
    uint64_t after_header;  // bit position (relative to word 0) right after outdegree + reference
    uint64_t resid_pos;     // bit position of the first residual code
    uint64_t rec_end;       // bit position one past the record (= end of the residual section)
};

struct LongIndex {
    const LongMeta* __restrict__ meta;
    const int32_t* __restrict__ cb_cum;    // copied count before copy block t (per node: ncb+1 entries, last = copied)
    const int32_t* __restrict__ cb_ppos;   // parent position of copy block t   (allocated with the same stride as cb_cum)
    const int32_t* __restrict__ iv_cum;    // interval elements before interval t (ic+1 entries, last = ilen)
    const int32_t* __restrict__ iv_left;   // left extreme of interval t          (same stride as iv_cum)
    const uint64_t* __restrict__ seg_pos;  // bit position of the first code of residual segment s
    const int64_t* __restrict__ seg_val;   // successor value just before it (unused for s = 0)
    int32_t seg, chunk;                    // residuals per sync point, outputs per merge-path item (LONG_SEG, LONG_CHUNK by default)
};

// Largest t in [0, n) with cum[t] <= q, for a non-decreasing cum[0..n] with cum[0] = 0 <= q < cum[n].
__device__ __forceinline__ int32_t upper_slot(const int32_t* __restrict__ cum, int32_t n, int32_t q) {
    int32_t lo = 0, hi = n;  // invariant: cum[lo] <= q < cum[hi]
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (cum[mid] <= q) lo = mid; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------------
// Open-time walk of one long record.  pass 0 counts (ncb, ic, ilen, copied, rc); pass 1 fills the arrays.
// ---------------------------------------------------------------------------------------------------
template <bool DEF>
__device__ void long_walk(const GraphDev& g, LongMeta& m, int pass, int32_t* cb_cum, int32_t* cb_ppos,
                          int32_t* iv_cum, int32_t* iv_left, uint64_t* seg_pos, int64_t* seg_val, int32_t seg = LONG_SEG) {
    const Codec& c = g.c;
    const int32_t x = m.x;
    BitBuf b = buffer_at(g, x);
    const int64_t d = (int64_t)Rd<DEF>::outdeg(b, c);
    int32_t r = 0;
    if (c.window > 0) r = (int32_t)Rd<DEF>::ref(b, c);
    m.d = (int32_t)d;
    m.ref = r;
    m.after_header = b.pos();
    int64_t copied = 0;
    int32_t ncb = 0;
    if (r > 0) {
        const int64_t bc = (int64_t)Rd<DEF>::bcount(b, c);
        const int64_t dp = g.outdeg[x - r - g.node_lo];
        int64_t p = 0;
        for (int64_t i = 0; i < bc; i++) {
            const int64_t blk = (int64_t)Rd<DEF>::block(b, c) + (i ? 1 : 0);
            if (!(i & 1)) {
                if (pass) { cb_cum[ncb] = (int32_t)copied; cb_ppos[ncb] = (int32_t)p; }
                ncb++;
                copied += blk;
            }
            p += blk;
        }
        if (!(bc & 1)) {  // implicit tail block (MaskedIntIterator: left = -1)
            if (pass) { cb_cum[ncb] = (int32_t)copied; cb_ppos[ncb] = (int32_t)p; }
            ncb++;
            copied += dp - p;
        }
        if (pass) cb_cum[ncb] = (int32_t)copied;
    }
    m.copied = (int32_t)copied;
    m.ncb = ncb;
    int64_t extra = d - copied;
    int64_t ic = 0, ilen = 0;
    if (extra > 0 && c.minlen != 0) {
        ic = (int64_t)b.gamma();
        int64_t prev = 0;
        for (int64_t i = 0; i < ic; i++) {
            int64_t left;
            if (i == 0) left = (int64_t)(int32_t)(nat2int(b.gamma()) + (int64_t)x);
            else left = (int64_t)b.gamma() + prev + 1;
            const int64_t len = (int64_t)b.gamma() + c.minlen;
            if (pass) { iv_cum[i] = (int32_t)ilen; iv_left[i] = (int32_t)left; }
            prev = left + len;
            ilen += len;
        }
        if (pass) iv_cum[ic] = (int32_t)ilen;
    }
    m.ic = (int32_t)ic;
    m.ilen = (int32_t)ilen;
    const int64_t rc = extra - ilen;
    m.rc = (int32_t)(rc > 0 ? rc : 0);
    m.resid_pos = b.pos();
    if (!pass || rc <= 0 || !seg_pos) return;
    // residual sync points (the one sequential pass over this record's residuals, paid once at open)
    int64_t v = 0;
    for (int64_t i = 0; i < rc; i++) {
        if (i % seg == 0) { seg_pos[i / seg] = b.pos(); seg_val[i / seg] = v; }
        if (i == 0) v = (int64_t)(int32_t)((int64_t)x + nat2int(Rd<DEF>::resid(b, c)));
        else v += (int64_t)Rd<DEF>::resid(b, c) + 1;
    }
}

// ---------------------------------------------------------------------------------------------------
// Sync points without a sequential pass.  Walking the residuals of every long record on one thread each costs
// ~10^6 dependent code reads for the largest record of the benchmark graph (120 ms at open).  The residual section of a
// long record is a run of zeta_k codes between two known bit positions, so it is cut into LSPEC_BITS sub-ranges that
// are entered speculatively and proven exactly as bvg_offsets.cuh does for the .offsets stream:
// speculate (count, sum of gap+1, exit) -> fix against the previous sub-range's exit until nothing moves ->
// per-record scan -> emit a sync point wherever the ordinal is a multiple of LONG_SEG.
// ---------------------------------------------------------------------------------------------------
#ifndef BVG_LSPEC_BITS
#define BVG_LSPEC_BITS 4096
#endif
constexpr int64_t LSPEC_BITS = BVG_LSPEC_BITS;

struct SpecItem {
    uint64_t lo, hi;      // bit range of the sub-range (hi of a record's last item = end of the record)
    uint64_t entry, exit; // proven entry / exit of the chain
    int64_t count;        // codes on the chain inside the sub-range
    int64_t sum;          // sum of (gap + 1) of those codes, the record's very first code excluded
    int32_t l;            // long record
    int32_t first;        // 1 = first sub-range of its record: entry is exact, first code is the long residual
};

template <bool DEF>
__device__ inline void lspec_walk(const GraphDev& g, uint64_t pos, uint64_t hi, bool skip_first, uint64_t& exit, int64_t& count, int64_t& sum) {
    BitBuf b;
    b.w = g.words; b.maxw = g.nwords - 3;
    b.seek(pos);
    count = 0; sum = 0;
    while (b.pos() < hi) {
        const uint64_t v = Rd<DEF>::resid(b, g.c);
        if (!(skip_first && count == 0)) sum += (int64_t)v + 1;
        count++;
    }
    exit = b.pos();
}

template <bool DEF>
__device__ inline void lspec_speculate_one(const GraphDev& g, SpecItem& it) {
    it.entry = it.lo;
    lspec_walk<DEF>(g, it.lo, it.hi, it.first != 0, it.exit, it.count, it.sum);
}

template <bool DEF>
__device__ inline void lspec_fix_one(const GraphDev& g, int64_t j, const SpecItem* __restrict__ in, SpecItem* __restrict__ out, int* changed) {
    SpecItem s = in[j];
    if (!s.first) {
        const uint64_t t = in[j - 1].exit;
        if (t != s.entry) {
            BitBuf a, b;
            a.w = b.w = g.words; a.maxw = b.maxw = g.nwords - 3;
            a.seek(t); b.seek(s.entry);
            int64_t ca = 0, cb = 0, sa = 0, sb = 0;
            while (a.pos() != b.pos() && a.pos() < s.hi && b.pos() < s.hi) {
                if (a.pos() < b.pos()) { sa += (int64_t)Rd<DEF>::resid(a, g.c) + 1; ca++; }
                else { sb += (int64_t)Rd<DEF>::resid(b, g.c) + 1; cb++; }
            }
            if (a.pos() == b.pos()) { s.count += ca - cb; s.sum += sa - sb; }
            else {
                while (a.pos() < s.hi) { sa += (int64_t)Rd<DEF>::resid(a, g.c) + 1; ca++; }
                s.count = ca; s.sum = sa;
                if (a.pos() != s.exit) { s.exit = a.pos(); *changed = 1; }
            }
            s.entry = t;
        }
    }
    out[j] = s;
}

// cbase: residual ordinal of the item's first code; sbase: sum of (gap+1) of the record's codes before it.
template <bool DEF>
__device__ inline void lspec_emit_one(const GraphDev& g, const SpecItem& it, const LongMeta& m, int64_t cbase, int64_t sbase,
                                      const int64_t* __restrict__ v0, int64_t* __restrict__ v0_out,
                                      uint64_t* __restrict__ seg_pos, int64_t* __restrict__ seg_val, int32_t seg = LONG_SEG) {
    BitBuf b;
    b.w = g.words; b.maxw = g.nwords - 3;
    b.seek(it.entry);
    int64_t ord = cbase;
    int64_t v = it.first ? 0 : v0[it.l] + sbase;
    while (b.pos() < it.hi && ord < m.rc) {
        if (ord % seg == 0) { seg_pos[m.seg_off + ord / seg] = b.pos(); seg_val[m.seg_off + ord / seg] = v; }
        const uint64_t code = Rd<DEF>::resid(b, g.c);
        if (ord == 0) { v = (int64_t)(int32_t)((int64_t)m.x + nat2int(code)); if (v0_out) v0_out[it.l] = v; }
        else v += (int64_t)code + 1;
        ord++;
    }
}

template <bool DEF>
__global__ void k_lspec_speculate(GraphDev g, SpecItem* __restrict__ items, int64_t n) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) { SpecItem it = items[j]; lspec_speculate_one<DEF>(g, it); items[j] = it; }
}

template <bool DEF>
__global__ void k_lspec_fix(GraphDev g, const SpecItem* __restrict__ in, SpecItem* __restrict__ out, int64_t n, int* __restrict__ changed) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) lspec_fix_one<DEF>(g, j, in, out, changed);
}

// The first residual of every record (its value anchors all the others): read by the record's first item.
template <bool DEF>
__global__ void k_lspec_first(GraphDev g, const LongMeta* __restrict__ meta, int32_t nlong, int64_t* __restrict__ v0) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlong) return;
    const LongMeta m = meta[l];
    if (m.rc <= 0) { v0[l] = 0; return; }
    BitBuf b;
    b.w = g.words; b.maxw = g.nwords - 3;
    b.seek(m.resid_pos);
    v0[l] = (int64_t)(int32_t)((int64_t)m.x + nat2int(Rd<DEF>::resid(b, g.c)));
}

template <bool DEF>
__global__ void k_lspec_emit(GraphDev g, const SpecItem* __restrict__ items, int64_t n, const LongMeta* __restrict__ meta,
                             const int64_t* __restrict__ cbase, const int64_t* __restrict__ sbase, const int64_t* __restrict__ v0,
                             uint64_t* __restrict__ seg_pos, int64_t* __restrict__ seg_val, int32_t seg) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const SpecItem it = items[j];
    lspec_emit_one<DEF>(g, it, meta[it.l], cbase[j], sbase[j], v0, nullptr, seg_pos, seg_val, seg);
}

// ---------------------------------------------------------------------------------------------------
// Per-scan items
// ---------------------------------------------------------------------------------------------------

// Residual segment s of long record m: LONG_SEG (or fewer) successors into dst[s * LONG_SEG ..].
template <bool DEF>
__device__ void long_resid_segment(const GraphDev& g, const LongMeta& m, const LongIndex& li, int32_t s, int32_t* __restrict__ dst) {
    const int32_t first = s * li.seg;
    const int32_t cnt = min(li.seg, m.rc - first);
    BitBuf b;
    b.w = g.words;
    b.maxw = g.nwords - 3;
    b.seek(li.seg_pos[m.seg_off + s]);
    int64_t v = li.seg_val[m.seg_off + s];
    dst += first;
    for (int32_t i = 0; i < cnt; i++) {
        if (first + i == 0) v = (int64_t)(int32_t)((int64_t)m.x + nat2int(Rd<DEF>::resid(b, g.c)));
        else v += (int64_t)Rd<DEF>::resid(b, g.c) + 1;
        dst[i] = (int32_t)v;
    }
}

// Virtual interval sequence: element q of the concatenated intervals.
struct IntervalSeq {
    const int32_t* __restrict__ cum;
    const int32_t* __restrict__ left;
    int32_t n, len;
    __device__ __forceinline__ int32_t at(int32_t q) const {
        const int32_t t = upper_slot(cum, n, q);
        return left[t] + (q - cum[t]);
    }
};

// Virtual copied sequence: element q of the parent's list seen through the copy blocks.
struct CopiedSeq {
    const int32_t* __restrict__ cum;
    const int32_t* __restrict__ ppos;
    const int32_t* __restrict__ parent;
    int32_t n, len;
    __device__ __forceinline__ int32_t at(int32_t q) const {
        const int32_t t = upper_slot(cum, n, q);
        return parent[ppos[t] + (q - cum[t])];
    }
};

// Merge path: how many of the first q0 outputs of merge(A, B) come from A (A, B ascending and disjoint).
template <class A>
__device__ __forceinline__ int32_t merge_path(const A& a, const int32_t* __restrict__ bv, int32_t blen, int32_t q0) {
    int32_t lo = max(0, q0 - blen), hi = min(q0, a.len);
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (a.at(mid) < bv[q0 - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Outputs [q0, q0 + cnt) of merge(A, B) into out, A being a virtual block-structured sequence walked by (slot, offset).
template <class A, class ValueOf, class F = Fold32>
__device__ void merge_chunk(const A& a, ValueOf value_of, const int32_t* __restrict__ bv, int32_t blen,
                            int32_t q0, int32_t cnt, int32_t* __restrict__ out, F* fold = nullptr) {
    int32_t i = merge_path(a, bv, blen, q0), j = q0 - i;
    int32_t t = 0, t_end = 0;  // current slot of A and the index where it ends
    if (i < a.len) { t = upper_slot(a.cum, a.n, i); t_end = a.cum[t + 1]; }
    for (int32_t k = 0; k < cnt; k++) {
        const bool has_a = i < a.len, has_b = j < blen;
        int32_t av = 0;
        if (has_a) {
            while (i >= t_end) { t++; t_end = a.cum[t + 1]; }  // skips empty slots
            av = value_of(t, i - a.cum[t]);
        }
        int32_t o;
        if (has_a && (!has_b || av < bv[j])) { o = av; i++; }
        else { o = bv[j]; j++; }
        out[q0 + k] = o;
        if (fold) fold->add((uint32_t)o);  // a consume-only scan folds a stored row where its final values are produced
    }
}

// ---------------------------------------------------------------------------------------------------
// Kernels.  Items are (long record, segment | chunk) pairs listed at open.
// ---------------------------------------------------------------------------------------------------
// Work items of every per-scan kernel are (long record, part) pairs.  They are not listed: item i belongs to the record l
// with cum[l] <= i < cum[l + 1] (cum = running item count per record, nlong + 1 entries built at open), found by
// binary search; a list would cost one host-side push per 128 successors of every long record at every open.
// hint (may be null): hint[j] = record of item j << ITEM_HINT_SHIFT, one more entry than needed past the last item: the
// search then starts in the handful of records those 64 items span instead of in all of them (20 dependent steps for 10^6
// long records, most of them the same for every thread but each an L1 / L2 round trip).
constexpr int ITEM_HINT_SHIFT = 6;
struct ItemMap {
    const int64_t* __restrict__ cum;
    int32_t nlong;
    const int32_t* __restrict__ hint;
    __device__ __forceinline__ void find(int64_t i, int32_t& l, int32_t& part) const {
        int32_t lo = 0, hi = nlong;  // invariant: cum[lo] <= i < cum[hi]
        if (hint) {
            lo = hint[i >> ITEM_HINT_SHIFT];
            const int32_t h = hint[(i >> ITEM_HINT_SHIFT) + 1] + 1;
            hi = h < nlong ? h : nlong;
        }
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (cum[mid] <= i) lo = mid; else hi = mid;
        }
        l = lo; part = (int32_t)(i - cum[lo]);
    }
};

#ifndef BVG_HOST_EMULATION
// hint[j] for j = 0 .. count - 1 (count = (total >> ITEM_HINT_SHIFT) + 2); items past the end map to the last record.
__global__ void k_item_hints(const int64_t* __restrict__ cum, int32_t nlong, int64_t total, int32_t* __restrict__ hint, int64_t count) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const int64_t i = j << ITEM_HINT_SHIFT;
    if (i >= total) { hint[j] = nlong > 0 ? nlong - 1 : 0; return; }
    ItemMap im{ cum, nlong, nullptr };
    int32_t l, part;
    im.find(i, l, part);
    hint[j] = l;
}
#endif

template <bool DEF>
__global__ void k_long_count(GraphDev g, const int32_t* __restrict__ long_nodes, int32_t nlong, const uint8_t* __restrict__ is_parent,
                             LongMeta* __restrict__ meta) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlong) return;
    LongMeta m{};   // every byte defined: the array is copied to the host (offsets are laid out there)
    m.x = long_nodes[l];
    m.level = g.depth[m.x - g.node_lo];
    m.flags = (is_parent && is_parent[m.x - g.node_lo]) ? 1 : 0;
    m.rec_end = g.offsets[m.x - g.node_lo + 1] - g.bit_base;
    long_walk<DEF>(g, m, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    m.cb_off = m.iv_off = m.seg_off = m.tmp_off = m.scan_off = 0;
    meta[l] = m;
}

template <bool DEF>
__global__ void k_long_fill(GraphDev g, int32_t nlong, LongMeta* __restrict__ meta, int32_t* cb_cum, int32_t* cb_ppos,
                            int32_t* iv_cum, int32_t* iv_left, uint64_t* seg_pos, int64_t* seg_val) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlong) return;
    LongMeta m = meta[l];
    long_walk<DEF>(g, m, 1, cb_cum + m.cb_off, cb_ppos + m.cb_off, iv_cum + m.iv_off, iv_left + m.iv_off,
                   seg_pos + m.seg_off, seg_val + m.seg_off);
}

// Where the parts of a long record go during a scan: residuals straight to their final place when nothing has to be
// merged with them, else to the temp.
struct LongDst {
    int32_t* tmp;
    __device__ __forceinline__ int32_t* resid(const LongMeta& m, int32_t* row) const {
        if (m.ic == 0) return m.copied == 0 ? row : tmp + m.tmp_off + m.d;
        return tmp + m.tmp_off;
    }
    __device__ __forceinline__ int32_t* extras(const LongMeta& m, int32_t* row) const {
        return m.copied == 0 ? row : tmp + m.tmp_off + m.d;
    }
};

// Consume-only scans (`fold` != nullptr) materialise a long record only when somebody copies from it; the successors of
// every other long record are folded where they are produced: residual segments here, interval elements in
// k_long_extras, copied elements in k_long_merge (the three parts of a list are disjoint, so the order they are
// consumed in is immaterial to the checksum).  The materialised ones are folded where their final rows are produced.
struct LongFold {
    unsigned long long* result;   // FOLD_SLOTS slot pairs; nullptr = materialise everything (range decode)
    int32_t from;                 // nodes below `from` are halo: they matter only as parents
    __device__ __forceinline__ bool only_consumed(const LongMeta& m) const { return result != nullptr && !(m.flags & 1); }
};

#ifndef BVG_HOST_EMULATION
template <bool DEF, class RM, bool HIST = false>
__global__ void k_long_resid(GraphDev g, LongIndex li, ItemMap im, int64_t item0, int64_t nitems, int32_t lo, int32_t hi, RM rm, LongDst dst, LongFold lf) {
    const int64_t i = item0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // items [item0, nitems): the long records of [lo, hi)
    unsigned long long acc = 0;
    long long arcs = 0;
    if (i < nitems) {
        int32_t l, part;
        im.find(i, l, part);
        const LongMeta m = li.meta[l];
        if (m.x >= lo && m.x < hi && rm.wanted(g, m.x)) {
            // one loop for both kinds of record, so that the lanes of a warp (segments of four to eight different records)
            // stay together: stored segments write their values, consumed ones fold them
            const bool consume = lf.only_consumed(m);
            if (!DEF) {
                if (!consume) long_resid_segment<DEF>(g, m, li, part, dst.resid(m, rm.row(g, m.x)));
                else if (m.x >= lf.from) {
                    const int32_t first = part * li.seg;
                    const int32_t cnt = min(li.seg, m.rc - first);
                    FoldT<HIST> f;
                    f.begin(m.x, g, true);
                    BitBuf b;
                    b.w = g.words; b.maxw = g.nwords - 3;
                    b.seek(li.seg_pos[m.seg_off + part]);
                    int64_t v = li.seg_val[m.seg_off + part];
                    for (int32_t t = 0; t < cnt; t++) {
                        if (first + t == 0) v = (int64_t)(int32_t)((int64_t)m.x + nat2int(Rd<DEF>::resid(b, g.c)));
                        else v += (int64_t)Rd<DEF>::resid(b, g.c) + 1;
                        f.add((uint32_t)v);
                    }
                    f.n = (uint32_t)cnt;
                    acc = f.finish(m.x);
                    arcs = cnt;
                }
            } else if (!consume || m.x >= lf.from) {
                const int32_t first = part * li.seg;
                const int32_t cnt = min(li.seg, m.rc - first);
                // stored records with neither intervals nor a copied part get their final row here: fold it now
                const bool final_here = (consume || (lf.result != nullptr && m.ic == 0 && m.copied == 0)) && m.x >= lf.from;
                FoldT<HIST> f;
                f.begin(m.x, g, final_here);
                const int k = g.c.zetak;
                Win b;
                b.seek(g, li.seg_pos[m.seg_off + part]);
                uint32_t v = (uint32_t)li.seg_val[m.seg_off + part];
                int32_t* out = consume ? nullptr : dst.resid(m, rm.row(g, m.x)) + first;
#pragma unroll 1
                for (int32_t t = 0; t < cnt; t++) {
                    if (first + t == 0) v = (uint32_t)(int32_t)((int64_t)m.x + nat2int(zeta_any<0>(b, g, k) - 1ull));
                    else {
                        uint32_t mm, len;
                        if (zeta_fast<0>(b.top(), k, mm, len)) b.skip(len);
                        else mm = (uint32_t)zeta_any<0>(b, g, k);
                        v += mm;
                    }
                    f.add(v);
                    if (!consume) out[t] = (int32_t)v;
                }
                if (final_here) {
                    f.n = (uint32_t)cnt;
                    acc = f.finish(m.x);
                    arcs = cnt;
                }
            }
        }
    }
    if (lf.result) warp_fold(acc, arcs, lf.result);
}

template <class RM, bool HIST = false>
__global__ void k_long_extras(GraphDev g, LongIndex li, ItemMap im, int64_t item0, int64_t nitems, int32_t lo, int32_t hi, RM rm, LongDst dst, LongFold lf) {
    const int64_t i = item0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    long long arcs = 0;
    if (i < nitems) {
        int32_t l, part;
        im.find(i, l, part);
        const LongMeta m = li.meta[l];
        if (m.x >= lo && m.x < hi && rm.wanted(g, m.x)) {
            IntervalSeq a{ li.iv_cum + m.iv_off, li.iv_left + m.iv_off, m.ic, m.ilen };
            const int32_t q0 = part * li.chunk;
            if (!lf.only_consumed(m)) {
                const int32_t total = m.ilen + m.rc;
                const int32_t* left = a.left;
                const int32_t cnt = min(li.chunk, total - q0);
                const bool final_row = lf.result != nullptr && m.copied == 0 && m.x >= lf.from;  // no copied part follows
                FoldT<HIST> f;
                f.begin(m.x, g, true);
                merge_chunk(a, [left](int32_t t, int32_t o) { return left[t] + o; }, dst.tmp + m.tmp_off, m.rc, q0,
                            cnt, dst.extras(m, rm.row(g, m.x)), final_row ? &f : nullptr);
                if (final_row) { f.n = (uint32_t)cnt; acc = f.finish(m.x); arcs = cnt; }
            } else if (m.x >= lf.from && q0 < m.ilen) {  // interval elements [q0, q0 + chunk) of the concatenated intervals
                const int32_t cnt = min(li.chunk, m.ilen - q0);
                FoldT<HIST> f;
                f.begin(m.x, g, true);
                int32_t t = upper_slot(a.cum, a.n, q0), t_end = a.cum[t + 1];
                for (int32_t q = q0; q < q0 + cnt; q++) {
                    while (q >= t_end) { t++; t_end = a.cum[t + 1]; }
                    f.add((uint32_t)(a.left[t] + (q - a.cum[t])));
                }
                f.n = (uint32_t)cnt;
                acc = f.finish(m.x);
                arcs = cnt;
            }
        }
    }
    if (lf.result) warp_fold(acc, arcs, lf.result);
}

template <class RM, bool HIST = false>
__global__ void k_long_merge(GraphDev g, LongIndex li, ItemMap im, int64_t item0, int64_t nitems, int32_t lo, int32_t hi, RM rm, LongDst dst, LongFold lf) {
    const int64_t i = item0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    long long arcs = 0;
    if (i < nitems) {
        int32_t l, part;
        im.find(i, l, part);
        const LongMeta m = li.meta[l];
        if (m.x >= lo && m.x < hi && rm.wanted(g, m.x)) {
            const int32_t* parent = rm.row(g, m.x - m.ref);
            CopiedSeq a{ li.cb_cum + m.cb_off, li.cb_ppos + m.cb_off, parent, m.ncb, m.copied };
            const int32_t* ppos = a.ppos;
            const int32_t q0 = part * li.chunk;
            if (!lf.only_consumed(m)) {
                const int32_t cnt = min(li.chunk, m.d - q0);
                const bool final_row = lf.result != nullptr && m.x >= lf.from;
                FoldT<HIST> f;
                f.begin(m.x, g, true);
                merge_chunk(a, [ppos, parent](int32_t t, int32_t o) { return parent[ppos[t] + o]; }, dst.tmp + m.tmp_off + m.d,
                            m.d - m.copied, q0, cnt, rm.row(g, m.x), final_row ? &f : nullptr);
                if (final_row) { f.n = (uint32_t)cnt; acc = f.finish(m.x); arcs = cnt; }
            } else if (m.x >= lf.from && q0 < m.copied) {  // copied elements [q0, q0 + chunk) seen through the copy blocks
                const int32_t cnt = min(li.chunk, m.copied - q0);
                FoldT<HIST> f;
                f.begin(m.x, g, true);
                int32_t t = upper_slot(a.cum, a.n, q0), t_end = a.cum[t + 1];
                for (int32_t q = q0; q < q0 + cnt; q++) {
                    while (q >= t_end) { t++; t_end = a.cum[t + 1]; }
                    f.add((uint32_t)parent[ppos[t] + (q - a.cum[t])]);
                }
                f.n = (uint32_t)cnt;
                acc = f.finish(m.x);
                arcs = cnt;
            }
        }
    }
    if (lf.result) warp_fold(acc, arcs, lf.result);
}

// Layout of the long index on the device (round 2; the host pass it replaces copied every LongMeta both ways and cost 6 ms per
// open of a 125 M-arc shard).  k_long_layout_vals writes, per long record, what it adds to each running sum (rows of `vals`,
// nl entries each): 0 copy blocks + 1, 1 intervals + 1, 2 residual segments, 3 outdegree (per-scan temp: 2 d), 4 outdegree when
// somebody copies from the record (tile scratch: 3 d), 5 extras chunks, 6 speculative sub-ranges, 7 .. 6 + levels merge chunks
// of each chain level, and three statistics for bvg_scan_bits: 7 + levels residual bits, 8 + levels block / interval bits.
// Exclusive scans of the rows give the offsets; k_long_layout_apply stores them into the metas.
__global__ void k_long_layout_vals(const LongMeta* __restrict__ meta, int32_t nl, int32_t levels, int32_t seg, int32_t chunk, int64_t lspec_bits,
                                   int32_t* __restrict__ vals) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    const LongMeta m = meta[l];
    const size_t n = (size_t)nl;
    vals[0 * n + l] = m.ncb + 1;
    vals[1 * n + l] = m.ic + 1;
    vals[2 * n + l] = (m.rc + seg - 1) / seg;
    vals[3 * n + l] = m.d;
    vals[4 * n + l] = (m.flags & 1) ? m.d : 0;
    vals[5 * n + l] = m.ic > 0 ? (int32_t)(((int64_t)m.ilen + m.rc + chunk - 1) / chunk) : 0;
    vals[6 * n + l] = m.rc > 0 ? (int32_t)((m.rec_end - m.resid_pos + (uint64_t)lspec_bits - 1) / (uint64_t)lspec_bits) : 0;
    for (int32_t lv = 1; lv <= levels; lv++)
        vals[(size_t)(6 + lv) * n + l] = (m.copied > 0 && m.level == lv) ? (int32_t)(((int64_t)m.d + chunk - 1) / chunk) : 0;
}
__global__ void k_long_layout_apply(LongMeta* __restrict__ meta, int32_t nl, const int64_t* __restrict__ cb, const int64_t* __restrict__ iv,
                                    const int64_t* __restrict__ sg, const int64_t* __restrict__ dsum, const int64_t* __restrict__ dstored) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    meta[l].cb_off = cb[l]; meta[l].iv_off = iv[l]; meta[l].seg_off = sg[l];
    meta[l].tmp_off = 2 * dsum[l]; meta[l].scan_off = 3 * dstored[l];
}
// what a scan reads of the long records (bvg_scan_bits): residual bits, block + interval bits, arcs
__global__ void k_long_stats(const LongMeta* __restrict__ meta, int32_t nl, unsigned long long* __restrict__ out) {
    unsigned long long r = 0, p = 0, a = 0;
    for (int32_t l = blockIdx.x * blockDim.x + threadIdx.x; l < nl; l += gridDim.x * blockDim.x) {
        r += meta[l].rec_end - meta[l].resid_pos; p += meta[l].resid_pos - meta[l].after_header; a += (unsigned long long)meta[l].d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); p += __shfl_xor_sync(0xffffffffu, p, o); a += __shfl_xor_sync(0xffffffffu, a, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, r); atomicAdd(out + 1, p); atomicAdd(out + 2, a); }
}
__global__ void k_gather_i64(const int64_t* const* __restrict__ ptrs, int32_t n, int64_t* __restrict__ out) {
    const int32_t i = threadIdx.x;
    if (i < n) out[i] = *ptrs[i];
}

// Speculative sub-ranges of the residual sections (see above), listed on the device: item j of record l covers
// LSPEC_BITS bits from resid_pos + part * LSPEC_BITS.
__global__ void k_lspec_init(const LongMeta* __restrict__ meta, ItemMap im, int64_t nitems, SpecItem* __restrict__ items) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nitems) return;
    int32_t l, part;
    im.find(j, l, part);
    const LongMeta m = meta[l];
    SpecItem it;
    it.lo = m.resid_pos + (uint64_t)part * (uint64_t)LSPEC_BITS;
    it.hi = min(it.lo + (uint64_t)LSPEC_BITS, m.rec_end);
    it.entry = it.lo; it.exit = it.lo; it.count = 0; it.sum = 0;
    it.l = l; it.first = part == 0 ? 1 : 0;
    items[j] = it;
}

// Per record: running (count, sum) before each of its sub-ranges; the counts must add up to the record's residuals.
__global__ void k_lspec_scan(GraphDev g, const LongMeta* __restrict__ meta, ItemMap im, const SpecItem* __restrict__ items,
                             int64_t* __restrict__ cbase, int64_t* __restrict__ sbase) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= im.nlong) return;
    int64_t c = 0, sum = 0;
    for (int64_t j = im.cum[l]; j < im.cum[l + 1]; j++) {
        cbase[j] = c; sbase[j] = sum;
        c += items[j].count; sum += items[j].sum;
    }
    if (im.cum[l + 1] > im.cum[l] && c != meta[l].rc) report(g.err, E_FORMAT, meta[l].x, meta[l].resid_pos + g.bit_base);
}
#endif  // BVG_HOST_EMULATION

}  // namespace bvg
