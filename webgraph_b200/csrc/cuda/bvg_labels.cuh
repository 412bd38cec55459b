// bvg_labels.cuh -- the label stream of a BitStreamArcLabelledImmutableGraph decoded on the device.
//
// <basename>.labels holds, for node 0, 1, ..., the labels of the node's arcs in successor order, written one after the other
// by Label.toBitStream with no separators; <basename>.labeloffsets holds gamma-coded gaps between the nodes' first bits
// (BitStreamArcLabelledImmutableGraph.java:139-141, 225-262, 330-358).  The reference reads the labels of node x by
// positioning a bit stream at offset(x) and calling fromBitStream once per successor.  Here the three label classes the
// reference ships are three different parallel shapes:
//   FixedWidthIntLabel      label k of node x sits at offset(x) + k * width: one thread per ARC, perfectly balanced whatever
//                           the outdegrees are (k_lab_fixed; the node of an arc found by a search in the row offsets, narrowed
//                           to the nodes the block's arcs touch).
//   GammaCodedIntLabel      the stream of a node range is nothing but gamma codes back to back, like .offsets: read by the
//                           same speculate / fix passes (bvg_offsets.cuh) over 2048-bit sub-ranges from the range's first
//                           bit, then k_lab_gamma_emit writes label number `ord` to slot `ord` -- the work per thread is a
//                           fixed number of BITS, not a node, so a node with 10^6 arcs costs nothing special.
//   FixedWidthIntListLabel  gamma(length) then length x width bits per arc: positions depend on the lengths read so far, so
//                           one thread per node walks its arcs twice (k_lab_list_count, then k_lab_list_decode after a scan of
//                           the counts).  Not split inside a node.
// All three either store the values (aligned with bvg_decode_range's successors of the same node range) or fold them into
// the label checksum (lab_fold below), which the oracle restates.
#pragma once
#include "bvg_device.cuh"
#include "bvg_offsets.cuh"

namespace bvg {

enum { LAB_GAMMA = 0, LAB_FIXED = 1, LAB_FIXED_LIST = 2 };

struct LabelsDev {
    const uint32_t* __restrict__ w;      // byte-swapped words of the loaded stretch of .labels
    uint64_t maxw;                       // BitBuf convention: nwords - 3
    uint64_t bit_base;                   // file bit position of bit 0 of w
    const uint64_t* __restrict__ off;    // label offsets (file bit positions) of nodes node_lo ..
    const int64_t* __restrict__ rowoff;  // the graph's row offsets of nodes node_lo ..
    int32_t node_lo;
    int32_t width;
};

// Checksum of a labelled range: sum over arcs j (numbered from the first arc of the range) of
// (2j + 1) * t_j mod 2^64, t_j = LAB_LEN_MUL * len_j + sum_i (v_i + 1) * (2i + 1); len = 1 for the integer labels.
// Order-sensitive inside a list and across arcs, so that a label attached to the wrong arc is seen.
constexpr uint64_t LAB_LEN_MUL = 0x9E3779B97F4A7C15ull;
__device__ __forceinline__ uint64_t lab_fold_int(int64_t j, uint32_t v) {
    return (2ull * (uint64_t)j + 1ull) * (LAB_LEN_MUL + (uint64_t)v + 1ull);
}

#ifndef BVG_HOST_EMULATION
__device__ __forceinline__ void lab_block_add(uint64_t acc, unsigned long long* __restrict__ result) {
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(result, (unsigned long long)acc);
}
#endif

// `width` bits at bit position p of the loaded words (0 <= width <= 31).
__device__ __forceinline__ uint32_t lab_bits_at(const LabelsDev& L, uint64_t p, int width) {
    if (width == 0) return 0u;
    uint64_t i = p >> 5;
    i = i < L.maxw ? i : L.maxw;
    const uint64_t two = ((uint64_t)L.w[i] << 32) | (uint64_t)L.w[i + 1];
    return (uint32_t)((two << (p & 31)) >> (64 - width));
}

// ---- FixedWidthIntLabel: one thread per arc -------------------------------------------------------------------------------
constexpr int LAB_FIXED_THREADS = 256, LAB_FIXED_ITEMS = 8, LAB_FIXED_TILE = LAB_FIXED_THREADS * LAB_FIXED_ITEMS;

// Last node x in [lo, hi] with rowoff[x] <= j (rowoff is non-decreasing; rowoff[lo] <= j is given).
__device__ __forceinline__ int32_t lab_node_of(const int64_t* __restrict__ rowoff, int32_t lo, int32_t hi, int64_t j) {
    while (lo < hi) {
        const int32_t mid = lo + (int32_t)(((int64_t)hi - lo + 1) >> 1);
        if (rowoff[mid] <= j) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Label of arc j, whose node is known to lie in [nlo, nhi].
__device__ __forceinline__ uint32_t lab_fixed_value(const LabelsDev& L, int32_t nlo, int32_t nhi, int64_t j) {
    const int32_t x = lab_node_of(L.rowoff, nlo, nhi, j);
    const uint64_t p = L.off[x] - L.bit_base + (uint64_t)(j - L.rowoff[x]) * (uint64_t)L.width;
    return lab_bits_at(L, p, L.width);
}

// ITEMS consecutive arcs from j0 (all inside [j0, jend), nodes inside [nlo, nhi]): one search for the first arc's node, then a
// walk -- the next arc's label starts where this one ends unless the node changes.  Returns how many were read.
__device__ __forceinline__ int lab_fixed_run(const LabelsDev& L, int32_t nlo, int32_t nhi, int64_t j0, int64_t jend, uint32_t* v /* ITEMS */) {
    int32_t x = lab_node_of(L.rowoff, nlo, nhi, j0);
    int64_t next = L.rowoff[x + 1];
    uint64_t p = L.off[x] - L.bit_base + (uint64_t)(j0 - L.rowoff[x]) * (uint64_t)L.width;
    int c = 0;
#pragma unroll
    for (int k = 0; k < LAB_FIXED_ITEMS; k++) {
        const int64_t j = j0 + k;
        if (j < jend) {
            if (j >= next) {
                do { x++; next = L.rowoff[x + 1]; } while (j >= next);   // empty nodes in between
                p = L.off[x] - L.bit_base;
            }
            v[k] = lab_bits_at(L, p, L.width);
            p += (uint64_t)L.width;
            c = k + 1;
        }
    }
    return c;
}

#ifndef BVG_HOST_EMULATION
// Arcs [ra, rb) of nodes [from, to) (node indices relative to node_lo); out[j - ra] = label of arc j.  A thread takes
// LAB_FIXED_ITEMS consecutive arcs: its reads are one contiguous stretch of the stream, its writes two 16-byte stores.
template <bool FOLD>
__global__ void __launch_bounds__(LAB_FIXED_THREADS)
k_lab_fixed(LabelsDev L, int32_t from, int32_t to, int64_t ra, int64_t rb, int32_t* __restrict__ out, unsigned long long* __restrict__ result) {
    __shared__ int32_t s_lo, s_hi;
    const int64_t tile = ra + (int64_t)blockIdx.x * LAB_FIXED_TILE;
    const int64_t tile_end = tile + LAB_FIXED_TILE < rb ? tile + LAB_FIXED_TILE : rb;
    if (threadIdx.x == 0) s_lo = lab_node_of(L.rowoff, from, to - 1, tile);
    if (threadIdx.x == 32) s_hi = lab_node_of(L.rowoff, from, to - 1, tile_end - 1);
    __syncthreads();
    const int64_t j0 = tile + (int64_t)threadIdx.x * LAB_FIXED_ITEMS;
    uint64_t acc = 0;
    if (j0 < tile_end) {
        uint32_t v[LAB_FIXED_ITEMS];
        const int c = lab_fixed_run(L, s_lo, s_hi, j0, tile_end, v);
        if (FOLD) {
#pragma unroll
            for (int k = 0; k < LAB_FIXED_ITEMS; k++) if (k < c) acc += lab_fold_int(j0 + k - ra, v[k]);
        } else {
            int32_t* o = out + (j0 - ra);
            if (c == LAB_FIXED_ITEMS && ((uintptr_t)o & 15) == 0) {
#pragma unroll
                for (int k = 0; k < LAB_FIXED_ITEMS; k += 4) *reinterpret_cast<uint4*>(o + k) = uint4{ v[k], v[k + 1], v[k + 2], v[k + 3] };
            } else {
#pragma unroll
                for (int k = 0; k < LAB_FIXED_ITEMS; k++) if (k < c) o[k] = (int32_t)v[k];
            }
        }
    }
    if (FOLD) lab_block_add(acc, result);
}
#endif

#ifndef BVG_LAB_SUB_BITS
#define BVG_LAB_SUB_BITS 8192
#endif
constexpr uint64_t LAB_SUB_BITS = BVG_LAB_SUB_BITS;   // sub-range pitch of the gamma labels (default; BVG_LAB_SUB_BITS in the environment overrides)

// ---- GammaCodedIntLabel: the proven sub-range chains (OffSub) emitted -----------------------------------------------------
// cbase: exclusive scan of the sub-ranges' counts; label number ord (from the first label of the range) belongs to arc ra + ord.
__device__ inline uint64_t lab_gamma_emit_one(int64_t j, const uint32_t* __restrict__ words, uint64_t nwords, uint64_t base, uint64_t end_bits,
                                              uint64_t sub_bits, const OffSub* __restrict__ sub, const int64_t* __restrict__ cbase, int64_t ra,
                                              int64_t narcs, int32_t* __restrict__ out, bool fold) {
    const uint64_t hi = off_min(base + (uint64_t)(j + 1) * sub_bits, end_bits);
    BitBuf b;
    b.w = words; b.maxw = nwords - 3;
    b.seek(sub[j].entry);
    int64_t ord = cbase[j];
    uint64_t acc = 0;
    (void)ra;
    if (fold) {
        while (b.pos() < hi && ord < narcs) { acc += lab_fold_int(ord, (uint32_t)b.gamma()); ord++; }
        return acc;
    }
    // a thread's labels are consecutive in `out`: groups of four go out as one 16-byte store (a full half sector instead of
    // four partial writes from 32 lanes that are ~100 labels apart)
    const bool vec = ((uintptr_t)out & 15) == 0;
    while (b.pos() < hi && ord < narcs) {
        if (vec && (ord & 3) == 0) {
            int32_t v[4];
            int c = 0;
            while (c < 4 && b.pos() < hi && ord + c < narcs) v[c++] = (int32_t)(uint32_t)b.gamma();
            if (c == 4) *reinterpret_cast<uint4*>(out + ord) = uint4{ (uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3] };
            else for (int i = 0; i < c; i++) out[ord + i] = v[i];
            ord += c;
        } else {
            out[ord++] = (int32_t)(uint32_t)b.gamma();
        }
    }
    return acc;
}

#ifndef BVG_HOST_EMULATION
template <bool FOLD>
__global__ void k_lab_gamma_emit(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t base, uint64_t end_bits, uint64_t sub_bits, int64_t nsub,
                                 const OffSub* __restrict__ sub, const int64_t* __restrict__ cbase, int64_t ra, int64_t narcs,
                                 int32_t* __restrict__ out, unsigned long long* __restrict__ result) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    if (j < nsub) acc = lab_gamma_emit_one(j, words, nwords, base, end_bits, sub_bits, sub, cbase, ra, narcs, out, FOLD);
    if (FOLD) lab_block_add(acc, result);
}
#endif

#ifndef BVG_HOST_EMULATION
__global__ void k_iota_i64(int64_t* __restrict__ out, int64_t n) {  // list offsets of the integer labels: one value per arc
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
#endif

#ifndef BVG_HOST_EMULATION
__global__ void k_lab_sub_counts(const OffSub* __restrict__ sub, int64_t nsub, int32_t* __restrict__ counts) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nsub) counts[j] = (int32_t)sub[j].count;
}
#endif

// ---- FixedWidthIntListLabel: one thread per node, two passes --------------------------------------------------------------
// Walks the lists of node x (relative index); calls f(arc ordinal k, list length, BitBuf positioned at the elements) and
// expects f to leave the buffer after the elements.
template <class F>
__device__ __forceinline__ void lab_list_walk(const LabelsDev& L, int32_t x, F f) {
    const int64_t d = L.rowoff[x + 1] - L.rowoff[x];
    if (d == 0) return;
    BitBuf b;
    b.w = L.w; b.maxw = L.maxw;
    b.seek(L.off[x] - L.bit_base);
    for (int64_t k = 0; k < d; k++) {
        const uint64_t len = b.gamma();
        f(k, len, b);
    }
}

__device__ __forceinline__ uint32_t lab_take(BitBuf& b, int width) {
    if (width == 0) return 0u;
    const uint32_t v = (uint32_t)(b.buf >> (64 - width));
    b.consume(width);
    return v;
}

// counts[x - from] = number of list elements of node x, clamped so that the scan cannot overflow on a corrupt stream (the
// decode pass reports the stream then: the walk is bounded by the node's bits).
__device__ inline void lab_list_count_one(const LabelsDev& L, int32_t x, int32_t from, int32_t* __restrict__ counts, ErrWord* err) {
    const uint64_t end = L.off[x + 1] - L.bit_base;
    int64_t total = 0;
    bool bad = false;
    lab_list_walk(L, x, [&](int64_t, uint64_t len, BitBuf& b) {
        if (bad) return;
        const uint64_t skip = len * (uint64_t)L.width;
        if (len > 0x7fffffffull || b.pos() + skip > end) { bad = true; return; }
        total += (int64_t)len;
        b.seek(b.pos() + skip);
    });
    if (bad || total > 0x7fffffff) { report(err, E_FORMAT, L.node_lo + x, L.off[x]); total = 0; }
    counts[x - from] = (int32_t)total;
}

#ifndef BVG_HOST_EMULATION
__global__ void k_lab_list_count(LabelsDev L, int32_t from, int32_t to, int32_t* __restrict__ counts, ErrWord* err) {
    const int64_t x = (int64_t)from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x < to) lab_list_count_one(L, (int32_t)x, from, counts, err);
}
#endif

// vbase: exclusive scan of counts (values before node x in the range).  list_off[j - ra] = first value of arc j's list,
// list_off[rb - ra] = total (written by the thread of the last node with arcs, or by thread 0 for an arc-less range).
__device__ inline uint64_t lab_list_decode_one(const LabelsDev& L, int32_t x, int32_t from, int32_t to, int64_t ra, const int64_t* __restrict__ vbase,
                                               const int32_t* __restrict__ counts, int64_t* __restrict__ list_off, int32_t* __restrict__ values,
                                               bool fold) {
    if (counts[x - from] == 0 && L.rowoff[x + 1] == L.rowoff[x]) return 0;
    int64_t vp = vbase[x - from];
    const int64_t vend = vp + counts[x - from];
    const int64_t j0 = L.rowoff[x];
    uint64_t acc = 0;
    lab_list_walk(L, x, [&](int64_t k, uint64_t len, BitBuf& b) {
        if (!fold && list_off) list_off[j0 + k - ra] = vp;
        uint64_t t = LAB_LEN_MUL * len;
        for (uint64_t i = 0; i < len && vp < vend; i++) {
            const uint32_t v = lab_take(b, L.width);
            if (fold) t += ((uint64_t)v + 1ull) * (2ull * i + 1ull);
            else values[vp] = (int32_t)v;
            vp++;
        }
        acc += (2ull * (uint64_t)(j0 + k - ra) + 1ull) * t;
    });
    (void)to;
    return acc;
}

#ifndef BVG_HOST_EMULATION
template <bool FOLD>
__global__ void k_lab_list_decode(LabelsDev L, int32_t from, int32_t to, int64_t ra, int64_t rb, const int64_t* __restrict__ vbase,
                                  const int32_t* __restrict__ counts, int64_t* __restrict__ list_off, int32_t* __restrict__ values,
                                  unsigned long long* __restrict__ result) {
    const int64_t x = (int64_t)from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    if (x < to) acc = lab_list_decode_one(L, (int32_t)x, from, to, ra, vbase, counts, list_off, values, FOLD);
    if (!FOLD && list_off && x == from) list_off[rb - ra] = vbase[to - from];
    if (FOLD) lab_block_add(acc, result);
}
#endif

}  // namespace bvg
