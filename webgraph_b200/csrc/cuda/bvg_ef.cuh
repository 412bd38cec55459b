// bvg_ef.cuh -- EFGraph, the reference's quasi-succinct format (EFGraph.java), decoded on the device (decode half).
//
// Per node the stream holds gamma(outdegree), then the Elias-Fano encoding of the d successors plus the terminator upperBound:
// skip pointers (numberOfPointers x pointerSize bits, not needed to enumerate), lower bits ((d + 1) x l), upper bits in unary
// (EFGraph.java:420-556, 1100-1145).  The stream is LSB-first in 64-bit words (LongWordBitReader, :892-1036); `.offsets` holds
// delta-coded gaps (MSB-first, decoded by bvg_offsets.cuh).  Successor k is ((position of the k-th one) - k) << l | lower[k]:
// selection in a bit vector, which the reference does with a running 64-bit window and here is
//   * one thread per node for lists of up to EF_SMALL = 64 successors (the same window walk: ctz, clear lowest one);
//   * one WARP per list up to EF_HEAVY: a step covers 256 upper bits, every lane takes one byte of them (about four ones at
//     the format's density of one half), a warp scan of the popcounts gives each byte the rank of its first one, and the
//     lanes emit their ones -- consecutive ranks land in consecutive lanes, so stores and lower-bit reads are contiguous.
//     A warp owns 32 consecutive nodes: its lanes first walk their own small lists, then the warp goes through the
//     medium ones among the 32 one after the other (geometry broadcast by shuffle);
//   * one block per node above that: every thread takes a word of the upper bits, a block-wide exclusive scan of the
//     popcounts gives each word the rank of its first one, and the ones are emitted in parallel.
// There are no reference chains and no intervals: every list decodes on its own.
#pragma once
#include "bvg_device.cuh"

namespace bvg {

constexpr int32_t EF_SMALL = 64;   // measured (117 M arcs, scan): 4: 1.86 ms, 8: 1.74, 16: 1.53, 32: 1.23, 64: 1.10, 96: 1.19, 128: 1.30, 256: 1.63, 1024: 2.44
constexpr int32_t EF_HEAVY = 8192;
constexpr int EF_BLOCK = 256;

struct EfDev {
    const uint64_t* __restrict__ w;       // long words in host order, >= 2 zero words of padding
    uint64_t nwords;                      // without the padding
    const uint64_t* __restrict__ offsets; // n + 1 bit positions
    int32_t n;
    uint32_t upper_bound;
    int log2_quantum;
};

__device__ __forceinline__ uint64_t ef_word(const EfDev& g, uint64_t i) { return g.w[i < g.nwords + 1 ? i : g.nwords + 1]; }

// `width` bits at bit position pos (0 <= width <= 32 here).  The long words are little-endian in device memory, so the
// stream is just as well a stream of 32-bit words: bit p is bit p % 32 of 32-bit word p / 32 -- two 4-byte loads and a funnel
// shift instead of 64-bit shifts.
__device__ __forceinline__ uint32_t ef_bits(const EfDev& g, uint64_t pos, int width) {
    if (width == 0) return 0u;
    const uint32_t* __restrict__ w32 = reinterpret_cast<const uint32_t*>(g.w);
    uint64_t i = pos >> 5;
    const uint64_t last = 2 * g.nwords + 2;   // the padding words are readable
    i = i < last ? i : last;
#ifdef BVG_HOST_EMULATION
    const uint64_t two = (uint64_t)w32[i] | ((uint64_t)w32[i + 1] << 32);
    const uint32_t v = (uint32_t)(two >> (pos & 31));
#else
    const uint32_t v = __funnelshift_r(w32[i], w32[i + 1], (uint32_t)pos & 31u);
#endif
    return width == 32 ? v : v & ((1u << width) - 1u);
}

// readGamma at *pos (:1002-1036): the number of zeros before the first one is the msb, then msb bits follow.
__device__ __forceinline__ uint64_t ef_gamma(const EfDev& g, uint64_t& pos) {
    int msb = 0;
    for (;;) {
        const uint64_t v = ef_word(g, pos >> 6) >> (pos & 63);
        if (v) { const int z = __ffsll((long long)v) - 1; msb += z; pos += (uint64_t)z + 1; break; }
        msb += 64 - (int)(pos & 63);
        pos = ((pos >> 6) + 1) << 6;
        if (msb > 64) return ~0ull;
    }
    if (msb > 32) { pos += (uint64_t)msb; return ~0ull; }   // an outdegree does not need more
    const uint64_t low = ef_bits(g, pos, msb);
    pos += (uint64_t)msb;
    return (low | (1ull << msb)) - 1ull;
}

// Geometry of node x's list: EFGraph.lowerBits / pointerSize / numberOfPointers (:145-171) for length d + 1.
struct EfList {
    uint64_t lower_start, upper_start;
    int64_t d;
    int l;
};
__device__ __forceinline__ bool ef_list(const EfDev& g, int64_t x, EfList& e) {
    uint64_t pos = g.offsets[x];
    const uint64_t d = ef_gamma(g, pos);
    if (d > 0x7fffffffull || pos > g.offsets[x + 1]) { e.d = 0; return false; }
    e.d = (int64_t)d;
    const uint64_t len = d + 1, ub = g.upper_bound;
    const uint32_t q = (uint32_t)ub / (uint32_t)len;   // both below 2^31 + 1: a 32-bit division
    e.l = q == 0 ? 0 : 31 - __clz((int)q);
    const uint64_t ulen = len + (ub >> e.l);
    const int psize = ulen <= 1 ? 0 : 64 - __clzll((long long)(ulen - 1));
    const uint64_t npointers = (ub >> e.l) >> g.log2_quantum;
    e.lower_start = pos + (uint64_t)psize * npointers;
    e.upper_start = e.lower_start + (uint64_t)e.l * len;
    (void)ulen;
    return true;
}

__device__ __forceinline__ unsigned long long ef_arc_hash(int64_t x, uint32_t y) { return (unsigned long long)x * 0x9E3779B97F4A7C15ull + (unsigned long long)y; }

// Outdegrees of nodes [from, to) (for the row offsets).
__device__ inline void ef_outdegree_one(const EfDev& g, int64_t x, int32_t* __restrict__ outdeg, int64_t at, ErrWord* err) {
    uint64_t pos = g.offsets[x];
    const uint64_t d = ef_gamma(g, pos);
    if (d > 0x7fffffffull || pos > g.offsets[x + 1]) { report(err, E_IO, (int)x, g.offsets[x]); outdeg[at] = 0; return; }
    outdeg[at] = (int32_t)d;
}

// One thread, one list whose geometry is known: out may be null (fold only).  Returns the XOR fold.
__device__ inline unsigned long long ef_walk(const EfDev& g, int64_t x, const EfList& e, int32_t* __restrict__ out, ErrWord* err) {
    unsigned long long acc = 0;
    uint64_t curr = e.upper_start >> 6;
    uint64_t window = ef_word(g, curr) & (~0ull << (e.upper_start & 63));
    const uint64_t end_word = (g.offsets[x + 1] + 63) >> 6;
    for (int64_t k = 0; k < e.d; k++) {
        while (window == 0) {
            if (++curr > end_word) { report(err, E_IO, (int)x, g.offsets[x]); return acc; }
            window = ef_word(g, curr);
        }
        const uint64_t upper = curr * 64 + (uint64_t)(__ffsll((long long)window) - 1) - (uint64_t)k - e.upper_start;
        window &= window - 1;
        const uint32_t v = (uint32_t)(upper << e.l) | ef_bits(g, e.lower_start + (uint64_t)e.l * (uint64_t)k, e.l);
        if (out) out[k] = (int32_t)v;
        acc ^= ef_arc_hash(x, v);
    }
    return acc;
}

__device__ inline unsigned long long ef_decode_one(const EfDev& g, int64_t x, int32_t* __restrict__ out, ErrWord* err) {
    EfList e;
    if (!ef_list(g, x, e)) { report(err, E_IO, (int)x, g.offsets[x]); return 0; }
    return ef_walk(g, x, e, out, err);
}

// The warp path, one lane's share of a step: the byte of upper bits at bit `at` (relative to upper_start, clipped to ulen).
__device__ __forceinline__ uint32_t ef_upper_byte(const EfDev& g, const EfList& e, uint64_t ulen, uint64_t at) {
    if (at >= ulen) return 0u;
    uint32_t b = ef_bits(g, e.upper_start + at, 8);
    const uint64_t left = ulen - at;
    if (left < 8) b &= (1u << left) - 1u;
    return b;
}
// Emits the ones of `byte` (first one has rank k): values ((at + j) - rank) << l | lower[rank].
__device__ __forceinline__ unsigned long long ef_emit_byte(const EfDev& g, int64_t x, const EfList& e, uint64_t at, uint32_t byte, int64_t k,
                                                           int32_t* __restrict__ out) {
    unsigned long long acc = 0;
    while (byte && k < e.d) {
        const int j = __ffs((int)byte) - 1;
        byte &= byte - 1;
        const uint32_t v = (uint32_t)((at + (uint64_t)j - (uint64_t)k) << e.l) | ef_bits(g, e.lower_start + (uint64_t)e.l * (uint64_t)k, e.l);
        if (out) out[k] = (int32_t)v;
        acc ^= ef_arc_hash(x, v);
        k++;
    }
    return acc;
}

#ifndef BVG_HOST_EMULATION
__global__ void k_ef_outdegrees(EfDev g, int32_t from, int32_t to, int32_t* __restrict__ outdeg, ErrWord* err) {
    const int64_t x = (int64_t)from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x < to) ef_outdegree_one(g, x, outdeg, x - from, err);
}

__device__ __forceinline__ void ef_block_fold(unsigned long long arcs, unsigned long long acc, unsigned long long* __restrict__ result) {
    for (int o = 16; o > 0; o >>= 1) { arcs += __shfl_down_sync(0xffffffffu, arcs, o); acc ^= __shfl_down_sync(0xffffffffu, acc, o); }
    if ((threadIdx.x & 31) == 0) {
        if (arcs) atomicAdd(result, arcs);
        if (acc) atomicXor(result + 1, acc);
    }
}

// rowoff: row offsets of nodes from.. (rowoff[0] = first arc of `from`); rows = out - rowoff[0].  Lists above EF_HEAVY go to
// heavy[] for k_ef_decode_heavy.  result (arcs, xor) may be null; out may be null (scan).
__global__ void __launch_bounds__(EF_BLOCK) k_ef_decode(EfDev g, int32_t from, int32_t to, const int64_t* __restrict__ rowoff, int32_t* __restrict__ out,
                                                        int32_t* __restrict__ heavy, int32_t* __restrict__ nheavy,
                                                        unsigned long long* __restrict__ result, ErrWord* err, int32_t small = EF_SMALL) {
    const int64_t x = (int64_t)from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    unsigned long long acc = 0, arcs = 0;
    EfList e;
    e.d = 0; e.l = 0; e.lower_start = e.upper_start = 0;
    int64_t a = 0;
    bool ok = false;
    if (x < to) {
        a = rowoff[x - from];
        const int64_t d = rowoff[x - from + 1] - a;
        if (d > 0) {
            ok = ef_list(g, x, e) && e.d == d;
            if (!ok) report(err, E_IO, (int)x, g.offsets[x]);
        }
    }
    if (ok) {
        if (e.d > EF_HEAVY) heavy[atomicAdd(nheavy, 1)] = (int32_t)x;
        else arcs = (unsigned long long)e.d;
        if (e.d <= small) acc = ef_walk(g, x, e, out ? out + (a - rowoff[0]) : nullptr, err);
    }
    // the medium lists of this warp's 32 nodes, one after the other, all lanes on each
    unsigned med = __ballot_sync(0xffffffffu, ok && e.d > small && e.d <= EF_HEAVY);
    while (med) {
        const int src = __ffs((int)med) - 1;
        med &= med - 1;
        EfList w;
        w.d = __shfl_sync(0xffffffffu, e.d, src);
        w.l = __shfl_sync(0xffffffffu, e.l, src);
        w.lower_start = __shfl_sync(0xffffffffu, e.lower_start, src);
        w.upper_start = __shfl_sync(0xffffffffu, e.upper_start, src);
        const int64_t wa = __shfl_sync(0xffffffffu, a, src);
        const int64_t wx = x - lane + src;
        int32_t* o = out ? out + (wa - rowoff[0]) : nullptr;
        const uint64_t ulen = (uint64_t)w.d + 1 + ((uint64_t)g.upper_bound >> w.l);
        int64_t carry = 0;
        for (uint64_t base = 0; base < ulen && carry < w.d; base += 256) {
            const uint64_t at = base + (uint64_t)lane * 8;
            const uint32_t byte = ef_upper_byte(g, w, ulen, at);
            const int pc = __popc(byte);
            int inc = pc;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, s); if (lane >= s) inc += t; }
            acc ^= ef_emit_byte(g, wx, w, at, byte, carry + inc - pc, o);
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (carry < w.d && lane == 0) report(err, E_IO, (int)wx, g.offsets[wx]);   // fewer ones than successors
    }
    if (result) ef_block_fold(arcs, acc, result);
}

// EF_SPLIT blocks per heavy list.  Every block of a list goes through all tiles of EF_BLOCK upper words (cheap: popcounts, the
// words sit in L2 after the first block) to keep the running rank, and emits the ones of every EF_SPLIT-th tile only: the
// emission -- a dependent chain of ~32 successors per word and thread -- is what takes the time, and it is spread S ways.
constexpr int EF_SPLIT = 8;
__global__ void __launch_bounds__(EF_BLOCK) k_ef_decode_heavy(EfDev g, int32_t from, const int32_t* __restrict__ heavy, const int64_t* __restrict__ rowoff,
                                                              int32_t* __restrict__ out, unsigned long long* __restrict__ result, ErrWord* err) {
    __shared__ int32_t warp_sums[EF_BLOCK / 32];
    const int64_t x = heavy[blockIdx.x / EF_SPLIT];
    const int part = blockIdx.x % EF_SPLIT;
    EfList e;
    const bool ok = ef_list(g, x, e);
    const int64_t a = rowoff[x - from];
    int32_t* o = out ? out + (a - rowoff[0]) : nullptr;
    if (!ok) { if (threadIdx.x == 0 && part == 0) report(err, E_IO, (int)x, g.offsets[x]); return; }
    const uint64_t first_word = e.upper_start >> 6;
    const uint64_t ulen = (uint64_t)e.d + 1 + ((uint64_t)g.upper_bound >> e.l);
    const uint64_t last_word = (e.upper_start + ulen - 1) >> 6;
    unsigned long long acc = 0, arcs = 0;
    int64_t carry = 0;
    int tile = 0;
    for (uint64_t base = first_word; base <= last_word && carry < e.d; base += EF_BLOCK, tile++) {
        const uint64_t wi = base + threadIdx.x;
        uint64_t word = 0;
        if (wi <= last_word) {
            word = ef_word(g, wi);
            if (wi == first_word) word &= ~0ull << (e.upper_start & 63);
        }
        const int pc = __popcll(word);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int inc = pc;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, s); if (lane >= s) inc += t; }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        int wbase = 0, tot = 0;
#pragma unroll
        for (int wv = 0; wv < EF_BLOCK / 32; wv++) { const int sv = warp_sums[wv]; if (wv < wid) wbase += sv; tot += sv; }
        __syncthreads();
        if (tile % EF_SPLIT == part) {
            int64_t k = carry + wbase + inc - pc;
            while (word && k < e.d) {
                const uint64_t p = wi * 64 + (uint64_t)(__ffsll((long long)word) - 1);
                word &= word - 1;
                const uint32_t v = (uint32_t)((p - (uint64_t)k - e.upper_start) << e.l) | ef_bits(g, e.lower_start + (uint64_t)e.l * (uint64_t)k, e.l);
                if (o) o[k] = (int32_t)v;
                acc ^= ef_arc_hash(x, v);
                arcs++;
                k++;
            }
        }
        carry += tot;   // every thread keeps the same running rank
    }
    if (threadIdx.x == 0 && part == 0 && carry < e.d) report(err, E_IO, (int)x, g.offsets[x]);   // fewer ones than successors
    if (result) ef_block_fold(arcs, acc, result);
}
#endif

}  // namespace bvg

// ---------------------------------------------------------------------------------------------------------------------------
// EFGraph.store on the device (EFGraph.java:812-888 with Accumulator :420-556): every field of the format is a function of one
// element of one list, so the stream is written element-parallel.  Elements are the successors plus one terminator per node
// (value upperBound); element k of node x has high part h = v >> l and contributes
//   * its l lower bits at lower_start + k * l,
//   * a one at upper position h + k,
//   * the skip pointers b with h_prev < b * quantum <= h (h_prev = high part of element k - 1, or 0 zeros before the first):
//     pointer b - 1 = b * quantum + k  (the position just after the (b * quantum)-th zero: that many zeros and k ones precede it),
//   * and, for k = 0, gamma(outdegree).
// Neighbouring nodes share words, so fields are ORed in with atomics on 64-bit words (the buffer starts zeroed).
// ---------------------------------------------------------------------------------------------------------------------------
namespace bvg {

struct EfcDev {
    const int64_t* __restrict__ off;    // CSR row offsets, n + 1
    const int32_t* __restrict__ succ;
    int32_t n;
    uint32_t upper_bound;
    int log2_quantum;
};

struct EfcGeom { int l, psize, gbits; uint64_t npointers, ulen, bits; };
__device__ __forceinline__ EfcGeom efc_geometry(uint64_t d, uint32_t ub, int q) {
    EfcGeom e;
    const uint64_t len = d + 1;
    const uint32_t quot = ub / (uint32_t)len;
    e.l = quot == 0 ? 0 : 31 - __clz((int)quot);
    e.ulen = len + ((uint64_t)ub >> e.l);
    e.psize = e.ulen <= 1 ? 0 : 64 - __clzll((long long)(e.ulen - 1));
    e.npointers = ((uint64_t)ub >> e.l) >> q;
    const int msb = 63 - __clzll((long long)(d + 1));
    e.gbits = 2 * msb + 1;
    e.bits = (uint64_t)e.gbits + (uint64_t)e.psize * e.npointers + (uint64_t)e.l * len + e.ulen;
    return e;
}

// ORs the low `width` bits of v into the stream at bit position pos (width <= 64).
__device__ __forceinline__ void efc_put(unsigned long long* __restrict__ w, uint64_t pos, uint64_t v, int width) {
    if (width == 0) return;
    if (width < 64) v &= (~0ull >> (64 - width));
    const uint64_t i = pos >> 6;
    const int s = (int)(pos & 63);
    if (v << s) atomicOr(w + i, (unsigned long long)(v << s));
    if (s && width > 64 - s && (v >> (64 - s))) atomicOr(w + i + 1, (unsigned long long)(v >> (64 - s)));
}

__device__ inline void efc_sizes_one(const EfcDev& c, int64_t x, int32_t* __restrict__ bits, int* __restrict__ bad) {
    const int64_t d = c.off[x + 1] - c.off[x];
    if (d < 0 || d > 0x7ffffffe) { *bad = 1; bits[x] = 0; return; }
    const EfcGeom e = efc_geometry((uint64_t)d, c.upper_bound, c.log2_quantum);
    if (e.bits > 0x7fffffffull) { *bad = 1; bits[x] = 0; return; }   // a record of 2^31 bits or more: not a case for this writer
    bits[x] = (int32_t)e.bits;
}

// Element e of the whole graph (elements of node x: eoff(x) = off[x] + x .. eoff(x + 1)); node_bits: bit offset of every node.
__device__ inline void efc_write_one(const EfcDev& c, int64_t x, int64_t k, const int64_t* __restrict__ node_bits,
                                     unsigned long long* __restrict__ w, int* __restrict__ bad) {
    const int64_t a = c.off[x], d = c.off[x + 1] - a;
    const EfcGeom e = efc_geometry((uint64_t)d, c.upper_bound, c.log2_quantum);
    const uint64_t start = (uint64_t)node_bits[x];
    const uint64_t pointer_start = start + (uint64_t)e.gbits;
    const uint64_t lower_start = pointer_start + (uint64_t)e.psize * e.npointers;
    const uint64_t upper_start = lower_start + (uint64_t)e.l * (uint64_t)(d + 1);
    if (k == 0) {   // gamma(d): the word 1 << msb in msb + 1 bits, then the msb low bits of d + 1 (LongWordOutputBitStream, :396-409)
        const uint64_t v = (uint64_t)d + 1;
        const int msb = (e.gbits - 1) / 2;
        efc_put(w, start, 1ull << msb, msb + 1);
        efc_put(w, start + (uint64_t)msb + 1, v ^ (1ull << msb), msb);
    }
    const uint64_t v = k < d ? (uint64_t)(uint32_t)c.succ[a + k] : (uint64_t)c.upper_bound;
    uint64_t prev_h = 0;
    if (k > 0) {
        const uint64_t pv = (uint64_t)(uint32_t)c.succ[a + k - 1];
        if (k < d && v <= pv) *bad = 1;   // lists are strictly increasing (Accumulator.add, :499)
        prev_h = pv >> e.l;
    }
    if (k < d && (c.succ[a + k] < 0 || v >= c.upper_bound)) *bad = 1;
    const uint64_t h = v >> e.l;
    if (e.l) efc_put(w, lower_start + (uint64_t)e.l * (uint64_t)k, v, e.l);
    efc_put(w, upper_start + h + (uint64_t)k, 1ull, 1);
    const uint64_t quantum = 1ull << c.log2_quantum;
    for (uint64_t b = prev_h / quantum + 1; b * quantum <= h; b++)   // pointers to the zeros this element is the first one after
        if (b <= e.npointers) efc_put(w, pointer_start + (b - 1) * (uint64_t)e.psize, b * quantum + (uint64_t)k, e.psize);
}

#ifndef BVG_HOST_EMULATION
__global__ void k_efc_sizes(EfcDev c, int32_t* __restrict__ bits, int* __restrict__ bad) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x < c.n) efc_sizes_one(c, x, bits, bad);
}

constexpr int EFC_THREADS = 256, EFC_ITEMS = 4, EFC_TILE = EFC_THREADS * EFC_ITEMS;

// Last node x in [lo, hi] with off[x] + x <= e.
__device__ __forceinline__ int32_t efc_node_of(const int64_t* __restrict__ off, int32_t lo, int32_t hi, int64_t e) {
    while (lo < hi) {
        const int32_t mid = lo + (int32_t)(((int64_t)hi - lo + 1) >> 1);
        if (off[mid] + mid <= e) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// One thread per EFC_ITEMS consecutive elements; a block's node range is found first so that the per-thread search is short.
__global__ void __launch_bounds__(EFC_THREADS) k_efc_write(EfcDev c, const int64_t* __restrict__ node_bits, unsigned long long* __restrict__ w,
                                                           int* __restrict__ bad) {
    __shared__ int32_t s_lo, s_hi;
    const int64_t total = c.off[c.n] + c.n;
    const int64_t tile = (int64_t)blockIdx.x * EFC_TILE;
    const int64_t tile_end = tile + EFC_TILE < total ? tile + EFC_TILE : total;
    if (threadIdx.x == 0) s_lo = efc_node_of(c.off, 0, c.n - 1, tile);
    if (threadIdx.x == 32) s_hi = efc_node_of(c.off, 0, c.n - 1, tile_end - 1);
    __syncthreads();
    int64_t e = tile + (int64_t)threadIdx.x * EFC_ITEMS;
    if (e >= tile_end) return;
    int32_t x = efc_node_of(c.off, s_lo, s_hi, e);
#pragma unroll
    for (int i = 0; i < EFC_ITEMS; i++, e++) {
        if (e >= tile_end) break;
        while (e >= c.off[x + 1] + x + 1) x++;
        efc_write_one(c, x, e - (c.off[x] + x), node_bits, w, bad);
    }
}
#endif

}  // namespace bvg
