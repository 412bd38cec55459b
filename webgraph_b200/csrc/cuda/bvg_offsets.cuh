// bvg_offsets.cuh -- the .offsets stream decoded on the device.
//
// .offsets is n+1 gamma (or delta) coded gaps, one after the other with no index (OffsetsLongIterator,
// BVGraph.java:907-935); the reference reads it with one sequential pass on the CPU, which for 32 M nodes costs more
// than the whole GPU scan of the graph.  Codeword k+1 starts where codeword k ends, so a parallel reader has to guess
// entry points and prove them right:
//   k_off_speculate  each thread takes a sub-range of OFF_SUB_BITS bits and decodes from its first bit as if a code
//                    started there (true for sub-range 0 only), until it leaves the sub-range; it records where it
//                    left (exit), how many codes it read and their sum.
//   k_off_fix        each thread compares its assumed entry with the exit of the sub-range before it.  If they
//                    differ it walks both chains in lockstep: instantaneous codes re-synchronise within a few code words,
//                    and from the meeting point on everything it had decoded is right, so only the count and sum
//                    before that point are corrected.  If the chains do not meet inside the sub-range its exit
//                    changes and the next sub-range has to look again: the host repeats the pass until nothing changes
//                    (sub-range 0 is exact, so pass k makes sub-ranges 0..k exact; in practice two passes).
//   k_off_emit       after an exclusive scan of counts and sums every thread re-reads its sub-range from its proven
//                    entry and writes the absolute offsets.
// The result is bit-exact by construction (every entry point has been walked to from bit 0), not probabilistic.
// `base` (default 0) moves bit 0 of the sub-range grid and `sub_bits` sets its pitch: the gamma-coded label stream
// (bvg_labels.cuh) is read the same way from the first label of a node range, a position its label offsets give exactly,
// with longer sub-ranges (label codes are longer than offset gaps, and two chains over long codes take longer to meet).
#pragma once
#include "bvg_device.cuh"

namespace bvg {

#ifndef BVG_OFF_SUB_BITS
#define BVG_OFF_SUB_BITS 2048
#endif
constexpr int64_t OFF_SUB_BITS = BVG_OFF_SUB_BITS;

struct OffSub {
    uint64_t entry, exit;   // bit positions: where decoding of this sub-range starts / where it left the sub-range
    int64_t count;          // codes starting inside the sub-range on the chain from `entry`
    uint64_t sum;           // their sum
};

__device__ __forceinline__ uint64_t off_code(BitBuf& b, int coding) { return coding == C_GAMMA ? b.gamma() : b.delta(); }

// Walks the chain from `pos` to the end of the sub-range.
__device__ __forceinline__ void off_walk(const uint32_t* __restrict__ words, uint64_t maxw, int coding, uint64_t pos, uint64_t hi,
                                         uint64_t& exit, int64_t& count, uint64_t& sum) {
    BitBuf b;
    b.w = words; b.maxw = maxw;
    b.seek(pos);
    count = 0; sum = 0;
    while (b.pos() < hi) { sum += off_code(b, coding); count++; }
    exit = b.pos();
}

__device__ __forceinline__ uint64_t off_min(uint64_t a, uint64_t b) { return a < b ? a : b; }

__device__ inline void off_speculate_one(int64_t j, const uint32_t* __restrict__ words, uint64_t nwords, uint64_t total_bits,
                                         int coding, OffSub* __restrict__ sub, uint64_t base = 0, uint64_t sub_bits = OFF_SUB_BITS) {
    const uint64_t lo = base + (uint64_t)j * sub_bits, hi = off_min(lo + sub_bits, total_bits);
    OffSub s;
    s.entry = lo;
    off_walk(words, nwords - 3, coding, lo, hi, s.exit, s.count, s.sum);
    sub[j] = s;
}

__global__ void k_off_speculate(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t total_bits, int coding,
                                int64_t nsub, OffSub* __restrict__ sub, uint64_t base = 0, uint64_t sub_bits = OFF_SUB_BITS) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nsub) off_speculate_one(j, words, nwords, total_bits, coding, sub, base, sub_bits);
}

__device__ inline void off_fix_one(int64_t j, const uint32_t* __restrict__ words, uint64_t nwords, uint64_t total_bits, int coding,
                                   const OffSub* __restrict__ in, OffSub* __restrict__ out, int* __restrict__ changed,
                                   uint64_t base = 0, uint64_t sub_bits = OFF_SUB_BITS) {
    OffSub s = in[j];
    if (j > 0) {
        const uint64_t t = in[j - 1].exit;  // the entry the previous sub-range's chain dictates
        const uint64_t hi = off_min(base + (uint64_t)(j + 1) * sub_bits, total_bits);
        if (t != s.entry) {
            BitBuf a, b;
            a.w = b.w = words; a.maxw = b.maxw = nwords - 3;
            a.seek(t); b.seek(s.entry);
            int64_t ca = 0, cb = 0;
            uint64_t sa = 0, sb = 0;
            // two-pointer walk: advance whichever chain is behind until they stand on the same bit
            while (a.pos() != b.pos() && a.pos() < hi && b.pos() < hi) {
                if (a.pos() < b.pos()) { sa += off_code(a, coding); ca++; }
                else { sb += off_code(b, coding); cb++; }
            }
            if (a.pos() == b.pos()) {  // met inside the sub-range: the tail (and the exit) was right all along
                s.count += ca - cb;
                s.sum += sa - sb;
            } else {  // no meeting point: finish the true chain, the exit moves
                while (a.pos() < hi) { sa += off_code(a, coding); ca++; }
                s.count = ca;
                s.sum = sa;
                if (a.pos() != s.exit) { s.exit = a.pos(); *changed = 1; }
            }
            s.entry = t;
        }
    }
    out[j] = s;
}

__global__ void k_off_fix(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t total_bits, int coding, int64_t nsub,
                          const OffSub* __restrict__ in, OffSub* __restrict__ out, int* __restrict__ changed, uint64_t base = 0,
                          uint64_t sub_bits = OFF_SUB_BITS) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nsub) off_fix_one(j, words, nwords, total_bits, coding, in, out, changed, base, sub_bits);
}

// cbase / sbase: exclusive scans of count / sum.  offsets[i] = sum of the first i+1 gaps, i = 0..n.
__device__ inline void off_emit_one(int64_t j, const uint32_t* __restrict__ words, uint64_t nwords, uint64_t total_bits, int coding,
                                    const OffSub* __restrict__ sub, const int64_t* __restrict__ cbase, const uint64_t* __restrict__ sbase,
                                    int64_t n, uint64_t* __restrict__ offsets) {
    const uint64_t hi = off_min((uint64_t)(j + 1) * OFF_SUB_BITS, total_bits);
    BitBuf b;
    b.w = words; b.maxw = nwords - 3;
    b.seek(sub[j].entry);
    int64_t ord = cbase[j];
    uint64_t acc = sbase[j];
    while (b.pos() < hi && ord <= n) {
        acc += off_code(b, coding);
        offsets[ord++] = acc;
    }
}

__global__ void k_off_emit(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t total_bits, int coding, int64_t nsub,
                           const OffSub* __restrict__ sub, const int64_t* __restrict__ cbase, const uint64_t* __restrict__ sbase,
                           int64_t n, uint64_t* __restrict__ offsets) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nsub) off_emit_one(j, words, nwords, total_bits, coding, sub, cbase, sbase, n, offsets);
}

}  // namespace bvg
