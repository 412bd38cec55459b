// bvg_scan.cuh -- the lean per-record walkers of the consume-only scan (default codings).
//
// The fused scan (k_scan_extras / k_scan_merge in bvg_kernels.cuh) is bound by the integer (alu) pipe of the SM, not by
// memory: ncu shows the alu pipe as the busiest unit and ~57 SASS instructions per residual in the first version.  What
// is here is the same arithmetic as ExtrasWalk / MergeWalk (bvg_device.cuh; reference BVGraph.java:1044-1100,
// ResidualIntIterator :939-991, MaskedIntIterator.java:65-97) written for the instruction count:
//   * Win: a sliding window of two stream words + one word of lookahead; the next code's first 32 bits are one funnel
//     shift, a refill is three register moves and one load, and all indices are 32-bit.
//   * zeta_k / gamma codes that fit 32 bits are decoded from that window with one count-leading-zeros and two shifts;
//     longer ones (gaps >= 2^24 for k = 3) fall back to the position-based reader `Bits`.
//   * the checksum fold keeps 32-bit halves: XOR of the low words plus a count of the carries into the high word
//     (the high word of x*MIX + y is hi(x*MIX) or hi(x*MIX) + 1).
//   * records whose list nobody copies from are only consumed, so their intervals need not be merged with their
//     residuals: interval elements are folded while the interval section is walked, then the residuals in a tight loop.
#pragma once
#include "bvg_device.cuh"

namespace bvg {

__device__ __forceinline__ int half_octave_bucket(uint64_t v) {  // quarter octaves: 0..255, monotone in v
    if (v == 0) return 0;
    const int l = 63 - __clzll((long long)v);
    const int frac = l >= 2 ? (int)((v >> (l - 2)) & 3) : (l == 1 ? (int)((v & 1) << 1) : 0);
    return 4 * l + frac;
}

__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }

// Codes that do not fit the 32-bit window go through the position-based reader, out of line: they are rare (gaps >= 2^24
// for zeta_3, values >= 2^16 - 1 for gamma) and inlining them at every call site triples the size of the hot loops.
__device__ BVG_NOINLINE uint64_t slow_gamma(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t* pos) {
    Bits t;
    t.w = words; t.maxw = nwords - 3; t.pos = *pos;
    const uint64_t r = t.gamma();
    *pos = t.pos;
    return r;
}
__device__ BVG_NOINLINE uint64_t slow_unary(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t* pos) {
    Bits t;
    t.w = words; t.maxw = nwords - 3; t.pos = *pos;
    const uint64_t r = t.unary();
    *pos = t.pos;
    return r;
}
__device__ BVG_NOINLINE uint64_t slow_zeta(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t* pos, int k) {
    Bits t;
    t.w = words; t.maxw = nwords - 3; t.pos = *pos;
    const uint64_t r = t.zeta(k);
    *pos = t.pos;
    return r;
}

// A ring is addressed by its 32-bit shared-memory address on the device (kept in a register; a generic pointer makes
// ptxas rebuild it from %tid at every refill) and by a plain pointer under host emulation.
#ifdef BVG_HOST_EMULATION
typedef unsigned char* ring_addr;
__device__ __forceinline__ ring_addr ring_address(void* p) { return (unsigned char*)p; }
__device__ __forceinline__ uint32_t ring_load(ring_addr a, uint32_t off) { uint32_t v; memcpy(&v, a + off, 4); return v; }
#define BVG_CP_ASYNC16(dst, off, src) memcpy((dst) + (off), (src), 16)
#define BVG_CP_COMMIT()
#define BVG_CP_WAIT_ALL()
#define BVG_CP_WAIT_1()
#else
typedef uint32_t ring_addr;
__device__ __forceinline__ ring_addr ring_address(void* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));  // opaque: keep it in a register
    return a;
}
__device__ __forceinline__ uint32_t ring_load(ring_addr a, uint32_t off) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + off) : "memory");
    return v;
}
#define BVG_CP_ASYNC16(dst, off, src) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((dst) + (off)), "l"(src) : "memory")
#define BVG_CP_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define BVG_CP_WAIT_ALL() asm volatile("cp.async.wait_group 0;" ::: "memory")
#define BVG_CP_WAIT_1() asm volatile("cp.async.wait_group 1;" ::: "memory")
#endif

// LA = words of lookahead (1..3).  A lane opens a new 32-byte sector of its record every eighth word and a warp-wide
// refill load therefore almost always carries a lane that misses L1; the loaded word has to be requested long enough
// before the funnel shift that needs it (ncu, LA = 1: 20 % of all stall samples sit on that shift).
template <int LA>
struct WinT {
    const uint32_t* __restrict__ p0;  // word holding the bit the window was opened at
    uint32_t idx, lim;                // index (from p0) of the last lookahead word; last readable index
    uint32_t w0, w1, q0, q1, q2;      // current two words, lookahead queue (q0 next; q1, q2 used when LA >= 2, 3)
    uint32_t s;                       // bits of w0 already consumed, 0..31

    __device__ __forceinline__ void seek(const GraphDev& g, uint64_t pos) {
        uint64_t i = pos >> 5;
        const uint64_t last = g.nwords - 1;
        if (i > last) i = last;  // a corrupt stream can point past the end: clamp, the caller reports E_IO
        p0 = g.words + i;
        const uint64_t room = last - i;
        lim = room > 0x7fffffffull ? 0x7fffffffu : (uint32_t)room;
        s = (uint32_t)pos & 31u;
        w0 = p0[0];
        w1 = p0[umin32(1u, lim)];
        q0 = p0[umin32(2u, lim)];
        q1 = q2 = 0;
        if (LA >= 2) q1 = p0[umin32(3u, lim)];
        if (LA >= 3) q2 = p0[umin32(4u, lim)];
        idx = umin32(1u + LA, lim);
    }
    // bit position relative to word 0 of the stream buffer (exact unless the window ran into the end of the buffer)
    __device__ __forceinline__ uint64_t pos(const GraphDev& g) const {
        return ((uint64_t)(p0 - g.words) + idx - (1u + LA)) * 32u + s;
    }
    __device__ __forceinline__ bool overrun() const { return idx >= lim; }
    __device__ __forceinline__ void attach(ring_addr) {}
    __device__ __forceinline__ void topup() {}
    __device__ __forceinline__ uint32_t top() const { return __funnelshift_l(w1, w0, s); }
    __device__ __forceinline__ uint64_t top64() const { return ((uint64_t)__funnelshift_l(w1, w0, s) << 32) | __funnelshift_l(q0, w1, s); }
    __device__ __forceinline__ void skip(uint32_t n) {  // n <= 32
        s += n;
        if (s >= 32u) {
            s -= 32u;
            w0 = w1;
            w1 = q0;
            idx = umin32(idx + 1u, lim);
            if (LA == 1) q0 = p0[idx];
            if (LA == 2) { q0 = q1; q1 = p0[idx]; }
            if (LA == 3) { q0 = q1; q1 = q2; q2 = p0[idx]; }
        }
    }
    __device__ __forceinline__ uint64_t gamma_slow(const GraphDev& g) {
        uint64_t p = pos(g);
        const uint64_t r = slow_gamma(g.words, g.nwords, &p);
        seek(g, p);
        return r;
    }
    __device__ __forceinline__ uint64_t zeta_slow(const GraphDev& g, int k) {
        uint64_t p = pos(g);
        const uint64_t r = slow_zeta(g.words, g.nwords, &p, k);
        seek(g, p);
        return r;
    }
    // gamma (BVGraph's block, block-count, interval codes): 32-bit window when the code has at most 31 bits
    __device__ __forceinline__ uint64_t gamma(const GraphDev& g) {
        const uint32_t t = top();
        const int m = __clz((int)t);
        if (m <= 15) {
            skip(2u * m + 1u);
            return (uint64_t)((t >> (31 - 2 * m)) - 1u);
        }
        return gamma_slow(g);
    }
    // unary (BVGraph's reference code): zeros before the first one; at most 31 of them in the 32-bit window
    __device__ __forceinline__ uint64_t unary(const GraphDev& g) {
        const uint32_t t = top();
        if (t != 0u) {
            const uint32_t z = (uint32_t)__clz((int)t);
            skip(z + 1u);
            return z;
        }
        uint64_t p = pos(g);
        const uint64_t r = slow_unary(g.words, g.nwords, &p);
        seek(g, p);
        return r;
    }
};
#ifndef BVG_WIN_LA
#define BVG_WIN_LA 1
#endif
typedef WinT<1> Win;  // block lists and interval sections: a few codes per record

// ---------------------------------------------------------------------------------------------------
// The same window fed from shared memory: every lane owns a ring of RING_GROUPS 16-byte groups of its record's stream,
// filled by asynchronous global->shared copies (cp.async, no register and no scoreboard involved) that run two to
// three groups ahead of the decoder; a refill is then a shared-memory load.  With refills straight from global memory
// the loop waits for an L1 miss almost every trip: each lane opens a new 32-byte sector every eighth word, a warp-wide
// load nearly always carries such a lane, and ptxas keeps all of the window's loads on one scoreboard, so no amount of
// register lookahead helps (measured: 1, 2 and 3 words of lookahead and L1/L2 prefetches all give the same time).
//   topup() must be called at least once every four words consumed (the residual loop calls it every fourth code, all
//   lanes together): it issues at most one group copy and waits for everything but the newest group.
//   Invariant (cg = group of the lookahead word, fg = next group to copy): after every topup fg >= cg + 3.  At most four
//   words, hence one group boundary, are crossed before the next topup, so there fg >= cg + 2: group cg + 2 is copied now
//   if it was not yet, and the groups up to cg + 1 -- all the window can reach before the topup after this one -- were
//   copied at least one topup ago and are complete once this topup's wait returns.  A slot is rewritten only by group
//   fg <= cg + 2 <= cg + RING_GROUPS - 1, whose previous tenant fg - RING_GROUPS < cg is dead.  seek() loads four groups.
// ---------------------------------------------------------------------------------------------------
constexpr int RING_GROUPS = 4;

template <int NT>  // threads per block: group slot i of a lane lives NT * 16 bytes after slot i - 1
struct WinRing {
    ring_addr ring;                   // this lane's 16 bytes of slot 0 (shared memory)
    const uint4* __restrict__ g0;     // group holding the bit the window was opened at
    uint32_t glim;                    // last readable group, from g0
    uint32_t r;                       // word index (from g0's first word) of the lookahead word
    uint32_t fg;                      // next group to copy, from g0
    uint32_t w0, w1, q;               // current two words, lookahead
    uint32_t s;                       // bits of w0 already consumed, 0..31

    __device__ __forceinline__ void attach(ring_addr lane_slot0) { ring = lane_slot0; }
    __device__ __forceinline__ uint32_t word(uint32_t i) const {
        return ring_load(ring, ((i >> 2) & (RING_GROUPS - 1)) * (NT * 16) + (i & 3u) * 4u);
    }
    __device__ __forceinline__ void copy_group(uint32_t gi) {
        BVG_CP_ASYNC16(ring, (gi & (RING_GROUPS - 1)) * (NT * 16), g0 + umin32(gi, glim));
    }
    __device__ __forceinline__ void seek(const GraphDev& g, uint64_t pos) {
        uint64_t wi = pos >> 5;
        const uint64_t last = g.nwords - 1;
        if (wi > last) wi = last;  // a corrupt stream can point past the end: clamp, the caller reports E_IO
        const uint64_t gi = wi >> 2, glast = (g.nwords >> 2) - 1;   // nwords is a multiple of 4
        g0 = (const uint4*)g.words + gi;
        const uint64_t room = glast - gi;
        glim = room > 0x3fffffffull ? 0x3fffffffu : (uint32_t)room;
        BVG_CP_WAIT_ALL();  // a re-seek (slow path) must not race with copies still in flight to the same slots
        copy_group(0); copy_group(1); copy_group(2); copy_group(3);
        BVG_CP_COMMIT();
        BVG_CP_WAIT_ALL();
        fg = RING_GROUPS;
        r = (uint32_t)wi & 3u;
        s = (uint32_t)pos & 31u;
        w0 = word(r);
        w1 = word(r + 1);
        q = word(r + 2);
        r += 2;
    }
    __device__ __forceinline__ void topup() {
        if (fg - (r >> 2) <= 2u) { copy_group(fg); fg++; }
        BVG_CP_COMMIT();
        BVG_CP_WAIT_1();
    }
    __device__ __forceinline__ uint64_t pos(const GraphDev& g) const {
        return ((uint64_t)(g0 - (const uint4*)g.words) * 4u + r - 2u) * 32u + s;
    }
    __device__ __forceinline__ bool overrun() const { return (r >> 2) > glim; }
    __device__ __forceinline__ uint32_t top() const { return __funnelshift_l(w1, w0, s); }
    __device__ __forceinline__ uint64_t top64() const { return ((uint64_t)__funnelshift_l(w1, w0, s) << 32) | __funnelshift_l(q, w1, s); }
    __device__ __forceinline__ void skip(uint32_t n) {  // n <= 32
        s += n;
        if (s >= 32u) {
            s -= 32u;
            w0 = w1;
            w1 = q;
            r++;
            q = word(r);
        }
    }
    __device__ __forceinline__ uint64_t gamma_slow(const GraphDev& g) {
        uint64_t p = pos(g);
        const uint64_t v = slow_gamma(g.words, g.nwords, &p);
        seek(g, p);
        return v;
    }
    __device__ __forceinline__ uint64_t zeta_slow(const GraphDev& g, int k) {
        uint64_t p = pos(g);
        const uint64_t v = slow_zeta(g.words, g.nwords, &p, k);
        seek(g, p);
        return v;
    }
    __device__ __forceinline__ uint64_t gamma(const GraphDev& g) {
        const uint32_t t = top();
        const int m = __clz((int)t);
        if (m <= 15) {
            skip(2u * m + 1u);
            return (uint64_t)((t >> (31 - 2 * m)) - 1u);
        }
        return gamma_slow(g);
    }
};

// zeta_k code that fits the 32-bit window: m = value + 1, len = code length.  K = 3 is BVGraph's default and gets
// constants; K = 0 takes k at run time.  With h = leading zeros, P = 2^(hk): the bits after the unary part, read with
// the stop bit as r = 1:m' (hk + k bits after the leading one), give the short form iff r < P(2^k + 2).
template <int K>
__device__ __forceinline__ bool zeta_fast(uint32_t t, int k, uint32_t& m, uint32_t& len) {
    const int h = __clz((int)t);
    if (K == 3) {
        if (h > 7) return false;
        const uint32_t r = t >> (28 - 4 * h);
        const uint32_t P8 = 8u << (3 * h);          // the leading one of r
        const bool sh = 4u * r < 5u * P8;           // r < 10 P
        m = sh ? (4u * r - 3u * P8) >> 3 : r - P8;  // (r >> 1) - 3 P : r - 8 P
        len = 4u * h + (sh ? 3u : 4u);
        return true;
    } else {
        const int hk = h * k;
        const int total = h + 1 + hk + k;  // long form
        if (total > 32) return false;
        const uint32_t r = t >> (32 - total);
        const uint32_t P = 1u << hk;
        const uint32_t two_k = 1u << k;
        const bool sh = r < P * (two_k + 2u);
        m = sh ? (r >> 1) - P * ((two_k >> 1) - 1u) : r - P * two_k;
        len = (uint32_t)total - (sh ? 1u : 0u);
        return true;
    }
}

// The same on a 64-bit window (codes of up to 64 bits: every zeta_3 code of a value below 2^48).  Used for the first
// residual of a record, x +- a distance that often exceeds 2^24 on a graph without locality; the out-of-line reader
// would re-open the window, which costs a full memory round trip.
template <int K>
__device__ __forceinline__ bool zeta_fast64(uint64_t t, int k, uint64_t& m, uint32_t& len) {
    if (t == 0) return false;
    const int h = __clzll((long long)t);
    const int kk = K == 3 ? 3 : k;
    const int hk = h * kk;
    const int total = h + 1 + hk + kk;  // long form
    if (total > 64) return false;
    const uint64_t r = t >> (64 - total);
    const uint64_t P = 1ull << hk;
    const uint64_t two_k = 1ull << kk;
    const bool sh = r < P * (two_k + 2ull);
    m = sh ? (r >> 1) - P * ((two_k >> 1) - 1ull) : r - P * two_k;
    len = (uint32_t)total - (sh ? 1u : 0u);
    return true;
}

template <int K, class W>
__device__ __forceinline__ uint64_t zeta_any(W& b, const GraphDev& g, int k) {  // value + 1
    uint32_t m, len;
    if (zeta_fast<K>(b.top(), k, m, len)) { b.skip(len); return m; }
    uint64_t m64;
    if (zeta_fast64<K>(b.top64(), k, m64, len)) {
        if (len > 32u) { b.skip(32u); len -= 32u; }
        b.skip(len);
        return m64;
    }
    return b.zeta_slow(g, k) + 1ull;
}

// 32-bit halves of the checksum of one record: XOR of lo(base + y), number of carries out of the low word.
// HIST: the fused in-degree count of bvg_indegrees (GraphDev.hist) -- a compile-time variant, so that the plain scan's loops
// carry neither the pointer nor the test (a run-time test cost the residual loop 20 %: 2.72 -> 3.26 ms)
template <bool HIST>
struct FoldT {
    uint32_t base_lo, xlo, carries, n;
    uint32_t* hist;
    uint32_t hist_len;
    __device__ __forceinline__ void begin(int32_t x) { base_lo = (uint32_t)((unsigned long long)(uint32_t)x * BVG_MIX); xlo = 0; carries = 0; n = 0; if (HIST) { hist = nullptr; hist_len = 0; } }
    // counted: the successors added here are consumed by the scan (not a halo record, not a fold that a later step repeats on
    // the final list): only then do they count in the fused histogram
    __device__ __forceinline__ void begin(int32_t x, const GraphDev& g, bool counted) {
        begin(x);
        if (HIST && counted) { hist = g.hist; hist_len = (uint32_t)(g.hist_len > 0xffffffffll ? 0xffffffffll : g.hist_len); }
    }
    __device__ __forceinline__ void add(uint32_t y) {
        const uint32_t lo = base_lo + y;
        carries += lo < y ? 1u : 0u;
        xlo ^= lo;
        if (HIST) { if (hist != nullptr && y < hist_len) atomicAdd(hist + y, 1u); }
    }
    // XOR over the n folded successors of (x*MIX + y): hi word is base_hi for the ones without a carry, base_hi + 1 with
    __device__ __forceinline__ unsigned long long finish(int32_t x) const {
        const uint32_t base_hi = (uint32_t)(((unsigned long long)(uint32_t)x * BVG_MIX) >> 32);
        uint32_t hi = 0;
        if ((n - carries) & 1u) hi ^= base_hi;
        if (carries & 1u) hi ^= base_hi + 1u;
        return ((unsigned long long)hi << 32) | xlo;
    }
};
typedef FoldT<false> Fold32;


// Sequential writer of one lane's row with 16-byte write combining.  The 32 lanes of a warp write 32 unrelated rows, so
// a 4-byte store per successor is 32 separate sector writes per instruction at the L2; measured on the benchmark graph
// they cost 37 % of the extras kernel (3.92 ms with, 2.48 ms without the row stores).  The writer keeps the last three
// values in registers and issues one 16-byte store whenever the write position reaches a 16-byte boundary; the unaligned
// head and tail of a row (at most three values each) are stored singly.
template <bool VEC>
struct RowWriter {
    int32_t* __restrict__ p;   // next position to write
    uint32_t b1, b2, b3;       // the last three values put (b3 newest)
    uint32_t pend;             // values put but not yet stored, 0..3
    uint32_t ph;               // (address of p / 4) & 3
    __device__ __forceinline__ void begin(int32_t* row) {
        p = row; b1 = b2 = b3 = 0; pend = 0;
        ph = (uint32_t)(((uintptr_t)row) >> 2) & 3u;
    }
    __device__ __forceinline__ void put(uint32_t v) {
        if (!VEC) { *p++ = (int32_t)v; return; }
        p++;
        ph = (ph + 1u) & 3u;
        if (ph == 0u) {  // p is 16-byte aligned again: the four values before it are one aligned group
            if (pend == 3u) {
#ifdef BVG_HOST_EMULATION
                p[-4] = (int32_t)b1; p[-3] = (int32_t)b2; p[-2] = (int32_t)b3; p[-1] = (int32_t)v;
#else
                *reinterpret_cast<uint4*>(p - 4) = make_uint4(b1, b2, b3, v);
#endif
            } else {       // head of the row: fewer than four values since it started
                p[-1] = (int32_t)v;
                if (pend >= 1u) p[-2] = (int32_t)b3;
                if (pend >= 2u) p[-3] = (int32_t)b2;
            }
            pend = 0;
        } else pend++;
        b1 = b2; b2 = b3; b3 = v;
    }
    __device__ __forceinline__ void flush() {  // tail of the row
        if (!VEC) return;
        if (pend >= 1u) p[-1] = (int32_t)b3;
        if (pend >= 2u) p[-2] = (int32_t)b2;
        if (pend >= 3u) p[-3] = (int32_t)b1;
        pend = 0;
    }
};

// ---------------------------------------------------------------------------------------------------
// Extras of one record (everything that is not copied), consume-only or into the row when `store`.
// Entry: the cursor at the extras section (ExtraRec.pos), nout = outdegree - copied.
//   phase iv()      interval section; fold-only records fold the elements here, storing records take the merge of
//                   ExtrasWalk::with_intervals instead (kept in separate warps by the schedule)
//   phase resid()   residuals, tight loop
// ---------------------------------------------------------------------------------------------------
template <int K, class W = Win, bool HIST = false>
struct ScanExtras {
    W b;
    FoldT<HIST> f;
    int32_t x, nout, rc;
    uint32_t v;
    int err;
    uint32_t ic;       // intervals (iv_fold)
    uint64_t iv_pos;   // bit position of the first interval's left extreme

    __device__ __forceinline__ void fail(const GraphDev& g, int code) {
        report(g.err, code, x, b.pos(g) + g.bit_base);
        err = code; rc = 0;
    }

    __device__ __forceinline__ void begin(const GraphDev& g, int32_t x_, int32_t nout_, uint64_t pos, bool active, ring_addr ring_slot = ring_addr(),
                                          bool counted = true) {
        x = x_; nout = 0; rc = 0; err = 0; v = 0; ic = 0; iv_pos = 0;
        f.begin(x_, g, counted);
        b.attach(ring_slot);
        if (!active) return;
        nout = nout_;
        rc = nout_;
        b.seek(g, pos);
    }

    // Interval section of a record that is only consumed (IntIntervalSequenceIterator.java:57-95, BVGraph.java:1076-1095)
    __device__ __forceinline__ void iv_fold(const GraphDev& g) {
        if (nout <= 0 || g.c.minlen == 0) return;
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
            const uint64_t len64 = b.gamma(g) + (uint64_t)g.c.minlen;
            total += (int64_t)len64;
            if (total > (int64_t)nout || b.overrun()) { fail(g, E_IO); return; }
            const uint32_t len = (uint32_t)len64;
            for (uint32_t j = 0; j < len; j++) f.add(left + j);
            prev = left + len;
            b.topup();
        }
        rc = nout - (int32_t)total;
    }

    // Stored records with intervals: the residuals were written right-aligned, row[nout - rc .. nout); walk the interval
    // section a second time and merge, forward and in place (Merged(IntIntervalSequenceIterator, ResidualIntIterator),
    // BVGraph.java:1103-1108; equal heads once, MergedIntIterator.java:70; a list that loses duplicates is padded with -1
    // as BVGraphNodeIterator does when it drains, :1210).  The write index never overtakes the unread residuals: it trails
    // them by the number of interval elements still to come.
    __device__ __forceinline__ void iv_merge(const GraphDev& g, int32_t* row) {
        if (ic == 0 || err) return;
        Win c;
        c.seek(g, iv_pos);
        int32_t k = 0, j = nout - rc;
        RowWriter<true> wr;
        wr.begin(row);
        uint32_t prev = 0;
        for (uint32_t i = 0; i < ic; i++) {
            uint32_t left;
            if (i == 0) left = (uint32_t)(int32_t)(nat2int(c.gamma(g)) + (int64_t)x);
            else left = prev + 1u + (uint32_t)c.gamma(g);
            const uint32_t len = (uint32_t)(c.gamma(g) + (uint64_t)g.c.minlen);
            while (j < nout && (uint32_t)row[j] < left) { wr.put((uint32_t)row[j++]); k++; }
            for (uint32_t e = 0; e < len; e++) { wr.put(left + e); k++; }
            prev = left + len;
            while (j < nout && (uint32_t)row[j] < prev) j++;  // residuals inside an interval: emitted once
        }
        if (k != j) {
            while (j < nout) { wr.put((uint32_t)row[j++]); k++; }
            while (k < nout) { wr.put(0xffffffffu); k++; }
        }
        wr.flush();
    }

    // Interval section skipped over (storing records without intervals still carry the count)
    __device__ __forceinline__ void iv_none(const GraphDev& g) {
        if (nout <= 0 || g.c.minlen == 0) return;
        const uint64_t ic = b.gamma(g);
        if (ic != 0) fail(g, E_FORMAT);  // the schedule said "no intervals"
    }

    // Residuals (ResidualIntIterator, BVGraph.java:939-972): first = x + nat2int(zeta), then += zeta + 1.
    template <bool STORE>
    __device__ __forceinline__ void resid(const GraphDev& g, int32_t* __restrict__ row, bool store) {
        if (rc <= 0) return;
        const int k = g.c.zetak;
        b.topup();
        v = (uint32_t)(int32_t)((int64_t)x + nat2int(zeta_any<K>(b, g, k) - 1ull));  // :954
        f.add(v);
        RowWriter<true> wr;
        if (STORE) wr.begin(row + (nout - rc));  // after the interval elements a later iv_merge puts in front (rc == nout without intervals)
        if (STORE && store) wr.put(v);
#pragma unroll 1
        for (int32_t i = 1; i < rc; i++) {
            if ((((uint32_t)i + 1u) & 3u) == 0u) b.topup();  // the first residual may have taken two words: 2 + 2, then every 4 codes
            uint32_t m, len;
            if (zeta_fast<K>(b.top(), k, m, len)) b.skip(len);
            else { b.topup(); m = (uint32_t)zeta_any<K>(b, g, k); b.topup(); }  // gap >= 2^24: up to two words, between two out-of-turn topups
            v += m;  // :966 (gap + 1)
            f.add(v);
            if (STORE && store) wr.put(v);
        }
        if (STORE && store) wr.flush();
        if (b.overrun() || b.pos(g) > g.bit_end - g.bit_base) fail(g, E_IO);
    }

    __device__ __forceinline__ unsigned long long finish() {
        f.n = (uint32_t)nout;
        return err ? 0ull : f.finish(x);
    }
};

// ---------------------------------------------------------------------------------------------------
// Copied part of one record (MaskedIntIterator.java:65-97): the copy-block list is parsed first, a few blocks at a
// time, into a short list of non-empty copy runs [start, end) of parent positions kept in shared memory; the element
// loops then only compare an index with the end of the current run.  With the parse inside the element loop (the flat
// loop of MergeWalk) every trip of a warp pays for a gamma decode as soon as one of its 32 lanes sits at a block
// boundary, which is almost always (ncu: 2.2 active lanes on the boundary path, 20 % of the kernel's instructions).
// ---------------------------------------------------------------------------------------------------
constexpr int COPY_RUNS = 8;  // runs staged per lane at a time (16 blocks); longer lists are staged in rounds

template <int NR>
struct CopyRunsT {
    Win b;
    int32_t* __restrict__ st;   // this lane's slots: start of run i at st[2 * i * stride], end at st[(2 * i + 1) * stride]
    int32_t stride;
    uint32_t bc, bi;            // blocks in the list, blocks parsed (bi == bc + 1: the implicit tail has been handled too)
    uint32_t ppos, dp;          // parent position after the parsed blocks, parent outdegree
    int32_t nr, r;              // runs staged, next run to take
    uint32_t pos, end;          // current run

    __device__ __forceinline__ void begin(const GraphDev& g, uint64_t recpos, int32_t bc_, int32_t dp_, int32_t* slots, int32_t stride_, bool active) {
        st = slots; stride = stride_;
        bc = (uint32_t)bc_; bi = 0; ppos = 0; dp = (uint32_t)dp_; nr = 0; r = 0; pos = 0; end = 0;
        if (!active) { bi = 1; bc = 0; dp = 0; return; }  // nothing to stage, nothing to copy
        if (bc) b.seek(g, recpos);
    }
    // Parses blocks until COPY_RUNS runs are staged or the list (and its implicit tail: an even number of blocks copies
    // the rest of the parent, MaskedIntIterator.java:76) is exhausted.  Positions are clamped to the parent's outdegree,
    // so a malformed list can make the result wrong but never the reads out of bounds.
    __device__ __forceinline__ void stage(const GraphDev& g) {
        nr = 0; r = 0;
        while (nr < NR && bi <= bc) {
            if (bi == bc) {
                if (!(bc & 1u) && ppos < dp) { st[2 * nr * stride] = (int32_t)ppos; st[(2 * nr + 1) * stride] = (int32_t)dp; nr++; ppos = dp; }
                bi++;
                break;
            }
            const uint64_t raw = b.gamma(g);
            uint32_t len = raw > 0x7fffffffull ? 0x7fffffffu : (uint32_t)raw;
            if (bi) len++;
            const uint32_t e = umin32(ppos + len, dp);
            if (!(bi & 1u) && e > ppos) { st[2 * nr * stride] = (int32_t)ppos; st[(2 * nr + 1) * stride] = (int32_t)e; nr++; }
            ppos = e;
            bi++;
        }
    }
    __device__ __forceinline__ bool done() const { return pos == end && r == nr && bi > bc; }
    // next parent position to copy; false when the list is exhausted
    __device__ __forceinline__ bool next(const GraphDev& g, uint32_t& at) {
        if (pos == end) {
            if (r == nr) {
                if (bi > bc) return false;
                stage(g);
                if (nr == 0) return false;
            }
            pos = (uint32_t)st[2 * r * stride];
            end = (uint32_t)st[(2 * r + 1) * stride];
            r++;
        }
        at = pos++;
        return true;
    }
};
typedef CopyRunsT<COPY_RUNS> CopyRuns;

// nobody copies from x: its copied successors are only consumed
template <int BATCH, class CR, bool HIST = false>
__device__ __forceinline__ unsigned long long copied_fold(const GraphDev& g, CR& c, int32_t x, const int32_t* __restrict__ parent) {
    FoldT<HIST> f;
    f.begin(x, g, true);
    // BATCH positions first, then their loads together: a lane opens a new sector of its parent's row every eighth element,
    // the 32 lanes read 32 unrelated rows, and with one load per trip the warp waits a memory round trip on every trip
#pragma unroll 1
    for (;;) {
        uint32_t at[BATCH], val[BATCH];
        bool has[BATCH];
        bool more = true;
#pragma unroll
        for (int u = 0; u < BATCH; u++) { at[u] = 0; has[u] = more && c.next(g, at[u]); more = has[u]; }
#pragma unroll
        for (int u = 0; u < BATCH; u++) val[u] = has[u] ? (uint32_t)parent[at[u]] : 0u;
#pragma unroll
        for (int u = 0; u < BATCH; u++) if (has[u]) { f.add(val[u]); f.n++; }
        if (!more) break;
    }
    return f.finish(x);
}

// somebody copies from x: merge the copied successors, forward and in place, with the extras at row[copied .. d)
// (MergedIntIterator.java:50-74: ascending union, equal heads once; a list that loses duplicates is padded with -1 as
// BVGraphNodeIterator does when it drains, BVGraph.java:1210).  Folds the copied successors only: the extras were
// folded when they were decoded.
template <class CR, bool HIST = false>
__device__ __forceinline__ unsigned long long copied_merge(const GraphDev& g, CR& c, int32_t x, int32_t d, int32_t copied,
                                                           int32_t* row, const int32_t* __restrict__ parent, bool counted = true) {
    FoldT<HIST> f;
    f.begin(x, g, counted);
    int32_t j = copied, k = 0;
    RowWriter<false> wr;  // measured: combining helps the extras kernel (3.92 -> 3.71 ms) and costs registers here (2.22 -> 2.49 ms)
    wr.begin(row);
    uint32_t at;
    bool have_a = c.next(g, at);
    uint32_t a = have_a ? (uint32_t)parent[at] : 0xffffffffu;
    uint32_t bv = j < d ? (uint32_t)row[j] : 0xffffffffu;
#pragma unroll 1
    while (k < d) {
        if (!have_a) {
            if (k == j) break;  // nothing was dropped: the remaining extras already sit in place
            if (j < d) { wr.put(bv); k++; j++; bv = j < d ? (uint32_t)row[j] : 0xffffffffu; continue; }
            wr.put(0xffffffffu); k++;  // duplicates were dropped
            continue;
        }
        if (a <= bv) {
            wr.put(a); k++;
            f.add(a); f.n++;
            if (a == bv) { j++; bv = j < d ? (uint32_t)row[j] : 0xffffffffu; }  // equal heads are emitted once
            have_a = c.next(g, at);
            if (have_a) a = (uint32_t)parent[at];
        } else {
            wr.put(bv); k++;
            j++;
            bv = j < d ? (uint32_t)row[j] : 0xffffffffu;
        }
    }
    wr.flush();
    return f.finish(x);
}

#ifndef BVG_HOST_EMULATION
// Blocks of the scan kernels fold into FOLD_SLOTS slot pairs (arcs, XOR) instead of two hot words; k_reduce_slots sums them.
constexpr int FOLD_SLOTS = 1024;

// The same without a block barrier: one atomic pair per warp (blocks of the dynamic-grid scan kernels live for one
// item per thread; a barrier at the end makes every warp wait for the block's longest record).
__device__ __forceinline__ void warp_fold(unsigned long long acc, long long arcs, unsigned long long* __restrict__ result) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
        arcs += __shfl_xor_sync(0xffffffffu, arcs, o);
    }
    if ((threadIdx.x & 31) == 0) {
        unsigned long long* slot = result + 2 * ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (FOLD_SLOTS - 1));
        if (acc) atomicXor(slot + 1, acc);
        if (arcs) atomicAdd(slot, (unsigned long long)arcs);
    }
}

#endif

}  // namespace bvg
