// bvg_device.cuh -- device-side bit reader, universal codes and per-record decode steps.
//
// Re-designs, for one GPU thread walking one record, what the reference does with
// it.unimi.dsi.io.InputBitStream + BVGraph.successors(x, ibs, window, outd) (reference
// src/it/unimi/dsi/webgraph/BVGraph.java:1032-1133) and its lazy iterators (MaskedIntIterator.java:65-97,
// IntIntervalSequenceIterator.java:57-95, MergedIntIterator.java:50-74, ResidualIntIterator BVGraph.java:939-991).
// There are no iterators here: a record is decoded in two streaming steps that write straight into the node's
// row of the output: (1) "extras" = intervals U residuals, merged on the fly from two bit cursors, written to the
// tail of the row; (2) the copied part, streamed from the parent's finished row through the copy-block list and
// merged forward, in place, with that tail.
//
// Layout: the .graph stream is held in HBM as 32-bit words in big-endian bit order (word i = stream bits
// [32i, 32i+32), MSB first), so a 64-bit window at any bit position is two funnel shifts over three words.
#pragma once
#include <cstdint>
#ifdef BVG_HOST_EMULATION   // tests/hostemu: the same logic compiled for the host, for sanitizer runs
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif
#ifdef BVG_HOST_EMULATION
#define BVG_NOINLINE inline
#else
#define BVG_NOINLINE __noinline__
#endif

namespace bvg {

// Zero words kept after the last stream word: readers look at most 5 words past a record.
constexpr int STREAM_PAD_WORDS = 8;

enum { C_DELTA = 1, C_GAMMA = 2, C_GOLOMB = 3, C_SKEWED_GOLOMB = 4, C_UNARY = 5, C_ZETA = 6, C_NIBBLE = 7 };

enum { E_OK = 0, E_INVAL = -1, E_STATE = -2, E_UNSUPPORTED = -3, E_IO = -4, E_FORMAT = -5, E_NOMEM = -6, E_CUDA = -7, E_END = -8 };

// Device-side error word: first failing record wins (the reference throws with node + bit position,
// BVGraph.java:705, 1129-1131).
struct ErrWord {
    int code;       // 0 = none
    int node;
    long long bitpos;
};

struct Codec {
    int outdeg, block, resid, ref, bcount;  // coding ids, BVGraph.java:525-541
    int zetak, window, minlen;
};

// Everything a kernel needs to decode nodes [node_lo, node_hi) held by this graph object.
struct GraphDev {
    const uint32_t* __restrict__ words;   // stream words, word 0 = stream bit `bit_base`
    uint64_t nwords;                      // incl. >= 4 padding words
    uint64_t bit_base;                    // global bit position of word 0 (multiple of 128)
    uint64_t bit_end;                     // global bit position one past the last loaded record
    const uint64_t* __restrict__ offsets; // offsets[x - node_lo], global bit positions, node_hi - node_lo + 1 entries
    int32_t node_lo, node_hi;
    Codec c;
    // decode index (built at open by k_header / scan / k_depth)
    const int32_t* __restrict__ outdeg;   // [x - node_lo]
    const int32_t* __restrict__ ref;      // [x - node_lo]  0 = no reference
    const int32_t* __restrict__ depth;    // [x - node_lo]  reference-chain depth, -1 = chain leaves the loaded window
    const int64_t* __restrict__ rowoff;   // [x - node_lo]  cumulative outdegree, node_hi - node_lo + 1 entries
    const int32_t* __restrict__ copied;   // [x - node_lo]  successors copied from the parent (k_order_keys); may be null
    ErrWord* err;
    // fused consumer of a scan (bvg_indegrees): when set, every successor y the scan folds also counts in hist[y]
    // (the transposition counting pass, reference Transform.java:977-987: numPred[a[d]]++)
    uint32_t* hist;
    int64_t hist_len;
};

__device__ __forceinline__ void report(ErrWord* e, int code, int node, uint64_t bitpos) {
    if (atomicCAS(&e->code, 0, code) == 0) { e->node = node; e->bitpos = (long long)bitpos; }
}

// ---------------------------------------------------------------------------------------------------
// Bit cursor over global/shared memory words.
// ---------------------------------------------------------------------------------------------------
struct Bits {
    const uint32_t* __restrict__ w;
    uint64_t maxw;   // last index for which w[i+2] is readable
    uint64_t pos;    // bit position relative to w[0]

    __device__ __forceinline__ uint64_t peek() const {
        uint64_t i = pos >> 5;
        i = i < maxw ? i : maxw;  // a corrupt stream can run past the end: clamp, the caller reports E_IO
        const uint32_t s = (uint32_t)pos & 31u;
        const uint32_t w0 = w[i], w1 = w[i + 1], w2 = w[i + 2];
        return ((uint64_t)__funnelshift_l(w1, w0, s) << 32) | (uint64_t)__funnelshift_l(w2, w1, s);
    }
    // n bits, MSB first, 0 <= n <= 64 (InputBitStream.readInt/readLong)
    __device__ __forceinline__ uint64_t bits(int n) {
        if (n == 0) return 0;
        const uint64_t v = peek() >> (64 - n);
        pos += (uint64_t)n;
        return v;
    }
    // number of zeros before the first one (InputBitStream.readUnary)
    __device__ __forceinline__ uint64_t unary() {
        uint64_t zeros = 0;
        for (;;) {
            const uint64_t v = peek();
            if (v == 0) {
                zeros += 64; pos += 64;
                if ((pos >> 5) > maxw + 4) return zeros;
                continue;
            }
            const int z = __clzll((long long)v);
            pos += (uint64_t)z + 1;
            return zeros + (uint64_t)z;
        }
    }
    // gamma: unary(msb) then the msb low bits of x+1. One window when the code fits 64 bits (x < 2^32 - 1).
    __device__ __forceinline__ uint64_t gamma() {
        const uint64_t v = peek();
        const int m = __clzll((long long)v);
        if (m < 32) {  // 2m+1 <= 63 bits: m zeros, a one, m bits == x+1 right-aligned
            pos += (uint64_t)(2 * m + 1);
            return (v >> (63 - 2 * m)) - 1;
        }
        const uint64_t msb = unary();
        if (msb > 63) return ~0ull;
        return ((1ull << msb) | bits((int)msb)) - 1;
    }
    __device__ __forceinline__ uint64_t delta() {
        const uint64_t msb = gamma();
        if (msb > 63) return ~0ull;
        return ((1ull << msb) | bits((int)msb)) - 1;
    }
    // zeta_k: unary(h), then the minimal binary code of x+1-2^{hk} over [0, 2^{(h+1)k} - 2^{hk})
    __device__ __forceinline__ uint64_t zeta(int k) {
        const uint64_t v = peek();
        const int h = __clzll((long long)v);
        const int nb = h * k + k - 1;
        if (v != 0 && h + 1 + nb + 1 <= 64) {
            const uint64_t left = 1ull << (h * k);
            const uint64_t t = v << (h + 1);
            uint64_t m = nb ? t >> (64 - nb) : 0;
            int len = h + 1 + nb;
            if (m < left) m += left;
            else { m = (m << 1) | ((t >> (63 - nb)) & 1ull); len++; }
            pos += (uint64_t)len;
            return m - 1;
        }
        const uint64_t hh = unary();
        if (hh * (uint64_t)k + (uint64_t)k > 64) return ~0ull;
        const uint64_t left = 1ull << (hh * k);
        const uint64_t m = bits((int)(hh * k) + k - 1);
        if (m < left) return m + left - 1;
        return ((m << 1) | bits(1)) - 1;
    }
    // Golomb with modulus b (InputBitStream.readGolomb(b), dsiutils; BVGraph passes zetaK as the modulus,
    // BVGraph.java:796, 812): unary(x / b), then x % b in minimal binary -- with l = msb(b) and m = 2^(l+1) - b,
    // remainders below m take l bits, the others l + 1 bits holding (remainder + m).  b == 0 reads nothing.
    __device__ BVG_NOINLINE uint64_t golomb(int b) {
        if (b <= 0) return 0;
        const uint64_t q = unary();
        const int l = 31 - __clz(b);
        const uint64_t m = (2ull << l) - (uint64_t)b;
        uint64_t r = bits(l);
        if (r >= m) r = ((r << 1) | bits(1)) - m;
        return q * (uint64_t)b + r;
    }
    // Nibble code (InputBitStream.readNibble): 4-bit groups, MSB group first, each a stop flag (1 = last group)
    // followed by three bits of the value.  At most 22 groups make a 64-bit value; more is a corrupt stream.
    __device__ BVG_NOINLINE uint64_t nibble() {
        uint64_t x = 0;
        for (int i = 0; i < 22; i++) {
            const uint64_t grp = bits(4);
            x = (x << 3) | (grp & 7ull);
            if (grp & 8ull) return x;
        }
        return ~0ull;
    }
    __device__ __forceinline__ uint64_t coded(int coding, int k) {
        switch (coding) {
            case C_GAMMA:  return gamma();
            case C_DELTA:  return delta();
            case C_UNARY:  return unary();
            case C_GOLOMB: return golomb(k);
            case C_NIBBLE: return nibble();
            default:       return zeta(k);
        }
    }
};


// ---------------------------------------------------------------------------------------------------
// Register bit buffer: the next 33..64 stream bits live in a register pair, refilled one 32-bit word at a time, so a
// code costs a count-leading-zeros, a few shifts and (amortised) 0.6 loads instead of three loads per code.
// Codes that do not fit the buffered bits (gaps >= 2^24, long gammas) fall back to the position-based reader.
// ---------------------------------------------------------------------------------------------------
struct BitBuf {
    const uint32_t* __restrict__ w;
    uint64_t maxw;   // as in Bits: w[maxw + 2] is the last readable word
    uint64_t widx;   // next word to enter the window
    uint64_t buf;    // MSB-aligned window; bits beyond `avail` are zero
    uint32_t q0, q1; // w[widx], w[widx + 1]: loaded two refills before they are used, so that the (compulsory)
    int avail;       // sector miss every eighth word overlaps with decoding instead of stalling the warp

    __device__ __forceinline__ uint32_t word(uint64_t i) const { return w[i < maxw + 2 ? i : maxw + 2]; }
    __device__ __forceinline__ void seek(uint64_t pos) {
        uint64_t i = pos >> 5;
        i = i < maxw ? i : maxw;
        const uint32_t s = (uint32_t)pos & 31u;
        buf = (((uint64_t)w[i] << 32) | (uint64_t)w[i + 1]) << s;
        avail = 64 - (int)s;
        widx = i + 2;
        q0 = word(widx);
        q1 = word(widx + 1);
    }
    __device__ __forceinline__ uint64_t pos() const { return widx * 32 - (uint64_t)avail; }
    __device__ __forceinline__ void consume(int n) {  // 0 <= n <= 63, n <= avail
        buf <<= n;
        avail -= n;
        if (avail <= 32) {
            buf |= (uint64_t)q0 << (32 - avail);
            avail += 32;
            q0 = q1;
            q1 = word(widx + 2);
            widx++;
        }
    }
    template <class F>
    __device__ __forceinline__ uint64_t slow(F f) {
        Bits t;
        t.w = w; t.maxw = maxw; t.pos = pos();
        const uint64_t r = f(t);
        seek(t.pos);
        return r;
    }
    __device__ __forceinline__ uint64_t unary() {
        const int z = __clzll((long long)buf);
        if (z < avail && z < 63) { consume(z + 1); return (uint64_t)z; }
        return slow([](Bits& t) { return t.unary(); });
    }
    __device__ __forceinline__ uint64_t gamma() {
        const uint64_t v = buf;
        const int m = __clzll((long long)v);
        if (m < 32 && 2 * m + 1 <= avail) {
            consume(2 * m + 1);
            return (v >> (63 - 2 * m)) - 1;
        }
        return slow([](Bits& t) { return t.gamma(); });
    }
    __device__ __forceinline__ uint64_t delta() { return slow([](Bits& t) { return t.delta(); }); }
    // zeta_k gap that fits 32 bits of code (k = 3: gaps < 2^24): all-32-bit arithmetic on the top half of the buffer.
    // Returns false (nothing consumed) when the code is longer; the caller then takes zeta().
    __device__ __forceinline__ bool zeta32(int k, uint32_t& out) {
        const uint32_t hi = (uint32_t)(buf >> 32);  // avail > 32 always holds
        const int h = __clz((int)hi);
        const int nb = h * k + k - 1;
        if (h + nb + 2 > 32) return false;
        const uint32_t left = 1u << (h * k);
        const uint32_t t = hi << (h + 1);
        uint32_t m = nb ? t >> (32 - nb) : 0;
        int len = h + 1 + nb;
        if (m < left) m += left;
        else { m = (m << 1) | ((t >> (31 - nb)) & 1u); len++; }
        consume(len);
        out = m - 1;
        return true;
    }
    __device__ __forceinline__ uint64_t zeta(int k) {
        const uint64_t v = buf;
        const int h = __clzll((long long)v);
        const int nb = h * k + k - 1;
        const int maxlen = h + 1 + nb + 1;
        if (maxlen <= avail && maxlen <= 63) {
            const uint64_t left = 1ull << (h * k);
            const uint64_t t = v << (h + 1);
            uint64_t m = nb ? t >> (64 - nb) : 0;
            int len = h + 1 + nb;
            if (m < left) m += left;
            else { m = (m << 1) | ((t >> (63 - nb)) & 1ull); len++; }
            consume(len);
            return m - 1;
        }
        return slow([k](Bits& t) { return t.zeta(k); });
    }
    __device__ __forceinline__ uint64_t coded(int coding, int k) {
        switch (coding) {
            case C_GAMMA:  return gamma();
            case C_DELTA:  return delta();
            case C_UNARY:  return unary();
            case C_GOLOMB: return slow([k](Bits& t) { return t.golomb(k); });
            case C_NIBBLE: return slow([](Bits& t) { return t.nibble(); });
            default:       return zeta(k);
        }
    }
};

// Fast.nat2int (dsiutils): even -> v/2, odd -> -(v+1)/2
__device__ __forceinline__ int64_t nat2int(uint64_t v) {
    return (v & 1) ? -(int64_t)((v + 1) >> 1) : (int64_t)(v >> 1);
}

// Coding policy: DEF = true hard-wires the defaults (gamma outdegrees/blocks/block counts, unary references,
// zeta residuals; BVGraph.java:525-541) so the switch disappears; k stays a runtime value.
template <bool DEF>
struct Rd {
    template <class B> static __device__ __forceinline__ uint64_t outdeg(B& b, const Codec& c) { return DEF ? b.gamma() : b.coded(c.outdeg, 0); }
    template <class B> static __device__ __forceinline__ uint64_t ref(B& b, const Codec& c)    { return DEF ? b.unary() : b.coded(c.ref, 0); }
    template <class B> static __device__ __forceinline__ uint64_t bcount(B& b, const Codec& c) { return DEF ? b.gamma() : b.coded(c.bcount, 0); }
    template <class B> static __device__ __forceinline__ uint64_t block(B& b, const Codec& c)  { return DEF ? b.gamma() : b.coded(c.block, 0); }
    template <class B> static __device__ __forceinline__ uint64_t resid(B& b, const Codec& c)  { return DEF ? b.zeta(c.zetak) : b.coded(c.resid, c.zetak); }
    // gap between consecutive residuals (fits 32 bits in any valid file: successors are int32)
    static __device__ __forceinline__ uint32_t gap(BitBuf& b, const Codec& c) {
        if (DEF) { uint32_t v; if (b.zeta32(c.zetak, v)) return v; }
        return (uint32_t)resid(b, c);
    }
};

__device__ __forceinline__ Bits cursor_at(const GraphDev& g, int32_t x) {
    Bits b;
    b.w = g.words;
    b.maxw = g.nwords - 3;
    b.pos = g.offsets[x - g.node_lo] - g.bit_base;
    return b;
}

__device__ __forceinline__ BitBuf buffer_at(const GraphDev& g, int32_t x) {
    BitBuf b;
    b.w = g.words;
    b.maxw = g.nwords - 3;
    b.seek(g.offsets[x - g.node_lo] - g.bit_base);
    return b;
}

#define BVG_INF 0x7fffffffffffffffll

// ---------------------------------------------------------------------------------------------------
// Step 1: outdegree, reference, copy-block totals, then intervals U residuals merged into row[copied .. d).
// BVGraph.java:1044-1100; the union order is MergedIntIterator's (equal heads once, :70).
//
// The walk is split into three phases with a single exit each so that a kernel can put a __syncwarp() between them:
// with early returns inside divergent branches the compiler's reconvergence point moves to the end of the function and
// the lanes of a warp, once apart, walk their residual loops one after the other instead of in lockstep.
//   header()          outdegree, reference, block list (copied count), interval section located, counts
//   residuals_only()  records without intervals: a tight loop in 32-bit arithmetic
//   with_intervals()  records with intervals: intervals and residuals merged on the fly from two bit cursors
// ---------------------------------------------------------------------------------------------------
#define BVG_MIX 0x9E3779B97F4A7C15ull

template <bool DEF>
struct ExtrasWalk {
    BitBuf b;            // residual cursor (after header())
    Bits ib;             // interval cursor (position-based: intervals are few, registers are not)
    int32_t x;
    int32_t d, copied;   // outdegree, successors copied from the parent
    int32_t ic, rc;      // intervals, residuals
    int32_t nout;        // d - copied: entries this step produces
    int err;             // 0 or a negative status; an erroneous or empty record has nout == 0

    __device__ __forceinline__ void fail(const GraphDev& g, int code) {
        report(g.err, code, x, b.pos() + g.bit_base);
        err = code; nout = 0; ic = 0; rc = 0;
    }

    __device__ __forceinline__ void header(const GraphDev& g, int32_t x_, bool active) {
        const Codec& c = g.c;
        x = x_; d = 0; copied = 0; ic = 0; rc = 0; nout = 0; err = 0;
        if (!active) return;
        b = buffer_at(g, x);
        ib.w = g.words; ib.maxw = g.nwords - 3; ib.pos = 0;
        const uint64_t limit = g.bit_end - g.bit_base;
        const uint64_t d64 = Rd<DEF>::outdeg(b, c);
        if (d64 > 0x7fffffffull || b.pos() > limit) { fail(g, E_IO); return; }
        d = (int32_t)d64;
        int64_t cp = 0;
        bool ok = d != 0;
        if (ok && c.window > 0) {
            const uint64_t r = Rd<DEF>::ref(b, c);
            if (r > (uint64_t)c.window) { fail(g, E_STATE); ok = false; }  // :705
            else if (r > 0) {
                if ((int64_t)r > (int64_t)x - g.node_lo) { fail(g, E_FORMAT); ok = false; }
                else {
                    const uint64_t bc = Rd<DEF>::bcount(b, c);
                    int64_t total = 0;
                    if (bc > 0x7fffffffull) { fail(g, E_IO); ok = false; }
                    for (uint64_t i = 0; ok && i < bc; i++) {  // :1062-1066
                        const int64_t blk = (int64_t)Rd<DEF>::block(b, c) + (i ? 1 : 0);
                        total += blk;
                        if (!(i & 1)) cp += blk;
                        if (b.pos() > limit) { fail(g, E_IO); ok = false; }
                    }
                    if (ok) {
                        const int64_t dp = g.outdeg[x - (int32_t)r - g.node_lo];
                        if (!(bc & 1)) cp += dp - total;  // :1069
                        if (total > dp || cp < 0 || cp > d) { fail(g, E_FORMAT); ok = false; }
                    }
                }
            }
        }
        if (!ok) return;
        copied = (int32_t)cp;
        nout = d - copied;
        sections(g);
    }

    // Entry from a schedule record (bvg_kernels.cuh, ExtraRec): outdegree, reference and block list were parsed at
    // open; the cursor starts at the extras section and the row pointer the caller passes is already past the copied part.
    __device__ __forceinline__ void header_rec(const GraphDev& g, int32_t x_, int32_t d_, int32_t nout_, uint64_t pos, bool active) {
        x = x_; d = 0; copied = 0; ic = 0; rc = 0; nout = 0; err = 0;
        if (!active) return;
        d = d_;
        nout = nout_;
        b.w = g.words; b.maxw = g.nwords - 3;
        b.seek(pos);
        ib.w = g.words; ib.maxw = g.nwords - 3; ib.pos = 0;
        sections(g);
    }

    // interval section located, residual count (:1076-1096)
    __device__ __forceinline__ void sections(const GraphDev& g) {
        const Codec& c = g.c;
        const uint64_t limit = g.bit_end - g.bit_base;
        int64_t extra = nout;
        bool ok = true;
        if (extra == 0) return;
        // interval section: remember where it starts, walk it once to find the residual section
        if (c.minlen != 0) {
            const int64_t n_iv = (int64_t)b.gamma();
            if (n_iv > extra || b.pos() > limit) { fail(g, E_IO); return; }
            ib.pos = b.pos();
            int64_t tot = 0;
            for (int64_t i = 0; ok && i < n_iv; i++) {
                (void)b.gamma();
                tot += (int64_t)b.gamma() + c.minlen;
                if (b.pos() > limit || tot > extra) { fail(g, E_IO); ok = false; }
            }
            if (!ok) return;
            ic = (int32_t)n_iv;
            extra -= tot;
        }
        rc = (int32_t)extra;
    }

    // Records without intervals (ResidualIntIterator, :939-972).
    template <bool FOLD>
    __device__ __forceinline__ void residuals_only(const GraphDev& g, int32_t* __restrict__ row, bool store, unsigned long long& fold) {
        if (ic != 0 || rc <= 0) return;
        const Codec& c = g.c;
        const unsigned long long fold_base = (unsigned long long)(uint32_t)x * BVG_MIX;
        row += copied;
        uint32_t v = (uint32_t)(int32_t)((int64_t)x + nat2int(Rd<DEF>::resid(b, c)));  // :954
        if (store) row[0] = (int32_t)v;
        if (FOLD) fold ^= fold_base + (unsigned long long)v;
        for (int32_t i = 1; i < rc; i++) {
            v += Rd<DEF>::gap(b, c) + 1u;  // :966
            if (store) row[i] = (int32_t)v;
            if (FOLD) fold ^= fold_base + (unsigned long long)v;
        }
        if (b.pos() > g.bit_end - g.bit_base) fail(g, E_IO);
    }

    // Records with intervals: Merged(IntIntervalSequenceIterator, ResidualIntIterator) streamed from two cursors.
    template <bool FOLD>
    __device__ __forceinline__ void with_intervals(const GraphDev& g, int32_t* __restrict__ row, bool store, unsigned long long& fold) {
        if (ic == 0) return;
        const Codec& c = g.c;
        const unsigned long long fold_base = (unsigned long long)(uint32_t)x * BVG_MIX;
        const uint64_t limit = g.bit_end - g.bit_base;
        row += copied;
        int64_t icur = 0, irem = 0, iprev = 0;
        int32_t ileft = ic, left_r = rc;
        bool ifirst = true;
        int64_t rnext = 0;
        if (left_r > 0) rnext = (int64_t)(int32_t)((int64_t)x + nat2int(Rd<DEF>::resid(b, c)));  // :954
        int32_t k = 0;
        for (;;) {
            if (irem == 0 && ileft > 0) {  // load the next interval (:1084-1095)
                if (ifirst) { icur = (int64_t)(int32_t)(nat2int(ib.gamma()) + (int64_t)x); ifirst = false; }
                else icur = (int64_t)ib.gamma() + iprev + 1;
                irem = (int64_t)ib.gamma() + c.minlen;
                iprev = icur + irem;
                ileft--;
            }
            const int64_t iv = irem > 0 ? icur : BVG_INF;
            const int64_t rv = left_r > 0 ? rnext : BVG_INF;
            if (iv == BVG_INF && rv == BVG_INF) break;
            int64_t out;
            if (iv < rv) { out = iv; icur++; irem--; }
            else {
                out = rv;
                if (iv == rv) { icur++; irem--; }
                if (--left_r > 0) rnext += (int64_t)Rd<DEF>::gap(b, c) + 1;  // :966
            }
            if (store) row[k] = (int32_t)out;
            if (FOLD) fold ^= fold_base + (unsigned long long)(uint32_t)out;
            k++;
            if (k >= nout) break;
        }
        while (k < nout) {  // only reachable for files with duplicated successors (:1210 drains -1)
            if (store) row[k] = -1;
            if (FOLD) fold ^= fold_base + 0xffffffffull;
            k++;
        }
        if (b.pos() > limit) fail(g, E_IO);
    }
};

// The three phases back to back (natural-order kernels, random access). Returns `copied` or a negative error.
template <bool DEF, bool FOLD = false>
__device__ int64_t decode_extras(const GraphDev& g, int32_t x, int32_t* __restrict__ row, bool store = true,
                                 unsigned long long* acc = nullptr) {
    ExtrasWalk<DEF> w;
    unsigned long long fold = 0;
    w.header(g, x, true);
    w.template residuals_only<FOLD>(g, row, store, fold);
    w.template with_intervals<FOLD>(g, row, store, fold);
    if (FOLD) *acc ^= fold;
    return w.err ? (int64_t)w.err : (int64_t)w.copied;
}

// ---------------------------------------------------------------------------------------------------
// Step 2: stream the parent's row through the copy blocks (MaskedIntIterator.java:65-97) and either merge it, forward
// and in place, with the extras sitting at row[copied .. d) -- the output position never overtakes the unread tail
// because exactly `copied` elements come from the parent, and once the parent stream is exhausted the rest is already
// in place -- or, when nobody copies from this node during a consume-only scan, just fold the copied successors.
// Same three-phase shape as ExtrasWalk.
// ---------------------------------------------------------------------------------------------------
template <bool DEF>
struct MergeWalk {
    BitBuf b;            // block-list cursor
    int32_t x;
    int32_t d, dp, copied;
    int32_t bc, bi;      // blocks, next block to read
    int32_t p, rem;      // next parent index, remaining length of the current copy block
    bool tail, active;
    const int32_t* __restrict__ parent;

    __device__ __forceinline__ void header(const GraphDev& g, int32_t x_, const int32_t* __restrict__ parent_, bool active_) {
        const Codec& c = g.c;
        x = x_; parent = parent_; active = active_;
        d = dp = copied = bc = bi = p = rem = 0;
        tail = false;
        if (!active) return;
        b = buffer_at(g, x);
        d = (int32_t)Rd<DEF>::outdeg(b, c);
        const int32_t r = (int32_t)Rd<DEF>::ref(b, c);
        bc = (int32_t)Rd<DEF>::bcount(b, c);
        dp = g.outdeg[x - r - g.node_lo];
        // copied count: from the index when it has one, else a first walk over the blocks (same arithmetic as step 1)
        if (g.copied) copied = g.copied[x - g.node_lo];
        else {
            BitBuf t = b;
            int64_t total = 0, cp = 0;
            for (int32_t i = 0; i < bc; i++) {
                const int64_t blk = (int64_t)Rd<DEF>::block(t, c) + (i ? 1 : 0);
                total += blk;
                if (!(i & 1)) cp += blk;
            }
            if (!(bc & 1)) cp += dp - total;
            copied = (int32_t)cp;
        }
    }

    // Entry from a schedule record (MergeRec): the cursor starts at the first block code.
    __device__ __forceinline__ void header_rec(const GraphDev& g, int32_t x_, int32_t d_, int32_t dp_, int32_t bc_, int32_t copied_,
                                               uint64_t pos, const int32_t* __restrict__ parent_, bool active_) {
        x = x_; parent = parent_; active = active_;
        d = d_; dp = dp_; copied = copied_; bc = bc_;
        bi = p = rem = 0;
        tail = false;
        if (!active) return;
        b.w = g.words; b.maxw = g.nwords - 3;
        b.seek(pos);
    }

    // next copied successor, or BVG_INF
    __device__ __forceinline__ int64_t next_a(const Codec& c) {
        for (;;) {
            if (rem > 0) { rem--; return parent[p++]; }
            if (tail) return p < dp ? (int64_t)parent[p++] : BVG_INF;
            if (bi == bc) { if (bc & 1) return BVG_INF; tail = true; continue; }  // even count: copy the tail
            const int32_t blk = (int32_t)Rd<DEF>::block(b, c) + (bi ? 1 : 0);
            if (bi & 1) p += blk; else rem = blk;
            bi++;
        }
    }

    // nobody copies from x: the copied successors are only consumed.  One flat loop whose every trip is either a block
    // boundary (parse the next length), a jump over a skip block, or one copied element -- so that the lanes of a warp,
    // which sit in different blocks, still share the loop.
    __device__ __forceinline__ void stream_only(const GraphDev& g, unsigned long long& fold) {
        if (!active) return;
        const Codec& c = g.c;
        const unsigned long long fold_base = (unsigned long long)(uint32_t)x * BVG_MIX;
        int32_t pos = 0, edge = dp, blk = 0;
        bool copying = true, in_tail = bc == 0;  // no blocks: everything is copied (MaskedIntIterator, left = -1)
        if (!in_tail) edge = (int32_t)Rd<DEF>::block(b, c);
        while (pos < dp) {
            if (pos == edge && !in_tail) {
                if (++blk < bc) { edge += (int32_t)Rd<DEF>::block(b, c) + 1; copying = !(blk & 1); }
                else { in_tail = true; copying = !(bc & 1); edge = dp; }
            } else if (!copying) {
                pos = edge;
            } else {
                fold ^= fold_base + (unsigned long long)(uint32_t)parent[pos++];
            }
        }
    }

    // Somebody copies from x: its list has to exist, sorted.  Same flat loop as stream_only, with one more kind of trip:
    // when the head of the parent stream is known, emit the smaller of it and the next unread extra.
    template <bool FOLD>
    __device__ __forceinline__ void merge_in_place(const GraphDev& g, int32_t* __restrict__ row, unsigned long long& fold) {
        if (!active) return;
        const Codec& c = g.c;
        const unsigned long long fold_base = (unsigned long long)(uint32_t)x * BVG_MIX;
        int32_t pos = 0, edge = dp, blk = 0;
        bool copying = true, in_tail = bc == 0;
        if (!in_tail) edge = (int32_t)Rd<DEF>::block(b, c);
        int32_t j = copied, k = 0;   // j: next unread extra, k: next output slot
        bool have_a = false;
        int32_t a = 0;
        for (;;) {
            if (!have_a && pos < dp) {           // bring the parent stream to its next copied element
                if (pos == edge && !in_tail) {
                    if (++blk < bc) { edge += (int32_t)Rd<DEF>::block(b, c) + 1; copying = !(blk & 1); }
                    else { in_tail = true; copying = !(bc & 1); edge = dp; }
                } else if (!copying) {
                    pos = edge;
                } else {
                    a = parent[pos++];
                    have_a = true;
                }
                continue;
            }
            if (!have_a) {                       // parent exhausted
                if (k == j) break;               // nothing was dropped: the remaining extras already sit in place
                if (j < d) { row[k++] = row[j++]; continue; }
                while (k < d) row[k++] = -1;     // duplicates were dropped (:1210)
                break;
            }
            const int32_t bv = j < d ? row[j] : 0x7fffffff;
            if (j >= d || a < bv) {
                row[k++] = a;
                if (FOLD) fold ^= fold_base + (unsigned long long)(uint32_t)a;
                have_a = false;
            } else {
                row[k++] = bv; j++;
                if (a == bv) have_a = false;     // equal heads are emitted once (MergedIntIterator.java:70)
            }
        }
    }
};

template <bool DEF, bool FOLD = false>
__device__ void merge_copied(const GraphDev& g, int32_t x, int32_t* __restrict__ row, const int32_t* __restrict__ parent,
                             bool store = true, unsigned long long* acc = nullptr) {
    MergeWalk<DEF> w;
    unsigned long long fold = 0;
    w.header(g, x, parent, true);
    if (FOLD && !store) w.stream_only(g, fold);
    else w.template merge_in_place<FOLD>(g, row, fold);
    if (FOLD) *acc ^= fold;
}

}  // namespace bvg
