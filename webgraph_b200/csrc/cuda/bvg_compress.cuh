// bvg_compress.cuh -- BVGraph.store on the device (compress half), default codings (gamma outdegrees / blocks /
// block counts / intervals, unary references, zeta_k residuals).
//
// The reference compresses node by node (CompressionThread, BVGraph.java:2221-2386): for node x every candidate reference
// ref = 0 .. W whose chain is shorter than maxRefCount is tried -- the list is split against the candidate into copy / skip
// blocks and extras (diffComp, :2049-2219), the extras into intervals and residuals (intervalize, :1631-1654) -- and the
// cheapest encoding wins, ties to the smaller ref.  Node x depends on its predecessors only through their chain lengths
// (refCount), and the reference's own multi-threaded store cuts the node range into pieces compressed with an empty window
// each and splices the bits (:2471-2550).  Here the pieces are RANGES of `range_nodes` nodes, three phases:
//   k_bvc_costs   what a candidate costs does not depend on the chain lengths, only whether it may be used does: so the cost of
//                 EVERY (node, candidate) pair is computed first, one thread per pair in node order (streaming walker: blocks,
//                 intervals and residuals are costed as the merge of the two lists produces them, nothing is materialised);
//   k_bvc_pick    one thread per range goes through its nodes in order and picks, per node, the cheapest candidate whose chain
//                 is shorter than maxRefCount (ties to the smaller ref); best ref and record length per node go out;
//   (scan)        record lengths -> bit position of every node;
//   k_bvc_write   one thread per node streams its list against the chosen reference once more and writes the three sections
//                 (blocks, intervals, residuals) through three cursors -- their lengths are known from the costing pass --
//                 ORing MSB-first fields into 32-bit big-endian words with atomics (neighbouring records share words).
// With the same ranges the output is byte-identical to the host writer's (webgraph_b200/csrc/tools/bvg_tools.cpp, itself
// byte-identical to the reference's files on cnr-2000), which is what the tests check.
#pragma once
#include <climits>
#include "bvg_device.cuh"

namespace bvg {

struct BvcCodec { int32_t window, maxref, minlen, zetak; };   // maxref < 0: unbounded

struct BvcDev {
    const int64_t* __restrict__ off;
    const int32_t* __restrict__ succ;
    int32_t n;
    BvcCodec c;
    int32_t range_nodes;
};

__device__ __forceinline__ int bvc_msb(uint64_t x) { return 63 - __clzll((long long)x); }
__device__ __forceinline__ int bvc_len_gamma(uint64_t x) { return 2 * bvc_msb(x + 1) + 1; }
__device__ __forceinline__ int bvc_len_zeta(uint64_t x, int k) {
    const uint64_t y = x + 1;
    const int h = bvc_msb(y) / k;
    const uint64_t left = 1ull << (h * k);
    return h + 1 + h * k + k - 1 + (y - left < left ? 0 : 1);
}
__device__ __forceinline__ uint64_t bvc_int2nat(int64_t v) { return v >= 0 ? (uint64_t)v << 1 : (((uint64_t)(-v)) << 1) - 1; }

// MSB-first field of `width` <= 64 bits at bit position pos, into big-endian 32-bit words (word i holds bits 32 i .. 32 i + 31,
// first bit in the most significant position).
__device__ __forceinline__ void bvc_put(uint32_t* __restrict__ w, uint64_t pos, uint64_t v, int width) {
    while (width > 0) {
        const uint64_t i = pos >> 5;
        const int room = 32 - (int)(pos & 31);
        const int take = width < room ? width : room;
        const uint32_t piece = (uint32_t)((v >> (width - take)) & (take == 32 ? 0xffffffffull : ((1ull << take) - 1ull)));
        if (piece) atomicOr(w + i, piece << (room - take));
        pos += (uint64_t)take;
        width -= take;
    }
}
__device__ __forceinline__ int bvc_put_gamma(uint32_t* w, uint64_t pos, uint64_t x) {   // msb zeros, then x + 1 in msb + 1 bits
    const int len = bvc_len_gamma(x);
    if (w) bvc_put(w, pos, x + 1, len);
    return len;
}
__device__ __forceinline__ int bvc_put_zeta(uint32_t* w, uint64_t pos, uint64_t x, int k) {
    const uint64_t y = x + 1;
    const int m = bvc_msb(y);
    const int h = k == 3 ? (m * 43) >> 7 : m / k;   // m / 3 for m < 64 without a division (the default zeta_3)
    const uint64_t left = 1ull << (h * k);
    if (w) bvc_put(w, pos, 1, h + 1);   // unary(h)
    if (y - left < left) { if (w) bvc_put(w, pos + (uint64_t)h + 1, y - left, h * k + k - 1); return h + 1 + h * k + k - 1; }
    if (w) bvc_put(w, pos + (uint64_t)h + 1, y, h * k + k);
    return h + 1 + h * k + k;
}

// What one encoding of a list costs / where its sections go.  The same walker costs (w == nullptr) and writes.
struct BvcEnc {
    int64_t block_bits, iv_bits, res_bits;   // blocks without the count; intervals without the count; residuals
    int32_t bc, ic, extras;
    bool bad;
    // writing: cursors of the three sections
    uint32_t* w;
    uint64_t pos_b, pos_i, pos_r;
    // running state
    int64_t run_left;      // current run of consecutive extras: first value, length
    int32_t run_len;
    int64_t iv_prev, res_prev;   // end of the previous interval (left + len); previous residual
    bool have_iv, have_res;
};

__device__ __forceinline__ void bvc_residual(BvcEnc& e, const BvcCodec& c, int64_t x, int64_t v) {
    uint64_t code;
    if (!e.have_res) { code = bvc_int2nat(v - x); e.have_res = true; }
    else {
        if (v <= e.res_prev) { e.bad = true; return; }   // repeated / unsorted successor (BVGraph.java:2201)
        code = (uint64_t)(v - e.res_prev - 1);
    }
    const int t = bvc_put_zeta(e.w, e.pos_r, code, c.zetak);
    e.pos_r += (uint64_t)t;
    e.res_bits += t;
    e.res_prev = v;
}
// The run of consecutive extras collected so far ends: an interval if long enough, residuals otherwise (:1631-1654).
__device__ __forceinline__ void bvc_flush_run(BvcEnc& e, const BvcCodec& c, int64_t x) {
    if (e.run_len == 0) return;
    if (c.minlen != 0 && e.run_len >= c.minlen) {
        const uint64_t lv = !e.have_iv ? bvc_int2nat(e.run_left - x) : (uint64_t)(e.run_left - e.iv_prev - 1);
        const int t = bvc_put_gamma(e.w, e.pos_i, lv);
        const int u = bvc_put_gamma(e.w, e.pos_i + (uint64_t)t, (uint64_t)(e.run_len - c.minlen));
        e.pos_i += (uint64_t)(t + u);
        e.iv_bits += t + u;
        e.iv_prev = e.run_left + e.run_len;
        e.have_iv = true;
        e.ic++;
    } else {
        for (int32_t t = 0; t < e.run_len; t++) bvc_residual(e, c, x, e.run_left + t);
    }
    e.run_len = 0;
}
__device__ __forceinline__ void bvc_extra(BvcEnc& e, const BvcCodec& c, int64_t x, int64_t v) {
    e.extras++;
    if (c.minlen == 0) { bvc_residual(e, c, x, v); return; }
    if (e.run_len > 0 && v == e.run_left + e.run_len) { e.run_len++; return; }
    bvc_flush_run(e, c, x);
    e.run_left = v; e.run_len = 1;
}
__device__ __forceinline__ void bvc_block(BvcEnc& e, int32_t run) {
    const uint64_t v = (uint64_t)(e.bc == 0 ? run : run - 1);
    const int t = bvc_put_gamma(e.w, e.pos_b, v);
    e.pos_b += (uint64_t)t;
    e.block_bits += t;
    e.bc++;
}

// diffComp's split of cur[0..d) against ref_list[0..ref_len) (:2066-2109), streamed into e.
__device__ inline void bvc_walk(BvcEnc& e, const BvcCodec& c, int64_t x, const int32_t* __restrict__ cur, int32_t d,
                                const int32_t* __restrict__ ref_list, int32_t ref_len) {
    int32_t j = 0, k = 0, run = 0;
    bool copying = true;
    while (j < d && k < ref_len) {
        const int32_t a = cur[j], b = ref_list[k];
        if (copying) {
            if (a > b) { bvc_block(e, run); copying = false; run = 0; }
            else if (a < b) { bvc_extra(e, c, x, a); j++; }
            else { j++; k++; run++; }
        } else {
            if (a < b) { bvc_extra(e, c, x, a); j++; }
            else if (a > b) { k++; run++; }
            else { bvc_block(e, run); copying = true; run = 0; }
        }
    }
    if (copying && k < ref_len) bvc_block(e, run);
    while (j < d) { bvc_extra(e, c, x, cur[j]); j++; }
    bvc_flush_run(e, c, x);
}

__device__ __forceinline__ void bvc_begin(BvcEnc& e, uint32_t* w, uint64_t pos_b, uint64_t pos_i, uint64_t pos_r) {
    e.block_bits = e.iv_bits = e.res_bits = 0;
    e.bc = e.ic = e.extras = 0;
    e.bad = false;
    e.w = w; e.pos_b = pos_b; e.pos_i = pos_i; e.pos_r = pos_r;
    e.run_left = 0; e.run_len = 0; e.iv_prev = 0; e.res_prev = 0;
    e.have_iv = e.have_res = false;
}

// Bits of everything after the outdegree for reference `ref` (encode(), host writer; BVGraph.java:2115-2205).
__device__ __forceinline__ int64_t bvc_total(const BvcEnc& e, const BvcCodec& c, int32_t ref) {
    int64_t bits = 0;
    if (c.window > 0) bits += ref + 1;                                        // unary reference
    if (ref != 0) bits += bvc_len_gamma((uint64_t)e.bc) + e.block_bits;
    if (e.extras > 0) {
        if (c.minlen != 0) bits += bvc_len_gamma((uint64_t)e.ic) + e.iv_bits;
        bits += e.res_bits;
    }
    return bits;
}

// Lists must be strictly increasing and non-negative (the reference throws otherwise, BVGraph.java:2201).
__device__ inline bool bvc_list_ok(const BvcDev& g, int64_t x) {
    const int64_t a = g.off[x], b = g.off[x + 1];
    if (b > a && g.succ[a] < 0) return false;
    for (int64_t i = a + 1; i < b; i++) if (g.succ[i] <= g.succ[i - 1]) return false;
    return true;
}

// Cost of node x with reference ref (ref == 0: none).  The caller has checked that x - ref is a legal candidate.
__device__ inline int64_t bvc_cost(const BvcDev& g, int64_t x, int32_t ref, bool& bad) {
    const int64_t a = g.off[x];
    const int32_t d = (int32_t)(g.off[x + 1] - a);
    BvcEnc e;
    bvc_begin(e, nullptr, 0, 0, 0);
    const int64_t ra = ref ? g.off[x - ref] : 0;
    const int32_t rl = ref ? (int32_t)(g.off[x - ref + 1] - ra) : 0;
    bvc_walk(e, g.c, x, g.succ + a, d, g.succ + ra, rl);
    bad = e.bad || (ref == 0 && !bvc_list_ok(g, x));
    return bvc_total(e, g.c, ref);
}

// The same cost with O(1) work per list element and two paths in the loop (an extra / an advance, a block boundary folded into
// the advance): what k_bvc_costs runs.  An extra is costed as a residual as it comes; when the run of consecutive extras it
// belongs to reaches minlen the run turns into an interval and the residual state of the run's start is restored (the residual
// codes of a run's second and later elements are gap 0, but restoring is simpler than subtracting).  Equal to bvc_cost for
// every pair (tests/hostemu/emu_bvc.cpp checks all of them); list validity is bvc_list_ok's business.
struct BvcSections { int32_t bc, ic, extras; int64_t block_bits, iv_bits; };   // what the writer needs to place its three cursors
__device__ inline int64_t bvc_cost_fast(const BvcDev& g, int64_t x, int32_t ref, BvcSections* sec = nullptr) {
    const BvcCodec& c = g.c;
    const int64_t ca = g.off[x];
    const int32_t d = (int32_t)(g.off[x + 1] - ca);
    const int32_t* __restrict__ cur = g.succ + ca;
    const int64_t ra = ref ? g.off[x - ref] : 0;
    const int32_t rl = ref ? (int32_t)(g.off[x - ref + 1] - ra) : 0;
    const int32_t* __restrict__ rlist = g.succ + ra;
    int64_t block_bits = 0;   // 64-bit sums: a list of 10^8 successors must not wrap into a plausible length
    int32_t bc = 0, run = 0;
    bool copying = true;
    // residuals
    int64_t res_bits = 0;
    bool have_res = false;
    int64_t res_prev = 0;
    // the current run of consecutive extras and the residual state at its start
    int64_t run_left = 0;
    int32_t run_len = 0;
    int64_t s_res_bits = 0;
    bool s_have_res = false;
    int64_t s_res_prev = 0;
    // intervals
    int64_t iv_bits = 0;
    int32_t ic = 0;
    bool have_iv = false;
    int64_t iv_prev = 0;
    int32_t extras = 0;
    auto residual = [&](int64_t v) {
        const uint64_t code = have_res ? (uint64_t)(v - res_prev - 1) : bvc_int2nat(v - x);
        res_bits += bvc_put_zeta(nullptr, 0, code, c.zetak);
        have_res = true;
        res_prev = v;
    };
    auto close_run = [&]() {
        if (run_len >= c.minlen && run_len > 0) {
            const uint64_t lv = have_iv ? (uint64_t)(run_left - iv_prev - 1) : bvc_int2nat(run_left - x);
            iv_bits += bvc_len_gamma(lv) + bvc_len_gamma((uint64_t)(run_len - c.minlen));
            iv_prev = run_left + run_len;
            have_iv = true;
            ic++;
        }
        run_len = 0;
    };
    auto extra = [&](int64_t v) {
        extras++;
        if (c.minlen == 0) { residual(v); return; }
        if (run_len > 0 && v == run_left + run_len) {
            run_len++;
            if (run_len < c.minlen) residual(v);
            else if (run_len == c.minlen) { res_bits = s_res_bits; have_res = s_have_res; res_prev = s_res_prev; }
        } else {
            close_run();
            s_res_bits = res_bits; s_have_res = have_res; s_res_prev = res_prev;
            run_left = v; run_len = 1;
            if (run_len < c.minlen) residual(v);
        }
    };
    int32_t j = 0, k = 0;
    while (j < d && k < rl) {
        const int32_t a = cur[j], b = rlist[k];
        if (a < b) { extra(a); j++; }
        else {
            const bool eq = a == b;
            if (copying != eq) {   // a block ends here: copying and a > b, or skipping and a == b
                block_bits += bvc_len_gamma((uint64_t)(bc == 0 ? run : run - 1));
                bc++;
                copying = !copying;
                run = 0;
            }
            k++; run++; j += eq ? 1 : 0;
        }
    }
    if (copying && k < rl) { block_bits += bvc_len_gamma((uint64_t)(bc == 0 ? run : run - 1)); bc++; }
    while (j < d) { extra(cur[j]); j++; }
    if (c.minlen != 0) close_run();
    if (sec) { sec->bc = bc; sec->ic = ic; sec->extras = extras; sec->block_bits = block_bits; sec->iv_bits = iv_bits; }
    int64_t bits = 0;
    if (c.window > 0) bits += ref + 1;
    if (ref != 0) bits += bvc_len_gamma((uint64_t)bc) + block_bits;
    if (extras > 0) {
        if (c.minlen != 0) bits += bvc_len_gamma((uint64_t)ic) + iv_bits;
        bits += res_bits;
    }
    return bits;
}

// Writes the record of node x (outdegree, reference, blocks, intervals, residuals) at bit position start.
__device__ inline void bvc_write_one(const BvcDev& g, int64_t x, int32_t ref, uint64_t start, uint32_t* __restrict__ w) {
    const int64_t a = g.off[x];
    const int32_t d = (int32_t)(g.off[x + 1] - a);
    uint64_t pos = start;
    pos += (uint64_t)bvc_put_gamma(w, pos, (uint64_t)d);
    if (d == 0) return;
    const int64_t ra = ref ? g.off[x - ref] : 0;
    const int32_t rl = ref ? (int32_t)(g.off[x - ref + 1] - ra) : 0;
    BvcSections e;
    bvc_cost_fast(g, x, ref, &e);   // the section lengths
    if (g.c.window > 0) { bvc_put(w, pos, 1, ref + 1); pos += (uint64_t)ref + 1; }
    uint64_t pos_b = pos;
    if (ref != 0) { pos_b += (uint64_t)bvc_put_gamma(w, pos, (uint64_t)e.bc); pos = pos_b + (uint64_t)e.block_bits; }
    uint64_t pos_i = pos, pos_r = pos;
    if (e.extras > 0 && g.c.minlen != 0) { pos_i += (uint64_t)bvc_put_gamma(w, pos, (uint64_t)e.ic); pos_r = pos_i + (uint64_t)e.iv_bits; }
    BvcEnc wr;
    bvc_begin(wr, w, pos_b, pos_i, pos_r);
    bvc_walk(wr, g.c, x, g.succ + a, d, g.succ + ra, rl);
}

// Phase 1 for one node, all candidates in turn (the host emulation's shape; the kernel spreads the candidates over lanes).
// refc: ring of W + 1 chain lengths indexed by node % (W + 1); returns the record length in bits or -1.
__device__ inline int64_t bvc_choose_one(const BvcDev& g, int64_t x, int64_t range_lo, int32_t* __restrict__ refc, int32_t* best_ref) {
    const int32_t size = g.c.window + 1;
    const int64_t d = g.off[x + 1] - g.off[x];
    if (d < 0 || d > 0x7ffffffe) return -1;
    int64_t bits = bvc_len_gamma((uint64_t)d);
    *best_ref = 0;
    if (d == 0) return bits;
    const int64_t maxref = g.c.maxref < 0 ? INT64_MAX : g.c.maxref;
    int64_t best = INT64_MAX;
    int32_t bref = 0;
    for (int32_t ref = 0; ref < size; ref++) {
        if (ref) {
            const int64_t y = x - ref;
            if (y < range_lo || g.off[y + 1] == g.off[y] || refc[y % size] >= maxref) continue;
        }
        bool bad = false;
        const int64_t cost = bvc_cost(g, x, ref, bad);
        if (bad) return -1;
        if (cost < best) { best = cost; bref = ref; }
    }
    refc[x % size] = bref ? refc[(x - bref) % size] + 1 : 0;
    *best_ref = bref;
    return bits + best;
}

// Phase 1b for one node given the cost table (cost[x * size + ref], LLONG_MAX where the candidate does not exist): the choice
// under the chain-length constraint.  Returns the record length in bits.
__device__ inline int64_t bvc_pick(const BvcDev& g, int64_t x, const long long* __restrict__ cost, int32_t* __restrict__ refc, int32_t* best_ref) {
    const int32_t size = g.c.window + 1;
    const int64_t d = g.off[x + 1] - g.off[x];
    *best_ref = 0;
    if (d == 0) return bvc_len_gamma(0);
    const int64_t maxref = g.c.maxref < 0 ? INT64_MAX : g.c.maxref;
    long long best = LLONG_MAX;
    int32_t bref = 0;
    for (int32_t ref = 0; ref < size; ref++) {
        const long long c = cost[x * size + ref];
        if (c == LLONG_MAX) continue;
        if (ref && refc[(x - ref) % size] >= maxref) continue;
        if (c < best) { best = c; bref = ref; }
    }
    refc[x % size] = bref ? refc[(x - bref) % size] + 1 : 0;
    *best_ref = bref;
    return (int64_t)bvc_len_gamma((uint64_t)d) + best;
}

#ifndef BVG_HOST_EMULATION
constexpr int BVC_MAX_WINDOW = 31;
// Phase 1a: the cost of every (node, candidate) pair, in node order (the lanes of a node share its list in L1; the order of a
// counting sort by outdegree class measured 574 ms against 400): cost[x * size + ref].  A candidate exists when x - ref lies in
// x's range and has a non-empty list; whether its chain is short enough is decided in phase 1b.
__global__ void __launch_bounds__(128) k_bvc_costs(BvcDev g, long long* __restrict__ cost, int* __restrict__ bad) {
    const int32_t size = g.c.window + 1;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)g.n * size) return;
    const int64_t x = t / size;
    const int32_t ref = (int32_t)(t % size);
    const int64_t d = g.off[x + 1] - g.off[x];
    long long c = LLONG_MAX;
    if (d < 0 || d > 0x7ffffffe) *bad = 1;
    else if (d > 0) {
        const int64_t lo = (x / g.range_nodes) * g.range_nodes, y = x - ref;
        if (ref == 0 || (y >= lo && g.off[y + 1] > g.off[y])) {
            c = (long long)bvc_cost_fast(g, x, ref);
            if (ref == 0 && !bvc_list_ok(g, x)) *bad = 1;
        }
    }
    cost[x * size + ref] = c;
}

// Phase 1b: one thread per range walks its nodes in order (the chain lengths are the only thing a node needs of its predecessors).
__global__ void k_bvc_pick(BvcDev g, int64_t nranges, const long long* __restrict__ cost, int8_t* __restrict__ best_ref, int32_t* __restrict__ bits,
                           int* __restrict__ bad) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nranges) return;
    int32_t refc[BVC_MAX_WINDOW + 1];
    for (int i = 0; i <= BVC_MAX_WINDOW; i++) refc[i] = 0;
    const int64_t lo = r * g.range_nodes, hi = lo + g.range_nodes < g.n ? lo + g.range_nodes : g.n;
    for (int64_t x = lo; x < hi; x++) {
        int32_t ref = 0;
        const int64_t tot = bvc_pick(g, x, cost, refc, &ref);
        best_ref[x] = (int8_t)ref;
        if (tot > 0x7fffffffll || tot < 0) { *bad = 1; bits[x] = 0; } else bits[x] = (int32_t)tot;
    }
}

__global__ void k_bvc_write(BvcDev g, const int8_t* __restrict__ best_ref, const int64_t* __restrict__ node_bits, uint32_t* __restrict__ w) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x < g.n) bvc_write_one(g, x, best_ref[x], (uint64_t)node_bits[x], w);
}
#endif

}  // namespace bvg
