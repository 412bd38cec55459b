// bvg_boundaries.cuh -- record boundaries found from the .graph stream alone (second half).
//
// The reference can iterate a graph that has no .offsets file (loadSequential / loadOffline: offsetType <= 0 never
// opens it, BVGraph.java:1516-1609; BVGraphNodeIterator just keeps reading, :1201-1213) and builds the file with one
// sequential pass (BVGraph.writeOffsets, :2662-2676, CLI -O).  Record x + 1 starts where record x ends, and where a
// record ends depends on its whole content and -- through the copied count of a reference with an even number of
// blocks, :1069 -- on the outdegree of a record up to `window` back.  A sequential walk on one GPU thread would take
// about a minute for 10^9 arcs, so the stream is cut into sub-ranges that are entered speculatively and then proven,
// the scheme of bvg_offsets.cuh with a richer state:
//   state of a chain   = (bit position of the next record, outdegrees of the `window` records before it)
//   k_bnd_walk pass 0  every thread assumes a record starts at the first bit of its sub-range after `window` empty
//                      records (true for sub-range 0 only), parses records until one starts at or beyond the end of the
//                      sub-range and keeps that state as its exit.  A parse that makes no sense (reference beyond the
//                      window, blocks longer than the parent's list, more copied or intervalised successors than the
//                      outdegree) is retried one bit later.  A wrong chain falls onto the right one after a few
//                      hundred records at most -- whenever it happens to land on a true boundary and the next records
//                      do not need an outdegree it has not seen -- and from there on parses what the right one parses.
//   later passes       every thread takes the exit of the sub-range before it as its entry; if that is the entry it
//                      already used nothing changes, else it walks again.  k_bnd_check marks the sub-ranges whose
//                      entry equals the previous exit; sub-range 0 is exact, so the leading run of marked sub-ranges
//                      is proven, and the host repeats the pass until all are (each pass proves at least one more:
//                      the first unproven sub-range then starts from a proven state).  In practice two or three passes.
//   k_bnd_emit         after an exclusive scan of the record counts every thread walks its sub-range once more from
//                      its proven entry and writes the bit position of every record start.
// Exact by construction (every entry state has been walked to from bit 0), never probabilistic.  Walks that are not
// yet known to start from a proven state are capped (a made-up outdegree of 2^30 would otherwise walk the stream to
// its end): a walk that runs `cap` bits past its sub-range gives up and its exit stays unknown until the sub-range is
// the first unproven one, which is walked without a cap.
#pragma once
#include "bvg_scan.cuh"

namespace bvg {

constexpr uint64_t BND_UNKNOWN = ~0ull;

struct BndSub {
    uint64_t lo, hi;      // bit range of the sub-range (hi of the last one = end of the stream buffer)
    uint64_t entry, exit; // start of the first record of the last walk; start of the first record at or beyond hi
    int64_t count;        // records parsed by the last walk
    uint64_t bad_pos;     // start of the first parse that made no sense (BND_UNKNOWN if none), retried one bit later
    int64_t bad_index;    // records parsed before it
    int32_t walked;       // 1 = the last pass walked this sub-range again
    int32_t capped;       // 1 = the last walk was capped (its entry was not known to be proven)
};

// Long residual runs are remembered: the position after `count` residual codes starting at `pos` is a function of the
// stream alone, and the same run is walked in pass 0, again when its sub-range adopts its proven entry, and by the emit
// step -- 0.2 s each time for the 858 018 residuals of the benchmark graph's largest record, all on one lane.  A small
// open-addressing table in global memory, filled as runs are completed; a slot is published by its `ready` word.
constexpr int64_t BND_MEMO_MIN = 8192;  // residuals from which a run is remembered
constexpr int BND_MEMO_SLOTS = 1024;
struct BndMemo {
    unsigned long long pos;   // bit position of the first residual code, 0 = free slot (a record never starts with its residuals)
    long long count;
    unsigned long long end;
    int ready;
    int pad_;
};

__device__ inline bool bnd_memo_find(const BndMemo* memo, uint64_t pos, int64_t count, uint64_t& end) {
    if (!memo) return false;
    unsigned h = (unsigned)((pos * 0x9E3779B97F4A7C15ull) >> 54);  // 10 bits
    for (int probe = 0; probe < 8; probe++, h = (h + 1) & (BND_MEMO_SLOTS - 1)) {
        const volatile BndMemo* m = memo + h;
        const unsigned long long p = m->pos;
        if (p == 0) return false;
        if (p == pos && m->ready) {
            __threadfence();
            if (m->count == count) { end = m->end; return true; }
        }
    }
    return false;
}

__device__ inline void bnd_memo_put(BndMemo* memo, uint64_t pos, int64_t count, uint64_t end) {
    if (!memo) return;
    unsigned h = (unsigned)((pos * 0x9E3779B97F4A7C15ull) >> 54);
    for (int probe = 0; probe < 8; probe++, h = (h + 1) & (BND_MEMO_SLOTS - 1)) {
        BndMemo* m = memo + h;
        const unsigned long long old = atomicCAS(&m->pos, 0ull, (unsigned long long)pos);
        if (old == 0ull) {
            m->count = count; m->end = end;
            __threadfence();
            *(volatile int*)&m->ready = 1;
            return;
        }
        if (old == pos) return;  // somebody else is publishing (or has published) a run from the same position
    }
}

// One record parsed for its length.  `ring` holds the outdegrees of the `window` records before it (record r back,
// r = 1 the latest, at ring[(head + r - 1) % window]).  Returns 0 and the outdegree, 1 if the parse makes no sense, 2 if it
// ran beyond `stop`.  Follows BVGraph.successors (:1044-1100) without producing anything.
template <bool DEF>
__device__ inline int bnd_record(BitBuf& b, const Codec& c, uint64_t stop, const int32_t* ring, int32_t head, int32_t& d_out, BndMemo* memo, int lean) {
    const uint64_t d64 = Rd<DEF>::outdeg(b, c);
    if (b.pos() > stop) return 2;
    if (d64 > 0x7fffffffull) return 1;
    const int64_t d = (int64_t)d64;
    d_out = (int32_t)d;
    if (d == 0) return 0;
    int64_t cp = 0;
    if (c.window > 0) {
        const uint64_t r = Rd<DEF>::ref(b, c);
        if (b.pos() > stop) return 2;
        if (r > (uint64_t)c.window) return 1;  // :705
        if (r > 0) {
            const int64_t dp = ring[(head + (int32_t)r - 1) % c.window];
            const uint64_t bc = Rd<DEF>::bcount(b, c);
            if (b.pos() > stop) return 2;
            if (bc > (uint64_t)dp + 1) return 1;  // every block but the first copies or skips at least one successor
            int64_t total = 0;
            for (uint64_t i = 0; i < bc; i++) {  // :1062-1066
                const uint64_t raw = Rd<DEF>::block(b, c);
                if (b.pos() > stop) return 2;
                if (raw > 0x7fffffffull) return 1;
                const int64_t blk = (int64_t)raw + (i ? 1 : 0);
                total += blk;
                if (total > dp) return 1;
                if (!(i & 1)) cp += blk;
            }
            if (!(bc & 1)) cp += dp - total;  // :1069
            if (cp > d) return 1;
        }
    }
    int64_t extra = d - cp;
    if (extra > 0 && c.minlen != 0) {  // :1076-1095
        const uint64_t ic = b.gamma();
        if (b.pos() > stop) return 2;
        if (ic > (uint64_t)extra) return 1;
        int64_t tot = 0;
        for (uint64_t i = 0; i < ic; i++) {
            (void)b.gamma();
            const uint64_t len = b.gamma();
            if (b.pos() > stop) return 2;
            if (len > 0x7fffffffull) return 1;
            tot += (int64_t)len + c.minlen;
            if (tot > extra) return 1;
        }
        extra -= tot;
    }
    const uint64_t run_pos = b.pos();
    uint64_t run_end;
    if (extra >= BND_MEMO_MIN && bnd_memo_find(memo, run_pos, extra, run_end)) {
        if (run_end > stop) return 2;
        b.seek(run_end);
        return 0;
    }
    if (extra > 0) (void)Rd<DEF>::resid(b, c);  // :939-972, values not needed; the first residual is often a long code
    if (DEF && lean && extra > 1) {
        // BVG_BND_LEAN=1 (to be measured): the run walked with the 32-bit sliding window of the scan kernels (bvg_scan.cuh)
        GraphDev gw{};
        gw.words = b.w; gw.nwords = b.maxw + 3;
        Win w;
        w.seek(gw, b.pos());
        const int k = c.zetak;
        for (int64_t i = 1; i < extra; i++) {
            uint32_t m, len;
            if (zeta_fast<0>(w.top(), k, m, len)) w.skip(len);
            else (void)w.zeta_slow(gw, k);
            if ((i & 7) == 7 && (w.overrun() || w.pos(gw) > stop)) return 2;
        }
        if (w.overrun()) return 2;
        b.seek(w.pos(gw));
    } else
    for (int64_t i = 1; i < extra; i++) {
        (void)Rd<DEF>::gap(b, c);               // the 32-bit fast path of the default zeta codes, else the same reader
        if ((i & 7) == 7 && b.pos() > stop) return 2;
    }
    if (b.pos() > stop) return 2;
    if (extra >= BND_MEMO_MIN) bnd_memo_put(memo, run_pos, extra, b.pos());
    return 0;
}

// Walks the chain that enters sub-range `s` at s.entry with the outdegree history hist_in (latest first; null = zeros)
// until a record starts at or beyond s.hi.  ring: `window` words of scratch; hist_out: the history at the exit.
// cap = 0 walks without a limit (the entry is proven: running off the stream is then a truncated file).
// starts != null: the bit position of record ord_base + i is written for every ordinal up to n (k_bnd_emit).
template <bool DEF>
__device__ inline void bnd_walk(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t stream_bits, const Codec& c, BndSub& s,
                                const int32_t* hist_in, int32_t* ring, int32_t* hist_out, uint64_t cap,
                                uint64_t* __restrict__ starts, int64_t ord_base, int64_t n, BndMemo* memo, int lean) {
    const int32_t W = c.window;
    for (int32_t k = 0; k < W; k++) ring[k] = hist_in ? hist_in[k] : 0;
    int32_t head = 0;
    const uint64_t stop = cap && s.hi + cap < stream_bits ? s.hi + cap : stream_bits;
    BitBuf b;
    b.w = words; b.maxw = nwords - 3;
    uint64_t pos = s.entry;
    s.count = 0; s.bad_pos = BND_UNKNOWN; s.bad_index = 0; s.walked = 1; s.capped = cap ? 1 : 0;
    bool reseek = true;
    while (pos < s.hi && pos < stream_bits) {
        // (a retry one bit later keeps the position of the first attempt: past the last record those are the padding bits)
        if (starts && !(reseek && pos != s.entry) && ord_base + s.count <= n) starts[ord_base + s.count] = pos;
        if (reseek) { b.seek(pos); reseek = false; }
        int32_t d = 0;
        const int st = bnd_record<DEF>(b, c, stop, ring, head, d, memo, lean);
        if (st == 0) {
            pos = b.pos();
            s.count++;
            if (W > 0) { head = head ? head - 1 : W - 1; ring[head] = d; }
            continue;
        }
        if (s.bad_pos == BND_UNKNOWN) { s.bad_pos = pos; s.bad_index = s.count; }
        if (st == 2) { pos = stop == stream_bits ? stream_bits : BND_UNKNOWN; break; }  // the stream ends inside a record (padding bits after the last one, or a truncated file) / gave up at the cap
        pos++;
        reseek = true;
    }
    s.exit = pos;
    for (int32_t k = 0; k < W; k++) hist_out[k] = ring[(head + k) % W];
}

// One pass.  pass 0: speculative entries.  Later: entry = exit of the sub-range before (as the pass before left it);
// unchanged entries are not walked again.  Sub-ranges up to `trusted` start from a proven state and are walked without a cap.
template <bool DEF>
__device__ inline void bnd_pass_one(int64_t j, const uint32_t* __restrict__ words, uint64_t nwords, uint64_t stream_bits, Codec c,
                                    const BndSub* __restrict__ in, BndSub* __restrict__ out,
                                    const int32_t* __restrict__ hist_entry_in, const int32_t* __restrict__ hist_exit_in,
                                    int32_t* __restrict__ hist_entry_out, int32_t* __restrict__ hist_exit_out, int32_t* __restrict__ ring,
                                    int pass, int64_t trusted, uint64_t cap, BndMemo* memo, int lean = 0) {
    const int32_t W = c.window;
    BndSub s = in[j];
    const int32_t* he = hist_entry_in + j * W;
    bool walk = pass == 0;
    const int32_t* hin = nullptr;
    bool proven_entry = j <= trusted;
    if (pass == 0) { s.entry = s.lo; }
    else if (j > trusted && trusted >= 1 && in[trusted - 1].exit != BND_UNKNOWN && in[trusted - 1].exit >= s.lo) {
        // the proven chain leaves sub-range trusted - 1 inside a record that reaches into (or beyond) this sub-range: every
        // sub-range it covers learns its true entry in this one pass instead of one sub-range per pass
        const BndSub& q = in[trusted - 1];
        const int32_t* hq = hist_exit_in + (trusted - 1) * W;
        bool same = q.exit == s.entry && !(s.capped && s.exit == BND_UNKNOWN);
        for (int32_t k = 0; same && k < W; k++) same = hq[k] == he[k];
        if (!same) { walk = true; s.entry = q.exit; hin = hq; }
        proven_entry = true;
    }
    else if (j > 0) {
        const BndSub& p = in[j - 1];
        // An exit beyond this whole sub-range is believed only from a proven sub-range: made-up outdegrees of wrong chains
        // produce such exits, and adopting them sends a wave of wrong states (each set right one pass later) down the stream.
        if (p.exit != BND_UNKNOWN && (p.exit < s.hi || j <= trusted)) {
            const int32_t* hp = hist_exit_in + (j - 1) * W;
            bool same = p.exit == s.entry;
            for (int32_t k = 0; same && k < W; k++) same = hp[k] == he[k];
            // the same entry is walked again only to lift a cap that made (or could have made) the walk give up
            if (!same || (s.capped && j <= trusted && s.exit == BND_UNKNOWN)) { walk = true; s.entry = p.exit; hin = hp; }
        }
    }
    if (!walk) {
        s.walked = 0;
        out[j] = s;
        for (int32_t k = 0; k < W; k++) { hist_entry_out[j * W + k] = he[k]; hist_exit_out[j * W + k] = hist_exit_in[j * W + k]; }
        return;
    }
    for (int32_t k = 0; k < W; k++) hist_entry_out[j * W + k] = hin ? hin[k] : 0;
    if (s.entry >= s.hi || s.entry >= stream_bits) {  // a record of an earlier sub-range covers this one entirely
        s.exit = s.entry; s.count = 0; s.bad_pos = BND_UNKNOWN; s.bad_index = 0; s.walked = 1; s.capped = 0;
        for (int32_t k = 0; k < W; k++) hist_exit_out[j * W + k] = hin ? hin[k] : 0;
    } else {
        bnd_walk<DEF>(words, nwords, stream_bits, c, s, hin, ring + j * (W > 0 ? W : 1), hist_exit_out + j * W, proven_entry ? 0 : cap, nullptr, 0, 0, memo, lean);
    }
    out[j] = s;
}

// ok[j] = sub-range j entered where (and with the history with which) sub-range j - 1 left, and both exits are known.
__device__ inline void bnd_check_one(int64_t j, const BndSub* __restrict__ sub, const int32_t* __restrict__ hist_entry,
                                     const int32_t* __restrict__ hist_exit, int32_t W, int32_t* __restrict__ ok) {
    bool good = sub[j].exit != BND_UNKNOWN;
    if (j > 0) {
        good = good && sub[j - 1].exit == sub[j].entry;
        for (int32_t k = 0; good && k < W; k++) good = hist_exit[(j - 1) * W + k] == hist_entry[j * W + k];
    } else good = good && sub[0].entry == sub[0].lo;
    ok[j] = good ? 1 : 0;
}

template <bool DEF>
__device__ inline void bnd_emit_one(int64_t j, const uint32_t* __restrict__ words, uint64_t nwords, uint64_t stream_bits, Codec c,
                                    const BndSub* __restrict__ sub, const int32_t* __restrict__ hist_entry, int32_t* __restrict__ ring,
                                    int32_t* __restrict__ hist_scratch, const int64_t* __restrict__ base, int64_t n, uint64_t* __restrict__ starts, BndMemo* memo, int lean = 0) {
    const int32_t W = c.window;
    BndSub s = sub[j];
    if (s.entry >= s.hi || s.entry >= stream_bits) return;
    bnd_walk<DEF>(words, nwords, stream_bits, c, s, hist_entry + j * W, ring + j * (W > 0 ? W : 1), hist_scratch + j * W, 0, starts, base[j], n, memo, lean);
}

#ifndef BVG_HOST_EMULATION
__global__ void k_bnd_init(BndSub* __restrict__ sub, int64_t nsub, uint64_t sub_bits, uint64_t stream_bits) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nsub) return;
    BndSub s;
    s.lo = (uint64_t)j * sub_bits;
    s.hi = j + 1 == nsub ? stream_bits : (uint64_t)(j + 1) * sub_bits;
    s.entry = s.lo; s.exit = BND_UNKNOWN; s.count = 0; s.bad_pos = BND_UNKNOWN; s.bad_index = 0; s.walked = 0; s.capped = 0;
    sub[j] = s;
}

template <bool DEF>
__global__ void k_bnd_walk(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t stream_bits, Codec c, int64_t nsub,
                           const BndSub* __restrict__ in, BndSub* __restrict__ out,
                           const int32_t* __restrict__ hist_entry_in, const int32_t* __restrict__ hist_exit_in,
                           int32_t* __restrict__ hist_entry_out, int32_t* __restrict__ hist_exit_out, int32_t* __restrict__ ring,
                           int pass, int64_t trusted, uint64_t cap, BndMemo* memo, int lanes, int lean) {
    // `lanes` walks per warp (default 1, lane 0 only; BVG_BND_LANES).  With 32 walks per warp there are too few warps to
    // hide latency and the walks of a warp diverge (2.4 s for the 1 B-arc graph); with one, every issue slot serves a
    // single lane (1.15 s).  The walks run the same loops, so a few per warp should converge most of the time.
    const int lane = threadIdx.x & 31;
    if (lane >= lanes) return;
    const int64_t j = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * lanes + lane;
    if (j < nsub) bnd_pass_one<DEF>(j, words, nwords, stream_bits, c, in, out, hist_entry_in, hist_exit_in, hist_entry_out, hist_exit_out, ring, pass, trusted, cap, memo, lean);
}

__global__ void k_bnd_check(const BndSub* __restrict__ sub, int64_t nsub, const int32_t* __restrict__ hist_entry,
                            const int32_t* __restrict__ hist_exit, int32_t W, int32_t* __restrict__ ok) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nsub) bnd_check_one(j, sub, hist_entry, hist_exit, W, ok);
}

template <bool DEF>
__global__ void k_bnd_emit(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t stream_bits, Codec c, int64_t nsub,
                           const BndSub* __restrict__ sub, const int32_t* __restrict__ hist_entry, int32_t* __restrict__ ring,
                           int32_t* __restrict__ hist_scratch, const int64_t* __restrict__ base, int64_t n, uint64_t* __restrict__ starts, BndMemo* memo, int lanes, int lean) {
    const int lane = threadIdx.x & 31;  // `lanes` walks per warp, as in k_bnd_walk
    if (lane >= lanes) return;
    const int64_t j = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * lanes + lane;
    if (j < nsub) bnd_emit_one<DEF>(j, words, nwords, stream_bits, c, sub, hist_entry, ring, hist_scratch, base, n, starts, memo, lean);
}
#endif

}  // namespace bvg
