// bvg_tile.cuh -- the tile kernel: one thread block decodes a tile of consecutive nodes end to end out of shared memory.
//
// What BVGraphNodeIterator does with its cyclic window of W + 1 decoded lists (reference
// src/it/unimi/dsi/webgraph/BVGraph.java:1136-1213: every list is kept only as long as a later node can copy from
// it), a block does here for a tile of a few hundred consecutive nodes:
//   * the tile's stretch of the .graph stream is brought into shared memory by bulk asynchronous copies
//     (cp.async.bulk + mbarrier, one copy per run of short records; SASS: UBLKCP), nothing of it is read twice from HBM;
//   * record headers (outdegree, reference: BVGraph.java:1048-1053) are parsed in the kernel, one thread per node;
//   * reference chains are resolved level by level inside the tile, levels separated by __syncthreads(); the lists somebody
//     in the tile copies from live in shared memory for the lifetime of the tile and never touch HBM; the planner (below)
//     cuts tiles where no reference crosses, or else prepends the few nodes a chain reaches back to (the halo, re-decoded
//     exactly as BVGraphNodeIterator's constructor re-reads the window, :1173-1183);
//   * records are handed to lanes from a work list ordered by record length (counting sort in shared memory), warps take
//     32 neighbouring entries at a time from an atomic ticket;
//   * records with more than `long_d` successors are not in the staged stream: their parts are split across all threads
//     of the block at the sync points of the long index (bvg_long.cuh), their lists -- when somebody copies from them --
//     live in a global scratch that only this block touches.
// No per-node schedule records, no CSR-sized row scratch: a scan reads the stream, the offsets and the tile plan.
#pragma once
#include "bvg_device.cuh"
#include "bvg_scan.cuh"
#include "bvg_long.cuh"

namespace bvg {

constexpr int TILE_MAX_NODES = 1024;   // local node indices are 16-bit
constexpr int TILE_MAX_LONG = 24;      // long records per tile (each starts a new run of the staged stream)
constexpr int TILE_BUCKETS = 100;      // quarter octaves of record bits (records of up to 2^25 bits)
constexpr int TILE_LEVELS = 30;        // chain levels with a work list of their own; deeper levels share the last one
constexpr int TILE_COPY_RUNS = 4;      // copy runs staged per lane (CopyRunsT)
constexpr int TILE_NODE_BYTES = 32;    // shared memory per node: five 32-bit, three 16-bit and four 8-bit arrays (30) rounded up
constexpr uint32_t TF_LONG = 1, TF_DEAD = 2;

struct TileEntry {
    int32_t lo, from, hi;        // nodes [lo, hi) are decoded, [from, hi) are consumed; [lo, from) is the halo
    int32_t long_lo, long_hi;    // long records of the tile: LongIndex.meta[long_lo .. long_hi)
    uint32_t cost;               // stream bits of the tile / 64 (launch order: heaviest first)
};

struct TileSh {
    unsigned long long mbar;
    unsigned long long run_src_bit[TILE_MAX_LONG + 1];  // bit position (from word 0 of the stream buffer) of the run's first staged bit
    uint32_t run_dst_bit[TILE_MAX_LONG + 1];            // where that bit sits in the staged stream
    int32_t long_local[TILE_MAX_LONG];                  // local index of every long record, ascending
    uint32_t next_item;
    int32_t maxlevel;
    uint32_t nE;
    uint32_t stream_words, rows_used;
    int32_t err;
    uint32_t bucket[TILE_BUCKETS + 1];
    uint32_t levcnt[TILE_LEVELS + 2];
    uint32_t long_cum[TILE_MAX_LONG + 1];               // running number of parts of the long records in the current phase
};

// Shared memory a tile needs: planner and kernel use the same arithmetic.
__host__ __device__ inline uint32_t tile_fixed_bytes(int nt) {
    return (uint32_t)((sizeof(TileSh) + 15) & ~(size_t)15) + (uint32_t)nt * 2u * TILE_COPY_RUNS * 4u + 64u;
}
__host__ __device__ inline uint32_t tile_node_bytes(int32_t nn) { return (uint32_t)TILE_NODE_BYTES * (uint32_t)((nn + 4) & ~3); }
// per node of a tile: its share of the node arrays, its record's bytes (short records), its list when somebody copies from it
__host__ __device__ inline uint32_t tile_node_cost(uint64_t bits, int32_t d, bool is_long, bool stored, uint32_t budget) {
    (void)budget;
    if (is_long) return (uint32_t)TILE_NODE_BYTES + 64u;   // not staged; each run of the stream is padded
    uint64_t c = (uint64_t)TILE_NODE_BYTES + (bits + 7) / 8 + (stored ? 4ull * (uint64_t)d : 0ull);
    return c > 0x7fffffffull ? 0x7fffffffu : (uint32_t)c;
}
// what is left for node costs in `smem_bytes` of dynamic shared memory
__host__ __device__ inline uint32_t tile_budget(uint32_t smem_bytes, int nt) {
    const uint32_t fixed = tile_fixed_bytes(nt) + (uint32_t)(TILE_MAX_LONG + 1) * 48u + 64u + tile_node_bytes(0);
    return smem_bytes > fixed + 1024u ? smem_bytes - fixed : 1024u;
}

struct TileArgs {
    const TileEntry* __restrict__ tiles;
    const int32_t* __restrict__ order;   // launch order (tile indices, heaviest first); nullptr = tiles[first + blockIdx.x]
    int32_t first, count;
    int32_t fold_lo, fold_hi;            // nodes outside [fold_lo, fold_hi) are decoded only as parents
    LongIndex li;
    int32_t* long_scr;                   // 3 d entries per long record somebody copies from, at LongMeta.scan_off
    unsigned long long* result;          // FOLD_SLOTS slot pairs (arcs, XOR)
    uint32_t smem_bytes;
    unsigned long long* timeline;        // debugging (BVG_TILE_TIMELINE): per block (tile, start ns, phase ends ..., sm), 8 words each
};

// ---------------------------------------------------------------------------------------------------
// Host emulation hooks: the kernel body is written as phases over (tid, nt) so that tests/hostemu can run a tile with a
// loop per phase.
// ---------------------------------------------------------------------------------------------------
#ifdef BVG_HOST_EMULATION
#define TILE_SYNCWARP()
static inline uint32_t tile_atomic_add(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
static inline void tile_atomic_max(int32_t* p, int32_t v) { if (v > *p) *p = v; }
#else
#define TILE_SYNCWARP() __syncwarp()
__device__ __forceinline__ uint32_t tile_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
__device__ __forceinline__ void tile_atomic_max(int32_t* p, int32_t v) { atomicMax(p, v); }
#endif

__device__ __forceinline__ int tile_bucket(uint64_t bits) {  // longest first: bucket 0 holds the longest records
    int b = half_octave_bucket(bits);
    if (b > TILE_BUCKETS - 1) b = TILE_BUCKETS - 1;
    return TILE_BUCKETS - 1 - b;
}

template <int K>
struct Tile {
    GraphDev g;        // the graph in global memory
    GraphDev tg;       // the staged stream seen as a graph: words in shared memory, bit positions tile-local
    TileSh* sh;
    uint32_t *pos, *deg, *rowo, *cop, *bpos;
    uint16_t *lvl, *ordE, *ordM;
    uint8_t *ref, *flg, *bkt, *sto;
    int32_t* runs;
    uint32_t* sw;
    uint32_t sw_cap;       // words available for the stream and the rows together
    int32_t* rows;
    uint32_t rows_cap;
    int32_t lo, from, hi, nn, long_lo, nlong;
    int32_t fold_lo, fold_hi;
    LongIndex li;
    int32_t* long_scr;

    __device__ __forceinline__ void fail(int code, int32_t x, uint64_t bitpos) {
        report(g.err, code, x, bitpos);
        sh->err = code;
    }

    // Carves the dynamic shared memory; identical in every thread.
    __device__ __forceinline__ void carve(unsigned char* smem, uint32_t smem_bytes, int nt, const TileEntry& e, const GraphDev& graph,
                                          const TileArgs& a) {
        g = graph;
        lo = e.lo; from = e.from; hi = e.hi; nn = e.hi - e.lo; long_lo = e.long_lo; nlong = e.long_hi - e.long_lo;
        fold_lo = a.fold_lo; fold_hi = a.fold_hi; li = a.li; long_scr = a.long_scr;
        uint32_t off = 0;
        sh = reinterpret_cast<TileSh*>(smem); off += (uint32_t)((sizeof(TileSh) + 15) & ~(size_t)15);
        runs = reinterpret_cast<int32_t*>(smem + off); off += (uint32_t)nt * 2u * TILE_COPY_RUNS * 4u;
        const uint32_t n4 = (uint32_t)((nn + 4) & ~3);
        pos = reinterpret_cast<uint32_t*>(smem + off); off += 4u * n4;
        deg = reinterpret_cast<uint32_t*>(smem + off); off += 4u * n4;
        rowo = reinterpret_cast<uint32_t*>(smem + off); off += 4u * n4;
        cop = reinterpret_cast<uint32_t*>(smem + off); off += 4u * n4;
        bpos = reinterpret_cast<uint32_t*>(smem + off); off += 4u * n4;
        lvl = reinterpret_cast<uint16_t*>(smem + off); off += 2u * n4;
        ordE = reinterpret_cast<uint16_t*>(smem + off); off += 2u * n4;
        ordM = reinterpret_cast<uint16_t*>(smem + off); off += 2u * n4;
        ref = smem + off; off += n4;
        flg = smem + off; off += n4;
        bkt = smem + off; off += n4;
        sto = smem + off; off += n4;
        off = (off + 15u) & ~15u;
        sw = reinterpret_cast<uint32_t*>(smem + off);
        sw_cap = smem_bytes > off ? (smem_bytes - off) / 4u : 0u;
        rows = nullptr; rows_cap = 0;
        tg = graph;
        tg.words = sw; tg.nwords = 0; tg.bit_base = 0; tg.bit_end = 0;
        tg.offsets = nullptr; tg.outdeg = nullptr; tg.ref = nullptr; tg.depth = nullptr; tg.rowoff = nullptr; tg.copied = nullptr;
    }

    // after the stream has been laid out (sh->stream_words known to everybody)
    __device__ __forceinline__ void bind_stream() {
        const uint32_t w = sh->stream_words;
        tg.nwords = w;                 // includes 4 zero words of padding
        tg.bit_end = (uint64_t)(w - 4u) * 32u;
        rows = reinterpret_cast<int32_t*>(sw + w);
        rows_cap = sw_cap - w;
    }

    __device__ __forceinline__ int long_slot(int32_t i) const {  // which long record of the tile local node i is, or -1
        for (int k = 0; k < nlong; k++) if (sh->long_local[k] == i) return k;
        return -1;
    }
    __device__ __forceinline__ int32_t* long_base(const LongMeta& m) const { return long_scr + m.scan_off; }
    // the finished list of local node p (somebody copies from it): shared memory, or the scratch of a long record
    __device__ __forceinline__ const int32_t* list_of(int32_t p) const {
        if (flg[p] & TF_LONG) {
            const LongMeta& m = li.meta[long_lo + long_slot(p)];
            return long_base(m) + 2 * (int64_t)m.d;
        }
        return rows + rowo[p];
    }

    // ---- phase 0a: one thread lays the stream out: one run of short records between every two long records ----
    // Returns the number of bulk copies it describes; copy r is (dst word offset, source byte offset from word 0, bytes).
    struct RunCopy { uint32_t dst_byte; uint64_t src_byte; uint32_t bytes; };
    __device__ __forceinline__ int layout(RunCopy* rc /* TILE_MAX_LONG + 1 */) {
        sh->next_item = 0; sh->maxlevel = 0; sh->nE = 0; sh->rows_used = 0; sh->err = 0;
        for (int b = 0; b <= TILE_BUCKETS; b++) sh->bucket[b] = 0;
        for (int l = 0; l < TILE_LEVELS + 2; l++) sh->levcnt[l] = 0;
        if (nlong > TILE_MAX_LONG || nn > TILE_MAX_NODES || nn < 0) { fail(E_NOMEM, lo, 0); sh->stream_words = 4; return 0; }
        for (int k = 0; k < nlong; k++) sh->long_local[k] = li.meta[long_lo + k].x - lo;
        uint32_t dst = 0;
        int32_t a = 0;
        int n = 0;
        for (int r = 0; r <= nlong; r++) {
            const int32_t b = r < nlong ? sh->long_local[r] : nn;
            rc[r].dst_byte = dst; rc[r].src_byte = 0; rc[r].bytes = 0;
            sh->run_src_bit[r] = 0; sh->run_dst_bit[r] = dst * 8u;
            if (b > a) {
                const uint64_t o0 = g.offsets[lo + a - g.node_lo] - g.bit_base, o1 = g.offsets[lo + b - g.node_lo] - g.bit_base;
                const uint64_t s0 = (o0 >> 3) & ~(uint64_t)15, s1 = (((o1 + 7) >> 3) + 15) & ~(uint64_t)15;
                const uint64_t bytes = s1 - s0;
                if (o1 < o0 || bytes + dst + 16 > (uint64_t)sw_cap * 4u || s1 > g.nwords * 4u) { fail(E_NOMEM, lo + a, o0 + g.bit_base); break; }
                rc[r].src_byte = s0; rc[r].bytes = (uint32_t)bytes;
                sh->run_src_bit[r] = s0 * 8u;
                dst += (uint32_t)bytes;
                n = r + 1;
            }
            a = b + 1;
        }
        sh->stream_words = dst / 4u + 4u;   // four zero words after the last run: the window looks two words ahead
        return sh->err ? 0 : n;
    }

    // ---- phase 0b: record positions in the staged stream, length buckets ----
    __device__ __forceinline__ void positions(int tid, int nt) {
        if (tid < 4) sw[sh->stream_words - 4u + (uint32_t)tid] = 0u;
        for (int32_t i = tid; i < nn; i += nt) {
            int r = 0;
            bool is_long = false;
            for (int k = 0; k < nlong; k++) { const int32_t L = sh->long_local[k]; r += L < i ? 1 : 0; is_long = is_long || L == i; }
            const uint64_t o = g.offsets[lo + i - g.node_lo] - g.bit_base, on = g.offsets[lo + i + 1 - g.node_lo] - g.bit_base;
            pos[i] = is_long ? 0u : (uint32_t)(o - sh->run_src_bit[r]) + sh->run_dst_bit[r];
            bkt[i] = (uint8_t)tile_bucket(on - o);
            flg[i] = is_long ? TF_LONG : 0;
            sto[i] = 0; ref[i] = 0; deg[i] = 0; rowo[i] = 0; cop[i] = 0; bpos[i] = 0; lvl[i] = 0;
        }
    }

    // ---- phase 1: headers (outdegree, reference), one thread per node; marks the lists somebody copies from ----
    __device__ __forceinline__ void headers(int tid, int nt) {
        for (int32_t i = tid; i < nn; i += nt) {
            uint64_t d = 0, r = 0;
            if (flg[i] & TF_LONG) {
                const LongMeta& m = li.meta[long_lo + long_slot(i)];
                d = (uint64_t)m.d; r = (uint64_t)m.ref;
            } else {
                Win b;
                b.seek(tg, pos[i]);
                d = b.gamma(tg);
                if (d != 0 && g.c.window > 0) r = b.unary(tg);
                bpos[i] = (uint32_t)b.pos(tg);
                if (d > 0x7fffffffull || b.overrun()) { fail(E_IO, lo + i, pos[i]); d = 0; r = 0; }
            }
            if (r > (uint64_t)g.c.window) { fail(E_STATE, lo + i, pos[i]); r = 0; d = 0; }        // BVGraph.java:705
            else if (r > (uint64_t)i) {
                // the chain leaves the tile: a halo node nobody in the tile needs (every ancestor of a consumed node is inside by
                // construction of the halo); anywhere else the plan and the stream disagree
                if (lo + i >= from) fail(E_FORMAT, lo + i, pos[i]);
                flg[i] |= TF_DEAD;
                r = 0; d = 0;
            }
            deg[i] = (uint32_t)d;
            ref[i] = (uint8_t)r;
            if (r) sto[i - (int32_t)r] = 1;
        }
    }

    // ---- phase 2: chain level of every node, sizes of the work lists ----
    __device__ __forceinline__ void levels(int tid, int nt) {
        for (int32_t i = tid; i < nn; i += nt) {
            int32_t y = i, lev = 0;
            while (ref[y]) { y -= ref[y]; lev++; }
            if (lev && (flg[y] & TF_DEAD)) {  // copies, through its chain, from a halo node whose chain leaves the tile: as dead as that one
                if (lo + i >= from) fail(E_FORMAT, lo + i, pos[i]);
                flg[i] |= TF_DEAD;
                deg[i] = 0;
                lev = 0;
            }
            lvl[i] = (uint16_t)lev;
            if (lev) tile_atomic_max(&sh->maxlevel, lev);
            if (deg[i] == 0 || (flg[i] & TF_LONG)) continue;
            tile_atomic_add(&sh->bucket[bkt[i]], 1u);
            if (lev) tile_atomic_add(&sh->levcnt[lev < TILE_LEVELS ? lev : TILE_LEVELS], 1u);
        }
    }

    // ---- phase 3: three exclusive scans (bucket sizes, level sizes, list lengths), one warp each ----
    // value(i) for i in [0, n) -> store(i, exclusive prefix); returns the total in every lane.
    template <class V, class S>
    __device__ __forceinline__ uint32_t warp_scan(int lane, int32_t n, V value, S store) {
#ifdef BVG_HOST_EMULATION
        (void)lane;
        uint32_t run = 0;
        for (int32_t i = 0; i < n; i++) { const uint32_t v = value(i); store(i, run); run += v; }
        return run;
#else
        const int32_t per = (n + 31) / 32, a = lane * per, b = min(n, a + per);
        uint32_t s = 0;
        for (int32_t i = a; i < b; i++) s += value(i);
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        uint32_t run = inc - s;
        for (int32_t i = a; i < b; i++) { const uint32_t v = value(i); store(i, run); run += v; }
        return __shfl_sync(0xffffffffu, inc, 31);
#endif
    }
    __device__ __forceinline__ void scan_buckets(int lane) {
        uint32_t* bk = sh->bucket;
        const uint32_t tot = warp_scan(lane, TILE_BUCKETS, [bk](int32_t i) { return bk[i]; }, [bk](int32_t i, uint32_t v) { bk[i] = v; });
        if (lane == 0) sh->nE = tot;
    }
    __device__ __forceinline__ void scan_levels(int lane) {
        uint32_t* lc = sh->levcnt;
        (void)warp_scan(lane, TILE_LEVELS + 1, [lc](int32_t i) { return lc[i]; }, [lc](int32_t i, uint32_t v) { lc[i] = v; });
    }
    __device__ __forceinline__ void scan_rows(int lane) {
        const uint8_t* st = sto; const uint8_t* fl = flg; const uint32_t* dg = deg; uint32_t* ro = rowo;
        const uint32_t tot = warp_scan(lane, nn, [st, fl, dg](int32_t i) { return (st[i] && !(fl[i] & TF_LONG)) ? dg[i] : 0u; },
                                       [ro](int32_t i, uint32_t v) { ro[i] = v; });
        if (lane == 0) {
            sh->rows_used = tot;
            if (tot > rows_cap) fail(E_NOMEM, lo, 0);
        }
    }

    // ---- phase 4: the work lists (counting sort; the buckets / level counters now hold running ends) ----
    __device__ __forceinline__ void scatter(int tid, int nt) {
        for (int32_t i = tid; i < nn; i += nt) {
            if (deg[i] == 0 || (flg[i] & TF_LONG)) continue;
            ordE[tile_atomic_add(&sh->bucket[bkt[i]], 1u)] = (uint16_t)i;
            const int32_t lev = lvl[i];
            if (lev) ordM[tile_atomic_add(&sh->levcnt[lev < TILE_LEVELS ? lev : TILE_LEVELS], 1u)] = (uint16_t)i;
        }
    }
    // entries of ordM holding level l (1 <= l < TILE_LEVELS), or every deeper level (l >= TILE_LEVELS), after scatter()
    __device__ __forceinline__ void level_range(int32_t l, uint32_t& a, uint32_t& b) const {
        const int32_t s = l < TILE_LEVELS ? l : TILE_LEVELS;
        a = sh->levcnt[s - 1]; b = sh->levcnt[s];
    }

    __device__ __forceinline__ bool consumed(int32_t x) const { return x >= from && x >= fold_lo && x < fold_hi; }

    // ---- phase 5: extras of one short record (BVGraph.java:1062-1100): copy blocks are walked for the copied count,
    // intervals and residuals are folded as they are decoded and, when somebody copies from the node, stored in the tail
    // of its list ----
    __device__ __forceinline__ void extras_item(int32_t i, unsigned long long& acc, long long& arcs) {
        const bool valid = i >= 0;
        const int32_t x = lo + (valid ? i : 0);
        const uint32_t d = valid ? deg[i] : 0u;
        const uint32_t r = valid ? ref[i] : 0u;
        const bool store = valid && sto[i] != 0;
        const bool fold = valid && consumed(x);
        bool active = valid && d > 0 && (fold || store) && sh->err == 0;
        Win b;
        b.p0 = tg.words; b.idx = 0; b.lim = 0; b.w0 = b.w1 = b.q0 = b.q1 = b.q2 = 0; b.s = 0;
        uint32_t copied = 0;
        if (active) {
            b.seek(tg, bpos[i]);
            if (r) {
                const uint64_t bc = b.gamma(tg);
                int64_t total = 0, cp = 0;
                bool ok = bc <= 0x7fffffffull;
                for (uint64_t k = 0; ok && k < bc; k++) {  // :1062-1066
                    const int64_t blk = (int64_t)b.gamma(tg) + (k ? 1 : 0);
                    total += blk;
                    if (!(k & 1)) cp += blk;
                    if (b.overrun()) ok = false;
                }
                const int64_t dp = deg[i - (int32_t)r];
                if (ok && !(bc & 1)) cp += dp - total;  // :1069
                if (!ok || total > dp || cp < 0 || cp > (int64_t)d) { fail(ok ? E_FORMAT : E_IO, x, b.pos(tg)); active = false; }
                else copied = (uint32_t)cp;
            }
            cop[i] = copied;
        }
        TILE_SYNCWARP();
        ScanExtras<K, Win> w;
        w.begin(tg, x, 0, 0, false);
        if (active) { w.nout = (int32_t)(d - copied); w.rc = w.nout; w.b = b; }
        w.iv_fold(tg);
        TILE_SYNCWARP();
        int32_t* row = store && active ? rows + rowo[i] + copied : nullptr;
        const bool st = store && active;
#ifdef BVG_HOST_EMULATION
        if (st) w.template resid<true>(tg, row, true); else w.template resid<false>(tg, row, false);
#else
        if (__any_sync(0xffffffffu, st)) w.template resid<true>(tg, row, st);
        else w.template resid<false>(tg, row, false);
#endif
        TILE_SYNCWARP();
        if (st && w.ic) w.iv_merge(tg, row);
        TILE_SYNCWARP();
        if (w.err) sh->err = w.err;
        if (active && fold && !w.err) { acc ^= w.finish(); arcs += (long long)d; }
    }

    // ---- phase 6: copied part of one short record of the current level (MaskedIntIterator.java:65-97) ----
    __device__ __forceinline__ void merge_item(int32_t i, int slot_tid, int nt, unsigned long long& acc) {
        const bool valid = i >= 0;
        const int32_t x = lo + (valid ? i : 0);
        const uint32_t d = valid ? deg[i] : 0u;
        const uint32_t r = valid ? ref[i] : 0u;
        const bool store = valid && sto[i] != 0;
        const bool fold = valid && consumed(x);
        const bool active = valid && r != 0 && (fold || store) && sh->err == 0;
        uint32_t dp = 0, bc = 0;
        uint64_t recpos = 0;
        const int32_t* parent = nullptr;
        if (active) {
            const int32_t p = i - (int32_t)r;
            dp = deg[p];
            parent = list_of(p);
            Win b;
            b.seek(tg, bpos[i]);
            bc = (uint32_t)b.gamma(tg);
            recpos = b.pos(tg);
        }
        CopyRunsT<TILE_COPY_RUNS> c;
        c.begin(tg, recpos, (int32_t)bc, (int32_t)dp, runs + slot_tid, nt, active);
        c.stage(tg);
        TILE_SYNCWARP();
        unsigned long long f = 0;
        if (active && !store) f = copied_fold<4>(tg, c, x, parent);
        TILE_SYNCWARP();
        if (active && store) f = copied_merge(tg, c, x, (int32_t)d, (int32_t)cop[i], rows + rowo[i], parent);
        TILE_SYNCWARP();
        if (active && fold) acc ^= f;
    }

    // ---- long records: parts split across the threads of the block (same arithmetic as k_long_* in bvg_long.cuh) ----
    struct LongRec {
        const LongMeta* m;
        int32_t local;
        bool stored, fold, active;
    };
    __device__ __forceinline__ LongRec long_rec(int k) const {
        LongRec L;
        L.m = li.meta + long_lo + k;
        L.local = sh->long_local[k];
        L.stored = sto[L.local] != 0;
        L.fold = consumed(lo + L.local);
        L.active = (L.stored || L.fold) && sh->err == 0 && deg[L.local] > 0;
        return L;
    }
    __device__ __forceinline__ int32_t* long_resid_dst(const LongMeta& m) const {
        int32_t* base = long_base(m);
        if (m.ic == 0) return m.copied == 0 ? base + 2 * (int64_t)m.d : base + m.d;
        return base;
    }
    __device__ __forceinline__ int32_t* long_extras_dst(const LongMeta& m) const {
        int32_t* base = long_base(m);
        return m.copied == 0 ? base + 2 * (int64_t)m.d : base + m.d;
    }

    // residual segment `part` (BVGraph.java:939-972)
    __device__ __forceinline__ void long_resid_part(const LongRec& L, int32_t part, unsigned long long& acc, long long& arcs) {
        const LongMeta& m = *L.m;
        const bool consume = !L.stored;
        if (consume && !L.fold) return;
        const int32_t first = part * li.seg;
        const int32_t cnt = min(li.seg, m.rc - first);
        Fold32 f;
        f.begin(m.x);
        const int k = g.c.zetak;
        Win b;
        b.seek(g, li.seg_pos[m.seg_off + part]);
        uint32_t v = (uint32_t)li.seg_val[m.seg_off + part];
        int32_t* out = consume ? nullptr : long_resid_dst(m) + first;
#pragma unroll 1
        for (int32_t t = 0; t < cnt; t++) {
            if (first + t == 0) v = (uint32_t)(int32_t)((int64_t)m.x + nat2int(zeta_any<K>(b, g, k) - 1ull));
            else {
                uint32_t mm, len;
                if (zeta_fast<K>(b.top(), k, mm, len)) b.skip(len);
                else mm = (uint32_t)zeta_any<K>(b, g, k);
                v += mm;
            }
            f.add(v);
            if (!consume) out[t] = (int32_t)v;
        }
        if (b.overrun()) fail(E_IO, m.x, b.pos(g) + g.bit_base);
        // stored records with neither intervals nor a copied part get their final list here
        if (L.fold && (consume || (m.ic == 0 && m.copied == 0))) {
            f.n = (uint32_t)cnt;
            acc ^= f.finish(m.x);
            arcs += cnt;
        }
    }
    __device__ __forceinline__ int32_t long_resid_parts(const LongMeta& m) const { return m.rc > 0 ? (m.rc + li.seg - 1) / li.seg : 0; }

    // intervals U residuals, chunk `part` of the union (stored) or of the interval elements (consumed)
    __device__ __forceinline__ void long_extras_part(const LongRec& L, int32_t part, unsigned long long& acc, long long& arcs) {
        const LongMeta& m = *L.m;
        IntervalSeq a{ li.iv_cum + m.iv_off, li.iv_left + m.iv_off, m.ic, m.ilen };
        const int32_t q0 = part * li.chunk;
        if (L.stored) {
            const int32_t total = m.ilen + m.rc;
            const int32_t* left = a.left;
            const int32_t cnt = min(li.chunk, total - q0);
            const bool final_row = L.fold && m.copied == 0;
            Fold32 f;
            f.begin(m.x);
            merge_chunk(a, [left](int32_t t, int32_t o) { return left[t] + o; }, long_base(m), m.rc, q0, cnt, long_extras_dst(m),
                        final_row ? &f : nullptr);
            if (final_row) { f.n = (uint32_t)cnt; acc ^= f.finish(m.x); arcs += cnt; }
        } else if (L.fold && q0 < m.ilen) {
            const int32_t cnt = min(li.chunk, m.ilen - q0);
            Fold32 f;
            f.begin(m.x);
            int32_t t = upper_slot(a.cum, a.n, q0), t_end = a.cum[t + 1];
            for (int32_t q = q0; q < q0 + cnt; q++) {
                while (q >= t_end) { t++; t_end = a.cum[t + 1]; }
                f.add((uint32_t)(a.left[t] + (q - a.cum[t])));
            }
            f.n = (uint32_t)cnt;
            acc ^= f.finish(m.x);
            arcs += cnt;
        }
    }
    __device__ __forceinline__ int32_t long_extras_parts(const LongRec& L) const {
        const LongMeta& m = *L.m;
        if (m.ic == 0) return 0;
        return ((L.stored ? m.ilen + m.rc : m.ilen) + li.chunk - 1) / li.chunk;
    }

    // masked parent U extras, chunk `part` of the list (stored) or of the copied elements (consumed)
    __device__ __forceinline__ void long_merge_part(const LongRec& L, int32_t part, unsigned long long& acc, long long& arcs) {
        const LongMeta& m = *L.m;
        const int32_t* parent = list_of(L.local - m.ref);
        CopiedSeq a{ li.cb_cum + m.cb_off, li.cb_ppos + m.cb_off, parent, m.ncb, m.copied };
        const int32_t* ppos = a.ppos;
        const int32_t q0 = part * li.chunk;
        if (L.stored) {
            const int32_t cnt = min(li.chunk, m.d - q0);
            Fold32 f;
            f.begin(m.x);
            merge_chunk(a, [ppos, parent](int32_t t, int32_t o) { return parent[ppos[t] + o]; }, long_base(m) + m.d, m.d - m.copied, q0, cnt,
                        long_base(m) + 2 * (int64_t)m.d, L.fold ? &f : nullptr);
            if (L.fold) { f.n = (uint32_t)cnt; acc ^= f.finish(m.x); arcs += cnt; }
        } else if (L.fold && q0 < m.copied) {
            const int32_t cnt = min(li.chunk, m.copied - q0);
            Fold32 f;
            f.begin(m.x);
            int32_t t = upper_slot(a.cum, a.n, q0), t_end = a.cum[t + 1];
            for (int32_t q = q0; q < q0 + cnt; q++) {
                while (q >= t_end) { t++; t_end = a.cum[t + 1]; }
                f.add((uint32_t)parent[ppos[t] + (q - a.cum[t])]);
            }
            f.n = (uint32_t)cnt;
            acc ^= f.finish(m.x);
            arcs += cnt;
        }
    }
    __device__ __forceinline__ int32_t long_merge_parts(const LongRec& L) const {
        const LongMeta& m = *L.m;
        if (m.copied <= 0) return 0;
        return ((L.stored ? m.d : m.copied) + li.chunk - 1) / li.chunk;
    }

    // The three long phases.  The parts of all long records of the tile form one list (long_cum = running part counts,
    // written by one thread before the phase): thread `tid` of `nt` takes parts tid, tid + nt, ... so that the parts of
    // different records run side by side.
    // kind 0: residual segments, 1: extras chunks, 2: merge chunks of the records at `level`
    __device__ __forceinline__ void long_count(int kind, int32_t level) {
        uint32_t run = 0;
        for (int k = 0; k < nlong; k++) {
            sh->long_cum[k] = run;
            const LongRec L = long_rec(k);
            if (!L.active) continue;
            if (kind == 0) run += (uint32_t)long_resid_parts(*L.m);
            else if (kind == 1) run += (uint32_t)long_extras_parts(L);
            else if ((int32_t)lvl[L.local] == level) run += (uint32_t)long_merge_parts(L);
        }
        sh->long_cum[nlong] = run;
    }
    __device__ __forceinline__ void long_phase(int kind, int tid, int nt, unsigned long long& acc, long long& arcs) {
        const uint32_t total = sh->long_cum[nlong];
        int k = 0;
        for (uint32_t it = (uint32_t)tid; it < total; it += (uint32_t)nt) {
            while (sh->long_cum[k + 1] <= it) k++;
            const LongRec L = long_rec(k);
            const int32_t part = (int32_t)(it - sh->long_cum[k]);
            if (kind == 0) long_resid_part(L, part, acc, arcs);
            else if (kind == 1) long_extras_part(L, part, acc, arcs);
            else long_merge_part(L, part, acc, arcs);
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// The planner: tiles are cut greedily along the running node cost (tile_node_cost), at most TILE_MAX_NODES nodes each,
// preferably where no reference crosses the cut (clean[c] = no node >= c copies from a node < c): then the tile needs no
// halo.  One thread walks one super-block of PLAN_SB nodes (cuts at super-block borders are unconditional), twice:
// first to count its tiles, then, after a scan of the counts, to write them.
// ---------------------------------------------------------------------------------------------------
constexpr int32_t PLAN_SB = 1 << 14;
constexpr int32_t PLAN_CLEAN_REACH = 48;   // how far below the greedy cut a clean cut is looked for

// cost[i] for the scan of costs; clean[c] for c in [0, n]
__device__ __forceinline__ void plan_node(const GraphDev& g, int64_t i, int64_t n, const uint8_t* __restrict__ is_parent, int32_t long_d,
                                          uint32_t budget, int32_t* __restrict__ cost, uint8_t* __restrict__ clean) {
    const int32_t d = g.outdeg[i];
    const bool is_long = d > long_d && g.depth[i] >= 0;
    cost[i] = (int32_t)tile_node_cost(g.offsets[i + 1] - g.offsets[i], d, is_long, is_parent[i] != 0, budget);
    bool ok = true;
    for (int64_t y = i; y < n && y < i + g.c.window; y++) if ((int64_t)g.ref[y] > y - i) { ok = false; break; }
    clean[i] = ok ? 1 : 0;
    if (i == n - 1) clean[n] = 1;
}

// One super-block: tiles written to out[] when it is not null; returns how many, or -1 when some node does not fit a tile.
__device__ inline int32_t plan_superblock(const GraphDev& g, int64_t n, const int64_t* __restrict__ cum /* n + 1 */, const uint8_t* __restrict__ clean,
                                          const int32_t* __restrict__ long_nodes, int32_t nlong, uint32_t budget, int32_t sb, int32_t ext_from,
                                          TileEntry* out) {
    // super-blocks are counted from the first node of the extent: halo nodes of a shard are only ever decoded as parents
    const int64_t first = (int64_t)ext_from - g.node_lo;
    const int64_t s0 = first + (int64_t)sb * PLAN_SB, s1 = s0 + PLAN_SB < n ? s0 + PLAN_SB : n;
    int32_t count = 0;
    int64_t s = s0;
    while (s < s1) {
        // the halo: the smallest chain root of the nodes whose chain crosses s.  A crossing chain enters through a node
        // below s + window, so after `window` nodes in a row with their chains inside nothing later reaches back.
        int64_t h = s;
        if (!clean[s]) {
            int64_t reach = s;
            for (int64_t y = s; y < s1; y++) {
                int64_t z = y;
                while (g.ref[z] != 0 && (int64_t)g.ref[z] <= z) z -= g.ref[z];
                if (z < h) h = z;
                if (z < s) reach = y;
                if (y - reach > g.c.window) break;
            }
        }
        // largest e in (s, s1] with cum[e] - cum[h] <= budget and e - h <= TILE_MAX_NODES
        const int64_t base = cum[h];
        int64_t a = s, b = h + TILE_MAX_NODES < s1 ? h + TILE_MAX_NODES : s1;
        while (a < b) {
            const int64_t mid = (a + b + 1) >> 1;
            if (cum[mid] - base <= (int64_t)budget) a = mid; else b = mid - 1;
        }
        {   // at most TILE_MAX_LONG long records in [h, a)
            int32_t la = 0, lb = nlong;
            const int32_t hx = (int32_t)(h + g.node_lo);
            while (la < lb) { const int32_t mid = (la + lb) >> 1; if (long_nodes[mid] < hx) la = mid + 1; else lb = mid; }
            if (la + TILE_MAX_LONG < nlong && (int64_t)long_nodes[la + TILE_MAX_LONG] - g.node_lo < a) a = (int64_t)long_nodes[la + TILE_MAX_LONG] - g.node_lo;
        }
        if (a <= s) return -1;  // a node and the ancestors it copies from do not fit one tile: the caller keeps the general kernels
        int64_t e = a;
        if (e < s1) {  // prefer a cut no reference crosses
            const int64_t floor_ = e - PLAN_CLEAN_REACH > s + 1 ? e - PLAN_CLEAN_REACH : s + 1;
            for (int64_t c = e; c >= floor_; c--) if (clean[c]) { e = c; break; }
        }
        if (out) {
            TileEntry t;
            t.lo = (int32_t)(h + g.node_lo); t.from = (int32_t)(s + g.node_lo); t.hi = (int32_t)(e + g.node_lo);
            // long records of [lo, hi)
            int32_t la = 0, lb = nlong;
            while (la < lb) { const int32_t mid = (la + lb) >> 1; if (long_nodes[mid] < t.lo) la = mid + 1; else lb = mid; }
            t.long_lo = la;
            lb = nlong;
            while (la < lb) { const int32_t mid = (la + lb) >> 1; if (long_nodes[mid] < t.hi) la = mid + 1; else lb = mid; }
            t.long_hi = la;
            const uint64_t bits = g.offsets[e] - g.offsets[h];
            t.cost = (bits >> 6) > 0xffffffffull ? 0xffffffffu : (uint32_t)(bits >> 6);
            out[count] = t;
        }
        count++;
        s = e;
    }
    return count;
}

#ifndef BVG_HOST_EMULATION
__global__ void k_plan_nodes(GraphDev g, const uint8_t* __restrict__ is_parent, int32_t long_d, uint32_t budget,
                             int32_t* __restrict__ cost, uint8_t* __restrict__ clean) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)g.node_hi - g.node_lo;
    if (i < n) plan_node(g, i, n, is_parent, long_d, budget, cost, clean);
}

__global__ void k_plan_tiles(GraphDev g, const int64_t* __restrict__ cum, const uint8_t* __restrict__ clean, const int32_t* __restrict__ long_nodes,
                             int32_t nlong, uint32_t budget, int32_t nsb, int32_t ext_from, int32_t* __restrict__ counts,
                             const int64_t* __restrict__ tile_base, TileEntry* __restrict__ tiles) {
    const int32_t sb = blockIdx.x * blockDim.x + threadIdx.x;
    if (sb >= nsb) return;
    const int64_t n = (int64_t)g.node_hi - g.node_lo;
    if (tiles) (void)plan_superblock(g, n, cum, clean, long_nodes, nlong, budget, sb, ext_from, tiles + tile_base[sb]);
    else counts[sb] = plan_superblock(g, n, cum, clean, long_nodes, nlong, budget, sb, ext_from, nullptr);
}

// ---------------------------------------------------------------------------------------------------
// Device glue
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// warps take `n` consecutive entries of a work list at a time from a ticket in shared memory
__device__ __forceinline__ uint32_t warp_ticket(uint32_t* ctr, uint32_t n) {
    uint32_t t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(ctr, n);
    return __shfl_sync(0xffffffffu, t, 0);
}

template <int K, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_tile_scan(GraphDev g, TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t t_idx = a.order ? a.order[a.first + (int32_t)blockIdx.x] : a.first + (int32_t)blockIdx.x;
    const TileEntry e = a.tiles[t_idx];
    Tile<K> T;
    T.carve(smem, a.smem_bytes, NT, e, g, a);
    unsigned long long acc = 0;
    long long arcs = 0;
    unsigned long long tl[8];
    auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (a.timeline && tid == 0) { tl[0] = (unsigned long long)t_idx; tl[1] = now(); }
    if (tid == 0) {
        mbar_init(&T.sh->mbar, 1);
        typename Tile<K>::RunCopy rc[TILE_MAX_LONG + 1];
        const int n = T.layout(rc);
        uint32_t total = 0;
        for (int r = 0; r < n; r++) total += rc[r].bytes;
        mbar_expect_tx(&T.sh->mbar, total);
        for (int r = 0; r < n; r++)
            if (rc[r].bytes) bulk_g2s(reinterpret_cast<unsigned char*>(T.sw) + rc[r].dst_byte, reinterpret_cast<const unsigned char*>(g.words) + rc[r].src_byte, rc[r].bytes, &T.sh->mbar);
    }
    __syncthreads();
    T.bind_stream();
    T.positions(tid, NT);          // reads the offsets while the stream is on its way
    mbar_wait(&T.sh->mbar, 0);
    __syncthreads();
    T.headers(tid, NT);
    __syncthreads();
    T.levels(tid, NT);
    __syncthreads();
    constexpr int NW = NT / 32;
    if (wid == 0) T.scan_buckets(lane);
    if (wid == 1 % NW) T.scan_levels(lane);
    if (wid == 2 % NW) T.scan_rows(lane);
    __syncthreads();
    T.scatter(tid, NT);
    __syncthreads();
    if (a.timeline && tid == 0) tl[2] = now();
    // level 0: residual segments of the long records, then the short records from the ticket
    if (T.nlong) {
        if (tid == 0) T.long_count(0, 0);
        __syncthreads();
        T.long_phase(0, tid, NT, acc, arcs);
    }
    {
        const uint32_t nE = T.sh->nE;
        unsigned long long w0 = 0, wmax = 0, wbase = 0, nit = 0;
        if (a.timeline) w0 = now();
        for (;;) {
            const uint32_t base = warp_ticket(&T.sh->next_item, 32u);
            if (base >= nE) break;
            const uint32_t it = base + (uint32_t)lane;
            unsigned long long i0 = 0;
            if (a.timeline) i0 = now();
            T.extras_item(it < nE ? (int32_t)T.ordE[it] : -1, acc, arcs);
            if (a.timeline) { __syncwarp(); const unsigned long long dt = now() - i0; nit++; if (dt > wmax) { wmax = dt; wbase = base; } }
        }
        if (a.timeline && lane == 0) {
            unsigned long long* w = a.timeline + (size_t)blockIdx.x * 40 + 8 + (wid & 7) * 4;
            w[0] = w0; w[1] = now(); w[2] = nit; w[3] = (wmax << 32) | wbase;
        }
    }
    if (a.timeline && tid == 0) tl[3] = now();
    __syncthreads();
    if (a.timeline && tid == 0) tl[4] = now();
    if (T.nlong) {
        if (tid == 0) T.long_count(1, 0);
        __syncthreads();
        T.long_phase(1, tid, NT, acc, arcs);
        __syncthreads();
    }
    if (a.timeline && tid == 0) tl[5] = now();
    const int32_t maxlevel = T.sh->maxlevel;
    for (int32_t level = 1; level <= maxlevel; level++) {
        if (T.nlong && tid == 0) T.long_count(2, level);   // read after the barrier that ends the short merges below
        uint32_t la, lb;
        T.level_range(level, la, lb);
        for (uint32_t base = la + (uint32_t)(tid & ~31); base < lb; base += NT) {
            const uint32_t it = base + (uint32_t)lane;
            int32_t i = it < lb ? (int32_t)T.ordM[it] : -1;
            if (level >= TILE_LEVELS && i >= 0 && (int32_t)T.lvl[i] != level) i = -1;  // the deep levels share one list
            T.merge_item(i, tid, NT, acc);
        }
        if (T.nlong) {
            __syncthreads();
            T.long_phase(2, tid, NT, acc, arcs);
        }
        __syncthreads();
    }
    if (a.timeline && tid == 0) {
        tl[6] = now();
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tl[7] = ((unsigned long long)smid << 32) | (unsigned)T.nn | ((unsigned long long)(unsigned)T.nlong << 16);
        for (int i = 0; i < 8; i++) a.timeline[(size_t)blockIdx.x * 40 + i] = tl[i];
    }
    if (a.result) warp_fold(acc, arcs, a.result);
}
#endif  // BVG_HOST_EMULATION

}  // namespace bvg
