// bvg_kernels.cuh -- the general ("any file") kernel set: index build at open, level-synchronous range decode,
// per-query chain decode for random access, consume-only checksum.  One thread walks one record; reference chains
// are resolved level by level (depth of the chain, <= maxrefcount for files the reference's writer produced,
// BVGraph.java:2315,2326), each level reading parents' finished rows.  Two alternatives to the per-record scan kernels
// were built and measured in round 2 and are kept, off by default: the tile kernel (bvg_tile.cuh, BVG_TILE=1) and the
// stream-position extras kernel (bvg_stream.cuh, BVG_STREAM=1); profiles/r02_kernels.md has the numbers.
#pragma once
#include <type_traits>
#include "bvg_device.cuh"
#include "bvg_offsets.cuh"
#include "bvg_scan.cuh"

namespace bvg {

// ---------------------------------------------------------------------------------------------------
// Load-time kernels
// ---------------------------------------------------------------------------------------------------

// The file's bytes -> big-endian 32-bit words, in place (one-time, at open).
__global__ void k_bswap(uint32_t* __restrict__ w, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = __byte_perm(w[i], 0, 0x0123);
}

// Small host-to-device transfers of an index build, pulled by a kernel out of pinned staging memory instead of queued on
// the copy engine, whose queue is shared by all streams.  Both pointers 16-byte aligned; bytes a multiple of 4.
__global__ void k_pull_copy(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t nwords) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) dst[i] = src[i];
}
__global__ void k_set_i32(int32_t* __restrict__ p, int32_t v) { *p = v; }

// Header pass: outdegree and reference of every loaded node (BVGraph.java:1048-1053; outdegree(x) :857-879).
template <bool DEF>
__global__ void k_header(GraphDev g, int32_t* __restrict__ outdeg, int32_t* __restrict__ ref) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)g.node_hi - g.node_lo;
    if (i >= n) return;
    const int32_t x = g.node_lo + (int32_t)i;
    Bits b = cursor_at(g, x);
    const uint64_t limit = g.bit_end - g.bit_base;
    const uint64_t d = Rd<DEF>::outdeg(b, g.c);
    int32_t r = 0;
    if (d > 0x7fffffffull || b.pos > limit) { report(g.err, E_IO, x, b.pos + g.bit_base); outdeg[i] = 0; ref[i] = 0; return; }
    if (d > 0 && g.c.window > 0) {
        const uint64_t rr = Rd<DEF>::ref(b, g.c);
        if (rr > (uint64_t)g.c.window) report(g.err, E_STATE, x, b.pos + g.bit_base);        // BVGraph.java:705
        else if ((int64_t)rr > (int64_t)x) report(g.err, E_FORMAT, x, b.pos + g.bit_base);    // would address node < 0
        else r = (int32_t)rr;
        if (b.pos > limit) report(g.err, E_IO, x, b.pos + g.bit_base);
    }
    outdeg[i] = (int32_t)d;
    ref[i] = r;
}

// Chain depth of every node; -1 when the chain leaves the loaded window (only possible in a shard's halo).
__global__ void k_depth(GraphDev g, int32_t* __restrict__ depth, int32_t* __restrict__ maxdepth, int32_t ext_from) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)g.node_hi - g.node_lo;
    if (i >= n) return;
    int64_t y = i;
    int32_t dep = 0;
    for (;;) {
        const int32_t r = g.ref[y];
        if (r == 0) break;
        if (r > y) { dep = -1; break; }
        y -= r;
        dep++;
    }
    depth[i] = dep;
    if (dep > 0) atomicMax(maxdepth, dep);
    // inside the extent every chain must close within the loaded window (it does when the file honours maxrefcount)
    if (dep < 0 && g.node_lo + (int32_t)i >= ext_from) report(g.err, E_FORMAT, g.node_lo + (int32_t)i, g.offsets[i]);
}

// Bit-balanced cuts of the node range [from, to) (bvg_plan_shards on the device): bounds[i] = first node whose record starts
// at or after offsets[from] + i * bits / pieces; one thread per cut.  taper != 0: the pieces shrink (> 0) or grow (< 0) linearly instead of
// being equal (bvg_scan_memory; measured).
__global__ void k_plan_cuts(const uint64_t* __restrict__ offsets, int32_t from, int32_t to, int32_t pieces, int32_t* __restrict__ bounds, int32_t taper = 0) {
    const uint64_t first = offsets[from], total = offsets[to] - first;
    for (int32_t i = threadIdx.x; i <= pieces; i += blockDim.x) {
        if (i == 0) { bounds[0] = from; continue; }
        if (i == pieces) { bounds[pieces] = to; continue; }
        uint64_t target = first + total / (uint64_t)pieces * (uint64_t)i;
        if (taper != 0) {
            // taper > 0: weights taper + pieces - 1, ..., taper (shrinking pieces); taper < 0: |taper|, ..., |taper| + pieces - 1 (growing)
            const uint64_t t = (uint64_t)(taper > 0 ? taper : -taper);
            const uint64_t wsum = (uint64_t)pieces * t + (uint64_t)pieces * (uint64_t)(pieces - 1) / 2;
            const uint64_t wi = taper > 0 ? (uint64_t)i * t + (uint64_t)i * (uint64_t)(2 * pieces - 1 - i) / 2
                                          : (uint64_t)i * t + (uint64_t)i * (uint64_t)(i - 1) / 2;
            target = first + (uint64_t)((double)total * ((double)wi / (double)wsum));
        }
        int32_t lo = from, hi = to;  // first node with offsets[node] >= target
        while (lo < hi) {
            const int32_t mid = lo + (hi - lo) / 2;
            if (offsets[mid] < target) lo = mid + 1; else hi = mid;
        }
        bounds[i] = lo;
    }
}

__global__ void k_max_i32(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
    int32_t m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, in[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// ---------------------------------------------------------------------------------------------------
// Length-bucketed schedules.  One thread walks one record, so a warp runs as long as its longest record; with
// power-law outdegrees natural node order wastes ~90 % of the lanes.  At open the nodes are counting-sorted by a
// half-octave bucket of their work (record bits for the extras step; chain level, then outdegree + parent outdegree
// for the merge step), longest first, so the 32 lanes of a warp get records of similar length.
// ---------------------------------------------------------------------------------------------------
constexpr int ORDER_BUCKETS = 256;
// Schedules are bucketed inside chunks of 2^ORDER_CHUNK_LOG consecutive nodes, chunks in node order: the lanes of a
// warp still get records of similar length, but everything in flight at one time comes from a few tens of MB of the
// stream, the index arrays and the rows, which the 126 MB L2 can hold.
constexpr int ORDER_CHUNK_LOG = 18;
// Heavy records (work >= ORDER_HEAVY) of all chunks are scheduled first, before chunk 0: a record is one lane's serial
// work, so a kernel cannot end before the heaviest record of its LAST chunk has been walked -- ~0.2 ms per launch on the
// benchmark graph when the heavy records start with their chunk.
constexpr int ORDER_HEAVY = 192;


// key_e: extras schedule (all nodes with successors); key_m: merge schedule (nodes with a reference), level-major.
// The merge work of a node is its block list (walked twice) plus the elements it copies, plus -- when its own list has
// to be materialised because somebody copies from it -- the extras it merges them with; the block list is parsed
// here once to learn that, and the copied count is kept for the merge step.
template <bool DEF>
__global__ void k_order_keys(GraphDev g, int32_t* __restrict__ key_e, int32_t* __restrict__ key_m, int32_t max_level_keys,
                             int32_t long_d, const uint8_t* __restrict__ is_parent, int32_t* __restrict__ copied_out,
                             uint64_t* __restrict__ extras_pos, uint64_t* __restrict__ blocks_pos, int32_t* __restrict__ bc_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)g.node_hi - g.node_lo;
    if (i >= n) return;
    const int32_t d = g.outdeg[i], dep = g.depth[i];
    const int32_t chunk = (int32_t)(i >> ORDER_CHUNK_LOG);
    int32_t ke = -1, km = -1, copied = 0;
    uint64_t epos = 0, bpos = 0;
    int64_t bc = 0;
    if (d > 0 && dep >= 0) {
        const int32_t x = g.node_lo + (int32_t)i;
        const uint64_t limit = g.bit_end - g.bit_base;
        BitBuf b = buffer_at(g, x);
        (void)Rd<DEF>::outdeg(b, g.c);
        if (g.c.window > 0) {
            const int32_t r = (int32_t)Rd<DEF>::ref(b, g.c);
            if (r > 0) {
                bc = (int64_t)Rd<DEF>::bcount(b, g.c);
                bpos = b.pos();
                int64_t total = 0, cp = 0;
                for (int64_t k = 0; k < bc && b.pos() <= limit; k++) {
                    const int64_t blk = (int64_t)Rd<DEF>::block(b, g.c) + (k ? 1 : 0);
                    total += blk;
                    if (!(k & 1)) cp += blk;
                }
                if (!(bc & 1)) cp += (int64_t)g.outdeg[i - r] - total;
                copied = (int32_t)(cp < 0 ? 0 : (cp > d ? d : cp));  // malformed records are reported by the decode step
            }
        }
        epos = b.pos();
        // records with intervals take a different loop than records without: keep the two kinds in separate warps
        const int has_iv = (d > copied && g.c.minlen != 0 && b.pos() <= limit && b.gamma() != 0) ? 1 : 0;
        if (d <= long_d) {  // longer records are split across threads (bvg_long.cuh)
            // what a lane's loop length is: residual count for the tight loop, record bits for the interval loop
            // (and records whose list somebody copies from store it: a third kind of loop, see k_scan_extras_lean)
            const uint64_t work_e = has_iv ? g.offsets[i + 1] - g.offsets[i] : (uint64_t)(d - copied);
            const uint64_t serial_e = (uint64_t)(d - copied);  // trips of the lane that walks it
            const int32_t slot_e = serial_e >= (uint64_t)ORDER_HEAVY ? 0 : chunk + 1;
            ke = (slot_e * 4 + has_iv * 2 + (is_parent[i] ? 1 : 0)) * ORDER_BUCKETS + (ORDER_BUCKETS - 1 - half_octave_bucket(work_e));
            if (dep >= 1 && dep <= max_level_keys) {
                const int par = is_parent[i] ? 1 : 0;  // parents merge in place, the others only stream: separate warps too
                const uint64_t work = 2 * (uint64_t)bc + (uint64_t)copied + (par ? (uint64_t)(d - copied) : 0);
                const int32_t nslots = (int32_t)((n + ((int64_t)1 << ORDER_CHUNK_LOG) - 1) >> ORDER_CHUNK_LOG) + 1;
                const int32_t slot_m = work >= (uint64_t)ORDER_HEAVY ? 0 : chunk + 1;
                km = (((dep - 1) * nslots + slot_m) * 2 + par) * ORDER_BUCKETS + (ORDER_BUCKETS - 1 - half_octave_bucket(work + 1));
            }
        }
    }
    key_e[i] = ke;
    key_m[i] = km;
    copied_out[i] = copied;
    extras_pos[i] = epos;
    blocks_pos[i] = bpos;
    bc_out[i] = (int32_t)bc;
}

// Schedule records: everything the scan kernels need about a node in one coalesced 32- / 56-byte load, in schedule
// order, instead of six scattered index reads and a second parse of the record header per node and scan.
struct ExtraRec {
    uint64_t pos;    // bit position (relative to word 0) of the extras section: interval count, else first residual
    int64_t row;     // CSR-space offset of the first extra: rowoff[x] + copied
    int32_t x, nout; // node, successors that are not copied (d - copied)
    int32_t d;       // outdegree
    uint32_t flags;  // bit 0: somebody copies from x (its list is materialised); bit 1: the record has intervals
};
struct MergeRec {
    uint64_t pos;    // bit position of the first copy-block code
    int64_t row;     // CSR-space offset of x's row
    int64_t prow;    // CSR-space offset of the parent's row
    int32_t x, px;   // node, parent
    int32_t d, dp;   // outdegrees of both
    int32_t bc, copied;
    uint32_t flags;  // bit 0 as above
    int32_t pad_;
};

__global__ void k_key_scatter_recs(GraphDev g, const int32_t* __restrict__ key_e, const int32_t* __restrict__ key_m, int64_t n,
                                   int32_t* __restrict__ cur_e, int32_t* __restrict__ cur_m, const uint8_t* __restrict__ is_parent,
                                   const int32_t* __restrict__ copied, const uint64_t* __restrict__ extras_pos,
                                   const uint64_t* __restrict__ blocks_pos, const int32_t* __restrict__ bc,
                                   int32_t* __restrict__ order_e, int32_t* __restrict__ order_m,
                                   ExtraRec* __restrict__ rec_e, MergeRec* __restrict__ rec_m) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t x = g.node_lo + (int32_t)i;
    const int32_t ke = key_e[i], km = key_m[i];
    if (ke >= 0) {
        const int32_t slot = atomicAdd(cur_e + ke, 1);
        order_e[slot] = x;
        ExtraRec r;
        r.pos = extras_pos[i]; r.row = g.rowoff[i] + copied[i]; r.x = x; r.d = g.outdeg[i]; r.nout = r.d - copied[i];
        r.flags = (is_parent[i] ? 1u : 0u) | ((((uint32_t)ke / ORDER_BUCKETS) & 2u) ? 2u : 0u);
        rec_e[slot] = r;
    }
    if (km >= 0) {
        const int32_t slot = atomicAdd(cur_m + km, 1);
        order_m[slot] = x;
        const int32_t rf = g.ref[i];
        MergeRec r;
        r.pos = blocks_pos[i]; r.row = g.rowoff[i]; r.prow = g.rowoff[i - rf]; r.x = x; r.px = x - rf;
        r.d = g.outdeg[i]; r.dp = g.outdeg[i - rf]; r.bc = bc[i]; r.copied = copied[i];
        r.flags = is_parent[i] ? 1u : 0u; r.pad_ = 0;
        rec_m[slot] = r;
    }
}

__global__ void k_key_hist(const int32_t* __restrict__ keys, int64_t n, int32_t* __restrict__ bins) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t k = keys[i];
    if (k >= 0) atomicAdd(bins + k, 1);
}

__global__ void k_key_scatter(const int32_t* __restrict__ keys, int64_t n, int32_t* __restrict__ cursors,
                              int32_t node_lo, int32_t* __restrict__ order) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t k = keys[i];
    if (k >= 0) order[atomicAdd(cursors + k, 1)] = node_lo + (int32_t)i;
}

__global__ void k_long_flags(GraphDev g, int32_t long_d, int32_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)g.node_hi - g.node_lo) return;
    flags[i] = (g.outdeg[i] > long_d && g.depth[i] >= 0) ? 1 : 0;
}

__global__ void k_long_compact(const int32_t* __restrict__ flags, const int64_t* __restrict__ pos, int64_t n, int32_t node_lo,
                               int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[pos[i]] = node_lo + (int32_t)i;
}

// Exclusive scan int32 -> int64, three phases, 2048 items per block.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* total) {
    __shared__ int64_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    int64_t base = 0;
    int64_t tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        const int64_t s = warp_sums[w];
        if (w < wid) base += s;
        tot += s;
    }
    __syncthreads();
    if (total) *total = tot;
    return base + inc - v;
}

__global__ void k_scan_sums(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ block_sums) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    int64_t tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void k_scan_blocks(int64_t* __restrict__ block_sums, int64_t nblocks) {  // one block
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < nblocks ? block_sums[i] : 0;
        int64_t tot;
        const int64_t ex = block_exclusive_scan(v, &tot);
        const int64_t c = carry;
        if (i < nblocks) block_sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
}

__global__ void k_scan_apply(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ block_sums,
                             int64_t* __restrict__ out /* n+1 */) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    int64_t run = block_sums[blockIdx.x] + block_exclusive_scan(s, nullptr);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
        if (base + k == n - 1) out[n] = run;
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

// Exclusive prefixes of the counts and sums of the .offsets sub-ranges (bvg_offsets.cuh), one block: every thread takes a
// contiguous stretch.  cbase / sbase get nsub entries, totals[0] / totals[1] the grand totals.  (The sums are 64-bit gaps, so
// the int32 scan above does not serve.)
__global__ void __launch_bounds__(SCAN_THREADS) k_off_prefix(const OffSub* __restrict__ sub, int64_t nsub, int64_t* __restrict__ cbase,
                                                             uint64_t* __restrict__ sbase, unsigned long long* __restrict__ totals) {
    const int64_t per = (nsub + SCAN_THREADS - 1) / SCAN_THREADS;
    const int64_t a = (int64_t)threadIdx.x * per, b = a + per < nsub ? a + per : nsub;
    int64_t c = 0;
    uint64_t sm = 0;
    for (int64_t j = a; j < b; j++) { c += sub[j].count; sm += sub[j].sum; }
    int64_t ctot, stot;
    int64_t cb = block_exclusive_scan(c, &ctot);
    uint64_t sb = (uint64_t)block_exclusive_scan((int64_t)sm, &stot);   // wrap-around arithmetic: exact modulo 2^64
    for (int64_t j = a; j < b; j++) {
        cbase[j] = cb; sbase[j] = sb;
        cb += sub[j].count; sb += sub[j].sum;
    }
    if (threadIdx.x == 0) { totals[0] = (unsigned long long)ctot; totals[1] = (unsigned long long)stot; }
}

// ---------------------------------------------------------------------------------------------------
// Sequential range decode (general path)
// ---------------------------------------------------------------------------------------------------

// Where the row of node x lives: nodes of the requested range in `out`, halo nodes (ancestors before `from`)
// in `halo` (self-decoded or imported from the previous shard).
struct RowMap {
    int32_t* out;
    int64_t out_base;          // rowoff of `from`
    int32_t from;
    int32_t* halo;
    const int64_t* halo_off;   // [x - halo_lo], arc offsets into halo
    int32_t halo_lo;
    __device__ __forceinline__ int32_t* row(const GraphDev& g, int32_t x) const {
        return x >= from ? out + (g.rowoff[x - g.node_lo] - out_base) : halo + halo_off[x - halo_lo];
    }
    int64_t halo_base;         // rowoff of halo_lo (both halo kinds are laid out like the CSR from there)
    const uint8_t* mask;       // when set: decode exactly the nodes with mask[x - node_lo] != 0 (indexed like the graph's arrays)
    // row pointer of node y from its CSR-space offset (schedule records carry the offsets)
    __device__ __forceinline__ int32_t* at(int32_t y, int64_t off) const {
        return y >= from ? out + (off - out_base) : halo + (off - halo_base);
    }
    // A halo node is decoded only if its whole chain lies inside the halo: halo_lo is the smallest chain root of the
    // requested range, so a halo node whose chain starts before it is nobody's ancestor (and its parent has no row).
    __device__ __forceinline__ bool wanted(const GraphDev& g, int32_t x) const {
        if (mask) return mask[x - g.node_lo] != 0;  // random-access batches: only the nodes on the queried chains
        if (x >= from) return true;
        int64_t y = (int64_t)x - g.node_lo;
        for (;;) {
            const int32_t r = g.ref[y];
            if (r == 0) break;
            if (r > y) return false;
            y -= r;
        }
        return y + g.node_lo >= halo_lo;
    }
};

// Level 0 for every node of [lo, hi): extras into the row tail; nodes without a reference are complete.
template <bool DEF>
__global__ void k_extras(GraphDev g, int32_t lo, int32_t hi, RowMap rm, int32_t skip_above) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)hi - lo) return;
    const int32_t x = lo + (int32_t)i;
    if (g.outdeg[x - g.node_lo] == 0 || g.depth[x - g.node_lo] < 0) return;  // depth < 0: chain leaves the shard
    if (g.outdeg[x - g.node_lo] > skip_above) return;                          // split across threads by the k_long_* kernels
    if (!rm.wanted(g, x)) return;
    decode_extras<DEF>(g, x, rm.row(g, x));
}

// Level l >= 1: nodes whose chain depth is l merge their parent's (finished) row into their own.
template <bool DEF>
__global__ void k_merge(GraphDev g, int32_t lo, int32_t hi, int32_t level, RowMap rm, int32_t skip_above) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)hi - lo) return;
    const int32_t x = lo + (int32_t)i;
    if (g.depth[x - g.node_lo] != level || !rm.wanted(g, x)) return;
    if (g.outdeg[x - g.node_lo] > skip_above) return;
    merge_copied<DEF>(g, x, rm.row(g, x), rm.row(g, x - g.ref[x - g.node_lo]));
}

// The same two steps over a length-bucketed schedule (order[0..count)): node ids outside [lo, hi) are skipped.
// Phases are separated by __syncwarp() so that the 32 lanes enter each loop together (see ExtrasWalk).
template <bool DEF>
__global__ void k_extras_ordered(GraphDev g, const int32_t* __restrict__ order, int64_t count, int32_t lo, int32_t hi, RowMap rm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t x = i < count ? order[i] : -1;
    const bool active = x >= lo && x < hi && rm.wanted(g, x);
    ExtrasWalk<DEF> w;
    w.header(g, x, active);
    int32_t* row = active ? rm.row(g, x) : nullptr;
    unsigned long long f = 0;
    __syncwarp();
    w.template residuals_only<false>(g, row, true, f);
    __syncwarp();
    w.template with_intervals<false>(g, row, true, f);
}

template <bool DEF>
__global__ void k_merge_ordered(GraphDev g, const int32_t* __restrict__ order, int64_t count, int32_t lo, int32_t hi, RowMap rm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t x = i < count ? order[i] : -1;
    const bool active = x >= lo && x < hi && rm.wanted(g, x);
    MergeWalk<DEF> w;
    w.header(g, x, active ? rm.row(g, x - g.ref[x - g.node_lo]) : nullptr, active);
    unsigned long long f = 0;
    __syncwarp();
    if (active) w.template merge_in_place<false>(g, rm.row(g, x), f);
}

// First node any chain starting in [from, to) reaches back to (the halo a range decode has to supply).
__global__ void k_halo_start(GraphDev g, int32_t from, int32_t count, int32_t* __restrict__ result) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int64_t y = (int64_t)from + i - g.node_lo;
    for (;;) {
        const int32_t r = g.ref[y];
        if (r == 0 || r > y) break;
        y -= r;
    }
    const int32_t root = (int32_t)(y + g.node_lo);
    if (root < from) atomicMin(result, root);
}

// Imported boundary lists (bvg_halo_import): offsets and lists copied on the device; the arc count is read here, not on the host.
__global__ void k_halo_copy(const int64_t* __restrict__ off, const int32_t* __restrict__ lists, int32_t count,
                            int64_t* __restrict__ dst_off, int32_t* __restrict__ dst_lists, int64_t cap, ErrWord* err) {
    const int64_t total = off[count];
    if (total < 0 || total > cap) { if (blockIdx.x == 0 && threadIdx.x == 0) report(err, E_NOMEM, -1, 0); return; }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t0; i <= count; i += stride) dst_off[i] = off[i];
    for (int64_t i = t0; i < total; i += stride) dst_lists[i] = lists[i];
}

__global__ void k_rel_offsets(const int64_t* __restrict__ rowoff, int64_t count, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= count) out[i] = rowoff[i] - rowoff[0];
}

// Consume: one warp per node folds the node's row into (arcs, xor checksum); one atomic pair per block.
__global__ void k_checksum(const int32_t* __restrict__ rows, const int64_t* __restrict__ rowoff /* at `from` */,
                           int32_t from, int64_t count, unsigned long long* __restrict__ result /* arcs, xor */) {
    __shared__ unsigned long long s_x[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned long long acc = 0;
    long long arcs = 0;
    for (int64_t i = (int64_t)blockIdx.x * nw + wid; i < count; i += (int64_t)gridDim.x * nw) {
        const int64_t a = rowoff[i] - rowoff[0], e = rowoff[i + 1] - rowoff[0];
        const unsigned long long base = (unsigned long long)(uint32_t)(from + (int32_t)i) * 0x9E3779B97F4A7C15ull;
        for (int64_t p = a + lane; p < e; p += 32) acc ^= base + (unsigned long long)(uint32_t)rows[p];
        if (lane == 0) arcs += e - a;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_x[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        unsigned long long v = lane < nw ? s_x[lane] : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicXor(result + 1, v);
    }
    if (lane == 0 && arcs) atomicAdd(result, (unsigned long long)arcs);
}

// ---------------------------------------------------------------------------------------------------
// Consume-only scan, fused: successors are folded into (arcs, XOR checksum) as they are decoded.  Only lists somebody
// copies from (is_parent) are materialised, in a scratch addressed like the CSR; every other node's successors -- its
// extras and the elements it copies through its block list -- are consumed straight out of registers.
// Persistent grid-stride kernels over the length-bucketed schedules; one atomic pair per block.
// ---------------------------------------------------------------------------------------------------
__global__ void k_mark_parents(GraphDev g, uint8_t* __restrict__ is_parent) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)g.node_hi - g.node_lo) return;
    const int32_t r = g.ref[i];
    if (r > 0 && r <= i) is_parent[i - r] = 1;
}

__device__ __forceinline__ void block_fold(unsigned long long acc, long long arcs, unsigned long long* __restrict__ result) {
    __shared__ unsigned long long s_x[32];
    __shared__ long long s_a[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
        arcs += __shfl_xor_sync(0xffffffffu, arcs, o);
    }
    if (lane == 0) { s_x[wid] = acc; s_a[wid] = arcs; }
    __syncthreads();
    if (wid == 0) {
        unsigned long long v = lane < nw ? s_x[lane] : 0;
        long long a = lane < nw ? s_a[lane] : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            v ^= __shfl_xor_sync(0xffffffffu, v, o);
            a += __shfl_xor_sync(0xffffffffu, a, o);
        }
        if (lane == 0) {  // spread over FOLD_SLOTS slot pairs: a dynamic grid has ~10^5 blocks, two hot words would serialise them
            unsigned long long* slot = result + 2 * (blockIdx.x & (FOLD_SLOTS - 1));
            if (v) atomicXor(slot + 1, v);
            if (a) atomicAdd(slot, (unsigned long long)a);
        }
    }
}

// slots[FOLD_SLOTS][2] -> out[2] (arcs added, checksum xored)
__global__ void k_reduce_slots(const unsigned long long* __restrict__ slots, unsigned long long* __restrict__ out) {
    unsigned long long a = 0, v = 0;
    for (int i = threadIdx.x; i < FOLD_SLOTS; i += blockDim.x) { a += slots[2 * i]; v ^= slots[2 * i + 1]; }
    __shared__ unsigned long long sa[32], sv[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); v ^= __shfl_xor_sync(0xffffffffu, v, o); }
    if (lane == 0) { sa[wid] = a; sv[wid] = v; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0; v = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += sa[w]; v ^= sv[w]; }
        atomicAdd(out, a);
        atomicXor(out + 1, v);
    }
}

constexpr int SCAN_BLOCK = 128;
constexpr int SCAN_BLOCKS_PER_SM = 8;

template <bool DEF>
__global__ void __launch_bounds__(SCAN_BLOCK, SCAN_BLOCKS_PER_SM) k_scan_extras(GraphDev g, const ExtraRec* __restrict__ recs, int64_t count,
                              int32_t lo, int32_t hi, int32_t from, RowMap rm, unsigned long long* __restrict__ result) {
    unsigned long long acc = 0;
    long long arcs = 0;
    // warp-uniform trip count and explicit reconvergence between the phases of each item: the lanes of a warp walk
    // 32 records of similar length, and they only do so in lockstep if they enter each loop together
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < count; base += stride) {
        const int64_t i = base + (threadIdx.x & 31);
        ExtraRec r;
        r.x = -1;
        if (i < count) r = recs[i];
        bool active = r.x >= lo && r.x < hi && rm.wanted(g, r.x);
        const bool store = active && (r.flags & 1u);
        const bool fold = active && r.x >= from;
        active = active && (fold || store);  // halo nodes matter only as parents
        ExtrasWalk<DEF> w;
        w.header_rec(g, r.x, r.d, r.nout, r.pos, active);
        int32_t* row = store ? rm.at(r.x, r.row) : nullptr;
        unsigned long long f = 0;
        __syncwarp();
        w.template residuals_only<true>(g, row, store, f);
        __syncwarp();
        w.template with_intervals<true>(g, row, store, f);
        __syncwarp();
        if (fold) { acc ^= f; arcs += w.d; }
    }
    block_fold(acc, arcs, result);
}

// The same step for the default codings with the lean walkers of bvg_scan.cuh.  The schedule keeps four kinds of
// records in separate warps: (no intervals | intervals) x (only consumed | stored because somebody copies from it).
// Every kind runs the same tight residual loop, preceded for records with intervals by a walk of the interval section
// that folds its elements; stored records with intervals write their residuals right-aligned and merge the intervals in
// front of them in a second walk (ScanExtras::iv_merge).
template <int K, bool RING, bool HIST = false>
__global__ void __launch_bounds__(SCAN_BLOCK, SCAN_BLOCKS_PER_SM) k_scan_extras_lean(GraphDev g, const ExtraRec* __restrict__ recs, int64_t count,
                              int32_t lo, int32_t hi, int32_t from, RowMap rm, unsigned long long* __restrict__ result, int debug_nostore, int store_all, int items) {
    __shared__ uint4 ring[RING ? RING_GROUPS * SCAN_BLOCK : 1];
    typedef typename std::conditional<RING, WinRing<SCAN_BLOCK>, Win>::type W;
    const ring_addr my_ring = ring_address(&ring[RING ? threadIdx.x : 0]);
    unsigned long long acc = 0;
    long long arcs = 0;
    // `items` consecutive groups of 128 schedule records per block (neighbours in the schedule: same chunk, similar length)
    for (int32_t it = 0; it < items; it++) {
        const int64_t base = ((int64_t)blockIdx.x * items + it) * blockDim.x + (threadIdx.x & ~31);
        if (base >= count) break;
        const int64_t i = base + (threadIdx.x & 31);
        ExtraRec r;
        r.x = -1; r.flags = 0;
        if (i < count) r = recs[i];
        bool active = r.x >= lo && r.x < hi && rm.wanted(g, r.x);
        const bool store = active && ((r.flags & 1u) || store_all) && debug_nostore != 1;  // store_all: range decode, every list is materialised
        const bool fold = active && r.x >= from && !store_all;
        active = active && (fold || store);  // halo nodes matter only as parents
        const bool has_iv = (r.flags & 2u) != 0;
        int32_t* row = store ? rm.at(r.x, r.row) : nullptr;
        unsigned long long f = 0;
        ScanExtras<K, W, HIST> w;
        w.begin(g, r.x, r.nout, r.pos, active, my_ring, fold);
        if (has_iv) w.iv_fold(g); else w.iv_none(g);
        __syncwarp();
        if (__any_sync(0xffffffffu, store)) w.template resid<true>(g, row, store);
        else w.template resid<false>(g, row, false);
        __syncwarp();
        if (store && has_iv) w.iv_merge(g, row);
        __syncwarp();
        if (active) f = w.finish();
        if (fold) { acc ^= f; arcs += r.d; }
    }
    if (result) warp_fold(acc, arcs, result);
}

template <bool DEF>
__global__ void __launch_bounds__(SCAN_BLOCK, SCAN_BLOCKS_PER_SM) k_scan_merge(GraphDev g, const MergeRec* __restrict__ recs, int64_t count,
                             int32_t lo, int32_t hi, int32_t from, RowMap rm, unsigned long long* __restrict__ result) {
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < count; base += stride) {
        const int64_t i = base + (threadIdx.x & 31);
        MergeRec r;
        r.x = -1;
        if (i < count) r = recs[i];
        bool active = r.x >= lo && r.x < hi && rm.wanted(g, r.x);
        const bool store = active && (r.flags & 1u);
        const bool fold = active && r.x >= from;
        active = active && (fold || store);
        MergeWalk<DEF> w;
        w.header_rec(g, r.x, r.d, r.dp, r.bc, r.copied, r.pos, active ? rm.at(r.px, r.prow) : nullptr, active);
        int32_t* row = store ? rm.at(r.x, r.row) : nullptr;
        unsigned long long f = 0;
        __syncwarp();
        if (!store) w.stream_only(g, f);
        __syncwarp();
        if (store) w.template merge_in_place<true>(g, row, f);
        __syncwarp();
        if (fold) acc ^= f;
    }
    block_fold(acc, 0, result);
}

// Merge step with the staged copy runs of bvg_scan.cuh (default codings).
template <int MINB, int BATCH, bool HIST = false>
__global__ void __launch_bounds__(SCAN_BLOCK, MINB) k_scan_merge_lean(GraphDev g, const MergeRec* __restrict__ recs, int64_t count,
                             int32_t lo, int32_t hi, int32_t from, RowMap rm, unsigned long long* __restrict__ result, int store_all) {
    __shared__ int32_t runs[2 * COPY_RUNS * SCAN_BLOCK];
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < count; base += stride) {
        const int64_t i = base + (threadIdx.x & 31);
        MergeRec r;
        r.x = -1; r.flags = 0;
        if (i < count) r = recs[i];
        bool active = r.x >= lo && r.x < hi && rm.wanted(g, r.x);
        const bool store = active && ((r.flags & 1u) || store_all);
        const bool fold = active && r.x >= from && !store_all;
        active = active && (fold || store);
        CopyRuns c;
        c.begin(g, r.pos, r.bc, r.dp, runs + threadIdx.x, SCAN_BLOCK, active);
        c.stage(g);
        __syncwarp();
        const int32_t* parent = active ? rm.at(r.px, r.prow) : nullptr;
        unsigned long long f = 0;
        if (active && !store) f = copied_fold<BATCH, CopyRuns, HIST>(g, c, r.x, parent);
        __syncwarp();
        if (store) f = copied_merge<CopyRuns, HIST>(g, c, r.x, r.d, r.copied, rm.at(r.x, r.row), parent, fold);
        __syncwarp();
        if (fold) acc ^= f;
    }
    if (result) warp_fold(acc, 0, result);
}

// ---------------------------------------------------------------------------------------------------
// Random access (general path): one thread per query decodes the query's whole reference chain, root first
// (the reference recurses lazily down the chain instead, BVGraph.java:1110-1121).
// ---------------------------------------------------------------------------------------------------

// Per query: outdegree of x and scratch needed for its strict ancestors' rows.
// A query whose chain holds a record of more than `long_d` successors is not walked by one thread: its chain is marked
// in `mask` (when given), decoded by the range kernels -- which split long records across threads -- into a scratch laid
// out like the CSR, and copied out by k_gather_rows.  heavy[q] = 1 for those queries, and they need no chain scratch.
__global__ void k_query_sizes(GraphDev g, const int32_t* __restrict__ xs, int64_t nx,
                              int32_t* __restrict__ dq, int32_t* __restrict__ need,
                              int32_t long_d, uint8_t* __restrict__ mask, uint8_t* __restrict__ heavy, int32_t* __restrict__ nheavy) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nx) return;
    const int32_t x = xs[q];
    if (heavy) heavy[q] = 0;
    if (x < g.node_lo || x >= g.node_hi) { report(g.err, E_INVAL, x, 0); dq[q] = 0; need[q] = 0; return; }
    int64_t y = x - g.node_lo;
    dq[q] = g.outdeg[y];
    if (g.depth[y] < 0) { report(g.err, E_FORMAT, x, 0); need[q] = 0; return; }
    int64_t s = 0;
    bool is_heavy = mask != nullptr && g.outdeg[y] > long_d;
    for (;;) {
        const int32_t r = g.ref[y];
        if (r == 0) break;
        y -= r;
        s += g.outdeg[y];
        is_heavy = is_heavy || (mask != nullptr && g.outdeg[y] > long_d);
    }
    if (is_heavy) {
        heavy[q] = 1;
        atomicAdd(nheavy, 1);
        y = x - g.node_lo;
        for (;;) {
            mask[y] = 1;
            const int32_t r = g.ref[y];
            if (r == 0) break;
            y -= r;
        }
        s = 0;
    }
    need[q] = (int32_t)(s > 0x7fffffff ? 0x7fffffff : s);
}

// rows of the heavy queries, decoded into the CSR-shaped scratch, copied to their place in the batch output: one warp per query
// (eight lanes per query with the wide rows left to the whole warp measured 2.7 x slower: 32-byte pieces instead of 128-byte ones)
__global__ void k_gather_rows(GraphDev g, const int32_t* __restrict__ xs, int64_t nx, const uint8_t* __restrict__ heavy,
                              const int64_t* __restrict__ out_off, int32_t* __restrict__ out, RowMap rm) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nx; q += nw) {
        if (!heavy[q]) continue;
        const int32_t x = xs[q];
        const int32_t d = g.outdeg[x - g.node_lo];
        const int32_t* row = rm.row(g, x);
        int32_t* dst = out + out_off[q];
        for (int32_t i = lane; i < d; i += 32) dst[i] = row[i];
    }
}

template <bool DEF>
__global__ void k_random(GraphDev g, const int32_t* __restrict__ xs, int64_t nx, const int64_t* __restrict__ out_off,
                         int32_t* __restrict__ out, const int64_t* __restrict__ scratch_off, int32_t* __restrict__ scratch,
                         const uint8_t* __restrict__ heavy) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nx) return;
    if (heavy && heavy[q]) return;  // decoded by the range kernels, copied by k_gather_rows
    const int32_t x = xs[q];
    if (x < g.node_lo || x >= g.node_hi) return;
    const int32_t dep = g.depth[x - g.node_lo];
    if (dep < 0 || g.outdeg[x - g.node_lo] == 0) return;
    int32_t* cur = scratch + scratch_off[q];
    const int32_t* parent = nullptr;
    for (int32_t level = dep; level >= 0; level--) {
        int32_t y = x;
        for (int32_t s = 0; s < level; s++) y -= g.ref[y - g.node_lo];  // ancestor at distance `level`
        int32_t* row = level == 0 ? out + out_off[q] : cur;
        const int64_t copied = decode_extras<DEF>(g, y, row);
        if (copied < 0) return;
        if (g.ref[y - g.node_lo] != 0) merge_copied<DEF>(g, y, row, parent);
        parent = row;
        cur += g.outdeg[y - g.node_lo];
    }
}

__global__ void k_gather_outdeg(GraphDev g, const int32_t* __restrict__ xs, int32_t from, int64_t nx, int32_t* __restrict__ d) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nx) return;
    const int32_t x = xs ? xs[q] : from + (int32_t)q;
    if (x < g.node_lo || x >= g.node_hi) { report(g.err, E_INVAL, x, 0); d[q] = 0; return; }
    d[q] = g.outdeg[x - g.node_lo];
}

}  // namespace bvg
