// bvg_capi.cu -- the C ABI of include/bvgraph_b200.h over the CUDA kernels.
//
// Host-side structure mirrors what the reference keeps per BVGraph (BVGraph.java:420-448: graphMemory, offsets,
// window/codec parameters), except that the bytes live in HBM and every decode entry point launches kernels.
// There is no CPU decode path: any compute call without a usable CUDA device returns BVG_ECUDA.
#include "../../../include/bvgraph_b200.h"
#include "bvg_format.hpp"
#include "bvg_kernels.cuh"
#include "bvg_long.cuh"
#include "bvg_offsets.cuh"
#include "bvg_labels.cuh"
#include "bvg_consumers.cuh"
#include "bvg_ef.cuh"
#include "bvg_compress.cuh"
#include "bvg_boundaries.cuh"
#include "bvg_tile.cuh"
#include "bvg_stream.cuh"

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace bvg;

static std::atomic<int64_t> g_launches{0};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_cuda_error(e_, #call); return BVG_ECUDA; } } while (0)
#define LAUNCH(kernel, grid, block, smem, stream, ...) do { \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); g_launches.fetch_add(1, std::memory_order_relaxed); } while (0)
// Same, bracketed by CUDA events on the launching stream when per-kernel profiling is on (bvg_profile).
#define LAUNCH_P(g, name, kernel, grid, block, smem, stream, ...) do { \
    ProfSpan* ps_ = (g)->prof_begin(name, stream); \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); g_launches.fetch_add(1, std::memory_order_relaxed); \
    (g)->prof_end(ps_, stream); } while (0)

struct ProfSpan { const char* name; cudaEvent_t e0, e1; };

static thread_local char t_cuda_msg[256];
static void set_cuda_error(cudaError_t e, const char* what) {
    snprintf(t_cuda_msg, sizeof t_cuda_msg, "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
}

// BVG_TRACE=1 prints host-side phase timings of open to stderr (each phase synchronised).
#include <chrono>
struct Trace {
    bool on;
    cudaStream_t s;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(cudaStream_t st) : on(getenv("BVG_TRACE") != nullptr), s(st), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(s);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[bvg] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ------------------------------------------------------------------------------------------------------------
// Device memory.  Graph arrays, open-time temporaries and per-scan scratch come from a size-keyed cache of cudaMalloc'd
// blocks owned by the library: an open-scan-close cycle asks for the same ~70 sizes every time, and the driver's
// stream-ordered pool (cudaMallocAsync), which served them before, occasionally takes 0.5 s to hand out a 256 MB block
// it had cached a moment earlier (measured: one open in seven on the 1 B-arc graph).  A freed block remembers an event
// on the stream it was last used on; reuse on that stream is ordered by the stream itself, reuse on another stream
// waits for the event.
// ------------------------------------------------------------------------------------------------------------
struct DevBlock { void* p; size_t bytes; int dev; cudaStream_t stream; cudaEvent_t ev; };
struct DevCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, DevBlock> idle;
    std::map<void*, DevBlock> live;
    size_t idle_bytes = 0;
    size_t cap() const {
        static const size_t c = [] { const char* v = getenv("BVG_CACHE_GB"); return (size_t)(v ? atol(v) : 32) << 30; }();
        return c;
    }
    void drop_idle(int dev) {  // caller holds mu
        for (auto it = idle.begin(); it != idle.end();) {
            if (it->second.dev == dev) { cudaEventDestroy(it->second.ev); cudaFree(it->second.p); idle_bytes -= it->second.bytes; it = idle.erase(it); }
            else ++it;
        }
        cudaGetLastError();
    }
};
static DevCache& dev_cache() { static DevCache* c = new DevCache(); return *c; }  // leaked on purpose: no CUDA calls at exit

static cudaError_t dev_alloc(void** out, size_t bytes, cudaStream_t s) {
    *out = nullptr;
    bytes = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
    {   // size classes, four per octave: requests of about the same size (the pieces of bvg_scan_memory, the shards of a
        // graph) share blocks exactly, and a cudaMalloc -- which waits for every stream, uploads in flight included --
        // happens only until each class has been seen
        int msb = 63 - __builtin_clzll((unsigned long long)bytes);
        if (msb >= 12) { const size_t step = (size_t)1 << (msb - 2); bytes = (bytes + step - 1) & ~(step - 1); }
    }
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.idle.lower_bound(std::make_pair(dev, bytes));
    if (it != c.idle.end() && it->first.first == dev && it->second.bytes == bytes) {
        DevBlock b = it->second;
        c.idle.erase(it);
        c.idle_bytes -= b.bytes;
        if (b.stream != s) { e = cudaStreamWaitEvent(s, b.ev, 0); if (e != cudaSuccess) { cudaGetLastError(); cudaEventSynchronize(b.ev); } }
        b.stream = s;
        c.live[b.p] = b;
        *out = b.p;
        return cudaSuccess;
    }
    DevBlock b{ nullptr, bytes, dev, s, nullptr };
    e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {  // give the idle blocks back and try once more
        cudaGetLastError();
        c.drop_idle(dev);
        e = cudaMalloc(&b.p, bytes);
        if (e != cudaSuccess) return e;
    }
    e = cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { cudaFree(b.p); return e; }
    c.live[b.p] = b;
    *out = b.p;
    return cudaSuccess;
}

static cudaError_t dev_free(void* p, cudaStream_t s) {
    if (!p) return cudaSuccess;
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) return cudaErrorInvalidValue;
    DevBlock b = it->second;
    c.live.erase(it);
    if (c.idle_bytes + b.bytes > c.cap() || cudaEventRecord(b.ev, s) != cudaSuccess) {
        cudaGetLastError();
        cudaEventDestroy(b.ev);
        cudaFree(b.p);  // synchronises with everything that may still use the block
        return cudaSuccess;
    }
    b.stream = s;
    c.idle.emplace(std::make_pair(b.dev, b.bytes), b);
    c.idle_bytes += b.bytes;
    return cudaSuccess;
}

// Pinned host memory of the cursors' batches comes from a cache as well: cudaHostAlloc / cudaFreeHost pin and unpin pages
// (~10 ms per 32 MB) and synchronise the whole device, which would stall every other cursor's stream.
struct PinCache {
    std::mutex mu;
    std::multimap<size_t, void*> idle;
    std::map<void*, size_t> live;
};
static PinCache& pin_cache() { static PinCache* c = new PinCache(); return *c; }
static cudaError_t pin_alloc(void** out, size_t bytes) {
    size_t cls = (size_t)1 << 20;
    while (cls < bytes) cls <<= 1;
    PinCache& c = pin_cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.idle.find(cls);
        if (it != c.idle.end()) { *out = it->second; c.idle.erase(it); c.live[*out] = cls; return cudaSuccess; }
    }
    const cudaError_t e = cudaHostAlloc(out, cls, cudaHostAllocDefault);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(c.mu);
    c.live[*out] = cls;
    return cudaSuccess;
}
static size_t pin_capacity(void* p) {
    PinCache& c = pin_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    return it == c.live.end() ? 0 : it->second;
}
static void pin_free(void* p) {
    if (!p) return;
    PinCache& c = pin_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) return;
    c.idle.emplace(it->second, p);
    c.live.erase(it);
}

static int env_int(const char* name, int dflt, int lo, int hi);
struct bvg_graph;
extern "C" { static int fetch_offsets(const bvg_graph* g, int32_t a, int32_t b, uint64_t* va, uint64_t* vb); }
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)std::max<int64_t>(1, (n + block - 1) / block); }

struct bvg_graph {
    int device = 0;
    cudaStream_t stream = nullptr;
    int offset_type = 1;
    // properties
    int32_t n_total = 0;
    int64_t m_total = 0;
    int32_t window = 0, maxref = 0, minlen = 0, zetak = 3;
    uint32_t flags = 0;
    uint64_t graph_bits_total = 0;
    Codec codec{};
    bool def_codec = true;
    // loaded window of nodes and the API extent
    int32_t node_lo = 0, node_hi = 0, ext_from = 0, ext_to = 0;
    // device state
    uint32_t* d_words = nullptr;
    uint64_t nwords = 0, bit_base = 0, bit_end = 0;
    uint64_t* d_offsets = nullptr;
    int32_t *d_outdeg = nullptr, *d_ref = nullptr, *d_depth = nullptr;
    int64_t* d_rowoff = nullptr;
    ErrWord* d_err = nullptr;
    int32_t max_depth = 0, max_outdeg = 0;
    // length-bucketed schedules (k_order_keys): extras order over all nodes with successors, merge order level-major
    int32_t *d_order_e = nullptr, *d_order_m = nullptr;
    ExtraRec* d_rec_e = nullptr;
    MergeRec* d_rec_m = nullptr;
    int64_t order_e_count = 0;
    std::vector<int64_t> level_start;  // merge schedule: nodes of chain level l+1 are order_m[level_start[l] .. level_start[l+1])
    uint8_t* d_is_parent = nullptr;    // nodes some other node copies from (k_mark_parents)
    int32_t* d_copied = nullptr;       // successors each node copies from its parent (k_order_keys)
    bool copied_ready = false;
    int32_t* d_long_nodes = nullptr;   // ids of the long records
    std::vector<int32_t> h_long_nodes; // the same on the host, ascending (does a small range hold a long record?)
    // long records split across threads (bvg_long.cuh)
    int32_t nlong = 0;
    LongMeta* d_long_meta = nullptr;
    int32_t *d_cb_cum = nullptr, *d_cb_ppos = nullptr, *d_iv_cum = nullptr, *d_iv_left = nullptr;
    uint64_t* d_seg_pos = nullptr;
    int64_t* d_seg_val = nullptr;
    // per-scan work items of the long records: running item counts per record (ItemMap, bvg_long.cuh), one array of
    // nlong + 1 entries per item family: [residual segments | row chunks | extras chunks | merge chunks of level 1, 2, ...]
    int64_t* d_long_cum = nullptr;
    int64_t n_items_resid = 0, n_items_extras = 0;
    std::vector<int64_t> n_items_merge;  // [level]
    int32_t* d_long_hint = nullptr;        // search hints of every family (ItemMap::hint), family f at hint_off[f]
    std::vector<int64_t> hint_off;
    ItemMap item_map(int family) const {
        return ItemMap{ d_long_cum + (size_t)family * ((size_t)nlong + 1), nlong,
                        d_long_hint && (size_t)family < hint_off.size() ? d_long_hint + hint_off[(size_t)family] : nullptr };
    }
    mutable std::vector<int64_t> h_long_cum;   // the same on the host (fetched on first use): which items belong to the long records of a node range
    size_t long_cum_entries = 0;
    std::vector<int64_t> fam_total;    // items of every family
    int64_t long_tmp_entries = 0;
    // what a scan reads of the long records, for the roofline arithmetic of bench.py (bvg_scan_bits)
    int64_t long_arcs = 0, long_resid_bits = 0, long_pre_bits = 0, long_index_bytes = 0;
    int64_t order_m_count = 0;
    // tunables (BVG_LONG_D / BVG_LONG_SEG / BVG_LONG_CHUNK override the defaults of bvg_long.cuh)
    int32_t long_d = LONG_D, long_seg = LONG_SEG, long_chunk = LONG_CHUNK;
    LongIndex long_index() const {
        LongIndex li;
        li.seg = long_seg; li.chunk = long_chunk;
        li.meta = d_long_meta; li.cb_cum = d_cb_cum; li.cb_ppos = d_cb_ppos; li.iv_cum = d_iv_cum; li.iv_left = d_iv_left;
        li.seg_pos = d_seg_pos; li.seg_val = d_seg_val;
        return li;
    }
    // tile plan of the scan (bvg_tile.cuh): tiles in node order, launch order (heaviest first), host copy of the first
    // consumed node of every tile (which tiles does a node range touch?)
    TileEntry* d_tiles = nullptr;
    int32_t* d_tile_order = nullptr;
    int32_t ntiles = 0;
    std::vector<int32_t> h_tile_from;
    bool tile_ok = false;
    int tile_nt = 256;
    uint32_t tile_smem = 0;
    int64_t long_scan_entries = 0;     // scratch of the long records somebody copies from (3 d entries each)
    // stream-position entries of the extras kernel (bvg_stream.cuh): one per STREAM_CHUNK_BITS bits of the loaded stream
    StreamEntry* d_stream_entries = nullptr;
    int64_t stream_chunks = 0;
    uint64_t stream_bit0 = 0;
    mutable std::map<int32_t, uint64_t> offset_seen;   // host-side memo of record positions (guarded by mu)
    // Threading (ImmutableGraph.java:157-165: an instance is not thread-safe, copy() is): the entry points that take a graph
    // lock call_mu, so concurrent callers are serialised instead of racing on the stream and the error word; cursors carry
    // their own stream and error word and run beside each other and beside everything else.
    mutable std::recursive_mutex call_mu;
    // where the records of each 2^ORDER_CHUNK_LOG-node chunk sit in the schedules (slot 0: the heavy records of all chunks), so
    // that a range decode walks only its own chunks: e_slot[s] .. e_slot[s + 1], m_slot[level - 1][s] .. [s + 1]
    std::vector<int64_t> e_slot;
    std::vector<std::vector<int64_t>> m_slot;
    mutable std::mutex sched_mu;       // the length-bucketed schedules of the range-decode kernels are built on first use
    bool schedules_ready = false;
    // halo imported from the previous shard (bvg_halo_import)
    int32_t* d_halo_lists = nullptr;
    int64_t* d_halo_off = nullptr;
    int32_t halo_count = 0;
    int64_t halo_off_cap = 0, halo_lists_cap = 0;
    mutable int32_t halo_first = 0;   // bvg_halo_needed, remembered
    mutable bool halo_first_known = false;
    int32_t halo_import_count = -1;   // shape of the last import (bvg_halo_import)
    int64_t halo_import_total = -1;
    // per-kernel timing (bench.py's roofline object): spans recorded while prof_on
    mutable bool prof_on = false;
    mutable std::vector<ProfSpan*> prof_spans;
    ProfSpan* prof_begin(const char* name, cudaStream_t s) const {
        if (!prof_on) return nullptr;
        ProfSpan* p = new ProfSpan{ name, nullptr, nullptr };
        cudaEventCreate(&p->e0); cudaEventCreate(&p->e1);
        cudaEventRecord(p->e0, s);
        return p;
    }
    void prof_end(ProfSpan* p, cudaStream_t s) const {
        if (!p) return;
        cudaEventRecord(p->e1, s);
        std::lock_guard<std::mutex> lk(mu);
        prof_spans.push_back(p);
    }
    // second stream + fork/join events: a consume-only scan runs the long-record kernels beside the short-record ones
    mutable cudaStream_t aux = nullptr;
    mutable cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    mutable std::mutex scan_mu;
    bool aux_ready() const {
        if (aux) return true;
        if (cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); aux = nullptr; return false; }
        return true;
    }
    // last device error (BVGraph.java:1129-1131 logs node + position)
    mutable std::mutex mu;
    mutable int32_t err_node = -1;
    mutable int64_t err_bitpos = -1;
    // host-side memo of immutable index values (guarded by mu): row offsets fetched so far, first chain root before `from`
    mutable std::map<int32_t, int64_t> rowoff_seen;
    mutable std::map<std::pair<int32_t, int32_t>, int32_t> halo_start_seen;

    GraphDev dev() const {
        GraphDev g;
        g.words = d_words; g.nwords = nwords; g.bit_base = bit_base; g.bit_end = bit_end;
        g.offsets = d_offsets; g.node_lo = node_lo; g.node_hi = node_hi; g.c = codec;
        g.outdeg = d_outdeg; g.ref = d_ref; g.depth = d_depth; g.rowoff = d_rowoff; g.err = d_err;
        g.copied = copied_ready ? d_copied : nullptr;
        g.hist = nullptr; g.hist_len = 0;
        return g;
    }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; cudaGetLastError(); return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) { ok = false; cudaGetLastError(); }
    }
    ~DeviceGuard() { if (ok && prev >= 0) cudaSetDevice(prev); }
};

// Small host-to-device transfers inside an index build (histogram offsets, long-record tables).  A cudaMemcpyAsync would
// queue on the host-to-device copy engine, which serves all streams in issue order: in bvg_scan_memory it would wait
// behind the next piece's 250 MB upload (measured: the index build of a piece takes 9 ms instead of 3.3).  The bytes are
// staged in pinned memory and pulled by a kernel on the graph's stream.
struct PinBlock { void* p; size_t bytes; cudaEvent_t ev; };
static int small_h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return BVG_OK;
    static std::mutex mu;
    static std::vector<PinBlock>* pool = new std::vector<PinBlock>();
    const size_t padded = (bytes + 3) & ~(size_t)3;
    PinBlock blk{ nullptr, 0, nullptr };
    {
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < pool->size(); i++)
            if ((*pool)[i].bytes >= padded && cudaEventQuery((*pool)[i].ev) == cudaSuccess) { blk = (*pool)[i]; pool->erase(pool->begin() + (long)i); break; }
        cudaGetLastError();
    }
    if (!blk.p) {
        blk.bytes = std::max<size_t>(((padded + ((size_t)1 << 20) - 1) >> 20) << 20, (size_t)1 << 20);
        CK(cudaHostAlloc(&blk.p, blk.bytes, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&blk.ev, cudaEventDisableTiming));
    }
    memcpy(blk.p, src, bytes);
    const uint64_t nwords = padded / 4;
    LAUNCH(k_pull_copy, (unsigned)std::min<uint64_t>(64, (nwords + 255) / 256), 256, 0, s, (const uint32_t*)blk.p, (uint32_t*)dst, nwords);
    CK(cudaGetLastError());
    CK(cudaEventRecord(blk.ev, s));
    std::lock_guard<std::mutex> lk(mu);
    pool->push_back(blk);
    return BVG_OK;
}

// Stream-ordered temporary.
template <class T>
struct Tmp {
    T* p = nullptr;
    cudaStream_t s;
    explicit Tmp(cudaStream_t st) : s(st) {}
    cudaError_t alloc(size_t count) { return dev_alloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T), s); }
    ~Tmp() { if (p) dev_free(p, s); }
};

static int set_codec(bvg_graph* g) {
    Codec& c = g->codec;
    c.outdeg = C_GAMMA; c.block = C_GAMMA; c.resid = C_ZETA; c.ref = C_UNARY; c.bcount = C_GAMMA;  // BVGraph.java:525-541
    const uint32_t f = g->flags;
    if (f & 0xF) c.outdeg = f & 0xF;
    if ((f >> 4) & 0xF) c.block = (f >> 4) & 0xF;
    if ((f >> 8) & 0xF) c.resid = (f >> 8) & 0xF;
    if ((f >> 12) & 0xF) c.ref = (f >> 12) & 0xF;
    if ((f >> 16) & 0xF) c.bcount = (f >> 16) & 0xF;
    c.zetak = g->zetak; c.window = g->window; c.minlen = g->minlen;
    auto gd = [](int x) { return x == C_GAMMA || x == C_DELTA; };
    auto gdu = [](int x) { return x == C_GAMMA || x == C_DELTA || x == C_UNARY; };
    // Exactly the codings the reference's readers accept (BVGraph.java:631-637, 658-664, 696-707, 732-739, 762-769,
    // 791-816): gamma | delta outdegrees; unary | gamma | delta blocks, block counts and references; gamma | zeta | delta |
    // Golomb (modulus zetaK) | nibble residuals.  Anything else (skewed Golomb anywhere) is the reference's
    // UnsupportedOperationException.  Golomb and nibble are covered by none of the reference's tests or fixtures
    // ("parity unpinned"): they follow the published dsiutils definitions.
    if (!gd(c.outdeg) || !gdu(c.block) || !gdu(c.ref) || !gdu(c.bcount) ||
        !(c.resid == C_GAMMA || c.resid == C_DELTA || c.resid == C_ZETA || c.resid == C_GOLOMB || c.resid == C_NIBBLE)) return BVG_EUNSUPPORTED;
    if (c.resid == C_GOLOMB && c.zetak < 0) return BVG_EINVAL;  // readGolomb: IllegalArgumentException on a negative modulus
    g->def_codec = c.outdeg == C_GAMMA && c.block == C_GAMMA && c.resid == C_ZETA && c.ref == C_UNARY && c.bcount == C_GAMMA;
    return BVG_OK;
}

// Where a call runs and where its kernels report: the graph's own stream and error word, or a cursor's.
struct Exec { cudaStream_t s; ErrWord* err; };
static inline Exec exec_of(const bvg_graph* g) { return Exec{ g->stream, g->d_err }; }

// Pulls the device error word; returns its code (0 if none) and remembers node/bitpos.
static int fetch_error(const bvg_graph* g, Exec ex) {
    ErrWord e{};
    if (cudaMemcpyAsync(&e, ex.err, sizeof e, cudaMemcpyDeviceToHost, ex.s) != cudaSuccess ||
        cudaStreamSynchronize(ex.s) != cudaSuccess) { cudaGetLastError(); return BVG_ECUDA; }
    if (e.code) {
        std::lock_guard<std::mutex> lk(g->mu);
        g->err_node = e.node; g->err_bitpos = e.bitpos;
        cudaMemsetAsync(ex.err, 0, sizeof(ErrWord), ex.s);
    }
    return e.code;
}
static int fetch_error(const bvg_graph* g) { return fetch_error(g, exec_of(g)); }

static int device_exclusive_scan(cudaStream_t s, const int32_t* d_in, int64_t n, int64_t* d_out /* n+1 */) {
    const int64_t nblocks = std::max<int64_t>(1, (n + SCAN_TILE - 1) / SCAN_TILE);
    Tmp<int64_t> sums(s);
    CK(sums.alloc((size_t)nblocks));
    LAUNCH(k_scan_sums, (unsigned)nblocks, SCAN_THREADS, 0, s, d_in, n, sums.p);
    LAUNCH(k_scan_blocks, 1, SCAN_THREADS, 0, s, sums.p, nblocks);
    LAUNCH(k_scan_apply, (unsigned)nblocks, SCAN_THREADS, 0, s, d_in, n, sums.p, d_out);
    CK(cudaGetLastError());
    return BVG_OK;
}

// The .offsets stream decoded on the device (bvg_offsets.cuh): upload, byte-swap, speculate, fix until stable, scan, emit.
// *d_full receives n+1 absolute bit offsets (device memory, cudaMalloc).
static int device_decode_offsets(cudaStream_t s, const uint8_t* stream, uint64_t nbytes, int coding, int64_t n, uint64_t** d_full) {
    if (coding != C_GAMMA && coding != C_DELTA) return BVG_EUNSUPPORTED;  // readOffset, BVGraph.java:631-637
    const uint64_t nwords = ((nbytes + 3) / 4 + 8 + 3) & ~(uint64_t)3;
    const uint64_t total_bits = nbytes * 8;
    const int64_t nsub = std::max<int64_t>(1, (int64_t)((total_bits + OFF_SUB_BITS - 1) / OFF_SUB_BITS));
    Trace tr(s);
    Tmp<uint32_t> words(s);
    Tmp<OffSub> sa(s), sb(s);
    Tmp<int> changed(s);
    Tmp<int64_t> cbase(s);
    Tmp<uint64_t> sbase(s);
    CK(words.alloc((size_t)nwords));
    CK(cudaMemsetAsync(words.p, 0, (size_t)nwords * 4, s));
    if (nbytes) CK(cudaMemcpyAsync(words.p, stream, (size_t)nbytes, cudaMemcpyHostToDevice, s));
    LAUNCH(k_bswap, grid_for((int64_t)nwords, 256), 256, 0, s, words.p, nwords);
    tr.mark("  offsets: upload + swap");
    CK(sa.alloc((size_t)nsub));
    CK(sb.alloc((size_t)nsub));
    CK(changed.alloc(1));
    LAUNCH(k_off_speculate, grid_for(nsub, 128), 128, 0, s, words.p, nwords, total_bits, coding, nsub, sa.p);
    OffSub *in = sa.p, *out = sb.p;
    for (int64_t pass = 0;; pass++) {
        CK(cudaMemsetAsync(changed.p, 0, sizeof(int), s));
        LAUNCH(k_off_fix, grid_for(nsub, 128), 128, 0, s, words.p, nwords, total_bits, coding, nsub, in, out, changed.p);
        int ch = 0;
        CK(cudaMemcpyAsync(&ch, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        std::swap(in, out);
        if (!ch) break;
        if (pass > nsub + 2) return BVG_EIO;
    }
    tr.mark("  offsets: speculate + fix");
    Tmp<unsigned long long> totals(s);
    CK(totals.alloc(2));
    CK(cbase.alloc((size_t)nsub));
    CK(sbase.alloc((size_t)nsub));
    LAUNCH(k_off_prefix, 1, SCAN_THREADS, 0, s, in, nsub, cbase.p, sbase.p, totals.p);
    unsigned long long ht[2] = { 0, 0 };
    CK(cudaMemcpyAsync(ht, totals.p, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if ((int64_t)ht[0] < n + 1) return BVG_EIO;  // the stream ends before n+1 gaps (EOFException in the reference)
    tr.mark("  offsets: prefix");
    CK(dev_alloc((void**)d_full, ((size_t)n + 1) * 8, s));
    LAUNCH(k_off_emit, grid_for(nsub, 128), 128, 0, s, words.p, nwords, total_bits, coding, nsub, in, cbase.p, sbase.p, n, *d_full);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));  // the temporaries die with this scope
    tr.mark("  offsets: emit");
    return BVG_OK;
}

// Record boundaries of a graph that comes without .offsets (bvg_boundaries.cuh): the whole stream goes to the device for
// the duration of the call; *d_full receives the n + 1 bit positions .offsets would have held.
static int device_offsets_from_graph(cudaStream_t s, const uint8_t* graph, uint64_t graph_bytes, const Codec& c, bool def_codec,
                                     int64_t n, uint64_t** d_full, int32_t* err_node, int64_t* err_bitpos) {
    const uint64_t stream_bits = graph_bytes * 8;
    const uint64_t nwords = ((graph_bytes + 3) / 4 + STREAM_PAD_WORDS + 3) & ~(uint64_t)3;
    Tmp<uint32_t> words(s);
    CK(words.alloc((size_t)nwords));
    CK(cudaMemsetAsync(words.p, 0, nwords * 4, s));
    if (graph_bytes) CK(cudaMemcpyAsync(words.p, graph, graph_bytes, cudaMemcpyHostToDevice, s));
    LAUNCH(k_bswap, grid_for((int64_t)nwords, 256), 256, 0, s, words.p, nwords);
    // Sub-ranges long enough for a wrong chain to fall onto the right one well inside them (a few hundred records),
    // few enough that their window-sized histories stay small.
    uint64_t sub_bits = 1ull << 21, cap = 1ull << 23;   // cap swept on the 1 B-arc graph: 2^25 1097 ms, 2^24 840, 2^23 697, 2^22 747, 2^21 701, 2^20 and below slower
    if (const char* e = getenv("BVG_BND_SUB_BITS")) sub_bits = std::max<uint64_t>(64, strtoull(e, nullptr, 10));
    if (const char* e = getenv("BVG_BND_CAP_BITS")) cap = std::max<uint64_t>(64, strtoull(e, nullptr, 10));
    const int lanes = env_int("BVG_BND_LANES", 1, 1, 32);  // walks per warp (k_bnd_walk)
    const int lean = env_int("BVG_BND_LEAN", 0, 0, 1);      // residual runs through the 32-bit window of the scan kernels (to be measured)
    const int32_t W = c.window;
    const int64_t max_sub = std::max<int64_t>(1, std::min<int64_t>(env_int("BVG_BND_MAX_SUB", 16384, 1, 1 << 24), ((int64_t)1 << 24) / std::max<int32_t>(W, 1)));
    while ((int64_t)((stream_bits + sub_bits - 1) / sub_bits) > max_sub) sub_bits *= 2;
    const int64_t nsub = std::max<int64_t>(1, (int64_t)((stream_bits + sub_bits - 1) / sub_bits));
    const size_t hw = (size_t)nsub * (size_t)std::max<int32_t>(W, 1);
    Tmp<BndSub> sa(s), sb(s);
    Tmp<int32_t> he_a(s), he_b(s), hx_a(s), hx_b(s), ring(s), ok(s);
    Tmp<BndMemo> memo(s);
    CK(memo.alloc(BND_MEMO_SLOTS));
    CK(cudaMemsetAsync(memo.p, 0, BND_MEMO_SLOTS * sizeof(BndMemo), s));
    CK(sa.alloc((size_t)nsub)); CK(sb.alloc((size_t)nsub));
    CK(he_a.alloc(hw)); CK(he_b.alloc(hw)); CK(hx_a.alloc(hw)); CK(hx_b.alloc(hw)); CK(ring.alloc(hw)); CK(ok.alloc((size_t)nsub));
    CK(cudaMemsetAsync(he_a.p, 0, hw * 4, s)); CK(cudaMemsetAsync(hx_a.p, 0, hw * 4, s));
    LAUNCH(k_bnd_init, grid_for(nsub, 128), 128, 0, s, sa.p, nsub, sub_bits, stream_bits);
    BndSub *in = sa.p, *out = sb.p;
    int32_t *he_in = he_a.p, *he_out = he_b.p, *hx_in = hx_a.p, *hx_out = hx_b.p;
    std::vector<int32_t> h_ok((size_t)nsub);
    int64_t trusted = 0;
    for (int64_t pass = 0;; pass++) {
        // one walk per warp (see k_bnd_walk), two warps per block
        if (def_codec) LAUNCH(k_bnd_walk<true>, grid_for((nsub + lanes - 1) / lanes * 32, 64), 64, 0, s, words.p, nwords, stream_bits, c, nsub, in, out, he_in, hx_in, he_out, hx_out, ring.p, (int)std::min<int64_t>(pass, 1), trusted, cap, memo.p, lanes, lean);
        else LAUNCH(k_bnd_walk<false>, grid_for((nsub + lanes - 1) / lanes * 32, 64), 64, 0, s, words.p, nwords, stream_bits, c, nsub, in, out, he_in, hx_in, he_out, hx_out, ring.p, (int)std::min<int64_t>(pass, 1), trusted, cap, memo.p, lanes, lean);
        LAUNCH(k_bnd_check, grid_for(nsub, 128), 128, 0, s, out, nsub, he_out, hx_out, W, ok.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_ok.data(), ok.p, (size_t)nsub * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        std::swap(in, out); std::swap(he_in, he_out); std::swap(hx_in, hx_out);
        int64_t first_bad = nsub;
        for (int64_t j = 0; j < nsub; j++) if (!h_ok[(size_t)j]) { first_bad = j; break; }
        if (getenv("BVG_TRACE")) fprintf(stderr, "[bvg] boundaries: pass %lld, %lld sub-ranges of %llu bits, first unproven %lld\n",
                                         (long long)pass, (long long)nsub, (unsigned long long)sub_bits, (long long)first_bad);
        if (first_bad == nsub) break;   // every entry is the exit before it, and sub-range 0 starts at bit 0: proven
        trusted = first_bad;
        if (pass > nsub + 2) return BVG_EIO;  // cannot happen: every pass proves at least one more sub-range
    }
    std::vector<BndSub> h_sub((size_t)nsub);
    CK(cudaMemcpyAsync(h_sub.data(), in, (size_t)nsub * sizeof(BndSub), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int64_t> base((size_t)nsub);
    int64_t total = 0;
    for (int64_t j = 0; j < nsub; j++) {
        base[(size_t)j] = total;
        const BndSub& b = h_sub[(size_t)j];
        if (b.bad_pos != BND_UNKNOWN && total + b.bad_index < n) {  // a record of the graph itself does not parse
            if (err_node) *err_node = (int32_t)(total + b.bad_index);
            if (err_bitpos) *err_bitpos = (int64_t)b.bad_pos;
            return BVG_EFORMAT;
        }
        total += b.count;
    }
    if (total < n) { if (err_node) *err_node = (int32_t)total; if (err_bitpos) *err_bitpos = (int64_t)stream_bits; return BVG_EIO; }  // the stream ends early
    Tmp<int64_t> d_base(s);
    CK(d_base.alloc((size_t)nsub));
    CK(cudaMemcpyAsync(d_base.p, base.data(), (size_t)nsub * 8, cudaMemcpyHostToDevice, s));
    CK(dev_alloc((void**)d_full, ((size_t)n + 1) * 8, s));
    if (total == n) {  // no padding bit after the last record: nobody tries to parse at its end
        const uint64_t end = h_sub[(size_t)nsub - 1].exit;
        CK(cudaMemcpyAsync(*d_full + n, &end, 8, cudaMemcpyHostToDevice, s));
    }
    if (def_codec) LAUNCH(k_bnd_emit<true>, grid_for((nsub + lanes - 1) / lanes * 32, 64), 64, 0, s, words.p, nwords, stream_bits, c, nsub, in, he_in, ring.p, hx_out, d_base.p, n, *d_full, memo.p, lanes, lean);
    else LAUNCH(k_bnd_emit<false>, grid_for((nsub + lanes - 1) / lanes * 32, 64), 64, 0, s, words.p, nwords, stream_bits, c, nsub, in, he_in, ring.p, hx_out, d_base.p, n, *d_full, memo.p, lanes, lean);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));  // base and end are host memory
    return BVG_OK;
}

// Index of the long records (bvg_long.cuh): list them, walk each once for sizes, lay out the arrays, walk again to fill
// them, and list the per-scan work items.
static int build_long_index(bvg_graph* g) {
    const int64_t nn = (int64_t)g->node_hi - g->node_lo;
    cudaStream_t s = g->stream;
    g->nlong = 0;
    if (nn == 0 || g->max_outdeg <= g->long_d || g->max_depth > 64) return BVG_OK;
    const int32_t LSEG = g->long_seg, LCHUNK = g->long_chunk;
    GraphDev gd = g->dev();
    Trace tr(s);
    Tmp<int32_t> flags(s);
    struct { int32_t* p; } long_nodes{ nullptr };
    Tmp<int64_t> pos(s);
    CK(flags.alloc((size_t)nn));
    CK(pos.alloc((size_t)nn + 1));
    LAUNCH(k_long_flags, grid_for(nn, 256), 256, 0, s, gd, g->long_d, flags.p);
    int rc = device_exclusive_scan(s, flags.p, nn, pos.p);
    if (rc) return rc;
    int64_t nl = 0;
    CK(cudaMemcpyAsync(&nl, pos.p + nn, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (nl == 0) return BVG_OK;
    CK(dev_alloc((void**)&g->d_long_nodes, (size_t)nl * 4, g->stream));
    long_nodes.p = g->d_long_nodes;
    LAUNCH(k_long_compact, grid_for(nn, 256), 256, 0, s, flags.p, pos.p, nn, g->node_lo, long_nodes.p);
    CK(dev_alloc((void**)&g->d_long_meta, (size_t)nl * sizeof(LongMeta), g->stream));
    if (g->def_codec) LAUNCH(k_long_count<true>, grid_for(nl, 64), 64, 0, s, gd, long_nodes.p, (int32_t)nl, g->d_is_parent, g->d_long_meta);
    else LAUNCH(k_long_count<false>, grid_for(nl, 64), 64, 0, s, gd, long_nodes.p, (int32_t)nl, g->d_is_parent, g->d_long_meta);
    tr.mark("    long: list + count walk");
    // Array offsets and running item counts, laid out on the device (k_long_layout_*): one scan per running sum.
    const size_t stride = (size_t)nl + 1;
    const int32_t levels = g->max_depth;
    const int fam_spec = 3 + levels;  // families: 0 residual segments, 1 unused, 2 extras chunks, 3.. merge chunks per level, then sub-ranges
    g->nlong = (int32_t)nl;  // item_map() below needs it; reset on failure by the caller's destroy
    const int nvals = 7 + levels;
    Tmp<int32_t> vals(s);
    Tmp<int64_t> offs(s), totals(s);   // scans that are not item families: 0 copy blocks, 1 intervals, 2 outdegree, 3 stored outdegree
    Tmp<unsigned long long> stats(s);
    CK(vals.alloc((size_t)nvals * (size_t)nl));
    CK(offs.alloc(4 * stride));
    CK(totals.alloc(8));
    CK(stats.alloc(3));
    CK(cudaMemsetAsync(stats.p, 0, 24, s));
    CK(dev_alloc((void**)&g->d_long_cum, (size_t)(fam_spec + 1) * stride * 8, g->stream));
    CK(cudaMemsetAsync(g->d_long_cum + stride, 0, stride * 8, s));   // family 1 is empty
    LAUNCH(k_long_layout_vals, grid_for(nl, 128), 128, 0, s, g->d_long_meta, (int32_t)nl, levels, LSEG, LCHUNK, (int64_t)LSPEC_BITS, vals.p);
    auto scan_row = [&](int row, int64_t* out) { return device_exclusive_scan(s, vals.p + (size_t)row * (size_t)nl, nl, out); };
    if ((rc = scan_row(0, offs.p)) || (rc = scan_row(1, offs.p + stride)) || (rc = scan_row(3, offs.p + 2 * stride)) || (rc = scan_row(4, offs.p + 3 * stride)) ||
        (rc = scan_row(2, g->d_long_cum)) || (rc = scan_row(5, g->d_long_cum + 2 * stride)) || (rc = scan_row(6, g->d_long_cum + (size_t)fam_spec * stride))) return rc;
    for (int32_t lv = 1; lv <= levels; lv++) if ((rc = scan_row(6 + lv, g->d_long_cum + (size_t)(2 + lv) * stride))) return rc;
    LAUNCH(k_long_layout_apply, grid_for(nl, 128), 128, 0, s, g->d_long_meta, (int32_t)nl, offs.p, offs.p + stride, g->d_long_cum, offs.p + 2 * stride, offs.p + 3 * stride);
    LAUNCH(k_long_stats, 64, 256, 0, s, g->d_long_meta, (int32_t)nl, stats.p);
    // totals: last entries of the scans, and the item counts of every family; the host copies of the node list and of the item
    // counts (which items belong to a node range: long_slice) come back in the same round trip
    // (the item counts per record stay on the device until a range decode asks which items belong to a node range: long_slice)
    std::vector<int64_t> h_tot(4), fam_total((size_t)fam_spec + 1, 0);
    g->h_long_nodes.resize((size_t)nl);
    g->h_long_cum.clear();
    g->long_cum_entries = (size_t)(fam_spec + 1) * stride;
    unsigned long long h_stats[3] = { 0, 0, 0 };
    for (int t = 0; t < 4; t++) CK(cudaMemcpyAsync(&h_tot[(size_t)t], offs.p + (size_t)t * stride + nl, 8, cudaMemcpyDeviceToHost, s));
    for (int f = 0; f <= fam_spec; f++) CK(cudaMemcpyAsync(&fam_total[(size_t)f], g->d_long_cum + (size_t)f * stride + nl, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(g->h_long_nodes.data(), g->d_long_nodes, (size_t)nl * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_stats, stats.p, 24, cudaMemcpyDeviceToHost, s));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    g->fam_total = fam_total;
    if (env_int("BVG_ITEM_HINTS", 1, 0, 1)) {   // search hints for the work items of every family
        std::vector<int64_t> hoff((size_t)fam_spec + 2, 0);
        for (int f = 0; f <= fam_spec; f++) hoff[(size_t)f + 1] = hoff[(size_t)f] + (fam_total[(size_t)f] >> ITEM_HINT_SHIFT) + 2;
        CK(dev_alloc((void**)&g->d_long_hint, (size_t)hoff[(size_t)fam_spec + 1] * 4, g->stream));
        for (int f = 0; f <= fam_spec; f++) {
            const int64_t cnt = hoff[(size_t)f + 1] - hoff[(size_t)f];
            LAUNCH(k_item_hints, grid_for(cnt, 256), 256, 0, s, g->d_long_cum + (size_t)f * stride, (int32_t)nl, fam_total[(size_t)f], g->d_long_hint + hoff[(size_t)f], cnt);
        }
        hoff.pop_back();
        g->hint_off = hoff;
    }
    const int64_t cb = h_tot[0], iv = h_tot[1], seg = fam_total[0], tmp = 2 * h_tot[2], scan = 3 * h_tot[3];
    CK(dev_alloc((void**)&g->d_cb_cum, (size_t)std::max<int64_t>(cb, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_cb_ppos, (size_t)std::max<int64_t>(cb, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_iv_cum, (size_t)std::max<int64_t>(iv, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_iv_left, (size_t)std::max<int64_t>(iv, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_seg_pos, (size_t)std::max<int64_t>(seg, 1) * 8, g->stream));
    CK(dev_alloc((void**)&g->d_seg_val, (size_t)std::max<int64_t>(seg, 1) * 8, g->stream));
    // copy blocks and intervals: one short walk per record; residual sync points: speculative sub-ranges (bvg_long.cuh)
    if (g->def_codec) LAUNCH(k_long_fill<true>, grid_for(nl, 64), 64, 0, s, gd, (int32_t)nl, g->d_long_meta, g->d_cb_cum, g->d_cb_ppos, g->d_iv_cum, g->d_iv_left, (uint64_t*)nullptr, (int64_t*)nullptr);
    else LAUNCH(k_long_fill<false>, grid_for(nl, 64), 64, 0, s, gd, (int32_t)nl, g->d_long_meta, g->d_cb_cum, g->d_cb_ppos, g->d_iv_cum, g->d_iv_left, (uint64_t*)nullptr, (int64_t*)nullptr);
    const int64_t ni = fam_total[(size_t)fam_spec];
    tr.mark("    long: layout + fill walk");
    if (ni > 0) {
        const ItemMap im = g->item_map(fam_spec);
        Tmp<SpecItem> ia(s), ib(s);
        Tmp<int> changed(s);
        Tmp<int64_t> v0(s), cbase(s), sbase(s);
        CK(ia.alloc((size_t)ni));
        CK(ib.alloc((size_t)ni));
        CK(changed.alloc(1));
        CK(v0.alloc((size_t)nl));
        CK(cbase.alloc((size_t)ni));
        CK(sbase.alloc((size_t)ni));
        LAUNCH(k_lspec_init, grid_for(ni, 128), 128, 0, s, g->d_long_meta, im, ni, ia.p);
        if (g->def_codec) {
            LAUNCH(k_lspec_first<true>, grid_for(nl, 128), 128, 0, s, gd, g->d_long_meta, (int32_t)nl, v0.p);
            LAUNCH(k_lspec_speculate<true>, grid_for(ni, 128), 128, 0, s, gd, ia.p, ni);
        } else {
            LAUNCH(k_lspec_first<false>, grid_for(nl, 128), 128, 0, s, gd, g->d_long_meta, (int32_t)nl, v0.p);
            LAUNCH(k_lspec_speculate<false>, grid_for(ni, 128), 128, 0, s, gd, ia.p, ni);
        }
        tr.mark("    long: speculate");
        SpecItem *in = ia.p, *out = ib.p;
        for (int64_t pass = 0;; pass++) {
            CK(cudaMemsetAsync(changed.p, 0, sizeof(int), s));
            if (g->def_codec) LAUNCH(k_lspec_fix<true>, grid_for(ni, 128), 128, 0, s, gd, in, out, ni, changed.p);
            else LAUNCH(k_lspec_fix<false>, grid_for(ni, 128), 128, 0, s, gd, in, out, ni, changed.p);
            int ch = 0;
            CK(cudaMemcpyAsync(&ch, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            std::swap(in, out);
            if (!ch) break;
            if (pass > ni + 2) return BVG_EIO;
        }
        tr.mark("    long: fix passes");
        LAUNCH(k_lspec_scan, grid_for(nl, 64), 64, 0, s, gd, g->d_long_meta, im, in, cbase.p, sbase.p);
        if (g->def_codec) LAUNCH(k_lspec_emit<true>, grid_for(ni, 128), 128, 0, s, gd, in, ni, g->d_long_meta, cbase.p, sbase.p, v0.p, g->d_seg_pos, g->d_seg_val, LSEG);
        else LAUNCH(k_lspec_emit<false>, grid_for(ni, 128), 128, 0, s, gd, in, ni, g->d_long_meta, cbase.p, sbase.p, v0.p, g->d_seg_pos, g->d_seg_val, LSEG);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(s));
        const int e = fetch_error(g);  // a record that does not hold rc residuals (k_lspec_scan)
        if (e) return e;
        tr.mark("    long: scan + emit");
    }
    g->n_items_resid = fam_total[0];
    g->n_items_extras = fam_total[2];
    g->n_items_merge.assign((size_t)levels + 1, 0);
    for (int32_t lv = 1; lv <= levels; lv++) g->n_items_merge[(size_t)lv] = fam_total[(size_t)(2 + lv)];
    g->long_tmp_entries = tmp;
    g->long_scan_entries = scan;
    g->long_resid_bits = (int64_t)h_stats[0]; g->long_pre_bits = (int64_t)h_stats[1]; g->long_arcs = (int64_t)h_stats[2];
    int64_t hint_entries = 0;
    if (g->d_long_hint) for (int f = 0; f <= fam_spec; f++) hint_entries += (fam_total[(size_t)f] >> ITEM_HINT_SHIFT) + 2;
    g->long_index_bytes = nl * (int64_t)sizeof(LongMeta) + 8 * (cb + iv) + 16 * seg + (int64_t)g->long_cum_entries * 8 + nl * 4 + 4 * hint_entries;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));  // cum / meta are host vectors
    return BVG_OK;
}

// Counting sort of the nodes into the two length-bucketed schedules (see bvg_kernels.cuh).  Chains deeper than
// MAX_LEVEL_KEYS levels (only files written with an unbounded maxrefcount) keep the natural-order kernels.
constexpr int32_t MAX_LEVEL_KEYS = 64;
static int build_parents(bvg_graph* g) {
    const int64_t nn = (int64_t)g->node_hi - g->node_lo;
    if (nn == 0) return BVG_OK;
    cudaStream_t s = g->stream;
    CK(dev_alloc((void**)&g->d_is_parent, (size_t)nn, g->stream));
    CK(cudaMemsetAsync(g->d_is_parent, 0, (size_t)nn, s));
    LAUNCH(k_mark_parents, grid_for(nn, 256), 256, 0, s, g->dev(), g->d_is_parent);
    CK(cudaGetLastError());
    return BVG_OK;
}

static int build_schedules(bvg_graph* g) {
    const int64_t nn = (int64_t)g->node_hi - g->node_lo;
    g->level_start.clear();
    if (nn == 0) return BVG_OK;
    cudaStream_t s = g->stream;
    Trace tr(s);
    const int32_t levels = std::min<int32_t>(g->max_depth, MAX_LEVEL_KEYS);
    const int64_t nchunks = ((nn + ((int64_t)1 << ORDER_CHUNK_LOG) - 1) >> ORDER_CHUNK_LOG) + 1;  // + the slot of the heavy records, scheduled first
    const int64_t per_level = nchunks * 2 * ORDER_BUCKETS;  // chunk x (parent | not) x half-octave bucket
    const int64_t nb_e = nchunks * 4 * ORDER_BUCKETS, nb_m = (int64_t)std::max(levels, 1) * per_level;
    Tmp<int32_t> key_e(s), key_m(s), bins(s), bcs(s);
    Tmp<uint64_t> epos(s), bpos(s);
    CK(key_e.alloc((size_t)nn));
    CK(key_m.alloc((size_t)nn));
    CK(bcs.alloc((size_t)nn));
    CK(epos.alloc((size_t)nn));
    CK(bpos.alloc((size_t)nn));
    CK(bins.alloc((size_t)(nb_e + nb_m)));
    CK(cudaMemsetAsync(bins.p, 0, (size_t)(nb_e + nb_m) * 4, s));
    GraphDev gd = g->dev();
    CK(dev_alloc((void**)&g->d_copied, (size_t)nn * 4, g->stream));
    if (g->def_codec) LAUNCH(k_order_keys<true>, grid_for(nn, 256), 256, 0, s, gd, key_e.p, key_m.p, levels, g->long_d, g->d_is_parent, g->d_copied, epos.p, bpos.p, bcs.p);
    else LAUNCH(k_order_keys<false>, grid_for(nn, 256), 256, 0, s, gd, key_e.p, key_m.p, levels, g->long_d, g->d_is_parent, g->d_copied, epos.p, bpos.p, bcs.p);
    tr.mark("  sched: parents + keys");
    LAUNCH(k_key_hist, grid_for(nn, 256), 256, 0, s, key_e.p, nn, bins.p);
    LAUNCH(k_key_hist, grid_for(nn, 256), 256, 0, s, key_m.p, nn, bins.p + nb_e);
    std::vector<int32_t> h((size_t)(nb_e + nb_m));
    CK(cudaMemcpyAsync(h.data(), bins.p, h.size() * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    tr.mark("  sched: histograms");
    int64_t run = 0;
    g->e_slot.assign((size_t)nchunks + 1, 0);
    for (int64_t i = 0; i < nb_e; i++) {
        if (i % (4 * ORDER_BUCKETS) == 0) g->e_slot[(size_t)(i / (4 * ORDER_BUCKETS))] = run;
        const int32_t c = h[(size_t)i]; h[(size_t)i] = (int32_t)run; run += c;
    }
    g->e_slot[(size_t)nchunks] = run;
    g->order_e_count = run;
    run = 0;
    g->level_start.assign((size_t)levels + 1, 0);
    g->m_slot.assign((size_t)std::max(levels, 1), std::vector<int64_t>((size_t)nchunks + 1, 0));
    for (int64_t i = 0; i < nb_m; i++) {
        if (i % per_level == 0 && i / per_level <= levels) g->level_start[(size_t)(i / per_level)] = run;
        if (i % (2 * ORDER_BUCKETS) == 0) g->m_slot[(size_t)(i / per_level)][(size_t)((i % per_level) / (2 * ORDER_BUCKETS))] = run;
        const int32_t c = h[(size_t)(nb_e + i)]; h[(size_t)(nb_e + i)] = (int32_t)run; run += c;
        if ((i + 1) % per_level == 0) g->m_slot[(size_t)(i / per_level)][(size_t)nchunks] = run;
    }
    g->level_start[(size_t)levels] = run;
    g->order_m_count = run;
    { const int r1 = small_h2d(bins.p, h.data(), h.size() * 4, s); if (r1) return r1; }
    CK(dev_alloc((void**)&g->d_order_e, std::max<size_t>((size_t)g->order_e_count, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_order_m, std::max<size_t>((size_t)run, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_rec_e, std::max<size_t>((size_t)g->order_e_count, 1) * sizeof(ExtraRec), g->stream));
    CK(dev_alloc((void**)&g->d_rec_m, std::max<size_t>((size_t)run, 1) * sizeof(MergeRec), g->stream));
    LAUNCH(k_key_scatter_recs, grid_for(nn, 256), 256, 0, s, gd, key_e.p, key_m.p, nn, bins.p, bins.p + nb_e, g->d_is_parent, g->d_copied,
           epos.p, bpos.p, bcs.p, g->d_order_e, g->d_order_m, g->d_rec_e, g->d_rec_m);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    tr.mark("  sched: scatter");
    g->copied_ready = true;
    g->schedules_ready = true;
    return BVG_OK;
}

// The range-decode kernels and the pre-tile scan run over the length-bucketed schedules; a graph that is only ever scanned
// through the tile kernel never builds them (2.3 GB on the 1 B-arc benchmark graph).
static int ensure_schedules(const bvg_graph* cg) {
    bvg_graph* g = const_cast<bvg_graph*>(cg);
    std::lock_guard<std::mutex> lk(g->sched_mu);
    if (g->schedules_ready) return BVG_OK;
    return build_schedules(g);
}

// Entries of the stream-position extras kernel (bvg_stream.cuh): one thread per chunk.
static int build_stream_entries(bvg_graph* g) {
    g->stream_chunks = 0;
    const int64_t nn = (int64_t)g->node_hi - g->node_lo;
    if (nn == 0 || !g->def_codec || env_int("BVG_STREAM", 0, 0, 1) == 0) return BVG_OK;
    if (g->max_outdeg > g->long_d && g->nlong == 0) return BVG_OK;  // no long index (chains too deep): the per-record kernels stay
    cudaStream_t s = g->stream;
    uint64_t o2[2];
    CK(cudaMemcpyAsync(&o2[0], g->d_offsets, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&o2[1], g->d_offsets + nn, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    g->stream_bit0 = o2[0] - g->bit_base;
    const int64_t nchunks = (int64_t)((o2[1] - o2[0] + STREAM_CHUNK_BITS - 1) / STREAM_CHUNK_BITS);
    if (nchunks <= 0) return BVG_OK;
    CK(dev_alloc((void**)&g->d_stream_entries, ((size_t)nchunks + 1) * sizeof(StreamEntry), g->stream));
    if (g->zetak == 3) LAUNCH(k_stream_entries<3>, grid_for(nchunks + 1, 128), 128, 0, s, g->dev(), nchunks, g->stream_bit0, g->node_lo, g->d_long_nodes, g->nlong, g->long_d, g->long_index(), g->d_stream_entries);
    else LAUNCH(k_stream_entries<0>, grid_for(nchunks + 1, 128), 128, 0, s, g->dev(), nchunks, g->stream_bit0, g->node_lo, g->d_long_nodes, g->nlong, g->long_d, g->long_index(), g->d_stream_entries);
    CK(cudaGetLastError());
    g->stream_chunks = nchunks;
    return BVG_OK;
}

// Tile plan of the scan (bvg_tile.cuh): node costs, their running sum, greedy cuts per super-block (count, scan, write),
// launch order by stream bits.
static int build_tile_plan(bvg_graph* g) {
    g->tile_ok = false; g->ntiles = 0;
    const int64_t nn = (int64_t)g->node_hi - g->node_lo;
    // off by default: measured 28.8 ms per scan of the 1 B-arc benchmark graph against 5.9 ms of the per-record kernels
    // (profiles/r02_tile.md: the phases of a tile are latency-bound, ~10 of 32 warps active)
    if (nn == 0 || !g->def_codec || g->window > 255 || env_int("BVG_TILE", 0, 0, 1) == 0) return BVG_OK;
    if (g->max_outdeg > g->long_d && g->nlong == 0) return BVG_OK;  // chains too deep for the long index: general kernels
    cudaStream_t s = g->stream;
    g->tile_nt = env_int("BVG_TILE_NT", 256, 64, 1024);
    if (g->tile_nt != 512 && g->tile_nt != 256 && g->tile_nt != 64 && g->tile_nt != 32) g->tile_nt = 256;
    g->tile_smem = (uint32_t)env_int("BVG_TILE_SMEM_KB", g->tile_nt == 512 ? 112 : g->tile_nt == 256 ? 55 : g->tile_nt == 64 ? 13 : 6, 4, 226) * 1024u;
    const uint32_t budget = tile_budget(g->tile_smem, g->tile_nt);
    GraphDev gd = g->dev();
    Tmp<int32_t> cost(s), counts(s);
    Tmp<uint8_t> clean(s);
    Tmp<int64_t> cum(s), base(s);
    CK(cost.alloc((size_t)nn));
    CK(clean.alloc((size_t)nn + 1));
    CK(cum.alloc((size_t)nn + 1));
    LAUNCH(k_plan_nodes, grid_for(nn, 256), 256, 0, s, gd, g->d_is_parent, g->long_d, budget, cost.p, clean.p);
    int rc = device_exclusive_scan(s, cost.p, nn, cum.p);
    if (rc) return rc;
    const int64_t first = (int64_t)g->ext_from - g->node_lo;
    const int32_t nsb = (int32_t)((nn - first + PLAN_SB - 1) / PLAN_SB);
    if (nsb <= 0) return BVG_OK;
    CK(counts.alloc((size_t)nsb));
    CK(base.alloc((size_t)nsb + 1));
    LAUNCH(k_plan_tiles, grid_for(nsb, 64), 64, 0, s, gd, cum.p, clean.p, g->d_long_nodes, g->nlong, budget, nsb, g->ext_from, counts.p, (const int64_t*)nullptr, (TileEntry*)nullptr);
    std::vector<int32_t> h_counts((size_t)nsb);
    CK(cudaMemcpyAsync(h_counts.data(), counts.p, (size_t)nsb * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int64_t> h_base((size_t)nsb + 1, 0);
    for (int32_t i = 0; i < nsb; i++) {
        if (h_counts[(size_t)i] < 0) return BVG_OK;  // some node and its ancestors do not fit a tile: the general kernels stay
        h_base[(size_t)i + 1] = h_base[(size_t)i] + h_counts[(size_t)i];
    }
    const int64_t nt = h_base[(size_t)nsb];
    if (nt <= 0 || nt > 0x7fffffff) return BVG_OK;
    { const int r1 = small_h2d(base.p, h_base.data(), ((size_t)nsb + 1) * 8, s); if (r1) return r1; }
    CK(dev_alloc((void**)&g->d_tiles, (size_t)nt * sizeof(TileEntry), g->stream));
    CK(dev_alloc((void**)&g->d_tile_order, (size_t)nt * 4, g->stream));
    LAUNCH(k_plan_tiles, grid_for(nsb, 64), 64, 0, s, gd, cum.p, clean.p, g->d_long_nodes, g->nlong, budget, nsb, g->ext_from, counts.p, base.p, g->d_tiles);
    CK(cudaGetLastError());
    std::vector<TileEntry> h_tiles((size_t)nt);
    CK(cudaMemcpyAsync(h_tiles.data(), g->d_tiles, (size_t)nt * sizeof(TileEntry), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    g->h_tile_from.resize((size_t)nt);
    std::vector<int32_t> order((size_t)nt);
    for (int64_t i = 0; i < nt; i++) { g->h_tile_from[(size_t)i] = h_tiles[(size_t)i].from; order[(size_t)i] = (int32_t)i; }
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return h_tiles[(size_t)a].cost > h_tiles[(size_t)b].cost; });
    { const int r2 = small_h2d(g->d_tile_order, order.data(), (size_t)nt * 4, s); if (r2) return r2; }
    CK(cudaStreamSynchronize(s));
    g->ntiles = (int32_t)nt;
    g->tile_ok = true;
    return BVG_OK;
}

// Uploads the stream bytes + offsets of nodes [node_lo, node_hi] and builds the decode index.
// Upload of the stream bytes of this graph object, asynchronous on its stream (truly so from pinned host memory).
static int upload_stream(bvg_graph* g, const uint8_t* bytes, uint64_t nbytes) {
    g->nwords = ((nbytes + 3) / 4 + STREAM_PAD_WORDS + 3) & ~(uint64_t)3;  // padding words, a whole number of 128-bit groups
    CK(dev_alloc((void**)&g->d_words, g->nwords * 4, g->stream));
    CK(cudaMemsetAsync(g->d_words, 0, g->nwords * 4, g->stream));
    if (nbytes) CK(cudaMemcpyAsync(g->d_words, bytes, nbytes, cudaMemcpyHostToDevice, g->stream));
    LAUNCH(k_bswap, grid_for((int64_t)g->nwords, 256), 256, 0, g->stream, g->d_words, g->nwords);
    CK(cudaGetLastError());
    return BVG_OK;
}

static int build_index(bvg_graph* g, uint64_t* d_offsets_full, int64_t n_full, bool keep_full);

static int build_device_state(bvg_graph* g, const uint8_t* bytes, uint64_t nbytes, uint64_t* d_offsets_full, int64_t n_full) {
    Trace tr(g->stream);
    const int rc = upload_stream(g, bytes, nbytes);
    if (rc) return rc;
    tr.mark("alloc + H2D stream + bswap");
    return build_index(g, d_offsets_full, n_full, false);
}

// Offsets slice and decode index of the nodes [node_lo, node_hi] of a graph object whose stream is uploaded (or on its way).
static int build_index(bvg_graph* g, uint64_t* d_offsets_full, int64_t n_full, bool keep_full) {
    const int64_t nn = (int64_t)g->node_hi - g->node_lo;
    Trace tr(g->stream);
    if (g->node_lo == 0 && nn == n_full && !keep_full) g->d_offsets = d_offsets_full;  // whole graph: adopt the decoded array
    else {
        CK(dev_alloc((void**)&g->d_offsets, ((size_t)nn + 1) * 8, g->stream));
        CK(cudaMemcpyAsync(g->d_offsets, d_offsets_full + g->node_lo, ((size_t)nn + 1) * 8, cudaMemcpyDeviceToDevice, g->stream));
        CK(cudaStreamSynchronize(g->stream));
        if (!keep_full) dev_free(d_offsets_full, g->stream);
    }
    CK(dev_alloc((void**)&g->d_outdeg, std::max<size_t>((size_t)nn, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_ref, std::max<size_t>((size_t)nn, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_depth, std::max<size_t>((size_t)nn, 1) * 4, g->stream));
    CK(dev_alloc((void**)&g->d_rowoff, ((size_t)nn + 1) * 8, g->stream));
    CK(dev_alloc((void**)&g->d_err, sizeof(ErrWord) + 2 * sizeof(int32_t), g->stream));
    CK(cudaMemsetAsync(g->d_err, 0, sizeof(ErrWord) + 2 * sizeof(int32_t), g->stream));
    int32_t* d_max = (int32_t*)(g->d_err + 1);  // [0] max depth, [1] max outdegree
    GraphDev gd = g->dev();
    tr.mark("offsets slice + allocs");
    if (nn > 0) {
        if (g->def_codec) LAUNCH(k_header<true>, grid_for(nn, 256), 256, 0, g->stream, gd, g->d_outdeg, g->d_ref);
        else LAUNCH(k_header<false>, grid_for(nn, 256), 256, 0, g->stream, gd, g->d_outdeg, g->d_ref);
    }
    int rc = device_exclusive_scan(g->stream, g->d_outdeg, nn, g->d_rowoff);
    if (rc) return rc;
    if (nn > 0) {
        LAUNCH(k_depth, grid_for(nn, 256), 256, 0, g->stream, gd, g->d_depth, d_max, g->ext_from);
        LAUNCH(k_max_i32, 512, 256, 0, g->stream, g->d_outdeg, nn, d_max + 1);
    }
    CK(cudaGetLastError());
    int32_t mx[2] = {0, 0};
    CK(cudaMemcpyAsync(mx, d_max, sizeof mx, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    g->max_depth = mx[0];
    g->max_outdeg = mx[1];
    const int e = fetch_error(g);
    if (e) return e;
    tr.mark("header + scan + depth");
    int rs = build_parents(g);
    if (!rs) rs = build_long_index(g);
    tr.mark("parents + long index");
    if (!rs) rs = build_stream_entries(g);
    tr.mark("stream entries");
    if (!rs) rs = build_tile_plan(g);
    tr.mark("tile plan");
    if (!rs && !g->tile_ok && env_int("BVG_LAZY_SCHEDULES", 0, 0, 1) == 0) { rs = build_schedules(g); tr.mark("schedules"); }
    return rs;
}

static void destroy(bvg_graph* g) {
    if (!g) return;
    DeviceGuard dg(g->device);
    // graph memory comes from the device's stream-ordered pool (kept warm): a later open reuses it without going back
    // to the driver, which is what makes open-scan-close cycles cheap
    void* ptrs[] = { g->d_words, g->d_offsets, g->d_outdeg, g->d_ref, g->d_depth, g->d_rowoff, g->d_err, g->d_halo_lists, g->d_halo_off,
                     g->d_tiles, g->d_tile_order, g->d_stream_entries, g->d_order_e, g->d_order_m, g->d_rec_e, g->d_rec_m, g->d_is_parent, g->d_long_nodes, g->d_copied, g->d_long_meta, g->d_cb_cum, g->d_cb_ppos,
                     g->d_iv_cum, g->d_iv_left, g->d_seg_pos, g->d_seg_val, g->d_long_cum, g->d_long_hint };
    for (void* p : ptrs) if (p) dev_free(p, g->stream);
    cudaStreamSynchronize(g->stream);
    if (g->aux) { cudaStreamSynchronize(g->aux); cudaStreamDestroy(g->aux); cudaEventDestroy(g->ev_fork); cudaEventDestroy(g->ev_join); }
    for (ProfSpan* p : g->prof_spans) { cudaEventDestroy(p->e0); cudaEventDestroy(p->e1); delete p; }
    cudaGetLastError();
    delete g;
}

static int pick_device(const int* devices, int ndev, int* out) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return BVG_ECUDA; }
    int dev = 0;
    if (devices && ndev > 0) dev = devices[0];
    else if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return BVG_ECUDA; }
    if (dev < 0 || dev >= count) return BVG_EINVAL;
    *out = dev;
    return BVG_OK;
}

// Stream-ordered temporaries come from the device's default pool; keep freed blocks cached instead of returning them to
// the driver at every synchronisation (a scan re-allocates the same scratch each call).
static void keep_pool_warm(int dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
}

static int env_int(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v) return dflt;
    const long x = atol(v);
    return (int)(x < lo ? lo : (x > hi ? hi : x));
}

// Records with more than long_d successors are split across threads (bvg_long.cuh).  Shorter records are one lane's
// serial work, and a kernel cannot end before its longest record has been walked: ~0.3 us per successor, once per launch
// (the extras step and every merge level).  That floor has to stay small against the scan itself (~5.5 ps per arc), so the
// threshold follows the arcs this graph object holds: 512 for the 1 B-arc benchmark graph, 128 for one eighth of it
// (measured on 125 M arcs: 2.64 ms per scan with 1024, 1.52 ms with 128; on 1 B arcs 512 and 1024 are level at 6.5 ms,
// 256 costs 0.7 ms more).  BVG_LONG_D / BVG_LONG_SEG / BVG_LONG_CHUNK override.
// Which records take the split path (bvg_long.cuh) and how finely they are split, by the arcs of the extent.  A launch of the
// per-record kernels lasts at least as long as its longest record takes one thread, and the fewer records a launch has the
// less the longest-first schedule can hide that; the split path spreads a record over threads.  Measured on the 1 B-arc
// power-law graph (profiles/shard_profile.py, ms per scan of a shard, long_d / segment / chunk):
//   whole graph  1024/128/128 5.91   512/128/128 5.72   512/32/32 5.63   256/64/64 5.72
//   1/2          512/128/128  3.30   256/64/64   3.16   128/32/32 3.19
//   1/4          256/64/64    1.91   128/64/64   1.83   96/32/32  1.74
//   1/8          1024/128/128 2.44   128/64/64   1.15   96/32/32  1.03   64/16/16 0.99   64/24/24 0.95   48/16/16 1.01
// and again with the search hints of the work items (ItemMap::hint), which make fine items cheap:
//   whole graph  512/32/32 5.58   256/32/32 5.61   256/16/16 5.21   256/12/12 5.23   192/16/16 5.25   128/16/16 5.36
//   1/2          256/32/32 3.10   128/32/32 3.06   128/16/16 2.82   192/16/16 2.85   96/16/16 2.87
//   1/4          96/32/32  1.70   96/16/16  1.57   64/24/24  1.57   64/16/16  1.59   48/16/16 1.63
//   1/8          64/24/24  0.896  64/16/16  0.909  48/16/16  0.918  32/16/16  0.955
// one_shot: the graph object lives for a single scan (bvg_scan_memory), so the index of the long records is paid per scan and
// a finer split costs more to build than it saves.
static void choose_long_threshold(bvg_graph* g, int32_t from, int32_t to, bool one_shot = false) {
    const double arcs = (double)g->m_total * (double)(to - from) / (double)std::max<int32_t>(g->n_total, 1);
    int32_t d, part;
    if (one_shot) { d = env_int("BVG_ONESHOT_D", 512, 2, 1 << 30); part = env_int("BVG_ONESHOT_PART", 128, 1, 1 << 20); }
    else if (arcs > 7.5e8) { d = 256; part = 16; }
    else if (arcs > 3.7e8) { d = 128; part = 16; }
    else if (arcs > 1.8e8) { d = 96; part = 16; }
    else { d = 64; part = 24; }
    g->long_d = env_int("BVG_LONG_D", d, 2, 1 << 30);
    g->long_seg = env_int("BVG_LONG_SEG", part, 1, 1 << 20);
    g->long_chunk = env_int("BVG_LONG_CHUNK", part, 1, 1 << 20);
}

static int open_common(bvg_graph* g, const Properties& p, int offset_type) {
    if (offset_type < -1 || offset_type > 2) return BVG_EINVAL;  // BVGraph.java:1545
    // the memory entry points get their parameters from the caller, not from a .properties file: the same checks
    // load_properties makes (a negative window would make the shard halo run backwards, zeta_0 shifts by -1)
    if (p.nodes < 0 || p.arcs < 0 || p.window < 0 || p.minlen < 0) return BVG_EINVAL;
    {
        const int resid = ((p.flags >> 8) & 0xF) ? (int)((p.flags >> 8) & 0xF) : C_ZETA;
        if (resid == C_ZETA && p.zetak < 1) return BVG_EINVAL;
        if (resid == C_GOLOMB && p.zetak < 0) return BVG_EINVAL;
    }
    keep_pool_warm(g->device);
    g->long_d = env_int("BVG_LONG_D", LONG_D, 2, 1 << 30);
    g->long_seg = env_int("BVG_LONG_SEG", LONG_SEG, 1, 1 << 20);
    g->long_chunk = env_int("BVG_LONG_CHUNK", LONG_CHUNK, 1, 1 << 20);
    g->offset_type = offset_type;
    g->n_total = (int32_t)p.nodes; g->m_total = p.arcs; g->window = p.window; g->maxref = p.maxref;
    g->minlen = p.minlen; g->zetak = p.zetak; g->flags = p.flags;
    return set_codec(g);
}

// Nodes before `from` a shard has to hold so that every reference chain stays inside it.
static int32_t shard_halo(const bvg_graph* g, int32_t from) {
    if (g->window == 0 || g->maxref == 0) return from;
    const int64_t reach = (int64_t)g->window * (int64_t)std::min<int32_t>(g->maxref < 0 ? INT32_MAX : g->maxref, 1 << 20);
    return (int32_t)std::max<int64_t>(0, (int64_t)from - reach);
}

extern "C" {

int bvg_open_memory_shard(const uint8_t* graph, uint64_t graph_bytes, const uint8_t* offsets_stream, uint64_t offsets_bytes,
                          int32_t nodes, int64_t arcs, int32_t window, int32_t maxref, int32_t minlen, int32_t zetak,
                          uint32_t flags, int offset_type, int device, int32_t from, int32_t to, bvg_graph** out) {
    if (!out || nodes < 0 || (!graph && graph_bytes)) return BVG_EINVAL;
    if (!offsets_stream && offset_type > 0) return BVG_EINVAL;  // random access needs .offsets (BVGraph.java:1581-1609)
    if (from < 0 || to < from || to > nodes) return BVG_EINVAL;
    int dev;
    int dl[1] = { device };
    int rc = pick_device(device >= 0 ? dl : nullptr, device >= 0 ? 1 : 0, &dev);
    if (rc) return rc;
    bvg_graph* g = new (std::nothrow) bvg_graph();
    if (!g) return BVG_ENOMEM;
    g->device = dev;
    DeviceGuard dg(dev);
    Properties p;
    p.nodes = nodes; p.arcs = arcs; p.window = window; p.maxref = maxref; p.minlen = minlen; p.zetak = zetak; p.flags = flags;
    rc = open_common(g, p, offset_type);
    if (rc) { destroy(g); return rc; }
    const int oc = ((flags >> 20) & 0xF) ? (int)((flags >> 20) & 0xF) : C_GAMMA;
    Trace tr(g->stream);
    uint64_t* d_full = nullptr;
    if (offsets_stream) rc = device_decode_offsets(g->stream, offsets_stream, offsets_bytes, oc, nodes, &d_full);
    else rc = device_offsets_from_graph(g->stream, graph, graph_bytes, g->codec, g->def_codec, nodes, &d_full, nullptr, nullptr);
    tr.mark(offsets_stream ? "device: decode .offsets" : "device: record boundaries from .graph");
    if (rc) { dev_free(d_full, g->stream); destroy(g); return rc; }
    g->ext_from = from; g->ext_to = to;
    choose_long_threshold(g, from, to);
    g->node_lo = shard_halo(g, from); g->node_hi = to;
    uint64_t o3[3];  // offsets of node_lo, to, nodes
    if (cudaMemcpy(&o3[0], d_full + g->node_lo, 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&o3[1], d_full + to, 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&o3[2], d_full + nodes, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); dev_free(d_full, g->stream); destroy(g); return BVG_ECUDA; }
    if (o3[2] > graph_bytes * 8) { dev_free(d_full, g->stream); destroy(g); return BVG_EIO; }
    g->graph_bits_total = o3[2];
    const uint64_t byte_lo = (o3[0] >> 3) & ~(uint64_t)15;
    const uint64_t byte_hi = std::min<uint64_t>(graph_bytes, (o3[1] + 7) >> 3);
    g->bit_base = byte_lo * 8; g->bit_end = o3[1];
    rc = build_device_state(g, graph + byte_lo, byte_hi - byte_lo, d_full, nodes);
    if (rc) { destroy(g); return rc; }
    *out = g;
    return BVG_OK;
}

int bvg_open_memory(const uint8_t* graph, uint64_t graph_bytes, const uint8_t* offsets_stream, uint64_t offsets_bytes,
                    int32_t nodes, int64_t arcs, int32_t window, int32_t maxref, int32_t minlen, int32_t zetak,
                    uint32_t flags, int offset_type, int device, bvg_graph** out) {
    return bvg_open_memory_shard(graph, graph_bytes, offsets_stream, offsets_bytes, nodes, arcs, window, maxref, minlen, zetak,
                                 flags, offset_type, device, 0, nodes, out);
}

// One pass over a graph held in host memory, pipelined: the node range is cut into `pieces` bit-balanced pieces; while the
// device builds the index of piece p and scans it, the bytes of piece p + 1 are on their way over PCIe (two streams).
int bvg_scan_memory(const uint8_t* graph, uint64_t graph_bytes, const uint8_t* offsets_stream, uint64_t offsets_bytes,
                    int32_t nodes, int64_t arcs, int32_t window, int32_t maxref, int32_t minlen, int32_t zetak,
                    uint32_t flags, int device, int32_t from, int32_t to, int pieces, int64_t* arcs_out, uint64_t* checksum_out) {
    if (nodes < 0 || (!graph && graph_bytes) || !offsets_stream || pieces < 1) return BVG_EINVAL;
    if (from < 0 || to < from || to > nodes) return BVG_EINVAL;
    int dev;
    int dl[1] = { device };
    int rc = pick_device(device >= 0 ? dl : nullptr, device >= 0 ? 1 : 0, &dev);
    if (rc) return rc;
    DeviceGuard dg(dev);
    Properties p;
    p.nodes = nodes; p.arcs = arcs; p.window = window; p.maxref = maxref; p.minlen = minlen; p.zetak = zetak; p.flags = flags;
    const int oc = ((flags >> 20) & 0xF) ? (int)((flags >> 20) & 0xF) : C_GAMMA;
    pieces = (int)std::max<int64_t>(1, std::min<int64_t>(pieces, ((int64_t)to - from) / 4096));
    cudaStream_t st[2] = { nullptr, nullptr };
    CK(cudaStreamCreateWithFlags(&st[0], cudaStreamNonBlocking));
    if (cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); cudaStreamDestroy(st[0]); return BVG_ECUDA; }
    uint64_t* d_full = nullptr;
    bvg_graph* gp[2] = { nullptr, nullptr };
    std::vector<int32_t> bounds((size_t)pieces + 1, 0);
    int64_t tot_arcs = 0;
    uint64_t tot_cs = 0;
    auto cleanup = [&]() {
        for (bvg_graph*& g : gp) if (g) { destroy(g); g = nullptr; }
        if (d_full) { dev_free(d_full, st[0]); d_full = nullptr; }
        cudaStreamSynchronize(st[0]); cudaStreamSynchronize(st[1]);
        cudaStreamDestroy(st[0]); cudaStreamDestroy(st[1]);
        cudaGetLastError();
    };
    rc = device_decode_offsets(st[0], offsets_stream, offsets_bytes, oc, nodes, &d_full);
    if (rc) { cleanup(); return rc; }
    {   // bit-balanced cuts (as bvg_plan_shards), found on the device
        Tmp<int32_t> d_bounds(st[0]);
        if (d_bounds.alloc((size_t)pieces + 1) != cudaSuccess) { cleanup(); return BVG_ECUDA; }
        LAUNCH(k_plan_cuts, 1, 64, 0, st[0], d_full, from, to, pieces, d_bounds.p, env_int("BVG_E2E_TAPER", 0, -(1 << 20), 1 << 20));
        if (cudaMemcpyAsync(bounds.data(), d_bounds.p, ((size_t)pieces + 1) * 4, cudaMemcpyDeviceToHost, st[0]) != cudaSuccess ||
            cudaStreamSynchronize(st[0]) != cudaSuccess) { cudaGetLastError(); cleanup(); return BVG_ECUDA; }
    }
    uint64_t total_bits = 0;
    if (cudaMemcpyAsync(&total_bits, d_full + nodes, 8, cudaMemcpyDeviceToHost, st[0]) != cudaSuccess || cudaStreamSynchronize(st[0]) != cudaSuccess) { cudaGetLastError(); cleanup(); return BVG_ECUDA; }
    if (total_bits > graph_bytes * 8) { cleanup(); return BVG_EIO; }
    auto prepare = [&](int q) -> int {  // graph object of piece q, its bytes enqueued for upload on its stream
        bvg_graph* g = new (std::nothrow) bvg_graph();
        if (!g) return BVG_ENOMEM;
        gp[q & 1] = g;
        g->device = dev;
        int r = open_common(g, p, 1);
        if (r) return r;
        g->stream = st[q & 1];
        const int32_t pf = bounds[(size_t)q], pt = bounds[(size_t)q + 1];
        g->ext_from = pf; g->ext_to = pt;
        choose_long_threshold(g, from, to, true);  // the index of the long records is paid at every open, here once per piece
        g->node_lo = shard_halo(g, pf); g->node_hi = pt;
        uint64_t o2[2];
        CK(cudaMemcpyAsync(&o2[0], d_full + g->node_lo, 8, cudaMemcpyDeviceToHost, g->stream));
        CK(cudaMemcpyAsync(&o2[1], d_full + pt, 8, cudaMemcpyDeviceToHost, g->stream));
        CK(cudaStreamSynchronize(g->stream));
        g->graph_bits_total = total_bits;
        const uint64_t byte_lo = (o2[0] >> 3) & ~(uint64_t)15;
        const uint64_t byte_hi = std::min<uint64_t>(graph_bytes, (o2[1] + 7) >> 3);
        g->bit_base = byte_lo * 8; g->bit_end = o2[1];
        return upload_stream(g, graph + byte_lo, byte_hi - byte_lo);
    };
    const bool trace = getenv("BVG_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t_start = now();
    rc = prepare(0);
    for (int q = 0; q < pieces && !rc; q++) {
        const auto t0 = now();
        if (q + 1 < pieces) rc = prepare(q + 1);  // its upload overlaps with everything below
        if (rc) break;
        const auto t1 = now();
        bvg_graph* g = gp[q & 1];
        rc = build_index(g, d_full, nodes, true);
        if (rc) break;
        const auto t2 = now();
        int64_t a = 0;
        uint64_t c = 0;
        rc = bvg_scan_range(g, g->ext_from, g->ext_to, &a, &c);
        if (rc) break;
        const auto t3 = now();
        tot_arcs += a; tot_cs ^= c;
        destroy(g);
        gp[q & 1] = nullptr;
        if (trace) fprintf(stderr, "[bvg] piece %d: at %.2f ms  prepare next %.2f  index %.2f  scan %.2f  close %.2f\n", q,
                           ms(t_start, t0), ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, now()));
    }
    cleanup();
    if (rc) return rc;
    if (arcs_out) *arcs_out = tot_arcs;
    if (checksum_out) *checksum_out = tot_cs;
    return BVG_OK;
}

// Shard planning (host only).  Cuts are placed at equal shares of a cost that is piecewise linear in the .graph bits -- plain
// bits for bvg_plan_shards, measured seconds per bit of every old shard for bvg_replan_shards -- and then moved to the
// nearest node no reference crosses (no node at or after the cut copies from a node before it), so that the shards need
// no boundary lists from each other at all; only when no such node exists within SHARD_CLEAN_REACH nodes does the cut stay
// where the cost puts it (and the shard re-decodes or imports its halo).
static const int64_t SHARD_CLEAN_REACH = 4096;

struct ShardPlanner {
    Properties p;
    std::vector<uint64_t> offs;
    FILE* graph = nullptr;
    int oc = C_GAMMA, rc = C_UNARY, dc = C_GAMMA;
    ~ShardPlanner() { if (graph) fclose(graph); }
    int open(const char* basename) {
        int r = load_properties(basename, p);
        if (r) return r;
        std::vector<uint8_t> ostream;
        if (!slurp_file(std::string(basename) + ".offsets", ostream)) return BVG_EIO;
        oc = ((p.flags >> 20) & 0xF) ? (int)((p.flags >> 20) & 0xF) : C_GAMMA;
        rc = ((p.flags >> 12) & 0xF) ? (int)((p.flags >> 12) & 0xF) : C_UNARY;
        dc = (p.flags & 0xF) ? (int)(p.flags & 0xF) : C_GAMMA;
        r = decode_offsets_stream(ostream.data(), ostream.size(), oc, p.nodes, offs);
        if (r) return r;
        graph = fopen((std::string(basename) + ".graph").c_str(), "rb");
        return BVG_OK;  // without the .graph file the cuts are simply not moved
    }
    static uint64_t code(HostBits& b, int coding) { return coding == C_GAMMA ? b.gamma() : coding == C_DELTA ? b.delta() : b.unary(); }
    // reference of node y (0 = none), from the first bytes of its record (BVGraph.java:1048-1053)
    int64_t ref_of(int64_t y) {
        if (!graph || p.window <= 0) return 0;
        uint8_t buf[64 + 16] = {0};
        const uint64_t bit = offs[(size_t)y];
        if (fseeko(graph, (off_t)(bit >> 3), SEEK_SET) != 0) return -1;
        const size_t got = fread(buf, 1, 64, graph);
        HostBits b{ buf, (uint64_t)got * 8 };
        b.pos = bit & 7;
        const uint64_t d = code(b, dc);
        if (d == 0) return 0;
        const uint64_t r = code(b, rc);
        return b.pos > b.nbits ? -1 : (int64_t)r;
    }
    bool clean(int64_t c) {
        for (int64_t y = c; y < p.nodes && y < c + p.window; y++) {
            const int64_t r = ref_of(y);
            if (r < 0 || r > y - c) return false;
        }
        return true;
    }
    int32_t cut_near(uint64_t target_bit, int64_t lo_limit) {
        const int64_t k = (int64_t)(std::lower_bound(offs.begin(), offs.end(), target_bit) - offs.begin());
        const int64_t k0 = std::max<int64_t>(lo_limit, std::min<int64_t>(k, p.nodes));
        for (int64_t dlt = 0; dlt <= SHARD_CLEAN_REACH; dlt++) {
            for (int sgn = 0; sgn < 2; sgn++) {
                const int64_t c = sgn ? k0 - dlt : k0 + dlt;
                if (dlt == 0 && sgn) continue;
                if (c <= lo_limit || c >= p.nodes) continue;
                if (clean(c)) return (int32_t)c;
            }
        }
        return (int32_t)k0;
    }
};

int bvg_plan_shards(const char* basename, int nshards, int32_t* bounds) {
    if (!basename || nshards < 1 || !bounds) return BVG_EINVAL;
    ShardPlanner sp;
    int rc = sp.open(basename);
    if (rc) return rc;
    const uint64_t total = sp.offs[(size_t)sp.p.nodes];
    bounds[0] = 0;
    for (int i = 1; i < nshards; i++) bounds[i] = sp.cut_near(total / (uint64_t)nshards * (uint64_t)i, bounds[i - 1]);  // equal bits
    bounds[nshards] = (int32_t)sp.p.nodes;
    return BVG_OK;
}

int bvg_replan_shards(const char* basename, int nshards, const int32_t* old_bounds, const double* old_cost, int32_t* bounds) {
    if (!basename || nshards < 1 || !bounds || !old_bounds || !old_cost) return BVG_EINVAL;
    ShardPlanner sp;
    int rc = sp.open(basename);
    if (rc) return rc;
    if (old_bounds[0] != 0 || old_bounds[nshards] != sp.p.nodes) return BVG_EINVAL;
    double total = 0;
    for (int j = 0; j < nshards; j++) { if (old_bounds[j + 1] < old_bounds[j] || !(old_cost[j] >= 0)) return BVG_EINVAL; total += old_cost[j]; }
    bounds[0] = 0;
    int j = 0;
    double before = 0;  // cost of the old shards 0 .. j-1
    for (int i = 1; i < nshards; i++) {
        const double want = total * i / nshards;
        while (j < nshards - 1 && before + old_cost[j] < want) { before += old_cost[j]; j++; }
        const uint64_t b0 = sp.offs[(size_t)old_bounds[j]], b1 = sp.offs[(size_t)old_bounds[j + 1]];
        const double f = old_cost[j] > 0 ? std::min(1.0, std::max(0.0, (want - before) / old_cost[j])) : 0.0;
        bounds[i] = sp.cut_near(b0 + (uint64_t)((double)(b1 - b0) * f), bounds[i - 1]);
    }
    bounds[nshards] = (int32_t)sp.p.nodes;
    return BVG_OK;
}

int bvg_open(const char* basename, int offset_type, const int* devices, int ndev, bvg_graph** out) {
    if (!basename || !out) return BVG_EINVAL;
    Properties p;
    int rc = load_properties(basename, p);
    if (rc) return rc;
    int dev;
    rc = pick_device(devices, ndev, &dev);
    if (rc) return rc;
    std::vector<uint8_t> graph, offs;
    if (!slurp_file(std::string(basename) + ".graph", graph)) return BVG_EIO;
    // The GPU needs record boundaries for every node, so .offsets is read for every offset_type when it exists.  The
    // reference streams a sequential / offline graph without it (offsetType <= 0, BVGraph.java:1516-1609, 1201-1213):
    // then the boundaries are found from the .graph stream itself (bvg_boundaries.cuh).
    const bool have_offsets = slurp_file(std::string(basename) + ".offsets", offs);
    if (!have_offsets && offset_type > 0) return BVG_EIO;
    return bvg_open_memory(graph.data(), graph.size(), have_offsets ? offs.data() : nullptr, offs.size(), (int32_t)p.nodes, p.arcs, p.window, p.maxref,
                           p.minlen, p.zetak, p.flags, offset_type, dev, out);
}

int bvg_open_shard(const char* basename, int device, int32_t from, int32_t to, bvg_graph** out) {
    if (!basename || !out) return BVG_EINVAL;
    Properties p;
    int rc = load_properties(basename, p);
    if (rc) return rc;
    if (from < 0 || to < from || to > p.nodes) return BVG_EINVAL;
    int dev;
    int dl[1] = { device };
    rc = pick_device(device >= 0 ? dl : nullptr, device >= 0 ? 1 : 0, &dev);
    if (rc) return rc;
    bvg_graph* g = new (std::nothrow) bvg_graph();
    if (!g) return BVG_ENOMEM;
    g->device = dev;
    DeviceGuard dg(dev);
    rc = open_common(g, p, 1);
    if (rc) { destroy(g); return rc; }
    std::vector<uint8_t> ostream;
    if (!slurp_file(std::string(basename) + ".offsets", ostream)) { destroy(g); return BVG_EIO; }
    const int oc = ((p.flags >> 20) & 0xF) ? (int)((p.flags >> 20) & 0xF) : C_GAMMA;
    uint64_t* d_full = nullptr;
    rc = device_decode_offsets(g->stream, ostream.data(), ostream.size(), oc, p.nodes, &d_full);
    if (rc) { dev_free(d_full, g->stream); destroy(g); return rc; }
    g->ext_from = from; g->ext_to = to;
    choose_long_threshold(g, from, to);
    g->node_lo = shard_halo(g, from); g->node_hi = to;
    uint64_t o3[3];
    if (cudaMemcpy(&o3[0], d_full + g->node_lo, 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&o3[1], d_full + to, 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&o3[2], d_full + p.nodes, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); dev_free(d_full, g->stream); destroy(g); return BVG_ECUDA; }
    g->graph_bits_total = o3[2];
    const uint64_t byte_lo = (o3[0] >> 3) & ~(uint64_t)15;
    const uint64_t byte_hi = (o3[1] + 7) >> 3;
    g->bit_base = byte_lo * 8; g->bit_end = o3[1];
    std::vector<uint8_t> bytes;
    if (!slurp_file(std::string(basename) + ".graph", bytes, byte_lo, byte_hi - byte_lo) || bytes.size() != byte_hi - byte_lo) { dev_free(d_full, g->stream); destroy(g); return BVG_EIO; }
    rc = build_device_state(g, bytes.data(), bytes.size(), d_full, p.nodes);
    if (rc) { destroy(g); return rc; }
    *out = g;
    return BVG_OK;
}

void bvg_close(bvg_graph* g) { destroy(g); }

int bvg_info(const bvg_graph* g, int32_t* nodes, int64_t* arcs, int32_t* window, int32_t* maxref,
             int32_t* minlen, int32_t* zetak, uint32_t* flags, int64_t* graph_bits) {
    if (!g) return BVG_EINVAL;
    if (nodes) *nodes = g->n_total;
    if (arcs) *arcs = g->m_total;
    if (window) *window = g->window;
    if (maxref) *maxref = g->maxref;
    if (minlen) *minlen = g->minlen;
    if (zetak) *zetak = g->zetak;
    if (flags) *flags = g->flags;
    if (graph_bits) *graph_bits = (int64_t)g->graph_bits_total;
    return BVG_OK;
}

int bvg_extent(const bvg_graph* g, int32_t* from, int32_t* to, int32_t* max_chain, int32_t* max_outdegree) {
    if (!g) return BVG_EINVAL;
    if (from) *from = g->ext_from;
    if (to) *to = g->ext_to;
    if (max_chain) *max_chain = g->max_depth;
    if (max_outdegree) *max_outdegree = g->max_outdeg;
    return BVG_OK;
}

int bvg_random_access(const bvg_graph* g) { return g && g->offset_type > 0 ? 1 : 0; }

int bvg_set_stream(bvg_graph* g, void* cuda_stream) {
    if (!g) return BVG_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(g->call_mu);
    g->stream = (cudaStream_t)cuda_stream;
    return BVG_OK;
}

int bvg_device(const bvg_graph* g) { return g ? g->device : BVG_EINVAL; }

int bvg_memory_footprint(const bvg_graph* g, int64_t* stream_bytes, int64_t* offsets_bytes, int64_t* index_bytes) {
    if (!g) return BVG_EINVAL;
    const int64_t nn = (int64_t)g->node_hi - g->node_lo;
    if (stream_bytes) *stream_bytes = (int64_t)g->nwords * 4;
    if (offsets_bytes) *offsets_bytes = (nn + 1) * 8;
    if (index_bytes) {
        int64_t b = nn * 12 + (nn + 1) * 8;                                         // outdegree, reference, depth; row offsets
        if (g->d_is_parent) b += nn;
        b += g->long_index_bytes;
        if (g->schedules_ready) b += nn * 4 + g->order_e_count * (4 + (int64_t)sizeof(ExtraRec)) + g->order_m_count * (4 + (int64_t)sizeof(MergeRec));
        if (g->d_tiles) b += (int64_t)g->ntiles * ((int64_t)sizeof(TileEntry) + 4);
        if (g->d_stream_entries) b += (g->stream_chunks + 1) * (int64_t)sizeof(StreamEntry);
        *index_bytes = b;
    }
    return BVG_OK;
}

int bvg_scan_bits(const bvg_graph* g, int64_t* out) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    if (!g || !out) return BVG_EINVAL;
    DeviceGuard dg(g->device);
    uint64_t oa = 0, ob = 0;
    if (g->ext_to > g->ext_from) { const int rc = fetch_offsets(g, g->ext_from, g->ext_to, &oa, &ob); if (rc) return rc; }
    out[0] = (int64_t)(ob - oa);
    out[1] = g->nlong; out[2] = g->long_resid_bits; out[3] = g->long_pre_bits; out[4] = g->long_arcs;
    out[5] = g->schedules_ready ? 1 : 0;
    return BVG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// sequential access
// ------------------------------------------------------------------------------------------------------------

static int range_check(const bvg_graph* g, int32_t from, int32_t to) {
    if (!g) return BVG_EINVAL;
    if (from < g->ext_from || to < from || to > g->ext_to) return BVG_EINVAL;  // BVGraph.java:1165
    return BVG_OK;
}

// Row offsets of two nodes on the host.  The graph is immutable, so what has been fetched once is remembered: the calls
// of a steady-state step (range decode of the boundary lists, scan of the shard) then cost no device round trip.
static int fetch_rowoff(const bvg_graph* g, int32_t a, int32_t b, int64_t* va, int64_t* vb, cudaStream_t st = nullptr) {
    {
        std::lock_guard<std::mutex> lk(g->mu);
        auto ia = g->rowoff_seen.find(a), ib = g->rowoff_seen.find(b);
        if (ia != g->rowoff_seen.end() && ib != g->rowoff_seen.end()) { *va = ia->second; *vb = ib->second; return BVG_OK; }
    }
    if (!st) st = g->stream;
    CK(cudaMemcpyAsync(va, g->d_rowoff + (a - g->node_lo), 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(vb, g->d_rowoff + (b - g->node_lo), 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->rowoff_seen.size() > 4096) g->rowoff_seen.clear();
    g->rowoff_seen[a] = *va; g->rowoff_seen[b] = *vb;
    return BVG_OK;
}

// Record start positions of two nodes on the host (memoised like the row offsets).
static int fetch_offsets(const bvg_graph* g, int32_t a, int32_t b, uint64_t* va, uint64_t* vb) {
    {
        std::lock_guard<std::mutex> lk(g->mu);
        auto ia = g->offset_seen.find(a), ib = g->offset_seen.find(b);
        if (ia != g->offset_seen.end() && ib != g->offset_seen.end()) { *va = ia->second; *vb = ib->second; return BVG_OK; }
    }
    CK(cudaMemcpyAsync(va, g->d_offsets + (a - g->node_lo), 8, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaMemcpyAsync(vb, g->d_offsets + (b - g->node_lo), 8, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->offset_seen.size() > 4096) g->offset_seen.clear();
    g->offset_seen[a] = *va; g->offset_seen[b] = *vb;
    return BVG_OK;
}

int bvg_range_arcs(const bvg_graph* g, int32_t from, int32_t to, int64_t* arcs) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    int rc = range_check(g, from, to);
    if (rc) return rc;
    DeviceGuard dg(g->device);
    int64_t a, b;
    rc = fetch_rowoff(g, from, to, &a, &b);
    if (rc) return rc;
    *arcs = b - a;
    return BVG_OK;
}

// Resolves where the parents before `from` come from (imported lists or a re-decoded halo) and fills the RowMap.
struct HaloPlan {
    Tmp<int32_t> halo;
    Tmp<int64_t> halo_off;
    int32_t lo;
    explicit HaloPlan(cudaStream_t s) : halo(s), halo_off(s), lo(0) {}
};

static int plan_halo(const bvg_graph* g, Exec ex, int32_t from, int32_t to, int32_t* d_out, int64_t row_from, RowMap& rm, HaloPlan& hp) {
    cudaStream_t s = ex.s;
    GraphDev gd = g->dev();
    gd.err = ex.err;
    rm.out = d_out; rm.out_base = row_from; rm.from = from; rm.halo = nullptr; rm.halo_off = nullptr; rm.halo_lo = from; rm.halo_base = row_from;
    rm.mask = nullptr;
    hp.lo = from;
    if (g->max_depth > 0 && from > g->node_lo) {
        if (g->halo_count > 0 && from == g->ext_from) {  // lists imported from the previous shard
            rm.halo = g->d_halo_lists; rm.halo_off = g->d_halo_off; rm.halo_lo = from - g->halo_count;
            int64_t hb, dummy;
            { const int rc = fetch_rowoff(g, rm.halo_lo, from, &hb, &dummy, s); if (rc) return rc; }
            rm.halo_base = hb;
        } else {  // re-decode the halo, as BVGraphNodeIterator's ctor re-reads the window (BVGraph.java:1173-1183)
            const int64_t reach = std::min<int64_t>((int64_t)to - from, (int64_t)g->window * g->max_depth);
            int32_t h = from;
            bool known = false;
            {
                std::lock_guard<std::mutex> lk(g->mu);
                auto it = g->halo_start_seen.find(std::make_pair(from, (int32_t)reach));
                if (it != g->halo_start_seen.end()) { h = it->second; known = true; }
            }
            if (!known) {
                Tmp<int32_t> hs(s);
                CK(hs.alloc(1));
                LAUNCH(k_set_i32, 1, 1, 0, s, hs.p, from);
                LAUNCH(k_halo_start, grid_for(reach, 128), 128, 0, s, gd, from, (int32_t)reach, hs.p);
                CK(cudaMemcpyAsync(&h, hs.p, 4, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                std::lock_guard<std::mutex> lk(g->mu);
                if (g->halo_start_seen.size() > 4096) g->halo_start_seen.clear();
                g->halo_start_seen[std::make_pair(from, (int32_t)reach)] = h;
            }
            if (h < from) {
                int64_t ra, rb;
                int rc = fetch_rowoff(g, h, from, &ra, &rb, s);
                if (rc) return rc;
                CK(hp.halo.alloc((size_t)(rb - ra)));
                CK(hp.halo_off.alloc((size_t)(from - h) + 1));
                LAUNCH(k_rel_offsets, grid_for((int64_t)from - h + 1, 128), 128, 0, s, g->d_rowoff + (h - g->node_lo), (int64_t)from - h, hp.halo_off.p);
                rm.halo = hp.halo.p; rm.halo_off = hp.halo_off.p; rm.halo_lo = h; rm.halo_base = ra;
                hp.lo = h;
            }
        }
    }
    return BVG_OK;
}

// The work items of one family that belong to the long records of the nodes [lo, to): a contiguous run of items, because the
// long records are listed in node order.
struct LongSlice { LongIndex li; ItemMap im; int64_t item0, end; int64_t count() const { return end - item0; } };
static LongSlice long_slice(const bvg_graph* g, int family, int32_t lo, int32_t to) {
    LongSlice sl{ g->long_index(), g->item_map(family), 0, 0 };
    if (g->nlong == 0) return sl;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        if (g->h_long_cum.empty() && g->long_cum_entries) {
            g->h_long_cum.assign(g->long_cum_entries, 0);
            if (cudaMemcpy(g->h_long_cum.data(), g->d_long_cum, g->long_cum_entries * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); g->h_long_cum.clear(); }
        }
    }
    if (g->h_long_cum.empty()) {   // could not fetch: every item of the family, filtered by the kernels
        sl.end = (size_t)family < g->fam_total.size() ? g->fam_total[(size_t)family] : 0;
        return sl;
    }
    const size_t l0 = (size_t)(std::lower_bound(g->h_long_nodes.begin(), g->h_long_nodes.end(), lo) - g->h_long_nodes.begin());
    const size_t l1 = (size_t)(std::lower_bound(g->h_long_nodes.begin(), g->h_long_nodes.end(), to) - g->h_long_nodes.begin());
    const size_t base = (size_t)family * ((size_t)g->nlong + 1);
    if (!sl.im.hint) {   // without search hints: narrow the search to the slice's records (the hints hold absolute record numbers)
        sl.li.meta += l0;
        sl.im.cum += l0; sl.im.nlong = (int32_t)(l1 - l0);
    }
    sl.item0 = g->h_long_cum[base + l0];
    sl.end = g->h_long_cum[base + l1];
    return sl;
}
#define LAUNCH_LONG(g, name, kernel, sl, stream, gd, lo, to, rm, ld, lf) do { \
    if ((sl).count() > 0) LAUNCH_P(g, name, kernel, grid_for((sl).count(), 64), 64, 0, stream, gd, (sl).li, (sl).im, (sl).item0, (sl).end, lo, to, rm, ld, lf); } while (0)

// Enqueues the decode of [from, to) into d_out (device, >= arcs entries). All temporaries are stream-ordered.
// Range decode over the length-bucketed schedules: extras of every wanted node, long records split across threads, one
// merge launch per chain level.  Only the schedule slices of the chunks [lo, to) touches are walked (plus the slot of the
// heavy records), so that the cost follows the range, not the graph.
static int run_ordered_decode(const bvg_graph* g, Exec ex, int32_t lo, int32_t to, int32_t from, const RowMap& rm) {
    cudaStream_t s = ex.s;
    GraphDev gd = g->dev();
    gd.err = ex.err;
    // default codings: the lean walkers of the consume-only scan with every list stored (bvg_scan.cuh)
    static const bool lean = !(getenv("BVG_DECODE_LEAN") && atoi(getenv("BVG_DECODE_LEAN")) == 0);
    const bool use_lean = g->def_codec && lean && g->d_rec_e && g->d_rec_m;
    unsigned long long* const no_fold = nullptr;
    const size_t c0 = (size_t)(((int64_t)lo - g->node_lo) >> ORDER_CHUNK_LOG) + 1, c1 = (size_t)(((int64_t)to - 1 - g->node_lo) >> ORDER_CHUNK_LOG) + 2;  // slots [c0, c1)
    auto extras = [&](int64_t a, int64_t c) {
        if (c <= 0) return;
        if (use_lean && g->zetak == 3) LAUNCH_P(g, "k_extras_lean", (k_scan_extras_lean<3, true>), grid_for(c, 128), 128, 0, s, gd, g->d_rec_e + a, c, lo, to, from, rm, no_fold, 0, 1, 1);
        else if (use_lean) LAUNCH_P(g, "k_extras_lean", (k_scan_extras_lean<0, true>), grid_for(c, 128), 128, 0, s, gd, g->d_rec_e + a, c, lo, to, from, rm, no_fold, 0, 1, 1);
        else if (g->def_codec) LAUNCH_P(g, "k_extras_ordered", k_extras_ordered<true>, grid_for(c, 128), 128, 0, s, gd, g->d_order_e + a, c, lo, to, rm);
        else LAUNCH_P(g, "k_extras_ordered", k_extras_ordered<false>, grid_for(c, 128), 128, 0, s, gd, g->d_order_e + a, c, lo, to, rm);
    };
    extras(g->e_slot[0], g->e_slot[1] - g->e_slot[0]);            // heavy records of all chunks, filtered by [lo, to)
    extras(g->e_slot[c0], g->e_slot[c1] - g->e_slot[c0]);
    Tmp<int32_t> long_tmp(s);
    LongDst ld{ nullptr };
    const LongFold lf{ nullptr, 0 };  // range decode: every long record is materialised
    const LongSlice sr = long_slice(g, 0, lo, to), se = long_slice(g, 2, lo, to);
    const bool any_long = g->nlong && (sr.count() > 0 || se.count() > 0 || long_slice(g, 2 + std::max(1, g->max_depth), lo, to).im.nlong > 0);
    if (any_long) {
        CK(long_tmp.alloc((size_t)g->long_tmp_entries));
        ld.tmp = long_tmp.p;
        if (g->def_codec) LAUNCH_LONG(g, "k_long_resid", (k_long_resid<true, RowMap>), sr, s, gd, lo, to, rm, ld, lf);
        else LAUNCH_LONG(g, "k_long_resid", (k_long_resid<false, RowMap>), sr, s, gd, lo, to, rm, ld, lf);
        LAUNCH_LONG(g, "k_long_extras", k_long_extras<RowMap>, se, s, gd, lo, to, rm, ld, lf);
    }
    for (int32_t level = 1; level <= g->max_depth; level++) {
        const std::vector<int64_t>& ms = g->m_slot[(size_t)level - 1];
        auto merge = [&](int64_t a, int64_t c) {
            if (c <= 0) return;
            if (use_lean) LAUNCH_P(g, "k_merge_lean", (k_scan_merge_lean<8, 4>), grid_for(c, 128), 128, 0, s, gd, g->d_rec_m + a, c, lo, to, from, rm, no_fold, 1);
            else if (g->def_codec) LAUNCH_P(g, "k_merge_ordered", k_merge_ordered<true>, grid_for(c, 128), 128, 0, s, gd, g->d_order_m + a, c, lo, to, rm);
            else LAUNCH_P(g, "k_merge_ordered", k_merge_ordered<false>, grid_for(c, 128), 128, 0, s, gd, g->d_order_m + a, c, lo, to, rm);
        };
        merge(ms[0], ms[1] - ms[0]);
        merge(ms[c0], ms[c1] - ms[c0]);
        if (any_long) {
            const LongSlice sm = long_slice(g, 2 + level, lo, to);
            LAUNCH_LONG(g, "k_long_merge", k_long_merge<RowMap>, sm, s, gd, lo, to, rm, ld, lf);
        }
    }
    CK(cudaGetLastError());
    return BVG_OK;
}

// Ranges of at least this many nodes are decoded over the schedules (built on first use), shorter ones in natural node order.
static const int64_t ORDERED_MIN_NODES = 32768;

static int enqueue_decode(const bvg_graph* g, Exec ex, int32_t from, int32_t to, int32_t* d_out, int64_t row_from) {
    if (to == from) return BVG_OK;
    cudaStream_t s = ex.s;
    GraphDev gd = g->dev();
    gd.err = ex.err;
    RowMap rm;
    HaloPlan hp(s);
    { const int rc = plan_halo(g, ex, from, to, d_out, row_from, rm, hp); if (rc) return rc; }
    const int32_t lo = hp.lo;
    const int64_t cnt = (int64_t)to - lo;
    // big ranges run over the length-bucketed schedules; small ones (halos, boundary lists) in natural node order
    const bool ordered = g->max_depth <= MAX_LEVEL_KEYS && (cnt >= ORDERED_MIN_NODES || cnt * 4 >= (int64_t)g->node_hi - g->node_lo);
    if (ordered) {
        const int rc = ensure_schedules(g);
        if (rc) return rc;
        return run_ordered_decode(g, ex, lo, to, from, rm);
    }
    // Small ranges: natural node order, one thread per record -- except the long records, which are split across threads
    // exactly as in a whole-graph decode.
    const LongSlice sr = long_slice(g, 0, lo, to), se = long_slice(g, 2, lo, to);
    const bool split = g->nlong > 0 &&
        std::lower_bound(g->h_long_nodes.begin(), g->h_long_nodes.end(), lo) != std::lower_bound(g->h_long_nodes.begin(), g->h_long_nodes.end(), to);
    const int32_t skip_above = split ? g->long_d : INT32_MAX;
    if (g->def_codec) LAUNCH_P(g, "k_extras", k_extras<true>, grid_for(cnt, 128), 128, 0, s, gd, lo, to, rm, skip_above);
    else LAUNCH_P(g, "k_extras", k_extras<false>, grid_for(cnt, 128), 128, 0, s, gd, lo, to, rm, skip_above);
    Tmp<int32_t> long_tmp(s);
    LongDst ld{ nullptr };
    const LongFold lf{ nullptr, 0 };
    if (split) {
        CK(long_tmp.alloc((size_t)g->long_tmp_entries));
        ld.tmp = long_tmp.p;
        if (g->def_codec) LAUNCH_LONG(g, "k_long_resid", (k_long_resid<true, RowMap>), sr, s, gd, lo, to, rm, ld, lf);
        else LAUNCH_LONG(g, "k_long_resid", (k_long_resid<false, RowMap>), sr, s, gd, lo, to, rm, ld, lf);
        LAUNCH_LONG(g, "k_long_extras", k_long_extras<RowMap>, se, s, gd, lo, to, rm, ld, lf);
    }
    for (int32_t level = 1; level <= g->max_depth; level++) {
        if (g->def_codec) LAUNCH_P(g, "k_merge", k_merge<true>, grid_for(cnt, 128), 128, 0, s, gd, lo, to, level, rm, skip_above);
        else LAUNCH_P(g, "k_merge", k_merge<false>, grid_for(cnt, 128), 128, 0, s, gd, lo, to, level, rm, skip_above);
        if (split) {
            const LongSlice sm = long_slice(g, 2 + level, lo, to);
            LAUNCH_LONG(g, "k_long_merge", k_long_merge<RowMap>, sm, s, gd, lo, to, rm, ld, lf);
        }
    }
    CK(cudaGetLastError());
    return BVG_OK;
}

int bvg_decode_range(const bvg_graph* g, int32_t from, int32_t to, int64_t* out_off, int32_t* out, int64_t cap, int on_device) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    int rc = range_check(g, from, to);
    if (rc) return rc;
    if (!out_off) return BVG_EINVAL;
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    int64_t ra, rb;
    rc = fetch_rowoff(g, from, to, &ra, &rb);
    if (rc) return rc;
    const int64_t arcs = rb - ra, cnt = (int64_t)to - from;
    if (out && cap < arcs) return BVG_ENOMEM;
    if (on_device) {
        LAUNCH(k_rel_offsets, grid_for(cnt + 1, 256), 256, 0, s, g->d_rowoff + (from - g->node_lo), cnt, out_off);
        if (out) { rc = enqueue_decode(g, exec_of(g), from, to, out, ra); if (rc) return rc; }
        CK(cudaGetLastError());
        return BVG_OK;
    }
    Tmp<int64_t> d_off(s);
    Tmp<int32_t> d_out(s);
    CK(d_off.alloc((size_t)cnt + 1));
    LAUNCH(k_rel_offsets, grid_for(cnt + 1, 256), 256, 0, s, g->d_rowoff + (from - g->node_lo), cnt, d_off.p);
    CK(cudaMemcpyAsync(out_off, d_off.p, ((size_t)cnt + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (out) {
        CK(d_out.alloc((size_t)arcs));
        rc = enqueue_decode(g, exec_of(g), from, to, d_out.p, ra);
        if (rc) return rc;
        if (arcs) CK(cudaMemcpyAsync(out, d_out.p, (size_t)arcs * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    const int e = fetch_error(g);
    return e ? e : BVG_OK;
}

// Fused consume-only scan over the length-bucketed schedules: parents' rows go to `rows` (addressed like the CSR of
// [from, to)), everything else is folded in registers; long records are materialised by the split path and folded by
// k_checksum_nodes.
static int enqueue_scan_fused(const bvg_graph* g, int32_t from, int32_t to, int32_t* rows, int64_t row_from, unsigned long long* d_result,
                              uint32_t* hist = nullptr, int64_t hist_len = 0) {
    cudaStream_t s = g->stream;
    GraphDev gd = g->dev();
    gd.hist = hist; gd.hist_len = hist_len;
    RowMap rm;
    HaloPlan hp(s);
    { const int rc = plan_halo(g, exec_of(g), from, to, rows, row_from, rm, hp); if (rc) return rc; }
    const int32_t lo = hp.lo;
    // One item per thread and as many blocks as that takes: the schedules are longest-first, so the hardware block
    // scheduler balances the SMs by itself (BVG_SCAN_PERSISTENT=1 keeps one resident wave looping instead).
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->device);
    static const bool persistent = getenv("BVG_SCAN_PERSISTENT") && atoi(getenv("BVG_SCAN_PERSISTENT")) != 0;
    const unsigned wave = (unsigned)(sms * SCAN_BLOCKS_PER_SM);
    // Extras step: BVG_SCAN_ITEMS records per thread, neighbours in the schedule (same chunk, similar length): 2.85 / 2.72 /
    // 2.82 / 3.07 ms for 1 / 2 / 4 / 8.  (With a thread's records a grid apart: 2.88 / 2.80 / 2.78 / 2.77 ms, but 4.8 instead
    // of 3.5 GB of DRAM reads.)  The merge levels are fastest with one record per thread (1.90 / 1.94 / 1.96 / 2.20 ms).
    static const int items = env_int("BVG_SCAN_ITEMS", 2, 1, 64);
    const unsigned grid = persistent ? wave : (unsigned)std::max<int64_t>(1, (g->order_e_count + (int64_t)SCAN_BLOCK * items - 1) / ((int64_t)SCAN_BLOCK * items));
    const unsigned grid_m = persistent ? wave : 0x7fffffffu;
    Tmp<int32_t> long_tmp(s);
    LongDst ld{ nullptr };
    const LongIndex li = g->long_index();
    const LongFold lf{ d_result, from };  // long records nobody copies from are folded where their parts are produced
    // The long-record kernels run on a second stream beside the short-record kernels of the same chain level; the two streams
    // meet before every level, because a short record may copy from a long one and the other way round.
    // Round 2: on by default (5.95 -> 5.67 ms at 1 GPU, 3.41 -> 3.09 ms per rank at 2 GPUs: the long-record kernels are chains of
    // dependent loads -- merge-path searches over prefix sums -- whose ~0.1 ms floors no longer add up with the short kernels');
    // off while per-kernel profiling is on, so that the kernel times of bvg_profile stay disjoint.
    static const bool overlap_default = env_int("BVG_SCAN_OVERLAP", 1, 0, 1) != 0;
    const bool overlap = overlap_default && !g->prof_on;
    std::unique_lock<std::mutex> scan_lock(g->scan_mu, std::defer_lock);
    cudaStream_t sa = s;
    if (g->nlong && overlap) { scan_lock.lock(); if (g->aux_ready()) sa = g->aux; else scan_lock.unlock(); }
    auto meet = [&]() -> cudaError_t {  // everything enqueued so far on either stream precedes everything enqueued after
        if (sa == s) return cudaSuccess;
        cudaError_t e = cudaEventRecord(g->ev_join, sa);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, g->ev_join, 0);
        if (e == cudaSuccess) e = cudaEventRecord(g->ev_fork, s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(sa, g->ev_fork, 0);
        return e;
    };
    if (g->nlong) {
        CK(long_tmp.alloc((size_t)g->long_tmp_entries));
        ld.tmp = long_tmp.p;
        if (sa != s) { CK(cudaEventRecord(g->ev_fork, s)); CK(cudaStreamWaitEvent(sa, g->ev_fork, 0)); }
    }
    static const bool lean = !(getenv("BVG_SCAN_LEAN") && atoi(getenv("BVG_SCAN_LEAN")) == 0);
    static const int dbg_nostore = env_int("BVG_DEBUG_NOSTORE", 0, 0, 1);  // timing experiments only: no row stores, results are wrong
    static const bool ring = env_int("BVG_SCAN_RING", 1, 0, 1) != 0;  // stream staged in shared memory by cp.async (bvg_scan.cuh, WinRing)
    const bool stream_extras = g->d_stream_entries != nullptr && g->stream_chunks > 0 && hist == nullptr;
    if (stream_extras) {
        // every record's extras and every residual run, long records included, by stream position (bvg_stream.cuh)
        uint64_t oa, ob;
        { const int rc = fetch_offsets(g, lo, to, &oa, &ob); if (rc) return rc; }
        StreamArgs a{};
        a.entries = g->d_stream_entries; a.nchunks = g->stream_chunks; a.bit0 = g->stream_bit0;
        a.first_chunk = (int64_t)((oa - g->bit_base - g->stream_bit0) / STREAM_CHUNK_BITS);
        const int64_t c1 = (int64_t)((ob - g->bit_base - g->stream_bit0 + STREAM_CHUNK_BITS - 1) / STREAM_CHUNK_BITS);
        a.count = std::max<int64_t>(0, std::min<int64_t>(c1, g->stream_chunks) - a.first_chunk);
        a.lo = lo; a.hi = to; a.from = from;
        a.is_parent = g->d_is_parent; a.long_nodes = g->d_long_nodes; a.nlong = g->nlong; a.long_d = g->long_d;
        a.li = li; a.long_tmp = ld.tmp; a.result = d_result; a.debug_nostore = dbg_nostore;
        if (a.count > 0) {
            Tmp<int32_t> defer(s);
            CK(defer.alloc((size_t)a.count + 2));   // a lane defers at most one record; [count] is the counter
            a.defer_list = defer.p; a.defer_count = (unsigned int*)(defer.p + a.count);
            CK(cudaMemsetAsync(a.defer_count, 0, 4, s));
            const int64_t lanes = (a.count + 32 * BVG_STREAM_GROUP - 1) / (32 * BVG_STREAM_GROUP) * 32;   // a warp per 32 * GROUP chunks
            if (g->zetak == 3) LAUNCH_P(g, "k_stream_extras", (k_stream_extras<3, RowMap>), grid_for(lanes, STREAM_BLOCK), STREAM_BLOCK, 0, s, gd, a, rm);
            else LAUNCH_P(g, "k_stream_extras", (k_stream_extras<0, RowMap>), grid_for(lanes, STREAM_BLOCK), STREAM_BLOCK, 0, s, gd, a, rm);
            if (g->minlen != 0) {
                if (g->zetak == 3) LAUNCH_P(g, "k_stream_ivfix", (k_stream_ivfix<3, RowMap>), sms * 2, 128, 0, s, gd, a, rm);
                else LAUNCH_P(g, "k_stream_ivfix", (k_stream_ivfix<0, RowMap>), sms * 2, 128, 0, s, gd, a, rm);
            }
        }
    }
    // the fused in-degree count (bvg_indegrees) is a compile-time variant of the lean kernels (default codings, enqueue_scan sees to that)
    else if (hist && g->zetak == 3) LAUNCH_P(g, "k_scan_extras", (k_scan_extras_lean<3, true, true>), grid, 128, 0, s, gd, g->d_rec_e, g->order_e_count, lo, to, from, rm, d_result, 0, 0, items);
    else if (hist) LAUNCH_P(g, "k_scan_extras", (k_scan_extras_lean<0, true, true>), grid, 128, 0, s, gd, g->d_rec_e, g->order_e_count, lo, to, from, rm, d_result, 0, 0, items);
    else if (g->def_codec && lean && g->zetak == 3 && ring) LAUNCH_P(g, "k_scan_extras", (k_scan_extras_lean<3, true>), grid, 128, 0, s, gd, g->d_rec_e, g->order_e_count, lo, to, from, rm, d_result, dbg_nostore, 0, items);
    else if (g->def_codec && lean && g->zetak == 3) LAUNCH_P(g, "k_scan_extras", (k_scan_extras_lean<3, false>), grid, 128, 0, s, gd, g->d_rec_e, g->order_e_count, lo, to, from, rm, d_result, dbg_nostore, 0, items);
    else if (g->def_codec && lean && ring) LAUNCH_P(g, "k_scan_extras", (k_scan_extras_lean<0, true>), grid, 128, 0, s, gd, g->d_rec_e, g->order_e_count, lo, to, from, rm, d_result, dbg_nostore, 0, items);
    else if (g->def_codec && lean) LAUNCH_P(g, "k_scan_extras", (k_scan_extras_lean<0, false>), grid, 128, 0, s, gd, g->d_rec_e, g->order_e_count, lo, to, from, rm, d_result, dbg_nostore, 0, items);
    else if (g->def_codec) LAUNCH_P(g, "k_scan_extras", k_scan_extras<true>, grid, 128, 0, s, gd, g->d_rec_e, g->order_e_count, lo, to, from, rm, d_result);
    else LAUNCH_P(g, "k_scan_extras", k_scan_extras<false>, grid, 128, 0, s, gd, g->d_rec_e, g->order_e_count, lo, to, from, rm, d_result);
    if (g->nlong) {
        if (g->n_items_resid && !stream_extras) {
            if (hist) LAUNCH_P(g, "k_long_resid", (k_long_resid<true, RowMap, true>), grid_for(g->n_items_resid, 64), 64, 0, sa, gd, li, g->item_map(0), (int64_t)0, g->n_items_resid, lo, to, rm, ld, lf);
            else if (g->def_codec) LAUNCH_P(g, "k_long_resid", (k_long_resid<true, RowMap>), grid_for(g->n_items_resid, 64), 64, 0, sa, gd, li, g->item_map(0), (int64_t)0, g->n_items_resid, lo, to, rm, ld, lf);
            else LAUNCH_P(g, "k_long_resid", (k_long_resid<false, RowMap>), grid_for(g->n_items_resid, 64), 64, 0, sa, gd, li, g->item_map(0), (int64_t)0, g->n_items_resid, lo, to, rm, ld, lf);
        }
        if (stream_extras) CK(meet());   // the residuals of the long records were written on the main stream
        if (g->n_items_extras && hist) LAUNCH_P(g, "k_long_extras", (k_long_extras<RowMap, true>), grid_for(g->n_items_extras, 64), 64, 0, sa, gd, li, g->item_map(2), (int64_t)0, g->n_items_extras, lo, to, rm, ld, lf);
        else if (g->n_items_extras) LAUNCH_P(g, "k_long_extras", k_long_extras<RowMap>, grid_for(g->n_items_extras, 64), 64, 0, sa, gd, li, g->item_map(2), (int64_t)0, g->n_items_extras, lo, to, rm, ld, lf);
    }
    for (int32_t level = 1; level <= g->max_depth; level++) {
        CK(meet());
        const int64_t a = g->level_start[(size_t)level - 1], c = g->level_start[(size_t)level] - a;
        if (c > 0) {
            const unsigned gm = (unsigned)std::min<int64_t>(grid_m, (c + 127) / 128);
            static const bool lean_m = !(getenv("BVG_MERGE_LEAN") && atoi(getenv("BVG_MERGE_LEAN")) == 0);
            // 8 resident blocks per SM (64 registers) and 4 parent loads in flight per lane: measured against 8/8, 6/8 and
            // 5/16 (blocks / batch): 1.92, 2.12, 2.19, 2.52 ms for the three levels -- occupancy beats deeper batching here
            // (10 / 12 resident blocks, 48 / 40 registers: 1.89 -> 2.16 / 2.50 ms, spills)
            // (measured and dropped in round 2: parent rows read in aligned 16-byte groups, 1.89 -> 3.57 ms)
            if (hist) LAUNCH_P(g, "k_scan_merge", (k_scan_merge_lean<8, 4, true>), gm, 128, 0, s, gd, g->d_rec_m + a, c, lo, to, from, rm, d_result, 0);
            else if (g->def_codec && lean_m) LAUNCH_P(g, "k_scan_merge", (k_scan_merge_lean<8, 4>), gm, 128, 0, s, gd, g->d_rec_m + a, c, lo, to, from, rm, d_result, 0);
            else if (g->def_codec) LAUNCH_P(g, "k_scan_merge", k_scan_merge<true>, gm, 128, 0, s, gd, g->d_rec_m + a, c, lo, to, from, rm, d_result);
            else LAUNCH_P(g, "k_scan_merge", k_scan_merge<false>, gm, 128, 0, s, gd, g->d_rec_m + a, c, lo, to, from, rm, d_result);
        }
        if (g->nlong) {
            const int64_t mc = g->n_items_merge[(size_t)level];
            if (mc > 0 && hist) LAUNCH_P(g, "k_long_merge", (k_long_merge<RowMap, true>), grid_for(mc, 64), 64, 0, sa, gd, li, g->item_map(2 + level), (int64_t)0, mc, lo, to, rm, ld, lf);
            else if (mc > 0) LAUNCH_P(g, "k_long_merge", k_long_merge<RowMap>, grid_for(mc, 64), 64, 0, sa, gd, li, g->item_map(2 + level), (int64_t)0, mc, lo, to, rm, ld, lf);
        }
    }
    // (stored long records are folded where their final rows are produced: k_long_resid / k_long_extras / k_long_merge)
    if (sa != s) { CK(cudaEventRecord(g->ev_join, sa)); CK(cudaStreamWaitEvent(s, g->ev_join, 0)); }
    CK(cudaGetLastError());
    return BVG_OK;
}

}  // extern "C" (templates cannot have C linkage)
// The scan through the tile kernel (bvg_tile.cuh): one launch over the tiles that touch [from, to).
template <int K, int NT, int MINB>
static int launch_tile_scan(const bvg_graph* g, const TileArgs& a, unsigned grid, cudaStream_t s) {
    CK(cudaFuncSetAttribute(k_tile_scan<K, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem_bytes));
    LAUNCH_P(g, "k_tile_scan", (k_tile_scan<K, NT, MINB>), grid, NT, a.smem_bytes, s, g->dev(), a);
    CK(cudaGetLastError());
    return BVG_OK;
}
extern "C" {

static int enqueue_scan_tiles(const bvg_graph* g, int32_t from, int32_t to, unsigned long long* d_result) {
    cudaStream_t s = g->stream;
    // tiles are in node order and partition the extent: the first one that ends after `from`, the last one that starts before `to`
    const auto& tf = g->h_tile_from;
    const int32_t t0 = (int32_t)(std::upper_bound(tf.begin(), tf.end(), from) - tf.begin()) - 1;
    const int32_t t1 = (int32_t)(std::lower_bound(tf.begin(), tf.end(), to) - tf.begin());
    if (t0 < 0 || t1 <= t0) return BVG_EINVAL;
    Tmp<int32_t> scr(s);
    CK(scr.alloc((size_t)g->long_scan_entries + 8));
    TileArgs a{};
    a.tiles = g->d_tiles;
    a.order = (t0 == 0 && t1 == g->ntiles) ? g->d_tile_order : nullptr;  // whole extent: heaviest tiles first
    a.first = t0; a.count = t1 - t0;
    a.fold_lo = from; a.fold_hi = to;
    a.li = g->long_index();
    a.long_scr = scr.p;
    a.result = d_result;
    a.smem_bytes = g->tile_smem;
    a.timeline = nullptr;
    const unsigned grid = (unsigned)(t1 - t0);
    Tmp<unsigned long long> tline(s);
    const char* tl_path = getenv("BVG_TILE_TIMELINE");  // debugging: per-block phase times dumped to this file after the scan
    if (tl_path) { CK(tline.alloc((size_t)grid * 40)); CK(cudaMemsetAsync(tline.p, 0, (size_t)grid * 40 * 8, s)); a.timeline = tline.p; }
    struct Dump {
        const char* path; unsigned long long* p; unsigned grid; cudaStream_t s;
        ~Dump() {
            if (!path) return;
            std::vector<unsigned long long> h((size_t)grid * 40);
            if (cudaStreamSynchronize(s) == cudaSuccess && cudaMemcpy(h.data(), p, h.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
                if (FILE* f = fopen(path, "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
            }
            cudaGetLastError();
        }
    } dump{ tl_path, tline.p, grid, s };
    if (g->tile_nt == 512) return g->zetak == 3 ? launch_tile_scan<3, 512, 2>(g, a, grid, s) : launch_tile_scan<0, 512, 2>(g, a, grid, s);
    if (g->tile_nt == 64) return g->zetak == 3 ? launch_tile_scan<3, 64, 16>(g, a, grid, s) : launch_tile_scan<0, 64, 16>(g, a, grid, s);
    if (g->tile_nt == 32) return g->zetak == 3 ? launch_tile_scan<3, 32, 32>(g, a, grid, s) : launch_tile_scan<0, 32, 32>(g, a, grid, s);
    return g->zetak == 3 ? launch_tile_scan<3, 256, 4>(g, a, grid, s) : launch_tile_scan<0, 256, 4>(g, a, grid, s);
}

// Scan = decode into a stream-ordered scratch + checksum kernel (general path); ranges are split so that the scratch
// stays below 2^30 arcs.
// k_checksum's companion for the general path of bvg_indegrees: counts every successor of a decoded CSR chunk
__global__ void k_hist_rows(const int32_t* __restrict__ rows, int64_t n, uint32_t* __restrict__ hist, int64_t hist_len) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t y = (uint32_t)rows[i];
        if ((int64_t)y < hist_len) atomicAdd(hist + y, 1u);
    }
}

static int enqueue_scan(const bvg_graph* g, int32_t from, int32_t to, unsigned long long* d_result, uint32_t* hist = nullptr, int64_t hist_len = 0) {
    if (to == from) return BVG_OK;
    if (g->tile_ok && !hist) return enqueue_scan_tiles(g, from, to, d_result);
    int64_t ra, rb;
    int rc = fetch_rowoff(g, from, to, &ra, &rb);
    if (rc) return rc;
    const int64_t arcs = rb - ra;
    if (arcs > ((int64_t)1 << 30) && to - from > 1) {
        const int32_t mid = from + (to - from) / 2;
        rc = enqueue_scan(g, from, mid, d_result, hist, hist_len);
        if (rc) return rc;
        return enqueue_scan(g, mid, to, d_result, hist, hist_len);
    }
    cudaStream_t s = g->stream;
    Tmp<int32_t> rows(s);
    CK(rows.alloc((size_t)arcs));
    // (the fused consumer counts in the lean walkers' fold: default codings; anything else decodes and counts the rows)
    if (g->d_is_parent && g->max_depth <= MAX_LEVEL_KEYS && ((int64_t)to - from) * 4 >= (int64_t)g->node_hi - g->node_lo && (!hist || g->def_codec)) {
        rc = ensure_schedules(g);
        if (rc) return rc;
        return enqueue_scan_fused(g, from, to, rows.p, ra, d_result, hist, hist_len);
    }
    rc = enqueue_decode(g, exec_of(g), from, to, rows.p, ra);
    if (rc) return rc;
    const int64_t cnt = (int64_t)to - from;
    const unsigned grid = (unsigned)std::min<int64_t>(148 * 8, std::max<int64_t>(1, (cnt + 7) / 8));
    LAUNCH_P(g, "k_checksum", k_checksum, grid, 256, 0, s, rows.p, g->d_rowoff + (from - g->node_lo), from, cnt, d_result);
    if (hist) LAUNCH_P(g, "k_hist_rows", k_hist_rows, 148 * 8, 256, 0, s, rows.p, arcs, hist, hist_len);
    CK(cudaGetLastError());
    return BVG_OK;
}

int bvg_scan_range_async(const bvg_graph* g, int32_t from, int32_t to, void* d_result) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    int rc = range_check(g, from, to);
    if (rc) return rc;
    if (!d_result) return BVG_EINVAL;
    DeviceGuard dg(g->device);
    CK(cudaMemsetAsync(d_result, 0, 16, g->stream));
    // blocks fold into FOLD_SLOTS slot pairs (bvg_kernels.cuh, block_fold); one small kernel reduces them into d_result
    Tmp<unsigned long long> slots(g->stream);
    CK(slots.alloc(2 * FOLD_SLOTS));
    CK(cudaMemsetAsync(slots.p, 0, 2 * FOLD_SLOTS * sizeof(unsigned long long), g->stream));
    rc = enqueue_scan(g, from, to, slots.p);
    if (rc) return rc;
    LAUNCH(k_reduce_slots, 1, 256, 0, g->stream, slots.p, (unsigned long long*)d_result);
    CK(cudaGetLastError());
    return BVG_OK;
}

int bvg_scan_range(const bvg_graph* g, int32_t from, int32_t to, int64_t* arcs, uint64_t* checksum) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    int rc = range_check(g, from, to);
    if (rc) return rc;
    DeviceGuard dg(g->device);
    Tmp<unsigned long long> res(g->stream);
    CK(res.alloc(2));
    rc = bvg_scan_range_async(g, from, to, res.p);
    if (rc) return rc;
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, res.p, 16, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    const int e = fetch_error(g);
    if (e) return e;
    if (arcs) *arcs = (int64_t)h[0];
    if (checksum) *checksum = h[1];
    return BVG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// fused consumers of the sequential scan
// ------------------------------------------------------------------------------------------------------------

// The counting pass of a transposition (reference Transform.java:977-987: for every arc (x, y) numPred[y]++): a scan whose
// consumer counts instead of only folding; the successors never leave the device.
int bvg_indegrees(const bvg_graph* g, int32_t from, int32_t to, uint32_t* counts, int64_t counts_len, int on_device, int64_t* arcs) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    int rc = range_check(g, from, to);
    if (rc) return rc;
    if (!counts || counts_len < 0) return BVG_EINVAL;
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    Tmp<uint32_t> d_counts(s);
    uint32_t* hist = counts;
    if (!on_device) {
        CK(d_counts.alloc((size_t)std::max<int64_t>(counts_len, 1)));
        CK(cudaMemcpyAsync(d_counts.p, counts, (size_t)counts_len * 4, cudaMemcpyHostToDevice, s));
        hist = d_counts.p;
    }
    Tmp<unsigned long long> slots(s), res(s);
    CK(slots.alloc(2 * FOLD_SLOTS));
    CK(res.alloc(2));
    CK(cudaMemsetAsync(slots.p, 0, 2 * FOLD_SLOTS * sizeof(unsigned long long), s));
    CK(cudaMemsetAsync(res.p, 0, 16, s));
    rc = enqueue_scan(g, from, to, slots.p, hist, counts_len);
    if (rc) return rc;
    LAUNCH(k_reduce_slots, 1, 256, 0, s, slots.p, res.p);
    CK(cudaGetLastError());
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, res.p, 16, cudaMemcpyDeviceToHost, s));
    if (!on_device) CK(cudaMemcpyAsync(counts, d_counts.p, (size_t)counts_len * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int e = fetch_error(g);
    if (e) return e;
    if (arcs) *arcs = (int64_t)h[0];
    return BVG_OK;
}

// One level of a breadth-first visit: every successor y of the frontier that has no distance yet gets `level` and joins the
// next frontier (the marker test-and-set of ParallelBreadthFirstVisit.java:163-172).
__global__ void k_bfs_expand(const int32_t* __restrict__ succ, int64_t n, int32_t* __restrict__ dist, int64_t dist_len, int32_t level,
                             int32_t* __restrict__ next, unsigned int* __restrict__ next_count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t y = succ[i];
        if (y >= 0 && y < dist_len && dist[y] < 0 && atomicCAS(dist + y, -1, level) == -1) next[atomicAdd(next_count, 1u)] = y;
    }
}
__global__ void k_fill_i32(int32_t* __restrict__ p, int64_t n, int32_t v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// One HyperBall iteration (reference algo/HyperBall.java:875-915, the non-systolic branch: every node enumerates its successors):
// out[x] = register-wise max(in[x], in[s] for every successor s of x), x in [from, to).  Rows are decoded chunk by chunk into
// a scratch CSR on the device and consumed there; nothing but the counters crosses PCIe.
int bvg_hyperball_step(const bvg_graph* g, int32_t from, int32_t to, int log2m, const uint8_t* in, uint8_t* out, int on_device, int64_t* modified) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    int rc = range_check(g, from, to);
    if (rc) return rc;
    if (!in || !out || log2m < 4 || log2m > 9) return BVG_EINVAL;   // 16 .. 512 registers of one byte
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    const int lanes_log = log2m - 4;
    const size_t m = (size_t)1 << log2m, total = (size_t)g->n_total * m;
    Tmp<uint8_t> d_in(s), d_out(s);
    const uint8_t* din = in;
    uint8_t* dout = out;
    if (!on_device) {
        CK(d_in.alloc(total));
        CK(d_out.alloc(total));
        CK(cudaMemcpyAsync(d_in.p, in, total, cudaMemcpyHostToDevice, s));
        din = d_in.p; dout = d_out.p;
    }
    if (((uintptr_t)din | (uintptr_t)dout) & 15) return BVG_EINVAL;   // counters are read and written 16 bytes at a time
    Tmp<unsigned long long> d_mod(s);
    Tmp<int32_t> d_heavy(s), d_nheavy(s);
    CK(d_mod.alloc(1));
    CK(d_nheavy.alloc(1));
    CK(cudaMemsetAsync(d_mod.p, 0, 8, s));
    // chunks of whole 2^ORDER_CHUNK_LOG-node schedule chunks holding about HB_CHUNK_ARCS arcs
    const int64_t chunk_arcs = (int64_t)env_int("BVG_HB_CHUNK_MARCS", 256, 1, 1 << 14) << 20;
    int32_t lo = from;
    while (lo < to) {
        int32_t hi = (int32_t)std::min<int64_t>(to, (((int64_t)lo >> ORDER_CHUNK_LOG) + 1) << ORDER_CHUNK_LOG);
        int64_t ra, rb;
        rc = fetch_rowoff(g, lo, hi, &ra, &rb);
        if (rc) return rc;
        while (hi < to) {   // extend while the chunk stays under the arc budget
            const int32_t nh = (int32_t)std::min<int64_t>(to, (int64_t)hi + ((int64_t)1 << ORDER_CHUNK_LOG));
            int64_t r2a, r2b;
            rc = fetch_rowoff(g, lo, nh, &r2a, &r2b);
            if (rc) return rc;
            if (r2b - r2a > chunk_arcs) break;
            hi = nh; rb = r2b;
        }
        const int64_t arcs = rb - ra, cnt = (int64_t)hi - lo;
        Tmp<int64_t> d_off(s);
        Tmp<int32_t> d_rows(s);
        CK(d_off.alloc((size_t)cnt + 1));
        CK(d_rows.alloc((size_t)std::max<int64_t>(arcs, 1)));
        CK(d_heavy.alloc((size_t)std::max<int64_t>(1, arcs / HB_HEAVY + 1)));
        CK(cudaMemsetAsync(d_nheavy.p, 0, 4, s));
        rc = bvg_decode_range(g, lo, hi, d_off.p, d_rows.p, arcs, 1);
        if (rc) return rc;
        LAUNCH_P(g, "k_hb_update", k_hb_update, grid_for(cnt << lanes_log, HB_THREADS), HB_THREADS, 0, s, d_rows.p, d_off.p, lo, cnt,
                 (const uint4*)din, (uint4*)dout, lanes_log, d_heavy.p, d_nheavy.p, d_mod.p);
        int32_t nheavy = 0;
        CK(cudaMemcpyAsync(&nheavy, d_nheavy.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (nheavy > 0) LAUNCH_P(g, "k_hb_update_heavy", k_hb_update_heavy, (unsigned)nheavy, HB_THREADS, 0, s, d_rows.p, d_off.p, lo, d_heavy.p,
                                 (const uint4*)din, (uint4*)dout, lanes_log, d_mod.p);
        CK(cudaGetLastError());
        lo = hi;
    }
    unsigned long long hm = 0;
    CK(cudaMemcpyAsync(&hm, d_mod.p, 8, cudaMemcpyDeviceToHost, s));
    if (!on_device) {
        const size_t a = (size_t)from * m, b = (size_t)to * m;
        if (b > a) CK(cudaMemcpyAsync(out + a, d_out.p + a, b - a, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    const int e = fetch_error(g);
    if (e) return e;
    if (modified) *modified = (int64_t)hm;
    return BVG_OK;
}

// Breadth-first visit from `source` (reference algo/ParallelBreadthFirstVisit.java:155-181: frontier by frontier, the
// successors of a frontier decoded by random access): dist[x] = distance from source, -1 when unreachable.  Frontier,
// successor lists and distances stay on the device; the host sees one counter per level.
int bvg_bfs(const bvg_graph* g, int32_t source, int32_t* dist, int on_device, int32_t* levels, int64_t* reached) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    if (!g || !dist) return BVG_EINVAL;
    if (source < g->ext_from || source >= g->ext_to) return BVG_EINVAL;
    if (g->offset_type <= 0) return BVG_EUNSUPPORTED;  // random access (BVGraph.java:901)
    if (g->ext_from != 0 || g->ext_to != g->n_total) return BVG_EUNSUPPORTED;  // a visit follows arcs anywhere: whole graphs (replicas) only
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    const int64_t n = g->n_total;
    Tmp<int32_t> d_dist(s), fa(s), fb(s), lists(s);
    Tmp<int64_t> off(s);
    Tmp<unsigned int> cnt(s);
    int32_t* dd = dist;
    if (!on_device) { CK(d_dist.alloc((size_t)n)); dd = d_dist.p; }
    CK(fa.alloc((size_t)n));
    CK(fb.alloc((size_t)n));
    CK(cnt.alloc(1));
    LAUNCH(k_fill_i32, 148 * 4, 256, 0, s, dd, n, -1);
    LAUNCH(k_set_i32, 1, 1, 0, s, dd + source, 0);
    LAUNCH(k_set_i32, 1, 1, 0, s, fa.p, source);
    CK(cudaGetLastError());
    int32_t *cur = fa.p, *nxt = fb.p;
    int64_t nf = 1, total = 1;
    int32_t level = 0;
    size_t off_cap = 0, lists_cap = 0;
    while (nf > 0) {
        if ((size_t)nf + 1 > off_cap) { off_cap = (size_t)nf + 1 + (size_t)nf / 2; Tmp<int64_t> t(s); CK(t.alloc(off_cap)); std::swap(t.p, off.p); }
        int rc = bvg_successors_batch(g, cur, nf, off.p, nullptr, 0, 1);   // sizes
        if (rc) return rc;
        int64_t arcs = 0;
        CK(cudaMemcpyAsync(&arcs, off.p + nf, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        unsigned int next_n = 0;
        if (arcs > 0) {
            if ((size_t)arcs > lists_cap) { lists_cap = (size_t)arcs + (size_t)arcs / 2; Tmp<int32_t> t(s); CK(t.alloc(lists_cap)); std::swap(t.p, lists.p); }
            rc = bvg_successors_batch(g, cur, nf, off.p, lists.p, arcs, 1);
            if (rc) return rc;
            CK(cudaMemsetAsync(cnt.p, 0, 4, s));
            LAUNCH(k_bfs_expand, 148 * 4, 256, 0, s, lists.p, arcs, dd, n, level + 1, nxt, cnt.p);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&next_n, cnt.p, 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        }
        nf = next_n;
        total += nf;
        if (nf > 0) level++;
        std::swap(cur, nxt);
    }
    if (!on_device) { CK(cudaMemcpyAsync(dist, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
    const int e = fetch_error(g);
    if (e) return e;
    if (levels) *levels = level;
    if (reached) *reached = total;
    return BVG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// random access
// ------------------------------------------------------------------------------------------------------------

int bvg_outdegree_batch(const bvg_graph* g, const int32_t* xs, int32_t from, int64_t nx, int32_t* d, int on_device) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    if (!g || nx < 0 || !d) return BVG_EINVAL;
    if (g->offset_type <= 0) return BVG_ESTATE;  // BVGraph.java:869
    if (nx == 0) return BVG_OK;
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    GraphDev gd = g->dev();
    if (on_device) {
        LAUNCH(k_gather_outdeg, grid_for(nx, 256), 256, 0, s, gd, xs, from, nx, d);
        CK(cudaGetLastError());
        return BVG_OK;
    }
    Tmp<int32_t> d_xs(s), d_d(s);
    if (xs) { CK(d_xs.alloc((size_t)nx)); CK(cudaMemcpyAsync(d_xs.p, xs, (size_t)nx * 4, cudaMemcpyHostToDevice, s)); }
    CK(d_d.alloc((size_t)nx));
    LAUNCH(k_gather_outdeg, grid_for(nx, 256), 256, 0, s, gd, xs ? d_xs.p : nullptr, from, nx, d_d.p);
    CK(cudaMemcpyAsync(d, d_d.p, (size_t)nx * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int e = fetch_error(g);
    return e ? e : BVG_OK;
}

int bvg_outdegree(const bvg_graph* g, int32_t x, int32_t* d) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    if (!g || !d) return BVG_EINVAL;
    if (x < g->ext_from || x >= g->ext_to) return BVG_EINVAL;  // BVGraph.java:860
    if (g->offset_type <= 0) return BVG_ESTATE;                 // :869
    DeviceGuard dg(g->device);
    CK(cudaMemcpyAsync(d, g->d_outdeg + (x - g->node_lo), 4, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    return BVG_OK;
}

int bvg_successors_batch(const bvg_graph* g, const int32_t* xs, int64_t nx, int64_t* out_off, int32_t* out, int64_t cap, int on_device) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    if (!g || nx < 0 || !out_off || (!xs && nx)) return BVG_EINVAL;
    if (g->offset_type <= 0) return BVG_EUNSUPPORTED;  // BVGraph.java:901
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    GraphDev gd = g->dev();
    Tmp<int32_t> d_xs(s), dq(s), need(s), d_out(s), scratch(s);
    Tmp<int64_t> d_off(s), scratch_off(s);
    const int32_t* xs_dev = xs;
    if (!on_device) {
        CK(d_xs.alloc((size_t)nx));
        if (nx) CK(cudaMemcpyAsync(d_xs.p, xs, (size_t)nx * 4, cudaMemcpyHostToDevice, s));
        xs_dev = d_xs.p;
    }
    CK(dq.alloc((size_t)nx));
    CK(need.alloc((size_t)nx));
    CK(scratch_off.alloc((size_t)nx + 1));
    int64_t* off_dev = out_off;
    if (!on_device) { CK(d_off.alloc((size_t)nx + 1)); off_dev = d_off.p; }
    // queries whose chain holds a long record go through the range kernels (k_query_sizes marks their chains)
    const int64_t nn = (int64_t)g->node_hi - g->node_lo;
    const bool split_long = g->nlong > 0 && g->max_depth <= MAX_LEVEL_KEYS;
    Tmp<uint8_t> mask(s), heavy(s);
    Tmp<int32_t> nheavy(s);
    if (split_long) {
        CK(mask.alloc((size_t)nn));
        CK(heavy.alloc((size_t)std::max<int64_t>(nx, 1)));
        CK(nheavy.alloc(1));
        CK(cudaMemsetAsync(mask.p, 0, (size_t)nn, s));
        CK(cudaMemsetAsync(nheavy.p, 0, 4, s));
    }
    // A batch that asks for a sizeable part of the graph goes through the range kernels as a whole (every chain marked, rows
    // gathered afterwards): their length-sorted schedules and lean walkers decode an arc several times cheaper than k_random's
    // one thread per query, which is the better deal only while the batch touches a small part of the schedules (measured on
    // the 32 M-node graph, profiles/c4_sweep.py: 10 M queries 27.2 -> 13.0 ms, 3 M 10.2 -> 8.0, 1 M 4.1 -> 4.4, 100 k 1.6 -> 1.8).
    static const int range_pct = env_int("BVG_RANDOM_RANGE_PCT", 5, 0, 100000);   // batch size in percent of the nodes from which on
    const bool all_range = split_long && nx > 0 && (double)nx * 100.0 >= (double)range_pct * (double)nn;
    if (nx) LAUNCH(k_query_sizes, grid_for(nx, 256), 256, 0, s, gd, xs_dev, nx, dq.p, need.p, all_range ? -1 : g->long_d, mask.p, heavy.p, nheavy.p);
    int rc = device_exclusive_scan(s, dq.p, nx, off_dev);
    if (rc) return rc;
    rc = device_exclusive_scan(s, need.p, nx, scratch_off.p);
    if (rc) return rc;
    int64_t tot[2];
    int32_t n_heavy = 0;
    CK(cudaMemcpyAsync(&tot[0], off_dev + nx, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&tot[1], scratch_off.p + nx, 8, cudaMemcpyDeviceToHost, s));
    if (split_long) CK(cudaMemcpyAsync(&n_heavy, nheavy.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    int e = fetch_error(g);
    if (e) return e;
    if (!on_device) CK(cudaMemcpyAsync(out_off, off_dev, ((size_t)nx + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (!out) { CK(cudaStreamSynchronize(s)); return BVG_OK; }
    if (cap < tot[0]) return BVG_ENOMEM;
    int32_t* out_dev = out;
    if (!on_device) { CK(d_out.alloc((size_t)tot[0])); out_dev = d_out.p; }
    CK(scratch.alloc((size_t)tot[1]));
    const uint8_t* heavy_flags = n_heavy > 0 ? heavy.p : nullptr;
    if (nx) {
        if (g->def_codec) LAUNCH_P(g, "k_random", k_random<true>, grid_for(nx, 128), 128, 0, s, gd, xs_dev, nx, off_dev, out_dev, scratch_off.p, scratch.p, heavy_flags);
        else LAUNCH_P(g, "k_random", k_random<false>, grid_for(nx, 128), 128, 0, s, gd, xs_dev, nx, off_dev, out_dev, scratch_off.p, scratch.p, heavy_flags);
    }
    Tmp<int32_t> rows(s);
    if (n_heavy > 0) {  // the marked chains, decoded by the range kernels into a scratch laid out like the CSR
        int64_t ra, rb;
        rc = fetch_rowoff(g, g->node_lo, g->node_hi, &ra, &rb);
        if (rc) return rc;
        CK(rows.alloc((size_t)(rb - ra)));
        RowMap rm;
        rm.out = rows.p; rm.out_base = ra; rm.from = g->node_lo; rm.halo = nullptr; rm.halo_off = nullptr; rm.halo_lo = g->node_lo; rm.halo_base = ra;
        rm.mask = mask.p;
        rc = ensure_schedules(g);
        if (rc) return rc;
        rc = run_ordered_decode(g, exec_of(g), g->node_lo, g->node_hi, g->node_lo, rm);
        if (rc) return rc;
        LAUNCH_P(g, "k_gather_rows", k_gather_rows, (unsigned)std::min<int64_t>(148 * env_int("BVG_GATHER_BLOCKS", 32, 1, 1024), std::max<int64_t>(1, (nx + 7) / 8)), 256, 0, s, gd, xs_dev, nx, heavy.p, off_dev, out_dev, rm);
    }
    CK(cudaGetLastError());
    if (on_device) return BVG_OK;
    if (tot[0]) CK(cudaMemcpyAsync(out, out_dev, (size_t)tot[0] * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    e = fetch_error(g);
    return e ? e : BVG_OK;
}

int bvg_successors(const bvg_graph* g, int32_t x, int32_t* out, int32_t cap, int32_t* d) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    if (!g || !d) return BVG_EINVAL;
    if (x < g->ext_from || x >= g->ext_to) return BVG_EINVAL;  // BVGraph.java:900
    if (g->offset_type <= 0) return BVG_EUNSUPPORTED;          // :901
    int32_t deg;
    int rc = bvg_outdegree(g, x, &deg);
    if (rc) return rc;
    if (deg > cap || (!out && deg)) return BVG_ENOMEM;
    int64_t off[2];
    rc = bvg_successors_batch(g, &x, 1, off, out, cap, 0);
    if (rc) return rc;
    *d = deg;
    return BVG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// NodeIterator
// ------------------------------------------------------------------------------------------------------------

// A cursor holds two batches of decoded nodes in pinned host memory: the host iterates over one while the device decodes
// the next one and copies it out (the reference's iterator decodes lazily, one successor per nextInt(); a per-node
// round trip to the device would cost a launch per node).
struct CursorBatch {
    int32_t lo = 0, hi = 0;              // nodes held (or being decoded)
    int64_t arcs = 0;
    int64_t* h_off = nullptr;            // pinned: hi - lo + 1 relative arc offsets
    int32_t* h_succ = nullptr;           // pinned: arcs successors
    size_t hcap_nodes = 0, hcap_arcs = 0;
    int64_t* d_off = nullptr;
    int32_t* d_succ = nullptr;
    size_t dcap_nodes = 0, dcap_arcs = 0;
    cudaEvent_t ready = nullptr;
    bool pending = false;                // enqueued, not yet waited for
};
struct bvg_cursor {
    const bvg_graph* g;
    cudaStream_t stream = nullptr;   // every cursor decodes on a stream of its own and reports into an error word of its own:
    ErrWord* d_err = nullptr;        // cursors of one graph run beside each other, one per host thread
    Exec ex() const { return Exec{ stream, d_err }; }
    int32_t next;        // next node to return
    int32_t upper;       // no node >= upper is returned (BVGraph.java:1185)
    CursorBatch b[2];
    int cur = 0;         // b[cur] is the batch being iterated; b[1 - cur] the one in flight
    bool have = false;   // b[cur] holds [lo, hi)
};
// A batch is one schedule chunk (2^18 nodes, cut at chunk boundaries): a range decode walks the schedule slices of the chunks it
// touches, so a chunk-aligned batch costs exactly its own records.
static const int32_t CURSOR_BATCH_NODES = 1 << ORDER_CHUNK_LOG;

static void cursor_release(bvg_cursor* c) {
    for (CursorBatch& b : c->b) {
        if (b.pending) { cudaEventSynchronize(b.ready); b.pending = false; }
        pin_free(b.h_off);
        pin_free(b.h_succ);
        dev_free(b.d_off, c->stream);
        dev_free(b.d_succ, c->stream);
        if (b.ready) cudaEventDestroy(b.ready);
        b = CursorBatch();
    }
    if (c->d_err) { dev_free(c->d_err, c->stream); c->d_err = nullptr; }
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); c->stream = nullptr; }
    cudaGetLastError();
}

static int cursor_init(bvg_cursor* c) {
    DeviceGuard dg(c->g->device);
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(dev_alloc((void**)&c->d_err, sizeof(ErrWord), c->stream));
    CK(cudaMemsetAsync(c->d_err, 0, sizeof(ErrWord), c->stream));
    return BVG_OK;
}

// Enqueues decode + copy-out of the batch starting at `lo` into b; returns without waiting.
static int cursor_enqueue(bvg_cursor* c, CursorBatch& b, int32_t lo) {
    const bvg_graph* g = c->g;
    cudaStream_t s = c->stream;
    // up to the next chunk boundary (counted from the first loaded node, as the schedules count them)
    const int64_t next_chunk = ((((int64_t)lo - g->node_lo) >> ORDER_CHUNK_LOG) + 1) << ORDER_CHUNK_LOG;
    int32_t hi = (int32_t)std::min<int64_t>(c->upper, next_chunk + g->node_lo);
    int64_t ra = 0, rb = 0;
    for (;;) {
        int rc = fetch_rowoff(g, lo, hi, &ra, &rb, s);
        if (rc) return rc;
        if (rb - ra <= ((int64_t)1 << 26) || hi - lo <= 1) break;
        hi = lo + (hi - lo) / 2;
    }
    const int64_t arcs = rb - ra, cnt = (int64_t)hi - lo;
    if ((size_t)cnt + 1 > b.hcap_nodes) {
        pin_free(b.h_off);
        b.h_off = nullptr; b.hcap_nodes = 0;
        CK(pin_alloc((void**)&b.h_off, ((size_t)CURSOR_BATCH_NODES + 1) * 8));
        b.hcap_nodes = (size_t)CURSOR_BATCH_NODES + 1;
    }
    if ((size_t)arcs > b.hcap_arcs) {
        pin_free(b.h_succ);
        b.h_succ = nullptr; b.hcap_arcs = 0;
        const size_t want = std::max<size_t>((size_t)arcs + (size_t)arcs / 4, (size_t)1 << 20);
        CK(pin_alloc((void**)&b.h_succ, want * 4));
        b.hcap_arcs = pin_capacity(b.h_succ) / 4;
    }
    if ((size_t)cnt + 1 > b.dcap_nodes) {
        dev_free(b.d_off, s); b.d_off = nullptr; b.dcap_nodes = 0;
        CK(dev_alloc((void**)&b.d_off, ((size_t)CURSOR_BATCH_NODES + 1) * 8, s));
        b.dcap_nodes = (size_t)CURSOR_BATCH_NODES + 1;
    }
    if ((size_t)arcs > b.dcap_arcs) {
        dev_free(b.d_succ, s); b.d_succ = nullptr; b.dcap_arcs = 0;
        const size_t want = std::max<size_t>((size_t)arcs + (size_t)arcs / 4, (size_t)1 << 20);
        CK(dev_alloc((void**)&b.d_succ, want * 4, s));
        b.dcap_arcs = want;
    }
    if (!b.ready) CK(cudaEventCreateWithFlags(&b.ready, cudaEventDisableTiming));
    LAUNCH(k_rel_offsets, grid_for(cnt + 1, 256), 256, 0, s, g->d_rowoff + (lo - g->node_lo), cnt, b.d_off);
    if (arcs) { const int rc = enqueue_decode(g, c->ex(), lo, hi, b.d_succ, ra); if (rc) return rc; }
    CK(cudaMemcpyAsync(b.h_off, b.d_off, ((size_t)cnt + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (arcs) CK(cudaMemcpyAsync(b.h_succ, b.d_succ, (size_t)arcs * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(b.ready, s));
    b.lo = lo; b.hi = hi; b.arcs = arcs; b.pending = true;
    return BVG_OK;
}

// Makes b[cur] the batch that holds c->next and starts the one after it.
static int cursor_refill(bvg_cursor* c) {
    const bvg_graph* g = c->g;
    DeviceGuard dg(g->device);
    CursorBatch& nb = c->b[1 - c->cur];
    if (!(nb.pending && nb.lo == c->next)) {  // first call, or the iteration did not run into the prefetched batch
        if (nb.pending) { CK(cudaEventSynchronize(nb.ready)); nb.pending = false; }
        const int rc = cursor_enqueue(c, nb, c->next);
        if (rc) return rc;
    }
    CK(cudaEventSynchronize(nb.ready));
    nb.pending = false;
    const int e = fetch_error(g, c->ex());
    if (e) return e;
    c->cur = 1 - c->cur;
    c->have = true;
    if (c->b[c->cur].hi < c->upper) {  // the host iterates over b[cur] while the device fills the other one
        const int rc = cursor_enqueue(c, c->b[1 - c->cur], c->b[c->cur].hi);
        if (rc) return rc;
    }
    return BVG_OK;
}

int bvg_cursor_open(const bvg_graph* g, int32_t from, int32_t upper, bvg_cursor** out) {
    if (!g || !out) return BVG_EINVAL;
    if (from < g->ext_from || from > g->ext_to) return BVG_EINVAL;  // BVGraph.java:1165
    if (from != 0 && g->offset_type <= 0) return BVG_ESTATE;        // :1174
    bvg_cursor* c = new (std::nothrow) bvg_cursor();
    if (!c) return BVG_ENOMEM;
    c->g = g; c->next = from; c->upper = std::min(upper, g->ext_to);
    const int rc = cursor_init(c);
    if (rc) { cursor_release(c); delete c; return rc; }
    *out = c;
    return BVG_OK;
}

int bvg_cursor_next(bvg_cursor* c, int32_t* node, int32_t* d, const int32_t** succ) {
    if (!c) return BVG_EINVAL;
    if (c->next >= c->upper) return BVG_EEND;  // BVGraph.java:1202
    if (!c->have || c->next >= c->b[c->cur].hi || c->next < c->b[c->cur].lo) {
        const int rc = cursor_refill(c);
        if (rc) return rc;
    }
    const CursorBatch& b = c->b[c->cur];
    const size_t i = (size_t)(c->next - b.lo);
    if (node) *node = c->next;
    if (d) *d = (int32_t)(b.h_off[i + 1] - b.h_off[i]);
    if (succ) *succ = b.h_succ + b.h_off[i];
    c->next++;
    return BVG_OK;
}

// The batch that holds the cursor's next node, as it sits in pinned host memory (no copy): nodes first .. first + count - 1,
// off[0 .. count] arc offsets relative to succ (off[0] is the first remaining node's), succ the successors.  The cursor moves
// past the batch; the pointers stay valid until the next call on this cursor.  What a binding iterates over without a
// per-node FFI call (a JNI direct ByteBuffer, a numpy view).
int bvg_cursor_next_batch(bvg_cursor* c, int32_t* first, int32_t* count, const int64_t** off, const int32_t** succ) {
    if (!c || !first || !count || !off || !succ) return BVG_EINVAL;
    if (c->next >= c->upper) return BVG_EEND;
    if (!c->have || c->next >= c->b[c->cur].hi || c->next < c->b[c->cur].lo) {
        const int rc = cursor_refill(c);
        if (rc) return rc;
    }
    const CursorBatch& b = c->b[c->cur];
    const size_t i = (size_t)(c->next - b.lo);
    *first = c->next;
    *count = (int32_t)(std::min(b.hi, c->upper) - c->next);
    *off = b.h_off + i;
    *succ = b.h_succ;
    c->next += *count;
    return BVG_OK;
}

int bvg_cursor_copy(const bvg_cursor* c, int32_t upper, bvg_cursor** out) {  // BVGraph.java:1252-1260
    if (!c || !out) return BVG_EINVAL;
    bvg_cursor* n = new (std::nothrow) bvg_cursor();
    if (!n) return BVG_ENOMEM;
    n->g = c->g; n->next = c->next; n->upper = std::min(upper, c->g->ext_to);
    const int rc = cursor_init(n);
    if (rc) { cursor_release(n); delete n; return rc; }
    *out = n;
    return BVG_OK;
}

void bvg_cursor_close(bvg_cursor* c) {
    if (!c) return;
    DeviceGuard dg(c->g->device);
    cursor_release(c);
    delete c;
}

int bvg_cursor_drain(bvg_cursor* c, int64_t max_nodes, int64_t* nodes, int64_t* arcs, uint64_t* checksum) {
    if (!c) return BVG_EINVAL;
    int64_t nn = 0, na = 0;
    uint64_t cs = 0;
    while (max_nodes < 0 || nn < max_nodes) {
        int32_t x, d;
        const int32_t* succ;
        const int rc = bvg_cursor_next(c, &x, &d, &succ);
        if (rc == BVG_EEND) break;
        if (rc) return rc;
        const uint64_t base = (uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ull;
        for (int32_t i = 0; i < d; i++) cs ^= base + (uint64_t)(uint32_t)succ[i];
        na += d;
        nn++;
    }
    if (nodes) *nodes = nn;
    if (arcs) *arcs = na;
    if (checksum) *checksum = cs;
    return BVG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// shard boundaries
// ------------------------------------------------------------------------------------------------------------

int bvg_boundary_count(const bvg_graph* g, int32_t* count) {
    if (!g || !count) return BVG_EINVAL;
    if (g->window == 0 || g->maxref == 0) { *count = 0; return BVG_OK; }
    if (g->maxref < 0 || g->maxref > (1 << 16)) { *count = 0; return BVG_EUNSUPPORTED; }  // unbounded chains: shards re-decode
    *count = (int32_t)std::min<int64_t>((int64_t)g->window * g->maxref, (int64_t)g->ext_to - g->ext_from);
    return BVG_OK;
}

int bvg_boundary_export(const bvg_graph* g, int64_t* out_off, int32_t* out, int64_t cap, int on_device) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    int32_t cnt;
    int rc = bvg_boundary_count(g, &cnt);
    if (rc) return rc;
    return bvg_decode_range(g, g->ext_to - cnt, g->ext_to, out_off, out, cap, on_device);
}

int bvg_halo_needed(const bvg_graph* g, int32_t* first_needed_node) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    if (!g || !first_needed_node) return BVG_EINVAL;
    if (g->halo_first_known) { *first_needed_node = g->halo_first; return BVG_OK; }  // the graph is immutable: asked once
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    int32_t h = g->ext_from;
    const int64_t reach = std::min<int64_t>((int64_t)g->ext_to - g->ext_from, (int64_t)g->window * g->max_depth);
    if (reach > 0 && g->ext_from > g->node_lo) {
        Tmp<int32_t> hs(s);
        CK(hs.alloc(1));
        CK(cudaMemcpyAsync(hs.p, &h, 4, cudaMemcpyHostToDevice, s));
        LAUNCH(k_halo_start, grid_for(reach, 128), 128, 0, s, g->dev(), g->ext_from, (int32_t)reach, hs.p);
        CK(cudaMemcpyAsync(&h, hs.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    *first_needed_node = h;
    g->halo_first = h; g->halo_first_known = true;
    return BVG_OK;
}

int bvg_halo_import(bvg_graph* g, int32_t count, const int64_t* off, const int32_t* lists, int on_device) {
    std::unique_lock<std::recursive_mutex> call_lock;
    if (g) call_lock = std::unique_lock<std::recursive_mutex>(g->call_mu);
    if (!g || count < 0 || (count && (!off || !lists))) return BVG_EINVAL;
    // the imported lists are those of the nodes [ext_from - count, ext_from): they must lie inside the loaded window (row
    // offsets of the halo are taken from this shard's own index) ...
    if (count > g->ext_from - g->node_lo) return BVG_EINVAL;
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    g->halo_count = 0;
    if (count == 0) return BVG_OK;
    {   // ... and reach back at least as far as this shard's chains do: a shorter import would leave a parent without a row
        int32_t first = g->ext_from;
        const int rc = bvg_halo_needed(g, &first);
        if (rc) return rc;
        if (first < g->ext_from - count) return BVG_EINVAL;
    }
    int64_t total = 0;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (on_device && g->halo_import_count == count && g->halo_import_total >= 0 && count + 1 <= g->halo_off_cap) {
        // same shape as the previous import (a shard re-imports its neighbour's boundary every step): copy by a kernel that
        // reads the arc count on the device and checks it against the capacity, no round trip to the host
        // (an import larger than the buffers is caught on the device: k_halo_copy reports E_NOMEM and copies nothing, the next
        // call that fetches the error word fails; the offsets of the previous import stay in place, so no read goes out of bounds)
        LAUNCH(k_halo_copy, 64, 256, 0, s, off, lists, count, g->d_halo_off, g->d_halo_lists, g->halo_lists_cap, g->d_err);
        CK(cudaGetLastError());
        g->halo_count = count;
        return BVG_OK;
    }
    if (on_device) { CK(cudaMemcpyAsync(&total, off + count, 8, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
    else total = off[count];
    g->halo_import_count = count; g->halo_import_total = total;
    if (count + 1 > g->halo_off_cap) {
        dev_free(g->d_halo_off, g->stream); g->d_halo_off = nullptr; g->halo_off_cap = 0;
        CK(dev_alloc((void**)&g->d_halo_off, ((size_t)count + 1) * 8, g->stream));
        g->halo_off_cap = count + 1;
    }
    if (total > g->halo_lists_cap) {
        dev_free(g->d_halo_lists, g->stream); g->d_halo_lists = nullptr; g->halo_lists_cap = 0;
        CK(dev_alloc((void**)&g->d_halo_lists, (size_t)total * 2 * 4, g->stream));
        g->halo_lists_cap = total * 2;
    }
    CK(cudaMemcpyAsync(g->d_halo_off, off, ((size_t)count + 1) * 8, kind, s));
    if (total) CK(cudaMemcpyAsync(g->d_halo_lists, lists, (size_t)total * 4, kind, s));
    if (!on_device) CK(cudaStreamSynchronize(s));
    g->halo_count = count;
    return BVG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------------------------

const char* bvg_strerror(int status) {
    switch (status) {
        case BVG_OK: return "ok";
        case BVG_EINVAL: return "node index out of range / invalid argument (IllegalArgumentException)";
        case BVG_ESTATE: return "illegal state: offsets not loaded, reference beyond the window, or iterator not advanced (IllegalStateException)";
        case BVG_EUNSUPPORTED: return "unsupported: random access without offsets, or a coding this build does not decode (UnsupportedOperationException)";
        case BVG_EIO: return "I/O error or truncated stream (IOException)";
        case BVG_EFORMAT: return "malformed properties or impossible record";
        case BVG_ENOMEM: return "out of memory or output buffer too small";
        case BVG_ECUDA: return t_cuda_msg[0] ? t_cuda_msg : "no usable CUDA device / CUDA runtime error (there is no CPU decode path)";
        case BVG_EEND: return "no more nodes (NoSuchElementException)";
        default: return "unknown status";
    }
}

int bvg_last_error_node(const bvg_graph* g, int32_t* node, int64_t* bitpos) {
    if (!g) return BVG_EINVAL;
    std::lock_guard<std::mutex> lk(g->mu);
    if (node) *node = g->err_node;
    if (bitpos) *bitpos = g->err_bitpos;
    return BVG_OK;
}

int64_t bvg_kernel_launches(void) { return g_launches.load(); }

int64_t bvg_release_cached_memory(int device) {
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    int64_t freed = 0;
    int prev = -1;
    cudaGetDevice(&prev);
    for (auto it = c.idle.begin(); it != c.idle.end();) {
        if (device < 0 || it->second.dev == device) {
            cudaSetDevice(it->second.dev);
            cudaEventDestroy(it->second.ev);
            cudaFree(it->second.p);  // waits for whatever may still use the block
            freed += (int64_t)it->second.bytes;
            c.idle_bytes -= it->second.bytes;
            it = c.idle.erase(it);
        } else ++it;
    }
    if (prev >= 0) cudaSetDevice(prev);
    {
        PinCache& pc = pin_cache();
        std::lock_guard<std::mutex> lk2(pc.mu);
        for (auto& kv : pc.idle) { cudaFreeHost(kv.second); freed += (int64_t)kv.first; }
        pc.idle.clear();
    }
    cudaGetLastError();
    return freed;
}

int bvg_profile(const bvg_graph* g, int enable) {
    if (!g) return BVG_EINVAL;
    g->prof_on = enable != 0;
    return BVG_OK;
}

int bvg_profile_read(const bvg_graph* g, char* buf, int cap) {
    if (!g || !buf || cap < 2) return BVG_EINVAL;
    DeviceGuard dg(g->device);
    std::vector<ProfSpan*> spans;
    { std::lock_guard<std::mutex> lk(g->mu); spans.swap(g->prof_spans); }
    std::map<std::string, std::pair<int64_t, double>> acc;
    for (ProfSpan* p : spans) {
        float ms = 0;
        if (cudaEventSynchronize(p->e1) == cudaSuccess && cudaEventElapsedTime(&ms, p->e0, p->e1) == cudaSuccess) {
            auto& a = acc[p->name];
            a.first++; a.second += ms;
        }
        cudaEventDestroy(p->e0); cudaEventDestroy(p->e1);
        delete p;
    }
    cudaGetLastError();
    std::string out = "{";
    for (auto& kv : acc) {
        char tmp[256];
        snprintf(tmp, sizeof tmp, "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", out.size() > 1 ? ", " : "", kv.first.c_str(),
                 (long long)kv.second.first, kv.second.second);
        out += tmp;
    }
    out += "}";
    if ((int)out.size() + 1 > cap) return BVG_ENOMEM;
    memcpy(buf, out.c_str(), out.size() + 1);
    return BVG_OK;
}

}  // extern "C"

#include "bvg_labels_capi.cuh"
#include "bvg_ef_capi.cuh"
