// bvg_ef_capi.cuh -- C ABI of the EFGraph decoder (include/bvgraph_b200.h, "EFGraph"); part of bvg_capi.cu.  Kernels: bvg_ef.cuh.
// Replaces EFGraph.loadInternal (EFGraph.java:709-790), outdegree (:1054-1060), successors (:1223-1225) and the sequential
// nodeIterator() an ImmutableGraph inherits, for graphs whose graphclass is it.unimi.dsi.webgraph.EFGraph.

struct bvg_efgraph {
    int device = 0;
    cudaStream_t stream = nullptr;
    int32_t n = 0;
    int64_t m = 0;
    int32_t upper_bound = 0;
    int log2_quantum = 0;
    uint64_t* d_words = nullptr;
    uint64_t nwords = 0, graph_bits = 0;
    uint64_t* d_offsets = nullptr;
    int64_t* d_rowoff = nullptr;   // n + 1
    ErrWord* d_err = nullptr;
    mutable std::recursive_mutex call_mu;
    mutable int32_t err_node = -1;
    mutable int64_t err_bitpos = -1;
    EfDev dev() const {
        EfDev g;
        g.w = d_words; g.nwords = nwords; g.offsets = d_offsets; g.n = n; g.upper_bound = (uint32_t)upper_bound; g.log2_quantum = log2_quantum;
        return g;
    }
};

// A pair of timing events that cannot leak on an early return.
struct EventPair {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t create() { cudaError_t r = cudaEventCreate(&e0); return r != cudaSuccess ? r : cudaEventCreate(&e1); }
    ~EventPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

static void ef_destroy(bvg_efgraph* g) {
    if (!g) return;
    DeviceGuard dg(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    void* ptrs[] = { g->d_words, g->d_offsets, g->d_rowoff, g->d_err };
    for (void* p : ptrs) if (p) dev_free(p, g->stream);
    if (g->stream) { cudaStreamSynchronize(g->stream); cudaStreamDestroy(g->stream); }
    cudaGetLastError();
    delete g;
}

static int ef_fetch_error(const bvg_efgraph* g) {
    ErrWord e{};
    if (cudaMemcpyAsync(&e, g->d_err, sizeof e, cudaMemcpyDeviceToHost, g->stream) != cudaSuccess ||
        cudaStreamSynchronize(g->stream) != cudaSuccess) { cudaGetLastError(); return BVG_ECUDA; }
    if (e.code) {
        g->err_node = e.node; g->err_bitpos = e.bitpos;
        cudaMemsetAsync(g->d_err, 0, sizeof(ErrWord), g->stream);
    }
    return e.code;
}

__global__ void k_bswap64(uint64_t* __restrict__ w, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t v = w[i];
        w[i] = ((uint64_t)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | (uint64_t)__byte_perm((uint32_t)(v >> 32), 0, 0x0123);
    }
}

extern "C" {

int bvg_ef_open_memory(const uint8_t* graph, uint64_t graph_bytes, const uint8_t* offsets_stream, uint64_t offsets_bytes, int32_t nodes, int64_t arcs,
                       int32_t upper_bound, int32_t quantum, int big_endian, int device, bvg_efgraph** out) {
    if (!out || nodes < 0 || arcs < 0 || (!graph && graph_bytes) || !offsets_stream) return BVG_EINVAL;
    if (quantum <= 0 || (quantum & (quantum - 1))) return BVG_EINVAL;   // "Illegal quantum (must be a power of 2)", :731
    if (upper_bound < 0) return BVG_EINVAL;
    int dev;
    int dl[1] = { device };
    int rc = pick_device(device >= 0 ? dl : nullptr, device >= 0 ? 1 : 0, &dev);
    if (rc) return rc;
    DeviceGuard dg(dev);
    bvg_efgraph* g = new (std::nothrow) bvg_efgraph();
    if (!g) return BVG_ENOMEM;
    g->device = dev; g->n = nodes; g->m = arcs; g->upper_bound = upper_bound;
    g->log2_quantum = 31 - __builtin_clz((unsigned)quantum);
    keep_pool_warm(dev);
    if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete g; return BVG_ECUDA; }
    cudaStream_t s = g->stream;
    auto fail = [&](int code) { ef_destroy(g); return code; };
    rc = device_decode_offsets(s, offsets_stream, offsets_bytes, C_DELTA, nodes, &g->d_offsets);   // OffsetsLongIterator, :641-672
    if (rc) return fail(rc);
    uint64_t last = 0;
    if (cudaMemcpyAsync(&last, g->d_offsets + nodes, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); return fail(BVG_ECUDA); }
    g->nwords = graph_bytes / 8;
    g->graph_bits = last;
    if (last > g->nwords * 64) return fail(BVG_EIO);   // offsets point past the stream
    if (dev_alloc((void**)&g->d_words, ((size_t)g->nwords + 2) * 8, s) != cudaSuccess) { cudaGetLastError(); return fail(BVG_ENOMEM); }
    if (dev_alloc((void**)&g->d_rowoff, ((size_t)nodes + 1) * 8, s) != cudaSuccess) { cudaGetLastError(); return fail(BVG_ENOMEM); }
    if (dev_alloc((void**)&g->d_err, sizeof(ErrWord), s) != cudaSuccess) { cudaGetLastError(); return fail(BVG_ENOMEM); }
    if (cudaMemsetAsync(g->d_err, 0, sizeof(ErrWord), s) != cudaSuccess || cudaMemsetAsync(g->d_words + g->nwords, 0, 16, s) != cudaSuccess) { cudaGetLastError(); return fail(BVG_ECUDA); }
    if (g->nwords && cudaMemcpyAsync(g->d_words, graph, (size_t)g->nwords * 8, cudaMemcpyHostToDevice, s) != cudaSuccess) { cudaGetLastError(); return fail(BVG_ECUDA); }
    if (big_endian && g->nwords) LAUNCH(k_bswap64, grid_for((int64_t)g->nwords, 256), 256, 0, s, g->d_words, g->nwords);
    rc = [&]() -> int {   // the temporaries of this scope must be gone before a failure destroys the stream
        Tmp<int32_t> outdeg(s);
        if (outdeg.alloc((size_t)std::max<int32_t>(nodes, 1)) != cudaSuccess) { cudaGetLastError(); return BVG_ENOMEM; }
        if (nodes) LAUNCH(k_ef_outdegrees, grid_for(nodes, 256), 256, 0, s, g->dev(), 0, nodes, outdeg.p, g->d_err);
        const int r = device_exclusive_scan(s, outdeg.p, nodes, g->d_rowoff);
        if (r) return r;
        int64_t total = 0;
        if (cudaMemcpyAsync(&total, g->d_rowoff + nodes, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); return BVG_ECUDA; }
        const int e = ef_fetch_error(g);
        if (e) return e;
        return total == arcs ? BVG_OK : BVG_EFORMAT;   // the outdegrees do not add up to the arcs property
    }();
    if (rc) return fail(rc);
    *out = g;
    return BVG_OK;
}

int bvg_ef_open(const char* basename, int device, bvg_efgraph** out) {
    if (!basename || !out) return BVG_EINVAL;
    std::map<std::string, std::string> kv;
    if (!read_properties_file(std::string(basename) + ".properties", kv)) return BVG_EIO;
    auto has = [&](const char* k) { return kv.find(k) != kv.end(); };
    if (!has("graphclass")) return BVG_EIO;
    const std::string gc = kv["graphclass"];
    if (gc != "it.unimi.dsi.webgraph.EFGraph" && gc != "it.unimi.dsi.big.webgraph.EFGraph") return BVG_EIO;   // IOException, :716-718
    if (!has("version") || atoi(kv["version"].c_str()) > 0) return BVG_EIO;                                   // :720-722
    if (!has("nodes") || !has("arcs") || !has("quantum") || !has("byteorder")) return BVG_EFORMAT;
    const long long nodes = atoll(kv["nodes"].c_str());
    if (nodes > 2147483647LL || nodes < 0) return BVG_EINVAL;                                                 // :724
    const long long arcs = atoll(kv["arcs"].c_str());
    const long long ub = has("upperbound") ? atoll(kv["upperbound"].c_str()) : nodes;
    const long long quantum = atoll(kv["quantum"].c_str());
    if (quantum <= 0 || quantum > (1LL << 30) || (quantum & (quantum - 1)) || ub < 0 || ub > 2147483647LL) return BVG_EINVAL;
    int big;
    if (kv["byteorder"] == "BIG_ENDIAN") big = 1;
    else if (kv["byteorder"] == "LITTLE_ENDIAN") big = 0;
    else return BVG_EINVAL;                                                                                   // "Unknown byte order", :736
    std::vector<uint8_t> graph, offs;
    if (!slurp_file(std::string(basename) + ".graph", graph)) return BVG_EIO;
    if (!slurp_file(std::string(basename) + ".offsets", offs)) return BVG_EIO;
    return bvg_ef_open_memory(graph.data(), graph.size(), offs.data(), offs.size(), (int32_t)nodes, arcs, (int32_t)ub, (int32_t)quantum, big, device, out);
}

void bvg_ef_close(bvg_efgraph* g) { ef_destroy(g); }

int bvg_ef_info(const bvg_efgraph* g, int32_t* nodes, int64_t* arcs, int32_t* upper_bound, int32_t* quantum, int64_t* graph_bits) {
    if (!g) return BVG_EINVAL;
    if (nodes) *nodes = g->n;
    if (arcs) *arcs = g->m;
    if (upper_bound) *upper_bound = g->upper_bound;
    if (quantum) *quantum = 1 << g->log2_quantum;
    if (graph_bits) *graph_bits = (int64_t)g->graph_bits;
    return BVG_OK;
}

// Shared body of decode / scan.  d_off (device, may be null) gets to - from + 1 relative row offsets, d_out (device, may be
// null: scan) the successors, d_result (device, may be null) (arcs, xor).
static int ef_run(const bvg_efgraph* g, int32_t from, int32_t to, int64_t* d_off, int32_t* d_out, unsigned long long* d_result) {
    cudaStream_t s = g->stream;
    const int64_t cnt = (int64_t)to - from;
    if (d_off) LAUNCH(k_rel_offsets, grid_for(cnt + 1, 256), 256, 0, s, g->d_rowoff + from, cnt, d_off);
    if (cnt == 0 || (!d_out && !d_result)) { CK(cudaGetLastError()); return BVG_OK; }
    Tmp<int32_t> heavy(s), nheavy(s);
    CK(heavy.alloc((size_t)std::max<int64_t>(1, g->m / EF_HEAVY + 1)));
    CK(nheavy.alloc(1));
    CK(cudaMemsetAsync(nheavy.p, 0, 4, s));
    LAUNCH(k_ef_decode, grid_for(cnt, EF_BLOCK), EF_BLOCK, 0, s, g->dev(), from, to, g->d_rowoff + from, d_out, heavy.p, nheavy.p, d_result, g->d_err, env_int("BVG_EF_SMALL", EF_SMALL, 0, EF_HEAVY));
    int32_t nh = 0;
    CK(cudaMemcpyAsync(&nh, nheavy.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (nh > 0) LAUNCH(k_ef_decode_heavy, (unsigned)nh * EF_SPLIT, EF_BLOCK, 0, s, g->dev(), from, heavy.p, g->d_rowoff + from, d_out, d_result, g->d_err);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));   // heavy / nheavy die with this scope
    return BVG_OK;
}

static int ef_range_check(const bvg_efgraph* g, int32_t from, int32_t to) {
    if (!g) return BVG_EINVAL;
    if (from < 0 || to < from || to > g->n) return BVG_EINVAL;
    return BVG_OK;
}

int bvg_ef_range_arcs(const bvg_efgraph* g, int32_t from, int32_t to, int64_t* arcs) {
    int rc = ef_range_check(g, from, to);
    if (rc) return rc;
    if (!arcs) return BVG_EINVAL;
    std::unique_lock<std::recursive_mutex> lk(g->call_mu);
    DeviceGuard dg(g->device);
    int64_t ab[2];
    CK(cudaMemcpyAsync(&ab[0], g->d_rowoff + from, 8, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaMemcpyAsync(&ab[1], g->d_rowoff + to, 8, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    *arcs = ab[1] - ab[0];
    return BVG_OK;
}

int bvg_ef_decode_range(const bvg_efgraph* g, int32_t from, int32_t to, int64_t* out_off, int32_t* out, int64_t cap, int on_device) {
    int rc = ef_range_check(g, from, to);
    if (rc) return rc;
    if (!out_off || cap < 0) return BVG_EINVAL;
    std::unique_lock<std::recursive_mutex> lk(g->call_mu);
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    int64_t arcs = 0;
    rc = bvg_ef_range_arcs(g, from, to, &arcs);
    if (rc) return rc;
    if (out && cap < arcs) return BVG_ENOMEM;
    if (on_device) {
        rc = ef_run(g, from, to, out_off, out, nullptr);
        if (rc) return rc;
        const int e = ef_fetch_error(g);
        return e ? e : BVG_OK;
    }
    const int64_t cnt = (int64_t)to - from;
    Tmp<int64_t> d_off(s);
    Tmp<int32_t> d_out(s);
    CK(d_off.alloc((size_t)cnt + 1));
    if (out) CK(d_out.alloc((size_t)std::max<int64_t>(arcs, 1)));
    rc = ef_run(g, from, to, d_off.p, out ? d_out.p : nullptr, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_off, d_off.p, ((size_t)cnt + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (out && arcs) CK(cudaMemcpyAsync(out, d_out.p, (size_t)arcs * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int e = ef_fetch_error(g);
    return e ? e : BVG_OK;
}

int bvg_ef_scan_range(const bvg_efgraph* g, int32_t from, int32_t to, int64_t* arcs, uint64_t* checksum) {
    int rc = ef_range_check(g, from, to);
    if (rc) return rc;
    std::unique_lock<std::recursive_mutex> lk(g->call_mu);
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    Tmp<unsigned long long> res(s);
    CK(res.alloc(2));
    CK(cudaMemsetAsync(res.p, 0, 16, s));
    rc = ef_run(g, from, to, nullptr, nullptr, res.p);
    if (rc) return rc;
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, res.p, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int e = ef_fetch_error(g);
    if (e) return e;
    if (arcs) *arcs = (int64_t)h[0];
    if (checksum) *checksum = h[1];
    return BVG_OK;
}

int bvg_ef_outdegree(const bvg_efgraph* g, int32_t x, int32_t* d) {
    if (!g || !d) return BVG_EINVAL;
    if (x < 0 || x >= g->n) return BVG_EINVAL;   // IllegalArgumentException as for every ImmutableGraph
    int64_t a = 0;
    const int rc = bvg_ef_range_arcs(g, x, x + 1, &a);
    if (rc) return rc;
    *d = (int32_t)a;
    return BVG_OK;
}

int bvg_ef_successors(const bvg_efgraph* g, int32_t x, int32_t* out, int32_t cap, int32_t* d) {
    if (!g || !d || cap < 0) return BVG_EINVAL;
    if (x < 0 || x >= g->n) return BVG_EINVAL;
    int64_t off[2] = { 0, 0 };
    const int rc = bvg_ef_decode_range(g, x, x + 1, off, out, cap, 0);
    if (rc) return rc;
    *d = (int32_t)off[1];
    return BVG_OK;
}

int bvg_ef_last_error_node(const bvg_efgraph* g, int32_t* node, int64_t* bitpos) {
    if (!g) return BVG_EINVAL;
    if (node) *node = g->err_node;
    if (bitpos) *bitpos = g->err_bitpos;
    return BVG_OK;
}

// EFGraph.store on the device (bvg_ef.cuh, second half).  off / succ: the CSR (host or device pointers); graph_out: host buffer
// for the long words in little-endian order (one trailing word as LongWordOutputBitStream.close() writes); node_bits: host
// buffer of n + 1 bit offsets (the caller writes .offsets' delta-coded gaps and .properties from them).
int bvg_ef_compress(const int64_t* off, const int32_t* succ, int32_t n, int32_t upper_bound, int log2_quantum, int on_device, int device,
                    uint8_t* graph_out, uint64_t graph_cap, uint64_t* graph_bytes, int64_t* node_bits, double* device_ms) {
    if (!off || n < 0 || log2_quantum < 0 || log2_quantum > 30 || !graph_bytes || !node_bits) return BVG_EINVAL;
    int dev;
    int dl[1] = { device };
    int rc = pick_device(device >= 0 ? dl : nullptr, device >= 0 ? 1 : 0, &dev);
    if (rc) return rc;
    DeviceGuard dg(dev);
    keep_pool_warm(dev);
    cudaStream_t s = nullptr;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    rc = [&]() -> int {
        Tmp<int64_t> d_off(s), d_bits(s);
        Tmp<int32_t> d_succ(s), d_sizes(s);
        Tmp<int> d_bad(s);
        Tmp<unsigned long long> d_words(s);
        int64_t m = 0;
        const int64_t* off_dev = off;
        const int32_t* succ_dev = succ;
        if (on_device) CK(cudaMemcpyAsync(&m, off + n, 8, cudaMemcpyDeviceToHost, s)); else m = off[n];
        CK(cudaStreamSynchronize(s));
        if (m < 0 || (m > 0 && !succ)) return BVG_EINVAL;
        if (!on_device) {
            CK(d_off.alloc((size_t)n + 1));
            CK(d_succ.alloc((size_t)std::max<int64_t>(m, 1)));
            CK(cudaMemcpyAsync(d_off.p, off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s));
            if (m) CK(cudaMemcpyAsync(d_succ.p, succ, (size_t)m * 4, cudaMemcpyHostToDevice, s));
            off_dev = d_off.p; succ_dev = d_succ.p;
        }
        EfcDev c;
        c.off = off_dev; c.succ = succ_dev; c.n = n; c.upper_bound = (uint32_t)(upper_bound > 0 ? upper_bound : n); c.log2_quantum = log2_quantum;
        CK(d_sizes.alloc((size_t)std::max<int32_t>(n, 1)));
        CK(d_bits.alloc((size_t)n + 1));
        CK(d_bad.alloc(1));
        CK(cudaMemsetAsync(d_bad.p, 0, 4, s));
        EventPair ev;
        CK(ev.create());
        CK(cudaEventRecord(ev.e0, s));
        if (n) LAUNCH(k_efc_sizes, grid_for(n, 256), 256, 0, s, c, d_sizes.p, d_bad.p);
        int r = device_exclusive_scan(s, d_sizes.p, n, d_bits.p);
        if (r) return r;
        int64_t total_bits = 0;
        int bad = 0;
        CK(cudaMemcpyAsync(&total_bits, d_bits.p + n, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (bad) return BVG_EINVAL;
        const uint64_t nwords = ((uint64_t)total_bits >> 6) + 1;
        *graph_bytes = nwords * 8;
        if (!graph_out || graph_cap < nwords * 8) return BVG_ENOMEM;   // *graph_bytes says how much is needed
        CK(d_words.alloc((size_t)nwords + 1));
        CK(cudaMemsetAsync(d_words.p, 0, ((size_t)nwords + 1) * 8, s));
        const int64_t elements = m + n;
        if (elements) LAUNCH(k_efc_write, grid_for(elements, EFC_TILE), EFC_THREADS, 0, s, c, d_bits.p, d_words.p, d_bad.p);
        CK(cudaEventRecord(ev.e1, s));
        CK(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(graph_out, d_words.p, (size_t)nwords * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(node_bits, d_bits.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev.e0, ev.e1);
        if (device_ms) *device_ms = ms;
        CK(cudaGetLastError());
        return bad ? BVG_EINVAL : BVG_OK;   // a list that is not strictly increasing or reaches the upper bound (Accumulator.add, :499-503)
    }();
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    cudaGetLastError();
    return rc;
}

// BVGraph.store on the device (bvg_compress.cuh), default codings.  off / succ: the CSR (host or device pointers).  graph_out
// (host) receives the bytes of .graph, node_bits (host, n + 1) the bit position of every node (the caller gamma-codes the gaps
// into .offsets).  range_nodes: nodes per independently compressed range (<= 0: 256).
int bvg_bv_compress(const int64_t* off, const int32_t* succ, int32_t n, int32_t window, int32_t maxref, int32_t minlen, int32_t zetak,
                    int32_t range_nodes, int on_device, int device, uint8_t* graph_out, uint64_t graph_cap, uint64_t* graph_bytes,
                    int64_t* node_bits, double* device_ms) {
    if (!off || n < 0 || window < 0 || minlen < 0 || zetak < 1 || !graph_bytes || !node_bits) return BVG_EINVAL;
    if (window > BVC_MAX_WINDOW) return BVG_EUNSUPPORTED;
    if (range_nodes <= 0) range_nodes = 256;
    int dev;
    int dl[1] = { device };
    int rc = pick_device(device >= 0 ? dl : nullptr, device >= 0 ? 1 : 0, &dev);
    if (rc) return rc;
    DeviceGuard dg(dev);
    keep_pool_warm(dev);
    cudaStream_t s = nullptr;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    rc = [&]() -> int {
        Tmp<int64_t> d_off(s), d_bits(s);
        Tmp<int32_t> d_succ(s), d_sizes(s);
        Tmp<int8_t> d_ref(s);
        Tmp<int> d_bad(s);
        Tmp<uint32_t> d_words(s);
        int64_t m = 0;
        const int64_t* off_dev = off;
        const int32_t* succ_dev = succ;
        if (on_device) CK(cudaMemcpyAsync(&m, off + n, 8, cudaMemcpyDeviceToHost, s)); else m = off[n];
        CK(cudaStreamSynchronize(s));
        if (m < 0 || (m > 0 && !succ)) return BVG_EINVAL;
        if (!on_device) {
            CK(d_off.alloc((size_t)n + 1));
            CK(d_succ.alloc((size_t)std::max<int64_t>(m, 1)));
            CK(cudaMemcpyAsync(d_off.p, off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s));
            if (m) CK(cudaMemcpyAsync(d_succ.p, succ, (size_t)m * 4, cudaMemcpyHostToDevice, s));
            off_dev = d_off.p; succ_dev = d_succ.p;
        }
        BvcDev g;
        g.off = off_dev; g.succ = succ_dev; g.n = n; g.c = BvcCodec{ window, maxref, minlen, zetak }; g.range_nodes = range_nodes;
        CK(d_sizes.alloc((size_t)std::max<int32_t>(n, 1)));
        CK(d_ref.alloc((size_t)std::max<int32_t>(n, 1)));
        CK(d_bits.alloc((size_t)n + 1));
        CK(d_bad.alloc(1));
        CK(cudaMemsetAsync(d_bad.p, 0, 4, s));
        EventPair ev;
        CK(ev.create());
        CK(cudaEventRecord(ev.e0, s));
        const int64_t nranges = ((int64_t)n + range_nodes - 1) / range_nodes;
        const int32_t size = window + 1;
        Tmp<long long> d_cost(s);
        CK(d_cost.alloc((size_t)std::max<int64_t>((int64_t)n * size, 1)));
        if (n) {
            LAUNCH(k_bvc_costs, grid_for((int64_t)n * size, 128), 128, 0, s, g, d_cost.p, d_bad.p);
            LAUNCH(k_bvc_pick, grid_for(nranges, 64), 64, 0, s, g, nranges, d_cost.p, d_ref.p, d_sizes.p, d_bad.p);
        }
        int r = device_exclusive_scan(s, d_sizes.p, n, d_bits.p);
        if (r) return r;
        int64_t total_bits = 0;
        int bad = 0;
        CK(cudaMemcpyAsync(&total_bits, d_bits.p + n, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (bad) return BVG_EINVAL;   // a list that is not strictly increasing (IllegalArgumentException, BVGraph.java:2201)
        const uint64_t nbytes = ((uint64_t)total_bits + 7) >> 3;
        *graph_bytes = nbytes;
        if (nbytes && (!graph_out || graph_cap < nbytes)) return BVG_ENOMEM;   // *graph_bytes says how much is needed
        const uint64_t nwords = nbytes / 4 + 2;
        CK(d_words.alloc((size_t)nwords));
        CK(cudaMemsetAsync(d_words.p, 0, (size_t)nwords * 4, s));
        if (n) LAUNCH(k_bvc_write, grid_for(n, 128), 128, 0, s, g, d_ref.p, d_bits.p, d_words.p);
        LAUNCH(k_bswap, grid_for((int64_t)nwords, 256), 256, 0, s, d_words.p, nwords);   // big-endian words -> bytes in stream order
        CK(cudaEventRecord(ev.e1, s));
        if (nbytes) CK(cudaMemcpyAsync(graph_out, d_words.p, (size_t)nbytes, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(node_bits, d_bits.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev.e0, ev.e1);
        if (device_ms) *device_ms = ms;
        CK(cudaGetLastError());
        return BVG_OK;
    }();
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    cudaGetLastError();
    return rc;
}

}  // extern "C"
