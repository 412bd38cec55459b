// bvg_labels_capi.cuh -- C ABI of the arc-label stream (include/bvgraph_b200.h, "Arc labels"); part of bvg_capi.cu (uses its
// allocator, launch macros and the device offsets decoder).  Kernels: bvg_labels.cuh.
//
// Replaces, for the three Label classes the reference ships, BitStreamArcLabelledImmutableGraph.load (:385-470: properties
// -> underlying graph + label prototype, .labels bytes, .labeloffsets gamma gaps), successors(x).label() (:225-262) and the
// sequential nodeIterator().labelArray() (ArcLabelledNodeIterator.java).  A bvg_labels belongs to the bvg_graph it was
// opened on (same device, same node window, same stream and call lock) and must be closed before it.

struct bvg_labels {
    const bvg_graph* g = nullptr;
    int kind = 0, width = 0;
    uint32_t* d_words = nullptr;
    uint64_t nwords = 0, bit_base = 0;
    uint64_t* d_off = nullptr;      // label offsets (file bit positions) of nodes node_lo .. node_hi
    uint64_t label_bits = 0;        // bits of the whole .labels stream (last label offset)
    uint64_t lo_bit = 0, hi_bit = 0;  // label offsets of node_lo / node_hi
    LabelsDev dev() const {
        LabelsDev L;
        L.w = d_words; L.maxw = nwords - 3; L.bit_base = bit_base; L.off = d_off; L.rowoff = g->d_rowoff;
        L.node_lo = g->node_lo; L.width = width;
        return L;
    }
};

static void labels_destroy(bvg_labels* l) {
    if (!l) return;
    DeviceGuard dg(l->g->device);
    cudaStreamSynchronize(l->g->stream);
    if (l->d_words) dev_free(l->d_words, l->g->stream);
    if (l->d_off) dev_free(l->d_off, l->g->stream);
    delete l;
}

// labelspec = <class>(<key>[,<width>]) (ObjectParser.fromSpec, :409-425).  The package prefix is not checked beyond the
// class's simple name: the three classes live in it.unimi.dsi.webgraph.labelling (and it.unimi.dsi.big.webgraph.labelling).
static int parse_labelspec(const std::string& spec, int* kind, int* width) {
    const size_t par = spec.find('(');
    std::string cls = spec.substr(0, par);
    while (!cls.empty() && isspace((unsigned char)cls.back())) cls.pop_back();
    const size_t dot = cls.rfind('.');
    if (dot != std::string::npos) cls = cls.substr(dot + 1);
    std::vector<std::string> args;
    if (par != std::string::npos) {
        const size_t close = spec.find(')', par);
        if (close == std::string::npos) return BVG_EFORMAT;
        std::string a = spec.substr(par + 1, close - par - 1), tok;
        for (size_t i = 0; i <= a.size(); i++) {
            if (i == a.size() || a[i] == ',') {
                size_t b = 0, e = tok.size();
                while (b < e && (isspace((unsigned char)tok[b]) || tok[b] == '"')) b++;
                while (e > b && (isspace((unsigned char)tok[e - 1]) || tok[e - 1] == '"')) e--;
                args.push_back(tok.substr(b, e - b));
                tok.clear();
            } else tok += a[i];
        }
    }
    if (cls == "GammaCodedIntLabel") { *kind = LAB_GAMMA; *width = 0; return args.empty() ? BVG_EFORMAT : BVG_OK; }
    if (cls == "FixedWidthIntLabel" || cls == "FixedWidthIntListLabel") {
        *kind = cls == "FixedWidthIntLabel" ? LAB_FIXED : LAB_FIXED_LIST;
        if (args.size() < 2) return BVG_EFORMAT;  // ArrayIndexOutOfBounds in the String... constructor
        char* end = nullptr;
        const long w = strtol(args[1].c_str(), &end, 10);
        if (end == args[1].c_str() || *end) return BVG_EFORMAT;  // NumberFormatException
        if (w < 0 || w > 31) return BVG_EINVAL;  // "Width out of range", FixedWidthIntLabel.java:41
        *width = (int)w;
        return BVG_OK;
    }
    return BVG_EUNSUPPORTED;  // a Label class this library has no kernel for
}

// Shared tail of the two open calls: the .labeloffsets stream decoded on the device, the window's slice kept, the window's
// stretch of the label stream (read by `fetch`) uploaded and byte-swapped.
template <class Fetch>
static int labels_build(bvg_labels* l, const uint8_t* offsets_stream, uint64_t offsets_bytes, uint64_t label_bytes, Fetch fetch) {
    const bvg_graph* g = l->g;
    cudaStream_t s = g->stream;
    const int64_t n = g->n_total;
    uint64_t* d_full = nullptr;
    int rc = device_decode_offsets(s, offsets_stream, offsets_bytes, C_GAMMA, n, &d_full);  // LabelOffsetsLongIterator, :330-358
    if (rc) { if (d_full) dev_free(d_full, s); return rc; }
    uint64_t o3[3];
    if (cudaMemcpyAsync(&o3[0], d_full + g->node_lo, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(&o3[1], d_full + g->node_hi, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(&o3[2], d_full + n, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); dev_free(d_full, s); return BVG_ECUDA; }
    if (o3[2] > label_bytes * 8 || o3[0] > o3[1] || o3[1] > o3[2]) { dev_free(d_full, s); return BVG_EIO; }  // offsets point past the stream
    l->label_bits = o3[2]; l->lo_bit = o3[0]; l->hi_bit = o3[1];
    const int64_t cnt = (int64_t)g->node_hi - g->node_lo;
    if (dev_alloc((void**)&l->d_off, ((size_t)cnt + 1) * 8, s) != cudaSuccess) { cudaGetLastError(); dev_free(d_full, s); return BVG_ENOMEM; }
    if (cudaMemcpyAsync(l->d_off, d_full + g->node_lo, ((size_t)cnt + 1) * 8, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
        cudaGetLastError(); dev_free(d_full, s); return BVG_ECUDA;
    }
    dev_free(d_full, s);
    const uint64_t byte_lo = (o3[0] >> 3) & ~(uint64_t)15, byte_hi = (o3[1] + 7) >> 3;
    l->bit_base = byte_lo * 8;
    const uint64_t nbytes = byte_hi - byte_lo;
    l->nwords = ((nbytes + 3) / 4 + STREAM_PAD_WORDS + 3) & ~(uint64_t)3;
    if (dev_alloc((void**)&l->d_words, (size_t)l->nwords * 4, s) != cudaSuccess) { cudaGetLastError(); return BVG_ENOMEM; }
    CK(cudaMemsetAsync(l->d_words, 0, (size_t)l->nwords * 4, s));
    if (nbytes) {
        const uint8_t* src = nullptr;
        std::vector<uint8_t> hold;
        rc = fetch(byte_lo, nbytes, &src, hold);
        if (rc) return rc;
        CK(cudaMemcpyAsync(l->d_words, src, (size_t)nbytes, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));  // `hold` dies with this scope
    }
    LAUNCH(k_bswap, grid_for((int64_t)l->nwords, 256), 256, 0, s, l->d_words, l->nwords);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return BVG_OK;
}

// The gamma labels of bits [base, end) (positions in the loaded words): speculate, fix until stable, scan of the counts.
// On return *sub / *cbase are valid until the Tmp objects die; *total = labels found.
static int labels_gamma_chains(const bvg_labels* l, cudaStream_t s, uint64_t base, uint64_t end, uint64_t sub_bits, int64_t nsub, Tmp<OffSub>& sa, Tmp<OffSub>& sb,
                               Tmp<int32_t>& counts, Tmp<int64_t>& cbase, Tmp<int>& changed, const OffSub** sub, int64_t* total) {
    CK(sa.alloc((size_t)nsub));
    CK(sb.alloc((size_t)nsub));
    CK(counts.alloc((size_t)nsub));
    CK(cbase.alloc((size_t)nsub + 1));
    CK(changed.alloc(1));
    LAUNCH_P(l->g, "k_lab_speculate", k_off_speculate, grid_for(nsub, 128), 128, 0, s, l->d_words, l->nwords, end, C_GAMMA, nsub, sa.p, base, sub_bits);
    OffSub *in = sa.p, *out = sb.p;
    for (int64_t pass = 0;; pass++) {
        CK(cudaMemsetAsync(changed.p, 0, sizeof(int), s));
        LAUNCH_P(l->g, "k_lab_fix", k_off_fix, grid_for(nsub, 128), 128, 0, s, l->d_words, l->nwords, end, C_GAMMA, nsub, in, out, changed.p, base, sub_bits);
        int ch = 0;
        CK(cudaMemcpyAsync(&ch, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        std::swap(in, out);
        if (!ch) break;
        if (pass > nsub + 2) return BVG_EIO;
    }
    LAUNCH(k_lab_sub_counts, grid_for(nsub, 256), 256, 0, s, in, nsub, counts.p);
    { const int rc = device_exclusive_scan(s, counts.p, nsub, cbase.p); if (rc) return rc; }
    CK(cudaMemcpyAsync(total, cbase.p + nsub, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *sub = in;
    return BVG_OK;
}

// Decodes (values != null or list_off != null) or folds (d_result != null) the labels of nodes [from, to).  All pointers
// are device pointers.  *nvalues receives the number of label values of the range.
static int labels_run(const bvg_labels* l, int32_t from, int32_t to, int64_t ra, int64_t rb, int64_t* d_list_off, int32_t* d_values, int64_t cap,
                      unsigned long long* d_result, int64_t* nvalues) {
    const bvg_graph* g = l->g;
    cudaStream_t s = g->stream;
    const LabelsDev L = l->dev();
    const int32_t rf = from - g->node_lo, rt = to - g->node_lo;
    const int64_t arcs = rb - ra;
    const bool fold = d_result != nullptr;
    if (l->kind != LAB_FIXED_LIST) {
        *nvalues = arcs;
        if (!fold && d_values && cap < arcs) return BVG_ENOMEM;
        if (d_list_off && !fold) LAUNCH(k_iota_i64, grid_for(arcs + 1, 256), 256, 0, s, d_list_off, arcs + 1);
        if (arcs == 0 || (!fold && !d_values)) { CK(cudaGetLastError()); return BVG_OK; }
    }
    if (l->kind == LAB_FIXED) {
        const unsigned grid = grid_for(arcs, LAB_FIXED_TILE);
        if (fold) LAUNCH_P(g, "k_lab_fixed", k_lab_fixed<true>, grid, LAB_FIXED_THREADS, 0, s, L, rf, rt, ra, rb, nullptr, d_result);
        else LAUNCH_P(g, "k_lab_fixed", k_lab_fixed<false>, grid, LAB_FIXED_THREADS, 0, s, L, rf, rt, ra, rb, d_values, nullptr);
    } else if (l->kind == LAB_GAMMA) {
        uint64_t oa, ob;
        CK(cudaMemcpyAsync(&oa, l->d_off + rf, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&ob, l->d_off + rt, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const uint64_t base = oa - l->bit_base, end = ob - l->bit_base;
        static const uint64_t sub_bits = (uint64_t)env_int("BVG_LAB_SUB_BITS", (int)LAB_SUB_BITS, 64, 1 << 24);
        const int64_t nsub = std::max<int64_t>(1, (int64_t)((end - base + sub_bits - 1) / sub_bits));
        Tmp<OffSub> sa(s), sb(s);
        Tmp<int32_t> counts(s);
        Tmp<int64_t> cbase(s);
        Tmp<int> changed(s);
        const OffSub* sub = nullptr;
        int64_t total = 0;
        const int rc = labels_gamma_chains(l, s, base, end, sub_bits, nsub, sa, sb, counts, cbase, changed, &sub, &total);
        if (rc) return rc;
        if (total != arcs) {  // the stretch between the two label offsets does not hold one label per arc
            std::lock_guard<std::mutex> lk(g->mu);
            g->err_node = from; g->err_bitpos = (int64_t)oa;
            return BVG_EFORMAT;
        }
        if (fold) LAUNCH_P(g, "k_lab_gamma_emit", k_lab_gamma_emit<true>, grid_for(nsub, 128), 128, 0, s, l->d_words, l->nwords, base, end, sub_bits, nsub, sub, cbase.p, ra, arcs, nullptr, d_result);
        else LAUNCH_P(g, "k_lab_gamma_emit", k_lab_gamma_emit<false>, grid_for(nsub, 128), 128, 0, s, l->d_words, l->nwords, base, end, sub_bits, nsub, sub, cbase.p, ra, arcs, d_values, nullptr);
    } else {
        const int64_t cnt = (int64_t)to - from;
        Tmp<int32_t> counts(s);
        Tmp<int64_t> vbase(s);
        CK(counts.alloc((size_t)cnt));
        CK(vbase.alloc((size_t)cnt + 1));
        if (cnt) LAUNCH_P(g, "k_lab_list_count", k_lab_list_count, grid_for(cnt, 128), 128, 0, s, L, rf, rt, counts.p, g->d_err);
        { const int rc = device_exclusive_scan(s, counts.p, cnt, vbase.p); if (rc) return rc; }
        CK(cudaMemcpyAsync(nvalues, vbase.p + cnt, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        { const int e = fetch_error(g); if (e) return e; }
        if (!fold && !d_values && !d_list_off) return BVG_OK;   // count only
        if (!fold && d_values && cap < *nvalues) return BVG_ENOMEM;
        if (cnt == 0) {
            if (!fold && d_list_off) CK(cudaMemsetAsync(d_list_off, 0, 8, s));
            return BVG_OK;
        }
        // without a values buffer the lists' offsets alone are wanted: the decode pass still walks, writing into a scratch
        Tmp<int32_t> scratch(s);
        int32_t* vals = d_values;
        if (!fold && !vals) { CK(scratch.alloc((size_t)*nvalues)); vals = scratch.p; }
        if (fold) LAUNCH_P(g, "k_lab_list_decode", k_lab_list_decode<true>, grid_for(cnt, 128), 128, 0, s, L, rf, rt, ra, rb, vbase.p, counts.p, nullptr, nullptr, d_result);
        else LAUNCH_P(g, "k_lab_list_decode", k_lab_list_decode<false>, grid_for(cnt, 128), 128, 0, s, L, rf, rt, ra, rb, vbase.p, counts.p, d_list_off, vals, nullptr);
        CK(cudaStreamSynchronize(s));  // counts / vbase / scratch die with this scope (stream-ordered frees would do; kept simple)
    }
    CK(cudaGetLastError());
    return BVG_OK;
}

extern "C" {

int bvg_labels_open_memory(const bvg_graph* g, const uint8_t* labels, uint64_t label_bytes, const uint8_t* label_offsets, uint64_t offsets_bytes,
                           int kind, int width, bvg_labels** out) {
    if (!g || !out || (!labels && label_bytes) || !label_offsets) return BVG_EINVAL;
    if (kind < LAB_GAMMA || kind > LAB_FIXED_LIST) return BVG_EUNSUPPORTED;
    if (kind != LAB_GAMMA && (width < 0 || width > 31)) return BVG_EINVAL;
    std::unique_lock<std::recursive_mutex> call_lock(g->call_mu);
    DeviceGuard dg(g->device);
    bvg_labels* l = new (std::nothrow) bvg_labels();
    if (!l) return BVG_ENOMEM;
    l->g = g; l->kind = kind; l->width = kind == LAB_GAMMA ? 0 : width;
    const int rc = labels_build(l, label_offsets, offsets_bytes, label_bytes,
                                [&](uint64_t from, uint64_t len, const uint8_t** src, std::vector<uint8_t>&) {
                                    if (from + len > label_bytes) return (int)BVG_EIO;
                                    *src = labels + from;
                                    return (int)BVG_OK;
                                });
    if (rc) { labels_destroy(l); return rc; }
    *out = l;
    return BVG_OK;
}

int bvg_labels_underlying(const char* basename, char* buf, int cap) {
    if (!basename || !buf || cap <= 0) return BVG_EINVAL;
    std::map<std::string, std::string> kv;
    if (!read_properties_file(std::string(basename) + ".properties", kv)) return BVG_EIO;
    if (kv.find("underlyinggraph") == kv.end()) return BVG_EIO;  // "does not contain an underlying graph basename", :391
    std::string name = kv["underlyinggraph"];
    if (name.empty() || name[0] != '/') {  // relative to the labelled graph's directory, :393-395
        const std::string b(basename);
        const size_t slash = b.rfind('/');
        if (slash != std::string::npos) name = b.substr(0, slash + 1) + name;
    }
    if ((int)name.size() + 1 > cap) return BVG_ENOMEM;
    memcpy(buf, name.c_str(), name.size() + 1);
    return BVG_OK;
}

int bvg_labels_open(const bvg_graph* g, const char* basename, bvg_labels** out) {
    if (!g || !basename || !out) return BVG_EINVAL;
    std::map<std::string, std::string> kv;
    if (!read_properties_file(std::string(basename) + ".properties", kv)) return BVG_EIO;
    if (kv.find("labelspec") == kv.end()) return BVG_EIO;  // "does not contain a label specification", :409
    int kind = 0, width = 0;
    int rc = parse_labelspec(kv["labelspec"], &kind, &width);
    if (rc) return rc;
    const std::string lpath = std::string(basename) + ".labels";
    const uint64_t label_bytes = file_size(lpath);
    if (label_bytes == ~0ull) return BVG_EIO;
    std::vector<uint8_t> offs;
    if (!slurp_file(std::string(basename) + ".labeloffsets", offs)) return BVG_EIO;
    std::unique_lock<std::recursive_mutex> call_lock(g->call_mu);
    DeviceGuard dg(g->device);
    bvg_labels* l = new (std::nothrow) bvg_labels();
    if (!l) return BVG_ENOMEM;
    l->g = g; l->kind = kind; l->width = width;
    rc = labels_build(l, offs.data(), offs.size(), label_bytes,
                      [&](uint64_t from, uint64_t len, const uint8_t** src, std::vector<uint8_t>& hold) {
                          if (!slurp_file(lpath, hold, from, len) || hold.size() != len) return (int)BVG_EIO;
                          *src = hold.data();
                          return (int)BVG_OK;
                      });
    if (rc) { labels_destroy(l); return rc; }
    *out = l;
    return BVG_OK;
}

void bvg_labels_close(bvg_labels* l) { labels_destroy(l); }

int bvg_labels_info(const bvg_labels* l, int* kind, int* width, int64_t* label_bits, int64_t* loaded_bytes) {
    if (!l) return BVG_EINVAL;
    if (kind) *kind = l->kind;
    if (width) *width = l->width;
    if (label_bits) *label_bits = (int64_t)l->label_bits;
    if (loaded_bytes) *loaded_bytes = (int64_t)(l->nwords * 4 + ((uint64_t)(l->g->node_hi - l->g->node_lo) + 1) * 8);
    return BVG_OK;
}

int bvg_labels_decode_range(const bvg_labels* l, int32_t from, int32_t to, int64_t* list_off, int32_t* values, int64_t cap, int on_device,
                            int64_t* nvalues) {
    if (!l) return BVG_EINVAL;
    const bvg_graph* g = l->g;
    std::unique_lock<std::recursive_mutex> call_lock(g->call_mu);
    int rc = range_check(g, from, to);
    if (rc) return rc;
    if (cap < 0) return BVG_EINVAL;
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    int64_t ra, rb, nv = 0;
    rc = fetch_rowoff(g, from, to, &ra, &rb);
    if (rc) return rc;
    const int64_t arcs = rb - ra;
    if (on_device) {
        rc = labels_run(l, from, to, ra, rb, list_off, values, cap, nullptr, &nv);
        if (rc) return rc;
        if (nvalues) *nvalues = nv;
        return BVG_OK;
    }
    if (!list_off && !values) {  // sizes only
        rc = labels_run(l, from, to, ra, rb, nullptr, nullptr, 0, nullptr, &nv);
        if (rc) return rc;
        if (nvalues) *nvalues = nv;
        return BVG_OK;
    }
    // host buffers: the number of values first (the lists' lengths are in the stream), then decode into device scratch
    rc = labels_run(l, from, to, ra, rb, nullptr, nullptr, 0, nullptr, &nv);
    if (rc) return rc;
    if (nvalues) *nvalues = nv;
    if (values && cap < nv) return BVG_ENOMEM;
    Tmp<int64_t> d_lo(s);
    Tmp<int32_t> d_v(s);
    if (list_off) CK(d_lo.alloc((size_t)arcs + 1));
    if (values) CK(d_v.alloc((size_t)nv));
    rc = labels_run(l, from, to, ra, rb, list_off ? d_lo.p : nullptr, values ? d_v.p : nullptr, nv, nullptr, &nv);
    if (rc) return rc;
    if (list_off) CK(cudaMemcpyAsync(list_off, d_lo.p, ((size_t)arcs + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (values && nv) CK(cudaMemcpyAsync(values, d_v.p, (size_t)nv * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int e = fetch_error(g);
    return e ? e : BVG_OK;
}

int bvg_labels_scan_range(const bvg_labels* l, int32_t from, int32_t to, int64_t* arcs, int64_t* nvalues, uint64_t* checksum) {
    if (!l) return BVG_EINVAL;
    const bvg_graph* g = l->g;
    std::unique_lock<std::recursive_mutex> call_lock(g->call_mu);
    int rc = range_check(g, from, to);
    if (rc) return rc;
    DeviceGuard dg(g->device);
    cudaStream_t s = g->stream;
    int64_t ra, rb, nv = 0;
    rc = fetch_rowoff(g, from, to, &ra, &rb);
    if (rc) return rc;
    Tmp<unsigned long long> res(s);
    CK(res.alloc(1));
    CK(cudaMemsetAsync(res.p, 0, 8, s));
    rc = labels_run(l, from, to, ra, rb, nullptr, nullptr, 0, res.p, &nv);
    if (rc) return rc;
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, res.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int e = fetch_error(g);
    if (e) return e;
    if (arcs) *arcs = rb - ra;
    if (nvalues) *nvalues = nv;
    if (checksum) *checksum = h;
    return BVG_OK;
}

}  // extern "C"
