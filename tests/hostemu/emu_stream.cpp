// emu_stream.cpp -- the stream-position extras kernel (bvg_stream.cuh: k_stream_entries + k_stream_extras) on the host:
// entries for every chunk, then every lane run to completion (the warp vote of the kernel only decides WHEN a lane takes
// its next step, not what the step does), then the copied parts level by level with the walkers of bvg_scan.cuh exactly as
// emu_scan.cpp does.  Returns (arcs, XOR checksum) for comparison with the oracle and the stored rows for comparison with
// the truth.  Long records are switched off here (long_d = INT_MAX); their interplay is covered on the GPU.
#define BVG_HOST_EMULATION
#include <algorithm>
using std::min;
using std::max;
#include "../../webgraph_b200/csrc/cuda/bvg_stream.cuh"
#include <vector>
#include <cstdlib>
#include <cstdio>
using namespace bvg;

namespace {
struct FlatRows {
    int32_t* out;
    const int64_t* rowoff;
    int32_t* row(const GraphDev&, int32_t x) const { return out + rowoff[x]; }
    bool wanted(const GraphDev&, int32_t) const { return true; }
};
}

extern "C" int emu_stream_scan(const uint8_t* graph, uint64_t nbytes, const uint64_t* offsets, int32_t n, int window, int minlen, int zetak,
                               int32_t lo, int32_t hi, int64_t* out_off, int32_t* rows, uint8_t* parent_flag, unsigned long long* out /* arcs, xor */,
                               int64_t* stats /* chunks, steps */) {
    std::vector<uint32_t> words((((nbytes + 3) / 4 + STREAM_PAD_WORDS + 3) / 4) * 4, 0);
    for (uint64_t i = 0; i < nbytes; i++) words[i >> 2] |= (uint32_t)graph[i] << (24 - 8 * (i & 3));
    std::vector<int32_t> outdeg(n), ref(n), depth(n);
    std::vector<int64_t> rowoff(n + 1, 0);
    ErrWord err{0, 0, 0};
    GraphDev g;
    g.words = words.data(); g.nwords = words.size(); g.bit_base = 0; g.bit_end = offsets[n];
    g.offsets = offsets; g.node_lo = 0; g.node_hi = n;
    g.c = Codec{ C_GAMMA, C_GAMMA, C_ZETA, C_UNARY, C_GAMMA, zetak, window, minlen };
    g.outdeg = outdeg.data(); g.ref = ref.data(); g.depth = depth.data(); g.rowoff = rowoff.data(); g.copied = nullptr; g.err = &err; g.hist = nullptr; g.hist_len = 0;
    int maxdepth = 0;
    for (int32_t x = 0; x < n; x++) {
        Bits b = cursor_at(g, x);
        const uint64_t d = Rd<true>::outdeg(b, g.c);
        int32_t r = 0;
        if (d > 0 && window > 0) r = (int32_t)Rd<true>::ref(b, g.c);
        outdeg[x] = (int32_t)d; ref[x] = r;
        depth[x] = r ? depth[x - r] + 1 : 0;
        if (depth[x] > maxdepth) maxdepth = depth[x];
        rowoff[x + 1] = rowoff[x] + (int64_t)d;
        if (r) parent_flag[x - r] = 1;
    }
    for (int32_t x = 0; x <= n; x++) out_off[x] = rowoff[x];
    const uint64_t bit0 = 0;
    const int64_t nchunks = (int64_t)((offsets[n] + STREAM_CHUNK_BITS - 1) / STREAM_CHUNK_BITS);
    std::vector<StreamEntry> entries((size_t)nchunks + 1);
    LongIndex li{};
    for (int64_t c = 0; c <= nchunks; c++) {
        if (zetak == 3) stream_entry_one<3>(g, c, nchunks, bit0, 0, nullptr, 0, 0x7fffffff, li, entries.data());
        else stream_entry_one<0>(g, c, nchunks, bit0, 0, nullptr, 0, 0x7fffffff, li, entries.data());
    }
    // entries are code boundaries in stream order
    for (int64_t c = 0; c < nchunks; c++) {
        const uint64_t p0 = (uint64_t)c * STREAM_CHUNK_BITS + entries[c].dpos, p1 = (uint64_t)(c + 1) * STREAM_CHUNK_BITS + entries[c + 1].dpos;
        if (p1 < p0) return -300;
    }
    StreamArgs a{};
    a.entries = entries.data(); a.nchunks = nchunks; a.first_chunk = 0; a.count = nchunks; a.bit0 = bit0;
    a.lo = lo; a.hi = hi; a.from = lo; a.is_parent = parent_flag; a.long_nodes = nullptr; a.nlong = 0; a.long_d = 0x7fffffff;
    a.li = li; a.long_tmp = nullptr; a.result = nullptr; a.debug_nostore = 0;
    std::vector<int32_t> defer((size_t)nchunks + 1);
    unsigned int ndefer = 0;
    a.defer_list = defer.data(); a.defer_count = &ndefer;
    FlatRows rm{ rows, rowoff.data() };
    unsigned long long acc = 0;
    long long arcs = 0, steps = 0;
    for (int64_t c = 0; c < nchunks; c++) {
        if (zetak == 3) {
            StreamLane<3, FlatRows> L;
            L.init();
            L.open(g, a, rm, c);
            while (!L.done) { if (L.r.rem > 0) L.step_resid(g, a); else L.prologue(g, a, rm); steps++; if (steps > (long long)offsets[n] + 1000) return -301; }
            acc ^= L.acc; arcs += L.arcs;
        } else {
            StreamLane<0, FlatRows> L;
            L.init();
            L.open(g, a, rm, c);
            while (!L.done) { if (L.r.rem > 0) L.step_resid(g, a); else L.prologue(g, a, rm); steps++; if (steps > (long long)offsets[n] + 1000) return -301; }
            acc ^= L.acc; arcs += L.arcs;
        }
        if (err.code) { fprintf(stderr, "emu_stream: error %d at node %d bit %lld (chunk %lld)\n", err.code, err.node, err.bitpos, (long long)c); return err.code; }
    }
    for (unsigned int i = 0; i < ndefer; i++) {
        if (zetak == 3) { StreamLane<3, FlatRows> L; L.init(); L.fix_intervals(g, a, rm, defer[i]); }
        else { StreamLane<0, FlatRows> L; L.init(); L.fix_intervals(g, a, rm, defer[i]); }
    }
    if (stats) stats[1] = ndefer;
    // copied parts, level by level (k_scan_merge_lean)
    for (int level = 1; level <= maxdepth; level++)
        for (int32_t x = lo; x < hi; x++) if (depth[x] == level && outdeg[x]) {
            const int32_t px = x - ref[x];
            BitBuf b = buffer_at(g, x);
            (void)Rd<true>::outdeg(b, g.c);
            (void)Rd<true>::ref(b, g.c);
            const int64_t bc = (int64_t)Rd<true>::bcount(b, g.c);
            const uint64_t bpos = b.pos();
            int64_t total = 0, cp = 0;
            for (int64_t k = 0; k < bc; k++) { const int64_t blk = (int64_t)Rd<true>::block(b, g.c) + (k ? 1 : 0); total += blk; if (!(k & 1)) cp += blk; }
            if (!(bc & 1)) cp += (int64_t)outdeg[px] - total;
            int32_t slots[2 * COPY_RUNS];
            CopyRuns c;
            c.begin(g, bpos, (int32_t)bc, outdeg[px], slots, 1, true);
            c.stage(g);
            if (parent_flag[x]) acc ^= copied_merge(g, c, x, outdeg[x], (int32_t)cp, rows + rowoff[x], rows + rowoff[px]);
            else acc ^= copied_fold<8>(g, c, x, rows + rowoff[px]);
        }
    out[0] = (unsigned long long)arcs; out[1] = acc;
    if (stats) stats[0] = nchunks;
    return err.code;
}
