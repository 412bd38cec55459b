// emu.cpp -- runs the device decode steps (decode_extras / merge_copied, level-synchronous like k_extras / k_merge, and
// the per-query chain walk of k_random) on the host, single-threaded, over a graph given as byte buffers.
// Debugging aid: compile with -fsanitize=address,undefined to catch out-of-bounds row writes before spending GPU time.
#define BVG_HOST_EMULATION
#include "../../webgraph_b200/csrc/cuda/bvg_device.cuh"
#include <vector>
#include <cstdio>
using namespace bvg;

extern "C" int emu_decode(const uint8_t* graph, uint64_t nbytes, const uint64_t* offsets, int32_t n,
                          int window, int minlen, int zetak, int c_outdeg, int c_block, int c_resid, int c_ref, int c_bcount,
                          int def_codec, int64_t* out_off, int32_t* out, int64_t cap, int random_mode) {
    std::vector<uint32_t> words((((nbytes + 3) / 4 + 8 + 3) / 4) * 4, 0);
    for (uint64_t i = 0; i < nbytes; i++) words[i >> 2] |= (uint32_t)graph[i] << (24 - 8 * (i & 3));
    std::vector<int32_t> outdeg(n), ref(n), depth(n);
    std::vector<int64_t> rowoff(n + 1, 0);
    ErrWord err{0, 0, 0};
    GraphDev g;
    g.words = words.data(); g.nwords = words.size(); g.bit_base = 0; g.bit_end = offsets[n];
    g.offsets = offsets; g.node_lo = 0; g.node_hi = n;
    g.c = Codec{ c_outdeg, c_block, c_resid, c_ref, c_bcount, zetak, window, minlen };
    g.outdeg = outdeg.data(); g.ref = ref.data(); g.depth = depth.data(); g.rowoff = rowoff.data(); g.copied = nullptr; g.err = &err; g.hist = nullptr; g.hist_len = 0;
    int maxdepth = 0;
    for (int32_t x = 0; x < n; x++) {
        Bits b = cursor_at(g, x);
        uint64_t d = def_codec ? Rd<true>::outdeg(b, g.c) : Rd<false>::outdeg(b, g.c);
        int32_t r = 0;
        if (d > 0 && window > 0) r = (int32_t)(def_codec ? Rd<true>::ref(b, g.c) : Rd<false>::ref(b, g.c));
        outdeg[x] = (int32_t)d; ref[x] = r;
        depth[x] = r ? depth[x - r] + 1 : 0;
        if (depth[x] > maxdepth) maxdepth = depth[x];
        rowoff[x + 1] = rowoff[x] + (int64_t)d;
    }
    if (rowoff[n] > cap) return -6;
    for (int32_t x = 0; x <= n; x++) out_off[x] = rowoff[x];
    if (!random_mode) {
        for (int32_t x = 0; x < n; x++) if (outdeg[x]) {
            int64_t c = def_codec ? decode_extras<true>(g, x, out + rowoff[x]) : decode_extras<false>(g, x, out + rowoff[x]);
            if (c < 0) return (int)c;
        }
        for (int level = 1; level <= maxdepth; level++)
            for (int32_t x = 0; x < n; x++) if (depth[x] == level) {
                if (def_codec) merge_copied<true>(g, x, out + rowoff[x], out + rowoff[x - ref[x]]);
                else merge_copied<false>(g, x, out + rowoff[x], out + rowoff[x - ref[x]]);
            }
    } else {
        for (int32_t x = 0; x < n; x++) if (outdeg[x]) {
            int64_t need = 0;
            for (int32_t y = x; ref[y]; ) { y -= ref[y]; need += outdeg[y]; }
            std::vector<int32_t> scratch((size_t)need + 1);
            int32_t* cur = scratch.data();
            const int32_t* parent = nullptr;
            for (int level = depth[x]; level >= 0; level--) {
                int32_t y = x;
                for (int s = 0; s < level; s++) y -= ref[y];
                int32_t* row = level == 0 ? out + rowoff[x] : cur;
                int64_t c = def_codec ? decode_extras<true>(g, y, row) : decode_extras<false>(g, y, row);
                if (c < 0) return (int)c;
                if (ref[y]) { if (def_codec) merge_copied<true>(g, y, row, parent); else merge_copied<false>(g, y, row, parent); }
                parent = row;
                cur += outdeg[y];
            }
        }
    }
    return err.code;
}

// Fold-only path (stream_only): checksum of the copied successors of every node with a reference, against the same
// value computed from the materialised rows through merge_in_place's companion (next_a).
extern "C" int emu_stream_fold(const uint8_t* graph, uint64_t nbytes, const uint64_t* offsets, int32_t n,
                               int window, int minlen, int zetak, int def_codec, const int64_t* row_off, const int32_t* rows,
                               unsigned long long* out_flat, unsigned long long* out_ref) {
    std::vector<uint32_t> words((((nbytes + 3) / 4 + 8 + 3) / 4) * 4, 0);
    for (uint64_t i = 0; i < nbytes; i++) words[i >> 2] |= (uint32_t)graph[i] << (24 - 8 * (i & 3));
    std::vector<int32_t> outdeg(n), ref(n), depth(n);
    ErrWord err{0, 0, 0};
    GraphDev g;
    g.words = words.data(); g.nwords = words.size(); g.bit_base = 0; g.bit_end = offsets[n];
    g.offsets = offsets; g.node_lo = 0; g.node_hi = n;
    g.c = Codec{ C_GAMMA, C_GAMMA, C_ZETA, C_UNARY, C_GAMMA, zetak, window, minlen };
    g.outdeg = outdeg.data(); g.ref = ref.data(); g.depth = depth.data(); g.rowoff = row_off; g.copied = nullptr; g.err = &err; g.hist = nullptr; g.hist_len = 0;
    if (!def_codec) return -3;
    unsigned long long flat = 0, refv = 0;
    for (int32_t x = 0; x < n; x++) {
        Bits b = cursor_at(g, x);
        uint64_t d = Rd<true>::outdeg(b, g.c);
        int32_t r = 0;
        if (d > 0 && window > 0) r = (int32_t)Rd<true>::ref(b, g.c);
        outdeg[x] = (int32_t)d; ref[x] = r;
    }
    for (int32_t x = 0; x < n; x++) if (ref[x]) {
        const int32_t* parent = rows + row_off[x - ref[x]];
        MergeWalk<true> w;
        w.header(g, x, parent, true);
        w.stream_only(g, flat);
        MergeWalk<true> v;
        v.header(g, x, parent, true);
        const unsigned long long base = (unsigned long long)(uint32_t)x * BVG_MIX;
        for (int64_t a = v.next_a(g.c); a != BVG_INF; a = v.next_a(g.c)) refv ^= base + (unsigned long long)(uint32_t)a;
    }
    *out_flat = flat; *out_ref = refv;
    return err.code;
}
