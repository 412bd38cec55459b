// emu_tile.cpp -- the tile kernel (bvg_tile.cuh: planner + k_tile_scan) on the host: the same phases, a loop over the
// threads of the block for every phase, a byte buffer standing in for the block's shared memory (with guard bytes), the
// bulk copies as memcpy.  Returns (arcs, XOR checksum) of a consume-only scan of [fold_lo, fold_hi) for comparison with the
// oracle, the number of tiles and how many of them carry a halo.  Test infrastructure only.
#define BVG_HOST_EMULATION
#include <algorithm>
using std::min;
using std::max;
#include "../../webgraph_b200/csrc/cuda/bvg_tile.cuh"
#include <vector>
#include <cstdlib>
#include <cstdio>
using namespace bvg;

namespace {

struct EmuGraph {
    std::vector<uint32_t> words;
    std::vector<int32_t> outdeg, ref, depth;
    std::vector<uint8_t> is_parent;
    std::vector<LongMeta> meta;
    std::vector<int32_t> long_nodes, cb_cum, cb_ppos, iv_cum, iv_left, long_scr;
    std::vector<uint64_t> seg_pos;
    std::vector<int64_t> seg_val;
    ErrWord err{0, 0, 0};
    GraphDev g;
    LongIndex li;
};

void build(EmuGraph& E, const uint8_t* graph, uint64_t nbytes, const uint64_t* offsets, int32_t n, int window, int minlen, int zetak,
           int32_t long_d, int32_t seg, int32_t chunk) {
    E.words.assign((((nbytes + 3) / 4 + STREAM_PAD_WORDS + 3) / 4) * 4, 0);
    for (uint64_t i = 0; i < nbytes; i++) E.words[i >> 2] |= (uint32_t)graph[i] << (24 - 8 * (i & 3));
    E.outdeg.assign(n, 0); E.ref.assign(n, 0); E.depth.assign(n, 0); E.is_parent.assign(n + 1, 0);
    GraphDev& g = E.g;
    g.words = E.words.data(); g.nwords = E.words.size(); g.bit_base = 0; g.bit_end = offsets[n];
    g.offsets = offsets; g.node_lo = 0; g.node_hi = n;
    g.c = Codec{ C_GAMMA, C_GAMMA, C_ZETA, C_UNARY, C_GAMMA, zetak, window, minlen };
    g.outdeg = E.outdeg.data(); g.ref = E.ref.data(); g.depth = E.depth.data(); g.rowoff = nullptr; g.copied = nullptr; g.err = &E.err; g.hist = nullptr; g.hist_len = 0;
    for (int32_t x = 0; x < n; x++) {
        Bits b = cursor_at(g, x);
        const uint64_t d = Rd<true>::outdeg(b, g.c);
        int32_t r = 0;
        if (d > 0 && window > 0) r = (int32_t)Rd<true>::ref(b, g.c);
        E.outdeg[x] = (int32_t)d; E.ref[x] = r;
        E.depth[x] = r ? E.depth[x - r] + 1 : 0;
        if (r) E.is_parent[x - r] = 1;
    }
    // long index, as build_long_index lays it out
    for (int32_t x = 0; x < n; x++) if (E.outdeg[x] > long_d) {
        LongMeta m{};
        m.x = x; m.level = E.depth[x]; m.flags = E.is_parent[x] ? 1 : 0; m.rec_end = offsets[x + 1];
        long_walk<true>(g, m, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        E.meta.push_back(m);
        E.long_nodes.push_back(x);
    }
    int64_t cb = 0, iv = 0, sg = 0, scan = 0;
    for (auto& m : E.meta) {
        m.cb_off = cb; cb += m.ncb + 1;
        m.iv_off = iv; iv += m.ic + 1;
        m.seg_off = sg; sg += (m.rc + seg - 1) / seg;
        m.tmp_off = 0;
        m.scan_off = scan; if (m.flags & 1) scan += 3 * (int64_t)m.d;
    }
    E.cb_cum.assign(cb + 1, -1); E.cb_ppos.assign(cb + 1, -1); E.iv_cum.assign(iv + 1, -1); E.iv_left.assign(iv + 1, -1);
    E.seg_pos.assign(sg + 1, 0); E.seg_val.assign(sg + 1, 0); E.long_scr.assign(scan + 8, -99);
    for (auto& m : E.meta)
        long_walk<true>(g, m, 1, E.cb_cum.data() + m.cb_off, E.cb_ppos.data() + m.cb_off, E.iv_cum.data() + m.iv_off, E.iv_left.data() + m.iv_off,
                        E.seg_pos.data() + m.seg_off, E.seg_val.data() + m.seg_off, seg);
    LongIndex& li = E.li;
    li.meta = E.meta.data(); li.cb_cum = E.cb_cum.data(); li.cb_ppos = E.cb_ppos.data(); li.iv_cum = E.iv_cum.data(); li.iv_left = E.iv_left.data();
    li.seg_pos = E.seg_pos.data(); li.seg_val = E.seg_val.data(); li.seg = seg; li.chunk = chunk;
}

template <int K>
int run_tile(EmuGraph& E, const TileEntry& e, const TileArgs& a, int nt, unsigned long long& acc, long long& arcs) {
    const size_t GUARD = 256;
    std::vector<unsigned char> buf(a.smem_bytes + 2 * GUARD + 16, 0xA5);
    unsigned char* smem = buf.data() + GUARD;
    smem += (16 - ((uintptr_t)smem & 15)) & 15;
    Tile<K> T;
    T.carve(smem, a.smem_bytes, nt, e, E.g, a);
    typename Tile<K>::RunCopy rc[TILE_MAX_LONG + 1];
    const int n = T.layout(rc);
    for (int r = 0; r < n; r++)
        if (rc[r].bytes) memcpy(reinterpret_cast<unsigned char*>(T.sw) + rc[r].dst_byte, reinterpret_cast<const unsigned char*>(E.g.words) + rc[r].src_byte, rc[r].bytes);
    T.bind_stream();
    for (int t = 0; t < nt; t++) T.positions(t, nt);
    for (int t = 0; t < nt; t++) T.headers(t, nt);
    for (int t = 0; t < nt; t++) T.levels(t, nt);
    T.scan_buckets(0); T.scan_levels(0); T.scan_rows(0);
    for (int t = 0; t < nt; t++) T.scatter(t, nt);
    if (T.nlong) { T.long_count(0, 0); for (int t = 0; t < nt; t++) T.long_phase(0, t, nt, acc, arcs); }
    for (uint32_t it = 0; it < T.sh->nE; it++) T.extras_item((int32_t)T.ordE[it], acc, arcs);
    if (T.nlong) { T.long_count(1, 0); for (int t = 0; t < nt; t++) T.long_phase(1, t, nt, acc, arcs); }
    const int32_t maxlevel = T.sh->maxlevel;
    for (int32_t level = 1; level <= maxlevel; level++) {
        uint32_t la, lb;
        T.level_range(level, la, lb);
        for (uint32_t it = la; it < lb; it++) {
            int32_t i = (int32_t)T.ordM[it];
            if (level >= TILE_LEVELS && (int32_t)T.lvl[i] != level) i = -1;
            T.merge_item(i, (int)(it % (uint32_t)nt), nt, acc);
        }
        if (T.nlong) { T.long_count(2, level); for (int t = 0; t < nt; t++) T.long_phase(2, t, nt, acc, arcs); }
    }
    const int err = T.sh->err;
    for (size_t i = 0; i < GUARD; i++) if (buf[i] != 0xA5) return -100;
    for (size_t i = (size_t)(smem - buf.data()) + a.smem_bytes; i < buf.size(); i++) if (buf[i] != 0xA5) return -101;
    return err;
}

}  // namespace

// stats: [0] tiles, [1] tiles with a halo, [2] halo nodes, [3] long records, [4] max tile nodes
extern "C" int emu_tile_scan(const uint8_t* graph, uint64_t nbytes, const uint64_t* offsets, int32_t n, int window, int minlen, int zetak,
                             int32_t long_d, int32_t seg, int32_t chunk, uint32_t smem_bytes, int nt, int32_t fold_lo, int32_t fold_hi,
                             unsigned long long* out /* arcs, xor */, int64_t* stats) {
    EmuGraph E;
    build(E, graph, nbytes, offsets, n, window, minlen, zetak, long_d, seg, chunk);
    const uint32_t budget = tile_budget(smem_bytes, nt);
    std::vector<int32_t> cost(n + 1, 0);
    std::vector<uint8_t> clean(n + 2, 0);
    for (int32_t i = 0; i < n; i++) plan_node(E.g, i, n, E.is_parent.data(), long_d, budget, cost.data(), clean.data());
    if (n == 0) clean[0] = 1;
    std::vector<int64_t> cum(n + 1, 0);
    for (int32_t i = 0; i < n; i++) cum[i + 1] = cum[i] + cost[i];
    const int32_t nsb = (n + PLAN_SB - 1) / PLAN_SB;
    std::vector<TileEntry> tiles;
    for (int32_t sb = 0; sb < nsb; sb++) {
        const int32_t c = plan_superblock(E.g, n, cum.data(), clean.data(), E.long_nodes.data(), (int32_t)E.long_nodes.size(), budget, sb, 0, nullptr);
        if (c < 0) return -300;
        const size_t at = tiles.size();
        tiles.resize(at + c);
        plan_superblock(E.g, n, cum.data(), clean.data(), E.long_nodes.data(), (int32_t)E.long_nodes.size(), budget, sb, 0, tiles.data() + at);
    }
    TileArgs a{};
    a.tiles = tiles.data(); a.order = nullptr; a.first = 0; a.count = (int32_t)tiles.size();
    a.fold_lo = fold_lo; a.fold_hi = fold_hi; a.li = E.li; a.long_scr = E.long_scr.data(); a.result = nullptr; a.smem_bytes = smem_bytes;
    unsigned long long acc = 0;
    long long arcs = 0;
    int64_t st[5] = { (int64_t)tiles.size(), 0, 0, (int64_t)E.meta.size(), 0 };
    int32_t expect = 0;
    for (const TileEntry& e : tiles) {
        if (e.from != expect || e.hi <= e.from || e.lo > e.from) return -200;  // the tiles partition [0, n)
        expect = e.hi;
        if (e.lo < e.from) { st[1]++; st[2] += e.from - e.lo; }
        st[4] = std::max<int64_t>(st[4], e.hi - e.lo);
        if (e.hi <= fold_lo || e.from >= fold_hi) continue;
        int rc = zetak == 3 ? run_tile<3>(E, e, a, nt, acc, arcs) : run_tile<0>(E, e, a, nt, acc, arcs);
        if (rc) { if (getenv("EMU_TILE_DEBUG")) fprintf(stderr, "emu_tile: tile [%d %d %d) long [%d %d): rc %d, error %d at node %d bit %lld\n", e.lo, e.from, e.hi, e.long_lo, e.long_hi, rc, E.err.code, E.err.node, E.err.bitpos); return rc; }
    }
    if (expect != n) return -201;
    out[0] = (unsigned long long)arcs; out[1] = acc;
    if (stats) for (int i = 0; i < 5; i++) stats[i] = st[i];
    if (E.err.code && getenv("EMU_TILE_DEBUG")) fprintf(stderr, "emu_tile: error %d at node %d bit %lld\n", E.err.code, E.err.node, E.err.bitpos);
    return E.err.code;
}
