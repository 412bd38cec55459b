// emu_long.cpp -- the split path for long records (bvg_long.cuh) on the host with tiny thresholds (every record with more
// than 2 successors is "long", sync point every 3 residuals, 4 outputs per merge-path chunk), so that the merge-path
// partition, the virtual copied/interval sequences and the sync points are exercised on thousands of records.
#define BVG_HOST_EMULATION
#define BVG_LONG_D 2
#define BVG_LONG_SEG 3
#define BVG_LONG_CHUNK 4
#define BVG_LSPEC_BITS 80
#include <algorithm>
using std::min;
using std::max;
#include "../../webgraph_b200/csrc/cuda/bvg_long.cuh"
#include <vector>
using namespace bvg;

struct FlatRows {  // RowMap stand-in: every row in one array
    int32_t* out;
    const int64_t* rowoff;
    int32_t* row(const GraphDev&, int32_t x) const { return out + rowoff[x]; }
    bool wanted(const GraphDev&, int32_t) const { return true; }
};

extern "C" int emu_decode_long(const uint8_t* graph, uint64_t nbytes, const uint64_t* offsets, int32_t n,
                               int window, int minlen, int zetak, int c_outdeg, int c_block, int c_resid, int c_ref, int c_bcount,
                               int def_codec, int64_t* out_off, int32_t* out, int64_t cap) {
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
        maxdepth = std::max(maxdepth, depth[x]);
        rowoff[x + 1] = rowoff[x] + (int64_t)d;
    }
    if (rowoff[n] > cap) return -6;
    for (int32_t x = 0; x <= n; x++) out_off[x] = rowoff[x];
    // long index
    std::vector<LongMeta> meta;
    for (int32_t x = 0; x < n; x++) if (outdeg[x] > LONG_D) {
        LongMeta m{};
        m.x = x; m.level = depth[x];
        if (def_codec) long_walk<true>(g, m, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        else long_walk<false>(g, m, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        meta.push_back(m);
    }
    int64_t cb = 0, iv = 0, seg = 0, tmp = 0;
    for (auto& m : meta) {
        m.cb_off = cb; cb += m.ncb + 1;
        m.iv_off = iv; iv += m.ic + 1;
        m.seg_off = seg; seg += (m.rc + LONG_SEG - 1) / LONG_SEG;
        m.tmp_off = tmp; tmp += 2 * (int64_t)m.d;
    }
    std::vector<int32_t> cb_cum(cb + 1, -1), cb_ppos(cb + 1, -1), iv_cum(iv + 1, -1), iv_left(iv + 1, -1), tmpbuf(tmp + 1, -99);
    std::vector<uint64_t> seg_pos(seg + 1);
    std::vector<int64_t> seg_val(seg + 1);
    for (auto& m : meta) {
        if (def_codec) long_walk<true>(g, m, 1, cb_cum.data() + m.cb_off, cb_ppos.data() + m.cb_off, iv_cum.data() + m.iv_off, iv_left.data() + m.iv_off, seg_pos.data() + m.seg_off, seg_val.data() + m.seg_off);
        else long_walk<false>(g, m, 1, cb_cum.data() + m.cb_off, cb_ppos.data() + m.cb_off, iv_cum.data() + m.iv_off, iv_left.data() + m.iv_off, seg_pos.data() + m.seg_off, seg_val.data() + m.seg_off);
    }
    {   // the same sync points through the speculative sub-range path (80-bit sub-ranges): must be identical
        std::vector<uint64_t> sp2(seg + 1, 0);
        std::vector<int64_t> sv2(seg + 1, 0);
        std::vector<SpecItem> items;
        for (size_t l = 0; l < meta.size(); l++) {
            const LongMeta& m = meta[l];
            if (m.rc <= 0) continue;
            const uint64_t end = offsets[m.x + 1];
            for (uint64_t lo = m.resid_pos; lo < end; lo += LSPEC_BITS) {
                SpecItem it{};
                it.lo = lo; it.hi = std::min<uint64_t>(lo + LSPEC_BITS, end); it.l = (int32_t)l; it.first = lo == m.resid_pos;
                items.push_back(it);
            }
        }
        for (auto& it : items) { if (def_codec) lspec_speculate_one<true>(g, it); else lspec_speculate_one<false>(g, it); }
        std::vector<SpecItem> tmp2(items.size());
        for (int pass = 0;; pass++) {
            int changed = 0;
            for (size_t j = 0; j < items.size(); j++) { if (def_codec) lspec_fix_one<true>(g, (int64_t)j, items.data(), tmp2.data(), &changed); else lspec_fix_one<false>(g, (int64_t)j, items.data(), tmp2.data(), &changed); }
            items.swap(tmp2);
            if (!changed) break;
            if (pass > (int)items.size() + 2) return -101;
        }
        std::vector<int64_t> v0(meta.size(), 0);
        int64_t cbase = 0, sbase = 0;
        for (size_t j = 0; j < items.size(); j++) {
            const SpecItem& it = items[j];
            if (it.first) { cbase = 0; sbase = 0; }
            if (def_codec) lspec_emit_one<true>(g, it, meta[it.l], cbase, sbase, v0.data(), v0.data(), sp2.data(), sv2.data());
            else lspec_emit_one<false>(g, it, meta[it.l], cbase, sbase, v0.data(), v0.data(), sp2.data(), sv2.data());
            cbase += it.count; sbase += it.sum;
            if ((j + 1 == items.size() || items[j + 1].first) && cbase != meta[it.l].rc) return -102;
        }
        for (int64_t k = 0; k < seg; k++) {
            if (sp2[k] != seg_pos[k]) return -103;
            if (sv2[k] != seg_val[k] && !(seg_val[k] == 0 && sp2[k] == seg_pos[k] && false)) {
                // seg_val of a record's first segment is unused (both paths leave their initial 0 there)
                bool first_of_record = false;
                for (auto& m : meta) if (m.seg_off == k) first_of_record = true;
                if (!first_of_record) return -104;
            }
        }
    }
    LongIndex li{ meta.data(), cb_cum.data(), cb_ppos.data(), iv_cum.data(), iv_left.data(), seg_pos.data(), seg_val.data(), LONG_SEG, LONG_CHUNK };
    FlatRows rm{ out, rowoff.data() };
    LongDst dst{ tmpbuf.data() };
    // level 0 work: short records sequentially, long records by segments / chunks
    for (int32_t x = 0; x < n; x++) if (outdeg[x] > 0 && outdeg[x] <= LONG_D) {
        int64_t c = def_codec ? decode_extras<true>(g, x, rm.row(g, x)) : decode_extras<false>(g, x, rm.row(g, x));
        if (c < 0) return (int)c;
    }
    for (auto& m : meta)
        for (int32_t s = 0; s * LONG_SEG < m.rc; s++) {
            if (def_codec) long_resid_segment<true>(g, m, li, s, dst.resid(m, rm.row(g, m.x)));
            else long_resid_segment<false>(g, m, li, s, dst.resid(m, rm.row(g, m.x)));
        }
    for (auto& m : meta) if (m.ic > 0) {
        IntervalSeq a{ li.iv_cum + m.iv_off, li.iv_left + m.iv_off, m.ic, m.ilen };
        const int32_t total = m.ilen + m.rc;
        const int32_t* left = a.left;
        for (int32_t q0 = 0; q0 < total; q0 += LONG_CHUNK)
            merge_chunk(a, [left](int32_t t, int32_t o) { return left[t] + o; }, dst.tmp + m.tmp_off, m.rc, q0,
                        std::min(LONG_CHUNK, total - q0), dst.extras(m, rm.row(g, m.x)));
    }
    for (int level = 1; level <= maxdepth; level++) {
        for (int32_t x = 0; x < n; x++) if (depth[x] == level && outdeg[x] <= LONG_D) {
            if (def_codec) merge_copied<true>(g, x, rm.row(g, x), rm.row(g, x - ref[x]));
            else merge_copied<false>(g, x, rm.row(g, x), rm.row(g, x - ref[x]));
        }
        for (auto& m : meta) if (m.level == level && m.copied > 0) {
            const int32_t* parent = rm.row(g, m.x - m.ref);
            CopiedSeq a{ li.cb_cum + m.cb_off, li.cb_ppos + m.cb_off, parent, m.ncb, m.copied };
            const int32_t* ppos = a.ppos;
            for (int32_t q0 = 0; q0 < m.d; q0 += LONG_CHUNK)
                merge_chunk(a, [ppos, parent](int32_t t, int32_t o) { return parent[ppos[t] + o]; }, dst.tmp + m.tmp_off + m.d,
                            m.d - m.copied, q0, std::min(LONG_CHUNK, m.d - q0), rm.row(g, m.x));
        }
    }
    return err.code;
}
