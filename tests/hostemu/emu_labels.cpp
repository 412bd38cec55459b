// emu_labels.cpp -- the label-stream decoders (bvg_labels.cuh) on the host: the per-arc fixed-width reader over the tiles a
// launch would use, the gamma labels through the speculate / fix / emit passes with 96-bit sub-ranges, the list labels
// through the count and decode passes.  Mirrors labels_run (bvg_labels_capi.cuh).
#define BVG_HOST_EMULATION
#define BVG_OFF_SUB_BITS 96
#include <algorithm>
using std::min;
using std::max;
#include "../../webgraph_b200/csrc/cuda/bvg_labels.cuh"
#include <vector>
using namespace bvg;

// labels: the .labels bytes; off / rowoff: n + 1 entries each.  first_byte simulates a shard whose words start there.
// Returns 0, -5 (gamma stretch does not hold one label per arc), -6 (cap), or the error a list walk reported.
extern "C" int emu_labels(const uint8_t* labels, uint64_t nbytes, const uint64_t* off, const int64_t* rowoff, int32_t n, int kind, int width,
                          uint64_t first_byte, int32_t from, int32_t to, int64_t* list_off, int32_t* values, int64_t cap,
                          unsigned long long* checksum, int64_t* nvalues) {
    first_byte &= ~(uint64_t)15;
    std::vector<uint32_t> words((((nbytes - first_byte + 3) / 4 + 8 + 3) / 4) * 4, 0);
    for (uint64_t i = first_byte; i < nbytes; i++) words[(i - first_byte) >> 2] |= (uint32_t)labels[i] << (24 - 8 * ((i - first_byte) & 3));
    LabelsDev L;
    L.w = words.data(); L.maxw = words.size() - 3; L.bit_base = first_byte * 8; L.off = off; L.rowoff = rowoff; L.node_lo = 0; L.width = width;
    (void)n;
    const int64_t ra = rowoff[from], rb = rowoff[to], arcs = rb - ra;
    uint64_t acc = 0;
    if (kind != LAB_FIXED_LIST) {
        *nvalues = arcs;
        if (cap < arcs) return -6;
        for (int64_t j = 0; j <= arcs; j++) list_off[j] = j;
        if (arcs == 0) { *checksum = 0; return 0; }
    }
    if (kind == LAB_FIXED) {
        for (int64_t tile = ra; tile < rb; tile += LAB_FIXED_TILE) {
            const int64_t tile_end = std::min<int64_t>(tile + LAB_FIXED_TILE, rb);
            const int32_t nlo = lab_node_of(rowoff, from, to - 1, tile), nhi = lab_node_of(rowoff, from, to - 1, tile_end - 1);
            for (int64_t j0 = tile; j0 < tile_end; j0 += LAB_FIXED_ITEMS) {  // one thread's run
                uint32_t v[LAB_FIXED_ITEMS];
                const int c = lab_fixed_run(L, nlo, nhi, j0, tile_end, v);
                if (c != (int)std::min<int64_t>(LAB_FIXED_ITEMS, tile_end - j0)) return -101;
                for (int k = 0; k < c; k++) {
                    if (v[k] != lab_fixed_value(L, nlo, nhi, j0 + k)) return -102;  // the per-arc search and the walk must agree
                    values[j0 + k - ra] = (int32_t)v[k];
                    acc += lab_fold_int(j0 + k - ra, v[k]);
                }
            }
        }
    } else if (kind == LAB_GAMMA) {
        const uint64_t base = off[from] - L.bit_base, end = off[to] - L.bit_base;
        const uint64_t sub_bits = 80;  // not the offsets' pitch: the run-time pitch is what the label path uses
        const int64_t nsub = std::max<int64_t>(1, (int64_t)((end - base + sub_bits - 1) / sub_bits));
        std::vector<OffSub> a((size_t)nsub), b((size_t)nsub);
        for (int64_t j = 0; j < nsub; j++) off_speculate_one(j, words.data(), words.size(), end, C_GAMMA, a.data(), base, sub_bits);
        for (int it = 0;; it++) {
            int changed = 0;
            for (int64_t j = 0; j < nsub; j++) off_fix_one(j, words.data(), words.size(), end, C_GAMMA, a.data(), b.data(), &changed, base, sub_bits);
            a.swap(b);
            if (!changed) break;
            if (it > nsub + 2) return -100;
        }
        std::vector<int64_t> cbase((size_t)nsub + 1, 0);
        for (int64_t j = 0; j < nsub; j++) cbase[j + 1] = cbase[j] + a[j].count;
        if (cbase[nsub] != arcs) return -5;
        for (int64_t j = 0; j < nsub; j++) lab_gamma_emit_one(j, words.data(), words.size(), base, end, sub_bits, a.data(), cbase.data(), ra, arcs, values, false);
        for (int64_t j = 0; j < nsub; j++) acc += lab_gamma_emit_one(j, words.data(), words.size(), base, end, sub_bits, a.data(), cbase.data(), ra, arcs, nullptr, true);
    } else {
        const int64_t cnt = (int64_t)to - from;
        std::vector<int32_t> counts((size_t)cnt + 1, 0);
        std::vector<int64_t> vbase((size_t)cnt + 1, 0);
        ErrWord err{};
        for (int32_t x = from; x < to; x++) lab_list_count_one(L, x, from, counts.data(), &err);
        if (err.code) return err.code;
        for (int64_t i = 0; i < cnt; i++) vbase[i + 1] = vbase[i] + counts[i];
        *nvalues = vbase[cnt];
        if (cap < *nvalues) return -6;
        for (int32_t x = from; x < to; x++) lab_list_decode_one(L, x, from, to, ra, vbase.data(), counts.data(), list_off, values, false);
        list_off[arcs] = vbase[cnt];
        for (int32_t x = from; x < to; x++) acc += lab_list_decode_one(L, x, from, to, ra, vbase.data(), counts.data(), nullptr, nullptr, true);
    }
    *checksum = acc;
    return 0;
}
