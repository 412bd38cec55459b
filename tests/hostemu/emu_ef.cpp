// emu_ef.cpp -- the per-thread EFGraph walkers (bvg_ef.cuh: ef_outdegree_one, ef_decode_one) on the host.  The block-per-node
// kernel for heavy lists is covered on the device only (tests/test_efgraph.py, -m gpu).
#define BVG_HOST_EMULATION
#include <algorithm>
using std::min;
using std::max;
#include "../../webgraph_b200/csrc/cuda/bvg_ef.cuh"
#include <vector>
using namespace bvg;

// words: nwords long words in host order followed by >= 2 zero words.  out_off: n + 1, out: cap entries.
extern "C" int emu_ef_decode(const uint64_t* words, uint64_t nwords, const uint64_t* offsets, int32_t n, uint32_t upper_bound, int log2_quantum,
                             int64_t* out_off, int32_t* out, int64_t cap, unsigned long long* checksum) {
    EfDev g;
    g.w = words; g.nwords = nwords; g.offsets = offsets; g.n = n; g.upper_bound = upper_bound; g.log2_quantum = log2_quantum;
    ErrWord err{};
    std::vector<int32_t> deg((size_t)n + 1, 0);
    for (int64_t x = 0; x < n; x++) ef_outdegree_one(g, x, deg.data(), x, &err);
    if (err.code) return err.code;
    out_off[0] = 0;
    for (int64_t x = 0; x < n; x++) out_off[x + 1] = out_off[x] + deg[(size_t)x];
    if (out_off[n] > cap) return -6;
    unsigned long long acc = 0;
    for (int64_t x = 0; x < n; x++) if (deg[(size_t)x]) acc ^= ef_decode_one(g, x, out + out_off[x], &err);
    if (err.code) return err.code;
    // the warp path's per-lane pieces (one byte of upper bits per lane, ranks by a prefix over the 32 lanes), lane loop spelled out
    std::vector<int32_t> again((size_t)out_off[n] + 1, -1);
    unsigned long long acc2 = 0;
    for (int64_t x = 0; x < n; x++) {
        EfList e;
        if (!ef_list(g, x, e) || e.d != deg[(size_t)x]) return -101;
        if (e.d == 0) continue;
        const uint64_t ulen = (uint64_t)e.d + 1 + ((uint64_t)upper_bound >> e.l);
        int64_t carry = 0;
        for (uint64_t base = 0; base < ulen && carry < e.d; base += 256) {
            uint32_t bytes[32];
            int64_t rank = carry;
            for (int lane = 0; lane < 32; lane++) bytes[lane] = ef_upper_byte(g, e, ulen, base + (uint64_t)lane * 8);
            for (int lane = 0; lane < 32; lane++) {
                acc2 ^= ef_emit_byte(g, x, e, base + (uint64_t)lane * 8, bytes[lane], rank, again.data() + out_off[x]);
                rank += __builtin_popcount(bytes[lane]);
            }
            carry = rank;
        }
        if (carry < e.d) return -102;
    }
    if (acc2 != acc) return -103;
    for (int64_t j = 0; j < out_off[n]; j++) if (again[(size_t)j] != out[j]) return -104;
    *checksum = acc;
    return 0;
}

// EFGraph.store's element-parallel writer (efc_*), every element in turn: words (zeroed, nwords + 1 entries) and node_bits out.
extern "C" int64_t emu_ef_compress(const int64_t* off, const int32_t* succ, int32_t n, uint32_t upper_bound, int log2_quantum,
                                   unsigned long long* words, uint64_t cap_words, int64_t* node_bits) {
    EfcDev c;
    c.off = off; c.succ = succ; c.n = n; c.upper_bound = upper_bound; c.log2_quantum = log2_quantum;
    int bad = 0;
    std::vector<int32_t> sizes((size_t)n + 1, 0);
    for (int64_t x = 0; x < n; x++) efc_sizes_one(c, x, sizes.data(), &bad);
    if (bad) return -1;
    node_bits[0] = 0;
    for (int64_t x = 0; x < n; x++) node_bits[x + 1] = node_bits[x] + sizes[(size_t)x];
    if ((uint64_t)(node_bits[n] >> 6) + 2 > cap_words) return -6;
    for (int64_t x = n - 1; x >= 0; x--)   // any order must do: the device runs the elements concurrently
        for (int64_t k = off[x + 1] - off[x]; k >= 0; k--) efc_write_one(c, x, k, node_bits, words, &bad);
    return bad ? -1 : node_bits[n];
}
