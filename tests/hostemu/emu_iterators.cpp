// emu_iterators.cpp -- the device-side counterparts of the reference's MaskedIntIterator (copy-block mask over the
// parent's list: CopyRuns, bvg_scan.cuh) and MergedIntIterator (ascending union, equal heads once: copied_merge) on the
// host, fed with explicit block lists and lists so that tests can restate the reference's own unit tests
// (test/it/unimi/dsi/webgraph/MaskedIntIteratorTest.java, MergedIntIteratorTest.java).  Test infrastructure only.
#define BVG_HOST_EMULATION
#include "../../webgraph_b200/csrc/cuda/bvg_scan.cuh"
#include <vector>
using namespace bvg;

namespace {
struct BitWriter {
    std::vector<uint32_t> w;
    uint64_t n = 0;
    void bit(int b) {
        if ((n >> 5) >= w.size()) w.push_back(0);
        if (b) w[n >> 5] |= 0x80000000u >> (n & 31);
        n++;
    }
    void gamma(uint64_t x) {  // unary(msb(x + 1)) then the msb low bits of x + 1
        const uint64_t y = x + 1;
        const int m = 63 - __builtin_clzll(y);
        for (int i = 0; i < m; i++) bit(0);
        bit(1);
        for (int i = m - 1; i >= 0; i--) bit((int)((y >> i) & 1));
    }
};
}  // namespace

// blocks[0..bc): the block lengths as MaskedIntIterator takes them (first block may be 0, the others >= 1); written to a
// stream the way BVGraph writes them (first as it is, the others minus one, gamma: BVGraph.java:1062-1066, 2177-2180).
// variant 0: the masked sequence pulled one position at a time (CopyRuns::next)            -> out[0 .. *out_len)
//         1: copied_fold<4>                                                                   -> *fold (out untouched)
//         3: copied_merge: row = [copied slots | extras], d = copied + ne                    -> out[0 .. d), *fold
extern "C" int emu_masked(const int32_t* parent, int32_t dp, const int32_t* blocks, int32_t bc, const int32_t* extras, int32_t ne,
                          int32_t copied, int32_t x, int variant, int32_t* out, int32_t* out_len, unsigned long long* fold) {
    BitWriter bw;
    for (int32_t i = 0; i < bc; i++) bw.gamma((uint64_t)(blocks[i] - (i ? 1 : 0)));
    bw.w.resize(((bw.w.size() + STREAM_PAD_WORDS + 3) / 4) * 4 + 4, 0);
    std::vector<uint64_t> offsets(2, 0);
    offsets[1] = bw.n;
    ErrWord err{0, 0, 0};
    GraphDev g{};
    g.words = bw.w.data(); g.nwords = bw.w.size(); g.bit_base = 0; g.bit_end = bw.n;
    g.offsets = offsets.data(); g.node_lo = 0; g.node_hi = 1;
    g.c = Codec{ C_GAMMA, C_GAMMA, C_ZETA, C_UNARY, C_GAMMA, 3, 7, 4 };
    g.err = &err; g.hist = nullptr; g.hist_len = 0;
    // the parent's list in a buffer with guard values on both sides (nothing may be read outside the list)
    std::vector<int32_t> pbuf((size_t)dp + 16, -12345);
    int32_t* prow = pbuf.data() + 5;  // deliberately not 16-byte aligned
    for (int32_t i = 0; i < dp; i++) prow[i] = parent[i];
    int32_t slots[2 * COPY_RUNS];
    CopyRuns c;
    c.begin(g, 0, bc, dp, slots, 1, true);
    c.stage(g);
    if (variant == 0) {
        int32_t n = 0;
        uint32_t at;
        while (c.next(g, at)) out[n++] = prow[at];
        *out_len = n;
        return err.code;
    }
    if (variant == 1) { *fold = copied_fold<4>(g, c, x, prow); return err.code; }
    const int32_t d = copied + ne;
    std::vector<int32_t> row((size_t)d + 8, -777);
    for (int32_t i = 0; i < ne; i++) row[(size_t)copied + i] = extras[i];
    *fold = copied_merge(g, c, x, d, copied, row.data(), prow);
    for (int32_t i = 0; i < d; i++) out[i] = row[(size_t)i];
    *out_len = d;
    for (size_t i = (size_t)d; i < row.size(); i++) if (row[i] != -777) return -99;  // wrote past the row
    return err.code;
}
