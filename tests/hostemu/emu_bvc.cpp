// emu_bvc.cpp -- BVGraph.store's device code (bvg_compress.cuh) on the host: the per-node choice (all candidates in turn) and the
// three-cursor writer, ranges of `range_nodes` nodes.  Output: MSB-first bytes of the stream and the bit position of every node.
#define BVG_HOST_EMULATION
#include <algorithm>
using std::min;
using std::max;
#include "../../webgraph_b200/csrc/cuda/bvg_compress.cuh"
#include <vector>
using namespace bvg;

extern "C" int64_t emu_bv_compress(const int64_t* off, const int32_t* succ, int32_t n, int32_t window, int32_t maxref, int32_t minlen, int32_t zetak,
                                   int32_t range_nodes, uint8_t* out, uint64_t cap, int64_t* node_bits, int8_t* refs) {
    BvcDev g;
    g.off = off; g.succ = succ; g.n = n; g.c = BvcCodec{ window, maxref, minlen, zetak }; g.range_nodes = range_nodes;
    std::vector<int32_t> refc((size_t)window + 1, 0), refc2((size_t)window + 1, 0);
    // the cost table of the device's phase 1a (every pair, any order), then phase 1b; checked against the all-in-one choice
    const int32_t size = window + 1;
    std::vector<long long> cost((size_t)std::max<int64_t>((int64_t)n * size, 1), LLONG_MAX);
    for (int64_t x = n - 1; x >= 0; x--) {
        const int64_t d = off[x + 1] - off[x];
        if (d < 0 || d > 0x7ffffffe) return -1;
        if (d == 0) continue;
        const int64_t lo = (x / range_nodes) * range_nodes;
        for (int32_t ref = 0; ref < size; ref++) {
            const int64_t y = x - ref;
            if (ref && (y < lo || off[y + 1] == off[y])) continue;
            bool b = false;
            cost[(size_t)(x * size + ref)] = (long long)bvc_cost(g, x, ref, b);
            if (b) return -1;
            BvcSections sec;
            if (bvc_cost_fast(g, x, ref, &sec) != cost[(size_t)(x * size + ref)]) return -300;   // the kernel's walker against the plain one
            {
                BvcEnc e;
                bvc_begin(e, nullptr, 0, 0, 0);
                const int64_t ra = ref ? off[y] : 0;
                bvc_walk(e, g.c, x, succ + off[x], (int32_t)d, succ + ra, ref ? (int32_t)(off[y + 1] - ra) : 0);
                if (sec.bc != e.bc || sec.ic != e.ic || sec.extras != e.extras || sec.block_bits != e.block_bits || sec.iv_bits != e.iv_bits) return -301;
            }
        }
    }
    node_bits[0] = 0;
    for (int64_t lo = 0; lo < n; lo += range_nodes) {
        std::fill(refc.begin(), refc.end(), 0);
        std::fill(refc2.begin(), refc2.end(), 0);
        const int64_t hi = std::min<int64_t>(n, lo + range_nodes);
        for (int64_t x = lo; x < hi; x++) {
            int32_t ref = 0, ref2 = 0;
            const int64_t bits = bvc_choose_one(g, x, lo, refc.data(), &ref);
            if (bits < 0) return -1;
            const int64_t bits2 = bvc_pick(g, x, cost.data(), refc2.data(), &ref2);
            if (bits2 != bits || ref2 != ref) return -200;
            refs[x] = (int8_t)ref;
            node_bits[x + 1] = node_bits[x] + bits;
        }
    }
    const uint64_t nbytes = ((uint64_t)node_bits[n] + 7) >> 3;
    if (nbytes > cap) return -6;
    std::vector<uint32_t> words((size_t)(nbytes / 4 + 2), 0);
    for (int64_t x = n - 1; x >= 0; x--) bvc_write_one(g, x, refs[x], (uint64_t)node_bits[x], words.data());   // any order
    for (uint64_t j = 0; j < nbytes; j++) out[j] = (uint8_t)(words[(size_t)(j >> 2)] >> (24 - 8 * (j & 3)));
    return node_bits[n];
}
