// emu_scan.cpp -- the fused consume-only scan (k_scan_extras_lean / k_scan_merge_lean, bvg_scan.cuh) on the host,
// one record at a time: the same per-record walkers, the same four record kinds, parents' rows materialised and every
// other successor folded.  Returns (arcs, XOR checksum) for comparison with the oracle's scan and the parents' rows for
// comparison with the truth.  Test infrastructure only.
#define BVG_HOST_EMULATION
#include "../../webgraph_b200/csrc/cuda/bvg_scan.cuh"
#include <vector>
#include <cstdlib>
using namespace bvg;

template <int K>
static int scan_all(const GraphDev& g, int32_t n, const std::vector<int32_t>& outdeg, const std::vector<int32_t>& ref,
                    const std::vector<int32_t>& depth, int maxdepth, const std::vector<int64_t>& rowoff,
                    int32_t* rows, uint8_t* parent_flag, unsigned long long* out) {
    unsigned long long acc = 0, arcs = 0;
    std::vector<uint64_t> blocks_pos(n, 0);
    std::vector<int32_t> copied(n, 0), bcs(n, 0);
    for (int32_t x = 0; x < n; x++) if (ref[x]) parent_flag[x - ref[x]] = 1;
    for (int32_t x = 0; x < n; x++) {
        const int32_t d = outdeg[x];
        if (d == 0) continue;
        // what k_order_keys leaves in the schedule record
        BitBuf b = buffer_at(g, x);
        (void)Rd<true>::outdeg(b, g.c);
        int64_t bc = 0, cp = 0;
        if (g.c.window > 0) {
            const int32_t r = (int32_t)Rd<true>::ref(b, g.c);
            if (r > 0) {
                bc = (int64_t)Rd<true>::bcount(b, g.c);
                blocks_pos[x] = b.pos();
                int64_t total = 0;
                for (int64_t k = 0; k < bc; k++) {
                    const int64_t blk = (int64_t)Rd<true>::block(b, g.c) + (k ? 1 : 0);
                    total += blk;
                    if (!(k & 1)) cp += blk;
                }
                if (!(bc & 1)) cp += (int64_t)outdeg[x - r] - total;
            }
        }
        copied[x] = (int32_t)cp; bcs[x] = (int32_t)bc;
        const uint64_t epos = b.pos();
        const bool has_iv = d > cp && g.c.minlen != 0 && b.gamma() != 0;
        const bool store = parent_flag[x] != 0;
        const int32_t nout = d - (int32_t)cp;
        int32_t* row = rows + rowoff[x] + cp;
        unsigned long long f = 0;
        {
            alignas(16) unsigned char ring[RING_GROUPS * 16];
            ScanExtras<K, WinRing<1>> w;
            w.begin(g, x, nout, epos, true, ring_address(ring));
            if (has_iv) w.iv_fold(g); else w.iv_none(g);
            if (store) w.template resid<true>(g, row, true); else w.template resid<false>(g, row, false);
            if (store && has_iv) w.iv_merge(g, row);
            if (w.err) return w.err;
            f = w.finish();
        }
        acc ^= f;
        arcs += (unsigned long long)d;
    }
    for (int level = 1; level <= maxdepth; level++)
        for (int32_t x = 0; x < n; x++) if (depth[x] == level && outdeg[x]) {
            const int32_t px = x - ref[x];
            int32_t slots[2 * COPY_RUNS];
            CopyRuns c;
            c.begin(g, blocks_pos[x], bcs[x], outdeg[px], slots, 1, true);
            c.stage(g);
            if (parent_flag[x]) acc ^= copied_merge(g, c, x, outdeg[x], copied[x], rows + rowoff[x], rows + rowoff[px]);
            else acc ^= copied_fold<8>(g, c, x, rows + rowoff[px]);
        }
    out[0] = arcs; out[1] = acc;
    return 0;
}

// ExtrasWalk::header_rec leaves copied = 0 and the caller passes the row already advanced; mirror k_scan_extras_lean.
extern "C" int emu_scan(const uint8_t* graph, uint64_t nbytes, const uint64_t* offsets, int32_t n,
                        int window, int minlen, int zetak, int64_t* out_off, int32_t* rows, uint8_t* parent_flag,
                        unsigned long long* out /* arcs, xor */) {
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
    }
    for (int32_t x = 0; x <= n; x++) out_off[x] = rowoff[x];
    int rc;
    if (zetak == 3) rc = scan_all<3>(g, n, outdeg, ref, depth, maxdepth, rowoff, rows, parent_flag, out);
    else rc = scan_all<0>(g, n, outdeg, ref, depth, maxdepth, rowoff, rows, parent_flag, out);
    return rc ? rc : err.code;
}
