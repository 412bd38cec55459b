// emu_boundaries.cpp -- record boundaries from the .graph stream alone (bvg_boundaries.cuh) on the host: the device
// functions of every pass run in a loop over the sub-ranges, the pass logic restates device_offsets_from_graph
// (bvg_capi.cu).  Sub-range and cap sizes are parameters so that tests can make wrong chains the rule.
#define BVG_HOST_EMULATION
#include <algorithm>
using std::min;
using std::max;
#include "../../webgraph_b200/csrc/cuda/bvg_boundaries.cuh"
#include <vector>
#include <cstdio>
#include <cstdlib>
using namespace bvg;

extern "C" int emu_boundaries(const uint8_t* graph, uint64_t nbytes, int64_t n,
                              int window, int minlen, int zetak, int c_outdeg, int c_block, int c_resid, int c_ref, int c_bcount,
                              int def_codec, uint64_t sub_bits, uint64_t cap, uint64_t* out, int* passes, int64_t* walks) {
    std::vector<uint32_t> words((((nbytes + 3) / 4 + 8 + 3) / 4) * 4, 0);
    for (uint64_t i = 0; i < nbytes; i++) words[i >> 2] |= (uint32_t)graph[i] << (24 - 8 * (i & 3));
    const Codec c{ c_outdeg, c_block, c_resid, c_ref, c_bcount, zetak, window, minlen };
    const uint64_t stream_bits = nbytes * 8;
    const int64_t nsub = std::max<int64_t>(1, (int64_t)((stream_bits + sub_bits - 1) / sub_bits));
    const int32_t W = window;
    const size_t hw = (size_t)nsub * (size_t)std::max(W, 1);
    std::vector<BndSub> a((size_t)nsub), b((size_t)nsub);
    std::vector<int32_t> he_a(hw, 0), he_b(hw, 0), hx_a(hw, 0), hx_b(hw, 0), ring(hw, 0), ok((size_t)nsub, 0);
    for (int64_t j = 0; j < nsub; j++) {
        BndSub s;
        s.lo = (uint64_t)j * sub_bits;
        s.hi = j + 1 == nsub ? stream_bits : (uint64_t)(j + 1) * sub_bits;
        s.entry = s.lo; s.exit = BND_UNKNOWN; s.count = 0; s.bad_pos = BND_UNKNOWN; s.bad_index = 0; s.walked = 0; s.capped = 0;
        a[(size_t)j] = s;
    }
    std::vector<BndMemo> memo(BND_MEMO_SLOTS, BndMemo{0, 0, 0, 0, 0});
    const int lean = getenv("EMU_BND_LEAN") && atoi(getenv("EMU_BND_LEAN")) != 0;  // BVG_BND_LEAN
    int64_t trusted = 0, nwalks = 0;
    int pass = 0;
    for (;; pass++) {
        for (int64_t j = 0; j < nsub; j++) {
            if (def_codec) bnd_pass_one<true>(j, words.data(), words.size(), stream_bits, c, a.data(), b.data(), he_a.data(), hx_a.data(), he_b.data(), hx_b.data(), ring.data(), std::min(pass, 1), trusted, cap, memo.data(), lean);
            else bnd_pass_one<false>(j, words.data(), words.size(), stream_bits, c, a.data(), b.data(), he_a.data(), hx_a.data(), he_b.data(), hx_b.data(), ring.data(), std::min(pass, 1), trusted, cap, memo.data(), lean);
            nwalks += b[(size_t)j].walked;
        }
        for (int64_t j = 0; j < nsub; j++) bnd_check_one(j, b.data(), he_b.data(), hx_b.data(), W, ok.data());
        a.swap(b); he_a.swap(he_b); hx_a.swap(hx_b);
        if (getenv("EMU_BND_TRACE")) {
            int64_t nw = 0, nbad = 0, unk = 0, firstw = -1, lastw = -1;
            for (int64_t j = 0; j < nsub; j++) { if (a[(size_t)j].walked) { nw++; if (firstw < 0) firstw = j; lastw = j; } nbad += !ok[(size_t)j]; unk += a[(size_t)j].exit == BND_UNKNOWN; }
            fprintf(stderr, "pass %d: walked %lld [%lld..%lld], not ok %lld, unknown exits %lld, trusted %lld\n", pass, (long long)nw, (long long)firstw, (long long)lastw, (long long)nbad, (long long)unk, (long long)trusted);
        }
        int64_t first_bad = nsub;
        for (int64_t j = 0; j < nsub; j++) if (!ok[(size_t)j]) { first_bad = j; break; }
        if (first_bad == nsub) break;
        trusted = first_bad;
        if (pass > nsub + 2) return -100;
    }
    *passes = pass + 1;
    *walks = nwalks;
    std::vector<int64_t> base((size_t)nsub);
    int64_t total = 0;
    for (int64_t j = 0; j < nsub; j++) {
        base[(size_t)j] = total;
        if (a[(size_t)j].bad_pos != BND_UNKNOWN && total + a[(size_t)j].bad_index < n) return -5;
        total += a[(size_t)j].count;
    }
    if (total < n) return -4;
    if (total == n) out[n] = a[(size_t)nsub - 1].exit;
    for (int64_t j = 0; j < nsub; j++) {
        if (def_codec) bnd_emit_one<true>(j, words.data(), words.size(), stream_bits, c, a.data(), he_a.data(), ring.data(), hx_b.data(), base.data(), n, out, memo.data(), lean);
        else bnd_emit_one<false>(j, words.data(), words.size(), stream_bits, c, a.data(), he_a.data(), ring.data(), hx_b.data(), base.data(), n, out, memo.data(), lean);
    }
    if (getenv("EMU_BND_TRACE")) { int used = 0; for (auto& m : memo) used += m.ready; fprintf(stderr, "memo slots used %d\n", used); }
    return 0;
}
