// emu_offsets.cpp -- the speculative parallel .offsets decoder (bvg_offsets.cuh) on the host, with tiny sub-ranges
// (96 bits) so that almost every sub-range starts in the middle of a code and the fix passes really have to work.
#define BVG_HOST_EMULATION
#define BVG_OFF_SUB_BITS 96
#include <algorithm>
using std::min;
using std::max;
#include "../../webgraph_b200/csrc/cuda/bvg_offsets.cuh"
#include <vector>
using namespace bvg;

extern "C" int emu_decode_offsets(const uint8_t* stream, uint64_t nbytes, int coding, int64_t n, uint64_t* out, int* passes) {
    std::vector<uint32_t> words((((nbytes + 3) / 4 + 8 + 3) / 4) * 4, 0);
    for (uint64_t i = 0; i < nbytes; i++) words[i >> 2] |= (uint32_t)stream[i] << (24 - 8 * (i & 3));
    const uint64_t total_bits = nbytes * 8;
    const int64_t nsub = (int64_t)((total_bits + OFF_SUB_BITS - 1) / OFF_SUB_BITS);
    std::vector<OffSub> a((size_t)nsub), b((size_t)nsub);
    for (int64_t j = 0; j < nsub; j++) off_speculate_one(j, words.data(), words.size(), total_bits, coding, a.data());
    int it = 0;
    for (;; it++) {
        int changed = 0;
        for (int64_t j = 0; j < nsub; j++) off_fix_one(j, words.data(), words.size(), total_bits, coding, a.data(), b.data(), &changed);
        a.swap(b);
        if (!changed) break;
        if (it > nsub + 2) return -100;
    }
    *passes = it + 1;
    std::vector<int64_t> cbase((size_t)nsub + 1, 0);
    std::vector<uint64_t> sbase((size_t)nsub + 1, 0);
    for (int64_t j = 0; j < nsub; j++) { cbase[j + 1] = cbase[j] + a[j].count; sbase[j + 1] = sbase[j] + a[j].sum; }
    if (cbase[nsub] < n + 1) return -4;
    for (int64_t j = 0; j < nsub; j++) off_emit_one(j, words.data(), words.size(), total_bits, coding, a.data(), cbase.data(), sbase.data(), n, out);
    return 0;
}
