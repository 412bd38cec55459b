// bvg_tools.cpp -- host-side BVGraph compressor + synthetic power-law generator (libbvgraph_tools.so).
//
// Cold side of the format, see include/bvgraph_tools.h.  Behaviour follows the reference's writer
// (src/it/unimi/dsi/webgraph/BVGraph.java): reference selection :2313-2327, differential compression
// :2049-2219, intervalisation :1631-1654, offsets stream :2285,2369, multi-range concatenation :2498-2550,
// properties :2557-2636.  Structure is our own: cost evaluation is arithmetic (no dry-run bit stream),
// ranges compress into in-memory bit buffers that are spliced at bit granularity.
#include "../../../include/bvgraph_tools.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------------------------
// MSB-first bit buffer (the write half of dsiutils OutputBitStream, restated).
// ----------------------------------------------------------------------------------------------
struct BitBuf {
    std::vector<uint64_t> w;
    uint64_t nbits = 0;

    inline void put(uint64_t v, int n) {  // low n bits of v, 0 <= n <= 64
        if (n == 0) return;
        if (n < 64) v &= (~0ULL >> (64 - n));
        const uint64_t idx = nbits >> 6;
        const int off = (int)(nbits & 63);
        if (w.size() < idx + 2) w.resize(std::max<uint64_t>(idx + 2, w.size() * 2), 0);
        const int room = 64 - off;
        if (n <= room) w[idx] |= (room == n ? v : v << (room - n));
        else { w[idx] |= v >> (n - room); w[idx + 1] |= v << (64 - (n - room)); }
        nbits += (uint64_t)n;
    }
    void append(const BitBuf& o) {
        const uint64_t full = o.nbits >> 6;
        for (uint64_t i = 0; i < full; i++) put(o.w[i], 64);
        const int tail = (int)(o.nbits & 63);
        if (tail) put(o.w[full] >> (64 - tail), tail);
    }
    bool write_file(const std::string& path) const {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) return false;
        const uint64_t nbytes = (nbits + 7) >> 3;
        std::vector<uint8_t> buf((size_t)std::min<uint64_t>(nbytes, 1u << 24));
        uint64_t done = 0;
        bool ok = true;
        while (done < nbytes && ok) {
            const uint64_t chunk = std::min<uint64_t>(nbytes - done, buf.size());
            for (uint64_t i = 0; i < chunk; i++) {
                const uint64_t b = done + i;
                buf[i] = (uint8_t)(w[b >> 3] >> (56 - 8 * (b & 7)));
            }
            ok = fwrite(buf.data(), 1, (size_t)chunk, f) == chunk;
            done += chunk;
        }
        return fclose(f) == 0 && ok;
    }
};

inline int msb64(uint64_t x) { return 63 - __builtin_clzll(x); }

// ----------------------------------------------------------------------------------------------
// LSB-first long-word stream (the write half of EFGraph.LongWordOutputBitStream, EFGraph.java:298-418, restated): bit i of the
// stream is bit i % 64 of word i / 64.
// ----------------------------------------------------------------------------------------------
struct LsbBuf {
    std::vector<uint64_t> w;
    uint64_t nbits = 0;
    inline void put(uint64_t v, int n) {  // low n bits of v, 0 <= n <= 64
        if (n == 0) return;
        if (n < 64) v &= (~0ULL >> (64 - n));
        const uint64_t idx = nbits >> 6;
        const int off = (int)(nbits & 63);
        if (w.size() < idx + 2) w.resize(std::max<uint64_t>(idx + 2, w.size() * 2), 0);
        w[idx] |= v << off;
        if (off && n > 64 - off) w[idx + 1] |= v >> (64 - off);
        nbits += (uint64_t)n;
    }
    inline void unary(uint64_t zeros) {  // zeros, then a one
        while (zeros >= 64) { put(0, 64); zeros -= 64; }
        put(1ULL << zeros, (int)zeros + 1);
    }
    inline void gamma(uint64_t x) {  // writeGamma: unary(msb) as the word 1 << msb, then the msb low bits of x + 1 (:396-409)
        const uint64_t v = x + 1;
        const int msb = msb64(v);
        put(1ULL << msb, msb + 1);
        put(v ^ (1ULL << msb), msb);
    }
    void append(const LsbBuf& o) {
        const uint64_t full = o.nbits >> 6;
        for (uint64_t i = 0; i < full; i++) put(o.w[i], 64);
        const int tail = (int)(o.nbits & 63);
        if (tail) put(o.w[full], tail);
    }
};

inline int ef_lower_bits(uint64_t length, uint64_t ub) {  // EFGraph.lowerBits, :145-147
    if (length == 0) return 0;
    const uint64_t q = ub / length;
    return q == 0 ? 0 : msb64(q);
}
inline int ef_ceil_log2(uint64_t x) { return x <= 1 ? 0 : 64 - __builtin_clzll(x - 1); }  // Fast.ceilLog2
inline int ef_pointer_size(uint64_t length, uint64_t ub) { return ef_ceil_log2(length + (ub >> ef_lower_bits(length, ub))); }  // :156-158


// Code lengths and writers (definitions: SURVEY Appendix A.2).
inline int len_unary(uint64_t x) { return (int)x + 1; }
inline int len_gamma(uint64_t x) { return 2 * msb64(x + 1) + 1; }
inline int len_delta(uint64_t x) { const int m = msb64(x + 1); return len_gamma((uint64_t)m) + m; }
inline int len_zeta(uint64_t x, int k) {
    const uint64_t y = x + 1;
    const int h = msb64(y) / k;
    const uint64_t left = 1ULL << (h * k);
    return h + 1 + h * k + k - 1 + (y - left < left ? 0 : 1);
}
inline void put_unary(BitBuf& b, uint64_t x) {
    while (x >= 64) { b.put(0, 64); x -= 64; }
    b.put(1, (int)x + 1);
}
inline void put_gamma(BitBuf& b, uint64_t x) {
    const uint64_t y = x + 1;
    const int m = msb64(y);
    put_unary(b, (uint64_t)m);
    b.put(y, m);  // low m bits
}
inline void put_delta(BitBuf& b, uint64_t x) {
    const uint64_t y = x + 1;
    const int m = msb64(y);
    put_gamma(b, (uint64_t)m);
    b.put(y, m);
}
inline void put_zeta(BitBuf& b, uint64_t x, int k) {
    const uint64_t y = x + 1;
    const int h = msb64(y) / k;
    put_unary(b, (uint64_t)h);
    const uint64_t left = 1ULL << (h * k);
    if (y - left < left) b.put(y - left, h * k + k - 1);
    else b.put(y, h * k + k);
}

// Golomb with modulus b (OutputBitStream.writeGolomb: unary(x / b), then x % b in minimal binary: with l = msb(b),
// m = 2^(l+1) - b, remainders below m in l bits, the others as remainder + m in l + 1 bits); BVGraph uses b = zetaK
// (BVGraph.java:779, 809-813).  b == 0 writes nothing (and only encodes 0).
inline int len_golomb(uint64_t x, int b) {
    if (b <= 0) return 0;
    const int l = 31 - __builtin_clz((unsigned)b);
    const uint64_t m = (2ULL << l) - (uint64_t)b;
    const uint64_t q = x / (uint64_t)b;
    return (int)std::min<uint64_t>(q + 1, 1u << 30) + l + (x % (uint64_t)b < m ? 0 : 1);
}
inline void put_golomb(BitBuf& bb, uint64_t x, int b) {
    if (b <= 0) return;
    const int l = 31 - __builtin_clz((unsigned)b);
    const uint64_t m = (2ULL << l) - (uint64_t)b;
    put_unary(bb, x / (uint64_t)b);
    const uint64_t r = x % (uint64_t)b;
    if (r < m) { if (l) bb.put(r, l); }
    else bb.put(r + m, l + 1);
}
// Nibble code (OutputBitStream.writeNibble): the value in 3-bit groups, most significant first, each preceded by a
// stop flag that is 1 on the last group only; 0 is the single group 1000.
inline int len_nibble(uint64_t x) { return x == 0 ? 4 : 4 * (msb64(x) / 3 + 1); }
inline void put_nibble(BitBuf& b, uint64_t x) {
    int h = x == 0 ? 0 : msb64(x) / 3;
    do b.put((h == 0 ? 8u : 0u) | ((x >> (3 * h)) & 7u), 4); while (h-- != 0);
}

inline uint64_t int2nat(int64_t v) { return v >= 0 ? (uint64_t)v << 1 : (((uint64_t)(-v)) << 1) - 1; }

struct Codec {
    int32_t window, maxref, minlen, zetak;
    uint32_t flags;
    int outdegree = BVGT_GAMMA, block = BVGT_GAMMA, residual = BVGT_ZETA, reference = BVGT_UNARY,
        block_count = BVGT_GAMMA, offset = BVGT_GAMMA;
    bool ok = true;
    void set_flags(uint32_t f) {  // BVGraph.java:1317-1325
        flags = f;
        if (f & 0xF) outdegree = f & 0xF;
        if ((f >> 4) & 0xF) block = (f >> 4) & 0xF;
        if ((f >> 8) & 0xF) residual = (f >> 8) & 0xF;
        if ((f >> 12) & 0xF) reference = (f >> 12) & 0xF;
        if ((f >> 16) & 0xF) block_count = (f >> 16) & 0xF;
        if ((f >> 20) & 0xF) offset = (f >> 20) & 0xF;
        auto gd = [](int c) { return c == BVGT_GAMMA || c == BVGT_DELTA; };
        auto gdu = [](int c) { return c == BVGT_GAMMA || c == BVGT_DELTA || c == BVGT_UNARY; };
        ok = gd(outdegree) && gdu(block) && gdu(reference) && gdu(block_count) && gd(offset) &&
             (residual == BVGT_GAMMA || residual == BVGT_DELTA || residual == BVGT_ZETA ||
              (residual == BVGT_GOLOMB && zetak > 0) || residual == BVGT_NIBBLE);
    }
    inline int len(int coding, uint64_t x) const {
        switch (coding) {
            case BVGT_GAMMA: return len_gamma(x);
            case BVGT_DELTA: return len_delta(x);
            case BVGT_UNARY: return len_unary(x);
            case BVGT_GOLOMB: return len_golomb(x, zetak);
            case BVGT_NIBBLE: return len_nibble(x);
            default: return len_zeta(x, zetak);
        }
    }
    inline void put(BitBuf& b, int coding, uint64_t x) const {
        switch (coding) {
            case BVGT_GAMMA: put_gamma(b, x); break;
            case BVGT_DELTA: put_delta(b, x); break;
            case BVGT_UNARY: put_unary(b, x); break;
            case BVGT_GOLOMB: put_golomb(b, x, zetak); break;
            case BVGT_NIBBLE: put_nibble(b, x); break;
            default: put_zeta(b, x, zetak); break;
        }
    }
};

// ----------------------------------------------------------------------------------------------
// One compression range (the job of the reference's CompressionThread, BVGraph.java:2221-2386).
// ----------------------------------------------------------------------------------------------
struct RangeCompressor {
    const Codec& c;
    BitBuf graph, offs;     // offs holds one gap code per node of the range (the record length of each node)
    bvgt_store_stats st{};
    std::vector<std::vector<int32_t>> list;  // cyclic window of W+1 lists
    std::vector<int32_t> ref_count;
    std::vector<int32_t> blocks, extras, left, len, residuals;
    bool bad = false;

    explicit RangeCompressor(const Codec& codec) : c(codec), list((size_t)codec.window + 1), ref_count((size_t)codec.window + 1, 0) {}

    // Splits cur against the candidate: copy/skip blocks over ref_list, everything else to extras (:2066-2109).
    void split(const int32_t* cur, int32_t d, const int32_t* ref_list, int32_t ref_len) {
        blocks.clear();
        extras.clear();
        int32_t j = 0, k = 0, run = 0;
        bool copying = true;
        while (j < d && k < ref_len) {
            const int32_t a = cur[j], b = ref_list[k];
            if (copying) {
                if (a > b) { blocks.push_back(run); copying = false; run = 0; }
                else if (a < b) extras.push_back(cur[j++]);
                else { j++; k++; run++; }
            } else {
                if (a < b) extras.push_back(cur[j++]);
                else if (a > b) { k++; run++; }
                else { blocks.push_back(run); copying = true; run = 0; }
            }
        }
        if (copying && k < ref_len) blocks.push_back(run);
        while (j < d) extras.push_back(cur[j++]);
    }

    // Maximal runs of consecutive integers of length >= minlen become intervals (:1631-1654).
    void intervalize() {
        left.clear(); len.clear(); residuals.clear();
        const int32_t vl = (int32_t)extras.size();
        const int32_t* v = extras.data();
        for (int32_t i = 0; i < vl;) {
            int32_t j = i + 1;
            while (j < vl && v[j] == v[j - 1] + 1) j++;
            const int32_t run = j - i;
            if (run >= c.minlen) { left.push_back(v[i]); len.push_back(run); }
            else for (int32_t t = i; t < j; t++) residuals.push_back(v[t]);
            i = j;
        }
    }

    // Bits of (or, with out != nullptr, the actual) encoding of everything after the outdegree (:2115-2205).
    int64_t encode(int32_t x, int32_t ref, const int32_t* /*cur*/, int32_t d, BitBuf* out, bool for_real) {
        int64_t bits = 0;
        int copied = 0;
        if (c.window > 0) {
            const int t = c.len(c.reference, (uint64_t)ref);
            if (out) c.put(*out, c.reference, (uint64_t)ref);
            bits += t;
            if (for_real) st.bits_references += t;
        }
        if (ref != 0) {
            const int32_t bc = (int32_t)blocks.size();
            int64_t t = c.len(c.block_count, (uint64_t)bc);
            if (out) c.put(*out, c.block_count, (uint64_t)bc);
            for (int32_t i = 0; i < bc; i++) {
                const uint64_t v = (uint64_t)(i == 0 ? blocks[0] : blocks[i] - 1);
                t += c.len(c.block, v);
                if (out) c.put(*out, c.block, v);
            }
            bits += t;
            if (for_real) st.bits_blocks += t;
            copied = d - (int32_t)extras.size();
        }
        if (!extras.empty()) {
            const std::vector<int32_t>* res = &extras;
            if (c.minlen != 0) {
                intervalize();
                res = &residuals;
                const int32_t ic = (int32_t)left.size();
                int64_t t = len_gamma((uint64_t)ic);
                if (out) put_gamma(*out, (uint64_t)ic);
                int64_t prev = 0;
                for (int32_t i = 0; i < ic; i++) {
                    const uint64_t lv = i == 0 ? int2nat((int64_t)left[0] - x) : (uint64_t)((int64_t)left[i] - prev - 1);
                    const uint64_t nv = (uint64_t)(len[i] - c.minlen);
                    t += len_gamma(lv) + len_gamma(nv);
                    if (out) { put_gamma(*out, lv); put_gamma(*out, nv); }
                    prev = (int64_t)left[i] + len[i];
                    if (for_real) st.intervalised_arcs += len[i];
                }
                bits += t;
                if (for_real) st.bits_intervals += t;
            }
            const int32_t rc = (int32_t)res->size();
            if (rc) {
                int64_t t = 0;
                int64_t prev = (*res)[0];
                const uint64_t first = int2nat(prev - x);
                t += c.len(c.residual, first);
                if (out) c.put(*out, c.residual, first);
                for (int32_t i = 1; i < rc; i++) {
                    const int64_t r = (*res)[i];
                    if (r <= prev) { bad = true; return bits; }  // repeated/unsorted successor (:2201)
                    const uint64_t v = (uint64_t)(r - prev - 1);
                    t += c.len(c.residual, v);
                    if (out) c.put(*out, c.residual, v);
                    prev = r;
                }
                bits += t;
                if (for_real) { st.bits_residuals += t; st.residual_arcs += rc; }
            }
        }
        if (for_real) st.copied_arcs += copied;
        return bits;
    }

    void add(int32_t x, const int32_t* succ, int32_t d) {
        const int32_t size = c.window + 1;
        const int32_t cur_idx = x % size;
        const uint64_t start = graph.nbits;
        {   // outdegree (:2292)
            const int t = c.len(c.outdegree, (uint64_t)d);
            c.put(graph, c.outdegree, (uint64_t)d);
            st.bits_outdegrees += t;
        }
        list[cur_idx].assign(succ, succ + d);
        const int32_t* cur = list[cur_idx].data();
        if (d > 0) {
            for (int32_t i = 1; i < d; i++) if (cur[i] <= cur[i - 1]) { bad = true; return; }
            if (cur[0] < 0) { bad = true; return; }
            const int64_t maxref = c.maxref < 0 ? INT64_MAX : c.maxref;
            int64_t best_cost = INT64_MAX;
            int32_t best_ref = -1, best_cand = -1;
            ref_count[cur_idx] = -1;
            for (int32_t ref = 0; ref < size; ref++) {  // :2313-2323
                const int32_t cand = (int32_t)(((int64_t)x - ref + size) % size);
                if (ref_count[cand] < maxref && !list[cand].empty()) {
                    split(cur, d, list[cand].data(), ref == 0 ? 0 : (int32_t)list[cand].size());
                    const int64_t cost = encode(x, ref, cur, d, nullptr, false);
                    if (bad) return;
                    if (cost < best_cost) { best_cost = cost; best_ref = ref; best_cand = cand; }
                }
            }
            ref_count[cur_idx] = ref_count[best_cand] + 1;  // :2326
            split(cur, d, list[best_cand].data(), best_ref == 0 ? 0 : (int32_t)list[best_cand].size());
            encode(x, best_ref, cur, d, &graph, true);
            st.tot_ref += ref_count[cur_idx];
            st.tot_dist += best_ref;
            st.max_ref_chain = std::max(st.max_ref_chain, ref_count[cur_idx]);
            st.max_outdegree = std::max(st.max_outdegree, d);
            const uint64_t base = (uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ULL;
            for (int32_t i = 0; i < d; i++) {
                st.xor_checksum ^= base + (uint64_t)(uint32_t)cur[i];
                st.sum_successors += (uint64_t)(uint32_t)cur[i];
            }
        }
        st.nodes++;
        st.arcs += d;
        c.put(offs, c.offset, graph.nbits - start);  // gap written when the NEXT node starts (:2285) / at the end (:2369)
    }
};

void merge_stats(bvgt_store_stats& a, const bvgt_store_stats& b) {
    a.nodes += b.nodes; a.arcs += b.arcs;
    a.bits_outdegrees += b.bits_outdegrees; a.bits_references += b.bits_references; a.bits_blocks += b.bits_blocks;
    a.bits_intervals += b.bits_intervals; a.bits_residuals += b.bits_residuals;
    a.copied_arcs += b.copied_arcs; a.intervalised_arcs += b.intervalised_arcs; a.residual_arcs += b.residual_arcs;
    a.tot_ref += b.tot_ref; a.tot_dist += b.tot_dist;
    a.max_outdegree = std::max(a.max_outdegree, b.max_outdegree);
    a.max_ref_chain = std::max(a.max_ref_chain, b.max_ref_chain);
    a.xor_checksum ^= b.xor_checksum; a.sum_successors += b.sum_successors;
}

const char* coding_name(int c) {
    static const char* names[] = { "", "DELTA", "GAMMA", "GOLOMB", "SKEWED_GOLOMB", "UNARY", "ZETA", "NIBBLE" };
    return names[c & 7];
}

std::string flags_string(uint32_t f) {  // BVGraph.java:1332-1344
    static const char* slot[] = { "OUTDEGREES_", "BLOCKS_", "RESIDUALS_", "REFERENCES_", "BLOCK_COUNT_", "OFFSETS_" };
    std::string s;
    for (int i = 0; i < 6; i++) {
        const int c = (f >> (4 * i)) & 0xF;
        if (!c) continue;
        if (!s.empty()) s += " | ";
        s += slot[i];
        s += coding_name(c);
    }
    return s;
}

// Splices the per-range buffers and writes the three files.
int finish(const std::string& basename, const Codec& c, std::vector<RangeCompressor*>& parts, bvgt_store_stats* stats) {
    bvgt_store_stats st{};
    BitBuf graph, offs;
    c.put(offs, c.offset, 0);  // the leading 0 (first node starts at bit 0); per-range zeros are dropped (:2518-2520)
    for (RangeCompressor* p : parts) {
        if (p->bad) return -1;
        if (parts.size() == 1) { graph = std::move(p->graph); offs.append(p->offs); }
        else { graph.append(p->graph); offs.append(p->offs); p->graph = BitBuf(); }
        merge_stats(st, p->st);
    }
    st.graph_bits = (int64_t)graph.nbits;
    st.offsets_bits = (int64_t)offs.nbits;
    if (!graph.write_file(basename + ".graph")) return -4;
    if (!offs.write_file(basename + ".offsets")) return -4;
    FILE* f = fopen((basename + ".properties").c_str(), "w");
    if (!f) return -4;
    fprintf(f, "#BVGraph properties\n#written by webgraph_b200 bvg_tools\n");
    fprintf(f, "graphclass=it.unimi.dsi.webgraph.BVGraph\nversion=0\n");
    fprintf(f, "nodes=%lld\narcs=%lld\n", (long long)st.nodes, (long long)st.arcs);
    fprintf(f, "windowsize=%d\nmaxrefcount=%d\nminintervallength=%d\n", c.window, c.maxref < 0 ? 2147483647 : c.maxref, c.minlen);
    if (c.residual == BVGT_ZETA) fprintf(f, "zetak=%d\n", c.zetak);
    fprintf(f, "compressionflags=%s\n", flags_string(c.flags).c_str());
    const double arcs = st.arcs ? (double)st.arcs : 1.0, nodes = st.nodes ? (double)st.nodes : 1.0;
    fprintf(f, "bitsperlink=%.3f\nbitspernode=%.3f\navgref=%.3f\navgdist=%.3f\n", st.graph_bits / arcs, st.graph_bits / nodes,
            st.tot_ref / nodes, st.tot_dist / nodes);
    fprintf(f, "copiedarcs=%lld\nintervalisedarcs=%lld\nresidualarcs=%lld\n", (long long)st.copied_arcs,
            (long long)st.intervalised_arcs, (long long)st.residual_arcs);
    fprintf(f, "bitsforoutdegrees=%lld\nbitsforreferences=%lld\nbitsforblocks=%lld\nbitsforintervals=%lld\nbitsforresiduals=%lld\n",
            (long long)st.bits_outdegrees, (long long)st.bits_references, (long long)st.bits_blocks,
            (long long)st.bits_intervals, (long long)st.bits_residuals);
    fclose(f);
    if (stats) *stats = st;
    return 0;
}

// ----------------------------------------------------------------------------------------------
// Synthetic generator (block-local copy-model power-law graph).
// ----------------------------------------------------------------------------------------------
inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t s;
    inline uint64_t next() { s += 0x9E3779B97F4A7C15ULL; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
    inline double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

struct Generator {
    bvgt_gen_params p;
    double scale = 1.0;

    // un-normalised outdegree weight of node x: genzipf's (n/r)^s for a hashed rank r, times u^2 so that small
    // and zero degrees exist
    inline double weight(int32_t x) const {
        const uint64_t h = mix64(p.seed ^ (0xA24BAED4963EE407ULL * (uint64_t)(x + 1)));
        const double r = 1.0 + (double)(h % (uint64_t)p.n);
        const double u = (double)(mix64(h) >> 11) * (1.0 / 9007199254740992.0);
        return std::exp(p.zipf_s * std::log((double)p.n / r)) * u * u;
    }
    inline int32_t degree(int32_t x) const {
        const double d = std::floor(scale * weight(x));
        const double cap = std::min<double>((double)p.n - 1.0, (double)p.max_degree);
        return (int32_t)std::min(d, cap);
    }
    void calibrate(int threads) {
        std::vector<double> part((size_t)threads, 0.0);
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back([&, t] {
            double s = 0;
            for (int64_t x = t; x < p.n; x += threads) s += weight((int32_t)x);
            part[(size_t)t] = s;
        });
        for (auto& t : th) t.join();
        double s = 0;
        for (double v : part) s += v;
        scale = s > 0 ? ((double)p.target_arcs + 0.5 * p.n) / s : 0.0;
    }
    // successors of x; block_lists[y - block_start] holds the lists of the earlier nodes of the same block
    void successors(int32_t x, int32_t block_start, const std::vector<std::vector<int32_t>>& block_lists, std::vector<int32_t>& out) const {
        out.clear();
        int32_t d = degree(x);
        if (d <= 0) return;
        Rng rng{ mix64(p.seed * 0x2545F4914F6CDD1DULL + (uint64_t)x) };
        const int32_t back = x - block_start;
        if (back > 0 && rng.unit() < p.p_copy) {
            const int32_t r = 1 + (int32_t)rng.below((uint32_t)std::min(7, back));
            const std::vector<int32_t>& proto = block_lists[(size_t)(back - r)];
            if (p.p_same_degree > 0.0 && !proto.empty() && rng.unit() < p.p_same_degree)
                d = (int32_t)std::min<int64_t>((int64_t)proto.size() + (int64_t)rng.below(4), (int64_t)p.n - 1);
            size_t pos = 0;
            bool copying = rng.unit() < 0.8;
            while (pos < proto.size() && (int32_t)out.size() < d) {
                const double mean = copying ? p.copy_run : p.skip_run;
                size_t run = 1 + (size_t)(-mean * std::log(1.0 - rng.unit()));
                run = std::min(run, proto.size() - pos);
                if (copying) for (size_t i = 0; i < run && (int32_t)out.size() < d; i++) out.push_back(proto[pos + i]);
                pos += run;
                copying = !copying;
            }
        }
        if ((int32_t)out.size() < d && rng.unit() < p.p_interval) {
            const int32_t k = 1 + (int32_t)rng.below((uint32_t)std::max(1, p.interval_max));
            for (int32_t i = 0; i < k && (int32_t)out.size() < d; i++) {
                int64_t start = (int64_t)x + (int64_t)rng.below(8192) - 4096;
                const int32_t len = 4 + (int32_t)rng.below(17);
                for (int32_t t = 0; t < len && (int32_t)out.size() < d; t++) {
                    const int64_t v = start + t;
                    if (v >= 0 && v < p.n) out.push_back((int32_t)v);
                }
            }
        }
        const int32_t missing = d - (int32_t)out.size();
        for (int32_t i = 0; i < missing; i++) {
            int64_t t;
            if (rng.unit() < p.p_local) {  // log-uniform distance up to 2^16, either side
                const uint32_t bits = 1 + rng.below((uint32_t)std::min(30, std::max(1, p.local_bits)));
                const int64_t dist = 1 + (int64_t)rng.below(1u << bits);
                t = (rng.next() & 1) ? (int64_t)x + dist : (int64_t)x - dist;
                if (t < 0) t = -t;
                if (t >= p.n) t = 2 * ((int64_t)p.n - 1) - t;
                if (t < 0) t = 0;
            } else {  // global, density ~ t^(-2/3): popular low ids
                const double u = rng.unit();
                t = (int64_t)(u * u * u * (double)p.n);
                if (t >= p.n) t = p.n - 1;
            }
            out.push_back((int32_t)t);
        }
        std::sort(out.begin(), out.end());
        out.erase(std::unique(out.begin(), out.end()), out.end());
    }
};

}  // namespace

extern "C" {

int bvgt_store_csr(const char* basename, int32_t n, const int64_t* off, const int32_t* succ,
                   int32_t window, int32_t maxref, int32_t minlen, int32_t zetak, uint32_t flags,
                   int threads, bvgt_store_stats* stats) {
    if (!basename || n < 0 || window < 0 || minlen < 0 || zetak < 1 || !off) return -1;
    Codec c{ window, maxref, minlen, zetak, flags };
    c.set_flags(flags);
    if (!c.ok) return -3;
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    std::vector<RangeCompressor*> parts;
    for (int t = 0; t < threads; t++) parts.push_back(new RangeCompressor(c));
    const int64_t step = ((int64_t)n + threads - 1) / threads;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t] {
        RangeCompressor& rc = *parts[(size_t)t];
        const int64_t a = t * step, b = std::min<int64_t>(n, a + step);
        for (int64_t x = a; x < b && !rc.bad; x++) {
            const int64_t d = off[x + 1] - off[x];
            if (d < 0 || d > 0x7fffffff) { rc.bad = true; break; }
            rc.add((int32_t)x, succ + off[x], (int32_t)d);
        }
    });
    for (auto& t : th) t.join();
    const int rc = finish(basename, c, parts, stats);
    for (auto* p : parts) delete p;
    return rc;
}

int64_t bvgt_write_codes(int coding, int32_t k, const uint64_t* values, int64_t count, uint8_t* out, int64_t cap) {
    if (!values || count < 0 || !out || cap < 0) return -1;
    if (coding < BVGT_DELTA || coding > BVGT_NIBBLE || coding == BVGT_SKEWED_GOLOMB) return -3;
    if ((coding == BVGT_ZETA || coding == BVGT_GOLOMB) && k < 1) return -1;
    Codec c{ 0, 0, 0, k, 0 };
    BitBuf b;
    for (int64_t i = 0; i < count; i++) c.put(b, coding, values[i]);
    const int64_t nbytes = (int64_t)((b.nbits + 7) >> 3);
    if (nbytes > cap) return -1;
    for (int64_t i = 0; i < nbytes; i++) out[i] = (uint8_t)(b.w[(size_t)(i >> 3)] >> (56 - 8 * (i & 7)));
    return (int64_t)b.nbits;
}

void bvgt_gen_defaults(bvgt_gen_params* p, int32_t n, int64_t target_arcs, uint64_t seed) {
    p->n = n; p->target_arcs = target_arcs; p->seed = seed;
    p->zipf_s = 0.65; p->p_copy = 0.5; p->p_interval = 0.1; p->p_local = 0.5; p->block = 1024; p->max_degree = 1 << 22;
    p->copy_run = 8.0; p->skip_run = 3.0; p->local_bits = 16; p->interval_max = 3; p->p_same_degree = 0.0;
}

int bvgt_generate_store(const char* basename, const bvgt_gen_params* gp,
                        int32_t window, int32_t maxref, int32_t minlen, int32_t zetak, uint32_t flags,
                        int threads, int64_t* out_off, int32_t* out_succ, int64_t succ_cap,
                        bvgt_store_stats* stats) {
    if (!basename || !gp || gp->n <= 0 || gp->block <= 0 || window < 0 || minlen < 0 || zetak < 1) return -1;
    Codec c{ window, maxref, minlen, zetak, flags };
    c.set_flags(flags);
    if (!c.ok) return -3;
    Generator g{ *gp };
    if (threads < 1) threads = 1;
    g.calibrate(threads);
    const int64_t nblocks = ((int64_t)gp->n + gp->block - 1) / gp->block;
    if (threads > nblocks) threads = (int)nblocks;
    const int64_t bstep = (nblocks + threads - 1) / threads;
    std::vector<RangeCompressor*> parts;
    for (int t = 0; t < threads; t++) parts.push_back(new RangeCompressor(c));
    const bool want_csr = out_off != nullptr;
    std::vector<std::vector<int32_t>> csr_parts((size_t)threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t] {
        RangeCompressor& rc = *parts[(size_t)t];
        std::vector<std::vector<int32_t>> lists((size_t)gp->block);
        for (int64_t b = t * bstep; b < std::min(nblocks, (t + 1) * bstep) && !rc.bad; b++) {
            const int32_t start = (int32_t)(b * gp->block);
            const int32_t end = (int32_t)std::min<int64_t>(gp->n, (int64_t)start + gp->block);
            for (int32_t x = start; x < end; x++) {
                std::vector<int32_t>& l = lists[(size_t)(x - start)];
                g.successors(x, start, lists, l);
                rc.add(x, l.data(), (int32_t)l.size());
                if (want_csr) {
                    out_off[x + 1] = (int64_t)l.size();  // degrees for now, prefix-summed below
                    if (out_succ) csr_parts[(size_t)t].insert(csr_parts[(size_t)t].end(), l.begin(), l.end());
                }
            }
        }
    });
    for (auto& t : th) t.join();
    int rc = finish(basename, c, parts, stats);
    for (auto* p : parts) delete p;
    if (rc == 0 && want_csr) {
        out_off[0] = 0;
        for (int64_t x = 0; x < gp->n; x++) out_off[x + 1] += out_off[x];
        if (out_succ) {
            if (out_off[gp->n] > succ_cap) return -1;
            int64_t pos = 0;
            for (auto& v : csr_parts) { std::memcpy(out_succ + pos, v.data(), v.size() * sizeof(int32_t)); pos += (int64_t)v.size(); }
        }
    }
    return rc;
}

int bvgt_store_labels(const char* basename, const char* underlying, const char* key, int32_t n, const int64_t* off,
                      const int64_t* list_off, const int32_t* values, int kind, int width, int threads, int64_t* label_bits) {
    if (!basename || !underlying || !key || n < 0 || !off) return -1;
    if (kind < BVGT_LABEL_GAMMA || kind > BVGT_LABEL_FIXED_LIST) return -1;
    if (kind != BVGT_LABEL_GAMMA && (width < 0 || width > 31)) return -1;
    if (kind == BVGT_LABEL_FIXED_LIST && !list_off) return -1;
    if (off[n] > 0 && !values && !(kind == BVGT_LABEL_FIXED_LIST && list_off[off[n]] == 0)) return -1;
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    struct Part { BitBuf labels, offs; bool bad = false; };
    std::vector<Part> parts((size_t)threads);
    const int64_t step = ((int64_t)n + threads - 1) / threads;
    const uint64_t lim = kind == BVGT_LABEL_GAMMA ? 0x7fffffffULL : 1ULL << width;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t] {
        Part& p = parts[(size_t)t];
        const int64_t a = t * step, b = std::min<int64_t>(n, a + step);
        for (int64_t x = a; x < b && !p.bad; x++) {
            const uint64_t before = p.labels.nbits;
            for (int64_t j = off[x]; j < off[x + 1]; j++) {
                if (kind == BVGT_LABEL_FIXED_LIST) {
                    const int64_t la = list_off[j], lb = list_off[j + 1];
                    if (lb < la || lb - la > 0x7ffffffe) { p.bad = true; break; }
                    put_gamma(p.labels, (uint64_t)(lb - la));
                    for (int64_t k = la; k < lb; k++) {
                        if (values[k] < 0 || (uint64_t)values[k] >= lim) { p.bad = true; break; }
                        p.labels.put((uint64_t)values[k], width);
                    }
                } else {
                    if (values[j] < 0 || (uint64_t)values[j] >= lim) { p.bad = true; break; }
                    if (kind == BVGT_LABEL_GAMMA) put_gamma(p.labels, (uint64_t)values[j]);
                    else p.labels.put((uint64_t)values[j], width);
                }
            }
            put_gamma(p.offs, p.labels.nbits - before);
        }
    });
    for (auto& t : th) t.join();
    BitBuf labels, offs;
    put_gamma(offs, 0);
    for (Part& p : parts) {
        if (p.bad) return -1;
        if (parts.size() == 1) labels = std::move(p.labels); else { labels.append(p.labels); p.labels = BitBuf(); }
        offs.append(p.offs);
    }
    const std::string base(basename);
    if (!labels.write_file(base + ".labels")) return -4;
    if (!offs.write_file(base + ".labeloffsets")) return -4;
    FILE* f = fopen((base + ".properties").c_str(), "w");
    if (!f) return -4;
    static const char* cls[] = { "GammaCodedIntLabel", "FixedWidthIntLabel", "FixedWidthIntListLabel" };
    fprintf(f, "graphclass = it.unimi.dsi.webgraph.labelling.BitStreamArcLabelledImmutableGraph\n");
    if (kind == BVGT_LABEL_GAMMA) fprintf(f, "labelspec = it.unimi.dsi.webgraph.labelling.%s(%s)\n", cls[kind], key);
    else fprintf(f, "labelspec = it.unimi.dsi.webgraph.labelling.%s(%s,%d)\n", cls[kind], key, width);
    fprintf(f, "underlyinggraph = %s\n", underlying);
    fclose(f);
    if (label_bits) *label_bits = (int64_t)labels.nbits;
    return 0;
}

int bvgt_store_ef(const char* basename, int32_t n, const int64_t* off, const int32_t* succ, int32_t upper_bound,
                  int log2_quantum, int big_endian, int threads, int64_t* graph_bits) {
    if (!basename || n < 0 || !off || log2_quantum < 0 || log2_quantum > 30) return -1;
    if (off[n] > 0 && !succ) return -1;
    const uint64_t ub = upper_bound > 0 ? (uint64_t)upper_bound : (uint64_t)n;
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    struct Part { LsbBuf graph; BitBuf offs; int64_t arcs = 0; uint64_t bits_out = 0, bits_succ = 0; bool bad = false; };
    std::vector<Part> parts((size_t)threads);
    const int64_t step = ((int64_t)n + threads - 1) / threads;
    const int q = log2_quantum;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t] {
        Part& p = parts[(size_t)t];
        LsbBuf pointers, lower, upper;
        const int64_t a = t * step, b = std::min<int64_t>(n, a + step);
        for (int64_t x = a; x < b && !p.bad; x++) {
            const int64_t d = off[x + 1] - off[x];
            if (d < 0 || d > 0x7ffffffe) { p.bad = true; break; }
            const uint64_t before = p.graph.nbits;
            p.graph.gamma((uint64_t)d);
            p.bits_out += p.graph.nbits - before;
            // Accumulator.init(outdegree, upperBound, strict = false, indexZeroes = true, log2Quantum), :472-497
            const uint64_t len = (uint64_t)d + 1;
            const int l = ef_lower_bits(len, ub), psize = ef_pointer_size(len, ub);
            const uint64_t mask = l ? (~0ULL >> (64 - l)) : 0;
            pointers.w.clear(); pointers.nbits = 0; lower.w.clear(); lower.nbits = 0; upper.w.clear(); upper.nbits = 0;
            int64_t last_one = -1, cur_len = 0;
            uint64_t prev = 0;
            for (int64_t k = 0; k <= d; k++) {   // Accumulator.add for every successor, then for the terminator (dump, :524-528)
                const uint64_t v = k < d ? (uint64_t)(uint32_t)succ[off[x] + k] : ub;
                // successors strictly increasing and below the upper bound (Accumulator.add throws otherwise, :499-503)
                if (k < d && (succ[off[x] + k] < 0 || v >= ub || (k > 0 && v <= prev))) { p.bad = true; break; }
                prev = v;
                if (l) lower.put(v & mask, l);
                const int64_t one_pos = (int64_t)(v >> l) + cur_len;
                upper.unary((uint64_t)(one_pos - last_one - 1));
                int64_t zeroes_before = last_one - cur_len + 1;
                for (int64_t pos = last_one + (zeroes_before & ~((1LL << q) - 1)) + (1LL << q) - zeroes_before; pos < one_pos; pos += 1LL << q, zeroes_before += 1LL << q)
                    pointers.put((uint64_t)(pos + 1), psize);
                last_one = one_pos;
                cur_len++;
            }
            if (p.bad) break;
            p.graph.append(pointers); p.graph.append(lower); p.graph.append(upper);
            put_delta(p.offs, p.graph.nbits - before);
            p.bits_succ += pointers.nbits + lower.nbits + upper.nbits;
            p.arcs += d;
        }
    });
    for (auto& t : th) t.join();
    LsbBuf graph;
    BitBuf offs;
    put_delta(offs, 0);
    int64_t arcs = 0;
    uint64_t bits_out = 0, bits_succ = 0;
    for (Part& p : parts) {
        if (p.bad) return -1;
        if (parts.size() == 1) graph = std::move(p.graph); else { graph.append(p.graph); p.graph = LsbBuf(); }
        offs.append(p.offs);
        arcs += p.arcs; bits_out += p.bits_out; bits_succ += p.bits_succ;
    }
    const std::string base(basename);
    {
        FILE* f = fopen((base + ".graph").c_str(), "wb");
        if (!f) return -4;
        const uint64_t nw = (graph.nbits >> 6) + 1;   // close() always writes the buffer, :413-418
        graph.w.resize((size_t)std::max<uint64_t>(nw, graph.w.size()), 0);
        bool ok = true;
        if (!big_endian) ok = fwrite(graph.w.data(), 8, (size_t)nw, f) == nw;
        else for (uint64_t i = 0; i < nw && ok; i++) { const uint64_t v = __builtin_bswap64(graph.w[(size_t)i]); ok = fwrite(&v, 8, 1, f) == 1; }
        if (fclose(f) != 0 || !ok) return -4;
    }
    if (!offs.write_file(base + ".offsets")) return -4;
    FILE* f = fopen((base + ".properties").c_str(), "w");
    if (!f) return -4;
    fprintf(f, "#EFGraph properties\n#written by webgraph_b200 bvg_tools\n");
    fprintf(f, "nodes=%d\narcs=%lld\n", n, (long long)arcs);
    if (ub != (uint64_t)n) fprintf(f, "upperbound=%llu\n", (unsigned long long)ub);
    fprintf(f, "quantum=%lld\nbyteorder=%s\n", 1LL << q, big_endian ? "BIG_ENDIAN" : "LITTLE_ENDIAN");
    fprintf(f, "bitsforoutdegrees=%llu\nbitsforsuccessors=%llu\n", (unsigned long long)bits_out, (unsigned long long)bits_succ);
    fprintf(f, "graphclass=it.unimi.dsi.webgraph.EFGraph\nversion=0\n");
    fclose(f);
    if (graph_bits) *graph_bits = (int64_t)graph.nbits;
    return 0;
}

}  // extern "C"
