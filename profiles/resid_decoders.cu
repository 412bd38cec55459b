// resid_decoders.cu -- micro-benchmark of three decoders of zeta_3 residual runs on the residual sections of a real .graph
// file ("next round" item 2: settle the per-code cost with a measurement).  Not part of the library: it includes
// the library's device headers for the bit window and the code readers, loads <basename>.graph/.offsets itself, finds every
// record's residual run on the device (k_runs) and then times, over exactly the same runs,
//   (a) k_lane      one lane per run, runs sorted by length (what k_scan_extras_lean does with its schedule);
//   (b) k_warp      one WARP per run (the shape BASELINE's north star sketches): the run's bits are cut into 32 sub-ranges,
//                   every lane decodes its sub-range speculatively (count, sum of gaps, exit), entries are corrected from the
//                   neighbour's exit (shuffle) until nothing moves, counts and sums go through an in-warp inclusive scan, and
//                   a second pass decodes every sub-range from its proven entry with its proven base value and folds;
//   (c) k_table     (a) with a 4096-entry shared-memory table indexed by the next 12 stream bits that yields up to three
//                   complete codes (their count, total length and running sums) per lookup, falling back to the arithmetic
//                   decoder when the next code is longer than 12 bits.
// All three fold every residual into the XOR checksum of the library (x * MIX + y) and must agree.  Output: one JSON line with
// residuals, ms and G residuals/s per kernel; run under `ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum`
// for warp instructions and lane slots per residual.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I webgraph_b200/csrc/cuda -I include \
//             profiles/resid_decoders.cu -o gpurun_out/resid_decoders
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "bvg_format.hpp"
#include "bvg_kernels.cuh"

using namespace bvg;

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

struct Run { uint64_t pos, end; int32_t x; uint32_t rc; };

// residual run of every node: parse outdegree, reference, copy blocks (copied count), intervals
__global__ void k_runs(GraphDev g, Run* __restrict__ runs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)g.node_hi - g.node_lo;
    if (i >= n) return;
    Run r{ 0, 0, (int32_t)i, 0 };
    Win b;
    b.seek(g, g.offsets[i]);
    const uint64_t d = b.gamma(g);
    int64_t copied = 0;
    if (d != 0 && g.c.window > 0) {
        const uint64_t ref = b.unary(g);
        if (ref) {
            const uint64_t bc = b.gamma(g);
            int64_t total = 0, cp = 0;
            for (uint64_t k = 0; k < bc; k++) { const int64_t blk = (int64_t)b.gamma(g) + (k ? 1 : 0); total += blk; if (!(k & 1)) cp += blk; }
            if (!(bc & 1)) cp += (int64_t)g.outdeg[i - (int64_t)ref] - total;
            copied = cp;
        }
    }
    int64_t extra = (int64_t)d - copied;
    if (extra > 0 && g.c.minlen != 0) {
        const int64_t ic = (int64_t)b.gamma(g);
        for (int64_t k = 0; k < ic; k++) { (void)b.gamma(g); extra -= (int64_t)b.gamma(g) + g.c.minlen; }
    }
    if (extra > 0) { r.pos = b.pos(g); r.end = g.offsets[i + 1]; r.rc = (uint32_t)extra; }
    runs[i] = r;
}

__device__ __forceinline__ void fold_into(unsigned long long acc, unsigned long long* out) {
#pragma unroll
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicXor(out + (blockIdx.x & 1023), acc);
}

// (a) one lane per run
__global__ void __launch_bounds__(128, 8) k_lane(GraphDev g, const Run* __restrict__ runs, const uint32_t* __restrict__ order, int64_t nruns, unsigned long long* out) {
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t - (threadIdx.x & 31) < nruns; t += stride) {
    Run r{ 0, 0, 0, 0 };
    if (t < nruns) r = runs[order[t]];
    Fold32 f;
    f.begin(r.x);
    Win b;
    b.seek(g, r.pos);
    uint32_t v = 0;
    __syncwarp();
    if (r.rc) v = (uint32_t)(int32_t)((int64_t)r.x + nat2int(zeta_any<3>(b, g, 3) - 1ull));
    if (r.rc) f.add(v);
#pragma unroll 1
    for (uint32_t i = 1; i < r.rc; i++) {
        uint32_t m, len;
        if (zeta_fast<3>(b.top(), 3, m, len)) b.skip(len);
        else m = (uint32_t)zeta_any<3>(b, g, 3);
        v += m;
        f.add(v);
    }
    f.n = r.rc;
    if (r.rc) acc ^= f.finish(r.x);
    __syncwarp();
    }
    fold_into(acc, out);
}

// (c) the same with a 12-bit lookup table in shared memory: entry = n codes (2 bits) | total length (4 bits) << 2 |
// s1 << 6 | s2 << 17 | s3 << 28 (running sums of gap + 1: each code of at most 12 bits is worth < 512)
__device__ __host__ inline unsigned long long table_entry(uint32_t w12) {
    uint32_t pos = 0, n = 0;
    unsigned long long s[3] = { 0, 0, 0 }, run = 0;
    while (n < 3) {
        if (pos >= 12) break;
        const uint32_t rest = (w12 << pos) & 0xfffu;   // bits from pos on, left-aligned in 12 bits
        if (rest == 0) break;
        int h = 0;
        while (!((rest >> (11 - h)) & 1u)) h++;
        const int longlen = 4 * h + 4;                 // long form
        if ((int)pos + longlen - 1 > 12) break;        // even the short form (4h + 3) does not fit
        // read the short form's bits; decide; needs up to 4h + 4 bits
        const int avail = 12 - (int)pos;
        if (avail < 4 * h + 3) break;
        const uint32_t rs = (rest >> (12 - (4 * h + 3)));          // h zeros, the one, 3h + 2 bits
        const uint32_t P = 1u << (3 * h);
        const uint32_t m_short = rs - (1u << (3 * h + 2));         // the 3h + 2 bits after the leading one
        uint32_t m, len;
        if (m_short < P) { m = m_short + P; len = 4 * h + 3; }
        else {
            if (avail < 4 * h + 4) break;
            const uint32_t rl = (rest >> (12 - (4 * h + 4)));
            m = rl - (1u << (3 * h + 3));
            len = 4 * h + 4;
        }
        run += m;                                                   // m = gap + 1
        s[n++] = run;
        pos += len;
    }
    return (unsigned long long)n | ((unsigned long long)pos << 2) | (s[0] << 6) | (s[1] << 17) | (s[2] << 28);
}

__global__ void __launch_bounds__(128, 6) k_table(GraphDev g, const Run* __restrict__ runs, const uint32_t* __restrict__ order, int64_t nruns, unsigned long long* out,
                                                  const unsigned long long* __restrict__ gtab) {
    __shared__ unsigned long long tab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = gtab[i];   // built once on the host; a block loops over many runs
    __syncthreads();
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t - (threadIdx.x & 31) < nruns; t += stride) {
    Run r{ 0, 0, 0, 0 };
    if (t < nruns) r = runs[order[t]];
    Fold32 f;
    f.begin(r.x);
    Win b;
    b.seek(g, r.pos);
    uint32_t v = 0;
    __syncwarp();
    if (r.rc) v = (uint32_t)(int32_t)((int64_t)r.x + nat2int(zeta_any<3>(b, g, 3) - 1ull));
    if (r.rc) f.add(v);
    uint32_t i = 1;
#pragma unroll 1
    while (i < r.rc) {
        const uint32_t top = b.top();
        const unsigned long long e = tab[top >> 20];
        const uint32_t n = (uint32_t)e & 3u;
        if (n != 0 && i + n <= r.rc) {
            f.add(v + (uint32_t)((e >> 6) & 0x7ffu));
            if (n > 1) f.add(v + (uint32_t)((e >> 17) & 0x7ffu));
            if (n > 2) f.add(v + (uint32_t)((e >> 28) & 0x7ffu));
            v += (uint32_t)((e >> (6 + 11 * (n - 1))) & 0x7ffu);
            b.skip((uint32_t)(e >> 2) & 15u);
            i += n;
        } else {
            uint32_t m, len;
            if (zeta_fast<3>(top, 3, m, len)) b.skip(len);
            else m = (uint32_t)zeta_any<3>(b, g, 3);
            v += m;
            f.add(v);
            i++;
        }
    }
    f.n = r.rc;
    if (r.rc) acc ^= f.finish(r.x);
    __syncwarp();
    }
    fold_into(acc, out);
}

// (b) one warp per run, speculative sub-ranges
struct Spec { uint64_t exit; uint32_t count; uint32_t sum; };
__device__ __forceinline__ Spec walk(const GraphDev& g, uint64_t from, uint64_t to, bool first_is_long) {
    Spec s{ from, 0, 0 };
    if (from >= to) return s;
    Win b;
    b.seek(g, from);
    uint64_t pos = from;
    while (pos < to) {
        uint32_t m, len;
        if (zeta_fast<3>(b.top(), 3, m, len)) { b.skip(len); pos += len; }
        else { m = (uint32_t)zeta_any<3>(b, g, 3); pos = b.pos(g); }
        if (!(first_is_long && s.count == 0)) s.sum += m;
        s.count++;
    }
    s.exit = pos;
    return s;
}

__global__ void __launch_bounds__(128, 8) k_warp(GraphDev g, const Run* __restrict__ runs, const uint32_t* __restrict__ order, int64_t nruns, unsigned long long* out) {
    const unsigned lane = threadIdx.x & 31u;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long acc = 0;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nruns; w += nwarps) {
        const Run r = runs[order[w]];
        if (!r.rc) continue;
        const uint64_t bits = r.end - r.pos;
        const uint64_t sub = (bits + 31) / 32;
        const uint64_t lo = r.pos + sub * lane < r.end ? r.pos + sub * lane : r.end, hi = lo + sub < r.end ? lo + sub : r.end;
        uint64_t entry = lo;
        Spec s = walk(g, entry, hi, lane == 0);
        for (;;) {   // entries from the neighbour's exit until nothing moves (lane 0 is exact, so round k proves lanes 0..k)
            const uint64_t prev_exit = __shfl_up_sync(0xffffffffu, s.exit, 1);
            bool changed = false;
            if (lane > 0 && prev_exit != entry) {
                entry = prev_exit;
                s = entry >= hi ? Spec{ entry, 0, 0 } : walk(g, entry, hi, false);
                changed = true;
            }
            if (!__any_sync(0xffffffffu, changed)) break;
        }
        // in-warp inclusive scan of the sums of gaps: value base of every sub-range
        uint32_t sum = s.sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t q = __shfl_up_sync(0xffffffffu, sum, o);
            if (lane >= (unsigned)o) sum += q;
        }
        // the first residual anchors all values: lane 0 decodes it, everybody gets it
        uint32_t first_val = 0;
        uint64_t start = entry;
        if (lane == 0) {
            Win b0;
            b0.seek(g, entry);
            first_val = (uint32_t)(int32_t)((int64_t)r.x + nat2int(zeta_any<3>(b0, g, 3) - 1ull));
            start = b0.pos(g);
        }
        first_val = __shfl_sync(0xffffffffu, first_val, 0);
        // second pass: decode from the proven entry with the proven base, fold
        Fold32 f;
        f.begin(r.x);
        uint32_t n = 0;
        uint32_t v = first_val + (sum - s.sum);
        if (lane == 0) { f.add(first_val); n = 1; v = first_val; }
        if (start < hi) {
            Win b;
            b.seek(g, start);
            uint64_t pos = start;
            while (pos < hi) {
                uint32_t m, len;
                if (zeta_fast<3>(b.top(), 3, m, len)) { b.skip(len); pos += len; }
                else { m = (uint32_t)zeta_any<3>(b, g, 3); pos = b.pos(g); }
                v += m;
                f.add(v); n++;
            }
        }
        f.n = n;
        if (n) acc ^= f.finish(r.x);
        __syncwarp();
    }
    fold_into(acc, out);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: resid_decoders <basename> [min_rc for the warp kernel = 1]\n"); return 1; }
    const std::string base = argv[1];
    const uint32_t warp_min = argc > 2 ? (uint32_t)atoi(argv[2]) : 1u;
    const uint32_t cap = argc > 3 ? (uint32_t)atoi(argv[3]) : 1024u;   // longer runs are split at sync points by the library (bvg_long.cuh): left out
    Properties p;
    if (load_properties(base, p)) { fprintf(stderr, "properties?\n"); return 1; }
    std::vector<uint8_t> graph, ostream;
    if (!slurp_file(base + ".graph", graph) || !slurp_file(base + ".offsets", ostream)) { fprintf(stderr, "files?\n"); return 1; }
    std::vector<uint64_t> offs;
    if (decode_offsets_stream(ostream.data(), ostream.size(), 2, p.nodes, offs)) { fprintf(stderr, "offsets?\n"); return 1; }
    const int64_t n = p.nodes;
    const uint64_t nwords = ((graph.size() + 3) / 4 + STREAM_PAD_WORDS + 3) & ~(uint64_t)3;
    uint32_t* d_words; uint64_t* d_off; int32_t *d_outdeg, *d_ref; ErrWord* d_err; Run* d_runs; uint32_t* d_order; unsigned long long* d_out;
    CKC(cudaMalloc(&d_words, nwords * 4)); CKC(cudaMemset(d_words, 0, nwords * 4));
    CKC(cudaMemcpy(d_words, graph.data(), graph.size(), cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&d_off, (n + 1) * 8)); CKC(cudaMemcpy(d_off, offs.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&d_outdeg, n * 4)); CKC(cudaMalloc(&d_ref, n * 4)); CKC(cudaMalloc(&d_err, sizeof(ErrWord))); CKC(cudaMemset(d_err, 0, sizeof(ErrWord)));
    CKC(cudaMalloc(&d_runs, n * sizeof(Run))); CKC(cudaMalloc(&d_order, n * 4)); CKC(cudaMalloc(&d_out, 1024 * 8));
    k_bswap<<<(unsigned)((nwords + 255) / 256), 256>>>(d_words, nwords);
    GraphDev g{};
    g.words = d_words; g.nwords = nwords; g.bit_base = 0; g.bit_end = offs[n]; g.offsets = d_off; g.node_lo = 0; g.node_hi = (int32_t)n;
    g.c = Codec{ C_GAMMA, C_GAMMA, C_ZETA, C_UNARY, C_GAMMA, p.zetak, p.window, p.minlen };
    g.outdeg = d_outdeg; g.ref = d_ref; g.err = d_err;
    k_header<true><<<(unsigned)((n + 255) / 256), 256>>>(g, d_outdeg, d_ref);
    k_runs<<<(unsigned)((n + 255) / 256), 256>>>(g, d_runs);
    CKC(cudaDeviceSynchronize());
    std::vector<Run> runs(n);
    CKC(cudaMemcpy(runs.data(), d_runs, n * sizeof(Run), cudaMemcpyDeviceToHost));
    // runs with residuals, longest first (what the library's schedule gives the lanes of a warp: runs of similar length)
    std::vector<uint32_t> order;
    order.reserve(n);
    uint64_t resid = 0, resid_bits = 0;
    uint64_t skipped_runs = 0, skipped_resid = 0;
    for (int64_t i = 0; i < n; i++) if (runs[i].rc) {
        if (runs[i].rc > cap) { skipped_runs++; skipped_resid += runs[i].rc; continue; }
        order.push_back((uint32_t)i); resid += runs[i].rc; resid_bits += runs[i].end - runs[i].pos;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return runs[a].rc > runs[b].rc; });
    const int64_t nruns = (int64_t)order.size();
    int64_t nwarp = nruns;   // the warp kernel takes the runs with at least warp_min residuals (they are first in the order)
    uint64_t resid_warp = resid;
    if (warp_min > 1) { nwarp = 0; resid_warp = 0; while (nwarp < nruns && runs[order[nwarp]].rc >= warp_min) { resid_warp += runs[order[nwarp]].rc; nwarp++; } }
    CKC(cudaMemcpy(d_order, order.data(), nruns * 4, cudaMemcpyHostToDevice));
    std::vector<unsigned long long> htab(4096);
    for (uint32_t i = 0; i < 4096; i++) htab[i] = table_entry(i);
    unsigned long long* d_tab;
    CKC(cudaMalloc(&d_tab, 4096 * 8));
    CKC(cudaMemcpy(d_tab, htab.data(), 4096 * 8, cudaMemcpyHostToDevice));
    const unsigned pgrid = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto reduce = [&]() { std::vector<unsigned long long> h(1024); CKC(cudaMemcpy(h.data(), d_out, 1024 * 8, cudaMemcpyDeviceToHost)); unsigned long long x = 0; for (auto v : h) x ^= v; return x; };
    double ms[4] = { 0, 0, 0, 0 };
    unsigned long long cs[4] = { 0, 0, 0, 0 };
    for (int k = 0; k < 4; k++) {
        for (int rep = 0; rep < 3; rep++) {
            CKC(cudaMemset(d_out, 0, 1024 * 8));
            cudaEventRecord(e0);
            if (k == 0) k_lane<<<pgrid, 128>>>(g, d_runs, d_order, nruns, d_out);
            else if (k == 1) k_table<<<148 * 6, 128>>>(g, d_runs, d_order, nruns, d_out, d_tab);
            else if (k == 2) k_warp<<<pgrid, 128>>>(g, d_runs, d_order, nwarp, d_out);
            else k_lane<<<pgrid, 128>>>(g, d_runs, d_order, nwarp, d_out);   // (a) on the warp kernel's subset
            cudaEventRecord(e1);
            CKC(cudaDeviceSynchronize());
            float t;
            cudaEventElapsedTime(&t, e0, e1);
            ms[k] = t;
        }
        cs[k] = reduce();
    }
    printf("{\"graph\": \"%s\", \"cap\": %u, \"runs_left_out\": %llu, \"residuals_left_out\": %llu, \"runs\": %lld, \"residuals\": %llu, \"residual_bits\": %llu, \"bits_per_residual\": %.2f, "
           "\"lane\": {\"ms\": %.3f, \"G_per_s\": %.1f}, \"table\": {\"ms\": %.3f, \"G_per_s\": %.1f, \"agrees\": %s}, "
           "\"warp\": {\"min_rc\": %u, \"runs\": %lld, \"residuals\": %llu, \"ms\": %.3f, \"G_per_s\": %.1f, \"agrees\": %s, \"lane_same_subset_ms\": %.3f, \"lane_same_subset_G_per_s\": %.1f}}\n",
           base.c_str(), cap, (unsigned long long)skipped_runs, (unsigned long long)skipped_resid, (long long)nruns, (unsigned long long)resid, (unsigned long long)resid_bits, (double)resid_bits / (double)resid,
           ms[0], resid / ms[0] / 1e6, ms[1], resid / ms[1] / 1e6, cs[1] == cs[0] ? "true" : "false",
           warp_min, (long long)nwarp, (unsigned long long)resid_warp, ms[2], resid_warp / ms[2] / 1e6, cs[2] == cs[3] ? "true" : "false", ms[3], resid_warp / ms[3] / 1e6);
    return (cs[1] == cs[0] && cs[2] == cs[3]) ? 0 : 3;
}
