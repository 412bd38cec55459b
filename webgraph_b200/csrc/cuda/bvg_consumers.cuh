// bvg_consumers.cuh -- kernels of the fused consumers that work on decoded rows.
//
// HyperBall's inner loop (reference algo/HyperBall.java:875-915): the counter of node x becomes the register-wise maximum of
// its old value and the old counters of its successors.  The reference packs registers of 5-7 bits into longs and takes the
// maximum with broadword arithmetic on the CPU; here a register is a byte, a counter is m = 2^log2m bytes (m >= 16), and a
// group of m / 16 lanes owns a node: each lane keeps 16 registers in a uint4 and takes __vmaxu4 against the same 16 bytes of
// every successor's counter (one 16-byte gather per lane and arc).  Rows come from the range decoder, chunk by chunk; a node
// with more than HB_HEAVY successors is taken by a whole block instead (its successors strided over the block's groups,
// partial maxima combined in shared memory).
#pragma once
#include "bvg_device.cuh"

namespace bvg {

constexpr int HB_THREADS = 256;
constexpr int32_t HB_HEAVY = 4096;

__device__ __forceinline__ uint4 hb_max(uint4 a, uint4 b) {
    return uint4{ __vmaxu4(a.x, b.x), __vmaxu4(a.y, b.y), __vmaxu4(a.z, b.z), __vmaxu4(a.w, b.w) };
}
__device__ __forceinline__ bool hb_differs(uint4 a, uint4 b) { return ((a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w)) != 0; }

// rows: successors of nodes [from, from + count), off[0..count] relative to rows.  in / out: counters of all nodes
// (node * m bytes).  lanes = m / 16 (a power of two, 1..32).  Heavy nodes are appended to heavy[] and left alone.
__global__ void __launch_bounds__(HB_THREADS) k_hb_update(const int32_t* __restrict__ rows, const int64_t* __restrict__ off, int32_t from, int64_t count,
                                                          const uint4* __restrict__ in, uint4* __restrict__ out, int lanes_log,
                                                          int32_t* __restrict__ heavy, int32_t* __restrict__ nheavy,
                                                          unsigned long long* __restrict__ modified) {
    const int lanes = 1 << lanes_log;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> lanes_log;
    const int sub = threadIdx.x & (lanes - 1);
    bool changed = false;
    if (gid < count) {
        const int64_t a = off[gid], b = off[gid + 1];
        const int64_t x = (int64_t)from + gid;
        if (b - a > HB_HEAVY) {
            if (sub == 0) heavy[atomicAdd(nheavy, 1)] = (int32_t)gid;
        } else {
            const uint4 old = in[x * lanes + sub];
            uint4 t = old;
            int64_t j = a;
            for (; j + 1 < b; j += 2) {   // two gathers in flight per lane
                const int64_t s0 = rows[j], s1 = rows[j + 1];
                const uint4 u0 = in[s0 * lanes + sub], u1 = in[s1 * lanes + sub];
                t = hb_max(t, hb_max(u0, u1));
            }
            if (j < b) t = hb_max(t, in[(int64_t)rows[j] * lanes + sub]);
            out[x * lanes + sub] = t;
            changed = hb_differs(t, old);
        }
    }
    // a node counts once: OR over its group, then one vote per group leader
    unsigned m = __ballot_sync(0xffffffffu, changed);
    if (lanes < 32) {
        const unsigned grp = ((lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u)) << ((threadIdx.x & 31) & ~(lanes - 1)));
        changed = (m & grp) != 0 && sub == 0;
    } else changed = m != 0 && sub == 0;
    m = __ballot_sync(0xffffffffu, changed);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(modified, (unsigned long long)__popc(m));
}

// One block per heavy node.
__global__ void __launch_bounds__(HB_THREADS) k_hb_update_heavy(const int32_t* __restrict__ rows, const int64_t* __restrict__ off, int32_t from,
                                                                const int32_t* __restrict__ heavy, const uint4* __restrict__ in, uint4* __restrict__ out,
                                                                int lanes_log, unsigned long long* __restrict__ modified) {
    __shared__ uint4 part[HB_THREADS];
    const int lanes = 1 << lanes_log;
    const int64_t gid = heavy[blockIdx.x];
    const int64_t a = off[gid], b = off[gid + 1];
    const int64_t x = (int64_t)from + gid;
    const int sub = threadIdx.x & (lanes - 1), grp = threadIdx.x >> lanes_log, ngrp = HB_THREADS >> lanes_log;
    const uint4 old = in[x * lanes + sub];
    uint4 t = old;
    for (int64_t j = a + grp; j < b; j += ngrp) t = hb_max(t, in[(int64_t)rows[j] * lanes + sub]);
    part[threadIdx.x] = t;
    __syncthreads();
    for (int stride = HB_THREADS >> 1; stride >= lanes; stride >>= 1) {   // threads with the same `sub` are `lanes` apart
        if (threadIdx.x < stride) part[threadIdx.x] = hb_max(part[threadIdx.x], part[threadIdx.x + stride]);
        __syncthreads();
    }
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    if (threadIdx.x < lanes) {
        out[x * lanes + threadIdx.x] = part[threadIdx.x];
        if (hb_differs(part[threadIdx.x], old)) atomicOr(&any, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0 && any) atomicAdd(modified, 1ull);
}

}  // namespace bvg
