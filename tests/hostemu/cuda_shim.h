// cuda_shim.h -- lets the device-side decode logic (bvg_device.cuh) compile as plain host C++ for debugging under
// ASan/UBSan (tests/test_device_logic_on_host.py).  Test infrastructure only; never part of the product build.
#pragma once
#include <cstdint>
#include <cstring>
struct uint4 { uint32_t x, y, z, w; };
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t shift) {
    shift &= 31;
    return shift ? (hi << shift) | (lo >> (32 - shift)) : hi;
}
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return v == 0 ? 0 : __builtin_ctz((unsigned)v) + 1; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffsll(long long v) { return v == 0 ? 0 : __builtin_ctzll((unsigned long long)v) + 1; }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int atomicCAS(int* p, int cmp, int val) { int old = *p; if (old == cmp) *p = val; return old; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long val) { unsigned long long old = *p; if (old == cmp) *p = val; return old; }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o | v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline void __threadfence() {}
// kernels in the included headers are compiled but never called on the host: give their index variables a meaning
struct Dim3Shim { unsigned x = 0, y = 0, z = 0; };
static Dim3Shim blockIdx, blockDim, threadIdx, gridDim;
