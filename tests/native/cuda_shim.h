// Host stand-ins for the CUDA constructs binder_b200/csrc/resolve_device.cuh uses, so that the per-query device
// code compiles with g++ and runs on the CPU (tests/native/emu_resolve.cpp).  Shared memory is one byte array;
// a "shared address" is an offset into it.  PTX semantics are kept where they differ from C++ (shift counts).
#ifndef BB_CUDA_SHIM_H
#define BB_CUDA_SHIM_H
#define BB_HOST_EMU 1

#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static const
#define __align__(n) alignas(n)

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }
struct dim3_emu { unsigned x = 0, y = 0, z = 0; };
static thread_local dim3_emu threadIdx, blockIdx;

using std::max;
using std::min;

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << (sh & 31)) >> 32); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t v = ((uint64_t)b << 32) | a; uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }

// emulated shared memory of the tile being processed
extern thread_local uint8_t* bb_emu_smem;
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)((const uint8_t*)p - bb_emu_smem); }
static inline uint32_t lds32(uint32_t a) { uint32_t v; memcpy(&v, bb_emu_smem + a, 4); return v; }
static inline uint32_t lds8(uint32_t a) { return bb_emu_smem[a]; }
static inline void sts32(uint32_t a, uint32_t v) { memcpy(bb_emu_smem + a, &v, 4); }
static inline void sts8(uint32_t a, uint32_t v) { bb_emu_smem[a] = (uint8_t)v; }
static inline void sts_or(uint32_t a, uint32_t v) { uint32_t o; memcpy(&o, bb_emu_smem + a, 4); o |= v; memcpy(bb_emu_smem + a, &o, 4); }
static inline void sts_or5(uint32_t base, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4) {
    if (w0) sts_or(base, w0); if (w1) sts_or(base + 4, w1); if (w2) sts_or(base + 8, w2); if (w3) sts_or(base + 12, w3); if (w4) sts_or(base + 16, w4);
}
static inline uint32_t shl_clamp(uint32_t v, uint32_t n) { return n > 31 ? 0u : v << n; }     // shl.b32 clamps its count
static inline unsigned long long gtime() { return 0; }
static inline unsigned long long gtime_early() { return 0; }
#endif
