// Microbenchmark: what does a random 32-byte (one sector) read from a table much larger than L2 cost in DRAM
// traffic, per load flavour and per cudaLimitMaxL2FetchGranularity?  (Why: ncu showed ~4 sectors of DRAM read per
// 32-byte cuckoo slot probe.)  Each thread reads two random 32-byte slots per iteration, like the resolve kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_gran probe_gran.cu ; ./probe_gran <granularity 0|32|64|128>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

struct alignas(32) Slot { uint32_t w[8]; };

__device__ __forceinline__ uint32_t mix(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }

template <int V>
__device__ __forceinline__ uint32_t load_slot(const Slot* s) {
    uint32_t a, b, c, d, e, f, g, h;
    if (V == 0) {        // two 16-byte read-only loads (what resolve_device.cuh does)
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(s));
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(s));
    } else if (V == 1) { // coherent loads
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(s));
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(s));
    } else if (V == 2) { // one 256-bit read-only load
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(s));
    } else if (V == 3) { // no L1 allocation
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(s));
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(s));
    } else if (V == 4) { // explicit 64-byte L2 prefetch hint (the smallest PTX offers)
        asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(s));
        asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(s));
    } else if (V == 5) { // cache-global (L2 only)
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(s));
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(s));
    } else {             // streaming / evict-first
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(s));
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(s));
    }
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int V>
__global__ void __launch_bounds__(128, 8) probe(const Slot* table, uint32_t mask, uint32_t seed, uint32_t* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t h1 = mix(t * 0x9E3779B1u + seed), h2 = mix(h1 + 0x7F4A7C15u);
    out[t] = load_slot<V>(table + (h1 & mask)) ^ load_slot<V>(table + (h2 & mask));
}

int main(int argc, char** argv) {
    const int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran); printf("set granularity %d: %s\n", gran, cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("cudaLimitMaxL2FetchGranularity = %zu\n", g);
    const uint32_t nslots = 1u << 25;                      // 1 GiB of 32-byte slots: nothing stays in the 126 MB L2
    Slot* table; cudaMalloc(&table, (size_t)nslots * sizeof(Slot)); cudaMemset(table, 1, (size_t)nslots * sizeof(Slot));
    const uint32_t n = 1u << 20;                           // 1M threads x 2 slots = 64 MiB of sectors asked for
    uint32_t* out; cudaMalloc(&out, n * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int v = 0; v < 7; v++) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(a);
            switch (v) {
            case 0: probe<0><<<n / 128, 128>>>(table, nslots - 1, rep * 977 + v, out); break;
            case 1: probe<1><<<n / 128, 128>>>(table, nslots - 1, rep * 977 + v, out); break;
            case 2: probe<2><<<n / 128, 128>>>(table, nslots - 1, rep * 977 + v, out); break;
            case 3: probe<3><<<n / 128, 128>>>(table, nslots - 1, rep * 977 + v, out); break;
            case 4: probe<4><<<n / 128, 128>>>(table, nslots - 1, rep * 977 + v, out); break;
            case 5: probe<5><<<n / 128, 128>>>(table, nslots - 1, rep * 977 + v, out); break;
            default: probe<6><<<n / 128, 128>>>(table, nslots - 1, rep * 977 + v, out); break;
            }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("variant %d: %.1f us for %u slot reads  (%.2f G slot reads/s, %.0f GB/s of 32-byte sectors)  %s\n", v, best * 1e3f, 2 * n,
               2.0 * n / (best * 1e-3) / 1e9, 2.0 * n * 32 / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
