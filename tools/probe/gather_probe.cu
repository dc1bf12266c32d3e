// gather_probe.cu -- what bounds a scattered 8-byte gather / atomic scatter on B200?
// Random float2 loads (LDG.64) and float2 atomic adds (RED.64) into a 48.8 MB table (= the L2-resident hash table of the
// field kernels), at full occupancy, with U independent accesses in flight per thread.  Reports lane-accesses per SM per
// cycle: the L1TEX wavefront rate that floors the field kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu && ./gather_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// mode 0: every lane its own random entry; mode 1: lanes 2k, 2k+1 share an aligned 16-byte pair (x-neighbour corners of an even
// cell); mode 2: groups of 8 lanes share one entry (coarse level: consecutive samples of a ray in one cell)
template <int U, int MODE>
__global__ void k_gather(const float2* __restrict__ tab, uint32_t mask, int iters, float* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = MODE == 0 ? tid : MODE == 1 ? (tid >> 1) : (tid >> 3);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t idx = mix(key * 0x9e3779b9u + it * U + u) & mask;
            if (MODE == 1) idx = (idx & ~1u) | (tid & 1u);
            v[u] = __ldg(tab + idx);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
    }
    if (acc == 123.456f) out[0] = acc;
}
template <int U, int MODE>
__global__ void k_scatter(float2* tab, uint32_t mask, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = MODE == 0 ? tid : MODE == 1 ? (tid >> 1) : (tid >> 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t idx = mix(key * 0x9e3779b9u + it * U + u) & mask;
            if (MODE == 1) idx = (idx & ~1u) | (tid & 1u);
            atomicAdd(tab + idx, make_float2(1e-9f, 1e-9f));
        }
    }
}
// 16-byte vector atomics on aligned pairs (red.global.add.v4.f32, sm_90+)
template <int U>
__global__ void k_scatter_v4(float4* tab, uint32_t mask, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t idx = mix(tid * 0x9e3779b9u + it * U + u) & (mask >> 1);
            float* p = reinterpret_cast<float*>(tab + idx);
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(1e-9f), "f"(1e-9f), "f"(1e-9f), "f"(1e-9f) : "memory");
        }
    }
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const uint32_t entries = 1u << 22;          // 4 Mi entries x 8 B = 32 MB (power of two, L2-resident like the 48.8 MB table)
    float2* tab; float* out;
    cudaMalloc(&tab, (size_t)entries * 8); cudaMemset(tab, 0, (size_t)entries * 8); cudaMalloc(&out, 4);
    printf("SMs %d, clock %.0f MHz, table %u entries (%.1f MB)\n", sms, khz / 1e3, entries, entries * 8 / 1e6);
    const int iters = 64;
#define RUN(NAME, KERNEL, U, TPB, CPS)                                                                              \
    do {                                                                                                            \
        const int grid = sms * (CPS);                                                                               \
        const float ms = timeit([&] { KERNEL<<<grid, TPB>>> ; });                                                   \
        (void)ms;                                                                                                   \
    } while (0)
    struct R { const char* name; float ms; double n; };
    auto report = [&](const char* name, float ms, double n_access) {
        const double per_sm_cyc = n_access / sms / (ms * 1e-3 * khz * 1e3);
        printf("%-58s %8.3f ms  %7.2f G lane-acc/s  %6.3f lane-acc/SM/cycle\n", name, ms, n_access / ms / 1e6, per_sm_cyc);
    };
    for (int cps : {2, 4, 8}) {
        const int tpb = 256, grid = sms * cps;
        const double n8 = (double)grid * tpb * iters * 8, n16 = (double)grid * tpb * iters * 16;
        char nm[128];
        snprintf(nm, sizeof nm, "gather  random, U=8,  %d x 256 thr/SM", cps);
        report(nm, timeit([&] { k_gather<8, 0><<<grid, tpb>>>(tab, entries - 1, iters, out); }), n8);
        snprintf(nm, sizeof nm, "gather  random, U=16, %d x 256 thr/SM", cps);
        report(nm, timeit([&] { k_gather<16, 0><<<grid, tpb>>>(tab, entries - 1, iters, out); }), n16);
        snprintf(nm, sizeof nm, "gather  16B pairs, U=8,  %d x 256 thr/SM", cps);
        report(nm, timeit([&] { k_gather<8, 1><<<grid, tpb>>>(tab, entries - 1, iters, out); }), n8);
        snprintf(nm, sizeof nm, "gather  8 lanes/entry, U=8,  %d x 256 thr/SM", cps);
        report(nm, timeit([&] { k_gather<8, 2><<<grid, tpb>>>(tab, entries - 1, iters, out); }), n8);
        snprintf(nm, sizeof nm, "scatter random RED.64, U=8,  %d x 256 thr/SM", cps);
        report(nm, timeit([&] { k_scatter<8, 0><<<grid, tpb>>>(tab, entries - 1, iters); }), n8);
        snprintf(nm, sizeof nm, "scatter 16B pairs RED.64, U=8,  %d x 256 thr/SM", cps);
        report(nm, timeit([&] { k_scatter<8, 1><<<grid, tpb>>>(tab, entries - 1, iters); }), n8);
        snprintf(nm, sizeof nm, "scatter 8 lanes/entry RED.64, U=8,  %d x 256 thr/SM", cps);
        report(nm, timeit([&] { k_scatter<8, 2><<<grid, tpb>>>(tab, entries - 1, iters); }), n8);
        snprintf(nm, sizeof nm, "scatter random RED.128 (v4), U=8,  %d x 256 thr/SM", cps);
        report(nm, timeit([&] { k_scatter_v4<8><<<grid, tpb>>>(reinterpret_cast<float4*>(tab), entries - 1, iters); }), n8);
    }
    return 0;
}
