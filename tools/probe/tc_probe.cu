// tc_probe.cu -- stand-alone probe of the tcgen05 TF32 path used by the round-2 MLP core:
// D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T with K-major, no-swizzle shared-memory operands written by ordinary
// threads, single-pass TF32 and 3xTF32 (hi/lo split), checked against an fp64 host reference and timed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu      Run: ./tc_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE canonical layout: core matrix = 8 rows x 16 B, contiguous (128 B);
// addr(r, k) = (k/4) * (ROWS/8*128) + (r/8)*128 + (r%8)*16 + (k%4)*4      => SBO = 128 B, LBO = ROWS*16 B
__device__ __forceinline__ int op_off(int r, int k, int rows) { return (k >> 2) * (rows * 16) + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;          // version = 1 (Blackwell)
    return d;                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D1,
                                                       float* __restrict__ D3, int reps, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Ahi = (float*)smem;                       // 128 x 64 x 4 = 32 KB
    float* Alo = Ahi + M * K;
    float* Bhi = Alo + M * K;                        // 64 x 64 x 4 = 16 KB
    float* Blo = Bhi + N * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_base_s)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    // operands: hi = value rounded to tf32, lo = value - hi
    for (int e = tid; e < M * K; e += 128) {
        const int r = e / K, k = e % K;
        const float v = A[e];
        uint32_t hb; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        *(float*)((char*)Ahi + op_off(r, k, M)) = hi;
        *(float*)((char*)Alo + op_off(r, k, M)) = v - hi;
    }
    for (int e = tid; e < N * K; e += 128) {
        const int r = e / K, k = e % K;
        const float v = B[e];
        uint32_t hb; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        *(float*)((char*)Bhi + op_off(r, k, N)) = hi;
        *(float*)((char*)Blo + op_off(r, k, N)) = v - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N, M
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            // columns [0,64): single pass;  columns [64,128): 3xTF32
            for (int kb = 0; kb < K / 8; ++kb) {
                const uint64_t a_hi = make_desc(smem_u32(Ahi) + kb * 2 * (M * 16), M * 16, 128);
                const uint64_t b_hi = make_desc(smem_u32(Bhi) + kb * 2 * (N * 16), N * 16, 128);
                mma_tf32(tmem, a_hi, b_hi, idesc, kb > 0);
            }
            for (int kb = 0; kb < K / 8; ++kb) {
                const uint64_t a_hi = make_desc(smem_u32(Ahi) + kb * 2 * (M * 16), M * 16, 128);
                const uint64_t a_lo = make_desc(smem_u32(Alo) + kb * 2 * (M * 16), M * 16, 128);
                const uint64_t b_hi = make_desc(smem_u32(Bhi) + kb * 2 * (N * 16), N * 16, 128);
                const uint64_t b_lo = make_desc(smem_u32(Blo) + kb * 2 * (N * 16), N * 16, 128);
                mma_tf32(tmem + 64, a_lo, b_hi, idesc, kb > 0);
                mma_tf32(tmem + 64, a_hi, b_lo, idesc, 1);
                mma_tf32(tmem + 64, a_hi, b_hi, idesc, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&bar)) : "memory");
        }
        // everybody waits for the accumulators
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        }
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n");
    }
    long long t1 = clock64();
    if (tid == 0) *cycles = t1 - t0;
    // epilogue: thread t owns TMEM lane t (row t); warp w may only touch lanes 32w..32w+31
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int half = 0; half < 2; ++half) {
        float* Dst = half ? D3 : D1;
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(lane_addr + half * 64 + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 16; ++j) Dst[tid * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "n"(128));
}

int main() {
    float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hD1 = (float*)malloc(M * N * 4), *hD3 = (float*)malloc(M * N * 4);
    srand(1);
    for (int i = 0; i < M * K; ++i) hA[i] = (float)rand() / RAND_MAX * 2 - 1;
    for (int i = 0; i < N * K; ++i) hB[i] = (float)rand() / RAND_MAX * 2 - 1;
    float *dA, *dB, *dD1, *dD3; long long* dC;
    CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dB, N * K * 4)); CK(cudaMalloc(&dD1, M * N * 4)); CK(cudaMalloc(&dD3, M * N * 4)); CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice));
    const int smem = (2 * M * K + 2 * N * K) * 4;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int reps : {1, 1000}) {
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD1, dD3, reps, dC);
        CK(cudaDeviceSynchronize());
        long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hD1, dD1, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hD3, dD3, M * N * 4, cudaMemcpyDeviceToHost));
        double e1 = 0, e3 = 0, ref_max = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double r = 0; for (int k = 0; k < K; ++k) r += (double)hA[m * K + k] * (double)hB[n * K + k];
            e1 = fmax(e1, fabs(hD1[m * N + n] - r)); e3 = fmax(e3, fabs(hD3[m * N + n] - r)); ref_max = fmax(ref_max, fabs(r));
        }
        printf("reps %d: max|ref| %.3f  err single-pass tf32 %.3e  err 3xTF32 %.3e  cycles %lld (%.1f per rep: 8 + 24 MMAs of 128x64x8)\n",
               reps, ref_max, e1, e3, cyc, (double)cyc / reps);
    }
    return 0;
}
