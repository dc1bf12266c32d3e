// tc_probe4.cu -- which shared-memory word does tcgen05.mma kind::tf32 fetch for operand element (row, k)?
// One K = 8 MMA with a one-hot partner operand: D[m][n] = (the probed operand's element).  The probed operand lives in a
// 16 KB shared-memory window filled with word-index patterns (two runs: index & 63, index >> 6 -- both exact in tf32), so the
// host can print the byte offset fetched for every (row, k) under a given descriptor (LBO, SBO, layout type, major-ness).
//   mode 0: probe B (N = 64 rows), A one-hot from TMEM            D[m][n] = B[n][m % 8]
//   mode 1: probe A from shared memory (M = 128 rows), B one-hot  D[m][n] = A[m][n % 8]   (B K-major no-swizzle, known good)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr int WIN_WORDS = 4096 * 2;     // 32 KB probed window
__global__ void __launch_bounds__(128, 1) probe4(int mode, int pattern, int lbo, int sbo, int layout, int mn, int start_off, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* win = (float*)smem;                       // probed operand window (1024-byte aligned)
    float* onehot = (float*)(smem + WIN_WORDS * 4);  // one-hot B (mode 1): [64 n][8 k] K-major no-swizzle: addr(n,k) = (k/4)*1024 + n*16 + (k%4)*4
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_base_s)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
    for (int e = tid; e < WIN_WORDS; e += 128) win[e] = pattern == 0 ? (float)((e & 63) + 1) : (float)((e >> 6) + 1);
    for (int e = tid; e < 64 * 8; e += 128) { const int n = e >> 3, k = e & 7; onehot[(k >> 2) * 256 + n * 4 + (k & 3)] = (k == (n & 7)) ? 1.f : 0.f; }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    {   // one-hot A in TMEM columns [0, 8): A[m][k] = (k == m % 8)
        uint32_t v[8];
        for (int q = 0; q < 8; ++q) v[q] = __float_as_uint(q == (tid & 7) ? 1.0f : 0.0f);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                     :: "r"(lane_addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        for (int q = 0; q < 8; ++q) v[q] = __float_as_uint(-7.f);     // poison D so "nothing written" is visible
        for (int c = 0; c < 64; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                         :: "r"(lane_addr + 64 + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        if (mode == 0) {
            const uint64_t b = make_desc(smem_u32(win) + start_off, lbo, sbo, layout);
            const uint32_t idesc = make_idesc(128, 64, 0, mn);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                         :: "r"(tmem + 64), "r"(tmem), "l"(b), "r"(idesc), "r"(0));
        } else {
            const uint64_t a = make_desc(smem_u32(win) + start_off, lbo, sbo, layout);
            const uint64_t b = make_desc(smem_u32(onehot), 1024, 128, 0);
            const uint32_t idesc = make_idesc(128, 64, mn, 0);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         :: "r"(tmem + 64), "l"(a), "l"(b), "r"(idesc), "r"(0));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    for (int c = 0; c < 64; c += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lane_addr + 64 + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int q = 0; q < 8; ++q) D[tid * 64 + c + q] = __uint_as_float(r[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "n"(128));
}
int main() {
    float* dD; CK(cudaMalloc(&dD, 128 * 64 * 4));
    static float h[2][128 * 64];
    const int smem = WIN_WORDS * 4 + 64 * 8 * 4;
    CK(cudaFuncSetAttribute(probe4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // {mode, lbo, sbo, layout, mn, start_off}
    const int combos[][6] = {
        {0, 1024, 128, 0, 0, 0},       // control: B K-major no-swizzle, 64 rows: addr(n,k) = (k/4)*1024 + (n/8)*128 + (n%8)*16 + (k%4)*4
        {0, 128, 1024, 0, 1, 0},       // B MN-major no-swizzle, canonical ((1,n),(8,k)):((X,SBO),(1,LBO))
        {0, 1024, 128, 0, 1, 0},
        {0, 4096, 1024, 2, 1, 0},      // B MN-major SW128: 32 n contiguous per k row (128 B), 8 k rows per 1024-B atom, n blocks LBO apart
        {0, 1024, 4096, 2, 1, 0},
        {0, 16, 1024, 2, 0, 0},        // B K-major SW128 (rows = n at 128 B, 8-row atoms SBO apart)
        {0, 16, 1024, 2, 0, 32},       // same, K advanced by 8 elements (start + 32 B)
        {1, 1024, 128, 0, 0, 0},       // A from smem K-major no-swizzle (control)
        {1, 128, 1024, 0, 1, 0},       // A MN-major no-swizzle
        {1, 4096, 1024, 2, 1, 0},      // A MN-major SW128
        {1, 1024, 4096, 2, 1, 0},
        {1, 16, 1024, 2, 0, 0},        // A K-major SW128
        {1, 16, 1024, 2, 0, 64},
    };
    for (auto& cb : combos) {
        for (int p = 0; p < 2; ++p) {
            probe4<<<1, 128, smem>>>(cb[0], p, cb[1], cb[2], cb[3], cb[4], cb[5], dD);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h[p], dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        }
        printf("mode %d LBO %4d SBO %4d layout %d mn %d start %3d\n", cb[0], cb[1], cb[2], cb[3], cb[4], cb[5]);
        const int rows[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 32, 33, 63, 64, 127};
        for (int r : rows) {
            if (cb[0] == 0 && r >= 64) continue;
            printf("  row %3d:", r);
            for (int k = 0; k < 8; ++k) {
                // mode 0: D[m][n] = B[n][m%8] -> element (row = n, k) at D[k][row];  mode 1: D[m][n] = A[m][n%8] -> D[row][k]
                const int idx = cb[0] == 0 ? k * 64 + r : r * 64 + k;
                const float lo = h[0][idx], hi2 = h[1][idx];
                if (lo < 0 || hi2 < 0) printf("  (none)"); else if (lo == 0 || hi2 == 0) printf("  (zero)"); else printf(" %6d", ((int)(hi2 - 1) * 64 + (int)(lo - 1)) * 4);   // byte offset fetched
            }
            printf("\n");
        }
    }
    return 0;
}
