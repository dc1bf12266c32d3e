// tc_probe3.cu -- reverse-engineers the MN-major (transposed) shared-memory operand addressing of tcgen05.mma kind::tf32.
// A (from TMEM) is one-hot: A[m][k] = (k == m % 64), so D[m][n] = B[n][k = m % 64] = whatever element the tensor core fetched for (n, k).
// The smem array holds W[j][i] in our fixed layout addr(j,i) = (i/4)*1024 + (j/8)*128 + (j%8)*16 + (i%4)*4 with value patterns j or i.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int op_off(int r, int k, int rows) { return (k >> 2) * (rows * 16) + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__global__ void __launch_bounds__(128, 1) probe3(int pattern, int lbo, int sbo, int kstep_bytes, int b_mn, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Ws = (float*)smem;    // [64 j][64 i]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_base_s)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
    for (int e = tid; e < 64 * 64; e += 128) { const int j = e / 64, i = e % 64; *(float*)((char*)Ws + op_off(j, i, 64)) = pattern == 0 ? (float)j : (float)i; }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c += 8) {
        uint32_t v[8];
        for (int q = 0; q < 8; ++q) v[q] = __float_as_uint((c + q) == (tid % 64) ? 1.0f : 0.0f);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                     :: "r"(lane_addr + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        for (int kb = 0; kb < 8; ++kb) {
            const uint64_t b = make_desc(smem_u32(Ws) + kb * kstep_bytes, lbo, sbo);
            const uint32_t idesc = make_idesc(128, 64, 0, b_mn);
            const uint32_t acc = kb > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                         :: "r"(tmem + 64), "r"(tmem + kb * 8), "l"(b), "r"(idesc), "r"(acc));
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
    float *dD; CK(cudaMalloc(&dD, 128 * 64 * 4));
    float* hj = (float*)malloc(128 * 64 * 4); float* hi = (float*)malloc(128 * 64 * 4);
    CK(cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 4));
    // {lbo, sbo, kstep, b_mn}; first row = K-major control: B[n = j][k = i] = W[j][i]  -> D[m][n] = W[n][m]
    const int combos[][4] = {{1024, 128, 2048, 0}, {128, 1024, 128, 1}, {1024, 128, 128, 1}, {128, 1024, 2048, 1}, {1024, 128, 2048, 1},
                             {16, 1024, 128, 1}, {1024, 16, 128, 1}, {128, 16, 1024, 1}, {16, 128, 1024, 1}};
    for (auto& cb : combos) {
        probe3<<<1, 128, 64 * 64 * 4>>>(0, cb[0], cb[1], cb[2], cb[3], dD); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(hj, dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        probe3<<<1, 128, 64 * 64 * 4>>>(1, cb[0], cb[1], cb[2], cb[3], dD); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(hi, dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        int ok = 0, okT = 0;
        for (int m = 0; m < 64; ++m) for (int n = 0; n < 64; ++n) {
            ok += (hj[m * 64 + n] == (float)m && hi[m * 64 + n] == (float)n);      // transposed read: fetched W[j = k][i = n]
            okT += (hj[m * 64 + n] == (float)n && hi[m * 64 + n] == (float)m);     // K-major read:   fetched W[j = n][i = k]
        }
        printf("b_mn %d LBO %4d SBO %4d kstep %4d : transposed-correct %4d, kmajor-correct %4d of 4096; (k,n)->(j,i): ", cb[3], cb[0], cb[1], cb[2], ok, okT);
        const int probes[][2] = {{0, 1}, {0, 4}, {1, 0}, {2, 0}, {8, 0}, {9, 4}, {17, 33}};
        for (auto& p : probes) printf("(%d,%d)->(%g,%g) ", p[0], p[1], hj[p[0] * 64 + p[1]], hi[p[0] * 64 + p[1]]);
        printf("\n");
    }
    return 0;
}
