// tc_probe2.cu -- probes the remaining tcgen05 building blocks of the round-2 MLP core:
//  T1: A operand from TMEM (written with tcgen05.st), B K-major from smem                      D1[128x64] = X * W^T
//  T2: A from TMEM, B = the SAME smem weight array read MN-major (transposed product), N = 48   D2[128x48] = Z * W   (W is [64 x 48])
//  T3: M = 64, A and B both MN-major from smem tiles stored [sample][feature]:                 D3[64x64]  = Z^T * X  (K = 128 samples)
//  T4: N = 32 output layer                                                                      D4[128x32] = X * V^T
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// [rows][k] tile, K-major canonical no-swizzle layout (core matrix 8 rows x 16 B)
__device__ __host__ __forceinline__ int op_off(int r, int k, int rows) { return (k >> 2) * (rows * 16) + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;\n");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                 :: "r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// X [128][64] (samples x in), Z [128][64] (samples x out), W [64][48] (out x in, in padded 48), V [32][64]
__global__ void __launch_bounds__(128, 1) probe2(const float* X, const float* Z, const float* W, const float* V,
                                                 float* D1, float* D2, float* D3, float* D4) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Wfull = (float*)smem;                 // [64 out][64 in] for T1 (X * Wsq^T) -- use first 64 in-cols of a square matrix: reuse Z-side? keep simple:
    float* Ws = Wfull;                           // [64][48]   12 KB  (T2: MN-major read)
    float* Wsq = Ws + 64 * 48;                   // [64][64]   16 KB  (T1)  = V-like square weights built from X columns (see host)
    float* Vs = Wsq + 64 * 64;                   // [32][64]    8 KB  (T4)
    float* Xs = Vs + 32 * 64;                    // [128][64]  32 KB  (T3 B operand)
    float* Zs = Xs + 128 * 64;                   // [128][64]  32 KB  (T3 A operand)
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
    for (int e = tid; e < 64 * 48; e += 128) *(float*)((char*)Ws + op_off(e / 48, e % 48, 64)) = W[e];
    for (int e = tid; e < 64 * 64; e += 128) *(float*)((char*)Wsq + op_off(e / 64, e % 64, 64)) = Z[e];           // Wsq := first 64 rows of Z
    for (int e = tid; e < 32 * 64; e += 128) *(float*)((char*)Vs + op_off(e / 64, e % 64, 32)) = V[e];
    for (int e = tid; e < 128 * 64; e += 128) { *(float*)((char*)Xs + op_off(e / 64, e % 64, 128)) = X[e]; *(float*)((char*)Zs + op_off(e / 64, e % 64, 128)) = Z[e]; }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    // TMEM columns: [0,64) X as A operand, [64,128) Z as A operand, D1 [128,192), D2 [192,240), D3 [256,320), D4 [320,352)
    for (int c = 0; c < 64; c += 8) { float v[8], z[8]; for (int j = 0; j < 8; ++j) { v[j] = X[tid * 64 + c + j]; z[j] = Z[tid * 64 + c + j]; } tmem_st8(lane_addr + c, v); tmem_st8(lane_addr + 64 + c, z); }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        // T1: D1 = X * Wsq^T : A tmem cols 0.., B K-major [64 rows][K 64]: LBO = rows*16, SBO = 128
        for (int kb = 0; kb < 8; ++kb) mma_ts(tmem + 128, tmem + kb * 8, make_desc(smem_u32(Wsq) + kb * 2 * (64 * 16), 64 * 16, 128), make_idesc(128, 64, 0, 0), kb > 0);
        // T2: D2[128 x 48] = Z * W (W [64 out][48 in]): A = Z (tmem cols 64..), K = out index j; B[n = i][k = j] = W[j][i] read MN-major:
        //     n-chunk (4 i) stride = 64*16 B -> SBO ; k-group (8 j) stride = 128 B -> LBO ; k-step advances the start address by 128 B
        for (int kb = 0; kb < 8; ++kb) mma_ts(tmem + 192, tmem + 64 + kb * 8, make_desc(smem_u32(Ws) + kb * 128, 128, 64 * 16), make_idesc(128, 48, 0, 1), kb > 0);
        // T3: D3[64 (j) x 64 (i)] = sum_s Z[s][j] X[s][i]: A[m = j][k = s] = Z tile [128 s][64 j] read MN-major, B[n = i][k = s] = X tile read MN-major;
        //     tiles are [rows = s][K-major feature]: chunk (4 features) stride = 128*16 B -> SBO, k-group (8 s) stride 128 B -> LBO
        for (int kb = 0; kb < 16; ++kb) mma_ss(tmem + 256, make_desc(smem_u32(Zs) + kb * 128, 128, 128 * 16), make_desc(smem_u32(Xs) + kb * 128, 128, 128 * 16), make_idesc(64, 64, 1, 1), kb > 0);
        // T4: D4[128 x 32] = X * V^T
        for (int kb = 0; kb < 8; ++kb) mma_ts(tmem + 320, tmem + kb * 8, make_desc(smem_u32(Vs) + kb * 2 * (32 * 16), 32 * 16, 128), make_idesc(128, 32, 0, 0), kb > 0);
        commit(&bar);
    }
    wait_bar(&bar, 0);
    for (int c = 0; c < 64; c += 8) { float v[8]; tmem_ld8(lane_addr + 128 + c, v); for (int j = 0; j < 8; ++j) D1[tid * 64 + c + j] = v[j]; }
    for (int c = 0; c < 48; c += 8) { float v[8]; tmem_ld8(lane_addr + 192 + c, v); for (int j = 0; j < 8; ++j) D2[tid * 48 + c + j] = v[j]; }
    for (int c = 0; c < 64; c += 8) { float v[8]; tmem_ld8(lane_addr + 256 + c, v); for (int j = 0; j < 8; ++j) D3[tid * 64 + c + j] = v[j]; }   // lanes >= 64 hold garbage for M = 64
    for (int c = 0; c < 32; c += 8) { float v[8]; tmem_ld8(lane_addr + 320 + c, v); for (int j = 0; j < 8; ++j) D4[tid * 32 + c + j] = v[j]; }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "n"(512));
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
#include <string.h>
int main() {
    const int nX = 128 * 64, nW = 64 * 48, nV = 32 * 64;
    float *X = (float*)malloc(nX * 4), *Z = (float*)malloc(nX * 4), *W = (float*)malloc(nW * 4), *V = (float*)malloc(nV * 4);
    srand(2);
    // values exactly representable in tf32 so the single-pass products are exact up to accumulation order
    for (int i = 0; i < nX; ++i) { X[i] = tf32r((float)rand() / RAND_MAX * 2 - 1); Z[i] = tf32r((float)rand() / RAND_MAX * 2 - 1); }
    for (int i = 0; i < nW; ++i) W[i] = tf32r((float)rand() / RAND_MAX * 2 - 1);
    for (int i = 0; i < nV; ++i) V[i] = tf32r((float)rand() / RAND_MAX * 2 - 1);
    float *dX, *dZ, *dW, *dV, *d1, *d2, *d3, *d4;
    CK(cudaMalloc(&dX, nX * 4)); CK(cudaMalloc(&dZ, nX * 4)); CK(cudaMalloc(&dW, nW * 4)); CK(cudaMalloc(&dV, nV * 4));
    CK(cudaMalloc(&d1, 128 * 64 * 4)); CK(cudaMalloc(&d2, 128 * 48 * 4)); CK(cudaMalloc(&d3, 128 * 64 * 4)); CK(cudaMalloc(&d4, 128 * 32 * 4));
    CK(cudaMemcpy(dX, X, nX * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dZ, Z, nX * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W, nW * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dV, V, nV * 4, cudaMemcpyHostToDevice));
    const int smem = (64 * 48 + 64 * 64 + 32 * 64 + 2 * 128 * 64) * 4;
    CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe2<<<1, 128, smem>>>(dX, dZ, dW, dV, d1, d2, d3, d4);
    CK(cudaDeviceSynchronize());
    float *h1 = (float*)malloc(128 * 64 * 4), *h2 = (float*)malloc(128 * 48 * 4), *h3 = (float*)malloc(128 * 64 * 4), *h4 = (float*)malloc(128 * 32 * 4);
    CK(cudaMemcpy(h1, d1, 128 * 64 * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2, d2, 128 * 48 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h3, d3, 128 * 64 * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h4, d4, 128 * 32 * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { double r = 0; for (int k = 0; k < 64; ++k) r += (double)X[m * 64 + k] * Z[n * 64 + k]; e1 = fmax(e1, fabs(h1[m * 64 + n] - r)); }
    for (int m = 0; m < 128; ++m) for (int i = 0; i < 48; ++i) { double r = 0; for (int j = 0; j < 64; ++j) r += (double)Z[m * 64 + j] * W[j * 48 + i]; e2 = fmax(e2, fabs(h2[m * 48 + i] - r)); }
    for (int j = 0; j < 64; ++j) for (int i = 0; i < 64; ++i) { double r = 0; for (int s = 0; s < 128; ++s) r += (double)Z[s * 64 + j] * X[s * 64 + i]; e3 = fmax(e3, fabs(h3[j * 64 + i] - r)); }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) { double r = 0; for (int k = 0; k < 64; ++k) r += (double)X[m * 64 + k] * V[n * 64 + k]; e4 = fmax(e4, fabs(h4[m * 32 + n] - r)); }
    printf("T1 A-from-TMEM, B K-major      : max err %.3e\n", e1);
    printf("T2 A-from-TMEM, B MN-major N=48: max err %.3e\n", e2);
    printf("T3 M=64, A and B MN-major      : max err %.3e\n", e3);
    printf("T4 N=32                        : max err %.3e\n", e4);
    return 0;
}
