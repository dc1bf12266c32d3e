// ls2fm_tc.cuh -- thin wrappers over the sm_100a tensor-core path (tcgen05 + TMEM) used by the MLP core.
//
// Model (validated on B200 by tools/probe/tc_probe*.cu before any kernel was built on it):
//   * D[128 x N] (fp32) lives in TMEM: lane = row (= the sample owned by thread `lane`), column = output feature.
//   * A[128 x K] comes from TMEM as well (lane = row, one 32-bit column per K element) -- the epilogue threads write the
//     next layer's input there with tcgen05.st, so activations never touch shared memory in the forward kernel.
//   * B[N x K] comes from shared memory, K-major, no swizzle: 8-row x 16-byte core matrices,
//       addr(n, k) = (k/4) * (N*16) + (n/8)*128 + (n%8)*16 + (k%4)*4     (LBO = N*16 bytes, SBO = 128 bytes).
//     (MN-major / transposed reads of tf32 operands returned zeros in every descriptor variant probed, so transposed
//     products use an explicitly transposed copy of the weights instead.)
//   * kind::tf32 keeps 10 mantissa bits; fp32-level accuracy comes from the 3xTF32 split
//       a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi,   x_hi = tf32(x) (round to nearest), x_lo = x - x_hi,
//     measured 5e-7 relative on 64-long dot products (tools/probe/tc_probe.cu) vs 3e-4 for a single pass.
//   * one thread issues the MMAs; completion is signalled through an mbarrier (tcgen05.commit).
//
// Under LS_HOSTSIM (tests/hostsim) the same entry points are emulated on the CPU with the same data layouts
// (TMEM = a per-block [128][512] array, operands truncated to tf32), so the kernels' indexing is testable without a GPU.
#pragma once

#include "ls2fm_common.cuh"

constexpr int LS_TC_M = 128;        // rows (samples) per tile = threads per CTA of the tensor-core kernels
constexpr int LS_TMEM_COLS = 512;

// byte offset of element (row n, k) of a K-major no-swizzle operand with `rows` rows
LS_DEV int ls_op_off(int n, int k, int rows) { return (k >> 2) * (rows * 16) + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4; }

#if defined(LS_HOSTSIM)
// ------------------------------------------------------------------------------------------------ CPU emulation
struct LsTmemSim { float v[LS_TC_M][LS_TMEM_COLS]; };
inline LsTmemSim& ls_tmem_sim() { static LsTmemSim t; return t; }
inline float ls_tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
inline float ls_tf32_round(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x00001000u;               // round to nearest (ties away), like cvt.rna.tf32.f32
    u &= 0xFFFFE000u;
    float r; memcpy(&r, &u, 4); return r;
}
struct LsTcBar { int arrived; };
LS_DEV uint32_t ls_tc_alloc(uint32_t*) { __syncthreads(); return 0; }
LS_DEV void ls_tc_dealloc(uint32_t) { __syncthreads(); }
LS_DEV void ls_tc_bar_init(LsTcBar* b) { if (threadIdx.x == 0) b->arrived = 0; __syncthreads(); }
LS_DEV int ls_tc_row() { return (int)(((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31)); }
LS_DEV void ls_tmem_ld(uint32_t, int col, float* v, int n) { for (int i = 0; i < n; ++i) v[i] = ls_tmem_sim().v[ls_tc_row()][col + i]; }
LS_DEV void ls_tmem_st(uint32_t, int col, const float* v, int n) { for (int i = 0; i < n; ++i) ls_tmem_sim().v[ls_tc_row()][col + i] = v[i]; }
LS_DEV void ls_tmem_ld2x16(uint32_t t, int col1, float* v1, int col2, float* v2) { ls_tmem_ld(t, col1, v1, 16); ls_tmem_ld(t, col2, v2, 16); }
LS_DEV void ls_tc_sync_before_mma() { __syncthreads(); }
// D[:, d_col + n] (+)= sum_k A[:, a_col + k] * B[n][k]   for n < N, k < K  (called by ONE thread)
LS_DEV void ls_tc_mma(uint32_t, int d_col, int a_col, const float* B, int N, int K, bool accumulate) {
    LsTmemSim& t = ls_tmem_sim();
    for (int m = 0; m < LS_TC_M; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = accumulate ? t.v[m][d_col + n] : 0.f;
            for (int k = 0; k < K; ++k)
                acc += ls_tf32_trunc(t.v[m][a_col + k]) * ls_tf32_trunc(*(const float*)((const char*)B + ls_op_off(n, k, N)));
            t.v[m][d_col + n] = acc;
        }
}
// one K = 8 step with BOTH operands in shared memory (K-major, SBO = 128 B, per-operand LBO): D[m][d_col+n] (+)= sum_k A(m,k) B(n,k)
LS_DEV void ls_tc_mma_ss(uint32_t, int d_col, const float* A, int a_lbo, const float* B, int b_lbo, int N, bool accumulate) {
    LsTmemSim& t = ls_tmem_sim();
    for (int m = 0; m < LS_TC_M; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = accumulate ? t.v[m][d_col + n] : 0.f;
            for (int k = 0; k < 8; ++k) {
                const float av = *(const float*)((const char*)A + (k >> 2) * a_lbo + (m >> 3) * 128 + (m & 7) * 16 + (k & 3) * 4);
                const float bv = *(const float*)((const char*)B + (k >> 2) * b_lbo + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4);
                acc += ls_tf32_trunc(av) * ls_tf32_trunc(bv);
            }
            t.v[m][d_col + n] = acc;
        }
}
LS_DEV void ls_tc_commit(LsTcBar* b) { b->arrived += 1; }
LS_DEV void ls_tc_wait(LsTcBar*, uint32_t& phase) { __syncthreads(); phase ^= 1; }
LS_DEV bool ls_elect() { return (threadIdx.x & 31) == 0; }
// ---- counted mbarriers and named barriers for warp-specialised kernels (fibers yield while they spin)
struct LsMbar { short pending; short count; int phase; };
LS_DEV void ls_mb_init(LsMbar* b, int count) { b->pending = (short)count; b->count = (short)count; b->phase = 0; }
LS_DEV void ls_mb_arrive(LsMbar* b) { if (--b->pending == 0) { b->pending = b->count; b->phase ^= 1; } }
LS_DEV void ls_mb_commit(LsMbar* b) { ls_mb_arrive(b); }          // (emulated MMAs complete at issue)
LS_DEV void ls_mb_wait(LsMbar* b, uint32_t parity) { while ((uint32_t)(b->phase & 1) == parity) simt::yield(simt::RUN); }
struct LsNamedBarSim { int count; int gen; };
inline LsNamedBarSim& ls_named_bar_sim(int id) { static LsNamedBarSim t[16]; return t[id]; }
LS_DEV void ls_named_sync(int id, int nthreads) {
    LsNamedBarSim& b = ls_named_bar_sim(id);
    const int g = b.gen;
    if (++b.count == nthreads) { b.count = 0; b.gen = g + 1; }
    else while (b.gen == g) simt::yield(simt::RUN);
}
LS_DEV void ls_ws_sync_before_mma(int id, int nthreads) { ls_named_sync(id, nthreads); }
LS_DEV float ls_tf32_lo(float x) { return x - ls_tf32_trunc(x); }
LS_DEV void ls_split_tf32(float v, float& hi, float& lo) { hi = ls_tf32_round(v); lo = v - hi; }
LS_DEV float ls_tf32_rna(float v) { return ls_tf32_round(v); }
LS_DEV void ls_fence_smem_to_async() {}
// bulk global -> shared copy completing on an mbarrier (emulated: plain copy at issue time, waits are no-ops)
LS_DEV void ls_bar_init1(LsTcBar* b) { b->arrived = 0; }
LS_DEV void ls_bulk_g2s(float* dst, const float* src, int bytes, LsTcBar*) { memcpy(dst, src, (size_t)bytes); }
LS_DEV void ls_bar_wait(LsTcBar*, uint32_t) {}
#else
// ------------------------------------------------------------------------------------------------ sm_100a
struct LsTcBar { unsigned long long v; };
LS_DEV uint32_t ls_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// warp 0 allocates all 512 columns; every thread returns the base address
LS_DEV uint32_t ls_tc_alloc(uint32_t* slot) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(ls_smem_u32(slot)), "n"(LS_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    return *slot;
}
LS_DEV void ls_tc_dealloc(uint32_t tmem) {
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "n"(LS_TMEM_COLS));
}
LS_DEV void ls_tc_bar_init(LsTcBar* b) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(ls_smem_u32(b)));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    __syncthreads();
}
// this thread's TMEM lane: a warp may only touch the lane quarter (warp_id % 4); warps w, w+4, w+8, ... share a quarter
LS_DEV int ls_tc_row() { return (int)(((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31)); }
LS_DEV uint32_t ls_tmem_lane_addr(uint32_t tmem, int col) { return tmem + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) + (uint32_t)col; }
LS_DEV void ls_tmem_ld16(uint32_t addr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
LS_DEV void ls_tmem_ld8(uint32_t addr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
LS_DEV void ls_tmem_st8(uint32_t addr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                 :: "r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
// two 16-column loads in flight behind ONE tcgen05.wait::ld
LS_DEV void ls_tmem_ld2x16(uint32_t tmem, int col1, float* v1, int col2, float* v2) {
    uint32_t r[16], q[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ls_tmem_lane_addr(tmem, col1)));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                   "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]) : "r"(ls_tmem_lane_addr(tmem, col2)));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) { v1[i] = __uint_as_float(r[i]); v2[i] = __uint_as_float(q[i]); }
}
// n must be a multiple of 8 (compile-time unrolled by the callers)
LS_DEV void ls_tmem_ld(uint32_t tmem, int col, float* v, int n) {
    int i = 0;
    for (; i + 16 <= n; i += 16) ls_tmem_ld16(ls_tmem_lane_addr(tmem, col + i), v + i);
    for (; i + 8 <= n; i += 8) ls_tmem_ld8(ls_tmem_lane_addr(tmem, col + i), v + i);
}
LS_DEV void ls_tmem_st(uint32_t tmem, int col, const float* v, int n) {
    for (int i = 0; i + 8 <= n; i += 8) ls_tmem_st8(ls_tmem_lane_addr(tmem, col + i), v + i);
}
// all threads: my TMEM stores are done and ordered before the barrier; after it one thread may issue MMAs that read them
LS_DEV void ls_tc_sync_before_mma() {
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
}
LS_DEV uint64_t ls_tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46);
}
// One lane of a converged warp.  MMA issue belongs inside   if (warp == W) { ...; if (ls_elect()) { issue; commit; } }   : under
// elect.sync the compiler keeps descriptors in uniform registers and emits back-to-back UTCHMMA; under  if (tid == 0)  it wraps
// every MMA in a 10-instruction R2UR waterfall (measured: ~50 cycles per MMA of pure issue).
LS_DEV bool ls_elect() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(p));
    return p != 0;
}
// D[:, d_col + n] (+)= sum_k A[:, a_col + k] * B[n][k]; A from TMEM, B K-major no-swizzle in smem; N % 16 == 0, K % 8 == 0
LS_DEV void ls_tc_mma(uint32_t tmem, int d_col, int a_col, const float* B, int N, int K, bool accumulate) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(LS_TC_M >> 4) << 24);
    uint64_t bdesc = ls_tc_desc(ls_smem_u32(B), N * 16, 128);
    const uint64_t bstep = (uint64_t)(2 * N);          // one K = 8 step = two 4-column groups of N rows x 16 B, in 16-byte units
    uint32_t a = tmem + (uint32_t)a_col;
    const uint32_t d = tmem + (uint32_t)d_col;
    for (int kb = 0; kb < K / 8; ++kb) {
        const uint32_t acc = (accumulate || kb > 0) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                     :: "r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc));
        bdesc += bstep;                                 // (only the 14-bit start-address field moves; it cannot carry out)
        a += 8;
    }
}
// one K = 8 step with BOTH operands in shared memory (K-major, SBO = 128 B, per-operand LBO)
LS_DEV void ls_tc_mma_ss(uint32_t tmem, int d_col, const float* A, int a_lbo, const float* B, int b_lbo, int N, bool accumulate) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(LS_TC_M >> 4) << 24);
    const uint64_t adesc = ls_tc_desc(ls_smem_u32(A), (uint32_t)a_lbo, 128);
    const uint64_t bdesc = ls_tc_desc(ls_smem_u32(B), (uint32_t)b_lbo, 128);
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem + (uint32_t)d_col), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
}
// what the tensor core drops when it reads an fp32 word as tf32 (truncation): x - trunc_tf32(x), exact in fp32
LS_DEV float ls_tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
LS_DEV void ls_tc_commit(LsTcBar* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(ls_smem_u32(b)) : "memory");
}
LS_DEV void ls_tc_wait(LsTcBar* b, uint32_t& phase) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(ls_smem_u32(b)), "r"(phase) : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n");
}
LS_DEV void ls_split_tf32(float v, float& hi, float& lo) {
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    hi = __uint_as_float(hb);
    lo = v - hi;
}
// ---- counted mbarriers and named barriers for warp-specialised kernels
struct LsMbar { unsigned long long v; };
LS_DEV void ls_mb_init(LsMbar* b, int count) {       // one thread, then fence + barrier by the caller
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(ls_smem_u32(b)), "r"(count));
}
LS_DEV void ls_mb_arrive(LsMbar* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(ls_smem_u32(b)) : "memory");
}
LS_DEV void ls_mb_commit(LsMbar* b) {                 // one arrival once every tcgen05 op issued so far by this thread is done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(ls_smem_u32(b)) : "memory");
}
LS_DEV void ls_mb_wait(LsMbar* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(ls_smem_u32(b)), "r"(parity) : "memory");
}
LS_DEV void ls_named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory"); }
// the MLP warps' version of ls_tc_sync_before_mma: a named barrier over themselves only
LS_DEV void ls_ws_sync_before_mma(int id, int nthreads) {
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    ls_named_sync(id, nthreads);
    asm volatile("tcgen05.fence::after_thread_sync;\n");
}
LS_DEV float ls_tf32_rna(float v) {
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    return __uint_as_float(hb);
}
LS_DEV void ls_fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// ---- bulk (TMA 1-D) global -> shared copy completing on an mbarrier; all three are called by ONE thread
LS_DEV void ls_bar_init1(LsTcBar* b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(ls_smem_u32(b)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
}
// bytes % 16 == 0, dst / src 16-byte aligned.  The barrier's phase completes when all bytes have landed.
LS_DEV void ls_bulk_g2s(float* dst, const float* src, int bytes, LsTcBar* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(ls_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(ls_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ls_smem_u32(bar)) : "memory");
}
LS_DEV void ls_bar_wait(LsTcBar* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(ls_smem_u32(b)), "r"(parity) : "memory");
}
#endif

// 3xTF32 product by the issuing thread: D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi (small terms first)
LS_DEV void ls_tc_mma_x3(uint32_t tmem, int d_col, int a_hi_col, int a_lo_col, const float* b_hi, const float* b_lo, int N, int K) {
    ls_tc_mma(tmem, d_col, a_lo_col, b_hi, N, K, false);
    ls_tc_mma(tmem, d_col, a_hi_col, b_lo, N, K, true);
    ls_tc_mma(tmem, d_col, a_hi_col, b_hi, N, K, true);
}
