// ls2fm_common.cuh -- shared definitions for the sm_100a kernels.
//
// The product library is built by nvcc for sm_100a only.  The same sources also compile with
// g++ -DLS_HOSTSIM against tests/hostsim/simt.h (a fiber-based SIMT emulator) so that the
// kernel arithmetic can be checked against the oracle in a container without a GPU; that is
// test infrastructure -- the python package never loads it and there is no CPU product path.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/ls2fm.h"

#if defined(LS_HOSTSIM)
#include "../../tests/hostsim/simt.h"
#define LS_HD inline
#define LS_DEV inline
#define LS_DYN_SMEM(name) float* name = reinterpret_cast<float*>(simt::S().smem)
#define LS_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define LS_FAST_EXP(x) expf(x)
#define LS_NOINLINE
#else
#include <cuda_runtime.h>
#define LS_HD __host__ __device__ __forceinline__
#define LS_DEV __device__ __forceinline__
#define LS_DYN_SMEM(name) extern __shared__ __align__(16) float name[]
#define LS_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
// MUFU-rate exp / log without __expf's denormal-range fix-up (two extra predicated multiplies and a compare per call): arguments
// below -126 flush to 0, which is what every user here wants (exp of a large negative number inside 1 - e, 1 + e, e / (1 + e)).
__device__ __forceinline__ float ls_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ls_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#define LS_FAST_EXP(x) ls_ex2((x) * 1.4426950408889634f)
#define LS_NOINLINE __device__ __noinline__
#endif

// ---------------------------------------------------------------- constants
constexpr int LS_H = LS2FM_HIDDEN;   // hidden width of the geometry MLP
constexpr int LS_WS = 8;             // samples per warp tile
constexpr int LS_EROWS = 36;         // rows of an encoding buffer (3 + 2*16 = 35 -> 36)
constexpr int LS_OROWS = LS2FM_MAX_OUT;  // rows of an output buffer (17 -> 20)
constexpr int LS_WPITCH = LS_H + 4;  // smem pitch of a [*, 64] weight matrix (== 4 mod 8: conflict-free rows)
constexpr int LS_OPITCH = LS2FM_MAX_OUT; // smem pitch of the [64, dout] output layer (20 == 4 mod 8)

// explicit single-rounding ops: the coordinate arithmetic must round exactly like the reference's
// eager torch ops (mul, then add -- never contracted into an fma) so hash cells agree bit for bit
LS_DEV float ls_fmul(float a, float b) { return __fmul_rn(a, b); }
LS_DEV float ls_fadd(float a, float b) { return __fadd_rn(a, b); }
LS_DEV float ls_fsub(float a, float b) { return __fsub_rn(a, b); }
LS_DEV float ls_fdiv(float a, float b) { return __fdiv_rn(a, b); }
LS_DEV float ls_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
LS_DEV double ls_dmul(double a, double b) { return __dmul_rn(a, b); }
LS_DEV double ls_dadd(double a, double b) { return __dadd_rn(a, b); }

// ---------------------------------------------------------------- softplus (beta, threshold)
// torch.nn.Softplus(beta=100, threshold=20) as the reference's Geometry MLP uses it (models/base.py:203).
LS_DEV float ls_softplus(float z, float beta, float thr) {
    const float bz = z * beta;
    return bz > thr ? z : log1pf(expf(bz)) / beta;
}
// phi'(z) recovered from a = phi(z):  exp(-beta a) = 1/(1+exp(beta z))  =>  phi' = 1 - exp(-beta a).
// In the threshold zone (a = z, beta z > 20) this gives 1 - 2e-9 == 1.0f, as autograd does.
// ex2.approx based: absolute error ~1e-7 on a value in [0, 1] -- fp32 rounding level for everything downstream.
LS_DEV float ls_softplus_d1_from_a(float a, float beta) { return 1.f - LS_FAST_EXP(-beta * a); }
LS_DEV float ls_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
// The tensor-core kernels' versions of the radiance tail: same functions on the MUFU units.
//   sin / cos: Cody-Waite reduction to [-pi, pi] (2 pi = hi + lo, both fma'd), then sin.approx / cos.approx: |error| < 5e-7 for the
//   arguments of the view-direction embedding (|v| 2^k < 16) -- sinf / cosf cost ~40 instructions each with a slow-path branch and
//   sat in only half of the warps (the Fourier lanes), holding everybody up at the next barrier.
//   sigmoid: ex2.approx + fast reciprocal, relative error ~2e-7.
LS_DEV void ls_sincos_fast(float x, float* s, float* c) {
#if defined(LS_HOSTSIM)
    *s = sinf(x); *c = cosf(x);
#else
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, x);
    r = fmaf(k, 1.7484555e-7f, r);
    *s = __sinf(r); *c = __cosf(r);
#endif
}
LS_DEV float ls_sigmoid_fast(float x) {
#if defined(LS_HOSTSIM)
    return 1.f / (1.f + expf(-x));
#else
    return __fdividef(1.f, 1.f + ls_ex2(-1.4426950408889634f * x));
#endif
}

// ---------------------------------------------------------------- hash grid (tcnn GridEncoding) [EXT]
struct LsCell {
    uint32_t g[3];   // (uint32_t)(int)floorf(p)
    float w[3];      // p - floorf(p)
};

LS_DEV LsCell ls_cell(float scale, const float u[3]) {
    LsCell c;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float p = ls_fma(scale, u[d], 0.5f);
        const float fl = floorf(p);
        c.g[d] = (uint32_t)(int)fl;
        c.w[d] = p - fl;
    }
    return c;
}

LS_DEV uint32_t ls_corner_index(uint32_t resolution, uint32_t size, uint32_t hashed, const LsCell& c, int corner) {
    const uint32_t q0 = c.g[0] + (corner & 1), q1 = c.g[1] + ((corner >> 1) & 1), q2 = c.g[2] + ((corner >> 2) & 1);
    uint32_t index;
    if (hashed) {
        index = q0 ^ (q1 * 2654435761u) ^ (q2 * 805459861u);
    } else {
        // dense level: all three strides fit (that is what hashed == 0 means); uint32 wrap as in tcnn
        index = q0 + q1 * resolution + q2 * resolution * resolution;
    }
    // tcnn: index % hashmap_size.  Hashed levels have power-of-two sizes (mask); dense levels are in range for every
    // point inside the box, so the (exact) modulo only runs for out-of-range coordinates.
    if (hashed && (size & (size - 1)) == 0) return index & (size - 1);
    return index < size ? index : index % size;
}

// All corner indices of a cell (or the 4 of one z-plane: corners first .. first + N - 1), deciding the path ONCE per level instead of
// once per corner: hashed power-of-two levels and in-range cells of dense levels are branch-free; anything else (out-of-range
// coordinates, odd table sizes) takes ls_corner_index corner by corner.  Bit-identical to ls_corner_index by construction: the
// same uint32 expressions, and for an in-range dense cell every index is < resolution^3 <= size, so tcnn's "% size" is the identity.
// (out of line: the rare path must not be replicated into every gather / scatter site of the big kernels)
LS_NOINLINE uint32_t ls_corner_index_slow(uint32_t resolution, uint32_t size, uint32_t hashed, uint32_t g0, uint32_t g1, uint32_t g2, int corner) {
    LsCell c;
    c.g[0] = g0; c.g[1] = g1; c.g[2] = g2;
    c.w[0] = c.w[1] = c.w[2] = 0.f;
    return ls_corner_index(resolution, size, hashed, c, corner);
}
template <int FIRST, int N>
LS_DEV void ls_corner_indices(uint32_t resolution, uint32_t size, uint32_t hashed, const LsCell& c, uint32_t (&idx)[N]) {
    if (hashed && (size & (size - 1)) == 0) {
        const uint32_t m = size - 1;
        const uint32_t y0 = c.g[1] * 2654435761u, y1 = (c.g[1] + 1) * 2654435761u;
        const uint32_t z0 = c.g[2] * 805459861u, z1 = (c.g[2] + 1) * 805459861u;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int corner = FIRST + k;
            idx[k] = ((c.g[0] + (corner & 1)) ^ ((corner & 2) ? y1 : y0) ^ ((corner & 4) ? z1 : z0)) & m;
        }
    } else if (!hashed && c.g[0] < resolution - 1 && c.g[1] < resolution - 1 && c.g[2] < resolution - 1) {
        const uint32_t r2 = resolution * resolution;
        const uint32_t base = c.g[0] + c.g[1] * resolution + c.g[2] * r2;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int corner = FIRST + k;
            idx[k] = base + (corner & 1) + ((corner & 2) ? resolution : 0u) + ((corner & 4) ? r2 : 0u);
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) idx[k] = ls_corner_index_slow(resolution, size, hashed, c.g[0], c.g[1], c.g[2], FIRST + k);
    }
}

// Table-gradient scatter of an x-neighbour corner pair (entries i0, i1 of one level; 2 floats each).  When the two entries form an
// aligned 16-byte pair -- always for an even cell of a hashed level (i1 == i0 ^ 1), for every other cell of a dense one -- ONE
// 16-byte vector atomic (red.global.add.v4.f32, sm_90+) replaces two 8-byte ones: the L1TEX / L2 atomic path costs the same per
// lane for either width (tools/probe/gather_probe.cu: 0.66 lane-atomics per SM per cycle for RED.64 and RED.128 alike), so the
// scatter's lane count -- what bounds it -- drops by a quarter.  tab must be 16-byte aligned for v4 (pass v4 = false otherwise).
LS_DEV void ls_red_pair(float* tab, uint32_t i0, uint32_t i1, float a0, float a1, float b0, float b1, bool v4) {
#if defined(LS_HOSTSIM)
    (void)v4;
    tab[2 * (size_t)i0] += a0; tab[2 * (size_t)i0 + 1] += a1;
    tab[2 * (size_t)i1] += b0; tab[2 * (size_t)i1 + 1] += b1;
#else
    if (v4 && ((i0 ^ i1) == 1u)) {
        const bool swap = (i0 & 1u) != 0u;
        float* p = tab + 2 * (size_t)(i0 & ~1u);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(swap ? b0 : a0), "f"(swap ? b1 : a1),
                     "f"(swap ? a0 : b0), "f"(swap ? a1 : b1) : "memory");
    } else {
        atomicAdd(reinterpret_cast<float2*>(tab) + i0, make_float2(a0, a1));
        atomicAdd(reinterpret_cast<float2*>(tab) + i1, make_float2(b0, b1));
    }
#endif
}

// world -> unit cube exactly as models/base.py:35: (x - bmin) / (bmax - bmin)
LS_DEV void ls_world_to_unit(const float bmin[3], const float bmax[3], const float x[3], float u[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = ls_fdiv(ls_fsub(x[d], bmin[d]), ls_fsub(bmax[d], bmin[d]));
}

// ---------------------------------------------------------------- ray / AABB (vren) [EXT]
// t_near/t_far of the slab test; (-1,-1) on a miss.  Mirrors oracle/aabb.py.
LS_DEV void ls_ray_aabb(const float o[3], const float d[3], const float center[3], const float half[3],
                        float* t_near, float* t_far) {
    float t1 = -INFINITY, t2 = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float inv = ls_fdiv(1.0f, d[k]);
        const float lo = ls_fmul(ls_fsub(ls_fsub(center[k], half[k]), o[k]), inv);
        const float hi = ls_fmul(ls_fsub(ls_fadd(center[k], half[k]), o[k]), inv);
        t1 = fmaxf(t1, fminf(lo, hi));
        t2 = fminf(t2, fmaxf(lo, hi));
    }
    if (t1 <= t2 && t2 > 0.f) { *t_near = fmaxf(t1, 0.f); *t_far = t2; }
    else { *t_near = -1.f; *t_far = -1.f; }
}

// Laplace-CDF density (models/SDF.py:84-87): sigma = alpha * (s >= 0 ? e : 1 - e), e = .5 exp(-|s|/beta)
LS_DEV float ls_sdf_to_sigma(float s, float alpha, float beta) {
    const float e = 0.5f * expf(-fabsf(s) / beta);
    return alpha * (s >= 0.f ? e : 1.f - e);
}

// ---------------------------------------------------------------- theta layout
// packed effective MLP parameters: for l: Wt_l [dims[l]][dims[l+1]] then b_l [dims[l+1]]
struct LsThetaLayout {
    int w_off[LS2FM_MAX_LAYERS];
    int b_off[LS2FM_MAX_LAYERS];
    int total;
};
inline LsThetaLayout ls_theta_layout(const ls2fm_field_t& f) {
    LsThetaLayout t;
    int off = 0;
    for (int l = 0; l < LS2FM_MAX_LAYERS; ++l) { t.w_off[l] = 0; t.b_off[l] = 0; }
    for (int l = 0; l < f.n_layers; ++l) {
        t.w_off[l] = off; off += f.dims[l] * f.dims[l + 1];
        t.b_off[l] = off; off += f.dims[l + 1];
    }
    t.total = off;
    return t;
}
