// ls2fm_field.cuh -- fused hash-grid + geometry-MLP (+ normals, + radiance) kernels.
//
// Replaces, per sample point, the reference's eager graph
//   Embedder_Hash.forward (models/base.py:23-40) -> tcnn grid [EXT] -> Geometry.forward (base.py:206-217)
//   SDF.infer_sdf (models/SDF.py:55-78), SDF.gradient (SDF.py:102-114, create_graph=True),
//   Fourier view embedding (base.py:75-97), the radiance input concat (Renderer.py:75) and
//   Radiance.forward (base.py:249-261)
// with one forward kernel and one backward kernel.  Nothing per-sample but the values the
// reference itself returns (sdf, normal, rgb) ever reaches HBM: encodings, hidden activations
// and their adjoints live in shared memory / registers only.
//
// Work decomposition
//   * a WARP owns a tile of LS_WS = 8 consecutive samples.
//   * gather phases map lane -> (sample = lane & 7, level group g = lane >> 3; levels g, g+4, g+8, g+12):
//     the 8 lanes of one level touch neighbouring cells of the table (samples are consecutive on a ray).
//   * matrix phases map lane -> (og = lane & 15, sg = lane >> 4): a 4-sample x 4-output register block,
//     activations in shared memory in "panel" layout  addr(row, sg, s) = ((sg * R + row) * 4 + s)
//     so a lane reads its 4 samples of one row with a single LDS.128 and the 32 lanes of a warp touch
//     only two distinct addresses (broadcast), weights come as LDS.128 rows of the transposed matrix.
//   * normals: reverse sweep for output 0 (one extra pass per layer) instead of 3 forward tangents;
//     its adjoint in the backward kernel is a forward *tangent* pass along the upstream normal
//     gradient, so the backward is an ordinary reverse pass over a 2-channel (primal, tangent) net.
//   * parameter gradients: the backward kernel is CTA-synchronous (8 warps = 64 samples per step);
//     each thread owns a fixed interleaved 4x4 block of every layer's weight-gradient matrix in
//     registers for the whole (persistent) CTA lifetime and flushes it once with atomics.
#pragma once

#include "ls2fm_common.cuh"
#include "ls2fm_tc.cuh"

// ---------------------------------------------------------------- kernel-side parameter blocks
struct LsNet {           // shared-memory placement of the staged MLP (floats)
    int sw_off[LS2FM_MAX_LAYERS];
    int sb_off[LS2FM_MAX_LAYERS];
    int pitch[LS2FM_MAX_LAYERS];
    int gw_off[LS2FM_MAX_LAYERS];   // offsets inside theta (global)
    int gb_off[LS2FM_MAX_LAYERS];
    int weff_off;        // radiance W_eff [3][rad_pitch] then b_eff[4]
    int rad_pitch;
    int warp_base;       // first float of the per-warp buffers
    int warp_stride;     // floats per warp
    int tc_misc;         // backward kernel: mbarrier + TMEM slot (8 floats)
    int total;           // floats of dynamic shared memory
};

struct LsFieldArgs {
    ls2fm_field_t f;
    ls2fm_points_t p;
    ls2fm_radiance_t r;   // r.w_eff == NULL: no radiance
    LsNet net;
    float inv_ext[3];     // 1 / (bound_max - bound_min)
    float s;              // sdf = s * y0  (sign / scale_mlp)
    // forward outputs (nullable)
    float* out_y; float* out_sdf; float* out_nrm; float* out_rgb;
    // backward inputs (nullable) / outputs
    const float* g_y; const float* g_sdf; const float* g_nrm; const float* g_rgb;
    const float* saved_nrm; const float* saved_rgb;
    float* d_table; float* d_theta; float* d_w_eff; float* d_b_eff; float* d_geo2;
    ls2fm_input_grads_t ig;   // gradients w.r.t. the sample positions (all NULL: not wanted)
    int dbg;                  // LS_ABLATE builds only (tools/ablate.py): bit 0 no scatter, 1 no gather loads, 2 no weight-gradient MMAs
};
#if defined(LS_ABLATE)
#define LS_DBG(a, bit) (((a).dbg >> (bit)) & 1)
#else
#define LS_DBG(a, bit) 0
#endif

inline int ls_round4(int v) { return (v + 3) & ~3; }

// Computes the shared-memory plan.  backward: per-warp buffers of the backward kernel.
inline LsNet ls_plan_net(const ls2fm_field_t& f, int rad_in_dim, int n_warps, bool backward) {
    LsNet n;
    memset(&n, 0, sizeof(n));
    const LsThetaLayout tl = ls_theta_layout(f);
    int off = 0;
    for (int l = 0; l < f.n_layers; ++l) {
        const bool last = l == f.n_layers - 1;
        n.pitch[l] = last ? LS_OPITCH : LS_WPITCH;
        n.sw_off[l] = off; off += f.dims[l] * n.pitch[l];
        n.sb_off[l] = off; off += last ? LS_OROWS : LS_H;
        n.gw_off[l] = tl.w_off[l];
        n.gb_off[l] = tl.b_off[l];
    }
    n.rad_pitch = ls_round4(rad_in_dim > 0 ? rad_in_dim : 4);
    n.weff_off = off; off += 3 * n.rad_pitch + 4;
    off = ls_round4(off);
    n.warp_base = off;
    const int hid = f.n_layers - 1;
    if (!backward) n.warp_stride = LS_WS * (LS_EROWS + LS_H * hid + LS_OROWS);
    else n.warp_stride = LS_WS * (2 * LS_EROWS + 2 * LS_H * hid + 2 * LS_H);
    n.tc_misc = off + n.warp_stride * n_warps;
    n.total = n.tc_misc + 8;
    return n;
}

// ---------------------------------------------------------------- small helpers
LS_DEV float4 ls_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
LS_DEV void ls_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
LS_DEV float* ls_row(float* buf, int R, int sg, int row) { return buf + (sg * R + row) * 4; }
LS_DEV const float* ls_row(const float* buf, int R, int sg, int row) { return buf + (sg * R + row) * 4; }
// element (row, sample s8 of the warp tile)
LS_DEV float& ls_el(float* buf, int R, int row, int s8) { return buf[((s8 >> 2) * R + row) * 4 + (s8 & 3)]; }

#define LS_FMA16(acc, w, a)                                                         \
    do {                                                                            \
        acc[0][0] = fmaf(w.x, a.x, acc[0][0]); acc[0][1] = fmaf(w.x, a.y, acc[0][1]); \
        acc[0][2] = fmaf(w.x, a.z, acc[0][2]); acc[0][3] = fmaf(w.x, a.w, acc[0][3]); \
        acc[1][0] = fmaf(w.y, a.x, acc[1][0]); acc[1][1] = fmaf(w.y, a.y, acc[1][1]); \
        acc[1][2] = fmaf(w.y, a.z, acc[1][2]); acc[1][3] = fmaf(w.y, a.w, acc[1][3]); \
        acc[2][0] = fmaf(w.z, a.x, acc[2][0]); acc[2][1] = fmaf(w.z, a.y, acc[2][1]); \
        acc[2][2] = fmaf(w.z, a.z, acc[2][2]); acc[2][3] = fmaf(w.z, a.w, acc[2][3]); \
        acc[3][0] = fmaf(w.w, a.x, acc[3][0]); acc[3][1] = fmaf(w.w, a.y, acc[3][1]); \
        acc[3][2] = fmaf(w.w, a.z, acc[3][2]); acc[3][3] = fmaf(w.w, a.w, acc[3][3]); \
    } while (0)

// acc[q][s] += sum_k Wt[k][4 og + q] * in[k][s]     (natural ownership: outputs 4og .. 4og+3)
template <int NCH>
LS_DEV void ls_prod_fwd(const float* Wt, int pitch, int n_in, const float* in0, const float* in1, int R,
                        int sg, int og, float (&acc0)[4][4], float (&acc1)[4][4]) {
    const float* wp = Wt + 4 * og;
    const float* p0 = in0 + sg * R * 4;
    const float* p1 = in1 + sg * R * 4;
#pragma unroll 4
    for (int k = 0; k < n_in; ++k) {
        const float4 w = ls_ld4(wp + k * pitch);
        const float4 a = ls_ld4(p0 + 4 * k);
        LS_FMA16(acc0, w, a);
        if (NCH == 2) {
            const float4 b = ls_ld4(p1 + 4 * k);
            LS_FMA16(acc1, w, b);
        }
    }
}

// acc[q][s] = sum_j Wt[og + 16 q][j] * z[j][s]      (interleaved ownership: inputs og, og+16, og+32, og+48)
// n_j: number of valid j (multiple of 4); rows >= n_rows are clamped (their results are discarded).
template <int NCH, int NQ>
LS_DEV void ls_prod_rev(const float* Wt, int pitch, int n_rows, int n_j, const float* z0, const float* z1, int R,
                        int sg, int og, float (&acc0)[4][4], float (&acc1)[4][4]) {
    const float* wr[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        int row = og + 16 * q;
        row = row < n_rows ? row : n_rows - 1;
        wr[q] = Wt + row * pitch;
    }
    const float* p0 = z0 + sg * R * 4;
    const float* p1 = z1 + sg * R * 4;
#pragma unroll 2
    for (int j = 0; j < n_j; j += 4) {
        float4 w[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) w[q] = ls_ld4(wr[q] + j);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float4 a = ls_ld4(p0 + 4 * (j + jj));
            float4 b = a;
            if (NCH == 2) b = ls_ld4(p1 + 4 * (j + jj));
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float wv = jj == 0 ? w[q].x : jj == 1 ? w[q].y : jj == 2 ? w[q].z : w[q].w;
                acc0[q][0] = fmaf(wv, a.x, acc0[q][0]); acc0[q][1] = fmaf(wv, a.y, acc0[q][1]);
                acc0[q][2] = fmaf(wv, a.z, acc0[q][2]); acc0[q][3] = fmaf(wv, a.w, acc0[q][3]);
                if (NCH == 2) {
                    acc1[q][0] = fmaf(wv, b.x, acc1[q][0]); acc1[q][1] = fmaf(wv, b.y, acc1[q][1]);
                    acc1[q][2] = fmaf(wv, b.z, acc1[q][2]); acc1[q][3] = fmaf(wv, b.w, acc1[q][3]);
                }
            }
        }
    }
}

// stage the effective MLP weights (and W_eff of the radiance decoder) into shared memory
LS_DEV void ls_stage_weights(const LsFieldArgs& a, float* smem) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int K = a.f.n_layers;
    for (int l = 0; l < K; ++l) {
        const int din = a.f.dims[l], dout = a.f.dims[l + 1], pitch = a.net.pitch[l];
        float* W = smem + a.net.sw_off[l];
        const float* G = a.f.theta + a.net.gw_off[l];
        for (int e = tid; e < din * pitch; e += nt) {
            const int k = e / pitch, j = e - k * pitch;
            W[e] = j < dout ? __ldg(G + k * dout + j) : 0.f;
        }
        float* B = smem + a.net.sb_off[l];
        const int nb = l == K - 1 ? LS_OROWS : LS_H;
        for (int e = tid; e < nb; e += nt) B[e] = e < dout ? __ldg(a.f.theta + a.net.gb_off[l] + e) : 0.f;
    }
    if (a.r.w_eff) {
        float* W = smem + a.net.weff_off;
        const int P = a.net.rad_pitch;
        for (int e = tid; e < 3 * P; e += nt) {
            const int c = e / P, i = e - c * P;
            W[e] = i < a.r.in_dim ? __ldg(a.r.w_eff + c * a.r.in_dim + i) : 0.f;
        }
        for (int e = tid; e < 4; e += nt) W[3 * P + e] = e < 3 ? __ldg(a.r.b_eff + e) : 0.f;
    }
}

// sample position: explicit xyz or center + ray * t (mul, then add: utils/camera.py:262-266 / torch eager).
// *out_i: where the per-sample outputs of sample i go (dense, or r * out_stride + out_offset + j for compacted rays)
LS_DEV void ls_sample_point(const ls2fm_points_t& p, int64_t i, float x[3], int* ray_id, int64_t* out_i) {
    *out_i = i;
    // (sample index -> ray: a 32-bit division whenever the launch has fewer than 2^31 samples -- a warp-uniform test; the 64-bit
    //  one is a ~100-instruction subroutine that every thread of the forward kernels paid once per tile)
    const bool small = p.n <= 0x7fffffffLL;
    if (p.xyz) {
        x[0] = __ldg(p.xyz + 3 * i); x[1] = __ldg(p.xyz + 3 * i + 1); x[2] = __ldg(p.xyz + 3 * i + 2);
        *ray_id = p.n_per_ray > 0 ? (small ? (int)((uint32_t)i / (uint32_t)p.n_per_ray) : (int)(i / p.n_per_ray)) : 0;
    } else {
        int r = small ? (int)((uint32_t)i / (uint32_t)p.n_per_ray) : (int)(i / p.n_per_ray);
        const int j = (int)(i - (int64_t)r * p.n_per_ray);
        if (p.ray_index) r = __ldg(p.ray_index + r);
        const float t = __ldg(p.t + (int64_t)r * p.t_stride + p.t_offset + j);
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = ls_fadd(__ldg(p.center + 3 * r + d), ls_fmul(__ldg(p.ray + 3 * r + d), t));
        *ray_id = r;
        if (p.out_stride > 0) *out_i = (int64_t)r * p.out_stride + p.out_offset + j;
    }
}
// number of samples to process: all, or (device-side count of compacted rays) * n_per_ray
LS_DEV int64_t ls_n_samples(const ls2fm_points_t& p) {
    if (p.n_active) {
        const int64_t m = (int64_t)__ldg(p.n_active) * p.n_per_ray;
        return m < p.n ? m : p.n;
    }
    return p.n;
}

// one level of the hash grid at u: features h[2] and dh/du [2][3]
LS_DEV void ls_level_eval(const ls2fm_field_t& f, int l, const float u[3], float h[2], float dh[2][3], int nogather = 0) {
    const float scale = f.levels[l].scale;
    const uint32_t res = f.levels[l].resolution, size = f.levels[l].size, hashed = f.levels[l].hashed;
    const float* tab = f.table + 2 * (size_t)f.levels[l].offset;
    const LsCell c = ls_cell(scale, u);
    float2 v[8];
    uint32_t ci[8];
    ls_corner_indices<0, 8>(res, size, hashed, c, ci);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = nogather ? make_float2(0.01f * (float)(ci[k] & 7), 0.02f) : __ldg(reinterpret_cast<const float2*>(tab) + ci[k]);
    const float w0 = c.w[0], w1 = c.w[1], w2 = c.w[2];
    const float m0 = 1.f - w0, m1 = 1.f - w1, m2 = 1.f - w2;
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        float q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = fi == 0 ? v[k].x : v[k].y;
        // interpolate along x first
        const float a00 = m0 * q[0] + w0 * q[1], a10 = m0 * q[2] + w0 * q[3];
        const float a01 = m0 * q[4] + w0 * q[5], a11 = m0 * q[6] + w0 * q[7];
        const float d00 = q[1] - q[0], d10 = q[3] - q[2], d01 = q[5] - q[4], d11 = q[7] - q[6];
        const float b0 = m1 * a00 + w1 * a10, b1 = m1 * a01 + w1 * a11;
        h[fi] = m2 * b0 + w2 * b1;
        dh[fi][0] = scale * (m2 * (m1 * d00 + w1 * d10) + w2 * (m1 * d01 + w1 * d11));
        dh[fi][1] = scale * (m2 * (a10 - a00) + w2 * (a11 - a01));
        dh[fi][2] = scale * (b1 - b0);
    }
}

// MLP forward of one warp tile: E (encoding rows) -> A_k = softplus(z_k) (k = 1..K-1, kept for the reverse sweep) -> Y
LS_DEV void ls_mlp_forward(const LsFieldArgs& a, const float* smem, const float* E, float* A, float* Y, int og, int sg) {
    const int K = a.f.n_layers;
    const float sp_beta = a.f.softplus_beta, sp_thr = a.f.softplus_threshold;
    const float* in = E;
    int R = LS_EROWS, n_in = a.f.dims[0];
    for (int l = 0; l < K - 1; ++l) {
        float acc[4][4], dummy[4][4];
        const float4 b = ls_ld4(smem + a.net.sb_off[l] + 4 * og);
#pragma unroll
        for (int s = 0; s < 4; ++s) { acc[0][s] = b.x; acc[1][s] = b.y; acc[2][s] = b.z; acc[3][s] = b.w; }
        ls_prod_fwd<1>(smem + a.net.sw_off[l], a.net.pitch[l], n_in, in, in, R, sg, og, acc, dummy);
        float* out = A + l * LS_WS * LS_H;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 o;
            o.x = ls_softplus(acc[q][0], sp_beta, sp_thr); o.y = ls_softplus(acc[q][1], sp_beta, sp_thr);
            o.z = ls_softplus(acc[q][2], sp_beta, sp_thr); o.w = ls_softplus(acc[q][3], sp_beta, sp_thr);
            ls_st4(ls_row(out, LS_H, sg, 4 * og + q), o);
        }
        __syncwarp();
        in = out; R = LS_H; n_in = LS_H;
    }
    if (og < LS_OROWS / 4) {
        float acc[4][4], dummy[4][4];
        const float4 b = ls_ld4(smem + a.net.sb_off[K - 1] + 4 * og);
#pragma unroll
        for (int s = 0; s < 4; ++s) { acc[0][s] = b.x; acc[1][s] = b.y; acc[2][s] = b.z; acc[3][s] = b.w; }
        ls_prod_fwd<1>(smem + a.net.sw_off[K - 1], a.net.pitch[K - 1], n_in, in, in, R, sg, og, acc, dummy);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            ls_st4(ls_row(Y, LS_OROWS, sg, 4 * og + q), make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]));
    }
    __syncwarp();
}

// hash-grid features of x into the encoding rows of a warp tile (no Jacobian): lane = (sample s8, level group g)
LS_DEV void ls_encode_tile(const LsFieldArgs& a, float* E, const float x[3], int s8, int g) {
    float u[3];
    ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int l = g + 4 * r;
        if (l < a.f.n_levels) {
            float h[2], dh[2][3];
            ls_level_eval(a.f, l, u, h, dh);
            ls_el(E, LS_EROWS, 3 + 2 * l, s8) = h[0];
            ls_el(E, LS_EROWS, 3 + 2 * l + 1, s8) = h[1];
        }
    }
    if (g == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) ls_el(E, LS_EROWS, d, s8) = ls_fdiv(x[d], a.f.rescale);
    }
    __syncwarp();
}

// ================================================================ forward kernel
// grid: persistent, blockDim = 32 * n_warps; every warp walks tiles independently.
__global__ void __launch_bounds__(512, 1) ls_field_forward_kernel(const LsFieldArgs a) {
    LS_DYN_SMEM(smem);
    if (ls_n_samples(a.p) == 0) return;
    ls_stage_weights(a, smem);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int K = a.f.n_layers, L = a.f.n_levels;
    const int dout = a.f.dims[K];
    const float sp_beta = a.f.softplus_beta;
    float* E = smem + a.net.warp_base + warp * a.net.warp_stride;
    float* A = E + LS_WS * LS_EROWS;                  // A_k = A + (k-1) * 8 * 64, k = 1..K-1
    float* Y = A + LS_WS * LS_H * (K - 1);
    const int s8 = lane & 7, g = lane >> 3;           // gather mapping
    const int og = lane & 15, sg = lane >> 4;         // matrix mapping
    const bool need_nrm = a.out_nrm != nullptr || a.r.w_eff != nullptr;

    const int64_t n_pts = ls_n_samples(a.p);
    const int64_t n_tiles = (n_pts + LS_WS - 1) / LS_WS;
    for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < n_tiles; tile += (int64_t)gridDim.x * nw) {
        // ------------------------------------------------ phase 1: points + hash-grid gather
        const int64_t i_in = tile * LS_WS + s8;
        const bool valid = i_in < n_pts;
        float x[3] = {0.f, 0.f, 0.f}, u[3];
        int ray_id = 0;
        int64_t i = i_in;
        if (valid) ls_sample_point(a.p, i_in, x, &ray_id, &i);
        ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u);
        float J[4][2][3];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int l = g + 4 * r;
            if (l < L) {
                float h[2], dh[2][3];
                ls_level_eval(a.f, l, u, h, dh);
                ls_el(E, LS_EROWS, 3 + 2 * l, s8) = h[0];
                ls_el(E, LS_EROWS, 3 + 2 * l + 1, s8) = h[1];
#pragma unroll
                for (int fi = 0; fi < 2; ++fi)
#pragma unroll
                    for (int d = 0; d < 3; ++d) J[r][fi][d] = dh[fi][d] * a.inv_ext[d];
            } else {
#pragma unroll
                for (int fi = 0; fi < 2; ++fi)
#pragma unroll
                    for (int d = 0; d < 3; ++d) J[r][fi][d] = 0.f;
            }
        }
        if (g == 0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) ls_el(E, LS_EROWS, d, s8) = ls_fdiv(x[d], a.f.rescale);
        }
        __syncwarp();

        // ------------------------------------------------ phase 2: MLP forward
        ls_mlp_forward(a, smem, E, A, Y, og, sg);

        // ------------------------------------------------ phase 3: reverse sweep for d(sdf)/dx
        float nrm[3] = {0.f, 0.f, 0.f};
        if (need_nrm) {
            {   // w_{K-1}[j] = phi'(z_{K-1}[j]) * s * W_{K-1}[0][j], in place over a_{K-1}
                float* AK = A + (K - 2) * LS_WS * LS_H;
                const float* WL = smem + a.net.sw_off[K - 1];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = og + 16 * q;
                    const float v = a.s * WL[j * a.net.pitch[K - 1]];
                    float* p = ls_row(AK, LS_H, sg, j);
                    float4 av = ls_ld4(p);
                    av.x = ls_softplus_d1_from_a(av.x, sp_beta) * v; av.y = ls_softplus_d1_from_a(av.y, sp_beta) * v;
                    av.z = ls_softplus_d1_from_a(av.z, sp_beta) * v; av.w = ls_softplus_d1_from_a(av.w, sp_beta) * v;
                    ls_st4(p, av);
                }
                __syncwarp();
            }
            for (int l = K - 2; l >= 0; --l) {
                float acc[4][4], dummy[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int s = 0; s < 4; ++s) acc[q][s] = 0.f;
                const float* Z = A + l * LS_WS * LS_H;         // w_{l+1}
                const int n_rows = a.f.dims[l];
                if (l > 0) ls_prod_rev<1, 4>(smem + a.net.sw_off[l], a.net.pitch[l], n_rows, LS_H, Z, Z, LS_H, sg, og, acc, dummy);
                else ls_prod_rev<1, 3>(smem + a.net.sw_off[l], a.net.pitch[l], n_rows, LS_H, Z, Z, LS_H, sg, og, acc, dummy);   // 35 rows < 48
                if (l > 0) {
                    float* AL = A + (l - 1) * LS_WS * LS_H;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float* p = ls_row(AL, LS_H, sg, og + 16 * q);
                        float4 av = ls_ld4(p);
                        av.x = ls_softplus_d1_from_a(av.x, sp_beta) * acc[q][0]; av.y = ls_softplus_d1_from_a(av.y, sp_beta) * acc[q][1];
                        av.z = ls_softplus_d1_from_a(av.z, sp_beta) * acc[q][2]; av.w = ls_softplus_d1_from_a(av.w, sp_beta) * acc[q][3];
                        ls_st4(p, av);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int row = og + 16 * q;
                        if (row < n_rows)
                            ls_st4(ls_row(E, LS_EROWS, sg, row), make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]));
                    }
                }
                __syncwarp();
            }
            // n = Je^T v0 : lanes (sample, level group) contract their register Jacobians, then butterfly
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int l = g + 4 * r;
                if (l < L) {
                    const float v0 = ls_el(E, LS_EROWS, 3 + 2 * l, s8), v1 = ls_el(E, LS_EROWS, 3 + 2 * l + 1, s8);
#pragma unroll
                    for (int d = 0; d < 3; ++d) nrm[d] += v0 * J[r][0][d] + v1 * J[r][1][d];
                }
            }
            if (g == 0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) nrm[d] += ls_el(E, LS_EROWS, d, s8) / a.f.rescale;
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                nrm[d] += __shfl_xor_sync(0xffffffffu, nrm[d], 8);
                nrm[d] += __shfl_xor_sync(0xffffffffu, nrm[d], 16);
            }
        }

        // ------------------------------------------------ phase 4: radiance (affine o sigmoid) + outputs
        if (a.r.w_eff) {
            const float* W = smem + a.net.weff_off;
            const int P = a.net.rad_pitch;
            const int nf = a.r.n_freq, kg = a.r.k_geo, kg2 = a.r.k_geo2;
            const int o_ray = 6, o_geo = 6 + 3 + 6 * nf, o_geo2 = o_geo + kg;
            float pr[3] = {0.f, 0.f, 0.f};
            float dir[3] = {0.f, 0.f, 0.f};
            if (valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) dir[d] = __ldg(a.p.ray + 3 * ray_id + d);
            }
            if (g == 0) {
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int c = 0; c < 3; ++c) pr[c] += W[c * P + d] * x[d] + W[c * P + o_ray + d] * dir[d];
            }
            if (g == 1) {
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int c = 0; c < 3; ++c) pr[c] += W[c * P + 3 + d] * nrm[d];
            }
            for (int k = g; k < nf; k += 4) {
                const float fr = (float)(1 << k);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float ang = dir[d] * fr;
                    const float sn = sinf(ang), cs = cosf(ang);
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        pr[c] += W[c * P + o_ray + 3 + 6 * k + d] * sn + W[c * P + o_ray + 6 + 6 * k + d] * cs;
                }
            }
            for (int k = g; k < kg; k += 4) {
                const float yv = ls_el(Y, LS_OROWS, 1 + k, s8);
#pragma unroll
                for (int c = 0; c < 3; ++c) pr[c] += W[c * P + o_geo + k] * yv;
            }
            if (kg2 > 0 && valid) {
                for (int k = g; k < kg2; k += 4) {
                    const float yv = __ldg(a.r.geo2 + i * (kg2 + 1) + 1 + k);
#pragma unroll
                    for (int c = 0; c < 3; ++c) pr[c] += W[c * P + o_geo2 + k] * yv;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                pr[c] += __shfl_xor_sync(0xffffffffu, pr[c], 8);
                pr[c] += __shfl_xor_sync(0xffffffffu, pr[c], 16);
            }
            if (g == 0 && valid && a.out_rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) a.out_rgb[3 * i + c] = ls_sigmoid(pr[c] + W[3 * P + c]);
            }
        }
        if (valid) {
            if (g == 0) {
                if (a.out_sdf) a.out_sdf[i] = a.s * ls_el(Y, LS_OROWS, 0, s8);
                if (a.out_nrm) { a.out_nrm[3 * i] = nrm[0]; a.out_nrm[3 * i + 1] = nrm[1]; a.out_nrm[3 * i + 2] = nrm[2]; }
            }
            if (a.out_y)
                for (int o = g; o < dout; o += 4) a.out_y[i * dout + o] = ls_el(Y, LS_OROWS, o, s8);
        }
        __syncwarp();
    }
}

// ================================================================ backward kernel
// blockDim = 256 (8 warps, 64 samples per CTA step), persistent grid.
constexpr int LS_BW_WARPS = 8;
constexpr int LS_BW_THREADS = 32 * LS_BW_WARPS;
constexpr int LS_IN_ROW0 = LS_OROWS;      // rows of the P|Pd area used for (pbar[3], in[...]) during the W_eff phase

template <bool TAN, int K>
__global__ void __launch_bounds__(LS_BW_THREADS, 1) ls_field_backward_kernel(const LsFieldArgs a) {
    LS_DYN_SMEM(smem);
    ls_stage_weights(a, smem);
    __syncthreads();
    // weight gradients of every layer but the last are accumulated by the tensor cores (tcgen05, 3xTF32) in TMEM for the
    // whole lifetime of the persistent CTA: D_l[j][i] += sum_s zbar_l[s][j] * a_{l-1}[s][i] (+ tangent channel).  The panel
    // layout of the activation buffers IS the K-major no-swizzle operand layout with K = sample (8-row x 16-byte core
    // matrices, LBO = rows*16 B between the two 4-sample panels of a warp, one K = 8 step per warp tile).
    LsTcBar* tc_bar = reinterpret_cast<LsTcBar*>(smem + a.net.tc_misc);
    const uint32_t tmem = ls_tc_alloc(reinterpret_cast<uint32_t*>(smem + a.net.tc_misc + 4));
    ls_tc_bar_init(tc_bar);
    uint32_t tc_phase = 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.f.n_levels;
    const int din0 = a.f.dims[0], dout = a.f.dims[K];
    const int hid = K - 1;
    const float sp_beta = a.f.softplus_beta, sp_thr = a.f.softplus_threshold;
    const int WSTR = a.net.warp_stride;
    float* WB = smem + a.net.warp_base;
    // per-warp buffer offsets (floats)
    const int oE = 0, oEd = LS_WS * LS_EROWS, oA = 2 * LS_WS * LS_EROWS;          // A_k at oA + (k-1) * 2 * 512, Ad_k right after A_k
    const int oP = oA + 2 * LS_WS * LS_H * hid, oPd = oP + LS_WS * LS_H;
    float* my = WB + warp * WSTR;
    float* E = my + oE; float* Ed = my + oEd; float* P = my + oP; float* Pd = my + oPd;
    const int s8 = lane & 7, g = lane >> 3;
    const int og = lane & 15, sg = lane >> 4;
    const int bj = tid & 15, bi = tid >> 4;           // weight-gradient block ownership (interleaved rows/cols)
    const bool rad = a.r.w_eff != nullptr;
    const int nf = a.r.n_freq, kg = a.r.k_geo, kg2 = a.r.k_geo2;
    const int o_ray = 6, o_geo = 6 + 3 + 6 * nf, o_geo2 = o_geo + kg;
    const int RP = a.net.rad_pitch;
    const float* Weff = smem + a.net.weff_off;

    // persistent register accumulators: bias gradients of every layer, weight gradient of the (small) output layer
    float wlast[4][4];
    float bacc[LS2FM_MAX_LAYERS];
#pragma unroll
    for (int l = 0; l < LS2FM_MAX_LAYERS; ++l) bacc[l] = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) wlast[p][q] = 0.f;
    float weff_acc = 0.f;

    const int64_t n_ctiles = (a.p.n + LS_WS * LS_BW_WARPS - 1) / (LS_WS * LS_BW_WARPS);
    for (int64_t ct = blockIdx.x; ct < n_ctiles; ct += gridDim.x) {
        const bool first_step = ct == (int64_t)blockIdx.x;
        // ------------------------------------------------ B1: upstream gradients of this sample
        const int64_t i_in = (ct * LS_BW_WARPS + warp) * LS_WS + s8;
        const bool valid = i_in < a.p.n;
        float x[3] = {0.f, 0.f, 0.f}, u[3];
        int ray_id = 0;
        int64_t i = i_in;
        if (valid) ls_sample_point(a.p, i_in, x, &ray_id, &i);
        ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u);
        float pbar[3] = {0.f, 0.f, 0.f};
        float nbar[3] = {0.f, 0.f, 0.f};
        if (valid) {
            if (rad && a.g_rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float rgb = __ldg(a.saved_rgb + 3 * i + c);
                    pbar[c] = __ldg(a.g_rgb + 3 * i + c) * rgb * (1.f - rgb);
                }
            }
            if (TAN) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float v = a.g_nrm ? __ldg(a.g_nrm + 3 * i + d) : 0.f;
                    if (rad) v += Weff[3 + d] * pbar[0] + Weff[RP + 3 + d] * pbar[1] + Weff[2 * RP + 3 + d] * pbar[2];
                    nbar[d] = v;
                }
            }
        }
        // ybar rows -> P[0..19]
        for (int o = g; o < LS_OROWS; o += 4) {
            float v = 0.f;
            if (valid && o < dout) {
                if (a.g_y) v += __ldg(a.g_y + i * dout + o);
                if (o == 0) { if (a.g_sdf) v += a.s * __ldg(a.g_sdf + i); }
                else if (rad && o - 1 < kg) {
                    const int col = o_geo + o - 1;
                    v += Weff[col] * pbar[0] + Weff[RP + col] * pbar[1] + Weff[2 * RP + col] * pbar[2];
                }
            }
            ls_el(P, LS_H, o, s8) = v;
        }
        // ------------------------------------------------ B2: gather, e and (TAN) edot = Je nbar
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            const int l = g + 4 * r;
            if (l < L) {
                float h[2], dh[2][3];
                ls_level_eval(a.f, l, u, h, dh);
                ls_el(E, LS_EROWS, 3 + 2 * l, s8) = h[0];
                ls_el(E, LS_EROWS, 3 + 2 * l + 1, s8) = h[1];
                if (TAN) {
#pragma unroll
                    for (int fi = 0; fi < 2; ++fi)
                        ls_el(Ed, LS_EROWS, 3 + 2 * l + fi, s8) =
                            dh[fi][0] * a.inv_ext[0] * nbar[0] + dh[fi][1] * a.inv_ext[1] * nbar[1] + dh[fi][2] * a.inv_ext[2] * nbar[2];
                }
            }
        }
        if (g == 0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                ls_el(E, LS_EROWS, d, s8) = ls_fdiv(x[d], a.f.rescale);
                if (TAN) ls_el(Ed, LS_EROWS, d, s8) = nbar[d] / a.f.rescale;
            }
        }
        __syncwarp();

        // ------------------------------------------------ B3: forward, primal + tangent
        {
            const float* in = E; const float* ind = Ed;
            int R = LS_EROWS, n_in = din0;
            for (int l = 0; l < K - 1; ++l) {
                float acc[4][4], accd[4][4];
                const float4 b = ls_ld4(smem + a.net.sb_off[l] + 4 * og);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    acc[0][s] = b.x; acc[1][s] = b.y; acc[2][s] = b.z; acc[3][s] = b.w;
                    accd[0][s] = accd[1][s] = accd[2][s] = accd[3][s] = 0.f;
                }
                ls_prod_fwd<TAN ? 2 : 1>(smem + a.net.sw_off[l], a.net.pitch[l], n_in, in, ind, R, sg, og, acc, accd);
                float* out = my + oA + l * 2 * LS_WS * LS_H;
                float* outd = out + LS_WS * LS_H;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float av[4], dv[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const float z = acc[q][s], bz = z * sp_beta;
                        if (bz > sp_thr) { av[s] = z; dv[s] = accd[q][s]; }
                        else {
                            const float e = expf(bz);
                            av[s] = log1pf(e) / sp_beta;
                            if (TAN) dv[s] = accd[q][s] * __fdividef(e, e + 1.f);
                        }
                    }
                    ls_st4(ls_row(out, LS_H, sg, 4 * og + q), make_float4(av[0], av[1], av[2], av[3]));
                    if (TAN) ls_st4(ls_row(outd, LS_H, sg, 4 * og + q), make_float4(dv[0], dv[1], dv[2], dv[3]));
                }
                __syncwarp();
                in = out; ind = outd; R = LS_H; n_in = LS_H;
            }
            if (rad) {   // geo features y[1..kg] feed the W_eff gradient
                if (og < LS_OROWS / 4) {
                    float acc[4][4], dummy[4][4];
                    const float4 b = ls_ld4(smem + a.net.sb_off[K - 1] + 4 * og);
#pragma unroll
                    for (int s = 0; s < 4; ++s) { acc[0][s] = b.x; acc[1][s] = b.y; acc[2][s] = b.z; acc[3][s] = b.w; }
                    ls_prod_fwd<1>(smem + a.net.sw_off[K - 1], a.net.pitch[K - 1], n_in, in, in, R, sg, og, acc, dummy);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int o = 4 * og + q;
                        if (o >= 1 && o <= kg)   // Pd rows 3.. hold the "in" vector: row 3 + index
                            ls_st4(ls_row(Pd, LS_H, sg, 3 + o_geo + o - 1 - 6), make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]));
                    }
                }
            }
        }
        if (rad) {
            // remaining rows of the "in" vector.  Pd rows: 0..2 pbar | 3 + (idx - 6) for idx >= 6 (ray enc, geo, geo2);
            // x and the saved normal (idx 0..5) go to P rows 20..25.
            float dir[3] = {0.f, 0.f, 0.f};
            if (valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) dir[d] = __ldg(a.p.ray + 3 * ray_id + d);
            }
            if (g == 0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    ls_el(P, LS_H, LS_IN_ROW0 + d, s8) = x[d];
                    ls_el(Pd, LS_H, d, s8) = pbar[d];
                    ls_el(Pd, LS_H, 3 + d, s8) = dir[d];
                }
            }
            if (g == 1) {
#pragma unroll
                for (int d = 0; d < 3; ++d) ls_el(P, LS_H, LS_IN_ROW0 + 3 + d, s8) = valid ? __ldg(a.saved_nrm + 3 * i + d) : 0.f;
            }
            for (int k = g; k < nf; k += 4) {
                const float fr = (float)(1 << k);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float ang = dir[d] * fr;
                    ls_el(Pd, LS_H, 3 + 3 + 6 * k + d, s8) = sinf(ang);
                    ls_el(Pd, LS_H, 3 + 6 + 6 * k + d, s8) = cosf(ang);
                }
            }
            for (int k = g; k < kg2; k += 4)
                ls_el(Pd, LS_H, 3 + o_geo2 - 6 + k, s8) = valid ? __ldg(a.r.geo2 + i * (kg2 + 1) + 1 + k) : 0.f;
            if (a.d_geo2 && valid) {
                for (int k = g; k <= kg2; k += 4) {
                    float v = 0.f;
                    if (k >= 1) {
                        const int col = o_geo2 + k - 1;
                        v = Weff[col] * pbar[0] + Weff[RP + col] * pbar[1] + Weff[2 * RP + col] * pbar[2];
                    }
                    a.d_geo2[i * (kg2 + 1) + k] = v;
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------ B4: W_eff / b_eff gradient (CTA-wide)
        if (rad) {
            const int in_dim = a.r.in_dim;
            if (tid < 3 * in_dim + 3) {
                const int c = tid < 3 * in_dim ? tid / in_dim : tid - 3 * in_dim;
                const int idx = tid < 3 * in_dim ? tid - c * in_dim : -1;
                float acc = 0.f;
#pragma unroll 1
                for (int w = 0; w < LS_BW_WARPS; ++w) {
                    const float* wP = WB + w * WSTR + oP;
                    const float* wPd = WB + w * WSTR + oPd;
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const float4 pb = ls_ld4(ls_row(wPd, LS_H, h2, c));
                        float4 iv = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (idx >= 0) iv = idx < 6 ? ls_ld4(ls_row(wP, LS_H, h2, LS_IN_ROW0 + idx)) : ls_ld4(ls_row(wPd, LS_H, h2, 3 + idx - 6));
                        acc += pb.x * iv.x + pb.y * iv.y + pb.z * iv.z + pb.w * iv.w;
                    }
                }
                weff_acc += acc;
            }
            __syncthreads();
        }

        // ------------------------------------------------ B5: reverse pass, primal + tangent
        // constant tangent adjoint of the output layer: zbar_dot_K = s * e_0
        if (TAN) {
            for (int o = g; o < LS_OROWS; o += 4) ls_el(Pd, LS_H, o, s8) = o == 0 ? a.s : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int l = LS2FM_MAX_LAYERS - 1; l >= 0; --l) {
            if (l >= K) continue;
            const bool last = l == K - 1;
            const int n_in = a.f.dims[l];
            const int R_in = l == 0 ? LS_EROWS : LS_H;
            const int o_in = l == 0 ? oE : oA + (l - 1) * 2 * LS_WS * LS_H;
            const int o_ind = l == 0 ? oEd : o_in + LS_WS * LS_H;
            const int o_z = last ? oP : oA + l * 2 * LS_WS * LS_H;
            const int o_zd = last ? oPd : o_z + LS_WS * LS_H;
            const int n_zrows = last ? LS_OROWS : LS_H;
            // ---- weight gradient of hidden / input layers on the tensor cores (tcgen05, 3xTF32), one batch per channel.
            //      The raw fp32 tiles are read as tf32 by truncation ("hi"); each warp writes what the truncation drops of its
            //      own 8 samples ("lo") into P (zbar) and Pd (layer input); thread 0 issues lo*hi + hi*lo + hi*hi for the 8 warp
            //      tiles (K = 64 samples).  The batches run under the SIMT work that follows them.
#define LS_BW_TC_BATCH(CH)                                                                                           \
            if (!last) {                                                                                             \
                const int d_col = 64 * l, n_cols = l == 0 ? 48 : 64;                                                 \
                const int oz = (CH) ? o_zd : o_z, oi = (CH) ? o_ind : o_in;                                          \
                for (int e = lane; e < 2 * LS_H; e += 32) {                                                          \
                    float4 v = ls_ld4(my + oz + 4 * e);                                                              \
                    v.x = ls_tf32_lo(v.x); v.y = ls_tf32_lo(v.y); v.z = ls_tf32_lo(v.z); v.w = ls_tf32_lo(v.w);      \
                    ls_st4(P + 4 * e, v);                                                                            \
                }                                                                                                    \
                for (int e = lane; e < 2 * R_in; e += 32) {                                                          \
                    const int h2 = e / R_in, row = e - h2 * R_in;                                                    \
                    float4 v = ls_ld4(ls_row(my + oi, R_in, h2, row));                                               \
                    v.x = ls_tf32_lo(v.x); v.y = ls_tf32_lo(v.y); v.z = ls_tf32_lo(v.z); v.w = ls_tf32_lo(v.w);      \
                    ls_st4(ls_row(Pd, LS_H, h2, row), v);                                                            \
                }                                                                                                    \
                ls_fence_smem_to_async();                                                                            \
                ls_tc_sync_before_mma();                                                                             \
                if (tid == 0) {                                                                                      \
                    for (int w = 0; w < LS_BW_WARPS; ++w) {                                                          \
                        const float* base = WB + w * WSTR;                                                           \
                        const bool acc0 = !(first_step && (CH) == 0 && w == 0);                                      \
                        ls_tc_mma_ss(tmem, d_col, base + oP, LS_H * 16, base + oi, R_in * 16, n_cols, acc0);         \
                        ls_tc_mma_ss(tmem, d_col, base + oz, LS_H * 16, base + oPd, LS_H * 16, n_cols, true);        \
                        ls_tc_mma_ss(tmem, d_col, base + oz, LS_H * 16, base + oi, R_in * 16, n_cols, true);         \
                    }                                                                                                \
                    ls_tc_commit(tc_bar);                                                                            \
                }                                                                                                    \
            }
            LS_BW_TC_BATCH(0)
            // ---- (b) input adjoints  abar = W^T zbar, abar_dot = W^T zbar_dot  (own warp tile)
            float acc[4][4], accd[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int s = 0; s < 4; ++s) { acc[q][s] = 0.f; accd[q][s] = 0.f; }
            if (l > 0) ls_prod_rev<TAN ? 2 : 1, 4>(smem + a.net.sw_off[l], a.net.pitch[l], n_in, last ? LS_OROWS : LS_H,
                                                   my + o_z, my + o_zd, LS_H, sg, og, acc, accd);
            else ls_prod_rev<TAN ? 2 : 1, 3>(smem + a.net.sw_off[l], a.net.pitch[l], n_in, last ? LS_OROWS : LS_H,
                                             my + o_z, my + o_zd, LS_H, sg, og, acc, accd);
            if (TAN && !last) {
                ls_tc_wait(tc_bar, tc_phase);          // primal batch done: P / Pd are free again
                LS_BW_TC_BATCH(1)
            }
#undef LS_BW_TC_BATCH
            // ---- (a) bias gradient (all layers) and, for the output layer, the weight gradient on the SIMT pipes
            {
                float bsum = 0.f;
                const int brow = tid < n_zrows ? tid : n_zrows - 1;
                if (last) {
                    int rin[4], rz[2];
#pragma unroll
                    for (int p = 0; p < 4; ++p) rin[p] = bi + 16 * p;
#pragma unroll
                    for (int q = 0; q < 2; ++q) { int r = bj + 16 * q; rz[q] = r < n_zrows ? r : n_zrows - 1; }
#pragma unroll 1
                    for (int w = 0; w < LS_BW_WARPS; ++w) {
                        const float* base = WB + w * WSTR;
#pragma unroll 1
                        for (int h2 = 0; h2 < 2; ++h2) {
                            float4 iv[4], zv[2];
#pragma unroll
                            for (int p = 0; p < 4; ++p) iv[p] = ls_ld4(ls_row(base + o_in, R_in, h2, rin[p]));
#pragma unroll
                            for (int q = 0; q < 2; ++q) zv[q] = ls_ld4(ls_row(base + o_z, LS_H, h2, rz[q]));
#pragma unroll
                            for (int p = 0; p < 4; ++p)
#pragma unroll
                                for (int q = 0; q < 2; ++q)
                                    wlast[p][q] = fmaf(iv[p].w, zv[q].w, fmaf(iv[p].z, zv[q].z, fmaf(iv[p].y, zv[q].y, fmaf(iv[p].x, zv[q].x, wlast[p][q]))));
                            if (TAN) {
#pragma unroll
                                for (int p = 0; p < 4; ++p) iv[p] = ls_ld4(ls_row(base + o_ind, R_in, h2, rin[p]));
#pragma unroll
                                for (int q = 0; q < 2; ++q) zv[q] = ls_ld4(ls_row(base + o_zd, LS_H, h2, rz[q]));
#pragma unroll
                                for (int p = 0; p < 4; ++p)
#pragma unroll
                                    for (int q = 0; q < 2; ++q)
                                        wlast[p][q] = fmaf(iv[p].w, zv[q].w, fmaf(iv[p].z, zv[q].z, fmaf(iv[p].y, zv[q].y, fmaf(iv[p].x, zv[q].x, wlast[p][q]))));
                            }
                            const float4 bz = ls_ld4(ls_row(base + o_z, LS_H, h2, brow));
                            bsum += bz.x + bz.y + bz.z + bz.w;
                        }
                    }
                } else {
#pragma unroll 1
                    for (int w = 0; w < LS_BW_WARPS; ++w) {
                        const float* base = WB + w * WSTR;
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const float4 bz = ls_ld4(ls_row(base + o_z, LS_H, h2, brow));
                            bsum += bz.x + bz.y + bz.z + bz.w;
                        }
                    }
                }
                bacc[l] += bsum;
            }
            if (!last) ls_tc_wait(tc_bar, tc_phase);     // the tensor cores are done with the tiles of layer l (and with P / Pd)
            __syncthreads();     // everybody is done reading the activations of layer l
            // ---- (c) through the activation (in place) or out to the encoding adjoints
            if (l > 0) {
                float* AL = my + o_in; float* ALd = my + o_ind;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int row = og + 16 * q;
                    float4 av = ls_ld4(ls_row(AL, LS_H, sg, row));
                    float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (TAN) dv = ls_ld4(ls_row(ALd, LS_H, sg, row));
                    float zb[4], zd[4];
                    const float aa[4] = {av.x, av.y, av.z, av.w};
                    const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const float qv = LS_FAST_EXP(-sp_beta * aa[s]);   // 1 - phi'
                        const float d1 = 1.f - qv;
                        zb[s] = d1 * acc[q][s];
                        if (TAN) { zb[s] += sp_beta * qv * dd[s] * accd[q][s]; zd[s] = d1 * accd[q][s]; }
                    }
                    ls_st4(ls_row(AL, LS_H, sg, row), make_float4(zb[0], zb[1], zb[2], zb[3]));
                    if (TAN) ls_st4(ls_row(ALd, LS_H, sg, row), make_float4(zd[0], zd[1], zd[2], zd[3]));
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int row = og + 16 * q;
                    if (row < n_in) {
                        ls_st4(ls_row(P, LS_H, sg, row), make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]));
                        if (TAN) ls_st4(ls_row(Pd, LS_H, sg, row), make_float4(accd[q][0], accd[q][1], accd[q][2], accd[q][3]));
                    }
                }
            }
            __syncthreads();
        }

        // ------------------------------------------------ B6: hash-table gradient scatter (+ gradient w.r.t. the sample position)
        // dL/dx = Je^T ebar + (d(Je nbar)/dx)^T ebar_dot (+ radiance's direct x term): the table values are gathered once more and
        // contracted with the first / mixed second derivatives of the trilinear weights (diagonal second derivatives
        // of a trilinear cell are zero).  P rows = ebar, Pd rows = ebar_dot (left there by layer 0's reverse step).
        const bool want_dx = a.ig.d_xyz || a.ig.d_center || a.ig.d_ray || a.ig.d_t;
        float dx[3] = {0.f, 0.f, 0.f};
        if ((a.d_table || want_dx) && valid) {
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                const int l = g + 4 * r;
                if (l < L) {
                    const float scale = a.f.levels[l].scale;
                    const uint32_t res = a.f.levels[l].resolution, size = a.f.levels[l].size, hashed = a.f.levels[l].hashed;
                    float* tab = a.d_table + 2 * (size_t)a.f.levels[l].offset;
                    const LsCell c = ls_cell(scale, u);
                    const float e0 = ls_el(P, LS_H, 3 + 2 * l, s8), e1 = ls_el(P, LS_H, 3 + 2 * l + 1, s8);
                    uint32_t ci[8];
                    ls_corner_indices<0, 8>(res, size, hashed, c, ci);
                    float t0 = 0.f, t1 = 0.f, ns[3] = {0.f, 0.f, 0.f};
                    if (TAN) {
                        t0 = ls_el(Pd, LS_H, 3 + 2 * l, s8); t1 = ls_el(Pd, LS_H, 3 + 2 * l + 1, s8);
#pragma unroll
                        for (int d = 0; d < 3; ++d) ns[d] = nbar[d] * scale * a.inv_ext[d];
                    }
                    if (a.d_table) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float f0 = (k & 1) ? c.w[0] : 1.f - c.w[0];
                            const float f1 = (k & 2) ? c.w[1] : 1.f - c.w[1];
                            const float f2 = (k & 4) ? c.w[2] : 1.f - c.w[2];
                            const float wgt = f0 * f1 * f2;
                            float g0 = wgt * e0, g1 = wgt * e1;
                            if (TAN) {
                                const float dw = ((k & 1) ? ns[0] : -ns[0]) * f1 * f2 + ((k & 2) ? ns[1] : -ns[1]) * f0 * f2 +
                                                 ((k & 4) ? ns[2] : -ns[2]) * f0 * f1;
                                g0 += dw * t0; g1 += dw * t1;
                            }
                            atomicAdd(reinterpret_cast<float2*>(tab) + ci[k], make_float2(g0, g1));
                        }
                    }
                    if (want_dx) {
                        const float2* vt = reinterpret_cast<const float2*>(a.f.table + 2 * (size_t)a.f.levels[l].offset);
                        float lx[3] = {0.f, 0.f, 0.f};
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float2 v = __ldg(vt + ci[k]);
                            const float fk[3] = {(k & 1) ? c.w[0] : 1.f - c.w[0], (k & 2) ? c.w[1] : 1.f - c.w[1], (k & 4) ? c.w[2] : 1.f - c.w[2]};
                            const float sg3[3] = {(k & 1) ? 1.f : -1.f, (k & 2) ? 1.f : -1.f, (k & 4) ? 1.f : -1.f};
                            const float Ak = v.x * e0 + v.y * e1;
                            const float Bk = TAN ? v.x * t0 + v.y * t1 : 0.f;
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
                                float term = fk[d1] * fk[d2] * Ak;
                                if (TAN) term += (sg3[d1] * ns[d1] * fk[d2] + sg3[d2] * ns[d2] * fk[d1]) * Bk;
                                lx[d] += sg3[d] * term;
                            }
                        }
#pragma unroll
                        for (int d = 0; d < 3; ++d) dx[d] += lx[d] * scale * a.inv_ext[d];
                    }
                }
            }
        }
        if (want_dx) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                dx[d] += __shfl_xor_sync(0xffffffffu, dx[d], 8);
                dx[d] += __shfl_xor_sync(0xffffffffu, dx[d], 16);
            }
            if (g == 0 && valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    dx[d] += ls_el(P, LS_H, d, s8) / a.f.rescale;
                    if (rad) dx[d] += Weff[d] * pbar[0] + Weff[RP + d] * pbar[1] + Weff[2 * RP + d] * pbar[2];
                }
                float dirg[3] = {0.f, 0.f, 0.f};     // gradient through the Fourier embedding of the direction
                if (rad && a.ig.d_ray) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float dv = __ldg(a.p.ray + 3 * ray_id + d);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float acc = Weff[c * RP + o_ray + d];
                            for (int k = 0; k < nf; ++k) {
                                const float fr = (float)(1 << k);
                                acc += fr * (Weff[c * RP + o_ray + 3 + 6 * k + d] * cosf(dv * fr) - Weff[c * RP + o_ray + 6 + 6 * k + d] * sinf(dv * fr));
                            }
                            dirg[d] += acc * pbar[c];
                        }
                    }
                }
                if (a.p.xyz) {
                    if (a.ig.d_xyz) { a.ig.d_xyz[3 * i] = dx[0]; a.ig.d_xyz[3 * i + 1] = dx[1]; a.ig.d_xyz[3 * i + 2] = dx[2]; }
                    if (a.ig.d_ray && rad) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) atomicAdd(a.ig.d_ray + 3 * ray_id + d, dirg[d]);
                    }
                } else {
                    const int j = (int)(i_in - (int64_t)ray_id * a.p.n_per_ray);
                    const float tv = __ldg(a.p.t + (int64_t)ray_id * a.p.t_stride + a.p.t_offset + j);
                    if (a.ig.d_center) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) atomicAdd(a.ig.d_center + 3 * ray_id + d, dx[d]);
                    }
                    if (a.ig.d_ray) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) atomicAdd(a.ig.d_ray + 3 * ray_id + d, tv * dx[d] + dirg[d]);
                    }
                    if (a.ig.d_t) {
                        float dt = 0.f;
#pragma unroll
                        for (int d = 0; d < 3; ++d) dt += __ldg(a.p.ray + 3 * ray_id + d) * dx[d];
                        a.ig.d_t[i_in] = dt;
                    }
                }
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------ flush the parameter gradients
    if (a.d_theta) {
        // tensor-core accumulators: TMEM lane j (warps 0, 1 own lanes 0..63) holds row j = output unit, column i = input unit
        if (warp < 2 && n_ctiles > (int64_t)blockIdx.x) {
#pragma unroll 1
            for (int l = 0; l < K - 1; ++l) {
                const int n_in = a.f.dims[l], n_out = a.f.dims[l + 1];
                const int jrow = tid;                      // 0..63
#pragma unroll 1
                for (int c = 0; c < 64; c += 16) {
                    if (c < n_in) {
                        float v[16];
                        ls_tmem_ld(tmem, 64 * l + c, v, 16);
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (c + q < n_in && jrow < n_out) atomicAdd(a.d_theta + a.net.gw_off[l] + (c + q) * n_out + jrow, v[q]);
                    }
                }
            }
        }
        {
            const int l = K - 1;
            const int n_in = a.f.dims[l], n_out = a.f.dims[l + 1];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int ii = bi + 16 * p, jj = bj + 16 * q;
                    if (ii < n_in && jj < n_out) atomicAdd(a.d_theta + a.net.gw_off[l] + ii * n_out + jj, wlast[p][q]);
                }
        }
#pragma unroll
        for (int l = 0; l < LS2FM_MAX_LAYERS; ++l) {
            if (l >= K) continue;
            const int n_out = a.f.dims[l + 1];
            if (tid < n_out) atomicAdd(a.d_theta + a.net.gb_off[l] + tid, bacc[l]);
        }
    }
    if (rad && a.d_w_eff) {
        const int in_dim = a.r.in_dim;
        if (tid < 3 * in_dim) atomicAdd(a.d_w_eff + tid, weff_acc);
        else if (tid < 3 * in_dim + 3 && a.d_b_eff) atomicAdd(a.d_b_eff + (tid - 3 * in_dim), weff_acc);
    }
    ls_tc_dealloc(tmem);
}
