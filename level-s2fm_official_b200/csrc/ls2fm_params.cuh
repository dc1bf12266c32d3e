// ls2fm_params.cuh -- parameter preparation of the two fields in ONE launch (and its backward in one more).
//
// What the reference does implicitly inside every nn.Linear call -- old-style weight norm W = g * v / ||v||_row
// (models/base.py:200,241) -- plus what our kernels want on top of it:
//   * theta   : the geometry MLP's effective weights packed as (W_l^T row-major, b_l) per layer   (ls2fm_field_t.theta)
//   * W_eff   : the radiance decoder composed into one affine map W3 W2 W1 (the reference applies no hidden activation,
//               models/base.py:230,257), b_eff = W3 (W2 b1 + b2) + b3                             (ls2fm_radiance_t)
// Eager PyTorch needs ~30 tiny kernels for the forward of this and ~50 for its backward; here one CTA does each.
#pragma once

#include "ls2fm_common.cuh"
#include "ls2fm_render.cuh"   // ls_warp_sum

constexpr int LS_PP_MAX_IN = 68;     // radiance input width (<= LS2FM_MAX_RAD_IN)
constexpr int LS_PP_THREADS = 512;

typedef ls2fm_param_layer_t LsParamLayer;

struct LsParamArgs {
    LsParamLayer geo[LS2FM_MAX_LAYERS];
    int n_geo;                      // 0: skip the geometry MLP
    LsParamLayer rad[3];
    int has_rad;                    // radiance decoder with exactly 3 weight-normed layers: in -> 64 -> 64 -> 3
    float* theta; float* w_eff; float* b_eff;                        // forward outputs
    const float* d_theta; const float* d_w_eff; const float* d_b_eff;   // backward inputs (nullable)
    int accumulate;                 // backward: dg / dv / db += instead of = (every element has exactly one writer thread)
};
// (accumulate: a reduction that does not return -- fire and forget -- instead of a read-modify-write whose load misses L2 after the
//  step's table traffic and stalls every round of the row loops; one writer per element, so the result is the same)
LS_DEV void ls_pp_out(float* p, float v, int acc) { if (acc) atomicAdd(p, v); else *p = v; }

// The decoder has no hidden activation, so everything downstream of W1 has rank 3: with P = W3 W2 [3][64],
//   forward : W_eff = P W1,  b_eff = P b1 + W3 b2 + b3                                        (no [64][in] product at all)
//   backward: Q = dW_eff W1^T [3][64];  dW1 = P^T dW_eff,  dW2 = W3^T Q + dt (x) b1,  dW3 = Q W2^T + d_b_eff (x) t,
//             dt = W3^T d_b_eff (= db2),  db1 = W2^T dt,  t = W2 b1 + b2
// i.e. ~50 k multiply-adds instead of the 3 x 200 k of the textbook chain (M = W2 W1, dM = W3^T dW_eff, dW1 = W2^T dM ...),
// which on ONE CTA was the long pole of both launches.
// shared memory: W1 [64][69] | W2 [64][65] | W3 [4][65] | dW1, dW2, dW3 likewise | raw V1, V2, V3 likewise | dWe [3][69] | P, Q [3][64] | vectors
constexpr int LS_PP_P1 = LS_PP_MAX_IN + 1, LS_PP_P2 = LS_H + 1;
constexpr int LS_PP_SMEM_FLOATS = 3 * (LS_H * LS_PP_P1 + LS_H * LS_PP_P2 + 4 * LS_PP_P2) + 3 * LS_PP_P1 + 6 * LS_H + 4 * LS_H + 2 * (2 * LS_H + 4) + 16;

constexpr int LS_PP_ROUNDS = 9;      // (64 + 64 + 3 decoder rows) / 16 warps
constexpr int LS_PP_ROWS = 2 * LS_H + 4;

// row -> (matrix m, row j inside it) of the decoder's three matrices stacked
LS_DEV void ls_pp_row(const LsParamArgs& a, int row, int* m, int* j) {
    const int r1 = a.rad[0].dout, r2 = r1 + a.rad[1].dout;
    *m = row < r1 ? 0 : (row < r2 ? 1 : 2);
    *j = row - (*m == 0 ? 0 : (*m == 1 ? r1 : r2));
}

// the three effective matrices of the decoder, one warp per row, the rows of all three in one sweep whose global loads are ALL
// issued before the first one is used (after the step's table traffic every parameter read is an L2 miss: nine dependent
// round trips otherwise).  V1..V3 (nullable): the raw rows, n2s / gs: ||v||^2 and g per stacked row -- for the backward.
LS_DEV void ls_pp_effective3(const LsParamArgs& a, float* W1, float* W2, float* W3, float* V1, float* V2, float* V3, float* n2s, float* gs,
                             int tid, int nt) {
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int r3 = a.rad[0].dout + a.rad[1].dout + a.rad[2].dout;
    for (int base = 0; base < r3; base += LS_PP_ROUNDS * nw) {
        float x[LS_PP_ROUNDS][3], gj[LS_PP_ROUNDS];     // din <= 68 < 96: a row is three registers per lane
#pragma unroll
        for (int u = 0; u < LS_PP_ROUNDS; ++u) {
            const int row = base + warp + u * nw;
            x[u][0] = x[u][1] = x[u][2] = 0.f; gj[u] = 0.f;
            if (row < r3) {
                int m, j;
                ls_pp_row(a, row, &m, &j);
                const LsParamLayer& L = a.rad[m];
#pragma unroll
                for (int w = 0; w < 3; ++w) { const int i = lane + 32 * w; if (i < L.din) x[u][w] = L.v[j * L.din + i]; }
                gj[u] = L.g[j];
            }
        }
#pragma unroll
        for (int u = 0; u < LS_PP_ROUNDS; ++u) {
            const int row = base + warp + u * nw;
            if (row < r3) {
                int m, j;
                ls_pp_row(a, row, &m, &j);
                const int din = a.rad[m].din, pitch = m == 0 ? LS_PP_P1 : LS_PP_P2;
                float* W = m == 0 ? W1 : (m == 1 ? W2 : W3);
                float* V = m == 0 ? V1 : (m == 1 ? V2 : V3);
                float n2 = 0.f;
#pragma unroll
                for (int w = 0; w < 3; ++w) n2 = fmaf(x[u][w], x[u][w], n2);      // (lanes past din hold zeros)
                n2 = ls_warp_sum(n2);
                const float sc = gj[u] / sqrtf(n2);
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const int i = lane + 32 * w;
                    if (i < din) { W[j * pitch + i] = x[u][w] * sc; if (V) V[j * pitch + i] = x[u][w]; }
                }
                if (n2s && lane == 0) { n2s[row] = n2; gs[row] = gj[u]; }
            }
        }
    }
}

// weight-norm backward of the three decoder matrices: dW (shared memory, same layout as W) -> dg, dv; one warp per stacked row,
// everything but the outputs comes from shared memory
LS_DEV void ls_pp_weightnorm_bwd3(const LsParamArgs& a, const float* dW1, const float* dW2, const float* dW3, const float* V1, const float* V2,
                                  const float* V3, const float* n2s, const float* gs, int tid, int nt) {
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int r3 = a.rad[0].dout + a.rad[1].dout + a.rad[2].dout;
    for (int row = warp; row < r3; row += nw) {
        int m, j;
        ls_pp_row(a, row, &m, &j);
        const LsParamLayer& L = a.rad[m];
        const int pitch = m == 0 ? LS_PP_P1 : LS_PP_P2;
        const float* dW = (m == 0 ? dW1 : (m == 1 ? dW2 : dW3)) + j * pitch;
        const float* V = (m == 0 ? V1 : (m == 1 ? V2 : V3)) + j * pitch;
        float dot = 0.f;
        for (int i = lane; i < L.din; i += 32) dot = fmaf(dW[i], V[i], dot);
        dot = ls_warp_sum(dot);
        const float inv = 1.f / sqrtf(n2s[row]);
        if (lane == 0) ls_pp_out(L.dg + j, dot * inv, a.accumulate);       // dW . v / ||v||
        const float c = gs[row] * inv, k = dot * inv * inv;               // dv = g/||v|| (dW - (dW . v) v / ||v||^2)
        for (int i = lane; i < L.din; i += 32) ls_pp_out(L.dv + j * L.din + i, c * (dW[i] - k * V[i]), a.accumulate);
    }
}

__global__ void __launch_bounds__(LS_PP_THREADS, 1) ls_params_forward_kernel(const LsParamArgs a) {
    LS_DYN_SMEM(smem);
    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- geometry MLP: theta = (W_l^T, b_l); CTA b >= 1 owns layer b - 1, one warp per output row
    if (blockIdx.x > 0) {
        const int l = (int)blockIdx.x - 1;
        if (l >= a.n_geo) return;
        int off = 0;
        for (int k = 0; k < l; ++k) off += a.geo[k].din * a.geo[k].dout + a.geo[k].dout;
        const LsParamLayer& L = a.geo[l];
        const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
        if (L.din <= 96) {      // rows in registers, the loads of four rounds in flight together (each one is an L2 miss)
            for (int base = 0; base < L.dout; base += 4 * nw) {
                float x[4][3], gj[4], bj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = base + warp + u * nw;
                    x[u][0] = x[u][1] = x[u][2] = 0.f; gj[u] = bj[u] = 0.f;
                    if (j < L.dout) {
#pragma unroll
                        for (int w = 0; w < 3; ++w) { const int i = lane + 32 * w; if (i < L.din) x[u][w] = L.v[j * L.din + i]; }
                        gj[u] = L.g[j]; bj[u] = L.b[j];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = base + warp + u * nw;
                    if (j < L.dout) {
                        float n2 = 0.f;
#pragma unroll
                        for (int w = 0; w < 3; ++w) n2 = fmaf(x[u][w], x[u][w], n2);
                        n2 = ls_warp_sum(n2);
                        const float sc = gj[u] / sqrtf(n2);
#pragma unroll
                        for (int w = 0; w < 3; ++w) { const int i = lane + 32 * w; if (i < L.din) a.theta[off + i * L.dout + j] = x[u][w] * sc; }
                        if (lane == 0) a.theta[off + L.din * L.dout + j] = bj[u];
                    }
                }
            }
            return;
        }
        for (int j = warp; j < L.dout; j += nw) {
            float n2 = 0.f;
            for (int i = lane; i < L.din; i += 32) { const float x = L.v[j * L.din + i]; n2 = fmaf(x, x, n2); }
            n2 = ls_warp_sum(n2);
            const float sc = L.g[j] / sqrtf(n2);
            for (int i = lane; i < L.din; i += 32) a.theta[off + i * L.dout + j] = L.v[j * L.din + i] * sc;
            if (lane == 0) a.theta[off + L.din * L.dout + j] = L.b[j];
        }
        return;
    }
    if (!a.has_rad) return;
    // ---- radiance decoder: W_eff = (W3 W2) W1, b_eff = (W3 W2) b1 + W3 b2 + b3
    float* W1 = smem; float* W2 = W1 + LS_H * LS_PP_P1; float* W3 = W2 + LS_H * LS_PP_P2; float* P = W3 + 4 * LS_PP_P2;
    float* b12 = P + 3 * LS_H;                               // b1 | b2
    const int din = a.rad[0].din;
    if (tid >= nt - 2 * LS_H) { const int e = tid - (nt - 2 * LS_H); b12[e] = e < LS_H ? a.rad[0].b[e] : a.rad[1].b[e - LS_H]; }
    ls_pp_effective3(a, W1, W2, W3, nullptr, nullptr, nullptr, nullptr, nullptr, tid, nt);
    __syncthreads();
    for (int e = tid; e < 3 * LS_H; e += nt) {               // P = W3 W2
        const int c = e / LS_H, k = e - c * LS_H;
        float acc0 = 0.f, acc1 = 0.f;
        for (int j = 0; j < LS_H; j += 2) {
            acc0 = fmaf(W3[c * LS_PP_P2 + j], W2[j * LS_PP_P2 + k], acc0);
            acc1 = fmaf(W3[c * LS_PP_P2 + j + 1], W2[(j + 1) * LS_PP_P2 + k], acc1);
        }
        P[e] = acc0 + acc1;
    }
    __syncthreads();
    for (int e = tid; e < 3 * din; e += nt) {                // W_eff = P W1
        const int c = e / din, i = e - c * din;
        float acc0 = 0.f, acc1 = 0.f;
        for (int k = 0; k < LS_H; k += 2) {
            acc0 = fmaf(P[c * LS_H + k], W1[k * LS_PP_P1 + i], acc0);
            acc1 = fmaf(P[c * LS_H + k + 1], W1[(k + 1) * LS_PP_P1 + i], acc1);
        }
        a.w_eff[c * din + i] = acc0 + acc1;
    }
    if (tid >= nt - 3) {                                     // (the last warp: idle in the loop above)
        const int c = tid - (nt - 3);
        float acc = a.rad[2].b[c];
        for (int k = 0; k < LS_H; ++k) acc = fmaf(P[c * LS_H + k], b12[k], acc);
        for (int j = 0; j < LS_H; ++j) acc = fmaf(W3[c * LS_PP_P2 + j], b12[LS_H + j], acc);
        a.b_eff[c] = acc;
    }
}

__global__ void __launch_bounds__(LS_PP_THREADS, 1) ls_params_backward_kernel(const LsParamArgs a) {
    LS_DYN_SMEM(smem);
    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- geometry MLP: dW_l[j][i] = d_theta[off + i*dout + j]; CTA b >= 1 owns layer b - 1, one warp per output row
    if (blockIdx.x > 0) {
        const int l = (int)blockIdx.x - 1;
        if (l >= a.n_geo || !a.d_theta) return;
        int off = 0;
        for (int k = 0; k < l; ++k) off += a.geo[k].din * a.geo[k].dout + a.geo[k].dout;
        const LsParamLayer& L = a.geo[l];
        const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
        if (L.din <= 96) {      // as in the forward: four rounds of loads in flight together
            for (int base = 0; base < L.dout; base += 4 * nw) {
                float x[4][3], dth[4][3], gj[4], dbj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = base + warp + u * nw;
                    gj[u] = dbj[u] = 0.f;
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        const int i = lane + 32 * w;
                        const bool on = j < L.dout && i < L.din;
                        x[u][w] = on ? L.v[j * L.din + i] : 0.f;
                        dth[u][w] = on ? a.d_theta[off + i * L.dout + j] : 0.f;
                    }
                    if (j < L.dout) { gj[u] = L.g[j]; dbj[u] = a.d_theta[off + L.din * L.dout + j]; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = base + warp + u * nw;
                    if (j < L.dout) {
                        float n2 = 0.f, dot = 0.f;
#pragma unroll
                        for (int w = 0; w < 3; ++w) { n2 = fmaf(x[u][w], x[u][w], n2); dot = fmaf(dth[u][w], x[u][w], dot); }
                        n2 = ls_warp_sum(n2);
                        dot = ls_warp_sum(dot);
                        const float inv = 1.f / sqrtf(n2);
                        const float c = gj[u] * inv, k = dot * inv * inv;
#pragma unroll
                        for (int w = 0; w < 3; ++w) {
                            const int i = lane + 32 * w;
                            if (i < L.din) ls_pp_out(L.dv + j * L.din + i, c * (dth[u][w] - k * x[u][w]), a.accumulate);
                        }
                        if (lane == 0) { ls_pp_out(L.dg + j, dot * inv, a.accumulate); ls_pp_out(L.db + j, dbj[u], a.accumulate); }
                    }
                }
            }
            return;
        }
        for (int j = warp; j < L.dout; j += nw) {
            float n2 = 0.f, dot = 0.f;
            for (int i = lane; i < L.din; i += 32) {
                const float x = L.v[j * L.din + i];
                n2 = fmaf(x, x, n2);
                dot = fmaf(a.d_theta[off + i * L.dout + j], x, dot);
            }
            n2 = ls_warp_sum(n2);
            dot = ls_warp_sum(dot);
            const float inv = 1.f / sqrtf(n2);
            const float c = L.g[j] * inv, k = dot * inv * inv;
            for (int i = lane; i < L.din; i += 32)
                ls_pp_out(L.dv + j * L.din + i, c * (a.d_theta[off + i * L.dout + j] - k * L.v[j * L.din + i]), a.accumulate);
            if (lane == 0) { ls_pp_out(L.dg + j, dot * inv, a.accumulate); ls_pp_out(L.db + j, a.d_theta[off + L.din * L.dout + j], a.accumulate); }
        }
        return;
    }
    if (!a.has_rad || !a.d_w_eff) return;
    float* W1 = smem; float* W2 = W1 + LS_H * LS_PP_P1; float* W3 = W2 + LS_H * LS_PP_P2;
    float* dW1 = W3 + 4 * LS_PP_P2;               // [64][P1]
    float* dW2 = dW1 + LS_H * LS_PP_P1;           // [64][P2]
    float* dW3 = dW2 + LS_H * LS_PP_P2;           // [4][P2]
    float* dWe = dW3 + 4 * LS_PP_P2;              // [3][P1]  d_w_eff
    float* P = dWe + 3 * LS_PP_P1;                // [3][64]  W3 W2
    float* Q = P + 3 * LS_H;                      // [3][64]  dW_eff W1^T
    float* tv = Q + 3 * LS_H;                     // [64]     t = W2 b1 + b2
    float* dt = tv + LS_H;                        // [64]     W3^T d_b_eff
    float* db1 = dt + LS_H;                       // [64]
    float* b1 = db1 + LS_H;                       // [64]
    float* dbe = b1 + LS_H;                       // d_b_eff [3] (+pad)
    float* V1 = dbe + 4; float* V2 = V1 + LS_H * LS_PP_P1; float* V3 = V2 + LS_H * LS_PP_P2;     // raw rows
    float* n2s = V3 + 4 * LS_PP_P2; float* gs = n2s + LS_PP_ROWS;                               // ||v||^2, g per stacked row
    const int din = a.rad[0].din;
    for (int e = tid; e < 3 * din; e += nt) { const int c = e / din, i = e - c * din; dWe[c * LS_PP_P1 + i] = a.d_w_eff[e]; }
    if (tid >= nt - 64) b1[tid - (nt - 64)] = a.rad[0].b[tid - (nt - 64)];
    if (tid < 4) dbe[tid] = (tid < 3 && a.d_b_eff) ? a.d_b_eff[tid] : 0.f;
    ls_pp_effective3(a, W1, W2, W3, V1, V2, V3, n2s, gs, tid, nt);
    __syncthreads();
    // ---- P, Q, t, dt: four independent jobs on four thread ranges
    if (tid < 3 * LS_H) {                                    // P = W3 W2
        const int c = tid / LS_H, k = tid - c * LS_H;
        float acc0 = 0.f, acc1 = 0.f;
        for (int j = 0; j < LS_H; j += 2) {
            acc0 = fmaf(W3[c * LS_PP_P2 + j], W2[j * LS_PP_P2 + k], acc0);
            acc1 = fmaf(W3[c * LS_PP_P2 + j + 1], W2[(j + 1) * LS_PP_P2 + k], acc1);
        }
        P[tid] = acc0 + acc1;
    } else if (tid < 6 * LS_H) {                             // Q = dW_eff W1^T
        const int e = tid - 3 * LS_H, c = e / LS_H, j = e - c * LS_H;
        float acc0 = 0.f, acc1 = 0.f;
        int i = 0;
        for (; i + 1 < din; i += 2) {
            acc0 = fmaf(dWe[c * LS_PP_P1 + i], W1[j * LS_PP_P1 + i], acc0);
            acc1 = fmaf(dWe[c * LS_PP_P1 + i + 1], W1[j * LS_PP_P1 + i + 1], acc1);
        }
        if (i < din) acc0 = fmaf(dWe[c * LS_PP_P1 + i], W1[j * LS_PP_P1 + i], acc0);
        Q[e] = acc0 + acc1;
    } else if (tid < 7 * LS_H) {                             // t = W2 b1 + b2
        const int j = tid - 6 * LS_H;
        float acc = a.rad[1].b[j];
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W2[j * LS_PP_P2 + k], b1[k], acc);
        tv[j] = acc;
    } else if (tid < 8 * LS_H) {                             // dt = W3^T d_b_eff
        const int j = tid - 7 * LS_H;
        dt[j] = W3[j] * dbe[0] + W3[LS_PP_P2 + j] * dbe[1] + W3[2 * LS_PP_P2 + j] * dbe[2];
    }
    __syncthreads();
    // ---- dW3 = Q W2^T + d_b_eff (x) t   |   db1 = W2^T dt   (256 threads), then the two rank-3 outer products (everyone)
    if (tid < 3 * LS_H) {
        const int c = tid / LS_H, k = tid - c * LS_H;
        float acc0 = dbe[c] * tv[k], acc1 = 0.f;
        for (int j = 0; j < LS_H; j += 2) {
            acc0 = fmaf(Q[c * LS_H + j], W2[k * LS_PP_P2 + j], acc0);
            acc1 = fmaf(Q[c * LS_H + j + 1], W2[k * LS_PP_P2 + j + 1], acc1);
        }
        dW3[c * LS_PP_P2 + k] = acc0 + acc1;
    } else if (tid < 4 * LS_H) {
        const int j = tid - 3 * LS_H;
        float acc = 0.f;
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W2[k * LS_PP_P2 + j], dt[k], acc);
        db1[j] = acc;
    }
    for (int e = tid; e < LS_H * LS_H; e += nt) {             // dW2 = W3^T Q + dt (x) b1
        const int j = e / LS_H, k = e - j * LS_H;
        float acc = dt[j] * b1[k];
        acc = fmaf(W3[j], Q[k], acc);
        acc = fmaf(W3[LS_PP_P2 + j], Q[LS_H + k], acc);
        acc = fmaf(W3[2 * LS_PP_P2 + j], Q[2 * LS_H + k], acc);
        dW2[j * LS_PP_P2 + k] = acc;
    }
    for (int e = tid; e < LS_H * din; e += nt) {              // dW1 = P^T dW_eff
        const int k = e / din, i = e - k * din;
        float acc = P[k] * dWe[i];
        acc = fmaf(P[LS_H + k], dWe[LS_PP_P1 + i], acc);
        acc = fmaf(P[2 * LS_H + k], dWe[2 * LS_PP_P1 + i], acc);
        dW1[k * LS_PP_P1 + i] = acc;
    }
    __syncthreads();
    ls_pp_weightnorm_bwd3(a, dW1, dW2, dW3, V1, V2, V3, n2s, gs, tid, nt);
    for (int j = tid; j < LS_H; j += nt) { ls_pp_out(a.rad[0].db + j, db1[j], a.accumulate); ls_pp_out(a.rad[1].db + j, dt[j], a.accumulate); }
    if (tid < 3) ls_pp_out(a.rad[2].db + tid, dbe[tid], a.accumulate);
}
