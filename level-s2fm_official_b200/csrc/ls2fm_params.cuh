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
LS_DEV void ls_pp_out(float* p, float v, int acc) { *p = acc ? *p + v : v; }

// effective (weight-normed) matrix of one layer into shared memory, row-major [dout][pitch]; one warp per row
LS_DEV void ls_pp_effective(const LsParamLayer& L, float* W, int pitch, int tid, int nt) {
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int j = warp; j < L.dout; j += nw) {
        float n2 = 0.f;
        for (int i = lane; i < L.din; i += 32) { const float x = L.v[j * L.din + i]; n2 = fmaf(x, x, n2); }
        n2 = ls_warp_sum(n2);
        const float sc = L.g[j] / sqrtf(n2);
        for (int i = lane; i < L.din; i += 32) W[j * pitch + i] = L.v[j * L.din + i] * sc;
    }
}

// weight-norm backward of one layer: dW (row-major [dout][pitch], shared memory) -> dg, dv; one warp per row
LS_DEV void ls_pp_weightnorm_bwd(const LsParamLayer& L, const float* dW, int pitch, int tid, int nt, int acc) {
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int j = warp; j < L.dout; j += nw) {
        float n2 = 0.f, dot = 0.f;
        for (int i = lane; i < L.din; i += 32) {
            const float x = L.v[j * L.din + i];
            n2 = fmaf(x, x, n2);
            dot = fmaf(dW[j * pitch + i], x, dot);
        }
        n2 = ls_warp_sum(n2);
        dot = ls_warp_sum(dot);
        const float inv = 1.f / sqrtf(n2);
        if (lane == 0) ls_pp_out(L.dg + j, dot * inv, acc);    // dW . v / ||v||
        const float c = L.g[j] * inv, k = dot * inv * inv;     // dv = g/||v|| (dW - (dW . v) v / ||v||^2)
        for (int i = lane; i < L.din; i += 32) ls_pp_out(L.dv + j * L.din + i, c * (dW[j * pitch + i] - k * L.v[j * L.din + i]), acc);
    }
}

// shared memory: W1 [64][69] | W2 [64][65] | W3 [4][65] | M = W2 W1 [64][69] | t [64] | scratch
constexpr int LS_PP_P1 = LS_PP_MAX_IN + 1, LS_PP_P2 = LS_H + 1;
constexpr int LS_PP_SMEM_FLOATS = LS_H * LS_PP_P1 * 3 + LS_H * LS_PP_P2 * 2 + 4 * LS_PP_P2 * 2 + 3 * LS_H + 16;   // backward layout (the larger one)

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
    // ---- radiance decoder: W_eff = W3 W2 W1, b_eff = W3 (W2 b1 + b2) + b3
    float* W1 = smem; float* W2 = W1 + LS_H * LS_PP_P1; float* W3 = W2 + LS_H * LS_PP_P2; float* M = W3 + 4 * LS_PP_P2;
    float* t = M + LS_H * LS_PP_P1;
    const int din = a.rad[0].din;
    ls_pp_effective(a.rad[0], W1, LS_PP_P1, tid, nt);
    ls_pp_effective(a.rad[1], W2, LS_PP_P2, tid, nt);
    ls_pp_effective(a.rad[2], W3, LS_PP_P2, tid, nt);
    __syncthreads();
    for (int e = tid; e < LS_H * din; e += nt) {            // M = W2 W1
        const int j = e / din, i = e - j * din;
        float acc = 0.f;
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W2[j * LS_PP_P2 + k], W1[k * LS_PP_P1 + i], acc);
        M[j * LS_PP_P1 + i] = acc;
    }
    for (int j = tid; j < LS_H; j += nt) {                   // t = W2 b1 + b2
        float acc = a.rad[1].b[j];
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W2[j * LS_PP_P2 + k], a.rad[0].b[k], acc);
        t[j] = acc;
    }
    __syncthreads();
    for (int e = tid; e < 3 * din; e += nt) {                // W_eff = W3 M
        const int c = e / din, i = e - c * din;
        float acc = 0.f;
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W3[c * LS_PP_P2 + k], M[k * LS_PP_P1 + i], acc);
        a.w_eff[c * din + i] = acc;
    }
    if (tid < 3) {
        float acc = a.rad[2].b[tid];
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W3[tid * LS_PP_P2 + k], t[k], acc);
        a.b_eff[tid] = acc;
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
    float* W1 = smem; float* W2 = W1 + LS_H * LS_PP_P1; float* W3 = W2 + LS_H * LS_PP_P2; float* M = W3 + 4 * LS_PP_P2;
    float* t = M + LS_H * LS_PP_P1;
    float* dM = t + LS_H;                         // [64][P1]  (reused as dW1 afterwards)
    float* dW2 = dM + LS_H * LS_PP_P1;            // [64][P2]
    float* dW3 = dW2 + LS_H * LS_PP_P2;           // [4][P2]
    float* dt = dW3 + 4 * LS_PP_P2;               // [64]
    float* db1 = dt + LS_H;                       // [64]
    float* dwe = db1 + LS_H;                      // d_b_eff [3] (+pad)
    const int din = a.rad[0].din;
    ls_pp_effective(a.rad[0], W1, LS_PP_P1, tid, nt);
    ls_pp_effective(a.rad[1], W2, LS_PP_P2, tid, nt);
    ls_pp_effective(a.rad[2], W3, LS_PP_P2, tid, nt);
    if (tid < 4) dwe[tid] = (tid < 3 && a.d_b_eff) ? a.d_b_eff[tid] : 0.f;
    __syncthreads();
    for (int e = tid; e < LS_H * din; e += nt) {            // M = W2 W1
        const int j = e / din, i = e - j * din;
        float acc = 0.f;
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W2[j * LS_PP_P2 + k], W1[k * LS_PP_P1 + i], acc);
        M[j * LS_PP_P1 + i] = acc;
    }
    for (int j = tid; j < LS_H; j += nt) {
        float acc = a.rad[1].b[j];
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W2[j * LS_PP_P2 + k], a.rad[0].b[k], acc);
        t[j] = acc;
        dt[j] = W3[j] * dwe[0] + W3[LS_PP_P2 + j] * dwe[1] + W3[2 * LS_PP_P2 + j] * dwe[2];      // dt = W3^T d_b_eff
    }
    __syncthreads();
    for (int e = tid; e < 3 * LS_H; e += nt) {               // dW3 = dWeff M^T + d_b_eff (x) t
        const int c = e / LS_H, k = e - c * LS_H;
        float acc = dwe[c] * t[k];
        for (int i = 0; i < din; ++i) acc = fmaf(a.d_w_eff[c * din + i], M[k * LS_PP_P1 + i], acc);
        dW3[c * LS_PP_P2 + k] = acc;
    }
    for (int e = tid; e < LS_H * din; e += nt) {             // dM = W3^T dWeff
        const int k = e / din, i = e - k * din;
        dM[k * LS_PP_P1 + i] = W3[k] * a.d_w_eff[i] + W3[LS_PP_P2 + k] * a.d_w_eff[din + i] + W3[2 * LS_PP_P2 + k] * a.d_w_eff[2 * din + i];
    }
    for (int j = tid; j < LS_H; j += nt) {                    // db1 = W2^T dt
        float acc = 0.f;
        for (int k = 0; k < LS_H; ++k) acc = fmaf(W2[k * LS_PP_P2 + j], dt[k], acc);
        db1[j] = acc;
    }
    __syncthreads();
    for (int e = tid; e < LS_H * LS_H; e += nt) {             // dW2 = dM W1^T + dt (x) b1
        const int j = e / LS_H, k = e - j * LS_H;
        float acc = dt[j] * a.rad[0].b[k];
        for (int i = 0; i < din; ++i) acc = fmaf(dM[j * LS_PP_P1 + i], W1[k * LS_PP_P1 + i], acc);
        dW2[j * LS_PP_P2 + k] = acc;
    }
    __syncthreads();
    // dW1 = W2^T dM  (into M's storage: M itself is dead now)
    for (int e = tid; e < LS_H * din; e += nt) {
        const int k = e / din, i = e - k * din;
        float acc = 0.f;
        for (int j = 0; j < LS_H; ++j) acc = fmaf(W2[j * LS_PP_P2 + k], dM[j * LS_PP_P1 + i], acc);
        M[k * LS_PP_P1 + i] = acc;
    }
    __syncthreads();
    ls_pp_weightnorm_bwd(a.rad[0], M, LS_PP_P1, tid, nt, a.accumulate);
    ls_pp_weightnorm_bwd(a.rad[1], dW2, LS_PP_P2, tid, nt, a.accumulate);
    ls_pp_weightnorm_bwd(a.rad[2], dW3, LS_PP_P2, tid, nt, a.accumulate);
    for (int j = tid; j < LS_H; j += nt) { ls_pp_out(a.rad[0].db + j, db1[j], a.accumulate); ls_pp_out(a.rad[1].db + j, dt[j], a.accumulate); }
    if (tid < 3) ls_pp_out(a.rad[2].db + tid, dwe[tid], a.accumulate);
}
