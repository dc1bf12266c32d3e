// ls2fm_pose.cuh -- camera-pose parametrisation in front of the ray generator.
//
// utils/camera.py:85-96 (Lie.se3_to_SE3) with its Taylor-series coefficients (camera.py:119-142, nth = 10):
//   wu = (w, u) in R^6,  theta = |w|,  wx = skew(w)
//   A = sum_i (-1)^i theta^(2i) / (2i+1)!        (sin x / x)
//   B = sum_i (-1)^i theta^(2i) / (2i+2)!        ((1 - cos x) / x^2)
//   C = sum_i (-1)^i theta^(2i) / (2i+3)!        ((x - sin x) / x^3)
//   R = I + A wx + B wx wx,   V = I + B wx + C wx wx,   Rt = [R | V u]         -> [3,4]
// The eager version is ~150 tiny launches (three 11-term loops of pow / div / add on a [..,1,1] tensor plus the matrix
// algebra) and as many again in its backward; in the BA "sfm" loop it runs on one row per tracked POINT
// (pipelines/BA.py:127: se3_to_SE3(self.se3_refine[self.pose_idx])).  Here: one thread per pose, forward and backward each one
// launch.  The backward carries the six partial derivatives of every intermediate along (forward-mode duals), so there is no
// hand-derived Jacobian to get wrong; the series are evaluated in s = theta^2, which is also what makes w = 0 regular.
#pragma once

#include "ls2fm_common.cuh"

template <int NP>
struct LsDual {
    float v;
    float d[NP];
};
template <int NP> LS_DEV LsDual<NP> ls_dconst(float c) { LsDual<NP> r; r.v = c; for (int i = 0; i < NP; ++i) r.d[i] = 0.f; return r; }
template <int NP> LS_DEV LsDual<NP> ls_dvar(float c, int k) { LsDual<NP> r = ls_dconst<NP>(c); if (k >= 0) r.d[k] = 1.f; return r; }
template <int NP> LS_DEV LsDual<NP> operator+(const LsDual<NP>& a, const LsDual<NP>& b) { LsDual<NP> r; r.v = a.v + b.v; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int NP> LS_DEV LsDual<NP> operator-(const LsDual<NP>& a, const LsDual<NP>& b) { LsDual<NP> r; r.v = a.v - b.v; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int NP> LS_DEV LsDual<NP> operator*(const LsDual<NP>& a, const LsDual<NP>& b) { LsDual<NP> r; r.v = a.v * b.v; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int NP> LS_DEV LsDual<NP> operator*(const LsDual<NP>& a, float c) { LsDual<NP> r; r.v = a.v * c; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] * c; return r; }
template <int NP> LS_DEV LsDual<NP> operator-(const LsDual<NP>& a) { LsDual<NP> r; r.v = -a.v; for (int i = 0; i < NP; ++i) r.d[i] = -a.d[i]; return r; }

// Rt[12] (row-major [3][4]) from wu[6]; T = float (forward) or LsDual<6> (backward)
template <class T, class MK>
LS_DEV void ls_se3_to_SE3_eval(const T (&wu)[6], T (&Rt)[12], MK mk) {
    const T s = wu[0] * wu[0] + wu[1] * wu[1] + wu[2] * wu[2];        // theta^2
    T A = mk(0.f), B = mk(0.f), C = mk(0.f), p = mk(1.f);             // p = s^i
    float dA = 1.f, dB = 1.f, dC = 1.f, sign = 1.f;
#pragma unroll
    for (int i = 0; i <= 10; ++i) {
        if (i > 0) dA *= (float)((2 * i) * (2 * i + 1));
        dB *= (float)((2 * i + 1) * (2 * i + 2));
        dC *= (float)((2 * i + 2) * (2 * i + 3));
        A = A + p * (sign / dA);
        B = B + p * (sign / dB);
        C = C + p * (sign / dC);
        p = p * s;
        sign = -sign;
    }
    // wx = [[0,-w2,w1],[w2,0,-w0],[-w1,w0,0]];  wx wx = w w^T - s I
    const T w0 = wu[0], w1 = wu[1], w2 = wu[2];
    const T zero = mk(0.f), one = mk(1.f);
    const T wx[9] = {zero, -w2, w1, w2, zero, -w0, -w1, w0, zero};
    const T ww[9] = {w0 * w0 - s, w0 * w1, w0 * w2, w1 * w0, w1 * w1 - s, w1 * w2, w2 * w0, w2 * w1, w2 * w2 - s};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T t = zero;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const T id = r == c ? one : zero;
            Rt[4 * r + c] = id + A * wx[3 * r + c] + B * ww[3 * r + c];
            const T V = id + B * wx[3 * r + c] + C * ww[3 * r + c];
            t = t + V * wu[3 + c];
        }
        Rt[4 * r + 3] = t;
    }
}

__global__ void ls_se3_to_SE3_kernel(const float* __restrict__ wu, int64_t n, float* __restrict__ Rt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float in[6], out[12];
#pragma unroll
    for (int k = 0; k < 6; ++k) in[k] = wu[6 * i + k];
    ls_se3_to_SE3_eval(in, out, [](float c) { return c; });
#pragma unroll
    for (int k = 0; k < 12; ++k) Rt[12 * i + k] = out[k];
}

// d_wu [n,6] (written) = J^T g_Rt, J from forward-mode duals
__global__ void ls_se3_to_SE3_backward_kernel(const float* __restrict__ wu, int64_t n, const float* __restrict__ g_Rt, float* __restrict__ d_wu) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    LsDual<6> in[6], out[12];
#pragma unroll
    for (int k = 0; k < 6; ++k) in[k] = ls_dvar<6>(wu[6 * i + k], k);
    ls_se3_to_SE3_eval(in, out, [](float c) { return ls_dconst<6>(c); });
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const float g = g_Rt[12 * i + k];
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[j] = fmaf(g, out[k].d[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) d_wu[6 * i + j] = acc[j];
}

// ---------------------------------------------------------------- reprojection residual of the BA "sfm" loop
// pipelines/BA.py:126-141 for n tracked points, each with its own world->camera pose Rt [3,4] (BA.py:127 gathers one pose per
// point) and the shared intrinsics K:
//   (a, b, c) = K (R x + t);  uv = (a, b) / (c + eps);  d = |uv - kp|
//   mask_surf = |sdf| < sdf_band (= 2 * sdf_threshold);  rows with an infinite uv component are dropped
//   loss = 0.5 * mean(2 log(1 + d^2 / 4)) + 0.5 * mean(d)   over the kept rows
// The eager version is ~40 launches (to_hom / matmuls / divisions / two boolean-mask gathers with their host syncs) and as many in
// its backward.  Pass 1: per-point residuals and the sums (sums: [0] sum 2 log(1 + d^2/4), [1] sum d, [2] kept rows, [3] mask_surf
// rows).  Pass 2: gradients of the loss w.r.t. the points and the poses, with the data-dependent count read from sums.
struct LsReproj { float uv[2]; float d; bool surf, kept; };
LS_DEV LsReproj ls_reproj_point(const float* x, const float* P, const float* K, const float* kp, float sdf, float band, float eps,
                                float* abc, float* pc) {
    LsReproj r;
#pragma unroll
    for (int i = 0; i < 3; ++i) pc[i] = P[4 * i] * x[0] + P[4 * i + 1] * x[1] + P[4 * i + 2] * x[2] + P[4 * i + 3];
#pragma unroll
    for (int i = 0; i < 3; ++i) abc[i] = K[3 * i] * pc[0] + K[3 * i + 1] * pc[1] + K[3 * i + 2] * pc[2];
    const float z = abc[2] + eps;
    r.uv[0] = abc[0] / z; r.uv[1] = abc[1] / z;
    const float dx = r.uv[0] - kp[0], dy = r.uv[1] - kp[1];
    r.d = sqrtf(dx * dx + dy * dy);
    r.surf = fabsf(sdf) < band;
    r.kept = r.surf && !isinf(r.uv[0]) && !isinf(r.uv[1]);
    return r;
}
__global__ void ls_reproj_sums_kernel(const float* __restrict__ xyz, const float* __restrict__ Rt, const float* __restrict__ K,
                                      const float* __restrict__ kp, const float* __restrict__ sdf, int64_t n, float band, float eps,
                                      float* __restrict__ sums, float* __restrict__ uv_out, unsigned char* __restrict__ mask_surf) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    float s1 = 0.f, s2 = 0.f, nk = 0.f, ns = 0.f;
    for (int64_t i = gid; i < n; i += stride) {
        float abc[3], pc[3];
        const LsReproj r = ls_reproj_point(xyz + 3 * i, Rt + 12 * i, K, kp + 2 * i, sdf[i], band, eps, abc, pc);
        if (uv_out) { uv_out[2 * i] = r.uv[0]; uv_out[2 * i + 1] = r.uv[1]; }
        if (mask_surf) mask_surf[i] = r.surf ? 1 : 0;
        if (r.surf) ns += 1.f;
        if (r.kept) { s1 += 2.f * logf(1.f + r.d * r.d / 4.f); s2 += r.d; nk += 1.f; }
    }
    // (warp sums by shuffle; every lane of every warp reaches this point)
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        nk += __shfl_xor_sync(0xffffffffu, nk, o); ns += __shfl_xor_sync(0xffffffffu, ns, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(sums, s1); atomicAdd(sums + 1, s2); atomicAdd(sums + 2, nk); atomicAdd(sums + 3, ns); }
}
__global__ void ls_reproj_grads_kernel(const float* __restrict__ xyz, const float* __restrict__ Rt, const float* __restrict__ K,
                                       const float* __restrict__ kp, const float* __restrict__ sdf, int64_t n, float band, float eps,
                                       const float* __restrict__ sums, float* __restrict__ g_xyz, float* __restrict__ g_Rt) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const float nk = sums[2];
    for (int64_t i = gid; i < n; i += stride) {
        const float* x = xyz + 3 * i;
        const float* P = Rt + 12 * i;
        float abc[3], pc[3];
        const LsReproj r = ls_reproj_point(x, P, K, kp + 2 * i, sdf[i], band, eps, abc, pc);
        float gp[3] = {0.f, 0.f, 0.f};
        if (r.kept && nk > 0.f && r.d > 0.f) {
            const float gd = 0.5f / nk * (r.d / (1.f + r.d * r.d / 4.f) + 1.f);
            const float gu = gd * (r.uv[0] - kp[2 * i]) / r.d, gv = gd * (r.uv[1] - kp[2 * i + 1]) / r.d;
            const float z = abc[2] + eps;
            const float gabc[3] = {gu / z, gv / z, -(gu * abc[0] + gv * abc[1]) / (z * z)};
#pragma unroll
            for (int j = 0; j < 3; ++j) gp[j] = K[j] * gabc[0] + K[3 + j] * gabc[1] + K[6 + j] * gabc[2];       // K^T g_abc
        }
        if (g_xyz) {
#pragma unroll
            for (int j = 0; j < 3; ++j) g_xyz[3 * i + j] = P[j] * gp[0] + P[4 + j] * gp[1] + P[8 + j] * gp[2];   // R^T g_p
        }
        if (g_Rt) {
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                g_Rt[12 * i + 4 * rr] = gp[rr] * x[0]; g_Rt[12 * i + 4 * rr + 1] = gp[rr] * x[1];
                g_Rt[12 * i + 4 * rr + 2] = gp[rr] * x[2]; g_Rt[12 * i + 4 * rr + 3] = gp[rr];
            }
        }
    }
}
