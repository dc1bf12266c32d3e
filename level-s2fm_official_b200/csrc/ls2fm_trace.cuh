// ls2fm_trace.cuh -- the no-grad march of SDF.sphere_tracing (models/SDF.py:116-200) as ONE kernel.
//
// Reference: a python `while True` with two SDF evaluations on boolean-masked subsets and two host synchronisations per
// iteration (`unfinished_mask_start.sum()`), all rays in lock step.  The only coupling between rays is the global stop
// test "no start front is unfinished", which merely decides how many track points K are used afterwards.  So here every
// ray marches on its own: a warp tile = 4 rays x (start, end) front = 8 points, the fused hash-grid + MLP evaluation of
// the field kernel runs inside the loop, per-iteration counts of unfinished start fronts go to a device array with one
// atomic per warp, and K = first iteration whose count is zero is read back once at the end (the per-ray state for every
// iteration < K is exactly the reference's; what happens after K is never used).
#pragma once

#include "ls2fm_field.cuh"

struct LsTraceArgs {
    LsFieldArgs fa;               // field, staged-network plan (forward layout)
    const float* ray0; const float* dir;
    int64_t m;
    float cx, cy, cz, hx, hy, hz;
    float thr;
    int iters_max;
    float* track;                 // [m][iters_max][3] start-front points x_k before step k
    int* cnt;                     // [iters_max + 1] unfinished start fronts at the top of iteration k (pre-zeroed)
    float* t_near; float* t_far;  // [m]
    float* acc_e_hist;            // [iters_max + 1][m] end-front depth at the top of iteration k
};

__global__ void __launch_bounds__(512, 1) ls_sphere_trace_kernel(const LsTraceArgs t) {
    LS_DYN_SMEM(smem);
    const LsFieldArgs& a = t.fa;
    ls_stage_weights(a, smem);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int K = a.f.n_layers;
    float* E = smem + a.net.warp_base + warp * a.net.warp_stride;
    float* A = E + LS_WS * LS_EROWS;
    float* Y = A + LS_WS * LS_H * (K - 1);
    const int s8 = lane & 7, g = lane >> 3, og = lane & 15, sg = lane >> 4;
    const int front = s8 & 1;
    const float c[3] = {t.cx, t.cy, t.cz}, h[3] = {t.hx, t.hy, t.hz};
    const int64_t n_tiles = (t.m + 3) / 4;
    for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < n_tiles; tile += (int64_t)gridDim.x * nw) {
        const int64_t ray = tile * 4 + (s8 >> 1);
        const bool valid = ray < t.m;
        float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f};
        if (valid) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { o[k] = __ldg(t.ray0 + 3 * ray + k); d[k] = __ldg(t.dir + 3 * ray + k); }
        }
        float tn, tf;
        ls_ray_aabb(o, d, c, h, &tn, &tf);
        if (valid && g == 0 && front == 0) { t.t_near[ray] = tn; t.t_far[ray] = tf; }
        float acc = front ? tf : tn;
        float p[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = ls_fadd(o[k], ls_fmul(acc, d[k]));
        float s = 0.f;
        bool un = true;
        for (int it = 0; it <= t.iters_max; ++it) {
            // evaluate the field only while something in this tile can still change
            const bool live = it == 0 || un || s != 0.f;
            if (__any_sync(0xffffffffu, live)) {
                ls_encode_tile(a, E, p, s8, g);
                ls_mlp_forward(a, smem, E, A, Y, og, sg);
                const float sdf_new = a.s * ls_el(Y, LS_OROWS, 0, s8);
                __syncwarp();
                if (it == 0 || un) s = sdf_new;           // only unfinished fronts take the new value (SDF.py:185-194)
            }
            const float acc_other = __shfl_xor_sync(0xffffffffu, acc, 1);
            if (it > 0) {
                const float acc_s = front ? acc_other : acc, acc_e = front ? acc : acc_other;
                un = un && (acc_s < acc_e);               // fronts that crossed are finished (SDF.py:199-200)
            }
            if (fabsf(s) <= t.thr) s = 0.f;               // SDF.py:153-157
            un = (it == 0 ? true : un) && (fabsf(s) > t.thr);
            const unsigned bal = __ballot_sync(0xffffffffu, valid && g == 0 && front == 0 && un);
            if (lane == 0 && bal) atomicAdd(t.cnt + it, __popc(bal));
            if (valid && g == 0 && front == 1) t.acc_e_hist[(int64_t)it * t.m + ray] = acc;
            if (it == t.iters_max) break;
            if (valid && g == 0 && front == 0) {
                float* tr = t.track + ((int64_t)ray * t.iters_max + it) * 3;
                tr[0] = p[0]; tr[1] = p[1]; tr[2] = p[2];
            }
            acc = fminf(ls_fadd(acc, s), tf);             // where(acc > t_far, t_far, acc)
#pragma unroll
            for (int k = 0; k < 3; ++k) p[k] = ls_fadd(o[k], ls_fmul(acc, d[k]));
        }
        __syncwarp();
    }
}
