// ls2fm_sampler.cuh -- the error-bounded (VolSDF) depth sampler of Renderer.volsdf_sampling
// (models/Renderer.py:186-328, with the three fixes of SURVEY.md 8(a) a5 that make the dormant branch runnable:
// opt.VolSDF -> opt.SDF.VolSDF, max_bisection_itr = 10, SDF.forward = SDF.infer_sdf).
//
// The reference runs the algorithm with boolean-mask gathers and a host synchronisation per decision
// (.sum() > 0 at Renderer.py:213,227,262,270,313).  Here one WARP owns one ray; the ray's samples live in shared
// memory; rays that still need samples are compacted on the device (atomic slot claim) so that the SDF evaluation
// between two rounds -- one launch of the fused field kernel on the compacted list -- touches active rays only.
// No host synchronisation anywhere: every round is launched unconditionally and exits at once when the device-side
// counter says nothing is left.
//
//   round it (0..max_upsample_iter), per active ray:
//     merge the N samples drawn in round it-1 (already sorted) into the ray's sorted list (rank merge),
//     error bound with the network beta (Renderer.error_bound, Renderer.py:330-360): converged -> 64 inverse-CDF
//     samples of the opacity (opacity_to_sample, 129-162) and done;
//     otherwise bisection on beta+ (281-291; round 0 uses the closed-form beta+_0 of 191-192), clamp the bounds,
//     then either (last round) sample with beta+ and give up, or draw N new samples from the bound pdf
//     (sample_pdf, 362-399, det=True, [1:-1]) and claim a slot in the next round's ray list.
#pragma once

#include "ls2fm_common.cuh"
#include "ls2fm_render.cuh"

struct LsSamplerArgs {
    const float* center; const float* ray; const float* beta_param;
    int n_rays, N, Nf, max_iter, max_bisect, Mmax;
    float eps, beta_speed;
    float cx, cy, cz, hx, hy, hz;
    // workspace
    float* D; float* S;            // [R][Mmax] depths / sdf values (sorted prefix + newly drawn tail)
    float* beta_plus;              // [R]
    float* fine;                   // [R][Nf]
    float* iters;                  // [R]
    int* state;                    // [R] 0 active, 1 converged, 2 gave up
    int* cnt;                      // [max_iter + 2] active rays per round
    int* ray_index;                // [max_iter + 2][R]
    float* hits;                   // [R][2]
};

// torch.linspace(0, 1, n)[i] on the CPU: i < n/2 ? step*i : 1 - step*(n-1-i), each with a single rounding
LS_DEV float ls_linspace01(int i, int n) {
    const float step = 1.0f / (float)(n - 1);
    return i < n / 2 ? ls_fma(step, (float)i, 0.f) : ls_fma(-step, (float)(n - 1 - i), 1.0f);
}

// Renderer.error_bound over the ray's M sorted samples; writes bound[0..M-2] (when out != nullptr), returns the max.
// clampit: torch.clamp(bounds, 0, 1e5) of Renderer.py:300.
// stop_above >= 0: the caller only compares the maximum with this threshold -- return +inf as soon as one chunk exceeds it (same decision,
// the remaining chunks are not evaluated).
LS_DEV float ls_error_bound(const float* d, const float* s, int M, float alpha, float beta, float* out, bool clampit, int lane,
                            float stop_above = -1.f) {
    float carryR = 0.f, carryE = 0.f, mx = -INFINITY;
    const float k = alpha / (4.f * beta);
    for (int base = 0; base < M - 1; base += 32) {
        const int i = base + lane;
        const bool on = i < M - 1;
        float sd = 0.f, err = 0.f;
        if (on) {
            const float delta = d[i + 1] - d[i];
            sd = ls_sdf_to_sigma(s[i], alpha, beta) * delta;
            const float dstar = fmaxf(0.5f * (fabsf(s[i]) + fabsf(s[i + 1]) - delta), 0.f);
            err = k * (delta * delta) * expf(-dstar / beta);
        }
        const float inR = ls_warp_incl_scan(sd, lane);
        const float inE = ls_warp_incl_scan(err, lane);
        float b = expf(-(carryR + inR - sd)) * (expf(carryE + inE) - 1.0f);
        if (b != b) b = INFINITY;
        if (clampit) b = fminf(fmaxf(b, 0.f), 1e5f);
        carryR += __shfl_sync(0xffffffffu, inR, 31);
        carryE += __shfl_sync(0xffffffffu, inE, 31);
        if (on) {
            mx = fmaxf(mx, b);
            if (out) out[i] = b;
        }
        if (stop_above >= 0.f && __any_sync(0xffffffffu, on && b > stop_above)) return INFINITY;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    return mx;
}

// first index in [0, n) with a[idx] >= v, n if none (torch.searchsorted(..., right=False))
LS_DEV int ls_lower_bound(const float* a, int n, float v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// opacity_to_sample + sample_depth_from_opacity: Nf inverse-CDF samples of the opacity; op is scratch [M]
LS_DEV void ls_opacity_samples(const float* d, const float* s, int M, float alpha, float beta, float* op, int Nf, float* out, int lane) {
    float carry = 0.f;
    if (lane == 0) op[0] = 0.f;
    for (int base = 0; base < M - 1; base += 32) {
        const int i = base + lane;
        const bool on = i < M - 1;
        float sd = 0.f;
        if (on) sd = ls_sdf_to_sigma(s[i], alpha, beta) * (d[i + 1] - d[i]);
        const float incl = ls_warp_incl_scan(sd, lane);
        if (on) op[i + 1] = 1.f - expf(-(carry + incl - sd));       // 1 - exp(-R_i), R exclusive
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    for (int j = lane; j < Nf; j += 32) {
        const float u = 0.5f * (ls_linspace01(j, Nf + 1) + ls_linspace01(j + 1, Nf + 1));
        const int idx = ls_lower_bound(op, M, u);
        const int lo = idx - 1 > 0 ? idx - 1 : 0, hi = idx < M - 1 ? idx : M - 1;
        const float t = (u - op[lo]) / (op[hi] - op[lo] + 1e-8f);
        out[j] = d[lo] + t * (d[hi] - d[lo]);
    }
    __syncwarp();
}

// ---------------------------------------------------------------- init: AABB, beta+_0, N uniform samples
__global__ void ls_sampler_init_kernel(const LsSamplerArgs a) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= a.n_rays) return;
    const float oo[3] = {a.center[3 * r], a.center[3 * r + 1], a.center[3 * r + 2]};
    const float dd[3] = {a.ray[3 * r], a.ray[3 * r + 1], a.ray[3 * r + 2]};
    const float c[3] = {a.cx, a.cy, a.cz}, h[3] = {a.hx, a.hy, a.hz};
    float tn, tf;
    ls_ray_aabb(oo, dd, c, h, &tn, &tf);
    const float ext = ls_fsub(tf, tn);
    for (int i = lane; i < a.N; i += 32)
        a.D[(int64_t)r * a.Mmax + i] = ls_fadd(ls_fmul(ls_fdiv((float)i + 0.5f, (float)a.N), ext), tn);
    if (lane == 0) {
        a.hits[2 * r] = tn; a.hits[2 * r + 1] = tf;
        // beta+_0 = sqrt(t_far^2 / (4 (N-1) log(1 + eps)))   (Renderer.py:191-192, float32)
        a.beta_plus[r] = sqrtf(ls_fdiv(ls_fmul(tf, tf), ls_fmul((float)(4 * (a.N - 1)), logf(ls_fadd(1.0f, a.eps)))));
        a.state[r] = 0;
        a.iters[r] = 0.f;
        a.ray_index[r] = r;
        if (r == 0) {
            a.cnt[0] = a.n_rays;
            for (int k = 1; k < a.max_iter + 2; ++k) a.cnt[k] = 0;
        }
    }
}

// ---------------------------------------------------------------- one round
// dynamic smem per warp: 5 * M floats, M = N (it + 1) (d, s, scratch b, merge buffers d2, s2)
__global__ void ls_sampler_round_kernel(const LsSamplerArgs a, int it) {
    LS_DYN_SMEM(smem);
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * wpb + warp;
    if (slot >= a.cnt[it]) return;
    const int r = a.ray_index[(int64_t)it * a.n_rays + slot];
    const int N = a.N, Mold = N * it, M = Mold + N;
    float* d = smem + warp * 5 * M;     // (arrays sized for THIS round's M, not Mmax: the early rounds -- the ones with many rays -- get 3-6x the warps per SM)
    float* s = d + M; float* b = s + M; float* d2 = b + M; float* s2 = d2 + M;
    float* Dg = a.D + (int64_t)r * a.Mmax;
    float* Sg = a.S + (int64_t)r * a.Mmax;
    // ---- load + rank-merge (old sorted prefix, new sorted tail)
    for (int i = lane; i < M; i += 32) { d2[i] = Dg[i]; s2[i] = Sg[i]; }
    __syncwarp();
    if (it == 0) {
        for (int i = lane; i < M; i += 32) { d[i] = d2[i]; s[i] = s2[i]; }
    } else {
        // stable merge that does not rely on the new tail being sorted (it is, up to an ulp of rounding in the
        // inverse-CDF interpolation; the reference simply torch.sort()s the concatenation)
        const float* nw = d2 + Mold;
        bool tail_sorted = true;                                // (it is, up to that ulp: then both ranks are binary searches)
        for (int j = lane; j < N - 1; j += 32) tail_sorted = tail_sorted && nw[j] <= nw[j + 1];
        tail_sorted = __all_sync(0xffffffffu, tail_sorted);
        for (int i = lane; i < Mold; i += 32) {                 // old element: i + #{new < old_i}
            const float v = d2[i];
            int c = 0;
            if (tail_sorted) c = ls_lower_bound(nw, N, v);
            else for (int j = 0; j < N; ++j) c += nw[j] < v ? 1 : 0;
            d[i + c] = v; s[i + c] = s2[i];
        }
        for (int j = lane; j < N; j += 32) {                    // new element: #{old <= new_j} + rank inside the tail
            const float v = nw[j];
            int lo = 0, hi = Mold;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (d2[mid] <= v) lo = mid + 1; else hi = mid; }
            int c = j;                                          // sorted tail: everything before j is <= v, nothing after j is < v
            if (!tail_sorted) { c = 0; for (int k = 0; k < N; ++k) c += (nw[k] < v || (nw[k] == v && k < j)) ? 1 : 0; }
            d[lo + c] = v; s[lo + c] = s2[Mold + j];
        }
    }
    __syncwarp();
    if (it > 0) for (int i = lane; i < M; i += 32) { Dg[i] = d[i]; Sg[i] = s[i]; }
    // ---- converged with the network's beta?
    const float beta_net = expf(__ldg(a.beta_param) * a.beta_speed);
    const float alpha_net = 1.f / beta_net;
    const float mx = ls_error_bound(d, s, M, alpha_net, beta_net, nullptr, false, lane, a.eps);
    if (!(mx > a.eps)) {
        ls_opacity_samples(d, s, M, alpha_net, beta_net, b, a.Nf, a.fine + (int64_t)r * a.Nf, lane);
        if (lane == 0) { a.state[r] = 1; a.iters[r] = (float)it; a.beta_plus[r] = beta_net; }
        return;
    }
    // ---- beta+: closed form in round 0, bisection afterwards
    float bp = a.beta_plus[r];
    if (it > 0) {
        float bl = beta_net, br = bp;
        for (int k = 0; k < a.max_bisect; ++k) {
            const float bm = 0.5f * (bl + br);
            const float m2 = ls_error_bound(d, s, M, 1.f / bm, bm, nullptr, false, lane, a.eps);
            if (m2 <= a.eps) br = bm; else bl = bm;
        }
        bp = br;
        if (lane == 0) a.beta_plus[r] = bp;
    }
    const float ap = 1.f / bp;
    if (it == a.max_iter) {     // out of rounds: sample with the last beta+ (Renderer.py:313-321)
        ls_opacity_samples(d, s, M, ap, bp, b, a.Nf, a.fine + (int64_t)r * a.Nf, lane);
        if (lane == 0) { a.state[r] = 2; a.iters[r] = -1.f; }
        return;
    }
    ls_error_bound(d, s, M, ap, bp, b, it > 0, lane);
    __syncwarp();
    // ---- sample_pdf(d, bound, N + 2, det)[1:-1]: cdf over the M-1 intervals, into d2 (scratch)
    float wsum = 0.f;
    for (int i = lane; i < M - 1; i += 32) wsum += b[i] + 1e-5f;
    wsum = ls_warp_sum(wsum);
    float carry = 0.f;
    if (lane == 0) d2[0] = 0.f;
    for (int base = 0; base < M - 1; base += 32) {
        const int i = base + lane;
        const bool on = i < M - 1;
        const float pdf = on ? (b[i] + 1e-5f) / wsum : 0.f;
        const float incl = ls_warp_incl_scan(pdf, lane);
        if (on) d2[i + 1] = carry + incl;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    for (int j = lane; j < N; j += 32) {
        const float u = ls_linspace01(j + 1, N + 2);
        const int idx = ls_lower_bound(d2, M, u);
        const int lo = idx - 1 > 0 ? idx - 1 : 0, hi = idx < M - 1 ? idx : M - 1;
        float den = d2[hi] - d2[lo];
        if (den < 1e-5f) den = 1.f;
        Dg[M + j] = d[lo] + (u - d2[lo]) / den * (d[hi] - d[lo]);
    }
    if (lane == 0) {
        const int k = atomicAdd(a.cnt + it + 1, 1);
        a.ray_index[(int64_t)(it + 1) * a.n_rays + k] = r;
    }
}

// ---------------------------------------------------------------- finalize: sort(cat(fine, coarse))
__global__ void ls_sampler_finalize_kernel(const LsSamplerArgs a, float* __restrict__ t_out, float* __restrict__ beta_out,
                                           float* __restrict__ iters_out) {
    LS_DYN_SMEM(smem);
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * wpb + warp;
    if (r >= a.n_rays) return;
    float* co = smem + warp * (a.N + a.Nf);
    float* fi = co + a.N;
    const float tn = a.hits[2 * r], tf = a.hits[2 * r + 1];
    const float ext = ls_fsub(tf, tn);
    for (int i = lane; i < a.N; i += 32) co[i] = ls_fadd(ls_fmul(ls_fdiv((float)i + 0.5f, (float)a.N), ext), tn);
    for (int j = lane; j < a.Nf; j += 32) fi[j] = a.fine[(int64_t)r * a.Nf + j];
    __syncwarp();
    float* out = t_out + (int64_t)r * (a.N + a.Nf);
    // rank of every element of the union by counting (ties broken by position): equal to a stable sort, and safe
    // for rays that miss the box (all coarse depths equal -1) or whose fine samples coincide
    const int T = a.N + a.Nf;
    // both halves sorted (the usual case): the rank is the position inside the own half plus a binary search in the other one
    // (coarse before fine among equals, like the stable sort of the concatenation)
    bool both_sorted = true;
    for (int i = lane; i < T - 1; i += 32) both_sorted = both_sorted && (i == a.N - 1 || co[i] <= co[i + 1]);
    both_sorted = __all_sync(0xffffffffu, both_sorted);
    for (int i = lane; i < T; i += 32) {
        const float v = co[i];
        int c = 0;
        if (both_sorted) {
            if (i < a.N) c = i + ls_lower_bound(fi, a.Nf, v);                      // #{fine < v}
            else {                                                                  // #{coarse <= v}
                int lo = 0, hi = a.N;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (co[mid] <= v) lo = mid + 1; else hi = mid; }
                c = lo + (i - a.N);
            }
        } else {
            for (int k = 0; k < T; ++k) c += (co[k] < v || (co[k] == v && k < i)) ? 1 : 0;
        }
        out[c] = v;
    }
    if (lane == 0) {
        if (beta_out) beta_out[r] = a.beta_plus[r];
        if (iters_out) iters_out[r] = a.iters[r];
    }
}
