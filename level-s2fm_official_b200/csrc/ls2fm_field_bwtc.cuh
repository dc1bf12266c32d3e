// ls2fm_field_bwtc.cuh -- field BACKWARD kernel with every matrix product on the tensor cores (tcgen05 + TMEM, 3xTF32).
//
// Same contract as ls_field_backward_kernel<true, K> (ls2fm_field.cuh): the adjoint of hash grid -> geometry MLP -> normals
// (-> radiance), i.e. a primal + tangent forward pass and a reverse pass over that 2-channel network, the weight / bias /
// W_eff gradients and the hash-table scatter.  What changes is where the arithmetic runs:
//
//   * tile = 64 samples x 2 channels = 128 TMEM lanes.  Lane 32 q + l (q = warp & 3): sample 16 q + (l & 15), channel l >> 4
//     (0 = primal row, 1 = tangent row).  Both channels multiply by the same weights, so ONE batch of tcgen05.mma serves both,
//     and the two rows of a sample sit 16 lanes apart in the same warp: everything the channels owe each other in an epilogue
//     (phi'(z) for the tangent, the phi'' cross term for the primal adjoint) travels by __shfl_xor(.., 16), which also lets the
//     two lanes split the transcendental work of their shared 16 columns evenly.  Warps w, w+4, w+8, w+12 share lane quarter
//     q and split the 64 columns (cg = w >> 2: 16 columns each) and the hash levels.
//   * activations live in TMEM as raw fp32 (A_k, 64 columns each, read by the tensor core as tf32 by truncation); the part
//     the truncation drops goes to a transient 64-column "lo" buffer, D = lo*W_hi + raw*W_lo + raw*W_hi.  The reverse pass
//     overwrites A_k with its adjoint in place.  TMEM map: weight-gradient accumulators [0, 192) | A_1..A_3 [192, 384) |
//     lo [384, 448) | accumulator D [448, 512).  The encoding (layer-0 input) aliases A_H.
//   * weight gradients contract over the tile's 128 rows: both operands come from shared memory, K-major with K = row (the
//     "panel" layout of ls2fm_field.cuh: 4 rows contiguous, feature stride 16 B), raw + lo copies written by the epilogue
//     that produced them; accumulators stay in TMEM for the whole persistent CTA.  The 2-channel sum of the weight gradient
//     IS the contraction over the stacked rows.  The output layer's gradient is accumulated transposed (M = hidden unit) with
//     three extra columns carrying the radiance pre-activation gradient, which yields the geo-feature block of dL/dW_eff
//     without ever computing the output layer in this kernel.
//   * weights: W_l (forward) and W_l^T (reverse) as hi/lo K-major operands would need 200 KB (tf32 operands cannot be read
//     transposed: every MN-major descriptor variant returns zeros, tools/probe/tc_probe4.cu); they stream instead from the
//     L2-resident operand image (ls2fm_field_prepare) through a 3-slot ring of 16 KB half-matrices (hi, then lo) with
//     cp.async.bulk + mbarrier, one matrix per MMA batch, fetched one to two batches ahead by the issuing lane.
//   * MMA issue: one elect.sync lane of the converged warp 0 (uniform-register descriptors, back-to-back UTCHMMA).
//   * per tile: H forward batches, 1 + H reverse batches.  A reverse batch commits the product into the layer below first
//     (critical path, mbarrier `bar`) and the layer's weight-gradient MMAs second on their own mbarrier (`bar2`), which is only
//     waited for before the staging arrays are rewritten.  No output-layer forward, no separate bias pass (layer 0's bias
//     gradient rides on the ones column of the encoding, the hidden layers' are 15-shuffle column sums of the adjoints, the
//     output layer's are per-thread partial sums).
//   * cross-tile software pipeline, same warps: the next tile's per-sample loads (PBN, 3 tiles deep) run under the last forward
//     batch, its hash gather -- four z-plane pieces of four corner loads -- under B5 and the reverse batches, and the previous
//     tile's table-gradient scatter (adjoints parked in 8 registers) under the next tile's forward batches.
//   * shared memory: staging 4 x 33 KB | encoding stash 25 KB | PBN 16 KB | ring 48 KB | biases, W_eff, barriers: 226 KB.
#pragma once

#include "ls2fm_field_tc.cuh"

constexpr int LS_BT_THREADS = 512;
constexpr int LS_BT_TILE = 64;                  // samples per tile
constexpr int64_t LS_BT_MIN_SAMPLES = 8192;     // automatic dispatch: smaller launches run the fp32-SIMT kernel
constexpr int LS_BT_KG = 32;                    // groups of 4 rows in a staging array (128 rows)
constexpr int LS_BT_LBO = LS_H * 16 + 16;       // bytes between row groups of a 64-feature staging array (+16: conflict-free stores)
constexpr int LS_BT_LBOF = LS_BT_LBO / 4;       // ... in floats (260)
constexpr int LS_BT_EROWS = 48;                 // feature rows of the encoding stash (35 + ones -> 40, MMA N = 48)
constexpr int LS_BT_LBO_E = LS_BT_EROWS * 16 + 16;
constexpr int LS_BT_LBOF_E = LS_BT_LBO_E / 4;   // 196
constexpr int LS_BT_SLOT = 4096;                // floats per ring slot (16 KB: the hi OR the lo half of a 64 x 64 operand)
constexpr int LS_BT_PBP = 21;                   // pitch of a per-sample staging row (odd: conflict-free across the 16 samples of a warp)
// TMEM columns
constexpr int LS_BT_WG = 0;                     // weight-gradient accumulators: last^T [0,32) | layer 0 [24,72) | layer l [64 l, +64)
constexpr int LS_BT_A = 192;                    // A_k at LS_BT_A + 64 (k - 1)
constexpr int LS_BT_LO = 384;
constexpr int LS_BT_D = 448;

struct LsBtNet {         // shared-memory plan (float offsets)
    int zr, zl, ar, al;  // staging: adjoint raw / lo, layer input raw / lo           [32 groups][64 features][4 rows] padded
    int es;              // encoding stash (raw), operand of the layer-0 weight gradient [32 groups][48 features][4 rows] padded
    int pbn;             // per-sample upstream values, 3 tiles deep (previous: deferred scatter, current, next: prefetch) [3][64][16]
    int ring;            // 3 slots
    int bias[LS2FM_MAX_LAYERS];
    int weff, rad_pitch;
    int gs;              // [3][64] scratch of the final flush
    int misc;            // barriers + TMEM slot
    int total;
};

inline LsBtNet ls_plan_bt(const ls2fm_field_t& f, int rad_in_dim) {
    LsBtNet n;
    memset(&n, 0, sizeof(n));
    int off = 0;
    const int stage = LS_BT_KG * LS_BT_LBOF;
    n.zr = off; off += stage;
    n.zl = off; off += stage;
    n.ar = off; off += stage;
    n.al = off; off += stage;
    n.es = off; off += LS_BT_KG * LS_BT_LBOF_E;
    n.pbn = ls_round4(off); off = n.pbn + 3 * LS_BT_TILE * LS_BT_PBP;
    n.ring = off; off += 3 * LS_BT_SLOT;
    for (int l = 0; l < f.n_layers; ++l) { n.bias[l] = off; off += LS_H; }
    n.rad_pitch = ls_round4(rad_in_dim > 0 ? rad_in_dim : 4);
    n.weff = off; off += 3 * n.rad_pitch + 4;
    n.gs = off; off += 3 * LS_H;
    n.misc = ls_round4(off); off = n.misc + 16;
    n.total = off;
    return n;
}

// the ring's schedule: batch b of a tile uses matrix b.  b < H: W_b (forward) | b == H: output layer transposed | b > H: W_l^T, l = 2H - b
LS_DEV void ls_bt_matrix(const LsTcNet& img, int H, int b, int* src, int* floats) {
    if (b < H) { *src = img.w_hi[b]; *floats = 2 * img.n_out_pad[b] * img.k_in_pad[b]; }
    else if (b == H) { *src = img.wtl_hi; *floats = 2 * LS_H * img.kl_pad; }
    else { const int l = 2 * H - b; *src = img.wt_hi[l]; *floats = 2 * img.n_in_pad[l] * LS_H; }
}

// weight-gradient batch: D[m][d_col + n] += sum_r A(m, r) B(n, r) over the tile's 128 rows, 3xTF32 (raw = hi by truncation).
// (Rounding one operand to tf32 instead of carrying its lo part was tried: the 2^-12 noise per term does not average out against
//  the gradient -- a sum of large cancelling terms -- and showed up as 3e-3 of the output layer's weight_g gradient on 16 k samples.)
LS_DEV void ls_bt_wgrad(uint32_t tmem, int d_col, const float* a_raw, const float* a_lo, int a_lbo, const float* b_raw, const float* b_lo,
                        int b_lbo, int N, int skip = 0) {
    if (skip) return;
#if defined(LS_HOSTSIM)
    for (int ks = 0; ks < LS_BT_KG / 2; ++ks) {      // 8 rows (two groups of 4) per MMA
        const int ao = ks * 2 * (a_lbo / 4), bo = ks * 2 * (b_lbo / 4);
        ls_tc_mma_ss(tmem, d_col, a_lo + ao, a_lbo, b_raw + bo, b_lbo, N, true);
        ls_tc_mma_ss(tmem, d_col, a_raw + ao, a_lbo, b_lo + bo, b_lbo, N, true);
        ls_tc_mma_ss(tmem, d_col, a_raw + ao, a_lbo, b_raw + bo, b_lbo, N, true);
    }
#else
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(LS_TC_M >> 4) << 24);
    uint64_t ar = ls_tc_desc(ls_smem_u32(a_raw), (uint32_t)a_lbo, 128), al = ls_tc_desc(ls_smem_u32(a_lo), (uint32_t)a_lbo, 128);
    uint64_t br = ls_tc_desc(ls_smem_u32(b_raw), (uint32_t)b_lbo, 128), bl = ls_tc_desc(ls_smem_u32(b_lo), (uint32_t)b_lbo, 128);
    const uint64_t as = (uint64_t)(2 * a_lbo >> 4), bs = (uint64_t)(2 * b_lbo >> 4);     // 8 rows (two groups of 4) per MMA
    const uint32_t d = tmem + (uint32_t)d_col;
#define LS_BT_SS(A, B) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" \
                                    :: "r"(d), "l"(A), "l"(B), "r"(idesc))
#pragma unroll
    for (int ks = 0; ks < LS_BT_KG / 2; ++ks) {
        LS_BT_SS(al, br);
        LS_BT_SS(ar, bl);
        LS_BT_SS(ar, br);
        ar += as; al += as; br += bs; bl += bs;
    }
#undef LS_BT_SS
#endif
}

// ================================================================ position gradients behind the tensor-core backward
// The tensor-core kernel has the adjoints of the encoding (ebar: primal channel, ebar_dot: tangent channel) in registers at the end of
// a tile but no registers left for another table gather; it parks them in a workspace (LS_PG_PITCH floats per sample:
// [0,32) ebar of the hash features, [32,64) ebar_dot, [64,67) ebar of x / rescale) and this kernel -- one thread per sample, 64 warps
// per SM, memory-bound like the stand-alone encoding -- gathers the table once more and contracts:
//   dL/dx = Je^T ebar + (d(Je nbar)/dx)^T ebar_dot + W_eff[:,0:3]^T pbar      (+ the Fourier-embedding term for d_ray)
// Same arithmetic as phase B6 of ls_field_backward_kernel.  Ray mode: the 32 samples of a warp usually belong to one ray, so the
// warp reduces first and issues one atomic per component.
constexpr int LS_PG_PITCH = 72;
__global__ void __launch_bounds__(256) ls_field_posgrad_kernel(const LsFieldArgs a) {
    const float* ws = a.ig.workspace;
    const int L = a.f.n_levels;
    const bool rad = a.r.w_eff != nullptr && a.g_rgb != nullptr;
    const int in_dim = a.r.in_dim, nf = a.r.n_freq;
    const int o_ray = 6;
    const int64_t n_pad = (a.p.n + 31) / 32 * 32;
    for (int64_t i_in = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i_in < n_pad; i_in += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = i_in < a.p.n;
        float x[3] = {0.f, 0.f, 0.f}, u[3];
        int ray_id = 0;
        int64_t i = i_in;
        if (valid) ls_sample_point(a.p, i_in, x, &ray_id, &i);
        ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u);
        float pbar[3] = {0.f, 0.f, 0.f}, nbar[3] = {0.f, 0.f, 0.f}, dx[3] = {0.f, 0.f, 0.f}, dirg[3] = {0.f, 0.f, 0.f};
        if (valid) {
            if (rad) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float rgb = __ldg(a.saved_rgb + 3 * i + c);
                    pbar[c] = __ldg(a.g_rgb + 3 * i + c) * rgb * (1.f - rgb);
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float v = a.g_nrm ? __ldg(a.g_nrm + 3 * i + d) : 0.f;
                if (rad) v += __ldg(a.r.w_eff + 3 + d) * pbar[0] + __ldg(a.r.w_eff + in_dim + 3 + d) * pbar[1] + __ldg(a.r.w_eff + 2 * in_dim + 3 + d) * pbar[2];
                nbar[d] = v;
            }
            const float* w = ws + i * LS_PG_PITCH;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const float scale = a.f.levels[l].scale;
                const uint32_t res = a.f.levels[l].resolution, size = a.f.levels[l].size, hashed = a.f.levels[l].hashed;
                const float2* vt = reinterpret_cast<const float2*>(a.f.table) + a.f.levels[l].offset;
                const LsCell c = ls_cell(scale, u);
                uint32_t ci[8];
                ls_corner_indices<0, 8>(res, size, hashed, c, ci);
                float2 tv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) tv[k] = __ldg(vt + ci[k]);
                const float e0 = w[2 * l], e1 = w[2 * l + 1], t0 = w[32 + 2 * l], t1 = w[32 + 2 * l + 1];
                float ns[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) ns[d] = nbar[d] * scale * a.inv_ext[d];
                float lx[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float fk[3] = {(k & 1) ? c.w[0] : 1.f - c.w[0], (k & 2) ? c.w[1] : 1.f - c.w[1], (k & 4) ? c.w[2] : 1.f - c.w[2]};
                    const float sg3[3] = {(k & 1) ? 1.f : -1.f, (k & 2) ? 1.f : -1.f, (k & 4) ? 1.f : -1.f};
                    const float Ak = tv[k].x * e0 + tv[k].y * e1, Bk = tv[k].x * t0 + tv[k].y * t1;
                    lx[0] += sg3[0] * (fk[1] * fk[2] * Ak + (sg3[1] * ns[1] * fk[2] + sg3[2] * ns[2] * fk[1]) * Bk);
                    lx[1] += sg3[1] * (fk[2] * fk[0] * Ak + (sg3[2] * ns[2] * fk[0] + sg3[0] * ns[0] * fk[2]) * Bk);
                    lx[2] += sg3[2] * (fk[0] * fk[1] * Ak + (sg3[0] * ns[0] * fk[1] + sg3[1] * ns[1] * fk[0]) * Bk);
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) dx[d] += lx[d] * scale * a.inv_ext[d];
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                dx[d] += w[64 + d] / a.f.rescale;
                if (rad) dx[d] += __ldg(a.r.w_eff + d) * pbar[0] + __ldg(a.r.w_eff + in_dim + d) * pbar[1] + __ldg(a.r.w_eff + 2 * in_dim + d) * pbar[2];
            }
            if (rad && a.ig.d_ray) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float dv = __ldg(a.p.ray + 3 * ray_id + d);
                    for (int c = 0; c < 3; ++c) {
                        float acc = __ldg(a.r.w_eff + c * in_dim + o_ray + d);
                        for (int k = 0; k < nf; ++k) {
                            const float fr = (float)(1 << k);
                            acc += fr * (__ldg(a.r.w_eff + c * in_dim + o_ray + 3 + 6 * k + d) * cosf(dv * fr) -
                                         __ldg(a.r.w_eff + c * in_dim + o_ray + 6 + 6 * k + d) * sinf(dv * fr));
                        }
                        dirg[d] += acc * pbar[c];
                    }
                }
            }
        }
        if (a.p.xyz) {
            if (valid && a.ig.d_xyz) { a.ig.d_xyz[3 * i] = dx[0]; a.ig.d_xyz[3 * i + 1] = dx[1]; a.ig.d_xyz[3 * i + 2] = dx[2]; }
            if (valid && a.ig.d_ray && rad) {
#pragma unroll
                for (int d = 0; d < 3; ++d) atomicAdd(a.ig.d_ray + 3 * ray_id + d, dirg[d]);
            }
        } else {
            float tv = 0.f;
            if (valid) {
                const int j = (int)(i_in - (int64_t)ray_id * a.p.n_per_ray);
                tv = __ldg(a.p.t + (int64_t)ray_id * a.p.t_stride + a.p.t_offset + j);
                if (a.ig.d_t) {
                    float dt = 0.f;
#pragma unroll
                    for (int d = 0; d < 3; ++d) dt += __ldg(a.p.ray + 3 * ray_id + d) * dx[d];
                    a.ig.d_t[i_in] = dt;
                }
            }
            // one ray per warp (the common case: n_per_ray a multiple of 32): reduce first, one atomic per component
            const int r0 = __shfl_sync(0xffffffffu, ray_id, 0);
            const bool uniform = __all_sync(0xffffffffu, !valid || ray_id == r0) && __shfl_sync(0xffffffffu, valid ? 1 : 0, 0);
            float acc[6] = {dx[0], dx[1], dx[2], tv * dx[0] + dirg[0], tv * dx[1] + dirg[1], tv * dx[2] + dirg[2]};
            if (uniform) {
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    float v = valid ? acc[q] : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    acc[q] = v;
                }
                if ((threadIdx.x & 31) == 0) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        if (a.ig.d_center) atomicAdd(a.ig.d_center + 3 * r0 + d, acc[d]);
                        if (a.ig.d_ray) atomicAdd(a.ig.d_ray + 3 * r0 + d, acc[3 + d]);
                    }
                }
            } else if (valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (a.ig.d_center) atomicAdd(a.ig.d_center + 3 * ray_id + d, acc[d]);
                    if (a.ig.d_ray) atomicAdd(a.ig.d_ray + 3 * ray_id + d, acc[3 + d]);
                }
            }
        }
    }
}

template <int K>
__global__ void __launch_bounds__(LS_BT_THREADS, 1) ls_field_backward_tc_kernel(const LsFieldArgs a, const LsTcNet img, const LsBtNet net) {
    LS_DYN_SMEM(smem);
    constexpr int H = K - 1;            // hidden layers
    constexpr int NB = 2 * H + 1;       // MMA batches (= ring matrices) per tile
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int q = warp & 3, cg = warp >> 2;
    const int isT = lane >> 4;                      // 0: primal row, 1: tangent row
    const int sl = 16 * q + (lane & 15);            // sample inside the tile
    const int row = 32 * q + lane;                  // TMEM lane = row index of the staging arrays
    const int L = a.f.n_levels, nh = 2 * L;
    const int dout = a.f.dims[K];
    const float sp_beta = a.f.softplus_beta, sp_thr = a.f.softplus_threshold, inv_beta = 1.f / a.f.softplus_beta;
    const bool rad = a.r.w_eff != nullptr;
    const int nf = a.r.n_freq, kg = a.r.k_geo, kg2 = a.r.k_geo2;
    const int o_geo = 6 + 3 + 6 * nf, o_geo2 = o_geo + kg;
    const int RP = net.rad_pitch;
    const int nin = rad ? a.r.in_dim - kg : 0;      // radiance inputs that are not geo features: x, n, dir, Fourier, (geo2)
    const float* Weff = smem + net.weff;
    float* ZR = smem + net.zr; float* ZL = smem + net.zl; float* AR = smem + net.ar; float* AL = smem + net.al;
    float* ES = smem + net.es;
    float* RIN = ZR;                                // [64][nin] radiance inputs of the tile (dead before the staging arrays are written)
    float* PB = ZL;                                 // [64][4]  radiance pre-activation gradients

    // ------------------------------------------------ one-time setup
    for (int l = 1; l < K - 1; ++l)
        for (int e = t; e < LS_H; e += LS_BT_THREADS) smem[net.bias[l] + e] = __ldg(a.f.tc_image + img.bias[l] + e);
    if (rad) {
        float* W = smem + net.weff;
        for (int e = t; e < 3 * RP; e += LS_BT_THREADS) {
            const int c = e / RP, ii = e - c * RP;
            W[e] = ii < a.r.in_dim ? __ldg(a.r.w_eff + c * a.r.in_dim + ii) : 0.f;
        }
    }
    for (int e = t; e < LS_BT_KG * LS_BT_LBOF_E; e += LS_BT_THREADS) ES[e] = 0.f;     // feature rows >= k_in_pad stay zero for ever
    LsTcBar* bar = reinterpret_cast<LsTcBar*>(smem + net.misc);
    LsTcBar* bar2 = reinterpret_cast<LsTcBar*>(smem + net.misc + 8);     // weight-gradient batches (off the critical path)
    const uint32_t tmem = ls_tc_alloc(reinterpret_cast<uint32_t*>(smem + net.misc + 10));
    ls_tc_bar_init(bar);
    ls_tc_bar_init(bar2);
    if (t == 0) { for (int s3 = 0; s3 < 3; ++s3) ls_bar_init1(reinterpret_cast<LsTcBar*>(smem + net.misc + 2 + 2 * s3)); }
    ls_fence_smem_to_async();
    __syncthreads();
    uint32_t phase = 0, phase2 = 0;
    bool wg_pending = false;            // a weight-gradient batch may still be reading the staging arrays / the stash
    {   // zero the weight-gradient accumulators (every lane: the upper 64 are scratch of the M = 128 instruction shape)
        float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 48; c += 8) ls_tmem_st(tmem, LS_BT_WG + 48 * cg + c, z8, 8);
    }

    const int64_t n_tiles = (a.p.n + LS_BT_TILE - 1) / LS_BT_TILE;
    const int64_t my_tiles = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_batches = my_tiles * NB;
    int64_t gb = 0;                     // running batch index
    // ring entries: 2 per batch (the hi half of its matrix, then the lo half), entry e lives in slot e % 3 and completes full[e % 3]
    // for the (e / 3)-th time.  Entries 0..2 are fetched up front; when batch g is done its two slots are refilled with entries
    // 2g+3 (the lo half of batch g+1's matrix) and 2g+4 (the hi half of batch g+2's).
    auto ring_fetch = [&](int64_t e) {  // one thread
        if (e < 2 * n_batches) {
            int src, floats;
            ls_bt_matrix(img, H, (int)((e >> 1) % NB), &src, &floats);
            const int half = floats / 2, s3 = (int)(e % 3);
            ls_bulk_g2s(smem + net.ring + s3 * LS_BT_SLOT, a.f.tc_image + src + (int)(e & 1) * half, half * 4,
                        reinterpret_cast<LsTcBar*>(smem + net.misc + 2 + 2 * s3));
        }
    };
    if (t == 0) { ring_fetch(0); ring_fetch(1); ring_fetch(2); }
    // the issuing lane: wait for entry e; returns its slot
    auto ring_slot = [&](int64_t e) -> const float* {
        const int s3 = (int)(e % 3);
        ls_bar_wait(reinterpret_cast<LsTcBar*>(smem + net.misc + 2 + 2 * s3), (uint32_t)((e / 3) & 1));
        return smem + net.ring + s3 * LS_BT_SLOT;
    };
    // 3xTF32 product of batch gb by the issuing lane: D = A_lo W_hi + A_raw W_hi, then (second ring entry) + A_raw W_lo
    auto issue_x3 = [&](int d_col, int a_raw_col, int a_lo_col, int N, int Kd) {
        const float* Wh = ring_slot(2 * gb);
        ls_tc_mma(tmem, d_col, a_lo_col, Wh, N, Kd, false);
        ls_tc_mma(tmem, d_col, a_raw_col, Wh, N, Kd, true);
        const float* Wl = ring_slot(2 * gb + 1);
        ls_tc_mma(tmem, d_col, a_raw_col, Wl, N, Kd, true);
    };
    // all threads: wait for batch gb, then one lane refills its two slots
    auto batch_done = [&]() {
        ls_tc_wait(bar, phase);
        if (warp == 0) { if (ls_elect()) { ring_fetch(2 * gb + 3); ring_fetch(2 * gb + 4); } }
        ++gb;
    };
    // persistent accumulators
    float bacc[LS2FM_MAX_LAYERS];
#pragma unroll
    for (int l = 0; l < LS2FM_MAX_LAYERS; ++l) bacc[l] = 0.f;
    float weff_acc = 0.f;
    float blast[8];                     // output-layer bias gradient: this thread's 8 columns of ybar, summed over its tiles
#pragma unroll
    for (int k = 0; k < 8; ++k) blast[k] = 0.f;
    const int st_off = (row >> 2) * LS_BT_LBOF + (row & 3);        // + feature * 4
    const int st_off_e = (row >> 2) * LS_BT_LBOF_E + (row & 3);
    const int colE = LS_BT_A + 64 * (H - 1);

    // ------------------------------------------------ software pipeline across tiles
    // The memory work of a tile is issued under the tensor-core batches of its neighbours, by the same warps:
    //   under B5 of tile n  : the per-sample loads of tile n+1 (position, upstream gradients, saved outputs) -> PBN[next]
    //   under B6 of tile n  : the hash-grid gather of tile n+1 (this thread's two levels) -> 8 registers
    //   under B4 of tile n+1: the hash-table gradient scatter of tile n (adjoints kept in 8 registers, cell from PBN[previous])
    // PBN row: 0..2 pbar | 3..5 nbar (upstream normal gradient incl. the radiance path) | 6..8 position | 9..11 ray direction | 12..14 saved normal | 15 valid |
    //          16..18 unit-cube coordinates (computed once here: every gather / scatter piece needs them)
    float* PBN = smem + net.pbn;
    int pb_prev = 2, pb_cur = 0, pb_nxt = 1;
    auto stage_sample = [&](int64_t tile, int buf) {        // raw per-sample loads, spread over the sample's 8 threads
        float* P = PBN + (buf * LS_BT_TILE + sl) * LS_BT_PBP;
        const int64_t i = tile * LS_BT_TILE + sl;           // (dense launches: output index == input index)
        const bool valid = tile < n_tiles && i < a.p.n;
        float v3[3] = {0.f, 0.f, 0.f};
        int base = -1;
        if (!isT) {
            if (cg == 0) {
                if (valid) { int ray_id; int64_t io; ls_sample_point(a.p, i, v3, &ray_id, &io); }
                base = 6;
                P[15] = valid ? 1.f : 0.f;
                float u3[3];
                ls_world_to_unit(a.f.bound_min, a.f.bound_max, v3, u3);
                P[16] = u3[0]; P[17] = u3[1]; P[18] = u3[2];
            } else if (cg == 1) {
                if (valid && rad && a.g_rgb) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float rgb = __ldg(a.saved_rgb + 3 * i + c);
                        v3[c] = __ldg(a.g_rgb + 3 * i + c) * rgb * (1.f - rgb);
                    }
                }
                base = 0;
            } else if (cg == 2) {       // nbar = upstream normal gradient + W_eff[:, 3:6]^T pbar (its own copy of pbar: no barrier needed)
                if (valid) {
                    float pb[3] = {0.f, 0.f, 0.f};
                    if (rad && a.g_rgb) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float rgb = __ldg(a.saved_rgb + 3 * i + c);
                            pb[c] = __ldg(a.g_rgb + 3 * i + c) * rgb * (1.f - rgb);
                        }
                    }
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        float v = a.g_nrm ? __ldg(a.g_nrm + 3 * i + d) : 0.f;
                        if (rad) v += Weff[3 + d] * pb[0] + Weff[RP + 3 + d] * pb[1] + Weff[2 * RP + 3 + d] * pb[2];
                        v3[d] = v;
                    }
                }
                base = 3;
            } else {
                if (valid && rad) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) v3[d] = __ldg(a.saved_nrm + 3 * i + d);
                }
                base = 12;
            }
        } else if (cg == 0) {
            if (valid && rad) {
                const int64_t r = a.p.n <= 0x7fffffffLL ? (int64_t)((uint32_t)i / (uint32_t)a.p.n_per_ray) : i / a.p.n_per_ray;
#pragma unroll
                for (int d = 0; d < 3; ++d) v3[d] = __ldg(a.p.ray + 3 * r + d);
            }
            base = 9;
        }
        if (base >= 0) { P[base] = v3[0]; P[base + 1] = v3[1]; P[base + 2] = v3[2]; }
    };
    auto load_nbar = [&](const float* P, float (&nbar)[3]) { nbar[0] = P[3]; nbar[1] = P[4]; nbar[2] = P[5]; };
    float e4[4], d4[4];                 // gather in flight (next tile)
    float pl[2][3];                     // ... its z = 0 plane of the level being evaluated: value, d/dx, d/dy per feature
    // The gather of a tile is cut into four pieces, one per tensor-core batch it hides under: piece p = (level r = p >> 1 of this
    // thread's two levels 4 cg + 2 isT + r, z-plane p & 1 of the cell: 4 of the 8 corner loads).  Same arithmetic, term by term,
    // as ls_level_eval.  Piece 2r+1 completes level r: features e and tangent features Je nbar.
    auto gather_piece = [&](int buf, int p) {
        const int r = p >> 1, plane = p & 1;
        const float* P = PBN + (buf * LS_BT_TILE + sl) * LS_BT_PBP;
        const float u[3] = {P[16], P[17], P[18]};
        const int l = 4 * cg + 2 * isT + r;
        const float scale = a.f.levels[l].scale;
        const uint32_t res = a.f.levels[l].resolution, size = a.f.levels[l].size, hashed = a.f.levels[l].hashed;
        const float2* tab = reinterpret_cast<const float2*>(a.f.table) + a.f.levels[l].offset;
        const LsCell c = ls_cell(scale, u);
        float2 v[4];
        uint32_t ci[4];
        if (plane) ls_corner_indices<4, 4>(res, size, hashed, c, ci);
        else ls_corner_indices<0, 4>(res, size, hashed, c, ci);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = LS_DBG(a, 1) ? make_float2(0.01f * (float)(ci[k] & 7), 0.02f) : __ldg(tab + ci[k]);
        const float w0 = c.w[0], w1 = c.w[1], w2 = c.w[2];
        const float m0 = 1.f - w0, m1 = 1.f - w1, m2 = 1.f - w2;
        float nbar[3];
        if (plane) load_nbar(P, nbar);
#pragma unroll
        for (int fi = 0; fi < 2; ++fi) {
            const float q0 = fi ? v[0].y : v[0].x, q1 = fi ? v[1].y : v[1].x, q2 = fi ? v[2].y : v[2].x, q3 = fi ? v[3].y : v[3].x;
            const float a0 = m0 * q0 + w0 * q1, a1 = m0 * q2 + w0 * q3;
            const float b = m1 * a0 + w1 * a1, gx = m1 * (q1 - q0) + w1 * (q3 - q2), gy = a1 - a0;
            if (!plane) { pl[fi][0] = b; pl[fi][1] = gx; pl[fi][2] = gy; }
            else {
                const float dh0 = scale * (m2 * pl[fi][1] + w2 * gx), dh1 = scale * (m2 * pl[fi][2] + w2 * gy), dh2 = scale * (b - pl[fi][0]);
                e4[2 * r + fi] = m2 * pl[fi][0] + w2 * b;
                d4[2 * r + fi] = dh0 * a.inv_ext[0] * nbar[0] + dh1 * a.inv_ext[1] * nbar[1] + dh2 * a.inv_ext[2] * nbar[2];
            }
        }
    };
    float c8n[8];                       // gathered chunk of the next tile: primal row e / tangent row edot, columns 8 cg .. 8 cg + 7
    auto gather_finish = [&]() {        // the two rows of a sample swap halves
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float r = __shfl_xor_sync(0xffffffffu, isT ? e4[k] : d4[k], 16);
            c8n[k] = isT ? r : e4[k];           // primal row: features of levels 4cg, 4cg+1 (own), 4cg+2, 4cg+3 (partner)
            c8n[4 + k] = isT ? d4[k] : r;       // tangent row: edot likewise
        }
    };
    float sb[4], sd[4];                 // deferred scatter (previous tile): adjoints of my two levels' features / tangent features
    bool sc_pending = false;
    auto scatter_level = [&](int buf, int r, float e0, float e1, float t0, float t1) {
        const float* P = PBN + (buf * LS_BT_TILE + sl) * LS_BT_PBP;
        if (!a.d_table || P[15] == 0.f || LS_DBG(a, 0)) return;
        const float u[3] = {P[16], P[17], P[18]};
        float nbar[3];
        load_nbar(P, nbar);
        const int l = 4 * cg + 2 * isT + r;
        const float scale = a.f.levels[l].scale;
        const uint32_t res = a.f.levels[l].resolution, size = a.f.levels[l].size, hashed = a.f.levels[l].hashed;
        float* tab = a.d_table + 2 * (size_t)a.f.levels[l].offset;
        const LsCell c = ls_cell(scale, u);
        float ns[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) ns[d] = nbar[d] * scale * a.inv_ext[d];
        uint32_t ci[8];
        ls_corner_indices<0, 8>(res, size, hashed, c, ci);
        const bool v4 = (reinterpret_cast<uintptr_t>(a.d_table) & 15) == 0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) {        // x-neighbour pairs: one 16-byte atomic when the two entries are an aligned pair
            const float f1 = (k & 2) ? c.w[1] : 1.f - c.w[1];
            const float f2 = (k & 4) ? c.w[2] : 1.f - c.w[2];
            const float m0 = 1.f - c.w[0], w0 = c.w[0];
            const float dyz = ((k & 2) ? ns[1] : -ns[1]) * f2 + ((k & 4) ? ns[2] : -ns[2]) * f1;       // d(f1 f2) along nbar
            const float wa = m0 * f1 * f2, wb = w0 * f1 * f2;
            const float da = -ns[0] * f1 * f2 + m0 * dyz, db = ns[0] * f1 * f2 + w0 * dyz;
            ls_red_pair(tab, ci[k], ci[k + 1], wa * e0 + da * t0, wa * e1 + da * t1, wb * e0 + db * t0, wb * e1 + db * t1, v4);
        }
    };
    const bool has_levels = 4 * cg < L;
    // prologue: the first tile's loads and gather, unhidden
    if (my_tiles > 0) {
        stage_sample(blockIdx.x, pb_cur);
        __syncthreads();
        if (has_levels) {
            gather_piece(pb_cur, 0); gather_piece(pb_cur, 1); gather_piece(pb_cur, 2); gather_piece(pb_cur, 3);
            gather_finish();
        }
    }

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const bool has_next = tile + gridDim.x < n_tiles;
        // ------------------------------------------------ B1: this tile's encoding (gathered under the previous tile) -> TMEM + stash
        if (wg_pending) { ls_tc_wait(bar2, phase2); wg_pending = false; }    // the stash / staging arrays are rewritten below
        const float* Pc = PBN + (pb_cur * LS_BT_TILE + sl) * LS_BT_PBP;
        const int64_t i = tile * LS_BT_TILE + sl;
        const bool valid = Pc[15] != 0.f;
        if (has_levels) {
            float lo8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                lo8[k] = ls_tf32_lo(c8n[k]);
                ES[st_off_e + (8 * cg + k) * 4] = c8n[k];
            }
            ls_tmem_st(tmem, colE + 8 * cg, c8n, 8);
            ls_tmem_st(tmem, LS_BT_LO + 8 * cg, lo8, 8);
        }
        if (cg == 0) {      // tail chunk: x / rescale (tangent row: nbar / rescale), the ones column (primal rows only), zero padding
            float nbar[3];
            load_nbar(Pc, nbar);
            float c8[8] = {0.f, 0.f, 0.f, isT ? 0.f : 1.f, 0.f, 0.f, 0.f, 0.f}, lo8[8];
#pragma unroll
            for (int d = 0; d < 3; ++d) c8[d] = isT ? nbar[d] / a.f.rescale : ls_fdiv(Pc[6 + d], a.f.rescale);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                lo8[k] = ls_tf32_lo(c8[k]);
                ES[st_off_e + (nh + k) * 4] = c8[k];
            }
            ls_tmem_st(tmem, colE + nh, c8, 8);
            ls_tmem_st(tmem, LS_BT_LO + nh, lo8, 8);
        }
        // ------------------------------------------------ B2/B3: W_eff gradient, non-geo inputs (x, n, dir, Fourier, geo2) and b_eff
        if (rad) {
            float* in = RIN + sl * nin;
            if (!isT) {
                if (cg == 0) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { in[d] = Pc[6 + d]; PB[4 * sl + d] = Pc[d]; }
                } else if (cg == 1) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) in[3 + d] = Pc[12 + d];
                } else if (cg == 2) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) in[6 + d] = Pc[9 + d];
                } else {
                    for (int k = 0; k < kg2; ++k) in[o_geo + k] = valid ? __ldg(a.r.geo2 + i * (kg2 + 1) + 1 + k) : 0.f;
                }
                if (a.d_geo2 && valid) {
                    for (int k = cg; k <= kg2; k += 4) {
                        float v = 0.f;
                        if (k >= 1) {
                            const int col = o_geo2 + k - 1;
                            v = Weff[col] * Pc[0] + Weff[RP + col] * Pc[1] + Weff[2 * RP + col] * Pc[2];
                        }
                        a.d_geo2[i * (kg2 + 1) + k] = v;
                    }
                }
            } else {
                for (int k = cg; k < nf; k += 4) {
                    const float fr = (float)(1 << k);
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float ang = Pc[9 + d] * fr;
                        ls_sincos_fast(ang, &in[9 + 6 * k + d], &in[12 + 6 * k + d]);
                    }
                }
            }
            // (reduced under the forward batches below, after the barrier in front of the first one)
        }
        // ------------------------------------------------ B4: forward, both channels per batch
#pragma unroll
        for (int l = 0; l < H; ++l) {
            const int a_col = l == 0 ? colE : LS_BT_A + 64 * (l - 1);
            ls_tc_sync_before_mma();
            if (warp == 0) {
                if (ls_elect()) {
                    issue_x3(LS_BT_D, a_col, LS_BT_LO, LS_H, img.k_in_pad[l]);
                    ls_tc_commit(bar);
                }
            }
            if (rad && l == 0 && t < 3 * (3 * nin + 3)) {
                // W_eff / b_eff gradient of this tile (RIN / PB were completed before this batch's barrier): (channel, input) pairs x
                // three sample ranges spread over the whole CTA, so no warp carries a 64-long dependent chain into the next barrier
                const int n_items = 3 * nin + 3;
                const int part = t / n_items, tt = t - part * n_items;
                const int c = tt < 3 * nin ? tt / nin : tt - 3 * nin;
                const int idx = tt < 3 * nin ? tt - c * nin : -1;
                const int s0 = part * 22, s1 = s0 + 22 < LS_BT_TILE ? s0 + 22 : LS_BT_TILE;
                float acc0 = 0.f, acc1 = 0.f;
                for (int s = s0; s + 1 < s1; s += 2) {      // (22, 22 and 20 samples: all even)
                    acc0 = fmaf(PB[4 * s + c], idx >= 0 ? RIN[s * nin + idx] : 1.f, acc0);
                    acc1 = fmaf(PB[4 * s + 4 + c], idx >= 0 ? RIN[(s + 1) * nin + idx] : 1.f, acc1);
                }
                weff_acc += acc0 + acc1;
            }
            if (l == H - 1 && has_next) stage_sample(tile + gridDim.x, pb_nxt);     // next tile's per-sample loads run under this batch
            if (sc_pending && has_levels) {     // the previous tile's table-gradient scatter runs under this batch
                if (l == 0) scatter_level(pb_prev, 0, sb[0], sb[1], sd[0], sd[1]);
                if (l == (H > 1 ? 1 : 0)) scatter_level(pb_prev, 1, sb[2], sb[3], sd[2], sd[3]);
            }
            batch_done();
            // epilogue: the primal lane takes columns 0..7 of the pair's 16, the tangent lane 8..15 (both channels of those)
            float xv[16];
            ls_tmem_ld(tmem, LS_BT_D + 16 * cg, xv, 16);
            float av[8], dv[8];
            const float* bias = smem + net.bias[l] + 16 * cg + 8 * isT;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float r = __shfl_xor_sync(0xffffffffu, isT ? xv[k] : xv[8 + k], 16);
                float z = isT ? r : xv[k];                       // primal pre-activation of my column
                const float zd = isT ? xv[8 + k] : r;            // tangent pre-activation
                if (l > 0) z += bias[k];                         // (layer 0: the bias rides on the ones column)
                const float bz = z * sp_beta;
#if defined(LS_HOSTSIM)
                if (bz > sp_thr) { av[k] = z; dv[k] = zd; }
                else {
                    const float e = expf(bz);
                    av[k] = log1pf(e) * inv_beta;
                    dv[k] = zd * (e / (e + 1.f));
                }
#else
                // branch-free (see ls_softplus_fast): phi = log(1 + e) / beta, phi' = e / (1 + e), threshold zone by select
                const float e = ls_ex2(fminf(bz, sp_thr) * 1.4426950408889634f);
                const float e1 = 1.f + e;
                const float soft = ls_lg2(e1) * (0.6931471805599453f * inv_beta);
                const float dphi = __fdividef(e, e1);
                av[k] = bz > sp_thr ? z : soft;
                dv[k] = bz > sp_thr ? zd : zd * dphi;
#endif
            }
            float out[16], lo[16];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float r = __shfl_xor_sync(0xffffffffu, isT ? av[k] : dv[k], 16);
                out[k] = isT ? r : av[k];
                out[8 + k] = isT ? dv[k] : r;
            }
            ls_tmem_st(tmem, LS_BT_A + 64 * l + 16 * cg, out, 16);
            if (l < H - 1) {    // (a_H is never a forward operand here; and B5 lets OTHER column groups write these lo columns)
#pragma unroll
                for (int k = 0; k < 16; ++k) lo[k] = ls_tf32_lo(out[k]);
                ls_tmem_st(tmem, LS_BT_LO + 16 * cg, lo, 16);
            }
        }
        if (H == 1 && rad) __syncthreads();     // (deeper nets: the barrier of forward batch 1 already separates the RIN readers from B5)
        // ------------------------------------------------ B5: reverse, output layer.  Row operand: primal [ybar (dout) | pbar (3) | 0],
        //                                                   tangent [s, 0, ...]: the adjoint of n = d(s y0)/dx is the tangent of y0
        {
            float c8[8], lo8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int o = 8 * cg + k;
                float v = 0.f;
                if (valid) {
                    if (isT) v = o == 0 ? a.s : 0.f;
                    else if (o < dout) {
                        if (a.g_y) v += __ldg(a.g_y + i * dout + o);
                        if (o == 0) { if (a.g_sdf) v += a.s * __ldg(a.g_sdf + i); }
                        else if (rad && o - 1 < kg) {
                            const int col = o_geo + o - 1;
                            v += Weff[col] * Pc[0] + Weff[RP + col] * Pc[1] + Weff[2 * RP + col] * Pc[2];
                        }
                    } else if (rad && o < dout + 3) v = Pc[o - dout];
                }
                c8[k] = v;
                lo8[k] = ls_tf32_lo(v);
                if (!isT) blast[k] += v;
                ZR[st_off + o * 4] = v;
                ZL[st_off + o * 4] = lo8[k];
            }
            ls_tmem_st(tmem, LS_BT_LO + 8 * cg, c8, 8);
            ls_tmem_st(tmem, LS_BT_LO + 32 + 8 * cg, lo8, 8);
            float ah[16];
            ls_tmem_ld(tmem, LS_BT_A + 64 * (H - 1) + 16 * cg, ah, 16);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                AR[st_off + (16 * cg + k) * 4] = ah[k];
                AL[st_off + (16 * cg + k) * 4] = ls_tf32_lo(ah[k]);
            }
            ls_fence_smem_to_async();
            ls_tc_sync_before_mma();
            if (warp == 0) {
                if (ls_elect()) {
                    // critical path first: the product into the layer below ...
                    issue_x3(LS_BT_D, LS_BT_LO, LS_BT_LO + 32, LS_H, img.kl_pad);
                    ls_tc_commit(bar);
                    // ... then the output-layer weight gradient, transposed: D[i][o] += sum_r a_H[r][i] ybar'[r][o] (columns
                    // dout..dout+2: G[c][i]); it runs under the next epilogue and is only waited for before the staging arrays change
                    ls_bt_wgrad(tmem, LS_BT_WG, AR, AL, LS_BT_LBO, ZR, ZL, LS_BT_LBO, 32, LS_DBG(a, 2));
                    ls_tc_commit(bar2);
                }
            }
            wg_pending = true;
            if (has_next && has_levels) {       // next tile's gather, piece by piece under B5 and the reverse batches: slot 0 of H + 1
#pragma unroll
                for (int pc = 0; pc < 4; ++pc)
                    if (pc * (H + 1) / 4 == 0) gather_piece(pb_nxt, pc);
            }
            batch_done();
        }
        // ------------------------------------------------ B6: reverse through the hidden layers, k = H .. 1
#pragma unroll
        for (int k = H; k >= 1; --k) {
            const int colA = LS_BT_A + 64 * (k - 1);
            float dvv[16], avv[16];
            ls_tmem_ld2x16(tmem, LS_BT_D + 16 * cg, dvv, colA + 16 * cg, avv);       // both in flight, one wait
            float zb[8], zd[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float r1 = __shfl_xor_sync(0xffffffffu, isT ? dvv[c] : dvv[8 + c], 16);
                const float r2 = __shfl_xor_sync(0xffffffffu, isT ? avv[c] : avv[8 + c], 16);
                const float abar = isT ? r1 : dvv[c];            // adjoint of a_k        (primal row)
                const float abard = isT ? dvv[8 + c] : r1;       // adjoint of adot_k     (tangent row)
                const float ak = isT ? r2 : avv[c];
                const float adk = isT ? avv[8 + c] : r2;
                const float qv = LS_FAST_EXP(-sp_beta * ak);     // 1 - phi'
                const float d1 = 1.f - qv;
                zb[c] = d1 * abar + sp_beta * qv * adk * abard;
                zd[c] = d1 * abard;
            }
            float out[16], lo[16];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float r = __shfl_xor_sync(0xffffffffu, isT ? zb[c] : zd[c], 16);
                out[c] = isT ? r : zb[c];
                out[8 + c] = isT ? zd[c] : r;
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) lo[c] = ls_tf32_lo(out[c]);
            ls_tmem_st(tmem, colA + 16 * cg, out, 16);
            ls_tmem_st(tmem, LS_BT_LO + 16 * cg, lo, 16);
            if (k > 1) {
                // bias gradient of layer k-1 = column sums of zbar_k over the primal rows: recursive halving over the 16 primal lanes
                // of the warp (15 shuffles), lane l ends with column 16 cg + (l & 15) of this warp's rows
                float w8[8], w4[4], w2[2];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float r = __shfl_xor_sync(0xffffffffu, (lane & 8) ? out[c] : out[8 + c], 8);
                    w8[c] = ((lane & 8) ? out[8 + c] : out[c]) + r;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float r = __shfl_xor_sync(0xffffffffu, (lane & 4) ? w8[c] : w8[4 + c], 4);
                    w4[c] = ((lane & 4) ? w8[4 + c] : w8[c]) + r;
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float r = __shfl_xor_sync(0xffffffffu, (lane & 2) ? w4[c] : w4[2 + c], 2);
                    w2[c] = ((lane & 2) ? w4[2 + c] : w4[c]) + r;
                }
                const float r = __shfl_xor_sync(0xffffffffu, (lane & 1) ? w2[0] : w2[1], 1);
                if (!isT) bacc[k - 1] += ((lane & 1) ? w2[1] : w2[0]) + r;
            }
            float ap[16];
            if (k > 1) ls_tmem_ld(tmem, colA - 64 + 16 * cg, ap, 16);
            // the previous weight-gradient batch still reads the staging arrays: everything above ran under it
            ls_tc_wait(bar2, phase2);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                ZR[st_off + (16 * cg + c) * 4] = out[c];
                ZL[st_off + (16 * cg + c) * 4] = lo[c];
            }
            if (k > 1) {        // input of layer k-1: a_{k-1} stack
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    AR[st_off + (16 * cg + c) * 4] = ap[c];
                    AL[st_off + (16 * cg + c) * 4] = ls_tf32_lo(ap[c]);
                }
            } else {            // input of layer 0: the stash is the raw operand, its lo part goes to AL (stash layout)
#pragma unroll
                for (int c = 0; c < LS_BT_EROWS / 4; ++c) {
                    const int f = (LS_BT_EROWS / 4) * cg + c;
                    AL[st_off_e + f * 4] = ls_tf32_lo(ES[st_off_e + f * 4]);
                }
            }
            ls_fence_smem_to_async();
            ls_tc_sync_before_mma();
            if (warp == 0) {
                if (ls_elect()) {
                    if (k > 1) {
                        issue_x3(LS_BT_D, colA, LS_BT_LO, LS_H, LS_H);
                        ls_tc_commit(bar);
                        ls_bt_wgrad(tmem, LS_BT_WG + 64 * (k - 1), ZR, ZL, LS_BT_LBO, AR, AL, LS_BT_LBO, LS_H, LS_DBG(a, 2));
                    } else {
                        issue_x3(LS_BT_D, colA, LS_BT_LO, img.n_in_pad[0], LS_H);
                        ls_tc_commit(bar);
                        ls_bt_wgrad(tmem, LS_BT_WG + 24, ZR, ZL, LS_BT_LBO, ES, AL, LS_BT_LBO_E, LS_BT_EROWS, LS_DBG(a, 2));
                    }
                    ls_tc_commit(bar2);
                }
            }
            if (has_next && has_levels) {       // slot H - k + 1 of H + 1
#pragma unroll
                for (int pc = 0; pc < 4; ++pc)
                    if (pc * (H + 1) / 4 == H - k + 1) gather_piece(pb_nxt, pc);
                if (k == 1) gather_finish();
            }
            batch_done();
        }
        // ------------------------------------------------ B7: encoding adjoints of this thread's two levels -> registers; the scatter
        //                                                   itself runs under the next tile's forward batches
        if (has_levels) {
            float c8[8];
            ls_tmem_ld(tmem, LS_BT_D + 8 * cg, c8, 8);
            if (a.ig.workspace && valid) {      // position gradients wanted: park the adjoints for ls_field_posgrad_kernel
                float* w = a.ig.workspace + i * LS_PG_PITCH + (isT ? 32 : 0) + 8 * cg;
                ls_st4(w, make_float4(c8[0], c8[1], c8[2], c8[3]));
                ls_st4(w + 4, make_float4(c8[4], c8[5], c8[6], c8[7]));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float r = __shfl_xor_sync(0xffffffffu, isT ? c8[k] : c8[4 + k], 16);
                sb[k] = isT ? r : c8[k];            // adjoint of the features of my levels
                sd[k] = isT ? c8[4 + k] : r;        // adjoint of their tangents
            }
        }
        if (a.ig.workspace && cg == 0) {             // ... and the adjoint of the x / rescale columns (primal rows; the TMEM load is
            float cx[8];                             //     warp-collective, so the whole warp executes it)
            ls_tmem_ld(tmem, LS_BT_D + nh, cx, 8);
            if (valid && !isT) ls_st4(a.ig.workspace + i * LS_PG_PITCH + 64, make_float4(cx[0], cx[1], cx[2], 0.f));
        }
        sc_pending = true;
        { const int o = pb_prev; pb_prev = pb_cur; pb_cur = pb_nxt; pb_nxt = o; }
        // (the next tile's first tcgen05.st / staging writes are ordered after this tile's reads by ls_tmem_ld's wait::ld and by
        //  the barrier in front of its first MMA; RIN / PB alias staging arrays whose last MMA reader is waited for at the tile top)
    }
    if (sc_pending && has_levels) {     // the last tile's scatter
        scatter_level(pb_prev, 0, sb[0], sb[1], sd[0], sd[1]);
        scatter_level(pb_prev, 1, sb[2], sb[3], sd[2], sd[3]);
    }

    // ------------------------------------------------ flush the parameter gradients
    if (wg_pending) ls_tc_wait(bar2, phase2);
    ls_tc_sync_before_mma();
    if (a.d_theta && my_tiles > 0) {
        float* GS = smem + net.gs;
        if (q < 2) {        // TMEM lanes 0..63 hold the accumulators: row = output unit (layers < H) / hidden unit (output layer)
            const int j = row;
#pragma unroll 1
            for (int l = 0; l < H; ++l) {
                const int n_in = a.f.dims[l], n_out = a.f.dims[l + 1];
                const int base = l == 0 ? LS_BT_WG + 24 : LS_BT_WG + 64 * l;
                const int n_cols = l == 0 ? img.k_in_pad[0] : LS_H;
                if (16 * cg < n_cols) {
                    float v[16];
                    ls_tmem_ld(tmem, base + 16 * cg, v, 16);
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int col = 16 * cg + c;
                        if (l == 0) {       // kernel column order: [hash features | x | ones | pad]
                            if (col < nh) atomicAdd(a.d_theta + a.net.gw_off[0] + (3 + col) * n_out + j, v[c]);
                            else if (col < nh + 3) atomicAdd(a.d_theta + a.net.gw_off[0] + (col - nh) * n_out + j, v[c]);
                            else if (col == nh + 3) atomicAdd(a.d_theta + a.net.gb_off[0] + j, v[c]);
                        } else if (col < n_in) atomicAdd(a.d_theta + a.net.gw_off[l] + col * n_out + j, v[c]);
                    }
                }
            }
            if (cg < 2) {   // output layer, transposed: lane = hidden unit, column = output (then the 3 G columns)
                float v[16];
                ls_tmem_ld(tmem, LS_BT_WG + 16 * cg, v, 16);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int o = 16 * cg + c;
                    if (o < dout) atomicAdd(a.d_theta + a.net.gw_off[K - 1] + j * dout + o, v[c]);
                    else if (o < dout + 3) GS[(o - dout) * LS_H + j] = v[c];
                }
            }
        }
        // bias gradients.  Hidden layers: primal lane l of every warp holds the column 16 cg + (l & 15) partial of its 16 rows.
#pragma unroll
        for (int l = 1; l < K - 1; ++l)
            if (!isT) atomicAdd(a.d_theta + a.net.gb_off[l] + 16 * cg + (lane & 15), bacc[l]);
        // output layer: 8 columns per thread, summed over the warp's primal lanes first
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = blast[k];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            if (lane == 0 && 8 * cg + k < dout) atomicAdd(a.d_theta + a.net.gb_off[K - 1] + 8 * cg + k, v);
        }
        __syncthreads();
        if (rad && a.d_w_eff && t < 3 * kg) {      // geo block of dL/dW_eff: sum_i W_last[1+k][i] G[c][i] + b_last[1+k] sum_s pbar[c]
            const int c = t / kg, k = t - c * kg;
            const float* G = a.f.theta + a.net.gw_off[K - 1];       // Wt [64][dout]
            float acc = 0.f;
            for (int ii = 0; ii < LS_H; ++ii) acc = fmaf(__ldg(G + ii * dout + 1 + k), GS[c * LS_H + ii], acc);
            atomicAdd(a.d_w_eff + c * a.r.in_dim + o_geo + k, acc);
        }
    }
    if (rad && a.d_w_eff && my_tiles > 0) {
        const int in_dim = a.r.in_dim;
        const int n_items = 3 * nin + 3;
        const int tt = t < 3 * n_items ? t % n_items : -1;         // three sample-range partials per (channel, input) pair
        if (tt >= 0 && tt < 3 * nin) {
            const int c = tt / nin, idx = tt - c * nin;
            atomicAdd(a.d_w_eff + c * in_dim + (idx < o_geo ? idx : idx + kg), weff_acc);
        } else if (tt >= 0 && tt < 3 * nin + 3) {
            const int c = tt - 3 * nin;
            if (a.d_b_eff) atomicAdd(a.d_b_eff + c, weff_acc);
            // the bias part of the geo block: b_last[1+k] * sum_s pbar[c]
            const float* Bl = a.f.theta + a.net.gb_off[K - 1];
            for (int k = 0; k < kg; ++k) atomicAdd(a.d_w_eff + c * in_dim + o_geo + k, __ldg(Bl + 1 + k) * weff_acc);
        }
    }
    ls_tc_dealloc(tmem);
}
