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
//   * weights: W_l (forward) and W_l^T (reverse) as hi/lo K-major operands would need 200 KB; they stream instead from the
//     L2-resident operand image (ls2fm_field_prepare) through a 2-slot ring of 32 KB with cp.async.bulk + mbarrier, one
//     matrix per MMA batch, fetched two batches ahead by the issuing thread.
//   * per tile: H forward batches, 1 + H reverse batches (each reverse batch = weight gradient of one layer + the transposed
//     product into the layer below); no output-layer forward, no separate bias pass (layer 0's bias gradient rides on the
//     ones column of the encoding, the others are column sums of the staged adjoints).
#pragma once

#include "ls2fm_field_tc.cuh"

constexpr int LS_BT_THREADS = 512;
constexpr int LS_BT_TILE = 64;                  // samples per tile
constexpr int64_t LS_BT_MIN_SAMPLES = 32768;    // automatic dispatch: smaller launches run the exact fp32-SIMT kernel
constexpr int LS_BT_KG = 32;                    // groups of 4 rows in a staging array (128 rows)
constexpr int LS_BT_LBO = LS_H * 16 + 16;       // bytes between row groups of a 64-feature staging array (+16: conflict-free stores)
constexpr int LS_BT_LBOF = LS_BT_LBO / 4;       // ... in floats (260)
constexpr int LS_BT_EROWS = 48;                 // feature rows of the encoding stash (35 + ones -> 40, MMA N = 48)
constexpr int LS_BT_LBO_E = LS_BT_EROWS * 16 + 16;
constexpr int LS_BT_LBOF_E = LS_BT_LBO_E / 4;   // 196
constexpr int LS_BT_SLOT = 8192;                // floats per ring slot (32 KB: a 64 x 64 hi/lo pair)
// TMEM columns
constexpr int LS_BT_WG = 0;                     // weight-gradient accumulators: last^T [0,32) | layer 0 [24,72) | layer l [64 l, +64)
constexpr int LS_BT_A = 192;                    // A_k at LS_BT_A + 64 (k - 1)
constexpr int LS_BT_LO = 384;
constexpr int LS_BT_D = 448;

struct LsBtNet {         // shared-memory plan (float offsets)
    int zr, zl, ar;      // staging: adjoint raw / lo, layer input (tf32-rounded)    [32 groups][64 features][4 rows] padded
    int es;              // encoding stash (raw), operand of the layer-0 weight gradient [32 groups][48 features][4 rows] padded
    int ring;            // 2 slots
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
    n.es = off; off += LS_BT_KG * LS_BT_LBOF_E;
    n.ring = off; off += 2 * LS_BT_SLOT;         // also absorbs the M = 128 over-read of the arrays above
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

// weight-gradient batch: D[m][d_col + n] += sum_r A(m, r) B(n, r) over the tile's 128 rows.  One operand ("split": the adjoint) is
// exact as raw + lo (raw read as tf32 by truncation); the other ("single": the layer input) was rounded to tf32 (rna) when it was
// staged -- an unbiased 2^-12 relative perturbation of each term of a sum over >= thousands of rows, far below the fp32 atomics'
// own reordering noise.  split_is_a: the split operand is A (M = its features), else B.
LS_DEV void ls_bt_wgrad(uint32_t tmem, int d_col, const float* p_raw, const float* p_lo, int p_lbo, const float* s_hi, int s_lbo, int N,
                        bool split_is_a) {
#if defined(LS_HOSTSIM)
    for (int ks = 0; ks < LS_BT_KG / 2; ++ks) {      // 8 rows (two groups of 4) per MMA
        const int po = ks * 2 * (p_lbo / 4), so = ks * 2 * (s_lbo / 4);
        if (split_is_a) {
            ls_tc_mma_ss(tmem, d_col, p_lo + po, p_lbo, s_hi + so, s_lbo, N, true);
            ls_tc_mma_ss(tmem, d_col, p_raw + po, p_lbo, s_hi + so, s_lbo, N, true);
        } else {
            ls_tc_mma_ss(tmem, d_col, s_hi + so, s_lbo, p_lo + po, p_lbo, N, true);
            ls_tc_mma_ss(tmem, d_col, s_hi + so, s_lbo, p_raw + po, p_lbo, N, true);
        }
    }
#else
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(LS_TC_M >> 4) << 24);
    uint64_t pr = ls_tc_desc(ls_smem_u32(p_raw), (uint32_t)p_lbo, 128), pl = ls_tc_desc(ls_smem_u32(p_lo), (uint32_t)p_lbo, 128);
    uint64_t sh = ls_tc_desc(ls_smem_u32(s_hi), (uint32_t)s_lbo, 128);
    const uint64_t ps = (uint64_t)(2 * p_lbo >> 4), ss = (uint64_t)(2 * s_lbo >> 4);     // 8 rows (two groups of 4) per MMA
    const uint32_t d = tmem + (uint32_t)d_col;
#define LS_BT_SS(A, B) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" \
                                    :: "r"(d), "l"(A), "l"(B), "r"(idesc))
#pragma unroll
    for (int ks = 0; ks < LS_BT_KG / 2; ++ks) {
        if (split_is_a) { LS_BT_SS(pl, sh); LS_BT_SS(pr, sh); }
        else { LS_BT_SS(sh, pl); LS_BT_SS(sh, pr); }
        pr += ps; pl += ps; sh += ss;
    }
#undef LS_BT_SS
#endif
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
    float* ZR = smem + net.zr; float* ZL = smem + net.zl; float* AR = smem + net.ar;
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
    LsTcBar* full0 = reinterpret_cast<LsTcBar*>(smem + net.misc + 2);
    LsTcBar* full1 = reinterpret_cast<LsTcBar*>(smem + net.misc + 4);
    LsTcBar* bar2 = reinterpret_cast<LsTcBar*>(smem + net.misc + 6);     // weight-gradient batches (off the critical path)
    const uint32_t tmem = ls_tc_alloc(reinterpret_cast<uint32_t*>(smem + net.misc + 8));
    ls_tc_bar_init(bar);
    ls_tc_bar_init(bar2);
    if (t == 0) { ls_bar_init1(full0); ls_bar_init1(full1); }
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
    int64_t gb = 0;                     // running batch index (ring position)
    auto ring_fetch = [&](int64_t g) {  // thread 0: start the copy of batch g's matrix into slot g & 1
        if (g < n_batches) {
            int src, floats;
            ls_bt_matrix(img, H, (int)(g % NB), &src, &floats);
            ls_bulk_g2s(smem + net.ring + (int)(g & 1) * LS_BT_SLOT, a.f.tc_image + src, floats * 4, (g & 1) ? full1 : full0);
        }
    };
    if (t == 0) { ring_fetch(0); ring_fetch(1); }
    // warp 0, after the barrier of batch g: wait for its matrix; returns the slot
    auto ring_slot = [&](int64_t g) -> const float* {
        ls_bar_wait((g & 1) ? full1 : full0, (uint32_t)((g >> 1) & 1));
        return smem + net.ring + (int)(g & 1) * LS_BT_SLOT;
    };
    // all threads: wait for batch g, then thread 0 refills its slot with the matrix of batch g + 2
    auto batch_done = [&]() {
        ls_tc_wait(bar, phase);
        if (warp == 0) { if (ls_elect()) ring_fetch(gb + 2); }
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

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ------------------------------------------------ B1: upstream gradients of this thread's sample
        if (wg_pending) { ls_tc_wait(bar2, phase2); wg_pending = false; }    // the stash / staging arrays are rewritten below
        const int64_t i_in = tile * LS_BT_TILE + sl;
        const bool valid = i_in < a.p.n;
        float x[3] = {0.f, 0.f, 0.f}, u[3];
        int ray_id = 0;
        int64_t i = i_in;
        if (valid) ls_sample_point(a.p, i_in, x, &ray_id, &i);
        ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u);
        float pbar[3] = {0.f, 0.f, 0.f}, nbar[3] = {0.f, 0.f, 0.f};
        if (valid) {
            if (rad && a.g_rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float rgb = __ldg(a.saved_rgb + 3 * i + c);
                    pbar[c] = __ldg(a.g_rgb + 3 * i + c) * rgb * (1.f - rgb);
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float v = a.g_nrm ? __ldg(a.g_nrm + 3 * i + d) : 0.f;
                if (rad) v += Weff[3 + d] * pbar[0] + Weff[RP + 3 + d] * pbar[1] + Weff[2 * RP + 3 + d] * pbar[2];
                nbar[d] = v;
            }
        }
        if (rad) {
            // radiance inputs of the W_eff gradient, spread over the 8 threads of the sample
            float* in = RIN + sl * nin;
            float dir[3] = {0.f, 0.f, 0.f};
            if (valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) dir[d] = __ldg(a.p.ray + 3 * ray_id + d);
            }
            if (!isT) {
                if (cg == 0) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { in[d] = x[d]; PB[4 * sl + d] = pbar[d]; }
                } else if (cg == 1) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) in[3 + d] = valid ? __ldg(a.saved_nrm + 3 * i + d) : 0.f;
                } else if (cg == 2) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) in[6 + d] = dir[d];
                } else {
                    for (int k = 0; k < kg2; ++k) in[o_geo + k] = valid ? __ldg(a.r.geo2 + i * (kg2 + 1) + 1 + k) : 0.f;
                }
                if (a.d_geo2 && valid) {
                    for (int k = cg; k <= kg2; k += 4) {
                        float v = 0.f;
                        if (k >= 1) {
                            const int col = o_geo2 + k - 1;
                            v = Weff[col] * pbar[0] + Weff[RP + col] * pbar[1] + Weff[2 * RP + col] * pbar[2];
                        }
                        a.d_geo2[i * (kg2 + 1) + k] = v;
                    }
                }
            } else {
                for (int k = cg; k < nf; k += 4) {
                    const float fr = (float)(1 << k);
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float ang = dir[d] * fr;
                        in[9 + 6 * k + d] = sinf(ang);
                        in[12 + 6 * k + d] = cosf(ang);
                    }
                }
            }
        }
        // ------------------------------------------------ B2: gather.  This thread: levels 4 cg + 2 isT + {0, 1}; features e and
        //                                                   tangent features edot = Je nbar; the two rows of the sample swap halves
        if (4 * cg < L) {
            float e4[4], d4[4];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int l = 4 * cg + 2 * isT + r;
                float h[2], dh[2][3];
                ls_level_eval(a.f, l, u, h, dh);
#pragma unroll
                for (int fi = 0; fi < 2; ++fi) {
                    e4[2 * r + fi] = h[fi];
                    d4[2 * r + fi] = dh[fi][0] * a.inv_ext[0] * nbar[0] + dh[fi][1] * a.inv_ext[1] * nbar[1] + dh[fi][2] * a.inv_ext[2] * nbar[2];
                }
            }
            float c8[8], lo8[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float r = __shfl_xor_sync(0xffffffffu, isT ? e4[k] : d4[k], 16);
                c8[k] = isT ? r : e4[k];            // primal row: features of levels 4cg, 4cg+1 (own), 4cg+2, 4cg+3 (partner)
                c8[4 + k] = isT ? d4[k] : r;        // tangent row: edot likewise
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                lo8[k] = ls_tf32_lo(c8[k]);
                ES[st_off_e + (8 * cg + k) * 4] = ls_tf32_rna(c8[k]);
            }
            ls_tmem_st(tmem, colE + 8 * cg, c8, 8);
            ls_tmem_st(tmem, LS_BT_LO + 8 * cg, lo8, 8);
        }
        if (cg == 0) {      // tail chunk: x / rescale (tangent row: nbar / rescale), the ones column (primal rows only), zero padding
            float c8[8] = {0.f, 0.f, 0.f, isT ? 0.f : 1.f, 0.f, 0.f, 0.f, 0.f}, lo8[8];
#pragma unroll
            for (int d = 0; d < 3; ++d) c8[d] = isT ? nbar[d] / a.f.rescale : ls_fdiv(x[d], a.f.rescale);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                lo8[k] = ls_tf32_lo(c8[k]);
                ES[st_off_e + (nh + k) * 4] = ls_tf32_rna(c8[k]);
            }
            ls_tmem_st(tmem, colE + nh, c8, 8);
            ls_tmem_st(tmem, LS_BT_LO + nh, lo8, 8);
        }
        // ------------------------------------------------ B3: W_eff gradient, non-geo inputs (x, n, dir, Fourier, geo2) and b_eff
        if (rad) {
            __syncthreads();
            if (t < 3 * nin + 3) {
                const int c = t < 3 * nin ? t / nin : t - 3 * nin;
                const int idx = t < 3 * nin ? t - c * nin : -1;
                float acc = 0.f;
#pragma unroll 4
                for (int s = 0; s < LS_BT_TILE; ++s) acc = fmaf(PB[4 * s + c], idx >= 0 ? RIN[s * nin + idx] : 1.f, acc);
                weff_acc += acc;
            }
        }
        // ------------------------------------------------ B4: forward, both channels per batch
#pragma unroll
        for (int l = 0; l < H; ++l) {
            const int a_col = l == 0 ? colE : LS_BT_A + 64 * (l - 1);
            ls_tc_sync_before_mma();
            if (warp == 0) {
                const float* W = ring_slot(gb);
                const int Kp = img.k_in_pad[l];
                if (ls_elect()) {
                    ls_tc_mma_x3(tmem, LS_BT_D, a_col, LS_BT_LO, W, W + LS_H * Kp, LS_H, Kp);
                    ls_tc_commit(bar);
                }
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
                if (bz > sp_thr) { av[k] = z; dv[k] = zd; }
                else {
#if defined(LS_HOSTSIM)
                    const float e = expf(bz);
                    av[k] = log1pf(e) * inv_beta;
                    dv[k] = zd * (e / (e + 1.f));
#else
                    const float e = __expf(bz);
                    av[k] = __logf(1.f + e) * inv_beta;
                    dv[k] = zd * __fdividef(e, e + 1.f);
#endif
                }
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
                            v += Weff[col] * pbar[0] + Weff[RP + col] * pbar[1] + Weff[2 * RP + col] * pbar[2];
                        }
                    } else if (rad && o < dout + 3) v = pbar[o - dout];
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
                AR[st_off + (16 * cg + k) * 4] = ls_tf32_rna(ah[k]);
            }
            ls_fence_smem_to_async();
            ls_tc_sync_before_mma();
            if (warp == 0) {
                const float* W = ring_slot(gb);
                if (ls_elect()) {
                    // critical path first: the product into the layer below ...
                    ls_tc_mma_x3(tmem, LS_BT_D, LS_BT_LO, LS_BT_LO + 32, W, W + LS_H * img.kl_pad, LS_H, img.kl_pad);
                    ls_tc_commit(bar);
                    // ... then the output-layer weight gradient, transposed: D[i][o] += sum_r a_H[r][i] ybar'[r][o] (columns
                    // dout..dout+2: G[c][i]); it runs under the next epilogue and is only waited for before the staging arrays change
                    ls_bt_wgrad(tmem, LS_BT_WG, ZR, ZL, LS_BT_LBO, AR, LS_BT_LBO, 32, false);
                    ls_tc_commit(bar2);
                }
            }
            wg_pending = true;
            batch_done();
        }
        // ------------------------------------------------ B6: reverse through the hidden layers, k = H .. 1
#pragma unroll
        for (int k = H; k >= 1; --k) {
            const int colA = LS_BT_A + 64 * (k - 1);
            float dvv[16], avv[16];
            ls_tmem_ld(tmem, LS_BT_D + 16 * cg, dvv, 16);
            ls_tmem_ld(tmem, colA + 16 * cg, avv, 16);
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
                    AR[st_off + (16 * cg + c) * 4] = ls_tf32_rna(ap[c]);
                }
            }
            ls_fence_smem_to_async();
            ls_tc_sync_before_mma();
            if (warp == 0) {
                const float* W = ring_slot(gb);
                if (ls_elect()) {
                    if (k > 1) {
                        ls_tc_mma_x3(tmem, LS_BT_D, colA, LS_BT_LO, W, W + LS_H * LS_H, LS_H, LS_H);
                        ls_tc_commit(bar);
                        ls_bt_wgrad(tmem, LS_BT_WG + 64 * (k - 1), ZR, ZL, LS_BT_LBO, AR, LS_BT_LBO, LS_H, true);
                    } else {
                        const int N0 = img.n_in_pad[0];
                        ls_tc_mma_x3(tmem, LS_BT_D, colA, LS_BT_LO, W, W + N0 * LS_H, N0, LS_H);
                        ls_tc_commit(bar);
                        ls_bt_wgrad(tmem, LS_BT_WG + 24, ZR, ZL, LS_BT_LBO, ES, LS_BT_LBO_E, LS_BT_EROWS, true);
                    }
                    ls_tc_commit(bar2);
                }
            }
            batch_done();
        }
        // ------------------------------------------------ B7: hash-table gradient scatter (this thread's two levels)
        if (4 * cg < L) {
            float c8[8];
            ls_tmem_ld(tmem, LS_BT_D + 8 * cg, c8, 8);
            float eb[4], ed[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float r = __shfl_xor_sync(0xffffffffu, isT ? c8[k] : c8[4 + k], 16);
                eb[k] = isT ? r : c8[k];            // adjoint of the features of my levels
                ed[k] = isT ? c8[4 + k] : r;        // adjoint of their tangents
            }
            if (a.d_table && valid) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int l = 4 * cg + 2 * isT + r;
                    const float scale = a.f.levels[l].scale;
                    const uint32_t res = a.f.levels[l].resolution, size = a.f.levels[l].size, hashed = a.f.levels[l].hashed;
                    float* tab = a.d_table + 2 * (size_t)a.f.levels[l].offset;
                    const LsCell c = ls_cell(scale, u);
                    const float e0 = eb[2 * r], e1 = eb[2 * r + 1], t0 = ed[2 * r], t1 = ed[2 * r + 1];
                    float ns[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) ns[d] = nbar[d] * scale * a.inv_ext[d];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float f0 = (k & 1) ? c.w[0] : 1.f - c.w[0];
                        const float f1 = (k & 2) ? c.w[1] : 1.f - c.w[1];
                        const float f2 = (k & 4) ? c.w[2] : 1.f - c.w[2];
                        const float wgt = f0 * f1 * f2;
                        const float dw = ((k & 1) ? ns[0] : -ns[0]) * f1 * f2 + ((k & 2) ? ns[1] : -ns[1]) * f0 * f2 +
                                         ((k & 4) ? ns[2] : -ns[2]) * f0 * f1;
                        const uint32_t idx = ls_corner_index(res, size, hashed, c, k);
                        atomicAdd(reinterpret_cast<float2*>(tab) + idx, make_float2(wgt * e0 + dw * t0, wgt * e1 + dw * t1));
                    }
                }
            }
        }
        // (the next tile's first tcgen05.st / staging writes are ordered after this tile's reads by ls_tmem_ld's wait::ld and by
        //  the barrier in front of its first MMA; RIN / PB alias staging arrays whose last readers finished before batch_done)
        if (rad) __syncthreads();
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
        if (t < 3 * nin) {
            const int c = t / nin, idx = t - c * nin;
            atomicAdd(a.d_w_eff + c * in_dim + (idx < o_geo ? idx : idx + kg), weff_acc);
        } else if (t < 3 * nin + 3) {
            const int c = t - 3 * nin;
            if (a.d_b_eff) atomicAdd(a.d_b_eff + c, weff_acc);
            // the bias part of the geo block: b_last[1+k] * sum_s pbar[c]
            const float* Bl = a.f.theta + a.net.gb_off[K - 1];
            for (int k = 0; k < kg; ++k) atomicAdd(a.d_w_eff + c * in_dim + o_geo + k, __ldg(Bl + 1 + k) * weff_acc);
        }
    }
    ls_tc_dealloc(tmem);
}
