// ls2fm_field_bwfeat.cuh -- field BACKWARD kernel on the tensor cores for launches whose normals carry NO gradient
// (RadF.Geometry_feat under dual_field: models/RadF.py:53-63 -- the second hash field only feeds features to the radiance
//  decoder; SDF.infer_sdf losses of BA, pipelines/BA.py:123-125, when they are large enough to leave the fp32-SIMT kernel).
//
// Without the tangent channel the adjoint network is the plain reverse of  hash grid -> MLP:  zbar_l = abar_l * phi'(z_l),
// abar_{l-1} = W_l^T zbar_l, dW_l += zbar_l (x) a_{l-1}, table scatter of ebar.  ls_field_backward_tc_kernel runs that case with
// a zero tangent channel, i.e. with half of every M = 128 instruction wasted; here the tile is 128 SAMPLES (TMEM lane = sample),
// every thread owns the 16 columns of its column group for its own row, and nothing is exchanged between lanes in the epilogues.
// Everything else is the design of ls2fm_field_bwtc.cuh, unchanged: raw fp32 activations in TMEM + a transient lo buffer (3xTF32),
// weights streamed from the operand image through the 3-slot ring, all weight-gradient accumulators resident in TMEM (both
// operands staged in shared memory with K = row), the cross-tile software pipeline (next tile's gather under the reverse
// batches, previous tile's scatter under the forward batches), the same shared-memory plan (LsBtNet; the per-sample staging
// rows are shorter, so three tiles of 128 fit where three of 64 did) and the same TMEM map.
#pragma once

#include "ls2fm_field_bwtc.cuh"

constexpr int LS_BF_TILE = 128;     // samples per tile
constexpr int LS_BF_PBP = 9;        // per-sample staging row: 0..2 position | 3 valid | 4..6 unit-cube coordinates (odd pitch)
static_assert(3 * LS_BF_TILE * LS_BF_PBP <= 3 * LS_BT_TILE * LS_BT_PBP, "per-sample staging of the feature-only kernel must fit LsBtNet.pbn");

template <int K>
__global__ void __launch_bounds__(LS_BT_THREADS, 1) ls_field_backward_feat_tc_kernel(const LsFieldArgs a, const LsTcNet img, const LsBtNet net) {
    LS_DYN_SMEM(smem);
    constexpr int H = K - 1;            // hidden layers
    constexpr int NB = 2 * H + 1;       // MMA batches (= ring matrices) per tile
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int q = warp & 3, cg = warp >> 2;
    const int row = 32 * q + lane;                  // TMEM lane = sample inside the tile = row of the staging arrays
    const int L = a.f.n_levels, nh = 2 * L;
    const int dout = a.f.dims[K];
    const float sp_beta = a.f.softplus_beta, sp_thr = a.f.softplus_threshold, inv_beta = 1.f / a.f.softplus_beta;
    float* ZR = smem + net.zr; float* ZL = smem + net.zl; float* AR = smem + net.ar; float* AL = smem + net.al;
    float* ES = smem + net.es;

    // ------------------------------------------------ one-time setup
    for (int l = 1; l < K - 1; ++l)
        for (int e = t; e < LS_H; e += LS_BT_THREADS) smem[net.bias[l] + e] = __ldg(a.f.tc_image + img.bias[l] + e);
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
    bool wg_pending = false;
    {   // zero the weight-gradient accumulators
        float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 48; c += 8) ls_tmem_st(tmem, LS_BT_WG + 48 * cg + c, z8, 8);
    }

    const int64_t n_tiles = (a.p.n + LS_BF_TILE - 1) / LS_BF_TILE;
    const int64_t my_tiles = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_batches = my_tiles * NB;
    int64_t gb = 0;                     // running batch index
    auto ring_fetch = [&](int64_t e) {  // one thread; ring entries as in ls_field_backward_tc_kernel
        if (e < 2 * n_batches) {
            int src, floats;
            ls_bt_matrix(img, H, (int)((e >> 1) % NB), &src, &floats);
            const int half = floats / 2, s3 = (int)(e % 3);
            ls_bulk_g2s(smem + net.ring + s3 * LS_BT_SLOT, a.f.tc_image + src + (int)(e & 1) * half, half * 4,
                        reinterpret_cast<LsTcBar*>(smem + net.misc + 2 + 2 * s3));
        }
    };
    if (t == 0) { ring_fetch(0); ring_fetch(1); ring_fetch(2); }
    auto ring_slot = [&](int64_t e) -> const float* {
        const int s3 = (int)(e % 3);
        ls_bar_wait(reinterpret_cast<LsTcBar*>(smem + net.misc + 2 + 2 * s3), (uint32_t)((e / 3) & 1));
        return smem + net.ring + s3 * LS_BT_SLOT;
    };
    auto issue_x3 = [&](int d_col, int a_raw_col, int a_lo_col, int N, int Kd) {
        const float* Wh = ring_slot(2 * gb);
        ls_tc_mma(tmem, d_col, a_lo_col, Wh, N, Kd, false);
        ls_tc_mma(tmem, d_col, a_raw_col, Wh, N, Kd, true);
        const float* Wl = ring_slot(2 * gb + 1);
        ls_tc_mma(tmem, d_col, a_raw_col, Wl, N, Kd, true);
    };
    auto batch_done = [&]() {
        ls_tc_wait(bar, phase);
        if (warp == 0) { if (ls_elect()) { ring_fetch(2 * gb + 3); ring_fetch(2 * gb + 4); } }
        ++gb;
    };
    // persistent accumulators
    float bacc[LS2FM_MAX_LAYERS];       // hidden-layer bias gradients: even lanes, column 16 cg + (lane >> 1)
#pragma unroll
    for (int l = 0; l < LS2FM_MAX_LAYERS; ++l) bacc[l] = 0.f;
    float blast[8];                     // output-layer bias gradient: this thread's 8 columns of ybar, summed over its tiles
#pragma unroll
    for (int k = 0; k < 8; ++k) blast[k] = 0.f;
    const int st_off = (row >> 2) * LS_BT_LBOF + (row & 3);        // + feature * 4
    const int st_off_e = (row >> 2) * LS_BT_LBOF_E + (row & 3);
    const int colE = LS_BT_A + 64 * (H - 1);

    // ------------------------------------------------ software pipeline across tiles (as in ls_field_backward_tc_kernel)
    float* PBN = smem + net.pbn;
    int pb_prev = 2, pb_cur = 0, pb_nxt = 1;
    auto stage_sample = [&](int64_t tile, int buf) {        // one thread per sample
        if (cg != 0) return;
        float* P = PBN + (buf * LS_BF_TILE + row) * LS_BF_PBP;
        const int64_t i = tile * LS_BF_TILE + row;          // (dense launches: output index == input index)
        const bool valid = tile < n_tiles && i < a.p.n;
        float x[3] = {0.f, 0.f, 0.f}, u3[3];
        if (valid) { int ray_id; int64_t io; ls_sample_point(a.p, i, x, &ray_id, &io); }
        ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u3);
        P[0] = x[0]; P[1] = x[1]; P[2] = x[2];
        P[3] = valid ? 1.f : 0.f;
        P[4] = u3[0]; P[5] = u3[1]; P[6] = u3[2];
    };
    // The gather of a tile: this thread's four levels 4 cg + r, each cut into its two z-planes (4 corner loads): 8 pieces spread over the
    // H + 1 reverse batches of the previous tile.  Same arithmetic, term by term, as ls_level_eval.
    float c8n[8];                       // gathered chunk of the next tile: columns 8 cg .. 8 cg + 7 of its row
    float pl[2];                        // z = 0 plane value of the level being evaluated, per feature
    auto gather_piece = [&](int buf, int p) {
        const int r = p >> 1, plane = p & 1;
        const float* P = PBN + (buf * LS_BF_TILE + row) * LS_BF_PBP;
        const float u[3] = {P[4], P[5], P[6]};
        const int l = 4 * cg + r;
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
#pragma unroll
        for (int fi = 0; fi < 2; ++fi) {
            const float q0 = fi ? v[0].y : v[0].x, q1 = fi ? v[1].y : v[1].x, q2 = fi ? v[2].y : v[2].x, q3 = fi ? v[3].y : v[3].x;
            const float a0 = m0 * q0 + w0 * q1, a1 = m0 * q2 + w0 * q3;
            const float b = m1 * a0 + w1 * a1;
            if (!plane) pl[fi] = b;
            else c8n[2 * r + fi] = m2 * pl[fi] + w2 * b;
        }
    };
    float sb[8];                        // deferred scatter (previous tile): adjoints of my four levels' features
    bool sc_pending = false;
    auto scatter_level = [&](int buf, int r, float e0, float e1) {
        const float* P = PBN + (buf * LS_BF_TILE + row) * LS_BF_PBP;
        if (!a.d_table || P[3] == 0.f || LS_DBG(a, 0)) return;
        const float u[3] = {P[4], P[5], P[6]};
        const int l = 4 * cg + r;
        const uint32_t res = a.f.levels[l].resolution, size = a.f.levels[l].size, hashed = a.f.levels[l].hashed;
        float* tab = a.d_table + 2 * (size_t)a.f.levels[l].offset;
        const LsCell c = ls_cell(a.f.levels[l].scale, u);
        uint32_t ci[8];
        ls_corner_indices<0, 8>(res, size, hashed, c, ci);
        const bool v4 = (reinterpret_cast<uintptr_t>(a.d_table) & 15) == 0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) {        // x-neighbour pairs: one 16-byte atomic when the two entries are an aligned pair
            const float f1 = (k & 2) ? c.w[1] : 1.f - c.w[1];
            const float f2 = (k & 4) ? c.w[2] : 1.f - c.w[2];
            const float wa = (1.f - c.w[0]) * f1 * f2, wb = c.w[0] * f1 * f2;
            ls_red_pair(tab, ci[k], ci[k + 1], wa * e0, wa * e1, wb * e0, wb * e1, v4);
        }
    };
    const bool has_levels = 4 * cg < L;
    // prologue: the first tile's loads and gather, unhidden
    if (my_tiles > 0) {
        stage_sample(blockIdx.x, pb_cur);
        __syncthreads();
        if (has_levels) {
#pragma unroll
            for (int pc = 0; pc < 8; ++pc) gather_piece(pb_cur, pc);
        }
    }

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const bool has_next = tile + gridDim.x < n_tiles;
        // ------------------------------------------------ B1: this tile's encoding (gathered under the previous tile) -> TMEM + stash
        if (wg_pending) { ls_tc_wait(bar2, phase2); wg_pending = false; }    // the stash / staging arrays are rewritten below
        const float* Pc = PBN + (pb_cur * LS_BF_TILE + row) * LS_BF_PBP;
        const int64_t i = tile * LS_BF_TILE + row;
        const bool valid = Pc[3] != 0.f;
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
        if (cg == 0) {      // tail chunk: x / rescale, the ones column (bias of layer 0), zero padding
            float c8[8] = {ls_fdiv(Pc[0], a.f.rescale), ls_fdiv(Pc[1], a.f.rescale), ls_fdiv(Pc[2], a.f.rescale), 1.f, 0.f, 0.f, 0.f, 0.f}, lo8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                lo8[k] = ls_tf32_lo(c8[k]);
                ES[st_off_e + (nh + k) * 4] = c8[k];
            }
            ls_tmem_st(tmem, colE + nh, c8, 8);
            ls_tmem_st(tmem, LS_BT_LO + nh, lo8, 8);
        }
        // ------------------------------------------------ B4: forward
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
            if (l == H - 1 && has_next) stage_sample(tile + gridDim.x, pb_nxt);     // next tile's per-sample loads run under this batch
            if (sc_pending && has_levels) {     // the previous tile's table-gradient scatter runs under the forward batches
                if (l == 0) { scatter_level(pb_prev, 0, sb[0], sb[1]); scatter_level(pb_prev, 1, sb[2], sb[3]); }
                if (l == (H > 1 ? 1 : 0)) { scatter_level(pb_prev, 2, sb[4], sb[5]); scatter_level(pb_prev, 3, sb[6], sb[7]); }
            }
            batch_done();
            float xv[16], out[16];
            ls_tmem_ld(tmem, LS_BT_D + 16 * cg, xv, 16);
            const float* bias = smem + net.bias[l] + 16 * cg;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float z = xv[k];
                if (l > 0) z += bias[k];                         // (layer 0: the bias rides on the ones column)
                const float bz = z * sp_beta;
#if defined(LS_HOSTSIM)
                out[k] = bz > sp_thr ? z : log1pf(expf(bz)) * inv_beta;
#else
                const float e = ls_ex2(fminf(bz, sp_thr) * 1.4426950408889634f);
                const float soft = ls_lg2(1.f + e) * (0.6931471805599453f * inv_beta);
                out[k] = bz > sp_thr ? z : soft;
#endif
            }
            ls_tmem_st(tmem, LS_BT_A + 64 * l + 16 * cg, out, 16);
            if (l < H - 1) {    // (a_H is never a forward operand here)
                float lo[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) lo[k] = ls_tf32_lo(out[k]);
                ls_tmem_st(tmem, LS_BT_LO + 16 * cg, lo, 16);
            }
        }
        // ------------------------------------------------ B5: reverse, output layer.  Row operand: [ybar (dout) | 0]
        {
            float c8[8], lo8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int o = 8 * cg + k;
                float v = 0.f;
                if (valid && o < dout) {
                    if (a.g_y) v += __ldg(a.g_y + i * dout + o);
                    if (o == 0 && a.g_sdf) v += a.s * __ldg(a.g_sdf + i);
                }
                c8[k] = v;
                lo8[k] = ls_tf32_lo(v);
                blast[k] += v;
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
                    issue_x3(LS_BT_D, LS_BT_LO, LS_BT_LO + 32, LS_H, img.kl_pad);       // critical path first
                    ls_tc_commit(bar);
                    // output-layer weight gradient, transposed: D[i][o] += sum_r a_H[r][i] ybar[r][o]
                    ls_bt_wgrad(tmem, LS_BT_WG, AR, AL, LS_BT_LBO, ZR, ZL, LS_BT_LBO, 32, LS_DBG(a, 2));
                    ls_tc_commit(bar2);
                }
            }
            wg_pending = true;
            if (has_next && has_levels) {       // next tile's gather: slot 0 of H + 1
#pragma unroll
                for (int pc = 0; pc < 8; ++pc)
                    if (pc * (H + 1) / 8 == 0) gather_piece(pb_nxt, pc);
            }
            batch_done();
        }
        // ------------------------------------------------ B6: reverse through the hidden layers, k = H .. 1
#pragma unroll
        for (int k = H; k >= 1; --k) {
            const int colA = LS_BT_A + 64 * (k - 1);
            float dvv[16], avv[16];
            ls_tmem_ld2x16(tmem, LS_BT_D + 16 * cg, dvv, colA + 16 * cg, avv);       // both in flight, one wait
            float out[16], lo[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float qv = LS_FAST_EXP(-sp_beta * avv[c]);     // 1 - phi'
                out[c] = (1.f - qv) * dvv[c];
                lo[c] = ls_tf32_lo(out[c]);
            }
            ls_tmem_st(tmem, colA + 16 * cg, out, 16);
            ls_tmem_st(tmem, LS_BT_LO + 16 * cg, lo, 16);
            if (k > 1) {
                // bias gradient of layer k-1 = column sums of zbar_k over the warp's 32 rows: recursive halving, lanes 2c and 2c+1 end
                // with column 16 cg + c
                float w8[8], w4[4], w2[2];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float r = __shfl_xor_sync(0xffffffffu, (lane & 16) ? out[c] : out[8 + c], 16);
                    w8[c] = ((lane & 16) ? out[8 + c] : out[c]) + r;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float r = __shfl_xor_sync(0xffffffffu, (lane & 8) ? w8[c] : w8[4 + c], 8);
                    w4[c] = ((lane & 8) ? w8[4 + c] : w8[c]) + r;
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float r = __shfl_xor_sync(0xffffffffu, (lane & 4) ? w4[c] : w4[2 + c], 4);
                    w2[c] = ((lane & 4) ? w4[2 + c] : w4[c]) + r;
                }
                const float r = __shfl_xor_sync(0xffffffffu, (lane & 2) ? w2[0] : w2[1], 2);
                float w1 = ((lane & 2) ? w2[1] : w2[0]) + r;
                w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
                if (!(lane & 1)) bacc[k - 1] += w1;
            }
            float ap[16];
            if (k > 1) ls_tmem_ld(tmem, colA - 64 + 16 * cg, ap, 16);
            ls_tc_wait(bar2, phase2);           // the previous weight-gradient batch still reads the staging arrays
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                ZR[st_off + (16 * cg + c) * 4] = out[c];
                ZL[st_off + (16 * cg + c) * 4] = lo[c];
            }
            if (k > 1) {        // input of layer k-1: a_{k-1}
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
                for (int pc = 0; pc < 8; ++pc)
                    if (pc * (H + 1) / 8 == H - k + 1) gather_piece(pb_nxt, pc);
            }
            batch_done();
        }
        // ------------------------------------------------ B7: encoding adjoints of this thread's four levels -> registers; the scatter
        //                                                   itself runs under the next tile's forward batches
        if (has_levels) {
            ls_tmem_ld(tmem, LS_BT_D + 8 * cg, sb, 8);
            if (a.ig.workspace && valid) {      // position gradients wanted: park the adjoints for ls_field_posgrad_kernel (no tangent part)
                float* w = a.ig.workspace + i * LS_PG_PITCH + 8 * cg;
                ls_st4(w, make_float4(sb[0], sb[1], sb[2], sb[3]));
                ls_st4(w + 4, make_float4(sb[4], sb[5], sb[6], sb[7]));
                ls_st4(w + 32, make_float4(0.f, 0.f, 0.f, 0.f));
                ls_st4(w + 36, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
        if (a.ig.workspace && cg == 0) {             // ... and the adjoint of the x / rescale columns
            float cx[8];
            ls_tmem_ld(tmem, LS_BT_D + nh, cx, 8);
            if (valid) ls_st4(a.ig.workspace + i * LS_PG_PITCH + 64, make_float4(cx[0], cx[1], cx[2], 0.f));
        }
        sc_pending = true;
        { const int o = pb_prev; pb_prev = pb_cur; pb_cur = pb_nxt; pb_nxt = o; }
    }
    if (sc_pending && has_levels) {     // the last tile's scatter
        scatter_level(pb_prev, 0, sb[0], sb[1]); scatter_level(pb_prev, 1, sb[2], sb[3]);
        scatter_level(pb_prev, 2, sb[4], sb[5]); scatter_level(pb_prev, 3, sb[6], sb[7]);
    }

    // ------------------------------------------------ flush the parameter gradients
    if (wg_pending) ls_tc_wait(bar2, phase2);
    ls_tc_sync_before_mma();
    if (a.d_theta && my_tiles > 0) {
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
            if (cg < 2) {   // output layer, transposed: lane = hidden unit, column = output
                float v[16];
                ls_tmem_ld(tmem, LS_BT_WG + 16 * cg, v, 16);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int o = 16 * cg + c;
                    if (o < dout) atomicAdd(a.d_theta + a.net.gw_off[K - 1] + j * dout + o, v[c]);
                }
            }
        }
        // bias gradients.  Hidden layers: even lane l of every warp holds the column 16 cg + (l >> 1) partial of its 32 rows.
#pragma unroll
        for (int l = 1; l < K - 1; ++l)
            if (!(lane & 1)) atomicAdd(a.d_theta + a.net.gb_off[l] + 16 * cg + (lane >> 1), bacc[l]);
        // output layer: 8 columns per thread, summed over the warp first
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = blast[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && 8 * cg + k < dout) atomicAdd(a.d_theta + a.net.gb_off[K - 1] + 8 * cg + k, v);
        }
    }
    ls_tc_dealloc(tmem);
}
