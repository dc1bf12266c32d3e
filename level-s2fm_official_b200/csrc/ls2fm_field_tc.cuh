// ls2fm_field_tc.cuh -- field FORWARD kernel with the MLP on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract as ls_field_forward_kernel (ls2fm_field.cuh): hash grid -> geometry MLP -> d(sdf)/dx by a reverse sweep ->
// radiance.  What changes is the execution model:
//   * CTA = 512 threads, one tile of 128 samples at a time (persistent over tiles).  Sample s = TMEM lane s.  A warp can
//     only touch the TMEM lane quarter (warp_id % 4), so warp w serves lanes 32 (w & 3) .. +31 and the "column group"
//     cg = w >> 2: four threads share a sample and split its columns (16 of the 64 hidden units each) and its hash levels
//     (levels 4 cg .. 4 cg + 3), which gives the SM 16 warps to hide gather / MUFU / TMEM latency behind.
//   * every matrix product is a 3xTF32 tcgen05.mma batch (fp32-level accuracy) issued by thread 0:
//       D[128 x N] (TMEM) = A[128 x K] (TMEM, written by the owning threads) * B[N x K] (weights in shared memory)
//   * activations never touch shared memory: the epilogue of a layer reads its accumulator slice with tcgen05.ld, applies
//     bias + softplus in registers, splits hi/lo and writes the next layer's A operand with tcgen05.st.  a_k stays in TMEM
//     for the reverse sweep (phi' = 1 - exp(-beta a_k)); the sweep's products use transposed weight copies (K-major).
//   * TMEM columns: A_k = [ (k-1)*128, +64 ) hi, [ (k-1)*128 + 64, +64 ) lo for k = 1..H;  D = [ H*128, +64 );
//     the encoding E (layer-0 input) aliases A_H (dead before A_H is written).
//   * encoding column order inside the kernel: [hash features (2L) | x/rescale (3) | 1 (bias of layer 0) | 0 pad]; the
//     staged W_0 / W_0^T are permuted accordingly, so every thread's columns are an aligned chunk of 8.
//   * shared memory holds only weights: W_l and W_l^T as hi/lo K-major operands (188 KB for 35-64-64-64-17).
#pragma once

#include "ls2fm_field.cuh"
#include "ls2fm_tc.cuh"

constexpr int LS_TC_THREADS = 512;

struct LsTcNet {
    int w_hi[LS2FM_MAX_LAYERS], w_lo[LS2FM_MAX_LAYERS];     // float offsets: W_l  [n_out_pad][k_in_pad]
    int wt_hi[LS2FM_MAX_LAYERS], wt_lo[LS2FM_MAX_LAYERS];   //                 W_l^T [n_in_pad][64]        (l < K-1)
    int bias[LS2FM_MAX_LAYERS];
    int n_out_pad[LS2FM_MAX_LAYERS], k_in_pad[LS2FM_MAX_LAYERS], n_in_pad[LS2FM_MAX_LAYERS];
    int wlast0;          // fp32 W_{K-1}[0, :]
    int weff, rad_pitch;
    int misc;            // mbarrier + tmem slot
    int red;             // [4][128][3] x 2 cross-column-group reduction scratch (normal, colour)
    int ring;            // streamed-weights variant: 3 slots of LS_TC_RING_SLOT floats
    int total;           // floats
    // image-only extension (global memory, never part of the forward kernels' shared-memory copy): the output layer as a
    // transposed operand B[n = hidden unit][k = output], hi then lo, for the reverse pass of the tensor-core backward kernel
    int wtl_hi, wtl_lo, kl_pad;
    int image_total;     // floats of the operand image
};

inline int ls_round_up(int v, int m) { return (v + m - 1) / m * m; }

// with_transposed = false: the values-only kernel keeps no W^T copies and no reduction scratch (about half the shared
// memory, which the hardware hands to L1 -- the coarse grid levels then stay L1-resident)
inline LsTcNet ls_plan_tc(const ls2fm_field_t& f, int rad_in_dim, bool with_transposed = true) {
    LsTcNet n;
    memset(&n, 0, sizeof(n));
    const int K = f.n_layers;
    int off = 0;
    for (int l = 0; l < K; ++l) {
        const bool last = l == K - 1;
        n.n_out_pad[l] = last ? 32 : LS_H;
        n.k_in_pad[l] = l == 0 ? ls_round_up(f.dims[0] + 1, 8) : LS_H;      // +1: the bias of layer 0 rides on a ones column
        n.n_in_pad[l] = l == 0 ? ls_round_up(f.dims[0], 16) : LS_H;
        n.w_hi[l] = off; off += n.n_out_pad[l] * n.k_in_pad[l];
        n.w_lo[l] = off; off += n.n_out_pad[l] * n.k_in_pad[l];
        if (!last && with_transposed) {
            n.wt_hi[l] = off; off += n.n_in_pad[l] * LS_H;
            n.wt_lo[l] = off; off += n.n_in_pad[l] * LS_H;
        }
        n.bias[l] = off; off += LS_H;
    }
    n.wlast0 = off; off += LS_H;
    n.rad_pitch = ls_round4(rad_in_dim > 0 ? rad_in_dim : 4);
    n.weff = off; off += 3 * n.rad_pitch + 4;
    n.misc = ls_round4(off); off = n.misc + 8;
    n.red = off; off += with_transposed ? 2 * 4 * LS_TC_M * 3 : 0;
    n.total = off;
    n.kl_pad = ls_round_up(f.dims[K], 8);
    n.wtl_hi = ls_round4(n.weff + 3 * ls_round4(LS2FM_MAX_RAD_IN) + 4) + 16;     // independent of the radiance block the image was built with
    n.wtl_lo = n.wtl_hi + LS_H * n.kl_pad;
    n.image_total = n.wtl_lo + LS_H * n.kl_pad;
    return n;
}

// Shared-memory plan of the streamed-weights forward kernel: only biases, W_last[0,:], W_eff, the reduction scratch and a 3-slot
// ring of 16 KB half-matrices (hi or lo) live in shared memory -- ~60 KB instead of 207 KB, and the hardware hands the rest to L1,
// which is what the hash gather wants.  Matrix offsets (w_hi, wt_hi, ...) of this plan are NOT used: they come from the image plan.
constexpr int LS_TC_RING_SLOT = 4096;
inline LsTcNet ls_plan_tc_ring(const ls2fm_field_t& f, int rad_in_dim) {
    LsTcNet n = ls_plan_tc(f, rad_in_dim);
    int off = 0;
    for (int l = 0; l < f.n_layers; ++l) { n.bias[l] = off; off += LS_H; }
    n.wlast0 = off; off += LS_H;
    n.weff = off; off += 3 * n.rad_pitch + 4;
    n.misc = ls_round4(off); off = n.misc + 16;        // +0 MMA barrier, +2/+4/+6 ring barriers, +12 TMEM slot
    n.red = off; off += 2 * 4 * LS_TC_M * 3;
    n.ring = ls_round_up(off, 32); off = n.ring + 3 * LS_TC_RING_SLOT;
    n.total = off;
    return n;
}
// biases, W_last[0, :] and W_eff straight from theta (the streamed-weights kernel keeps nothing else resident)
LS_DEV void ls_stage_small_tc(const LsFieldArgs& a, const LsTcNet& net, float* smem, int tid, int nt) {
    const int K = a.f.n_layers;
    for (int l = 0; l < K; ++l) {
        const int dout = a.f.dims[l + 1];
        const float* Bg = a.f.theta + a.net.gb_off[l];
        for (int e = tid; e < LS_H; e += nt) smem[net.bias[l] + e] = e < dout ? __ldg(Bg + e) : 0.f;
    }
    {
        const int dout = a.f.dims[K];
        const float* G = a.f.theta + a.net.gw_off[K - 1];
        for (int e = tid; e < LS_H; e += nt) smem[net.wlast0 + e] = __ldg(G + e * dout);
    }
    if (a.r.w_eff) {
        float* W = smem + net.weff;
        const int P = net.rad_pitch;
        for (int e = tid; e < 3 * P; e += nt) {
            const int c = e / P, i = e - c * P;
            W[e] = i < a.r.in_dim ? __ldg(a.r.w_eff + c * a.r.in_dim + i) : 0.f;
        }
        for (int e = tid; e < 4; e += nt) W[3 * P + e] = e < 3 ? __ldg(a.r.b_eff + e) : 0.f;
    }
}

// weights -> shared memory operands (hi/lo, K-major no-swizzle), zero padded
LS_DEV void ls_stage_weights_tc(const LsFieldArgs& a, const LsTcNet& net, float* smem, int tid, int nt, bool with_transposed = true) {
    const int K = a.f.n_layers;
    for (int l = 0; l < K; ++l) {
        const int din = a.f.dims[l], dout = a.f.dims[l + 1];
        const float* G = a.f.theta + a.net.gw_off[l];          // Wt [din][dout]
        const float* Bg = a.f.theta + a.net.gb_off[l];
        const int N = net.n_out_pad[l], Kp = net.k_in_pad[l];
        const int nh = din - 3;                                     // layer 0: kernel column k <-> reference input index
        for (int e = tid; e < N * Kp; e += nt) {
            const int j = e / Kp, i = e - j * Kp;
            float v = 0.f;
            if (j < dout) {
                if (l == 0) {
                    if (i < nh) v = __ldg(G + (3 + i) * dout + j);          // hash features first
                    else if (i < nh + 3) v = __ldg(G + (i - nh) * dout + j);   // then x / rescale
                    else if (i == nh + 3) v = __ldg(Bg + j);                 // ones column: bias
                } else if (i < din) v = __ldg(G + i * dout + j);
            }
            float hi, lo;
            ls_split_tf32(v, hi, lo);
            const int o = ls_op_off(j, i, N) >> 2;
            smem[net.w_hi[l] + o] = hi;
            smem[net.w_lo[l] + o] = lo;
        }
        if (l < K - 1 && with_transposed) {
            const int Nt = net.n_in_pad[l];
            for (int e = tid; e < Nt * LS_H; e += nt) {
                const int i = e / LS_H, j = e - i * LS_H;
                int src = i;
                if (l == 0) src = i < nh ? 3 + i : (i < nh + 3 ? i - nh : -1);
                const float v = (src >= 0 && src < din && j < dout) ? __ldg(G + src * dout + j) : 0.f;
                float hi, lo;
                ls_split_tf32(v, hi, lo);
                const int o = ls_op_off(i, j, Nt) >> 2;
                smem[net.wt_hi[l] + o] = hi;
                smem[net.wt_lo[l] + o] = lo;
            }
        }
        for (int e = tid; e < LS_H; e += nt) smem[net.bias[l] + e] = e < dout ? __ldg(Bg + e) : 0.f;
    }
    {
        const int dout = a.f.dims[K];
        const float* G = a.f.theta + a.net.gw_off[K - 1];
        for (int e = tid; e < LS_H; e += nt) smem[net.wlast0 + e] = __ldg(G + e * dout);     // W_{K-1}[0][e]
    }
    if (a.r.w_eff) {
        float* W = smem + net.weff;
        const int P = net.rad_pitch;
        for (int e = tid; e < 3 * P; e += nt) {
            const int c = e / P, i = e - c * P;
            W[e] = i < a.r.in_dim ? __ldg(a.r.w_eff + c * a.r.in_dim + i) : 0.f;
        }
        for (int e = tid; e < 4; e += nt) W[3 * P + e] = e < 3 ? __ldg(a.r.b_eff + e) : 0.f;
    }
}

// builds the operand image in global memory (same layout as the kernel's shared memory, floats [0, net.misc))
__global__ void ls_field_prepare_kernel(const LsFieldArgs a, const LsTcNet net, float* __restrict__ image) {
    const int tid = (int)(blockIdx.x * blockDim.x + threadIdx.x), nt = (int)(gridDim.x * blockDim.x);
    ls_stage_weights_tc(a, net, image, tid, nt);
    // extension: W_{K-1} as B[n = i][k = o] (K-major, rows o >= dout zero)
    const int K = a.f.n_layers, dout = a.f.dims[K], kl = net.kl_pad;
    const float* G = a.f.theta + a.net.gw_off[K - 1];           // Wt [64][dout]
    for (int e = tid; e < LS_H * kl; e += nt) {
        const int i = e / kl, o = e - i * kl;
        float hi, lo;
        ls_split_tf32(o < dout ? __ldg(G + i * dout + o) : 0.f, hi, lo);
        const int off = ls_op_off(i, o, LS_H) >> 2;
        image[net.wtl_hi + off] = hi;
        image[net.wtl_lo + off] = lo;
    }
}

// fast softplus for the tensor-core epilogue: ex2/lg2 based (MUFU rate), absolute error ~1e-9 at beta = 100
LS_DEV float ls_softplus_fast(float z, float beta, float inv_beta, float thr) {
    const float bz = z * beta;
#if defined(LS_HOSTSIM)
    return bz > thr ? z : log1pf(expf(bz)) * inv_beta;
#else
    // branch-free: the exponential is evaluated for every element (argument clamped at the threshold) and the threshold zone is a
    // select -- a per-element branch costs BSSY/BRA/BSYNC plus the serialisation of mixed warps, more than the two MUFU ops it skips
    const float e = ls_ex2(fminf(bz, thr) * 1.4426950408889634f);
    const float s = ls_lg2(1.f + e) * (0.6931471805599453f * inv_beta);
    return bz > thr ? z : s;
#endif
}

// RING: the weights stream from the operand image (img: its plan) through a 3-slot ring instead of living in shared memory; one
// matrix per MMA batch in the order W_0 .. W_{K-1}, W_{H-1}^T .. W_0^T, as hi / lo halves fetched 1-2 batches ahead (cp.async.bulk).
template <bool RING>
__global__ void __launch_bounds__(LS_TC_THREADS, 1) ls_field_forward_tc_kernel(const LsFieldArgs a, const LsTcNet net, const LsTcNet img) {
    LS_DYN_SMEM(smem);
    if (ls_n_samples(a.p) == 0) return;          // compacted launch with nothing left (sampler rounds): skip the weight staging
    const int t = threadIdx.x;
    const int cg = t >> 7;                       // column group 0..3 (warps cg*4 .. cg*4+3)
    const int row = ls_tc_row();                 // sample row in the tile = TMEM lane
    const int K = a.f.n_layers, H = K - 1, L = a.f.n_levels;
    const int din = a.f.dims[0], dout = a.f.dims[K], nh = din - 3;
    const float sp_beta = a.f.softplus_beta, sp_thr = a.f.softplus_threshold, inv_beta = 1.f / a.f.softplus_beta;
    if (RING) {
        ls_stage_small_tc(a, net, smem, threadIdx.x, blockDim.x);
    } else if (a.f.tc_image) {      // operand image prepared once per step: plain 16-byte copies
        const float4* src = reinterpret_cast<const float4*>(a.f.tc_image);
        float4* dst = reinterpret_cast<float4*>(smem);
        for (int e = threadIdx.x; e < net.misc / 4; e += blockDim.x) dst[e] = __ldg(src + e);
    } else {
        ls_stage_weights_tc(a, net, smem, threadIdx.x, blockDim.x);
    }
    ls_fence_smem_to_async();
    LsTcBar* bar = reinterpret_cast<LsTcBar*>(smem + net.misc);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + net.misc + (RING ? 12 : 4));
    const uint32_t tmem = ls_tc_alloc(slot);
    ls_tc_bar_init(bar);
    if (RING) {
        if (t == 0) { for (int s3 = 0; s3 < 3; ++s3) ls_bar_init1(reinterpret_cast<LsTcBar*>(smem + net.misc + 2 + 2 * s3)); }
        ls_fence_smem_to_async();
        __syncthreads();
    }
    uint32_t phase = 0;
    const bool need_nrm = a.out_nrm != nullptr || a.r.w_eff != nullptr;
    const int colD = H * 128;
    const int colE_hi = (H - 1) * 128, colE_lo = colE_hi + 64;
    float* red_n = smem + net.red;               // [4][128][3]
    float* red_c = red_n + 4 * LS_TC_M * 3;

    const int64_t n_pts = ls_n_samples(a.p);
    const int64_t n_tiles = (n_pts + LS_TC_M - 1) / LS_TC_M;
    // ---- weight ring (RING): entry e = (batch e >> 1, hi / lo half e & 1) lives in slot e % 3
    const int NB = 2 * H + 1;
    const int64_t my_tiles = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_entries = 2 * my_tiles * NB;
    int64_t gb = 0;
    auto ring_fetch = [&](int64_t e) {      // one thread
        if (e < n_entries) {
            const int b = (int)((e >> 1) % NB);
            int src, half;
            if (b < K) { src = img.w_hi[b]; half = img.n_out_pad[b] * img.k_in_pad[b]; }
            else { const int l = 2 * H - b; src = img.wt_hi[l]; half = img.n_in_pad[l] * LS_H; }
            const int s3 = (int)(e % 3);
            ls_bulk_g2s(smem + net.ring + s3 * LS_TC_RING_SLOT, a.f.tc_image + src + (int)(e & 1) * half, half * 4,
                        reinterpret_cast<LsTcBar*>(smem + net.misc + 2 + 2 * s3));
        }
    };
    auto ring_slot = [&](int64_t e) -> const float* {
        const int s3 = (int)(e % 3);
        ls_bar_wait(reinterpret_cast<LsTcBar*>(smem + net.misc + 2 + 2 * s3), (uint32_t)((e / 3) & 1));
        return smem + net.ring + s3 * LS_TC_RING_SLOT;
    };
    if (RING && t == 0) { ring_fetch(0); ring_fetch(1); ring_fetch(2); }
    // the issuing lane: 3xTF32 batch + commit
    auto issue = [&](int d_col, int a_hi, int a_lo, const float* w_hi, const float* w_lo, int N, int Kd) {
        if (RING) {
            const float* Wh = ring_slot(2 * gb);
            ls_tc_mma(tmem, d_col, a_lo, Wh, N, Kd, false);
            ls_tc_mma(tmem, d_col, a_hi, Wh, N, Kd, true);
            const float* Wl = ring_slot(2 * gb + 1);
            ls_tc_mma(tmem, d_col, a_hi, Wl, N, Kd, true);
        } else {
            ls_tc_mma_x3(tmem, d_col, a_hi, a_lo, w_hi, w_lo, N, Kd);
        }
        ls_tc_commit(bar);
    };
    auto batch_done = [&]() {
        ls_tc_wait(bar, phase);
        if (RING) {
            if ((t >> 5) == 0) { if (ls_elect()) { ring_fetch(2 * gb + 3); ring_fetch(2 * gb + 4); } }
            ++gb;
        }
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ------------------------------------------------ gather: (sample = row, levels 4cg .. 4cg+3)
        const int64_t i_in = tile * LS_TC_M + row;
        const bool valid = i_in < n_pts;
        float x[3] = {0.f, 0.f, 0.f}, u[3];
        int ray_id = 0;
        int64_t i = i_in;
        if (valid) ls_sample_point(a.p, i_in, x, &ray_id, &i);
        ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u);
        float J[4][2][3];
        if (4 * cg < L) {
            float e8[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int l = 4 * cg + r;
                float h[2], dh[2][3];
                ls_level_eval(a.f, l, u, h, dh, LS_DBG(a, 1));        // L % 4 == 0 on this path: all four levels of the group exist
                e8[2 * r] = h[0]; e8[2 * r + 1] = h[1];
#pragma unroll
                for (int fi = 0; fi < 2; ++fi)
#pragma unroll
                    for (int d = 0; d < 3; ++d) J[r][fi][d] = dh[fi][d] * a.inv_ext[d];
            }
            float hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ls_split_tf32(e8[q], hi[q], lo[q]);
            ls_tmem_st(tmem, colE_hi + 8 * cg, hi, 8);
            ls_tmem_st(tmem, colE_lo + 8 * cg, lo, 8);
        }
        if (cg == 0) {      // tail chunk: x / rescale, the ones column, zero padding
            float e8[8] = {ls_fdiv(x[0], a.f.rescale), ls_fdiv(x[1], a.f.rescale), ls_fdiv(x[2], a.f.rescale), 1.f, 0.f, 0.f, 0.f, 0.f};
            float hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ls_split_tf32(e8[q], hi[q], lo[q]);
            ls_tmem_st(tmem, colE_hi + nh, hi, 8);
            ls_tmem_st(tmem, colE_lo + nh, lo, 8);
        }
        // ------------------------------------------------ hidden layers
        for (int l = 0; l < H; ++l) {
            const int a_hi = l == 0 ? colE_hi : (l - 1) * 128, a_lo = a_hi + 64;
            ls_tc_sync_before_mma();
            if ((t >> 5) == 0 && ls_elect())        // one lane of a converged warp: uniform-register descriptors, back-to-back MMAs
                issue(colD, a_hi, a_lo, smem + net.w_hi[l], smem + net.w_lo[l], LS_H, net.k_in_pad[l]);
            batch_done();
            const float* bias = smem + net.bias[l] + 16 * cg;
            float v[16], hi[16], lo[16];
            ls_tmem_ld(tmem, colD + 16 * cg, v, 16);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float z = l == 0 ? v[q] : v[q] + bias[q];
                ls_split_tf32(ls_softplus_fast(z, sp_beta, inv_beta, sp_thr), hi[q], lo[q]);
            }
            ls_tmem_st(tmem, l * 128 + 16 * cg, hi, 16);
            ls_tmem_st(tmem, l * 128 + 64 + 16 * cg, lo, 16);
        }
        // ------------------------------------------------ output layer: cg 0 holds y[0..15], cg 1 holds y[16..23]
        float y[16];
        {
            ls_tc_sync_before_mma();
            if ((t >> 5) == 0 && ls_elect())
                issue(colD, (H - 1) * 128, (H - 1) * 128 + 64, smem + net.w_hi[K - 1], smem + net.w_lo[K - 1], 32, LS_H);
            batch_done();
#pragma unroll
            for (int q = 0; q < 16; ++q) y[q] = 0.f;
            const float* bias = smem + net.bias[K - 1];
            if (cg == 0) {
                ls_tmem_ld(tmem, colD, y, 16);
#pragma unroll
                for (int q = 0; q < 16; ++q) y[q] += bias[q];
            } else if (cg == 1) {
                ls_tmem_ld(tmem, colD + 16, y, 8);
#pragma unroll
                for (int q = 0; q < 8; ++q) y[q] += bias[16 + q];
            }
        }
        // ------------------------------------------------ reverse sweep: n = d(s*y0)/dx
        float nrm[3] = {0.f, 0.f, 0.f};
        if (need_nrm) {
            {   // w_H = phi'(a_H) * s * W_last[0, :]   (in place over A_H), my 16 columns
                const float* wl = smem + net.wlast0 + 16 * cg;
                const int cA = (H - 1) * 128 + 16 * cg;
                float ah[16], al[16], hi[16], lo[16];
                ls_tmem_ld2x16(tmem, cA, ah, cA + 64, al);
#pragma unroll
                for (int q = 0; q < 16; ++q) ls_split_tf32(ls_softplus_d1_from_a(ah[q] + al[q], sp_beta) * (a.s * wl[q]), hi[q], lo[q]);
                ls_tmem_st(tmem, cA, hi, 16);
                ls_tmem_st(tmem, cA + 64, lo, 16);
            }
            for (int l = H - 1; l >= 0; --l) {
                ls_tc_sync_before_mma();
                if ((t >> 5) == 0 && ls_elect())
                    issue(colD, l * 128, l * 128 + 64, smem + net.wt_hi[l], smem + net.wt_lo[l], net.n_in_pad[l], LS_H);
                batch_done();
                if (l > 0) {
                    const int cA = (l - 1) * 128 + 16 * cg;
                    float v[16], ah[16], al[16], hi[16], lo[16];
                    ls_tmem_ld2x16(tmem, cA, ah, cA + 64, al);
                    ls_tmem_ld(tmem, colD + 16 * cg, v, 16);
#pragma unroll
                    for (int q = 0; q < 16; ++q) ls_split_tf32(ls_softplus_d1_from_a(ah[q] + al[q], sp_beta) * v[q], hi[q], lo[q]);
                    ls_tmem_st(tmem, cA, hi, 16);
                    ls_tmem_st(tmem, cA + 64, lo, 16);
                } else {
                    // n = Je^T v0: my hash levels from the register Jacobians, cg 0 adds the x / rescale part
                    if (4 * cg < L) {
                        float v0[8];
                        ls_tmem_ld(tmem, colD + 8 * cg, v0, 8);
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int d = 0; d < 3; ++d) nrm[d] += v0[2 * r] * J[r][0][d] + v0[2 * r + 1] * J[r][1][d];
                    }
                    if (cg == 0) {
                        float vx[8];
                        ls_tmem_ld(tmem, colD + nh, vx, 8);
#pragma unroll
                        for (int d = 0; d < 3; ++d) nrm[d] += vx[d] / a.f.rescale;
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) red_n[(cg * LS_TC_M + row) * 3 + d] = nrm[d];
            __syncthreads();
#pragma unroll
            for (int d = 0; d < 3; ++d)
                nrm[d] = red_n[row * 3 + d] + red_n[(LS_TC_M + row) * 3 + d] + red_n[(2 * LS_TC_M + row) * 3 + d] + red_n[(3 * LS_TC_M + row) * 3 + d];
        }
        // ------------------------------------------------ radiance (affine o sigmoid): the input vector is split over the column groups
        if (a.r.w_eff) {
            const float* W = smem + net.weff;
            const int P = net.rad_pitch;
            const int nf = a.r.n_freq, kg = a.r.k_geo, kg2 = a.r.k_geo2;
            const int o_ray = 6, o_geo = 6 + 3 + 6 * nf, o_geo2 = o_geo + kg;
            float pr[3] = {0.f, 0.f, 0.f};
            float dir[3] = {0.f, 0.f, 0.f};
            if (valid) {
#pragma unroll
                for (int d = 0; d < 3; ++d) dir[d] = __ldg(a.p.ray + 3 * ray_id + d);
            }
            if (cg == 0) {          // bias, x, raw direction, geo features y[1..15]
#pragma unroll
                for (int c = 0; c < 3; ++c) pr[c] = W[3 * P + c];
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int c = 0; c < 3; ++c) pr[c] += W[c * P + d] * x[d] + W[c * P + o_ray + d] * dir[d];
#pragma unroll
                for (int k = 0; k < 15; ++k) {
                    if (k < kg) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) pr[c] += W[c * P + o_geo + k] * y[1 + k];
                    }
                }
            } else if (cg == 1) {   // normal, geo features y[16..]
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int c = 0; c < 3; ++c) pr[c] += W[c * P + 3 + d] * nrm[d];
#pragma unroll
                for (int k = 15; k < 23; ++k) {
                    if (k < kg) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) pr[c] += W[c * P + o_geo + k] * y[k - 15];
                    }
                }
            }
            // Fourier features of the view direction: frequencies split over cg 2, 3 (and 0, 1 when there are more)
            for (int k = (cg + 2) & 3; k < nf; k += 4) {
                const float fr = (float)(1 << k);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float ang = dir[d] * fr;
                    float sn, cs;
                    ls_sincos_fast(ang, &sn, &cs);
#pragma unroll
                    for (int c = 0; c < 3; ++c) pr[c] += W[c * P + o_ray + 3 + 6 * k + d] * sn + W[c * P + o_ray + 6 + 6 * k + d] * cs;
                }
            }
            if (kg2 > 0 && valid) {
                for (int k = cg; k < kg2; k += 4) {
                    const float yv = __ldg(a.r.geo2 + i * (kg2 + 1) + 1 + k);
#pragma unroll
                    for (int c = 0; c < 3; ++c) pr[c] += W[c * P + o_geo2 + k] * yv;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) red_c[(cg * LS_TC_M + row) * 3 + c] = pr[c];
            __syncthreads();
            if (cg == 0 && valid && a.out_rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    a.out_rgb[3 * i + c] = ls_sigmoid_fast(red_c[row * 3 + c] + red_c[(LS_TC_M + row) * 3 + c] + red_c[(2 * LS_TC_M + row) * 3 + c] +
                                                      red_c[(3 * LS_TC_M + row) * 3 + c]);
            }
        }
        if (valid) {
            if (cg == 0) {
                if (a.out_sdf) a.out_sdf[i] = a.s * y[0];
                if (a.out_nrm) { a.out_nrm[3 * i] = nrm[0]; a.out_nrm[3 * i + 1] = nrm[1]; a.out_nrm[3 * i + 2] = nrm[2]; }
            }
            if (a.out_y) {
                if (cg == 0) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) if (o < dout) a.out_y[i * dout + o] = y[o];
                } else if (cg == 1) {
#pragma unroll
                    for (int o = 0; o < 8; ++o) if (16 + o < dout) a.out_y[i * dout + 16 + o] = y[o];
                }
            }
        }
    }
    ls_tc_dealloc(tmem);
}

// ================================================================ SDF-only forward, two tiles in flight
// The sampler rounds, sphere tracing and SDF.infer_sdf need only y (no normals, no radiance): nothing has to survive the
// layer that consumes it, so a tile needs 192 TMEM columns (operand hi/lo 128 + accumulator 64) and TWO tiles fit.  All 512
// threads run one instruction stream that alternates between the two tiles: while the tensor core works on a batch of tile A
// the threads gather / run the epilogue of tile B, so the MMA round trips that serialise the one-tile kernel are hidden.
struct LsSdfTile {
    int64_t i;          // output index of this thread's sample
    bool valid;
    int col;            // TMEM column base of the tile context: [col, +64) operand hi, [col+64, +64) lo, [col+128, +64) accumulator
};

// net: compact plan (no W^T); img: the full plan the operand image was built with
__global__ void __launch_bounds__(LS_TC_THREADS, 1) ls_field_sdf_tc_kernel(const LsFieldArgs a, const LsTcNet net, const LsTcNet img) {
    LS_DYN_SMEM(smem);
    if (ls_n_samples(a.p) == 0) return;
    const int t = threadIdx.x;
    const int cg = t >> 7;
    const int row = ls_tc_row();
    const int K = a.f.n_layers, H = K - 1, L = a.f.n_levels;
    const int dout = a.f.dims[K], nh = a.f.dims[0] - 3;
    const float sp_beta = a.f.softplus_beta, sp_thr = a.f.softplus_threshold, inv_beta = 1.f / a.f.softplus_beta;
    if (a.f.tc_image) {      // copy W_l (hi | lo, contiguous) and the biases of every layer out of the full image
        for (int l = 0; l < K; ++l) {
            const int nw4 = (2 * net.n_out_pad[l] * net.k_in_pad[l]) / 4;
            const float4* src = reinterpret_cast<const float4*>(a.f.tc_image + img.w_hi[l]);
            float4* dst = reinterpret_cast<float4*>(smem + net.w_hi[l]);
            for (int e = threadIdx.x; e < nw4; e += blockDim.x) dst[e] = __ldg(src + e);
            for (int e = threadIdx.x; e < LS_H; e += blockDim.x) smem[net.bias[l] + e] = __ldg(a.f.tc_image + img.bias[l] + e);
        }
    } else {
        ls_stage_weights_tc(a, net, smem, threadIdx.x, blockDim.x, false);
    }
    ls_fence_smem_to_async();
    LsTcBar* bar0 = reinterpret_cast<LsTcBar*>(smem + net.misc);
    LsTcBar* bar1 = reinterpret_cast<LsTcBar*>(smem + net.misc + 2);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + net.misc + 4);
    const uint32_t tmem = ls_tc_alloc(slot);
    ls_tc_bar_init(bar0);
    ls_tc_bar_init(bar1);
    uint32_t ph[2] = {0, 0};

    const int64_t n_pts = ls_n_samples(a.p);
    const int64_t n_tiles = (n_pts + LS_TC_M - 1) / LS_TC_M;
    const int64_t n_pairs = (n_tiles + 1) / 2;

    // gather of one tile into its operand columns
    auto gather = [&](LsSdfTile& T, int64_t tile) {
        const int64_t i_in = tile * LS_TC_M + row;
        T.valid = tile < n_tiles && i_in < n_pts;
        T.i = i_in;
        float x[3] = {0.f, 0.f, 0.f}, u[3];
        int ray_id = 0;
        if (T.valid) ls_sample_point(a.p, i_in, x, &ray_id, &T.i);
        ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u);
        if (4 * cg < L) {
            float e8[8], hi[8], lo[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int l = 4 * cg + r;
                float h[2], dh[2][3];
                ls_level_eval(a.f, l, u, h, dh, LS_DBG(a, 1));        // L % 4 == 0 on this path: all four levels of the group exist
                e8[2 * r] = h[0]; e8[2 * r + 1] = h[1];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) ls_split_tf32(e8[q], hi[q], lo[q]);
            ls_tmem_st(tmem, T.col + 8 * cg, hi, 8);
            ls_tmem_st(tmem, T.col + 64 + 8 * cg, lo, 8);
        }
        if (cg == 0) {
            float e8[8] = {ls_fdiv(x[0], a.f.rescale), ls_fdiv(x[1], a.f.rescale), ls_fdiv(x[2], a.f.rescale), 1.f, 0.f, 0.f, 0.f, 0.f};
            float hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ls_split_tf32(e8[q], hi[q], lo[q]);
            ls_tmem_st(tmem, T.col + nh, hi, 8);
            ls_tmem_st(tmem, T.col + 64 + nh, lo, 8);
        }
    };
    // issue the MMA batch of layer l (l == H: output layer) for a tile; every thread passes the barrier inside
    auto issue = [&](const LsSdfTile& T, int l, LsTcBar* bar) {
        ls_tc_sync_before_mma();
        if ((t >> 5) == 0 && ls_elect()) {      // one lane of a converged warp: uniform-register descriptors, back-to-back MMAs
            ls_tc_mma_x3(tmem, T.col + 128, T.col, T.col + 64, smem + net.w_hi[l], smem + net.w_lo[l], l == H ? 32 : LS_H, net.k_in_pad[l]);
            ls_tc_commit(bar);
        }
    };
    // epilogue of hidden layer l: accumulator slice -> bias + softplus -> operand slice of the next layer
    auto epilogue = [&](const LsSdfTile& T, int l) {
        const float* bias = smem + net.bias[l] + 16 * cg;
        float v[16], hi[16], lo[16];
        ls_tmem_ld(tmem, T.col + 128 + 16 * cg, v, 16);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float z = l == 0 ? v[q] : v[q] + bias[q];
            ls_split_tf32(ls_softplus_fast(z, sp_beta, inv_beta, sp_thr), hi[q], lo[q]);
        }
        ls_tmem_st(tmem, T.col + 16 * cg, hi, 16);
        ls_tmem_st(tmem, T.col + 64 + 16 * cg, lo, 16);
    };
    auto output = [&](const LsSdfTile& T) {
        const float* bias = smem + net.bias[K - 1];
        if (cg == 0) {
            float y[16];
            ls_tmem_ld(tmem, T.col + 128, y, 16);
            if (T.valid) {
                if (a.out_sdf) a.out_sdf[T.i] = a.s * (y[0] + bias[0]);
                if (a.out_y) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) if (o < dout) a.out_y[T.i * dout + o] = y[o] + bias[o];
                }
            }
        } else if (cg == 1 && a.out_y && dout > 16) {
            float y[8];
            ls_tmem_ld(tmem, T.col + 128 + 16, y, 8);
            if (T.valid) {
#pragma unroll
                for (int o = 0; o < 8; ++o) if (16 + o < dout) a.out_y[T.i * dout + 16 + o] = y[o] + bias[16 + o];
            }
        }
    };

    LsSdfTile A, B;
    A.col = 0; B.col = 192;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        gather(A, 2 * pair);
        issue(A, 0, bar0);
        gather(B, 2 * pair + 1);
        issue(B, 0, bar1);
        for (int l = 0; l < H; ++l) {
            ls_tc_wait(bar0, ph[0]);
            epilogue(A, l);
            issue(A, l + 1, bar0);
            ls_tc_wait(bar1, ph[1]);
            epilogue(B, l);
            issue(B, l + 1, bar1);
        }
        ls_tc_wait(bar0, ph[0]);
        output(A);
        ls_tc_wait(bar1, ph[1]);
        output(B);
    }
    ls_tc_dealloc(tmem);
}
