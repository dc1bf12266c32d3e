// ls2fm_field_ws.cuh -- EXPERIMENTAL warp-specialised values-only field kernel (round-2 groundwork, opt-in through
// ls2fm_field_forward_ws).  Validated against the oracle under the host emulator; one run on a B200 at the very end of round 1
// (tools/ws_probe.py): bit-identical to the default kernel, 185 us vs 165 us on 262 144 samples -- it works, and four gather
// warps (16 loads in flight per thread) do not yet feed the MLP warps fast enough: next are 8-12 gather warps / deeper batching.
//
// Same contract as ls_field_sdf_tc_kernel (ls2fm_field_tc.cuh): hash grid -> geometry MLP -> y / sdf, nothing kept.  The field
// kernels are bound by two things that never overlap when every warp does both: the L1 tag stage of the scattered 8-byte table
// reads (one line per cycle: 128 reads per sample) and the serial MMA -> epilogue chain.  Here they run side by side:
//   * warps 0..15 (512 threads) are the MLP warps of the two-tiles-in-flight kernel, minus the gather: epilogues + MMA issue.
//   * warps 16..19 (128 threads) are gather warps: thread g evaluates all levels of sample g of the NEXT tile and writes the
//     encoding straight into a shared-memory A operand (K-major, hi and lo arrays, double-buffered), so layer 0 becomes an
//     smem x smem MMA; the other layers keep their A operand in TMEM.
//   * hand-over per buffer b: full[b] (128 gather arrivals, after fence.proxy.async) -> the issuing lane starts layer 0;
//     empty[b] (1 tcgen05.commit arrival when layer 0's MMAs have read the operand + 512 MLP-thread arrivals once they have read
//     their sample's output index) -> the gather warps may refill b.
//   * the MLP warps synchronise among themselves with a named barrier (bar.sync 1, 512), never with __syncthreads.
#pragma once

#include "ls2fm_field_tc.cuh"

// GW gather warps: 4 (one thread per sample, all levels) or 8 (two threads per sample, half of the levels each)
constexpr int ls_ws_threads(int gw) { return LS_TC_THREADS + 32 * gw; }

struct LsWsPlan {
    int eb;         // [2 buffers][hi | lo][128 x k_in_pad0] A operand of layer 0:  addr(m, k) = (k / 4) * 512 + m * 4 + k % 4
    int oi;         // [2][128] output index of each sample (int64), -1: past the end
    int mb;         // full[2], empty[2], mma[2] (8 bytes each)
    int total;      // floats
};
inline LsWsPlan ls_plan_ws(const LsTcNet& cnet) {
    LsWsPlan w;
    int off = ls_round_up(cnet.total, 32);
    w.eb = off; off += 2 * 2 * LS_TC_M * cnet.k_in_pad[0];
    w.oi = off; off += 2 * LS_TC_M * 2;
    w.mb = off; off += 12;
    w.total = off;
    return w;
}

// DEPTH: levels evaluated back to back before their results are stored (2 or 4): 16 or 32 table reads in flight per gather thread
template <int GW, int DEPTH>
__global__ void __launch_bounds__(ls_ws_threads(GW), 1) ls_field_sdf_ws_kernel(const LsFieldArgs a, const LsTcNet net, const LsTcNet img,
                                                                               const LsWsPlan ws) {
    static_assert(GW == 4 || GW == 8, "4 or 8 gather warps");
    static_assert(DEPTH == 2 || DEPTH == 4, "2 or 4 levels per batch");
    LS_DYN_SMEM(smem);
    if (ls_n_samples(a.p) == 0) return;
    const int t = threadIdx.x, warp = t >> 5;
    const bool is_gather = warp >= LS_TC_THREADS / 32;
    const int K = a.f.n_layers, H = K - 1, L = a.f.n_levels;
    const int dout = a.f.dims[K], nh = a.f.dims[0] - 3, Kp0 = net.k_in_pad[0];
    const float sp_beta = a.f.softplus_beta, sp_thr = a.f.softplus_threshold, inv_beta = 1.f / a.f.softplus_beta;
    // ---- weights (compact plan: W_l hi | lo and the biases), by everybody
    if (a.f.tc_image) {
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
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + net.misc + 4);
    LsMbar* full = reinterpret_cast<LsMbar*>(smem + ws.mb);          // full[0], full[1]
    LsMbar* empty = full + 2;                                        // empty[0], empty[1]
    LsMbar* bar0 = full + 4;                                         // MMA batches of tile context A / B (the MLP warps cannot use
    LsMbar* bar1 = full + 5;                                         //  a CTA-wide barrier anywhere in their loop)
    const uint32_t tmem = ls_tc_alloc(slot);
    if (t == 0) {
        ls_mb_init(bar0, 1); ls_mb_init(bar1, 1);
        ls_mb_init(full + 0, 32 * GW); ls_mb_init(full + 1, 32 * GW);
        ls_mb_init(empty + 0, 1 + LS_TC_THREADS); ls_mb_init(empty + 1, 1 + LS_TC_THREADS);
#if !defined(LS_HOSTSIM)
        asm volatile("fence.mbarrier_init.release.cluster;\n");
#endif
    }
    __syncthreads();

    const int64_t n_pts = ls_n_samples(a.p);
    const int64_t n_tiles = (n_pts + LS_TC_M - 1) / LS_TC_M;
    const int64_t my_tiles = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    float* EB = smem + ws.eb;
    long long* OI = reinterpret_cast<long long*>(smem + ws.oi);
    const int ebuf = 2 * LS_TC_M * Kp0;         // floats per buffer (hi then lo)

    if (is_gather) {
        // ================================================================ gather warps: one sample per thread, one tile ahead
        const int g = (t - LS_TC_THREADS) & (LS_TC_M - 1);         // sample row
        const int part = (t - LS_TC_THREADS) / LS_TC_M;            // which share of the levels (GW == 8: two threads per sample)
        const int l_per = L / (GW / 4), l_begin = part * l_per, l_end = l_begin + l_per;     // (L % 4 == 0: both shares are even)
        for (int64_t n = 0; n < my_tiles; ++n) {
            const int b = (int)(n & 1);
            ls_mb_wait(empty + b, (uint32_t)(((n >> 1) & 1) ^ 1));      // (first use of a buffer: passes at once)
            const int64_t tile = blockIdx.x + n * gridDim.x;
            const int64_t i_in = tile * LS_TC_M + g;
            const bool valid = i_in < n_pts;
            float x[3] = {0.f, 0.f, 0.f}, u[3];
            int ray_id = 0;
            int64_t io = i_in;
            if (valid) ls_sample_point(a.p, i_in, x, &ray_id, &io);
            ls_world_to_unit(a.f.bound_min, a.f.bound_max, x, u);
            float* Eh = EB + b * ebuf;
            float* El = Eh + LS_TC_M * Kp0;
            for (int l = l_begin; l < l_end; l += DEPTH) {  // two levels = four consecutive K columns = one 16-byte store per array
                float h[DEPTH][2], dh[2][3];
#pragma unroll
                for (int j = 0; j < DEPTH; ++j) ls_level_eval(a.f, l + j, u, h[j], dh);      // (all loads of the batch issue first)
#pragma unroll
                for (int j = 0; j < DEPTH; j += 2) {
                    float hi[4], lo[4];
                    ls_split_tf32(h[j][0], hi[0], lo[0]); ls_split_tf32(h[j][1], hi[1], lo[1]);
                    ls_split_tf32(h[j + 1][0], hi[2], lo[2]); ls_split_tf32(h[j + 1][1], hi[3], lo[3]);
                    const int o = ((2 * (l + j)) >> 2) * (4 * LS_TC_M) + g * 4;
                    ls_st4(Eh + o, make_float4(hi[0], hi[1], hi[2], hi[3]));
                    ls_st4(El + o, make_float4(lo[0], lo[1], lo[2], lo[3]));
                }
            }
            if (part == 0) {    // tail: x / rescale, the ones column (bias of layer 0), zero padding; the sample's output index
                float hi[4], lo[4];
                for (int d = 0; d < 3; ++d) ls_split_tf32(ls_fdiv(x[d], a.f.rescale), hi[d], lo[d]);
                hi[3] = 1.f; lo[3] = 0.f;
                const int o = (nh >> 2) * (4 * LS_TC_M) + g * 4;
                ls_st4(Eh + o, make_float4(hi[0], hi[1], hi[2], hi[3]));
                ls_st4(El + o, make_float4(lo[0], lo[1], lo[2], lo[3]));
                for (int k = nh + 4; k < Kp0; k += 4) {
                    const int oz = (k >> 2) * (4 * LS_TC_M) + g * 4;
                    ls_st4(Eh + oz, make_float4(0.f, 0.f, 0.f, 0.f));
                    ls_st4(El + oz, make_float4(0.f, 0.f, 0.f, 0.f));
                }
                OI[b * LS_TC_M + g] = valid ? (long long)io : -1;
            }
            ls_fence_smem_to_async();
            ls_mb_arrive(full + b);
        }
    } else {
        // ================================================================ MLP warps
        const int cg = t >> 7;
        const int row = ls_tc_row();
        struct Ctx { int64_t i; bool valid; int col; };
        Ctx A, B;
        A.col = 0; B.col = 192;
        uint32_t ph[2] = {0, 0};
        // layer 0 of tile n: operands are in shared memory (gather warps), accumulator in the context's D
        auto wait_mma = [&](LsMbar* bar, uint32_t& phase) {
            ls_mb_wait(bar, phase);
            phase ^= 1;
#if !defined(LS_HOSTSIM)
            asm volatile("tcgen05.fence::after_thread_sync;\n");
#endif
        };
        auto issue0 = [&](const Ctx& T, int64_t n, LsMbar* bar) {
            if ((t >> 5) == 0 && ls_elect()) {
                const int b = (int)(n & 1);
                ls_mb_wait(full + b, (uint32_t)((n >> 1) & 1));
                const float* Eh = EB + b * ebuf;
                const float* El = Eh + LS_TC_M * Kp0;
                const float* Wh = smem + net.w_hi[0];
                const float* Wl = smem + net.w_lo[0];
                const int a_lbo = LS_TC_M * 16, b_lbo = LS_H * 16;           // bytes between 4-column groups
                for (int pass = 0; pass < 3; ++pass) {                      // lo*hi, hi*lo, hi*hi
                    const float* Ao = pass == 0 ? El : Eh;
                    const float* Bo = pass == 1 ? Wl : Wh;
                    for (int ks = 0; ks < Kp0 / 8; ++ks)
                        ls_tc_mma_ss(tmem, T.col + 128, Ao + ks * 2 * (a_lbo / 4), a_lbo, Bo + ks * 2 * (b_lbo / 4), b_lbo, LS_H, pass > 0 || ks > 0);
                }
                ls_mb_commit(bar);
                ls_mb_commit(empty + b);
            }
        };
        // layer l >= 1 (l == H: output layer): A operand in TMEM, written by all MLP warps
        auto issue = [&](const Ctx& T, int l, LsMbar* bar) {
            ls_ws_sync_before_mma(1, LS_TC_THREADS);
            if ((t >> 5) == 0 && ls_elect()) {
                ls_tc_mma_x3(tmem, T.col + 128, T.col, T.col + 64, smem + net.w_hi[l], smem + net.w_lo[l], l == H ? 32 : LS_H, net.k_in_pad[l]);
                ls_mb_commit(bar);
            }
        };
        auto epilogue = [&](Ctx& T, int l, int64_t n) {
            if (l == 0) {       // first touch of the tile by this thread: its sample's output index, then release the buffer
                const int b = (int)(n & 1);
                ls_mb_wait(full + b, (uint32_t)((n >> 1) & 1));         // (already complete: acquires the gather warps' writes)
                const long long io = OI[b * LS_TC_M + row];
                T.valid = io >= 0;
                T.i = io;
                ls_mb_arrive(empty + b);
            }
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
        auto output = [&](const Ctx& T) {
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
        for (int64_t nA = 0; nA < my_tiles; nA += 2) {
            const int64_t nB = nA + 1;
            const bool hasB = nB < my_tiles;
            ls_ws_sync_before_mma(1, LS_TC_THREADS);        // everybody is done reading the accumulators of the previous pair
            issue0(A, nA, bar0);
            if (hasB) issue0(B, nB, bar1);
            for (int l = 0; l < H; ++l) {
                wait_mma(bar0, ph[0]);
                epilogue(A, l, nA);
                issue(A, l + 1, bar0);
                if (hasB) {
                    wait_mma(bar1, ph[1]);
                    epilogue(B, l, nB);
                    issue(B, l + 1, bar1);
                }
            }
            wait_mma(bar0, ph[0]);
            output(A);
            if (hasB) {
                wait_mma(bar1, ph[1]);
                output(B);
            }
        }
    }
    ls_tc_dealloc(tmem);
}
