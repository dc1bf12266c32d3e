/* Oracle: scalar C restatement of the tiny-cuda-nn 1.7 hash-grid arithmetic. [EXT]
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: tiny-cuda-nn is
 * not under /root/reference; this follows its published algorithm (SURVEY.md Appendix
 * A.1/A.2) and is anchored on the reference call sites models/base.py:17,37 and the
 * config built at models/base.py:124-139.
 *
 * Purpose: pin the INTEGER part (corner indices, bit-exact) and the fmaf/floorf weight
 * arithmetic of oracle/hashgrid.py and of the CUDA kernels with real single-precision
 * fused multiply-add, which PyTorch cannot express directly.
 *
 * Build (oracle/Makefile):  gcc -O2 -ffp-contract=off -shared -fPIC hashgrid_ref.c -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

typedef struct {
    float scale;
    uint32_t resolution;
    uint32_t offset;   /* entries */
    uint32_t size;     /* entries (hashmap_size of the level) */
    uint32_t hashed;
} ref_level_t;

/* per-level table; returns total number of entries */
uint32_t ref_grid_meta(int n_levels, int log2_hashmap_size, int base_resolution,
                       float per_level_scale, ref_level_t* out) {
    const float log2b = log2f(per_level_scale);
    uint32_t offset = 0;
    for (int l = 0; l < n_levels; ++l) {
        const float scale = exp2f((float)l * log2b) * (float)base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1u;
        const uint32_t max_params = UINT32_MAX / 2u;
        uint32_t params = powf((float)res, 3.0f) > (float)max_params ? max_params : res * res * res;
        params = (params + 7u) / 8u * 8u;
        const uint32_t cap = 1u << log2_hashmap_size;
        if (params > cap) params = cap;
        uint32_t stride = 1;
        for (int d = 0; d < 3 && stride <= params; ++d) stride *= res;
        out[l].scale = scale;
        out[l].resolution = res;
        out[l].offset = offset;
        out[l].size = params;
        out[l].hashed = params < stride;
        offset += params;
    }
    return offset;
}

static uint32_t corner_index(const ref_level_t* L, const uint32_t q[3]) {
    uint32_t stride = 1, index = 0;
    for (int d = 0; d < 3 && stride <= L->size; ++d) {
        index += q[d] * stride;
        stride *= L->resolution;
    }
    if (L->size < stride) index = q[0] ^ (q[1] * 2654435761u) ^ (q[2] * 805459861u);
    return index % L->size;
}

/* u [n,3] -> idx [n,8] (entry index inside the level, without offset), w [n,3] */
void ref_grid_corners(const ref_level_t* L, const float* u, int64_t n, uint32_t* idx, float* w) {
    for (int64_t i = 0; i < n; ++i) {
        uint32_t g[3];
        for (int d = 0; d < 3; ++d) {
            const float p = fmaf(L->scale, u[3 * i + d], 0.5f);
            const float fl = floorf(p);
            g[d] = (uint32_t)(int)fl;
            w[3 * i + d] = p - fl;
        }
        for (int c = 0; c < 8; ++c) {
            uint32_t q[3];
            for (int d = 0; d < 3; ++d) q[d] = g[d] + ((c >> d) & 1u);
            idx[8 * i + c] = corner_index(L, q);
        }
    }
}

/* full encode: u [n,3], table [n_entries*F] -> out [n, n_levels*F] */
void ref_grid_encode(const ref_level_t* levels, int n_levels, int F, const float* table,
                     const float* u, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        for (int l = 0; l < n_levels; ++l) {
            const ref_level_t* L = &levels[l];
            uint32_t idx[8];
            float w[3];
            ref_grid_corners(L, u + 3 * i, 1, idx, w);
            for (int f = 0; f < F; ++f) {
                float acc = 0.f;
                for (int c = 0; c < 8; ++c) {
                    float wc = 1.f;
                    for (int d = 0; d < 3; ++d) wc *= ((c >> d) & 1) ? w[d] : (1.f - w[d]);
                    acc += wc * table[((size_t)L->offset + idx[c]) * F + f];
                }
                out[i * (int64_t)(n_levels * F) + l * F + f] = acc;
            }
        }
    }
}
