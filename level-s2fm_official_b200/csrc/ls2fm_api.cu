// ls2fm_api.cu -- the C ABI of libls2fm_sm100.so (include/ls2fm.h): argument checking + kernel launches.
// Built by nvcc for sm_100a (see level-s2fm_official_b200/build.py).  No host synchronisation, no device
// allocation, no global state besides a thread-local error string and the cached SM count.
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <string>

#include "ls2fm_field.cuh"
#include "ls2fm_field_tc.cuh"
#include "ls2fm_field_bwtc.cuh"
#include "ls2fm_field_bwfeat.cuh"
#include "ls2fm_field_ws.cuh"
#include "ls2fm_render.cuh"
#include "ls2fm_sampler.cuh"
#include "ls2fm_trace.cuh"
#include "ls2fm_params.cuh"
#include "ls2fm_pose.cuh"

static thread_local std::string g_err;

static int ls_fail(const std::string& msg) {
    g_err = msg;
    return 1;
}

#if defined(LS_HOSTSIM)
static int ls_sm_count() { return 2; }
static int ls_max_smem() { return 227 * 1024; }
template <class K> static int ls_opt_in_smem(K, int) { return 0; }
static int ls_check_launch(const char*) { return 0; }
static void ls_memset_async(void* p, int v, size_t n, void*) { memset(p, v, n); }
#else
static void ls_memset_async(void* p, int v, size_t n, void* stream) { cudaMemsetAsync(p, v, n, (cudaStream_t)stream); }
static int ls_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
static int ls_max_smem() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (n <= 0) n = 48 * 1024;
    }
    return n;
}
// opt in to > 48 KB of dynamic shared memory, once per (kernel, device) and only ever upwards: the attribute call costs
// microseconds of host time in front of every launch otherwise.  The cache is PROCESS-wide (autograd runs backward launches on its
// own thread): the attribute is state of the function, and a second thread re-setting it to a smaller size would pull the limit
// down under a thread that cached the larger one.
struct LsSmemOptIn { const void* fn; int dev; int bytes; };
static LsSmemOptIn g_optin[128];
static int g_n_optin = 0;
static std::mutex g_optin_mutex;
static int ls_opt_in_smem_impl(const void* fn, int bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_optin_mutex);
    LsSmemOptIn* e = nullptr;
    for (int i = 0; i < g_n_optin; ++i)
        if (g_optin[i].fn == fn && g_optin[i].dev == dev) { e = &g_optin[i]; break; }
    if (e && e->bytes >= bytes) return 0;
    cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (err != cudaSuccess) return ls_fail(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
    if (e) e->bytes = bytes;
    else if (g_n_optin < 128) g_optin[g_n_optin++] = LsSmemOptIn{fn, dev, bytes};
    return 0;
}
template <class K> static int ls_opt_in_smem(K kernel, int bytes) { return ls_opt_in_smem_impl(reinterpret_cast<const void*>(kernel), bytes); }
static int ls_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ls_fail(std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
}
#endif

static int ls_check_field(const ls2fm_field_t* f) {
    if (!f) return ls_fail("field is NULL");
    if (!f->table || !f->theta) return ls_fail("field.table / field.theta is NULL");
    if (f->n_levels < 1 || f->n_levels > LS2FM_MAX_LEVELS) return ls_fail("field.n_levels out of range (1..16)");
    if (f->n_layers < 2 || f->n_layers > LS2FM_MAX_LAYERS) return ls_fail("field.n_layers out of range (2..4)");
    if (f->dims[0] != 3 + 2 * f->n_levels) return ls_fail("field.dims[0] must be 3 + 2*n_levels");
    for (int l = 1; l < f->n_layers; ++l)
        if (f->dims[l] != LS2FM_HIDDEN) return ls_fail("hidden width must be 64");
    if (f->dims[f->n_layers] < 1 || f->dims[f->n_layers] > LS2FM_MAX_OUT) return ls_fail("output width must be 1..20");
    if (!(f->scale_mlp != 0.f)) return ls_fail("field.scale_mlp must be non-zero");
    for (int d = 0; d < 3; ++d)
        if (!(f->bound_max[d] > f->bound_min[d])) return ls_fail("field bounds are empty");
    return 0;
}

static int ls_check_points(const ls2fm_points_t* p) {
    if (!p) return ls_fail("points is NULL");
    if (p->n < 0) return ls_fail("points.n < 0");
    if (p->n == 0) return 0;
    if (!p->xyz) {
        if (!p->center || !p->ray || !p->t) return ls_fail("ray mode needs center, ray and t");
        if (p->n_per_ray <= 0 || p->n_rays < 0) return ls_fail("ray mode needs n_per_ray > 0");
        if (p->n != (int64_t)p->n_rays * p->n_per_ray) return ls_fail("points.n != n_rays * n_per_ray");
        if (p->t_stride < p->n_per_ray + p->t_offset) return ls_fail("points.t_stride too small");
    }
    return 0;
}

static int ls_check_rad(const ls2fm_radiance_t* r, const ls2fm_field_t* f, const ls2fm_points_t* p) {
    if (!r->w_eff || !r->b_eff) return ls_fail("radiance.w_eff / b_eff is NULL");
    if (r->in_dim != 6 + 3 + 6 * r->n_freq + r->k_geo + r->k_geo2) return ls_fail("radiance.in_dim inconsistent");
    if (r->in_dim > LS2FM_MAX_RAD_IN) return ls_fail("radiance.in_dim too large");
    if (r->k_geo != f->dims[f->n_layers] - 1) return ls_fail("radiance.k_geo must equal field output width - 1");
    if (r->k_geo2 > 0 && !r->geo2) return ls_fail("radiance.geo2 is NULL but k_geo2 > 0");
    if (!p->ray || p->n_per_ray <= 0) return ls_fail("radiance needs the ray directions (points.ray, n_per_ray)");
    return 0;
}

static void ls_fill_args(LsFieldArgs& a, const ls2fm_field_t* field, const ls2fm_points_t* pts, const ls2fm_radiance_t* rad) {
    memset(&a, 0, sizeof(a));
    a.f = *field;
    a.p = *pts;
    if (rad) a.r = *rad;
    for (int d = 0; d < 3; ++d) a.inv_ext[d] = 1.0f / (field->bound_max[d] - field->bound_min[d]);
    a.s = field->sdf_sign / field->scale_mlp;
#if defined(LS_ABLATE)
    { const char* e = getenv("LS2FM_ABLATE"); a.dbg = e ? atoi(e) : 0; }
#endif
}

extern "C" {

int ls2fm_abi_version(void) { return LS2FM_ABI_VERSION; }

const char* ls2fm_last_error(void) { return g_err.c_str(); }

int ls2fm_grid_meta(const ls2fm_grid_cfg_t* cfg, ls2fm_level_t* levels, uint32_t* n_entries) {
    if (!cfg || !levels) return ls_fail("grid_meta: NULL argument");
    if (cfg->n_levels < 1 || cfg->n_levels > LS2FM_MAX_LEVELS) return ls_fail("grid_meta: n_levels out of range");
    if (cfg->n_features != 2) return ls_fail("grid_meta: n_features_per_level must be 2");
    // tiny-cuda-nn GridEncoding constructor [EXT]: float32 arithmetic throughout
    const float log2b = log2f(cfg->per_level_scale);
    uint32_t offset = 0;
    for (int l = 0; l < cfg->n_levels; ++l) {
        const float scale = exp2f((float)l * log2b) * (float)cfg->base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1u;
        const uint32_t max_params = UINT32_MAX / 2u;
        uint32_t params = powf((float)res, 3.0f) > (float)max_params ? max_params : res * res * res;
        params = (params + 7u) / 8u * 8u;
        const uint32_t cap = 1u << cfg->log2_hashmap_size;
        if (params > cap) params = cap;
        uint32_t stride = 1;
        for (int d = 0; d < 3 && stride <= params; ++d) stride *= res;
        levels[l].scale = scale;
        levels[l].resolution = res;
        levels[l].offset = offset;
        levels[l].size = params;
        levels[l].hashed = params < stride ? 1u : 0u;
        offset += params;
    }
    if (n_entries) *n_entries = offset;
    return 0;
}

int ls2fm_smem_bytes(const ls2fm_field_t* field, int backward, int with_radiance) {
    if (!field || field->n_layers < 2 || field->n_layers > LS2FM_MAX_LAYERS) return 0;
    const LsNet n = ls_plan_net(*field, with_radiance ? LS2FM_MAX_RAD_IN : 0, backward ? LS_BW_WARPS : 4, backward != 0);
    return n.total * (int)sizeof(float);
}

int ls2fm_ray_aabb(const float* rays_o, const float* rays_d, int64_t m, const float center[3], const float half_size[3],
                   float* hits_t, int32_t* hit_cnt, void* stream) {
    if (m < 0 || (m > 0 && (!rays_o || !rays_d || !hits_t))) return ls_fail("ray_aabb: bad arguments");
    if (m == 0) return 0;
    const int bs = 256;
    LS_LAUNCH(ls_ray_aabb_kernel, (unsigned)((m + bs - 1) / bs), bs, 0, stream, rays_o, rays_d, m, center[0], center[1],
              center[2], half_size[0], half_size[1], half_size[2], hits_t, hit_cnt);
    return ls_check_launch("ray_aabb");
}

int ls2fm_sample_uniform(const float* center, const float* ray, int32_t n_rays, int32_t n_samples, const float bound_min[3],
                         const float bound_max[3], float* t, float* hits_t, void* stream) {
    if (n_rays < 0 || n_samples <= 0 || (n_rays > 0 && (!center || !ray || !t))) return ls_fail("sample_uniform: bad arguments");
    if (n_rays == 0) return 0;
    // center / half size as models/Renderer.py:18-19: (max+min)/2, (max-min)/2
    float c[3], h[3];
    for (int d = 0; d < 3; ++d) { c[d] = (bound_max[d] + bound_min[d]) / 2.f; h[d] = (bound_max[d] - bound_min[d]) / 2.f; }
    const int wpb = 8;
    LS_LAUNCH(ls_sample_uniform_kernel, (unsigned)((n_rays + wpb - 1) / wpb), wpb * 32, 0, stream, center, ray, n_rays, n_samples,
              c[0], c[1], c[2], h[0], h[1], h[2], t, hits_t);
    return ls_check_launch("sample_uniform");
}

int ls2fm_ray_aabb_backward(const float* center, const float* ray, int32_t n_rays, int32_t n_samples, const float bound_min[3],
                            const float bound_max[3], const float* hits_t, const float* g, float* d_center, float* d_ray, void* stream) {
    if (n_rays < 0 || n_samples < 0) return ls_fail("ray_aabb_backward: bad sizes");
    if (n_rays > 0 && (!center || !ray || !hits_t || !g || !d_center || !d_ray)) return ls_fail("ray_aabb_backward: NULL argument");
    if (n_rays == 0) return 0;
    float c[3], h[3];
    for (int d = 0; d < 3; ++d) { c[d] = (bound_max[d] + bound_min[d]) / 2.f; h[d] = (bound_max[d] - bound_min[d]) / 2.f; }
    const int wpb = 8;
    LS_LAUNCH(ls_aabb_vjp_kernel, (unsigned)((n_rays + wpb - 1) / wpb), wpb * 32, 0, stream, center, ray, n_rays, n_samples, c[0], c[1], c[2],
              h[0], h[1], h[2], hits_t, g, d_center, d_ray);
    return ls_check_launch("ray_aabb_backward");
}

int ls2fm_grid_encode(const ls2fm_field_t* field, const float* u, int64_t m, float* enc, uint32_t* idx, void* stream) {
    if (!field || !field->table) return ls_fail("grid_encode: field/table is NULL");
    if (field->n_levels < 1 || field->n_levels > LS2FM_MAX_LEVELS) return ls_fail("grid_encode: n_levels out of range");
    if (m < 0 || (m > 0 && !u)) return ls_fail("grid_encode: bad arguments");
    if (m == 0) return 0;
    const int bs = 256;
    const int64_t total = m * field->n_levels;
    LS_LAUNCH(ls_grid_encode_kernel, (unsigned)((total + bs - 1) / bs), bs, 0, stream, *field, u, m, enc, idx);
    return ls_check_launch("grid_encode");
}

int ls2fm_grid_encode_backward(const ls2fm_field_t* field, const float* u, int64_t m, const float* g_enc, float* d_table,
                               float* d_u, void* stream) {
    if (!field || !field->table) return ls_fail("grid_encode_backward: field/table is NULL");
    if (field->n_levels < 1 || field->n_levels > LS2FM_MAX_LEVELS) return ls_fail("grid_encode_backward: n_levels out of range");
    if (m < 0 || (m > 0 && (!u || !g_enc))) return ls_fail("grid_encode_backward: bad arguments");
    if (m == 0) return 0;
    const int bs = 256;
    const int64_t total = m * field->n_levels;
    LS_LAUNCH(ls_grid_encode_backward_kernel, (unsigned)((total + bs - 1) / bs), bs, 0, stream, *field, u, m, g_enc, d_table, d_u);
    return ls_check_launch("grid_encode_backward");
}

int ls2fm_grid_encode_tangent(const ls2fm_field_t* field, const float* u, int64_t m, const float* v, const float* g_enc, float* t_enc,
                              float* d_table, float* d_u2, void* stream) {
    if (!field || !field->table) return ls_fail("grid_encode_tangent: field/table is NULL");
    if (field->n_levels < 1 || field->n_levels > LS2FM_MAX_LEVELS) return ls_fail("grid_encode_tangent: n_levels out of range");
    if (m < 0 || (m > 0 && (!u || !v))) return ls_fail("grid_encode_tangent: bad arguments");
    if ((d_table || d_u2) && !g_enc) return ls_fail("grid_encode_tangent: d_table / d_u2 need g_enc");
    if (m == 0) return 0;
    const int bs = 256;
    const int64_t total = m * field->n_levels;
    LS_LAUNCH(ls_grid_encode_tangent_kernel, (unsigned)((total + bs - 1) / bs), bs, 0, stream, *field, u, m, v, g_enc, t_enc, d_table, d_u2);
    return ls_check_launch("grid_encode_tangent");
}

static int ls_fill_params(LsParamArgs& a, const ls2fm_param_layer_t* geo, int32_t n_geo, const ls2fm_param_layer_t* rad, bool backward) {
    memset(&a, 0, sizeof(a));
    if (n_geo < 0 || n_geo > LS2FM_MAX_LAYERS || (n_geo > 0 && !geo)) return ls_fail("params: bad geometry layer list");
    for (int l = 0; l < n_geo; ++l) {
        a.geo[l] = geo[l];
        if (!geo[l].g || !geo[l].v || !geo[l].b || geo[l].din < 1 || geo[l].dout < 1) return ls_fail("params: NULL / empty geometry layer");
        if (backward && (!geo[l].dg || !geo[l].dv || !geo[l].db)) return ls_fail("params: NULL geometry gradient output");
    }
    a.n_geo = n_geo;
    if (rad) {
        for (int l = 0; l < 3; ++l) {
            a.rad[l] = rad[l];
            if (!rad[l].g || !rad[l].v || !rad[l].b) return ls_fail("params: NULL radiance layer");
            if (backward && (!rad[l].dg || !rad[l].dv || !rad[l].db)) return ls_fail("params: NULL radiance gradient output");
        }
        if (rad[0].dout != LS_H || rad[1].din != LS_H || rad[1].dout != LS_H || rad[2].din != LS_H || rad[2].dout != 3 ||
            rad[0].din < 1 || rad[0].din > LS_PP_MAX_IN)
            return ls_fail("params: the radiance decoder must be in -> 64 -> 64 -> 3 with in <= 68");
        a.has_rad = 1;
    }
    return 0;
}

int ls2fm_params_forward(const ls2fm_param_layer_t* geo, int32_t n_geo, const ls2fm_param_layer_t* rad, float* theta, float* w_eff,
                         float* b_eff, void* stream) {
    LsParamArgs a;
    if (ls_fill_params(a, geo, n_geo, rad, false)) return 1;
    if ((n_geo > 0 && !theta) || (rad && (!w_eff || !b_eff))) return ls_fail("params_forward: NULL output");
    a.theta = theta; a.w_eff = w_eff; a.b_eff = b_eff;
    const int smem = LS_PP_SMEM_FLOATS * (int)sizeof(float);
    if (ls_opt_in_smem(ls_params_forward_kernel, smem)) return 1;
    LS_LAUNCH(ls_params_forward_kernel, (unsigned)(1 + n_geo), LS_PP_THREADS, smem, stream, a);
    return ls_check_launch("params_forward");
}

int ls2fm_params_backward(const ls2fm_param_layer_t* geo, int32_t n_geo, const ls2fm_param_layer_t* rad, const float* d_theta,
                          const float* d_w_eff, const float* d_b_eff, int32_t accumulate, void* stream) {
    LsParamArgs a;
    if (ls_fill_params(a, geo, n_geo, rad, true)) return 1;
    a.d_theta = d_theta; a.d_w_eff = d_w_eff; a.d_b_eff = d_b_eff; a.accumulate = accumulate ? 1 : 0;
    const int smem = LS_PP_SMEM_FLOATS * (int)sizeof(float);
    if (ls_opt_in_smem(ls_params_backward_kernel, smem)) return 1;
    LS_LAUNCH(ls_params_backward_kernel, (unsigned)(1 + n_geo), LS_PP_THREADS, smem, stream, a);
    return ls_check_launch("params_backward");
}

int64_t ls2fm_field_image_floats(const ls2fm_field_t* field, const ls2fm_radiance_t* rad) {
    if (!field || field->n_layers < 2 || field->n_layers > LS2FM_MAX_LAYERS) return -1;
    return (int64_t)ls_plan_tc(*field, rad ? rad->in_dim : 0).image_total;
}

int ls2fm_field_prepare(const ls2fm_field_t* field, const ls2fm_radiance_t* rad, float* image, void* stream) {
    if (ls_check_field(field)) return 1;
    if (!image) return ls_fail("field_prepare: image is NULL");
    if (rad && (!rad->w_eff || !rad->b_eff)) return ls_fail("field_prepare: radiance.w_eff / b_eff is NULL");
    LsFieldArgs a;
    ls2fm_points_t pts;
    memset(&pts, 0, sizeof(pts));
    ls_fill_args(a, field, &pts, rad);
    a.net = ls_plan_net(*field, rad ? rad->in_dim : 0, 1, false);
    const LsTcNet net = ls_plan_tc(*field, rad ? rad->in_dim : 0);
    LS_LAUNCH(ls_field_prepare_kernel, 32, 256, 0, stream, a, net, image);
    return ls_check_launch("field_prepare");
}

static int ls_field_forward_checks(const ls2fm_field_t* field, const ls2fm_points_t* pts, const ls2fm_radiance_t* rad, float* out_rgb) {
    if (ls_check_field(field) || ls_check_points(pts)) return 1;
    if (pts->n == 0) return 0;
    if (rad && ls_check_rad(rad, field, pts)) return 1;
    if (out_rgb && !rad) return ls_fail("field_forward: out_rgb needs the radiance block");
    return 0;
}

int ls2fm_field_forward_simt(const ls2fm_field_t* field, const ls2fm_points_t* pts, const ls2fm_radiance_t* rad, float* out_y,
                             float* out_sdf, float* out_nrm, float* out_rgb, void* stream) {
    if (ls_field_forward_checks(field, pts, rad, out_rgb)) return 1;
    if (pts->n == 0) return 0;
    LsFieldArgs a;
    ls_fill_args(a, field, pts, rad);
    a.out_y = out_y; a.out_sdf = out_sdf; a.out_nrm = out_nrm; a.out_rgb = out_rgb;
    // as many warps per CTA as shared memory allows (one persistent CTA per SM)
    int nw = 16;
    const int64_t n_tiles = (pts->n + LS_WS - 1) / LS_WS;
    for (;; nw -= 4) {
        a.net = ls_plan_net(*field, rad ? rad->in_dim : 0, nw, false);
        if (a.net.total * (int)sizeof(float) <= ls_max_smem()) break;
        if (nw <= 4) return ls_fail("field_forward: network does not fit in shared memory");
    }
    const int smem = a.net.total * (int)sizeof(float);
    if (ls_opt_in_smem(ls_field_forward_kernel, smem)) return 1;
    int64_t grid = (n_tiles + nw - 1) / nw;
    if (grid > ls_sm_count()) grid = ls_sm_count();
    LS_LAUNCH(ls_field_forward_kernel, (unsigned)grid, nw * 32, smem, stream, a);
    return ls_check_launch("field_forward_simt");
}

int ls2fm_field_forward(const ls2fm_field_t* field, const ls2fm_points_t* pts, const ls2fm_radiance_t* rad, float* out_y,
                        float* out_sdf, float* out_nrm, float* out_rgb, void* stream) {
    if (ls_field_forward_checks(field, pts, rad, out_rgb)) return 1;
    if (pts->n == 0) return 0;
    LsFieldArgs a;
    ls_fill_args(a, field, pts, rad);
    a.out_y = out_y; a.out_sdf = out_sdf; a.out_nrm = out_nrm; a.out_rgb = out_rgb;
    a.net = ls_plan_net(*field, rad ? rad->in_dim : 0, 1, false);        // theta offsets
    const LsTcNet net = ls_plan_tc(*field, rad ? rad->in_dim : 0);
    const int smem = net.total * (int)sizeof(float);
    if (smem > ls_max_smem() || (field->n_levels & 3))     // operands too large / level groups not chunk-aligned: SIMT kernel
        return ls2fm_field_forward_simt(field, pts, rad, out_y, out_sdf, out_nrm, out_rgb, stream);
    const int64_t n_tiles = (pts->n + LS_TC_M - 1) / LS_TC_M;
    if (!rad && !out_nrm) {      // values only: the two-tiles-in-flight kernel, weights without their transposed copies
        const LsTcNet cnet = ls_plan_tc(*field, 0, false);
        const int csmem = cnet.total * (int)sizeof(float);
        if (ls_opt_in_smem(ls_field_sdf_tc_kernel, csmem)) return 1;
        const int64_t n_pairs = (n_tiles + 1) / 2;
        const int64_t grid = n_pairs < ls_sm_count() ? n_pairs : ls_sm_count();
        LS_LAUNCH(ls_field_sdf_tc_kernel, (unsigned)grid, LS_TC_THREADS, csmem, stream, a, cnet, net);
        return ls_check_launch("field_forward(sdf)");
    }
    int64_t grid = n_tiles < ls_sm_count() ? n_tiles : ls_sm_count();
    bool ring_ok = field->tc_image != nullptr;      // streamed weights: every half-matrix must fit a ring slot
    for (int l = 0; l < field->n_layers; ++l) ring_ok = ring_ok && net.n_out_pad[l] * net.k_in_pad[l] <= LS_TC_RING_SLOT;
    for (int l = 0; l < field->n_layers - 1; ++l) ring_ok = ring_ok && net.n_in_pad[l] * LS_H <= LS_TC_RING_SLOT;
    if (ring_ok) {
        const LsTcNet rnet = ls_plan_tc_ring(*field, rad ? rad->in_dim : 0);
        const int rsmem = rnet.total * (int)sizeof(float);
        if (ls_opt_in_smem(ls_field_forward_tc_kernel<true>, rsmem)) return 1;
        LS_LAUNCH(ls_field_forward_tc_kernel<true>, (unsigned)grid, LS_TC_THREADS, rsmem, stream, a, rnet, net);
    } else {
        if (ls_opt_in_smem(ls_field_forward_tc_kernel<false>, smem)) return 1;
        LS_LAUNCH(ls_field_forward_tc_kernel<false>, (unsigned)grid, LS_TC_THREADS, smem, stream, a, net, net);
    }
    return ls_check_launch("field_forward");
}

int ls2fm_field_forward_ws(const ls2fm_field_t* field, const ls2fm_points_t* pts, float* out_y, float* out_sdf, void* stream) {
    if (ls_field_forward_checks(field, pts, nullptr, nullptr)) return 1;
    if (pts->n == 0) return 0;
    if (field->n_levels & 3) return ls_fail("field_forward_ws: n_levels must be a multiple of 4");
    LsFieldArgs a;
    ls_fill_args(a, field, pts, nullptr);
    a.out_y = out_y; a.out_sdf = out_sdf;
    a.net = ls_plan_net(*field, 0, 1, false);
    const LsTcNet img = ls_plan_tc(*field, 0);
    const LsTcNet cnet = ls_plan_tc(*field, 0, false);
    const LsWsPlan ws = ls_plan_ws(cnet);
    const int smem = ws.total * (int)sizeof(float);
    if (smem > ls_max_smem()) return ls_fail("field_forward_ws: network does not fit in shared memory");
    const int64_t n_tiles = (pts->n + LS_TC_M - 1) / LS_TC_M;
    const int64_t n_pairs = (n_tiles + 1) / 2;
    const int64_t grid = n_pairs < ls_sm_count() ? n_pairs : ls_sm_count();
    // experiment knobs: LS2FM_WS_GATHER_WARPS = 4 (default) | 8, LS2FM_WS_DEPTH = 2 (default) | 4 levels per load batch
    const char* gw = getenv("LS2FM_WS_GATHER_WARPS");
    const char* dp = getenv("LS2FM_WS_DEPTH");
    const bool gw8 = gw && atoi(gw) == 8, dp4 = dp && atoi(dp) == 4 && (field->n_levels / (gw8 ? 2 : 1)) % 4 == 0;
#define LS_WS_LAUNCH(GWV, DPV)                                                                                          \
    do {                                                                                                                \
        if (ls_opt_in_smem(ls_field_sdf_ws_kernel<GWV, DPV>, smem)) return 1;                                           \
        LS_LAUNCH((ls_field_sdf_ws_kernel<GWV, DPV>), (unsigned)grid, ls_ws_threads(GWV), smem, stream, a, cnet, img, ws); \
    } while (0)
    if (gw8) { if (dp4) LS_WS_LAUNCH(8, 4); else LS_WS_LAUNCH(8, 2); }
    else { if (dp4) LS_WS_LAUNCH(4, 4); else LS_WS_LAUNCH(4, 2); }
#undef LS_WS_LAUNCH
    return ls_check_launch("field_forward_ws");
}

static int ls_field_backward_impl(const ls2fm_field_t* field, const ls2fm_points_t* pts, const ls2fm_radiance_t* rad, const float* g_y,
                                  const float* g_sdf, const float* g_nrm, const float* g_rgb, const float* saved_nrm,
                                  const float* saved_rgb, float* d_table, float* d_theta, float* d_w_eff, float* d_b_eff, float* d_geo2,
                                  const ls2fm_input_grads_t* in_grads, void* stream, int mode /* 0 auto, 1 simt, 2 tensor core or fail */) {
    if (ls_check_field(field) || ls_check_points(pts)) return 1;
    if (pts->n == 0) return 0;
    if (rad && ls_check_rad(rad, field, pts)) return 1;
    if (pts->ray_index || pts->n_active || pts->out_stride) return ls_fail("field_backward: compacted ray lists are forward-only");
    if (rad && g_rgb && (!saved_nrm || !saved_rgb)) return ls_fail("field_backward: saved_nrm / saved_rgb required with radiance");
    if (!rad && (g_rgb || d_w_eff || d_b_eff || d_geo2)) return ls_fail("field_backward: radiance gradients need the radiance block");
    LsFieldArgs a;
    ls_fill_args(a, field, pts, (rad && g_rgb) ? rad : nullptr);
    a.g_y = g_y; a.g_sdf = g_sdf; a.g_nrm = g_nrm; a.g_rgb = g_rgb;
    a.saved_nrm = saved_nrm; a.saved_rgb = saved_rgb;
    a.d_table = d_table; a.d_theta = d_theta; a.d_w_eff = d_w_eff; a.d_b_eff = d_b_eff; a.d_geo2 = d_geo2;
    if (d_table && (reinterpret_cast<uintptr_t>(d_table) & 7)) return ls_fail("field_backward: d_table must be 8-byte aligned (vector atomics)");
    if (in_grads) a.ig = *in_grads;
    const bool want_dx = a.ig.d_xyz || a.ig.d_center || a.ig.d_ray || a.ig.d_t;
    if (pts->xyz && (a.ig.d_center || a.ig.d_t)) return ls_fail("field_backward: d_center / d_t need ray mode (points.xyz == NULL)");
    if (!pts->xyz && a.ig.d_xyz) return ls_fail("field_backward: d_xyz needs explicit points; ray mode yields d_center / d_ray / d_t");
    if (a.ig.d_ray && !pts->ray) return ls_fail("field_backward: d_ray needs points.ray");
    const bool with_rad = rad && g_rgb;
    const bool tan = with_rad || g_nrm;
    const int KL = field->n_layers;
    // ---- tensor-core kernel: needs the operand image (weights stream from it), the 2-channel form (normals carry gradient),
    //      chunk-aligned level groups and matrices that fit a ring slot; everything else runs the fp32-SIMT kernel below
    //      Launches below LS_BT_MIN_SAMPLES stay on the SIMT kernel: a few tiles do not amortise the tensor-core kernel's set-up.
    // (a launch without a gradient on the normals -- RadF.Geo_enc under dual_field: g_y only -- runs the single-channel variant
    //  ls_field_backward_feat_tc_kernel: 128 samples per tile instead of 64 samples x 2 channels)
    //  position gradients: the tensor-core kernel parks the encoding adjoints in ig.workspace and ls_field_posgrad_kernel finishes)
    const bool tc_ok = field->tc_image && (field->n_levels & 3) == 0 && (!want_dx || a.ig.workspace) && (tan || g_y || g_sdf);
    if (!want_dx) a.ig.workspace = nullptr;
    if (mode == 2 && !tc_ok) return ls_fail("field_backward_tc: needs field.tc_image, an upstream gradient, n_levels % 4 == 0 and (for position gradients) a workspace");
    if ((mode == 2 || (mode == 0 && pts->n >= LS_BT_MIN_SAMPLES)) && tc_ok) {
        const LsTcNet img = ls_plan_tc(*field, with_rad ? rad->in_dim : 0);
        const LsBtNet net = ls_plan_bt(*field, with_rad ? rad->in_dim : 0);
        const int smem = net.total * (int)sizeof(float);
        bool fits = smem <= ls_max_smem() && img.k_in_pad[0] <= LS_BT_EROWS && img.n_in_pad[0] <= LS_BT_EROWS;
        if (with_rad) fits = fits && 3 * (3 * (rad->in_dim - rad->k_geo) + 3) <= LS_BT_THREADS;      // W_eff-gradient reducers: 3 sample ranges per pair
        for (int l = 0; l < KL - 1; ++l) fits = fits && img.n_out_pad[l] * img.k_in_pad[l] <= LS_BT_SLOT && img.n_in_pad[l] * LS_H <= LS_BT_SLOT;
        if (fits) {
            a.net = ls_plan_net(*field, with_rad ? rad->in_dim : 0, 1, false);        // theta offsets
            // no gradient on the normals (RadF.Geo_enc under dual_field: g_y only): the single-channel kernel, 128 samples per tile
            const int64_t n_tiles = tan ? (pts->n + LS_BT_TILE - 1) / LS_BT_TILE : (pts->n + LS_BF_TILE - 1) / LS_BF_TILE;
            const int64_t grid = n_tiles < ls_sm_count() ? n_tiles : ls_sm_count();
#define LS_BT_LAUNCH(KERNEL, KV)                                                                           \
            do {                                                                                           \
                if (ls_opt_in_smem(KERNEL<KV>, smem)) return 1;                                            \
                LS_LAUNCH((KERNEL<KV>), (unsigned)grid, LS_BT_THREADS, smem, stream, a, img, net);         \
            } while (0)
            if (tan) { if (KL == 2) LS_BT_LAUNCH(ls_field_backward_tc_kernel, 2); else if (KL == 3) LS_BT_LAUNCH(ls_field_backward_tc_kernel, 3); else LS_BT_LAUNCH(ls_field_backward_tc_kernel, 4); }
            else { if (KL == 2) LS_BT_LAUNCH(ls_field_backward_feat_tc_kernel, 2); else if (KL == 3) LS_BT_LAUNCH(ls_field_backward_feat_tc_kernel, 3); else LS_BT_LAUNCH(ls_field_backward_feat_tc_kernel, 4); }
#undef LS_BT_LAUNCH
            if (ls_check_launch("field_backward(tc)")) return 1;
            if (want_dx) {
                int64_t pg = (pts->n + 255) / 256;
                if (pg > 8 * ls_sm_count()) pg = 8 * ls_sm_count();
                LS_LAUNCH(ls_field_posgrad_kernel, (unsigned)pg, 256, 0, stream, a);
                return ls_check_launch("field_backward(posgrad)");
            }
            return 0;
        }
        if (mode == 2) return ls_fail("field_backward_tc: the network does not fit the tensor-core kernel");
    }
    a.ig.workspace = nullptr;       // (the fp32-SIMT kernel computes the position gradient itself)
    a.net = ls_plan_net(*field, with_rad ? rad->in_dim : 0, LS_BW_WARPS, true);
    const int smem = a.net.total * (int)sizeof(float);
    if (smem > ls_max_smem()) return ls_fail("field_backward: network does not fit in shared memory");
    const int64_t n_ct = (pts->n + LS_WS * LS_BW_WARPS - 1) / (LS_WS * LS_BW_WARPS);
    int64_t grid = n_ct < ls_sm_count() ? n_ct : ls_sm_count();
#define LS_BW_LAUNCH(TANV, KV)                                                                          \
    do {                                                                                                \
        if (ls_opt_in_smem(ls_field_backward_kernel<TANV, KV>, smem)) return 1;                         \
        LS_LAUNCH((ls_field_backward_kernel<TANV, KV>), (unsigned)grid, LS_BW_THREADS, smem, stream, a); \
    } while (0)
    if (tan) { if (KL == 2) LS_BW_LAUNCH(true, 2); else if (KL == 3) LS_BW_LAUNCH(true, 3); else LS_BW_LAUNCH(true, 4); }
    else { if (KL == 2) LS_BW_LAUNCH(false, 2); else if (KL == 3) LS_BW_LAUNCH(false, 3); else LS_BW_LAUNCH(false, 4); }
#undef LS_BW_LAUNCH
    return ls_check_launch("field_backward");
}

int64_t ls2fm_field_backward_workspace_floats(int64_t n_samples) { return n_samples < 0 ? -1 : n_samples * (int64_t)LS_PG_PITCH; }

int ls2fm_field_backward(const ls2fm_field_t* field, const ls2fm_points_t* pts, const ls2fm_radiance_t* rad, const float* g_y,
                         const float* g_sdf, const float* g_nrm, const float* g_rgb, const float* saved_nrm,
                         const float* saved_rgb, float* d_table, float* d_theta, float* d_w_eff, float* d_b_eff, float* d_geo2,
                         const ls2fm_input_grads_t* in_grads, void* stream) {
    return ls_field_backward_impl(field, pts, rad, g_y, g_sdf, g_nrm, g_rgb, saved_nrm, saved_rgb, d_table, d_theta, d_w_eff, d_b_eff,
                                  d_geo2, in_grads, stream, 0);
}

int ls2fm_field_backward_simt(const ls2fm_field_t* field, const ls2fm_points_t* pts, const ls2fm_radiance_t* rad, const float* g_y,
                              const float* g_sdf, const float* g_nrm, const float* g_rgb, const float* saved_nrm,
                              const float* saved_rgb, float* d_table, float* d_theta, float* d_w_eff, float* d_b_eff, float* d_geo2,
                              const ls2fm_input_grads_t* in_grads, void* stream) {
    return ls_field_backward_impl(field, pts, rad, g_y, g_sdf, g_nrm, g_rgb, saved_nrm, saved_rgb, d_table, d_theta, d_w_eff, d_b_eff,
                                  d_geo2, in_grads, stream, 1);
}

int ls2fm_field_backward_tc(const ls2fm_field_t* field, const ls2fm_points_t* pts, const ls2fm_radiance_t* rad, const float* g_y,
                            const float* g_sdf, const float* g_nrm, const float* g_rgb, const float* saved_nrm,
                            const float* saved_rgb, float* d_table, float* d_theta, float* d_w_eff, float* d_b_eff, float* d_geo2,
                            const ls2fm_input_grads_t* in_grads, void* stream) {
    return ls_field_backward_impl(field, pts, rad, g_y, g_sdf, g_nrm, g_rgb, saved_nrm, saved_rgb, d_table, d_theta, d_w_eff, d_b_eff,
                                  d_geo2, in_grads, stream, 2);
}

int ls2fm_composite_forward(const float* ray, const float* t, const float* sdf, const float* rgbs, const float* nrm,
                            const float* beta_param, float beta_speed, const float bgcolor[3], int32_t n_rays, int32_t n_samples,
                            float* rgb, float* depth, float* normal, float* opacity, void* stream) {
    if (n_rays < 0 || n_samples < 2) return ls_fail("composite_forward: need n_samples >= 2");
    if (n_rays > 0 && (!ray || !t || !sdf || !beta_param)) return ls_fail("composite_forward: NULL input");
    if (n_rays == 0) return 0;
    const int wpb = 4;
    LS_LAUNCH(ls_composite_forward_kernel, (unsigned)((n_rays + wpb - 1) / wpb), wpb * 32, 0, stream, ray, t, sdf, rgbs, nrm,
              beta_param, beta_speed, bgcolor[0], bgcolor[1], bgcolor[2], n_rays, n_samples, rgb, depth, normal, opacity);
    return ls_check_launch("composite_forward");
}

int ls2fm_composite_backward(const float* ray, const float* t, const float* sdf, const float* rgbs, const float* nrm,
                             const float* beta_param, float beta_speed, const float bgcolor[3], int32_t n_rays, int32_t n_samples,
                             const float* g_rgb, const float* g_depth, const float* g_normal, float* d_sdf, float* d_rgbs,
                             float* d_nrm, float* d_beta_param, float* d_ray, float* d_t, void* stream) {
    if (n_rays < 0 || n_samples < 2) return ls_fail("composite_backward: need n_samples >= 2");
    if (n_samples - 1 > 32 * LS_MAX_CHUNKS) return ls_fail("composite_backward: at most 257 samples per ray");
    if (n_rays > 0 && (!ray || !t || !sdf || !beta_param)) return ls_fail("composite_backward: NULL input");
    if (n_rays == 0) return 0;
    const int wpb = 4;
    LS_LAUNCH(ls_composite_backward_kernel, (unsigned)((n_rays + wpb - 1) / wpb), wpb * 32, 0, stream, ray, t, sdf, rgbs, nrm,
              beta_param, beta_speed, bgcolor[0], bgcolor[1], bgcolor[2], n_rays, n_samples, g_rgb, g_depth, g_normal, d_sdf,
              d_rgbs, d_nrm, d_beta_param, d_ray, d_t);
    return ls_check_launch("composite_backward");
}

// workspace carving for the error-bounded sampler (all 16-byte aligned)
struct LsSamplerWs { size_t D, S, beta_plus, fine, iters, state, cnt, ray_index, hits, total; };
static LsSamplerWs ls_sampler_ws(const ls2fm_sampler_cfg_t* c, int32_t R) {
    LsSamplerWs w;
    const size_t Mmax = (size_t)c->n_samples * (c->max_upsample_iter + 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    w.D = take((size_t)R * Mmax * 4); w.S = take((size_t)R * Mmax * 4);
    w.beta_plus = take((size_t)R * 4); w.fine = take((size_t)R * c->n_final * 4); w.iters = take((size_t)R * 4);
    w.state = take((size_t)R * 4); w.cnt = take((size_t)(c->max_upsample_iter + 2) * 4);
    w.ray_index = take((size_t)(c->max_upsample_iter + 2) * R * 4); w.hits = take((size_t)R * 8);
    w.total = off;
    return w;
}

int64_t ls2fm_sampler_workspace_bytes(const ls2fm_sampler_cfg_t* cfg, int32_t n_rays) {
    if (!cfg || n_rays < 0 || cfg->n_samples < 2 || cfg->max_upsample_iter < 0) return -1;
    return (int64_t)ls_sampler_ws(cfg, n_rays).total;
}

int ls2fm_sample_error_bounded(const ls2fm_field_t* sdf_field, const float* beta_param, const ls2fm_sampler_cfg_t* cfg,
                               const float* center, const float* ray, int32_t n_rays, void* workspace, float* t_out,
                               float* beta_plus, float* iters, void* stream) {
    if (ls_check_field(sdf_field)) return 1;
    if (!cfg || cfg->n_samples < 2 || cfg->n_final < 1 || cfg->max_upsample_iter < 0 || cfg->max_bisection_itr < 0)
        return ls_fail("sample_error_bounded: bad sampler config");
    if (n_rays < 0 || (n_rays > 0 && (!center || !ray || !beta_param || !workspace || !t_out)))
        return ls_fail("sample_error_bounded: NULL argument");
    if (n_rays == 0) return 0;
    const LsSamplerWs w = ls_sampler_ws(cfg, n_rays);
    char* base = (char*)workspace;
    LsSamplerArgs a;
    memset(&a, 0, sizeof(a));
    a.center = center; a.ray = ray; a.beta_param = beta_param;
    a.n_rays = n_rays; a.N = cfg->n_samples; a.Nf = cfg->n_final; a.max_iter = cfg->max_upsample_iter;
    a.max_bisect = cfg->max_bisection_itr; a.Mmax = cfg->n_samples * (cfg->max_upsample_iter + 1);
    a.eps = cfg->eps; a.beta_speed = cfg->beta_speed;
    a.cx = (sdf_field->bound_max[0] + sdf_field->bound_min[0]) / 2.f; a.hx = (sdf_field->bound_max[0] - sdf_field->bound_min[0]) / 2.f;
    a.cy = (sdf_field->bound_max[1] + sdf_field->bound_min[1]) / 2.f; a.hy = (sdf_field->bound_max[1] - sdf_field->bound_min[1]) / 2.f;
    a.cz = (sdf_field->bound_max[2] + sdf_field->bound_min[2]) / 2.f; a.hz = (sdf_field->bound_max[2] - sdf_field->bound_min[2]) / 2.f;
    a.D = (float*)(base + w.D); a.S = (float*)(base + w.S); a.beta_plus = (float*)(base + w.beta_plus);
    a.fine = (float*)(base + w.fine); a.iters = (float*)(base + w.iters); a.state = (int*)(base + w.state);
    a.cnt = (int*)(base + w.cnt); a.ray_index = (int*)(base + w.ray_index); a.hits = (float*)(base + w.hits);

    const int wpb = 4;
    const unsigned grid = (unsigned)((n_rays + wpb - 1) / wpb);
    const int smem_round = wpb * 5 * a.Mmax * (int)sizeof(float);
    const int smem_final = wpb * (a.N + a.Nf) * (int)sizeof(float);
    if (smem_round > ls_max_smem()) return ls_fail("sample_error_bounded: too many samples per ray for shared memory");
    if (ls_opt_in_smem(ls_sampler_round_kernel, smem_round)) return 1;
    LS_LAUNCH(ls_sampler_init_kernel, grid, wpb * 32, 0, stream, a);
    if (ls_check_launch("sampler_init")) return 1;
    ls2fm_points_t p;
    memset(&p, 0, sizeof(p));
    p.center = center; p.ray = ray; p.t = a.D;
    p.n_rays = n_rays; p.n_per_ray = a.N; p.n = (int64_t)n_rays * a.N;
    p.t_stride = a.Mmax; p.out_stride = a.Mmax;
    for (int it = 0; it <= a.max_iter; ++it) {
        // SDF of the samples drawn for this round, on the compacted list of rays that asked for them
        p.t_offset = a.N * it; p.out_offset = a.N * it;
        p.ray_index = a.ray_index + (size_t)it * n_rays;
        p.n_active = a.cnt + it;
        if (ls2fm_field_forward(sdf_field, &p, nullptr, nullptr, a.S, nullptr, nullptr, stream)) return 1;
        LS_LAUNCH(ls_sampler_round_kernel, grid, wpb * 32, wpb * 5 * a.N * (it + 1) * (int)sizeof(float), stream, a, it);
        if (ls_check_launch("sampler_round")) return 1;
    }
    LS_LAUNCH(ls_sampler_finalize_kernel, grid, wpb * 32, smem_final, stream, a, t_out, beta_plus, iters);
    return ls_check_launch("sampler_finalize");
}

int ls2fm_sphere_trace(const ls2fm_field_t* sdf_field, const float* ray0, const float* ray_dir, int64_t m, float sdf_threshold,
                       int32_t iters_max, float* track, int32_t* n_unfinished, float* t_near, float* t_far, float* acc_end,
                       void* stream) {
    if (ls_check_field(sdf_field)) return 1;
    if (m < 0 || iters_max < 1 || iters_max > 1024) return ls_fail("sphere_trace: bad m / iters_max");
    if (m > 0 && (!ray0 || !ray_dir || !track || !n_unfinished || !t_near || !t_far || !acc_end))
        return ls_fail("sphere_trace: NULL argument");
    ls_memset_async(n_unfinished, 0, sizeof(int32_t) * (size_t)(iters_max + 1), stream);
    if (m == 0) return 0;
    LsTraceArgs t;
    memset(&t, 0, sizeof(t));
    ls2fm_points_t pts;
    memset(&pts, 0, sizeof(pts));
    ls_fill_args(t.fa, sdf_field, &pts, nullptr);
    t.ray0 = ray0; t.dir = ray_dir; t.m = m;
    t.cx = (sdf_field->bound_max[0] + sdf_field->bound_min[0]) / 2.f; t.hx = (sdf_field->bound_max[0] - sdf_field->bound_min[0]) / 2.f;
    t.cy = (sdf_field->bound_max[1] + sdf_field->bound_min[1]) / 2.f; t.hy = (sdf_field->bound_max[1] - sdf_field->bound_min[1]) / 2.f;
    t.cz = (sdf_field->bound_max[2] + sdf_field->bound_min[2]) / 2.f; t.hz = (sdf_field->bound_max[2] - sdf_field->bound_min[2]) / 2.f;
    t.thr = sdf_threshold; t.iters_max = iters_max;
    t.track = track; t.cnt = n_unfinished; t.t_near = t_near; t.t_far = t_far; t.acc_e_hist = acc_end;
    const int64_t n_tiles = (m + 3) / 4;
    // few rays: spread the tiles over the SMs (one warp tile = 4 rays), at most 16 warps per CTA
    int nw = 16;
    while (nw > 1 && (n_tiles + nw - 1) / nw < ls_sm_count()) nw >>= 1;
    for (;; nw >>= 1) {
        t.fa.net = ls_plan_net(*sdf_field, 0, nw, false);
        if (t.fa.net.total * (int)sizeof(float) <= ls_max_smem()) break;
        if (nw <= 1) return ls_fail("sphere_trace: network does not fit in shared memory");
    }
    const int smem = t.fa.net.total * (int)sizeof(float);
    if (ls_opt_in_smem(ls_sphere_trace_kernel, smem)) return 1;
    int64_t grid = (n_tiles + nw - 1) / nw;
    if (grid > ls_sm_count()) grid = ls_sm_count();
    LS_LAUNCH(ls_sphere_trace_kernel, (unsigned)grid, nw * 32, smem, stream, t);
    return ls_check_launch("sphere_trace");
}

int ls2fm_grid_points(int32_t n, double step, const double origin[3], int64_t begin, int64_t count, float* xyz, void* stream) {
    if (n < 2 || count < 0 || begin < 0 || begin + count > (int64_t)n * n * n) return ls_fail("grid_points: bad range");
    if (count > 0 && (!xyz || !origin)) return ls_fail("grid_points: NULL argument");
    if (count == 0) return 0;
    const int bs = 256;
    LS_LAUNCH(ls_grid_points_kernel, (unsigned)((count + bs - 1) / bs), bs, 0, stream, n, step, origin[0], origin[1], origin[2], begin, count, xyz);
    return ls_check_launch("grid_points");
}

int ls2fm_render_loss(const float* rgb, const float* gt, int64_t n_rays, const float* normals, int64_t n_samples, float w_rgb,
                      float w_eik, float* sums, float* g_rgb, float* g_normals, void* stream) {
    if (n_rays < 0 || n_samples < 0 || !sums) return ls_fail("render_loss: bad arguments");
    if ((n_rays > 0 && (!rgb || !gt)) || (n_samples > 0 && !normals)) return ls_fail("render_loss: NULL input");
    ls_memset_async(sums, 0, 2 * sizeof(float), stream);
    if (n_rays == 0 && n_samples == 0) return 0;
    const int bs = 256;
    int64_t work = n_samples > 3 * n_rays ? n_samples : 3 * n_rays;
    int64_t grid = (work + bs - 1) / bs;
    if (grid > 4 * ls_sm_count()) grid = 4 * ls_sm_count();
    LS_LAUNCH(ls_render_loss_kernel, (unsigned)grid, bs, 0, stream, rgb, gt, 3 * n_rays, normals, n_samples, w_rgb, w_eik, sums, g_rgb,
              g_normals);
    return ls_check_launch("render_loss");
}

int ls2fm_render_tail(const float* rgb, const float* gt, const float* depth_mlp, const float* d_points, const uint8_t* mask_finish_in,
                      const float* normals, int64_t n_rays, int32_t n_per_ray, int32_t eik_masked, float w_rgb, float w_eik, float w_dc,
                      float* sums, uint8_t* mask_bg, uint8_t* mask_finish, float* g_rgb, float* g_depth, float* g_dpoints, float* g_normals,
                      void* stream) {
    if (n_rays < 0 || n_per_ray < 0 || !sums) return ls_fail("render_tail: bad arguments");
    if (n_rays > 0 && (!rgb || !gt || !mask_bg || !mask_finish)) return ls_fail("render_tail: rgb / gt / mask outputs are required");
    if ((depth_mlp == nullptr) != (d_points == nullptr)) return ls_fail("render_tail: depth_mlp and d_points go together");
    if (normals && n_per_ray < 1) return ls_fail("render_tail: n_per_ray must be >= 1 with normals");
    ls_memset_async(sums, 0, 8 * sizeof(float), stream);
    if (n_rays == 0) return 0;
    const int bs = 256;
    int64_t grid = (n_rays + bs - 1) / bs;
    if (grid > 4 * ls_sm_count()) grid = 4 * ls_sm_count();
    LS_LAUNCH(ls_render_tail_rays_kernel, (unsigned)grid, bs, 0, stream, rgb, gt, depth_mlp, d_points, mask_finish_in, n_rays, w_rgb, sums,
              mask_bg, mask_finish, g_rgb);
    if (ls_check_launch("render_tail(rays)")) return 1;
    const int64_t work = normals ? n_rays * n_per_ray : n_rays;
    grid = (work + bs - 1) / bs;
    if (grid > 8 * ls_sm_count()) grid = 8 * ls_sm_count();
    LS_LAUNCH(ls_render_tail_grads_kernel, (unsigned)grid, bs, 0, stream, normals, n_rays, n_per_ray, eik_masked, depth_mlp, d_points, mask_bg,
              mask_finish, w_eik, w_dc, sums, g_normals, g_depth, g_dpoints);
    return ls_check_launch("render_tail(grads)");
}

int ls2fm_se3_to_SE3(const float* wu, int64_t n, float* Rt, void* stream) {
    if (n < 0 || (n > 0 && (!wu || !Rt))) return ls_fail("se3_to_SE3: bad arguments");
    if (n == 0) return 0;
    const int bs = 128;
    LS_LAUNCH(ls_se3_to_SE3_kernel, (unsigned)((n + bs - 1) / bs), bs, 0, stream, wu, n, Rt);
    return ls_check_launch("se3_to_SE3");
}

int ls2fm_se3_to_SE3_backward(const float* wu, int64_t n, const float* g_Rt, float* d_wu, void* stream) {
    if (n < 0 || (n > 0 && (!wu || !g_Rt || !d_wu))) return ls_fail("se3_to_SE3_backward: bad arguments");
    if (n == 0) return 0;
    const int bs = 128;
    LS_LAUNCH(ls_se3_to_SE3_backward_kernel, (unsigned)((n + bs - 1) / bs), bs, 0, stream, wu, n, g_Rt, d_wu);
    return ls_check_launch("se3_to_SE3_backward");
}

int ls2fm_reproj_loss(const float* xyz, const float* Rt, const float* K, const float* kypts, const float* sdf, int64_t n, float sdf_band,
                      float eps, float* sums, float* uv, uint8_t* mask_surf, float* g_xyz, float* g_Rt, void* stream) {
    if (n < 0 || !sums) return ls_fail("reproj_loss: bad arguments");
    if (n > 0 && (!xyz || !Rt || !K || !kypts || !sdf)) return ls_fail("reproj_loss: NULL input");
    ls_memset_async(sums, 0, 4 * sizeof(float), stream);
    if (n == 0) return 0;
    const int bs = 128;
    int64_t grid = (n + bs - 1) / bs;
    if (grid > 4 * ls_sm_count()) grid = 4 * ls_sm_count();
    LS_LAUNCH(ls_reproj_sums_kernel, (unsigned)grid, bs, 0, stream, xyz, Rt, K, kypts, sdf, n, sdf_band, eps, sums, uv, mask_surf);
    if (ls_check_launch("reproj_loss(sums)")) return 1;
    if (g_xyz || g_Rt) {
        LS_LAUNCH(ls_reproj_grads_kernel, (unsigned)grid, bs, 0, stream, xyz, Rt, K, kypts, sdf, n, sdf_band, eps, sums, g_xyz, g_Rt);
        return ls_check_launch("reproj_loss(grads)");
    }
    return 0;
}

int ls2fm_generate_rays(const float* pose, const float* kinv, const float* xy, int32_t n_cams, int64_t n_pix, float* center, float* ray,
                        void* stream) {
    if (n_cams < 0 || n_pix < 0) return ls_fail("generate_rays: bad sizes");
    if (n_cams * n_pix > 0 && (!pose || !kinv || !xy || !center || !ray)) return ls_fail("generate_rays: NULL argument");
    if ((int64_t)n_cams * n_pix == 0) return 0;
    const int bs = 256;
    LS_LAUNCH(ls_generate_rays_kernel, (unsigned)(((int64_t)n_cams * n_pix + bs - 1) / bs), bs, 0, stream, pose, kinv, xy, n_cams, n_pix,
              center, ray);
    return ls_check_launch("generate_rays");
}

int ls2fm_generate_rays_backward(const float* pose, const float* kinv, const float* xy, int32_t n_cams, int64_t n_pix,
                                 const float* g_center, const float* g_ray, float* d_pose, void* stream) {
    if (n_cams < 0 || n_pix < 0) return ls_fail("generate_rays_backward: bad sizes");
    if (n_cams * n_pix > 0 && (!pose || !kinv || !xy || !d_pose)) return ls_fail("generate_rays_backward: NULL argument");
    if ((int64_t)n_cams * n_pix == 0) return 0;
    const int bs = 256;
    int64_t gx = (n_pix + bs - 1) / bs;
    if (gx > 64) gx = 64;
    LS_LAUNCH(ls_generate_rays_backward_kernel, dim3((unsigned)gx, (unsigned)n_cams), bs, 0, stream, pose, kinv, xy, n_cams, n_pix, g_center,
              g_ray, d_pose);
    return ls_check_launch("generate_rays_backward");
}

}  // extern "C"
