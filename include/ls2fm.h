/* ls2fm.h -- C ABI of libls2fm_sm100.so
 *
 * B200-native (sm_100a) replacement for the two native dependencies the Level-S2fM
 * hot path sits on -- tiny-cuda-nn's hash-grid encoding (reference call sites
 * models/base.py:17,37) and vren's ray/AABB test (utils/custom_functions.py:31) --
 * plus the fused per-sample / per-ray kernels that replace the eager PyTorch graph of
 * models/Renderer.py:51-116, models/SDF.py:55-226 and models/RadF.py:66-86.
 *
 * Conventions:
 *  - plain C, no torch types; every pointer marked "device" is a CUDA device pointer
 *    owned by the caller (PyTorch); the library never allocates or frees device memory
 *    and keeps no global state besides a thread-local error string;
 *  - every entry point is asynchronous on the given cudaStream_t (passed as void*),
 *    performs no host synchronisation and is re-entrant across streams;
 *  - return value 0 = success, non-zero = error (message via ls2fm_last_error());
 *  - all arrays are contiguous, row-major, fp32 unless stated; "nullable" pointers may
 *    be NULL to skip that output / input.
 */
#ifndef LS2FM_H_
#define LS2FM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LS2FM_ABI_VERSION 2
#define LS2FM_MAX_LEVELS 16
#define LS2FM_MAX_LAYERS 4   /* linear layers of the geometry MLP: 1..3 hidden (width 64) + output */
#define LS2FM_HIDDEN 64
#define LS2FM_MAX_OUT 20     /* k_geo + 1 <= 20 */
#define LS2FM_MAX_RAD_IN 68  /* 3 + 3 + 27 + 16 (+16 dual) rounded up to a multiple of 4 */

/* One level of the multiresolution hash grid (tiny-cuda-nn GridEncoding). */
typedef struct {
    float scale;          /* exp2f(l*log2f(b))*N_min - 1 */
    uint32_t resolution;  /* ceilf(scale)+1 */
    uint32_t offset;      /* first entry of this level in the table (entries, not floats) */
    uint32_t size;        /* entries in this level (hashmap_size) */
    uint32_t hashed;      /* 1: prime-xor hash, 0: dense strides */
} ls2fm_level_t;

/* Hash-grid configuration as the reference builds it (models/base.py:130-139). */
typedef struct {
    int32_t n_levels;
    int32_t n_features;        /* must be 2 */
    int32_t log2_hashmap_size;
    int32_t base_resolution;
    float per_level_scale;
} ls2fm_grid_cfg_t;

/* A hash-grid + geometry-MLP field: SDF (models/SDF.py:41-53) or RadF's Geo_enc
 * (models/RadF.py:28-45).
 * theta packs the EFFECTIVE (weight-normalised) MLP parameters, layer after layer:
 *   for l in 0..n_layers-1:  Wt_l [dims[l]][dims[l+1]]  (= W_l transposed, row-major),  b_l [dims[l+1]]
 * dims[0] = 3 + 2*n_levels; hidden dims must be 64; dims[n_layers] = k_geo+1 <= 20. */
typedef struct {
    const float* table;   /* device, [n_entries*2] -- embed_fn.embedder_obj.params */
    const float* theta;   /* device, packed as above */
    int32_t n_levels;
    ls2fm_level_t levels[LS2FM_MAX_LEVELS];
    float bound_min[3];
    float bound_max[3];
    float rescale;        /* opt.SDF.VolSDF.rescale: enc[0:3] = x / rescale */
    int32_t n_layers;
    int32_t dims[LS2FM_MAX_LAYERS + 1];
    float softplus_beta;       /* 100 */
    float softplus_threshold;  /* 20 */
    float sdf_sign;       /* +1 when opt.data.inside else -1:  sdf = sdf_sign * (y0 / scale_mlp) */
    float scale_mlp;      /* opt.SDF.NN_Init.scale_mlp */
    const float* tc_image; /* device, nullable: tensor-core operand image of theta (+ radiance) built by ls2fm_field_prepare;
                              when NULL every forward launch converts the weights itself */
} ls2fm_field_t;

/* Where the sample points come from. */
typedef struct {
    const float* xyz;     /* device [n,3] explicit points, or NULL for ray mode */
    const float* center;  /* device [n_rays,3]  (ray mode) */
    const float* ray;     /* device [n_rays,3]  un-normalised direction (ray mode) */
    const float* t;       /* device [n_rays,n_per_ray] depths; x = center + ray*t (mul then add, as torch) */
    const int32_t* ray_index; /* device [n_rays] nullable: compacted list of ray ids to process (ray mode) */
    const int32_t* n_active;  /* device scalar nullable: number of rays to process (entries of ray_index), read on the device */
    int64_t n;            /* explicit mode: number of points; ray mode: n_rays * n_per_ray */
    int32_t n_rays;
    int32_t n_per_ray;
    int32_t t_stride;     /* row stride of t (>= n_per_ray) */
    int32_t t_offset;     /* first column of t to use */
    int32_t out_stride;   /* ray mode, 0 = dense: per-sample outputs of sample j of ray r go to index r*out_stride + out_offset + j */
    int32_t out_offset;
} ls2fm_points_t;

/* The radiance decoder (models/base.py:221-261).  The reference applies NO hidden
 * activation (len(self.mlp) is taken on an empty ModuleList, base.py:230,257), so the
 * decoder is affine o sigmoid; the host composes W_eff = W3 W2 W1, b_eff accordingly
 * and the kernel evaluates rgb = sigmoid(W_eff . in + b_eff),
 * in = [x(3), normal(3), fourier(ray)(3+6*n_freq), geo(k_geo) (, geo2(k_geo))]  (Renderer.py:75). */
typedef struct {
    const float* w_eff;   /* device [3][in_dim] */
    const float* b_eff;   /* device [3] */
    int32_t in_dim;
    int32_t n_freq;       /* 4 */
    int32_t k_geo;        /* 16 */
    int32_t k_geo2;       /* 0 or 16 (dual_field) */
    const float* geo2;    /* device [n, k_geo2+1] nullable: RadF.Geometry_feat output (column 0 dropped) */
} ls2fm_radiance_t;

/* ------------------------------------------------------------------ host helpers */
int ls2fm_abi_version(void);
const char* ls2fm_last_error(void);
/* fills levels[0..n_levels) and *n_entries; shared table for oracle and kernels */
int ls2fm_grid_meta(const ls2fm_grid_cfg_t* cfg, ls2fm_level_t* levels, uint32_t* n_entries);
/* dynamic shared memory (bytes) the fused kernels need for this field; 0 if unsupported */
int ls2fm_smem_bytes(const ls2fm_field_t* field, int backward, int with_radiance);

/* ------------------------------------------------------------------ vren replacement
 * replaces vren.ray_aabb_intersect (utils/custom_functions.py:31) for N_voxels = 1, max_hits = 1.
 * hits_t [m,2] = (max(t1,0), t2) or (-1,-1); hit_cnt [m] int32 nullable. */
int ls2fm_ray_aabb(const float* rays_o, const float* rays_d, int64_t m,
                   const float center[3], const float half_size[3],
                   float* hits_t, int32_t* hit_cnt, void* stream);

/* The backward the reference's RayAABBIntersector lacks (utils/custom_functions.py:10-31; SURVEY 8a defect iii): on the slab axis that
 * decides t_near / t_far, dt/do_k = -1/d_k and dt/dd_k = -t/d_k; zero where t_near was clamped to 0 or the ray misses.
 * n_samples > 0: g [R, n_samples] is the gradient on the uniform depths of ls2fm_sample_uniform; n_samples == 0: g [R,2] is the
 * gradient on hits_t itself.  hits_t [R,2] as returned by the forward.  d_center / d_ray [R,3] are written. */
int ls2fm_ray_aabb_backward(const float* center, const float* ray, int32_t n_rays, int32_t n_samples, const float bound_min[3],
                            const float bound_max[3], const float* hits_t, const float* g, float* d_center, float* d_ray, void* stream);

/* ------------------------------------------------------------------ tcnn replacement (unfused)
 * replaces tcnn.Encoding.forward (models/base.py:37): u [m,3] -> enc [m, 2*n_levels].
 * idx (nullable) [m, n_levels, 8] uint32 receives the table entry index of every corner
 * (offset included) -- the bit-exact parity hook. */
int ls2fm_grid_encode(const ls2fm_field_t* field, const float* u, int64_t m,
                      float* enc, uint32_t* idx, void* stream);
/* backward of the above: d_table += scatter(g_enc), d_u (nullable) [m,3] */
int ls2fm_grid_encode_backward(const ls2fm_field_t* field, const float* u, int64_t m,
                               const float* g_enc, float* d_table, float* d_u, void* stream);

/* second-order pieces (tcnn's encoding is double-differentiable and the reference uses it: SDF.gradient, models/SDF.py:102-114,
 * differentiates through the encoding's input gradient with create_graph=True).  v [m,3] = the direction arriving at d_u:
 *   t_enc [m, 2L] (written)  = J_enc(u) v               -- gradient of d_u w.r.t. g_enc
 *   d_table (+=)             = scatter((dw/du . v) g_enc) -- gradient of d_u w.r.t. the table
 *   d_u2 [m,3] (+=)          = mixed second derivatives of the trilinear weights along v, contracted with table . g_enc
 * each output nullable (d_table / d_u2 need g_enc). */
int ls2fm_grid_encode_tangent(const ls2fm_field_t* field, const float* u, int64_t m, const float* v, const float* g_enc,
                              float* t_enc, float* d_table, float* d_u2, void* stream);

/* ------------------------------------------------------------------ parameter preparation
 * One weight-normed linear layer as the reference stores it (nn.utils.weight_norm, models/base.py:200,241):
 * W = g * v / ||v||_row.  dg / dv / db are the backward outputs (written, same shapes); ignored by the forward. */
typedef struct {
    const float* g;       /* device [dout]      weight_g */
    const float* v;       /* device [dout, din] weight_v */
    const float* b;       /* device [dout]      bias */
    float* dg; float* dv; float* db;
    int32_t din, dout;
} ls2fm_param_layer_t;
/* forward: geometry MLP layers (n_geo <= 4, nullable) -> theta (packed W_l^T, b_l); radiance decoder (exactly 3 layers
 * in -> 64 -> 64 -> 3, nullable) -> w_eff [3, in] = W3 W2 W1 and b_eff [3] (the reference applies no hidden activation,
 * models/base.py:230,257).  One launch instead of ~30 eager kernels. */
int ls2fm_params_forward(const ls2fm_param_layer_t* geo, int32_t n_geo, const ls2fm_param_layer_t* rad,
                         float* theta, float* w_eff, float* b_eff, void* stream);
/* backward: d_theta / d_w_eff / d_b_eff (nullable) -> dg, dv, db of every layer: written, or added to when accumulate != 0
 * (the caller points dg / dv / db straight at the parameters' .grad buffers: no temporaries, no autograd accumulation kernels). */
int ls2fm_params_backward(const ls2fm_param_layer_t* geo, int32_t n_geo, const ls2fm_param_layer_t* rad,
                          const float* d_theta, const float* d_w_eff, const float* d_b_eff, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------ tensor-core operand image
 * The tcgen05 forward kernel wants the weights as hi/lo TF32 operands in its shared-memory layout (W_l and W_l^T, K-major
 * core matrices, biases, W_eff).  ls2fm_field_prepare converts theta (+ the radiance block, nullable) ONCE into `image`
 * (ls2fm_field_image_floats floats); kernels launched with field.tc_image = image copy it with 16-byte loads instead of
 * re-deriving it per launch (the sampler alone launches the field kernel once per up-sampling round). */
int64_t ls2fm_field_image_floats(const ls2fm_field_t* field, const ls2fm_radiance_t* rad);
int ls2fm_field_prepare(const ls2fm_field_t* field, const ls2fm_radiance_t* rad, float* image, void* stream);

/* ------------------------------------------------------------------ fused field evaluation
 * replaces SDF.infer_sdf (+ SDF.gradient) / RadF.Geometry_feat (+ RadF.infer_app when rad != NULL).
 *   out_y   [n, dout]  raw MLP output                                  (nullable)
 *   out_sdf [n]        sign * (y0 / scale_mlp)                         (nullable)
 *   out_nrm [n,3]      d(out_sdf)/dx  (analytic, reverse sweep)        (nullable)
 *   out_rgb [n,3]      radiance (needs rad != NULL and ray mode)       (nullable) */
int ls2fm_field_forward(const ls2fm_field_t* field, const ls2fm_points_t* pts,
                        const ls2fm_radiance_t* rad,
                        float* out_y, float* out_sdf, float* out_nrm, float* out_rgb, void* stream);
/* Same contract, MLP on the fp32 SIMT pipes instead of tcgen05 (3xTF32) tensor cores: the cross-check of the
 * tensor-core kernel and the path for networks whose weight operands do not fit in shared memory. */
int ls2fm_field_forward_simt(const ls2fm_field_t* field, const ls2fm_points_t* pts,
                             const ls2fm_radiance_t* rad,
                             float* out_y, float* out_sdf, float* out_nrm, float* out_rgb, void* stream);

/* EXPERIMENTAL (round-2 groundwork; emulator-validated, one hardware run: bit-identical to the default kernel but 12 % slower
 * with its four gather warps; called by nothing in the package unless ops.FORWARD_WS is set): the values-only evaluation (out_y / out_sdf) by a warp-specialised kernel --
 * four gather warps write the next tile's encoding into a shared-memory MMA operand while sixteen MLP warps run the tensor-core
 * chain of the current one.  Same results as ls2fm_field_forward(field, pts, NULL, out_y, out_sdf, NULL, NULL). */
int ls2fm_field_forward_ws(const ls2fm_field_t* field, const ls2fm_points_t* pts, float* out_y, float* out_sdf, void* stream);

/* Gradients w.r.t. the sample POSITIONS (all nullable; pass NULL for the whole struct when none is wanted).  tiny-cuda-nn
 * propagates input gradients and the reference consumes them: SDF.gradient (models/SDF.py:102-114, create_graph=True) is
 * differentiated again w.r.t. p, and BA feeds get_surface_pts' output back into infer_sdf (pipelines/BA.py:123-125), so
 * d loss / d x flows into the first evaluation's normals.  With the upstream gradients of ls2fm_field_backward,
 *   dL/dx = Je^T ebar  +  (d(Je nbar)/dx)^T ebar_dot  +  W_eff[:,0:3]^T pbar
 * (Je = d enc / dx; ebar / ebar_dot the adjoints of the encoding in the primal / tangent channel; the middle term is the
 * hash grid's mixed second derivative; the last one the radiance decoder's direct dependence on x).
 *   explicit mode: d_xyz [n,3] is WRITTEN.
 *   ray mode (x = center + ray * t): d_center [n_rays,3] += sum_j dL/dx,  d_ray [n_rays,3] += sum_j t_j dL/dx + the gradient
 *   through the Fourier embedding of the direction (radiance),  d_t [n_rays, n_per_ray] is WRITTEN (= ray . dL/dx).
 *   d_ray also accepts explicit points that carry rays (pts.xyz and pts.ray set): only the Fourier term then. */
typedef struct {
    float* d_xyz;
    float* d_center;
    float* d_ray;
    float* d_t;
    float* workspace;   /* device, nullable: ls2fm_field_backward_workspace_floats(n) floats of scratch.  With it, launches that qualify
                           for the tcgen05 backward kernel keep running on it: the kernel parks the per-sample encoding adjoints
                           (72 floats / sample) here and a light second kernel gathers the table once more and contracts them into the
                           position gradient.  Without it such launches run the fp32-SIMT kernel (2.6x slower at 4096 x 128 samples). */
} ls2fm_input_grads_t;
int64_t ls2fm_field_backward_workspace_floats(int64_t n_samples);

/* backward of ls2fm_field_forward.  Upstream gradients (all nullable): g_y [n,dout], g_sdf [n],
 * g_nrm [n,3], g_rgb [n,3].  saved_nrm/saved_rgb: forward outputs (required when rad != NULL).
 * Accumulates (+=, atomics) into d_table [n_entries*2], d_theta [len(theta)], d_w_eff [3*in_dim],
 * d_b_eff [3]; writes d_geo2 [n,k_geo2+1]; each nullable.  The second-order path (gradient of the
 * normals w.r.t. the parameters, SDF.gradient's create_graph=True) is included.  in_grads (nullable): gradients
 * w.r.t. the sample positions, see ls2fm_input_grads_t. */
int ls2fm_field_backward(const ls2fm_field_t* field, const ls2fm_points_t* pts,
                         const ls2fm_radiance_t* rad,
                         const float* g_y, const float* g_sdf, const float* g_nrm, const float* g_rgb,
                         const float* saved_nrm, const float* saved_rgb,
                         float* d_table, float* d_theta, float* d_w_eff, float* d_b_eff, float* d_geo2,
                         const ls2fm_input_grads_t* in_grads, void* stream);
/* ls2fm_field_backward dispatches: launches of >= 8192 samples (field.tc_image set, n_levels % 4 == 0; position gradients
 * additionally need in_grads->workspace; launches without a gradient on the normals use the single-channel variant, 128 samples per tile) run the tcgen05 kernel -- every matrix product of the 2-channel forward / reverse pass and every weight
 * gradient as 3xTF32 tensor-core batches (fp32-level), weights streamed from the operand image, weight gradients accumulated in
 * TMEM, the next tile's loads and the previous tile's scatter hidden under the batches.  _simt forces the fp32-SIMT kernel
 * (cross-check and fallback), _tc forces the tensor-core kernel and fails when its preconditions do not hold. */
int ls2fm_field_backward_simt(const ls2fm_field_t* field, const ls2fm_points_t* pts,
                              const ls2fm_radiance_t* rad,
                              const float* g_y, const float* g_sdf, const float* g_nrm, const float* g_rgb,
                              const float* saved_nrm, const float* saved_rgb,
                              float* d_table, float* d_theta, float* d_w_eff, float* d_b_eff, float* d_geo2,
                              const ls2fm_input_grads_t* in_grads, void* stream);
int ls2fm_field_backward_tc(const ls2fm_field_t* field, const ls2fm_points_t* pts,
                            const ls2fm_radiance_t* rad,
                            const float* g_y, const float* g_sdf, const float* g_nrm, const float* g_rgb,
                            const float* saved_nrm, const float* saved_rgb,
                            float* d_table, float* d_theta, float* d_w_eff, float* d_b_eff, float* d_geo2,
                            const ls2fm_input_grads_t* in_grads, void* stream);

/* ------------------------------------------------------------------ compositing
 * replaces SDF.sdf_to_sigma + Renderer.composite + the background tail of Renderer.forward
 * (models/Renderer.py:33-49, 80-107).  Per ray: ray [R,3], t [R,N], sdf [R,N], rgbs [R,N,3],
 * nrm [R,N,3]; beta_param = SDF.beta (device scalar), beta = exp(beta_param*beta_speed).
 * Outputs rgb [R,3], depth [R], normal [R,3], opacity [R]. */
int ls2fm_composite_forward(const float* ray, const float* t, const float* sdf, const float* rgbs,
                            const float* nrm, const float* beta_param, float beta_speed,
                            const float bgcolor[3], int32_t n_rays, int32_t n_samples,
                            float* rgb, float* depth, float* normal, float* opacity, void* stream);
/* upstream g_rgb [R,3], g_depth [R], g_normal [R,3] (nullable each) ->
 * d_sdf [R,N], d_rgbs [R,N,3], d_nrm [R,N,3] (written), d_beta_param (+=, scalar), d_ray [R,3] (+=, nullable: through the
 * interval lengths |ray| (t_{i+1} - t_i)), d_t [R,N] (written, nullable: depths enter the depth output and the intervals) */
int ls2fm_composite_backward(const float* ray, const float* t, const float* sdf, const float* rgbs,
                             const float* nrm, const float* beta_param, float beta_speed,
                             const float bgcolor[3], int32_t n_rays, int32_t n_samples,
                             const float* g_rgb, const float* g_depth, const float* g_normal,
                             float* d_sdf, float* d_rgbs, float* d_nrm, float* d_beta_param, float* d_ray,
                             float* d_t, void* stream);

/* ------------------------------------------------------------------ depth samplers
 * Renderer.sample_depth after the AABB test (models/Renderer.py:118-127,178-185):
 * t[r,i] = (i+0.5)/N * (t_far - t_near) + t_near.  hits_t [R,2] nullable output. */
int ls2fm_sample_uniform(const float* center, const float* ray, int32_t n_rays, int32_t n_samples,
                         const float bound_min[3], const float bound_max[3],
                         float* t, float* hits_t, void* stream);

/* Error-bounded (VolSDF) sampler, models/Renderer.py:186-328 with the SURVEY 8(a) a5 fixes.
 * The caller provides the device workspace (ls2fm_sampler_workspace_bytes).  Outputs: t [R, N+Nf] sorted,
 * beta_plus [R] (network beta for converged rays), iters [R] (round of convergence, -1 = gave up); the last two
 * nullable.  No host synchronisation: active rays are compacted on the device between rounds and the SDF of the
 * newly drawn samples is evaluated by the fused field kernel on the compacted list only. */
typedef struct {
    int32_t n_samples;        /* sample_intvs N */
    int32_t n_final;          /* final_sample_intvs Nf */
    int32_t max_upsample_iter;
    int32_t max_bisection_itr;
    float eps;
    float beta_speed;
} ls2fm_sampler_cfg_t;
int64_t ls2fm_sampler_workspace_bytes(const ls2fm_sampler_cfg_t* cfg, int32_t n_rays);
int ls2fm_sample_error_bounded(const ls2fm_field_t* sdf_field, const float* beta_param,
                               const ls2fm_sampler_cfg_t* cfg,
                               const float* center, const float* ray, int32_t n_rays,
                               void* workspace, float* t_out, float* beta_plus, float* iters,
                               void* stream);

/* ------------------------------------------------------------------ sphere tracing
 * The no-grad march of SDF.sphere_tracing (models/SDF.py:116-200) as one kernel: both fronts, thresholding, clamping to
 * t_far, crossing test; every ray marches independently with the fused field evaluation inside the loop.
 *   track        [m, iters_max, 3]  start-front point before step k
 *   n_unfinished [iters_max + 1]    unfinished start fronts at the top of iteration k (zeroed by the call);
 *                                   K = first k with n_unfinished[k] == 0 (else iters_max) is the reference's iteration count
 *   t_near, t_far [m];  acc_end [iters_max + 1, m]  end-front depth at the top of iteration k (use row K). */
int ls2fm_sphere_trace(const ls2fm_field_t* sdf_field, const float* ray0, const float* ray_dir, int64_t m,
                       float sdf_threshold, int32_t iters_max, float* track, int32_t* n_unfinished,
                       float* t_near, float* t_far, float* acc_end, void* stream);

/* ------------------------------------------------------------------ rendering-loss tail
 * The rendering losses every stage computes on the renderer's outputs (pipelines/rendering_refine.py:99-121,
 * BA.py:190-204, Initialization.py:251-261): sums[0] = sum |rgb - gt| over [R,3], sums[1] = sum | ||n|| - 1 | over the
 * per-sample normals [S,3], and -- in the same pass -- the gradients of w_rgb * mean_L1 + w_eik * mean_eikonal
 * w.r.t. rgb (g_rgb [R,3]) and the normals (g_normals [S,3]); both nullable. */
int ls2fm_render_loss(const float* rgb, const float* gt, int64_t n_rays, const float* normals, int64_t n_samples,
                      float w_rgb, float w_eik, float* sums, float* g_rgb, float* g_normals, void* stream);

/* The whole loss tail of CameraSet.render + the stage's compute_loss (pipelines/Camera.py:506-537, BA.py:190-204,
 * rendering_refine.py:99-107) in two launches:
 *   mask_bg [R] u8 (written) = 0.05 < mean(gt) < 0.95;  mask_finish [R] u8 (written) = mask_finish_in & mask_bg (0 when mask_finish_in is NULL)
 *   sums [8] (zeroed by the call): [0] sum |rgb - gt| over R*3, [1] sum | ||n|| - 1 | over the counted samples, [2] #mask_bg rays,
 *        [3] sum (rgb - gt)^2 over the mask_bg rays (PSNR = -10 log10(sums[3] / (3 sums[2]))), [4] sum smooth_l1(d_points - depth_mlp)
 *        over the mask_finish rays, [5] #mask_finish rays
 *   eik_masked = 1: the eikonal term runs over the samples of mask_bg rays only (BA.py:192-193); 0: over all samples
 *   gradients of  w_rgb * mean_L1 + w_eik * mean_eikonal + w_dc * mean_smooth_l1  (each mean over its own count; the DC term is 0
 *   when no ray is finished): g_rgb [R,3], g_depth [R], g_dpoints [R], g_normals [R*n_per_ray,3]; all nullable.
 * depth_mlp / d_points (both or neither) and normals are nullable inputs. */
int ls2fm_render_tail(const float* rgb, const float* gt, const float* depth_mlp, const float* d_points, const uint8_t* mask_finish_in,
                      const float* normals, int64_t n_rays, int32_t n_per_ray, int32_t eik_masked, float w_rgb, float w_eik, float w_dc,
                      float* sums, uint8_t* mask_bg, uint8_t* mask_finish, float* g_rgb, float* g_depth, float* g_dpoints, float* g_normals,
                      void* stream);

/* ------------------------------------------------------------------ marching-cubes query grid
 * The N^3 query points of utils/util.py:392-411 (extract_mesh) for flat indices [begin, begin + count), written to
 * xyz [count,3]: same float64 arithmetic as the reference's numpy code (true division included), cast to float32.
 * step = volume_size / (N - 1); origin = (voxel_grid_origin[2], [1], [0]) -- the reference's own order.  Feed the chunk
 * to ls2fm_field_forward (values only) -- replaces 8192 host-side chunks + H2D/D2H copies of a 512^3 grid. */
int ls2fm_grid_points(int32_t n, double step, const double origin[3], int64_t begin, int64_t count, float* xyz, void* stream);

/* ------------------------------------------------------------------ ray generation
 * utils/camera.py:230-252 get_center_and_ray: pose [B,3,4] = [R|t] (world -> camera), kinv [B,3,3] = K^-1, pixel centres
 * xy [N,2] shared by the B cameras -> center [B,N,3] = -R^T t, ray [B,N,3] = R^T K^-1 [x,y,1] (un-normalised).
 * The backward accumulates (+=) the pose gradient d_pose [B,3,4] from g_center / g_ray (nullable each). */
int ls2fm_generate_rays(const float* pose, const float* kinv, const float* xy, int32_t n_cams, int64_t n_pix,
                        float* center, float* ray, void* stream);
/* utils/camera.py:85-96,119-142 Lie.se3_to_SE3 with its 11-term Taylor coefficients: wu [n,6] = (w, u) -> Rt [n,3,4] = [R | V u];
 * one launch instead of ~150 eager ones (BA evaluates it once per tracked POINT per iteration, pipelines/BA.py:127).
 * backward: d_wu [n,6] (written) = J^T g_Rt. */
int ls2fm_se3_to_SE3(const float* wu, int64_t n, float* Rt, void* stream);
int ls2fm_se3_to_SE3_backward(const float* wu, int64_t n, const float* g_Rt, float* d_wu, void* stream);
int ls2fm_generate_rays_backward(const float* pose, const float* kinv, const float* xy, int32_t n_cams, int64_t n_pix,
                                 const float* g_center, const float* g_ray, float* d_pose, void* stream);

/* ------------------------------------------------------------------ BA "sfm" reprojection residual
 * pipelines/BA.py:126-141 for n tracked points, each with its own world->camera pose Rt [n,3,4] (BA.py:127) and shared
 * intrinsics K [3,3]: (a,b,c) = K (R x + t), uv = (a,b) / (c + eps), d = |uv - kypts|, rows kept where |sdf| < sdf_band and uv is
 * finite; loss = 0.5 mean(2 log(1 + d^2/4)) + 0.5 mean(d) over the kept rows.
 * sums [4] (zeroed by the call): sum 2 log(1 + d^2/4), sum d, kept rows, mask_surf rows.  uv [n,2], mask_surf [n] u8, and the
 * gradients of the loss g_xyz [n,3], g_Rt [n,3,4] are written when non-NULL.  Two launches, no host synchronisation. */
int ls2fm_reproj_loss(const float* xyz, const float* Rt, const float* K, const float* kypts, const float* sdf, int64_t n, float sdf_band,
                      float eps, float* sums, float* uv, uint8_t* mask_surf, float* g_xyz, float* g_Rt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LS2FM_H_ */
