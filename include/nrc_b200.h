/*
 * nrc_b200.h -- C ABI of the B200-native radiance-cache query path.
 *
 * This is the drop-in boundary.  The reference
 * (benattal/neural-radiance-caching) is pure JAX and has no FFI of its own;
 * each entry point below names the reference Python function whose *body* it
 * replaces (file:line relative to the reference root).  A JAX maintainer binds
 * them through an XLA-FFI custom call + jax.custom_vjp; this
 * repository binds them through ctypes on torch device memory
 * (neural_radiance_caching_b200/_lib.py).
 *
 * Conventions
 *   - Every pointer named d_* is a DEVICE pointer owned by the caller; nothing
 *     is allocated, freed or retained by the library.  `stream` is a
 *     cudaStream_t passed as void*; all work is enqueued asynchronously on it.
 *   - All arrays are dense row-major fp32 unless stated (int32 where stated).
 *   - Hash/grid tables stay in the reference's checkpoint layout:
 *       dense level  [N,N,N,F] indexed grid[ix][iy][iz][f]  (internal/grid_utils.py:703-711)
 *       hash  level  [T,F]                                    (internal/grid_utils.py:835-841)
 *   - Gradient outputs are ACCUMULATED INTO (caller zero-fills).
 *   - Return value: 0 on success, negative nrc_status_t otherwise.  No
 *     exceptions cross the ABI; numerics never trap (safe_* guards reproduced).
 *   - Entry points are re-entrant and thread-safe.
 *   - Randomness is never generated here: kernels take uniforms / Gumbel
 *     noise as input tensors (the JAX host keeps its threefry streams).
 */
#ifndef NRC_B200_H_
#define NRC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRC_MAX_LEVELS 16
#define NRC_ABI_VERSION 1

typedef enum {
  NRC_OK = 0,
  NRC_E_INVALID_ARG = -1,   /* bad shape / null pointer / unsupported F, L, width  */
  NRC_E_UNSUPPORTED = -2,   /* valid request outside the compiled configurations   */
  NRC_E_CUDA = -3           /* launch failure; see nrc_last_cuda_error()           */
} nrc_status_t;

/* One resolution level of a HashEncoding (internal/grid_utils.py:830-852). */
typedef struct {
  const float* d_table;   /* [N,N,N,F] (dense) or [T,F] (hash), fp32            */
  float* d_grad;          /* same shape, accumulated into by *_bwd; may be NULL  */
  int32_t grid_size;      /* N                                                   */
  int32_t is_hash;        /* 0: dense grid (N^3 <= T), 1: hash table             */
  uint32_t table_size;    /* T for hash levels, N^3 for dense levels             */
  uint32_t reserved;
} nrc_level_t;

/* A HashEncoding instance (internal/grid_utils.py:738-905). */
typedef struct {
  int32_t num_levels;           /* L, 1..NRC_MAX_LEVELS                           */
  int32_t num_features;         /* F in {1,2,4,8}                                 */
  float bbox_min[3];            /* bbox[0]                        (:800-805)      */
  float bbox_max[3];            /* bbox[1]                                        */
  float bbox_span[3];           /* float32(bbox[1]-bbox[0])       (:820)          */
  float precondition_scaling;   /* 10.0                            (:903)         */
  nrc_level_t levels[NRC_MAX_LEVELS];
} nrc_encoding_t;

/* Fused density MLP parameters (internal/geometry.py:123-168): Dense kernels are
 * Flax layout [in,out] fp32, biases [out].  hidden = 64, depth = 2. */
typedef struct {
  const float* d_w0; const float* d_b0;   /* [in_dim,64], [64]                     */
  const float* d_w1; const float* d_b1;   /* [64,64], [64]                         */
  const float* d_wd; const float* d_bd;   /* output_density_layer [64,1], [1]      */
  const float* d_wn; const float* d_bn;   /* pred_normals_layer [64,3], [3] or NULL*/
  int32_t in_dim;                          /* L*F (6, 7 or 32 in the configs)      */
  int32_t width;                           /* 64                                   */
} nrc_density_mlp_t;

typedef struct {
  float* d_w0; float* d_b0; float* d_w1; float* d_b1;
  float* d_wd; float* d_bd; float* d_wn; float* d_bn;
} nrc_density_mlp_grad_t;

/* ---------------------------------------------------------------- misc ---- */
int32_t nrc_abi_version(void);
/* SHA-256 of the sources this library was compiled from (neural_radiance_caching_b200/build.py): the loader refuses a
 * library whose digest differs from the sources beside it. */
const char* nrc_build_digest(void);
const char* nrc_error_string(int32_t status);
/* cudaError_t of the most recent failed launch on the calling thread. */
int32_t nrc_last_cuda_error(void);

/* ---------------------------------------------------- K1/K2: encoding ---- */
/* Replaces HashEncoding.__call__ (internal/grid_utils.py:807-905) with
 * x_scale=None, per_level_fn=mean over a size-1 multisample axis,
 * feature_aggregator='concatenate': trilerp (:679-726) over
 * jax_hash_resample_3d (:41-121) / jax_resample_3d (:352-445).
 *   d_x   [P,3]   coordinates in the encoding's bbox frame (already warped)
 *   d_out [P,L*F] concatenated features * precondition_scaling
 */
int32_t nrc_encode_fwd(void* stream, const nrc_encoding_t* enc, const float* d_x,
                       int64_t num_points, float* d_out);

/* Parity aid: the integer half of the same path for one level.  d_idx [P,8]
 * int32 in the reference's corner order: hash levels -> table row index
 * (grid_utils.py:101-111); dense levels -> flat index into the zero-padded
 * (N+2)^3 volume, ((ix*(N+2))+iy)*(N+2)+iz  (grid_utils.py:384-438). */
int32_t nrc_encode_indices(void* stream, const nrc_encoding_t* enc, int32_t level,
                           const float* d_x, int64_t num_points, int32_t* d_idx);

/* VJP of nrc_encode_fwd (XLA's transpose of the gathers).
 *   d_g_out [P,L*F]  upstream gradient
 *   levels[l].d_grad += scatter-add of g*w          (skipped when NULL)
 *   d_g_x   [P,3]    written (not accumulated) when non-NULL: dL/dx
 */
int32_t nrc_encode_bwd(void* stream, const nrc_encoding_t* enc, const float* d_x,
                       const float* d_g_out, int64_t num_points, float* d_g_x);
/* The table scatter of a fused point query's VJP (nrc_density_query_fwd saved d_enc_out, nrc_density_mlp_bwd produced
 * d_g_out): x = contract(d_means / warp_c) is recomputed in the kernel (warp_c <= 0: identity), so no
 * nrc_contract_fwd launch or [P,3] buffer sits in front of it.  levels[l].d_grad += scatter-add of g*w. */
int32_t nrc_encode_bwd_warped(void* stream, const nrc_encoding_t* enc, const float* d_means, float warp_c,
                              const float* d_g_out, int64_t num_points);

/* -------------------------------------------------- contraction (row 7) ---- */
/* coord.contract(x / c) (internal/coord.py:33-38,63-69); c <= 0 means identity. */
int32_t nrc_contract_fwd(void* stream, const float* d_x, int64_t num_points, float c, float* d_z);
/* d_g_x[P,3] = J^T d_g_z. */
int32_t nrc_contract_bwd(void* stream, const float* d_x, const float* d_g_z, int64_t num_points,
                         float c, float* d_g_x);

/* ------------------------------------------------ K3: density MLP (row 8) -- */
/* BaseDensityMLP.run_network (internal/geometry.py:155-168) for depth 2/width 64,
 * plus pred_normals_layer (:467) when mlp->d_wn != NULL.
 *   d_enc [P,in_dim] -> d_raw [P], d_feat [P,64] (may be NULL), d_grad_pred [P,3] (may be NULL)
 * bf16 != 0 selects the tensor-core variant (bf16 operands, fp32 accumulate). */
int32_t nrc_density_mlp_fwd(void* stream, const nrc_density_mlp_t* mlp, const float* d_enc,
                            int64_t num_points, int32_t bf16, float* d_raw, float* d_feat,
                            float* d_grad_pred);
/* VJP.  Upstream grads d_g_raw [P], d_g_feat [P,64] or NULL, d_g_grad_pred [P,3] or NULL.
 * d_density [P] (may be NULL): when given, d_g_raw is the gradient w.r.t. the *activated*
 * density of nrc_density_query_fwd and is multiplied by d_density here -- the VJP of
 * safe_exp (y * g, internal/math.py:186-192) and of the bbox mask (density == 0 outside).
 * Outputs: d_g_enc [P,in_dim] (written, may be NULL); weight grads accumulated into `grads`
 * (may be NULL to skip). */
int32_t nrc_density_mlp_bwd(void* stream, const nrc_density_mlp_t* mlp, const float* d_enc,
                            const float* d_g_raw, const float* d_density, const float* d_g_feat,
                            const float* d_g_grad_pred, int64_t num_points, int32_t bf16, float* d_g_enc,
                            const nrc_density_mlp_grad_t* grads);

/* Fused point query: means -> contract(x/c) -> encode -> MLP ->
 * density = safe_exp(raw + density_bias) masked to the bbox
 * (internal/geometry.py:199-341), optionally the analytic normals' raw gradient
 * d raw / d means (:442-460).
 *   d_means [P,3]; outputs (each may be NULL): d_density [P], d_raw [P],
 *   d_feat [P,64], d_grad_pred [P,3], d_raw_grad [P,3], d_enc_out [P,L*F] (the encoded
 *   features, saved for nrc_density_mlp_bwd / nrc_encode_bwd in training). */
int32_t nrc_density_query_fwd(void* stream, const nrc_encoding_t* enc, const nrc_density_mlp_t* mlp,
                              const float* d_means, int64_t num_points, float warp_c,
                              float density_bias, int32_t bf16, float* d_density, float* d_raw,
                              float* d_feat, float* d_grad_pred, float* d_raw_grad, float* d_enc_out);

/* Second-order path of the analytic normals.  The reference back-propagates losses through
 * normals = -l2_normalize(d raw / d means) (internal/geometry.py:442-460; predicted_normal_loss with
 * pred='normals', internal/loss_utils.py:169-199, internal/train_utils.py:1049-1070).  Given
 * d_g_raw_grad [P,3] = dL / d(d raw / d means), accumulates the parameter gradient of < g, d raw / d means >
 * (= of the forward-mode tangent of raw along g) into grads->d_w0 / d_w1 / d_wd (biases and the predicted-normal
 * head do not enter) and into the level gradient tables enc->levels[l].d_grad.  Sample positions are constants
 * (stop_level_grad, internal/sampling.py:353-354).  fp32 arithmetic. */
int32_t nrc_density_normals_bwd(void* stream, const nrc_encoding_t* enc, const nrc_density_mlp_t* mlp,
                                const float* d_means, const float* d_g_raw_grad, int64_t num_points, float warp_c,
                                const nrc_density_mlp_grad_t* grads);

/* The same second-order term on tensor cores (bf16-MLP variant), as three launches that reuse the first-order kernels'
 * structure.  With zdot = J_contract d_g and dw_c the derivative of corner c's trilinear weight along zdot:
 *   nrc_encode_tangent_fwd       d_edot [P, L*F]   = scale * sum_c dw_c T[c]                     (fp32)
 *   nrc_density_mlp_bwd_tangent  the MLP's backward pass on the TANGENT network (same ReLU masks as the primal, recomputed
 *                                from d_enc = the primal features saved by nrc_density_query_fwd; no biases; upstream 1):
 *                                grads->d_w0 += edot (x) a1, d_w1 += h1dot (x) a2, d_wd += h2dot (accumulated; biases and
 *                                the normal head untouched), d_g_enc [P, L*F] = W0 a1                (bf16 mma.sync)
 *   nrc_encode_tangent_bwd       levels[l].d_grad[c] += dw_c * scale * d_ge                         (fp32 atomics)
 * d_means, d_g [P,3] as for nrc_density_normals_bwd. */
int32_t nrc_encode_tangent_fwd(void* stream, const nrc_encoding_t* enc, const float* d_means, const float* d_g,
                               int64_t num_points, float warp_c, float* d_edot);
int32_t nrc_density_mlp_bwd_tangent(void* stream, const nrc_density_mlp_t* mlp, const float* d_enc, const float* d_enc_dot,
                                    int64_t num_points, float* d_g_enc, const nrc_density_mlp_grad_t* grads);
int32_t nrc_encode_tangent_bwd(void* stream, const nrc_encoding_t* enc, const float* d_means, const float* d_g,
                               const float* d_ge, int64_t num_points, float warp_c);

/* ------------------------------------------------------ K4: ray kernels ---- */
/* render.compute_alpha_weights (internal/render.py:134-169), delta=None.
 *   d_density [R,n], d_tdist [R,n+1], d_dirs [R,3] -> d_weights, d_alpha, d_trans [R,n]
 *   (d_alpha / d_trans may be NULL). */
int32_t nrc_ray_alpha_weights_fwd(void* stream, const float* d_density, const float* d_tdist,
                                  const float* d_dirs, int64_t num_rays, int32_t n,
                                  int32_t opaque_background, float* d_weights, float* d_alpha,
                                  float* d_trans);
/* VJP wrt density: d_g_density [R,n] written. Upstream d_g_weights (may be NULL),
 * d_g_alpha (may be NULL), d_g_trans (may be NULL). */
int32_t nrc_ray_alpha_weights_bwd(void* stream, const float* d_density, const float* d_tdist,
                                  const float* d_dirs, const float* d_g_weights,
                                  const float* d_g_alpha, const float* d_g_trans, int64_t num_rays,
                                  int32_t n, float* d_g_density);

/* stepfun.sample_intervals (internal/stepfun.py:207-250) with single_jitter=True,
 * preceded by the sampler's annealed logits (internal/sampling.py:340):
 *   logits = anneal * safe_log(weights + padding).
 *   d_t [R,m+1] current fenceposts, d_w [R,m] current weights, d_u01 [R] uniforms in [0,1),
 *   d_u_base [n] = linspace part of u (host computes it in fp32, see stepfun.sample_u_base)
 *   -> d_t_new [R,n+1] sorted, clipped to [dom_lo,dom_hi]; d_bin_idx [R,n] int32 (idx0 of
 *   math.sorted_lookup, internal/math.py:433-439; may be NULL). */
int32_t nrc_ray_sample_intervals(void* stream, const float* d_t, const float* d_w,
                                 const float* d_u01, const float* d_u_base, int64_t num_rays,
                                 int32_t m, int32_t n, float anneal, float padding,
                                 float max_jitter, float dom_lo, float dom_hi, float* d_t_new,
                                 int32_t* d_bin_idx);

/* s_to_t (internal/coord.py:223-260) + render.cast_rays means (internal/render.py:106-131,
 * cone, stable mip-NeRF mean).  warp_kind 0: linear t = s*far + (1-s)*near; 1: power ladder
 * (internal/math.py:295-341) with (p, premult).
 *   d_sdist [R,n+1] -> d_tdist [R,n+1], d_means [R,n,3]. */
int32_t nrc_ray_cast(void* stream, const float* d_sdist, const float* d_origins,
                     const float* d_directions, const float* d_near, const float* d_far,
                     int64_t num_rays, int32_t n, int32_t warp_kind, float p, float premult,
                     float* d_tdist, float* d_means);
/* The COVARIANCES of render.cast_rays (internal/render.py:26-131; `diag=False` at internal/sampling.py:361-368), on request:
 * the BASELINE configs never read them ('mean' unscented basis), so the sampler's launches do not compute them.
 *   ray_shape 0 'cone' (conical_frustum_to_gaussian :62-81), 1 'cylinder' (cylinder_to_gaussian :84-103);
 *   d_tdist [R,n+1] metric fenceposts, d_directions [R,3], d_radii [R] (base radius at distance 1 / cylinder radius);
 *   -> d_covs [R,n,3,3] (diag = 0) or [R,n,3] (diag = 1), may be NULL; d_means [R,n,3] (may be NULL; needs d_origins) -
 *      the cone means equal nrc_ray_cast's, the cylinder means (t0+t1)/2 exist only here. */
int32_t nrc_ray_cast_covs(void* stream, const float* d_tdist, const float* d_origins, const float* d_directions,
                          const float* d_radii, int64_t num_rays, int32_t n, int32_t ray_shape, int32_t diag,
                          float* d_covs, float* d_means);
/* nrc_ray_sample_intervals followed by nrc_ray_cast in ONE launch (the sampler's per-level pair sampling.py:340-349 ->
 * render.cast_rays, internal/render.py:26-131): same arguments and bit-identical outputs; d_sdist_new [R,n+1] are the
 * resampled normalised fenceposts, d_tdist [R,n+1] their metric distances, d_means [R,n,3] (may be NULL) the
 * Gaussian means. */
int32_t nrc_ray_sample_cast(void* stream, const float* d_t, const float* d_w, const float* d_u01, const float* d_u_base,
                            int64_t num_rays, int32_t m, int32_t n, float anneal, float padding, float max_jitter,
                            float dom_lo, float dom_hi, const float* d_origins, const float* d_directions,
                            const float* d_near, const float* d_far, int32_t warp_kind, float p, float premult,
                            float* d_sdist_new, float* d_tdist, float* d_means);
/* nrc_ray_sample_cast with compute_alpha_weights (internal/render.py:134-169) of the level being resampled folded into
 * its head: the weights of the step function (d_t [R,m+1]) are computed from that level's densities d_density [R,m], its
 * metric fenceposts d_tdist_prev [R,m+1] and the ray directions - bit-identical to nrc_ray_alpha_weights_fwd - used for
 * the resampling and written to d_weights_out [R,m] (may be NULL).  Densities and weights make no HBM round trip between
 * the density query and the resampling. */
int32_t nrc_ray_weights_sample_cast(void* stream, const float* d_t, const float* d_density, const float* d_tdist_prev,
                                    int32_t opaque_background, float* d_weights_out, const float* d_u01,
                                    const float* d_u_base, int64_t num_rays, int32_t m, int32_t n, float anneal,
                                    float padding, float max_jitter, float dom_lo, float dom_hi,
                                    const float* d_origins, const float* d_directions, const float* d_near,
                                    const float* d_far, int32_t warp_kind, float p, float premult,
                                    float* d_sdist_new, float* d_tdist, float* d_means);

/* render.volumetric_rendering (internal/render.py:172-247): acc, rgb (+bg), C channels
 * composited with `weights`, distance mean and percentiles (5,50,95) from
 * `weights_no_filter`.
 *   d_values [R,k,C] (rgb first when has_rgb), d_weights [R,k],
 *   d_weights_nf [R,n] (NULL: same as d_weights, k == n), d_tdist [R,n+1], d_bg [R,3] or NULL
 *   -> d_out [R,C], d_acc [R] (may be NULL), d_dist [R,4] = (mean, p5, median, p95; may be NULL). */
int32_t nrc_ray_composite_fwd(void* stream, const float* d_values, const float* d_weights, int32_t k,
                              const float* d_weights_nf, const float* d_tdist, const float* d_bg,
                              int64_t num_rays, int32_t n, int32_t channels, int32_t has_rgb,
                              float* d_out, float* d_acc, float* d_dist);
/* VJP wrt values, weights and weights_no_filter (rgb/extras/acc terms; the distance
 * statistics carry no gradient here).  Outputs are written; d_g_weights_nf NULL with
 * d_weights_nf NULL means weights == weights_no_filter and the acc term goes to d_g_weights. */
int32_t nrc_ray_composite_bwd(void* stream, const float* d_values, const float* d_weights, int32_t k,
                              const float* d_weights_nf, const float* d_bg, const float* d_g_out,
                              const float* d_g_acc, int64_t num_rays, int32_t n, int32_t channels,
                              int32_t has_rgb, float* d_g_values, float* d_g_weights,
                              float* d_g_weights_nf);

/* Model.maybe_resample (internal/models.py:193-292), resample_argmax=False:
 *   logits = safe_log(w + bias) * mult; inds = argmax(logits + gumbel) per draw;
 *   w_new = w[inds] / (num_resample * softmax(logits)[inds] + 1e-8).
 *   d_weights [R,n], d_gumbel [R,n,k] -> d_inds [R,k] int32, d_w_new [R,k]. */
int32_t nrc_ray_resample(void* stream, const float* d_weights, const float* d_gumbel,
                         int64_t num_rays, int32_t n, int32_t k, float bias, float mult,
                         int32_t* d_inds, float* d_w_new);
/* take_along_axis of a per-sample field: d_field [R,n,C] -> d_out [R,k,C]. */
int32_t nrc_ray_resample_gather(void* stream, const float* d_field, const int32_t* d_inds,
                                int64_t num_rays, int32_t n, int32_t k, int32_t channels,
                                float* d_out);

/* ------------------------------------------- cache shader building blocks ---- */
/* flax.linen.Dense (internal/geometry.py:127-139, nerf.py:232-345,
 * surface_light_field.py:352-403): y = act(x @ kernel + bias), kernel [in,out] fp32.
 * Row strides ldx/ldy (in floats) let a layer read/write a column slice of a wider
 * activation buffer, which is how the reference's concatenate skip connections
 * (surface_light_field.py:480-500) are realised without copies.  relu != 0 applies ReLU;
 * bf16 != 0 selects the tensor-core variant (bf16 operands, fp32 accumulate). */
int32_t nrc_dense_fwd(void* stream, const float* d_x, int64_t ldx, const float* d_kernel,
                      const float* d_bias, int64_t num_rows, int32_t in_dim, int32_t out_dim,
                      int32_t relu, int32_t bf16, float* d_y, int64_t ldy);
/* ReLU VJP: d_g_pre[r,c] = d_y[r,c] > 0 ? d_g_y[r,c] : 0 (all with row strides). */
int32_t nrc_relu_bwd(void* stream, const float* d_y, int64_t ldy, const float* d_g_y, int64_t ldgy,
                     int64_t num_rows, int32_t num_cols, float* d_g_pre, int64_t ldgp);
/* VJP of nrc_dense_fwd with respect to the PRE-activation output (apply nrc_relu_bwd first for
 * ReLU layers).
 *   d_g_x [rows,in] (stride ldgx) written, or accumulated into when accumulate_g_x != 0
 *   (second consumer of a skip connection); may be NULL.
 *   d_g_kernel [in,out] and d_g_bias [out] are ACCUMULATED INTO; may be NULL (both). */
int32_t nrc_dense_bwd(void* stream, const float* d_x, int64_t ldx, const float* d_kernel,
                      const float* d_g_y, int64_t ldgy, int64_t num_rows, int32_t in_dim,
                      int32_t out_dim, int32_t bf16, float* d_g_x, int64_t ldgx,
                      int32_t accumulate_g_x, float* d_g_kernel, float* d_g_bias);

/* Integrated directional encoding, ref_utils.generate_ide_fn (internal/ref_utils.py:131-192).
 * Host tables built like the reference: ml_m/ml_l [n_sh] from get_ml_array (:117-128),
 * sigma[i] = 0.5*l*(l+1), d_mat [(l_max+1), n_sh] fp32 (device) from sph_harm_coeff (:107-114).
 *   d_xyz [P,3], d_kappa_inv [P] -> d_out [P, 2*n_sh] (row stride ldo): real parts then imaginary. */
int32_t nrc_ide_fwd(void* stream, int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l,
                    const float* sigma, const float* d_mat, const float* d_xyz,
                    const float* d_kappa_inv, int64_t num_points, float* d_out, int64_t ldo);
/* VJP: d_g_xyz [P,3] and d_g_kappa_inv [P] written (either may be NULL). */
int32_t nrc_ide_bwd(void* stream, int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l,
                    const float* sigma, const float* d_mat, const float* d_xyz,
                    const float* d_kappa_inv, const float* d_g_out, int64_t ldg, int64_t num_points,
                    float* d_g_xyz, float* d_g_kappa_inv);

/* ------------------------------- fused MLP chains on tcgen05 tensor cores ---- */
/* The Dense stacks of the cache shader / material / light-field MLPs
 * (internal/nerf.py:232-345,561-689, internal/surface_light_field.py:352-403,480-500,
 * internal/material.py:2073-2123) executed per 128-point tile as one PROGRAM, activations
 * staying in shared memory (bf16, 128-byte-swizzled 128x64 "atoms") and accumulators in
 * tensor memory.  Used for the forward pass and, with transposed weights and ReLU masks, for
 * the data-gradient pass; nrc_chain_wgrad turns the saved tile images into weight gradients.
 *
 * Pointer table: every device pointer an op refers to is an index into `d_ptrs` (host array of
 * device pointers, copied at launch).  A "tile image" is a bf16 buffer
 * [ceil(rows/128)][img_atoms][16384 bytes] holding whole atoms in their shared-memory layout. */
#define NRC_CHAIN_MAX_OPS 24
#define NRC_CHAIN_MAX_PTRS 40
#define NRC_CHAIN_MAX_ATOMS 8
#define NRC_CHAIN_MAX_PROGRAMS 3
#define NRC_PACK_MAX_ENTRIES 80
#define NRC_WGRAD_MAX_LAYERS 12
#define NRC_WGRAD_MAX_X_ATOMS 6
#define NRC_WGRAD_MAX_SEGS 5

typedef enum { NRC_OP_LOAD = 0, NRC_OP_GEMM = 1, NRC_OP_EPI = 2, NRC_OP_SAVE = 3, NRC_OP_GATHER = 4, NRC_OP_LOADIMG = 5 } nrc_chain_op_kind_t;
#define NRC_GEMM_ACCUMULATE 1     /* accumulate onto the accumulator's current contents (skip VJP) */
#define NRC_EPI_RELU 1            /* max(0, .) after the bias                                      */
#define NRC_EPI_OUT_ACCUMULATE 2  /* fp32 output: += instead of =                                  */
#define NRC_EPI_DENSITY 4         /* density head (nrc_chain_query): see below                     */

typedef struct {
  int32_t kind;       /* nrc_chain_op_kind_t                                                       */
  int32_t slot;       /* LOAD/LOADIMG/EPI/SAVE: first atom slot (EPI: -1 = no bf16 result)         */
  int32_t ptr;        /* LOAD: fp32 source (-1 = zeros); EPI: bias [ncols] or -1; SAVE / LOADIMG:   */
                      /* tile image (LOADIMG = the inverse of SAVE: npad atoms starting at atom     */
                      /* col0 of the tile's image are bulk-copied into the slots; the producer of   */
                      /* the image - another chain's SAVE or a per-point kernel - wrote bf16 atoms, */
                      /* so no fp32 row is read or converted; CTA-pair kernel only)                 */
  int32_t ld;         /* LOAD: source row stride; EPI: fp32 output row stride (floats)             */
  int32_t col0;       /* LOAD: first destination column (x8); EPI: first output column;            */
                      /* SAVE: first atom inside the image                                         */
  int32_t ncols;      /* LOAD: source columns; EPI: valid result columns                           */
  int32_t npad;       /* LOAD: columns written, zero padded (x8); EPI: columns processed (x16);    */
                      /* SAVE: number of atoms                                                     */
  int32_t tmem_col;   /* GEMM/EPI: first accumulator column (0..255 per context; up to 511 when the */
                      /* program runs with one tile context per CTA: CTA-pair kernel only)          */
  int32_t n;          /* GEMM: N (x16, <= 128).  EPI with out_ptr: staging slot + 1 (0 = direct stores): the
                       * fp32 rows leave through that free slot as coalesced 16-byte stores              */
  int32_t flags;      /* NRC_GEMM_* / NRC_EPI_*                                                    */
  int32_t out_ptr;    /* EPI: fp32 output matrix or -1                                             */
  int32_t mask_ptr;   /* EPI: tile image of the forward activation; result zeroed where it is <= 0 */
  int32_t mask_atom0; /* EPI: first atom of that activation inside the image                       */
  int32_t img_atoms;  /* SAVE / EPI mask: atoms per tile of the image                              */
  int32_t w_chunk;    /* GEMM: first 16 KB chunk of the packed weight image (one per K atom)       */
  int32_t n_atoms;    /* GEMM: number of K atoms                                                   */
  uint8_t a_slot[NRC_CHAIN_MAX_ATOMS];  /* GEMM: slot holding each K atom                          */
  uint8_t a_klen[NRC_CHAIN_MAX_ATOMS];  /* GEMM: K extent used in each atom (x16, <= 64)           */
  float fparam;       /* EPI with NRC_EPI_DENSITY: density_bias                                   */
} nrc_chain_op_t;

typedef struct {
  int32_t num_ops;
  int32_t slots_per_ctx;   /* atom slots per tile context (2 contexts per CTA), 1..7 */
  nrc_chain_op_t ops[NRC_CHAIN_MAX_OPS];
} nrc_chain_program_t;

/* One rectangular piece of a Flax kernel [in,out] (row stride ld) written into a 16 KB chunk:
 *   transpose == 0: chunk[n0+n][k0+k] = W[row0+k][col0+n]  (forward: B = kernel^T, K = in)
 *   transpose == 1: chunk[n0+n][k0+k] = W[row0+n][col0+k]  (data gradient: B = kernel, K = out)
 * with n < (transpose ? nrows : ncols) and k < (transpose ? ncols : nrows). */
typedef struct {
  int32_t ptr, ld, row0, nrows, col0, ncols, chunk, n0, k0, transpose;
} nrc_pack_entry_t;

/* One Dense layer's weight gradient: kernel rows are gathered from up to 6 X atoms (skip
 * connections concatenate atoms of two images), columns are split into up to 5 segments so
 * that several reference layers sharing one input (shader heads) are one GEMM. */
typedef struct {
  int32_t n_x_atoms;
  int32_t x_ptr[NRC_WGRAD_MAX_X_ATOMS];        /* tile image holding the atom                  */
  int32_t x_img_atoms[NRC_WGRAD_MAX_X_ATOMS];  /* atoms per tile of that image                 */
  int32_t x_atom[NRC_WGRAD_MAX_X_ATOMS];       /* atom index inside the image                  */
  int32_t x_rows[NRC_WGRAD_MAX_X_ATOMS];       /* valid input features in the atom (<= 64)     */
  int32_t w_row0[NRC_WGRAD_MAX_X_ATOMS];       /* kernel row of the atom's first feature       */
  int32_t dy_ptr, dy_img_atoms, dy_atom0;      /* pre-activation gradient image                */
  int32_t n;                                   /* columns of dY processed (x16, <= 128)        */
  int32_t n_seg;
  int32_t seg_col0[NRC_WGRAD_MAX_SEGS], seg_ncols[NRC_WGRAD_MAX_SEGS];
  int32_t seg_w_ptr[NRC_WGRAD_MAX_SEGS];       /* fp32 [in, seg_ncols] gradient, accumulated   */
  int32_t seg_b_ptr[NRC_WGRAD_MAX_SEGS];       /* fp32 [seg_ncols] bias gradient or -1         */
} nrc_wgrad_layer_t;

/* Run `prog` over ceil(num_rows/128) tiles.  d_weights_packed: chunks from nrc_chain_pack_weights. */
int32_t nrc_chain_run(void* stream, const nrc_chain_program_t* prog, void* const* d_ptrs, int32_t num_ptrs,
                      const void* d_weights_packed, int64_t num_rows);
/* Several INDEPENDENT programs over the same rows in one launch (e.g. the integrated-BRDF, SurfaceLightField and EnvMap
 * stacks of internal/nerf.py:940-1090, which all read the bottleneck / reflection encoding of one shaded point and do
 * not read each other): their (program, tile) work items are dealt to the CTA pairs in cost-balanced contiguous
 * ranges.  progs / d_ptrs / num_ptrs / d_weights_packed are arrays of num_programs (<= NRC_CHAIN_MAX_PROGRAMS) entries
 * with the meaning of nrc_chain_run's arguments.  Equivalent to num_programs calls of nrc_chain_run. */
int32_t nrc_chain_run_multi(void* stream, int32_t num_programs, const nrc_chain_program_t* const* progs,
                            void* const* const* d_ptrs, const int32_t* num_ptrs, const void* const* d_weights_packed,
                            int64_t num_rows);
/* nrc_chain_run with a hash-grid front end: the fused point query (predict_density + convert_raw_density,
 * internal/geometry.py:199-341) on the tensor-core chain kernel.  Two extra op behaviours:
 *   GATHER  slot = destination atom, ptr = means [P,3], out_ptr = encoded features fp32 [P, L*F] or -1,
 *           ncols = L*F (<= 32), npad = ncols rounded up to 16: every tile row contracts its point
 *           (coord.contract(x / warp_c)), gathers `enc` (same arithmetic as nrc_encode_fwd) and writes the bf16
 *           feature row; the bbox mask of the point stays in a register of the row's thread.
 *   EPI with NRC_EPI_DENSITY: column 0 -> density = inside ? safe_exp(raw + fparam) : 0 to out_ptr [P];
 *           columns 1..3 -> predicted-normal gradient to mask_ptr [P,3] (-1: none); bias = ptr [4]. */
int32_t nrc_chain_query(void* stream, const nrc_chain_program_t* prog, void* const* d_ptrs, int32_t num_ptrs,
                        const void* d_weights_packed, int64_t num_rows, const nrc_encoding_t* enc, float warp_c);

/* (Re)build the packed bf16 weight image (num_chunks * 16 KB).  keep_existing == 0 zero-fills the
 * whole image first; further calls with keep_existing != 0 add more pieces to the same image. */
int32_t nrc_chain_pack_weights(void* stream, const nrc_pack_entry_t* entries, int32_t num_entries,
                               void* const* d_ptrs, int32_t num_ptrs, void* d_packed, int32_t num_chunks,
                               int32_t keep_existing);
/* dW += X^T dY, db += 1^T dY for `num_layers` layers from saved tile images. */
int32_t nrc_chain_wgrad(void* stream, const nrc_wgrad_layer_t* layers, int32_t num_layers,
                        void* const* d_ptrs, int32_t num_ptrs, int64_t num_rows);

/* --------------------------- cache shader: per-point stages between the stacks ---- */
/* internal/nerf.py:940-1090 (_predict_appearance_passive), :461-482, :1344-1358,
 * internal/ref_utils.py:25-42,131-192 with the activations of configs/nerf_ngp_yobo.gin:491-506.
 * IDE tables as in nrc_ide_fwd (degree-5 list; the first n_sh_env harmonics are the EnvMap's
 * degree-4 encoding).  d_heads [P,ldh]: column 0 roughness, 1-3 ambient irradiance, 4-6 irradiance,
 * 7-9 tint (raw Dense outputs).  d_viewdirs [P/samples_per_ray, 3].
 *   mid: roughness = softplus(raw + roughness_bias); dotprod = n.(-v); refdirs = reflect(-v, n);
 *        d_ide_slf [P,2*n_sh] = IDE(refdirs, roughness); d_ide_env [P,2*n_sh_env] (may be NULL). */
/* Optional bf16 outputs of the mid stage in the chains' operand layout (tile images, see above): IDE_5 -> the
 * SurfaceLightField stack's input atoms, IDE_4 -> the EnvMap stack's, n.v -> the integrated-BRDF stack's last atom
 * (column 0; 16 columns written).  With an image given the matching fp32 output may be NULL. */
typedef struct {
  void* slf_img; int32_t slf_img_atoms, slf_atom0;
  void* env_img; int32_t env_img_atoms, env_atom0;
  void* dot_img; int32_t dot_img_atoms, dot_atom0;
} nrc_shader_images_t;
int32_t nrc_shader_mid_fwd(void* stream, int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l,
                           const float* sigma, const float* d_mat, int32_t n_sh_env, const float* d_heads,
                           int64_t ldh, const float* d_normals, const float* d_viewdirs, int64_t num_points,
                           int32_t samples_per_ray, float roughness_bias, float* d_roughness,
                           float* d_dotprod, float* d_refdirs, float* d_ide_slf, float* d_ide_env,
                           const nrc_shader_images_t* images);
/* VJP: writes d_g_heads[:,0] (roughness raw) and d_g_normals [P,3] from the gradients of dotprod,
 * the SLF encoding and (optionally) the EnvMap encoding; all with row strides. */
int32_t nrc_shader_mid_bwd(void* stream, int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l,
                           const float* sigma, const float* d_mat, int32_t n_sh_env, const float* d_heads,
                           int64_t ldh, const float* d_normals, const float* d_viewdirs, int64_t num_points,
                           int32_t samples_per_ray, float roughness_bias, const float* d_g_dotprod,
                           int64_t ldgd, const float* d_g_ide_slf, int64_t ldgs, const float* d_g_ide_env,
                           int64_t ldge, float* d_g_heads, int64_t ldgh, float* d_g_normals);
/*   out: rgb = ambient_diffuse + indirect_diffuse + tint*F*(env*(1-acc) + ref*acc), acc = 1 for the
 *        IDE-form light field; d_extras [P,22] (may be NULL) = diffuse 3, specular 3, ambient 3,
 *        indirect 3, albedo (tint) 3, integrated BRDF 1, env_rgb 3, ref_rgb 3. */
int32_t nrc_shader_out_fwd(void* stream, const float* d_heads, int64_t ldh, const float* d_f_raw, int64_t ldf,
                           const float* d_slf_raw, int64_t lds, const float* d_env_raw, int64_t lde,
                           int64_t num_points, float rgb_max, float diffuse_bias, float light_bias,
                           float brdf_bias, float* d_rgb, float* d_extras);
/* VJP of rgb: writes d_g_heads[:,1:10], d_g_f_raw[:,0], d_g_slf_raw[:,0:3]; the EnvMap receives an
 * exactly-zero gradient (1 - acc == 0). */
int32_t nrc_shader_out_bwd(void* stream, const float* d_heads, int64_t ldh, const float* d_f_raw, int64_t ldf,
                           const float* d_slf_raw, int64_t lds, int64_t num_points, float rgb_max,
                           float diffuse_bias, float light_bias, float brdf_bias, const float* d_g_rgb,
                           float* d_g_heads, int64_t ldgh, float* d_g_f_raw, int64_t ldgf,
                           float* d_g_slf_raw, int64_t ldgs);

/* ------------------------------------------- training-step glue (fused) ---- */
/* normals = nan_to_num(-l2_normalize(grad)) (internal/ref_utils.py:45-70, internal/geometry.py:442-479):
 * d_grad [P,3] -> d_normals [P,3]. */
int32_t nrc_normals_fwd(void* stream, const float* d_grad, int64_t num_points, float* d_normals);
/* VJP with the reference's gradient override (denominator clamped at float32 eps in the backward pass). */
int32_t nrc_normals_bwd(void* stream, const float* d_grad, const float* d_g_normals, int64_t num_points,
                        float* d_g_grad);
/* Cache-stage objective of the benchmark step and its gradients in one pass (the loss is the root of
 * the backward pass): mean Charbonnier(linear_to_srgb(rgb) - target) (internal/image.py:192-200,
 * configs/ngp_yobo.gin:35-37) + prop_weight * sum_l mean((sum w_l - stop_grad(sum w_2))^2), l = 0,1.
 *   d_rgb, d_target [R,3]; d_w{0,1,2} [R,n{0,1,2}] -> d_loss [1], d_g_rgb [R,3], d_g_w0, d_g_w1. */
int32_t nrc_cache_loss(void* stream, const float* d_rgb, const float* d_target, const float* d_w0, int32_t n0,
                       const float* d_w1, int32_t n1, const float* d_w2, int32_t n2, int64_t num_rays,
                       float charb_padding, float prop_weight, float* d_loss, float* d_g_rgb, float* d_g_w0,
                       float* d_g_w1);

/* coord.pos_enc (internal/coord.py:298-312): d_x [P,dim] -> d_out [P, (append_identity ? dim : 0) +
 * 2*dim*(max_deg-min_deg)] (row stride ldo): [x | sin(2^j x) | sin(2^j x + pi/2)]. */
int32_t nrc_pos_enc(void* stream, const float* d_x, int64_t num_points, int32_t dim, int32_t min_deg,
                    int32_t max_deg, int32_t append_identity, float* d_out, int64_t ldo);

/* ------------------------------------------- material stage (rows 18-19) ---- */
/* importance_sample_rays + get_secondary_rays (internal/inverse_render/render_utils.py:722-1056) for ONE
 * sampler set: n_microfacet MicrofacetSampler (:485-546) + n_cosine CosineSampler (:417-444) + n_light
 * LightSampler (vMF mixture, :1335-1490) samples per shaded point, in this order, with the MIS power
 * heuristic (:817-853) over the samplers present.  Local frame from get_rotation_matrix (:145-168).
 *   d_means, d_viewdirs (camera view directions), d_normals [R,3]; d_roughness [R];
 *   d_u [R,S,2] uniforms in [0,1) (a light sample's own uniform travels in its first slot);
 *   light sampler (NULL when n_light == 0): d_vmf_means [R,K,3], d_vmf_kappas, d_vmf_logits [R,K],
 *   d_latent [R] int32 (mixture component drawn per point), d_normal2 [R,n_light,2] Gaussian pairs.
 *   -> d_origins = means + normal * normal_eps, d_directions (global) [R,S,3], d_local_lightdirs [R,S,3],
 *      d_local_viewdirs [R,3], d_pdf, d_weight [R,S]. */
int32_t nrc_secondary_sample(void* stream, const float* d_means, const float* d_viewdirs, const float* d_normals,
                             const float* d_roughness, int64_t num_points, int32_t n_microfacet, int32_t n_cosine,
                             int32_t n_light, const float* d_u, const float* d_vmf_means, const float* d_vmf_kappas,
                             const float* d_vmf_logits, int32_t num_lobes, const int32_t* d_latent,
                             const float* d_normal2, float normal_eps, float* d_origins, float* d_directions,
                             float* d_local_lightdirs, float* d_local_viewdirs, float* d_pdf, float* d_weight);
/* _get_microfacet_material (internal/material.py:1290-1322, table :957-1023, configs/ngp_yobo.gin:256-303):
 * d_brdf_params [P,ld] (10 raw channels of pred_brdf_layer) -> albedo [P,3], roughness, metalness, F_0 [P]
 * (constant Fresnel), specular_albedo [P] (may be NULL). */
int32_t nrc_material_head(void* stream, const float* d_brdf_params, int64_t ld, int64_t num_points,
                          float min_roughness, float default_f0, float* d_albedo, float* d_roughness,
                          float* d_metalness, float* d_f0, float* d_specular_albedo);

/* ------------------------------------ transient (time-resolved) rendering, row 22 ---- */
/* render.volumetric_transient_rendering (internal/render.py:250-449: shift_direct :452-490,
 * shift_map_coordinates :493-507) fused with the transient heads' post-processing
 * (internal/nerf.py:1660-1777: softplus(raw + bias) * indirect_scale, specular = tint*F * ref * indirect_scale,
 * clip to [0, rgb_max]) and render_utils.zero_invalid_bins (internal/inverse_render/render_utils.py:1699-1767).
 *   d_direct_rgbs [R,n,C]; d_diffuse_raw [R,n,B,C] raw irradiance-head output or NULL; d_specular [R,n,B,C]
 *   activated light-field output or NULL with d_spec_scale [R,n,C] (tint * integrated BRDF);
 *   d_weights, d_ray_dists, d_light_dists, d_cam_dists [R,n] (cam_dists = |o - x| + |o - cam_origin|)
 *   -> d_transient_direct, d_transient_indirect, d_rgb [R,B,C]  (rgb = direct + indirect + dark_level).
 * The [R,n,B,C] inputs are read once; nothing of that size is written. */
int32_t nrc_transient_render_fwd(void* stream, const float* d_direct_rgbs, const float* d_diffuse_raw,
                                 const float* d_specular, const float* d_spec_scale, const float* d_weights,
                                 const float* d_ray_dists, const float* d_light_dists, const float* d_cam_dists,
                                 int64_t num_rays, int32_t n, int32_t n_bins, int32_t channels, float exposure_time,
                                 float shift, float diffuse_bias, float indirect_scale, float bin_zero_threshold_light,
                                 int32_t light_zero, float light_near, float rgb_max, float dark_level,
                                 float* d_transient_direct, float* d_transient_indirect, float* d_rgb);

/* Proposal supervision: one level of loss_utils.spline_interlevel_loss
 * (internal/loss_utils.py:74-108; blur_and_resample_weights, internal/stepfun.py:463-483, over
 * internal/linspline.py:95-141,187-221):  w_blur = stop_gradient(blur(t, w, blur_halfwidth) resampled into tq),
 * loss += mult * mean(max(0, w_blur - wp)^2 / (wp + eps)) (ACCUMULATED into d_loss[0]), d_g_wp [R,nq] = d loss / d wp.
 *   d_t [R,m+1], d_w [R,m] final-level step function (m <= 64); d_tq [R,nq+1], d_wp [R,nq] proposal level (nq <= 128);
 *   d_w_blur [R,nq] optional output. */
int32_t nrc_interlevel_loss(void* stream, const float* d_t, const float* d_w, int32_t m, const float* d_tq,
                            const float* d_wp, int32_t nq, int64_t num_rays, float blur_halfwidth, float mult,
                            float eps, float* d_loss, float* d_g_wp, float* d_w_blur);
/* Tail of the cache training step in one launch: volumetric_rendering's rgb [R,3] and acc [R] from the shaded samples
 * d_values [R,n,3] and d_weights [R,n] over the background d_bg [R,3] (may be NULL) (internal/render.py:172-224), the
 * Charbonnier-sRGB data term against d_target [R,3], compute_mask_loss on acc (use_mask != 0; d_mask NULL = all ones;
 * internal/train_utils.py:785-836) and the VJP of the compositing: loss += both terms (accumulated), d_g_values [R,n,3]
 * and d_g_weights [R,n] written.  Equivalent to nrc_ray_composite_fwd + nrc_charb_srgb_loss + nrc_mask_loss +
 * nrc_ray_composite_bwd (without the distance statistics, which the objective does not consume). */
int32_t nrc_render_loss(void* stream, const float* d_values, const float* d_weights, const float* d_bg,
                        const float* d_target, const float* d_mask, int64_t num_rays, int32_t n, float charb_padding,
                        int32_t use_mask, float opaque_weight, float empty_weight, float* d_loss, float* d_out_rgb,
                        float* d_acc, float* d_g_values, float* d_g_weights);
/* The same tail with the cache shader's per-point `out` stage folded in on both sides (training path of the fused step):
 * heads / f_raw / slf_raw / env_raw are the raw stack outputs nrc_shader_out_fwd takes (same column meaning and biases),
 * d_weights [R,n] the final level's weights.  Per ray: out stage of its n samples -> rgb, acc -> data + mask terms ->
 * VJP of the compositing (d_g_weights [R,n]) -> VJP of the out stage (d_g_heads [P,ldgh] columns 1-9, d_g_f_raw [P,ldgf]
 * column 0, d_g_slf_raw [P,ldgs] columns 0-2; P = R*n).  d_rgb_samples [P,3] (may be NULL) receives the per-sample colours.
 * Replaces nrc_shader_out_fwd + nrc_render_loss + nrc_shader_out_bwd (three dependent launches) with one; n <= 128. */
int32_t nrc_shade_render_loss(void* stream, const float* d_heads, int64_t ldh, const float* d_f_raw, int64_t ldf,
                              const float* d_slf_raw, int64_t lds, const float* d_env_raw, int64_t lde,
                              float rgb_max, float diffuse_bias, float light_bias, float brdf_bias,
                              const float* d_weights, const float* d_bg, const float* d_target, const float* d_mask,
                              int64_t num_rays, int32_t n, float charb_padding, int32_t use_mask,
                              float opaque_weight, float empty_weight, float* d_loss, float* d_rgb_samples,
                              float* d_out_rgb, float* d_acc, float* d_g_weights, float* d_g_heads, int64_t ldgh,
                              float* d_g_f_raw, int64_t ldgf, float* d_g_slf_raw, int64_t ldgs);
/* Data term: loss += mean Charbonnier(linear_to_srgb(rgb) - target) (accumulated), d_g_rgb [R,3] written. */
int32_t nrc_charb_srgb_loss(void* stream, const float* d_rgb, const float* d_target, int64_t num_rays,
                            float charb_padding, float* d_loss, float* d_g_rgb);
/* compute_mask_loss (internal/train_utils.py:785-836) with lossmult == 1: loss += mean(Charbonnier(acc - mask) *
 * (mask > 0.5 ? opaque_weight : empty_weight)); d_mask NULL = all ones; d_g_acc [R] written.  n > 0: d_acc is the
 * ray's weights [R,n] (acc = their sum, the weights_only pass) and d_g_acc [R,n] their gradient.  The backward-mask
 * pass (train_utils.py:2929-2945,3348-3401) calls it with a zero mask, opaque_weight 0 and
 * empty_weight = backward_mask_loss_weight on the accumulation of the extra rays. */
int32_t nrc_mask_loss(void* stream, const float* d_acc, int32_t n, const float* d_mask, int64_t num_rays,
                      float charb_padding, float opaque_weight, float empty_weight, float* d_loss, float* d_g_acc);
/* param_regularizer_loss (internal/train_utils.py:1169-1216) for one grid module with the common setting
 * (mult, jnp.mean, alpha=2, scale=1; Config.param_regularizers, configs/nerf_ngp_yobo.gin:47-51): for every level
 * table T of `enc`: loss += mult * 0.5 * mean(T^2), levels[l].d_grad += mult * T / numel(T) (atomic reductions: may run
 * concurrently with nrc_encode_bwd on the same tables). */
int32_t nrc_grid_regularizer(void* stream, const nrc_encoding_t* enc, float mult, float* d_loss);
/* Same loss, but levels[l].d_grad is OVERWRITTEN with mult * T / numel(T) (plain stores): the launch doubles as the
 * per-step zero-fill of these gradient tables.  Must be ordered BEFORE every scatter into them. */
int32_t nrc_grid_regularizer_init(void* stream, const nrc_encoding_t* enc, float mult, float* d_loss);
/* Zero-fill of up to 8 ranges [lo[k], hi[k]) (in floats, multiples of 4; d_base 16-byte aligned) of one buffer in ONE
 * launch: the per-step clearing of the gradient arena minus the tables nrc_grid_regularizer_init overwrites.  lo / hi are
 * HOST arrays (copied into the launch parameters).  streaming != 0: evict-first stores, so that the fill does not displace
 * the level tables from L2 while the sampler's forward gathers them. */
int32_t nrc_zero_ranges(void* stream, float* d_base, const int64_t* lo, const int64_t* hi, int32_t num_ranges,
                        int32_t streaming);
/* Gradient all-reduce (MEAN over ranks) of the flat gradient arena over NVLink / NVSwitch peer memory: the reference's
 * lax.pmean over the gradient pytree (internal/train_utils.py:3132-3136).  The arena is symmetric memory (same size on
 * every rank); rank r reduces the r-th slice of floats [offset, offset+count) and writes the mean into EVERY rank's copy.
 * `mc_base` = multicast (NVLS) address of the arena: in-switch reduction (multimem.ld_reduce / multimem.st);
 * `peer_bases` = HOST array of the `world` peer mappings of the arena: plain peer loads / stores (no multicast object).
 * offset and count are multiples of 4 floats.  The caller brackets the launch with a cross-rank barrier on `stream`
 * (all producers done before, all slices written after); num_ctas <= 0 picks the default. */
int32_t nrc_allreduce_mean_multicast(void* stream, float* mc_base, int64_t offset, int64_t count, int32_t rank,
                                     int32_t world, int32_t num_ctas);
int32_t nrc_allreduce_mean_peer(void* stream, float* const* peer_bases, int64_t offset, int64_t count, int32_t rank,
                                int32_t world, int32_t num_ctas);
/* Measurement aid ("L2 roofline denominator"): every thread issues `per_thread` independent, uniformly random
 * row reads (row_bytes = 4 or 16, 8 in flight) from a table of `table_rows` rows and adds them up; d_sink [1] keeps the
 * loads alive.  bench.py times it on an L2-resident table to get the B200's random-gather rate, the bound the
 * hash-level gathers of HashEncoding (internal/grid_utils.py:41-121) are reported against beside the HBM peak. */
int32_t nrc_probe_gather(void* stream, const float* d_table, int64_t table_rows, int32_t row_bytes, int64_t num_threads,
                         int32_t per_thread, float* d_sink);
/* Distortion loss of mip-NeRF 360 on the final level (internal/loss_utils.py:108-123, internal/stepfun.py:253-269;
 * Config.distortion_loss_target='tdist', curve_fn = math.power_ladder(p, premult), configs/ngp_yobo.gin:250-253,
 * mult configs/nerf_ngp_yobo_lego.gin:10): d_t [R,n+1] metric fenceposts, d_weights [R,n], n <= 128.
 * loss += mult * mean_r(...); d_g_weights [R,n] is ACCUMULATED. */
int32_t nrc_distortion_loss(void* stream, const float* d_t, const float* d_weights, int32_t n, int64_t num_rays, float p,
                            float premult, float mult, float* d_loss, float* d_g_weights);
/* Geometry losses on the final sampler level (internal/train_utils.py:3255-3311, internal/loss_utils.py:127-199):
 * orientation loss on the predicted normals (orientation_loss_target='normals_pred'), predicted-normal loss
 * (gt = stop_gradient(normals_pred), pred = the analytic normals; weights through stopgrad_with_weight) and its
 * reverse (gt = stop_gradient(normals), pred = normals_pred, weights stop-gradiented); the ease / decay schedule
 * factors are folded into the mults by the caller.
 *   d_weights [R,n], d_normals [R,n,3] (analytic; NULL disables both predicted-normal terms), d_normals_pred
 *   [R,n,3], d_viewdirs [R,3].  loss += ...; d_g_weights [R,n] and d_g_normals_pred [R,n,3] are ACCUMULATED,
 *   d_g_normals [R,n,3] is written (feed it to nrc_normals_bwd, then nrc_density_normals_bwd). */
int32_t nrc_geometry_losses(void* stream, const float* d_weights, const float* d_normals, const float* d_normals_pred,
                            const float* d_viewdirs, int64_t num_rays, int32_t n, float orientation_mult,
                            float predicted_normal_mult, float predicted_normal_reverse_mult, float stopgrad_weight,
                            float* d_loss, float* d_g_weights, float* d_g_normals_pred, float* d_g_normals);

/* VJP of nrc_transient_render_fwd (training the time-resolved cache, config 4): d_g_transient_direct /
 * d_g_transient_indirect [R,B,C] are the gradients of the two histograms (rgb = direct + indirect + dark_level: add rgb's
 * gradient to both; either may be NULL = zero) -> d_g_direct_rgbs [R,n,C], d_g_diffuse_raw / d_g_specular [R,n,B,C] (NULL
 * when that head is absent or its gradient is not wanted), d_g_spec_scale [R,n,C] (may be NULL), d_g_weights [R,n] - all
 * WRITTEN.  The distances are stop-gradient inputs (internal/sampling.py:353-354).  Same constants as the forward. */
int32_t nrc_transient_render_bwd(void* stream, const float* d_direct_rgbs, const float* d_diffuse_raw,
                                 const float* d_specular, const float* d_spec_scale, const float* d_weights,
                                 const float* d_ray_dists, const float* d_light_dists, const float* d_cam_dists,
                                 int64_t num_rays, int32_t n, int32_t n_bins, int32_t channels, float exposure_time,
                                 float shift, float diffuse_bias, float indirect_scale, float bin_zero_threshold_light,
                                 int32_t light_zero, float light_near, float rgb_max, const float* d_g_transient_direct,
                                 const float* d_g_transient_indirect, float* d_g_direct_rgbs, float* d_g_diffuse_raw,
                                 float* d_g_specular, float* d_g_spec_scale, float* d_g_weights);

/* ------------------------------------------- time-resolved path, fused (row 22, config 4) ---- */
/* nrc_transient_render_fwd with the LAST LAYER of both transient heads inside: the per-sample histograms
 * [R, n, n_bins, C] the reference materialises (internal/nerf.py:1660-1777 get_indirect / transient SurfaceLightField
 * output, internal/surface_light_field.py:1033-1041) never reach HBM.
 *   d_h_diffuse  [R*n, 64]  hidden activation of the irradiance stack (after its last ReLU), d_w_diffuse Flax kernel
 *                [64, >= n_bins*C] (row stride ld_w_diffuse), d_b_diffuse [n_bins*C]:
 *                diffuse = clip(softplus(h W + b + diffuse_bias) * indirect_scale, 0, rgb_max)
 *   d_h_specular [R*n, 128] hidden activation of the transient SurfaceLightField stack, d_w_specular [128, >= n_bins*C]
 *                (the 2101-wide output layer: row stride ld_w_specular, the alpha column is not read), d_b_specular:
 *                specular = clip(spec_scale * clip(softplus(spec_premult * (h W + b) + spec_bias), 0, spec_max)
 *                                * indirect_scale, 0, rgb_max),   d_spec_scale [R,n,C] = tint * integrated BRDF
 *   either head may be NULL.  Validity masks, sub-bin shift, weighted reduction and outputs as nrc_transient_render_fwd.
 *   The head GEMMs run on the tensor cores with bf16 operands (north-star bf16-MLP variant); the activated values of a
 *   ray's samples are staged in shared memory as bf16 before the reduction.
 *   d_w_packed: caller-owned scratch of n_bins*C*192 bf16 holding both kernels in the kernel's operand order; written
 *   when repack != 0 (first call / after a parameter update), reused otherwise. */
int32_t nrc_transient_head_render_fwd(
    void* stream, const float* d_direct_rgbs, const float* d_h_diffuse, int32_t k_diffuse, const float* d_w_diffuse,
    int64_t ld_w_diffuse, const float* d_b_diffuse, const float* d_h_specular, int32_t k_specular, const float* d_w_specular,
    int64_t ld_w_specular, const float* d_b_specular, const float* d_spec_scale, const float* d_weights, const float* d_ray_dists,
    const float* d_light_dists, const float* d_cam_dists, int64_t num_rays, int32_t n, int32_t n_bins, int32_t channels,
    float exposure_time, float shift, float diffuse_bias, float spec_premult, float spec_bias, float spec_max,
    float indirect_scale, float bin_zero_threshold_light, int32_t light_zero, float light_near, float rgb_max, float dark_level,
    void* d_w_packed, int32_t repack, float* d_transient_direct, float* d_transient_indirect, float* d_rgb);
/* The temporal filter of volumetric_transient_rendering (internal/render.py:397-415): convolution along the bin axis with
 * the impulse response / normalised Gaussian d_filter [taps], mode 'same'.  d_x, d_y [R, n_bins, C] (not in place). */
int32_t nrc_transient_filter(void* stream, const float* d_x, const float* d_filter, int32_t taps, int64_t num_rays,
                             int32_t n_bins, int32_t channels, float* d_y);
/* transient_integrate_reflect_rays, direct=False (internal/inverse_render/render_utils.py:1195-1302): like
 * nrc_ggx_integrate_fwd with a histogram of incoming radiance per secondary ray, d_radiance [R,S,n_bins,3]
 *   -> d_radiance_out, d_irradiance [R,n_bins,3] (the second may be NULL), d_occ_out [R] (with d_occ [R,S]). */
int32_t nrc_ggx_integrate_transient_fwd(void* stream, const float* d_wi, const float* d_wo, const float* d_radiance,
                                        const float* d_weight, const float* d_pdf, const float* d_occ, const float* d_albedo,
                                        const float* d_roughness, const float* d_metalness, const float* d_f0,
                                        int64_t num_points, int32_t num_samples, int32_t n_bins, int32_t lobe_kind,
                                        float rgb_max, float* d_radiance_out, float* d_irradiance, float* d_occ_out);

/* ------------------------------- camera rays and the whole-image chunk loop (rows 23, 8f-3) ---- */
/* camera_utils.pixels_to_rays (internal/camera_utils.py:896-1073) for the perspective camera without distortion, NDC or
 * jitter + the near / far broadcast of cast_ray_batch (:1225-1330), for `num_rays` consecutive pixels of a row-major
 * image starting at flat index first_pixel (y * width + x); d_first_pixel != NULL reads that index from device memory
 * (the chunk counter of a graph-replayed render).  Indices beyond last_pixel repeat last_pixel (edge padding of
 * internal/models.py:2434-2445).  pixtocam [9] / camtoworld [12] are HOST arrays (row major).
 *   -> d_origins, d_directions (not normalised), d_viewdirs [N,3], d_radii [N], d_imageplane [N,2] (may be NULL),
 *      d_near, d_far [N] (may be NULL). */
int32_t nrc_camera_rays(void* stream, const float* pixtocam, const float* camtoworld, int32_t width, int32_t height,
                        int64_t first_pixel, const int64_t* d_first_pixel, int64_t last_pixel, int64_t num_rays,
                        float near, float far, float* d_origins, float* d_directions, float* d_viewdirs,
                        float* d_radii, float* d_imageplane, float* d_near, float* d_far);
/* The chunk loop of models.render_image (internal/models.py:2361-2525) without host work per chunk: *d_counter += step
 * (last node of a chunk's CUDA graph) ... */
int32_t nrc_chunk_advance(void* stream, int64_t* d_counter, int64_t step);
/* ... and the chunk's results [chunk, channels] stored at the band position of its first pixel (*d_first_pixel -
 * band_first_pixel); rows outside [0, band_pixels) - the padding rays of the last chunk - are dropped. */
int32_t nrc_band_store(void* stream, const float* d_src, int32_t channels, const int64_t* d_first_pixel,
                       int64_t band_first_pixel, int64_t band_pixels, int64_t chunk, float* d_band);

/* --------------------------------------------- light sampler ---- */
/* vMF head of LightMLP: get_vmfs + the recentring of predict_lighting (internal/light_sampler.py:135-160,203-204).
 *   d_raw [P, K*5] (output layer), d_means_random [P,K,3] (means_random_per_point = 1) or [K,3] (0), d_positions [P,3]
 *   -> d_means [P,K,3] = raw[0:3]*vmf_scale + means_random - position, d_kappas [P,K] = min(softplus(raw[3]+1), 50),
 *      d_logits [P,K] = max(raw[4]+1, -50).  The reference draws means_random from a fixed JAX key: an input here. */
int32_t nrc_vmf_head_fwd(void* stream, const float* d_raw, const float* d_means_random, int32_t means_random_per_point,
                         const float* d_positions, int64_t num_points, int32_t num_lobes, float vmf_scale,
                         float* d_means, float* d_kappas, float* d_logits);
int32_t nrc_vmf_head_bwd(void* stream, const float* d_raw, const float* d_g_means, const float* d_g_kappas,
                         const float* d_g_logits, int64_t num_points, int32_t num_lobes, float vmf_scale, float* d_g_raw);
/* render_utils.vmf_loss_fn (internal/inverse_render/render_utils.py:1493-1550) as called by
 * train_utils.light_sampling_loss (internal/train_utils.py:1985-2071), function_vals_nocorr == function_vals:
 *   d_means [P,K,3], d_kappas [P,K], d_logits [P,K] (K <= 128), d_normals [P,3], d_dirs [P,S,3], d_pdf / d_weight /
 *   d_function_vals [P,S], lossmult (1/S), linear_to_srgb flag.  loss += mean over (P,S); gradients w.r.t. means
 *   (through l2_normalize, grad_eps 1e-5), kappas and logits are written. */
int32_t nrc_vmf_loss(void* stream, const float* d_means, const float* d_kappas, const float* d_logits,
                     const float* d_normals, const float* d_dirs, const float* d_pdf, const float* d_weight,
                     const float* d_function_vals, int64_t num_points, int32_t num_lobes, int32_t num_samples,
                     float lossmult, int32_t linear_to_srgb, float* d_loss, float* d_g_means, float* d_g_kappas,
                     float* d_g_logits);

/* ------------------------------------------------- K5: GGX integration ---- */
/* render_utils.get_lobe (internal/inverse_render/render_utils.py:566-695) +
 * integrate_reflect_rays (:1102-1193): Disney-GGX D*F*G and Lambert lobes evaluated in the
 * local shading frame (normal = +z), MIS-weighted Monte-Carlo mean of cache radiance.
 *   lobe_kind 0 'microfacet', 1 'microfacet_diffuse', 2 'microfacet_specular', 3 'lambertian'
 *   d_wi [R,S,3] local incoming dirs, d_wo [R,S,3] local outgoing dirs, d_radiance [R,S,3],
 *   d_weight [R,S] MIS weights, d_pdf [R,S], d_occ [R,S] or NULL; material: d_albedo [R,3],
 *   d_roughness [R], d_metalness [R], d_f0 [R] (NULL allowed for 'lambertian').
 *   -> d_radiance_out [R,3], d_irradiance [R,3] (may be NULL), d_occ_out [R] (may be NULL). */
int32_t nrc_ggx_integrate_fwd(void* stream, const float* d_wi, const float* d_wo,
                              const float* d_radiance, const float* d_weight, const float* d_pdf,
                              const float* d_occ, const float* d_albedo, const float* d_roughness,
                              const float* d_metalness, const float* d_f0, int64_t num_points,
                              int32_t num_samples, int32_t lobe_kind, float rgb_max,
                              float* d_radiance_out, float* d_irradiance, float* d_occ_out);
/* VJP with respect to the incoming cache radiance: d_g_radiance [R,S,3] written from
 * d_g_out [R,3] (may be NULL) and d_g_irradiance [R,3] (may be NULL). */
int32_t nrc_ggx_integrate_bwd(void* stream, const float* d_wi, const float* d_wo,
                              const float* d_radiance, const float* d_weight, const float* d_pdf,
                              const float* d_albedo, const float* d_roughness,
                              const float* d_metalness, const float* d_f0, const float* d_g_out,
                              const float* d_g_irradiance, int64_t num_points, int32_t num_samples,
                              int32_t lobe_kind, float rgb_max, float* d_g_radiance);

/* ------------------- surface-light-field memory variant (second half) ---- */
/* Constants of BaseSurfaceLightFieldMLP.predict_points (internal/surface_light_field.py:594-780) as `surface_lf_mem` is
 * configured (internal/models.py:813-833, configs/nerf_ngp_yobo.gin:97-165, ngp_yobo.gin:232-236). */
typedef struct nrc_slf_points_t {
  int32_t num_distance_samples; /* n: the distance network emits 8 n + 4 columns */
  int32_t warp_kind;            /* raydist_fn: 0 identity, 1 math.power_ladder(p, premult) */
  float distance_near, distance_far; /* the module's own range (ray warps, mask, clip) */
  float near, far;              /* the call's `near` / `far` keyword arguments (mask only); 0 / +inf by default */
  float distance_scale, distance_bias;
  float rgb_premultiplier, rgb_bias, alpha_bias;
  float warp_p, warp_premult;
  float ref_warp_c;             /* ref_warp_fn = coord.contract_radius_c; <= 0: identity */
} nrc_slf_points_t;

/* predict_points for use_voxel_grid = use_sorted_distances = use_point_offsets = False, num_far_samples = 0,
 * use_env_alpha = True, and the head of __call__ that follows it (:899-913): ref_warp_fn on the points, softmax of the raw
 * weights, the weighted s-distance, weights * mask * env_alpha.
 *   d_raw [P, 8n+4] (row stride ld_raw floats): per sample [offset, sigma, -, -, raw_weight, -, -, -], then env rgb (3) and
 *   env alpha (1);  d_origins, d_refdirs [P,3]
 *   -> d_points [P,n,3] (warped), d_weights [P,n], d_s_dist [P], d_distances [P,n] (clipped), d_env_rgba [P,4]. */
int32_t nrc_slf_points_fwd(void* stream, const nrc_slf_points_t* cfg, const float* d_raw, int64_t ld_raw,
                           const float* d_origins, const float* d_refdirs, int64_t num_points, float* d_points,
                           float* d_weights, float* d_s_dist, float* d_distances, float* d_env_rgba);
/* VJP of the above with respect to d_raw (origins and directions are stop-gradient inputs of the light field: models.py:854,
 * utils.partial_stopgrad_rays).  Upstream gradients may be NULL (= zero).  d_g_raw [P, 8n+4] dense, every column written. */
int32_t nrc_slf_points_bwd(void* stream, const nrc_slf_points_t* cfg, const float* d_raw, int64_t ld_raw,
                           const float* d_origins, const float* d_refdirs, int64_t num_points, const float* d_g_points,
                           const float* d_g_weights, const float* d_g_s_dist, const float* d_g_distances,
                           const float* d_g_env_rgba, float* d_g_raw);
/* Weighted feature sum over the predicted points (:981): d_feat [P,n,F], d_weights [P,n] -> d_out [P,F]. */
int32_t nrc_slf_reduce_fwd(void* stream, const float* d_feat, const float* d_weights, int64_t num_points,
                           int32_t num_samples, int32_t num_features, float* d_out);
/* ... and its VJP: d_g_out [P,F] -> d_g_feat [P,n,F], d_g_weights [P,n]. */
int32_t nrc_slf_reduce_bwd(void* stream, const float* d_feat, const float* d_weights, const float* d_g_out,
                           int64_t num_points, int32_t num_samples, int32_t num_features, float* d_g_feat,
                           float* d_g_weights);

#ifdef __cplusplus
}
#endif
#endif /* NRC_B200_H_ */
