/* XLA custom-call entry points of libnrc_b200.so: the seam the reference's JAX code calls through.
 *
 * The reference pins jax==0.4.16 (requirements.txt:2), which predates the typed XLA FFI: GPU custom calls there have
 * the ORIGINAL ABI
 *     void target(cudaStream_t stream, void** buffers, const char* opaque, size_t opaque_len);
 * `buffers` = operand device pointers followed by result device pointers (in the order of the custom call's operands /
 * result tuple), `opaque` = a byte string fixed at trace time.  Every target below unpacks one of the plain-C
 * descriptors of this header from `opaque` and forwards to the entry point of nrc_b200.h named in its comment; the
 * Python side that packs the descriptors and wraps the calls in jax.custom_vjp is
 * neural_radiance_caching_b200/jax_binding/nrc_jax.py (registration through xla_client.register_custom_call_target,
 * the dispatch seam being ResampleOpMode, internal/grid_utils.py:651-676).  The same handlers serve a current jaxlib
 * through jax.ffi.ffi_call(..., custom_call_api_version=2 / legacy mode).
 *
 * The original ABI cannot return a status: a failing call records its nrc status (nrc_xla_last_status(), sticky until
 * read) and leaves the outputs untouched.  tests/test_xla_gpu.py drives every target through ctypes with exactly these
 * arguments - the ABI XLA uses. */
#ifndef NRC_XLA_H_
#define NRC_XLA_H_

#include <stddef.h>
#include <stdint.h>

#include "nrc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NRC_XLA_DESC_VERSION 1

/* HashEncoding: the level tables live in ONE flat fp32 arena operand; level l starts at float offset level_offset[l]
 * (enc.levels[l].d_table / d_grad are ignored and filled in from the arena operands). */
typedef struct {
  int32_t version;
  int32_t reserved;
  int64_t num_points;
  int64_t arena_floats;
  int64_t level_offset[NRC_MAX_LEVELS];
  nrc_encoding_t enc;
} nrc_xla_encode_desc_t;

/* buffers: x [P,3], arena -> out [P, L*F]                                            (nrc_encode_fwd) */
void nrc_xla_encode_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: x [P,3], arena (the d/dx term reads the tables), g_out [P, L*F]
 *          -> g_arena [arena_floats] (zero-filled here, then accumulated), g_x [P,3]  (nrc_encode_bwd) */
void nrc_xla_encode_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);

typedef struct {
  int32_t version;
  int32_t reserved;
  int64_t num_points;
  float c;
  int32_t pad;
} nrc_xla_contract_desc_t;
/* buffers: x [P,3] -> z [P,3]                                                        (nrc_contract_fwd) */
void nrc_xla_contract_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: x [P,3], g_z [P,3] -> g_x [P,3]                                           (nrc_contract_bwd) */
void nrc_xla_contract_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* DensityMLP.predict_density + convert_raw_density (+ analytic raw gradient): encoding descriptor as above plus the
 * network shape and which optional results the call has. */
typedef struct {
  nrc_xla_encode_desc_t grid;
  int32_t in_dim, width;
  int32_t has_pred_normals;   /* operands w_n, b_n present; result grad_pred present */
  int32_t want_raw_grad;      /* result raw_grad present */
  int32_t bf16;
  float warp_c, density_bias;
  int32_t pad;
} nrc_xla_density_query_desc_t;
/* buffers: means [P,3], arena, w0, b0, w1, b1, wd, bd, (wn, bn) -> density [P], raw [P], feat [P,64], (grad_pred [P,3]),
 * (raw_grad [P,3])                                                                   (nrc_density_query_fwd) */
void nrc_xla_density_query_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);

typedef struct {
  int32_t version;
  int32_t n;                  /* samples per ray */
  int64_t num_rays;
  int32_t opaque_background;
  int32_t m;                  /* sample_intervals: current number of intervals */
  float anneal, padding, max_jitter, dom_lo, dom_hi;
  int32_t warp_kind;          /* ray_cast: 0 linear, 1 power ladder */
  float p, premult;
  int32_t k, channels, has_rgb, has_bg, has_weights_nf;   /* composite / resample */
  float bias, mult;           /* resample */
  int32_t pad;
} nrc_xla_ray_desc_t;
/* buffers: density [R,n], tdist [R,n+1], dirs [R,3] -> weights, alpha, trans [R,n]    (nrc_ray_alpha_weights_fwd) */
void nrc_xla_ray_alpha_weights_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: density, tdist, dirs, g_weights, g_alpha, g_trans -> g_density [R,n]      (nrc_ray_alpha_weights_bwd) */
void nrc_xla_ray_alpha_weights_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: t [R,m+1], w [R,m], u01 [R], u_base [n] -> t_new [R,n+1]                   (nrc_ray_sample_intervals) */
void nrc_xla_ray_sample_intervals(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: sdist [R,n+1], origins, directions [R,3], near, far [R,1] -> tdist [R,n+1], means [R,n,3]  (nrc_ray_cast) */
void nrc_xla_ray_cast(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: values [R,k,C], weights [R,k], (weights_nf [R,n]), tdist [R,n+1], (bg [R,3]) -> out [R,C], acc [R], dist [R,4]
 *                                                                                    (nrc_ray_composite_fwd) */
void nrc_xla_ray_composite_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: values, weights, (weights_nf), (bg), g_out [R,C], g_acc [R] -> g_values, g_weights, (g_weights_nf)
 *                                                                                    (nrc_ray_composite_bwd) */
void nrc_xla_ray_composite_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: weights [R,n], gumbel [R,n,k] -> inds [R,k] int32, w_new [R,k]             (nrc_ray_resample) */
void nrc_xla_ray_resample(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: field [R,n,C], inds [R,k] int32 -> out [R,k,C]                             (nrc_ray_resample_gather) */
void nrc_xla_ray_resample_gather(void* stream, void** buffers, const char* opaque, size_t opaque_len);

typedef struct {
  int32_t version;
  int32_t num_samples;
  int64_t num_points;
  int32_t lobe_kind, has_occ;
  float rgb_max;
  int32_t pad;
} nrc_xla_ggx_desc_t;
/* buffers: wi, wo [R,S,3], radiance [R,S,3], weight, pdf [R,S], (occ [R,S]), albedo [R,3], roughness, metalness, f0 [R]
 *          -> radiance_out [R,3], irradiance [R,3], (occ_out [R])                     (nrc_ggx_integrate_fwd) */
void nrc_xla_ggx_integrate_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: wi, wo, radiance, weight, pdf, albedo, roughness, metalness, f0, g_out [R,3], g_irradiance [R,3]
 *          -> g_radiance [R,S,3]                                                      (nrc_ggx_integrate_bwd) */
void nrc_xla_ggx_integrate_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* Surface-light-field memory variant (surface_lf_mem, internal/surface_light_field.py:594-780,899-913,981): the point stage
 * between the distance network and the reflectance grid, and the weighted feature sum. */
typedef struct {
  int32_t version;
  int32_t num_features;      /* reduce targets: F */
  int64_t num_points;
  int64_t ld_raw;            /* row stride of the distance-network output (>= 8 n + 4) */
  nrc_slf_points_t cfg;
} nrc_xla_slf_desc_t;
/* buffers: raw [P, ld_raw], origins [P,3], refdirs [P,3]
 *          -> points [P,n,3], weights [P,n], s_dist [P], distances [P,n], env_rgba [P,4]      (nrc_slf_points_fwd) */
void nrc_xla_slf_points_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: raw, origins, refdirs, g_points, g_weights, g_s_dist, g_distances, g_env_rgba -> g_raw [P, 8n+4]
 *                                                                                             (nrc_slf_points_bwd) */
void nrc_xla_slf_points_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: feat [P,n,F], weights [P,n] -> out [P,F]                                            (nrc_slf_reduce_fwd) */
void nrc_xla_slf_reduce_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);
/* buffers: feat, weights, g_out [P,F] -> g_feat [P,n,F], g_weights [P,n]                       (nrc_slf_reduce_bwd) */
void nrc_xla_slf_reduce_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* Status of the last failing target on this thread (NRC_OK if none); reading clears it. */
int32_t nrc_xla_last_status(void);
/* NULL-terminated table of (name, function) pairs for registration loops. */
typedef struct { const char* name; void (*fn)(void*, void**, const char*, size_t); } nrc_xla_target_t;
const nrc_xla_target_t* nrc_xla_targets(void);

#ifdef __cplusplus
}
#endif
#endif /* NRC_XLA_H_ */
