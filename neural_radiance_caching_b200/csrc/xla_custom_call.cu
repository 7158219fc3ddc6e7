// XLA custom-call targets (include/nrc_xla.h): the legacy GPU custom-call ABI of the jaxlib the reference pins
// (jax==0.4.16, requirements.txt:2) - void(cudaStream_t, void** buffers, const char* opaque, size_t opaque_len) - over
// the entry points of nrc_b200.h.  Each target checks the descriptor's size and version, resolves its operands / results
// from `buffers` and forwards; nothing is computed here.
#include <cstring>

#include "nrc_common.cuh"
#include "nrc_xla.h"

namespace {

thread_local int32_t g_xla_status = NRC_OK;

template <typename Desc>
bool unpack(const char* opaque, size_t len, Desc& d) {
  if (!opaque || len != sizeof(Desc)) { g_xla_status = NRC_E_INVALID_ARG; return false; }
  std::memcpy(&d, opaque, sizeof(Desc));   // the opaque string has no alignment guarantee
  return true;
}
inline void done(int32_t st) { if (st != NRC_OK) g_xla_status = st; }
inline const float* F(void** b, int i) { return static_cast<const float*>(b[i]); }
inline float* O(void** b, int i) { return static_cast<float*>(b[i]); }

// level tables (and gradient tables) of an arena operand
bool bind_arena(nrc_xla_encode_desc_t& d, const float* arena, float* g_arena) {
  if (d.version != NRC_XLA_DESC_VERSION || d.enc.num_levels < 1 || d.enc.num_levels > NRC_MAX_LEVELS) {
    g_xla_status = NRC_E_INVALID_ARG;
    return false;
  }
  for (int l = 0; l < d.enc.num_levels; ++l) {
    if (d.level_offset[l] < 0 || d.level_offset[l] >= d.arena_floats) { g_xla_status = NRC_E_INVALID_ARG; return false; }
    d.enc.levels[l].d_table = arena ? const_cast<float*>(arena) + d.level_offset[l] : nullptr;
    d.enc.levels[l].d_grad = g_arena ? g_arena + d.level_offset[l] : nullptr;
  }
  return true;
}

}  // namespace

extern "C" {

int32_t nrc_xla_last_status(void) {
  const int32_t s = g_xla_status;
  g_xla_status = NRC_OK;
  return s;
}

void nrc_xla_encode_fwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_encode_desc_t d;
  if (!unpack(opaque, len, d) || !bind_arena(d, F(b, 1), nullptr)) return;
  done(nrc_encode_fwd(stream, &d.enc, F(b, 0), d.num_points, O(b, 2)));
}

void nrc_xla_encode_bwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_encode_desc_t d;
  if (!unpack(opaque, len, d)) return;
  float* g_arena = O(b, 3);
  if (!bind_arena(d, F(b, 1), g_arena)) return;
  // XLA hands out uninitialised result buffers; nrc_encode_bwd accumulates
  if (cudaMemsetAsync(g_arena, 0, static_cast<size_t>(d.arena_floats) * sizeof(float), static_cast<cudaStream_t>(stream)) !=
      cudaSuccess) {
    g_xla_status = NRC_E_CUDA;
    return;
  }
  done(nrc_encode_bwd(stream, &d.enc, F(b, 0), F(b, 2), d.num_points, O(b, 4)));
}

void nrc_xla_contract_fwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_contract_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_contract_fwd(stream, F(b, 0), d.num_points, d.c, O(b, 1)));
}

void nrc_xla_contract_bwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_contract_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_contract_bwd(stream, F(b, 0), F(b, 1), d.num_points, d.c, O(b, 2)));
}

void nrc_xla_density_query_fwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_density_query_desc_t d;
  if (!unpack(opaque, len, d) || !bind_arena(d.grid, F(b, 1), nullptr)) return;
  nrc_density_mlp_t m{};
  int i = 2;
  m.d_w0 = F(b, i++); m.d_b0 = F(b, i++); m.d_w1 = F(b, i++); m.d_b1 = F(b, i++); m.d_wd = F(b, i++); m.d_bd = F(b, i++);
  if (d.has_pred_normals) { m.d_wn = F(b, i++); m.d_bn = F(b, i++); }
  m.in_dim = d.in_dim; m.width = d.width;
  float* density = O(b, i++);
  float* raw = O(b, i++);
  float* feat = O(b, i++);
  float* grad_pred = d.has_pred_normals ? O(b, i++) : nullptr;
  float* raw_grad = d.want_raw_grad ? O(b, i++) : nullptr;
  done(nrc_density_query_fwd(stream, &d.grid.enc, &m, F(b, 0), d.grid.num_points, d.warp_c, d.density_bias, d.bf16, density,
                             raw, feat, grad_pred, raw_grad, nullptr));
}

void nrc_xla_ray_alpha_weights_fwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ray_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_ray_alpha_weights_fwd(stream, F(b, 0), F(b, 1), F(b, 2), d.num_rays, d.n, d.opaque_background, O(b, 3), O(b, 4),
                                 O(b, 5)));
}

void nrc_xla_ray_alpha_weights_bwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ray_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_ray_alpha_weights_bwd(stream, F(b, 0), F(b, 1), F(b, 2), F(b, 3), F(b, 4), F(b, 5), d.num_rays, d.n, O(b, 6)));
}

void nrc_xla_ray_sample_intervals(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ray_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_ray_sample_intervals(stream, F(b, 0), F(b, 1), F(b, 2), F(b, 3), d.num_rays, d.m, d.n, d.anneal, d.padding,
                                d.max_jitter, d.dom_lo, d.dom_hi, O(b, 4), nullptr));
}

void nrc_xla_ray_cast(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ray_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_ray_cast(stream, F(b, 0), F(b, 1), F(b, 2), F(b, 3), F(b, 4), d.num_rays, d.n, d.warp_kind, d.p, d.premult, O(b, 5),
                    O(b, 6)));
}

void nrc_xla_ray_composite_fwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ray_desc_t d;
  if (!unpack(opaque, len, d)) return;
  int i = 0;
  const float* values = F(b, i++);
  const float* weights = F(b, i++);
  const float* wnf = d.has_weights_nf ? F(b, i++) : nullptr;
  const float* tdist = F(b, i++);
  const float* bg = d.has_bg ? F(b, i++) : nullptr;
  float* out = O(b, i++);
  float* acc = O(b, i++);
  float* dist = O(b, i++);
  done(nrc_ray_composite_fwd(stream, values, weights, d.k, wnf, tdist, bg, d.num_rays, d.n, d.channels, d.has_rgb, out, acc, dist));
}

void nrc_xla_ray_composite_bwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ray_desc_t d;
  if (!unpack(opaque, len, d)) return;
  int i = 0;
  const float* values = F(b, i++);
  const float* weights = F(b, i++);
  const float* wnf = d.has_weights_nf ? F(b, i++) : nullptr;
  const float* bg = d.has_bg ? F(b, i++) : nullptr;
  const float* g_out = F(b, i++);
  const float* g_acc = F(b, i++);
  float* g_values = O(b, i++);
  float* g_weights = O(b, i++);
  float* g_wnf = d.has_weights_nf ? O(b, i++) : nullptr;
  done(nrc_ray_composite_bwd(stream, values, weights, d.k, wnf, bg, g_out, g_acc, d.num_rays, d.n, d.channels, d.has_rgb,
                             g_values, g_weights, g_wnf));
}

void nrc_xla_ray_resample(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ray_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_ray_resample(stream, F(b, 0), F(b, 1), d.num_rays, d.n, d.k, d.bias, d.mult, static_cast<int32_t*>(b[2]), O(b, 3)));
}

void nrc_xla_ray_resample_gather(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ray_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_ray_resample_gather(stream, F(b, 0), static_cast<const int32_t*>(b[1]), d.num_rays, d.n, d.k, d.channels, O(b, 2)));
}

void nrc_xla_ggx_integrate_fwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ggx_desc_t d;
  if (!unpack(opaque, len, d)) return;
  int i = 0;
  const float* wi = F(b, i++);
  const float* wo = F(b, i++);
  const float* rad = F(b, i++);
  const float* weight = F(b, i++);
  const float* pdf = F(b, i++);
  const float* occ = d.has_occ ? F(b, i++) : nullptr;
  const float* albedo = F(b, i++);
  const float* rough = F(b, i++);
  const float* metal = F(b, i++);
  const float* f0 = F(b, i++);
  float* out = O(b, i++);
  float* irr = O(b, i++);
  float* occ_out = d.has_occ ? O(b, i++) : nullptr;
  done(nrc_ggx_integrate_fwd(stream, wi, wo, rad, weight, pdf, occ, albedo, rough, metal, f0, d.num_points, d.num_samples,
                             d.lobe_kind, d.rgb_max, out, irr, occ_out));
}

void nrc_xla_ggx_integrate_bwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_ggx_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_ggx_integrate_bwd(stream, F(b, 0), F(b, 1), F(b, 2), F(b, 3), F(b, 4), F(b, 5), F(b, 6), F(b, 7), F(b, 8), F(b, 9),
                             F(b, 10), d.num_points, d.num_samples, d.lobe_kind, d.rgb_max, O(b, 11)));
}

void nrc_xla_slf_points_fwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_slf_desc_t d;
  if (!unpack(opaque, len, d)) return;
  if (d.version != NRC_XLA_DESC_VERSION) { g_xla_status = NRC_E_INVALID_ARG; return; }
  done(nrc_slf_points_fwd(stream, &d.cfg, F(b, 0), d.ld_raw, F(b, 1), F(b, 2), d.num_points, O(b, 3), O(b, 4), O(b, 5), O(b, 6),
                          O(b, 7)));
}

void nrc_xla_slf_points_bwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_slf_desc_t d;
  if (!unpack(opaque, len, d)) return;
  if (d.version != NRC_XLA_DESC_VERSION) { g_xla_status = NRC_E_INVALID_ARG; return; }
  done(nrc_slf_points_bwd(stream, &d.cfg, F(b, 0), d.ld_raw, F(b, 1), F(b, 2), d.num_points, F(b, 3), F(b, 4), F(b, 5), F(b, 6),
                          F(b, 7), O(b, 8)));
}

void nrc_xla_slf_reduce_fwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_slf_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_slf_reduce_fwd(stream, F(b, 0), F(b, 1), d.num_points, d.cfg.num_distance_samples, d.num_features, O(b, 2)));
}

void nrc_xla_slf_reduce_bwd(void* stream, void** b, const char* opaque, size_t len) {
  nrc_xla_slf_desc_t d;
  if (!unpack(opaque, len, d)) return;
  done(nrc_slf_reduce_bwd(stream, F(b, 0), F(b, 1), F(b, 2), d.num_points, d.cfg.num_distance_samples, d.num_features, O(b, 3),
                          O(b, 4)));
}

const nrc_xla_target_t* nrc_xla_targets(void) {
  static const nrc_xla_target_t table[] = {
      {"nrc_xla_encode_fwd", nrc_xla_encode_fwd},
      {"nrc_xla_encode_bwd", nrc_xla_encode_bwd},
      {"nrc_xla_contract_fwd", nrc_xla_contract_fwd},
      {"nrc_xla_contract_bwd", nrc_xla_contract_bwd},
      {"nrc_xla_density_query_fwd", nrc_xla_density_query_fwd},
      {"nrc_xla_ray_alpha_weights_fwd", nrc_xla_ray_alpha_weights_fwd},
      {"nrc_xla_ray_alpha_weights_bwd", nrc_xla_ray_alpha_weights_bwd},
      {"nrc_xla_ray_sample_intervals", nrc_xla_ray_sample_intervals},
      {"nrc_xla_ray_cast", nrc_xla_ray_cast},
      {"nrc_xla_ray_composite_fwd", nrc_xla_ray_composite_fwd},
      {"nrc_xla_ray_composite_bwd", nrc_xla_ray_composite_bwd},
      {"nrc_xla_ray_resample", nrc_xla_ray_resample},
      {"nrc_xla_ray_resample_gather", nrc_xla_ray_resample_gather},
      {"nrc_xla_ggx_integrate_fwd", nrc_xla_ggx_integrate_fwd},
      {"nrc_xla_ggx_integrate_bwd", nrc_xla_ggx_integrate_bwd},
      {"nrc_xla_slf_points_fwd", nrc_xla_slf_points_fwd},
      {"nrc_xla_slf_points_bwd", nrc_xla_slf_points_bwd},
      {"nrc_xla_slf_reduce_fwd", nrc_xla_slf_reduce_fwd},
      {"nrc_xla_slf_reduce_bwd", nrc_xla_slf_reduce_bwd},
      {nullptr, nullptr},
  };
  return table;
}

}  // extern "C"
