// K5: material-stage Monte-Carlo integration of cache radiance against the Disney-GGX BRDF.
// One warp per shaded point, lanes over the secondary-ray samples; the lobe, the MIS
// weighting, the clip and the mean never leave registers.
//
// Reference: render_utils.get_lobe (internal/inverse_render/render_utils.py:566-695),
// GGX_D (:480-482), integrate_reflect_rays (:1102-1193), math.normalize/dot
// (internal/inverse_render/math.py:72-88); material flags per configs/ngp_yobo.gin:256-303
// (no diffuseness / mirrorness / specular albedo, brdf_correction = 1).
#include "nrc_common.cuh"

namespace nrc {

constexpr int kGgxWarps = 4;
constexpr float kPi = 3.14159265358979323846f;
constexpr float kDenEps = 1e-5f;  // DENOMINATOR_EPS (render_utils.py:41)

enum LobeKind { kMicrofacet = 0, kMicrofacetDiffuse = 1, kMicrofacetSpecular = 2, kLambertian = 3 };

struct Material { float albedo[3]; float rough; float metal; float f0; };

// Lobe value for one (wi, wo) pair, normal = +z in the local frame.
__device__ __forceinline__ void eval_lobe(int kind, const float wi[3], const float wo[3],
                                          const Material& m, float lobe[3]) {
  const float eps = f32_eps();
  if (kind == kLambertian) {
    float c = fmaxf(0.f, wi[2]);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) lobe[ch] = c * m.albedo[ch] / kPi;
    return;
  }
  float h[3] = {wi[0] + wo[0], wi[1] + wo[1], wi[2] + wo[2]};
  float hn = sqrtf(1e-10f + (h[0] * h[0] + h[1] * h[1] + h[2] * h[2]));
  h[0] /= hn; h[1] /= hn; h[2] /= hn;
  float n_dot_v = fmaxf(0.f, wo[2]);
  float n_dot_l = fmaxf(0.f, wi[2]);
  float n_dot_h = fmaxf(0.f, h[2]);
  float l_dot_h = fmaxf(0.f, wi[0] * h[0] + wi[1] * h[1] + wi[2] * h[2]);
  float a = m.rough;
  float a2 = a * a;
  float t = n_dot_h * n_dot_h * (a2 - 1.f) + 1.f;
  float D = a2 / fmaxf(eps, kPi * (t * t));
  float k = a / 2.f;
  float G = (n_dot_v / fmaxf(eps, n_dot_v * (1.f - k) + k)) * (n_dot_l / fmaxf(eps, n_dot_l * (1.f - k) + k));
  float om = fminf(fmaxf(1.f - l_dot_h, 0.f), 1.f);
  float om5 = (om * om) * (om * om) * om;
  float diffuseness = 1.f - m.metal;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float F0 = m.albedo[ch] * m.metal + m.f0 * (1.f - m.metal);
    float F = F0 + (1.f - F0) * om5;
    float ggx = D * F * G / fmaxf(eps, 4.f * n_dot_v);
    float lam = n_dot_l * m.albedo[ch] / kPi;
    float v;
    if (kind == kMicrofacet) v = ggx + lam * diffuseness;
    else if (kind == kMicrofacetDiffuse) v = lam * diffuseness;
    else v = ggx;
    lobe[ch] = v;
  }
}

__global__ void __launch_bounds__(kGgxWarps * 32)
ggx_integrate_fwd_kernel(const float* __restrict__ wi, const float* __restrict__ wo,
                         const float* __restrict__ radiance, const float* __restrict__ weight,
                         const float* __restrict__ pdf, const float* __restrict__ occ,
                         const float* __restrict__ albedo, const float* __restrict__ roughness,
                         const float* __restrict__ metalness, const float* __restrict__ f0, int64_t R,
                         int S, int kind, float rgb_max, float* __restrict__ radiance_out,
                         float* __restrict__ irradiance, float* __restrict__ occ_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kGgxWarps + warp;
  if (r >= R) return;
  Material m;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) m.albedo[ch] = albedo[3 * r + ch];
  m.rough = roughness ? roughness[r] : 1.f;
  m.metal = metalness ? metalness[r] : 0.f;
  m.f0 = f0 ? f0[r] : 0.04f;
  float ro[3] = {0.f, 0.f, 0.f}, ir[3] = {0.f, 0.f, 0.f}, oc = 0.f;
  for (int s = lane; s < S; s += 32) {
    const int64_t q = r * S + s;
    float li[3] = {wi[3 * q], wi[3 * q + 1], wi[3 * q + 2]};
    float lo[3] = {wo[3 * q], wo[3 * q + 1], wo[3 * q + 2]};
    float lobe[3];
    eval_lobe(kind, li, lo, m, lobe);
    float den = fmaxf(pdf[q], kDenEps);
    float w = fmaxf(weight[q], 0.f);
    if (!(li[2] > 0.f)) w = 0.f;
    float dl = fmaxf(0.f, li[2]) / kPi;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float L = radiance[3 * q + ch];
      ro[ch] += fminf(fmaxf(L * lobe[ch], 0.f), rgb_max) * w / den;
      ir[ch] += fminf(fmaxf(L * dl, 0.f), rgb_max) * w / den;
    }
    if (occ) oc += occ[q];
  }
  const float inv = 1.f / static_cast<float>(S);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float a = ro[ch], b = ir[ch];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      radiance_out[3 * r + ch] = a * inv;
      if (irradiance) irradiance[3 * r + ch] = b * inv;
    }
  }
  if (occ && occ_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) oc += __shfl_xor_sync(0xffffffffu, oc, o);
    if (lane == 0) occ_out[r] = oc * inv;
  }
}

// VJP with respect to the incoming radiance (the path through which the material stage
// trains the cache): d radiance_out / d L = lobe * w / den inside the clip, 0 outside.
__global__ void __launch_bounds__(kGgxWarps * 32)
ggx_integrate_bwd_radiance_kernel(const float* __restrict__ wi, const float* __restrict__ wo,
                                  const float* __restrict__ radiance, const float* __restrict__ weight,
                                  const float* __restrict__ pdf, const float* __restrict__ albedo,
                                  const float* __restrict__ roughness, const float* __restrict__ metalness,
                                  const float* __restrict__ f0, const float* __restrict__ g_out,
                                  const float* __restrict__ g_irr, int64_t R, int S, int kind, float rgb_max,
                                  float* __restrict__ g_radiance) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kGgxWarps + warp;
  if (r >= R) return;
  Material m;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) m.albedo[ch] = albedo[3 * r + ch];
  m.rough = roughness ? roughness[r] : 1.f;
  m.metal = metalness ? metalness[r] : 0.f;
  m.f0 = f0 ? f0[r] : 0.04f;
  const float inv = 1.f / static_cast<float>(S);
  for (int s = lane; s < S; s += 32) {
    const int64_t q = r * S + s;
    float li[3] = {wi[3 * q], wi[3 * q + 1], wi[3 * q + 2]};
    float lo[3] = {wo[3 * q], wo[3 * q + 1], wo[3 * q + 2]};
    float lobe[3];
    eval_lobe(kind, li, lo, m, lobe);
    float den = fmaxf(pdf[q], kDenEps);
    float w = fmaxf(weight[q], 0.f);
    if (!(li[2] > 0.f)) w = 0.f;
    float dl = fmaxf(0.f, li[2]) / kPi;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float L = radiance[3 * q + ch];
      float g = 0.f;
      float v = L * lobe[ch];
      if (g_out && v > 0.f && v < rgb_max) g += g_out[3 * r + ch] * lobe[ch] * w / den * inv;
      float u = L * dl;
      if (g_irr && u > 0.f && u < rgb_max) g += g_irr[3 * r + ch] * dl * w / den * inv;
      g_radiance[3 * q + ch] = g;
    }
  }
}


// transient_integrate_reflect_rays with direct=False (internal/inverse_render/render_utils.py:1195-1302): the incoming
// radiance of every secondary ray is a HISTOGRAM [n_bins, 3]; lobe, weight and pdf are per sample.  One CTA per shaded
// point: the per-sample factors are computed once (warp 0 style: threads over samples) into shared memory, then threads
// over (bin, channel) reduce over the samples.
__global__ void __launch_bounds__(256)
ggx_integrate_transient_fwd_kernel(const float* __restrict__ wi, const float* __restrict__ wo, const float* __restrict__ radiance,
                                   const float* __restrict__ weight, const float* __restrict__ pdf, const float* __restrict__ occ,
                                   const float* __restrict__ albedo, const float* __restrict__ roughness,
                                   const float* __restrict__ metalness, const float* __restrict__ f0, int64_t R, int S, int n_bins,
                                   int kind, float rgb_max, float* __restrict__ radiance_out, float* __restrict__ irradiance,
                                   float* __restrict__ occ_out) {
  extern __shared__ float sm[];       // per sample: lobe[3], diffuse lobe, weight / denominator
  const int64_t r = blockIdx.x;
  float* s_lobe = sm;
  float* s_dl = sm + 3 * S;
  float* s_wd = s_dl + S;
  Material m;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) m.albedo[ch] = albedo[3 * r + ch];
  m.rough = roughness ? roughness[r] : 1.f;
  m.metal = metalness ? metalness[r] : 0.f;
  m.f0 = f0 ? f0[r] : 0.04f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int64_t q = r * S + s;
    float li[3] = {wi[3 * q], wi[3 * q + 1], wi[3 * q + 2]};
    float lo[3] = {wo[3 * q], wo[3 * q + 1], wo[3 * q + 2]};
    float lobe[3];
    eval_lobe(kind, li, lo, m, lobe);
    float w = fmaxf(weight[q], 0.f);
    if (!(li[2] > 0.f)) w = 0.f;
    s_lobe[3 * s] = lobe[0]; s_lobe[3 * s + 1] = lobe[1]; s_lobe[3 * s + 2] = lobe[2];
    s_dl[s] = fmaxf(0.f, li[2]) / kPi;
    s_wd[s] = w / fmaxf(pdf[q], kDenEps);
  }
  __syncthreads();
  const float inv = 1.f / static_cast<float>(S);
  for (int e = threadIdx.x; e < n_bins * 3; e += blockDim.x) {
    const int ch = e % 3;
    float ro = 0.f, ir = 0.f;
    for (int s = 0; s < S; ++s) {
      const float L = radiance[((r * S + s) * n_bins) * 3 + e];
      // (clip(L * lobe) * weight) / denominator in the reference: weight / denominator folded, as in the static kernel
      ro += fminf(fmaxf(L * s_lobe[3 * s + ch], 0.f), rgb_max) * s_wd[s];
      ir += fminf(fmaxf(L * s_dl[s], 0.f), rgb_max) * s_wd[s];
    }
    radiance_out[r * n_bins * 3 + e] = ro * inv;
    if (irradiance) irradiance[r * n_bins * 3 + e] = ir * inv;
  }
  if (occ && occ_out && threadIdx.x == 0) {
    float oc = 0.f;
    for (int s = 0; s < S; ++s) oc += occ[r * S + s];
    occ_out[r] = oc * inv;
  }
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_ggx_integrate_fwd(void* stream, const float* d_wi, const float* d_wo,
                                         const float* d_radiance, const float* d_weight, const float* d_pdf,
                                         const float* d_occ, const float* d_albedo, const float* d_roughness,
                                         const float* d_metalness, const float* d_f0, int64_t num_points,
                                         int32_t num_samples, int32_t lobe_kind, float rgb_max,
                                         float* d_radiance_out, float* d_irradiance, float* d_occ_out) {
  if (num_points < 0 || num_samples < 1 || lobe_kind < 0 || lobe_kind > 3) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_wi || !d_wo || !d_radiance || !d_weight || !d_pdf || !d_albedo || !d_radiance_out)
    return NRC_E_INVALID_ARG;
  if (lobe_kind != kLambertian && (!d_roughness || !d_metalness || !d_f0)) return NRC_E_INVALID_ARG;
  unsigned grid = static_cast<unsigned>((num_points + kGgxWarps - 1) / kGgxWarps);
  ggx_integrate_fwd_kernel<<<grid, kGgxWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      d_wi, d_wo, d_radiance, d_weight, d_pdf, d_occ, d_albedo, d_roughness, d_metalness, d_f0, num_points,
      num_samples, lobe_kind, rgb_max, d_radiance_out, d_irradiance, d_occ_out);
  return check_launch();
}

extern "C" int32_t nrc_ggx_integrate_bwd(void* stream, const float* d_wi, const float* d_wo,
                                         const float* d_radiance, const float* d_weight, const float* d_pdf,
                                         const float* d_albedo, const float* d_roughness,
                                         const float* d_metalness, const float* d_f0, const float* d_g_out,
                                         const float* d_g_irradiance, int64_t num_points, int32_t num_samples,
                                         int32_t lobe_kind, float rgb_max, float* d_g_radiance) {
  if (num_points < 0 || num_samples < 1 || lobe_kind < 0 || lobe_kind > 3) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_wi || !d_wo || !d_radiance || !d_weight || !d_pdf || !d_albedo || !d_g_radiance)
    return NRC_E_INVALID_ARG;
  if (lobe_kind != kLambertian && (!d_roughness || !d_metalness || !d_f0)) return NRC_E_INVALID_ARG;
  unsigned grid = static_cast<unsigned>((num_points + kGgxWarps - 1) / kGgxWarps);
  ggx_integrate_bwd_radiance_kernel<<<grid, kGgxWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      d_wi, d_wo, d_radiance, d_weight, d_pdf, d_albedo, d_roughness, d_metalness, d_f0, d_g_out,
      d_g_irradiance, num_points, num_samples, lobe_kind, rgb_max, d_g_radiance);
  return check_launch();
}

extern "C" int32_t nrc_ggx_integrate_transient_fwd(void* stream, const float* d_wi, const float* d_wo, const float* d_radiance,
                                                   const float* d_weight, const float* d_pdf, const float* d_occ,
                                                   const float* d_albedo, const float* d_roughness, const float* d_metalness,
                                                   const float* d_f0, int64_t num_points, int32_t num_samples, int32_t n_bins,
                                                   int32_t lobe_kind, float rgb_max, float* d_radiance_out, float* d_irradiance,
                                                   float* d_occ_out) {
  if (num_points < 0 || num_samples < 1 || num_samples > 4096 || n_bins < 1 || lobe_kind < 0 || lobe_kind > 3)
    return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_wi || !d_wo || !d_radiance || !d_weight || !d_pdf || !d_albedo || !d_radiance_out) return NRC_E_INVALID_ARG;
  ggx_integrate_transient_fwd_kernel<<<static_cast<unsigned>(num_points), 256, static_cast<size_t>(num_samples) * 5 * sizeof(float),
                                       static_cast<cudaStream_t>(stream)>>>(
      d_wi, d_wo, d_radiance, d_weight, d_pdf, d_occ, d_albedo, d_roughness, d_metalness, d_f0, num_points, num_samples, n_bins,
      lobe_kind, rgb_max, d_radiance_out, d_irradiance, d_occ_out);
  return check_launch();
}
