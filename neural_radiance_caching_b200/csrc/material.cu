// Material stage, rows 18-19 of SURVEY 8a: importance sampling of secondary rays with the multiple
// importance sampling power heuristic, and the microfacet material head.
//   internal/inverse_render/render_utils.py: get_rotation_matrix (:145-168), global_to_local /
//   local_to_global (:698-710), CosineSampler (:417-444), MicrofacetSampler (:485-546), eval_vmf /
//   sample_vmf / LightSampler (:1335-1490), importance_sample_rays (:722-924), get_secondary_rays (:927-1056);
//   internal/material.py: _get_microfacet_material (:1290-1322, property table :957-1023).
// One thread per (shaded point, secondary sample): local frame, the sample's own sampler, the pdfs of
// every sampler of the set for the MIS weight, ray origin / direction.  Random draws are inputs.
#include "nrc_common.cuh"

namespace nrc {

constexpr float kDenEps = 1e-5f;   // DENOMINATOR_EPS (render_utils.py:41)
constexpr float kPi = 3.14159265358979323846f;

struct Frame {   // columns new_x, new_y, new_z = normal
  float x[3], y[3], z[3];
  __device__ __forceinline__ void build(float nx, float ny, float nz) {
    z[0] = nx; z[1] = ny; z[2] = nz;
    const bool use_z = fabsf(nz) < 0.9f;
    const float u0 = 0.f, u1 = use_z ? 0.f : 1.f, u2 = use_z ? 1.f : 0.f;
    // new_x = cross(up, normal)
    float ax = u1 * nz - u2 * ny, ay = u2 * nx - u0 * nz, az = u0 * ny - u1 * nx;
    float inv = 1.0f / (sqrtf(ax * ax + ay * ay + az * az) + 1e-10f);
    x[0] = ax * inv; x[1] = ay * inv; x[2] = az * inv;
    // new_y = cross(new_z, new_x)
    float bx = ny * x[2] - nz * x[1], by = nz * x[0] - nx * x[2], bz = nx * x[1] - ny * x[0];
    inv = 1.0f / (sqrtf(bx * bx + by * by + bz * bz) + 1e-10f);
    y[0] = bx * inv; y[1] = by * inv; y[2] = bz * inv;
  }
  __device__ __forceinline__ void to_local(const float d[3], float o[3]) const {
    o[0] = d[0] * x[0] + d[1] * x[1] + d[2] * x[2];
    o[1] = d[0] * y[0] + d[1] * y[1] + d[2] * y[2];
    o[2] = d[0] * z[0] + d[1] * z[1] + d[2] * z[2];
  }
  __device__ __forceinline__ void to_global(const float d[3], float o[3]) const {
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = d[0] * x[a] + d[1] * y[a] + d[2] * z[a];
  }
};

__device__ __forceinline__ float ggx_d(float c, float a) {
  const float t = c * c * (a * a - 1.f) + 1.f;
  return a * a / fmaxf(f32_eps(), kPi * t * t);
}

__device__ __forceinline__ float cosine_pdf(const float wi[3]) { return fmaxf(wi[2] < 0.f ? 0.f : wi[2] / kPi, 0.f); }

__device__ __forceinline__ float microfacet_pdf(const float wo[3], const float wi[3], float a) {
  float h[3] = {wo[0] + wi[0], wo[1] + wi[1], wo[2] + wi[2]};
  const float inv = 1.0f / sqrtf(1e-10f + h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);
  h[0] *= inv; h[1] *= inv; h[2] *= inv;
  const float dotp = wo[0] * h[0] + wo[1] * h[1] + wo[2] * h[2];
  const float pdf = ggx_d(h[2], a) * fabsf(h[2]) * (1.0f / fmaxf(4.0f * dotp, f32_eps()));
  return fmaxf(dotp <= 0.f ? 0.f : pdf, 0.f);
}

struct VmfMixture {
  const float* means;   // [K,3] (un-normalised)
  const float* kappas;  // [K]
  const float* logits;  // [K]
  int K;
  // l2_normalize forward value (ref_utils.py:45-70)
  __device__ __forceinline__ void mean(int k, float m[3]) const {
    const float a = means[3 * k], b = means[3 * k + 1], c = means[3 * k + 2];
    const float d = a * a + b * b + c * c;
    const float inv = 1.0f / sqrtf(fmaxf(f32_tiny(), d));
    const bool zero = d < f32_tiny();
    m[0] = zero ? 0.f : a * inv; m[1] = zero ? 0.f : b * inv; m[2] = zero ? 0.f : c * inv;
  }
  // sum_k softmax(logits)_k vmf(x; mean_k, kappa_k)   (render_utils.py:1335-1346, 1470-1490)
  __device__ float pdf(const float x[3]) const {
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, logits[k]);
    float den = 0.f, acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const float e = expf(logits[k] - mx);
      den += e;
      float m[3];
      mean(k, m);
      const float kp = kappas[k];
      float v;
      if (kp <= f32_eps()) v = 1.0f / (4.0f * kPi);
      else v = kp * safe_exp(kp * (x[0] * m[0] + x[1] * m[1] + x[2] * m[2])) / (4.0f * kPi * sinhf(kp));
      acc += e * v;
    }
    return fmaxf(acc / den, 0.f);
  }
};

// The same mixture with everything that does not depend on the query direction computed ONCE per shaded point and kept
// in shared memory (the point's S sample threads evaluate it 1-2 times each over all K lobes): per lobe the normalised
// mean, the exponent scale (kappa, or 0 for the uniform lobe) and coef = softmax(logits)_k kappa_k / (4 pi sinh kappa_k)
// (1 / (4 pi) for the uniform lobe), so that pdf(x) = sum_k coef_k safe_exp(kappa_k <x, mean_k>).
struct VmfTable {
  const float* t;   // [K][5]: mean (3), exponent scale, coef
  int K;
  __device__ __forceinline__ float pdf(const float x[3]) const {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const float* e = t + 5 * k;
      acc += e[4] * safe_exp(e[3] * (x[0] * e[0] + x[1] * e[1] + x[2] * e[2]));
    }
    return fmaxf(acc, 0.f);
  }
};

__device__ __forceinline__ void l2n3(float v[3]) {
  const float d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const float inv = 1.0f / sqrtf(fmaxf(f32_tiny(), d));
  const bool zero = d < f32_tiny();
#pragma unroll
  for (int a = 0; a < 3; ++a) v[a] = zero ? 0.f : v[a] * inv;
}

template <bool kTable>
__global__ void secondary_sample_kernel(const float* __restrict__ means, const float* __restrict__ viewdirs,
                                        const float* __restrict__ normals, const float* __restrict__ roughness,
                                        int64_t R, int n_micro, int n_cos, int n_light, const float* __restrict__ u,
                                        const float* __restrict__ vmf_means, const float* __restrict__ vmf_kappas,
                                        const float* __restrict__ vmf_logits, int K, const int32_t* __restrict__ latent,
                                        const float* __restrict__ normal2, float normal_eps,
                                        float* __restrict__ origins, float* __restrict__ dirs,
                                        float* __restrict__ local_lightdirs, float* __restrict__ local_viewdirs,
                                        float* __restrict__ pdf_out, float* __restrict__ weight_out) {
  const int S = n_micro + n_cos + n_light;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  extern __shared__ float s_tab[];          // kTable: [points per CTA][K][5]
  if constexpr (kTable) {
    // blockDim.x is a multiple of S (the launcher checks): this CTA owns whole points p0 .. p0 + ppc - 1
    const int ppc = blockDim.x / S;
    const int64_t p0 = static_cast<int64_t>(blockIdx.x) * ppc;
    // stage 1: per lobe - normalised mean, exponent scale, kappa / (4 pi sinh kappa); the logit waits in the coef slot
    for (int j = threadIdx.x; j < ppc * K; j += blockDim.x) {
      const int lp = j / K, k = j - lp * K;
      const int64_t pp = p0 + lp;
      float* e = s_tab + 5 * j;
      if (pp >= R) { e[0] = e[1] = e[2] = e[3] = 0.f; e[4] = -INFINITY; continue; }
      const float* mk = vmf_means + (pp * K + k) * 3;
      const float a0 = mk[0], a1 = mk[1], a2 = mk[2];
      const float d = a0 * a0 + a1 * a1 + a2 * a2;
      const float inv = 1.0f / sqrtf(fmaxf(f32_tiny(), d));
      const bool zero = d < f32_tiny();
      e[0] = zero ? 0.f : a0 * inv; e[1] = zero ? 0.f : a1 * inv; e[2] = zero ? 0.f : a2 * inv;
      const float kp = vmf_kappas[pp * K + k];
      e[3] = kp <= f32_eps() ? 0.f : kp;
      e[4] = vmf_logits[pp * K + k];
    }
    __syncthreads();
    // stage 2: softmax over the point's K logits (the point's S threads share the lobes), folded into coef
    {
      const int lp = threadIdx.x / S, sl = threadIdx.x - lp * S;
      float* tp = s_tab + 5 * lp * K;
      float mx = -INFINITY, den = 0.f;
      if (S == 32) {                         // one warp per point: lanes split the lobes, two shuffle reductions
        for (int k = sl; k < K; k += 32) mx = fmaxf(mx, tp[5 * k + 4]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        for (int k = sl; k < K; k += 32) den += expf(tp[5 * k + 4] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
      } else {
        for (int k = 0; k < K; ++k) mx = fmaxf(mx, tp[5 * k + 4]);
        for (int k = 0; k < K; ++k) den += expf(tp[5 * k + 4] - mx);
      }
      __syncthreads();                       // every thread has read the logits before they are overwritten
      for (int k = sl; k < K; k += S) {
        const float kp = tp[5 * k + 3];
        const float norm = kp == 0.f ? 1.0f / (4.0f * kPi) : kp / (4.0f * kPi * sinhf(kp));
        tp[5 * k + 4] = expf(tp[5 * k + 4] - mx) / den * norm;
      }
      __syncthreads();
    }
  }
  if (idx >= R * S) return;
  const int64_t p = idx / S;
  const int s = static_cast<int>(idx - p * S);
  const VmfTable tab{s_tab + 5 * (threadIdx.x / S) * K, K};
  Frame f;
  f.build(normals[3 * p], normals[3 * p + 1], normals[3 * p + 2]);
  const float gv[3] = {-viewdirs[3 * p], -viewdirs[3 * p + 1], -viewdirs[3 * p + 2]};   // global_viewdirs
  float wo[3];
  f.to_local(gv, wo);
  const float a = roughness[p];
  const float u1 = u[2 * idx], u2 = u[2 * idx + 1];
  VmfMixture mix{vmf_means ? vmf_means + p * K * 3 : nullptr, vmf_kappas ? vmf_kappas + p * K : nullptr,
                 vmf_logits ? vmf_logits + p * K : nullptr, K};
  float wi[3];
  float pdf;
  int own_count;
  if (s < n_micro) {
    own_count = n_micro;
    const float tan2 = a * a * u1 / fmaxf(1.0f - u1, f32_eps());
    const float cost = 1.0f / sqrtf(fmaxf(1.0f + tan2, f32_eps()));
    const float sint = sqrtf(fmaxf(kDenEps, 1.0f - cost * cost));
    const float phi = u2 * 2.0f * kPi - kPi;
    const float h[3] = {sint * cosf(phi), sint * sinf(phi), cost};
    const float npdf = fmaxf(ggx_d(cost, a) * fabsf(cost), 0.f);
    const float dotp = wo[0] * h[0] + wo[1] * h[1] + wo[2] * h[2];
    float d[3] = {2.f * dotp * h[0] - wo[0], 2.f * dotp * h[1] - wo[1], 2.f * dotp * h[2] - wo[2]};
    pdf = npdf * (1.0f / fmaxf(4.0f * dotp, f32_eps()));
    pdf = fmaxf(dotp <= 0.f ? 0.f : pdf, 0.f);
    const float inv = 1.0f / sqrtf(1e-10f + d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    wi[0] = d[0] * inv; wi[1] = d[1] * inv; wi[2] = d[2] * inv;
  } else if (s < n_micro + n_cos) {
    own_count = n_cos;
    const float r = sqrtf(u1);
    const float phi = u2 * 2.0f * kPi - kPi;
    wi[0] = r * cosf(phi); wi[1] = r * sinf(phi);
    wi[2] = sqrtf(fmaxf(kDenEps, 1.0f - wi[0] * wi[0] - wi[1] * wi[1]));
    pdf = fmaxf(wi[2] / kPi, 0.f);
  } else {
    own_count = n_light;
    const int ls = s - n_micro - n_cos;
    const int k = latent[p];
    float m[3];
    mix.mean(k, m);
    const float kp = mix.kappas[k];
    float t[3] = {-m[1], m[0], 0.f};
    l2n3(t);
    float b[3] = {m[1] * t[2] - m[2] * t[1], m[2] * t[0] - m[0] * t[2], m[0] * t[1] - m[1] * t[0]};
    l2n3(b);
    const float* n2 = normal2 + (p * n_light + ls) * 2;
    float v0 = n2[0], v1 = n2[1];
    {
      const float d = v0 * v0 + v1 * v1;
      const float inv = 1.0f / sqrtf(fmaxf(f32_tiny(), d));
      const bool zero = d < f32_tiny();
      v0 = zero ? 0.f : v0 * inv; v1 = zero ? 0.f : v1 * inv;
    }
    const float tmp = u1;   // the sampler's own uniform travels in the u1 slot of its samples
    const float w = 1.0f + (1.0f / fmaxf(kp, f32_eps())) * safe_log(tmp + (1.0f - tmp) * expf(-2.0f * kp));
    const float sq = sqrtf(fminf(fmaxf(1.0f - w * w, 0.f), f32_max()));
    float g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = t[c] * (sq * v0) + b[c] * (sq * v1) + m[c] * w;
    pdf = kTable ? tab.pdf(g) : mix.pdf(g);
    f.to_local(g, wi);    // global_dirs sampler: back to the local frame
  }
  // ---- MIS power heuristic over the samplers of the set (render_utils.py:817-853)
  float weight = 1.0f;
  const int n_sets = (n_micro > 0) + (n_cos > 0) + (n_light > 0);
  if (n_sets > 1) {
    float den = 0.f;
    if (n_micro > 0) { const float q = microfacet_pdf(wo, wi, a) * n_micro; den += q * q; }
    if (n_cos > 0) { const float q = cosine_pdf(wi) * n_cos; den += q * q; }
    if (n_light > 0) {
      float gl[3];
      f.to_global(wi, gl);
      const float q = (kTable ? tab.pdf(gl) : mix.pdf(gl)) * n_light;
      den += q * q;
    }
    den = fmaxf(den, kDenEps);
    const float q = own_count * pdf;
    weight = q * q / den * (static_cast<float>(S) / static_cast<float>(own_count));
  }
  float gl[3];
  f.to_global(wi, gl);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    origins[3 * idx + c] = means[3 * p + c] + normals[3 * p + c] * normal_eps;
    dirs[3 * idx + c] = gl[c];
    local_lightdirs[3 * idx + c] = wi[c];
  }
  if (s == 0) {
    local_viewdirs[3 * p] = wo[0]; local_viewdirs[3 * p + 1] = wo[1]; local_viewdirs[3 * p + 2] = wo[2];
  }
  pdf_out[idx] = pdf;
  weight_out[idx] = weight;
}

// brdf_params [P,ld] (10 raw channels) -> albedo [P,3], roughness [P], metalness [P], F_0 [P], specular_albedo [P]
__global__ void material_head_kernel(const float* __restrict__ raw, int64_t ld, int64_t P, float min_roughness,
                                     float default_f0, float* __restrict__ albedo, float* __restrict__ rough,
                                     float* __restrict__ metal, float* __restrict__ f0, float* __restrict__ spec_albedo) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* r = raw + p * ld;
  auto sig = [](float x) { return 1.f / (1.f + expf(-x)); };
#pragma unroll
  for (int c = 0; c < 3; ++c) albedo[3 * p + c] = sig(r[c] - 1.0f);
  const float mr2 = min_roughness * min_roughness;
  rough[p] = sig(r[6] - 1.0f) * (1.0f - mr2) + mr2;
  metal[p] = sig(r[8]);
  f0[p] = default_f0;
  if (spec_albedo) spec_albedo[p] = sig(r[5] - 1.0f);
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_secondary_sample(void* stream, const float* d_means, const float* d_viewdirs, const float* d_normals,
                                        const float* d_roughness, int64_t num_points, int32_t n_microfacet, int32_t n_cosine,
                                        int32_t n_light, const float* d_u, const float* d_vmf_means,
                                        const float* d_vmf_kappas, const float* d_vmf_logits, int32_t num_lobes,
                                        const int32_t* d_latent, const float* d_normal2, float normal_eps, float* d_origins,
                                        float* d_directions, float* d_local_lightdirs, float* d_local_viewdirs, float* d_pdf,
                                        float* d_weight) {
  const int S = n_microfacet + n_cosine + n_light;
  if (num_points < 0 || n_microfacet < 0 || n_cosine < 0 || n_light < 0 || S < 1) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_means || !d_viewdirs || !d_normals || !d_roughness || !d_u || !d_origins || !d_directions || !d_local_lightdirs ||
      !d_local_viewdirs || !d_pdf || !d_weight)
    return NRC_E_INVALID_ARG;
  if (n_light > 0 && (!d_vmf_means || !d_vmf_kappas || !d_vmf_logits || !d_latent || !d_normal2 || num_lobes < 1))
    return NRC_E_INVALID_ARG;
  const int64_t total = num_points * S;
  // whole points per CTA and a per-point lobe table in shared memory when the samples tile a 128-thread CTA (S = 32 in the
  // configured material stage: 4 points, 10 KB at 128 lobes); otherwise every thread walks the raw lobes itself
  const size_t tab_bytes = n_light > 0 ? static_cast<size_t>(128 / (S > 128 ? 128 : S)) * num_lobes * 5 * sizeof(float) : 0;
  if (n_light > 0 && S <= 128 && 128 % S == 0 && tab_bytes <= 48u * 1024u) {
    const int64_t ppc = 128 / S;
    secondary_sample_kernel<true><<<static_cast<unsigned>((num_points + ppc - 1) / ppc), 128, tab_bytes,
                                    static_cast<cudaStream_t>(stream)>>>(
        d_means, d_viewdirs, d_normals, d_roughness, num_points, n_microfacet, n_cosine, n_light, d_u, d_vmf_means,
        d_vmf_kappas, d_vmf_logits, num_lobes, d_latent, d_normal2, normal_eps, d_origins, d_directions, d_local_lightdirs,
        d_local_viewdirs, d_pdf, d_weight);
  } else {
    secondary_sample_kernel<false><<<static_cast<unsigned>((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_means, d_viewdirs, d_normals, d_roughness, num_points, n_microfacet, n_cosine, n_light, d_u, d_vmf_means,
        d_vmf_kappas, d_vmf_logits, num_lobes, d_latent, d_normal2, normal_eps, d_origins, d_directions, d_local_lightdirs,
        d_local_viewdirs, d_pdf, d_weight);
  }
  return check_launch();
}

extern "C" int32_t nrc_material_head(void* stream, const float* d_brdf_params, int64_t ld, int64_t num_points,
                                     float min_roughness, float default_f0, float* d_albedo, float* d_roughness,
                                     float* d_metalness, float* d_f0, float* d_specular_albedo) {
  if (num_points < 0 || ld < 10) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_brdf_params || !d_albedo || !d_roughness || !d_metalness || !d_f0) return NRC_E_INVALID_ARG;
  material_head_kernel<<<static_cast<unsigned>((num_points + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_brdf_params, ld, num_points, min_roughness, default_f0, d_albedo, d_roughness, d_metalness, d_f0,
      d_specular_albedo);
  return check_launch();
}
