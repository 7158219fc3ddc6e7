// Light sampler: the vMF head of LightMLP and the light-sampling loss.
//
//   vmf head  LightMLP.get_vmfs + the recentring in predict_lighting (internal/light_sampler.py:135-160,203-204):
//       means  = raw[0:3] * vmf_scale + means_random - position
//       kappas = min(softplus(raw[3] + 1), 50)
//       logits = max(raw[4] + 1, -50)
//     thread per (point, lobe); the VJP maps (g_means, g_kappas, g_logits) back to g_raw.
//   vmf loss  render_utils.vmf_loss_fn (internal/inverse_render/render_utils.py:1493-1550) as called by
//     train_utils.light_sampling_loss (internal/train_utils.py:1985-2071): per point a K-lobe mixture likelihood at
//     each of its S secondary-sample directions,
//       L_s = sum_k safe_exp(logit_k) vmf(d_s; l2_normalize(mean_k), kappa_k)
//       loss += mean_{p,s} (f_s - L_s) stop_grad(f_s - L_s) clip(w_s, 0, 10) [d_s.n > 0] lossmult / max(pdf_s, 1e-2)
//     (f and L through linear_to_srgb(max(., 1e-5)) when `srgb`), with gradients w.r.t. means / kappas / logits.
//     One CTA per point, one thread per lobe: the lobes' values at every sample stay in registers between the
//     likelihood pass and the gradient pass.
#include "nrc_common.cuh"

namespace nrc {

constexpr int kMaxLobes = 128;
constexpr int kSampleChunk = 32;
constexpr float kFourPi = 12.566370614359172f;

__global__ void vmf_head_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ means_random,
                                    int64_t mr_point_stride, const float* __restrict__ pos, int64_t P, int K,
                                    float vmf_scale, float* __restrict__ means, float* __restrict__ kappas,
                                    float* __restrict__ logits) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= P * K) return;
  const int64_t p = i / K;
  const int k = static_cast<int>(i - p * K);
  const float* r = raw + i * 5;
  const float* mr = means_random + p * mr_point_stride + 3 * k;
#pragma unroll
  for (int a = 0; a < 3; ++a) means[3 * i + a] = r[a] * vmf_scale + mr[a] - pos[3 * p + a];
  const float x = r[3] + 1.0f;
  const float sp = x > 20.f ? x : log1pf(expf(x));          // jax.nn.softplus = logaddexp(x, 0)
  kappas[i] = fminf(sp, 50.0f);
  logits[i] = fmaxf(r[4] + 1.0f, -50.0f);
}

__global__ void vmf_head_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ g_means,
                                    const float* __restrict__ g_kappas, const float* __restrict__ g_logits, int64_t N,
                                    float vmf_scale, float* __restrict__ g_raw) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* r = raw + i * 5;
  float* g = g_raw + i * 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) g[a] = g_means[3 * i + a] * vmf_scale;
  const float x = r[3] + 1.0f;
  const float sp = x > 20.f ? x : log1pf(expf(x));
  const float sig = 1.0f / (1.0f + expf(-x));
  g[3] = sp < 50.0f ? g_kappas[i] * sig : 0.f;               // minimum(): gradient to the smaller argument
  g[4] = (r[4] + 1.0f) > -50.0f ? g_logits[i] : 0.f;
}

__device__ __forceinline__ float srgb_fwd(float x, float& d) {
  // image.linear_to_srgb (internal/image.py:192-200) of max(x, 1e-5) and its derivative w.r.t. x
  const bool clipped = x < 1e-5f;
  const float xc = fmaxf(x, 1e-5f);
  const float eps = f32_eps();
  const float xe = fmaxf(xc, eps);
  const float p512 = powf(xe, 5.0f / 12.0f);
  const bool lin = xc <= 0.0031308f;
  const float y = lin ? (323.0f / 25.0f) * xc : (211.0f * p512 - 11.0f) / 200.0f;
  const float dy = lin ? (323.0f / 25.0f) : (211.0f / 200.0f) * (5.0f / 12.0f) * p512 / xe;
  d = clipped ? 0.f : dy;
  return y;
}

__global__ void __launch_bounds__(kMaxLobes)
vmf_loss_kernel(const float* __restrict__ means, const float* __restrict__ kappas, const float* __restrict__ logits,
                const float* __restrict__ normals, const float* __restrict__ dirs, const float* __restrict__ pdf,
                const float* __restrict__ weight, const float* __restrict__ fvals, int64_t P, int K, int S, float lossmult,
                int srgb, float* __restrict__ loss, float* __restrict__ g_means, float* __restrict__ g_kappas,
                float* __restrict__ g_logits) {
  __shared__ float sd[kSampleChunk][3], sL[kSampleChunk], sc[kSampleChunk], sred[4];
  const int64_t p = blockIdx.x;
  const int k = threadIdx.x;
  const bool live = k < K;
  // this thread's lobe
  float mx = 0.f, my = 0.f, mz = 0.f, kap = 0.f, wk = 0.f, q = 0.f;
  if (live) {
    const int64_t i = p * K + k;
    mx = means[3 * i]; my = means[3 * i + 1]; mz = means[3 * i + 2];
    kap = kappas[i];
    wk = safe_exp(logits[i]);
    q = mx * mx + my * my + mz * mz;
  }
  const float inv_n = q < f32_tiny() ? 0.f : 1.0f / sqrtf(fmaxf(q, f32_tiny()));   // l2_normalize forward
  const float nx = mx * inv_n, ny = my * inv_n, nz = mz * inv_n;
  const bool uniform = kap <= f32_eps();
  const float norm = uniform ? 0.f : kap / (kFourPi * sinhf(kap));
  // langevin = coth(kappa) - 1/kappa; the closed form cancels catastrophically for small kappa: series below 0.5
  float langevin = 0.f;
  if (!uniform) {
    if (kap < 0.5f) {
      const float k2 = kap * kap;
      langevin = kap * (1.0f / 3.0f + k2 * (-1.0f / 45.0f + k2 * (2.0f / 945.0f - k2 * (1.0f / 4725.0f))));
    } else {
      langevin = coshf(kap) / sinhf(kap) - 1.0f / kap;
    }
  }
  const float n0 = normals[3 * p], n1 = normals[3 * p + 1], n2 = normals[3 * p + 2];
  const float inv_count = 1.0f / (static_cast<float>(P) * static_cast<float>(S));
  float g_w = 0.f, g_k = 0.f, gnx = 0.f, gny = 0.f, gnz = 0.f, loss_acc = 0.f;
  for (int s0 = 0; s0 < S; s0 += kSampleChunk) {
    const int ns = min(kSampleChunk, S - s0);
    __syncthreads();
    if (k < ns) {
      const float* d = dirs + (p * S + s0 + k) * 3;
      sd[k][0] = d[0]; sd[k][1] = d[1]; sd[k][2] = d[2];
      sL[k] = 0.f;
    }
    __syncthreads();
    float v[kSampleChunk];
#pragma unroll
    for (int s = 0; s < kSampleChunk; ++s) {
      v[s] = 0.f;
      if (s < ns) {
        const float t = sd[s][0] * nx + sd[s][1] * ny + sd[s][2] * nz;
        float val = uniform ? 1.0f / kFourPi : norm * safe_exp(kap * t);
        if (!live) val = 0.f;
        v[s] = val;
        float part = wk * val;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((k & 31) == 0) atomicAdd(&sL[s], part);
      }
    }
    __syncthreads();
    if (k < ns) {   // coefficient c_s = d term_s / d L_s and the sample's loss term
      const int64_t j = p * S + s0 + k;
      float f = fvals[j], L = sL[k], dL = 1.f, df;
      if (srgb) { f = srgb_fwd(f, df); L = srgb_fwd(L, dL); }
      const float dot = sd[k][0] * n0 + sd[k][1] * n1 + sd[k][2] * n2;
      const float wgt = dot > 0.f ? fminf(fmaxf(weight[j], 0.f), 10.f) : 0.f;
      const float c = wgt * lossmult / fmaxf(pdf[j], 1e-2f) * inv_count;
      const float diff = f - L;
      loss_acc += diff * diff * c;
      sc[k] = -diff * c * dL;      // d/dL of (f - L) * stop_grad(f - L) * c
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int s = 0; s < kSampleChunk; ++s) {
        if (s < ns) {
          const float c = sc[s];
          g_w = fmaf(c, v[s], g_w);
          if (!uniform) {
            const float t = sd[s][0] * nx + sd[s][1] * ny + sd[s][2] * nz;
            const float cw = c * wk * v[s];
            g_k = fmaf(cw, t - langevin, g_k);
            const float ck = cw * kap;
            gnx = fmaf(ck, sd[s][0], gnx); gny = fmaf(ck, sd[s][1], gny); gnz = fmaf(ck, sd[s][2], gnz);
          }
        }
      }
    }
  }
  if (live) {
    const int64_t i = p * K + k;
    g_logits[i] = g_w * wk;                     // d safe_exp(logit) = exp(clip(logit)) (internal/math.py:186-192)
    g_kappas[i] = g_k;
    // l2_normalize VJP with grad_eps = 1e-5 (internal/ref_utils.py:45-70)
    float ox, oy, oz;
    if (q < f32_tiny()) {
      ox = oy = oz = 0.f;
    } else if (q > 1e-5f) {
      const float inv = 1.0f / sqrtf(q);
      const float ux = mx * inv, uy = my * inv, uz = mz * inv;
      const float dot = ux * gnx + uy * gny + uz * gnz;
      ox = (gnx - ux * dot) * inv; oy = (gny - uy * dot) * inv; oz = (gnz - uz * dot) * inv;
    } else {
      const float inv = 1.0f / sqrtf(1e-5f);
      ox = gnx * inv; oy = gny * inv; oz = gnz * inv;
    }
    g_means[3 * i] = ox; g_means[3 * i + 1] = oy; g_means[3 * i + 2] = oz;
  }
  // block-reduce the loss terms (held by the first `ns` threads of each chunk)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
  __syncthreads();
  if ((k & 31) == 0) sred[k >> 5] = loss_acc;
  __syncthreads();
  if (k == 0) atomicAdd(loss, sred[0] + sred[1] + sred[2] + sred[3]);
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_vmf_head_fwd(void* stream, const float* d_raw, const float* d_means_random,
                                    int32_t means_random_per_point, const float* d_positions, int64_t num_points,
                                    int32_t num_lobes, float vmf_scale, float* d_means, float* d_kappas, float* d_logits) {
  if (num_points < 0 || num_lobes < 1) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_raw || !d_means_random || !d_positions || !d_means || !d_kappas || !d_logits) return NRC_E_INVALID_ARG;
  const int64_t n = num_points * num_lobes;
  vmf_head_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_raw, d_means_random, means_random_per_point ? static_cast<int64_t>(num_lobes) * 3 : 0, d_positions, num_points,
      num_lobes, vmf_scale, d_means, d_kappas, d_logits);
  return check_launch();
}

extern "C" int32_t nrc_vmf_head_bwd(void* stream, const float* d_raw, const float* d_g_means, const float* d_g_kappas,
                                    const float* d_g_logits, int64_t num_points, int32_t num_lobes, float vmf_scale,
                                    float* d_g_raw) {
  if (num_points < 0 || num_lobes < 1) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_raw || !d_g_means || !d_g_kappas || !d_g_logits || !d_g_raw) return NRC_E_INVALID_ARG;
  const int64_t n = num_points * num_lobes;
  vmf_head_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_raw, d_g_means, d_g_kappas, d_g_logits, n, vmf_scale, d_g_raw);
  return check_launch();
}

extern "C" int32_t nrc_vmf_loss(void* stream, const float* d_means, const float* d_kappas, const float* d_logits,
                                const float* d_normals, const float* d_dirs, const float* d_pdf, const float* d_weight,
                                const float* d_function_vals, int64_t num_points, int32_t num_lobes, int32_t num_samples,
                                float lossmult, int32_t linear_to_srgb, float* d_loss, float* d_g_means,
                                float* d_g_kappas, float* d_g_logits) {
  if (num_points < 1 || num_lobes < 1 || num_lobes > kMaxLobes || num_samples < 1) return NRC_E_INVALID_ARG;
  if (!d_means || !d_kappas || !d_logits || !d_normals || !d_dirs || !d_pdf || !d_weight || !d_function_vals || !d_loss ||
      !d_g_means || !d_g_kappas || !d_g_logits)
    return NRC_E_INVALID_ARG;
  vmf_loss_kernel<<<static_cast<unsigned>(num_points), kMaxLobes, 0, static_cast<cudaStream_t>(stream)>>>(
      d_means, d_kappas, d_logits, d_normals, d_dirs, d_pdf, d_weight, d_function_vals, num_points, num_lobes,
      num_samples, lossmult, linear_to_srgb, d_loss, d_g_means, d_g_kappas, d_g_logits);
  return check_launch();
}
