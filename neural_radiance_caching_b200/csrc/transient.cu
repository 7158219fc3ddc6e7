// Time-resolved (transient) cache rendering, SURVEY 8a row 22 / BASELINE config 4:
//   internal/nerf.py:1660-1777 (_compute_indirect_lighting / get_indirect: softplus head, indirect_scale, clip),
//   internal/inverse_render/render_utils.py:1699-1767 (zero_invalid_bins),
//   internal/render.py:250-449 (volumetric_transient_rendering), :452-490 (shift_direct), :493-507
//   (shift_map_coordinates = order-1 map_coordinates along the bin axis, mode='constant').
// The reference materialises [R, n, n_bins, 3] tensors three to four times (activation, masks, shift)
// before reducing over the n samples (275 MB each at R = 1024, n_bins = 700).  Here the raw head outputs are
// read ONCE: activation, validity masks, clip, the per-sample sub-bin shift and the weighted reduction over
// samples happen in registers; only [R, n_bins, 3] leaves the SM.
#include <cuda_bf16.h>

#include "nrc_common.cuh"

namespace nrc {

struct TransientParams {
  int n, n_bins, C;
  float exposure_time, shift, diffuse_bias, indirect_scale, bin_zero_threshold_light, light_near, rgb_max, dark_level;
  int light_zero;
};

// shift_direct (render.py:452-490): every sample splats weights*direct into bins floor(d) and ceil(d) of a
// FLAT [R*n_bins] histogram (index = ray*n_bins + bin, exactly like the reference's .at[].add on the flattened
// array: a bin >= n_bins lands in the next ray's histogram, indices outside the array are dropped).
__global__ void transient_direct_kernel(const float* __restrict__ direct, const float* __restrict__ weights,
                                        const float* __restrict__ ray_dists, const float* __restrict__ light_dists,
                                        int64_t R, TransientParams p, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * p.n) return;
  const int64_t ray = i / p.n;
  const float d = (light_dists[i] + ray_dists[i]) / p.exposure_time + p.shift / p.exposure_time;
  const float lo = fmaxf(floorf(d), 0.f), hi = ceilf(d);
  const float w_hi = d - lo, w_lo = 1.0f - w_hi;
  const int64_t total = R * p.n_bins;
  const int64_t i_lo = ray * p.n_bins + static_cast<int64_t>(static_cast<int32_t>(lo));
  const int64_t i_hi = ray * p.n_bins + static_cast<int64_t>(static_cast<int32_t>(hi));
  const float w = weights[i];
  for (int c = 0; c < p.C; ++c) {
    const float v = w * direct[i * p.C + c];
    if (i_lo >= 0 && i_lo < total) atomicAdd(out + i_lo * p.C + c, v * w_lo);
    if (i_hi >= 0 && i_hi < total) atomicAdd(out + i_hi * p.C + c, v * w_hi);
  }
}

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// One CTA per ray; thread e owns output element (bin, channel) = (e / C, e % C).
__global__ void transient_indirect_kernel(const float* __restrict__ diffuse_raw, const float* __restrict__ specular,
                                          const float* __restrict__ spec_scale, const float* __restrict__ weights,
                                          const float* __restrict__ ray_dists, const float* __restrict__ light_dists,
                                          const float* __restrict__ cam_dists, int64_t R, TransientParams p,
                                          const float* __restrict__ t_direct, float* __restrict__ t_indirect,
                                          float* __restrict__ rgb) {
  const int64_t ray = blockIdx.x;
  const int BC = p.n_bins * p.C;
  const float max_dists = static_cast<float>(p.n_bins - 1) * p.exposure_time;
  extern __shared__ float sm[];   // per sample: weight, bins_move, light_dist, cam_dist, spec_scale[C]
  float* s_w = sm;
  float* s_move = sm + p.n;
  float* s_light = sm + 2 * p.n;
  float* s_cam = sm + 3 * p.n;
  float* s_scale = sm + 4 * p.n;
  for (int s = threadIdx.x; s < p.n; s += blockDim.x) {
    const int64_t i = ray * p.n + s;
    s_w[s] = weights[i];
    s_move[s] = (ray_dists[i] + p.shift) / p.exposure_time;
    s_light[s] = light_dists[i];
    s_cam[s] = cam_dists[i];
    for (int c = 0; c < p.C; ++c) s_scale[s * p.C + c] = spec_scale ? spec_scale[i * p.C + c] : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < BC; e += blockDim.x) {
    const int b = e / p.C, c = e - b * p.C;
    float acc = 0.f;
    for (int s = 0; s < p.n; ++s) {
      // map_coordinates(order=1, mode='constant'): value at y = b - bins_move between bins y0 and y0+1
      const float y = static_cast<float>(b) - s_move[s];
      const float y0f = floorf(y);
      const float t = y - y0f;
      const int y0 = static_cast<int>(y0f);
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int bin = y0 + k;
        const float wk = k ? t : 1.0f - t;
        if (bin < 0 || bin >= p.n_bins || wk == 0.f) continue;
        // zero_invalid_bins (render_utils.py:1699-1767)
        const float fb = static_cast<float>(bin);
        bool ok = !((fb + p.bin_zero_threshold_light) * p.exposure_time < s_light[s]);
        ok = ok && !((fb * p.exposure_time + s_cam[s]) > max_dists);
        if (p.light_zero) ok = ok && !(s_light[s] < p.light_near);
        if (!ok) continue;
        const int64_t idx = ((ray * p.n + s) * p.n_bins + bin) * p.C + c;
        float val = 0.f;
        if (diffuse_raw) val += fminf(fmaxf(softplus_t(diffuse_raw[idx] + p.diffuse_bias) * p.indirect_scale, 0.f), p.rgb_max);
        if (specular) val += fminf(fmaxf(s_scale[s * p.C + c] * specular[idx] * p.indirect_scale, 0.f), p.rgb_max);
        v += wk * val;
      }
      acc += s_w[s] * v;
    }
    const int64_t o = ray * BC + e;
    t_indirect[o] = acc;
    rgb[o] = t_direct[o] + acc + p.dark_level;
  }
}


// ------------------------------------------------------------------------------------------------
// The transient heads FUSED with the time-resolved integration (BASELINE config 4): the reference materialises the
// per-sample histograms [R, n, n_bins, 3] of the diffuse head (irradiance stack 64 -> n_bins * 3, nerf.py:1757-1777) and
// of the transient SurfaceLightField (128 -> n_bins * 3 + 1, surface_light_field.py:1033-1041) - 275 MB each at R = 1024,
// n = 32, 700 bins - and then activates, masks, shifts and reduces them (render.py:250-449).  Here one CTA owns one ray:
// the last layer of both heads runs on the tensor cores (mma.sync m16n8k16, bf16 operands, fp32 accumulate) 16 samples
// at a time, the activated and masked values of those 16 samples are staged in shared memory in fp32 (134 KB), and the
// sub-bin shift + weighted reduction over the samples reads them from there in the order of
// transient_indirect_kernel (same taps, same summation order).  Only [R, n_bins, 3] leaves the SM.
constexpr int kTrhThreads = 512;
constexpr int kTrhRows = 16;          // samples per pass (one MMA row tile)

struct TrHeadParams {
  int n, n_bins, C, Kd, Ks;
  int64_t ld_wd, ld_ws;
  float exposure_time, shift, diffuse_bias, spec_bias, spec_premult, spec_max, indirect_scale, bin_zero_threshold_light,
      light_near, rgb_max, dark_level;
  int light_zero;
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// One head's last layer for the 16 staged samples: stage[row][e] (+)= activate(h[row] . W[:, e] + b[e]).
// h_s: [16][K + 8] bf16 in shared memory; W: Flax kernel [K][ld] fp32 in global memory (L2 resident: every CTA reads it).
template <int K, bool SPECULAR>
__device__ __forceinline__ void head_pass(const TrHeadParams& p, const __nv_bfloat16* __restrict__ h_s, const float* __restrict__ W,
                                          int64_t ldw, const float* __restrict__ bias, int row0, const float* s_light,
                                          const float* s_cam, const float* s_scale, float* __restrict__ stage, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr int KS = K / 16, HS = K + 8;
  uint32_t a[KS][4];
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    const __nv_bfloat16* r0 = h_s + g * HS + kk * 16 + 2 * t;
    const __nv_bfloat16* r1 = h_s + (g + 8) * HS + kk * 16 + 2 * t;
    a[kk][0] = *reinterpret_cast<const uint32_t*>(r0);
    a[kk][1] = *reinterpret_cast<const uint32_t*>(r1);
    a[kk][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
    a[kk][3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
  }
  const float max_dists = static_cast<float>(p.n_bins - 1) * p.exposure_time;
  const int n_tiles = (N + 7) >> 3;
  for (int tile = warp; tile < n_tiles; tile += kTrhThreads / 32) {
    const int nb = tile * 8 + g;                 // output column this thread's B fragment feeds
    const bool nb_ok = nb < N;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const float* w = W + static_cast<int64_t>(kk * 16 + 2 * t) * ldw + nb;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
      if (nb_ok) { w0 = __ldg(w); w1 = __ldg(w + ldw); w2 = __ldg(w + 8 * ldw); w3 = __ldg(w + 9 * ldw); }
      mma_bf16_16816(acc, a[kk], pack_bf16x2(w0, w1), pack_bf16x2(w2, w3));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = g + (i >= 2 ? 8 : 0), e = tile * 8 + 2 * t + (i & 1);
      if (e >= N) continue;
      const int s = row0 + row;
      float val = 0.f;
      if (s < p.n) {
        const int bin = e / p.C, c = e - bin * p.C;
        // zero_invalid_bins (render_utils.py:1699-1767)
        const float fb = static_cast<float>(bin);
        bool ok = !((fb + p.bin_zero_threshold_light) * p.exposure_time < s_light[s]);
        ok = ok && !((fb * p.exposure_time + s_cam[s]) > max_dists);
        if (p.light_zero) ok = ok && !(s_light[s] < p.light_near);
        if (ok) {
          const float raw = acc[i] + __ldg(bias + e);
          if (!SPECULAR) {
            val = fminf(fmaxf(softplus_t(raw + p.diffuse_bias) * p.indirect_scale, 0.f), p.rgb_max);
          } else {
            const float ref = fminf(fmaxf(softplus_t(p.spec_premult * raw + p.spec_bias), 0.f), p.spec_max);
            val = fminf(fmaxf(s_scale[s * p.C + c] * ref * p.indirect_scale, 0.f), p.rgb_max);
          }
        }
      }
      float* dst = stage + static_cast<size_t>(row) * N + e;
      if (SPECULAR) *dst += val; else *dst = val;      // the same thread owns (row, e) in both heads
    }
  }
}

template <int KD, int KS_>
__global__ void __launch_bounds__(kTrhThreads, 1)
transient_head_render_kernel(const float* __restrict__ h_d, const float* __restrict__ w_d, const float* __restrict__ b_d,
                             const float* __restrict__ h_s, const float* __restrict__ w_s, const float* __restrict__ b_s,
                             const float* __restrict__ spec_scale, const float* __restrict__ weights,
                             const float* __restrict__ ray_dists, const float* __restrict__ light_dists,
                             const float* __restrict__ cam_dists, int64_t R, TrHeadParams p, const float* __restrict__ t_direct,
                             float* __restrict__ t_indirect, float* __restrict__ rgb) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int64_t ray = blockIdx.x;
  const int N = p.n_bins * p.C;
  float* stage = reinterpret_cast<float*>(smem);                               // [16][N]
  float* s_w = stage + static_cast<size_t>(kTrhRows) * N;                      // per sample: weight, bins_move, light, cam
  float* s_move = s_w + p.n;
  float* s_light = s_move + p.n;
  float* s_cam = s_light + p.n;
  float* s_scale = s_cam + p.n;                                                // [n][C]
  __nv_bfloat16* hd_s = reinterpret_cast<__nv_bfloat16*>(s_scale + p.n * p.C + ((p.n * p.C) & 1));
  __nv_bfloat16* hs_s = hd_s + kTrhRows * (KD + 8);
  for (int s = threadIdx.x; s < p.n; s += blockDim.x) {
    const int64_t i = ray * p.n + s;
    s_w[s] = weights[i];
    s_move[s] = (ray_dists[i] + p.shift) / p.exposure_time;
    s_light[s] = light_dists[i];
    s_cam[s] = cam_dists[i];
    for (int c = 0; c < p.C; ++c) s_scale[s * p.C + c] = spec_scale ? spec_scale[i * p.C + c] : 0.f;
  }
  constexpr int kMaxPer = 8;    // output elements per thread (N <= 8 * 512)
  float out[kMaxPer];
#pragma unroll
  for (int q = 0; q < kMaxPer; ++q) out[q] = 0.f;
  for (int row0 = 0; row0 < p.n; row0 += kTrhRows) {
    __syncthreads();            // the previous pass's gather is done with `stage`; the per-sample arrays are written
    for (int idx = threadIdx.x; idx < kTrhRows * KD; idx += blockDim.x) {
      const int r = idx / KD, k = idx - r * KD, s = row0 + r;
      hd_s[r * (KD + 8) + k] = __float2bfloat16_rn((h_d && s < p.n) ? h_d[(ray * p.n + s) * KD + k] : 0.f);
    }
    if (h_s)
      for (int idx = threadIdx.x; idx < kTrhRows * KS_; idx += blockDim.x) {
        const int r = idx / KS_, k = idx - r * KS_, s = row0 + r;
        hs_s[r * (KS_ + 8) + k] = __float2bfloat16_rn(s < p.n ? h_s[(ray * p.n + s) * KS_ + k] : 0.f);
      }
    __syncthreads();
    if (h_d) head_pass<KD, false>(p, hd_s, w_d, p.ld_wd, b_d, row0, s_light, s_cam, s_scale, stage, N);
    else
      for (int idx = threadIdx.x; idx < kTrhRows * N; idx += blockDim.x) stage[idx] = 0.f;
    if (h_s) {
      if (!h_d) __syncthreads();
      head_pass<KS_, true>(p, hs_s, w_s, p.ld_ws, b_s, row0, s_light, s_cam, s_scale, stage, N);
    }
    __syncthreads();
    // shift_map_coordinates (order-1 map_coordinates along the bin axis, mode 'constant') + the weighted reduction,
    // samples in ascending order like transient_indirect_kernel
#pragma unroll
    for (int q = 0; q < kMaxPer; ++q) {
      const int e = threadIdx.x + q * kTrhThreads;
      if (e >= N) continue;
      const int b = e / p.C, c = e - b * p.C;
      float acc = out[q];
      const int s_end = min(kTrhRows, p.n - row0);
      for (int r = 0; r < s_end; ++r) {
        const int s = row0 + r;
        const float y = static_cast<float>(b) - s_move[s];
        const float y0f = floorf(y);
        const float tt = y - y0f;
        const int y0 = static_cast<int>(y0f);
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int bin = y0 + k;
          const float wk = k ? tt : 1.0f - tt;
          if (bin < 0 || bin >= p.n_bins || wk == 0.f) continue;
          v += wk * stage[static_cast<size_t>(r) * N + bin * p.C + c];
        }
        acc += s_w[s] * v;
      }
      out[q] = acc;
    }
  }
#pragma unroll
  for (int q = 0; q < kMaxPer; ++q) {
    const int e = threadIdx.x + q * kTrhThreads;
    if (e >= N) continue;
    const int64_t o = ray * N + e;
    t_indirect[o] = out[q];
    rgb[o] = t_direct[o] + out[q] + p.dark_level;
  }
}

// Temporal filter of volumetric_transient_rendering (internal/render.py:397-415): jax.scipy.signal.convolve(x,
// filter[None, :, None], mode='same') along the bin axis; `filt` is the impulse response or the normalised Gaussian the
// reference builds from tfilter_sigma.  One thread per output element.
__global__ void transient_filter_kernel(const float* __restrict__ x, const float* __restrict__ filt, int taps, int64_t R,
                                        int n_bins, int C, float* __restrict__ y) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * n_bins * C) return;
  const int c = static_cast<int>(i % C);
  const int b = static_cast<int>((i / C) % n_bins);
  const int64_t ray = i / (static_cast<int64_t>(C) * n_bins);
  // full convolution z[m] = sum_k x[m - k] f[k], 'same' keeps m in [(taps - 1) / 2, (taps - 1) / 2 + n_bins)
  const int m = b + (taps - 1) / 2;
  float acc = 0.f;
  for (int k = 0; k < taps; ++k) {
    const int j = m - k;
    if (j >= 0 && j < n_bins) acc += x[(ray * n_bins + j) * C + c] * filt[k];
  }
  y[i] = acc;
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_transient_render_fwd(void* stream, const float* d_direct_rgbs, const float* d_diffuse_raw,
                                            const float* d_specular, const float* d_spec_scale, const float* d_weights,
                                            const float* d_ray_dists, const float* d_light_dists, const float* d_cam_dists,
                                            int64_t num_rays, int32_t n, int32_t n_bins, int32_t channels,
                                            float exposure_time, float shift, float diffuse_bias, float indirect_scale,
                                            float bin_zero_threshold_light, int32_t light_zero, float light_near,
                                            float rgb_max, float dark_level, float* d_transient_direct,
                                            float* d_transient_indirect, float* d_rgb) {
  if (num_rays < 0 || n < 1 || n > 1024 || n_bins < 1 || channels < 1 || channels > 4 || !(exposure_time > 0.f))
    return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_direct_rgbs || !d_weights || !d_ray_dists || !d_light_dists || !d_cam_dists || !d_transient_direct ||
      !d_transient_indirect || !d_rgb || (d_specular && !d_spec_scale))
    return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TransientParams p{n, n_bins, channels, exposure_time, shift, diffuse_bias, indirect_scale, bin_zero_threshold_light,
                    light_near, rgb_max, dark_level, light_zero};
  const size_t bytes = static_cast<size_t>(num_rays) * n_bins * channels * sizeof(float);
  if (cudaMemsetAsync(d_transient_direct, 0, bytes, s) != cudaSuccess) return check_launch();
  const int64_t tot = num_rays * n;
  transient_direct_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(d_direct_rgbs, d_weights, d_ray_dists,
                                                                                   d_light_dists, num_rays, p,
                                                                                   d_transient_direct);
  const size_t smem = static_cast<size_t>(n) * (4 + channels) * sizeof(float);
  transient_indirect_kernel<<<static_cast<unsigned>(num_rays), 256, smem, s>>>(
      d_diffuse_raw, d_specular, d_spec_scale, d_weights, d_ray_dists, d_light_dists, d_cam_dists, num_rays, p,
      d_transient_direct, d_transient_indirect, d_rgb);
  return check_launch();
}

extern "C" int32_t nrc_transient_head_render_fwd(
    void* stream, const float* d_direct_rgbs, const float* d_h_diffuse, int32_t k_diffuse, const float* d_w_diffuse,
    int64_t ld_w_diffuse, const float* d_b_diffuse, const float* d_h_specular, int32_t k_specular, const float* d_w_specular,
    int64_t ld_w_specular, const float* d_b_specular, const float* d_spec_scale, const float* d_weights, const float* d_ray_dists,
    const float* d_light_dists, const float* d_cam_dists, int64_t num_rays, int32_t n, int32_t n_bins, int32_t channels,
    float exposure_time, float shift, float diffuse_bias, float spec_premult, float spec_bias, float spec_max,
    float indirect_scale, float bin_zero_threshold_light, int32_t light_zero, float light_near, float rgb_max, float dark_level,
    float* d_transient_direct, float* d_transient_indirect, float* d_rgb) {
  if (num_rays < 0 || n < 1 || n > 1024 || n_bins < 1 || channels < 1 || channels > 4 || !(exposure_time > 0.f))
    return NRC_E_INVALID_ARG;
  const int64_t N = static_cast<int64_t>(n_bins) * channels;
  if (N > 8 * kTrhThreads) return NRC_E_UNSUPPORTED;
  if ((d_h_diffuse && k_diffuse != 64) || (d_h_specular && k_specular != 128)) return NRC_E_UNSUPPORTED;
  if (num_rays == 0) return NRC_OK;
  if (!d_direct_rgbs || !d_weights || !d_ray_dists || !d_light_dists || !d_cam_dists || !d_transient_direct ||
      !d_transient_indirect || !d_rgb || (!d_h_diffuse && !d_h_specular) || (d_h_diffuse && (!d_w_diffuse || !d_b_diffuse)) ||
      (d_h_specular && (!d_w_specular || !d_b_specular || !d_spec_scale)) || ld_w_diffuse < (d_h_diffuse ? N : 0) ||
      ld_w_specular < (d_h_specular ? N : 0))
    return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TransientParams dp{n, n_bins, channels, exposure_time, shift, diffuse_bias, indirect_scale, bin_zero_threshold_light,
                     light_near, rgb_max, dark_level, light_zero};
  const size_t bytes = static_cast<size_t>(num_rays) * N * sizeof(float);
  if (cudaMemsetAsync(d_transient_direct, 0, bytes, s) != cudaSuccess) return check_launch();
  const int64_t tot = num_rays * n;
  transient_direct_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(d_direct_rgbs, d_weights, d_ray_dists,
                                                                                   d_light_dists, num_rays, dp,
                                                                                   d_transient_direct);
  TrHeadParams p{};
  p.n = n; p.n_bins = n_bins; p.C = channels; p.Kd = 64; p.Ks = 128; p.ld_wd = ld_w_diffuse; p.ld_ws = ld_w_specular;
  p.exposure_time = exposure_time; p.shift = shift; p.diffuse_bias = diffuse_bias; p.spec_bias = spec_bias;
  p.spec_premult = spec_premult; p.spec_max = spec_max; p.indirect_scale = indirect_scale;
  p.bin_zero_threshold_light = bin_zero_threshold_light; p.light_near = light_near; p.rgb_max = rgb_max;
  p.dark_level = dark_level; p.light_zero = light_zero;
  const size_t smem = static_cast<size_t>(kTrhRows) * N * sizeof(float) + static_cast<size_t>(n) * (4 + channels + 1) * sizeof(float) +
                      static_cast<size_t>(kTrhRows) * (64 + 8 + 128 + 8) * 2 + 16;
  if (smem > 227 * 1024) return NRC_E_UNSUPPORTED;
  if (const int32_t st_attr = ensure_dynamic_smem<transient_head_render_kernel<64, 128>>(227 * 1024)   /* the size varies with n_bins: opt in to the maximum once */; st_attr != NRC_OK)
    return st_attr;
  transient_head_render_kernel<64, 128><<<static_cast<unsigned>(num_rays), kTrhThreads, smem, s>>>(
      d_h_diffuse, d_w_diffuse, d_b_diffuse, d_h_specular, d_w_specular, d_b_specular, d_spec_scale, d_weights, d_ray_dists,
      d_light_dists, d_cam_dists, num_rays, p, d_transient_direct, d_transient_indirect, d_rgb);
  return check_launch();
}

extern "C" int32_t nrc_transient_filter(void* stream, const float* d_x, const float* d_filter, int32_t taps, int64_t num_rays,
                                        int32_t n_bins, int32_t channels, float* d_y) {
  if (num_rays < 0 || taps < 1 || n_bins < 1 || channels < 1) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_x || !d_filter || !d_y || d_x == d_y) return NRC_E_INVALID_ARG;
  const int64_t tot = num_rays * n_bins * channels;
  transient_filter_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_x, d_filter, taps, num_rays, n_bins, channels, d_y);
  return check_launch();
}
