// Time-resolved (transient) cache rendering, SURVEY 8a row 22 / BASELINE config 4:
//   internal/nerf.py:1660-1777 (_compute_indirect_lighting / get_indirect: softplus head, indirect_scale, clip),
//   internal/inverse_render/render_utils.py:1699-1767 (zero_invalid_bins),
//   internal/render.py:250-449 (volumetric_transient_rendering), :452-490 (shift_direct), :493-507
//   (shift_map_coordinates = order-1 map_coordinates along the bin axis, mode='constant').
// The reference materialises [R, n, n_bins, 3] tensors three to four times (activation, masks, shift)
// before reducing over the n samples (275 MB each at R = 1024, n_bins = 700).  Here the raw head outputs are
// read ONCE: activation, validity masks, clip, the per-sample sub-bin shift and the weighted reduction over
// samples happen in registers; only [R, n_bins, 3] leaves the SM.
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
