// Per-ray loss terms shared by the render + loss kernels (ray.cu: nrc_render_loss, shader.cu: nrc_shade_render_loss):
// the Charbonnier-sRGB data term (internal/image.py:192-200, configs/ngp_yobo.gin:35-37) and compute_mask_loss on the
// accumulation (internal/train_utils.py:785-836, lossmult == 1).  Each returns its loss contribution and the gradient.
#pragma once
#include "nrc_common.cuh"

namespace nrc {

// x: rendered channel value, inv = 1 / (3 R).
__device__ __forceinline__ float charb_srgb_term(float x, float target, float charb_padding, float inv, float& g) {
  const float eps = f32_eps();
  const float xc = fmaxf(x, eps);
  const float p512 = powf(xc, 5.0f / 12.0f);
  const bool lin = x <= 0.0031308f;
  const float srgb = lin ? (323.0f / 25.0f) * x : (211.0f * p512 - 11.0f) / 200.0f;
  const float dsrgb = lin ? (323.0f / 25.0f) : (x > eps ? (211.0f / 200.0f) * (5.0f / 12.0f) * p512 / xc : 0.f);
  const float diff = srgb - target;
  const float ch = sqrtf(diff * diff + charb_padding * charb_padding);
  g = (diff / ch) * dsrgb * inv;
  return ch * inv;
}

// acc: sum of the ray's weights, mk: mask value, invR = 1 / R.
__device__ __forceinline__ float mask_term(float acc, float mk, float opaque_w, float empty_w, float invR, float charb_padding,
                                           float& g) {
  const float wt = (mk > 0.5f ? opaque_w : empty_w) * invR;
  const float diff = acc - mk;
  const float ch = sqrtf(diff * diff + charb_padding * charb_padding);
  g = wt * diff / ch;
  return wt * ch;
}

}  // namespace nrc
