// Training-step glue of the cache stage, fused so that the step is a static schedule of kernels with
// no elementwise round trips:
//   * predicted / analytic normals:  n = nan_to_num(-l2_normalize(g))  and its VJP
//     (internal/ref_utils.py:45-70 incl. the forward-tiny / backward-eps override, internal/geometry.py:442-479)
//   * the benchmark's cache objective in one pass over the rays: Charbonnier data term on the sRGB-mapped
//     render (internal/image.py:192-200, MaterialModel.cache_loss='charb', configs/ngp_yobo.gin:35-37)
//     plus the proposal-weight stand-in of workload.cache_loss; emits the loss AND its gradients wrt the
//     rendered rgb and the proposal levels' weights (the loss is the root of the backward pass).
#include "nrc_common.cuh"

namespace nrc {

__global__ void normals_fwd_kernel(const float* __restrict__ g, int64_t P, float* __restrict__ n) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float x = g[3 * p], y = g[3 * p + 1], z = g[3 * p + 2];
  const float d = x * x + y * y + z * z;
  const float inv = 1.0f / sqrtf(fmaxf(f32_tiny(), d));
  float o[3] = {-(x * inv), -(y * inv), -(z * inv)};
  const bool zero = d < f32_tiny();
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float v = zero ? -0.0f : o[a];
    // jnp.nan_to_num: nan -> 0, +-inf -> +-float32 max
    if (isnan(v)) v = 0.f;
    else if (isinf(v)) v = v > 0.f ? f32_max() : -f32_max();
    n[3 * p + a] = v;
  }
}

// VJP through n = -(x / sqrt(max(eps, |x|^2)))  (the backward-pass value of l2_normalize)
__global__ void normals_bwd_kernel(const float* __restrict__ g, const float* __restrict__ g_n, int64_t P,
                                   float* __restrict__ g_g) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float x = g[3 * p], y = g[3 * p + 1], z = g[3 * p + 2];
  const float u = -g_n[3 * p], v = -g_n[3 * p + 1], w = -g_n[3 * p + 2];   // through the negation
  const float d = x * x + y * y + z * z;
  float ox, oy, oz;
  if (d < f32_tiny()) {
    ox = oy = oz = 0.f;                    // where() selected the zeros branch
  } else if (d > f32_eps()) {
    const float inv = 1.0f / sqrtf(d);     // d/dx (x / |x|) = (I - n n^T) / |x|
    const float nx = x * inv, ny = y * inv, nz = z * inv;
    const float dot = nx * u + ny * v + nz * w;
    ox = (u - nx * dot) * inv; oy = (v - ny * dot) * inv; oz = (w - nz * dot) * inv;
  } else {
    const float inv = 1.0f / sqrtf(f32_eps());   // clamped denominator is a constant
    ox = u * inv; oy = v * inv; oz = w * inv;
  }
  g_g[3 * p] = ox; g_g[3 * p + 1] = oy; g_g[3 * p + 2] = oz;
}

// One warp per ray.
__global__ void cache_loss_kernel(const float* __restrict__ rgb, const float* __restrict__ target,
                                  const float* __restrict__ w0, int32_t n0, const float* __restrict__ w1, int32_t n1,
                                  const float* __restrict__ w2, int32_t n2, int64_t R, float charb_padding,
                                  float prop_weight, float* __restrict__ loss, float* __restrict__ g_rgb,
                                  float* __restrict__ g_w0, float* __restrict__ g_w1) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  float contrib = 0.f;
  if (ray < R) {
    auto row_sum = [&](const float* w, int n) {
      float s = 0.f;
      for (int i = lane; i < n; i += 32) s += w[ray * n + i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      return s;
    };
    const float a0 = row_sum(w0, n0), a1 = row_sum(w1, n1), a2 = row_sum(w2, n2);
    const float invR = 1.0f / static_cast<float>(R);
    const float d0 = a0 - a2, d1 = a1 - a2;
    const float gw0 = 2.f * prop_weight * d0 * invR, gw1 = 2.f * prop_weight * d1 * invR;
    for (int i = lane; i < n0; i += 32) g_w0[ray * n0 + i] = gw0;
    for (int i = lane; i < n1; i += 32) g_w1[ray * n1 + i] = gw1;
    if (lane == 0) contrib = prop_weight * (d0 * d0 + d1 * d1) * invR;
    if (lane < 3) {
      const float x = rgb[3 * ray + lane];
      const float eps = f32_eps();
      // image.linear_to_srgb
      const float xc = fmaxf(x, eps);
      const float p512 = powf(xc, 5.0f / 12.0f);
      const bool lin = x <= 0.0031308f;
      const float srgb = lin ? (323.0f / 25.0f) * x : (211.0f * p512 - 11.0f) / 200.0f;
      const float dsrgb = lin ? (323.0f / 25.0f) : (x > eps ? (211.0f / 200.0f) * (5.0f / 12.0f) * p512 / xc : 0.f);
      const float diff = srgb - target[3 * ray + lane];
      const float ch = sqrtf(diff * diff + charb_padding * charb_padding);
      const float inv3R = invR * (1.0f / 3.0f);
      g_rgb[3 * ray + lane] = (diff / ch) * dsrgb * inv3R;
      contrib += ch * inv3R;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  }
  __shared__ float part[8];
  if (lane == 0) part[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) s += part[i];
    atomicAdd(loss, s);
  }
}


// ------------------------------------------------------------------------------------------------
// Proposal supervision: loss_utils.spline_interlevel_loss (internal/loss_utils.py:74-108) =
// stepfun.blur_and_resample_weights (internal/stepfun.py:463-483) over linspline.blur_stepfun /
// compute_integral / interpolate_integral (internal/linspline.py:187-221, 95-109, 124-141), then the truncated
// chi-squared loss max(0, w_blur - wp)^2 / (wp + eps).  w_blur is a stop_gradient in the reference, so the
// only gradient is d loss / d wp, produced here together with the loss.  One warp per ray: the 2(m+1)-knot merge is a
// rank computation, the three running sums are chunked warp scans carried in fp64 (rounded to fp32 per element: they
// agree with a left-to-right fp64 sum except at rounding ties), all lanes evaluate the piecewise quadratic at the
// proposal fenceposts.
constexpr int kMaxKnots = 2 * (64 + 1);

__device__ __forceinline__ float plus_eps_f(float x) {
  return fabsf(x) < f32_tiny() ? f32_tiny() : nextafterf(x, INFINITY);
}
__device__ __forceinline__ float minus_eps_f(float x) {
  return fabsf(x) < f32_tiny() ? -f32_tiny() : nextafterf(x, -INFINITY);
}

struct InterlevelSmem {
  float tp[kMaxKnots], dyp[kMaxKnots], yp[kMaxKnots], a[kMaxKnots], c[kMaxKnots], acc[130];
  float lo[66], hi[66], dy[66], pdf[66];
};

__global__ void interlevel_loss_kernel(const float* __restrict__ t, const float* __restrict__ w, int m,
                                       const float* __restrict__ tq, const float* __restrict__ wp, int nq, int64_t R,
                                       float halfwidth, float mult, float eps_loss, float* __restrict__ loss,
                                       float* __restrict__ g_wp, float* __restrict__ w_blur_out) {
  __shared__ InterlevelSmem sm[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  float contrib = 0.f;
  if (ray < R) {
    InterlevelSmem& s = sm[warp];
    const float* ts = t + ray * (m + 1);
    const float* ws = w + ray * m;
    const int K = 2 * (m + 1);
    // weight_to_pdf (stepfun.py:75-79, math.safe_div) and the dilated knots of blur_stepfun, all lanes
    for (int i = lane; i <= m; i += 32) {
      const float ti = ts[i];
      s.lo[i] = fminf(minus_eps_f(ti), __fsub_rn(ti, halfwidth));
      s.hi[i] = fmaxf(plus_eps_f(ti), __fadd_rn(ti, halfwidth));
      float p = 0.f;
      if (i < m) {
        const float td = __fsub_rn(ts[i + 1], ti);
        if (!(td < f32_tiny())) p = fminf(fmaxf(__fdiv_rn(ws[i], td), -f32_max()), f32_max());
      }
      s.pdf[i] = p;   // pdf[m] = 0 (zero padding)
    }
    __syncwarp();
    for (int i = lane; i <= m; i += 32)
      s.dy[i] = __fdiv_rn(__fsub_rn(s.pdf[i], i > 0 ? s.pdf[i - 1] : 0.f), __fsub_rn(s.hi[i], s.lo[i]));
    __syncwarp();
    // Merge of the (individually sorted) knots ts_lo, ts_hi by rank: every knot's position is its own index plus the
    // number of knots of the OTHER list in front of it (binary search); stable like jnp.argsort - on ties the ts_lo
    // element (lower original index) comes first.  (A serial merge + three serial running sums on lane 0 were ~400
    // dependent shared-memory iterations per ray: 25-70 us per launch on the proposal branch of the step.)
    for (int i = lane; i <= m; i += 32) {
      const float vl = s.lo[i], vh = s.hi[i];
      int a0 = 0, a1 = m + 1;            // number of hi[j] <  lo[i]
      while (a0 < a1) { const int mid = (a0 + a1) >> 1; if (s.hi[mid] < vl) a0 = mid + 1; else a1 = mid; }
      int b0 = 0, b1 = m + 1;            // number of lo[j] <= hi[i]
      while (b0 < b1) { const int mid = (b0 + b1) >> 1; if (s.lo[mid] <= vh) b0 = mid + 1; else b1 = mid; }
      s.tp[i + a0] = vl; s.dyp[i + a0] = s.dy[i];
      s.tp[i + b0] = vh; s.dyp[i + b0] = -s.dy[i];
    }
    __syncwarp();
    // yp = [0, cumsum(diff(tp)[:-1] * cumsum(dyp[:K-2])), 0].  The double running sums are ill conditioned
    // (O(p / halfwidth) terms that cancel back to zero at the last knot): they are carried in fp64 and rounded
    // to fp32 per element -- what the oracle's torch.cumsum does on the host, and tighter than any fp32 order.
    // Each lane owns a contiguous chunk; a warp scan of the chunk totals (fp64) gives its starting value.
    auto excl_scan = [&](double v) {
      double inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const double ex = __shfl_up_sync(0xffffffffu, inc, 1);
      return lane ? ex : 0.0;
    };
    {
      const int n1 = K - 2, ch = (n1 + 31) / 32;
      const int k0 = min(lane * ch, n1), k1 = min(k0 + ch, n1);
      double tot = 0.0;
      for (int k = k0; k < k1; ++k) tot += static_cast<double>(s.dyp[k]);
      double cs = excl_scan(tot);
      double tot2 = 0.0;
      for (int k = k0; k < k1; ++k) {
        cs += static_cast<double>(s.dyp[k]);
        const float term = __fmul_rn(__fsub_rn(s.tp[k + 1], s.tp[k]), static_cast<float>(cs));
        s.c[k] = term;                       // scratch until the integral coefficients are written below
        tot2 += static_cast<double>(term);
      }
      double ys = excl_scan(tot2);
      for (int k = k0; k < k1; ++k) {
        ys += static_cast<double>(s.c[k]);
        s.yp[k + 1] = static_cast<float>(ys);
      }
      if (lane == 0) { s.yp[0] = 0.f; s.yp[K - 1] = 0.f; }
    }
    __syncwarp();
    {
      // compute_integral: a, b = yp[:-1], c
      const float e2 = f32_eps() * f32_eps();
      const int n2 = K - 1, ch = (n2 + 31) / 32;
      const int k0 = min(lane * ch, n2), k1 = min(k0 + ch, n2);
      double tot = 0.0;
      for (int k = k0; k < k1; ++k)
        if (k < K - 2) tot += static_cast<double>(__fmul_rn(__fsub_rn(s.tp[k + 1], s.tp[k]), __fadd_rn(s.yp[k], s.yp[k + 1])));
      double cc = excl_scan(tot);
      for (int k = k0; k < k1; ++k) {
        const float dt = __fsub_rn(s.tp[k + 1], s.tp[k]);
        s.a[k] = __fdiv_rn(__fsub_rn(s.yp[k + 1], s.yp[k]), fmaxf(e2, __fmul_rn(2.f, dt)));
        s.c[k] = __fmul_rn(0.5f, static_cast<float>(cc));
        if (k < K - 2) cc += static_cast<double>(__fmul_rn(dt, __fadd_rn(s.yp[k], s.yp[k + 1])));
      }
    }
    __syncwarp();
    // interpolate_integral at the nq + 1 proposal fenceposts
    const float t_first = s.tp[0], t_last = minus_eps_f(s.tp[K - 1]);
    for (int q = lane; q <= nq; q += 32) {
      float x = tq[ray * (nq + 1) + q];
      x = fmaxf(fminf(x, t_last), t_first);
      int lo_i = 0, hi_i = K;            // searchsorted(side='right'): first index with tp > x
      while (lo_i < hi_i) {
        const int mid = (lo_i + hi_i) >> 1;
        if (s.tp[mid] <= x) lo_i = mid + 1; else hi_i = mid;
      }
      const int i0 = max(lo_i - 1, 0);
      const float td = __fsub_rn(x, s.tp[i0]);
      s.acc[q] = __fadd_rn(__fadd_rn(__fmul_rn(s.a[i0], __fmul_rn(td, td)), __fmul_rn(s.yp[i0], td)), s.c[i0]);
    }
    __syncwarp();
    const float scale = mult / (static_cast<float>(R) * static_cast<float>(nq));
    for (int j = lane; j < nq; j += 32) {
      const float wb = fmaxf(0.f, __fsub_rn(s.acc[j + 1], s.acc[j]));
      const float p = wp[ray * nq + j];
      const float d = fmaxf(0.f, wb - p);
      const float den = p + eps_loss;
      contrib += d * d / den;
      // d/dp [ d^2 / (p + eps) ] = -2 d / (p + eps) - d^2 / (p + eps)^2   (d > 0)
      g_wp[ray * nq + j] = scale * (-2.f * d / den - d * d / (den * den));
      if (w_blur_out) w_blur_out[ray * nq + j] = wb;
    }
    contrib *= scale;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  }
  __shared__ float part[4];
  if (lane == 0) part[warp] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(loss, part[0] + part[1] + part[2] + part[3]);
}

// data term only: mean Charbonnier(linear_to_srgb(rgb) - target); thread per (ray, channel)
__global__ void charb_srgb_loss_kernel(const float* __restrict__ rgb, const float* __restrict__ target, int64_t R,
                                       float charb_padding, float* __restrict__ loss, float* __restrict__ g_rgb) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (i < R * 3) {
    const float x = rgb[i];
    const float eps = f32_eps();
    const float xc = fmaxf(x, eps);
    const float p512 = powf(xc, 5.0f / 12.0f);
    const bool lin = x <= 0.0031308f;
    const float srgb = lin ? (323.0f / 25.0f) * x : (211.0f * p512 - 11.0f) / 200.0f;
    const float dsrgb = lin ? (323.0f / 25.0f) : (x > eps ? (211.0f / 200.0f) * (5.0f / 12.0f) * p512 / xc : 0.f);
    const float diff = srgb - target[i];
    const float ch = sqrtf(diff * diff + charb_padding * charb_padding);
    const float inv = 1.0f / (3.0f * static_cast<float>(R));
    g_rgb[i] = (diff / ch) * dsrgb * inv;
    contrib = ch * inv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) sum += part[k];
    atomicAdd(loss, sum);
  }
}

// compute_mask_loss (internal/train_utils.py:785-836): mean(lossmult * Charbonnier(acc - mask) * weight),
// weight = opaque_w where mask > 0.5 else empty_w; lossmult == 1.  mask == nullptr: all ones.
__global__ void mask_loss_kernel(const float* __restrict__ acc, int n, const float* __restrict__ mask, int64_t R,
                                 float charb_padding, float opaque_w, float empty_w, float* __restrict__ loss,
                                 float* __restrict__ g_acc) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (i < R) {
    float a;
    if (n > 0) {   // acc = sum of the ray's weights (render.py: acc = weights.sum(-1))
      a = 0.f;
      for (int j = 0; j < n; ++j) a += acc[i * n + j];
    } else {
      a = acc[i];
    }
    const float mk = mask ? mask[i] : 1.0f;
    const float wt = (mk > 0.5f ? opaque_w : empty_w) / static_cast<float>(R);
    const float diff = a - mk;
    const float ch = sqrtf(diff * diff + charb_padding * charb_padding);
    const float g = wt * diff / ch;
    if (n > 0) {
      for (int j = 0; j < n; ++j) g_acc[i * n + j] = g;
    } else {
      g_acc[i] = g;
    }
    contrib = wt * ch;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) sum += part[k];
    atomicAdd(loss, sum);
  }
}

// distortion loss (internal/loss_utils.py:108-123 over stepfun.lossfun_distortion, internal/stepfun.py:253-269) on the
// final level, target 'tdist' through curve_fn = power_ladder(p, premult) (configs/ngp_yobo.gin:250-253):
//   c = curve(t);  u = midpoints(c);  loss_ray = sum_i w_i sum_j w_j |u_i - u_j| + sum_i w_i^2 (c_{i+1} - c_i) / 3
//   loss += mult * mean_r loss_ray;  g_w (ACCUMULATED) += mult / R * (2 sum_j w_j |u_i - u_j| + 2 w_i dc_i / 3).
// One warp per ray, n <= 128; positions carry no gradient (stop_level_grad).
__device__ __forceinline__ float power_ladder_pos(float x, float p, float premult) {
  // internal/math.py:295-316, general branch, x >= 0:  |p-1|/p * ((x/|p-1| + 1)^p - 1)
  x *= premult;
  const float a = fabsf(p - 1.0f);
  return (a / p) * (powf(fabsf(x) / a + 1.0f, p) - 1.0f) * (x < 0.f ? -1.f : 1.f);
}

__global__ void distortion_loss_kernel(const float* __restrict__ t, const float* __restrict__ w, int n, int64_t R,
                                       float p, float premult, float mult, float* __restrict__ loss,
                                       float* __restrict__ g_w) {
  __shared__ float su[8][128], sw[8][128];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  float dc[4], wi[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = lane + 32 * q;
    dc[q] = wi[q] = 0.f;
    if (i < n) {
      const float c0 = power_ladder_pos(t[r * (n + 1) + i], p, premult);
      const float c1 = power_ladder_pos(t[r * (n + 1) + i + 1], p, premult);
      su[wib][i] = 0.5f * (c0 + c1);
      dc[q] = c1 - c0;
      wi[q] = w[r * n + i];
      sw[wib][i] = wi[q];
    }
  }
  __syncwarp();
  float total = 0.f;
  const float k = mult / static_cast<float>(R);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = lane + 32 * q;
    if (i < n) {
      const float ui = su[wib][i];
      float s = 0.f;
      for (int j = 0; j < n; ++j) s = fmaf(sw[wib][j], fabsf(ui - su[wib][j]), s);
      total += wi[q] * s + wi[q] * wi[q] * dc[q] * (1.0f / 3.0f);
      g_w[r * n + i] += k * (2.f * s + 2.f * wi[q] * dc[q] * (1.0f / 3.0f));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if (lane == 0) atomicAdd(loss, k * total);
}

// Geometry losses of the final sampler level, one warp per ray (internal/train_utils.py:3255-3311):
//   orientation (internal/loss_utils.py:127-166, target 'normals_pred'):
//       mult_o * mean_r | sum_i |w_i min(0, n_i . v)^2| + 1e-5 |,  v = -viewdirs
//   predicted normals (loss_utils.py:169-199 called with gt='normals_pred' (stop-gradient), pred='normals'
//       (the ANALYTIC normals: second-order path), w = stopgrad_with_weight(w, sg_w)):
//       mult_p * mean_r | sum_i |w_i (1 - n_pred_i . n_i)| + 1e-5 |
//   reverse (gt='normals' stop-gradient, pred='normals_pred', w stop-gradient): same value, mult_r.
// Gradients: g_w and g_npred are ACCUMULATED (they already hold the compositing / shader terms), g_n is written.
__device__ __forceinline__ float nan0(float x) { return x == x ? x : 0.f; }
__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__global__ void geometry_losses_kernel(const float* __restrict__ w, const float* __restrict__ nrm,
                                       const float* __restrict__ npred, const float* __restrict__ viewdirs, int64_t R,
                                       int n, float mult_o, float mult_p, float mult_r, float sg_w,
                                       float* __restrict__ loss, float* __restrict__ g_w, float* __restrict__ g_npred,
                                       float* __restrict__ g_n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  const float invR = 1.0f / static_cast<float>(R);
  const float v0 = -viewdirs[3 * r], v1 = -viewdirs[3 * r + 1], v2 = -viewdirs[3 * r + 2];
  float so = 0.f, sp = 0.f;
  for (int i = lane; i < n; i += 32) {
    const int64_t q = r * n + i;
    const float wi = w[q];
    const float p0 = nan0(npred[3 * q]), p1 = nan0(npred[3 * q + 1]), p2 = nan0(npred[3 * q + 2]);
    float gw = 0.f, gp0 = 0.f, gp1 = 0.f, gp2 = 0.f;
    // orientation on the predicted normals
    const float ndv = p0 * v0 + p1 * v1 + p2 * v2;
    const float mn = fminf(0.f, ndv);
    const float to = wi * mn * mn;
    so += fabsf(to);
    const float so_sign = sgn(to) * mult_o * invR;
    gw += so_sign * mn * mn;
    const float co = so_sign * wi * 2.f * mn;   // mn == 0 when n.v >= 0: no gradient
    gp0 += co * v0; gp1 += co * v1; gp2 += co * v2;
    if (nrm) {
      const float a0 = nan0(nrm[3 * q]), a1 = nan0(nrm[3 * q + 1]), a2 = nan0(nrm[3 * q + 2]);
      const float one_m = 1.0f - (p0 * a0 + p1 * a1 + p2 * a2);
      const float tp = wi * one_m;
      sp += fabsf(tp);
      const float s = sgn(tp) * invR;
      gw += s * mult_p * sg_w * one_m;                       // stopgrad_with_weight(w, sg_w)
      const float ca = -s * mult_p * wi;                     // predicted-normal loss -> analytic normals
      g_n[3 * q] = ca * p0; g_n[3 * q + 1] = ca * p1; g_n[3 * q + 2] = ca * p2;
      const float cp = -s * mult_r * wi;                     // reverse loss -> predicted normals
      gp0 += cp * a0; gp1 += cp * a1; gp2 += cp * a2;
    }
    g_w[q] += gw;
    g_npred[3 * q] += gp0; g_npred[3 * q + 1] += gp1; g_npred[3 * q + 2] += gp2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    so += __shfl_xor_sync(0xffffffffu, so, o);
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
  }
  if (lane == 0) {
    float l = mult_o * fabsf(so + 1e-5f);
    if (nrm) l += (mult_p + mult_r) * fabsf(sp + 1e-5f);
    atomicAdd(loss, l * invR);
  }
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_mask_loss(void* stream, const float* d_acc, int32_t n, const float* d_mask, int64_t num_rays,
                                 float charb_padding, float opaque_weight, float empty_weight, float* d_loss,
                                 float* d_g_acc) {
  if (num_rays < 1 || n < 0 || !d_acc || !d_loss || !d_g_acc) return NRC_E_INVALID_ARG;
  mask_loss_kernel<<<static_cast<unsigned>((num_rays + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_acc, n, d_mask, num_rays, charb_padding, opaque_weight, empty_weight, d_loss, d_g_acc);
  return check_launch();
}

extern "C" int32_t nrc_distortion_loss(void* stream, const float* d_t, const float* d_weights, int32_t n, int64_t num_rays,
                                       float p, float premult, float mult, float* d_loss, float* d_g_weights) {
  if (num_rays < 1 || n < 1 || n > 128) return NRC_E_INVALID_ARG;
  if (!d_t || !d_weights || !d_loss || !d_g_weights) return NRC_E_INVALID_ARG;
  if (p == 0.f || p == 1.f) return NRC_E_UNSUPPORTED;   // the special branches of power_ladder are not configured
  const int64_t threads = num_rays * 32;
  distortion_loss_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_t, d_weights, n, num_rays, p, premult, mult, d_loss, d_g_weights);
  return check_launch();
}

extern "C" int32_t nrc_geometry_losses(void* stream, const float* d_weights, const float* d_normals,
                                       const float* d_normals_pred, const float* d_viewdirs, int64_t num_rays,
                                       int32_t n, float orientation_mult, float predicted_normal_mult,
                                       float predicted_normal_reverse_mult, float stopgrad_weight, float* d_loss,
                                       float* d_g_weights, float* d_g_normals_pred, float* d_g_normals) {
  if (num_rays < 1 || n < 1) return NRC_E_INVALID_ARG;
  if (!d_weights || !d_normals_pred || !d_viewdirs || !d_loss || !d_g_weights || !d_g_normals_pred) return NRC_E_INVALID_ARG;
  if (d_normals && !d_g_normals) return NRC_E_INVALID_ARG;
  const int64_t threads = num_rays * 32;
  geometry_losses_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_weights, d_normals, d_normals_pred, d_viewdirs, num_rays, n, orientation_mult, predicted_normal_mult,
      predicted_normal_reverse_mult, stopgrad_weight, d_loss, d_g_weights, d_g_normals_pred, d_g_normals);
  return check_launch();
}

extern "C" int32_t nrc_normals_fwd(void* stream, const float* d_grad, int64_t num_points, float* d_normals) {
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_grad || !d_normals) return NRC_E_INVALID_ARG;
  normals_fwd_kernel<<<static_cast<unsigned>((num_points + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_grad, num_points, d_normals);
  return check_launch();
}

extern "C" int32_t nrc_normals_bwd(void* stream, const float* d_grad, const float* d_g_normals, int64_t num_points,
                                   float* d_g_grad) {
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_grad || !d_g_normals || !d_g_grad) return NRC_E_INVALID_ARG;
  normals_bwd_kernel<<<static_cast<unsigned>((num_points + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_grad, d_g_normals, num_points, d_g_grad);
  return check_launch();
}

extern "C" int32_t nrc_cache_loss(void* stream, const float* d_rgb, const float* d_target, const float* d_w0, int32_t n0,
                                  const float* d_w1, int32_t n1, const float* d_w2, int32_t n2, int64_t num_rays,
                                  float charb_padding, float prop_weight, float* d_loss, float* d_g_rgb, float* d_g_w0,
                                  float* d_g_w1) {
  if (num_rays < 1 || n0 < 1 || n1 < 1 || n2 < 1) return NRC_E_INVALID_ARG;
  if (!d_rgb || !d_target || !d_w0 || !d_w1 || !d_w2 || !d_loss || !d_g_rgb || !d_g_w0 || !d_g_w1)
    return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(d_loss, 0, sizeof(float), s) != cudaSuccess) return check_launch();
  const unsigned grid = static_cast<unsigned>((num_rays * 32 + 255) / 256);
  cache_loss_kernel<<<grid, 256, 0, s>>>(d_rgb, d_target, d_w0, n0, d_w1, n1, d_w2, n2, num_rays, charb_padding,
                                         prop_weight, d_loss, d_g_rgb, d_g_w0, d_g_w1);
  return check_launch();
}

extern "C" int32_t nrc_interlevel_loss(void* stream, const float* d_t, const float* d_w, int32_t m, const float* d_tq,
                                       const float* d_wp, int32_t nq, int64_t num_rays, float blur_halfwidth, float mult,
                                       float eps, float* d_loss, float* d_g_wp, float* d_w_blur) {
  if (num_rays < 1 || m < 1 || m > 64 || nq < 1 || nq > 128) return NRC_E_INVALID_ARG;
  if (!d_t || !d_w || !d_tq || !d_wp || !d_loss || !d_g_wp) return NRC_E_INVALID_ARG;
  interlevel_loss_kernel<<<static_cast<unsigned>((num_rays + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_t, d_w, m, d_tq, d_wp, nq, num_rays, blur_halfwidth, mult, eps, d_loss, d_g_wp, d_w_blur);
  return check_launch();
}

extern "C" int32_t nrc_charb_srgb_loss(void* stream, const float* d_rgb, const float* d_target, int64_t num_rays,
                                       float charb_padding, float* d_loss, float* d_g_rgb) {
  if (num_rays < 1) return NRC_E_INVALID_ARG;
  if (!d_rgb || !d_target || !d_loss || !d_g_rgb) return NRC_E_INVALID_ARG;
  const int64_t n = num_rays * 3;
  charb_srgb_loss_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_rgb, d_target, num_rays, charb_padding, d_loss, d_g_rgb);
  return check_launch();
}
