// K4: per-ray kernels -- alpha compositing weights, CDF inversion / interval
// resampling, ray casting (s->t warp + Gaussian means), compositing with
// distance statistics, categorical resampling.  One warp owns one ray; all
// per-ray intermediates live in registers / shared memory, nothing but the
// named outputs goes back to HBM.
//
// Reference: internal/render.py:106-247, internal/stepfun.py:125-250,306-314,
// internal/math.py:295-341,412-457, internal/sampling.py:326-368,
// internal/models.py:193-292.
#include "loss_terms.cuh"
#include "nrc_common.cuh"

namespace nrc {

constexpr int kMaxN = 128;          // max samples per ray per level
constexpr int kRayWarps = 4;        // rays (warps) per CTA
constexpr int kRayThreads = kRayWarps * 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Inclusive prefix sums with the reduce_window semantics of jnp.cumsum on XLA:CPU:
// every output is the left-to-right sum of its own window.  src/dst in shared memory.
__device__ __forceinline__ void window_cumsum(const float* src, float* dst, int n, int lane) {
  for (int i = lane; i < n; i += 32) {
    float s = 0.f;
    for (int j = 0; j <= i; ++j) s = __fadd_rn(s, src[j]);
    dst[i] = s;
  }
}

// Inclusive prefix sums / exclusive suffix sums of a shared-memory row by chunked warp scans (32 entries per pass): used
// where the result feeds a gradient, not a fencepost (summation order differs from window_cumsum in the last bits).
__device__ __forceinline__ void scan_prefix_inclusive(const float* src, float* dst, int n, int lane) {
  float carry = 0.f;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int i = c0 + lane;
    float x = i < n ? src[i] : 0.f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (i < n) dst[i] = carry + x;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
}
__device__ __forceinline__ void scan_suffix_exclusive(const float* src, float* dst, int n, int lane) {
  float carry = 0.f;
  for (int c0 = ((n - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
    const int i = c0 + lane;
    const float v = i < n ? src[i] : 0.f;
    float x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float y = __shfl_down_sync(0xffffffffu, x, o);
      if (lane + o < 32) x += y;
    }
    if (i < n) dst[i] = carry + (x - v);
    carry += __shfl_sync(0xffffffffu, x, 0);
  }
}

// ------------------------------------------------------------ alpha weights --
__global__ void __launch_bounds__(kRayThreads)
alpha_weights_fwd_kernel(const float* __restrict__ density, const float* __restrict__ tdist,
                         const float* __restrict__ dirs, int64_t R, int n, int opaque,
                         float* __restrict__ weights, float* __restrict__ alpha,
                         float* __restrict__ trans) {
  __shared__ float s_dd[kRayWarps][kMaxN];
  __shared__ float s_cs[kRayWarps][kMaxN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  if (r >= R) return;
  const float d0 = dirs[3 * r], d1 = dirs[3 * r + 1], d2 = dirs[3 * r + 2];
  const float dn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
  const float* t = tdist + r * (n + 1);
  const float* den = density + r * n;
  for (int i = lane; i < n; i += 32) {
    float delta = __fmul_rn(__fsub_rn(t[i + 1], t[i]), dn);
    float dd = __fmul_rn(den[i], fabsf(delta));
    if (opaque && i == n - 1) dd = INFINITY;
    s_dd[warp][i] = dd;
  }
  __syncwarp();
  window_cumsum(s_dd[warp], s_cs[warp], n - 1, lane);
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    float a = __fsub_rn(1.0f, expf(-s_dd[warp][i]));
    float tr = expf(-(i == 0 ? 0.f : s_cs[warp][i - 1]));
    weights[r * n + i] = __fmul_rn(a, tr);
    if (alpha) alpha[r * n + i] = a;
    if (trans) trans[r * n + i] = tr;
  }
}

__global__ void __launch_bounds__(kRayThreads)
alpha_weights_bwd_kernel(const float* __restrict__ density, const float* __restrict__ tdist,
                         const float* __restrict__ dirs, const float* __restrict__ g_w,
                         const float* __restrict__ g_a, const float* __restrict__ g_t, int64_t R,
                         int n, float* __restrict__ g_density) {
  __shared__ float s_dd[kRayWarps][kMaxN];
  __shared__ float s_cs[kRayWarps][kMaxN];
  __shared__ float s_gt[kRayWarps][kMaxN];  // gT_k * T_k
  __shared__ float s_tr[kRayWarps][kMaxN];  // T_k
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  if (r >= R) return;
  const float d0 = dirs[3 * r], d1 = dirs[3 * r + 1], d2 = dirs[3 * r + 2];
  const float dn = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
  const float* t = tdist + r * (n + 1);
  for (int i = lane; i < n; i += 32) {
    float delta = (t[i + 1] - t[i]) * dn;
    s_dd[warp][i] = density[r * n + i] * fabsf(delta);
  }
  __syncwarp();
  scan_prefix_inclusive(s_dd[warp], s_cs[warp], n - 1, lane);
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    float a = 1.0f - expf(-s_dd[warp][i]);
    float tr = expf(-(i == 0 ? 0.f : s_cs[warp][i - 1]));
    float gw = g_w ? g_w[r * n + i] : 0.f;
    float gT = (g_t ? g_t[r * n + i] : 0.f) + gw * a;
    s_gt[warp][i] = gT * tr;
    s_tr[warp][i] = tr;
  }
  __syncwarp();
  scan_suffix_exclusive(s_gt[warp], s_cs[warp], n, lane);     // sum_{k > i} gT_k T_k (the transmittances are not needed any more)
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    float e = expf(-s_dd[warp][i]);
    float gw = g_w ? g_w[r * n + i] : 0.f;
    float gA = (g_a ? g_a[r * n + i] : 0.f) + gw * s_tr[warp][i];
    float g_dd = gA * e - s_cs[warp][i];
    float delta = (t[i + 1] - t[i]) * dn;
    g_density[r * n + i] = g_dd * fabsf(delta);
  }
}

// power-ladder ray warp: power_ladder_fwd / power_ladder_inv live in nrc_common.cuh (shared with slf.cu)


// --------------------------------------------------------- sample_intervals --
// sampling.py:340 (annealed logits) + stepfun.sample_intervals :207-250
// (sample :158-204 -> invert_cdf :147-155 -> integrate_weights :125-144 ->
// math.sorted_interp :447-457), single_jitter=True.
__global__ void __launch_bounds__(kRayThreads)
sample_intervals_kernel(const float* __restrict__ t_in, const float* __restrict__ w_in,
                        const float* __restrict__ u01, const float* __restrict__ u_base, int64_t R,
                        int m, int n, float anneal, float padding, float max_jitter, float dom_lo,
                        float dom_hi, float* __restrict__ t_new, int32_t* __restrict__ bin_idx,
                        // optional fused cast_rays (nrc_ray_sample_cast): tdist == nullptr -> sampling only
                        const float* __restrict__ origins, const float* __restrict__ directions,
                        const float* __restrict__ near, const float* __restrict__ far, int warp_kind, float p,
                        float premult, float* __restrict__ tdist, float* __restrict__ means,
                        // optional fused compute_alpha_weights of the level being resampled (nrc_ray_weights_sample_cast):
                        // w_in is ignored, the weights come from (density_prev, tdist_prev) and are written to weights_out
                        const float* __restrict__ density_prev = nullptr, const float* __restrict__ tdist_prev = nullptr,
                        int opaque = 0, float* __restrict__ weights_out = nullptr) {
  __shared__ float s_t[kRayWarps][kMaxN + 1];
  __shared__ float s_w[kRayWarps][kMaxN];       // softmax weights
  __shared__ float s_cw[kRayWarps][kMaxN + 1];  // integrated weights
  __shared__ float s_c[kRayWarps][kMaxN];       // sampled centres
  __shared__ float s_s[kRayWarps][kMaxN + 1];   // fenceposts before sorting
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  if (r >= R) return;
  for (int i = lane; i <= m; i += 32) s_t[warp][i] = t_in[r * (m + 1) + i];
  if (density_prev) {
    // same expressions as alpha_weights_fwd_kernel (render.compute_alpha_weights, internal/render.py:134-169); the
    // weights never leave the SM between the density query and the resampling (they are also written out: the
    // interlevel loss and the backward pass read them)
    const float e0 = directions[3 * r], e1 = directions[3 * r + 1], e2 = directions[3 * r + 2];
    const float dn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2)));
    const float* tp = tdist_prev + r * (m + 1);
    for (int i = lane; i < m; i += 32) {
      float delta = __fmul_rn(__fsub_rn(tp[i + 1], tp[i]), dn);
      float dd = __fmul_rn(density_prev[r * m + i], fabsf(delta));
      if (opaque && i == m - 1) dd = INFINITY;
      s_c[warp][i] = dd;
    }
    __syncwarp();
    window_cumsum(s_c[warp], s_s[warp], m - 1, lane);
    __syncwarp();
    for (int i = lane; i < m; i += 32) {
      float a = __fsub_rn(1.0f, expf(-s_c[warp][i]));
      float tr = expf(-(i == 0 ? 0.f : s_s[warp][i - 1]));
      const float wgt = __fmul_rn(a, tr);
      s_cw[warp][i] = wgt;
      if (weights_out) weights_out[r * m + i] = wgt;
    }
    __syncwarp();
  }
  // logits and softmax (jax.nn.softmax: exp(x - max) / sum)
  float lmax = -INFINITY;
  for (int i = lane; i < m; i += 32) {
    float lg = __fmul_rn(anneal, safe_log(__fadd_rn(density_prev ? s_cw[warp][i] : w_in[r * m + i], padding)));
    s_w[warp][i] = lg;
    lmax = fmaxf(lmax, lg);
  }
  lmax = warp_max(lmax);
  float lsum = 0.f;
  for (int i = lane; i < m; i += 32) {
    float e = expf(__fsub_rn(s_w[warp][i], lmax));
    s_w[warp][i] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  for (int i = lane; i < m; i += 32) s_w[warp][i] = __fdiv_rn(s_w[warp][i], lsum);
  __syncwarp();
  // cw = [0, min(1, cumsum(w[:-1])), 1]
  window_cumsum(s_w[warp], s_cw[warp] + 1, m - 1, lane);
  __syncwarp();
  for (int i = lane; i < m - 1; i += 32) s_cw[warp][i + 1] = fminf(1.0f, s_cw[warp][i + 1]);
  if (lane == 0) { s_cw[warp][0] = 0.f; s_cw[warp][m] = 1.0f; }
  __syncwarp();
  // u = linspace(0, 1-u_max, n) + uniform(maxval=max_jitter)   (one jitter per ray)
  const float jitter = fmaxf(0.f, __fmul_rn(u01[r], max_jitter));
  const float eps2 = f32_eps() * f32_eps();
  for (int j = lane; j < n; j += 32) {
    float u = __fadd_rn(u_base[j], jitter);
    // searchsorted(cw, u, side='right'): number of entries <= u (cw is sorted)
    int lo = 0, hi = m + 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (s_cw[warp][mid] <= u) lo = mid + 1; else hi = mid;
    }
    int idx1 = min(lo, m);
    int idx0 = max(lo - 1, 0);
    float xp0 = s_cw[warp][idx0], xp1 = s_cw[warp][idx1];
    float fp0 = s_t[warp][idx0], fp1 = s_t[warp][idx1];
    float off = __fdiv_rn(__fsub_rn(u, xp0), fmaxf(eps2, __fsub_rn(xp1, xp0)));
    off = fminf(fmaxf(off, 0.f), 1.f);
    s_c[warp][j] = __fadd_rn(fp0, __fmul_rn(off, __fsub_rn(fp1, fp0)));
    if (bin_idx) bin_idx[r * n + j] = idx0;
  }
  __syncwarp();
  // fenceposts: reflect the end midpoints about the end centres, clip to the domain
  for (int j = lane; j <= n; j += 32) {
    float v;
    if (j == 0) {
      float mid0 = __fdiv_rn(__fadd_rn(s_c[warp][1], s_c[warp][0]), 2.0f);
      v = __fsub_rn(__fmul_rn(2.0f, s_c[warp][0]), mid0);
    } else if (j == n) {
      float midl = __fdiv_rn(__fadd_rn(s_c[warp][n - 1], s_c[warp][n - 2]), 2.0f);
      v = __fsub_rn(__fmul_rn(2.0f, s_c[warp][n - 1]), midl);
    } else {
      v = __fdiv_rn(__fadd_rn(s_c[warp][j], s_c[warp][j - 1]), 2.0f);
    }
    s_s[warp][j] = fminf(fmaxf(v, dom_lo), dom_hi);
  }
  __syncwarp();
  // jnp.sort.  The fenceposts are sorted up to last-ulp inversions: a warp vote finds the (usual) already-sorted
  // case, where the stable sort is the identity; otherwise a stable rank sort (O(n^2), same result as before).
  bool ok = true;
  for (int j = lane; j < n; j += 32) ok = ok && (s_s[warp][j] <= s_s[warp][j + 1]);
  const float* sorted = s_s[warp];
  if (!__all_sync(0xffffffffu, ok)) {
    for (int j = lane; j <= n; j += 32) {
      float v = s_s[warp][j];
      int rank = 0;
      for (int k = 0; k <= n; ++k) {
        float o = s_s[warp][k];
        rank += (o < v) || (o == v && k < j);
      }
      s_cw[warp][rank] = v;   // the integrated weights are no longer needed
    }
    __syncwarp();
    sorted = s_cw[warp];
  }
  for (int j = lane; j <= n; j += 32) t_new[r * (n + 1) + j] = sorted[j];
  if (tdist == nullptr) return;
  // ---- fused cast_rays (same expressions as ray_cast_kernel; every metric distance is computed once) ----
  float s_near = near[r], s_far = far[r];
  if (warp_kind == 1) {
    s_near = power_ladder_fwd(s_near, p, premult);
    s_far = power_ladder_fwd(s_far, p, premult);
  }
  for (int j = lane; j <= n; j += 32) {
    const float sv = sorted[j];
    float v = __fadd_rn(__fmul_rn(sv, s_far), __fmul_rn(__fsub_rn(1.0f, sv), s_near));
    if (warp_kind == 1) v = power_ladder_inv(v, p, premult);
    tdist[r * (n + 1) + j] = v;
    s_t[warp][j] = v;           // the input fenceposts are no longer needed
  }
  if (means == nullptr) return;
  __syncwarp();
  const float o0 = origins[3 * r], o1 = origins[3 * r + 1], o2 = origins[3 * r + 2];
  const float d0 = directions[3 * r], d1 = directions[3 * r + 1], d2 = directions[3 * r + 2];
  const float eps2c = f32_eps() * f32_eps();
  for (int i = lane; i < n; i += 32) {
    const float t0 = s_t[warp][i], t1 = s_t[warp][i + 1];
    const float sm = __fadd_rn(t0, t1), d = __fsub_rn(t1, t0);
    const float dsq = __fmul_rn(d, d);
    const float ratio = __fdiv_rn(dsq, fmaxf(eps2c, __fadd_rn(__fmul_rn(3.0f, __fmul_rn(sm, sm)), dsq)));
    const float t_mean = __fmul_rn(sm, __fadd_rn(0.5f, ratio));
    float* mo = means + (r * n + i) * 3;
    mo[0] = __fadd_rn(__fmul_rn(d0, t_mean), o0);
    mo[1] = __fadd_rn(__fmul_rn(d1, t_mean), o1);
    mo[2] = __fadd_rn(__fmul_rn(d2, t_mean), o2);
  }
}

// ------------------------------------------------------------------ ray cast --
__global__ void ray_cast_kernel(const float* __restrict__ sdist, const float* __restrict__ origins,
                                const float* __restrict__ directions, const float* __restrict__ near,
                                const float* __restrict__ far, int64_t R, int n, int warp_kind, float p,
                                float premult, float* __restrict__ tdist, float* __restrict__ means) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= R * (n + 1)) return;
  const int64_t r = gid / (n + 1);
  const int i = static_cast<int>(gid - r * (n + 1));
  float s_near = near[r], s_far = far[r];
  if (warp_kind == 1) {
    s_near = power_ladder_fwd(s_near, p, premult);
    s_far = power_ladder_fwd(s_far, p, premult);
  }
  auto s_to_t = [&](float s) {
    float v = __fadd_rn(__fmul_rn(s, s_far), __fmul_rn(__fsub_rn(1.0f, s), s_near));
    return warp_kind == 1 ? power_ladder_inv(v, p, premult) : v;
  };
  const float t0 = s_to_t(sdist[gid]);
  tdist[gid] = t0;
  if (i == n || means == nullptr) return;
  const float t1 = s_to_t(sdist[gid + 1]);
  // gaussianize_frustum t_mean (internal/render.py:49-59)
  float s = __fadd_rn(t0, t1), d = __fsub_rn(t1, t0);
  float eps2 = f32_eps() * f32_eps();
  float d2 = __fmul_rn(d, d);
  float ratio = __fdiv_rn(d2, fmaxf(eps2, __fadd_rn(__fmul_rn(3.0f, __fmul_rn(s, s)), d2)));
  float t_mean = __fmul_rn(s, __fadd_rn(0.5f, ratio));
  float* mo = means + (r * n + i) * 3;
#pragma unroll
  for (int a = 0; a < 3; ++a) mo[a] = __fadd_rn(__fmul_rn(directions[3 * r + a], t_mean), origins[3 * r + a]);
}

// ------------------------------------------------------------ ray cast: covs --
// render.cast_rays covariances (internal/render.py:26-103): gaussianize_frustum (cone) or cylinder_to_gaussian, lifted by
// lift_gaussian.  One thread per (ray, interval); diag: 3 floats, else the full symmetric 3 x 3 (9 floats).  Also the
// cylinder MEAN (t0 + t1) / 2, which nrc_ray_cast (cone mean) does not produce.
__global__ void ray_cast_covs_kernel(const float* __restrict__ tdist, const float* __restrict__ origins,
                                     const float* __restrict__ directions, const float* __restrict__ radii, int64_t R, int n,
                                     int cylinder, int diag, float* __restrict__ covs, float* __restrict__ means) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= R * n) return;
  const int64_t r = gid / n;
  const int i = static_cast<int>(gid - r * n);
  const float t0 = tdist[r * (n + 1) + i], t1 = tdist[r * (n + 1) + i + 1];
  const float rad = radii[r];
  float t_mean, t_var, r_var;
  if (cylinder) {
    t_mean = (t0 + t1) / 2.0f;                                   // :100-103
    r_var = rad * rad / 4.0f;
    const float dd = t1 - t0;
    t_var = dd * dd / 12.0f;
  } else {
    const float s = t0 + t1, d = t1 - t0;                        // :49-59
    const float eps2 = f32_eps() * f32_eps();
    const float d2 = d * d, s2 = s * s;
    const float ratio = d2 / fmaxf(eps2, 3.0f * s2 + d2);
    t_mean = s * (0.5f + ratio);
    t_var = (1.0f / 12.0f) * d2 - (1.0f / 15.0f) * (ratio * ratio) * (12.0f * s2 - d2);
    r_var = (1.0f / 16.0f) * s2 + d2 * (5.0f / 48.0f - (1.0f / 15.0f) * ratio);
    r_var *= rad * rad;                                          // :79
  }
  const float d0 = directions[3 * r], d1 = directions[3 * r + 1], d2v = directions[3 * r + 2];
  const float dv[3] = {d0, d1, d2v};
  const float mag = fmaxf(1e-10f, d0 * d0 + d1 * d1 + d2v * d2v);   // :30
  if (means != nullptr) {
    float* mo = means + gid * 3;
#pragma unroll
    for (int a = 0; a < 3; ++a) mo[a] = __fadd_rn(__fmul_rn(dv[a], t_mean), origins[3 * r + a]);
  }
  if (covs == nullptr) return;
  if (diag) {
    float* co = covs + gid * 3;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float outer = dv[a] * dv[a];
      co[a] = t_var * outer + r_var * (1.0f - outer / mag);       // :33-38
    }
  } else {
    float* co = covs + gid * 9;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const float null_outer = (a == b ? 1.0f : 0.0f) - dv[a] * (dv[b] / mag);   // :41-42
        co[3 * a + b] = t_var * (dv[a] * dv[b]) + r_var * null_outer;              // :43-45
      }
  }
}

// ---------------------------------------------------------------- composite --
__global__ void __launch_bounds__(kRayThreads)
composite_fwd_kernel(const float* __restrict__ values, const float* __restrict__ weights, int k,
                     const float* __restrict__ weights_nf, const float* __restrict__ tdist,
                     const float* __restrict__ bg, int64_t R, int n, int C, int has_rgb,
                     float* __restrict__ out, float* __restrict__ acc_out, float* __restrict__ dist) {
  __shared__ float s_w[kRayWarps][kMaxN];
  __shared__ float s_cw[kRayWarps][kMaxN + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  if (r >= R) return;
  // acc = weights_no_filter.sum(-1)   (internal/render.py:200-201)
  float a = 0.f;
  for (int i = lane; i < n; i += 32) {
    float w = weights_nf[r * n + i];
    s_w[warp][i] = w;
    a += w;
  }
  const float acc = warp_sum(a);
  if (lane == 0 && acc_out) acc_out[r] = acc;
  __syncwarp();
  const float bg_w = fmaxf(0.f, 1.0f - acc);
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (int i = lane; i < k; i += 32) s += weights[r * k + i] * values[(r * k + i) * C + c];
    s = warp_sum(s);
    if (lane == 0) {
      if (has_rgb && c < 3 && bg) s += bg_w * bg[3 * r + c];
      out[r * C + c] = s;
    }
  }
  if (!dist) return;
  const float* t = tdist + r * (n + 1);
  // distance_mean = clip(nan_to_num(exp(sum(w log t_mid) / max(eps, acc))), t0, tn)  (:226-238)
  float e = 0.f;
  for (int i = lane; i < n; i += 32) e += s_w[warp][i] * logf(0.5f * (t[i] + t[i + 1]));
  e = warp_sum(e);
  const float denom = fmaxf(f32_eps(), acc);
  // percentiles: cw = integrate_weights(w / max(eps, acc)); interp(ps/100, cw, t)  (:240-246)
  for (int i = lane; i < n; i += 32) s_w[warp][i] = s_w[warp][i] / denom;
  __syncwarp();
  window_cumsum(s_w[warp], s_cw[warp] + 1, n - 1, lane);
  __syncwarp();
  for (int i = lane; i < n - 1; i += 32) s_cw[warp][i + 1] = fminf(1.0f, s_cw[warp][i + 1]);
  if (lane == 0) { s_cw[warp][0] = 0.f; s_cw[warp][n] = 1.0f; }
  __syncwarp();
  if (lane == 0) {
    float dm = expf(e / denom);
    if (isnan(dm)) dm = 0.f;
    dm = fminf(fmaxf(dm, t[0]), t[n]);
    dist[r * 4 + 0] = dm;
  } else if (lane <= 3) {
    const float ps = lane == 1 ? 0.05f : (lane == 2 ? 0.5f : 0.95f);
    // jnp.interp: i = clip(searchsorted(xp, x, 'right'), 1, len-1)
    int lo = 0, hi = n + 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (s_cw[warp][mid] <= ps) lo = mid + 1; else hi = mid;
    }
    int i = min(max(lo, 1), n);
    float dx = s_cw[warp][i] - s_cw[warp][i - 1];
    float df = t[i] - t[i - 1];
    float delta = ps - s_cw[warp][i - 1];
    const float epsilon = 1.4210855e-14f;  // np.spacing(finfo(float32).eps)
    float f = fabsf(dx) <= epsilon ? t[i - 1] : t[i - 1] + (delta / dx) * df;
    if (ps < s_cw[warp][0]) f = t[0];
    if (ps > s_cw[warp][n]) f = t[n];
    dist[r * 4 + lane] = f;
  }
}

__global__ void __launch_bounds__(kRayThreads)
composite_bwd_kernel(const float* __restrict__ values, const float* __restrict__ weights, int k,
                     const float* __restrict__ weights_nf, const float* __restrict__ bg,
                     const float* __restrict__ g_out, const float* __restrict__ g_acc, int64_t R, int n,
                     int C, int has_rgb, float* __restrict__ g_values, float* __restrict__ g_weights,
                     float* __restrict__ g_weights_nf) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  if (r >= R) return;
  float a = 0.f;
  for (int i = lane; i < n; i += 32) a += weights_nf[r * n + i];
  const float acc = warp_sum(a);
  // out_c = sum_i w_i v_ic + [c<3] max(0, 1-acc) bg_c ;  acc = sum_j wnf_j
  float gacc = g_acc ? g_acc[r] : 0.f;
  if (has_rgb && bg && (1.0f - acc) > 0.f) {
    for (int c = 0; c < 3 && c < C; ++c) gacc -= g_out[r * C + c] * bg[3 * r + c];
  }
  const bool shared_w = (g_weights_nf == nullptr) || (g_weights_nf == g_weights);
  for (int i = lane; i < k; i += 32) {
    float w = weights[r * k + i];
    float gw = shared_w ? gacc : 0.f;
    for (int c = 0; c < C; ++c) {
      float go = g_out[r * C + c];
      gw += go * values[(r * k + i) * C + c];
      if (g_values) g_values[(r * k + i) * C + c] = go * w;
    }
    if (g_weights) g_weights[r * k + i] = gw;
  }
  if (!shared_w)
    for (int i = lane; i < n; i += 32) g_weights_nf[r * n + i] = gacc;
}

// ---------------------------------------------------- render + data/mask loss --
// The tail of the cache training step in ONE launch (one warp per ray): volumetric_rendering's rgb and acc
// (internal/render.py:172-224; the distance statistics are not consumed by the objective - XLA eliminates them from
// the reference's training step too), the Charbonnier-sRGB data term (internal/image.py:192-200,
// configs/ngp_yobo.gin:35-37), compute_mask_loss on acc (internal/train_utils.py:785-836, lossmult == 1) and the VJP
// of the compositing.  Same expressions and summation order as composite_fwd_kernel / charb_srgb_loss_kernel /
// mask_loss_kernel / composite_bwd_kernel, which remain the reference points of the parity tests.
__global__ void __launch_bounds__(kRayThreads)
render_loss_kernel(const float* __restrict__ values, const float* __restrict__ weights, const float* __restrict__ bg,
                   const float* __restrict__ target, const float* __restrict__ mask, int64_t R, int n,
                   float charb_padding, int use_mask, float opaque_w, float empty_w, float* __restrict__ loss,
                   float* __restrict__ out_rgb, float* __restrict__ acc_out, float* __restrict__ g_values,
                   float* __restrict__ g_weights) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  float contrib = 0.f;
  if (r < R) {
    float a = 0.f;
    for (int i = lane; i < n; i += 32) a += weights[r * n + i];
    const float acc = warp_sum(a);
    const float bg_w = fmaxf(0.f, 1.0f - acc);
    float out[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sc = 0.f;
      for (int i = lane; i < n; i += 32) sc += weights[r * n + i] * values[(r * n + i) * 3 + c];
      sc = warp_sum(sc);
      if (bg) sc += bg_w * bg[3 * r + c];
      out[c] = sc;
    }
    if (lane < 3) out_rgb[3 * r + lane] = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
    if (lane == 0 && acc_out) acc_out[r] = acc;
    // lanes 0..2: data term of one channel; lane 3: mask term
    float g = 0.f;
    const float invR = 1.0f / static_cast<float>(R);
    if (lane < 3) {
      const float x = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
      contrib = charb_srgb_term(x, target[3 * r + lane], charb_padding, 1.0f / (3.0f * static_cast<float>(R)), g);
    } else if (lane == 3 && use_mask) {
      contrib = mask_term(acc, mask ? mask[r] : 1.0f, opaque_w, empty_w, invR, charb_padding, g);
    }
    const float go0 = __shfl_sync(0xffffffffu, g, 0), go1 = __shfl_sync(0xffffffffu, g, 1),
                go2 = __shfl_sync(0xffffffffu, g, 2);
    float gacc = __shfl_sync(0xffffffffu, g, 3);
    if (bg && (1.0f - acc) > 0.f) {
      gacc -= go0 * bg[3 * r];
      gacc -= go1 * bg[3 * r + 1];
      gacc -= go2 * bg[3 * r + 2];
    }
    for (int i = lane; i < n; i += 32) {
      const float w = weights[r * n + i];
      const float* v = values + (r * n + i) * 3;
      float gw = gacc;
      gw += go0 * v[0]; gw += go1 * v[1]; gw += go2 * v[2];
      float* gv = g_values + (r * n + i) * 3;
      gv[0] = go0 * w; gv[1] = go1 * w; gv[2] = go2 * w;
      g_weights[r * n + i] = gw;
    }
    contrib = warp_sum(contrib);
  }
  __shared__ float part[kRayWarps];
  if (lane == 0) part[warp] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kRayWarps; ++k) sum += part[k];
    atomicAdd(loss, sum);
  }
}

// ----------------------------------------------------------------- resample --
__global__ void __launch_bounds__(kRayThreads)
resample_kernel(const float* __restrict__ weights, const float* __restrict__ gumbel, int64_t R, int n,
                int k, float bias, float mult, int32_t* __restrict__ inds, float* __restrict__ w_new) {
  __shared__ float s_l[kRayWarps][kMaxN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kRayWarps + warp;
  if (r >= R) return;
  float lmax = -INFINITY;
  for (int i = lane; i < n; i += 32) {
    float lg = __fmul_rn(safe_log(__fadd_rn(weights[r * n + i], bias)), mult);
    s_l[warp][i] = lg;
    lmax = fmaxf(lmax, lg);
  }
  lmax = warp_max(lmax);
  float lsum = 0.f;
  for (int i = lane; i < n; i += 32) lsum += expf(s_l[warp][i] - lmax);
  lsum = warp_sum(lsum);
  __syncwarp();
  for (int d = 0; d < k; ++d) {
    // jax.random.categorical = argmax(logits + gumbel), first index on ties
    float best = -INFINITY;
    int bi = -1;
    for (int i = lane; i < n; i += 32) {  // ascending i per lane: strict > keeps the first max
      float v = __fadd_rn(s_l[warp][i], gumbel[(r * n + i) * k + d]);
      if (bi < 0 || v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      inds[r * k + d] = bi;
      float prob = expf(s_l[warp][bi] - lmax) / lsum;
      w_new[r * k + d] = weights[r * n + bi] / (static_cast<float>(k) * prob + 1e-8f);
    }
  }
}

__global__ void resample_gather_kernel(const float* __restrict__ field, const int32_t* __restrict__ inds,
                                       int64_t R, int n, int k, int C, float* __restrict__ out) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= R * k * C) return;
  const int c = static_cast<int>(gid % C);
  const int64_t rk = gid / C;
  const int64_t r = rk / k;
  const int i = inds[rk];
  out[gid] = field[(r * n + i) * C + c];
}

inline unsigned ray_grid(int64_t R) { return static_cast<unsigned>((R + kRayWarps - 1) / kRayWarps); }

}  // namespace nrc

using namespace nrc;

#define NRC_STREAM static_cast<cudaStream_t>(stream)

extern "C" int32_t nrc_ray_alpha_weights_fwd(void* stream, const float* d_density, const float* d_tdist,
                                             const float* d_dirs, int64_t num_rays, int32_t n,
                                             int32_t opaque_background, float* d_weights, float* d_alpha,
                                             float* d_trans) {
  if (num_rays < 0 || n < 1 || n > kMaxN) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_density || !d_tdist || !d_dirs || !d_weights) return NRC_E_INVALID_ARG;
  alpha_weights_fwd_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(
      d_density, d_tdist, d_dirs, num_rays, n, opaque_background, d_weights, d_alpha, d_trans);
  return check_launch();
}

extern "C" int32_t nrc_ray_alpha_weights_bwd(void* stream, const float* d_density, const float* d_tdist,
                                             const float* d_dirs, const float* d_g_weights,
                                             const float* d_g_alpha, const float* d_g_trans,
                                             int64_t num_rays, int32_t n, float* d_g_density) {
  if (num_rays < 0 || n < 1 || n > kMaxN) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_density || !d_tdist || !d_dirs || !d_g_density) return NRC_E_INVALID_ARG;
  alpha_weights_bwd_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(
      d_density, d_tdist, d_dirs, d_g_weights, d_g_alpha, d_g_trans, num_rays, n, d_g_density);
  return check_launch();
}

extern "C" int32_t nrc_ray_sample_intervals(void* stream, const float* d_t, const float* d_w,
                                            const float* d_u01, const float* d_u_base, int64_t num_rays,
                                            int32_t m, int32_t n, float anneal, float padding,
                                            float max_jitter, float dom_lo, float dom_hi, float* d_t_new,
                                            int32_t* d_bin_idx) {
  // stepfun.py:230-231: num_samples <= 1 raises ValueError
  if (num_rays < 0 || m < 1 || m > kMaxN || n <= 1 || n > kMaxN) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_t || !d_w || !d_u01 || !d_u_base || !d_t_new) return NRC_E_INVALID_ARG;
  sample_intervals_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(
      d_t, d_w, d_u01, d_u_base, num_rays, m, n, anneal, padding, max_jitter, dom_lo, dom_hi, d_t_new,
      d_bin_idx, nullptr, nullptr, nullptr, nullptr, 0, 0.f, 0.f, nullptr, nullptr);
  return check_launch();
}

extern "C" int32_t nrc_ray_sample_cast(void* stream, const float* d_t, const float* d_w, const float* d_u01,
                                       const float* d_u_base, int64_t num_rays, int32_t m, int32_t n, float anneal,
                                       float padding, float max_jitter, float dom_lo, float dom_hi,
                                       const float* d_origins, const float* d_directions, const float* d_near,
                                       const float* d_far, int32_t warp_kind, float p, float premult, float* d_sdist_new,
                                       float* d_tdist, float* d_means) {
  if (num_rays < 0 || m < 1 || m > kMaxN || n <= 1 || n > kMaxN || (warp_kind != 0 && warp_kind != 1))
    return NRC_E_INVALID_ARG;
  if (warp_kind == 1 && (p == 1.0f || p == 0.0f || isinf(p))) return NRC_E_UNSUPPORTED;
  if (num_rays == 0) return NRC_OK;
  if (!d_t || !d_w || !d_u01 || !d_u_base || !d_sdist_new || !d_origins || !d_directions || !d_near || !d_far || !d_tdist)
    return NRC_E_INVALID_ARG;
  sample_intervals_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(
      d_t, d_w, d_u01, d_u_base, num_rays, m, n, anneal, padding, max_jitter, dom_lo, dom_hi, d_sdist_new, nullptr,
      d_origins, d_directions, d_near, d_far, warp_kind, p, premult, d_tdist, d_means);
  return check_launch();
}

// nrc_ray_sample_cast with compute_alpha_weights of the level being resampled folded into its head: the step function's
// weights come from that level's densities (d_density [R,m], d_tdist_prev [R,m+1], ray directions) and are written to
// d_weights_out [R,m] (may be NULL).  One launch instead of nrc_ray_alpha_weights_fwd + nrc_ray_sample_cast.
extern "C" int32_t nrc_ray_weights_sample_cast(void* stream, const float* d_t, const float* d_density, const float* d_tdist_prev,
                                               int32_t opaque_background, float* d_weights_out, const float* d_u01,
                                               const float* d_u_base, int64_t num_rays, int32_t m, int32_t n, float anneal,
                                               float padding, float max_jitter, float dom_lo, float dom_hi,
                                               const float* d_origins, const float* d_directions, const float* d_near,
                                               const float* d_far, int32_t warp_kind, float p, float premult,
                                               float* d_sdist_new, float* d_tdist, float* d_means) {
  if (num_rays < 0 || m < 1 || m > kMaxN || n <= 1 || n > kMaxN || (warp_kind != 0 && warp_kind != 1))
    return NRC_E_INVALID_ARG;
  if (warp_kind == 1 && (p == 1.0f || p == 0.0f || isinf(p))) return NRC_E_UNSUPPORTED;
  if (num_rays == 0) return NRC_OK;
  if (!d_t || !d_density || !d_tdist_prev || !d_u01 || !d_u_base || !d_sdist_new || !d_origins || !d_directions || !d_near ||
      !d_far || !d_tdist)
    return NRC_E_INVALID_ARG;
  sample_intervals_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(
      d_t, nullptr, d_u01, d_u_base, num_rays, m, n, anneal, padding, max_jitter, dom_lo, dom_hi, d_sdist_new, nullptr,
      d_origins, d_directions, d_near, d_far, warp_kind, p, premult, d_tdist, d_means, d_density, d_tdist_prev,
      opaque_background, d_weights_out);
  return check_launch();
}

extern "C" int32_t nrc_ray_cast(void* stream, const float* d_sdist, const float* d_origins,
                                const float* d_directions, const float* d_near, const float* d_far,
                                int64_t num_rays, int32_t n, int32_t warp_kind, float p, float premult,
                                float* d_tdist, float* d_means) {
  if (num_rays < 0 || n < 1 || (warp_kind != 0 && warp_kind != 1)) return NRC_E_INVALID_ARG;
  if (warp_kind == 1 && (p == 1.0f || p == 0.0f || isinf(p))) return NRC_E_UNSUPPORTED;
  if (num_rays == 0) return NRC_OK;
  if (!d_sdist || !d_origins || !d_directions || !d_near || !d_far || !d_tdist) return NRC_E_INVALID_ARG;
  int64_t total = num_rays * (n + 1);
  ray_cast_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, NRC_STREAM>>>(
      d_sdist, d_origins, d_directions, d_near, d_far, num_rays, n, warp_kind, p, premult, d_tdist, d_means);
  return check_launch();
}

extern "C" int32_t nrc_ray_cast_covs(void* stream, const float* d_tdist, const float* d_origins, const float* d_directions,
                                     const float* d_radii, int64_t num_rays, int32_t n, int32_t ray_shape, int32_t diag,
                                     float* d_covs, float* d_means) {
  using namespace nrc;
  if (!d_tdist || !d_directions || !d_radii || num_rays < 0 || n < 1 || (ray_shape != 0 && ray_shape != 1)) return NRC_E_INVALID_ARG;
  if (d_means && !d_origins) return NRC_E_INVALID_ARG;
  if (!d_covs && !d_means) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  const int64_t total = num_rays * n;
  ray_cast_covs_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, NRC_STREAM>>>(
      d_tdist, d_origins, d_directions, d_radii, num_rays, n, ray_shape, diag != 0, d_covs, d_means);
  return check_launch();
}

extern "C" int32_t nrc_ray_composite_fwd(void* stream, const float* d_values, const float* d_weights,
                                         int32_t k, const float* d_weights_nf, const float* d_tdist,
                                         const float* d_bg, int64_t num_rays, int32_t n, int32_t channels,
                                         int32_t has_rgb, float* d_out, float* d_acc, float* d_dist) {
  if (num_rays < 0 || n < 1 || n > kMaxN || k < 1 || channels < 0) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_weights || (channels > 0 && (!d_values || !d_out)) || (d_dist && !d_tdist)) return NRC_E_INVALID_ARG;
  if (!d_weights_nf) {
    if (k != n) return NRC_E_INVALID_ARG;
    d_weights_nf = d_weights;
  }
  composite_fwd_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(
      d_values, d_weights, k, d_weights_nf, d_tdist, d_bg, num_rays, n, channels, has_rgb, d_out, d_acc,
      d_dist);
  return check_launch();
}

extern "C" int32_t nrc_ray_composite_bwd(void* stream, const float* d_values, const float* d_weights,
                                         int32_t k, const float* d_weights_nf, const float* d_bg,
                                         const float* d_g_out, const float* d_g_acc, int64_t num_rays,
                                         int32_t n, int32_t channels, int32_t has_rgb, float* d_g_values,
                                         float* d_g_weights, float* d_g_weights_nf) {
  if (num_rays < 0 || n < 1 || n > kMaxN || k < 1 || channels < 0) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_weights || (channels > 0 && (!d_values || !d_g_out))) return NRC_E_INVALID_ARG;
  if (!d_weights_nf) {
    if (k != n || d_g_weights_nf) return NRC_E_INVALID_ARG;
    d_weights_nf = d_weights;
  }
  composite_bwd_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(
      d_values, d_weights, k, d_weights_nf, d_bg, d_g_out, d_g_acc, num_rays, n, channels, has_rgb,
      d_g_values, d_g_weights, d_g_weights_nf);
  return check_launch();
}

extern "C" int32_t nrc_ray_resample(void* stream, const float* d_weights, const float* d_gumbel,
                                    int64_t num_rays, int32_t n, int32_t k, float bias, float mult,
                                    int32_t* d_inds, float* d_w_new) {
  if (num_rays < 0 || n < 1 || n > kMaxN || k < 1) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_weights || !d_gumbel || !d_inds || !d_w_new) return NRC_E_INVALID_ARG;
  resample_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(d_weights, d_gumbel, num_rays, n, k,
                                                                      bias, mult, d_inds, d_w_new);
  return check_launch();
}

extern "C" int32_t nrc_ray_resample_gather(void* stream, const float* d_field, const int32_t* d_inds,
                                           int64_t num_rays, int32_t n, int32_t k, int32_t channels,
                                           float* d_out) {
  if (num_rays < 0 || n < 1 || k < 1 || channels < 1) return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_field || !d_inds || !d_out) return NRC_E_INVALID_ARG;
  int64_t total = num_rays * k * channels;
  resample_gather_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, NRC_STREAM>>>(
      d_field, d_inds, num_rays, n, k, channels, d_out);
  return check_launch();
}

extern "C" int32_t nrc_render_loss(void* stream, const float* d_values, const float* d_weights, const float* d_bg,
                                   const float* d_target, const float* d_mask, int64_t num_rays, int32_t n,
                                   float charb_padding, int32_t use_mask, float opaque_weight, float empty_weight,
                                   float* d_loss, float* d_out_rgb, float* d_acc, float* d_g_values, float* d_g_weights) {
  if (num_rays < 1 || n < 1) return NRC_E_INVALID_ARG;
  if (!d_values || !d_weights || !d_target || !d_loss || !d_out_rgb || !d_g_values || !d_g_weights) return NRC_E_INVALID_ARG;
  render_loss_kernel<<<ray_grid(num_rays), kRayThreads, 0, NRC_STREAM>>>(
      d_values, d_weights, d_bg, d_target, d_mask, num_rays, n, charb_padding, use_mask, opaque_weight, empty_weight,
      d_loss, d_out_rgb, d_acc, d_g_values, d_g_weights);
  return check_launch();
}
