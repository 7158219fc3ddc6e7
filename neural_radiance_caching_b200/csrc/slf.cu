// Surface-light-field MEMORY variant (second half): the per-ray stage between the distance network and
// the reflectance grid of `surface_lf_mem` (internal/models.py:813-833; configs/nerf_ngp_yobo.gin:97-165).
//
//   points   BaseSurfaceLightFieldMLP.predict_points (internal/surface_light_field.py:594-780) for the configured
//            branch (no voxel grid, no far samples, unsorted, no point offsets, env alpha on) + the head of __call__
//            (:899-913): per ray and distance sample i
//              off_i = raw[8i] * distance_scale / n * sigmoid(raw[8i+1] + distance_bias)
//              s_i   = fold(off_i + linspace(1e-8, 1 - 1e-8, n)[i])      fold: frac, mirrored on odd floors
//              t_i   = s_to_t(s_i)   (coord.construct_ray_warps over [distance_near, distance_far], power ladder)
//              mask  = [dn < t < df] [near < t < far];   t_i <- clip(t_i, dn, df)
//              x_i   = ref_warp_fn(origin + t_i dir)
//              w     = softmax(raw[8i+4]);  s_dist = sum s_i w_i;  w_i <- w_i mask_i env_alpha
//            env_rgb = softplus(premult raw[-4:-1] + rgb_bias), env_alpha = sigmoid(raw[-1] + alpha_bias).
//            One thread per ray; a CTA stages its rays' network rows in shared memory (coalesced loads, odd row
//            pitch), the backward writes the gradient rows back through the same tile.
//   reduce   the weighted feature sum over a ray's n reflectance-grid rows (:981) and its VJP.
#include "nrc_common.cuh"

namespace nrc {

constexpr int kSlfMaxSamples = 32;

struct SlfSample {
  float s, t, mask, sgn, dtds;   // folded s, clipped distance, validity, d s / d (off + start), d t / d s (0 where clipped)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }   // logaddexp(x, 0)

__device__ __forceinline__ float slf_start(int i, int n) {
  // jnp.linspace(1e-8, 1 - 1e-8, n) (:715-718)
  if (n == 1) return 1e-8f;
  const double step = (1.0 - 2e-8) / static_cast<double>(n - 1);
  return static_cast<float>(1e-8 + static_cast<double>(i) * step);
}

template <bool kGrad>
__device__ __forceinline__ SlfSample slf_sample(const nrc_slf_points_t& c, float o0, float o1, int i, float s_near,
                                                float s_far) {
  SlfSample r;
  const float n = static_cast<float>(c.num_distance_samples);
  const float sig = sigmoidf_(__fadd_rn(o1, c.distance_bias));
  const float off = __fmul_rn(__fdiv_rn(__fmul_rn(o0, c.distance_scale), n), sig);          // :652-657
  const float sp = __fadd_rn(off, slf_start(i, c.num_distance_samples));
  const float fl = floorf(sp);
  const float frac = __fsub_rn(sp, fl);
  const bool even = (static_cast<int>(fl) & 1) == 0;                                        // floor-mod 2 (:724-728)
  r.s = even ? frac : __fsub_rn(1.0f, frac);
  r.sgn = even ? 1.0f : -1.0f;
  const float u = __fadd_rn(__fmul_rn(r.s, s_far), __fmul_rn(__fsub_rn(1.0f, r.s), s_near));   // coord.py:256
  float t = c.warp_kind == 1 ? power_ladder_inv(u, c.warp_p, c.warp_premult) : u;
  r.mask = (t > c.distance_near && t < c.distance_far && t > c.near && t < c.far) ? 1.0f : 0.0f;   // :747-750
  r.dtds = 0.f;
  if (kGrad) {
    // jnp.clip passes the gradient strictly inside the range
    if (t > c.distance_near && t < c.distance_far) {
      float dtdu = 1.0f;
      if (c.warp_kind == 1) {
        // x = |p-1| ((ratio |y| + 1)^(1/p) - 1) / premult, ratio = p / |p-1|:  dx/dy = (ratio |y| + 1)^(1/p - 1) / premult
        const float p = c.warp_p;
        const float ratio = p / fabsf(p - 1.0f);
        float ymax = nextafterf((p - 1.0f) / p, -INFINITY);
        if (p >= 0.f) ymax = f32_max();
        const float yp = fabsf(u);
        dtdu = yp < ymax ? powf(ratio * yp + 1.0f, 1.0f / p - 1.0f) / c.warp_premult : 0.f;
      }
      r.dtds = dtdu * (s_far - s_near);
    }
  }
  r.t = fminf(fmaxf(t, c.distance_near), c.distance_far);                                   // :753
  return r;
}

__device__ __forceinline__ void slf_warp_ends(const nrc_slf_points_t& c, float& s_near, float& s_far) {
  s_near = c.distance_near; s_far = c.distance_far;
  if (c.warp_kind == 1) {
    s_near = power_ladder_fwd(s_near, c.warp_p, c.warp_premult);
    s_far = power_ladder_fwd(s_far, c.warp_p, c.warp_premult);
  }
}

__device__ __forceinline__ void load_tile(float* tile, const float* __restrict__ raw, int64_t ld, int64_t row0, int rows,
                                          int W) {
  const int pitch = W + 1;
  for (int idx = threadIdx.x; idx < rows * W; idx += blockDim.x) {
    const int r = idx / W, col = idx - r * W;
    tile[r * pitch + col] = raw[(row0 + r) * ld + col];
  }
}

__global__ void slf_points_fwd_kernel(nrc_slf_points_t c, const float* __restrict__ raw, int64_t ld,
                                      const float* __restrict__ origins, const float* __restrict__ dirs, int64_t P,
                                      float* __restrict__ points, float* __restrict__ weights, float* __restrict__ s_dist,
                                      float* __restrict__ distances, float* __restrict__ env_rgba) {
  extern __shared__ float tile[];
  const int n = c.num_distance_samples, W = 8 * n + 4, pitch = W + 1;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * blockDim.x;
  const int rows = static_cast<int>(min(static_cast<int64_t>(blockDim.x), P - row0));
  load_tile(tile, raw, ld, row0, rows, W);
  __syncthreads();
  if (static_cast<int>(threadIdx.x) >= rows) return;
  const int64_t p = row0 + threadIdx.x;
  const float* row = tile + threadIdx.x * pitch;
  float s_near, s_far;
  slf_warp_ends(c, s_near, s_far);
  const float ox = origins[3 * p], oy = origins[3 * p + 1], oz = origins[3 * p + 2];
  const float dx = dirs[3 * p], dy = dirs[3 * p + 1], dz = dirs[3 * p + 2];
  const float ea = sigmoidf_(__fadd_rn(row[W - 1], c.alpha_bias));                          // :630-633
#pragma unroll
  for (int a = 0; a < 3; ++a)
    env_rgba[4 * p + a] = softplusf_(__fadd_rn(__fmul_rn(c.rgb_premultiplier, row[W - 4 + a]), c.rgb_bias));   // :626-629
  env_rgba[4 * p + 3] = ea;
  float mx = -INFINITY;
  for (int i = 0; i < n; ++i) mx = fmaxf(mx, row[8 * i + 4]);
  float den = 0.f;
  for (int i = 0; i < n; ++i) den += expf(row[8 * i + 4] - mx);
  float sd = 0.f;
  for (int i = 0; i < n; ++i) {
    const SlfSample q = slf_sample<false>(c, row[8 * i], row[8 * i + 1], i, s_near, s_far);
    const float sm = expf(row[8 * i + 4] - mx) / den;                                       // jax.nn.softmax (:908)
    sd += q.s * sm;                                                                         // :909
    weights[p * n + i] = sm * q.mask * ea;                                                  // :910
    distances[p * n + i] = q.t;
    const float x0 = __fadd_rn(ox, __fmul_rn(q.t, dx)), x1 = __fadd_rn(oy, __fmul_rn(q.t, dy)),
                x2 = __fadd_rn(oz, __fmul_rn(q.t, dz));                                     // :771
    float z0, z1, z2;
    contract_point(c.ref_warp_c, x0, x1, x2, z0, z1, z2);                                   // ref_warp_fn (:905)
    float* o = points + (p * n + i) * 3;
    o[0] = z0; o[1] = z1; o[2] = z2;
  }
  s_dist[p] = sd;
}

__global__ void slf_points_bwd_kernel(nrc_slf_points_t c, const float* __restrict__ raw, int64_t ld,
                                      const float* __restrict__ origins, const float* __restrict__ dirs, int64_t P,
                                      const float* __restrict__ g_points, const float* __restrict__ g_weights,
                                      const float* __restrict__ g_s_dist, const float* __restrict__ g_distances,
                                      const float* __restrict__ g_env, float* __restrict__ g_raw) {
  extern __shared__ float tile[];
  const int n = c.num_distance_samples, W = 8 * n + 4, pitch = W + 1;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * blockDim.x;
  const int rows = static_cast<int>(min(static_cast<int64_t>(blockDim.x), P - row0));
  load_tile(tile, raw, ld, row0, rows, W);
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < rows) {
    const int64_t p = row0 + threadIdx.x;
    float* row = tile + threadIdx.x * pitch;
    float s_near, s_far;
    slf_warp_ends(c, s_near, s_far);
    const float ox = origins[3 * p], oy = origins[3 * p + 1], oz = origins[3 * p + 2];
    const float dx = dirs[3 * p], dy = dirs[3 * p + 1], dz = dirs[3 * p + 2];
    const float ea = sigmoidf_(row[W - 1] + c.alpha_bias);
    const float gsd = g_s_dist ? g_s_dist[p] : 0.f;
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) mx = fmaxf(mx, row[8 * i + 4]);
    float den = 0.f;
    for (int i = 0; i < n; ++i) den += expf(row[8 * i + 4] - mx);
    // first pass: sum_j sm_j g_sm_j (softmax VJP) and the env-alpha gradient
    float dot = 0.f, g_ea = g_env ? g_env[4 * p + 3] : 0.f;
    for (int i = 0; i < n; ++i) {
      const SlfSample q = slf_sample<false>(c, row[8 * i], row[8 * i + 1], i, s_near, s_far);
      const float sm = expf(row[8 * i + 4] - mx) / den;
      const float gw = g_weights ? g_weights[p * n + i] : 0.f;
      const float g_sm = gw * q.mask * ea + gsd * q.s;
      dot += sm * g_sm;
      g_ea += gw * sm * q.mask;
    }
    // second pass: per-sample columns
    for (int i = 0; i < n; ++i) {
      const float o0 = row[8 * i], o1 = row[8 * i + 1], rw = row[8 * i + 4];
      const SlfSample q = slf_sample<true>(c, o0, o1, i, s_near, s_far);
      const float sm = expf(rw - mx) / den;
      const float gw = g_weights ? g_weights[p * n + i] : 0.f;
      const float g_sm = gw * q.mask * ea + gsd * q.s;
      const float g_rw = sm * (g_sm - dot);
      // distance path: points -> t, plus the distances output
      float g_t = g_distances ? g_distances[p * n + i] : 0.f;
      if (g_points) {
        const float* gp = g_points + (p * n + i) * 3;
        float a0, a1, a2;
        contract_vjp(c.ref_warp_c, ox + q.t * dx, oy + q.t * dy, oz + q.t * dz, gp[0], gp[1], gp[2], a0, a1, a2);
        g_t += a0 * dx + a1 * dy + a2 * dz;
      }
      const float g_s = gsd * sm + g_t * q.dtds;
      const float g_off = g_s * q.sgn;
      const float sig = sigmoidf_(o1 + c.distance_bias);
      const float k = c.distance_scale / static_cast<float>(n);
#pragma unroll
      for (int col = 0; col < 8; ++col) row[8 * i + col] = 0.f;
      row[8 * i] = g_off * k * sig;
      row[8 * i + 1] = g_off * o0 * k * sig * (1.0f - sig);
      row[8 * i + 4] = g_rw;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float x = c.rgb_premultiplier * row[W - 4 + a] + c.rgb_bias;
      row[W - 4 + a] = g_env ? g_env[4 * p + a] * sigmoidf_(x) * c.rgb_premultiplier : 0.f;
    }
    row[W - 1] = g_ea * ea * (1.0f - ea);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < rows * W; idx += blockDim.x) {
    const int r = idx / W, col = idx - r * W;
    g_raw[(row0 + r) * W + col] = tile[r * pitch + col];
  }
}

__global__ void slf_reduce_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ w, int64_t P, int n, int F,
                                      float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= P * F) return;
  const int64_t p = i / F;
  const int f = static_cast<int>(i - p * F);
  float acc = 0.f;
  for (int s = 0; s < n; ++s) acc += feat[(p * n + s) * F + f] * w[p * n + s];
  out[i] = acc;
}

// one warp per (ray, sample) row: lanes stride the features, the weight gradient is a warp sum
__global__ void slf_reduce_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                                      const float* __restrict__ g_out, int64_t P, int n, int F, float* __restrict__ g_feat,
                                      float* __restrict__ g_w) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= P * n) return;
  const int64_t p = row / n;
  const float wi = w[row];
  float acc = 0.f;
  for (int f = lane; f < F; f += 32) {
    const float g = g_out[p * F + f];
    acc += g * feat[row * F + f];
    g_feat[row * F + f] = g * wi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) g_w[row] = acc;
}

static int slf_rows_per_cta(int n) {
  const int W = 8 * n + 4;
  for (int rows = 128; rows >= 32; rows >>= 1)
    if (static_cast<size_t>(rows) * (W + 1) * sizeof(float) <= 48u * 1024u) return rows;
  return 0;
}

static int32_t slf_check(const nrc_slf_points_t* cfg, const void* a, const void* b, const void* c2, int64_t ld, int64_t P) {
  if (!cfg || !a || !b || !c2 || P < 0) return NRC_E_INVALID_ARG;
  if (cfg->num_distance_samples < 1 || ld < 8 * static_cast<int64_t>(cfg->num_distance_samples) + 4) return NRC_E_INVALID_ARG;
  if (cfg->num_distance_samples > kSlfMaxSamples || (cfg->warp_kind != 0 && cfg->warp_kind != 1)) return NRC_E_UNSUPPORTED;
  return NRC_OK;
}

}  // namespace nrc

extern "C" {

int32_t nrc_slf_points_fwd(void* stream, const nrc_slf_points_t* cfg, const float* d_raw, int64_t ld_raw,
                           const float* d_origins, const float* d_refdirs, int64_t num_points, float* d_points,
                           float* d_weights, float* d_s_dist, float* d_distances, float* d_env_rgba) {
  using namespace nrc;
  int32_t st = slf_check(cfg, d_raw, d_origins, d_refdirs, ld_raw, num_points);
  if (st != NRC_OK) return st;
  if (!d_points || !d_weights || !d_s_dist || !d_distances || !d_env_rgba) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  const int rows = slf_rows_per_cta(cfg->num_distance_samples);
  const size_t smem = static_cast<size_t>(rows) * (8 * cfg->num_distance_samples + 5) * sizeof(float);
  const unsigned grid = static_cast<unsigned>((num_points + rows - 1) / rows);
  slf_points_fwd_kernel<<<grid, rows, smem, static_cast<cudaStream_t>(stream)>>>(
      *cfg, d_raw, ld_raw, d_origins, d_refdirs, num_points, d_points, d_weights, d_s_dist, d_distances, d_env_rgba);
  return check_launch();
}

int32_t nrc_slf_points_bwd(void* stream, const nrc_slf_points_t* cfg, const float* d_raw, int64_t ld_raw,
                           const float* d_origins, const float* d_refdirs, int64_t num_points, const float* d_g_points,
                           const float* d_g_weights, const float* d_g_s_dist, const float* d_g_distances,
                           const float* d_g_env_rgba, float* d_g_raw) {
  using namespace nrc;
  int32_t st = slf_check(cfg, d_raw, d_origins, d_refdirs, ld_raw, num_points);
  if (st != NRC_OK) return st;
  if (!d_g_raw) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  const int rows = slf_rows_per_cta(cfg->num_distance_samples);
  const size_t smem = static_cast<size_t>(rows) * (8 * cfg->num_distance_samples + 5) * sizeof(float);
  const unsigned grid = static_cast<unsigned>((num_points + rows - 1) / rows);
  slf_points_bwd_kernel<<<grid, rows, smem, static_cast<cudaStream_t>(stream)>>>(
      *cfg, d_raw, ld_raw, d_origins, d_refdirs, num_points, d_g_points, d_g_weights, d_g_s_dist, d_g_distances,
      d_g_env_rgba, d_g_raw);
  return check_launch();
}

int32_t nrc_slf_reduce_fwd(void* stream, const float* d_feat, const float* d_weights, int64_t num_points,
                           int32_t num_samples, int32_t num_features, float* d_out) {
  using namespace nrc;
  if (!d_feat || !d_weights || !d_out || num_points < 0 || num_samples < 1 || num_features < 1) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  const int64_t total = num_points * num_features;
  slf_reduce_fwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_feat, d_weights, num_points, num_samples, num_features, d_out);
  return check_launch();
}

int32_t nrc_slf_reduce_bwd(void* stream, const float* d_feat, const float* d_weights, const float* d_g_out,
                           int64_t num_points, int32_t num_samples, int32_t num_features, float* d_g_feat,
                           float* d_g_weights) {
  using namespace nrc;
  if (!d_feat || !d_weights || !d_g_out || !d_g_feat || !d_g_weights || num_points < 0 || num_samples < 1 ||
      num_features < 1)
    return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  const int64_t warps = num_points * num_samples;
  slf_reduce_bwd_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_feat, d_weights, d_g_out, num_points, num_samples, num_features, d_g_feat, d_g_weights);
  return check_launch();
}

}  // extern "C"
