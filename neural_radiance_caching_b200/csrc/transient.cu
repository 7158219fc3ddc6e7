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
constexpr int kTrhThreads = 512;     // 16 warps: 8 column-tile lanes x 2 row tiles (one 16-row MMA tile per warp)
constexpr int kTrhColWarps = 8;
constexpr int kTrhRows = 32;          // samples per pass (two MMA row tiles)
constexpr int kTrhK = 64 + 128;       // packed K: diffuse hidden | specular hidden

struct TrHeadParams {
  int n, n_bins, C;
  float exposure_time, shift, diffuse_bias, spec_bias, spec_premult, spec_max, indirect_scale, bin_zero_threshold_light,
      light_near, rgb_max, dark_level;
  int light_zero;
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Both heads' last-layer kernels as ONE bf16 operand image [N][192]: row e = output column, K = [diffuse 64 | specular
// 128]; inside every 16-wide K block the elements are stored in the order an mma.sync B fragment wants them
// (thread t of a quad: k = 2t, 2t+1, 2t+8, 2t+9), so a fragment is one 8-byte load and a quad reads a whole 32-byte sector.
// The fp32 kernels are read 0.8 M times per call otherwise (every CTA needs all of both layers): 3.3 GB of L2 traffic
// at 1024 rays, which bound the first version of the fused kernel at 1.36 ms.
__global__ void transient_pack_heads_kernel(const float* __restrict__ w_d, int64_t ld_d, const float* __restrict__ w_s, int64_t ld_s,
                                            int N, __nv_bfloat16* __restrict__ packed) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(N) * kTrhK) return;
  const int e = static_cast<int>(i % N);            // consecutive threads: consecutive columns (coalesced reads)
  const int k = static_cast<int>(i / N);
  float v = 0.f;
  if (k < 64) { if (w_d) v = w_d[static_cast<int64_t>(k) * ld_d + e]; }
  else if (w_s) v = w_s[static_cast<int64_t>(k - 64) * ld_s + e];
  const int blk = k >> 4, r = k & 15;
  const int t = (r & 7) >> 1, pos = 4 * t + (r & 1) + (r >= 8 ? 2 : 0);
  packed[static_cast<int64_t>(e) * kTrhK + blk * 16 + pos] = __float2bfloat16_rn(v);
}

// softplus on the special-function units (the fused kernel evaluates it 134 k times per ray; its results are staged as
// bf16): log(1 + e^x) with __expf / __logf, absolute error <= 1.2e-7 (the rounding of 1 + e^x) - below bf16 resolution
// of every value that survives the weighted reduction.
__device__ __forceinline__ float softplus_fast(float x) { return x > 15.f ? x : __logf(1.f + __expf(x)); }

template <int C_>
__global__ void __launch_bounds__(kTrhThreads, 1)
transient_head_render_kernel(const float* __restrict__ h_d, const float* __restrict__ b_d, const float* __restrict__ h_s,
                             const float* __restrict__ b_s, const __nv_bfloat16* __restrict__ w_packed,
                             const float* __restrict__ spec_scale, const float* __restrict__ weights,
                             const float* __restrict__ ray_dists, const float* __restrict__ light_dists,
                             const float* __restrict__ cam_dists, int64_t R, TrHeadParams p, const float* __restrict__ t_direct,
                             float* __restrict__ t_indirect, float* __restrict__ rgb) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int64_t ray = blockIdx.x;
  constexpr int C = C_;
  const int N = p.n_bins * C;
  const int Ns = (N + 1) & ~1;
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(smem);                 // [32][Ns] activated, masked values
  float* s_w = reinterpret_cast<float*>(stage + static_cast<size_t>(kTrhRows) * Ns);
  float* s_move = s_w + p.n;                                                     // per sample: weight, bins_move, light, cam
  float* s_light = s_move + p.n;
  float* s_cam = s_light + p.n;
  float* s_scale = s_cam + p.n;                                                  // [n][C]
  float* s_wa = s_scale + p.n * C + ((p.n * C) & 1);                             // weight x (1 - t): tap at bin b + off
  float* s_wb = s_wa + p.n;                                                      // weight x t:       tap at bin b + off + 1
  int* s_off = reinterpret_cast<int*>(s_wb + p.n);
  int* s_lo = s_off + p.n;                                                       // valid bins of zero_invalid_bins: [lo, hi]
  int* s_hi = s_lo + p.n;
  float* s_bd = reinterpret_cast<float*>(s_hi + p.n + (p.n & 1));                // [N] diffuse bias + diffuse_bias
  float* s_bs = s_bd + Ns;                                                       // [N] specular bias
  __nv_bfloat16* h_sm = reinterpret_cast<__nv_bfloat16*>(s_bs + Ns);             // [32][192 + 8]
  constexpr int HS = kTrhK + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float max_dists = static_cast<float>(p.n_bins - 1) * p.exposure_time;
  for (int e = threadIdx.x; e < N; e += blockDim.x) {
    s_bd[e] = h_d ? __ldg(b_d + e) + p.diffuse_bias : 0.f;
    s_bs[e] = h_s ? __ldg(b_s + e) : 0.f;
  }
  for (int s = threadIdx.x; s < p.n; s += blockDim.x) {
    const int64_t i = ray * p.n + s;
    const float w = weights[i], light = light_dists[i], cam = cam_dists[i];
    const float move = (ray_dists[i] + p.shift) / p.exposure_time;
    s_w[s] = w; s_move[s] = move; s_light[s] = light; s_cam[s] = cam;
    for (int c = 0; c < C; ++c) s_scale[s * C + c] = spec_scale ? spec_scale[i * C + c] : 0.f;
    // order-1 shift: the value at y = b - move lies between bins b + off and b + off + 1 with the SAME fraction for every b
    const float yf = floorf(-move);
    const float tt = -move - yf;
    s_off[s] = static_cast<int>(yf);
    s_wa[s] = w * (1.0f - tt);
    s_wb[s] = w * tt;
    // zero_invalid_bins (render_utils.py:1699-1767): both predicates are monotone in the bin -> one valid range per sample,
    // found with the exact float predicates of the unfused kernel
    int lo = 0, hi = p.n_bins;
    while (lo < hi) {                                       // first bin that is NOT too close to the light
      const int mid = (lo + hi) >> 1;
      if ((static_cast<float>(mid) + p.bin_zero_threshold_light) * p.exposure_time < light) lo = mid + 1; else hi = mid;
    }
    int flo = 0, fhi = p.n_bins;
    while (flo < fhi) {                                     // first bin that is too far
      const int mid = (flo + fhi) >> 1;
      if ((static_cast<float>(mid) * p.exposure_time + cam) > max_dists) fhi = mid; else flo = mid + 1;
    }
    if (p.light_zero && light < p.light_near) lo = p.n_bins;
    s_lo[s] = lo;
    s_hi[s] = flo - 1;
  }
  constexpr int kMaxPer = 8;     // output elements per thread (N <= 8 * 512)
  float out[kMaxPer];
#pragma unroll
  for (int q = 0; q < kMaxPer; ++q) out[q] = 0.f;
  const int n_tiles = (N + 7) >> 3;
  for (int row0 = 0; row0 < p.n; row0 += kTrhRows) {
    __syncthreads();            // the previous pass's gather is done with `stage`; the per-sample arrays are written
    {
      // hidden activations of the pass's 32 samples -> bf16 rows: every load issued before the first conversion (one L2
      // round trip instead of one per element)
      constexpr int kPer = kTrhRows * kTrhK / kTrhThreads;
      static_assert(kTrhRows * kTrhK % kTrhThreads == 0, "staging loop is unrolled");
      float hv[kPer];
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int idx = threadIdx.x + q * kTrhThreads;
        const int r = idx / kTrhK, k = idx - r * kTrhK, s = row0 + r;
        float v = 0.f;
        if (s < p.n) {
          if (k < 64) { if (h_d) v = __ldg(h_d + (ray * p.n + s) * 64 + k); }
          else if (h_s) v = __ldg(h_s + (ray * p.n + s) * 128 + (k - 64));
        }
        hv[q] = v;
      }
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int idx = threadIdx.x + q * kTrhThreads;
        const int r = idx / kTrhK, k = idx - r * kTrhK;
        h_sm[r * HS + k] = __float2bfloat16_rn(hv[q]);
      }
    }
    __syncthreads();
    // A fragments of the pass: this warp's row tile (16 samples) x 12 K blocks, kept in registers over all column tiles
    const int mt = warp / kTrhColWarps, cw = warp % kTrhColWarps;
    uint32_t a[kTrhK / 16][4];
#pragma unroll
    for (int kk = 0; kk < kTrhK / 16; ++kk) {
      const __nv_bfloat16* r0 = h_sm + (mt * 16 + g) * HS + kk * 16 + 2 * t;
      const __nv_bfloat16* r1 = r0 + 8 * HS;
      a[kk][0] = *reinterpret_cast<const uint32_t*>(r0);
      a[kk][1] = *reinterpret_cast<const uint32_t*>(r1);
      a[kk][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
      a[kk][3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
    }
    // rows of this thread in the pass: valid-bin range (empty beyond the last sample) and the specular scale
    int row_lo[2], row_hi[2];
    float row_sc[2][C];
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
      const int sr = row0 + mt * 16 + g + 8 * ri;
      const bool live = sr < p.n;
      row_lo[ri] = live ? s_lo[sr] : 1;
      row_hi[ri] = live ? s_hi[sr] : 0;
#pragma unroll
      for (int c = 0; c < C; ++c) row_sc[ri][c] = live ? s_scale[sr * C + c] : 0.f;
    }
    // B fragments of a column tile: 12 8-byte loads per thread (its output column nb, K blocks 0..11), software-pipelined -
    // the next tile's loads are in flight (L2 latency) while this tile's MMAs and epilogue run
    uint2 bnext[kTrhK / 16];
    auto load_b = [&](int tile) {
      const int nb = min(tile * 8 + g, N - 1);     // output column this thread's B fragments feed (clamped: unused beyond N)
      const uint2* wrow = reinterpret_cast<const uint2*>(w_packed + static_cast<int64_t>(nb) * kTrhK) + t;
#pragma unroll
      for (int kk = 0; kk < kTrhK / 16; ++kk) bnext[kk] = __ldg(wrow + kk * 4);
    };
    if (cw < n_tiles) load_b(cw);
    for (int tile = cw; tile < n_tiles; tile += kTrhColWarps) {
      uint2 bcur[kTrhK / 16];
#pragma unroll
      for (int kk = 0; kk < kTrhK / 16; ++kk) bcur[kk] = bnext[kk];
      if (tile + kTrhColWarps < n_tiles) load_b(tile + kTrhColWarps);
      float accd[4], accs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) accd[i] = accs[i] = 0.f;
#pragma unroll
      for (int kk = 0; kk < kTrhK / 16; ++kk) {
        if (kk < 4) mma_bf16_16816(accd, a[kk], bcur[kk].x, bcur[kk].y);
        else mma_bf16_16816(accs, a[kk], bcur[kk].x, bcur[kk].y);
      }
      // epilogue: this thread owns columns e0, e0 + 1 of rows g, g + 8, 16 + g, 24 + g; the per-column and per-row terms are
      // hoisted, the two bf16 values of a row leave as one 4-byte store
      const int e0 = tile * 8 + 2 * t;
      if (e0 < N) {
        int binj[2], cj[2];
        float bdj[2], bsj[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int e = min(e0 + j, N - 1);
          binj[j] = e / C;
          cj[j] = e - binj[j] * C;
          bdj[j] = s_bd[e];
          bsj[j] = s_bs[e];
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int ri = half, row = mt * 16 + g + 8 * half;
            float val[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float v = 0.f;
              if (binj[j] >= row_lo[ri] && binj[j] <= row_hi[ri]) {
                if (h_d) v += fminf(fmaxf(softplus_fast(accd[half * 2 + j] + bdj[j]) * p.indirect_scale, 0.f), p.rgb_max);
                if (h_s) {
                  const float ref = fminf(fmaxf(softplus_fast(p.spec_premult * (accs[half * 2 + j] + bsj[j]) + p.spec_bias), 0.f), p.spec_max);
                  const float sc = cj[j] == 0 ? row_sc[ri][0] : (cj[j] == 1 ? row_sc[ri][1] : row_sc[ri][C - 1]);
                  v += fminf(fmaxf(sc * ref * p.indirect_scale, 0.f), p.rgb_max);
                }
              }
              val[j] = v;
            }
            __nv_bfloat16* dst = stage + static_cast<size_t>(row) * Ns + e0;
            if (e0 + 1 < N) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(val[0], val[1]);
            else *dst = __float2bfloat16_rn(val[0]);
          }
      }
    }
    __syncthreads();
    // shift_map_coordinates (order-1 map_coordinates along the bin axis, mode 'constant') + the weighted reduction,
    // samples in ascending order like transient_indirect_kernel
#pragma unroll
    for (int q = 0; q < kMaxPer; ++q) {
      const int e = threadIdx.x + q * kTrhThreads;
      if (e >= N) continue;
      const int b = e / C;
      float acc = out[q];
      const int s_end = min(kTrhRows, p.n - row0);
      for (int r = 0; r < s_end; ++r) {
        const int sidx = row0 + r;
        const int bin0 = b + s_off[sidx];
        const __nv_bfloat16* st = stage + static_cast<size_t>(r) * Ns + e + s_off[sidx] * C;
        const float v0 = (bin0 >= 0 && bin0 < p.n_bins) ? __bfloat162float(st[0]) : 0.f;
        const float v1 = (bin0 >= -1 && bin0 + 1 < p.n_bins) ? __bfloat162float(st[C]) : 0.f;
        acc = fmaf(s_wa[sidx], v0, fmaf(s_wb[sidx], v1, acc));
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

// ------------------------------------------------------------------------------------------------
// VJP of nrc_transient_render_fwd (training of the time-resolved cache): given the gradients of transient_direct and
// transient_indirect [R, n_bins, C] (rgb = direct + indirect + dark_level: the caller adds rgb's gradient to both),
//   direct splat   g_direct_rgbs[i] = w_i (w_lo G[i_lo] + w_hi G[i_hi]),  g_w_i += direct_i . (w_lo G[i_lo] + w_hi G[i_hi])
//   indirect       out[b] = sum_s w_s ((1 - t_s) val_s[b - m_s] + t_s val_s[b - m_s + 1])   (order-1 shift, constant per sample)
//                  => A_s[bin] = sum over the output bins b whose taps hit `bin` of tap weight x G[b]
//                     g_val_s[bin] = w_s A_s[bin] [bin valid],   g_w_s += sum_bin,c val_s A_s
//                  val = clip(softplus(raw + bias) scale, 0, max) + clip(spec_scale specular scale, 0, max)
// The sample distances are stop-gradient inputs (the sampler detaches its fenceposts, internal/sampling.py:353-354).
__global__ void transient_direct_bwd_kernel(const float* __restrict__ direct, const float* __restrict__ weights,
                                            const float* __restrict__ ray_dists, const float* __restrict__ light_dists,
                                            int64_t R, TransientParams p, const float* __restrict__ g_out,
                                            float* __restrict__ g_direct, float* __restrict__ g_w) {
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
  float gw = 0.f;
  for (int c = 0; c < p.C; ++c) {
    float a = 0.f;
    if (g_out) {
      if (i_lo >= 0 && i_lo < total) a += w_lo * g_out[i_lo * p.C + c];
      if (i_hi >= 0 && i_hi < total) a += w_hi * g_out[i_hi * p.C + c];
    }
    g_direct[i * p.C + c] = w * a;
    gw += direct[i * p.C + c] * a;
  }
  g_w[i] = gw;     // the indirect kernel (launched after this one) adds its part
}

__global__ void __launch_bounds__(256)
transient_indirect_bwd_kernel(const float* __restrict__ diffuse_raw, const float* __restrict__ specular,
                              const float* __restrict__ spec_scale, const float* __restrict__ weights,
                              const float* __restrict__ ray_dists, const float* __restrict__ light_dists,
                              const float* __restrict__ cam_dists, int64_t R, TransientParams p,
                              const float* __restrict__ g_out, float* __restrict__ g_diffuse_raw,
                              float* __restrict__ g_specular, float* __restrict__ g_spec_scale, float* __restrict__ g_w) {
  const int64_t ray = blockIdx.x;
  const int BC = p.n_bins * p.C;
  const float max_dists = static_cast<float>(p.n_bins - 1) * p.exposure_time;
  extern __shared__ float sm[];
  float* s_g = sm;                       // [n_bins * C] upstream gradient of this ray
  float* s_w = s_g + BC;
  float* s_move = s_w + p.n;
  float* s_light = s_move + p.n;
  float* s_cam = s_light + p.n;
  float* s_scale = s_cam + p.n;          // [n * C]
  float* s_acc = s_scale + p.n * p.C;    // [n * (1 + C)]: g_w, g_spec_scale
  for (int e = threadIdx.x; e < BC; e += blockDim.x) s_g[e] = g_out ? g_out[ray * BC + e] : 0.f;
  for (int s = threadIdx.x; s < p.n; s += blockDim.x) {
    const int64_t i = ray * p.n + s;
    s_w[s] = weights[i];
    s_move[s] = (ray_dists[i] + p.shift) / p.exposure_time;
    s_light[s] = light_dists[i];
    s_cam[s] = cam_dists[i];
    for (int c = 0; c < p.C; ++c) s_scale[s * p.C + c] = spec_scale ? spec_scale[i * p.C + c] : 0.f;
  }
  for (int k = threadIdx.x; k < p.n * (1 + p.C); k += blockDim.x) s_acc[k] = 0.f;
  __syncthreads();
  for (int s = 0; s < p.n; ++s) {
    const float move = s_move[s];
    const int m0 = static_cast<int>(ceilf(move));
    bool sample_ok = true;
    if (p.light_zero) sample_ok = !(s_light[s] < p.light_near);
    float part_w = 0.f, part_ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = threadIdx.x; e < BC; e += blockDim.x) {
      const int bin = e / p.C, c = e - bin * p.C;
      const int64_t idx = ((ray * p.n + s) * p.n_bins + bin) * p.C + c;
      const float fb = static_cast<float>(bin);
      bool ok = sample_ok && !((fb + p.bin_zero_threshold_light) * p.exposure_time < s_light[s]);
      ok = ok && !((fb * p.exposure_time + s_cam[s]) > max_dists);
      float gd = 0.f, gs = 0.f;
      if (ok) {
        // output bins whose two taps (y0, y0 + 1) of the forward pass land on `bin`
        float A = 0.f;
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
          const int b = bin + m0 + j;
          if (b < 0 || b >= p.n_bins) continue;
          const float y = static_cast<float>(b) - move;
          const float y0f = floorf(y);
          const float t = y - y0f;
          const int y0 = static_cast<int>(y0f);
          if (y0 == bin) A += (1.0f - t) * s_g[b * p.C + c];
          else if (y0 + 1 == bin) A += t * s_g[b * p.C + c];
        }
        float val = 0.f;
        if (diffuse_raw) {
          const float x = diffuse_raw[idx] + p.diffuse_bias;
          const float v = softplus_t(x) * p.indirect_scale;
          val += fminf(fmaxf(v, 0.f), p.rgb_max);
          if (v >= 0.f && v <= p.rgb_max) gd = s_w[s] * A * p.indirect_scale / (1.0f + expf(-x));
        }
        if (specular) {
          const float sc = s_scale[s * p.C + c], sp = specular[idx];
          const float v = sc * sp * p.indirect_scale;
          val += fminf(fmaxf(v, 0.f), p.rgb_max);
          if (v >= 0.f && v <= p.rgb_max) {
            gs = s_w[s] * A * sc * p.indirect_scale;
            const float gsc = s_w[s] * A * sp * p.indirect_scale;
            part_ss[0] += c == 0 ? gsc : 0.f; part_ss[1] += c == 1 ? gsc : 0.f;
            part_ss[2] += c == 2 ? gsc : 0.f; part_ss[3] += c == 3 ? gsc : 0.f;
          }
        }
        part_w += val * A;
      }
      if (g_diffuse_raw) g_diffuse_raw[idx] = gd;
      if (g_specular) g_specular[idx] = gs;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      part_w += __shfl_xor_sync(0xffffffffu, part_w, o);
#pragma unroll
      for (int c = 0; c < 4; ++c) part_ss[c] += __shfl_xor_sync(0xffffffffu, part_ss[c], o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&s_acc[s * (1 + p.C)], part_w);
      for (int c = 0; c < p.C; ++c) atomicAdd(&s_acc[s * (1 + p.C) + 1 + c], part_ss[c]);
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < p.n; s += blockDim.x) {
    const int64_t i = ray * p.n + s;
    g_w[i] += s_acc[s * (1 + p.C)];
    if (g_spec_scale)
      for (int c = 0; c < p.C; ++c) g_spec_scale[i * p.C + c] = s_acc[s * (1 + p.C) + 1 + c];
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
    void* d_w_packed, int32_t repack, float* d_transient_direct, float* d_transient_indirect, float* d_rgb) {
  if (num_rays < 0 || n < 1 || n > 1024 || n_bins < 1 || channels < 1 || channels > 4 || !(exposure_time > 0.f))
    return NRC_E_INVALID_ARG;
  const int64_t N = static_cast<int64_t>(n_bins) * channels;
  if (N > 8 * kTrhThreads) return NRC_E_UNSUPPORTED;
  if ((d_h_diffuse && k_diffuse != 64) || (d_h_specular && k_specular != 128)) return NRC_E_UNSUPPORTED;
  if (num_rays == 0) return NRC_OK;
  if (!d_direct_rgbs || !d_weights || !d_ray_dists || !d_light_dists || !d_cam_dists || !d_transient_direct ||
      !d_transient_indirect || !d_rgb || !d_w_packed || (!d_h_diffuse && !d_h_specular) ||
      (d_h_diffuse && (!d_w_diffuse || !d_b_diffuse)) || (d_h_specular && (!d_w_specular || !d_b_specular || !d_spec_scale)) ||
      ld_w_diffuse < (d_h_diffuse ? N : 0) || ld_w_specular < (d_h_specular ? N : 0))
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
  if (repack) {
    const int64_t tw = N * kTrhK;
    transient_pack_heads_kernel<<<static_cast<unsigned>((tw + 255) / 256), 256, 0, s>>>(
        d_h_diffuse ? d_w_diffuse : nullptr, ld_w_diffuse, d_h_specular ? d_w_specular : nullptr, ld_w_specular,
        static_cast<int>(N), static_cast<__nv_bfloat16*>(d_w_packed));
  }
  TrHeadParams p{};
  p.n = n; p.n_bins = n_bins; p.C = channels;
  p.exposure_time = exposure_time; p.shift = shift; p.diffuse_bias = diffuse_bias; p.spec_bias = spec_bias;
  p.spec_premult = spec_premult; p.spec_max = spec_max; p.indirect_scale = indirect_scale;
  p.bin_zero_threshold_light = bin_zero_threshold_light; p.light_near = light_near; p.rgb_max = rgb_max;
  p.dark_level = dark_level; p.light_zero = light_zero;
  const size_t Ns = static_cast<size_t>((N + 1) & ~1);
  const size_t smem = static_cast<size_t>(kTrhRows) * Ns * 2 + static_cast<size_t>(n) * (4 + channels + 1 + 5 + 1) * sizeof(float) +
                      2 * Ns * sizeof(float) + static_cast<size_t>(kTrhRows) * (kTrhK + 8) * 2 + 16;
  if (smem > 227 * 1024) return NRC_E_UNSUPPORTED;
  if (channels != 3) return NRC_E_UNSUPPORTED;   /* the fused kernel is compiled for RGB histograms */
  if (const int32_t st_attr = ensure_dynamic_smem<transient_head_render_kernel<3>>(227 * 1024); st_attr != NRC_OK)
    return st_attr;   /* the size varies with n_bins: opt in to the maximum once */
  transient_head_render_kernel<3><<<static_cast<unsigned>(num_rays), kTrhThreads, smem, s>>>(
      d_h_diffuse, d_b_diffuse, d_h_specular, d_b_specular, static_cast<const __nv_bfloat16*>(d_w_packed), d_spec_scale, d_weights,
      d_ray_dists, d_light_dists, d_cam_dists, num_rays, p, d_transient_direct, d_transient_indirect, d_rgb);
  return check_launch();
}

extern "C" int32_t nrc_transient_render_bwd(void* stream, const float* d_direct_rgbs, const float* d_diffuse_raw,
                                            const float* d_specular, const float* d_spec_scale, const float* d_weights,
                                            const float* d_ray_dists, const float* d_light_dists, const float* d_cam_dists,
                                            int64_t num_rays, int32_t n, int32_t n_bins, int32_t channels,
                                            float exposure_time, float shift, float diffuse_bias, float indirect_scale,
                                            float bin_zero_threshold_light, int32_t light_zero, float light_near,
                                            float rgb_max, const float* d_g_transient_direct,
                                            const float* d_g_transient_indirect, float* d_g_direct_rgbs,
                                            float* d_g_diffuse_raw, float* d_g_specular, float* d_g_spec_scale,
                                            float* d_g_weights) {
  if (num_rays < 0 || n < 1 || n > 1024 || n_bins < 1 || channels < 1 || channels > 4 || !(exposure_time > 0.f))
    return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_direct_rgbs || !d_weights || !d_ray_dists || !d_light_dists || !d_cam_dists || !d_g_direct_rgbs || !d_g_weights ||
      (d_specular && !d_spec_scale) || (d_g_diffuse_raw && !d_diffuse_raw) || (d_g_specular && !d_specular) ||
      (d_g_spec_scale && !d_specular))
    return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TransientParams p{n, n_bins, channels, exposure_time, shift, diffuse_bias, indirect_scale, bin_zero_threshold_light,
                    light_near, rgb_max, 0.f, light_zero};
  const int64_t tot = num_rays * n;
  transient_direct_bwd_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(
      d_direct_rgbs, d_weights, d_ray_dists, d_light_dists, num_rays, p, d_g_transient_direct, d_g_direct_rgbs, d_g_weights);
  const size_t smem = (static_cast<size_t>(n_bins) * channels + static_cast<size_t>(n) * (5 + 2 * channels)) * sizeof(float);
  if (smem > 48u * 1024u) {
    int32_t st = ensure_dynamic_smem<transient_indirect_bwd_kernel>(static_cast<int>(smem));
    if (st != NRC_OK) return st;
  }
  transient_indirect_bwd_kernel<<<static_cast<unsigned>(num_rays), 256, smem, s>>>(
      d_diffuse_raw, d_specular, d_spec_scale, d_weights, d_ray_dists, d_light_dists, d_cam_dists, num_rays, p,
      d_g_transient_indirect, d_g_diffuse_raw, d_g_specular, d_g_spec_scale, d_g_weights);
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
