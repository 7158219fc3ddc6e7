// Second-order path of the analytic normals.
//
// The reference differentiates the loss through normals = -l2_normalize(d raw / d means)
// (internal/geometry.py:442-460, jax.vjp of predict_density w.r.t. the means), which makes the predicted-normal
// loss (internal/loss_utils.py:169-199 with pred='normals') a function of the parameters THROUGH a gradient.
// With g = dL / d(d raw / d means) given per point, that contribution is
//
//     d/d theta  < g , d raw / d x >  =  d/d theta  JVP(raw; x, xdot = g),
//
// i.e. the parameter gradient of the forward-mode tangent of raw along g.  For
// raw = wd . relu(W1^T relu(W0^T e(z(x)) + b0) + b1) with masks M1, M2 (piecewise linear: no mask derivative):
//     zdot  = J_contract xdot                      (J is symmetric: the VJP routine gives the JVP)
//     edot  = scale * sum_corners (grad_z w_c . zdot) T[c]          tangent of the trilinear interpolation
//     h1dot = M1 (W0^T edot),  h2dot = M2 (W1^T h1dot),  rawdot = wd . h2dot
// and, with the ordinary adjoints a2 = M2 wd, a1 = M1 (W1 a2), ge = W0 a1 (= d raw / d e):
//     dW0 += edot (x) a1,  dW1 += h1dot (x) a2,  dwd += h2dot,  dT[c] += (grad_z w_c . zdot) * scale * ge.
// Biases get nothing (the tangent map has no bias term); sample positions are constants
// (stop_level_grad, internal/sampling.py:353-354).  fp32 FFMA, one point per thread: this runs on the 32 final
// samples per ray only (32 768 points per 1024-ray batch), 19 k MAC per point.
#include <cstdint>

#include "encode.cuh"
#include "mlp.cuh"

namespace nrc {

struct Normals2Smem {      // 109 KB: two CTAs per SM
  MlpWeights w;
  float x[kMaxIn * kPad];    // e, then ge, then edot
  float b1[kW * kPad];       // h1, then a1
  float h1d[kW * kPad];      // h1dot
  uint32_t m2[2 * kT];       // M2 of every point of the tile as two bit words (a2 = M2 wd is rebuilt from it)
  float dwd[kW];             // per-CTA partial of d wd
};

template <int F>
__global__ void __launch_bounds__(kT, 2)
density_normals_bwd_kernel(const __grid_constant__ EncDev enc, const nrc_density_mlp_t m,
                           const float* __restrict__ means, const float* __restrict__ g_raw_grad, int64_t P,
                           float warp_c, const nrc_density_mlp_grad_t grads) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Normals2Smem& s = *reinterpret_cast<Normals2Smem*>(smem_raw);
  load_weights(s.w, m);
  __syncthreads();
  const int tid = threadIdx.x;
  const int in_dim = enc.L * F;
  const int rk = tid >> 3, cj = tid & 7;
  float aW1[4][8], aW0[2][8];
  float wdc[8];              // wd[j * 8 + cj]: this thread's columns of a2 in the dW1 update
#pragma unroll
  for (int j = 0; j < 8; ++j) wdc[j] = s.w.wo[4 * (j * 8 + cj)];
  if (tid < kW) s.dwd[tid] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) aW1[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) aW0[i][j] = 0.f;

  const int64_t num_tiles = (P + kT - 1) / kT;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t p = tile * kT + tid;
    const bool valid = p < P;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, xd[3] = {0.f, 0.f, 0.f};
    if (valid) {
      x0 = __ldg(means + 3 * p); x1 = __ldg(means + 3 * p + 1); x2 = __ldg(means + 3 * p + 2);
      xd[0] = __ldg(g_raw_grad + 3 * p); xd[1] = __ldg(g_raw_grad + 3 * p + 1); xd[2] = __ldg(g_raw_grad + 3 * p + 2);
    }
    float z[3], xn[3];
    contract_point(warp_c, x0, x1, x2, z[0], z[1], z[2]);
    normalise_point(enc, z, xn);
    for (int l = 0; l < enc.L; ++l) {
      Corners c = level_setup(enc.lv[l], xn);
      FeatVec<F> v = level_interp<F>(enc.lv[l], c);
#pragma unroll
      for (int f = 0; f < F; ++f) s.x[(l * F + f) * kPad + tid] = __fmul_rn(v.v[f], enc.scale);
    }
    for (int k = in_dim; k < kMaxIn; ++k) s.x[k * kPad + tid] = 0.f;
    float acc[kW];
    mlp_forward_point(s.w, in_dim, s.x + tid, kPad, s.b1 + tid, kPad, acc);
    // a2 = M2 wd (in acc); M2 itself goes to two bit words for the tangent pass and the dW1 update
    uint32_t m2lo = 0u, m2hi = 0u;
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      const bool on = acc[j] > 0.f;
      if (on) { if (j < 32) m2lo |= 1u << j; else m2hi |= 1u << (j - 32); }
      acc[j] = on ? s.w.wo[4 * j] : 0.f;
    }
    s.m2[tid] = m2lo;
    s.m2[kT + tid] = m2hi;
    // a1[k] = M1[k] sum_j W1[k][j] a2[j]   (overwrites the h1 column; M1 kept as the sign of what is stored: the
    // tangent pass needs it too, so keep it in a bit mask)
    uint32_t m1lo = 0u, m1hi = 0u;
    for (int k = 0; k < kW; ++k) {
      const float4* w4 = reinterpret_cast<const float4*>(s.w.w1 + k * kW);
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < kW / 4; ++q) {
        const float4 w = w4[q];
        g = fmaf(w.x, acc[4 * q + 0], g); g = fmaf(w.y, acc[4 * q + 1], g);
        g = fmaf(w.z, acc[4 * q + 2], g); g = fmaf(w.w, acc[4 * q + 3], g);
      }
      const bool on = s.b1[k * kPad + tid] > 0.f;
      if (on) { if (k < 32) m1lo |= 1u << k; else m1hi |= 1u << (k - 32); }
      s.b1[k * kPad + tid] = on ? g : 0.f;
    }
    // ge[i] = sum_k W0[i][k] a1[k]   (overwrites the feature column)
    for (int i = 0; i < in_dim; ++i) {
      const float4* w4 = reinterpret_cast<const float4*>(s.w.w0 + i * kW);
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < kW / 4; ++q) {
        const float4 w = w4[q];
        g = fmaf(w.x, s.b1[(4 * q + 0) * kPad + tid], g); g = fmaf(w.y, s.b1[(4 * q + 1) * kPad + tid], g);
        g = fmaf(w.z, s.b1[(4 * q + 2) * kPad + tid], g); g = fmaf(w.w, s.b1[(4 * q + 3) * kPad + tid], g);
      }
      s.x[i * kPad + tid] = g;
    }
    // tangent of the contraction, then per level: edot and the table-gradient scatter
    float zd[3];
    contract_vjp(warp_c, x0, x1, x2, xd[0], xd[1], xd[2], zd[0], zd[1], zd[2]);
    for (int l = 0; l < enc.L; ++l) {
      const LevelDev& lv = enc.lv[l];
      const Corners c = level_setup(lv, xn);
      const float fN = static_cast<float>(lv.N);
      const float t0 = zd[0] * (fN / enc.span[0]), t1 = zd[1] * (fN / enc.span[1]), t2 = zd[2] * (fN / enc.span[2]);
      float g[F], ed[F];
#pragma unroll
      for (int f = 0; f < F; ++f) { g[f] = s.x[(l * F + f) * kPad + tid] * enc.scale; ed[f] = 0.f; }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int bx, by, bz;
        corner_bits(lv.is_hash, k, bx, by, bz);
        // no early-out: every lane stays for the shuffles of the run-aggregated scatter
        const int32_t row = valid ? corner_row(lv, c, bx, by, bz) : -1;
        const float wx = bx ? c.cw[0] : c.fw[0];
        const float wy = by ? c.cw[1] : c.fw[1];
        const float wz = bz ? c.cw[2] : c.fw[2];
        const float dw = (bx ? t0 : -t0) * (wy * wz) + (by ? t1 : -t1) * (wx * wz) + (bz ? t2 : -t2) * (wx * wy);
        FeatVec<F> v;
        if (row >= 0) v = load_row<F>(lv.table, row);
        else {
#pragma unroll
          for (int f = 0; f < F; ++f) v.v[f] = 0.f;
        }
        float gw[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          ed[f] = fmaf(dw, v.v[f], ed[f]);
          gw[f] = dw * g[f];
        }
        // one vector reduction per corner row; on dense levels runs of lanes in the same cell are summed first
        if (lv.grad) {
          if (!lv.is_hash) warp_run_atomic_add<F>(lv.grad, row, gw, tid & 31);
          else if (row >= 0) atomic_add_row<F>(lv.grad, row, gw);
        }
      }
#pragma unroll
      for (int f = 0; f < F; ++f) s.x[(l * F + f) * kPad + tid] = ed[f] * enc.scale;
    }
    // h1dot = M1 (W0^T edot)
    {
      float t[kW];
#pragma unroll
      for (int j = 0; j < kW; ++j) t[j] = 0.f;
      for (int k = 0; k < in_dim; ++k) axpy64(t, s.x[k * kPad + tid], s.w.w0 + k * kW);
#pragma unroll
      for (int j = 0; j < kW; ++j) {
        const bool on = j < 32 ? ((m1lo >> j) & 1u) : ((m1hi >> (j - 32)) & 1u);
        s.h1d[j * kPad + tid] = on ? t[j] : 0.f;
      }
      // h2dot = M2 (W1^T h1dot)
#pragma unroll
      for (int j = 0; j < kW; ++j) t[j] = 0.f;
#pragma unroll 4
      for (int k = 0; k < kW; ++k) axpy64(t, s.h1d[k * kPad + tid], s.w.w1 + k * kW);
      // d wd += h2dot = M2 t: reduced over the warp's 32 points, then one shared-memory atomic per warp and column
#pragma unroll
      for (int j = 0; j < kW; ++j) {
        const bool on = j < 32 ? ((m2lo >> j) & 1u) : ((m2hi >> (j - 32)) & 1u);
        float v = (valid && on) ? t[j] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) atomicAdd(&s.dwd[j], v);
      }
    }
    __syncthreads();
    // weight gradients of this tile, reduced over its points in registers
    for (int pp = 0; pp < kT; ++pp) {
      float a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = s.h1d[(i * 16 + rk) * kPad + pp];
      const uint32_t lo = s.m2[pp], hi = s.m2[kT + pp];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = j * 8 + cj;
        const bool on = col < 32 ? ((lo >> col) & 1u) : ((hi >> (col - 32)) & 1u);
        b[j] = on ? wdc[j] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) aW1[i][j] = fmaf(a[i], b[j], aW1[i][j]);
      float xa[2], gb[8];
#pragma unroll
      for (int i = 0; i < 2; ++i) xa[i] = s.x[(i * 16 + rk) * kPad + pp];
#pragma unroll
      for (int j = 0; j < 8; ++j) gb[j] = s.b1[(j * 8 + cj) * kPad + pp];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) aW0[i][j] = fmaf(xa[i], gb[j], aW0[i][j]);
    }
    __syncthreads();
  }
  // Adjacent lanes hold adjacent columns (cj = tid & 7): the lanes with cj % 4 == 0 collect their three neighbours'
  // values and issue ONE 16-byte reduction per four columns.  Scalar fallback for unaligned gradient buffers.
  const bool vec = ((reinterpret_cast<uintptr_t>(grads.d_w1) | reinterpret_cast<uintptr_t>(grads.d_w0)) & 15) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = aW1[i][j];
      const float v1 = __shfl_down_sync(0xffffffffu, v, 1), v2 = __shfl_down_sync(0xffffffffu, v, 2),
                  v3 = __shfl_down_sync(0xffffffffu, v, 3);
      float* dst = grads.d_w1 + (i * 16 + rk) * kW + j * 8 + cj;
      if (vec) { if ((cj & 3) == 0) red_add_v4(dst, v, v1, v2, v3); }
      else atomicAdd(dst, v);
    }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = i * 16 + rk;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = aW0[i][j];
      const float v1 = __shfl_down_sync(0xffffffffu, v, 1), v2 = __shfl_down_sync(0xffffffffu, v, 2),
                  v3 = __shfl_down_sync(0xffffffffu, v, 3);
      if (row >= in_dim) continue;
      float* dst = grads.d_w0 + row * kW + j * 8 + cj;
      if (vec) { if ((cj & 3) == 0) red_add_v4(dst, v, v1, v2, v3); }
      else atomicAdd(dst, v);
    }
  }
  if (tid < kW) atomicAdd(grads.d_wd + tid, s.dwd[tid]);
}

template <int F>
int32_t launch_normals2(cudaStream_t s, const EncDev& d, const nrc_density_mlp_t* mlp, const float* means,
                        const float* g, int64_t P, float warp_c, const nrc_density_mlp_grad_t& grads) {
  if (const int32_t st_attr = ensure_dynamic_smem<density_normals_bwd_kernel<F>>(static_cast<int>(sizeof(Normals2Smem))); st_attr != NRC_OK) return st_attr;
  const int64_t tiles = (P + kT - 1) / kT;
  const unsigned grid = static_cast<unsigned>(tiles < 2 * num_sms() ? tiles : 2 * num_sms());
  density_normals_bwd_kernel<F><<<grid, kT, sizeof(Normals2Smem), s>>>(d, *mlp, means, g, P, warp_c, grads);
  return check_launch();
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_density_normals_bwd(void* stream, const nrc_encoding_t* enc, const nrc_density_mlp_t* mlp,
                                           const float* d_means, const float* d_g_raw_grad, int64_t num_points,
                                           float warp_c, const nrc_density_mlp_grad_t* grads) {
  EncDev d;
  int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  st = validate_mlp(mlp);
  if (st != NRC_OK) return st;
  if (mlp->in_dim != d.L * d.F || num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_means || !d_g_raw_grad || !grads || !grads->d_w0 || !grads->d_w1 || !grads->d_wd) return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (d.F) {
    case 1: return launch_normals2<1>(s, d, mlp, d_means, d_g_raw_grad, num_points, warp_c, *grads);
    case 2: return launch_normals2<2>(s, d, mlp, d_means, d_g_raw_grad, num_points, warp_c, *grads);
    case 4: return launch_normals2<4>(s, d, mlp, d_means, d_g_raw_grad, num_points, warp_c, *grads);
    case 8: return NRC_E_UNSUPPORTED;
  }
  return NRC_E_UNSUPPORTED;
}
