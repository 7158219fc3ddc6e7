// Fused radiance-cache point query:
//   means -> contract(x/c) -> multires hash/dense encode -> Dense64 ReLU Dense64 ReLU ->
//   raw density (+ pred-normals head) -> density = safe_exp(raw + bias) masked to the bbox,
// and optionally the analytic-normal gradient d raw / d means (geometry.py:442-460) by
// back-propagating through the MLP, the encoding and the contraction inside the same kernel.
// Features never round-trip HBM.  Reference: internal/geometry.py:199-341,381-518.
#include "encode.cuh"
#include "mlp.cuh"

namespace nrc {

template <int F>
__global__ void __launch_bounds__(kT)
density_query_fwd_kernel(const __grid_constant__ EncDev enc, const nrc_density_mlp_t m,
                         const float* __restrict__ means, int64_t P, float warp_c, float density_bias,
                         float* __restrict__ density, float* __restrict__ raw, float* __restrict__ feat,
                         float* __restrict__ grad_pred, float* __restrict__ raw_grad,
                         float* __restrict__ enc_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(smem_raw);
  load_weights(s.w, m);
  __syncthreads();
  const int tid = threadIdx.x;
  const int in_dim = enc.L * F;
  const int64_t num_tiles = (P + kT - 1) / kT;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t p = tile * kT + tid;
    if (p >= P) continue;
    const float x0 = __ldg(means + 3 * p), x1 = __ldg(means + 3 * p + 1), x2 = __ldg(means + 3 * p + 2);
    float z[3];
    contract_point(warp_c, x0, x1, x2, z[0], z[1], z[2]);
    float xn[3];
    normalise_point(enc, z, xn);
    for (int l = 0; l < enc.L; ++l) {
      Corners c = level_setup(enc.lv[l], xn);
      FeatVec<F> v = level_interp<F>(enc.lv[l], c);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float e = __fmul_rn(v.v[f], enc.scale);
        s.x[(l * F + f) * kT + tid] = e;
        if (enc_out) enc_out[p * in_dim + l * F + f] = e;
      }
    }
    float acc[kW];
    mlp_forward_point(s.w, in_dim, s.x + tid, kT, s.h1 + tid, kT, acc);
    float o[4] = {s.w.bo[0], s.w.bo[1], s.w.bo[2], s.w.bo[3]};
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      float4 w = *reinterpret_cast<const float4*>(s.w.wo + 4 * j);
      o[0] = fmaf(acc[j], w.x, o[0]); o[1] = fmaf(acc[j], w.y, o[1]);
      o[2] = fmaf(acc[j], w.z, o[2]); o[3] = fmaf(acc[j], w.w, o[3]);
    }
    if (raw) raw[p] = o[0];
    if (density) {
      // convert_raw_density (geometry.py:318-341): bbox test on the warped mean, strict.
      bool inside = true;
#pragma unroll
      for (int a = 0; a < 3; ++a)
        inside = inside && (z[a] > enc.b0[a]) && (z[a] < enc.b1[a]);
      density[p] = inside ? safe_exp(o[0] + density_bias) : 0.f;
    }
    if (grad_pred) { grad_pred[3 * p] = o[1]; grad_pred[3 * p + 1] = o[2]; grad_pred[3 * p + 2] = o[3]; }
    if (feat) {
      float4* f4 = reinterpret_cast<float4*>(feat + p * kW);
#pragma unroll
      for (int q = 0; q < kW / 4; ++q) f4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
    if (!raw_grad) continue;
    // ---- analytic normals: d raw / d means ---------------------------------------
    // g_h2[j] = wd[j] * [h2[j] > 0]
#pragma unroll
    for (int j = 0; j < kW; ++j) acc[j] = acc[j] > 0.f ? s.w.wo[4 * j] : 0.f;
    // g_h1[k] = [h1[k] > 0] * sum_j W1[k][j] g_h2[j]   (overwrites the h1 column)
    for (int k = 0; k < kW; ++k) {
      const float4* w4 = reinterpret_cast<const float4*>(s.w.w1 + k * kW);
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < kW / 4; ++q) {
        float4 w = w4[q];
        g = fmaf(w.x, acc[4 * q + 0], g); g = fmaf(w.y, acc[4 * q + 1], g);
        g = fmaf(w.z, acc[4 * q + 2], g); g = fmaf(w.w, acc[4 * q + 3], g);
      }
      float h = s.h1[k * kT + tid];
      s.h1[k * kT + tid] = h > 0.f ? g : 0.f;
    }
    // g_enc[i] = sum_k W0[i][k] g_h1[k]   (overwrites the feature column)
    for (int i = 0; i < in_dim; ++i) {
      const float4* w4 = reinterpret_cast<const float4*>(s.w.w0 + i * kW);
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < kW / 4; ++q) {
        float4 w = w4[q];
        g = fmaf(w.x, s.h1[(4 * q + 0) * kT + tid], g); g = fmaf(w.y, s.h1[(4 * q + 1) * kT + tid], g);
        g = fmaf(w.z, s.h1[(4 * q + 2) * kT + tid], g); g = fmaf(w.w, s.h1[(4 * q + 3) * kT + tid], g);
      }
      s.x[i * kT + tid] = g;
    }
    // VJP through the encoding (re-gathers the corners; L1/L2 hits) and the contraction.
    float gz[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < enc.L; ++l) {
      const LevelDev& lv = enc.lv[l];
      Corners c = level_setup(lv, xn);
      float g[F];
#pragma unroll
      for (int f = 0; f < F; ++f) g[f] = s.x[(l * F + f) * kT + tid] * enc.scale;
      float gl[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int bx, by, bz;
        corner_bits(lv.is_hash, k, bx, by, bz);
        int32_t row = corner_row(lv, c, bx, by, bz);
        if (row < 0) continue;
        FeatVec<F> v = load_row<F>(lv.table, row);
        float dot = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) dot = fmaf(g[f], v.v[f], dot);
        float wx = bx ? c.cw[0] : c.fw[0];
        float wy = by ? c.cw[1] : c.fw[1];
        float wz = bz ? c.cw[2] : c.fw[2];
        gl[0] += (bx ? dot : -dot) * (wy * wz);
        gl[1] += (by ? dot : -dot) * (wx * wz);
        gl[2] += (bz ? dot : -dot) * (wx * wy);
      }
      const float fN = static_cast<float>(lv.N);
#pragma unroll
      for (int a = 0; a < 3; ++a) gz[a] += gl[a] * (fN / enc.span[a]);
    }
    float o0, o1, o2;
    contract_vjp(warp_c, x0, x1, x2, gz[0], gz[1], gz[2], o0, o1, o2);
    raw_grad[3 * p] = o0; raw_grad[3 * p + 1] = o1; raw_grad[3 * p + 2] = o2;
  }
}

template <int F>
int32_t launch_query(cudaStream_t s, const EncDev& d, const nrc_density_mlp_t* mlp, const float* means,
                     int64_t P, float warp_c, float bias, float* density, float* raw, float* feat,
                     float* gp, float* rg, float* eo) {
  if (const int32_t st_attr = ensure_dynamic_smem<density_query_fwd_kernel<F>>(static_cast<int>(sizeof(FwdSmem))); st_attr != NRC_OK) return st_attr;
  int64_t tiles = (P + kT - 1) / kT;
  unsigned grid = static_cast<unsigned>(tiles < num_sms() * 3 ? tiles : num_sms() * 3);
  density_query_fwd_kernel<F><<<grid, kT, sizeof(FwdSmem), s>>>(d, *mlp, means, P, warp_c, bias, density,
                                                               raw, feat, gp, rg, eo);
  return check_launch();
}

int32_t density_query_fwd_bf16(cudaStream_t s, const EncDev& d, const nrc_density_mlp_t* mlp,
                               const float* d_means, int64_t P, float warp_c, float bias, float* density,
                               float* raw, float* feat, float* gp, float* rg, float* enc_out);

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_density_query_fwd(void* stream, const nrc_encoding_t* enc,
                                         const nrc_density_mlp_t* mlp, const float* d_means,
                                         int64_t num_points, float warp_c, float density_bias,
                                         int32_t bf16, float* d_density, float* d_raw, float* d_feat,
                                         float* d_grad_pred, float* d_raw_grad, float* d_enc_out) {
  EncDev d;
  int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  st = validate_mlp(mlp);
  if (st != NRC_OK) return st;
  if (mlp->in_dim != d.L * d.F) return NRC_E_INVALID_ARG;
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_means) return NRC_E_INVALID_ARG;
  if (d_grad_pred && !mlp->d_wn) return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bf16)
    return density_query_fwd_bf16(s, d, mlp, d_means, num_points, warp_c, density_bias, d_density, d_raw,
                                  d_feat, d_grad_pred, d_raw_grad, d_enc_out);
  switch (d.F) {
    case 1: return launch_query<1>(s, d, mlp, d_means, num_points, warp_c, density_bias, d_density, d_raw, d_feat, d_grad_pred, d_raw_grad, d_enc_out);
    case 2: return launch_query<2>(s, d, mlp, d_means, num_points, warp_c, density_bias, d_density, d_raw, d_feat, d_grad_pred, d_raw_grad, d_enc_out);
    case 4: return launch_query<4>(s, d, mlp, d_means, num_points, warp_c, density_bias, d_density, d_raw, d_feat, d_grad_pred, d_raw_grad, d_enc_out);
    case 8: return launch_query<8>(s, d, mlp, d_means, num_points, warp_c, density_bias, d_density, d_raw, d_feat, d_grad_pred, d_raw_grad, d_enc_out);
  }
  return NRC_E_UNSUPPORTED;
}
