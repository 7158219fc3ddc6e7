// Shared device helpers of the fused density MLP (fp32 variant): shared-memory weight
// image and the one-point-per-thread forward evaluation.  See mlp.cu.
#pragma once
#include "nrc_common.cuh"

namespace nrc {

constexpr int kW = 64;        // hidden width
constexpr int kMaxIn = 32;    // max in_dim (L*F)
constexpr int kT = 128;       // points per tile == threads per CTA
constexpr int kPad = kT + 1;  // padded row stride for the transposed tiles (bank = row + p)

struct MlpWeights {           // shared-memory image of the parameters
  float w0[kMaxIn * kW];
  float b0[kW];
  float w1[kW * kW];
  float b1[kW];
  float wo[kW * 4];           // [j][0] = output_density_layer, [j][1..3] = pred_normals_layer
  float bo[4];
};

__device__ __forceinline__ void load_weights(MlpWeights& s, const nrc_density_mlp_t& m) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < kMaxIn * kW; i += nt) s.w0[i] = (i < m.in_dim * kW) ? m.d_w0[i] : 0.f;
  for (int i = tid; i < kW * kW; i += nt) s.w1[i] = m.d_w1[i];
  for (int i = tid; i < kW; i += nt) {
    s.b0[i] = m.d_b0[i];
    s.b1[i] = m.d_b1[i];
    s.wo[i * 4 + 0] = m.d_wd[i];
    for (int c = 0; c < 3; ++c) s.wo[i * 4 + 1 + c] = m.d_wn ? m.d_wn[i * 3 + c] : 0.f;
  }
  if (tid < 4) s.bo[tid] = tid == 0 ? m.d_bd[0] : (m.d_bn ? m.d_bn[tid - 1] : 0.f);
}

// acc[j] += x * w[j], j = 0..63, weights broadcast from shared memory as float4.
__device__ __forceinline__ void axpy64(float (&acc)[kW], float x, const float* __restrict__ w) {
  const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
  for (int q = 0; q < kW / 4; ++q) {
    float4 v = w4[q];
    acc[4 * q + 0] = fmaf(x, v.x, acc[4 * q + 0]);
    acc[4 * q + 1] = fmaf(x, v.y, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(x, v.z, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(x, v.w, acc[4 * q + 3]);
  }
}

// One point per thread.  hcol: this thread's column of a [kW][stride] shared tile that
// receives relu(layer0); on return acc holds relu(layer1).
__device__ __forceinline__ void mlp_forward_point(const MlpWeights& s, int in_dim, const float* xcol,
                                                  int xstride, float* hcol, int hstride,
                                                  float (&acc)[kW]) {
#pragma unroll
  for (int j = 0; j < kW; ++j) acc[j] = s.b0[j];
  for (int k = 0; k < in_dim; ++k) axpy64(acc, xcol[k * xstride], s.w0 + k * kW);
#pragma unroll
  for (int j = 0; j < kW; ++j) hcol[j * hstride] = fmaxf(acc[j], 0.f);
#pragma unroll
  for (int j = 0; j < kW; ++j) acc[j] = s.b1[j];
#pragma unroll 4
  for (int k = 0; k < kW; ++k) axpy64(acc, hcol[k * hstride], s.w1 + k * kW);
#pragma unroll
  for (int j = 0; j < kW; ++j) acc[j] = fmaxf(acc[j], 0.f);
}

struct FwdSmem {
  MlpWeights w;
  float x[kMaxIn * kT];
  float h1[kW * kT];
};

inline int32_t validate_mlp(const nrc_density_mlp_t* m) {
  if (!m) return NRC_E_INVALID_ARG;
  if (m->width != kW) return NRC_E_UNSUPPORTED;
  if (m->in_dim < 1 || m->in_dim > kMaxIn) return NRC_E_UNSUPPORTED;
  if (!m->d_w0 || !m->d_b0 || !m->d_w1 || !m->d_b1 || !m->d_wd || !m->d_bd) return NRC_E_INVALID_ARG;
  if ((m->d_wn == nullptr) != (m->d_bn == nullptr)) return NRC_E_INVALID_ARG;
  return NRC_OK;
}

}  // namespace nrc
