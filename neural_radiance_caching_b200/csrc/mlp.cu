// K3 (fp32 parity variant): fused density MLP  enc -> Dense64+ReLU -> Dense64+ReLU ->
// {Dense1 raw density, Dense3 pred-normals}, forward and backward, activations kept
// in shared memory / registers (never in HBM).
//
// Reference: BaseDensityMLP.run_network (internal/geometry.py:155-168),
// pred_normals_layer (:467), widths from configs/ngp_yobo.gin:139-140,206-230.
//
// fp32 FFMA on purpose: TF32 (10-bit mantissa) cannot meet the 1e-5 parity bar; the
// tensor-core (bf16) variant lives in mlp_bf16.cu.
#include "mlp.cuh"

namespace nrc {

__global__ void __launch_bounds__(kT)
density_mlp_fwd_kernel(const nrc_density_mlp_t m, const float* __restrict__ enc, int64_t P,
                       float* __restrict__ raw, float* __restrict__ feat,
                       float* __restrict__ grad_pred) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(smem_raw);
  load_weights(s.w, m);
  __syncthreads();
  const int tid = threadIdx.x;
  const int in_dim = m.in_dim;
  const int64_t num_tiles = (P + kT - 1) / kT;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t p = tile * kT + tid;
    const bool valid = p < P;
    for (int k = 0; k < in_dim; ++k) s.x[k * kT + tid] = valid ? __ldg(enc + p * in_dim + k) : 0.f;
    float acc[kW];
    mlp_forward_point(s.w, in_dim, s.x + tid, kT, s.h1 + tid, kT, acc);
    if (!valid) continue;
    float o[4] = {s.w.bo[0], s.w.bo[1], s.w.bo[2], s.w.bo[3]};
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      float4 w = *reinterpret_cast<const float4*>(s.w.wo + 4 * j);
      o[0] = fmaf(acc[j], w.x, o[0]); o[1] = fmaf(acc[j], w.y, o[1]);
      o[2] = fmaf(acc[j], w.z, o[2]); o[3] = fmaf(acc[j], w.w, o[3]);
    }
    raw[p] = o[0];
    if (grad_pred) { grad_pred[3 * p] = o[1]; grad_pred[3 * p + 1] = o[2]; grad_pred[3 * p + 2] = o[3]; }
    if (feat) {
      float4* f4 = reinterpret_cast<float4*>(feat + p * kW);
#pragma unroll
      for (int q = 0; q < kW / 4; ++q) f4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
  }
}

struct BwdSmem {
  MlpWeights w;
  float x[kMaxIn * kPad];
  float h1[kW * kPad];
  float h2[kW * kPad];
  float g2[kW * kPad];
  float g1[kW * kPad];
  float go[4 * kPad];
};

// Persistent CTAs: each loops over point tiles, keeps its share of every weight gradient
// in registers, and issues one atomic pass at the end.
__global__ void __launch_bounds__(kT)
density_mlp_bwd_kernel(const nrc_density_mlp_t m, const float* __restrict__ enc,
                       const float* __restrict__ g_raw, const float* __restrict__ density,
                       const float* __restrict__ g_feat, const float* __restrict__ g_gp, int64_t P,
                       float* __restrict__ g_enc,
                       const nrc_density_mlp_grad_t grads, int want_wgrad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& s = *reinterpret_cast<BwdSmem*>(smem_raw);
  load_weights(s.w, m);
  __syncthreads();
  const int tid = threadIdx.x;
  const int in_dim = m.in_dim;
  const int rk = tid >> 3;   // 0..15
  const int cj = tid & 7;    // 0..7
  float aW1[4][8], aW0[2][8], aWo[2], aB = 0.f, aBo = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) aW1[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) aW0[i][j] = 0.f;
  aWo[0] = aWo[1] = 0.f;

  const int64_t num_tiles = (P + kT - 1) / kT;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t p = tile * kT + tid;
    const bool valid = p < P;
    for (int k = 0; k < in_dim; ++k) s.x[k * kPad + tid] = valid ? __ldg(enc + p * in_dim + k) : 0.f;
    float acc[kW];
    mlp_forward_point(s.w, in_dim, s.x + tid, kPad, s.h1 + tid, kPad, acc);
    // upstream gradient of the 4 heads
    float go[4];
    go[0] = valid ? (density ? g_raw[p] * density[p] : g_raw[p]) : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) go[1 + c] = (valid && g_gp) ? g_gp[3 * p + c] : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) s.go[c * kPad + tid] = go[c];
    // g_h2 = (Wo go + g_feat) * [h2 > 0]
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      float4 w = *reinterpret_cast<const float4*>(s.w.wo + 4 * j);
      float g = go[0] * w.x + go[1] * w.y + go[2] * w.z + go[3] * w.w;
      if (g_feat && valid) g += g_feat[p * kW + j];
      s.h2[j * kPad + tid] = acc[j];
      s.g2[j * kPad + tid] = acc[j] > 0.f ? g : 0.f;
    }
    // g_h1[k] = sum_j W1[k][j] g_h2[j], masked by relu
    for (int k = 0; k < kW; ++k) {
      const float4* w4 = reinterpret_cast<const float4*>(s.w.w1 + k * kW);
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < kW / 4; ++q) {
        float4 w = w4[q];
        g = fmaf(w.x, s.g2[(4 * q + 0) * kPad + tid], g);
        g = fmaf(w.y, s.g2[(4 * q + 1) * kPad + tid], g);
        g = fmaf(w.z, s.g2[(4 * q + 2) * kPad + tid], g);
        g = fmaf(w.w, s.g2[(4 * q + 3) * kPad + tid], g);
      }
      s.g1[k * kPad + tid] = s.h1[k * kPad + tid] > 0.f ? g : 0.f;
    }
    // g_enc[i] = sum_k W0[i][k] g_h1[k]
    if (g_enc && valid) {
      for (int i = 0; i < in_dim; ++i) {
        const float4* w4 = reinterpret_cast<const float4*>(s.w.w0 + i * kW);
        float g = 0.f;
#pragma unroll
        for (int q = 0; q < kW / 4; ++q) {
          float4 w = w4[q];
          g = fmaf(w.x, s.g1[(4 * q + 0) * kPad + tid], g);
          g = fmaf(w.y, s.g1[(4 * q + 1) * kPad + tid], g);
          g = fmaf(w.z, s.g1[(4 * q + 2) * kPad + tid], g);
          g = fmaf(w.w, s.g1[(4 * q + 3) * kPad + tid], g);
        }
        g_enc[p * in_dim + i] = g;
      }
    }
    if (want_wgrad) {
      __syncthreads();
      // dW1[k][j] += sum_p H1[k][p] G2[j][p];  k = i*16 + rk, j = jj*8 + cj
      for (int pp = 0; pp < kT; ++pp) {
        float a[4], b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = s.h1[(i * 16 + rk) * kPad + pp];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = s.g2[(j * 8 + cj) * kPad + pp];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) aW1[i][j] = fmaf(a[i], b[j], aW1[i][j]);
        // dW0[i][k] += sum_p X[i][p] G1[k][p];  i = ii*16 + rk, k = jj*8 + cj
        float xa[2], gb[8];
#pragma unroll
        for (int i = 0; i < 2; ++i) xa[i] = s.x[(i * 16 + rk) * kPad + pp];
#pragma unroll
        for (int j = 0; j < 8; ++j) gb[j] = s.g1[(j * 8 + cj) * kPad + pp];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) aW0[i][j] = fmaf(xa[i], gb[j], aW0[i][j]);
        // dWo[j][c]: j = tid % 64, c in {2*(tid/64), 2*(tid/64)+1}
        float h2v = s.h2[(tid & 63) * kPad + pp];
        aWo[0] = fmaf(h2v, s.go[(2 * (tid >> 6)) * kPad + pp], aWo[0]);
        aWo[1] = fmaf(h2v, s.go[(2 * (tid >> 6) + 1) * kPad + pp], aWo[1]);
        // biases: tid < 64 -> db1[tid]; tid >= 64 -> db0[tid-64]
        aB += (tid < 64) ? s.g2[tid * kPad + pp] : s.g1[(tid - 64) * kPad + pp];
        if (tid < 4) aBo += s.go[tid * kPad + pp];
      }
      __syncthreads();
    }
  }
  if (!want_wgrad) return;
  // note: rows of x beyond in_dim are zero-filled only up to what was loaded; mask on in_dim.
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(grads.d_w1 + (i * 16 + rk) * kW + j * 8 + cj, aW1[i][j]);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int row = i * 16 + rk;
    if (row < in_dim) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(grads.d_w0 + row * kW + j * 8 + cj, aW0[i][j]);
    }
  }
  {
    int j = tid & 63, c0 = 2 * (tid >> 6);
    if (c0 == 0) {
      atomicAdd(grads.d_wd + j, aWo[0]);
      if (grads.d_wn) atomicAdd(grads.d_wn + j * 3 + 0, aWo[1]);
    } else if (grads.d_wn) {
      atomicAdd(grads.d_wn + j * 3 + 1, aWo[0]);
      atomicAdd(grads.d_wn + j * 3 + 2, aWo[1]);
    }
  }
  if (tid < 64) atomicAdd(grads.d_b1 + tid, aB); else atomicAdd(grads.d_b0 + (tid - 64), aB);
  if (tid == 0) atomicAdd(grads.d_bd, aBo);
  else if (tid < 4 && grads.d_bn) atomicAdd(grads.d_bn + (tid - 1), aBo);
}

int32_t density_mlp_fwd_f32(cudaStream_t s, const nrc_density_mlp_t* mlp, const float* d_enc, int64_t P,
                            float* d_raw, float* d_feat, float* d_gp) {
  if (const int32_t st_attr = ensure_dynamic_smem<density_mlp_fwd_kernel>(static_cast<int>(sizeof(FwdSmem))); st_attr != NRC_OK) return st_attr;
  int64_t tiles = (P + kT - 1) / kT;
  unsigned grid = static_cast<unsigned>(tiles < num_sms() * 3 ? tiles : num_sms() * 3);
  density_mlp_fwd_kernel<<<grid, kT, sizeof(FwdSmem), s>>>(*mlp, d_enc, P, d_raw, d_feat, d_gp);
  return check_launch();
}

}  // namespace nrc

using namespace nrc;

namespace nrc {
int32_t density_mlp_fwd_bf16(cudaStream_t s, const nrc_density_mlp_t* mlp, const float* d_enc, int64_t P,
                             float* d_raw, float* d_feat, float* d_gp);
int32_t density_mlp_bwd_bf16(cudaStream_t s, const nrc_density_mlp_t* mlp, const float* d_enc,
                             const float* d_g_raw, const float* d_density, const float* d_g_feat,
                             const float* d_g_gp, int64_t P, float* d_g_enc, const nrc_density_mlp_grad_t* grads);
}

extern "C" int32_t nrc_density_mlp_fwd(void* stream, const nrc_density_mlp_t* mlp, const float* d_enc,
                                       int64_t num_points, int32_t bf16, float* d_raw, float* d_feat,
                                       float* d_grad_pred) {
  int32_t st = validate_mlp(mlp);
  if (st != NRC_OK) return st;
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_enc || !d_raw) return NRC_E_INVALID_ARG;
  if (d_grad_pred && !mlp->d_wn) return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bf16) return density_mlp_fwd_bf16(s, mlp, d_enc, num_points, d_raw, d_feat, d_grad_pred);
  return density_mlp_fwd_f32(s, mlp, d_enc, num_points, d_raw, d_feat, d_grad_pred);
}

extern "C" int32_t nrc_density_mlp_bwd(void* stream, const nrc_density_mlp_t* mlp, const float* d_enc,
                                       const float* d_g_raw, const float* d_density, const float* d_g_feat,
                                       const float* d_g_grad_pred, int64_t num_points, int32_t bf16,
                                       float* d_g_enc, const nrc_density_mlp_grad_t* grads) {
  int32_t st = validate_mlp(mlp);
  if (st != NRC_OK) return st;
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_enc || !d_g_raw) return NRC_E_INVALID_ARG;
  if (d_g_grad_pred && !mlp->d_wn) return NRC_E_INVALID_ARG;
  nrc_density_mlp_grad_t g{};
  int want = 0;
  if (grads) {
    g = *grads;
    if (!g.d_w0 || !g.d_b0 || !g.d_w1 || !g.d_b1 || !g.d_wd || !g.d_bd) return NRC_E_INVALID_ARG;
    if (mlp->d_wn && (!g.d_wn || !g.d_bn)) return NRC_E_INVALID_ARG;
    if (!mlp->d_wn) { g.d_wn = nullptr; g.d_bn = nullptr; }
    want = 1;
  }
  if (!want && !d_g_enc) return NRC_OK;
  if (bf16)
    return density_mlp_bwd_bf16(static_cast<cudaStream_t>(stream), mlp, d_enc, d_g_raw, d_density, d_g_feat,
                                d_g_grad_pred, num_points, d_g_enc, want ? &g : nullptr);
  if (const int32_t st_attr = ensure_dynamic_smem<density_mlp_bwd_kernel>(static_cast<int>(sizeof(BwdSmem))); st_attr != NRC_OK) return st_attr;
  int64_t tiles = (num_points + kT - 1) / kT;
  unsigned grid = static_cast<unsigned>(tiles < num_sms() ? tiles : num_sms());
  density_mlp_bwd_kernel<<<grid, kT, sizeof(BwdSmem), static_cast<cudaStream_t>(stream)>>>(
      *mlp, d_enc, d_g_raw, d_density, d_g_feat, d_g_grad_pred, num_points, d_g_enc, g, want);
  return check_launch();
}
