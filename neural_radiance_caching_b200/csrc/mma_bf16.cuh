// Warp-level bf16 tensor-core building blocks (mma.sync m16n8k16, fp32 accumulate) for the
// fused density MLP.  The layers are 6/7/32 -> 64 -> 64 -> 4 wide: far below a tcgen05 tile
// (M=128, TMEM accumulators, TMA-staged operands), and their A operand is produced in
// registers by the hash-grid gather of the same warp, so the layers are chained in registers
// with warp-synchronous MMAs: the accumulator fragment of layer l *is* the A fragment of
// layer l+1 (no shared-memory or TMEM round trip between layers).
#pragma once
#include <cuda_bf16.h>

#include "nrc_common.cuh"

namespace nrc {

constexpr int kHid = 64;          // hidden width
constexpr int kWStride = 72;      // bf16 row stride of 64-wide tiles (144 B: conflict-free ldmatrix)
constexpr int kXStride = 40;      // bf16 row stride of the <=32-wide input tile (80 B)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}

// D += A(16x16, row) * B(16x8, col), bf16 in, fp32 accumulate.
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// A fragment (16 rows x 16 k) from a row-major bf16 tile [row][k].
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], const __nv_bfloat16* tile, int stride,
                                            int row0, int k0, int lane) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = k0 + (lane >> 4) * 8;
  ldmatrix_x4(a, tile + r * stride + c);
}
// B fragments of two adjacent n-tiles (16 n x 16 k) from a tile stored [n][k] (k contiguous):
// b[0],b[1] -> n-tile n0/8, b[2],b[3] -> n-tile n0/8 + 1.
__device__ __forceinline__ void load_b_frag2(uint32_t (&b)[4], const __nv_bfloat16* tile, int stride,
                                             int n0, int k0, int lane) {
  const int n = n0 + (lane & 7) + (lane >> 4) * 8;
  const int c = k0 + ((lane >> 3) & 1) * 8;
  ldmatrix_x4(b, tile + n * stride + c);
}
// B fragments of ONE n-tile (8 n) for two adjacent k-steps (32 k): b[0],b[1] -> k0, b[2],b[3] -> k0+16.
__device__ __forceinline__ void load_b_frag_k32(uint32_t (&b)[4], const __nv_bfloat16* tile, int stride,
                                                int n0, int k0, int lane) {
  const int n = n0 + (lane & 7);
  const int c = k0 + (lane >> 3) * 8;
  ldmatrix_x4(b, tile + n * stride + c);
}
// A fragment of M^T where the tile is stored [k][m] (i.e. rows are the reduction index):
// used for weight gradients dW = A^T G with the activations stored point-major.
__device__ __forceinline__ void load_a_frag_trans(uint32_t (&a)[4], const __nv_bfloat16* tile, int stride,
                                                  int k0, int m0, int lane) {
  const int r = k0 + (lane & 7) + (lane >> 4) * 8;
  const int c = m0 + ((lane >> 3) & 1) * 8;
  ldmatrix_x4_trans(a, tile + r * stride + c);
}
// B fragments of two adjacent n-tiles from a tile stored [k][n] (n contiguous).
__device__ __forceinline__ void load_b_frag2_trans(uint32_t (&b)[4], const __nv_bfloat16* tile, int stride,
                                                   int k0, int n0, int lane) {
  const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = n0 + (lane >> 4) * 8;
  ldmatrix_x4_trans(b, tile + r * stride + c);
}
// B fragment of one n-tile from a tile stored [k][n], one k-step.
__device__ __forceinline__ void load_b_frag1_trans(uint32_t (&b)[2], const __nv_bfloat16* tile, int stride,
                                                   int k0, int n0, int lane) {
  const int r = k0 + (lane & 15);
  ldmatrix_x2_trans(b, tile + r * stride + n0);
}

// Shared-memory image of the MLP parameters in bf16, both orientations.  The forward-only part and the
// part needed for data gradients are separate structs so that a forward-only launch can leave the second
// one (and its per-warp scratch) out of its dynamic shared memory.
struct __align__(16) MlpWeightsFwdBf16 {
  __nv_bfloat16 w0t[kHid][kXStride];   // [out j][in i]   forward layer 0   (B of X W0)
  __nv_bfloat16 w1t[kHid][kWStride];   // [out j][in k]   forward layer 1
  __nv_bfloat16 wot[8][kWStride];      // [head c][in j]  heads: c=0 density, 1..3 pred normals
  float b0[kHid];
  float b1[kHid];
  float bo[4];
};
// The data-gradient GEMMs (g W1^T, g W0^T) read the SAME tiles through transposing ldmatrix loads
// (mma_layer64_t / load_b_frag2_trans): no second copy of the weights.
struct __align__(16) MlpWeightsGradBf16 {
  float wo[kHid][4];                   // fp32 heads for the elementwise g_h2 term
};
struct __align__(16) MlpWeightsBf16 : MlpWeightsFwdBf16, MlpWeightsGradBf16 {};

// Global reads run in the parameters' own (Flax [in][out]) order, i.e. coalesced; the transposed copies are
// scattered into shared memory instead.  This prologue runs once per CTA, and at 1024 rays a CTA has ONE tile: the
// scalar version below was 20-30 % of the density kernels' warp samples (profiles/r02i, long-scoreboard stalls on
// 50 dependent 4-byte loads per thread).  The vector version issues every 16-byte load of a thread (two rows of the
// matrix x four columns per item) before the first conversion, so the whole image costs about one L2 round trip.
__device__ __forceinline__ void load_weights_bf16_scalar(MlpWeightsFwdBf16& s, MlpWeightsGradBf16* g,
                                                         const nrc_density_mlp_t& m) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int in_dim = m.in_dim;
  for (int idx = tid; idx < kXStride * kHid; idx += nt) {          // idx = i * 64 + j
    const int i = idx / kHid, j = idx % kHid;
    s.w0t[j][i] = __float2bfloat16((i < in_dim) ? m.d_w0[idx] : 0.f);
  }
  for (int idx = tid; idx < kHid * kHid; idx += nt) {              // idx = k * 64 + j
    const int k = idx / kHid, j = idx % kHid;
    s.w1t[j][k] = __float2bfloat16(m.d_w1[idx]);
  }
  for (int idx = tid; idx < kHid * (kWStride - kHid); idx += nt) { // zero the padding columns
    const int r = idx / (kWStride - kHid), c = kHid + idx % (kWStride - kHid);
    s.w1t[r][c] = __float2bfloat16(0.f);
  }
  for (int idx = tid; idx < 8 * kWStride; idx += nt) {
    int c = idx / kWStride, j = idx % kWStride;
    float v = 0.f;
    if (j < kHid) {
      if (c == 0) v = m.d_wd[j];
      else if (c < 4 && m.d_wn) v = m.d_wn[j * 3 + (c - 1)];
    }
    s.wot[c][j] = __float2bfloat16(v);
  }
  for (int j = tid; j < kHid; j += nt) {
    s.b0[j] = m.d_b0[j];
    s.b1[j] = m.d_b1[j];
    if (g) {
      g->wo[j][0] = m.d_wd[j];
      for (int c = 0; c < 3; ++c) g->wo[j][1 + c] = m.d_wn ? m.d_wn[j * 3 + c] : 0.f;
    }
  }
  if (tid < 4) s.bo[tid] = tid == 0 ? m.d_bd[0] : (m.d_bn ? m.d_bn[tid - 1] : 0.f);
}

// kThreads = blockDim.x (128 in every density kernel).
template <int kThreads>
__device__ __forceinline__ void load_weights_bf16(MlpWeightsFwdBf16& s, MlpWeightsGradBf16* g,
                                                  const nrc_density_mlp_t& m) {
  if (((reinterpret_cast<uintptr_t>(m.d_w0) | reinterpret_cast<uintptr_t>(m.d_w1)) & 15) != 0) {
    load_weights_bf16_scalar(s, g, m);
    return;
  }
  const int tid = threadIdx.x;
  const int in_dim = m.in_dim;
  constexpr int kQ = kHid / 4;                                   // column quads per matrix row
  constexpr int kItems1 = (kHid / 2) * kQ / kThreads;            // (row pair, column quad) items of W1 per thread
  constexpr int kItems0 = ((kXStride / 2) * kQ + kThreads - 1) / kThreads;
  static_assert((kHid / 2) * kQ % kThreads == 0, "W1 items must divide evenly");
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 r1[kItems1][2], r0[kItems0][2];
#pragma unroll
  for (int n = 0; n < kItems1; ++n) {
    const int it = tid + n * kThreads, kp = it / kQ, jq = it % kQ;
    r1[n][0] = __ldg(reinterpret_cast<const float4*>(m.d_w1 + (2 * kp) * kHid) + jq);
    r1[n][1] = __ldg(reinterpret_cast<const float4*>(m.d_w1 + (2 * kp + 1) * kHid) + jq);
  }
#pragma unroll
  for (int n = 0; n < kItems0; ++n) {
    const int it = tid + n * kThreads, ip = it / kQ, jq = it % kQ;
    r0[n][0] = (2 * ip < in_dim) ? __ldg(reinterpret_cast<const float4*>(m.d_w0 + (2 * ip) * kHid) + jq) : zero4;
    r0[n][1] = (2 * ip + 1 < in_dim) ? __ldg(reinterpret_cast<const float4*>(m.d_w0 + (2 * ip + 1) * kHid) + jq) : zero4;
  }
  float hd[4] = {0.f, 0.f, 0.f, 0.f}, bb0 = 0.f, bb1 = 0.f, bbo = 0.f;
  if (tid < kHid) {
    hd[0] = __ldg(m.d_wd + tid);
    if (m.d_wn) { hd[1] = __ldg(m.d_wn + 3 * tid); hd[2] = __ldg(m.d_wn + 3 * tid + 1); hd[3] = __ldg(m.d_wn + 3 * tid + 2); }
    bb0 = __ldg(m.d_b0 + tid);
    bb1 = __ldg(m.d_b1 + tid);
  }
  if (tid < 4) bbo = tid == 0 ? __ldg(m.d_bd) : (m.d_bn ? __ldg(m.d_bn + tid - 1) : 0.f);
#pragma unroll
  for (int n = 0; n < kItems1; ++n) {
    const int it = tid + n * kThreads, kp = it / kQ, jq = it % kQ;
    const float a[4] = {r1[n][0].x, r1[n][0].y, r1[n][0].z, r1[n][0].w};
    const float b[4] = {r1[n][1].x, r1[n][1].y, r1[n][1].z, r1[n][1].w};
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint32_t*>(&s.w1t[4 * jq + c][2 * kp]) = pack_bf16(a[c], b[c]);
  }
#pragma unroll
  for (int n = 0; n < kItems0; ++n) {
    const int it = tid + n * kThreads, ip = it / kQ, jq = it % kQ;
    if (ip >= kXStride / 2) continue;
    const float a[4] = {r0[n][0].x, r0[n][0].y, r0[n][0].z, r0[n][0].w};
    const float b[4] = {r0[n][1].x, r0[n][1].y, r0[n][1].z, r0[n][1].w};
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint32_t*>(&s.w0t[4 * jq + c][2 * ip]) = pack_bf16(a[c], b[c]);
  }
  if (tid < kHid) {
    *reinterpret_cast<uint4*>(&s.w1t[tid][kHid]) = make_uint4(0u, 0u, 0u, 0u);      // padding columns 64..71
#pragma unroll
    for (int c = 0; c < 8; ++c) s.wot[c][tid] = __float2bfloat16(c < 4 ? hd[c] : 0.f);
    s.b0[tid] = bb0;
    s.b1[tid] = bb1;
    if (g) *reinterpret_cast<float4*>(&g->wo[tid][0]) = make_float4(hd[0], hd[1], hd[2], hd[3]);
  } else if (tid < kWStride) {
#pragma unroll
    for (int c = 0; c < 8; ++c) s.wot[c][tid] = __float2bfloat16(0.f);
  }
  if (tid < 4) s.bo[tid] = bbo;
}
template <int kThreads>
__device__ __forceinline__ void load_weights_bf16(MlpWeightsBf16& s, const nrc_density_mlp_t& m) {
  load_weights_bf16<kThreads>(s, &s, m);
}

// acc[nt][e] (16 rows x 64 cols, fp32) = bias + A(16 x 16*KS) * W^T, W^T stored [64][stride].
template <int KS>
__device__ __forceinline__ void mma_layer64(float (&acc)[8][4], const uint32_t (&a)[KS][4],
                                            const __nv_bfloat16* wt, int stride, const float* bias,
                                            int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float2 bv = bias ? *reinterpret_cast<const float2*>(bias + nt * 8 + (lane & 3) * 2) : make_float2(0.f, 0.f);
    acc[nt][0] = bv.x; acc[nt][1] = bv.y; acc[nt][2] = bv.x; acc[nt][3] = bv.y;
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_b_frag2(b, wt, stride, np * 16, ks * 16, lane);
      mma_bf16(acc[2 * np], a[ks], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}

// acc (16 x 64) = A(16 x 64) * W with W^T stored [j][k] (the forward tile): the reduction runs over the
// tile's ROW index j, i.e. the data-gradient GEMM g W^T of a layer whose forward GEMM is mma_layer64.
template <int KS>
__device__ __forceinline__ void mma_layer64_t(float (&acc)[8][4], const uint32_t (&a)[KS][4],
                                              const __nv_bfloat16* wt, int stride, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_b_frag2_trans(b, wt, stride, ks * 16, np * 16, lane);
      mma_bf16(acc[2 * np], a[ks], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}

// Accumulator fragment (16 x 64 fp32) -> A fragments of the next layer (4 k-steps), with an
// optional ReLU applied first.
template <bool kRelu>
__device__ __forceinline__ void acc_to_afrag(const float (&acc)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float* c = acc[2 * kk + h];
      float v0 = c[0], v1 = c[1], v2 = c[2], v3 = c[3];
      if (kRelu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
      a[kk][2 * h + 0] = pack_bf16(v0, v1);
      a[kk][2 * h + 1] = pack_bf16(v2, v3);
    }
  }
}

}  // namespace nrc
