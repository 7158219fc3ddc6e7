// Dense layers of the cache shader: flax.linen.Dense y = act(x @ kernel + bias)
// and its VJPs, as ONE tiled GEMM template
//     C[M,N] (=|+=) epilogue( opA(A)[M,K] * opB(B)[K,N] )
// instantiated for fp32 FFMA (parity variant, 1e-5) and bf16 mma.sync (tensor-core variant, 2e-2):
//   forward          : A = x [M,K]           B = kernel [K,N]         bias + ReLU epilogue
//   backward (data)  : A = g_y [M,N'] masked  B = kernel^T (transB)    optional accumulate (skips)
//   backward (weight): A = x^T (transA)      B = g_y masked           split over M, atomic epilogue
// Row strides (lda/ldb/ldc) let layers read/write column slices of wider activation buffers, so the
// reference's jnp.concatenate skip connections (surface_light_field.py:480-500) never copy.
// Reference: geometry.py:127-139 (Dense), nerf.py:232-345, surface_light_field.py:352-403.
#include "mma_bf16.cuh"

namespace nrc {

constexpr int BM = 128, BN = 64, BK = 32;
constexpr int kGemmThreads = 128;

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  const float* bias;          // [N] or null (forward)
  const float* maskA; int64_t ldma;  // relu mask source for A (same indexing as A) or null
  const float* maskB; int64_t ldmb;  // relu mask source for B (same indexing as B) or null
  float* colsum;              // [N]: += column sums of opB(B) rows (bias gradient) or null
  int M, N, K;
  int relu;                   // epilogue activation
  int accumulate;             // C += result instead of C = result
  int k_splits;               // >1: split the K loop over blockIdx.z, atomicAdd epilogue
};

// element (r, c) of op(X): X stored row-major with stride ld; trans => stored [c][r].
template <bool TRANS>
__device__ __forceinline__ float load_elem(const float* __restrict__ X, int64_t ld, const float* __restrict__ mask,
                                           int64_t ldm, int r, int c, int R, int Cc) {
  if (r >= R || c >= Cc) return 0.f;
  const int64_t off = TRANS ? static_cast<int64_t>(c) * ld + r : static_cast<int64_t>(r) * ld + c;
  float v = __ldg(X + off);
  if (mask) {
    const int64_t moff = TRANS ? static_cast<int64_t>(c) * ldm + r : static_cast<int64_t>(r) * ldm + c;
    if (!(__ldg(mask + moff) > 0.f)) v = 0.f;
  }
  return v;
}

__device__ __forceinline__ void store_out(const GemmArgs& g, int row, int col, float v) {
  if (row >= g.M || col >= g.N) return;
  if (g.bias) v += g.bias[col];
  if (g.relu) v = fmaxf(v, 0.f);
  float* p = g.C + static_cast<int64_t>(row) * g.ldc + col;
  if (g.k_splits > 1) atomicAdd(p, v);
  else if (g.accumulate) *p += v;
  else *p = v;
}

// ------------------------------------------------------------------------------ bf16 tensor cores
// smem tiles keep the global orientation: A [BM][BK] (or [BK][BM] when TRANS_A), B [BK][BN]
// (or [BN][BK] when TRANS_B); ldmatrix(.trans) produces the fragments for every combination.
template <bool TRANS_A, bool TRANS_B>
__global__ void __launch_bounds__(kGemmThreads)
gemm_bf16_kernel(const GemmArgs g) {
  constexpr int kAStride = TRANS_A ? BM + 8 : BK + 8;
  constexpr int kARows = TRANS_A ? BK : BM;
  constexpr int kBStride = TRANS_B ? BK + 8 : BN + 8;
  constexpr int kBRows = TRANS_B ? BN : BK;
  __shared__ __align__(16) __nv_bfloat16 sA[kARows * kAStride];
  __shared__ __align__(16) __nv_bfloat16 sB[kBRows * kBStride];
  __shared__ float s_colsum[BN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_chunks = (g.K + BK - 1) / BK;
  const int per = (k_chunks + g.k_splits - 1) / g.k_splits;
  const int kc_begin = blockIdx.z * per, kc_end = min(k_chunks, kc_begin + per);
  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  float csum = 0.f;  // column sum of B for column (tid % BN), rows handled by tid / BN
  if (tid < BN) s_colsum[tid] = 0.f;

  for (int kc = kc_begin; kc < kc_end; ++kc) {
    const int k0 = kc * BK;
    __syncthreads();
    // A tile
    if constexpr (!TRANS_A) {
      for (int idx = tid; idx < BM * BK; idx += kGemmThreads) {
        int r = idx / BK, c = idx % BK;
        sA[r * kAStride + c] = __float2bfloat16(load_elem<false>(g.A, g.lda, g.maskA, g.ldma, m0 + r, k0 + c, g.M, g.K));
      }
    } else {
      for (int idx = tid; idx < BK * BM; idx += kGemmThreads) {
        int kk = idx / BM, r = idx % BM;   // stored [k][m]
        sA[kk * kAStride + r] = __float2bfloat16(load_elem<true>(g.A, g.lda, g.maskA, g.ldma, m0 + r, k0 + kk, g.M, g.K));
      }
    }
    // B tile
    if constexpr (!TRANS_B) {
      for (int idx = tid; idx < BK * BN; idx += kGemmThreads) {
        int kk = idx / BN, c = idx % BN;   // stored [k][n]
        float v = load_elem<false>(g.B, g.ldb, g.maskB, g.ldmb, k0 + kk, n0 + c, g.K, g.N);
        sB[kk * kBStride + c] = __float2bfloat16(v);
        csum += v;                          // c == tid % BN for every idx of this thread
      }
    } else {
      for (int idx = tid; idx < BN * BK; idx += kGemmThreads) {
        int c = idx / BK, kk = idx % BK;   // stored [n][k]
        sB[c * kBStride + kk] = __float2bfloat16(load_elem<true>(g.B, g.ldb, g.maskB, g.ldmb, k0 + kk, n0 + c, g.K, g.N));
      }
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if constexpr (!TRANS_A) load_a_frag(a[mt], sA, kAStride, warp * 32 + mt * 16, ks * 16, lane);
        else load_a_frag_trans(a[mt], sA, kAStride, ks * 16, warp * 32 + mt * 16, lane);
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        if constexpr (!TRANS_B) load_b_frag2_trans(b, sB, kBStride, ks * 16, np * 16, lane);
        else load_b_frag2(b, sB, kBStride, np * 16, ks * 16, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16(acc[mt][2 * np], a[mt], b[0], b[1]);
          mma_bf16(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
        }
      }
    }
  }
  const int r = lane >> 2, cq = (lane & 3) * 2;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        store_out(g, m0 + warp * 32 + mt * 16 + r + (e >= 2 ? 8 : 0), n0 + nt * 8 + cq + (e & 1), acc[mt][nt][e]);
  if (g.colsum && !TRANS_B && blockIdx.x == 0) {
    // bias gradient: every CTA of the first M-tile row holds the sums of its K range
    __syncthreads();
    atomicAdd(&s_colsum[tid % BN], csum);
    __syncthreads();
    if (tid < BN && n0 + tid < g.N) atomicAdd(g.colsum + n0 + tid, s_colsum[tid]);
  }
}

// ------------------------------------------------------------------------------ fp32 FFMA
template <bool TRANS_A, bool TRANS_B>
__global__ void __launch_bounds__(kGemmThreads)
gemm_f32_kernel(const GemmArgs g) {
  __shared__ __align__(16) float sA[BK][BM + 4];   // [k][m]
  __shared__ __align__(16) float sB[BK][BN + 4];   // [k][n]
  __shared__ float s_colsum[BN];
  const int tid = threadIdx.x;
  const int ty = tid >> 3, tx = tid & 7;           // 16 x 8 threads, 8 x 8 outputs each
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_chunks = (g.K + BK - 1) / BK;
  const int per = (k_chunks + g.k_splits - 1) / g.k_splits;
  const int kc_begin = blockIdx.z * per, kc_end = min(k_chunks, kc_begin + per);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float csum = 0.f;
  if (tid < BN) s_colsum[tid] = 0.f;
  for (int kc = kc_begin; kc < kc_end; ++kc) {
    const int k0 = kc * BK;
    __syncthreads();
    if constexpr (!TRANS_A) {
      for (int idx = tid; idx < BM * BK; idx += kGemmThreads) {
        int r = idx / BK, c = idx % BK;
        sA[c][r] = load_elem<false>(g.A, g.lda, g.maskA, g.ldma, m0 + r, k0 + c, g.M, g.K);
      }
    } else {
      for (int idx = tid; idx < BK * BM; idx += kGemmThreads) {
        int kk = idx / BM, r = idx % BM;
        sA[kk][r] = load_elem<true>(g.A, g.lda, g.maskA, g.ldma, m0 + r, k0 + kk, g.M, g.K);
      }
    }
    if constexpr (!TRANS_B) {
      for (int idx = tid; idx < BK * BN; idx += kGemmThreads) {
        int kk = idx / BN, c = idx % BN;
        float v = load_elem<false>(g.B, g.ldb, g.maskB, g.ldmb, k0 + kk, n0 + c, g.K, g.N);
        sB[kk][c] = v;
        csum += v;
      }
    } else {
      for (int idx = tid; idx < BN * BK; idx += kGemmThreads) {
        int c = idx / BK, kk = idx % BK;
        sB[kk][c] = load_elem<true>(g.B, g.ldb, g.maskB, g.ldmb, k0 + kk, n0 + c, g.K, g.N);
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&sA[kk][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&sA[kk][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&sB[kk][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&sB[kk][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) store_out(g, m0 + ty * 8 + i, n0 + tx * 8 + j, acc[i][j]);
  if (g.colsum && !TRANS_B && blockIdx.x == 0) {
    __syncthreads();
    atomicAdd(&s_colsum[tid % BN], csum);
    __syncthreads();
    if (tid < BN && n0 + tid < g.N) atomicAdd(g.colsum + n0 + tid, s_colsum[tid]);
  }
}

template <bool TA, bool TB>
int32_t launch_gemm(cudaStream_t st, const GemmArgs& g, int bf16) {
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN, g.k_splits);
  if (bf16) gemm_bf16_kernel<TA, TB><<<grid, kGemmThreads, 0, st>>>(g);
  else gemm_f32_kernel<TA, TB><<<grid, kGemmThreads, 0, st>>>(g);
  return check_launch();
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_dense_fwd(void* stream, const float* d_x, int64_t ldx, const float* d_kernel,
                                 const float* d_bias, int64_t num_rows, int32_t in_dim, int32_t out_dim,
                                 int32_t relu, int32_t bf16, float* d_y, int64_t ldy) {
  if (num_rows < 0 || in_dim < 1 || out_dim < 1 || ldx < in_dim || ldy < out_dim) return NRC_E_INVALID_ARG;
  if (num_rows == 0) return NRC_OK;
  if (!d_x || !d_kernel || !d_y) return NRC_E_INVALID_ARG;
  if (num_rows > 0x7fffffff) return NRC_E_UNSUPPORTED;
  GemmArgs g{};
  g.A = d_x; g.lda = ldx; g.B = d_kernel; g.ldb = out_dim; g.C = d_y; g.ldc = ldy; g.bias = d_bias;
  g.M = static_cast<int>(num_rows); g.N = out_dim; g.K = in_dim; g.relu = relu; g.k_splits = 1;
  return launch_gemm<false, false>(static_cast<cudaStream_t>(stream), g, bf16);
}

extern "C" int32_t nrc_dense_bwd(void* stream, const float* d_x, int64_t ldx, const float* d_kernel,
                                 const float* d_y, int64_t ldy, const float* d_g_y, int64_t ldgy,
                                 int64_t num_rows, int32_t in_dim, int32_t out_dim, int32_t relu, int32_t bf16,
                                 float* d_g_x, int64_t ldgx, int32_t accumulate_g_x, float* d_g_kernel,
                                 float* d_g_bias) {
  if (num_rows < 0 || in_dim < 1 || out_dim < 1) return NRC_E_INVALID_ARG;
  if (num_rows == 0) return NRC_OK;
  if (!d_g_y || !d_kernel || (relu && !d_y)) return NRC_E_INVALID_ARG;
  if (num_rows > 0x7fffffff) return NRC_E_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* mask = relu ? d_y : nullptr;
  if (d_g_x) {
    // g_x[M,K] = (g_y * [y > 0]) @ kernel^T
    GemmArgs g{};
    g.A = d_g_y; g.lda = ldgy; g.maskA = mask; g.ldma = ldy;
    g.B = d_kernel; g.ldb = out_dim;
    g.C = d_g_x; g.ldc = ldgx; g.accumulate = accumulate_g_x;
    g.M = static_cast<int>(num_rows); g.N = in_dim; g.K = out_dim; g.k_splits = 1;
    int32_t s = launch_gemm<false, true>(st, g, bf16);
    if (s != NRC_OK) return s;
  }
  if (d_g_kernel) {
    if (!d_x) return NRC_E_INVALID_ARG;
    // g_kernel[K,N] += x^T @ (g_y * [y > 0]); g_bias[N] += column sums.  Reduction over rows is
    // split across CTAs (atomic epilogue) so that the grid covers the 148 SMs.
    GemmArgs g{};
    g.A = d_x; g.lda = ldx;
    g.B = d_g_y; g.ldb = ldgy; g.maskB = mask; g.ldmb = ldy;
    g.C = d_g_kernel; g.ldc = out_dim; g.colsum = d_g_bias;
    g.M = in_dim; g.N = out_dim; g.K = static_cast<int>(num_rows);
    int tiles = ((in_dim + BM - 1) / BM) * ((out_dim + BN - 1) / BN);
    int chunks = (g.K + BK - 1) / BK;
    int splits = (2 * kNumSMs + tiles - 1) / tiles;
    g.k_splits = splits < 1 ? 1 : (splits > chunks ? chunks : splits);
    if (g.k_splits == 1) g.accumulate = 1;
    int32_t s = launch_gemm<true, false>(st, g, bf16);
    if (s != NRC_OK) return s;
  }
  return NRC_OK;
}
