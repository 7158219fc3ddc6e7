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
constexpr int kGemmThreads = 256;   // 8 warps: 4 (M) x 2 (N), warp tile 32 x 32

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  const float* bias;          // [N] or null (forward)
  float* colsum;              // [N]: += column sums of opB(B) rows (bias gradient) or null
  int M, N, K;
  int relu;                   // epilogue activation
  int accumulate;             // C += result instead of C = result
  int k_splits;               // >1: split the K loop over blockIdx.z, atomicAdd epilogue
};

// ---------------------------------------------------------------------------- tile loader
// A tile is ROWS x COLS in the operand's *global* orientation (row-major, stride ld); each thread
// moves NV = ROWS*COLS/4/threads float4 units.  The 16-byte path needs ld % 4 == 0 and a
// 16-byte aligned base (the host pads odd widths); columns beyond the logical limit are zeroed.
template <int ROWS, int COLS>
struct TileLoader {
  static constexpr int kUnits = ROWS * COLS / 4;
  static constexpr int NV = kUnits / kGemmThreads;
  static_assert(kUnits % kGemmThreads == 0, "tile must divide evenly");
  float4 v[NV];

  __device__ __forceinline__ void load(const float* __restrict__ X, int64_t ld, int row0, int col0, int row_lim,
                                       int col_lim, bool vec_ok, int tid) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int u = tid + j * kGemmThreads;
      const int r = u / (COLS / 4), c = (u % (COLS / 4)) * 4;
      const int gr = row0 + r, gc = col0 + c;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < row_lim && gc < col_lim) {
        const float* p = X + static_cast<int64_t>(gr) * ld + gc;
        if (vec_ok && gc + 4 <= ld) {
          t = __ldg(reinterpret_cast<const float4*>(p));
          if (gc + 1 >= col_lim) t.y = 0.f;
          if (gc + 2 >= col_lim) t.z = 0.f;
          if (gc + 3 >= col_lim) t.w = 0.f;
        } else {
          t.x = __ldg(p);
          if (gc + 1 < col_lim) t.y = __ldg(p + 1);
          if (gc + 2 < col_lim) t.z = __ldg(p + 2);
          if (gc + 3 < col_lim) t.w = __ldg(p + 3);
        }
      }
      v[j] = t;
    }
  }
  // bf16 tile, same orientation, row stride `stride` halves
  __device__ __forceinline__ void store_bf16(__nv_bfloat16* s, int stride, int tid) const {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int u = tid + j * kGemmThreads;
      const int r = u / (COLS / 4), c = (u % (COLS / 4)) * 4;
      *reinterpret_cast<uint2*>(s + r * stride + c) = make_uint2(pack_bf16(v[j].x, v[j].y), pack_bf16(v[j].z, v[j].w));
    }
  }
  // fp32 tile; transposed => s[c][r]
  template <bool TRANSPOSE>
  __device__ __forceinline__ void store_f32(float* s, int stride, int tid) const {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int u = tid + j * kGemmThreads;
      const int r = u / (COLS / 4), c = (u % (COLS / 4)) * 4;
      if constexpr (!TRANSPOSE) {
        *reinterpret_cast<float4*>(s + r * stride + c) = v[j];
      } else {
        s[(c + 0) * stride + r] = v[j].x; s[(c + 1) * stride + r] = v[j].y;
        s[(c + 2) * stride + r] = v[j].z; s[(c + 3) * stride + r] = v[j].w;
      }
    }
  }
  __device__ __forceinline__ float sum() const {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) t += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    return t;
  }
};

__device__ __forceinline__ bool vec_ok(const float* p, int64_t ld) {
  return (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
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

// Column sums of the B operand (bias gradient), only for !TRANS_B tiles [BK][BN]: every float4
// unit of a thread lies in columns (u % (BN/4))*4.. with u % 16 == tid % 16 fixed.
template <typename Loader>
__device__ __forceinline__ void add_colsum(const Loader& L, float (&cs)[4]) {
#pragma unroll
  for (int j = 0; j < Loader::NV; ++j) { cs[0] += L.v[j].x; cs[1] += L.v[j].y; cs[2] += L.v[j].z; cs[3] += L.v[j].w; }
}

// ------------------------------------------------------------------------------ bf16 tensor cores
// smem tiles keep the global orientation: A [BM][BK] (or [BK][BM] when TRANS_A), B [BK][BN]
// (or [BN][BK] when TRANS_B); ldmatrix(.trans) produces the fragments for every combination.
// Global loads of chunk k+1 are issued into registers before the MMAs of chunk k.
template <bool TRANS_A, bool TRANS_B>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_bf16_kernel(const GemmArgs g) {
  constexpr int kARows = TRANS_A ? BK : BM, kACols = TRANS_A ? BM : BK;
  constexpr int kBRows = TRANS_B ? BN : BK, kBCols = TRANS_B ? BK : BN;
  constexpr int kAStride = kACols + 8, kBStride = kBCols + 8;
  __shared__ __align__(16) __nv_bfloat16 sA[kARows * kAStride];
  __shared__ __align__(16) __nv_bfloat16 sB[kBRows * kBStride];
  __shared__ float s_colsum[BN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;   // warp tile origin inside the CTA tile
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_chunks = (g.K + BK - 1) / BK;
  const int per = (k_chunks + g.k_splits - 1) / g.k_splits;
  const int kc_begin = blockIdx.z * per, kc_end = min(k_chunks, kc_begin + per);
  const bool va = vec_ok(g.A, g.lda), vb = vec_ok(g.B, g.ldb);
  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  if (tid < BN) s_colsum[tid] = 0.f;
  TileLoader<kARows, kACols> la;
  TileLoader<kBRows, kBCols> lb;
  auto issue = [&](int kc) {
    const int k0 = kc * BK;
    if constexpr (!TRANS_A) la.load(g.A, g.lda, m0, k0, g.M, g.K, va, tid);
    else la.load(g.A, g.lda, k0, m0, g.K, g.M, va, tid);
    if constexpr (!TRANS_B) lb.load(g.B, g.ldb, k0, n0, g.K, g.N, vb, tid);
    else lb.load(g.B, g.ldb, n0, k0, g.N, g.K, vb, tid);
  };
  if (kc_begin < kc_end) issue(kc_begin);
  for (int kc = kc_begin; kc < kc_end; ++kc) {
    __syncthreads();                       // previous chunk's fragments are consumed
    la.store_bf16(sA, kAStride, tid);
    lb.store_bf16(sB, kBStride, tid);
    if (g.colsum && !TRANS_B) add_colsum(lb, cs);
    __syncthreads();
    if (kc + 1 < kc_end) issue(kc + 1);    // in flight during the MMAs below
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if constexpr (!TRANS_A) load_a_frag(a[mt], sA, kAStride, wm + mt * 16, ks * 16, lane);
        else load_a_frag_trans(a[mt], sA, kAStride, ks * 16, wm + mt * 16, lane);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        if constexpr (!TRANS_B) load_b_frag2_trans(b, sB, kBStride, ks * 16, wn + np * 16, lane);
        else load_b_frag2(b, sB, kBStride, wn + np * 16, ks * 16, lane);
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
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        store_out(g, m0 + wm + mt * 16 + r + (e >= 2 ? 8 : 0), n0 + wn + nt * 8 + cq + (e & 1), acc[mt][nt][e]);
  if (g.colsum && !TRANS_B && blockIdx.x == 0) {
    // bias gradient: the CTAs of the first M-tile row hold the column sums of their K range
    __syncthreads();
    const int c = (tid % (BN / 4)) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(&s_colsum[c + i], cs[i]);
    __syncthreads();
    if (tid < BN && n0 + tid < g.N) atomicAdd(g.colsum + n0 + tid, s_colsum[tid]);
  }
}

// ------------------------------------------------------------------------------ fp32 FFMA
template <bool TRANS_A, bool TRANS_B>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_f32_kernel(const GemmArgs g) {
  constexpr int kARows = TRANS_A ? BK : BM, kACols = TRANS_A ? BM : BK;
  constexpr int kBRows = TRANS_B ? BN : BK, kBCols = TRANS_B ? BK : BN;
  __shared__ __align__(16) float sA[BK][BM + 4];   // [k][m]
  __shared__ __align__(16) float sB[BK][BN + 4];   // [k][n]
  __shared__ float s_colsum[BN];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;          // 16 x 16 threads, 8 x 4 outputs each
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_chunks = (g.K + BK - 1) / BK;
  const int per = (k_chunks + g.k_splits - 1) / g.k_splits;
  const int kc_begin = blockIdx.z * per, kc_end = min(k_chunks, kc_begin + per);
  const bool va = vec_ok(g.A, g.lda), vb = vec_ok(g.B, g.ldb);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  if (tid < BN) s_colsum[tid] = 0.f;
  TileLoader<kARows, kACols> la;
  TileLoader<kBRows, kBCols> lb;
  auto issue = [&](int kc) {
    const int k0 = kc * BK;
    if constexpr (!TRANS_A) la.load(g.A, g.lda, m0, k0, g.M, g.K, va, tid);
    else la.load(g.A, g.lda, k0, m0, g.K, g.M, va, tid);
    if constexpr (!TRANS_B) lb.load(g.B, g.ldb, k0, n0, g.K, g.N, vb, tid);
    else lb.load(g.B, g.ldb, n0, k0, g.N, g.K, vb, tid);
  };
  if (kc_begin < kc_end) issue(kc_begin);
  for (int kc = kc_begin; kc < kc_end; ++kc) {
    __syncthreads();
    la.template store_f32<!TRANS_A>(&sA[0][0], BM + 4, tid);   // smem is [k][m]: transpose when A is [m][k]
    lb.template store_f32<TRANS_B>(&sB[0][0], BN + 4, tid);    // smem is [k][n]: transpose when B is [n][k]
    if (g.colsum && !TRANS_B) add_colsum(lb, cs);
    __syncthreads();
    if (kc + 1 < kc_end) issue(kc + 1);
#pragma unroll 8
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[4];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&sA[kk][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&sA[kk][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) store_out(g, m0 + ty * 8 + i, n0 + tx * 4 + j, acc[i][j]);
  if (g.colsum && !TRANS_B && blockIdx.x == 0) {
    __syncthreads();
    const int c = (tid % (BN / 4)) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(&s_colsum[c + i], cs[i]);
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

namespace nrc {
__global__ void relu_bwd_kernel(const float* __restrict__ y, int64_t ldy, const float* __restrict__ g, int64_t ldg,
                                int64_t rows, int cols, float* __restrict__ out, int64_t ldo) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int64_t r = i / cols;
  const int c = static_cast<int>(i - r * cols);
  out[r * ldo + c] = y[r * ldy + c] > 0.f ? g[r * ldg + c] : 0.f;
}
}  // namespace nrc

extern "C" int32_t nrc_relu_bwd(void* stream, const float* d_y, int64_t ldy, const float* d_g_y, int64_t ldgy,
                                int64_t num_rows, int32_t num_cols, float* d_g_pre, int64_t ldgp) {
  if (num_rows < 0 || num_cols < 1) return NRC_E_INVALID_ARG;
  if (num_rows == 0) return NRC_OK;
  if (!d_y || !d_g_y || !d_g_pre) return NRC_E_INVALID_ARG;
  int64_t total = num_rows * num_cols;
  relu_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_y, ldy, d_g_y, ldgy, num_rows, num_cols, d_g_pre, ldgp);
  return check_launch();
}

extern "C" int32_t nrc_dense_bwd(void* stream, const float* d_x, int64_t ldx, const float* d_kernel,
                                 const float* d_g_y, int64_t ldgy, int64_t num_rows, int32_t in_dim,
                                 int32_t out_dim, int32_t bf16, float* d_g_x, int64_t ldgx,
                                 int32_t accumulate_g_x, float* d_g_kernel, float* d_g_bias) {
  if (num_rows < 0 || in_dim < 1 || out_dim < 1) return NRC_E_INVALID_ARG;
  if (num_rows == 0) return NRC_OK;
  if (!d_g_y || !d_kernel) return NRC_E_INVALID_ARG;
  if (num_rows > 0x7fffffff) return NRC_E_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_g_x) {
    // g_x[M,K] = g_pre @ kernel^T
    GemmArgs g{};
    g.A = d_g_y; g.lda = ldgy;
    g.B = d_kernel; g.ldb = out_dim;
    g.C = d_g_x; g.ldc = ldgx; g.accumulate = accumulate_g_x;
    g.M = static_cast<int>(num_rows); g.N = in_dim; g.K = out_dim; g.k_splits = 1;
    int32_t s = launch_gemm<false, true>(st, g, bf16);
    if (s != NRC_OK) return s;
  }
  if (d_g_kernel) {
    if (!d_x) return NRC_E_INVALID_ARG;
    // g_kernel[K,N] += x^T @ g_pre; g_bias[N] += column sums.  The reduction over rows is split
    // across CTAs (atomic epilogue) so that the grid covers the 148 SMs.
    GemmArgs g{};
    g.A = d_x; g.lda = ldx;
    g.B = d_g_y; g.ldb = ldgy;
    g.C = d_g_kernel; g.ldc = out_dim; g.colsum = d_g_bias;
    g.M = in_dim; g.N = out_dim; g.K = static_cast<int>(num_rows);
    int tiles = ((in_dim + BM - 1) / BM) * ((out_dim + BN - 1) / BN);
    int chunks = (g.K + BK - 1) / BK;
    int splits = (2 * num_sms() + tiles - 1) / tiles;
    g.k_splits = splits < 1 ? 1 : (splits > chunks ? chunks : splits);
    if (g.k_splits == 1) g.accumulate = 1;
    int32_t s = launch_gemm<true, false>(st, g, bf16);
    if (s != NRC_OK) return s;
  }
  return NRC_OK;
}
