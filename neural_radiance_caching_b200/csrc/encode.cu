// K1 / K2: multiresolution hash-grid encoding, forward gather and backward
// scatter-add.  Replaces HashEncoding.__call__ (internal/grid_utils.py:807-905)
// and XLA's transpose of its gathers.
//
// Work decomposition: one thread per (point, level); blockIdx.y = level so that
// the CTAs resident at any instant mostly touch ONE level's table (2-8 MB), which
// stays L2/L1 resident; threadIdx.x runs over consecutive points (= consecutive
// samples of one ray), so coarse-level corner fetches of a warp coalesce.
#include <cstdlib>

#include "encode.cuh"

namespace nrc {

thread_local int g_last_cuda_error = 0;

constexpr int kEncThreads = 256;

template <int F>
__global__ void __launch_bounds__(kEncThreads)
encode_fwd_kernel(const __grid_constant__ EncDev enc, const float* __restrict__ x, int64_t P,
                  float* __restrict__ out) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kEncThreads + threadIdx.x;
  if (p >= P) return;
  const int l = blockIdx.y;
  const LevelDev& lv = enc.lv[l];
  float xi[3] = {__ldg(x + 3 * p), __ldg(x + 3 * p + 1), __ldg(x + 3 * p + 2)};
  float xn[3];
  normalise_point(enc, xi, xn);
  Corners c = level_setup(lv, xn);
  FeatVec<F> acc = level_interp<F>(lv, c);
  float* o = out + p * (enc.L * F) + l * F;
  if constexpr (F == 4) {
    float4 r = make_float4(__fmul_rn(acc.v[0], enc.scale), __fmul_rn(acc.v[1], enc.scale),
                           __fmul_rn(acc.v[2], enc.scale), __fmul_rn(acc.v[3], enc.scale));
    *reinterpret_cast<float4*>(o) = r;
  } else {
#pragma unroll
    for (int f = 0; f < F; ++f) o[f] = __fmul_rn(acc.v[f], enc.scale);
  }
}

__global__ void __launch_bounds__(kEncThreads)
encode_indices_kernel(const __grid_constant__ EncDev enc, int level, const float* __restrict__ x,
                      int64_t P, int32_t* __restrict__ idx) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kEncThreads + threadIdx.x;
  if (p >= P) return;
  const LevelDev& lv = enc.lv[level];
  float xi[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
  float xn[3];
  normalise_point(enc, xi, xn);
  Corners c = level_setup(lv, xn);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int bx, by, bz;
    corner_bits(lv.is_hash, k, bx, by, bz);
    idx[p * 8 + k] = lv.is_hash ? corner_row(lv, c, bx, by, bz) : corner_padded_index(lv, c, bx, by, bz);
  }
}

// Backward: scatter-add into the level's gradient table and (optionally) the VJP
// with respect to x, accumulated over levels with atomics into a zeroed [P,3].
template <int F, bool kTableGrad, bool kXGrad>
__global__ void __launch_bounds__(kEncThreads)
encode_bwd_kernel(const __grid_constant__ EncDev enc, const float* __restrict__ x,
                  const float* __restrict__ g_out, int64_t P, float* __restrict__ g_x, const int aggregate,
                  const float warp_c) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kEncThreads + threadIdx.x;
  const int l = blockIdx.y;
  const LevelDev& lv = enc.lv[l];
  if constexpr (kTableGrad && !kXGrad) {
    // aggregate: 1 = dense levels only, 2 = hash levels too (consecutive samples of a ray share fine cells wherever the
    // sampler has concentrated them at a surface)
    if (aggregate && lv.grad && (aggregate > 1 || !lv.is_hash)) {   // block-uniform: every lane stays for the shuffles
      const bool valid = p < P;
      const int64_t pc = valid ? p : P - 1;
      float xi[3] = {__ldg(x + 3 * pc), __ldg(x + 3 * pc + 1), __ldg(x + 3 * pc + 2)};
      contract_point(warp_c, xi[0], xi[1], xi[2], xi[0], xi[1], xi[2]);   // identity when warp_c <= 0
      float xn[3];
      normalise_point(enc, xi, xn);
      const Corners c = level_setup(lv, xn);
      float g[F];
      const float* gp = g_out + pc * (enc.L * F) + l * F;
#pragma unroll
      for (int f = 0; f < F; ++f) g[f] = valid ? __ldg(gp + f) * enc.scale : 0.f;
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int bx, by, bz;
        corner_bits(lv.is_hash, k, bx, by, bz);
        const int32_t row = valid ? corner_row(lv, c, bx, by, bz) : -1;
        const float w = (bx ? c.cw[0] : c.fw[0]) * (by ? c.cw[1] : c.fw[1]) * (bz ? c.cw[2] : c.fw[2]);
        float gw[F];
#pragma unroll
        for (int f = 0; f < F; ++f) gw[f] = g[f] * w;
        warp_run_atomic_add<F>(lv.grad, row, gw, lane);
      }
      return;
    }
  }
  if (p >= P) return;
  float xi[3] = {__ldg(x + 3 * p), __ldg(x + 3 * p + 1), __ldg(x + 3 * p + 2)};
  contract_point(warp_c, xi[0], xi[1], xi[2], xi[0], xi[1], xi[2]);
  float xn[3];
  normalise_point(enc, xi, xn);
  Corners c = level_setup(lv, xn);
  float g[F];
  const float* gp = g_out + p * (enc.L * F) + l * F;
#pragma unroll
  for (int f = 0; f < F; ++f) g[f] = __ldg(gp + f) * enc.scale;

  float gx[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int bx, by, bz;
    corner_bits(lv.is_hash, k, bx, by, bz);
    int32_t row = corner_row(lv, c, bx, by, bz);
    if (row < 0) continue;
    float wx = bx ? c.cw[0] : c.fw[0];
    float wy = by ? c.cw[1] : c.fw[1];
    float wz = bz ? c.cw[2] : c.fw[2];
    if constexpr (kTableGrad) {
      if (lv.grad) {
        float w = wx * wy * wz;
        float gw[F];
#pragma unroll
        for (int f = 0; f < F; ++f) gw[f] = g[f] * w;
        atomic_add_row<F>(lv.grad, row, gw);
      }
    }
    if constexpr (kXGrad) {
      FeatVec<F> v = load_row<F>(lv.table, row);
      float dot = 0.f;
#pragma unroll
      for (int f = 0; f < F; ++f) dot += g[f] * v.v[f];
      gx[0] += (bx ? dot : -dot) * (wy * wz);
      gx[1] += (by ? dot : -dot) * (wx * wz);
      gx[2] += (bz ? dot : -dot) * (wx * wy);
    }
  }
  if constexpr (kXGrad) {
    const float fN = static_cast<float>(lv.N);
#pragma unroll
    for (int a = 0; a < 3; ++a) atomicAdd(g_x + 3 * p + a, gx[a] * (fN / enc.span[a]));
  }
}

// ---------------------------------------------------------------------------------------------------
// Tangent of the encoding along a per-point direction (second-order path of the analytic normals, SURVEY 8f-1,
// internal/geometry.py:442-460 differentiated again by the predicted-normal loss).  With x the sample mean,
// z = contract(x / c), zdot = J_contract g (J is symmetric: the VJP routine gives the JVP) and, per level and corner,
//     dw_c = sum_a (+-) zdot_a * N / span_a * prod_{b != a} w_b          (the derivative of the trilinear weight),
// kTangentFwd:  edot[l*F+f]  = scale * sum_c dw_c T[c][f]                 (tangent features, [P, L*F])
// else       :  dT[c][f]    += dw_c * scale * ge[l*F+f]                   (table gradient of <ge, edot>)
// Same thread mapping as encode_fwd/bwd (thread per point, blockIdx.y = level); the scatter uses the dense-level
// run aggregation.  fp32 throughout; the MLP between the two is nrc_density_mlp_bwd_tangent (bf16 tensor cores).
template <int F, bool kTangentFwd>
__global__ void __launch_bounds__(kEncThreads)
encode_tangent_kernel(const __grid_constant__ EncDev enc, const float* __restrict__ means, const float* __restrict__ g,
                      const float* __restrict__ ge, int64_t P, float warp_c, float* __restrict__ edot) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kEncThreads + threadIdx.x;
  const int l = blockIdx.y;
  const LevelDev& lv = enc.lv[l];
  const bool valid = p < P;
  const int64_t pc = valid ? p : P - 1;
  const float x0 = __ldg(means + 3 * pc), x1 = __ldg(means + 3 * pc + 1), x2 = __ldg(means + 3 * pc + 2);
  float z[3], xn[3], zd[3];
  contract_point(warp_c, x0, x1, x2, z[0], z[1], z[2]);
  normalise_point(enc, z, xn);
  contract_vjp(warp_c, x0, x1, x2, __ldg(g + 3 * pc), __ldg(g + 3 * pc + 1), __ldg(g + 3 * pc + 2), zd[0], zd[1], zd[2]);
  const Corners c = level_setup(lv, xn);
  const float fN = static_cast<float>(lv.N);
  const float t0 = zd[0] * (fN / enc.span[0]), t1 = zd[1] * (fN / enc.span[1]), t2 = zd[2] * (fN / enc.span[2]);
  const int LF = enc.L * F;
  float gl[F], ed[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    ed[f] = 0.f;
    gl[f] = 0.f;
    if constexpr (!kTangentFwd) gl[f] = valid ? __ldg(ge + pc * LF + l * F + f) * enc.scale : 0.f;
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int bx, by, bz;
    corner_bits(lv.is_hash, k, bx, by, bz);
    const int32_t row = valid ? corner_row(lv, c, bx, by, bz) : -1;
    const float wx = bx ? c.cw[0] : c.fw[0];
    const float wy = by ? c.cw[1] : c.fw[1];
    const float wz = bz ? c.cw[2] : c.fw[2];
    const float dw = (bx ? t0 : -t0) * (wy * wz) + (by ? t1 : -t1) * (wx * wz) + (bz ? t2 : -t2) * (wx * wy);
    if constexpr (kTangentFwd) {
      if (row >= 0) {
        const FeatVec<F> v = load_row<F>(lv.table, row);
#pragma unroll
        for (int f = 0; f < F; ++f) ed[f] = fmaf(dw, v.v[f], ed[f]);
      }
    } else {
      float gw[F];
#pragma unroll
      for (int f = 0; f < F; ++f) gw[f] = dw * gl[f];
      if (!lv.is_hash) warp_run_atomic_add<F>(lv.grad, row, gw, lane);     // block-uniform branch
      else if (row >= 0) atomic_add_row<F>(lv.grad, row, gw);
    }
  }
  if constexpr (kTangentFwd) {
    if (valid) {
#pragma unroll
      for (int f = 0; f < F; ++f) edot[p * LF + l * F + f] = ed[f] * enc.scale;
    }
  }
}

template <int F>
int32_t launch_tangent(cudaStream_t s, const EncDev& d, const float* means, const float* g, const float* ge, int64_t P,
                       float warp_c, float* edot) {
  dim3 grid(static_cast<unsigned>((P + kEncThreads - 1) / kEncThreads), d.L);
  if (edot) encode_tangent_kernel<F, true><<<grid, kEncThreads, 0, s>>>(d, means, g, nullptr, P, warp_c, edot);
  else encode_tangent_kernel<F, false><<<grid, kEncThreads, 0, s>>>(d, means, g, ge, P, warp_c, nullptr);
  return check_launch();
}

template <int F>
int32_t launch_fwd(cudaStream_t s, const EncDev& d, const float* x, int64_t P, float* out) {
  dim3 grid(static_cast<unsigned>((P + kEncThreads - 1) / kEncThreads), d.L);
  encode_fwd_kernel<F><<<grid, kEncThreads, 0, s>>>(d, x, P, out);
  return check_launch();
}

template <int F>
int32_t launch_bwd(cudaStream_t s, const EncDev& d, const float* x, const float* g, int64_t P,
                   float* g_x, bool table_grad, float warp_c = 0.f) {
  dim3 grid(static_cast<unsigned>((P + kEncThreads - 1) / kEncThreads), d.L);
  static const int agg = getenv("NRC_ENC_BWD_AGG") ? atoi(getenv("NRC_ENC_BWD_AGG")) : 2;
  if (table_grad && g_x) encode_bwd_kernel<F, true, true><<<grid, kEncThreads, 0, s>>>(d, x, g, P, g_x, 0, warp_c);
  else if (table_grad) encode_bwd_kernel<F, true, false><<<grid, kEncThreads, 0, s>>>(d, x, g, P, g_x, agg, warp_c);
  else if (g_x) encode_bwd_kernel<F, false, true><<<grid, kEncThreads, 0, s>>>(d, x, g, P, g_x, 0, warp_c);
  return check_launch();
}

// param_regularizer_loss (internal/train_utils.py:1169-1216) for one grid module with the common setting
// (mult, jnp.mean, alpha = 2, scale = 1): per level table  loss += mult * 0.5 * mean(T^2),  dT += mult * T / numel.
// blockIdx.y = level; grid-stride over the table; one atomic per block for the loss.
// overwrite != 0: the gradient tables are INITIALISED with the regularizer's gradient (plain 16-byte stores) instead of
// being zero-filled first and then atomically added to: the launch doubles as the memset of these tables.
__global__ void __launch_bounds__(256) grid_regularizer_kernel(const __grid_constant__ EncDev d, float mult,
                                                                 float* __restrict__ loss, const int overwrite) {
  const LevelDev& lv = d.lv[blockIdx.y];
  const size_t n = static_cast<size_t>(lv.T) * d.F;
  const float inv_n = 1.0f / static_cast<float>(n);
  const float* __restrict__ t = lv.table;
  float* __restrict__ g = lv.grad;
  float acc = 0.f;
  const size_t n4 = n >> 2;   // tables are 16-byte aligned slices of the arena with numel % 4 == 0 in every config
  const bool vec = ((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(g)) & 15) == 0 && (n & 3) == 0;
  if (vec) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(t) + i);
      const float k = mult * inv_n;
      if (overwrite == 2) {
        // evict-first stores: the gradient tables are next touched by the backward scatters, the level TABLES are what the
        // forward gathers (running beside this launch) want to keep in L2
        __stcs(reinterpret_cast<float4*>(g + 4 * i), make_float4(k * v.x, k * v.y, k * v.z, k * v.w));
      } else if (overwrite) {
        *reinterpret_cast<float4*>(g + 4 * i) = make_float4(k * v.x, k * v.y, k * v.z, k * v.w);
      } else {
        // 16-byte reduction: commutes with the scatter kernels' atomics on the same tables, so no stream ordering
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + 4 * i), "f"(k * v.x), "f"(k * v.y),
                     "f"(k * v.z), "f"(k * v.w)
                     : "memory");
      }
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const float v = __ldg(t + i);
      if (overwrite) g[i] = mult * inv_n * v;
      else atomicAdd(g + i, mult * inv_n * v);
      acc += v * v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int k = 0; k < 8; ++k) sum += part[k];
    atomicAdd(loss, 0.5f * mult * inv_n * sum);
  }
}

// Zero-fill of up to 8 float ranges of one buffer in ONE launch (the gradient arena minus the tables the regularizer
// initialises): 16-byte stores, optionally evict-first (streaming) so that the fill does not push the level tables out of L2
// while the sampler's forward gathers them.
struct ZeroRanges { int n; long long lo4[8], cnt4[8]; };
__global__ void __launch_bounds__(256) zero_ranges_kernel(float* __restrict__ base, const ZeroRanges r, const int streaming) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < r.n; ++k) {
    float4* p = reinterpret_cast<float4*>(base) + r.lo4[k];
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < r.cnt4[k];
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      if (streaming) __stcs(p + i, z); else p[i] = z;
    }
  }
}

// Random-gather probe (measurement aid, see nrc_probe_gather in the header): the per-thread row sequence is a
// counter-based integer hash, so consecutive lanes and consecutive iterations hit unrelated rows.
template <int F>
__global__ void __launch_bounds__(256) probe_gather_kernel(const float* __restrict__ table, uint32_t rows,
                                                            int64_t n, int per_thread, float* __restrict__ sink) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  uint32_t state = static_cast<uint32_t>(tid) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < per_thread; it += 8) {
    uint32_t r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      state ^= state << 13; state ^= state >> 17; state ^= state << 5;   // xorshift32
      r[k] = state % rows;
    }
    FeatVec<F> v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = load_row<F>(table, static_cast<int32_t>(r[k]));
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int f = 0; f < F; ++f) acc += v[k].v[f];
  }
  if (acc == 123456.789f) atomicAdd(sink, acc);   // never true for the bench tables; keeps the loads alive
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_probe_gather(void* stream, const float* d_table, int64_t table_rows, int32_t row_bytes,
                                    int64_t num_threads, int32_t per_thread, float* d_sink) {
  if (!d_table || !d_sink || table_rows < 1 || table_rows > 0x7fffffff || num_threads < 0 || per_thread < 8 ||
      (per_thread & 7))
    return NRC_E_INVALID_ARG;
  if (num_threads == 0) return NRC_OK;
  const unsigned grid = static_cast<unsigned>((num_threads + 255) / 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (row_bytes == 4) probe_gather_kernel<1><<<grid, 256, 0, s>>>(d_table, (uint32_t)table_rows, num_threads, per_thread, d_sink);
  else if (row_bytes == 16) probe_gather_kernel<4><<<grid, 256, 0, s>>>(d_table, (uint32_t)table_rows, num_threads, per_thread, d_sink);
  else return NRC_E_UNSUPPORTED;
  return check_launch();
}

static int32_t grid_regularizer(void* stream, const nrc_encoding_t* enc, float mult, float* d_loss, int overwrite);

extern "C" int32_t nrc_grid_regularizer(void* stream, const nrc_encoding_t* enc, float mult, float* d_loss) {
  return grid_regularizer(stream, enc, mult, d_loss, 0);
}
extern "C" int32_t nrc_grid_regularizer_init(void* stream, const nrc_encoding_t* enc, float mult, float* d_loss) {
  // NRC_STREAM_FILLS=0: plain stores (the gradient lines then compete with the level tables for the L2 during the forward)
  static const int mode = (getenv("NRC_STREAM_FILLS") && getenv("NRC_STREAM_FILLS")[0] == '0') ? 1 : 2;
  return grid_regularizer(stream, enc, mult, d_loss, mode);
}

extern "C" int32_t nrc_zero_ranges(void* stream, float* d_base, const int64_t* lo, const int64_t* hi, int32_t num_ranges,
                                   int32_t streaming) {
  if (!d_base || num_ranges < 0 || num_ranges > 8 || (num_ranges > 0 && (!lo || !hi))) return NRC_E_INVALID_ARG;
  nrc::ZeroRanges r;
  r.n = 0;
  long long most = 0;
  for (int k = 0; k < num_ranges; ++k) {
    if (lo[k] < 0 || hi[k] < lo[k] || (lo[k] & 3) || (hi[k] & 3)) return NRC_E_INVALID_ARG;
    if (hi[k] == lo[k]) continue;
    r.lo4[r.n] = lo[k] >> 2;
    r.cnt4[r.n] = (hi[k] - lo[k]) >> 2;
    most = r.cnt4[r.n] > most ? r.cnt4[r.n] : most;
    ++r.n;
  }
  if ((reinterpret_cast<uintptr_t>(d_base) & 15) != 0) return NRC_E_INVALID_ARG;
  if (r.n == 0) return NRC_OK;
  const long long want = (most + 255) / 256;
  const unsigned grid = static_cast<unsigned>(want < 4LL * nrc::num_sms() ? (want > 0 ? want : 1) : 4LL * nrc::num_sms());
  nrc::zero_ranges_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_base, r, streaming);
  return nrc::check_launch();
}

static int32_t grid_regularizer(void* stream, const nrc_encoding_t* enc, float mult, float* d_loss, int overwrite) {
  EncDev d;
  const int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  if (!d_loss) return NRC_E_INVALID_ARG;
  for (int l = 0; l < d.L; ++l)
    if (!d.lv[l].grad) return NRC_E_INVALID_ARG;
  grid_regularizer_kernel<<<dim3(2 * num_sms(), d.L), 256, 0, static_cast<cudaStream_t>(stream)>>>(d, mult, d_loss, overwrite);
  return check_launch();
}

extern "C" int32_t nrc_encode_fwd(void* stream, const nrc_encoding_t* enc, const float* d_x,
                                  int64_t num_points, float* d_out) {
  EncDev d;
  int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  if (num_points < 0 || (num_points > 0 && (!d_x || !d_out))) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (d.F) {
    case 1: return launch_fwd<1>(s, d, d_x, num_points, d_out);
    case 2: return launch_fwd<2>(s, d, d_x, num_points, d_out);
    case 4: return launch_fwd<4>(s, d, d_x, num_points, d_out);
    case 8: return launch_fwd<8>(s, d, d_x, num_points, d_out);
  }
  return NRC_E_UNSUPPORTED;
}

extern "C" int32_t nrc_encode_indices(void* stream, const nrc_encoding_t* enc, int32_t level,
                                      const float* d_x, int64_t num_points, int32_t* d_idx) {
  EncDev d;
  int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  if (level < 0 || level >= d.L || num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_x || !d_idx) return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned grid = static_cast<unsigned>((num_points + kEncThreads - 1) / kEncThreads);
  encode_indices_kernel<<<grid, kEncThreads, 0, s>>>(d, level, d_x, num_points, d_idx);
  return check_launch();
}

extern "C" int32_t nrc_encode_bwd(void* stream, const nrc_encoding_t* enc, const float* d_x,
                                  const float* d_g_out, int64_t num_points, float* d_g_x) {
  EncDev d;
  int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_x || !d_g_out) return NRC_E_INVALID_ARG;
  bool table_grad = false;
  for (int l = 0; l < d.L; ++l) table_grad |= (d.lv[l].grad != nullptr);
  if (!table_grad && !d_g_x) return NRC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d_g_x) {
    cudaError_t e = cudaMemsetAsync(d_g_x, 0, sizeof(float) * 3 * num_points, s);
    if (e != cudaSuccess) { g_last_cuda_error = e; return NRC_E_CUDA; }
  }
  switch (d.F) {
    case 1: return launch_bwd<1>(s, d, d_x, d_g_out, num_points, d_g_x, table_grad);
    case 2: return launch_bwd<2>(s, d, d_x, d_g_out, num_points, d_g_x, table_grad);
    case 4: return launch_bwd<4>(s, d, d_x, d_g_out, num_points, d_g_x, table_grad);
    case 8: return launch_bwd<8>(s, d, d_x, d_g_out, num_points, d_g_x, table_grad);
  }
  return NRC_E_UNSUPPORTED;
}

// Table scatter of a query's VJP straight from the sample means: the contraction is recomputed per thread (a dozen
// flops) instead of a nrc_contract_fwd launch and a [P,3] round trip in front of every nrc_encode_bwd of the step.
extern "C" int32_t nrc_encode_bwd_warped(void* stream, const nrc_encoding_t* enc, const float* d_means, float warp_c,
                                         const float* d_g_out, int64_t num_points) {
  EncDev d;
  int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_means || !d_g_out) return NRC_E_INVALID_ARG;
  bool table_grad = false;
  for (int l = 0; l < d.L; ++l) table_grad |= (d.lv[l].grad != nullptr);
  if (!table_grad) return NRC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (d.F) {
    case 1: return launch_bwd<1>(s, d, d_means, d_g_out, num_points, nullptr, true, warp_c);
    case 2: return launch_bwd<2>(s, d, d_means, d_g_out, num_points, nullptr, true, warp_c);
    case 4: return launch_bwd<4>(s, d, d_means, d_g_out, num_points, nullptr, true, warp_c);
    case 8: return launch_bwd<8>(s, d, d_means, d_g_out, num_points, nullptr, true, warp_c);
  }
  return NRC_E_UNSUPPORTED;
}

extern "C" int32_t nrc_abi_version(void) { return NRC_ABI_VERSION; }
#ifndef NRC_BUILD_DIGEST
#define NRC_BUILD_DIGEST "nrc-build-digest:unknown"
#endif
// the build stamps "nrc-build-digest:<sha256 of the sources>"; the prefix lets build.py find it in the file's bytes
extern "C" const char* nrc_build_digest(void) { return NRC_BUILD_DIGEST + sizeof("nrc-build-digest:") - 1; }

extern "C" const char* nrc_error_string(int32_t status) {
  switch (status) {
    case NRC_OK: return "ok";
    case NRC_E_INVALID_ARG: return "invalid argument";
    case NRC_E_UNSUPPORTED: return "unsupported configuration";
    case NRC_E_CUDA: return "CUDA launch failure (see nrc_last_cuda_error)";
  }
  return "unknown status";
}

extern "C" int32_t nrc_last_cuda_error(void) { return g_last_cuda_error; }

// ------------------------------------------------------------- contraction --
namespace nrc {
__global__ void contract_fwd_kernel(const float* __restrict__ x, int64_t P, float c,
                                    float* __restrict__ z) {
  int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float z0, z1, z2;
  contract_point(c, x[3 * p], x[3 * p + 1], x[3 * p + 2], z0, z1, z2);
  z[3 * p] = z0; z[3 * p + 1] = z1; z[3 * p + 2] = z2;
}
__global__ void contract_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gz,
                                    int64_t P, float c, float* __restrict__ gx) {
  int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float o0, o1, o2;
  contract_vjp(c, x[3 * p], x[3 * p + 1], x[3 * p + 2], gz[3 * p], gz[3 * p + 1], gz[3 * p + 2], o0,
               o1, o2);
  gx[3 * p] = o0; gx[3 * p + 1] = o1; gx[3 * p + 2] = o2;
}
}  // namespace nrc

extern "C" int32_t nrc_contract_fwd(void* stream, const float* d_x, int64_t num_points, float c,
                                    float* d_z) {
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_x || !d_z) return NRC_E_INVALID_ARG;
  unsigned grid = static_cast<unsigned>((num_points + 255) / 256);
  contract_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, num_points, c, d_z);
  return check_launch();
}

extern "C" int32_t nrc_contract_bwd(void* stream, const float* d_x, const float* d_g_z,
                                    int64_t num_points, float c, float* d_g_x) {
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_x || !d_g_z || !d_g_x) return NRC_E_INVALID_ARG;
  unsigned grid = static_cast<unsigned>((num_points + 255) / 256);
  contract_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_g_z, num_points, c,
                                                                          d_g_x);
  return check_launch();
}

static int32_t encode_tangent(void* stream, const nrc_encoding_t* enc, const float* d_means, const float* d_g,
                              const float* d_ge, int64_t num_points, float warp_c, float* d_edot) {
  EncDev d;
  const int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_means || !d_g || (!d_ge && !d_edot)) return NRC_E_INVALID_ARG;
  if (!d_edot)
    for (int l = 0; l < d.L; ++l)
      if (!d.lv[l].grad) return NRC_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (d.F) {
    case 1: return launch_tangent<1>(s, d, d_means, d_g, d_ge, num_points, warp_c, d_edot);
    case 2: return launch_tangent<2>(s, d, d_means, d_g, d_ge, num_points, warp_c, d_edot);
    case 4: return launch_tangent<4>(s, d, d_means, d_g, d_ge, num_points, warp_c, d_edot);
    case 8: return launch_tangent<8>(s, d, d_means, d_g, d_ge, num_points, warp_c, d_edot);
  }
  return NRC_E_UNSUPPORTED;
}

extern "C" int32_t nrc_encode_tangent_fwd(void* stream, const nrc_encoding_t* enc, const float* d_means, const float* d_g,
                                          int64_t num_points, float warp_c, float* d_edot) {
  if (!d_edot) return NRC_E_INVALID_ARG;
  return encode_tangent(stream, enc, d_means, d_g, nullptr, num_points, warp_c, d_edot);
}

extern "C" int32_t nrc_encode_tangent_bwd(void* stream, const nrc_encoding_t* enc, const float* d_means, const float* d_g,
                                          const float* d_ge, int64_t num_points, float warp_c) {
  if (!d_ge) return NRC_E_INVALID_ARG;
  return encode_tangent(stream, enc, d_means, d_g, d_ge, num_points, warp_c, nullptr);
}
