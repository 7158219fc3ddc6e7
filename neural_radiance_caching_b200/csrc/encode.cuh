// Device-side corner enumeration of one HashEncoding level, shared by the
// stand-alone encode kernels (encode.cu) and the fused query kernels (query.cu).
//
// Follows internal/grid_utils.py: hash levels :41-121, dense levels :352-445 via
// trilerp :679-726.  Index arithmetic uses individually rounded fp32 ops in the
// reference's order (no FMA contraction), so corner indices are bit-exact.
#pragma once
#include "nrc_common.cuh"

namespace nrc {

struct LevelDev {
  const float* table;
  float* grad;
  int32_t N;
  int32_t is_hash;
  uint32_t T;
  uint32_t pow2_mask;  // T-1 when T is a power of two, else 0
};

struct EncDev {
  int32_t L;
  int32_t F;
  float b0[3];
  float b1[3];
  float span[3];
  float scale;
  LevelDev lv[NRC_MAX_LEVELS];
};

inline int32_t make_enc_dev(const nrc_encoding_t* enc, EncDev& d) {
  if (!enc) return NRC_E_INVALID_ARG;
  if (enc->num_levels < 1 || enc->num_levels > NRC_MAX_LEVELS) return NRC_E_INVALID_ARG;
  int F = enc->num_features;
  if (!(F == 1 || F == 2 || F == 4 || F == 8)) return NRC_E_UNSUPPORTED;
  d.L = enc->num_levels;
  d.F = F;
  for (int a = 0; a < 3; ++a) {
    d.b0[a] = enc->bbox_min[a];
    d.b1[a] = enc->bbox_max[a];
    d.span[a] = enc->bbox_span[a];
    if (!(enc->bbox_span[a] > 0.f)) return NRC_E_INVALID_ARG;
  }
  d.scale = enc->precondition_scaling;
  for (int l = 0; l < d.L; ++l) {
    const nrc_level_t& s = enc->levels[l];
    if (!s.d_table || s.grid_size < 1 || s.table_size == 0) return NRC_E_INVALID_ARG;
    if (!s.is_hash && (uint64_t)s.grid_size * s.grid_size * s.grid_size != s.table_size)
      return NRC_E_INVALID_ARG;
    d.lv[l].table = s.d_table;
    d.lv[l].grad = s.d_grad;
    d.lv[l].N = s.grid_size;
    d.lv[l].is_hash = s.is_hash;
    d.lv[l].T = s.table_size;
    d.lv[l].pow2_mask = ((s.table_size & (s.table_size - 1)) == 0) ? s.table_size - 1 : 0u;
  }
  return NRC_OK;
}

constexpr uint32_t kPi2 = 19349663u;  // internal/grid_utils.py:102
constexpr uint32_t kPi3 = 83492791u;  // internal/grid_utils.py:103

// Per-level interpolation set-up for one point.
struct Corners {
  int32_t fl[3];   // floor index per original axis (x,y,z); dense: padded-grid index
  float cw[3];     // ceil weights  (loc - floor)
  float fw[3];     // floor weights (1 - cw)
};

// xn: x mapped to [0,1]^3 (already (x-b0)/span).  HashEncoding.__call__ :863 multiplies
// by N; trilerp / hash_resample subtract 0.5; the dense path then adds 1.0 (:390).
__device__ __forceinline__ Corners level_setup(const LevelDev& lv, const float xn[3]) {
  Corners c;
  const float fN = static_cast<float>(lv.N);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float pos = __fmul_rn(xn[a], fN);
    float loc = __fsub_rn(pos, 0.5f);
    if (!lv.is_hash) loc = __fadd_rn(loc, 1.0f);
    float f = floorf(loc);
    c.cw[a] = __fsub_rn(loc, f);
    c.fw[a] = __fsub_rn(1.0f, c.cw[a]);
    c.fl[a] = __float2int_rz(f);  // saturating f32->s32, like XLA's convert
  }
  return c;
}

// Row index of corner (bx,by,bz) in {0,1}^3, or -1 for a zero-padding voxel.
__device__ __forceinline__ int32_t corner_row(const LevelDev& lv, const Corners& c, int bx, int by,
                                              int bz) {
  if (lv.is_hash) {
    // int32 -> uint32 wrap-around (:98-101), uint32 multiplies wrap (:105-108).
    uint32_t ux = static_cast<uint32_t>(c.fl[0] + bx);
    uint32_t uy = static_cast<uint32_t>(c.fl[1] + by);
    uint32_t uz = static_cast<uint32_t>(c.fl[2] + bz);
    uint32_t h = ux ^ ((uy * kPi2) ^ (uz * kPi3));
    return static_cast<int32_t>(lv.pow2_mask ? (h & lv.pow2_mask) : (h % lv.T));
  }
  const int N = lv.N;
  int ix = min(max(c.fl[0] + bx, 0), N + 1);  // clamp into the padded volume (:437-438)
  int iy = min(max(c.fl[1] + by, 0), N + 1);
  int iz = min(max(c.fl[2] + bz, 0), N + 1);
  if (ix == 0 || iy == 0 || iz == 0 || ix == N + 1 || iy == N + 1 || iz == N + 1) return -1;
  return ((ix - 1) * N + (iy - 1)) * N + (iz - 1);
}

// Flat index into the padded (N+2)^3 volume (parity aid for dense levels).
__device__ __forceinline__ int32_t corner_padded_index(const LevelDev& lv, const Corners& c, int bx,
                                                       int by, int bz) {
  const int N = lv.N;
  int ix = min(max(c.fl[0] + bx, 0), N + 1);
  int iy = min(max(c.fl[1] + by, 0), N + 1);
  int iz = min(max(c.fl[2] + bz, 0), N + 1);
  return (ix * (N + 2) + iy) * (N + 2) + iz;
}

// Reference corner order k = 0..7 and weight product order.
//   hash : x outermost, z innermost; w = (wx*wy)*wz          (:68-89)
//   dense: operates on flipped coords -> z outermost, x innermost; w = (wz*wy)*wx
__device__ __forceinline__ void corner_bits(int is_hash, int k, int& bx, int& by, int& bz) {
  int hi = (k >> 2) & 1, mid = (k >> 1) & 1, lo = k & 1;
  if (is_hash) { bx = hi; by = mid; bz = lo; }
  else         { bz = hi; by = mid; bx = lo; }
}

__device__ __forceinline__ float corner_weight(int is_hash, const Corners& c, int bx, int by,
                                               int bz) {
  float wx = bx ? c.cw[0] : c.fw[0];
  float wy = by ? c.cw[1] : c.fw[1];
  float wz = bz ? c.cw[2] : c.fw[2];
  return is_hash ? __fmul_rn(__fmul_rn(wx, wy), wz) : __fmul_rn(__fmul_rn(wz, wy), wx);
}

template <int F>
struct FeatVec { float v[F]; };

template <int F>
__device__ __forceinline__ FeatVec<F> load_row(const float* __restrict__ table, int32_t row) {
  FeatVec<F> r;
  if constexpr (F == 1) {
    r.v[0] = __ldg(table + row);
  } else if constexpr (F == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(table) + row);
    r.v[0] = t.x; r.v[1] = t.y;
  } else if constexpr (F == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(table) + row);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    static_assert(F == 8, "F in {1,2,4,8}");
    float4 t0 = __ldg(reinterpret_cast<const float4*>(table) + 2 * row);
    float4 t1 = __ldg(reinterpret_cast<const float4*>(table) + 2 * row + 1);
    r.v[0] = t0.x; r.v[1] = t0.y; r.v[2] = t0.z; r.v[3] = t0.w;
    r.v[4] = t1.x; r.v[5] = t1.y; r.v[6] = t1.z; r.v[7] = t1.w;
  }
  return r;
}

// Interpolated features of one level (before precondition scaling); products and
// sums individually rounded in the reference's corner order so that the result
// is bit-identical to the fp32 oracle.
template <int F>
__device__ __forceinline__ FeatVec<F> level_interp(const LevelDev& lv, const Corners& c) {
  // Per-axis factorisation of the eight corners: two indices / hash terms / validity flags / weights per axis,
  // combined per corner with two integer ops and one multiply.  The weight products keep the reference's
  // association ((wx*wy)*wz for hash levels, (wz*wy)*wx for dense levels) and the corner order, so the result
  // is bit-identical to the per-corner evaluation (corner_row / corner_weight above).
  int32_t rows[8];
  float w[8];
  const float wx[2] = {c.fw[0], c.cw[0]}, wy[2] = {c.fw[1], c.cw[1]}, wz[2] = {c.fw[2], c.cw[2]};
  if (lv.is_hash) {
    const uint32_t hx[2] = {static_cast<uint32_t>(c.fl[0]), static_cast<uint32_t>(c.fl[0] + 1)};
    const uint32_t hy[2] = {static_cast<uint32_t>(c.fl[1]) * kPi2, static_cast<uint32_t>(c.fl[1] + 1) * kPi2};
    const uint32_t hz[2] = {static_cast<uint32_t>(c.fl[2]) * kPi3, static_cast<uint32_t>(c.fl[2] + 1) * kPi3};
    float wxy[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) wxy[i][j] = __fmul_rn(wx[i], wy[j]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int bx = (k >> 2) & 1, by = (k >> 1) & 1, bz = k & 1;
      const uint32_t h = hx[bx] ^ (hy[by] ^ hz[bz]);
      rows[k] = static_cast<int32_t>(lv.pow2_mask ? (h & lv.pow2_mask) : (h % lv.T));
      w[k] = __fmul_rn(wxy[bx][by], wz[bz]);
    }
  } else {
    const int N = lv.N;
    int ox[2], oy[2], oz[2];
    bool vx[2], vy[2], vz[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ix = min(max(c.fl[0] + b, 0), N + 1), iy = min(max(c.fl[1] + b, 0), N + 1),
                iz = min(max(c.fl[2] + b, 0), N + 1);
      vx[b] = ix >= 1 && ix <= N; vy[b] = iy >= 1 && iy <= N; vz[b] = iz >= 1 && iz <= N;
      ox[b] = (ix - 1) * N * N; oy[b] = (iy - 1) * N; oz[b] = iz - 1;
    }
    float wzy[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) wzy[i][j] = __fmul_rn(wz[i], wy[j]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int bz = (k >> 2) & 1, by = (k >> 1) & 1, bx = k & 1;
      rows[k] = (vx[bx] && vy[by] && vz[bz]) ? ox[bx] + oy[by] + oz[bz] : -1;
      w[k] = __fmul_rn(wzy[bz][by], wx[bx]);
    }
  }
  FeatVec<F> vals[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {  // issue all gathers before consuming any
    if (rows[k] >= 0) vals[k] = load_row<F>(lv.table, rows[k]);
    else {
#pragma unroll
      for (int f = 0; f < F; ++f) vals[k].v[f] = 0.f;
    }
  }
  FeatVec<F> acc;
#pragma unroll
  for (int f = 0; f < F; ++f) acc.v[f] = __fmul_rn(vals[0].v[f], w[0]);
#pragma unroll
  for (int k = 1; k < 8; ++k)
#pragma unroll
    for (int f = 0; f < F; ++f) acc.v[f] = __fadd_rn(acc.v[f], __fmul_rn(vals[k].v[f], w[k]));
  return acc;
}

// G consecutive HASH levels at once, branch-free: every gather of the group (8*G rows) is issued before the
// first one is consumed, so a lane keeps 16-32 independent L2 requests in flight instead of 8 (the fused
// query kernels are latency bound on these gathers).  Arithmetic and its order are those of level_interp.
// Levels beyond enc.L repeat the last level (results discarded by the caller).
template <int F, int G>
__device__ __forceinline__ void hash_interp_group(const EncDev& enc, int l0, const float xn[3], FeatVec<F> (&out)[G]) {
  uint32_t rows[G][8];
  float w[G][8];
  const float* tables[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const LevelDev& lv = enc.lv[min(l0 + g, enc.L - 1)];
    tables[g] = lv.table;
    const float fN = static_cast<float>(lv.N);
    uint32_t u[3][2];
    float wt[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float loc = __fsub_rn(__fmul_rn(xn[a], fN), 0.5f);
      const float f = floorf(loc);
      wt[a][1] = __fsub_rn(loc, f);
      wt[a][0] = __fsub_rn(1.0f, wt[a][1]);
      const int32_t fi = __float2int_rz(f);
      u[a][0] = static_cast<uint32_t>(fi);
      u[a][1] = static_cast<uint32_t>(fi + 1);
    }
    const uint32_t hy[2] = {u[1][0] * kPi2, u[1][1] * kPi2};
    const uint32_t hz[2] = {u[2][0] * kPi3, u[2][1] * kPi3};
#pragma unroll
    for (int k = 0; k < 8; ++k) {   // hash corner order: x outermost, z innermost; w = (wx*wy)*wz
      const int bx = (k >> 2) & 1, by = (k >> 1) & 1, bz = k & 1;
      const uint32_t h = u[0][bx] ^ (hy[by] ^ hz[bz]);
      rows[g][k] = lv.pow2_mask ? (h & lv.pow2_mask) : (h % lv.T);
      w[g][k] = __fmul_rn(__fmul_rn(wt[0][bx], wt[1][by]), wt[2][bz]);
    }
  }
  FeatVec<F> vals[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int k = 0; k < 8; ++k) vals[g][k] = load_row<F>(tables[g], static_cast<int32_t>(rows[g][k]));
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int f = 0; f < F; ++f) out[g].v[f] = __fmul_rn(vals[g][0].v[f], w[g][0]);
#pragma unroll
    for (int k = 1; k < 8; ++k)
#pragma unroll
      for (int f = 0; f < F; ++f) out[g].v[f] = __fadd_rn(out[g].v[f], __fmul_rn(vals[g][k].v[f], w[g][k]));
  }
}

// Lane-pair gather.  Two adjacent lanes (side = lane & 1) share ONE point and split its eight corners along
// the axis that is outermost in the reference's summation order: x for hash levels, z for dense levels.
// That axis is also the one whose two corners sit next to each other in memory - x and x+1 differ in their
// trailing bits only, so the two hash rows h ^ x, h ^ (x+1) fall into the same aligned 2^(k+1)-row block
// (k = trailing ones of x; same 128-byte line 97 % of the time at F = 1, 87.5 % at F = 4); z is the
// contiguous axis of the dense [N,N,N,F] layout.  One warp-wide load instruction then touches ~16 lines
// instead of ~32: the fused query is bound by L1TEX wavefronts (one per distinct line per instruction,
// ~2 cycles each), not by bytes.  H points per lane are processed together so that 4*H gathers are in
// flight per lane.  The even lane sums its four products in the reference's order and hands the partial
// sum to the odd lane, which continues the same left-to-right sum: bit-identical to level_interp.
// All 32 lanes must call this (full-mask shuffle); the result is valid on ODD lanes.
// kExact = false (bf16-MLP variant, whose features are rounded to 8 mantissa bits right after): the weighted
// sum is contracted into FMAs (5 instead of 11 FP instructions per feature); corner INDICES and weights are
// computed exactly as above in both modes.
template <int F, int H, bool kExact>
__device__ __forceinline__ void level_interp_pair(const LevelDev& lv, const float (&xn)[H][3], const int side,
                                                  FeatVec<F> (&out)[H]) {
  int32_t rows[H][4];
  float w[H][4];
  const float fN = static_cast<float>(lv.N);
#pragma unroll
  for (int h = 0; h < H; ++h) {
    int32_t fl[3];
    float wt[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float loc = __fsub_rn(__fmul_rn(xn[h][a], fN), 0.5f);
      if (!lv.is_hash) loc = __fadd_rn(loc, 1.0f);
      const float f = floorf(loc);
      wt[a][1] = __fsub_rn(loc, f);
      wt[a][0] = __fsub_rn(1.0f, wt[a][1]);
      fl[a] = __float2int_rz(f);
    }
    if (lv.is_hash) {
      const uint32_t hx = static_cast<uint32_t>(fl[0] + side);
      const uint32_t hy[2] = {static_cast<uint32_t>(fl[1]) * kPi2, static_cast<uint32_t>(fl[1] + 1) * kPi2};
      const uint32_t hz[2] = {static_cast<uint32_t>(fl[2]) * kPi3, static_cast<uint32_t>(fl[2] + 1) * kPi3};
      const float wo = side ? wt[0][1] : wt[0][0];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int by = j >> 1, bz = j & 1;
        const uint32_t hh = hx ^ (hy[by] ^ hz[bz]);
        rows[h][j] = static_cast<int32_t>(lv.pow2_mask ? (hh & lv.pow2_mask) : (hh % lv.T));
        w[h][j] = __fmul_rn(__fmul_rn(wo, wt[1][by]), wt[2][bz]);
      }
    } else {
      const int N = lv.N;
      const int iz = min(max(fl[2] + side, 0), N + 1);
      const bool vz = iz >= 1 && iz <= N;
      int ox[2], oy[2];
      bool vx[2], vy[2];
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ix = min(max(fl[0] + b, 0), N + 1), iy = min(max(fl[1] + b, 0), N + 1);
        vx[b] = ix >= 1 && ix <= N; vy[b] = iy >= 1 && iy <= N;
        ox[b] = (ix - 1) * N * N; oy[b] = (iy - 1) * N;
      }
      const float wo = side ? wt[2][1] : wt[2][0];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int by = j >> 1, bx = j & 1;
        rows[h][j] = (vz && vy[by] && vx[bx]) ? ox[bx] + oy[by] + (iz - 1) : -1;
        w[h][j] = __fmul_rn(__fmul_rn(wo, wt[1][by]), wt[0][bx]);
      }
    }
  }
  FeatVec<F> vals[H][4];
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (rows[h][j] >= 0) vals[h][j] = load_row<F>(lv.table, rows[h][j]);
      else {
#pragma unroll
        for (int f = 0; f < F; ++f) vals[h][j].v[f] = 0.f;
      }
    }
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int f = 0; f < F; ++f) {
      if constexpr (kExact) {
        float p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = __fmul_rn(vals[h][j].v[f], w[h][j]);
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[2]), p[3]);
        const float lo = __shfl_xor_sync(0xffffffffu, s, 1);
        out[h].v[f] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(lo, p[0]), p[1]), p[2]), p[3]);
      } else {
        float s = vals[h][0].v[f] * w[h][0];
#pragma unroll
        for (int j = 1; j < 4; ++j) s = fmaf(vals[h][j].v[f], w[h][j], s);
        out[h].v[f] = s + __shfl_xor_sync(0xffffffffu, s, 1);
      }
    }
}

// ---- table-gradient scatter helpers (encode.cu, normals2.cu) ----
template <int F>
__device__ __forceinline__ void atomic_add_row(float* grad, int32_t row, const float (&g)[F]) {
  if constexpr (F == 4) {
    atomicAdd(reinterpret_cast<float4*>(grad) + row, make_float4(g[0], g[1], g[2], g[3]));
  } else if constexpr (F == 2) {
    atomicAdd(reinterpret_cast<float2*>(grad) + row, make_float2(g[0], g[1]));
  } else if constexpr (F == 8) {
    atomicAdd(reinterpret_cast<float4*>(grad) + 2 * row, make_float4(g[0], g[1], g[2], g[3]));
    atomicAdd(reinterpret_cast<float4*>(grad) + 2 * row + 1, make_float4(g[4], g[5], g[6], g[7]));
  } else {
    atomicAdd(grad + row, g[0]);
  }
}

// Contiguous runs of lanes that hit the SAME row are summed into the run's first lane, which issues one
// atomic for the run.  Consecutive lanes are consecutive samples of a ray, and on the coarse dense levels
// (16^3, 32^3, 64^3) whole stretches of a ray fall into one cell: without this the cells around the scene
// centre receive thousands of same-address atomics per step, which the L2 serialises.  Segmented reduction
// by shuffles: after the step with distance d lane i holds the sum of lanes [i, i+2d) of its run.
template <int F>
__device__ __forceinline__ void warp_run_atomic_add(float* grad, int32_t row, float (&gw)[F], int lane) {
  // a run starts wherever the row differs from the previous lane's; lanes i and i+d belong to the same run
  // iff no run starts in (i, i+d]  (equal rows in DIFFERENT runs stay separate: each run issues its own atomic)
  const int32_t rp = __shfl_up_sync(0xffffffffu, row, 1);
  const bool head = lane == 0 || rp != row;
  const uint32_t heads = __ballot_sync(0xffffffffu, head);
  const uint32_t after = lane == 31 ? 0xffffffffu : (heads >> (lane + 1));   // bit j: a run starts at lane+1+j
  // longest run in the warp (warp-uniform, from the ballot): only ceil(log2(longest)) shuffle rounds are needed, and none
  // when every lane heads its own run (the usual case on fine hash levels).  The fixed five rounds were 30 % of the
  // F = 1 scatter's warp samples (profiles/r02i).
  int longest = 1;
  for (uint32_t t = ~heads; t; t &= t << 1) ++longest;
  for (int d = 1; d < longest; d <<= 1) {
    float o[F];
#pragma unroll
    for (int f = 0; f < F; ++f) o[f] = __shfl_down_sync(0xffffffffu, gw[f], d);
    if (lane + d < 32 && (after & ((1u << d) - 1u)) == 0u) {
#pragma unroll
      for (int f = 0; f < F; ++f) gw[f] += o[f];
    }
  }
  if (head && row >= 0) atomic_add_row<F>(grad, row, gw);
}

__device__ __forceinline__ void normalise_point(const EncDev& enc, const float x[3], float xn[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) xn[a] = __fdiv_rn(__fsub_rn(x[a], enc.b0[a]), enc.span[a]);
}

}  // namespace nrc
