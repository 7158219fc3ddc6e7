// K3 (tensor-core variant): the fused density MLP with bf16 operands / fp32 accumulation on
// mma.sync m16n8k16, layers chained in registers (see mma_bf16.cuh), optionally fused with the
// contraction + hash-grid gather in front (features never reach HBM) and the analytic-normal
// back-propagation behind.  One warp owns 32 points (two 16-row MMA tiles); a CTA is four
// independent warps, so there is no block-level barrier on the forward path.
//
// Reference: internal/geometry.py:155-168,199-341,442-460 (same maths as mlp.cu / query.cu;
// parity bar for this variant: rel 2e-2).
#include <cstddef>
#include <cstdlib>

#include "encode.cuh"
#include "mma_bf16.cuh"

namespace nrc {

constexpr int kBfWarps = 4;
constexpr int kBfThreads = kBfWarps * 32;
constexpr int kGStride = 33;  // fp32 row stride of the per-warp g_enc scratch

struct WarpScratch {
  __nv_bfloat16 x[32][kXStride];  // encoded features of the warp's 32 points (bf16, zero padded)
  int inside[32];                 // bbox mask per point
};
struct WarpGradScratch {
  float g[32][kGStride];          // d raw / d enc (normals path)
};

// Dynamic shared memory of the forward kernel.  Everything from `wg` on is only needed when the analytic
// normals (raw_grad) are requested; a launch without them allocates offsetof(FwdSmemBf16, wg) bytes
// (27 KB instead of 45 KB: more of the 256 KB L1/shared array left to L1; four CTAs per SM either way,
// bounded by registers).
struct FwdSmemBf16 {
  MlpWeightsFwdBf16 w;
  WarpScratch ws[kBfWarps];
  MlpWeightsGradBf16 wg;
  WarpGradScratch gs[kBfWarps];
};

struct QueryOut {
  float* density; float* raw; float* feat; float* grad_pred; float* raw_grad; float* enc_out;
};

// Two 16-row tiles through one 64-wide layer with every B fragment loaded ONCE (half the ldmatrix wavefronts
// of two mma_layer64 calls: shared-memory wavefronts were half of the kernel's L1TEX traffic).
template <int KS>
__device__ __forceinline__ void mma_layer64_x2(float (&acc)[2][8][4], const uint32_t (&a)[2][KS][4],
                                               const __nv_bfloat16* wt, int stride, const float* bias, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float2 bv = *reinterpret_cast<const float2*>(bias + nt * 8 + (lane & 3) * 2);
#pragma unroll
    for (int t = 0; t < 2; ++t) { acc[t][nt][0] = bv.x; acc[t][nt][1] = bv.y; acc[t][nt][2] = bv.x; acc[t][nt][3] = bv.y; }
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_b_frag2(b, wt, stride, np * 16, ks * 16, lane);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mma_bf16(acc[t][2 * np], a[t][ks], b[0], b[1]);
        mma_bf16(acc[t][2 * np + 1], a[t][ks], b[2], b[3]);
      }
    }
  }
}

// Head outputs of one 16-point tile (accumulator layout) -> global memory.
__device__ __forceinline__ void store_tile_outputs(const QueryOut& out, const float (&o)[4], const float (&acc)[8][4],
                                                   const int* inside, int64_t base, int mt, int64_t P,
                                                   float density_bias, int lane) {
  const int r = lane >> 2;
  const int64_t pr[2] = {base + mt * 16 + r, base + mt * 16 + r + 8};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (pr[h] >= P) continue;
    if ((lane & 3) == 0) {
      float rawv = o[2 * h];
      if (out.raw) out.raw[pr[h]] = rawv;
      if (out.density) out.density[pr[h]] = inside[mt * 16 + r + 8 * h] ? safe_exp(rawv + density_bias) : 0.f;
      if (out.grad_pred) out.grad_pred[3 * pr[h]] = o[2 * h + 1];
    } else if ((lane & 3) == 1 && out.grad_pred) {
      out.grad_pred[3 * pr[h] + 1] = o[2 * h];
      out.grad_pred[3 * pr[h] + 2] = o[2 * h + 1];
    }
    if (out.feat) {
      float* f = out.feat + pr[h] * kHid + (lane & 3) * 2;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<float2*>(f + nt * 8) = make_float2(acc[nt][2 * h], acc[nt][2 * h + 1]);
    }
  }
}

// kFused: `in` = means [P,3], features are gathered here; else `in` = enc [P,in_dim].
template <int F, int KS0, bool kFused, int kMinCtas>
__global__ void __launch_bounds__(kBfThreads, kMinCtas)
mlp_bf16_fwd_kernel(const __grid_constant__ EncDev enc, const nrc_density_mlp_t m,
                    const float* __restrict__ in, int64_t P, float warp_c, float density_bias,
                    const QueryOut out, const int g_group, const int ppw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmemBf16& s = *reinterpret_cast<FwdSmemBf16*>(smem_raw);
  const bool want_grad = kFused && out.raw_grad != nullptr;   // the launcher sizes the allocation accordingly
  load_weights_bf16<kBfThreads>(s.w, want_grad ? &s.wg : nullptr, m);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpScratch& ws = s.ws[warp];
  WarpGradScratch& gs = s.gs[warp];
  const int in_dim = m.in_dim;
  // ppw = points per warp: 32 (two 16-row MMA tiles), or 16 for launches too small to fill the machine with
  // 128-point CTAs (32 768 points = 256 CTAs on 148 SMs): twice the CTAs, half the serial work per warp
  const int64_t pts_per_cta = static_cast<int64_t>(kBfWarps) * ppw;
  const int64_t num_tiles = (P + pts_per_cta - 1) / pts_per_cta;
  if constexpr (kFused) {
    uint32_t* xz = reinterpret_cast<uint32_t*>(&ws.x[0][0]);
    for (int i = lane; i < 32 * kXStride / 2; i += 32) xz[i] = 0u;
    __syncwarp();
  }
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t base = tile * pts_per_cta + warp * ppw;
    if (base >= P) continue;  // warp-uniform
    const int64_t p = base + lane;
    const bool valid = lane < ppw && p < P;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, xn[3] = {0.f, 0.f, 0.f};
    // ---------------- front end: one point per lane -> bf16 feature row -------------------
    if constexpr (kFused) {
      int inside = 0;
      if (valid) {
        x0 = __ldg(in + 3 * p); x1 = __ldg(in + 3 * p + 1); x2 = __ldg(in + 3 * p + 2);
        float z[3];
        contract_point(warp_c, x0, x1, x2, z[0], z[1], z[2]);
        normalise_point(enc, z, xn);
        inside = 1;
#pragma unroll
        for (int a = 0; a < 3; ++a) inside = inside && (z[a] > enc.b0[a]) && (z[a] < enc.b1[a]);
      }
      ws.inside[lane] = inside;
      if ((g_group & 2) && ppw == 16) {
        // Lane-pair gather, one 16-point tile per warp
        const int side = lane & 1, q = lane >> 1;
        float xq[1][3];
#pragma unroll
        for (int a = 0; a < 3; ++a) xq[0][a] = __shfl_sync(0xffffffffu, xn[a], q);
        for (int l = 0; l < enc.L; ++l) {
          FeatVec<F> v[1];
          level_interp_pair<F, 1, false>(enc.lv[l], xq, side, v);
          if (side && base + q < P) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
              float e = __fmul_rn(v[0].v[f], enc.scale);
              ws.x[q][l * F + f] = __float2bfloat16(e);
              if (out.enc_out) out.enc_out[(base + q) * in_dim + l * F + f] = e;
            }
          }
        }
        __syncwarp();
      } else if (g_group & 2) {
        // Lane-pair gather (encode.cuh: level_interp_pair): lanes 2i, 2i+1 share point i of each 16-point half.
        const int side = lane & 1;
        float xq[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int a = 0; a < 3; ++a) xq[h][a] = __shfl_sync(0xffffffffu, xn[a], h * 16 + (lane >> 1));
        for (int l = 0; l < enc.L; ++l) {
          FeatVec<F> v[2];
          level_interp_pair<F, 2, false>(enc.lv[l], xq, side, v);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = h * 16 + (lane >> 1);
            if (side && base + q < P) {
#pragma unroll
              for (int f = 0; f < F; ++f) {
                float e = __fmul_rn(v[h].v[f], enc.scale);
                ws.x[q][l * F + f] = __float2bfloat16(e);
                if (out.enc_out) out.enc_out[(base + q) * in_dim + l * F + f] = e;
              }
            }
          }
        }
        __syncwarp();
      } else if (valid) {
        auto emit = [&](int l, const FeatVec<F>& v) {
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float e = __fmul_rn(v.v[f], enc.scale);
            ws.x[lane][l * F + f] = __float2bfloat16(e);
            if (out.enc_out) out.enc_out[p * in_dim + l * F + f] = e;
          }
        };
        // dense levels come first in the level schedule (N^3 <= T), hash levels after them
        int l = 0;
        for (; l < enc.L && !enc.lv[l].is_hash; ++l) {
          Corners c = level_setup(enc.lv[l], xn);
          emit(l, level_interp<F>(enc.lv[l], c));
        }
        constexpr int G = 2;   // hash levels gathered two at a time (16 rows in flight per lane)
        for (; l < enc.L; l += G) {
          bool all_hash = true;
          for (int g = 0; g < G && l + g < enc.L; ++g) all_hash = all_hash && enc.lv[l + g].is_hash;
          if (all_hash && (g_group & 1)) {
            FeatVec<F> v[G];
            hash_interp_group<F, G>(enc, l, xn, v);
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (l + g < enc.L) emit(l + g, v[g]);
          } else {
            for (int g = 0; g < G && l + g < enc.L; ++g) {
              Corners c = level_setup(enc.lv[l + g], xn);
              emit(l + g, level_interp<F>(enc.lv[l + g], c));
            }
          }
        }
      }
      // padding columns [in_dim, 16*KS0) were zeroed once before the tile loop and are never written again
      if (!valid && lane < ppw)
        for (int k = 0; k < in_dim; ++k) ws.x[lane][k] = __float2bfloat16(0.f);
    } else {
      for (int k = 0; k < KS0 * 16; ++k)
        ws.x[lane][k] = __float2bfloat16((valid && k < in_dim) ? __ldg(in + p * in_dim + k) : 0.f);
      ws.inside[lane] = 1;
    }
    __syncwarp();
    // ---------------- tensor-core MLP: two 16-point tiles per warp ------------------------
    if (!want_grad && (g_group & 4) && ppw == 32) {
      // forward only: both tiles together, B fragments shared
      uint32_t af[2][KS0][4];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int ks = 0; ks < KS0; ++ks) load_a_frag(af[t][ks], &ws.x[0][0], kXStride, t * 16, ks * 16, lane);
      float acc[2][8][4];
      mma_layer64_x2<KS0>(acc, af, &s.w.w0t[0][0], kXStride, s.w.b0, lane);
      uint32_t hf[2][4][4];
      acc_to_afrag<true>(acc[0], hf[0]);
      acc_to_afrag<true>(acc[1], hf[1]);
      mma_layer64_x2<4>(acc, hf, &s.w.w1t[0][0], kWStride, s.w.b1, lane);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[t][nt][e] = fmaxf(acc[t][nt][e], 0.f);
        acc_to_afrag<false>(acc[t], hf[t]);
      }
      float o[2][4];
      {
        const int c = (lane & 3) * 2;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          o[t][0] = o[t][2] = c < 4 ? s.w.bo[c] : 0.f;
          o[t][1] = o[t][3] = c + 1 < 4 ? s.w.bo[c + 1] : 0.f;
        }
        uint32_t b[4];
        load_b_frag_k32(b, &s.w.wot[0][0], kWStride, 0, 0, lane);
#pragma unroll
        for (int t = 0; t < 2; ++t) { mma_bf16(o[t], hf[t][0], b[0], b[1]); mma_bf16(o[t], hf[t][1], b[2], b[3]); }
        load_b_frag_k32(b, &s.w.wot[0][0], kWStride, 0, 32, lane);
#pragma unroll
        for (int t = 0; t < 2; ++t) { mma_bf16(o[t], hf[t][2], b[0], b[1]); mma_bf16(o[t], hf[t][3], b[2], b[3]); }
      }
      store_tile_outputs(out, o[0], acc[0], ws.inside, base, 0, P, density_bias, lane);
      store_tile_outputs(out, o[1], acc[1], ws.inside, base, 1, P, density_bias, lane);
      __syncwarp();
      continue;
    }
#pragma unroll 1
    for (int mt = 0; mt < (ppw >> 4); ++mt) {
      uint32_t a0[KS0][4];
#pragma unroll
      for (int ks = 0; ks < KS0; ++ks) load_a_frag(a0[ks], &ws.x[0][0], kXStride, mt * 16, ks * 16, lane);
      float acc[8][4];
      mma_layer64<KS0>(acc, a0, &s.w.w0t[0][0], kXStride, s.w.b0, lane);
      uint32_t h1f[4][4];
      acc_to_afrag<true>(acc, h1f);
      mma_layer64<4>(acc, h1f, &s.w.w1t[0][0], kWStride, s.w.b1, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = fmaxf(acc[nt][e], 0.f);
      uint32_t h2f[4][4];
      acc_to_afrag<false>(acc, h2f);
      // heads (density + 3 pred-normal channels), N = 8 tile
      float o[4];
      {
        const int c = (lane & 3) * 2;
        o[0] = o[2] = c < 4 ? s.w.bo[c] : 0.f;
        o[1] = o[3] = c + 1 < 4 ? s.w.bo[c + 1] : 0.f;
        uint32_t b[4];
        load_b_frag_k32(b, &s.w.wot[0][0], kWStride, 0, 0, lane);
        mma_bf16(o, h2f[0], b[0], b[1]);
        mma_bf16(o, h2f[1], b[2], b[3]);
        load_b_frag_k32(b, &s.w.wot[0][0], kWStride, 0, 32, lane);
        mma_bf16(o, h2f[2], b[0], b[1]);
        mma_bf16(o, h2f[3], b[2], b[3]);
      }
      store_tile_outputs(out, o, acc, ws.inside, base, mt, P, density_bias, lane);
      const int r = lane >> 2;
      if constexpr (kFused) {
        if (out.raw_grad) {
          // g_h2 = wd * [h2 > 0]
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + (lane & 3) * 2;
            const float w0 = s.wg.wo[c][0], w1 = s.wg.wo[c + 1][0];
            acc[nt][0] = acc[nt][0] > 0.f ? w0 : 0.f;
            acc[nt][1] = acc[nt][1] > 0.f ? w1 : 0.f;
            acc[nt][2] = acc[nt][2] > 0.f ? w0 : 0.f;
            acc[nt][3] = acc[nt][3] > 0.f ? w1 : 0.f;
          }
          uint32_t gf[4][4];
          acc_to_afrag<false>(acc, gf);
          // g_h1 = (g_h2 W1^T) * [h1 > 0]
          mma_layer64_t<4>(acc, gf, &s.w.w1t[0][0], kWStride, lane);
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            float2 lo = unpack_bf16(h1f[nt >> 1][2 * (nt & 1)]);
            float2 hi = unpack_bf16(h1f[nt >> 1][2 * (nt & 1) + 1]);
            acc[nt][0] = lo.x > 0.f ? acc[nt][0] : 0.f;
            acc[nt][1] = lo.y > 0.f ? acc[nt][1] : 0.f;
            acc[nt][2] = hi.x > 0.f ? acc[nt][2] : 0.f;
            acc[nt][3] = hi.y > 0.f ? acc[nt][3] : 0.f;
          }
          acc_to_afrag<false>(acc, gf);
          // g_enc = g_h1 W0^T  (N = 16 * KS0 columns, of which in_dim are real)
          float ge[2 * KS0][4];
#pragma unroll
          for (int nt = 0; nt < 2 * KS0; ++nt) ge[nt][0] = ge[nt][1] = ge[nt][2] = ge[nt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < KS0; ++np) {
              uint32_t b[4];
              load_b_frag2_trans(b, &s.w.w0t[0][0], kXStride, ks * 16, np * 16, lane);
              mma_bf16(ge[2 * np], gf[ks], b[0], b[1]);
              mma_bf16(ge[2 * np + 1], gf[ks], b[2], b[3]);
            }
          }
#pragma unroll
          for (int nt = 0; nt < 2 * KS0; ++nt) {
            const int c = nt * 8 + (lane & 3) * 2;
            if (c < in_dim) { gs.g[mt * 16 + r][c] = ge[nt][0]; gs.g[mt * 16 + r + 8][c] = ge[nt][2]; }
            if (c + 1 < in_dim) { gs.g[mt * 16 + r][c + 1] = ge[nt][1]; gs.g[mt * 16 + r + 8][c + 1] = ge[nt][3]; }
          }
        }
      }
    }
    __syncwarp();
    if constexpr (kFused) {
      if (out.raw_grad && valid) {
        // VJP through the encoding (corner re-gather hits L1/L2) and the contraction.
        float gz[3] = {0.f, 0.f, 0.f};
        for (int l = 0; l < enc.L; ++l) {
          const LevelDev& lv = enc.lv[l];
          Corners c = level_setup(lv, xn);
          float g[F];
#pragma unroll
          for (int f = 0; f < F; ++f) g[f] = gs.g[lane][l * F + f] * enc.scale;
          float gl[3] = {0.f, 0.f, 0.f};
          // all eight corner rows are requested before the first one is used (one exposed latency per level, not eight)
          int32_t rows8[8];
          FeatVec<F> v8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            int bx, by, bz;
            corner_bits(lv.is_hash, k, bx, by, bz);
            rows8[k] = corner_row(lv, c, bx, by, bz);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (rows8[k] >= 0) {
              v8[k] = load_row<F>(lv.table, rows8[k]);
            } else {
#pragma unroll
              for (int f = 0; f < F; ++f) v8[k].v[f] = 0.f;
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (rows8[k] < 0) continue;
            int bx, by, bz;
            corner_bits(lv.is_hash, k, bx, by, bz);
            float dot = 0.f;
#pragma unroll
            for (int f = 0; f < F; ++f) dot = fmaf(g[f], v8[k].v[f], dot);
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
        out.raw_grad[3 * p] = o0; out.raw_grad[3 * p + 1] = o1; out.raw_grad[3 * p + 2] = o2;
      }
    }
    __syncwarp();
  }
}

template <int F, int KS0, bool kFused, int kMinCtas>
int32_t launch_bf16_fwd_occ(cudaStream_t st, const EncDev& d, const nrc_density_mlp_t* mlp, const float* in,
                            int64_t P, float warp_c, float bias, const QueryOut& out) {
  if (const int32_t st_attr = ensure_dynamic_smem<mlp_bf16_fwd_kernel<F, KS0, kFused, kMinCtas>>(static_cast<int>(sizeof(FwdSmemBf16))); st_attr != NRC_OK) return st_attr;
  static const int ppw_env = getenv("NRC_QUERY_PPW") ? atoi(getenv("NRC_QUERY_PPW")) : 0;
  static const int mult_env = getenv("NRC_QUERY_GRID_MULT") ? atoi(getenv("NRC_QUERY_GRID_MULT")) : 0;
  // bit 0: hash levels gathered two at a time; bit 1: lane-pair gather (default); bit 2: forward-only launches
  // run both 16-point tiles together (default)
  static const int group = (getenv("NRC_QUERY_GROUP") ? atoi(getenv("NRC_QUERY_GROUP")) : 0) |
                           ((getenv("NRC_QUERY_PAIR") ? atoi(getenv("NRC_QUERY_PAIR")) : 1) ? 2 : 0) |
                           ((getenv("NRC_QUERY_DUAL") ? atoi(getenv("NRC_QUERY_DUAL")) : 1) ? 4 : 0);
  const bool grad = kFused && out.raw_grad != nullptr;
  const size_t smem = grad ? sizeof(FwdSmemBf16) : offsetof(FwdSmemBf16, wg);
  // resident CTAs per SM: kMinCtas by registers; the gradient scratch (45 KB per CTA) caps it at 4
  const int resident = (grad && kMinCtas > 4) ? 4 : kMinCtas;
  const int64_t cap = static_cast<int64_t>(num_sms()) * (mult_env > 0 ? mult_env : resident);
  // 32 points per warp.  The 16-point mode (NRC_QUERY_PPW=16: twice the CTAs for launches that leave SMs idle, e.g.
  // 32 768 points = 256 CTAs) measured SLOWER on the config-2 step (0.979 vs 0.951 ms, profiles/r01j_ab_runs.txt, block j7): the weight
  // staging per CTA is paid twice as often and the side streams already fill the idle SMs.
  int ppw = 32;
  if (ppw_env == 16 && kFused) ppw = 16;
  const int64_t tiles = (P + kBfWarps * ppw - 1) / (kBfWarps * ppw);
  const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
  mlp_bf16_fwd_kernel<F, KS0, kFused, kMinCtas><<<grid, kBfThreads, smem, st>>>(d, *mlp, in, P, warp_c, bias, out,
                                                                               group, ppw);
  return check_launch();
}

// Resident CTAs per SM: 4 (128 registers).  Measured on B200 (profiles/r01j_ab_runs.txt, block j2): capping registers for 5 or 6 CTAs
// per SM (96 / 80 registers, a few spills) is 3-8 % SLOWER on every workload - the kernel is bound by L1TEX
// wavefronts (gathers + ldmatrix), and more resident shared memory leaves less L1.
template <int F, int KS0, bool kFused>
int32_t launch_bf16_fwd(cudaStream_t st, const EncDev& d, const nrc_density_mlp_t* mlp, const float* in,
                        int64_t P, float warp_c, float bias, const QueryOut& out) {
  return launch_bf16_fwd_occ<F, KS0, kFused, 4>(st, d, mlp, in, P, warp_c, bias, out);
}

int32_t density_mlp_fwd_bf16(cudaStream_t s, const nrc_density_mlp_t* mlp, const float* d_enc, int64_t P,
                             float* d_raw, float* d_feat, float* d_gp) {
  EncDev d{};
  QueryOut out{nullptr, d_raw, d_feat, d_gp, nullptr, nullptr};
  if (mlp->in_dim <= 16) return launch_bf16_fwd<1, 1, false>(s, d, mlp, d_enc, P, 0.f, 0.f, out);
  return launch_bf16_fwd<1, 2, false>(s, d, mlp, d_enc, P, 0.f, 0.f, out);
}

int32_t density_query_fwd_bf16(cudaStream_t s, const EncDev& d, const nrc_density_mlp_t* mlp,
                               const float* d_means, int64_t P, float warp_c, float bias, float* density,
                               float* raw, float* feat, float* gp, float* rg, float* enc_out) {
  QueryOut out{density, raw, feat, gp, rg, enc_out};
  const bool small = mlp->in_dim <= 16;
  switch (d.F) {
    case 1: return small ? launch_bf16_fwd<1, 1, true>(s, d, mlp, d_means, P, warp_c, bias, out)
                         : launch_bf16_fwd<1, 2, true>(s, d, mlp, d_means, P, warp_c, bias, out);
    case 2: return small ? launch_bf16_fwd<2, 1, true>(s, d, mlp, d_means, P, warp_c, bias, out)
                         : launch_bf16_fwd<2, 2, true>(s, d, mlp, d_means, P, warp_c, bias, out);
    case 4: return small ? launch_bf16_fwd<4, 1, true>(s, d, mlp, d_means, P, warp_c, bias, out)
                         : launch_bf16_fwd<4, 2, true>(s, d, mlp, d_means, P, warp_c, bias, out);
    case 8: return small ? launch_bf16_fwd<8, 1, true>(s, d, mlp, d_means, P, warp_c, bias, out)
                         : launch_bf16_fwd<8, 2, true>(s, d, mlp, d_means, P, warp_c, bias, out);
  }
  return NRC_E_UNSUPPORTED;
}

}  // namespace nrc
