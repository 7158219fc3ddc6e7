// K3 backward (tensor-core variant).  Persistent CTAs of four warps loop over 128-point tiles:
//   phase 1 (per warp, 32 points): recompute the forward activations from the saved encoded
//     features with register-chained MMAs, form g_h2 / g_h1 / g_enc (data gradients through the
//     transposed weights), and leave bf16 copies of x, h1, h2, g2, g1, g_out in shared memory;
//   phase 2 (per CTA): weight gradients dW = A^T G as MMAs over the 128-point tile with the
//     transposed-ldmatrix trick, accumulated in fp32 registers across all tiles of the CTA.
// One atomic pass per CTA at the end.  Reference: XLA autodiff of geometry.py:155-168,467.
#include <cstdint>
#include <cstdlib>

#include "mma_bf16.cuh"

namespace nrc {

constexpr int kBwWarps = 4;
constexpr int kBwThreads = kBwWarps * 32;
constexpr int kBwTile = kBwThreads;  // points per tile

struct BwdSmemBf16 {
  MlpWeightsBf16 w;
  __nv_bfloat16 x[kBwTile][kXStride];
  __nv_bfloat16 xd[kBwTile][kXStride];   // tangent mode: edot (the A operand of dW0 there); 113.8 KB total, 2 CTAs/SM
  __nv_bfloat16 h1[kBwTile][kWStride];
  __nv_bfloat16 h2[kBwTile][kWStride];
  __nv_bfloat16 g2[kBwTile][kWStride];
  __nv_bfloat16 g1[kBwTile][kWStride];
  __nv_bfloat16 go[kBwTile][8];
};

// Store an A-fragment image (16 rows x 64 cols, bf16) back to a row-major tile.
__device__ __forceinline__ void store_afrag(__nv_bfloat16* tile, int stride, int row0, const uint32_t (&a)[4][4],
                                            int lane) {
  const int r = lane >> 2, c = (lane & 3) * 2;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = (2 * kk + h) * 8 + c;
      *reinterpret_cast<uint32_t*>(tile + (row0 + r) * stride + col) = a[kk][2 * h + 0];
      *reinterpret_cast<uint32_t*>(tile + (row0 + r + 8) * stride + col) = a[kk][2 * h + 1];
    }
  }
}

// The warp's 32 feature rows are one contiguous block of 32 * in_dim floats: read it with 16-byte loads (all issued
// before the first conversion) and scatter the bf16 values into the row-major tile.  Rows past `nvalid` are zeros.
// The per-lane row-wise version (32 scalar loads of a 128-byte stride per lane) was 13 % of the kernel's warp
// samples, stalled on the LSU queue (profiles/r02i).
template <int KS0>
__device__ __forceinline__ void load_rows_bf16(__nv_bfloat16* tile, const float* __restrict__ src, int in_dim, int nvalid,
                                               int lane) {
  const int total = nvalid * in_dim, span = 32 * in_dim;
  const bool vec = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  float4 v[4 * KS0];
#pragma unroll
  for (int it = 0; it < 4 * KS0; ++it) {
    const int e = 4 * (lane + 32 * it);
    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e >= span) continue;
    if (vec && e + 3 < total) {
      v[it] = __ldg(reinterpret_cast<const float4*>(src + e));
    } else {
      if (e < total) v[it].x = __ldg(src + e);
      if (e + 1 < total) v[it].y = __ldg(src + e + 1);
      if (e + 2 < total) v[it].z = __ldg(src + e + 2);
      if (e + 3 < total) v[it].w = __ldg(src + e + 3);
    }
  }
#pragma unroll
  for (int it = 0; it < 4 * KS0; ++it) {
    const int e = 4 * (lane + 32 * it);
    if (e >= span) continue;
    int row = e / in_dim, col = e - row * in_dim;
    const float f[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (e + c < span) tile[row * kXStride + col] = __float2bfloat16(f[c]);
      if (++col == in_dim) { col = 0; ++row; }
    }
  }
}

template <int KS0>
__global__ void __launch_bounds__(kBwThreads)
mlp_bf16_bwd_kernel(const nrc_density_mlp_t m, const float* __restrict__ enc, const float* __restrict__ g_raw,
                    const float* __restrict__ density, const float* __restrict__ g_feat,
                    const float* __restrict__ g_gp, int64_t P, float* __restrict__ g_enc,
                    const nrc_density_mlp_grad_t grads, int want_wgrad, const float* __restrict__ enc_dot) {
  // enc_dot != nullptr: TANGENT mode (second-order path of the analytic normals).  The MLP is piecewise linear, so
  // d/d theta <g, d raw / d x> is the ordinary backward pass of the TANGENT network - same ReLU masks as the primal,
  // activations replaced by the tangents h1dot = M1 (W0^T edot), h2dot = M2 (W1^T h1dot), no biases - with upstream
  // d rawdot = 1:  dW0 += edot (x) a1, dW1 += h1dot (x) a2, dwd += h2dot, g_enc = W0 a1 (= d raw / d e, for the table
  // scatter).  The primal forward is recomputed from `enc` for the masks only.
  const bool tangent = enc_dot != nullptr;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmemBf16& s = *reinterpret_cast<BwdSmemBf16*>(smem_raw);
  load_weights_bf16<kBwThreads>(s.w, m);
  {
    // the padding columns [in_dim, 16 KS0) of the feature tiles stay zero for the whole launch
    uint32_t* z = reinterpret_cast<uint32_t*>(&s.x[0][0]);
    for (int i = threadIdx.x; i < kBwTile * kXStride / 2; i += kBwThreads) z[i] = 0u;
    if (tangent) {
      z = reinterpret_cast<uint32_t*>(&s.xd[0][0]);
      for (int i = threadIdx.x; i < kBwTile * kXStride / 2; i += kBwThreads) z[i] = 0u;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int in_dim = m.in_dim;
  const int r = lane >> 2, cq = (lane & 3) * 2;

  float accW1[8][4], accW0[KS0][2][4], accWo[4];
  // bias gradients = column sums of g2 / g1 / g_out over the points: the same tiles times an all-ones A operand in
  // phase 2 (every row of the product is the column sum).  Replaces per-warp shuffle reductions + shared-memory
  // atomics in phase 1 (ATOMS.CAST.SPIN loops, ~14 % of the warp samples in profiles/r02i).
  float accB1[2][4], accB0[2][4], accBo[4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) { accB1[i][e] = 0.f; accB0[i][e] = 0.f; }
#pragma unroll
  for (int e = 0; e < 4; ++e) accBo[e] = 0.f;
  const uint32_t ones[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) accW1[i][e] = 0.f;
#pragma unroll
  for (int i = 0; i < KS0; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 4; ++e) accW0[i][h][e] = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) accWo[e] = 0.f;

  const int64_t num_tiles = (P + kBwTile - 1) / kBwTile;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t base = tile * kBwTile + warp * 32;
    // ------------------------------ phase 1 ------------------------------------------------
    {
      const int64_t p = base + lane;
      const bool valid = p < P;
      const int row = warp * 32 + lane;
      if (base < P) {
        const int nvalid = static_cast<int>(P - base < 32 ? P - base : 32);
        load_rows_bf16<KS0>(&s.x[warp * 32][0], enc + base * in_dim, in_dim, nvalid, lane);
        if (tangent) load_rows_bf16<KS0>(&s.xd[warp * 32][0], enc_dot + base * in_dim, in_dim, nvalid, lane);
      } else {
        for (int k = 0; k < in_dim; ++k) { s.x[row][k] = __float2bfloat16(0.f); if (tangent) s.xd[row][k] = __float2bfloat16(0.f); }
      }
      float go[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid && tangent) {
        go[0] = 1.0f;                                   // upstream of the tangent network: d rawdot = 1
      } else if (valid) {
        go[0] = density ? g_raw[p] * density[p] : g_raw[p];
        if (g_gp) { go[1] = g_gp[3 * p]; go[2] = g_gp[3 * p + 1]; go[3] = g_gp[3 * p + 2]; }
      }
      *reinterpret_cast<uint2*>(&s.go[row][0]) = make_uint2(pack_bf16(go[0], go[1]), pack_bf16(go[2], go[3]));
      *reinterpret_cast<uint2*>(&s.go[row][4]) = make_uint2(0u, 0u);
    }
    __syncwarp();
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int row0 = warp * 32 + mt * 16;
      const int64_t pr[2] = {base + mt * 16 + r, base + mt * 16 + r + 8};
      // upstream feature gradients of this 16-point tile: requested now, used after the recomputed forward
      float2 gfr[8][2];
      const bool has_gf = g_feat && !tangent;
      if (has_gf) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            gfr[nt][h] = pr[h] < P ? __ldg(reinterpret_cast<const float2*>(g_feat + pr[h] * kHid + nt * 8 + cq))
                                   : make_float2(0.f, 0.f);
      }
      uint32_t a0[KS0][4];
#pragma unroll
      for (int ks = 0; ks < KS0; ++ks) load_a_frag(a0[ks], &s.x[0][0], kXStride, row0, ks * 16, lane);
      float acc[8][4];
      mma_layer64<KS0>(acc, a0, &s.w.w0t[0][0], kXStride, s.w.b0, lane);
      uint32_t h1f[4][4];
      acc_to_afrag<true>(acc, h1f);
      uint32_t f2[4][4];
      uint32_t m1 = 0u, m2 = 0u;     // ReLU masks of the primal in accumulator layout (bit nt*4+e), tangent mode
      if (!tangent) {
        store_afrag(&s.h1[0][0], kWStride, row0, h1f, lane);
        mma_layer64<4>(acc, h1f, &s.w.w1t[0][0], kWStride, s.w.b1, lane);
        acc_to_afrag<true>(acc, f2);
        store_afrag(&s.h2[0][0], kWStride, row0, f2, lane);
      } else {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) m1 |= (acc[nt][e] > 0.f ? 1u : 0u) << (nt * 4 + e);
        mma_layer64<4>(acc, h1f, &s.w.w1t[0][0], kWStride, s.w.b1, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) m2 |= (acc[nt][e] > 0.f ? 1u : 0u) << (nt * 4 + e);
        // tangent forward on edot (staged with the primal rows above)
#pragma unroll
        for (int ks = 0; ks < KS0; ++ks) load_a_frag(a0[ks], &s.xd[0][0], kXStride, row0, ks * 16, lane);
        mma_layer64<KS0>(acc, a0, &s.w.w0t[0][0], kXStride, nullptr, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[nt][e] = ((m1 >> (nt * 4 + e)) & 1u) ? acc[nt][e] : 0.f;
        acc_to_afrag<false>(acc, f2);
        store_afrag(&s.h1[0][0], kWStride, row0, f2, lane);                 // h1 tile <- h1dot
        mma_layer64<4>(acc, f2, &s.w.w1t[0][0], kWStride, nullptr, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[nt][e] = ((m2 >> (nt * 4 + e)) & 1u) ? acc[nt][e] : 0.f;
        acc_to_afrag<false>(acc, f2);
        store_afrag(&s.h2[0][0], kWStride, row0, f2, lane);                 // h2 tile <- h2dot
      }
      // g_h2 = (Wo go + g_feat) * [h2 > 0], in accumulator layout
      float gor[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool v = pr[h] < P;
        gor[h][0] = v ? (tangent ? 1.0f : (density ? g_raw[pr[h]] * density[pr[h]] : g_raw[pr[h]])) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) gor[h][1 + c] = (v && g_gp && !tangent) ? g_gp[3 * pr[h] + c] : 0.f;
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + cq;
        const float4 w0 = *reinterpret_cast<const float4*>(&s.w.wo[col][0]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s.w.wo[col + 1][0]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float ga = gor[h][0] * w0.x + gor[h][1] * w0.y + gor[h][2] * w0.z + gor[h][3] * w0.w;
          float gb = gor[h][0] * w1.x + gor[h][1] * w1.y + gor[h][2] * w1.z + gor[h][3] * w1.w;
          if (has_gf) { ga += gfr[nt][h].x; gb += gfr[nt][h].y; }
          const bool on0 = tangent ? ((m2 >> (nt * 4 + 2 * h)) & 1u) : (acc[nt][2 * h] > 0.f);
          const bool on1 = tangent ? ((m2 >> (nt * 4 + 2 * h + 1)) & 1u) : (acc[nt][2 * h + 1] > 0.f);
          acc[nt][2 * h] = on0 ? ga : 0.f;
          acc[nt][2 * h + 1] = on1 ? gb : 0.f;
        }
      }
      acc_to_afrag<false>(acc, f2);
      store_afrag(&s.g2[0][0], kWStride, row0, f2, lane);
      // g_h1 = (g_h2 W1^T) * [h1 > 0]
      mma_layer64_t<4>(acc, f2, &s.w.w1t[0][0], kWStride, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float2 lo = unpack_bf16(h1f[nt >> 1][2 * (nt & 1)]);
        float2 hi = unpack_bf16(h1f[nt >> 1][2 * (nt & 1) + 1]);
        // h1f holds relu(h1) of the primal in A-fragment layout, i.e. the same (row, column) pairs as acc[nt][0..3]
        acc[nt][0] = lo.x > 0.f ? acc[nt][0] : 0.f;
        acc[nt][1] = lo.y > 0.f ? acc[nt][1] : 0.f;
        acc[nt][2] = hi.x > 0.f ? acc[nt][2] : 0.f;
        acc[nt][3] = hi.y > 0.f ? acc[nt][3] : 0.f;
      }
      acc_to_afrag<false>(acc, f2);
      store_afrag(&s.g1[0][0], kWStride, row0, f2, lane);
      if (g_enc) {
        float ge[2 * KS0][4];
#pragma unroll
        for (int nt = 0; nt < 2 * KS0; ++nt) ge[nt][0] = ge[nt][1] = ge[nt][2] = ge[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int np = 0; np < KS0; ++np) {
            uint32_t b[4];
            load_b_frag2_trans(b, &s.w.w0t[0][0], kXStride, ks * 16, np * 16, lane);
            mma_bf16(ge[2 * np], f2[ks], b[0], b[1]);
            mma_bf16(ge[2 * np + 1], f2[ks], b[2], b[3]);
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2 * KS0; ++nt) {
          const int c = nt * 8 + cq;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (pr[h] >= P) continue;
            if (c < in_dim) g_enc[pr[h] * in_dim + c] = ge[nt][2 * h];
            if (c + 1 < in_dim) g_enc[pr[h] * in_dim + c + 1] = ge[nt][2 * h + 1];
          }
        }
      }
    }
    if (!want_wgrad) { __syncwarp(); continue; }
    __syncthreads();
    // ------------------------------ phase 2 ------------------------------------------------
#pragma unroll 2
    for (int ks = 0; ks < kBwTile / 16; ++ks) {
      const int k0 = ks * 16;
      uint32_t a[4], b[4];
      // dW1[k][j]: rows k = 16*warp.., all 64 columns
      load_a_frag_trans(a, &s.h1[0][0], kWStride, k0, warp * 16, lane);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        load_b_frag2_trans(b, &s.g2[0][0], kWStride, k0, np * 16, lane);
        mma_bf16(accW1[2 * np], a, b[0], b[1]);
        mma_bf16(accW1[2 * np + 1], a, b[2], b[3]);
        if (np == warp) { mma_bf16(accB1[0], ones, b[0], b[1]); mma_bf16(accB1[1], ones, b[2], b[3]); }
      }
      // dW0[i][k]: all rows i, columns k = 16*warp..
      load_b_frag2_trans(b, &s.g1[0][0], kWStride, k0, warp * 16, lane);
      mma_bf16(accB0[0], ones, b[0], b[1]);
      mma_bf16(accB0[1], ones, b[2], b[3]);
#pragma unroll
      for (int mi = 0; mi < KS0; ++mi) {
        load_a_frag_trans(a, tangent ? &s.xd[0][0] : &s.x[0][0], kXStride, k0, mi * 16, lane);
        mma_bf16(accW0[mi][0], a, b[0], b[1]);
        mma_bf16(accW0[mi][1], a, b[2], b[3]);
      }
      // dWo[j][c]: rows j = 16*warp.., 8 head columns
      load_a_frag_trans(a, &s.h2[0][0], kWStride, k0, warp * 16, lane);
      uint32_t bo[2];
      load_b_frag1_trans(bo, &s.go[0][0], 8, k0, 0, lane);
      mma_bf16(accWo, a, bo[0], bo[1]);
      if (warp == 0) mma_bf16(accBo, ones, bo[0], bo[1]);
    }
    __syncthreads();
  }
  if (!want_wgrad) return;
  // Each lane holds column pairs (cq, cq+1); the odd lane of a pair hands its two values to the even one, which issues
  // ONE 16-byte reduction for columns cq..cq+3 (cq = 0 or 4 there).  Scalar fallback for unaligned gradient buffers.
  const bool vec = ((reinterpret_cast<uintptr_t>(grads.d_w1) | reinterpret_cast<uintptr_t>(grads.d_w0)) & 15) == 0;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = warp * 16 + r + 8 * h, col = nt * 8 + cq;
      const float v0 = accW1[nt][2 * h], v1 = accW1[nt][2 * h + 1];
      const float o0 = __shfl_xor_sync(0xffffffffu, v0, 1), o1 = __shfl_xor_sync(0xffffffffu, v1, 1);
      if (vec) {
        if ((lane & 1) == 0) red_add_v4(grads.d_w1 + row * kHid + col, v0, v1, o0, o1);
      } else {
        atomicAdd(grads.d_w1 + row * kHid + col, v0);
        atomicAdd(grads.d_w1 + row * kHid + col + 1, v1);
      }
    }
#pragma unroll
  for (int mi = 0; mi < KS0; ++mi)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = mi * 16 + r + 8 * h, col = warp * 16 + hh * 8 + cq;
        const float v0 = accW0[mi][hh][2 * h], v1 = accW0[mi][hh][2 * h + 1];
        const float o0 = __shfl_xor_sync(0xffffffffu, v0, 1), o1 = __shfl_xor_sync(0xffffffffu, v1, 1);
        if (row >= in_dim) continue;   // uniform within a lane pair (same row)
        if (vec) {
          if ((lane & 1) == 0) red_add_v4(grads.d_w0 + row * kHid + col, v0, v1, o0, o1);
        } else {
          atomicAdd(grads.d_w0 + row * kHid + col, v0);
          atomicAdd(grads.d_w0 + row * kHid + col + 1, v1);
        }
      }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int j = warp * 16 + r + (e >= 2 ? 8 : 0), c = cq + (e & 1);
    if (c == 0) atomicAdd(grads.d_wd + j, accWo[e]);
    else if (c < 4 && grads.d_wn) atomicAdd(grads.d_wn + j * 3 + (c - 1), accWo[e]);
  }
  if (tangent) return;          // no bias terms in the tangent network (and the caller may not pass bias buffers)
  if (lane < 4) {               // row 0 of the all-ones products: columns cq, cq + 1 of each 8-wide tile
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        atomicAdd(grads.d_b1 + warp * 16 + t * 8 + cq + e, accB1[t][e]);
        atomicAdd(grads.d_b0 + warp * 16 + t * 8 + cq + e, accB0[t][e]);
      }
    if (warp == 0) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = cq + e;
        if (c == 0) atomicAdd(grads.d_bd, accBo[e]);
        else if (c < 4 && grads.d_bn) atomicAdd(grads.d_bn + (c - 1), accBo[e]);
      }
    }
  }
}

template <int KS0>
int32_t launch_bf16_bwd(cudaStream_t st, const nrc_density_mlp_t* mlp, const float* enc, const float* g_raw,
                        const float* density, const float* g_feat, const float* g_gp, int64_t P, float* g_enc,
                        const nrc_density_mlp_grad_t* grads, const float* enc_dot = nullptr) {
  if (const int32_t st_attr = ensure_dynamic_smem<mlp_bf16_bwd_kernel<KS0>>(static_cast<int>(sizeof(BwdSmemBf16))); st_attr != NRC_OK) return st_attr;
  nrc_density_mlp_grad_t g{};
  if (grads) g = *grads;
  int64_t tiles = (P + kBwTile - 1) / kBwTile;
  // persistent CTAs, two per SM (103 KB of shared memory each, since the data-gradient GEMMs read the forward weight
  // tiles through transposing loads instead of a second copy).  An earlier "2-4 CTAs per SM are slower" measurement
  // was taken when only ONE 117 KB CTA fitted per SM, i.e. it measured extra staging passes, not extra warps.
  static const int mult = getenv("NRC_MLP_BWD_GRID_MULT") ? atoi(getenv("NRC_MLP_BWD_GRID_MULT")) : 2;
  unsigned grid = static_cast<unsigned>(tiles < num_sms() * mult ? tiles : num_sms() * mult);
  mlp_bf16_bwd_kernel<KS0><<<grid, kBwThreads, sizeof(BwdSmemBf16), st>>>(*mlp, enc, g_raw, density, g_feat, g_gp,
                                                                          P, g_enc, g, grads ? 1 : 0, enc_dot);
  return check_launch();
}

int32_t density_mlp_bwd_bf16(cudaStream_t s, const nrc_density_mlp_t* mlp, const float* d_enc,
                             const float* d_g_raw, const float* d_density, const float* d_g_feat,
                             const float* d_g_gp, int64_t P, float* d_g_enc, const nrc_density_mlp_grad_t* grads) {
  if (mlp->in_dim <= 16)
    return launch_bf16_bwd<1>(s, mlp, d_enc, d_g_raw, d_density, d_g_feat, d_g_gp, P, d_g_enc, grads);
  return launch_bf16_bwd<2>(s, mlp, d_enc, d_g_raw, d_density, d_g_feat, d_g_gp, P, d_g_enc, grads);
}

// Tangent mode (see the kernel): d_enc = primal features (masks), d_enc_dot = tangent features, both [P, in_dim];
// d_g_enc [P, in_dim] <- W0 a1; grads: d_w0, d_w1, d_wd accumulated (biases and the normal head receive nothing).
int32_t density_mlp_bwd_tangent_bf16(cudaStream_t s, const nrc_density_mlp_t* mlp, const float* d_enc,
                                     const float* d_enc_dot, int64_t P, float* d_g_enc,
                                     const nrc_density_mlp_grad_t* grads) {
  if (mlp->in_dim <= 16)
    return launch_bf16_bwd<1>(s, mlp, d_enc, nullptr, nullptr, nullptr, nullptr, P, d_g_enc, grads, d_enc_dot);
  return launch_bf16_bwd<2>(s, mlp, d_enc, nullptr, nullptr, nullptr, nullptr, P, d_g_enc, grads, d_enc_dot);
}

}  // namespace nrc

// Second-order path of the analytic normals on tensor cores: the MLP part (see the kernel's tangent mode).  Reference:
// internal/geometry.py:442-460 (jax.vjp of predict_density w.r.t. the means) differentiated by the predicted-normal loss
// (internal/loss_utils.py:169-199); fp32 counterpart: nrc_density_normals_bwd.
extern "C" int32_t nrc_density_mlp_bwd_tangent(void* stream, const nrc_density_mlp_t* mlp, const float* d_enc,
                                               const float* d_enc_dot, int64_t num_points, float* d_g_enc,
                                               const nrc_density_mlp_grad_t* grads) {
  if (!mlp || !mlp->d_w0 || !mlp->d_b0 || !mlp->d_w1 || !mlp->d_b1 || !mlp->d_wd || !mlp->d_bd) return NRC_E_INVALID_ARG;
  if (mlp->width != nrc::kHid || mlp->in_dim < 1 || mlp->in_dim > 32) return NRC_E_UNSUPPORTED;
  if (num_points < 0) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_enc || !d_enc_dot || !grads || !grads->d_w0 || !grads->d_w1 || !grads->d_wd) return NRC_E_INVALID_ARG;
  return nrc::density_mlp_bwd_tangent_bf16(static_cast<cudaStream_t>(stream), mlp, d_enc, d_enc_dot, num_points, d_g_enc,
                                           grads);
}
