// Fused MLP chains of the cache shader on the 5th-generation
// tensor cores: tcgen05.mma with accumulators in tensor memory, operands in 128-byte-swizzled
// shared-memory atoms (tc05.cuh), weights streamed by bulk async copies through an mbarrier ring.
//
// A chain is a small PROGRAM interpreted per 128-point tile (nrc_chain_program_t):
//   LOAD  fp32 rows from global memory -> bf16 atom slots (concatenation = several LOADs)
//   GEMM  D[tmem] (+)= A[slots] * W[packed chunks]            (consecutive GEMMs form a group)
//   EPI   D -> (+bias, ReLU | ReLU-mask) -> bf16 atom slots and / or fp32 global output
//   SAVE  atom slots -> global "tile image" (kept for the backward pass / weight gradients)
// so that a whole MLP (all layers, skip connections as extra K atoms, forward or data-gradient)
// runs without its activations leaving the SM.  Reference bodies replaced: flax.linen.Dense stacks of
// internal/nerf.py:232-345,561-689, internal/surface_light_field.py:352-403,480-500,
// internal/geometry.py:127-168, internal/material.py:2073-2123.
//
// Roles inside a CTA (320 threads, 1 CTA / SM, persistent over tile pairs):
//   warp 0      weight producer  (cp.async.bulk global -> ring, one elected lane)
//   warp 1      UMMA issuer      (one elected lane)
//   warps 2-5   epilogue / loader warpgroup of context 0   (TMEM lane quadrant = warp % 4)
//   warps 6-9   epilogue / loader warpgroup of context 1
// Two contexts = two independent tiles in flight: while one context's warpgroup drains its
// accumulator (TMEM -> registers -> bf16 -> shared memory), the tensor core runs the other
// context's layer.
#include <cuda_bf16.h>

#include "encode.cuh"
#include "nrc_common.cuh"
#include "tc05.cuh"

namespace nrc {
using namespace tc;

constexpr int kChainThreads = 320;
constexpr int kCtxTmemCols = 256;   // TMEM columns per tile context of ordinary programs (queries allocate fewer)

struct ChainParams {
  nrc_chain_program_t prog;
  void* ptrs[NRC_CHAIN_MAX_PTRS];
  const uint8_t* weights;
  int64_t num_rows;
  int32_t num_tiles;
  int32_t ring_stages;
  int32_t ctx_tmem_cols;   // TMEM columns per tile context (power of two; the CTA allocates twice this)
  EncDev enc;       // hash-grid front end of GATHER ops (nrc_chain_query); unused otherwise
  float warp_c;
};

// GATHER op: the thread owning tile row r contracts point row0 + r, gathers the multiresolution features
// (level_interp: bit-identical to nrc_encode_fwd) and writes them as one bf16 row of the destination atom.
template <int F>
__device__ __forceinline__ bool gather_row(const ChainParams& p, const nrc_chain_op_t& op, int64_t pt, bool valid,
                                           uint32_t slot_base, int r) {
  constexpr int kMaxL = 32 / F > 8 ? 8 : 32 / F;
  float feat[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) feat[i] = 0.f;
  bool inside = false;
  if (valid) {
    const float* m = static_cast<const float*>(p.ptrs[op.ptr]) + 3 * pt;
    const float x0 = __ldg(m), x1 = __ldg(m + 1), x2 = __ldg(m + 2);
    float z[3], xn[3];
    contract_point(p.warp_c, x0, x1, x2, z[0], z[1], z[2]);
    normalise_point(p.enc, z, xn);
    inside = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) inside = inside && (z[a] > p.enc.b0[a]) && (z[a] < p.enc.b1[a]);
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      if (l < p.enc.L) {
        const Corners c = level_setup(p.enc.lv[l], xn);
        const FeatVec<F> v = level_interp<F>(p.enc.lv[l], c);
#pragma unroll
        for (int f = 0; f < F; ++f) feat[l * F + f] = __fmul_rn(v.v[f], p.enc.scale);
      }
    }
    if (op.out_ptr >= 0) {
      float* eo = static_cast<float*>(p.ptrs[op.out_ptr]) + pt * op.ncols;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < op.ncols) eo[i] = feat[i];
    }
  }
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (ch * 8 < op.npad) {
      const uint32_t dst = slot_base + atom_chunk_offset(r, ch);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack2_bf16(feat[8 * ch], feat[8 * ch + 1])),
                   "r"(pack2_bf16(feat[8 * ch + 2], feat[8 * ch + 3])), "r"(pack2_bf16(feat[8 * ch + 4], feat[8 * ch + 5])),
                   "r"(pack2_bf16(feat[8 * ch + 6], feat[8 * ch + 7]))
                   : "memory");
    }
  }
  return inside;
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Epilogue of one accumulator region for the tile row this thread owns: 16 columns at a time
// TMEM -> registers -> (+bias, ReLU | ReLU mask) -> bf16 atom slots and / or fp32 global row.
// Specialised at compile time so the per-element work is 2-3 instructions.
struct EpiArgs {
  uint32_t taddr;            // TMEM address (lane quadrant + first column)
  const float* bias;         // [ncols] or nullptr
  float* out_row;            // fp32 output row (already offset to col0) or nullptr
  const uint8_t* mask_tile;  // forward-activation image of this tile or nullptr
  uint32_t slot0_addr;       // shared-memory address of the first destination slot
  int ncols, npad, mask_atom0, r;
  bool out_vec, accum, has_slot;
};

template <bool BIAS, bool RELU, bool MASK, bool FULL>
__device__ __forceinline__ void epi_chunk(const EpiArgs& a, int j0, const uint32_t (&v)[16], const uint32_t (&mw)[8],
                                          const float (&b)[16]) {
  float x[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    float t = __uint_as_float(v[e]);
    if (BIAS) t += b[e];
    if (RELU) t = fmaxf(t, 0.f);
    if (MASK) {   // bf16 activation > 0  <=>  its 16 bits, read as a signed short, are > 0
      const short h = static_cast<short>((mw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu);
      t = h > 0 ? t : 0.f;
    }
    if (!FULL && j0 + e >= a.ncols) t = 0.f;
    x[e] = t;
  }
  if (a.has_slot) {
    const uint32_t sa = a.slot0_addr + static_cast<uint32_t>(j0 >> 6) * kAtomBytes;
    const uint32_t d0 = sa + atom_chunk_offset(a.r, (j0 & 63) >> 3);
    const uint32_t d1 = sa + atom_chunk_offset(a.r, ((j0 & 63) >> 3) + 1);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d0), "r"(pack2_bf16(x[0], x[1])),
                 "r"(pack2_bf16(x[2], x[3])), "r"(pack2_bf16(x[4], x[5])), "r"(pack2_bf16(x[6], x[7]))
                 : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d1), "r"(pack2_bf16(x[8], x[9])),
                 "r"(pack2_bf16(x[10], x[11])), "r"(pack2_bf16(x[12], x[13])), "r"(pack2_bf16(x[14], x[15]))
                 : "memory");
  }
  if (a.out_row && (FULL || j0 < a.ncols)) {
    float* o = a.out_row + j0;
    if (a.out_vec && (FULL || j0 + 16 <= a.ncols)) {
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        float4 w = make_float4(x[e], x[e + 1], x[e + 2], x[e + 3]);
        if (a.accum) {
          const float4 old = *reinterpret_cast<const float4*>(o + e);
          w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
        }
        *reinterpret_cast<float4*>(o + e) = w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (j0 + e < a.ncols) o[e] = a.accum ? o[e] + x[e] : x[e];
    }
  }
}

template <bool BIAS, bool RELU, bool MASK, bool FULL>
__device__ __forceinline__ void epi_loads(const EpiArgs& a, int j0, uint32_t (&mw)[8], float (&b)[16]) {
  if (MASK) {
    const int mc = a.mask_atom0 * 64 + j0;
    const uint8_t* ma = a.mask_tile + static_cast<size_t>(mc >> 6) * kAtomBytes;
    const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(ma + atom_chunk_offset(a.r, (mc & 63) >> 3)));
    const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(ma + atom_chunk_offset(a.r, ((mc & 63) >> 3) + 1)));
    mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
    mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
  }
  if (BIAS) {
    if (FULL) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.bias + j0) + q);
        b[4 * q] = t.x; b[4 * q + 1] = t.y; b[4 * q + 2] = t.z; b[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) b[e] = (j0 + e < a.ncols) ? __ldg(a.bias + j0 + e) : 0.f;
    }
  }
}

// 16 columns per iteration; the mask / bias loads are issued between the TMEM load and its wait.
template <bool BIAS, bool RELU, bool MASK, bool FULL>
__device__ __forceinline__ void epi_run(const EpiArgs& a) {
  for (int j0 = 0; j0 < a.npad; j0 += 16) {
    uint32_t v[16], mw[8];
    float b[16];
    tmem_ld16(a.taddr + j0, v);
    epi_loads<BIAS, RELU, MASK, FULL>(a, j0, mw, b);
    tmem_ld_wait();
    epi_chunk<BIAS, RELU, MASK, FULL>(a, j0, v, mw, b);
  }
}

// ------------------------------------------------------------------------------------------------
template <int kMinCtas>
__global__ void __launch_bounds__(kChainThreads, kMinCtas) chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.prog.slots_per_ctx, R = p.ring_stages, nops = p.prog.num_ops;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring_base = base + 2u * S * kAtomBytes;
  auto slot_addr = [&](int ctx, int s) { return base + static_cast<uint32_t>(ctx * S + s) * kAtomBytes; };
  // barriers: [0,R) full, [4,4+R) empty, 8+c a_ready, 10+c acc_ready
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };
  auto a_ready = [&](int c) { return bar0 + 8u * (8 + c); };
  auto acc_ready = [&](int c) { return bar0 + 8u * (10 + c); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int c = 0; c < 2; ++c) {
      mbar_init(a_ready(c), 128);
      mbar_init(acc_ready(c), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 2u * p.ctx_tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int num_pairs = (p.num_tiles + 1) >> 1;

  if (warp == 0) {
    // ===================================================================== weight producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int q = blockIdx.x; q < num_pairs; q += gridDim.x) {
        for (int i = 0; i < nops;) {
          if (p.prog.ops[i].kind != NRC_OP_GEMM) { ++i; continue; }
          int j = i;
          while (j < nops && p.prog.ops[j].kind == NRC_OP_GEMM) ++j;
          for (int c = 0; c < 2; ++c) {
            if (2 * q + c >= p.num_tiles) continue;
            for (int o = i; o < j; ++o) {
              const nrc_chain_op_t& op = p.prog.ops[o];
              const uint32_t bytes = static_cast<uint32_t>(op.n) * 128u;
              for (int a = 0; a < op.n_atoms; ++a) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_arrive_expect_tx(full_bar(stage), bytes);
                bulk_g2s(ring_base + static_cast<uint32_t>(stage) * kAtomBytes,
                         p.weights + static_cast<size_t>(op.w_chunk + a) * kAtomBytes, bytes, full_bar(stage));
                if (++stage == R) { stage = 0; phase ^= 1u; }
              }
            }
          }
          i = j;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== UMMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_par[2] = {0u, 0u};
      for (int q = blockIdx.x; q < num_pairs; q += gridDim.x) {
        for (int i = 0; i < nops;) {
          if (p.prog.ops[i].kind != NRC_OP_GEMM) { ++i; continue; }
          int j = i;
          while (j < nops && p.prog.ops[j].kind == NRC_OP_GEMM) ++j;
          for (int c = 0; c < 2; ++c) {
            if (2 * q + c >= p.num_tiles) continue;
            mbar_wait(a_ready(c), a_par[c]);
            a_par[c] ^= 1u;
            tc_fence_after();
            for (int o = i; o < j; ++o) {
              const nrc_chain_op_t& op = p.prog.ops[o];
              const uint32_t idesc = make_idesc(128, op.n, 0, 0);
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(c * p.ctx_tmem_cols + op.tmem_col);
              for (int a = 0; a < op.n_atoms; ++a) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a_addr = slot_addr(c, op.a_slot[a]);
                const uint32_t b_addr = ring_base + static_cast<uint32_t>(stage) * kAtomBytes;
                const int nk = op.a_klen[a] >> 4;
                for (int k = 0; k < nk; ++k) {
                  const uint32_t acc = ((op.flags & NRC_GEMM_ACCUMULATE) || a > 0 || k > 0) ? 1u : 0u;
                  umma_bf16(d_tmem, kmajor_desc(a_addr, k), kmajor_desc(b_addr, k), idesc, acc);
                }
                umma_commit(empty_bar(stage));
                if (++stage == R) { stage = 0; phase ^= 1u; }
              }
            }
            umma_commit(acc_ready(c));
          }
          i = j;
        }
      }
    }
  } else {
    // ===================================================================== loader / epilogue warpgroups
    const int c = (warp - 2) >> 2;                 // context
    const int wg_tid = threadIdx.x - 64 - 128 * c;
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;                // tile row owned in epilogues
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    uint32_t acc_par = 0;
    bool store_pending = false;
    bool inside_reg = false;   // bbox mask of the point this thread gathered (GATHER -> NRC_EPI_DENSITY)

    auto guard_slots = [&]() {  // before overwriting slots a bulk store may still be reading
      if (store_pending) {
        if (wg_tid == 0) bulk_wait_read0();
        named_barrier_sync(1 + c, 128);
        store_pending = false;
      }
    };

    for (int q = blockIdx.x; q < num_pairs; q += gridDim.x) {
      const int tile = 2 * q + c;
      if (tile >= p.num_tiles) continue;
      const int64_t row0 = static_cast<int64_t>(tile) * 128;
      for (int i = 0; i < nops;) {
        const nrc_chain_op_t& op = p.prog.ops[i];
        if (op.kind == NRC_OP_GEMM) {
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(a_ready(c));
          while (i < nops && p.prog.ops[i].kind == NRC_OP_GEMM) ++i;
          mbar_wait(acc_ready(c), acc_par);
          acc_par ^= 1u;
          tc_fence_after();
          continue;
        }
        if (op.kind == NRC_OP_GATHER) {
          guard_slots();
          const int64_t pt = row0 + r;
          const bool valid = pt < p.num_rows;
          const uint32_t sb = slot_addr(c, op.slot);
          switch (p.enc.F) {
            case 1: inside_reg = gather_row<1>(p, op, pt, valid, sb, r); break;
            case 2: inside_reg = gather_row<2>(p, op, pt, valid, sb, r); break;
            default: inside_reg = gather_row<4>(p, op, pt, valid, sb, r); break;
          }
        } else if (op.kind == NRC_OP_LOAD) {
          guard_slots();
          const float* src = op.ptr >= 0 ? static_cast<const float*>(p.ptrs[op.ptr]) : nullptr;
          const int nch = op.npad >> 3;
          const bool vec = src && (op.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
          // four independent 32-byte loads in flight per thread before the first conversion
          constexpr int kU = 4;
          const int total = 128 * nch;
          for (int item0 = wg_tid; item0 < total; item0 += 128 * kU) {
            float v[kU][8];
            int rrs[kU], cols[kU];
#pragma unroll
            for (int q = 0; q < kU; ++q) {
              const int item = item0 + q * 128;
              const int rr = item / nch;
              rrs[q] = rr;
              cols[q] = (item - rr * nch) * 8;
#pragma unroll
              for (int e = 0; e < 8; ++e) v[q][e] = 0.f;
              if (item < total && src && row0 + rr < p.num_rows && cols[q] < op.ncols) {
                const float* s = src + (row0 + rr) * op.ld + cols[q];
                if (vec && cols[q] + 8 <= op.ncols) {
                  const float4 x0 = __ldg(reinterpret_cast<const float4*>(s));
                  const float4 x1 = __ldg(reinterpret_cast<const float4*>(s) + 1);
                  v[q][0] = x0.x; v[q][1] = x0.y; v[q][2] = x0.z; v[q][3] = x0.w;
                  v[q][4] = x1.x; v[q][5] = x1.y; v[q][6] = x1.z; v[q][7] = x1.w;
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e)
                    if (cols[q] + e < op.ncols) v[q][e] = __ldg(s + e);
                }
              }
            }
#pragma unroll
            for (int q = 0; q < kU; ++q) {
              if (item0 + q * 128 >= total) continue;
              const int dcol = op.col0 + cols[q];
              const uint32_t dst = slot_addr(c, op.slot + (dcol >> 6)) + atom_chunk_offset(rrs[q], (dcol & 63) >> 3);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack2_bf16(v[q][0], v[q][1])),
                           "r"(pack2_bf16(v[q][2], v[q][3])), "r"(pack2_bf16(v[q][4], v[q][5])),
                           "r"(pack2_bf16(v[q][6], v[q][7]))
                           : "memory");
            }
          }
        } else if (op.kind == NRC_OP_SAVE) {
          fence_proxy_async_smem();
          named_barrier_sync(1 + c, 128);
          if (wg_tid == 0) {
            uint8_t* img = static_cast<uint8_t*>(p.ptrs[op.ptr]);
            for (int a = 0; a < op.npad; ++a)
              bulk_s2g(img + (static_cast<size_t>(tile) * op.img_atoms + op.col0 + a) * kAtomBytes,
                       slot_addr(c, op.slot + a), kAtomBytes);
            bulk_commit();
          }
          store_pending = true;
        } else if (op.kind == NRC_OP_EPI && (op.flags & NRC_EPI_DENSITY)) {
          // density head: column 0 -> safe_exp(raw + density_bias) masked to the bbox; columns 1..3 -> grad_pred
          uint32_t v[16];
          tmem_ld16(tmem_base + t_lane + static_cast<uint32_t>(c * p.ctx_tmem_cols + op.tmem_col), v);
          tmem_ld_wait();
          const int64_t pt = row0 + r;
          if (pt < p.num_rows) {
            const float* bias = op.ptr >= 0 ? static_cast<const float*>(p.ptrs[op.ptr]) : nullptr;
            const float raw = __uint_as_float(v[0]) + (bias ? __ldg(bias) : 0.f);
            static_cast<float*>(p.ptrs[op.out_ptr])[pt] = inside_reg ? safe_exp(raw + op.fparam) : 0.f;
            if (op.mask_ptr >= 0) {
              float* gp = static_cast<float*>(p.ptrs[op.mask_ptr]) + 3 * pt;
#pragma unroll
              for (int j = 0; j < 3; ++j) gp[j] = __uint_as_float(v[1 + j]) + (bias ? __ldg(bias + 1 + j) : 0.f);
            }
          }
        } else {  // NRC_OP_EPI
          if (op.slot >= 0) guard_slots();
          EpiArgs a;
          a.bias = op.ptr >= 0 ? static_cast<const float*>(p.ptrs[op.ptr]) : nullptr;
          float* out = op.out_ptr >= 0 ? static_cast<float*>(p.ptrs[op.out_ptr]) : nullptr;
          a.out_row = (out && row0 + r < p.num_rows) ? out + (row0 + r) * op.ld + op.col0 : nullptr;
          a.out_vec = out && (op.ld % 4 == 0) && (op.col0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
          a.accum = (op.flags & NRC_EPI_OUT_ACCUMULATE) != 0;
          a.mask_tile = op.mask_ptr >= 0 ? static_cast<const uint8_t*>(p.ptrs[op.mask_ptr]) +
                                               static_cast<size_t>(tile) * op.img_atoms * kAtomBytes
                                         : nullptr;
          a.mask_atom0 = op.mask_atom0;
          a.taddr = tmem_base + t_lane + static_cast<uint32_t>(c * p.ctx_tmem_cols + op.tmem_col);
          a.has_slot = op.slot >= 0;
          a.slot0_addr = a.has_slot ? slot_addr(c, op.slot) : 0u;
          a.ncols = op.ncols; a.npad = op.npad; a.r = r;
          const bool full = (op.ncols == op.npad) && (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0);
          const bool relu = (op.flags & NRC_EPI_RELU) != 0;
          if (a.mask_tile) {
            if (full) epi_run<false, false, true, true>(a); else epi_run<false, false, true, false>(a);
          } else if (a.bias) {
            if (relu) { if (full) epi_run<true, true, false, true>(a); else epi_run<true, true, false, false>(a); }
            else      { if (full) epi_run<true, false, false, true>(a); else epi_run<true, false, false, false>(a); }
          } else {
            if (relu) { if (full) epi_run<false, true, false, true>(a); else epi_run<false, true, false, false>(a); }
            else      { if (full) epi_run<false, false, false, true>(a); else epi_run<false, false, false, false>(a); }
          }
        }
        ++i;
      }
    }
    if (wg_tid == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 2u * p.ctx_tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// Weight packing: fp32 Flax kernels [in,out] -> bf16 16 KB chunks in the swizzled K-major atom
// layout, chunk[n][k]:
//   transpose == 0 : chunk[n0+n][k0+k] = W[row0+k][col0+n]   (forward, B = W^T)
//   transpose == 1 : chunk[n0+n][k0+k] = W[row0+n][col0+k]   (data gradient, B = W)
struct PackParams {
  nrc_pack_entry_t e[NRC_PACK_MAX_ENTRIES];
  void* ptrs[NRC_CHAIN_MAX_PTRS];
  uint8_t* packed;
};

__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackParams p) {
  const nrc_pack_entry_t& e = p.e[blockIdx.x];
  const float* W = static_cast<const float*>(p.ptrs[e.ptr]);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.packed + static_cast<size_t>(e.chunk) * kAtomBytes);
  const int nn = e.transpose ? e.nrows : e.ncols;   // destination rows
  const int nk = e.transpose ? e.ncols : e.nrows;   // destination k extent
  for (int idx = threadIdx.x; idx < nn * nk; idx += blockDim.x) {
    int n, k;
    float w;
    if (e.transpose) {   // source row-major [n][k]: k fastest for coalesced reads
      n = idx / nk; k = idx - n * nk;
      w = W[static_cast<size_t>(e.row0 + n) * e.ld + e.col0 + k];
    } else {             // source [k][n]: n fastest
      k = idx / nn; n = idx - k * nn;
      w = W[static_cast<size_t>(e.row0 + k) * e.ld + e.col0 + n];
    }
    const int dn = e.n0 + n, dk = e.k0 + k;
    dst[(atom_chunk_offset(dn, dk >> 3) >> 1) + (dk & 7)] = __float2bfloat16_rn(w);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradients dW[in,out] += X^T dY and db[out] += 1^T dY over all points, from the bf16 tile
// images the forward (X) and data-gradient (dY) chains saved.  Both operands are read MN-major
// straight from the [point][feature] atoms (no transpose anywhere).  One CTA = (layer, tile range);
// accumulators for every 128-feature pair of X atoms (+ one for the bias, A = ones) live in TMEM over
// the whole range and are flushed once with vector reductions into the fp32 gradient sinks.
struct WgradParams {
  nrc_wgrad_layer_t layers[NRC_WGRAD_MAX_LAYERS];
  void* ptrs[NRC_CHAIN_MAX_PTRS];
  int32_t num_tiles;
  int32_t tiles_per_cta;
};

constexpr int kWgradThreads = 192;   // warp 0 producer, warp 1 UMMA issuer, warps 2-5 flush
constexpr int kWgradRingAtoms = 12;

__global__ void __launch_bounds__(kWgradThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_base_s;

  const nrc_wgrad_layer_t& L = p.layers[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.num_tiles, t_begin + p.tiles_per_cta);
  const int nB = (L.n + 63) >> 6;                       // dY atoms
  const int nXp = (L.n_x_atoms + 1) & ~1;               // X atoms padded to whole pairs
  const int G = ((nB + 1) & ~1) + nXp;                  // atoms per stage (pairs stay adjacent)
  const int R = min(4, kWgradRingAtoms / G);            // stages
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ones_addr = base;                      // one atom of bf16 1.0
  const uint32_t ring_base = base + kAtomBytes;
  auto stage_addr = [&](int s) { return ring_base + static_cast<uint32_t>(s * G) * kAtomBytes; };
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };
  const uint32_t done_bar = bar0 + 8u * 8;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < kAtomBytes / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ones_addr + 4u * i), "r"(0x3F803F80u) : "memory");
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int n_pairs = nXp >> 1;
  const int Npad = L.n;                                  // multiple of 16

  if (t_begin < t_end) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>(nB + L.n_x_atoms) * kAtomBytes);
          const uint32_t sa = stage_addr(stage);
          const uint8_t* dy = static_cast<const uint8_t*>(p.ptrs[L.dy_ptr]);
          for (int a = 0; a < nB; ++a)
            bulk_g2s(sa + static_cast<uint32_t>(a) * kAtomBytes,
                     dy + (static_cast<size_t>(t) * L.dy_img_atoms + L.dy_atom0 + a) * kAtomBytes, kAtomBytes,
                     full_bar(stage));
          for (int a = 0; a < L.n_x_atoms; ++a) {
            const uint8_t* x = static_cast<const uint8_t*>(p.ptrs[L.x_ptr[a]]);
            bulk_g2s(sa + static_cast<uint32_t>(((nB + 1) & ~1) + a) * kAtomBytes,
                     x + (static_cast<size_t>(t) * L.x_img_atoms[a] + L.x_atom[a]) * kAtomBytes, kAtomBytes,
                     full_bar(stage));
          }
          if (++stage == R) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t idesc = make_idesc(128, Npad, 1, 1);
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage_addr(stage);
          const uint32_t b_addr = sa;
          const uint32_t x_addr = sa + static_cast<uint32_t>((nB + 1) & ~1) * kAtomBytes;
          const uint32_t first = (t == t_begin) ? 0u : 1u;
          for (int pr = 0; pr < n_pairs; ++pr)
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem_base + static_cast<uint32_t>(pr * Npad),
                        mnmajor_desc(x_addr + static_cast<uint32_t>(2 * pr) * kAtomBytes, k, kAtomBytes),
                        mnmajor_desc(b_addr, k, kAtomBytes), idesc, (first | (k > 0)) ? 1u : 0u);
          for (int k = 0; k < 8; ++k)   // bias gradient: ones^T dY
            umma_bf16(tmem_base + static_cast<uint32_t>(n_pairs * Npad), mnmajor_desc(ones_addr, k, 0),
                      mnmajor_desc(b_addr, k, kAtomBytes), idesc, (first | (k > 0)) ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (++stage == R) { stage = 0; phase ^= 1u; }
        }
        umma_commit(done_bar);
      }
    } else {
      // ------------------------------------------------------------------ flush
      const int quad = warp & 3;
      const int r = quad * 32 + lane;
      const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      for (int pr = 0; pr <= n_pairs; ++pr) {
        const bool is_bias = pr == n_pairs;
        const int atom = 2 * pr + (r >> 6);
        const int fr = r & 63;
        const bool row_ok = is_bias ? (r == 0) : (atom < L.n_x_atoms && fr < L.x_rows[atom]);
        const int krow = (!is_bias && atom < L.n_x_atoms) ? L.w_row0[atom] + fr : 0;
        for (int j0 = 0; j0 < Npad; j0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + t_lane + static_cast<uint32_t>(pr * Npad + j0), v);
          tmem_ld_wait();
          if (!row_ok) continue;
          for (int sgi = 0; sgi < L.n_seg; ++sgi) {
            const int c0 = L.seg_col0[sgi], nc = L.seg_ncols[sgi];
            const int pidx = is_bias ? L.seg_b_ptr[sgi] : L.seg_w_ptr[sgi];
            if (pidx < 0) continue;
            float* g = static_cast<float*>(p.ptrs[pidx]) + (is_bias ? 0 : static_cast<size_t>(krow) * nc);
            const int lo = max(j0, c0), hi = min(j0 + 16, c0 + nc);
            if (lo >= hi) continue;
            if (hi - lo == 16 && ((nc & 3) == 0) && (((lo - c0) & 3) == 0) &&
                ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
#pragma unroll
              for (int e = 0; e < 16; e += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + (lo - c0) + e),
                             "f"(__uint_as_float(v[e])), "f"(__uint_as_float(v[e + 1])),
                             "f"(__uint_as_float(v[e + 2])), "f"(__uint_as_float(v[e + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int j = j0 + e;
                if (j >= lo && j < hi) atomicAdd(g + (j - c0), __uint_as_float(v[e]));
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace nrc

// ================================================================================================
using namespace nrc;

static int32_t validate_program(const nrc_chain_program_t* prog, int32_t num_ptrs) {
  if (!prog || prog->num_ops < 1 || prog->num_ops > NRC_CHAIN_MAX_OPS) return NRC_E_INVALID_ARG;
  const int S = prog->slots_per_ctx;
  if (S < 1 || S > 7) return NRC_E_INVALID_ARG;
  auto ptr_ok = [&](int32_t i, bool optional) { return (optional && i < 0) || (i >= 0 && i < num_ptrs); };
  for (int i = 0; i < prog->num_ops; ++i) {
    const nrc_chain_op_t& op = prog->ops[i];
    switch (op.kind) {
      case NRC_OP_LOAD:
        if (!ptr_ok(op.ptr, true) || op.slot < 0 || (op.col0 & 7) || (op.npad & 7) || op.npad < op.ncols ||
            op.npad <= 0 || op.slot + ((op.col0 + op.npad + 63) >> 6) > S)
          return NRC_E_INVALID_ARG;
        break;
      case NRC_OP_GEMM:
        if (op.n < 16 || op.n > 128 || (op.n & 15) || op.n_atoms < 1 || op.n_atoms > NRC_CHAIN_MAX_ATOMS ||
            op.tmem_col < 0 || op.tmem_col + op.n > 256 || op.w_chunk < 0)
          return NRC_E_INVALID_ARG;
        for (int a = 0; a < op.n_atoms; ++a)
          if (op.a_slot[a] >= S || op.a_klen[a] < 16 || op.a_klen[a] > 64 || (op.a_klen[a] & 15)) return NRC_E_INVALID_ARG;
        break;
      case NRC_OP_EPI:
        if ((op.flags & NRC_EPI_DENSITY) && (!ptr_ok(op.out_ptr, false) || op.npad != 16)) return NRC_E_INVALID_ARG;
        if (op.npad <= 0 || (op.npad & 15) || op.ncols > op.npad || op.tmem_col < 0 || op.tmem_col + op.npad > 256 ||
            !ptr_ok(op.ptr, true) || !ptr_ok(op.out_ptr, true) || !ptr_ok(op.mask_ptr, true) ||
            (op.slot >= 0 && op.slot + ((op.npad + 63) >> 6) > S))
          return NRC_E_INVALID_ARG;
        break;
      case NRC_OP_GATHER:
        if (!ptr_ok(op.ptr, false) || !ptr_ok(op.out_ptr, true) || op.slot < 0 || op.slot >= S || op.ncols < 1 ||
            op.ncols > 32 || op.npad < op.ncols || (op.npad & 15) || op.npad > 32)
          return NRC_E_INVALID_ARG;
        break;
      case NRC_OP_SAVE:
        if (!ptr_ok(op.ptr, false) || op.slot < 0 || op.npad < 1 || op.slot + op.npad > S || op.col0 < 0 ||
            op.col0 + op.npad > op.img_atoms)
          return NRC_E_INVALID_ARG;
        break;
      default:
        return NRC_E_INVALID_ARG;
    }
  }
  return NRC_OK;
}

static int32_t chain_launch(void* stream, const nrc_chain_program_t* prog, void* const* d_ptrs, int32_t num_ptrs,
                            const void* d_weights_packed, int64_t num_rows, const EncDev* enc, float warp_c) {
  if (!d_ptrs || num_ptrs < 0 || num_ptrs > NRC_CHAIN_MAX_PTRS || num_rows < 0) return NRC_E_INVALID_ARG;
  const int32_t st = validate_program(prog, num_ptrs);
  if (st != NRC_OK) return st;
  if (num_rows == 0) return NRC_OK;
  static thread_local ChainParams hp;
  hp.prog = *prog;
  for (int i = 0; i < NRC_CHAIN_MAX_PTRS; ++i) hp.ptrs[i] = i < num_ptrs ? d_ptrs[i] : nullptr;
  hp.weights = static_cast<const uint8_t*>(d_weights_packed);
  hp.num_rows = num_rows;
  hp.num_tiles = static_cast<int32_t>((num_rows + 127) / 128);
  bool has_gather = false;
  for (int i = 0; i < prog->num_ops; ++i) has_gather = has_gather || prog->ops[i].kind == NRC_OP_GATHER;
  if (has_gather && !enc) return NRC_E_INVALID_ARG;
  if (enc) hp.enc = *enc;
  hp.warp_c = warp_c;
  const int S = prog->slots_per_ctx;
  const int max_atoms = (227 * 1024 - 2048) / kAtomBytes;   // 14
  int R = max_atoms - 2 * S;
  if (R < 1) return NRC_E_UNSUPPORTED;
  if (R > 4) R = 4;
  if (has_gather && R > 2) R = 2;   // queries: small weights, two CTAs per SM so gathers of one hide behind the other
  hp.ring_stages = R;
  const size_t smem = static_cast<size_t>(2 * S + R) * kAtomBytes + 1024;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024) != cudaSuccess ||
        cudaFuncSetAttribute(chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024) != cudaSuccess)
      return check_launch();
    attr_set = true;
  }
  const int pairs = (hp.num_tiles + 1) / 2;
  hp.ctx_tmem_cols = kCtxTmemCols;
  if (has_gather && smem <= 113 * 1024) {
    int need = 32;
    for (int i = 0; i < prog->num_ops; ++i)
      if (prog->ops[i].kind == NRC_OP_GEMM)
        while (need < prog->ops[i].tmem_col + prog->ops[i].n) need *= 2;
    if (need > 128) return NRC_E_UNSUPPORTED;
    hp.ctx_tmem_cols = need;
    const int grid = pairs < 2 * kNumSMs ? pairs : 2 * kNumSMs;
    chain_kernel<2><<<grid, kChainThreads, smem, static_cast<cudaStream_t>(stream)>>>(hp);
  } else {
    const int grid = pairs < kNumSMs ? pairs : kNumSMs;
    chain_kernel<1><<<grid, kChainThreads, smem, static_cast<cudaStream_t>(stream)>>>(hp);
  }
  return check_launch();
}

extern "C" int32_t nrc_chain_run(void* stream, const nrc_chain_program_t* prog, void* const* d_ptrs, int32_t num_ptrs,
                                 const void* d_weights_packed, int64_t num_rows) {
  return chain_launch(stream, prog, d_ptrs, num_ptrs, d_weights_packed, num_rows, nullptr, 0.f);
}

extern "C" int32_t nrc_chain_query(void* stream, const nrc_chain_program_t* prog, void* const* d_ptrs, int32_t num_ptrs,
                                   const void* d_weights_packed, int64_t num_rows, const nrc_encoding_t* enc,
                                   float warp_c) {
  EncDev d;
  const int32_t st = make_enc_dev(enc, d);
  if (st != NRC_OK) return st;
  if (d.L * d.F > 32 || d.F == 8) return NRC_E_UNSUPPORTED;
  return chain_launch(stream, prog, d_ptrs, num_ptrs, d_weights_packed, num_rows, &d, warp_c);
}

extern "C" int32_t nrc_chain_pack_weights(void* stream, const nrc_pack_entry_t* entries, int32_t num_entries,
                                          void* const* d_ptrs, int32_t num_ptrs, void* d_packed, int32_t num_chunks,
                                          int32_t keep_existing) {
  if (!entries || num_entries < 1 || num_entries > NRC_PACK_MAX_ENTRIES || !d_ptrs || num_ptrs < 1 ||
      num_ptrs > NRC_CHAIN_MAX_PTRS || !d_packed || num_chunks < 1)
    return NRC_E_INVALID_ARG;
  static thread_local PackParams hp;
  for (int i = 0; i < num_entries; ++i) {
    const nrc_pack_entry_t& e = entries[i];
    const int nn = e.transpose ? e.nrows : e.ncols, nk = e.transpose ? e.ncols : e.nrows;
    if (e.ptr < 0 || e.ptr >= num_ptrs || e.chunk < 0 || e.chunk >= num_chunks || e.n0 < 0 || e.k0 < 0 || nn < 1 ||
        nk < 1 || e.n0 + nn > 128 || e.k0 + nk > 64)
      return NRC_E_INVALID_ARG;
    hp.e[i] = e;
  }
  for (int i = 0; i < NRC_CHAIN_MAX_PTRS; ++i) hp.ptrs[i] = i < num_ptrs ? d_ptrs[i] : nullptr;
  hp.packed = static_cast<uint8_t*>(d_packed);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!keep_existing &&
      cudaMemsetAsync(d_packed, 0, static_cast<size_t>(num_chunks) * kAtomBytes, s) != cudaSuccess)
    return check_launch();
  pack_kernel<<<num_entries, 256, 0, s>>>(hp);
  return check_launch();
}

extern "C" int32_t nrc_chain_wgrad(void* stream, const nrc_wgrad_layer_t* layers, int32_t num_layers,
                                   void* const* d_ptrs, int32_t num_ptrs, int64_t num_rows) {
  if (!layers || num_layers < 1 || num_layers > NRC_WGRAD_MAX_LAYERS || !d_ptrs || num_ptrs < 1 ||
      num_ptrs > NRC_CHAIN_MAX_PTRS || num_rows < 0)
    return NRC_E_INVALID_ARG;
  if (num_rows == 0) return NRC_OK;
  static thread_local WgradParams hp;
  auto ok = [&](int32_t i) { return i >= 0 && i < num_ptrs; };
  for (int l = 0; l < num_layers; ++l) {
    const nrc_wgrad_layer_t& L = layers[l];
    if (L.n < 16 || L.n > 128 || (L.n & 15) || L.n_x_atoms < 1 || L.n_x_atoms > NRC_WGRAD_MAX_X_ATOMS || !ok(L.dy_ptr) ||
        L.n_seg < 1 || L.n_seg > NRC_WGRAD_MAX_SEGS)
      return NRC_E_INVALID_ARG;
    const int n_pairs = (L.n_x_atoms + 1) / 2;
    if ((n_pairs + 1) * L.n > 512) return NRC_E_UNSUPPORTED;
    for (int a = 0; a < L.n_x_atoms; ++a)
      if (!ok(L.x_ptr[a]) || L.x_rows[a] < 1 || L.x_rows[a] > 64) return NRC_E_INVALID_ARG;
    for (int s = 0; s < L.n_seg; ++s)
      if (L.seg_w_ptr[s] >= num_ptrs || L.seg_b_ptr[s] >= num_ptrs || L.seg_ncols[s] < 1) return NRC_E_INVALID_ARG;
    hp.layers[l] = L;
  }
  for (int i = 0; i < NRC_CHAIN_MAX_PTRS; ++i) hp.ptrs[i] = i < num_ptrs ? d_ptrs[i] : nullptr;
  hp.num_tiles = static_cast<int32_t>((num_rows + 127) / 128);
  int splits = (2 * kNumSMs + num_layers - 1) / num_layers;
  if (splits > hp.num_tiles) splits = hp.num_tiles;
  hp.tiles_per_cta = (hp.num_tiles + splits - 1) / splits;
  splits = (hp.num_tiles + hp.tiles_per_cta - 1) / hp.tiles_per_cta;
  static thread_local bool attr_set = false;
  const size_t smem = static_cast<size_t>(1 + kWgradRingAtoms) * kAtomBytes + 1024;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return check_launch();
    attr_set = true;
  }
  wgrad_kernel<<<dim3(splits, num_layers), kWgradThreads, smem, static_cast<cudaStream_t>(stream)>>>(hp);
  return check_launch();
}
