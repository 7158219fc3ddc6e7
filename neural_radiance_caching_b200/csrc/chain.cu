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
//   warps 2-9   epilogue / loader warps of context 0   (TMEM lane quadrant = warp % 4; the two warps of a
//   warps 10-17 epilogue / loader warps of context 1    quadrant take alternate 16-column chunks)
// Two contexts = two independent tiles in flight: while one context's warps drain their accumulator
// (TMEM -> registers -> bf16 -> shared memory), the tensor core runs the other context's layer.  (Staging each
// weight atom once for both tiles was measured: running the contexts in phase costs more than the halved ring
// traffic saves.)
#include <cuda_bf16.h>

#include <cstdlib>

#include "encode.cuh"
#include "nrc_common.cuh"
#include "tc05.cuh"

namespace nrc {
using namespace tc;

constexpr int kCtxThreads = 256;                   // loader / epilogue threads per tile context
constexpr int kChainThreads = 64 + 2 * kCtxThreads;
constexpr int kCtxTmemCols = 256;                  // TMEM columns per tile context
constexpr int kTailBytes = 3072;   // shared-memory tail: mbarriers (128 B), TMEM base (16 B), the program, staged biases

// The program as the kernel reads it: built once per CTA in shared memory from the launch parameters (pointers
// resolved, fields narrowed), because the per-tile interpreter touches it constantly and indexed constant-bank
// loads cost hundreds of cycles each.
struct __align__(8) DevOp {
  int8_t kind, slot;
  uint8_t flags, n_atoms;
  int16_t ncols, npad, tmem_col, n;
  int32_t ld;
  int16_t col0, mask_atom0, img_atoms, bias_off;   // bias_off: float index of the staged bias, -1 = read global
  int32_t w_chunk;
  float fparam;
  void* ptr;
  void* out;
  void* mask;
  uint8_t a_src[NRC_CHAIN_MAX_ATOMS];   // low nibble: slot, high nibble: K extent / 16
};
static_assert(sizeof(DevOp) == 64, "DevOp layout");

struct ChainParams {
  nrc_chain_program_t prog;
  void* ptrs[NRC_CHAIN_MAX_PTRS];
  const uint8_t* weights;
  int64_t num_rows;
  int32_t num_tiles;
  int32_t ring_stages;
  EncDev enc;       // hash-grid front end of GATHER ops (nrc_chain_query); unused otherwise
  float warp_c;
};

#ifdef NRC_CHAIN_TRACE
// Debug builds only (NRC_EXTRA_NVCC_FLAGS=-DNRC_CHAIN_TRACE): CTA 0 stamps clock64() at the end of every op of
// every tile it processes; tools/trace_chain.py prints the per-op timeline.
constexpr int kTraceTiles = 48, kTraceOps = 32;
__device__ long long g_chain_trace[2][kTraceTiles][kTraceOps];
__device__ long long g_chain_marks[16];
__device__ long long g_chain_mma[kTraceTiles][8][2];   // v2: UMMA issuer of CTA 0, per tile and GEMM group: operands ready, issued
#define TRACE_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 64) g_chain_marks[i] = clock64(); } while (0)
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TRACE_NS(i) do { if (blockIdx.x == 0 && threadIdx.x == 64) g_chain_marks[i] = global_ns(); } while (0)
#else
#define TRACE_NS(i) do {} while (0)
#define TRACE_MARK(i) do {} while (0)
#endif

// GATHER op: the two threads owning tile row r (one per half) contract point row0 + r and gather alternate levels
// of the multiresolution features (level_interp: bit-identical to nrc_encode_fwd); each writes its levels' bf16
// values into the row of the destination atom, half 0 also the zero padding up to npad.
template <int F>
__device__ __forceinline__ bool gather_row(const ChainParams& p, const DevOp& op, int64_t pt, bool valid,
                                           uint32_t slot_base, int r, int half) {
  constexpr int kMaxL = 32 / F > 8 ? 8 : 32 / F;
  bool inside = false;
  float xn[3] = {0.f, 0.f, 0.f};
  if (valid) {
    const float* m = static_cast<const float*>(op.ptr) + 3 * pt;
    const float x0 = __ldg(m), x1 = __ldg(m + 1), x2 = __ldg(m + 2);
    float z[3];
    contract_point(p.warp_c, x0, x1, x2, z[0], z[1], z[2]);
    normalise_point(p.enc, z, xn);
    inside = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) inside = inside && (z[a] > p.enc.b0[a]) && (z[a] < p.enc.b1[a]);
  }
  float* eo = (valid && op.out) ? static_cast<float*>(op.out) + pt * op.ncols : nullptr;
#pragma unroll
  for (int k = 0; k < kMaxL / 2; ++k) {
    const int l = 2 * k + half;
    if (l < p.enc.L) {
      float e[F];
#pragma unroll
      for (int f = 0; f < F; ++f) e[f] = 0.f;
      if (valid) {
        const Corners c = level_setup(p.enc.lv[l], xn);
        const FeatVec<F> v = level_interp<F>(p.enc.lv[l], c);
#pragma unroll
        for (int f = 0; f < F; ++f) e[f] = __fmul_rn(v.v[f], p.enc.scale);
        if (eo) {
#pragma unroll
          for (int f = 0; f < F; ++f) eo[l * F + f] = e[f];
        }
      }
      const int col = l * F;
      const uint32_t dst = slot_base + atom_chunk_offset(r, col >> 3) + static_cast<uint32_t>(col & 7) * 2u;
      if (F == 4) {
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(pack2_bf16(e[0], e[1])),
                     "r"(pack2_bf16(e[F > 2 ? 2 : 0], e[F > 2 ? 3 : 0])) : "memory");
      } else if (F == 2) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(pack2_bf16(e[0], e[F > 1 ? 1 : 0])) : "memory");
      } else {
        const unsigned short hbits = __bfloat16_as_ushort(__float2bfloat16_rn(e[0]));
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst), "h"(hbits) : "memory");
      }
    }
  }
  if (half == 0) {
    for (int col = op.ncols; col < op.npad; ++col) {
      const uint32_t dst = slot_base + atom_chunk_offset(r, col >> 3) + static_cast<uint32_t>(col & 7) * 2u;
      const unsigned short zero = 0;
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst), "h"(zero) : "memory");
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
  const float* bias_s;       // the same bias staged in shared memory, zero padded to npad (or nullptr)
  float* out_row;            // fp32 output row (already offset to col0) or nullptr
  const uint8_t* mask_tile;  // forward-activation image of this tile or nullptr
  uint32_t slot0_addr;       // shared-memory address of the first destination slot
  int ncols, npad, mask_atom0, r, half;
  int nparts = 2;           // warps sharing this thread's TMEM lane quadrant: warp `half` takes every nparts-th chunk
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

template <bool MASK>
__device__ __forceinline__ void epi_mask_load(const EpiArgs& a, int j0, uint32_t (&mw)[8]) {
  if (MASK) {
    const int mc = a.mask_atom0 * 64 + j0;
    const uint8_t* ma = a.mask_tile + static_cast<size_t>(mc >> 6) * kAtomBytes;
    const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(ma + atom_chunk_offset(a.r, (mc & 63) >> 3)));
    const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(ma + atom_chunk_offset(a.r, ((mc & 63) >> 3) + 1)));
    mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
    mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
  }
}

template <bool BIAS, bool FULL>
__device__ __forceinline__ void epi_bias_load(const EpiArgs& a, int j0, float (&b)[16]) {
  if (BIAS) {
    if (a.bias_s) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = *(reinterpret_cast<const float4*>(a.bias_s + j0) + q);
        b[4 * q] = t.x; b[4 * q + 1] = t.y; b[4 * q + 2] = t.z; b[4 * q + 3] = t.w;
      }
    } else if (FULL) {
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

// This thread's 16-column chunks (every other one: the quadrant's second warp takes the rest), kEpiBatch chunks
// per TMEM round trip: all their tcgen05.ld (and the mask loads) are issued before the one wait.
template <bool BIAS, bool RELU, bool MASK, bool FULL, int kEpiBatch = 1>
__device__ __forceinline__ void epi_run(const EpiArgs& a) {
  const int step = 16 * a.nparts;
  for (int j0 = 16 * a.half; j0 < a.npad; j0 += step * kEpiBatch) {
    uint32_t v[kEpiBatch][16], mw[kEpiBatch][8];
#pragma unroll
    for (int k = 0; k < kEpiBatch; ++k)
      if (j0 + step * k < a.npad) tmem_ld16(a.taddr + j0 + step * k, v[k]);
    TRACE_MARK(2);
#pragma unroll
    for (int k = 0; k < kEpiBatch; ++k)
      if (j0 + step * k < a.npad) epi_mask_load<MASK>(a, j0 + step * k, mw[k]);
    tmem_ld_wait();
    TRACE_MARK(3);
#pragma unroll
    for (int k = 0; k < kEpiBatch; ++k) {
      if (j0 + step * k < a.npad) {
        float b[16];
        epi_bias_load<BIAS, FULL>(a, j0 + step * k, b);
        epi_chunk<BIAS, RELU, MASK, FULL>(a, j0 + step * k, v[k], mw[k], b);
      }
    }
    TRACE_MARK(4);
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kChainThreads, 1) chain_kernel(const __grid_constant__ ChainParams p) {
  // dynamic shared memory only: [2S slot atoms][R ring atoms][tail: mbarriers, TMEM base, staged biases]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.prog.slots_per_ctx, R = p.ring_stages, nops = p.prog.num_ops;
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();   // swizzled atoms need 1024-byte alignment
  const uint32_t ring_base = base + 2u * S * kAtomBytes;
  uint8_t* tail = smem_raw + static_cast<size_t>(2 * S + R) * kAtomBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(tail + 128);
  DevOp* sops = reinterpret_cast<DevOp*>(tail + 144);
  float* sbias = reinterpret_cast<float*>(tail + 144 + sizeof(DevOp) * nops);
  const int bias_cap = (kTailBytes - 144 - static_cast<int>(sizeof(DevOp)) * nops) / 4;
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
      mbar_init(a_ready(c), kCtxThreads);
      mbar_init(acc_ready(c), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (threadIdx.x < nops) {
    const nrc_chain_op_t& o = p.prog.ops[threadIdx.x];
    DevOp d;
    d.kind = static_cast<int8_t>(o.kind); d.slot = static_cast<int8_t>(o.slot);
    d.flags = static_cast<uint8_t>(o.flags); d.n_atoms = static_cast<uint8_t>(o.n_atoms);
    d.ncols = static_cast<int16_t>(o.ncols); d.npad = static_cast<int16_t>(o.npad);
    d.tmem_col = static_cast<int16_t>(o.tmem_col); d.n = static_cast<int16_t>(o.n);
    d.ld = o.ld;
    d.col0 = static_cast<int16_t>(o.col0); d.mask_atom0 = static_cast<int16_t>(o.mask_atom0);
    d.img_atoms = static_cast<int16_t>(o.img_atoms);
    d.w_chunk = o.w_chunk; d.fparam = o.fparam;
    d.ptr = o.ptr >= 0 ? p.ptrs[o.ptr] : nullptr;
    d.out = o.out_ptr >= 0 ? p.ptrs[o.out_ptr] : nullptr;
    d.mask = o.mask_ptr >= 0 ? p.ptrs[o.mask_ptr] : nullptr;
#pragma unroll
    for (int a = 0; a < NRC_CHAIN_MAX_ATOMS; ++a) d.a_src[a] = static_cast<uint8_t>((o.a_slot[a] & 15) | ((o.a_klen[a] >> 4) << 4));
    // biases of the epilogues are staged behind the program, in op order, while they fit
    int cur = 0, off = -1;
    for (int k = 0; k <= static_cast<int>(threadIdx.x); ++k) {
      const nrc_chain_op_t& e = p.prog.ops[k];
      if (e.kind != NRC_OP_EPI || e.ptr < 0 || e.mask_ptr >= 0 || (e.flags & NRC_EPI_DENSITY)) continue;
      if (cur + e.npad > bias_cap) break;
      if (k == static_cast<int>(threadIdx.x)) off = cur;
      cur += e.npad;
    }
    d.bias_off = static_cast<int16_t>(off);
    sops[threadIdx.x] = d;
  }
  __syncthreads();
  for (int i = 0; i < nops; ++i) {
    const DevOp& op = sops[i];
    if (op.bias_off < 0) continue;
    const float* b = static_cast<const float*>(op.ptr);
    for (int k = threadIdx.x; k < op.npad; k += kChainThreads) sbias[op.bias_off + k] = k < op.ncols ? __ldg(b + k) : 0.f;
  }
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
          if (sops[i].kind != NRC_OP_GEMM) { ++i; continue; }
          int j = i;
          while (j < nops && sops[j].kind == NRC_OP_GEMM) ++j;
          for (int c = 0; c < 2; ++c) {
            if (2 * q + c >= p.num_tiles) continue;
            for (int o = i; o < j; ++o) {
              const DevOp& op = sops[o];
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
          if (sops[i].kind != NRC_OP_GEMM) { ++i; continue; }
          int j = i;
          while (j < nops && sops[j].kind == NRC_OP_GEMM) ++j;
          for (int c = 0; c < 2; ++c) {
            if (2 * q + c >= p.num_tiles) continue;
            mbar_wait(a_ready(c), a_par[c]);
            a_par[c] ^= 1u;
            tc_fence_after();
            for (int o = i; o < j; ++o) {
              const DevOp& op = sops[o];
              const uint32_t idesc = make_idesc(128, op.n, 0, 0);
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(c * kCtxTmemCols + op.tmem_col);
              for (int a = 0; a < op.n_atoms; ++a) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a_addr = slot_addr(c, op.a_src[a] & 15);
                const uint32_t b_addr = ring_base + static_cast<uint32_t>(stage) * kAtomBytes;
                const int nk = op.a_src[a] >> 4;
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
    const int c = (warp - 2) >> 3;                 // context
    const int wg_tid = threadIdx.x - 64 - kCtxThreads * c;
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int half = ((warp - 2) >> 2) & 1;        // which of the quadrant's two warps: alternate 16-column chunks
    const int r = quad * 32 + lane;                // tile row owned in epilogues
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    uint32_t acc_par = 0;
    bool store_pending = false;
    bool inside_reg = false;   // bbox mask of the point this thread gathered (GATHER -> NRC_EPI_DENSITY)

    auto guard_slots = [&]() {  // before overwriting slots a bulk store may still be reading
      if (store_pending) {
        if (wg_tid == 0) bulk_wait_read0();
        named_barrier_sync(1 + c, kCtxThreads);
        store_pending = false;
      }
    };

    for (int q = blockIdx.x; q < num_pairs; q += gridDim.x) {
      const int tile = 2 * q + c;
      if (tile >= p.num_tiles) continue;
      const int64_t row0 = static_cast<int64_t>(tile) * 128;
#ifdef NRC_CHAIN_TRACE
      const int trace_it = (q - blockIdx.x) / gridDim.x;
      const bool tracing = blockIdx.x == 0 && wg_tid == 0 && trace_it < kTraceTiles;
      if (tracing) g_chain_trace[c][trace_it][kTraceOps - 1] = clock64();
#endif
      for (int i = 0; i < nops;) {
        const DevOp& op = sops[i];
        if (op.kind == NRC_OP_GEMM) {
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(a_ready(c));
          while (i < nops && sops[i].kind == NRC_OP_GEMM) ++i;
          mbar_wait(acc_ready(c), acc_par);
          acc_par ^= 1u;
          tc_fence_after();
#ifdef NRC_CHAIN_TRACE
          if (tracing) g_chain_trace[c][trace_it][i - 1] = clock64();
#endif
          continue;
        }
        if (op.kind == NRC_OP_GATHER) {
          guard_slots();
          const int64_t pt = row0 + r;
          const bool valid = pt < p.num_rows;
          const uint32_t sb = slot_addr(c, op.slot);
          switch (p.enc.F) {
            case 1: inside_reg = gather_row<1>(p, op, pt, valid, sb, r, half); break;
            case 2: inside_reg = gather_row<2>(p, op, pt, valid, sb, r, half); break;
            default: inside_reg = gather_row<4>(p, op, pt, valid, sb, r, half); break;
          }
        } else if (op.kind == NRC_OP_LOAD) {
          guard_slots();
          const float* src = static_cast<const float*>(op.ptr);
          const int nch = op.npad >> 3;
          const bool vec = src && (op.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
          // six independent 32-byte loads in flight per thread before the first conversion
          constexpr int kU = 4;
          const int total = 128 * nch;
          for (int item0 = wg_tid; item0 < total; item0 += kCtxThreads * kU) {
            float v[kU][8];
#pragma unroll
            for (int q = 0; q < kU; ++q) {
              const int item = item0 + q * kCtxThreads;
              const int rr = item / nch;
              const int col = (item - rr * nch) * 8;
#pragma unroll
              for (int e = 0; e < 8; ++e) v[q][e] = 0.f;
              if (item < total && src && row0 + rr < p.num_rows && col < op.ncols) {
                const float* s = src + (row0 + rr) * op.ld + col;
                if (vec && col + 8 <= op.ncols) {
                  const float4 x0 = __ldg(reinterpret_cast<const float4*>(s));
                  const float4 x1 = __ldg(reinterpret_cast<const float4*>(s) + 1);
                  v[q][0] = x0.x; v[q][1] = x0.y; v[q][2] = x0.z; v[q][3] = x0.w;
                  v[q][4] = x1.x; v[q][5] = x1.y; v[q][6] = x1.z; v[q][7] = x1.w;
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e)
                    if (col + e < op.ncols) v[q][e] = __ldg(s + e);
                }
              }
            }
#pragma unroll
            for (int q = 0; q < kU; ++q) {
              const int item = item0 + q * kCtxThreads;
              if (item >= total) continue;
              const int rr = item / nch;
              const int dcol = op.col0 + (item - rr * nch) * 8;
              const uint32_t dst = slot_addr(c, op.slot + (dcol >> 6)) + atom_chunk_offset(rr, (dcol & 63) >> 3);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack2_bf16(v[q][0], v[q][1])),
                           "r"(pack2_bf16(v[q][2], v[q][3])), "r"(pack2_bf16(v[q][4], v[q][5])),
                           "r"(pack2_bf16(v[q][6], v[q][7]))
                           : "memory");
            }
          }
        } else if (op.kind == NRC_OP_SAVE) {
          fence_proxy_async_smem();
          named_barrier_sync(1 + c, kCtxThreads);
          if (wg_tid == 0) {
            uint8_t* img = static_cast<uint8_t*>(op.ptr);
            for (int a = 0; a < op.npad; ++a)
              bulk_s2g(img + (static_cast<size_t>(tile) * op.img_atoms + op.col0 + a) * kAtomBytes,
                       slot_addr(c, op.slot + a), kAtomBytes);
            bulk_commit();
          }
          store_pending = true;
        } else if (op.kind == NRC_OP_EPI && (op.flags & NRC_EPI_DENSITY)) {
          // density head: column 0 -> safe_exp(raw + density_bias) masked to the bbox; columns 1..3 -> grad_pred
          if (half == 0) {
          uint32_t v[16];
          tmem_ld16(tmem_base + t_lane + static_cast<uint32_t>(c * kCtxTmemCols + op.tmem_col), v);
          tmem_ld_wait();
          const int64_t pt = row0 + r;
          if (pt < p.num_rows) {
            const float* bias = static_cast<const float*>(op.ptr);
            const float raw = __uint_as_float(v[0]) + (bias ? __ldg(bias) : 0.f);
            static_cast<float*>(op.out)[pt] = inside_reg ? safe_exp(raw + op.fparam) : 0.f;
            if (op.mask) {
              float* gp = static_cast<float*>(op.mask) + 3 * pt;
#pragma unroll
              for (int j = 0; j < 3; ++j) gp[j] = __uint_as_float(v[1 + j]) + (bias ? __ldg(bias + 1 + j) : 0.f);
            }
          }
          }
        } else {  // NRC_OP_EPI
          TRACE_MARK(0);
          if (op.slot >= 0) guard_slots();
          EpiArgs a;
          a.bias = static_cast<const float*>(op.ptr);
          a.bias_s = op.bias_off >= 0 ? sbias + op.bias_off : nullptr;
          float* out = static_cast<float*>(op.out);
          a.out_row = (out && row0 + r < p.num_rows) ? out + (row0 + r) * op.ld + op.col0 : nullptr;
          a.out_vec = out && (op.ld % 4 == 0) && (op.col0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
          a.accum = (op.flags & NRC_EPI_OUT_ACCUMULATE) != 0;
          a.mask_tile = op.mask ? static_cast<const uint8_t*>(op.mask) + static_cast<size_t>(tile) * op.img_atoms * kAtomBytes
                                : nullptr;
          a.mask_atom0 = op.mask_atom0;
          a.taddr = tmem_base + t_lane + static_cast<uint32_t>(c * kCtxTmemCols + op.tmem_col);
          a.has_slot = op.slot >= 0;
          a.slot0_addr = a.has_slot ? slot_addr(c, op.slot) : 0u;
          a.ncols = op.ncols; a.npad = op.npad; a.r = r; a.half = half;
          const bool full = (op.ncols == op.npad) && (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0);
          const bool relu = (op.flags & NRC_EPI_RELU) != 0;
          TRACE_MARK(1);
          if (a.mask_tile) {
            if (full) epi_run<false, false, true, true>(a); else epi_run<false, false, true, false>(a);
          } else if (a.bias) {
            if (relu) { if (full) epi_run<true, true, false, true>(a); else epi_run<true, true, false, false>(a); }
            else      { if (full) epi_run<true, false, false, true>(a); else epi_run<true, false, false, false>(a); }
          } else {
            if (relu) { if (full) epi_run<false, true, false, true>(a); else epi_run<false, true, false, false>(a); }
            else      { if (full) epi_run<false, false, false, true>(a); else epi_run<false, false, false, false>(a); }
          }
          TRACE_MARK(5);
        }
#ifdef NRC_CHAIN_TRACE
        if (tracing) g_chain_trace[c][trace_it][i] = clock64();
#endif
        ++i;
      }
    }
    if (wg_tid == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// Fast epilogue of the common case - the accumulator becomes the next layer's bf16 operand and nothing else: TMEM ->
// (+ staged bias) -> bf16 pairs -> ReLU / ReLU mask applied on the PACKED pairs (max(.,0) commutes with the rounding;
// the mask is an AND) -> two 16-byte shared-memory stores per 16 columns.  ~45 instructions per chunk instead of ~80.
// sig_bar != 0: the NEXT layer's UMMAs were split by K atom (chain2_build): every time this context has finished an
// operand atom (64 columns, all 128 rows) it tells the issuer - proxy fence, the context's named barrier, one
// arrival per CTA on the issuing CTA's operand barrier - so the k-steps that read atom 0 run while atom 1 is still
// being drained from tensor memory.  (16 * part + k * 16 * nparts crosses a 64-column boundary at the same k for
// every part: the barrier count is uniform over the context.)
template <bool MASK>
__device__ __forceinline__ void epi_slot_fast(uint32_t taddr, const float* bias_s, bool relu, uint32_t slot0_addr, int r, int npad,
                                              int part, int nparts, const uint8_t* mask_tile, int mask_atom0,
                                              int sig_bar = 0, int sig_threads = 0, bool sig_t0 = false, uint32_t sig_remote = 0u) {
  // (two chunks per tensor-memory round trip were measured: the second 16-register buffer spills at the 96-register
  // cap of 18 warps and the SurfaceLightField forward went from 34.6 to 49 us)
  const bool has_bias = bias_s != nullptr;
  const uint32_t row_off = static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128);
  const uint32_t rx = static_cast<uint32_t>(r & 7);
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  for (int j0 = 16 * part; j0 < npad; j0 += 16 * nparts) {
    uint32_t v[16], mw[8];
    tmem_ld16(taddr + j0, v);
    if (MASK) {
      const int mc = mask_atom0 * 64 + j0;
      const uint8_t* ma = mask_tile + static_cast<size_t>(mc >> 6) * kAtomBytes + row_off;
      const uint32_t c8 = static_cast<uint32_t>((mc & 63) >> 3);
      const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(ma + ((c8 ^ rx) << 4)));
      const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(ma + (((c8 + 1) ^ rx) << 4)));
      mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
      mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
    }
    const uint32_t c8 = static_cast<uint32_t>((j0 & 63) >> 3);
    const uint32_t d = slot0_addr + static_cast<uint32_t>(j0 >> 6) * kAtomBytes + row_off;
    tmem_ld_wait();
    uint32_t o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float x0 = __uint_as_float(v[2 * e]), x1 = __uint_as_float(v[2 * e + 1]);
      if (has_bias) {
        const float2 bb = *reinterpret_cast<const float2*>(bias_s + j0 + 2 * e);
        x0 += bb.x; x1 += bb.y;
      }
      __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
      if (relu) h = __hmax2(h, zero2);
      uint32_t w = *reinterpret_cast<uint32_t*>(&h);
      if (MASK) {
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&mw[e]);
        w &= __hgt2_mask(a, zero2);
      }
      o[e] = w;
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + ((c8 ^ rx) << 4)), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                 "r"(o[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + (((c8 + 1) ^ rx) << 4)), "r"(o[4]), "r"(o[5]),
                 "r"(o[6]), "r"(o[7]) : "memory");
    if (sig_bar && (((j0 + 16 * nparts) ^ j0) >> 6)) {
      fence_proxy_async_smem();
      tc_fence_before();
      named_barrier_sync(sig_bar, sig_threads);
      // one operand barrier per atom index: an arrival may only run ONE phase ahead of the issuer's wait (two
      // completed phases look like none to a parity wait), and atom k+1 of this layer is announced before the issuer
      // has necessarily consumed atom k
      const uint32_t stage = static_cast<uint32_t>(j0 >> 6);
      if (sig_t0) mbar_arrive_cluster(sig_remote + (stage ? 32u + 16u * stage : 0u));
    }
  }
}

// Epilogue with an fp32 ROW output (head results, input gradients) staged through a free shared-memory slot: a thread
// owns one tile row in tensor memory, so direct stores put 32 different 128-byte lines behind every store instruction
// (measured ~4 cycles per 16 bytes per SM: 5-9.5 k cycles for a 72-column gradient).  Here 32 columns at a time go
// TMEM -> registers -> (+bias, ReLU) -> an XOR-swizzled [128 rows][32 floats] staging tile -> coalesced 16-byte
// stores (eight consecutive lanes write 128 contiguous bytes of one output row).  Optionally the same values are also
// written as bf16 into the destination atom slots.
__device__ __forceinline__ void epi_staged(uint32_t taddr, const float* bias_s, bool relu, uint32_t stage_addr, int r, int part,
                                           int ncols, int npad, float* out, int64_t row0, int64_t num_rows, int ld,
                                           bool accum, bool has_slot, uint32_t slot0_addr, int bar_id, int nthreads,
                                           int tid) {
  const uint32_t rx = static_cast<uint32_t>(r & 7);
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int g0 = 0; g0 < npad; g0 += 32) {
    const int j0 = g0 + 16 * part;
    if (part < 2 && j0 < npad) {
      uint32_t v[16];
      tmem_ld16(taddr + j0, v);
      tmem_ld_wait();
      float x[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float t = __uint_as_float(v[e]);
        if (bias_s) t += bias_s[j0 + e];
        if (relu) t = fmaxf(t, 0.f);
        x[e] = t;
      }
      const uint32_t srow = stage_addr + static_cast<uint32_t>(r) * 128u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t c = static_cast<uint32_t>(4 * part + q);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((c ^ rx) << 4)), "f"(x[4 * q]), "f"(x[4 * q + 1]),
                     "f"(x[4 * q + 2]), "f"(x[4 * q + 3]) : "memory");
      }
      if (has_slot) {
        const uint32_t row_off = static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128);
        const uint32_t c8 = static_cast<uint32_t>((j0 & 63) >> 3);
        const uint32_t d = slot0_addr + static_cast<uint32_t>(j0 >> 6) * kAtomBytes + row_off;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + ((c8 ^ rx) << 4)), "r"(pack2_bf16(x[0], x[1])),
                     "r"(pack2_bf16(x[2], x[3])), "r"(pack2_bf16(x[4], x[5])), "r"(pack2_bf16(x[6], x[7])) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + (((c8 + 1) ^ rx) << 4)), "r"(pack2_bf16(x[8], x[9])),
                     "r"(pack2_bf16(x[10], x[11])), "r"(pack2_bf16(x[12], x[13])), "r"(pack2_bf16(x[14], x[15])) : "memory");
      }
    }
    named_barrier_sync(bar_id, nthreads);
    for (int piece = tid; piece < 1024; piece += nthreads) {
      const int row = piece >> 3, c = piece & 7;
      const int col = g0 + 4 * c;
      if (col >= ncols || row0 + row >= num_rows) continue;
      float4 w;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w)
                   : "r"(stage_addr + static_cast<uint32_t>(row) * 128u + ((static_cast<uint32_t>(c) ^ static_cast<uint32_t>(row & 7)) << 4)));
      float* o = out + (row0 + row) * ld + col;
      if (vec && col + 4 <= ncols) {
        if (accum) {
          const float4 old = *reinterpret_cast<const float4*>(o);
          w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
        }
        *reinterpret_cast<float4*>(o) = w;
      } else {
        const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < ncols) o[e] = accum ? o[e] + ww[e] : ww[e];
      }
    }
    if (g0 + 32 < npad) named_barrier_sync(bar_id, nthreads);   // the staging tile is rewritten by the next round
  }
}

// ------------------------------------------------------------------------------------------------
// The CTA-PAIR kernel: stack weights RESIDENT in shared memory, up to three PROGRAMS per launch.
//   * a cluster of two CTAs executes one 256-row tile (128 rows per CTA) with tcgen05.mma.cta_group::2: the B operand
//     (weights) is split by N between the two CTAs, so each SM holds HALF of every layer's weights - the whole
//     SurfaceLightField stack (228 KB of bf16 operands) fits beside the activations and is loaded ONCE per CTA
//     instead of being streamed through a ring for every tile (the streaming kernel above moves 256 KB of weights per
//     128-row tile through a 2-4 stage ring: its GEMMs wait on L2, not on the tensor core);
//   * no ring, no producer: the UMMA issuer never waits for operands after the first tile; a tile's layer costs one
//     cluster-scope barrier round trip + the MMAs + the epilogue;
//   * one launch runs several independent programs (the integrated-BRDF, SurfaceLightField and EnvMap stacks of the
//     forward pass; the SurfaceLightField and integrated-BRDF data gradients of the backward pass): the (program, tile)
//     work items of all of them are dealt to the CTA pairs in contiguous, cost-balanced ranges, so that at 32 768
//     points (128 pair tiles per program on 74 pairs) the launch is ONE balanced wave instead of three launches of
//     one-and-three-quarter waves each with its own prologue; a pair that crosses a program boundary reloads its
//     weights (2-3 us) behind a cluster barrier.
// Programs, weight images, tile images and the op semantics are those of the streaming kernel (a 256-row pair tile is
// the two consecutive 128-row tiles 2t and 2t+1 of every tile image).
constexpr uint8_t kFlagSplit = 0x80;   // DevOp.flags, set by chain2_build: EPI = announce every finished operand atom; GEMM = group announced that way
constexpr int kTail2Bytes = 6144;   // shared-memory head: mbarriers (128 B), TMEM base (16 B), program, UMMA list, staged biases
constexpr int kMaxMma = 96;     // UMMA instructions of one program (precomputed descriptor list in shared memory)
constexpr int kMaxWcopy = 48;   // bulk copies that make a program's weights resident (one per K atom of every GEMM)
constexpr int kMaxBias = 12;    // epilogue biases staged in shared memory
constexpr int kMaxProgs = NRC_CHAIN_MAX_PROGRAMS;    // programs per launch
constexpr int kMaxPairs = 96;   // CTA pairs per launch (SM count / 2)

// Everything the kernel needs about a program, RESOLVED ON THE HOST (pointers, bias offsets, the UMMA descriptor list,
// the list of weight copies) and copied from the parameter space to shared memory with warp-uniform constant loads.
// Decoding the program in the kernel cost 4-10 us per launch (thread-indexed reads of the parameter space serialise;
// measured with the trace build: `decode` + `prologue` marks of tools/trace_chain.py), more than the tile work of the
// small stacks.
struct __align__(16) Chain2Plan {
  DevOp ops[NRC_CHAIN_MAX_OPS];
  uint4 mma[kMaxMma];        // x: A descriptor low word, address relative to the CTA's dynamic shared memory (context 0);
                             // y: B descriptor low word (same); z: accumulator column | accumulate << 31; w: idesc
  uint4 wcopy[kMaxWcopy];    // x: destination byte offset in the resident region, y: source byte offset inside the packed
                             // weights of CTA rank 0, z: bytes, w: extra source offset of CTA rank 1
  uint4 bias[kMaxBias];      // x, y: source pointer (lo, hi), z: valid floats | padded floats << 16, w: destination float offset
  int32_t n_ops, n_mma, n_wcopy, n_bias;
};
static_assert(sizeof(Chain2Plan) % 16 == 0, "plan is copied as 16-byte words");
struct Chain2Prog {
  Chain2Plan plan;
  const uint8_t* weights;
  int32_t w_bytes;         // resident weight bytes per CTA (multiple of 1024)
  int32_t slots_per_ctx;
  int32_t nctx;            // tile contexts: 2 (8 + 8 epilogue warps) when two fit, else 1 (16 epilogue warps)
  int32_t pad;
};
struct Chain2Params {
  Chain2Prog prog[kMaxProgs];
  int64_t num_rows;
  int32_t num_tiles;     // 128-row tiles
  int32_t num_ptiles;    // 256-row pair tiles
  int32_t n_progs;
  int32_t pad;
  uint16_t seg_t0[kMaxProgs][kMaxPairs];   // per program and CTA pair: first super tile (nctx pair tiles) ...
  uint16_t seg_n[kMaxProgs][kMaxPairs];    // ... and how many
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kChainThreads, 1)
chain2_kernel(const __grid_constant__ Chain2Params p) {
  // dynamic shared memory: [head: mbarriers, TMEM base, program, UMMA list, staged biases][resident weights][nctx * S slot atoms]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TRACE_NS(14);
  TRACE_MARK(8);
  constexpr int kThreads = kChainThreads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smem_raw + 128);
  DevOp* sops = reinterpret_cast<DevOp*>(smem_raw + 144);
  uint4* smma = reinterpret_cast<uint4*>(smem_raw + 144 + sizeof(DevOp) * NRC_CHAIN_MAX_OPS);
  float* sbias = reinterpret_cast<float*>(smem_raw + 144 + sizeof(DevOp) * NRC_CHAIN_MAX_OPS + sizeof(uint4) * kMaxMma);
  const uint32_t w_base = base + kTail2Bytes;
  // barriers: 0 weights resident, 2+c operands of context c ready (count 2: one arrival per CTA, used in the
  // issuing CTA only), 4+c accumulator of context c ready (multicast commit), 6+c LOADIMG bulk copies of context c landed
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t w_ready = bar0;
  auto a_ready = [&](int c) { return bar0 + 8u * (2 + c); };
  // operand atoms 1 and 2 of a split group (see chain2_build): barriers 8+c and 10+c, i.e. a_ready(c) + 48 / + 64
  auto a_stage = [&](int c, uint32_t stage) { return bar0 + 8u * (2 + c) + (stage ? 32u + 16u * stage : 0u); };
  auto acc_ready = [&](int c) { return bar0 + 8u * (4 + c); };
  auto img_ready = [&](int c) { return bar0 + 8u * (6 + c); };

  bool first = true;
  uint32_t tmem_base = 0;
  for (int k = 0; k < p.n_progs; ++k) {
    const int n_super = p.seg_n[k][pair];
    if (n_super == 0) continue;            // the same for both CTAs of the pair
    const int q0 = p.seg_t0[k][pair];
    const Chain2Prog& pg = p.prog[k];
    const int S = pg.slots_per_ctx, nops = pg.plan.n_ops, nctx = pg.nctx;
    const int ctxT = 2 * kCtxThreads / nctx;   // loader / epilogue threads per tile context: 16 warps (one context) or 8
    const int parts = ctxT / 128;              // warps per TMEM lane quadrant
    const uint32_t slot_base = w_base + static_cast<uint32_t>(pg.w_bytes);
    auto slot_addr = [&](int ctx, int s) { return slot_base + static_cast<uint32_t>(ctx * S + s) * kAtomBytes; };

    // ------------------------------------------------------------------ program set-up, every role in parallel:
    // warp 0 allocates tensor memory (first program), warp 1 (re)initialises the tile barriers, warp 2 arms the weight
    // barrier and issues the bulk copies that make this CTA's half of every weight atom resident (rows [rank n/2,
    // (rank+1) n/2) of a K-major chunk are the contiguous bytes [rank n/2 * 128, ...) of the swizzled atom), warps 3-6
    // stage the epilogue biases, the rest copy the program + UMMA list out of the parameter space.
    if (!first) {
      tc_fence_before();
      cluster_sync_all();   // both CTAs drained the previous program: its weights, slots and barriers can be reused
      tc_fence_after();
    }
    if (warp == 0) {
      if (first) tmem_alloc2(smem_u32(&tmem_base_s), 512);
    } else if (warp == 1) {
      if (lane == 0) {
        for (int c = 0; c < 2; ++c) {
          if (!first) {
            mbar_inval(a_ready(c)); mbar_inval(acc_ready(c)); mbar_inval(img_ready(c));
            mbar_inval(a_stage(c, 1)); mbar_inval(a_stage(c, 2));
          }
          mbar_init(a_ready(c), 2);
          mbar_init(a_stage(c, 1), 2);
          mbar_init(a_stage(c, 2), 2);
          mbar_init(acc_ready(c), 1);
          mbar_init(img_ready(c), 1);
        }
        fence_barrier_init();
      }
    } else if (warp == 2) {
      if (lane == 0) {
        if (!first) mbar_inval(w_ready);
        mbar_init(w_ready, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(w_ready, static_cast<uint32_t>(pg.w_bytes));
      }
      __syncwarp();
      for (int i = lane; i < pg.plan.n_wcopy; i += 32) {
        // every CTA pair starts at another entry: 74 CTAs otherwise ask the same L2 lines at the same moment
        const uint4 e = pg.plan.wcopy[(i + pair * 5) % pg.plan.n_wcopy];
        bulk_g2s(w_base + e.x, pg.weights + e.y + rank * e.w, e.z, w_ready);
      }
    } else if (warp < 7) {
      // biases of the epilogues (<= 256 floats each): thread t of these 128 takes elements t and t + 128 of every bias,
      // eight loads in flight
      const int t = threadIdx.x - 96;
      for (int b0 = 0; b0 < pg.plan.n_bias; b0 += 4) {
        float v[4][2];
        uint32_t dst[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            dst[u][h] = 0xFFFFFFFFu;
            v[u][h] = 0.f;
            if (b0 + u < pg.plan.n_bias) {
              const uint4 e = pg.plan.bias[b0 + u];
              const float* bsrc = reinterpret_cast<const float*>(static_cast<uintptr_t>(e.x) | (static_cast<uintptr_t>(e.y) << 32));
              const int nvalid = static_cast<int>(e.z & 0xFFFFu), npad = static_cast<int>(e.z >> 16);
              const int kk = t + 128 * h;
              if (kk < npad) dst[u][h] = e.w + static_cast<uint32_t>(kk);
              if (kk < nvalid) v[u][h] = __ldg(bsrc + kk);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (dst[u][h] != 0xFFFFFFFFu) sbias[dst[u][h]] = v[u][h];
      }
    } else {
      constexpr int kWords = static_cast<int>((sizeof(DevOp) * NRC_CHAIN_MAX_OPS + sizeof(uint4) * kMaxMma) / 16);
      constexpr int kCopyWarps = kThreads / 32 - 7;
      constexpr int kPerWarp = (kWords + kCopyWarps - 1) / kCopyWarps;
      const uint4* src = reinterpret_cast<const uint4*>(&pg.plan);
      uint4* dst = reinterpret_cast<uint4*>(sops);
      const int w0 = (warp - 7) * kPerWarp;
#pragma unroll
      for (int j = 0; j < kPerWarp; ++j) {
        const int i = w0 + j;                 // warp-uniform index: one constant load serves the warp
        if (i < kWords) {
          const uint4 v = src[i];
          if (lane == (j & 31)) dst[i] = v;
        }
      }
    }
    __syncthreads();
    if (first) TRACE_MARK(9);
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated, program staged
    tc_fence_after();
    if (first) TRACE_MARK(10);
    tmem_base = tmem_base_s;

    if (warp == 1) {
      // ===================================================================== UMMA issuer (even CTA of the pair)
      if (rank == 0 && lane == 0) {
        uint32_t a_par[2][3] = {{0u, 0u, 0u}, {0u, 0u, 0u}};
        for (int q = q0; q < q0 + n_super; ++q) {
#ifdef NRC_CHAIN_TRACE
          const int trace_it = q - q0;
          int trace_g = 0;
#endif
          for (int i = 0; i < nops;) {
            if (sops[i].kind != NRC_OP_GEMM) { ++i; continue; }
            int j = i;
            while (j < nops && sops[j].kind == NRC_OP_GEMM) ++j;
            for (int c = 0; c < nctx; ++c) {
              if (nctx * q + c >= p.num_ptiles) continue;
              {
                const int e0 = sops[i].ld, e1 = sops[j - 1].ld + sops[j - 1].col0;
                const uint32_t d0 = tmem_base + static_cast<uint32_t>(c * kCtxTmemCols);
                const uint32_t a_off = static_cast<uint32_t>(c * S) * (kAtomBytes >> 4) + (base >> 4);
                const uint32_t b_off = base >> 4;
                constexpr uint64_t kDescHi = static_cast<uint64_t>(64u | (1u << 14) | (2u << 29)) << 32;   // SBO 1024, v1, SW128
                // The list is a sequence of SEGMENTS: the first instruction of a segment names the operand barrier to
                // wait for (bits 29-30: atom index + 1) and the segment's length (bits 16-23); the instructions of a
                // segment are issued by a tight loop with no barrier code in it (with the wait inside the loop the
                // compiler stopped unrolling it and the issue rate dropped from ~120 to ~160 cycles per UMMA).
                for (int e = e0; e < e1;) {
                  uint4 m = smma[e];
                  const uint32_t ws = (m.z >> 29) & 3u;
                  const int seg_end = e + static_cast<int>((m.z >> 16) & 0xFFu);
                  if (ws) {
                    mbar_wait_cluster(a_stage(c, ws - 1u), a_par[c][ws - 1u]);
                    a_par[c][ws - 1u] ^= 1u;
                    tc_fence_after();
                  }
#ifdef NRC_CHAIN_TRACE
                  if (e == e0 && blockIdx.x == 0 && first && c == 0 && trace_it < kTraceTiles && trace_g < 8) g_chain_mma[trace_it][trace_g][0] = clock64();
#endif
                  const uint4* seg = smma + e;
                  const int n_seg = seg_end - e;
#pragma unroll 4
                  for (int t = 0; t < n_seg; ++t) {
                    const uint4 cur = m;
                    if (t + 1 < n_seg) m = seg[t + 1];
                    umma2_bf16(d0 + (cur.z & 0x1FFu), kDescHi | (cur.x + a_off), kDescHi | (cur.y + b_off), cur.w, cur.z >> 31);
                  }
                  e = seg_end;
                }
              }
              umma2_commit_mc(acc_ready(c), 3);
#ifdef NRC_CHAIN_TRACE
              if (blockIdx.x == 0 && first && c == 0 && trace_it < kTraceTiles && trace_g < 8) g_chain_mma[trace_it][trace_g][1] = clock64();
              if (c == 0) ++trace_g;
#endif
            }
            i = j;
          }
        }
      }
    } else if (warp >= 2) {
      // ===================================================================== loader / epilogue warpgroups
      const int c = (warp - 2) / (ctxT / 32);
      const int wg_tid = threadIdx.x - 64 - ctxT * c;
      const int quad = warp & 3;
      const int half = ((warp - 2) >> 2) % parts;
      const int r = quad * 32 + lane;
      const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
      const uint32_t a_ready_remote = mapa_shared(a_ready(c), 0);
      uint32_t acc_par = 0, img_par = 0;
      bool store_pending = false, w_waited = false, img_pending = false;

      auto guard_slots = [&]() {
        if (store_pending) {
          if (wg_tid == 0) bulk_wait_read0();
          named_barrier_sync(1 + c, ctxT);
          store_pending = false;
        }
      };

      for (int q = q0; q < q0 + n_super; ++q) {
        const int ptile = nctx * q + c;
        if (ptile >= p.num_ptiles) continue;
        const int tile = 2 * ptile + static_cast<int>(rank);     // 128-row tile of every tile image
        const bool tile_ok = tile < p.num_tiles;
        const int64_t row0 = static_cast<int64_t>(tile) * 128;
#ifdef NRC_CHAIN_TRACE
        const int trace_it = q - q0;
        const bool tracing = blockIdx.x == 0 && first && wg_tid == 0 && trace_it < kTraceTiles;
        if (tracing) g_chain_trace[c][trace_it][kTraceOps - 1] = clock64();
#endif
        for (int i = 0; i < nops;) {
          const DevOp& op = sops[i];
          if (op.kind == NRC_OP_GEMM) {
            if (!(op.flags & kFlagSplit)) {   // (a split group's operands were announced atom by atom by the epilogue before it)
              fence_proxy_async_smem();
              tc_fence_before();
              named_barrier_sync(1 + c, ctxT);
              if (wg_tid == 0) {
                if (!w_waited) { mbar_wait(w_ready, 0); w_waited = true; if (first) TRACE_MARK(11); }
                if (img_pending) { mbar_arrive(img_ready(c)); mbar_wait(img_ready(c), img_par); img_par ^= 1u; img_pending = false; }
                mbar_arrive_cluster(a_ready_remote);
              }
            }
            while (i < nops && sops[i].kind == NRC_OP_GEMM) ++i;
            mbar_wait(acc_ready(c), acc_par);
            acc_par ^= 1u;
            tc_fence_after();
#ifdef NRC_CHAIN_TRACE
            if (tracing) g_chain_trace[c][trace_it][i - 1] = clock64();
#endif
            continue;
          }
          if (op.kind == NRC_OP_LOADIMG) {
            // bf16 atoms written by the producer (another chain's SAVE, a per-point kernel) straight into the slots;
            // consecutive LOADIMG ops before a GEMM share one barrier phase
            guard_slots();
            if (wg_tid == 0 && tile_ok) {
              // every LOADIMG adds its bytes to the phase; the one arrival follows at the GEMM that consumes them
              mbar_expect_tx(img_ready(c), static_cast<uint32_t>(op.npad) * kAtomBytes);
              img_pending = true;
              const uint8_t* img = static_cast<const uint8_t*>(op.ptr);
              for (int a = 0; a < op.npad; ++a)
                bulk_g2s(slot_addr(c, op.slot + a), img + (static_cast<size_t>(tile) * op.img_atoms + op.col0 + a) * kAtomBytes,
                         kAtomBytes, img_ready(c));
            }
          } else if (op.kind == NRC_OP_LOAD) {
            guard_slots();
            const float* src = static_cast<const float*>(op.ptr);
            const int nch = op.npad >> 3;
            const bool vec = src && (op.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
            constexpr int kU = 4;
            const int total = 128 * nch;
            for (int item0 = wg_tid; item0 < total; item0 += ctxT * kU) {
              float v[kU][8];
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                const int item = item0 + u * ctxT;
                const int rr = item / nch;
                const int col = (item - rr * nch) * 8;
#pragma unroll
                for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
                if (item < total && src && row0 + rr < p.num_rows && col < op.ncols) {
                  const float* sp = src + (row0 + rr) * op.ld + col;
                  if (vec && col + 8 <= op.ncols) {
                    const float4 x0 = __ldg(reinterpret_cast<const float4*>(sp));
                    const float4 x1 = __ldg(reinterpret_cast<const float4*>(sp) + 1);
                    v[u][0] = x0.x; v[u][1] = x0.y; v[u][2] = x0.z; v[u][3] = x0.w;
                    v[u][4] = x1.x; v[u][5] = x1.y; v[u][6] = x1.z; v[u][7] = x1.w;
                  } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                      if (col + e < op.ncols) v[u][e] = __ldg(sp + e);
                  }
                }
              }
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                const int item = item0 + u * ctxT;
                if (item >= total) continue;
                const int rr = item / nch;
                const int dcol = op.col0 + (item - rr * nch) * 8;
                const uint32_t dst = slot_addr(c, op.slot + (dcol >> 6)) + atom_chunk_offset(rr, (dcol & 63) >> 3);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack2_bf16(v[u][0], v[u][1])),
                             "r"(pack2_bf16(v[u][2], v[u][3])), "r"(pack2_bf16(v[u][4], v[u][5])),
                             "r"(pack2_bf16(v[u][6], v[u][7]))
                             : "memory");
              }
            }
          } else if (op.kind == NRC_OP_SAVE) {
            fence_proxy_async_smem();
            named_barrier_sync(1 + c, ctxT);
            if (wg_tid == 0 && tile_ok) {
              uint8_t* img = static_cast<uint8_t*>(op.ptr);
              for (int a = 0; a < op.npad; ++a)
                bulk_s2g(img + (static_cast<size_t>(tile) * op.img_atoms + op.col0 + a) * kAtomBytes,
                         slot_addr(c, op.slot + a), kAtomBytes);
              bulk_commit();
            }
            store_pending = true;
          } else {  // NRC_OP_EPI
            if (op.slot >= 0) guard_slots();
            bool split_signalled = false;
            EpiArgs a;
            a.bias = static_cast<const float*>(op.ptr);
            a.bias_s = op.bias_off >= 0 ? sbias + op.bias_off : nullptr;
            float* out = static_cast<float*>(op.out);
            a.out_row = (out && row0 + r < p.num_rows) ? out + (row0 + r) * op.ld + op.col0 : nullptr;
            a.out_vec = out && (op.ld % 4 == 0) && (op.col0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
            a.accum = (op.flags & NRC_EPI_OUT_ACCUMULATE) != 0;
            a.mask_tile = (op.mask && tile_ok)
                              ? static_cast<const uint8_t*>(op.mask) + static_cast<size_t>(tile) * op.img_atoms * kAtomBytes
                              : nullptr;
            a.mask_atom0 = op.mask_atom0;
            a.taddr = tmem_base + t_lane + static_cast<uint32_t>(c * kCtxTmemCols + op.tmem_col);
            a.has_slot = op.slot >= 0;
            a.slot0_addr = a.has_slot ? slot_addr(c, op.slot) : 0u;
            a.ncols = op.ncols; a.npad = op.npad; a.r = r; a.half = half; a.nparts = parts;
            const bool full = (op.ncols == op.npad) && (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0);
            const bool relu = (op.flags & NRC_EPI_RELU) != 0;
            if (out && op.n > 0 && !op.mask && (!a.bias || a.bias_s)) {
              // fp32 row output through the staging slot the program names (op.n - 1)
              guard_slots();
              float* obase = out + op.col0;
              const uint32_t st = slot_addr(c, op.n - 1);
              epi_staged(a.taddr, a.bias ? a.bias_s : nullptr, relu, st, r, half, op.ncols, op.npad, obase, row0, p.num_rows, op.ld,
                         a.accum, a.has_slot, a.slot0_addr, 1 + c, ctxT, wg_tid);
            } else if (a.has_slot && !out && op.ncols == op.npad && (!a.bias || a.bias_s) && (tile_ok || !op.mask)) {
              // the accumulator only becomes the next operand: tight path
              const int sb = (op.flags & kFlagSplit) ? 1 + c : 0;
              if (a.mask_tile) epi_slot_fast<true>(a.taddr, nullptr, false, a.slot0_addr, r, op.npad, half, parts, a.mask_tile, a.mask_atom0,
                                                   sb, ctxT, wg_tid == 0, a_ready_remote);
              else             epi_slot_fast<false>(a.taddr, a.bias ? a.bias_s : nullptr, relu, a.slot0_addr, r, op.npad, half, parts, nullptr, 0,
                                                    sb, ctxT, wg_tid == 0, a_ready_remote);
              split_signalled = sb != 0;
            } else if (a.mask_tile) {
              if (full) epi_run<false, false, true, true>(a); else epi_run<false, false, true, false>(a);
            } else if (a.bias) {
              if (relu) { if (full) epi_run<true, true, false, true>(a); else epi_run<true, true, false, false>(a); }
              else      { if (full) epi_run<true, false, false, true>(a); else epi_run<true, false, false, false>(a); }
            } else {
              if (relu) { if (full) epi_run<false, true, false, true>(a); else epi_run<false, true, false, false>(a); }
              else      { if (full) epi_run<false, false, false, true>(a); else epi_run<false, false, false, false>(a); }
            }
            if ((op.flags & kFlagSplit) && !split_signalled) {
              // a general-path epilogue in front of a split group (e.g. the rows past the end in the odd last tile):
              // the issuer still expects one phase per operand atom
              for (uint32_t k = 0; k < static_cast<uint32_t>(op.npad >> 6); ++k) {
                fence_proxy_async_smem();
                tc_fence_before();
                named_barrier_sync(1 + c, ctxT);
                if (wg_tid == 0) mbar_arrive_cluster(a_ready_remote + (k ? 32u + 16u * k : 0u));
              }
            }
          }
#ifdef NRC_CHAIN_TRACE
          if (tracing) g_chain_trace[c][trace_it][i] = clock64();
#endif
          ++i;
        }
      }
      if (first) TRACE_MARK(12);
      if (wg_tid == 0) bulk_wait0();
    }
    __syncwarp();
    first = false;
  }

  tc_fence_before();
  TRACE_MARK(13);
  cluster_sync_all();   // no CTA leaves (or frees tensor memory) while its partner may still use its memories
  if (warp == 0 && !first) tmem_dealloc2(tmem_base, 512);
  TRACE_NS(15);
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
            op.tmem_col < 0 || op.tmem_col + op.n > 512 || op.w_chunk < 0)
          return NRC_E_INVALID_ARG;
        for (int a = 0; a < op.n_atoms; ++a)
          if (op.a_slot[a] >= S || op.a_klen[a] < 16 || op.a_klen[a] > 64 || (op.a_klen[a] & 15)) return NRC_E_INVALID_ARG;
        break;
      case NRC_OP_EPI:
        if ((op.flags & NRC_EPI_DENSITY) && (!ptr_ok(op.out_ptr, false) || op.npad != 16)) return NRC_E_INVALID_ARG;
        if (op.npad <= 0 || (op.npad & 15) || op.ncols > op.npad || op.tmem_col < 0 || op.tmem_col + op.npad > 512 ||
            !ptr_ok(op.ptr, true) || !ptr_ok(op.out_ptr, true) || !ptr_ok(op.mask_ptr, true) ||
            (op.slot >= 0 && op.slot + ((op.npad + 63) >> 6) > S))
          return NRC_E_INVALID_ARG;
        // fp32 output staged through slot n - 1: inside the context and not one of the destination slots
        if (op.n < 0 || op.n > S || (op.n > 0 && op.slot >= 0 && op.n - 1 >= op.slot && op.n - 1 < op.slot + ((op.npad + 63) >> 6)))
          return NRC_E_INVALID_ARG;
        break;
      case NRC_OP_GATHER:
        if (!ptr_ok(op.ptr, false) || !ptr_ok(op.out_ptr, true) || op.slot < 0 || op.slot >= S || op.ncols < 1 ||
            op.ncols > 32 || op.npad < op.ncols || (op.npad & 15) || op.npad > 32)
          return NRC_E_INVALID_ARG;
        break;
      case NRC_OP_SAVE:
      case NRC_OP_LOADIMG:
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

// v2 launch (CTA pairs, resident weights): returns NRC_E_UNSUPPORTED when the program does not fit, so the caller
// can fall back to the streaming kernel.
// Resolve one program into the plan the pair kernel executes.  NRC_E_UNSUPPORTED when it does not fit (the caller
// then falls back to the streaming kernel).
static int32_t chain2_build(const nrc_chain_program_t* prog, void* const* d_ptrs, const void* d_weights_packed,
                            Chain2Prog& out, int& cost) {
  Chain2Plan& pl = out.plan;
  const int nops = prog->num_ops;
  const int S = prog->slots_per_ctx;
  // resident weight layout: byte offset of every GEMM's first K atom inside the CTA's weight region
  int w_off[NRC_CHAIN_MAX_OPS];
  int w_bytes = 0;
  pl.n_wcopy = 0;
  for (int i = 0; i < nops; ++i) {
    const nrc_chain_op_t& op = prog->ops[i];
    w_off[i] = 0;
    if (op.kind == NRC_OP_GATHER || (op.kind == NRC_OP_EPI && (op.flags & NRC_EPI_DENSITY))) return NRC_E_UNSUPPORTED;
    if (op.kind != NRC_OP_GEMM) continue;
    // a GEMM that multiplies further operand atoms with weights an earlier op already holds (summed upstream
    // gradients) reuses the resident copy: its chunks are a sub-range of the earlier op's
    int shared = -1;
    for (int k = 0; k < i && shared < 0; ++k) {
      const nrc_chain_op_t& e = prog->ops[k];
      if (e.kind == NRC_OP_GEMM && e.n == op.n && op.w_chunk >= e.w_chunk && op.w_chunk + op.n_atoms <= e.w_chunk + e.n_atoms)
        shared = k;
    }
    if (shared >= 0) {
      w_off[i] = w_off[shared] + (op.w_chunk - prog->ops[shared].w_chunk) * op.n * 64;
      continue;
    }
    w_off[i] = w_bytes;
    const uint32_t half_bytes = static_cast<uint32_t>(op.n) * 64u;   // n/2 rows of 128 bytes per K atom
    for (int a = 0; a < op.n_atoms; ++a) {
      if (pl.n_wcopy >= kMaxWcopy) return NRC_E_UNSUPPORTED;
      pl.wcopy[pl.n_wcopy++] = make_uint4(static_cast<uint32_t>(w_bytes) + a * half_bytes,
                                          static_cast<uint32_t>(op.w_chunk + a) * kAtomBytes, half_bytes, half_bytes);
    }
    w_bytes += op.n_atoms * static_cast<int>(half_bytes);
  }
  const int budget = 227 * 1024 - kTail2Bytes - 1024;
  int nctx = 2;
  if (w_bytes + 2 * S * kAtomBytes > budget) nctx = 1;
  for (int i = 0; i < nops; ++i) {   // accumulators beyond a context's 256 columns: one context owns all 512
    const nrc_chain_op_t& op = prog->ops[i];
    if ((op.kind == NRC_OP_GEMM && op.tmem_col + op.n > 256) || (op.kind == NRC_OP_EPI && op.tmem_col + op.npad > 256)) nctx = 1;
  }
  if (w_bytes + nctx * S * kAtomBytes > budget) return NRC_E_UNSUPPORTED;
  const char* force = getenv("NRC_CHAIN_NCTX");
  if (force && force[0] == '1') nctx = 1;

  // the program with resolved pointers, staged biases and the UMMA descriptor list
  const int bias_cap = (kTail2Bytes - 144 - static_cast<int>(sizeof(DevOp)) * NRC_CHAIN_MAX_OPS - static_cast<int>(sizeof(uint4)) * kMaxMma) / 4;
  int bias_cur = 0;
  pl.n_ops = nops; pl.n_mma = 0; pl.n_bias = 0;
  auto ptr_of = [&](int32_t i) -> void* { return i >= 0 ? d_ptrs[i] : nullptr; };
  const uint32_t slot0 = static_cast<uint32_t>(kTail2Bytes + w_bytes);   // context 0's first slot, relative to the CTA's shared memory
  int n_groups = 0, epi_cols = 0;
  for (int i = 0; i < nops; ++i) {
    const nrc_chain_op_t& o = prog->ops[i];
    DevOp& d = pl.ops[i];
    d.kind = static_cast<int8_t>(o.kind); d.slot = static_cast<int8_t>(o.slot);
    d.flags = static_cast<uint8_t>(o.flags); d.n_atoms = static_cast<uint8_t>(o.n_atoms);
    d.ncols = static_cast<int16_t>(o.ncols); d.npad = static_cast<int16_t>(o.npad);
    d.tmem_col = static_cast<int16_t>(o.tmem_col); d.n = static_cast<int16_t>(o.n);
    d.ld = o.ld;
    d.col0 = static_cast<int16_t>(o.col0); d.mask_atom0 = static_cast<int16_t>(o.mask_atom0);
    d.img_atoms = static_cast<int16_t>(o.img_atoms);
    d.w_chunk = o.w_chunk;
    d.fparam = o.fparam;
    d.ptr = ptr_of(o.ptr); d.out = ptr_of(o.out_ptr); d.mask = ptr_of(o.mask_ptr);
    for (int a = 0; a < NRC_CHAIN_MAX_ATOMS; ++a) d.a_src[a] = static_cast<uint8_t>((o.a_slot[a] & 15) | ((o.a_klen[a] >> 4) << 4));
    d.bias_off = -1;
    if (o.kind == NRC_OP_EPI) epi_cols += o.npad;
    if (o.kind == NRC_OP_EPI && o.ptr >= 0 && o.mask_ptr < 0 && pl.n_bias < kMaxBias && bias_cur + o.npad <= bias_cap) {
      const uintptr_t bp = reinterpret_cast<uintptr_t>(d.ptr);
      pl.bias[pl.n_bias++] = make_uint4(static_cast<uint32_t>(bp), static_cast<uint32_t>(static_cast<uint64_t>(bp) >> 32),
                                        static_cast<uint32_t>(o.ncols) | (static_cast<uint32_t>(o.npad) << 16),
                                        static_cast<uint32_t>(bias_cur));
      d.bias_off = static_cast<int16_t>(bias_cur);
      bias_cur += o.npad;
    }
  }
  // UMMA instructions of every GEMM group as ready-made descriptor words (context 0, addresses relative to the CTA's
  // dynamic shared memory; the issuer adds the base and the context's slot offset): the issuing thread spends a handful
  // of instructions per UMMA instead of re-deriving everything from the program.
  //
  // SPLIT groups.  A hidden layer costs [epilogue of layer l: ~1.9 k cycles] -> barrier -> [UMMAs of layer l+1: ~1 k
  // cycles of issue + completion latency] -> barrier -> ... strictly one after the other (trace build: the GEMM ops
  // are 55 % of a SurfaceLightField tile and only a third of that is UMMA issue).  When the operand of a group is
  // produced by the epilogue directly in front of it (only SAVEs in between), that epilogue announces every finished
  // 64-column atom and the group's k-steps are ordered by the atom they read - operands that were resident before
  // (skip connections) and atom 0 first - with a barrier wait where the next atom's k-steps begin.  The group
  // accumulates in the OTHER 128-column half of the context's tensor memory, because the epilogue is still draining
  // the previous accumulator.  Conditions: the working accumulators of the whole program live in columns [0, 128)
  // (persistent input-gradient accumulators at >= 256 are left alone), 2-3 operand atoms, tight epilogue path.
  static const bool split_on = !(getenv("NRC_CHAIN_SPLIT") && getenv("NRC_CHAIN_SPLIT")[0] == '0');
  bool regions_free = true;     // nobody uses columns [128, 256)
  for (int i = 0; i < nops; ++i) {
    const nrc_chain_op_t& o = prog->ops[i];
    const int w = o.kind == NRC_OP_GEMM ? o.n : (o.kind == NRC_OP_EPI ? o.npad : 0);
    if (w > 0 && !(o.tmem_col + w <= 128 || o.tmem_col >= 256)) regions_free = false;
  }
  int region_prev = 0;
  for (int i = 0; i < nops;) {
    if (prog->ops[i].kind != NRC_OP_GEMM) { ++i; continue; }
    int j = i;
    while (j < nops && prog->ops[j].kind == NRC_OP_GEMM) ++j;
    int jc = j;                                       // consumers: everything up to the next GEMM group
    while (jc < nops && prog->ops[jc].kind != NRC_OP_GEMM) ++jc;
    ++n_groups;
    // the epilogue that produces this group's operand
    int E = i - 1;
    while (E >= 0 && prog->ops[E].kind == NRC_OP_SAVE) --E;
    bool split = split_on && regions_free && E >= 0 && prog->ops[E].kind == NRC_OP_EPI;
    int K = 0;
    if (split) {
      const nrc_chain_op_t& e = prog->ops[E];
      K = e.npad >> 6;
      split = e.slot >= 0 && e.out_ptr < 0 && e.ncols == e.npad && (e.npad & 63) == 0 && K >= 2 && K <= 3 && e.n == 0 &&
              !(e.flags & NRC_EPI_DENSITY) && (e.ptr < 0 || pl.ops[E].bias_off >= 0);
    }
    struct Ent { int op, a, k, stage; };
    Ent ents[kMaxMma];
    int ne = 0;
    bool stage_seen[4] = {false, false, false, false};
    for (int g = i; g < j; ++g) {
      const nrc_chain_op_t& o = prog->ops[g];
      for (int a = 0; a < o.n_atoms; ++a)
        for (int k = 0; k < (o.a_klen[a] >> 4); ++k) {
          if (ne >= kMaxMma) return NRC_E_UNSUPPORTED;
          int stage = 0;
          if (split) {
            const int rel = o.a_slot[a] - prog->ops[E].slot;
            stage = (rel >= 0 && rel < K) ? rel : 0;
            stage_seen[stage] = true;
          }
          ents[ne++] = Ent{g, a, k, stage};
        }
    }
    if (split)
      for (int st = 0; st < K; ++st) split = split && stage_seen[st];
    // accumulator targets of the group: equal or disjoint column ranges
    for (int g = i; g < j && split; ++g)
      for (int h = i; h < g; ++h) {
        const nrc_chain_op_t &x = prog->ops[g], &y = prog->ops[h];
        const bool same = x.tmem_col == y.tmem_col && x.n == y.n;
        const bool disjoint = x.tmem_col + x.n <= y.tmem_col || y.tmem_col + y.n <= x.tmem_col;
        if (!same && !disjoint) split = false;
      }
    const int region = split ? (region_prev ^ 1) : 0;
    if (split) {
      // stable order by stage
      Ent sorted[kMaxMma];
      int ns = 0;
      for (int st = 0; st < K; ++st)
        for (int e = 0; e < ne; ++e)
          if (ents[e].stage == st) sorted[ns++] = ents[e];
      for (int e = 0; e < ne; ++e) ents[e] = sorted[e];
      pl.ops[E].flags |= kFlagSplit;
      pl.ops[i].flags |= kFlagSplit;
    }
    if (pl.n_mma + ne > kMaxMma) return NRC_E_UNSUPPORTED;
    const int first = pl.n_mma;
    int stage_cur = -1;
    for (int e = 0; e < ne; ++e) {
      const nrc_chain_op_t& o = prog->ops[ents[e].op];
      const int a = ents[e].a, k = ents[e].k;
      const uint32_t half_bytes = static_cast<uint32_t>(o.n) * 64u;
      const uint32_t a_addr = slot0 + static_cast<uint32_t>(o.a_slot[a]) * kAtomBytes;
      const uint32_t b_addr = static_cast<uint32_t>(kTail2Bytes + w_off[ents[e].op]) + static_cast<uint32_t>(a) * half_bytes;
      // accumulate unless this is the first instruction (in issue order) into a target that some op of the group resets
      bool acc = true;
      {
        bool earlier = false, resets = false;
        for (int f = 0; f < e; ++f) earlier = earlier || prog->ops[ents[f].op].tmem_col == o.tmem_col;
        for (int g = i; g < j; ++g)
          resets = resets || (prog->ops[g].tmem_col == o.tmem_col && !(prog->ops[g].flags & NRC_GEMM_ACCUMULATE));
        if (!earlier && resets) acc = false;
      }
      uint32_t waits = 0;     // barrier of the atom index this stage waits for, + 1
      if (ents[e].stage != stage_cur) { stage_cur = ents[e].stage; waits = static_cast<uint32_t>(stage_cur) + 1u; }
      const uint32_t col = static_cast<uint32_t>(o.tmem_col + ((o.tmem_col < 128) ? region * 128 : 0));
      pl.mma[pl.n_mma++] = make_uint4(((a_addr + 32u * k) >> 4) | (1u << 16), ((b_addr + 32u * k) >> 4) | (1u << 16),
                                      col | (waits << 29) | (acc ? 0x80000000u : 0u), make_idesc(256, o.n, 0, 0));
    }
    for (int e = first; e < pl.n_mma;) {          // segment lengths into the segment heads (bits 16-23)
      int e1 = e + 1;
      while (e1 < pl.n_mma && ((pl.mma[e1].z >> 29) & 3u) == 0u) ++e1;
      pl.mma[e].z |= static_cast<uint32_t>(e1 - e) << 16;
      e = e1;
    }
    for (int g = i; g < j; ++g) { pl.ops[g].ld = first + (g == i ? 0 : ne); pl.ops[g].col0 = static_cast<int16_t>(g == i ? ne : 0); }
    if (region)
      for (int c = j; c < jc; ++c)
        if (prog->ops[c].kind == NRC_OP_EPI && prog->ops[c].tmem_col < 128) pl.ops[c].tmem_col = static_cast<int16_t>(prog->ops[c].tmem_col + 128);
    region_prev = region;
    i = j;
  }
  out.weights = static_cast<const uint8_t*>(d_weights_packed);
  out.w_bytes = w_bytes;
  out.slots_per_ctx = S;
  out.nctx = nctx;
  // cycles one super tile (nctx pair tiles in flight together) takes, from the per-op timelines of the trace build:
  // ~100 per UMMA, ~11 per accumulator column drained, ~900 per GEMM group (barrier round trip), ~100 per op
  cost = pl.n_mma * 100 + epi_cols * 11 + n_groups * 900 + nops * 100 + 1500;
  if (nctx == 2) cost += cost / 4;
  return NRC_OK;
}

// Launch of up to kMaxProgs programs over the same rows on the CTA-pair kernel.
static int32_t chain2_launch(void* stream, int32_t n_progs, const nrc_chain_program_t* const* progs, void* const* const* d_ptrs,
                             const void* const* d_weights, int64_t num_rows) {
  static thread_local Chain2Params hp;
  if (n_progs < 1 || n_progs > kMaxProgs) return NRC_E_INVALID_ARG;
  int cost[kMaxProgs], n_super[kMaxProgs];
  hp.num_rows = num_rows;
  hp.num_tiles = static_cast<int32_t>((num_rows + 127) / 128);
  hp.num_ptiles = (hp.num_tiles + 1) / 2;
  hp.n_progs = n_progs;
  size_t smem = 0;
  long long total = 0;
  int total_units = 0;
  for (int k = 0; k < n_progs; ++k) {
    const int32_t st = chain2_build(progs[k], d_ptrs[k], d_weights[k], hp.prog[k], cost[k]);
    if (st != NRC_OK) return st;
    const Chain2Prog& pg = hp.prog[k];
    const size_t need = static_cast<size_t>(kTail2Bytes) + pg.w_bytes + static_cast<size_t>(pg.nctx * pg.slots_per_ctx) * kAtomBytes;
    if (need > smem) smem = need;
    n_super[k] = (hp.num_ptiles + pg.nctx - 1) / pg.nctx;
    if (n_super[k] > 65535) return NRC_E_UNSUPPORTED;
    total += static_cast<long long>(n_super[k]) * cost[k];
    total_units += n_super[k];
  }
  int pairs = num_sms() / 2;
  if (pairs > kMaxPairs) pairs = kMaxPairs;
  if (pairs > total_units) pairs = total_units;
  // deal the (program, super tile) items to the pairs in contiguous ranges of equal estimated time; a pair that
  // starts a further program pays its set-up (weights + barriers) on top
  for (int k = 0; k < n_progs; ++k)
    for (int i = 0; i < kMaxPairs; ++i) hp.seg_t0[k][i] = hp.seg_n[k][i] = 0;
  {
    int k = 0, t = 0;
    long long done = 0;
    for (int i = 0; i < pairs; ++i) {
      const long long goal = total * (i + 1) / pairs;
      bool any = false;
      while (k < n_progs) {
        if (t >= n_super[k]) { ++k; t = 0; continue; }
        // take the item if the pair has none yet or taking it stays closer to the goal than leaving it
        if (any && done + cost[k] / 2 > goal && i + 1 < pairs) break;
        if (hp.seg_n[k][i] == 0) hp.seg_t0[k][i] = static_cast<uint16_t>(t);
        ++hp.seg_n[k][i];
        ++t;
        done += cost[k];
        any = true;
      }
    }
  }
  if (const int32_t st_attr = ensure_dynamic_smem<chain2_kernel>(227 * 1024 - 1024); st_attr != NRC_OK) return st_attr;
  chain2_kernel<<<2 * pairs, kChainThreads, smem, static_cast<cudaStream_t>(stream)>>>(hp);
  return check_launch();
}

static int32_t chain_launch(void* stream, const nrc_chain_program_t* prog, void* const* d_ptrs, int32_t num_ptrs,
                            const void* d_weights_packed, int64_t num_rows, const EncDev* enc, float warp_c) {
  if (!d_ptrs || num_ptrs < 0 || num_ptrs > NRC_CHAIN_MAX_PTRS || num_rows < 0) return NRC_E_INVALID_ARG;
  const int32_t st = validate_program(prog, num_ptrs);
  if (st != NRC_OK) return st;
  if (num_rows == 0) return NRC_OK;
  if (!enc) {   // CTA-pair kernel with resident weights whenever the program fits (NRC_CHAIN_V1=1: streaming kernel)
    const char* v1 = getenv("NRC_CHAIN_V1");
    if (!(v1 && v1[0] == '1')) {
      const int32_t s2 = chain2_launch(stream, 1, &prog, &d_ptrs, &d_weights_packed, num_rows);
      if (s2 != NRC_E_UNSUPPORTED) return s2;
    }
  }
  for (int i = 0; i < prog->num_ops; ++i) {   // features of the CTA-pair kernel only
    const nrc_chain_op_t& op = prog->ops[i];
    if (op.kind == NRC_OP_LOADIMG) return NRC_E_UNSUPPORTED;
    if (op.kind == NRC_OP_GEMM && op.tmem_col + op.n > 256) return NRC_E_UNSUPPORTED;
    if (op.kind == NRC_OP_EPI && op.tmem_col + op.npad > 256) return NRC_E_UNSUPPORTED;
  }
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
  hp.ring_stages = R;
  const size_t smem = static_cast<size_t>(2 * S + R) * kAtomBytes + kTailBytes;
  if (const int32_t st_attr = ensure_dynamic_smem<chain_kernel>(14 * kAtomBytes + kTailBytes); st_attr != NRC_OK) return st_attr;
  const int pairs = (hp.num_tiles + 1) / 2;
  const int grid = pairs < num_sms() ? pairs : num_sms();
  chain_kernel<<<grid, kChainThreads, smem, static_cast<cudaStream_t>(stream)>>>(hp);
  return check_launch();
}

#ifdef NRC_CHAIN_TRACE
extern "C" int32_t nrc_chain_trace_dump(long long* host_out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, g_chain_trace, sizeof(long long) * 2 * kTraceTiles * kTraceOps) == cudaSuccess
             ? NRC_OK : NRC_E_CUDA;
}
extern "C" int32_t nrc_chain_mma_dump(long long* host_out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, g_chain_mma, sizeof(long long) * kTraceTiles * 8 * 2) == cudaSuccess ? NRC_OK : NRC_E_CUDA;
}
extern "C" int32_t nrc_chain_marks_dump(long long* host_out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, g_chain_marks, sizeof(long long) * 16) == cudaSuccess ? NRC_OK : NRC_E_CUDA;
}
#endif

extern "C" int32_t nrc_chain_run(void* stream, const nrc_chain_program_t* prog, void* const* d_ptrs, int32_t num_ptrs,
                                 const void* d_weights_packed, int64_t num_rows) {
  return chain_launch(stream, prog, d_ptrs, num_ptrs, d_weights_packed, num_rows, nullptr, 0.f);
}

extern "C" int32_t nrc_chain_run_multi(void* stream, int32_t num_programs, const nrc_chain_program_t* const* progs,
                                       void* const* const* d_ptrs, const int32_t* num_ptrs, const void* const* d_weights_packed,
                                       int64_t num_rows) {
  if (num_programs < 1 || num_programs > NRC_CHAIN_MAX_PROGRAMS || !progs || !d_ptrs || !num_ptrs || !d_weights_packed || num_rows < 0)
    return NRC_E_INVALID_ARG;
  for (int k = 0; k < num_programs; ++k) {
    if (!d_ptrs[k] || num_ptrs[k] < 0 || num_ptrs[k] > NRC_CHAIN_MAX_PTRS) return NRC_E_INVALID_ARG;
    const int32_t st = validate_program(progs[k], num_ptrs[k]);
    if (st != NRC_OK) return st;
  }
  if (num_rows == 0) return NRC_OK;
  const char* v1 = getenv("NRC_CHAIN_V1");
  if (!(v1 && v1[0] == '1')) {
    const int32_t s2 = chain2_launch(stream, num_programs, progs, d_ptrs, d_weights_packed, num_rows);
    if (s2 != NRC_E_UNSUPPORTED) return s2;
  }
  for (int k = 0; k < num_programs; ++k) {   // a program the pair kernel cannot hold: one launch per program
    const int32_t st = chain_launch(stream, progs[k], d_ptrs[k], num_ptrs[k], d_weights_packed[k], num_rows, nullptr, 0.f);
    if (st != NRC_OK) return st;
  }
  return NRC_OK;
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
  int splits = (2 * num_sms() + num_layers - 1) / num_layers;
  if (splits > hp.num_tiles) splits = hp.num_tiles;
  hp.tiles_per_cta = (hp.num_tiles + splits - 1) / splits;
  splits = (hp.num_tiles + hp.tiles_per_cta - 1) / hp.tiles_per_cta;
  const size_t smem = static_cast<size_t>(1 + kWgradRingAtoms) * kAtomBytes + 1024;
  if (const int32_t st_attr = ensure_dynamic_smem<wgrad_kernel>(static_cast<int>(smem)); st_attr != NRC_OK) return st_attr;
  wgrad_kernel<<<dim3(splits, num_layers), kWgradThreads, smem, static_cast<cudaStream_t>(stream)>>>(hp);
  return check_launch();
}
