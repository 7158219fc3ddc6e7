// Micro-benchmarks behind the chain kernel's design: cycles per tcgen05.mma by shape and
// operand source (A from shared memory vs. A from tensor memory, one CTA vs. a CTA pair), tensor-memory read / write
// bandwidth by warp count, and the latency of one GEMM -> epilogue -> GEMM dependency round trip.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural_radiance_caching_b200/csrc tools/umma_probe.cu -o tools/umma_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"
using namespace nrc::tc;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc, bool pair) {
  if (pair)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS (A, B from shared memory); mode 1: TS (A from tensor memory)
template <int CG>
__global__ void __launch_bounds__(128, 1) mma_rate(int n, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + kAtomBytes;   // B: up to 256 rows of 128 bytes
  for (int i = threadIdx.x; i < (kAtomBytes + 32768) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4u * i), "r"(0x3C003C00u) : "memory");
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { if (CG == 2) tmem_alloc2(smem_u32(&tbase), 512); else tmem_alloc(smem_u32(&tbase), 512); }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t t0 = tbase;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  long long c0 = 0, c1 = 0;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = make_idesc(128 * CG, n, 0, 0);
    c0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = i & 3;
      if (mode == 0 || mode == 2) {
        const uint32_t d = (mode == 2 && (i & 1)) ? t0 + 256 : t0;   // mode 2: two independent accumulators, alternating
        if (CG == 2) umma2_bf16(d, kmajor_desc(a_addr, k), kmajor_desc(b_addr, k), idesc, i > 1);
        else umma_bf16(d, kmajor_desc(a_addr, k), kmajor_desc(b_addr, k), idesc, i > 1);
      } else {
        umma_ts(t0, t0 + 256 + 8 * k, kmajor_desc(b_addr, k), idesc, i > 0, CG == 2);
      }
    }
    if (CG == 2) umma2_commit_mc(smem_u32(&bar), 3); else umma_commit(smem_u32(&bar));
    c1 = clock64();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (threadIdx.x == 0 && rank == 0) {
    const long long c2 = clock64();
    if (blockIdx.x == 0) { out[0] = c1 - c0; out[1] = c2 - c0; }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) { if (CG == 2) tmem_dealloc2(t0, 512); else tmem_dealloc(t0, 512); }
}

// tensor-memory read / write bandwidth: `warps` warps, each reading (or writing) `cols` columns of its lane quadrant
__global__ void __launch_bounds__(512, 1) tmem_bw(int cols, int iters, int write, long long* out) {
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tbase), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5, parts = nw / 4;
  const uint32_t lane_base = tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t v[16], acc = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = threadIdx.x + e;
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int j = 16 * (warp >> 2); j < cols; j += 16 * parts) {
      if (write) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(lane_base + j),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                     "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
      } else {
        tmem_ld16(lane_base + j, v);
      }
    }
    if (write) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); else tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) acc += v[e];
  }
  __syncthreads();
  const long long c1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = acc; }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

// one dependency round trip of the MLP chain with a single tile: UMMA group (ksteps k-steps, N = n) -> commit -> all
// `warps` epilogue warps wait, read the accumulator (ncols/parts columns each), write bf16 to shared memory (mode 0)
// or to tensor memory (mode 1), fence, barrier, signal the issuer -> next group.
__global__ void __launch_bounds__(576, 1) round_trip(int n, int ksteps, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tbase;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 2 * kAtomBytes;
  for (int i = threadIdx.x; i < (6 * kAtomBytes) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4u * i), "r"(0u) : "memory");
  fence_proxy_async_smem();
  const int nepi = blockDim.x - 64;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), nepi); mbar_init(smem_u32(&bars[1]), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tbase), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t0 = tbase;
  const int warp = threadIdx.x >> 5;
  const uint32_t a_ready = smem_u32(&bars[0]), acc_ready = smem_u32(&bars[1]);
  const long long c0 = clock64();
  if (warp == 1) {
    if ((threadIdx.x & 31) == 0) {
      const uint32_t idesc = make_idesc(128, n, 0, 0);
      uint32_t par = 0;
      for (int it = 0; it < iters; ++it) {
        if (it > 0) { mbar_wait(a_ready, par); par ^= 1u; tc_fence_after(); }
        for (int k = 0; k < ksteps; ++k) {
          if (mode == 0) umma_bf16(t0, kmajor_desc(a_addr + (k >> 2) * kAtomBytes, k & 3), kmajor_desc(b_addr + (k >> 2) * kAtomBytes, k & 3), idesc, k > 0);
          else umma_ts(t0, t0 + 256 + 8 * k, kmajor_desc(b_addr + (k >> 2) * kAtomBytes, k & 3), idesc, k > 0, false);
        }
        umma_commit(acc_ready);
      }
    }
  } else if (warp >= 2) {
    const int ew = warp - 2, parts = (nepi / 32) / 4;
    const int quad = warp & 3, part = ew >> 2, r = quad * 32 + (threadIdx.x & 31);
    const uint32_t lane_base = t0 + (static_cast<uint32_t>(quad * 32) << 16);
    uint32_t par = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(acc_ready, par); par ^= 1u;
      tc_fence_after();
      for (int j = 16 * part; j < n; j += 16 * parts) {
        uint32_t v[16];
        tmem_ld16(lane_base + j, v);
        tmem_ld_wait();
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = pack2_bf16(fmaxf(__uint_as_float(v[2 * e]), 0.f), fmaxf(__uint_as_float(v[2 * e + 1]), 0.f));
        if (mode == 0) {
          const uint32_t d = a_addr + static_cast<uint32_t>(j >> 6) * kAtomBytes;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + atom_chunk_offset(r, (j & 63) >> 3)), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + atom_chunk_offset(r, ((j & 63) >> 3) + 1)), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
        } else {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + 256 + (j >> 1)),
                       "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
        }
      }
      if (mode == 0) fence_proxy_async_smem(); else asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(a_ready);
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = c1 - c0;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(t0, 512);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

int main() {
  long long* d_out;
  long long h[2];
  CK(cudaMalloc(&d_out, 16));
  const int smem = 8 * kAtomBytes;
  CK(cudaFuncSetAttribute(mma_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(mma_rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(round_trip, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 512;
  for (int grid : {1, 148}) {
    for (int cg : {1, 2}) {
      for (int mode : {0, 1, 2}) {
        for (int n : {16, 64, 128, 256}) {
          if (grid == 148 && mode != 2) continue;
          if (cg == 1) mma_rate<1><<<grid, 128, smem>>>(n, mode, iters, d_out);
          else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid == 1 ? 2 : 148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, mma_rate<2>, n, mode, iters, d_out));
          }
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
          const double cyc = double(h[1]) / iters;
          const double macs = 128.0 * cg * n * 16;
          printf("mma_rate grid=%3d cta_group=%d %s M=%d N=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma -> %.0f MAC/cyc/SM (%.0f%% of 4096)\n",
                 grid, cg, mode == 1 ? "TS" : (mode == 2 ? "SS-2acc" : "SS"), 128 * cg, n, double(h[0]) / iters, cyc, macs / cyc / cg, 100.0 * macs / cyc / cg / 4096);
        }
      }
    }
  }
  for (int warps : {4, 8, 16}) {
    for (int write : {0, 1}) {
      tmem_bw<<<1, warps * 32>>>(128, 256, write, d_out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
      printf("tmem %s warps=%2d: 128 lanes x 128 cols fp32 (64 KB) in %.0f cyc -> %.1f B/cyc/SM\n", write ? "st" : "ld", warps,
             double(h[0]) / 256, 65536.0 * 256 / double(h[0]));
    }
  }
  for (int warps : {4, 8, 16}) {
    for (int mode : {0, 1}) {
      for (int ks : {8, 16}) {
        round_trip<<<1, 64 + warps * 32, smem>>>(128, ks, mode, 256, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("round_trip epi_warps=%2d A->%s N=128 K=%3d: %.0f cyc per layer (MMA floor %d)\n", warps, mode ? "tmem" : "smem", ks * 16,
               double(h[0]) / 256, ks * 64);
      }
    }
  }
  return 0;
}
