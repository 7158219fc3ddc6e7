// Gradient all-reduce (mean) over NVLink / NVSwitch peer memory: the reference's lax.pmean over the gradient
// pytree (internal/train_utils.py:3132-3136) as ONE kernel per bucket of the flat gradient arena.
//
// Two-shot over symmetric memory: rank r owns the r-th slice of the bucket, reduces it across all ranks and
// writes the mean back into every rank's copy.
//   * multicast variant (NVSwitch, NVLS): `multimem.ld_reduce` lets the switch add the N copies on the way in
//     and `multimem.st` lets it fan the result out - every byte crosses this GPU's links once per direction;
//   * peer variant (no multicast object): plain loads from / stores to the N peer mappings.
// The caller brackets the launch with a cross-rank barrier on the same stream (producers done / consumers may
// read); no host synchronisation.  Host plumbing: neural_radiance_caching_b200/dist.py (PeerArena).
#include "nrc_common.cuh"

namespace nrc {

constexpr int kArThreads = 512;
constexpr int kMaxPeers = 16;

struct PeerPtrs { float* p[kMaxPeers]; };

__device__ __forceinline__ float4 mc_ld_reduce(const float* addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* addr, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// [v0, v1): this rank's slice in float4 units
__global__ void __launch_bounds__(kArThreads) allreduce_mc_kernel(float* __restrict__ mc, int64_t v0, int64_t v1,
                                                                   float inv) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kArThreads;
  int64_t i = v0 + static_cast<int64_t>(blockIdx.x) * kArThreads + threadIdx.x;
  for (; i + 3 * stride < v1; i += 4 * stride) {     // four independent 16-byte reductions in flight per thread
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = mc_ld_reduce(mc + 4 * (i + k * stride));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k].x *= inv; v[k].y *= inv; v[k].z *= inv; v[k].w *= inv;
      mc_st(mc + 4 * (i + k * stride), v[k]);
    }
  }
  for (; i < v1; i += stride) {
    float4 v = mc_ld_reduce(mc + 4 * i);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    mc_st(mc + 4 * i, v);
  }
}

__global__ void __launch_bounds__(kArThreads) allreduce_peer_kernel(const PeerPtrs peers, int world, int rank,
                                                                     int64_t v0, int64_t v1, float inv) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kArThreads;
  for (int64_t i = v0 + static_cast<int64_t>(blockIdx.x) * kArThreads + threadIdx.x; i < v1; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world) v[r] = __ldcg(reinterpret_cast<const float4*>(peers.p[r]) + i);   // fixed rank order: every
#pragma unroll                                                                          // rank adds in the same order
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world) __stcg(reinterpret_cast<float4*>(peers.p[(rank + r) % world]) + i, acc);
  }
}

static inline void slice(int64_t offset, int64_t count, int rank, int world, int64_t& v0, int64_t& v1) {
  const int64_t n4 = count / 4, per = (n4 + world - 1) / world;
  v0 = offset / 4 + (per * rank < n4 ? per * rank : n4);
  v1 = offset / 4 + (per * (rank + 1) < n4 ? per * (rank + 1) : n4);
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_allreduce_mean_multicast(void* stream, float* mc_base, int64_t offset, int64_t count, int32_t rank,
                                                int32_t world, int32_t num_ctas) {
  if (!mc_base || world < 1 || rank < 0 || rank >= world || offset < 0 || count < 0 || (offset & 3) || (count & 3) ||
      (reinterpret_cast<uintptr_t>(mc_base) & 15))
    return NRC_E_INVALID_ARG;
  if (count == 0) return NRC_OK;
  int64_t v0, v1;
  slice(offset, count, rank, world, v0, v1);
  if (v1 <= v0) return NRC_OK;
  const int ctas = num_ctas > 0 ? num_ctas : 2 * num_sms();
  allreduce_mc_kernel<<<ctas, kArThreads, 0, static_cast<cudaStream_t>(stream)>>>(mc_base, v0, v1, 1.0f / world);
  return check_launch();
}

extern "C" int32_t nrc_allreduce_mean_peer(void* stream, float* const* peer_bases, int64_t offset, int64_t count,
                                           int32_t rank, int32_t world, int32_t num_ctas) {
  if (!peer_bases || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || offset < 0 || count < 0 ||
      (offset & 3) || (count & 3))
    return NRC_E_INVALID_ARG;
  PeerPtrs pp{};
  for (int r = 0; r < world; ++r) {
    if (!peer_bases[r] || (reinterpret_cast<uintptr_t>(peer_bases[r]) & 15)) return NRC_E_INVALID_ARG;
    pp.p[r] = peer_bases[r];
  }
  if (count == 0) return NRC_OK;
  int64_t v0, v1;
  slice(offset, count, rank, world, v0, v1);
  if (v1 <= v0) return NRC_OK;
  const int ctas = num_ctas > 0 ? num_ctas : 2 * num_sms();
  allreduce_peer_kernel<<<ctas, kArThreads, 0, static_cast<cudaStream_t>(stream)>>>(pp, world, rank, v0, v1, 1.0f / world);
  return check_launch();
}
