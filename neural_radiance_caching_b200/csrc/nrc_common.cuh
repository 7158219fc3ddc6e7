// Shared device/host helpers for the nrc_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nrc_b200.h"

namespace nrc {

extern thread_local int g_last_cuda_error;

inline int32_t check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = static_cast<int>(e);
    return NRC_E_CUDA;
  }
  return NRC_OK;
}

// Multiprocessors of the CURRENT device (148 on a B200), asked once per device: grids are sized from it.
inline int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cache[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cache[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once per (kernel, device) and report
// a failure (the launch that follows would otherwise fail with an unrelated message).
template <auto Kernel>
inline int32_t ensure_dynamic_smem(int bytes) {
  static unsigned long long done = 0;   // one bit per device ordinal; a racing second call only repeats the (idempotent) set
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return check_launch();
  const unsigned long long bit = 1ull << (dev & 63);
  if (done & bit) return NRC_OK;
  if (cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return check_launch();
  done |= bit;
  return NRC_OK;
}

// float32 limits used by the reference's safe_* guards (internal/math.py:24-26).
__device__ __forceinline__ float f32_tiny() { return 1.17549435e-38f; }
__device__ __forceinline__ float f32_max() { return 3.40282347e+38f; }
__device__ __forceinline__ float f32_eps() { return 1.1920929e-07f; }

// math.safe_exp (internal/math.py:186-192): exp(clip(x, min, 70)).
__device__ __forceinline__ float safe_exp(float x) { return expf(fminf(x, 70.0f)); }
// math.safe_log (internal/math.py:177-183): log(clip(x, tiny, max)).
__device__ __forceinline__ float safe_log(float x) {
  return logf(fminf(fmaxf(x, f32_tiny()), f32_max()));
}

// coord.contract(x / c) (internal/coord.py:33-38,63-69).  Each op individually
// rounded (no FMA contraction) so that the coordinates fed to floor() follow the
// oracle's op order.  c <= 0: identity.
__device__ __forceinline__ void contract_point(float c, float x0, float x1, float x2, float& z0,
                                               float& z1, float& z2, float* scale_out = nullptr,
                                               float* mag_sq_out = nullptr) {
  if (c <= 0.f) {
    z0 = x0; z1 = x1; z2 = x2;
    if (scale_out) *scale_out = 1.f;
    if (mag_sq_out) *mag_sq_out = 0.f;
    return;
  }
  float y0 = __fdiv_rn(x0, c), y1 = __fdiv_rn(x1, c), y2 = __fdiv_rn(x2, c);
  float m = __fadd_rn(__fadd_rn(__fmul_rn(y0, y0), __fmul_rn(y1, y1)), __fmul_rn(y2, y2));
  float ms = fmaxf(1.0f, m);
  float scale = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, __fsqrt_rn(ms)), 1.0f), ms);
  z0 = __fmul_rn(scale, y0); z1 = __fmul_rn(scale, y1); z2 = __fmul_rn(scale, y2);
  if (scale_out) *scale_out = scale;
  if (mag_sq_out) *mag_sq_out = m;
}

// J^T g for z = contract(x / c).
__device__ __forceinline__ void contract_vjp(float c, float x0, float x1, float x2, float g0,
                                             float g1, float g2, float& o0, float& o1, float& o2) {
  if (c <= 0.f) { o0 = g0; o1 = g1; o2 = g2; return; }
  float y0 = x0 / c, y1 = x1 / c, y2 = x2 / c;
  float m = y0 * y0 + y1 * y1 + y2 * y2;
  float a0, a1, a2;
  if (m <= 1.0f) {
    // maximum(1, m) passes no gradient to m when m < 1 (tie at m == 1: jnp.maximum
    // splits 0.5/0.5; measure-zero, treated as the constant branch here).
    a0 = g0; a1 = g1; a2 = g2;
  } else {
    float r = sqrtf(m);
    float s = (2.0f * r - 1.0f) / m;
    // ds/dm = 1/(r m) - (2r - 1)/m^2
    float dsdm = 1.0f / (r * m) - (2.0f * r - 1.0f) / (m * m);
    float gy = g0 * y0 + g1 * y1 + g2 * y2;
    a0 = s * g0 + 2.0f * dsdm * gy * y0;
    a1 = s * g1 + 2.0f * dsdm * gy * y1;
    a2 = s * g2 + 2.0f * dsdm * gy * y2;
  }
  o0 = a0 / c; o1 = a1 / c; o2 = a2 / c;
}

// power-ladder ray warp (ray.cu: cast and fused sample + cast; slf.cu: predicted distances)
__device__ __forceinline__ float power_ladder_fwd(float x, float p, float premult) {
  // math.power_ladder general branch (internal/math.py:295-316)
  x = __fmul_rn(x, premult);
  float xp = fabsf(x);
  float xs = __fdiv_rn(xp, fmaxf(f32_tiny(), fabsf(p - 1.0f)));
  float y = __fmul_rn(__fdiv_rn(fabsf(p - 1.0f), p), __fsub_rn(powf(__fadd_rn(xs, 1.0f), p), 1.0f));
  return x < 0.f ? -y : y;
}
__device__ __forceinline__ float power_ladder_inv(float y, float p, float premult) {
  // math.inv_power_ladder general branch (internal/math.py:319-341)
  float yp = fabsf(y);
  float ymax = nextafterf((p - 1.0f) / p, -INFINITY);  // minus_eps(power_ladder_max_output(p)), p < 0
  if (p >= 0.f) ymax = f32_max();
  yp = fminf(fmaxf(yp, -ymax), ymax);
  float ratio = __fdiv_rn(p, fabsf(p - 1.0f));
  float base = __fadd_rn(__fmul_rn(ratio, yp), 1.0f);
  float x = __fmul_rn(fabsf(p - 1.0f), __fsub_rn(powf(base, __fdiv_rn(1.0f, p)), 1.0f));
  x = y < 0.f ? -x : x;
  return __fdiv_rn(x, premult);
}

// 16-byte vector reduction (sm_90+): one L2 atomic transaction for four adjacent floats.  The per-CTA weight-gradient
// flushes put a few hundred CTAs' worth of partial sums onto the SAME few thousand addresses at the same moment, and the
// L2 serialises same-address atomics: four floats per transaction is four times fewer of them.  addr must be 16-byte
// aligned.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace nrc
