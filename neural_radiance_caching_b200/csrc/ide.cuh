// Integrated directional encoding core (Ref-NeRF eqs. 6-8): ref_utils.generate_ide_fn
// (internal/ref_utils.py:131-192).  Shared by ide.cu (stand-alone entry points) and shader.cu (fused
// per-point stages of the cache shader).
#pragma once
#include "nrc_common.cuh"

namespace nrc {

constexpr int kMaxL = 16;      // deg_view <= 5
constexpr int kMaxSh = 36;     // 2+3+5+9+17

struct IdeTable {
  int n_sh;
  int l_max;
  int m[kMaxSh];
  int l[kMaxSh];
  float sigma[kMaxSh];
};

// Powers shared by the forward and the VJP.  The l = 16 Legendre polynomials are alternating sums
// with coefficients up to ~1e5: in fp32 (the reference) they carry ~1e-3 relative noise.  The
// z-polynomial is therefore accumulated in fp64 (222 DFMA per point; B200 runs fp64 at full rate),
// which lands within fp32 rounding of the exact value.
struct IdePowers {
  double zp[kMaxL + 1];
  float cr[kMaxL + 1], ci[kMaxL + 1];
  __device__ __forceinline__ void init(int l_max, float x, float y, float z) {
    zp[0] = 1.0; cr[0] = 1.f; ci[0] = 0.f;
    for (int k = 1; k <= l_max; ++k) {
      zp[k] = zp[k - 1] * static_cast<double>(z);
      cr[k] = cr[k - 1] * x - ci[k - 1] * y;
      ci[k] = cr[k - 1] * y + ci[k - 1] * x;
    }
  }
};

// i-th harmonic: (real, imaginary) parts.
__device__ __forceinline__ void ide_term(const IdeTable& tab, const float* __restrict__ mat, const IdePowers& pw,
                                         float kinv, int i, float& re, float& im) {
  const int m = tab.m[i], l = tab.l[i];
  double polyd = 0.0;
  for (int k = 0; k <= l - m; ++k) polyd = fma(pw.zp[k], static_cast<double>(__ldg(mat + k * tab.n_sh + i)), polyd);
  const float poly = static_cast<float>(polyd);
  const float att = expf(-tab.sigma[i] * kinv);
  re = pw.cr[m] * poly * att;
  im = pw.ci[m] * poly * att;
}

// VJP contribution of the i-th harmonic given upstream (gr, gi).
__device__ __forceinline__ void ide_term_vjp(const IdeTable& tab, const float* __restrict__ mat, const IdePowers& pw,
                                             float kinv, int i, float gr, float gi, float& gx, float& gy, float& gz,
                                             float& gk) {
  const int m = tab.m[i], l = tab.l[i];
  double polyd = 0.0, dpolyd = 0.0;
  for (int k = 0; k <= l - m; ++k) {
    const double c = static_cast<double>(__ldg(mat + k * tab.n_sh + i));
    polyd = fma(pw.zp[k], c, polyd);
    if (k > 0) dpolyd = fma(static_cast<double>(k) * pw.zp[k - 1], c, dpolyd);
  }
  const float poly = static_cast<float>(polyd), dpoly = static_cast<float>(dpolyd);
  const float att = expf(-tab.sigma[i] * kinv);
  // out_r = cr[m] poly att, out_i = ci[m] poly att
  const float s = gr * pw.cr[m] + gi * pw.ci[m];
  gz += s * dpoly * att;
  gk += -tab.sigma[i] * s * poly * att;
  if (m > 0) {
    // d (x+iy)^m / dx = m (x+iy)^(m-1);  d/dy = i m (x+iy)^(m-1)
    const float fm = static_cast<float>(m) * poly * att;
    const float pr = pw.cr[m - 1], pi = pw.ci[m - 1];
    gx += fm * (gr * pr + gi * pi);
    gy += fm * (-gr * pi + gi * pr);
  }
}

inline int32_t make_ide_table(int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l, const float* sigma,
                              IdeTable& t) {
  if (n_sh < 1 || n_sh > kMaxSh || !ml_m || !ml_l || !sigma) return NRC_E_INVALID_ARG;
  t.n_sh = n_sh;
  t.l_max = 0;
  for (int i = 0; i < n_sh; ++i) {
    if (ml_l[i] < 0 || ml_l[i] > kMaxL || ml_m[i] < 0 || ml_m[i] > ml_l[i]) return NRC_E_INVALID_ARG;
    t.m[i] = ml_m[i]; t.l[i] = ml_l[i]; t.sigma[i] = sigma[i];
    if (ml_l[i] > t.l_max) t.l_max = ml_l[i];
  }
  return NRC_OK;
}

}  // namespace nrc
