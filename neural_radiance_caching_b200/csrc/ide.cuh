// Integrated directional encoding core (Ref-NeRF eqs. 6-8): ref_utils.generate_ide_fn
// (internal/ref_utils.py:131-192).  Shared by ide.cu (stand-alone entry points) and shader.cu (fused
// per-point stages of the cache shader).
#pragma once
#include "nrc_common.cuh"

namespace nrc {

constexpr int kMaxL = 16;      // deg_view <= 5
constexpr int kMaxSh = 36;     // 2+3+5+9+17

constexpr int kIdeLanes = 8;   // threads that share one point in the fused shader stages (shader.cu)
constexpr int kIdePerLane = 8;

struct IdeTable {
  int n_sh;
  int l_max;
  int m[kMaxSh];
  int l[kMaxSh];
  float sigma[kMaxSh];
  // harmonics each of the kIdeLanes threads of a point evaluates, most expensive first (0xFF = none): the lanes of a
  // warp walk their lists in lockstep, step j costs about the same in every lane
  unsigned char lane_list[kIdeLanes][kIdePerLane];
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

// The z-polynomial of harmonic i and its derivative by Horner's rule in fp64 (no power table: l - m fused multiply-adds
// in a thread that owns only a few harmonics).
__device__ __forceinline__ void ide_poly_horner(const IdeTable& tab, const float* __restrict__ mat, int i, double z,
                                                float& poly, float& dpoly) {
  const int n = tab.l[i] - tab.m[i];
  double pv = static_cast<double>(__ldg(mat + n * tab.n_sh + i)), dv = 0.0;
  for (int k = n - 1; k >= 0; --k) {
    dv = fma(dv, z, pv);
    pv = fma(pv, z, static_cast<double>(__ldg(mat + k * tab.n_sh + i)));
  }
  poly = static_cast<float>(pv);
  dpoly = static_cast<float>(dv);
}

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
  // longest-processing-time assignment of the harmonics to the lanes of a point
  int load[kIdeLanes] = {0}, cnt[kIdeLanes] = {0};
  bool done[kMaxSh] = {false};
  for (int q = 0; q < kIdeLanes; ++q)
    for (int j = 0; j < kIdePerLane; ++j) t.lane_list[q][j] = 0xFF;
  for (int it = 0; it < n_sh; ++it) {
    int best = -1;
    for (int i = 0; i < n_sh; ++i)
      if (!done[i] && (best < 0 || ml_l[i] - ml_m[i] > ml_l[best] - ml_m[best])) best = i;
    int lane = -1;
    for (int q = 0; q < kIdeLanes; ++q)
      if (cnt[q] < kIdePerLane && (lane < 0 || load[q] < load[lane])) lane = q;
    if (lane < 0) return NRC_E_UNSUPPORTED;
    t.lane_list[lane][cnt[lane]++] = static_cast<unsigned char>(best);
    load[lane] += ml_l[best] - ml_m[best] + 6;
    done[best] = true;
  }
  return NRC_OK;
}

}  // namespace nrc
