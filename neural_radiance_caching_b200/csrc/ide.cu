// Integrated directional encoding (Ref-NeRF eqs. 6-8) used by the cache shader's view-dependent
// MLPs: ref_utils.generate_ide_fn (internal/ref_utils.py:131-192).  One thread per direction.
// The (l, m) list, the Legendre/SH coefficient matrix `mat` [(l_max+1), n_sh] and sigma are built
// on the host exactly like the reference (float64 -> float32) and passed in.
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

__global__ void ide_fwd_kernel(const __grid_constant__ IdeTable tab, const float* __restrict__ mat,
                               const float* __restrict__ xyz, const float* __restrict__ kappa_inv, int64_t P,
                               float* __restrict__ out, int64_t ldo) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float x = xyz[3 * p], y = xyz[3 * p + 1], z = xyz[3 * p + 2];
  const float kinv = kappa_inv[p];
  // The l = 16 Legendre polynomials are alternating sums with coefficients up to ~1e5: in fp32
  // (the reference) they carry ~1e-3 relative noise.  The z-polynomial is therefore accumulated
  // in fp64 here (222 DFMA per point; B200 runs fp64 at full rate), which lands within fp32
  // rounding of the exact value -- closer to the truth than any fp32 evaluation order.
  double zp[kMaxL + 1];
  float cr[kMaxL + 1], ci[kMaxL + 1];
  zp[0] = 1.0; cr[0] = 1.f; ci[0] = 0.f;
  for (int k = 1; k <= tab.l_max; ++k) {
    zp[k] = zp[k - 1] * static_cast<double>(z);
    cr[k] = cr[k - 1] * x - ci[k - 1] * y;
    ci[k] = cr[k - 1] * y + ci[k - 1] * x;
  }
  float* o = out + p * ldo;
  for (int i = 0; i < tab.n_sh; ++i) {
    const int m = tab.m[i], l = tab.l[i];
    double polyd = 0.0;
    for (int k = 0; k <= l - m; ++k) polyd = fma(zp[k], static_cast<double>(__ldg(mat + k * tab.n_sh + i)), polyd);
    const float poly = static_cast<float>(polyd);
    const float att = expf(-tab.sigma[i] * kinv);
    o[i] = cr[m] * poly * att;
    o[tab.n_sh + i] = ci[m] * poly * att;
  }
}

// VJP: g_xyz [P,3] and g_kappa_inv [P] (written).
__global__ void ide_bwd_kernel(const __grid_constant__ IdeTable tab, const float* __restrict__ mat,
                               const float* __restrict__ xyz, const float* __restrict__ kappa_inv,
                               const float* __restrict__ g_out, int64_t ldg, int64_t P,
                               float* __restrict__ g_xyz, float* __restrict__ g_kappa) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float x = xyz[3 * p], y = xyz[3 * p + 1], z = xyz[3 * p + 2];
  const float kinv = kappa_inv[p];
  double zp[kMaxL + 1];
  float cr[kMaxL + 1], ci[kMaxL + 1];
  zp[0] = 1.0; cr[0] = 1.f; ci[0] = 0.f;
  for (int k = 1; k <= tab.l_max; ++k) {
    zp[k] = zp[k - 1] * static_cast<double>(z);
    cr[k] = cr[k - 1] * x - ci[k - 1] * y;
    ci[k] = cr[k - 1] * y + ci[k - 1] * x;
  }
  const float* g = g_out + p * ldg;
  float gx = 0.f, gy = 0.f, gz = 0.f, gk = 0.f;
  for (int i = 0; i < tab.n_sh; ++i) {
    const int m = tab.m[i], l = tab.l[i];
    double polyd = 0.0, dpolyd = 0.0;
    for (int k = 0; k <= l - m; ++k) {
      const double c = static_cast<double>(__ldg(mat + k * tab.n_sh + i));
      polyd = fma(zp[k], c, polyd);
      if (k > 0) dpolyd = fma(static_cast<double>(k) * zp[k - 1], c, dpolyd);
    }
    const float poly = static_cast<float>(polyd), dpoly = static_cast<float>(dpolyd);
    const float att = expf(-tab.sigma[i] * kinv);
    const float gr = g[i], gi = g[tab.n_sh + i];
    // out_r = cr[m] poly att, out_i = ci[m] poly att
    const float s = gr * cr[m] + gi * ci[m];
    gz += s * dpoly * att;
    gk += -tab.sigma[i] * s * poly * att;
    if (m > 0) {
      // d (x+iy)^m / dx = m (x+iy)^(m-1);  d/dy = i m (x+iy)^(m-1)
      const float fm = static_cast<float>(m) * poly * att;
      const float pr = cr[m - 1], pi = ci[m - 1];
      gx += fm * (gr * pr + gi * pi);
      gy += fm * (-gr * pi + gi * pr);
    }
  }
  if (g_xyz) { g_xyz[3 * p] = gx; g_xyz[3 * p + 1] = gy; g_xyz[3 * p + 2] = gz; }
  if (g_kappa) g_kappa[p] = gk;
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

using namespace nrc;

extern "C" int32_t nrc_ide_fwd(void* stream, int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l,
                               const float* sigma, const float* d_mat, const float* d_xyz,
                               const float* d_kappa_inv, int64_t num_points, float* d_out, int64_t ldo) {
  IdeTable t;
  int32_t st = make_ide_table(n_sh, ml_m, ml_l, sigma, t);
  if (st != NRC_OK) return st;
  if (num_points < 0 || ldo < 2 * n_sh) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_mat || !d_xyz || !d_kappa_inv || !d_out) return NRC_E_INVALID_ARG;
  unsigned grid = static_cast<unsigned>((num_points + 127) / 128);
  ide_fwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, d_mat, d_xyz, d_kappa_inv, num_points,
                                                                      d_out, ldo);
  return check_launch();
}

extern "C" int32_t nrc_ide_bwd(void* stream, int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l,
                               const float* sigma, const float* d_mat, const float* d_xyz,
                               const float* d_kappa_inv, const float* d_g_out, int64_t ldg, int64_t num_points,
                               float* d_g_xyz, float* d_g_kappa_inv) {
  IdeTable t;
  int32_t st = make_ide_table(n_sh, ml_m, ml_l, sigma, t);
  if (st != NRC_OK) return st;
  if (num_points < 0 || ldg < 2 * n_sh) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_mat || !d_xyz || !d_kappa_inv || !d_g_out) return NRC_E_INVALID_ARG;
  unsigned grid = static_cast<unsigned>((num_points + 127) / 128);
  ide_bwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, d_mat, d_xyz, d_kappa_inv, d_g_out, ldg,
                                                                      num_points, d_g_xyz, d_g_kappa_inv);
  return check_launch();
}
