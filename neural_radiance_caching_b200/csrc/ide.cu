// Integrated directional encoding (Ref-NeRF eqs. 6-8) used by the cache shader's view-dependent
// MLPs: ref_utils.generate_ide_fn (internal/ref_utils.py:131-192).  One thread per direction.
// The (l, m) list, the Legendre/SH coefficient matrix `mat` [(l_max+1), n_sh] and sigma are built
// on the host exactly like the reference (float64 -> float32) and passed in.
#include "ide.cuh"

namespace nrc {

__global__ void ide_fwd_kernel(const __grid_constant__ IdeTable tab, const float* __restrict__ mat,
                               const float* __restrict__ xyz, const float* __restrict__ kappa_inv, int64_t P,
                               float* __restrict__ out, int64_t ldo) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float kinv = kappa_inv[p];
  IdePowers pw;
  pw.init(tab.l_max, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2]);
  float* o = out + p * ldo;
  for (int i = 0; i < tab.n_sh; ++i) {
    float re, im;
    ide_term(tab, mat, pw, kinv, i, re, im);
    o[i] = re;
    o[tab.n_sh + i] = im;
  }
}

// VJP: g_xyz [P,3] and g_kappa_inv [P] (written).
__global__ void ide_bwd_kernel(const __grid_constant__ IdeTable tab, const float* __restrict__ mat,
                               const float* __restrict__ xyz, const float* __restrict__ kappa_inv,
                               const float* __restrict__ g_out, int64_t ldg, int64_t P,
                               float* __restrict__ g_xyz, float* __restrict__ g_kappa) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float kinv = kappa_inv[p];
  IdePowers pw;
  pw.init(tab.l_max, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2]);
  const float* g = g_out + p * ldg;
  float gx = 0.f, gy = 0.f, gz = 0.f, gk = 0.f;
  for (int i = 0; i < tab.n_sh; ++i) ide_term_vjp(tab, mat, pw, kinv, i, g[i], g[tab.n_sh + i], gx, gy, gz, gk);
  if (g_xyz) { g_xyz[3 * p] = gx; g_xyz[3 * p + 1] = gy; g_xyz[3 * p + 2] = gz; }
  if (g_kappa) g_kappa[p] = gk;
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
