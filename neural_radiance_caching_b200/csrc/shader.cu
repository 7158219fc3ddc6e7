// Per-point stages of the cache shader between its Dense stacks:
// internal/nerf.py:940-1090 (_predict_appearance_passive), :461-482 (integrated BRDF input),
// :1344-1358 (_get_refdirs), internal/ref_utils.py:25-42 (reflect), :131-192 (IDE) and the
// activations of configs/nerf_ngp_yobo.gin:491-506.  One thread per shaded point; everything a point
// needs between two MLP stacks happens in one kernel, so the only tensors that touch HBM are the
// stacks' inputs and outputs.
//   mid : head outputs + normal + view direction -> roughness, n.v, reflection direction, IDE_5 (SLF
//         input) and IDE_4 (EnvMap input; the degree-4 harmonics are a prefix of the degree-5 list)
//   out : raw outputs of the heads / integrated-BRDF / SLF / EnvMap stacks -> rgb and the extras
#include <cuda_bf16.h>

#include "ide.cuh"
#include "loss_terms.cuh"
#include "tc05.cuh"

namespace nrc {

// Row `p` of a bf16 tile image [tiles][img_atoms][16 KB] (the operand layout of the tensor-core chains, tc05.cuh):
// columns [0, npad) of the atoms starting at `atom0`, taken from vals[0, ncols) and zero padded; npad % 8 == 0.
// The per-point kernels hand their results to the chains in this form, so the chains bulk-copy operands instead of
// reading and converting fp32 rows.
__device__ __forceinline__ void store_row_atoms(uint8_t* __restrict__ img, int img_atoms, int atom0, int64_t p,
                                                const float* vals, int ncols, int npad) {
  const int64_t tile = p >> 7;
  const int r = static_cast<int>(p & 127);
  uint8_t* base = img + (tile * img_atoms + atom0) * static_cast<int64_t>(tc::kAtomBytes);
  for (int c0 = 0; c0 < npad; c0 += 8) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c0 + 2 * e;
      w[e] = tc::pack2_bf16(c < ncols ? vals[c] : 0.f, c + 1 < ncols ? vals[c + 1] : 0.f);
    }
    *reinterpret_cast<uint4*>(base + static_cast<size_t>(c0 >> 6) * tc::kAtomBytes + tc::atom_chunk_offset(r, (c0 & 63) >> 3)) =
        make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

struct ShaderGeom {
  float nx, ny, nz, wx, wy, wz, dot, rx, ry, rz;
  __device__ __forceinline__ void load(const float* __restrict__ normals, const float* __restrict__ viewdirs,
                                       int64_t p, int32_t spr) {
    nx = normals[3 * p]; ny = normals[3 * p + 1]; nz = normals[3 * p + 2];
    const int64_t ray = p / spr;
    wx = -viewdirs[3 * ray]; wy = -viewdirs[3 * ray + 1]; wz = -viewdirs[3 * ray + 2];
    dot = nx * wx + ny * wy + nz * wz;
    // reflect(w, n) = 2 (n.w) n - w
    rx = 2.f * dot * nx - wx; ry = 2.f * dot * ny - wy; rz = 2.f * dot * nz - wz;
  }
};

// mid stages: kIdeLanes threads per point (a 128-thread-per-CTA, thread-per-point version was latency bound: 222
// dependent fp64 multiply-adds per thread at ~7 resident warps per SM took 30 us (forward) / 40 us (VJP) for 32 768
// points).  Each lane evaluates the harmonics the table assigns to it (Horner in fp64), the point's values meet in
// shared memory, and the bf16 operand rows of the downstream stacks leave as whole 16-byte chunks.
constexpr int kMidPts = 32;                      // points per CTA
constexpr int kMidThreads = kMidPts * kIdeLanes;
constexpr int kMidRow = 2 * kMaxSh + 4;          // staged floats per point: [re | im] of the degree-5 list, n.v

__global__ void __launch_bounds__(kMidThreads)
shader_mid_fwd_kernel(const __grid_constant__ IdeTable tab, int n_sh_env, const float* __restrict__ mat,
                      const float* __restrict__ heads, int64_t ldh, const float* __restrict__ normals,
                      const float* __restrict__ viewdirs, int64_t P, int32_t spr, float rough_bias,
                      float* __restrict__ roughness, float* __restrict__ dotprod, float* __restrict__ refdirs,
                      float* __restrict__ ide_slf, float* __restrict__ ide_env, const nrc_shader_images_t im) {
  __shared__ float sv[kMidPts][kMidRow];
  __shared__ float2 spw[kMidPts][kMaxL + 1];
  const int pl = threadIdx.x / kIdeLanes, ln = threadIdx.x % kIdeLanes;
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kMidPts + pl;
  const bool valid = p < P;
  const int64_t pc = valid ? p : P - 1;   // out-of-range lanes follow the last point and store nothing
  const int n_sh = tab.n_sh;
  ShaderGeom g;
  g.load(normals, viewdirs, pc, spr);
  const float rough = softplus_f(heads[pc * ldh] + rough_bias);
  if (ln == 0) {
    if (valid) {
      if (roughness) roughness[p] = rough;
      if (dotprod) dotprod[p] = g.dot;
      if (refdirs) { refdirs[3 * p] = g.rx; refdirs[3 * p + 1] = g.ry; refdirs[3 * p + 2] = g.rz; }
    }
    sv[pl][2 * kMaxSh] = g.dot;
    // (x + iy)^k of the reflection direction, shared by the point's lanes
    float cr = 1.f, ci = 0.f;
    spw[pl][0] = make_float2(1.f, 0.f);
    for (int k = 1; k <= tab.l_max; ++k) {
      const float nr = cr * g.rx - ci * g.ry;
      ci = cr * g.ry + ci * g.rx;
      cr = nr;
      spw[pl][k] = make_float2(cr, ci);
    }
  }
  __syncwarp();
  const double z = static_cast<double>(g.rz);
  for (int j = 0; j < kIdePerLane; ++j) {
    const int i = tab.lane_list[ln][j];
    if (i == 0xFF) break;
    float poly, dpoly;
    ide_poly_horner(tab, mat, i, z, poly, dpoly);
    const float att = expf(-tab.sigma[i] * rough);
    const float2 c = spw[pl][tab.m[i]];
    sv[pl][i] = c.x * poly * att;
    sv[pl][n_sh + i] = c.y * poly * att;
  }
  __syncwarp();   // the kIdeLanes threads of a point sit in one warp
  if (valid) {    // optional fp32 copies (exact-mode stacks read these)
    if (ide_slf)
      for (int j = ln; j < 2 * n_sh; j += kIdeLanes) ide_slf[p * (2 * n_sh) + j] = sv[pl][j];
    if (ide_env)
      for (int j = ln; j < 2 * n_sh_env; j += kIdeLanes)
        ide_env[p * (2 * n_sh_env) + j] = sv[pl][j < n_sh_env ? j : n_sh + (j - n_sh_env)];
  }
  // bf16 operand rows: 16-byte chunks dealt to the lanes of the warp (its 32 / kIdeLanes points)
  const int wl = threadIdx.x & 31;
  const int pts_per_warp = 32 / kIdeLanes;
  const int n_slf = im.slf_img ? (((2 * n_sh + 15) & ~15) >> 3) : 0;
  const int n_env = im.env_img ? (((2 * n_sh_env + 15) & ~15) >> 3) : 0;
  const int n_dot = im.dot_img ? 2 : 0;
  const int per_pt = n_slf + n_env + n_dot;
  const int pl0 = (threadIdx.x >> 5) * pts_per_warp;
  for (int id = wl; id < per_pt * pts_per_warp; id += 32) {
    const int q = id / per_pt;
    int ch = id - q * per_pt;
    const int64_t pp = static_cast<int64_t>(blockIdx.x) * kMidPts + pl0 + q;
    if (pp >= P) continue;
    const float* row = sv[pl0 + q];
    float v[8];
    uint8_t* img;
    int img_atoms, atom0;
    if (ch < n_slf) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (8 * ch + e < 2 * n_sh) ? row[8 * ch + e] : 0.f;
      img = static_cast<uint8_t*>(im.slf_img); img_atoms = im.slf_img_atoms; atom0 = im.slf_atom0;
    } else if (ch < n_slf + n_env) {
      ch -= n_slf;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = 8 * ch + e;
        v[e] = j < n_sh_env ? row[j] : (j < 2 * n_sh_env ? row[n_sh + (j - n_sh_env)] : 0.f);
      }
      img = static_cast<uint8_t*>(im.env_img); img_atoms = im.env_img_atoms; atom0 = im.env_atom0;
    } else {
      ch -= n_slf + n_env;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (ch == 0) v[0] = row[2 * kMaxSh];
      img = static_cast<uint8_t*>(im.dot_img); img_atoms = im.dot_img_atoms; atom0 = im.dot_atom0;
    }
    const int64_t tile = pp >> 7;
    const int r = static_cast<int>(pp & 127);
    uint8_t* dst = img + (tile * img_atoms + atom0 + (ch >> 3)) * static_cast<int64_t>(tc::kAtomBytes) + tc::atom_chunk_offset(r, ch & 7);
    *reinterpret_cast<uint4*>(dst) = make_uint4(tc::pack2_bf16(v[0], v[1]), tc::pack2_bf16(v[2], v[3]), tc::pack2_bf16(v[4], v[5]),
                                                tc::pack2_bf16(v[6], v[7]));
  }
}

__global__ void __launch_bounds__(kMidThreads)
shader_mid_bwd_kernel(const __grid_constant__ IdeTable tab, int n_sh_env, const float* __restrict__ mat,
                      const float* __restrict__ heads, int64_t ldh, const float* __restrict__ normals,
                      const float* __restrict__ viewdirs, int64_t P, int32_t spr, float rough_bias,
                      const float* __restrict__ g_dot, int64_t ldgd, const float* __restrict__ g_ide_slf, int64_t ldgs,
                      const float* __restrict__ g_ide_env, int64_t ldge, float* __restrict__ g_heads, int64_t ldgh,
                      float* __restrict__ g_normals) {
  const int pl = threadIdx.x / kIdeLanes, ln = threadIdx.x % kIdeLanes;
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kMidPts + pl;
  const bool valid = p < P;
  const int64_t pc = valid ? p : P - 1;   // out-of-range lanes follow the last point (the shuffles below are warp wide)
  ShaderGeom g;
  g.load(normals, viewdirs, pc, spr);
  const float r_raw = heads[pc * ldh] + rough_bias;
  const float rough = softplus_f(r_raw);
  __shared__ float2 spw[kMidPts][kMaxL + 1];
  if (ln == 0) {
    float cr = 1.f, ci = 0.f;
    spw[pl][0] = make_float2(1.f, 0.f);
    for (int k = 1; k <= tab.l_max; ++k) {
      const float nr = cr * g.rx - ci * g.ry;
      ci = cr * g.ry + ci * g.rx;
      cr = nr;
      spw[pl][k] = make_float2(cr, ci);
    }
  }
  __syncwarp();
  const double z = static_cast<double>(g.rz);
  const float* g5 = g_ide_slf + pc * ldgs;
  const float* g4 = g_ide_env ? g_ide_env + pc * ldge : nullptr;
  const int n_sh = tab.n_sh;
  float grx = 0.f, gry = 0.f, grz = 0.f, gk = 0.f;
  for (int j = 0; j < kIdePerLane; ++j) {
    const int i = tab.lane_list[ln][j];
    if (i == 0xFF) break;
    float gr = g5[i], gi = g5[n_sh + i];
    if (g4 && i < n_sh_env) { gr += g4[i]; gi += g4[n_sh_env + i]; }
    float poly, dpoly;
    ide_poly_horner(tab, mat, i, z, poly, dpoly);
    const float sig = tab.sigma[i];
    const float att = expf(-sig * rough);
    const int m = tab.m[i];
    const float2 c = spw[pl][m];
    // out_r = c.x poly att, out_i = c.y poly att
    const float s = gr * c.x + gi * c.y;
    grz += s * dpoly * att;
    gk += -sig * s * poly * att;
    if (m > 0) {
      const float2 pm = spw[pl][m - 1];   // d (x+iy)^m / dx = m (x+iy)^(m-1);  d/dy = i m (x+iy)^(m-1)
      const float fm = static_cast<float>(m) * poly * att;
      grx += fm * (gr * pm.x + gi * pm.y);
      gry += fm * (-gr * pm.y + gi * pm.x);
    }
  }
#pragma unroll
  for (int o = 1; o < kIdeLanes; o <<= 1) {
    grx += __shfl_xor_sync(0xffffffffu, grx, o);
    gry += __shfl_xor_sync(0xffffffffu, gry, o);
    grz += __shfl_xor_sync(0xffffffffu, grz, o);
    gk += __shfl_xor_sync(0xffffffffu, gk, o);
  }
  if (!valid || ln != 0) return;
  // softplus'(x) = sigmoid(x)
  g_heads[p * ldgh] = gk * sigmoid_f(r_raw);
  // dot = n.w ; r = 2 dot n - w  =>  dL/dn = g_dot w + 2 dot g_r + 2 (g_r.n) w
  const float gd = g_dot[p * ldgd];
  const float grn = grx * g.nx + gry * g.ny + grz * g.nz;
  const float s = gd + 2.f * grn;
  g_normals[3 * p] = s * g.wx + 2.f * g.dot * grx;
  g_normals[3 * p + 1] = s * g.wy + 2.f * g.dot * gry;
  g_normals[3 * p + 2] = s * g.wz + 2.f * g.dot * grz;
}

// heads columns: 0 roughness, 1-3 ambient irradiance, 4-6 irradiance, 7-9 tint
struct ShadeConsts { float rgb_max, diffuse_bias, light_bias, brdf_bias; };

// `out` stage of one point: rgb (and the 22 diagnostic channels when e != nullptr).
__device__ __forceinline__ void shade_point_fwd(const float* __restrict__ h, float f_raw, const float* __restrict__ slf,
                                                const float* __restrict__ envr, const ShadeConsts k, float (&rgb)[3], float* e) {
  const float F = sigmoid_f(f_raw + k.brdf_bias);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float amb_d = fminf(fmaxf(softplus_f(h[1 + c] + k.diffuse_bias), 0.f), k.rgb_max);
    const float ind_d = fminf(fmaxf(softplus_f(h[4 + c] + k.diffuse_bias), 0.f), k.rgb_max);
    const float tint = sigmoid_f(h[7 + c]);
    const float env = fmaxf(softplus_f(envr[c] + k.light_bias), 0.f);
    const float ref = fmaxf(softplus_f(slf[c] + k.light_bias), 0.f);
    const float acc = 1.0f;   // incoming_acc of the IDE-form light field (surface_light_field.py:1046-1069)
    const float amb_s = fminf(fmaxf(tint * F * (env * (1.0f - acc)), 0.f), k.rgb_max);
    const float ind_s = fminf(fmaxf(tint * F * (ref * acc), 0.f), k.rgb_max);
    const float ambient = amb_d + amb_s, indirect = ind_d + ind_s;
    rgb[c] = ambient + indirect;
    if (e) {
      e[c] = amb_d + ind_d;       // diffuse_rgb
      e[3 + c] = amb_s + ind_s;   // specular_rgb
      e[6 + c] = ambient;         // ambient_rgb
      e[9 + c] = indirect;        // indirect_rgb
      e[12 + c] = tint;           // albedo_rgb
      e[16 + c] = env;            // env_rgb
      e[19 + c] = ref;            // ref_rgb
    }
  }
  if (e) e[15] = F;
}

// VJP of rgb only (the extras are diagnostics; the cache loss reads rgb).
__device__ __forceinline__ void shade_point_bwd(const float* __restrict__ h, float f_raw, const float* __restrict__ slf,
                                                const ShadeConsts k, const float (&g_rgb)[3], float* __restrict__ g_heads,
                                                float* __restrict__ g_f, float* __restrict__ g_slf) {
  const float F = sigmoid_f(f_raw + k.brdf_bias);
  float gF = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float g = g_rgb[c];
    const float xa = h[1 + c] + k.diffuse_bias, xi = h[4 + c] + k.diffuse_bias;
    // clip passes the gradient on the closed interval [0, rgb_max] (jnp.clip / torch.clamp)
    g_heads[1 + c] = softplus_f(xa) <= k.rgb_max ? g * sigmoid_f(xa) : 0.f;
    g_heads[4 + c] = softplus_f(xi) <= k.rgb_max ? g * sigmoid_f(xi) : 0.f;
    const float tint = sigmoid_f(h[7 + c]);
    const float xs = slf[c] + k.light_bias;
    const float ref = fmaxf(softplus_f(xs), 0.f);
    const float spec = tint * F * ref;
    const float gs = (spec >= 0.f && spec <= k.rgb_max) ? g : 0.f;
    g_heads[7 + c] = gs * F * ref * tint * (1.f - tint);
    gF += gs * tint * ref;
    g_slf[c] = gs * tint * F * sigmoid_f(xs);
  }
  g_f[0] = gF * F * (1.f - F);
}

__global__ void shader_out_fwd_kernel(const float* __restrict__ heads, int64_t ldh, const float* __restrict__ f_raw,
                                      int64_t ldf, const float* __restrict__ slf_raw, int64_t lds,
                                      const float* __restrict__ env_raw, int64_t lde, int64_t P, const ShadeConsts k,
                                      float* __restrict__ rgb, float* __restrict__ extras) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float v[3];
  shade_point_fwd(heads + p * ldh, f_raw[p * ldf], slf_raw + p * lds, env_raw + p * lde, k, v, extras ? extras + p * 22 : nullptr);
  rgb[3 * p] = v[0]; rgb[3 * p + 1] = v[1]; rgb[3 * p + 2] = v[2];
}

__global__ void shader_out_bwd_kernel(const float* __restrict__ heads, int64_t ldh, const float* __restrict__ f_raw,
                                      int64_t ldf, const float* __restrict__ slf_raw, int64_t lds, int64_t P,
                                      const ShadeConsts k, const float* __restrict__ g_rgb, float* __restrict__ g_heads,
                                      int64_t ldgh, float* __restrict__ g_f, int64_t ldgf, float* __restrict__ g_slf,
                                      int64_t ldgs) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float g[3] = {g_rgb[3 * p], g_rgb[3 * p + 1], g_rgb[3 * p + 2]};
  shade_point_bwd(heads + p * ldh, f_raw[p * ldf], slf_raw + p * lds, k, g, g_heads + p * ldgh, g_f + p * ldgf, g_slf + p * ldgs);
}

// The tail of the cache training step's forward and the head of its backward in ONE launch, one warp per ray:
// `out` stage of the ray's samples -> volumetric rendering (rgb, acc; internal/render.py:172-224) -> Charbonnier-sRGB data
// term + mask loss -> VJP of the compositing -> VJP of the `out` stage.  Same expressions and summation order as
// shader_out_fwd_kernel, render_loss_kernel (ray.cu) and shader_out_bwd_kernel, which it replaces on the training path
// (three launches in a row on the step's critical path, each a fraction of a wave) and which remain the reference
// points of the parity tests.  The per-sample colours stay in registers between the two halves (n <= 128).
constexpr int kShadeLossWarps = 4;
__global__ void __launch_bounds__(kShadeLossWarps * 32)
shade_render_loss_kernel(const float* __restrict__ heads, int64_t ldh, const float* __restrict__ f_raw, int64_t ldf,
                         const float* __restrict__ slf_raw, int64_t lds, const float* __restrict__ env_raw, int64_t lde,
                         const ShadeConsts k, const float* __restrict__ weights, const float* __restrict__ bg,
                         const float* __restrict__ target, const float* __restrict__ mask, int64_t R, int n,
                         float charb_padding, int use_mask, float opaque_w, float empty_w, float* __restrict__ loss,
                         float* __restrict__ rgb_s, float* __restrict__ out_rgb, float* __restrict__ acc_out,
                         float* __restrict__ g_weights, float* __restrict__ g_heads, int64_t ldgh, float* __restrict__ g_f,
                         int64_t ldgf, float* __restrict__ g_slf, int64_t ldgs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kShadeLossWarps + warp;
  float contrib = 0.f;
  if (r < R) {
    float col[4][3], w[4];
    float a = 0.f, sc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = lane + 32 * q;
      w[q] = 0.f;
      col[q][0] = col[q][1] = col[q][2] = 0.f;
      if (i < n) {
        const int64_t p = r * n + i;
        shade_point_fwd(heads + p * ldh, f_raw[p * ldf], slf_raw + p * lds, env_raw + p * lde, k, col[q], nullptr);
        w[q] = weights[p];
        if (rgb_s) { rgb_s[3 * p] = col[q][0]; rgb_s[3 * p + 1] = col[q][1]; rgb_s[3 * p + 2] = col[q][2]; }
        a += w[q];
#pragma unroll
        for (int c = 0; c < 3; ++c) sc[c] += w[q] * col[q][c];
      }
    }
    auto wsum = [](float v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      return v;
    };
    const float acc = wsum(a);
    const float bg_w = fmaxf(0.f, 1.0f - acc);
    float out[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      out[c] = wsum(sc[c]);
      if (bg) out[c] += bg_w * bg[3 * r + c];
    }
    if (lane < 3) out_rgb[3 * r + lane] = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
    if (lane == 0 && acc_out) acc_out[r] = acc;
    float g = 0.f;
    const float invR = 1.0f / static_cast<float>(R);
    if (lane < 3) {
      const float x = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
      contrib = charb_srgb_term(x, target[3 * r + lane], charb_padding, 1.0f / (3.0f * static_cast<float>(R)), g);
    } else if (lane == 3 && use_mask) {
      contrib = mask_term(acc, mask ? mask[r] : 1.0f, opaque_w, empty_w, invR, charb_padding, g);
    }
    const float go[3] = {__shfl_sync(0xffffffffu, g, 0), __shfl_sync(0xffffffffu, g, 1), __shfl_sync(0xffffffffu, g, 2)};
    float gacc = __shfl_sync(0xffffffffu, g, 3);
    if (bg && (1.0f - acc) > 0.f) {
      gacc -= go[0] * bg[3 * r];
      gacc -= go[1] * bg[3 * r + 1];
      gacc -= go[2] * bg[3 * r + 2];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = lane + 32 * q;
      if (i < n) {
        const int64_t p = r * n + i;
        float gw = gacc;
        gw += go[0] * col[q][0]; gw += go[1] * col[q][1]; gw += go[2] * col[q][2];
        g_weights[p] = gw;
        const float gv[3] = {go[0] * w[q], go[1] * w[q], go[2] * w[q]};
        shade_point_bwd(heads + p * ldh, f_raw[p * ldf], slf_raw + p * lds, k, gv, g_heads + p * ldgh, g_f + p * ldgf,
                        g_slf + p * ldgs);
      }
    }
    contrib = wsum(contrib);
  }
  __shared__ float part[kShadeLossWarps];
  if (lane == 0) part[warp] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < kShadeLossWarps; ++q) sum += part[q];
    atomicAdd(loss, sum);
  }
}

// coord.pos_enc (internal/coord.py:298-312): [x, sin(2^j x), sin(2^j x + pi/2)], j = min_deg..max_deg-1
__global__ void pos_enc_kernel(const float* __restrict__ x, int64_t P, int dim, int min_deg, int max_deg,
                               int append_identity, float* __restrict__ out, int64_t ldo) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float* o = out + p * ldo;
  const int nd = max_deg - min_deg;
  int base = 0;
  if (append_identity) {
    for (int c = 0; c < dim; ++c) o[c] = x[p * dim + c];
    base = dim;
  }
  for (int j = 0; j < nd; ++j) {
    const float scale = exp2f(static_cast<float>(min_deg + j));
    for (int c = 0; c < dim; ++c) {
      const float v = __fmul_rn(x[p * dim + c], scale);
      o[base + j * dim + c] = sinf(v);
      o[base + nd * dim + j * dim + c] = sinf(__fadd_rn(v, 0.5f * 3.14159265358979323846f));
    }
  }
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_pos_enc(void* stream, const float* d_x, int64_t num_points, int32_t dim, int32_t min_deg,
                               int32_t max_deg, int32_t append_identity, float* d_out, int64_t ldo) {
  if (num_points < 0 || dim < 1 || max_deg < min_deg || ldo < (append_identity ? dim : 0) + 2 * dim * (max_deg - min_deg))
    return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_x || !d_out) return NRC_E_INVALID_ARG;
  pos_enc_kernel<<<static_cast<unsigned>((num_points + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_x, num_points, dim, min_deg, max_deg, append_identity, d_out, ldo);
  return check_launch();
}

extern "C" int32_t nrc_shader_mid_fwd(void* stream, int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l,
                                      const float* sigma, const float* d_mat, int32_t n_sh_env, const float* d_heads,
                                      int64_t ldh, const float* d_normals, const float* d_viewdirs, int64_t num_points,
                                      int32_t samples_per_ray, float roughness_bias, float* d_roughness,
                                      float* d_dotprod, float* d_refdirs, float* d_ide_slf, float* d_ide_env,
                                      const nrc_shader_images_t* images) {
  IdeTable t;
  int32_t st = make_ide_table(n_sh, ml_m, ml_l, sigma, t);
  if (st != NRC_OK) return st;
  if (num_points < 0 || samples_per_ray < 1 || ldh < 1 || n_sh_env < 0 || n_sh_env > n_sh) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  nrc_shader_images_t im = {};
  if (images) im = *images;
  if (!d_mat || !d_heads || !d_normals || !d_viewdirs || (!d_dotprod && !im.dot_img) || (!d_ide_slf && !im.slf_img))
    return NRC_E_INVALID_ARG;
  const unsigned grid = static_cast<unsigned>((num_points + kMidPts - 1) / kMidPts);
  shader_mid_fwd_kernel<<<grid, kMidThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      t, n_sh_env, d_mat, d_heads, ldh, d_normals, d_viewdirs, num_points, samples_per_ray, roughness_bias, d_roughness,
      d_dotprod, d_refdirs, d_ide_slf, d_ide_env, im);
  return check_launch();
}

extern "C" int32_t nrc_shader_mid_bwd(void* stream, int32_t n_sh, const int32_t* ml_m, const int32_t* ml_l,
                                      const float* sigma, const float* d_mat, int32_t n_sh_env, const float* d_heads,
                                      int64_t ldh, const float* d_normals, const float* d_viewdirs, int64_t num_points,
                                      int32_t samples_per_ray, float roughness_bias, const float* d_g_dotprod,
                                      int64_t ldgd, const float* d_g_ide_slf, int64_t ldgs, const float* d_g_ide_env,
                                      int64_t ldge, float* d_g_heads, int64_t ldgh, float* d_g_normals) {
  IdeTable t;
  int32_t st = make_ide_table(n_sh, ml_m, ml_l, sigma, t);
  if (st != NRC_OK) return st;
  if (num_points < 0 || samples_per_ray < 1 || ldh < 1 || ldgh < 1 || ldgd < 1 || ldgs < 2 * n_sh || n_sh_env < 0 ||
      n_sh_env > n_sh || (d_g_ide_env && ldge < 2 * n_sh_env))
    return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_mat || !d_heads || !d_normals || !d_viewdirs || !d_g_dotprod || !d_g_ide_slf || !d_g_heads || !d_g_normals)
    return NRC_E_INVALID_ARG;
  const unsigned grid = static_cast<unsigned>((num_points + kMidPts - 1) / kMidPts);
  shader_mid_bwd_kernel<<<grid, kMidThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      t, n_sh_env, d_mat, d_heads, ldh, d_normals, d_viewdirs, num_points, samples_per_ray, roughness_bias, d_g_dotprod,
      ldgd, d_g_ide_slf, ldgs, d_g_ide_env, ldge, d_g_heads, ldgh, d_g_normals);
  return check_launch();
}

extern "C" int32_t nrc_shader_out_fwd(void* stream, const float* d_heads, int64_t ldh, const float* d_f_raw, int64_t ldf,
                                      const float* d_slf_raw, int64_t lds, const float* d_env_raw, int64_t lde,
                                      int64_t num_points, float rgb_max, float diffuse_bias, float light_bias,
                                      float brdf_bias, float* d_rgb, float* d_extras) {
  if (num_points < 0 || ldh < 10 || ldf < 1 || lds < 3 || lde < 3) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_heads || !d_f_raw || !d_slf_raw || !d_env_raw || !d_rgb) return NRC_E_INVALID_ARG;
  const unsigned grid = static_cast<unsigned>((num_points + 127) / 128);
  shader_out_fwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_heads, ldh, d_f_raw, ldf, d_slf_raw, lds, d_env_raw, lde, num_points,
      nrc::ShadeConsts{rgb_max, diffuse_bias, light_bias, brdf_bias}, d_rgb, d_extras);
  return check_launch();
}

extern "C" int32_t nrc_shader_out_bwd(void* stream, const float* d_heads, int64_t ldh, const float* d_f_raw, int64_t ldf,
                                      const float* d_slf_raw, int64_t lds, int64_t num_points, float rgb_max,
                                      float diffuse_bias, float light_bias, float brdf_bias, const float* d_g_rgb,
                                      float* d_g_heads, int64_t ldgh, float* d_g_f_raw, int64_t ldgf,
                                      float* d_g_slf_raw, int64_t ldgs) {
  if (num_points < 0 || ldh < 10 || ldf < 1 || lds < 3 || ldgh < 10 || ldgf < 1 || ldgs < 3) return NRC_E_INVALID_ARG;
  if (num_points == 0) return NRC_OK;
  if (!d_heads || !d_f_raw || !d_slf_raw || !d_g_rgb || !d_g_heads || !d_g_f_raw || !d_g_slf_raw)
    return NRC_E_INVALID_ARG;
  const unsigned grid = static_cast<unsigned>((num_points + 127) / 128);
  shader_out_bwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_heads, ldh, d_f_raw, ldf, d_slf_raw, lds, num_points, nrc::ShadeConsts{rgb_max, diffuse_bias, light_bias, brdf_bias},
      d_g_rgb, d_g_heads, ldgh, d_g_f_raw, ldgf, d_g_slf_raw, ldgs);
  return check_launch();
}

extern "C" int32_t nrc_shade_render_loss(void* stream, const float* d_heads, int64_t ldh, const float* d_f_raw, int64_t ldf,
                                         const float* d_slf_raw, int64_t lds, const float* d_env_raw, int64_t lde,
                                         float rgb_max, float diffuse_bias, float light_bias, float brdf_bias,
                                         const float* d_weights, const float* d_bg, const float* d_target, const float* d_mask,
                                         int64_t num_rays, int32_t n, float charb_padding, int32_t use_mask,
                                         float opaque_weight, float empty_weight, float* d_loss, float* d_rgb_samples,
                                         float* d_out_rgb, float* d_acc, float* d_g_weights, float* d_g_heads, int64_t ldgh,
                                         float* d_g_f_raw, int64_t ldgf, float* d_g_slf_raw, int64_t ldgs) {
  if (num_rays < 1 || n < 1 || n > 128 || ldh < 10 || ldf < 1 || lds < 3 || lde < 3 || ldgh < 10 || ldgf < 1 || ldgs < 3)
    return NRC_E_INVALID_ARG;
  if (!d_heads || !d_f_raw || !d_slf_raw || !d_env_raw || !d_weights || !d_target || !d_loss || !d_out_rgb || !d_g_weights ||
      !d_g_heads || !d_g_f_raw || !d_g_slf_raw)
    return NRC_E_INVALID_ARG;
  const unsigned grid = static_cast<unsigned>((num_rays + nrc::kShadeLossWarps - 1) / nrc::kShadeLossWarps);
  nrc::shade_render_loss_kernel<<<grid, nrc::kShadeLossWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      d_heads, ldh, d_f_raw, ldf, d_slf_raw, lds, d_env_raw, lde, nrc::ShadeConsts{rgb_max, diffuse_bias, light_bias, brdf_bias},
      d_weights, d_bg, d_target, d_mask, num_rays, n, charb_padding, use_mask, opaque_weight, empty_weight, d_loss,
      d_rgb_samples, d_out_rgb, d_acc, d_g_weights, d_g_heads, ldgh, d_g_f_raw, ldgf, d_g_slf_raw, ldgs);
  return check_launch();
}
