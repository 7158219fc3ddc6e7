// Camera rays on the device and the chunk bookkeeping of a whole-image render:
//   internal/camera_utils.py:896-1073 pixels_to_rays for the PERSPECTIVE camera without distortion, NDC or jitter (the
//   configuration of every render in BASELINE.md), in the reference's op order: (x + 0.5, y + 0.5, 1) and its two
//   neighbours -> inverse intrinsics -> OpenCV-to-OpenGL flip -> camera rotation; viewdirs = d / |d|;
//   radii = 0.5 (|d_x - d| + |d_y - d|) * 2 / sqrt(12);  near / far broadcast (cast_ray_batch :1225-1330).
//   internal/models.py:2361-2525 render_image's chunk loop: the chunk's first pixel lives in a DEVICE counter that the
//   last kernel of a chunk advances, so that one captured CUDA graph replayed n times renders n consecutive chunks of a
//   row band without any host work per chunk; a pixel index past the band repeats the band's last pixel (the reference's
//   edge padding, models.py:2434-2445) and its results are dropped by nrc_band_store.
#include "nrc_common.cuh"

namespace nrc {

struct CamParams {
  float k[9];      // pixtocam, row major
  float c2w[12];   // camtoworld [3,4], row major
  int32_t width;
  int64_t first_pixel, last_pixel;   // flat pixel index (y * width + x) of the chunk's first pixel; clamp
  float near, far;
};

__device__ __forceinline__ void mat3_vec(const float* m, float x, float y, float z, float& o0, float& o1, float& o2) {
  // numpy.matmul's inner loop: products rounded, summed left to right (no FMA contraction)
  o0 = __fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z));
  o1 = __fadd_rn(__fadd_rn(__fmul_rn(m[3], x), __fmul_rn(m[4], y)), __fmul_rn(m[5], z));
  o2 = __fadd_rn(__fadd_rn(__fmul_rn(m[6], x), __fmul_rn(m[7], y)), __fmul_rn(m[8], z));
}

__global__ void camera_rays_kernel(const __grid_constant__ CamParams p, const int64_t* __restrict__ d_first_pixel,
                                   int64_t num_rays, float* __restrict__ origins, float* __restrict__ directions,
                                   float* __restrict__ viewdirs, float* __restrict__ radii, float* __restrict__ imageplane,
                                   float* __restrict__ near, float* __restrict__ far) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= num_rays) return;
  int64_t pix = (d_first_pixel ? *d_first_pixel : p.first_pixel) + i;
  if (pix > p.last_pixel) pix = p.last_pixel;
  const float px = static_cast<float>(pix % p.width), py = static_cast<float>(pix / p.width);
  float rot[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) rot[3 * r + c] = p.c2w[4 * r + c];
  float d[3][3];   // the pixel and its +x / +y neighbours
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const float x = __fadd_rn(__fadd_rn(px, s == 1 ? 1.f : 0.f), 0.5f);
    const float y = __fadd_rn(__fadd_rn(py, s == 2 ? 1.f : 0.f), 0.5f);
    float c0, c1, c2;
    mat3_vec(p.k, x, y, 1.f, c0, c1, c2);
    c1 = -c1; c2 = -c2;   // OpenCV -> OpenGL (a matmul with diag(1, -1, -1): exact)
    if (s == 0 && imageplane) { imageplane[2 * i] = c0; imageplane[2 * i + 1] = c1; }
    mat3_vec(rot, c0, c1, c2, d[s][0], d[s][1], d[s][2]);
  }
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0][0], d[0][0]), __fmul_rn(d[0][1], d[0][1])), __fmul_rn(d[0][2], d[0][2])));
  float dn[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const float a = __fsub_rn(d[s + 1][0], d[0][0]), b = __fsub_rn(d[s + 1][1], d[0][1]), c = __fsub_rn(d[s + 1][2], d[0][2]);
    dn[s] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    origins[3 * i + a] = p.c2w[4 * a + 3];
    directions[3 * i + a] = d[0][a];
    viewdirs[3 * i + a] = __fdiv_rn(d[0][a], nrm);
  }
  // (0.5 * (dx_norm + dy_norm)) * 2 / sqrt(12), evaluated left to right in float32
  radii[i] = __fdiv_rn(__fmul_rn(__fmul_rn(0.5f, __fadd_rn(dn[0], dn[1])), 2.f), 3.4641016151377544f);
  if (near) near[i] = p.near;
  if (far) far[i] = p.far;
}

__global__ void chunk_advance_kernel(int64_t* counter, int64_t step) { *counter += step; }

__global__ void band_store_kernel(const float* __restrict__ src, int32_t channels, const int64_t* __restrict__ d_first_pixel,
                                  int64_t band_first, int64_t band_pixels, int64_t chunk, float* __restrict__ dst) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= chunk * channels) return;
  const int64_t i = e / channels;
  const int64_t pix = *d_first_pixel + i - band_first;
  if (pix < 0 || pix >= band_pixels) return;    // padding rays of the band's last chunk
  dst[pix * channels + (e - i * channels)] = src[e];
}

}  // namespace nrc

using namespace nrc;

extern "C" int32_t nrc_camera_rays(void* stream, const float* pixtocam, const float* camtoworld, int32_t width, int32_t height,
                                   int64_t first_pixel, const int64_t* d_first_pixel, int64_t last_pixel, int64_t num_rays,
                                   float near, float far, float* d_origins, float* d_directions, float* d_viewdirs,
                                   float* d_radii, float* d_imageplane, float* d_near, float* d_far) {
  if (!pixtocam || !camtoworld || width < 1 || height < 1 || num_rays < 0 || first_pixel < 0 || last_pixel < 0 ||
      last_pixel >= static_cast<int64_t>(width) * height)
    return NRC_E_INVALID_ARG;
  if (num_rays == 0) return NRC_OK;
  if (!d_origins || !d_directions || !d_viewdirs || !d_radii) return NRC_E_INVALID_ARG;
  CamParams p;
  for (int i = 0; i < 9; ++i) p.k[i] = pixtocam[i];
  for (int i = 0; i < 12; ++i) p.c2w[i] = camtoworld[i];
  p.width = width; p.first_pixel = first_pixel; p.last_pixel = last_pixel; p.near = near; p.far = far;
  camera_rays_kernel<<<static_cast<unsigned>((num_rays + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, d_first_pixel, num_rays, d_origins, d_directions, d_viewdirs, d_radii, d_imageplane, d_near, d_far);
  return check_launch();
}

extern "C" int32_t nrc_chunk_advance(void* stream, int64_t* d_counter, int64_t step) {
  if (!d_counter) return NRC_E_INVALID_ARG;
  chunk_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(d_counter, step);
  return check_launch();
}

extern "C" int32_t nrc_band_store(void* stream, const float* d_src, int32_t channels, const int64_t* d_first_pixel,
                                  int64_t band_first_pixel, int64_t band_pixels, int64_t chunk, float* d_band) {
  if (channels < 1 || chunk < 0 || band_pixels < 0 || !d_first_pixel) return NRC_E_INVALID_ARG;
  if (chunk == 0 || band_pixels == 0) return NRC_OK;
  if (!d_src || !d_band) return NRC_E_INVALID_ARG;
  const int64_t n = chunk * channels;
  band_store_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_src, channels, d_first_pixel, band_first_pixel, band_pixels, chunk, d_band);
  return check_launch();
}
