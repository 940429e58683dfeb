// Inverse camera pipeline of the synthetic burst generator (the parameter-deterministic part):
//   unprocess: invert_smoothstep -> gamma_expansion -> apply_ccm -> safe_invert_gains -> clamp
//              (reference data/synthetic_burst_generation.py:59-79 calling data/camera_pipeline.py:78-136), one pass;
//   mosaic + noise: RGGB mosaic -> shot / read noise -> clamp (synthetic_burst_generation.py:88-99, camera_pipeline.py:139-183).
// Both are HBM-bound element-wise passes (24 B and 16 / 32 B per pixel); the reference runs them as ~20 separate torch ops.
// The random affine warps / downsampling between the two (cv2.warpAffine / cv2.resize on uint8, fixed-point) are host code
// in the reference and stay out of scope.
#include "common.cuh"

namespace dbsr {

struct UnprocessParams {
  float ccm[9];      // rgb2cam, row-major
  float gains[3];    // [1 / red_gain, 1, 1 / blue_gain] / rgb_gain  (camera_pipeline.py:125)
  int smoothstep, gamma;
};

__global__ void __launch_bounds__(256) unprocess_kernel(const float* __restrict__ img, float* __restrict__ out, long long plane,
                                                        long long total, const UnprocessParams p) {
  griddep_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / plane, px = i - b * plane;
    const float* src = img + b * 3 * plane + px;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = __ldg(src + c * plane);
      if (p.smoothstep) {                                   // camera_pipeline.py:78-81
        x = fminf(fmaxf(x, 0.0f), 1.0f);
        x = 0.5f - sinf(asinf(1.0f - 2.0f * x) / 3.0f);
      }
      if (p.gamma) x = powf(fmaxf(x, 1e-8f), 2.2f);         // :84-87
      v[c] = x;
    }
    float cam[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)                              // :96-107 (torch.mm, k ascending)
      cam[c] = __fadd_rn(__fadd_rn(__fmul_rn(p.ccm[3 * c], v[0]), __fmul_rn(p.ccm[3 * c + 1], v[1])), __fmul_rn(p.ccm[3 * c + 2], v[2]));
    // :121-136: gains masked near white so that saturated pixels are not dimmed
    const float gray = __fdiv_rn(__fadd_rn(__fadd_rn(cam[0], cam[1]), cam[2]), 3.0f);
    const float m0 = __fdiv_rn(fmaxf(__fsub_rn(gray, 0.9f), 0.0f), (float)(1.0 - 0.9));
    const float mask = __fmul_rn(m0, m0);
    float* dst = out + b * 3 * plane + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float g = p.gains[c];
      const float safe = fmaxf(__fadd_rn(mask, __fmul_rn(__fsub_rn(1.0f, mask), g)), g);
      dst[c * plane] = fminf(fmaxf(__fmul_rn(cam[c], safe), 0.0f), 1.0f);       // synthetic_burst_generation.py:79
    }
  }
}

// rgb [n, 3, h, w] -> raw [n, 4, h/2, w/2] = (R(0,0), G(0,1), G(1,0), B(1,1)) (+ z * sqrt(raw * shot + read)), clamped to [0, 1]
__global__ void __launch_bounds__(256) mosaic_noise_kernel(const float* __restrict__ rgb, const float* __restrict__ z, float* __restrict__ raw,
                                                           int h, int w, long long total, float shot, float read) {
  griddep_wait();
  const int h2 = h / 2, w2 = w / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w2);
    long long t = i / w2;
    const int y = (int)(t % h2);
    t /= h2;
    const int ch = (int)(t & 3);
    const long long n = t >> 2;
    const int src_c = ch == 0 ? 0 : (ch == 3 ? 2 : 1);
    const int sy = 2 * y + (ch >> 1), sx = 2 * x + (ch & 1);
    float v = __ldg(rgb + ((n * 3 + src_c) * h + sy) * (long long)w + sx);
    if (z) v = __fadd_rn(v, __fmul_rn(__ldg(z + i), sqrtf(__fadd_rn(__fmul_rn(v, shot), read))));    // camera_pipeline.py:178-183
    raw[i] = fminf(fmaxf(v, 0.0f), 1.0f);
  }
}

static inline int grid_for_cam(long long total) {
  long long b = (total + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace dbsr

using namespace dbsr;

extern "C" int dbsr_unprocess_rgb(const float* image, float* out, int32_t batch, int32_t h, int32_t w, const float* rgb2cam9,
                                  const float* gains3, int32_t smoothstep, int32_t gamma, void* stream) {
  DBSR_REQUIRE(image && out && rgb2cam9 && gains3 && batch > 0 && h > 0 && w > 0, "unprocess_rgb: bad arguments");
  UnprocessParams p;
  for (int i = 0; i < 9; ++i) p.ccm[i] = rgb2cam9[i];
  for (int i = 0; i < 3; ++i) p.gains[i] = gains3[i];
  p.smoothstep = smoothstep; p.gamma = gamma;
  const long long plane = (long long)h * w, total = plane * batch;
  launch_pdl(unprocess_kernel, dim3(grid_for_cam(total)), dim3(256), 0, (cudaStream_t)stream, image, out, plane, total, p);
  return check_launch("unprocess_rgb");
}

extern "C" int dbsr_mosaic_noise(const float* rgb, const float* noise, float* raw, int32_t n, int32_t h, int32_t w, float shot_noise,
                                 float read_noise, void* stream) {
  DBSR_REQUIRE(rgb && raw && n > 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0, "mosaic_noise: needs even image sizes");
  const long long total = (long long)n * 4 * (h / 2) * (w / 2);
  launch_pdl(mosaic_noise_kernel, dim3(grid_for_cam(total)), dim3(256), 0, (cudaStream_t)stream, rgb, noise, raw, h, w, total, shot_noise,
             read_noise);
  return check_launch("mosaic_noise");
}
