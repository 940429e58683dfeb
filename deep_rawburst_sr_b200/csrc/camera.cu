// Inverse camera pipeline of the synthetic burst generator (the parameter-deterministic part):
//   unprocess: invert_smoothstep -> gamma_expansion -> apply_ccm -> safe_invert_gains -> clamp
//              (reference data/synthetic_burst_generation.py:59-79 calling data/camera_pipeline.py:78-136), one pass;
//   mosaic + noise: RGGB mosaic -> shot / read noise -> clamp (synthetic_burst_generation.py:88-99, camera_pipeline.py:139-183).
// Both are HBM-bound element-wise passes (24 B and 16 / 32 B per pixel); the reference runs them as ~20 separate torch ops.
// The random affine warps / downsampling between the two (cv2.warpAffine / cv2.resize on uint8, fixed-point) are host code
// in the reference and stay out of scope.
#include "common.cuh"

#include <string.h>

namespace dbsr {

struct UnprocessParams {
  float ccm[9];      // rgb2cam, row-major
  float gains[3];    // [1 / red_gain, 1, 1 / blue_gain] / rgb_gain  (camera_pipeline.py:125)
  int smoothstep, gamma;
};

// per_image (optional, device): [batch][12] = rgb2cam (9, row-major) + gains (3) of every image; else the launch-wide p.ccm / p.gains
__global__ void __launch_bounds__(256) unprocess_kernel(const float* __restrict__ img, float* __restrict__ out, long long plane,
                                                        long long total, const UnprocessParams p, const float* __restrict__ per_image) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / plane, px = i - b * plane;
    float ccm[9], gains[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) ccm[k] = per_image ? __ldg(per_image + b * 12 + k) : p.ccm[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) gains[k] = per_image ? __ldg(per_image + b * 12 + 9 + k) : p.gains[k];
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
      cam[c] = __fadd_rn(__fadd_rn(__fmul_rn(ccm[3 * c], v[0]), __fmul_rn(ccm[3 * c + 1], v[1])), __fmul_rn(ccm[3 * c + 2], v[2]));
    // :121-136: gains masked near white so that saturated pixels are not dimmed
    const float gray = __fdiv_rn(__fadd_rn(__fadd_rn(cam[0], cam[1]), cam[2]), 3.0f);
    const float m0 = __fdiv_rn(fmaxf(__fsub_rn(gray, 0.9f), 0.0f), (float)(1.0 - 0.9));
    const float mask = __fmul_rn(m0, m0);
    float* dst = out + b * 3 * plane + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float g = gains[c];
      const float safe = fmaxf(__fadd_rn(mask, __fmul_rn(__fsub_rn(1.0f, mask), g)), g);
      dst[c * plane] = fminf(fmaxf(__fmul_rn(cam[c], safe), 0.0f), 1.0f);       // synthetic_burst_generation.py:79
    }
  }
}

// rgb [n, 3, h, w] -> raw [n, 4, h/2, w/2] = (R(0,0), G(0,1), G(1,0), B(1,1)) (+ z * sqrt(raw * shot + read)), clamped to [0, 1]
__global__ void __launch_bounds__(256) mosaic_noise_kernel(const float* __restrict__ rgb, const float* __restrict__ z, float* __restrict__ raw,
                                                           int h, int w, long long total, float shot, float read,
                                                           const float* __restrict__ levels, int per) {
  // levels (optional, device): [bursts][2] = (shot, read) noise level of every burst of `per` frames; else the launch-wide pair
  griddep_launch_dependents_if_small();
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
    if (levels) { shot = __ldg(levels + 2 * (n / per)); read = __ldg(levels + 2 * (n / per) + 1); }
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
  launch_pdl(unprocess_kernel, dim3(grid_for_cam(total)), dim3(256), 0, (cudaStream_t)stream, image, out, plane, total, p, (const float*)nullptr);
  return check_launch("unprocess_rgb");
}

extern "C" int dbsr_unprocess_rgb_batch(const float* image, float* out, int32_t batch, int32_t h, int32_t w, const float* params12,
                                        int32_t smoothstep, int32_t gamma, void* stream) {
  DBSR_REQUIRE(image && out && params12 && batch > 0 && h > 0 && w > 0, "unprocess_rgb_batch: bad arguments");
  UnprocessParams p;
  memset(&p, 0, sizeof(p));
  p.smoothstep = smoothstep; p.gamma = gamma;
  const long long plane = (long long)h * w, total = plane * batch;
  launch_pdl(unprocess_kernel, dim3(grid_for_cam(total)), dim3(256), 0, (cudaStream_t)stream, image, out, plane, total, p, params12);
  return check_launch("unprocess_rgb_batch");
}

extern "C" int dbsr_mosaic_noise(const float* rgb, const float* noise, float* raw, int32_t n, int32_t h, int32_t w, float shot_noise,
                                 float read_noise, void* stream) {
  DBSR_REQUIRE(rgb && raw && n > 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0, "mosaic_noise: needs even image sizes");
  const long long total = (long long)n * 4 * (h / 2) * (w / 2);
  launch_pdl(mosaic_noise_kernel, dim3(grid_for_cam(total)), dim3(256), 0, (cudaStream_t)stream, rgb, noise, raw, h, w, total, shot_noise,
             read_noise, (const float*)nullptr, 1);
  return check_launch("mosaic_noise");
}

extern "C" int dbsr_mosaic_noise_batch(const float* rgb, const float* noise, float* raw, int32_t n, int32_t h, int32_t w,
                                       const float* levels, int32_t frames_per_burst, void* stream) {
  DBSR_REQUIRE(rgb && raw && levels && n > 0 && frames_per_burst > 0 && n % frames_per_burst == 0 && h >= 2 && w >= 2 && h % 2 == 0 &&
                   w % 2 == 0, "mosaic_noise_batch: needs even image sizes and n = bursts * frames_per_burst");
  const long long total = (long long)n * 4 * (h / 2) * (w / 2);
  launch_pdl(mosaic_noise_kernel, dim3(grid_for_cam(total)), dim3(256), 0, (cudaStream_t)stream, rgb, noise, raw, h, w, total, 0.0f, 0.0f,
             levels, (int)frames_per_burst);
  return check_launch("mosaic_noise_batch");
}

// ---------------------------------------------------------------------------------------------------------
// single2lrburst (data/synthetic_burst_generation.py:131-246) for given frame transforms: uint8 quantisation of the image,
// cv2.warpAffine (8-bit, INTER_LINEAR, BORDER_CONSTANT), border crop, cv2.resize by 1 / factor (8-bit, INTER_LINEAR), / 255 --
// fused: only the warped pixels the resize reads are ever computed (2 x 2 of every factor x factor block for an even factor),
// straight from the fp32 image.  Byte / integer work: BIT-EXACT against OpenCV's fixed-point scheme (restated in
// oracle/lrburst_oracle.py: inverse map in 1/1024 px with separately rounded column / row terms + 16, reduced to 1/32 px; tap
// weights (32 - fx)(32 - fy) * 32 ..., (sum + 2^14) >> 15; resize (sum of the 2 x 2 block + 2) >> 2).  The flow vectors
// (sampling-position maps, :214-218, 229-246) are fp32.
// ---------------------------------------------------------------------------------------------------------
namespace dbsr {

struct LrBurstParams {
  const float* image;      // [3, H, W]
  const double* inv;       // [n][6] inverse affine maps (double, as cv::warpAffine inverts them)
  const float* pos;        // [n][6] fp32 inverse of the 3x3 forward matrix, rows 0..1 (torch .inverse(), :215)
  float* burst;            // [n, 3, h, w]
  float* flow;             // [n, 2, h, w] or null
  int H, W, n, f, crop, h, w, normalize;
  int per;                 // frames per source image: frame k warps image k / per, flows are relative to frame (k / per) * per
};

__device__ __forceinline__ int warped_u8x3(const LrBurstParams& p, const double* M, int x, int y, int out[3]) {
  // cv::warpAffine: X0 = round((M1 y + M2) 1024) + 16, adelta = round(M0 x 1024); 1/32 px after >> 5
  // (explicit IEEE mul / add: an FMA contraction would round the half-way cases differently from the host code)
  const double dx = (double)x, dy = (double)y;
  const long long X = (__double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], dy), M[2]), 1024.0)) + 16 +
                       __double2ll_rn(__dmul_rn(__dmul_rn(M[0], dx), 1024.0))) >> 5;
  const long long Y = (__double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], dy), M[5]), 1024.0)) + 16 +
                       __double2ll_rn(__dmul_rn(__dmul_rn(M[3], dx), 1024.0))) >> 5;
  long long sxl = X >> 5, syl = Y >> 5;
  sxl = sxl < -32768 ? -32768 : (sxl > 32767 ? 32767 : sxl);        // saturate_cast<short>
  syl = syl < -32768 ? -32768 : (syl > 32767 ? 32767 : syl);
  const int sx = (int)sxl, sy = (int)syl, fx = (int)(X & 31), fy = (int)(Y & 31);
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const bool x0 = sx >= 0 && sx < p.W, x1 = sx + 1 >= 0 && sx + 1 < p.W, y0 = sy >= 0 && sy < p.H, y1 = sy + 1 >= 0 && sy + 1 < p.H;
  const long long plane = (long long)p.H * p.W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* im = p.image + c * plane;      // p.image: this frame's source image (set by the caller)
    auto q = [&](bool ok, int yy, int xx) -> int {       // (image * 255).astype(uint8): fp32 product, truncation
      if (!ok) return 0;
      float v = __ldg(im + (long long)yy * p.W + xx);
      if (p.normalize) v = __fmul_rn(v, 255.0f);
      return (int)fminf(fmaxf(v, 0.0f), 255.0f);
    };
    const int acc = q(y0 && x0, sy, sx) * w00 + q(y0 && x1, sy, sx + 1) * w01 + q(y1 && x0, sy + 1, sx) * w10 + q(y1 && x1, sy + 1, sx + 1) * w11;
    out[c] = (acc + 16384) >> 15;
  }
  return 0;
}

__global__ void __launch_bounds__(256) lrburst_kernel(const LrBurstParams p) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const long long total = (long long)p.n * p.h * p.w;
  const int k = (p.f & 1) ? 1 : 2, o = (p.f & 1) ? (p.f - 1) / 2 : p.f / 2 - 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % p.w);
    const long long t = i / p.w;
    const int oy = (int)(t % p.h), frame = (int)(t / p.h);
    const double* M = p.inv + 6 * frame;
    LrBurstParams q = p;
    q.image = p.image + (long long)(frame / p.per) * 3 * p.H * p.W;
    const int bx = p.crop + p.f * ox + o, by = p.crop + p.f * oy + o;
    int s[3] = {0, 0, 0};
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        int v[3];
        warped_u8x3(q, M, bx + dx, by + dy, v);
        s[0] += v[0]; s[1] += v[1]; s[2] += v[2];
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int v = k == 2 ? (s[c] + 2) >> 2 : s[c];
      const float fv = (float)v;
      p.burst[(((long long)frame * 3 + c) * p.h + oy) * p.w + ox] = p.normalize ? __fdiv_rn(fv, 255.0f) : fv;
    }
    if (p.flow) {
      float d[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        float pos[2];
#pragma unroll
        for (int which = 0; which < 2; ++which) {          // this frame, frame 0
          const float* T = p.pos + 6 * (which == 0 ? frame : (frame / p.per) * p.per) + 3 * a;
          auto at = [&](int xx, int yy) { return __fadd_rn(__fadd_rn(__fmul_rn((float)xx, T[0]), __fmul_rn((float)yy, T[1])), T[2]); };
          float r;
          if (k == 2) {
            const float h0 = __fadd_rn(__fmul_rn(at(bx, by), 0.5f), __fmul_rn(at(bx + 1, by), 0.5f));
            const float h1 = __fadd_rn(__fmul_rn(at(bx, by + 1), 0.5f), __fmul_rn(at(bx + 1, by + 1), 0.5f));
            r = __fadd_rn(__fmul_rn(h0, 0.5f), __fmul_rn(h1, 0.5f));
          } else {
            r = at(bx, by);
          }
          pos[which] = __fdiv_rn(r, (float)p.f);
        }
        d[a] = __fsub_rn(pos[0], pos[1]);
      }
      p.flow[(((long long)frame * 2 + 0) * p.h + oy) * p.w + ox] = d[0];
      p.flow[(((long long)frame * 2 + 1) * p.h + oy) * p.w + ox] = d[1];
    }
  }
}

}  // namespace dbsr

extern "C" int dbsr_single2lrburst(const float* image, int32_t H, int32_t W, const double* inverse_maps, const float* position_maps,
                                   int32_t n, int32_t factor, int32_t border_crop, int32_t normalize, float* burst, float* flow,
                                   void* stream) {
  DBSR_REQUIRE(image && inverse_maps && burst && H > 0 && W > 0 && n > 0 && factor >= 1 && border_crop >= 0, "single2lrburst: bad arguments");
  DBSR_REQUIRE(!flow || position_maps, "single2lrburst: flow vectors need the fp32 position maps");
  const int hc = H - 2 * border_crop, wc = W - 2 * border_crop;
  DBSR_REQUIRE(hc > 0 && wc > 0 && hc % factor == 0 && wc % factor == 0,
               "single2lrburst: the cropped image (%d x %d) must be a multiple of the down-sampling factor %d", hc, wc, factor);
  LrBurstParams p;
  p.image = image; p.inv = inverse_maps; p.pos = position_maps; p.burst = burst; p.flow = flow;
  p.H = H; p.W = W; p.n = n; p.f = factor; p.crop = border_crop; p.h = hc / factor; p.w = wc / factor; p.normalize = normalize;
  p.per = n;
  const long long total = (long long)n * p.h * p.w;
  launch_pdl(lrburst_kernel, dim3(grid_for_cam(total)), dim3(256), 0, (cudaStream_t)stream, p);
  return check_launch("single2lrburst");
}

extern "C" int dbsr_single2lrburst_batch(const float* images, int32_t batch, int32_t H, int32_t W, const double* inverse_maps,
                                         const float* position_maps, int32_t frames_per_image, int32_t factor, int32_t border_crop,
                                         int32_t normalize, float* burst, float* flow, void* stream) {
  DBSR_REQUIRE(images && inverse_maps && burst && batch > 0 && H > 0 && W > 0 && frames_per_image > 0 && factor >= 1 && border_crop >= 0,
               "single2lrburst_batch: bad arguments");
  DBSR_REQUIRE(!flow || position_maps, "single2lrburst_batch: flow vectors need the fp32 position maps");
  const int hc = H - 2 * border_crop, wc = W - 2 * border_crop;
  DBSR_REQUIRE(hc > 0 && wc > 0 && hc % factor == 0 && wc % factor == 0,
               "single2lrburst_batch: the cropped image (%d x %d) must be a multiple of the down-sampling factor %d", hc, wc, factor);
  LrBurstParams p;
  p.image = images; p.inv = inverse_maps; p.pos = position_maps; p.burst = burst; p.flow = flow;
  p.H = H; p.W = W; p.n = batch * frames_per_image; p.f = factor; p.crop = border_crop; p.h = hc / factor; p.w = wc / factor;
  p.normalize = normalize; p.per = frames_per_image;
  const long long total = (long long)p.n * p.h * p.w;
  launch_pdl(lrburst_kernel, dim3(grid_for_cam(total)), dim3(256), 0, (cudaStream_t)stream, p);
  return check_launch("single2lrburst_batch");
}
