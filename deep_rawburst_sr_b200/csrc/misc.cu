// Error plumbing, probes, layout converters and the small HBM-bound kernels of the DBSR forward path
// (burst preparation, transposed conv, flow head, offsets modulo, weight-predictor input, blur, predictor).
#include "common.cuh"

#include <string.h>

namespace dbsr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

// -------------------------------------------------------------------------------------------------------
// NCHW fp32 <-> NHWC view.  Tiled transpose through shared memory so both sides stay coalesced.
// -------------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, View dst) {
  __shared__ float tile[32][33];
  const int HW = dst.h * dst.w;
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < dst.c && p < HW) ? __ldg(src + ((long long)n * dst.c + c) * HW + p) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < dst.c && p < HW) view_st(dst, (long long)n * HW + p, c, tile[threadIdx.x][i]);
  }
}

__global__ void nhwc_to_nchw_kernel(View src, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int HW = src.h * src.w;
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < src.c && p < HW) ? view_ld(src, (long long)n * HW + p, c) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < src.c && p < HW) dst[((long long)n * src.c + c) * HW + p] = tile[threadIdx.x][i];
  }
}

__global__ void copy_channels_kernel(View src, View dst, int group, int src_group, int src_first) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  const int HW = dst.h * dst.w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dst.c);
    const long long pix = i / dst.c;
    const int n = (int)(pix / HW);
    const int rem = (int)(pix - (long long)n * HW);
    const int sn = group > 0 ? (n / group) * src_group + src_first : n;
    view_st(dst, pix, c, view_ld(src, (long long)sn * HW + rem, c));
  }
}

// same dtype, channel counts / offsets / pitches multiples of 8 (bf16) and 16-byte aligned views: one thread moves 16
// bytes; grid.y walks the images so that the per-thread decode is 32-bit (the generic kernel above spends its time in
// 64-bit divisions and per-element dtype dispatch)
__global__ void copy_channels_v8_kernel(View src, View dst, int group, int src_group, int src_first) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int HW = dst.h * dst.w, C8 = dst.c >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)(HW * C8)) return;
  const int rem = (int)(idx / (unsigned)C8), c8 = (int)(idx - (unsigned)rem * (unsigned)C8);
  const __nv_bfloat16* sb = reinterpret_cast<const __nv_bfloat16*>(src.data) + src.c_off + c8 * 8;
  __nv_bfloat16* db = reinterpret_cast<__nv_bfloat16*>(dst.data) + dst.c_off + c8 * 8;
  for (int n = blockIdx.y; n < dst.n; n += gridDim.y) {
    const int sn = group > 0 ? (n / group) * src_group + src_first : n;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(sb + ((long long)sn * HW + rem) * src.c_pitch));
    *reinterpret_cast<uint4*>(db + ((long long)n * HW + rem) * dst.c_pitch) = v;
  }
}

// -------------------------------------------------------------------------------------------------------
// Space to depth (2x2): y[n, Y, X, (p*2 + q)*C + c] = x[n, 2Y + p, 2X + q, c]  (zero beyond the image: odd sizes).
// A 3x3 / stride-2 / pad-1 convolution of x is then a 3x3 / stride-1 / pad-1 convolution of y whose taps
// (ky', kx') in {0, 1}^2 carry the weights (engine.pack_s2d_weight) -- which is how the PWC-Net extractor's
// stride-2 layers (pwcnet.py:49-97) reach the tensor-core kernel.  Converts dtype on the way.
// -------------------------------------------------------------------------------------------------------
__global__ void space_to_depth2_kernel(View x, View y) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int C = x.c, C4 = 4 * C;
  const long long total = (long long)y.n * y.h * y.w * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % C4);
    const long long pix = i / C4;
    const int X = (int)(pix % y.w);
    const long long t = pix / y.w;
    const int Y = (int)(t % y.h), n = (int)(t / y.h);
    const int pq = k / C, c = k - pq * C;
    const int sy = 2 * Y + (pq >> 1), sx = 2 * X + (pq & 1);
    float v = 0.0f;
    if (sy < x.h && sx < x.w) v = view_ld(x, ((long long)n * x.h + sy) * x.w + sx, c);
    view_st(y, pix, k, v);
  }
}
// bf16 -> bf16 with C % 8 == 0 and 16-byte aligned views: one thread moves 8 channels (16 bytes)
__global__ void space_to_depth2_v8_kernel(View x, View y) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int C8 = x.c >> 3, G = 4 * C8;
  const long long total = (long long)y.n * y.h * y.w * G;
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x.data) + x.c_off;
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y.data) + y.c_off;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % G);
    const long long pix = i / G;
    const int X = (int)(pix % y.w);
    const long long t = pix / y.w;
    const int Y = (int)(t % y.h), n = (int)(t / y.h);
    const int pq = k / C8, c8 = k - pq * C8;
    const int sy = 2 * Y + (pq >> 1), sx = 2 * X + (pq & 1);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (sy < x.h && sx < x.w) v = __ldg(reinterpret_cast<const uint4*>(xb + (((long long)n * x.h + sy) * x.w + sx) * x.c_pitch + c8 * 8));
    *reinterpret_cast<uint4*>(yb + pix * y.c_pitch + pq * x.c + c8 * 8) = v;
  }
}

// -------------------------------------------------------------------------------------------------------
// Burst preparation.  One thread per output pixel of either destination.
//   enc_in: channels-last copy of the packed RAW frame (zero-padded channels)
//   pwc_in: RGGB->RGB (encoders.py:52) then bilinear resize to (Hp, Wp), align_corners=False
//           (pwcnet.py:266-271): src = (dst + 0.5) * in/out - 0.5, clamped at 0, neighbour clamped.
// -------------------------------------------------------------------------------------------------------
__global__ void prep_burst_kernel(const float* __restrict__ burst, int H, int W, View enc_in, View pwc_in) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int HW = H * W;
  const long long n_enc = (long long)enc_in.n * HW;
  const long long n_pwc = (long long)pwc_in.n * pwc_in.h * pwc_in.w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_enc + n_pwc;
       i += (long long)gridDim.x * blockDim.x) {
    if (i < n_enc) {
      const int f = (int)(i / HW);
      const int rem = (int)(i - (long long)f * HW);
      const float* b = burst + (long long)f * 4 * HW + rem;
      for (int c = 0; c < enc_in.c; ++c) view_st(enc_in, i, c, c < 4 ? __ldg(b + (long long)c * HW) : 0.0f);
    } else {
      const long long j = i - n_enc;
      const int Hp = pwc_in.h, Wp = pwc_in.w;
      const int f = (int)(j / ((long long)Hp * Wp));
      const int rem = (int)(j - (long long)f * Hp * Wp);
      const int oy = rem / Wp, ox = rem - oy * Wp;
      const float sy = (float)H / (float)Hp, sx = (float)W / (float)Wp;
      float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.0f);
      float fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.0f);
      int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
      const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
      const float wy = fy - (float)y0, wx = fx - (float)x0;
      const float* b = burst + (long long)f * 4 * HW;
      float rgb[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v[4];
        const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int pix = ys[t >> 1] * W + xs[t & 1];
          if (c == 0) v[t] = __ldg(b + pix);
          else if (c == 1) v[t] = (__ldg(b + HW + pix) + __ldg(b + 2 * HW + pix)) / 2.0f;
          else v[t] = __ldg(b + 3 * HW + pix);
        }
        // same association as ATen's upsample_bilinear2d: rows first (wy), then columns (wx)
        const float top = v[0] * (1.0f - wx) + v[1] * wx;
        const float bot = v[2] * (1.0f - wx) + v[3] * wx;
        rgb[c] = top * (1.0f - wy) + bot * wy;
      }
      for (int c = 0; c < pwc_in.c; ++c) view_st(pwc_in, j, c, c < 3 ? rgb[c] : 0.0f);
    }
  }
}

// The same preparation for the bf16 tensor-core PWC-Net path, with the space-to-depth of the extractor's first stride-2
// conv folded in: instead of the fp32 [Hp, Wp, 4] image (29 MB at B=32) and a second kernel that re-reads it, the
// resized RGB image is written directly as pwc_s2d[n, Y, X, (2p + q) * 3 + c] = rgb[n, 2Y + p, 2X + q, c] in bf16
// (12 channels + 4 zero pad = two 16-byte stores per thread).  grid.y walks the frames (32-bit decode per thread).
__device__ __forceinline__ void resized_rgb(const float* __restrict__ b, int H, int W, int HW, float sy, float sx, int oy, int ox,
                                            float (&rgb)[3]) {
  const float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.0f);
  const float fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.0f);
  const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float wy = fy - (float)y0, wx = fx - (float)x0;
  const int p00 = y0 * W + x0, p01 = y0 * W + x1, p10 = y1 * W + x0, p11 = y1 * W + x1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4];
    const int pix[4] = {p00, p01, p10, p11};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (c == 0) v[t] = __ldg(b + pix[t]);
      else if (c == 1) v[t] = (__ldg(b + HW + pix[t]) + __ldg(b + 2 * HW + pix[t])) / 2.0f;
      else v[t] = __ldg(b + 3 * HW + pix[t]);
    }
    const float top = v[0] * (1.0f - wx) + v[1] * wx;       // same association as prep_burst_kernel / ATen
    const float bot = v[2] * (1.0f - wx) + v[3] * wx;
    rgb[c] = top * (1.0f - wy) + bot * wy;
  }
}
__device__ __forceinline__ uint32_t bf16x2_bits(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__global__ void __launch_bounds__(256) prep_burst_s2d_kernel(const float* __restrict__ burst, int H, int W, int Hp, int Wp,
                                                             View enc_in, View pwc_s2d) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int HW = H * W, H2 = Hp >> 1, W2 = Wp >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= HW + H2 * W2) return;
  const bool enc_vec = enc_in.dtype == DBSR_BF16 && enc_in.c_off == 0 && enc_in.c_pitch == 8 && ((uintptr_t)enc_in.data % 16) == 0;
  const float sy = (float)H / (float)Hp, sx = (float)W / (float)Wp;
  for (int f = blockIdx.y; f < enc_in.n; f += gridDim.y) {
    const float* b = burst + (long long)f * 4 * HW;
    if (idx < HW) {
      const float v0 = __ldg(b + idx), v1 = __ldg(b + HW + idx), v2 = __ldg(b + 2 * HW + idx), v3 = __ldg(b + 3 * HW + idx);
      const long long i = (long long)f * HW + idx;
      if (enc_vec) {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(enc_in.data) + i * 8) =
            make_uint4(bf16x2_bits(v0, v1), bf16x2_bits(v2, v3), 0u, 0u);
      } else {
        const float v[4] = {v0, v1, v2, v3};
        for (int c = 0; c < enc_in.c; ++c) view_st(enc_in, i, c, c < 4 ? v[c] : 0.0f);
      }
    } else {
      const int j = idx - HW;
      const int Y = j / W2, X = j - Y * W2;
      float r[4][3];
#pragma unroll
      for (int pq = 0; pq < 4; ++pq) resized_rgb(b, H, W, HW, sy, sx, 2 * Y + (pq >> 1), 2 * X + (pq & 1), r[pq]);
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(pwc_s2d.data) + ((long long)f * H2 * W2 + j) * 16);
      dst[0] = make_uint4(bf16x2_bits(r[0][0], r[0][1]), bf16x2_bits(r[0][2], r[1][0]), bf16x2_bits(r[1][1], r[1][2]),
                          bf16x2_bits(r[2][0], r[2][1]));
      dst[1] = make_uint4(bf16x2_bits(r[2][2], r[3][0]), bf16x2_bits(r[3][1], r[3][2]), 0u, 0u);
    }
  }
}

// -------------------------------------------------------------------------------------------------------
// ConvTranspose2d(k=4, s=2, p=1), Cout = 2 (pwcnet.py:119-120).  One warp per output pixel; lanes stride
// over input channels (coalesced NHWC reads), 2x2 valid taps per output parity, shuffle reduction.
//   out[oy,ox,oc] = b[oc] + sum_{ic} sum_{ky = (oy+1)&1 (+2)} sum_{kx} x[(oy+1-ky)/2, (ox+1-kx)/2, ic] w[ky][kx][oc][ic]
// -------------------------------------------------------------------------------------------------------
__global__ void deconv4x4s2_kernel(View x, const float* __restrict__ w, const float* __restrict__ bias, View y,
                                   View y2) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int Ho = 2 * x.h, Wo = 2 * x.w;
  const long long total = (long long)x.n * Ho * Wo;
  if (warp >= total) return;
  const int n = (int)(warp / ((long long)Ho * Wo));
  const int rem = (int)(warp - (long long)n * Ho * Wo);
  const int oy = rem / Wo, ox = rem - oy * Wo;
  float acc0 = 0.0f, acc1 = 0.0f;
  const int Cin = x.c;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int ky = ((oy + 1) & 1) + 2 * a;
    const int iy2 = oy + 1 - ky;
    if (iy2 < 0) continue;
    const int iy = iy2 >> 1;
    if (iy >= x.h) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int kx = ((ox + 1) & 1) + 2 * b;
      const int ix2 = ox + 1 - kx;
      if (ix2 < 0) continue;
      const int ix = ix2 >> 1;
      if (ix >= x.w) continue;
      const long long pix = ((long long)n * x.h + iy) * x.w + ix;
      const float* w0 = w + ((ky * 4 + kx) * 2 + 0) * (long long)Cin;
      const float* w1 = w0 + Cin;
      for (int ic = lane; ic < Cin; ic += 32) {
        const float v = view_ld(x, pix, ic);
        acc0 = fmaf(v, __ldg(w0 + ic), acc0);
        acc1 = fmaf(v, __ldg(w1 + ic), acc1);
      }
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc0 += __shfl_xor_sync(0xffffffffu, acc0, s);
    acc1 += __shfl_xor_sync(0xffffffffu, acc1, s);
  }
  if (lane == 0) {
    acc0 += bias[0];
    acc1 += bias[1];
    view_st(y, warp, 0, acc0);
    view_st(y, warp, 1, acc1);
    if (y2.data) {
      view_st(y2, warp, 0, acc0);
      view_st(y2, warp, 1, acc1);
    }
  }
}

// -------------------------------------------------------------------------------------------------------
// Second half of ConvTranspose2d(k=4, s=2, p=1), Cout = 2 when the channel contraction ran as a 1x1 convolution:
//   taps[n, iy, ix, (ky*4 + kx)*2 + oc] = sum_ic x[n, iy, ix, ic] * w[ic, oc, ky, kx]        (tensor cores / conv_direct)
//   out[n, oy, ox, oc] = b[oc] + sum over the (up to) 2x2 (ky, kx) with iy = (oy + 1 - ky) / 2, ix = (ox + 1 - kx) / 2
// The 2-channel flow deconvolution of the same level (netUpflow) is computed directly in the same pass.
// One thread per output pixel.
// -------------------------------------------------------------------------------------------------------
// A 3x3 convolution with TWO output channels whose channel contraction ran as a 1x1 convolution Cin -> 18 planes
//   ftaps[n, y, x, (ky*3 + kx)*2 + oc] = sum_ic x[n, y, x, ic] * w[oc, ic, ky, kx]
// (the flow heads of PWC-Net's decoders, pwcnet.py:150: nine N = 16 tensor-core MMAs per K step become one):
//   out[n, y, x, oc] = b[oc] + sum over the taps inside the map of ftaps[n, y + ky - 1, x + kx - 1, (ky*3 + kx)*2 + oc]
__device__ __forceinline__ void flow_from_taps(const View& ftaps, const float* __restrict__ bias6, int n, int h, int w, int y,
                                               int x, float& u, float& v) {
  u = bias6[0]; v = bias6[1];
  // fp32 planes whose (oc = 0, 1) pair is 8-byte aligned (the engine's layout): one 8-byte load per tap
  const bool pair_ld = ftaps.dtype == DBSR_F32 && ((ftaps.c_off | ftaps.c_pitch) & 1) == 0 &&
                       (reinterpret_cast<uintptr_t>(ftaps.data) & 7) == 0;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= h) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x + kx - 1;
      if (xx < 0 || xx >= w) continue;
      const long long q = ((long long)n * h + yy) * w + xx;
      if (pair_ld) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(ftaps.data) + q * ftaps.c_pitch +
                                                                ftaps.c_off + (ky * 3 + kx) * 2));
        u += t.x; v += t.y;
      } else {
        u += view_ld(ftaps, q, (ky * 3 + kx) * 2 + 0);
        v += view_ld(ftaps, q, (ky * 3 + kx) * 2 + 1);
      }
    }
  }
}

__global__ void flow_from_taps_kernel(View ftaps, const float* __restrict__ bias6, View y) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int h = ftaps.h, w = ftaps.w;
  const long long total = (long long)ftaps.n * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ((long long)h * w));
    const int rem = (int)(i - (long long)n * h * w);
    float u, v;
    flow_from_taps(ftaps, bias6, n, h, w, rem / w, rem % w, u, v);
    view_st(y, i, 0, u); view_st(y, i, 1, v);
  }
}

// ftaps.data != NULL: the flow is not read from `flow` but summed from the 18 tap planes of its convolution on the fly
// (the flow map of the coarser level is then never written)
__global__ void deconv_col2im_kernel(View taps, const float* __restrict__ bias_t, View y_t, View flow,
                                     const float* __restrict__ wf, const float* __restrict__ bias_f, View y_f, View y_f2,
                                     View ftaps, const float* __restrict__ bias6) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int h = taps.h, w = taps.w, Ho = 2 * h, Wo = 2 * w;
  const long long total = (long long)taps.n * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ((long long)Ho * Wo));
    const int rem = (int)(i - (long long)n * Ho * Wo);
    const int oy = rem / Wo, ox = rem - oy * Wo;
    float t0 = bias_t[0], t1 = bias_t[1];
    float f0 = 0.0f, f1 = 0.0f;
    const bool from_taps = ftaps.data != nullptr;
    const bool has_flow = flow.data != nullptr || from_taps;
    if (has_flow) { f0 = bias_f[0]; f1 = bias_f[1]; }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ((oy + 1) & 1) + 2 * a;
      const int iy2 = oy + 1 - ky;
      if (iy2 < 0 || (iy2 >> 1) >= h) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kx = ((ox + 1) & 1) + 2 * b;
        const int ix2 = ox + 1 - kx;
        if (ix2 < 0 || (ix2 >> 1) >= w) continue;
        const long long pix = ((long long)n * h + (iy2 >> 1)) * w + (ix2 >> 1);
        const int tap = ky * 4 + kx;
        t0 += view_ld(taps, pix, tap * 2 + 0);
        t1 += view_ld(taps, pix, tap * 2 + 1);
        if (has_flow) {
          float u, v;
          if (from_taps) flow_from_taps(ftaps, bias6, n, h, w, iy2 >> 1, ix2 >> 1, u, v);
          else { u = view_ld(flow, pix, 0); v = view_ld(flow, pix, 1); }
          const float* q = wf + tap * 4;      // [ky][kx][oc][ic]
          f0 = fmaf(u, __ldg(q + 0), fmaf(v, __ldg(q + 1), f0));
          f1 = fmaf(u, __ldg(q + 2), fmaf(v, __ldg(q + 3), f1));
        }
      }
    }
    view_st(y_t, i, 0, t0); view_st(y_t, i, 1, t1);
    if (has_flow) {
      view_st(y_f, i, 0, f0); view_st(y_f, i, 1, f1);
      if (y_f2.data) { view_st(y_f2, i, 0, f0); view_st(y_f2, i, 1, f1); }
    }
  }
}

// -------------------------------------------------------------------------------------------------------
// Flow head (pwcnet.py:274-279): bilinear resize of the quarter-resolution flow to (H, W) with
// align_corners=False, x20, x(W/Wp, H/Hp).  Output NCHW fp32 (the public `offsets`).
// -------------------------------------------------------------------------------------------------------
__global__ void flow_head_kernel(View f4, float* __restrict__ offsets, int H, int W, float mulx, float muly) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const long long total = (long long)f4.n * H * W;
  const int h4 = f4.h, w4 = f4.w;
  const float sy = (float)h4 / (float)H, sx = (float)w4 / (float)W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ((long long)H * W));
    const int rem = (int)(i - (long long)n * H * W);
    const int oy = rem / W, ox = rem - oy * W;
    const float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.0f);
    const float fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.0f);
    const int y0 = min((int)fy, h4 - 1), x0 = min((int)fx, w4 - 1);
    const int y1 = min(y0 + 1, h4 - 1), x1 = min(x0 + 1, w4 - 1);
    const float wy = fy - (float)y0, wx = fx - (float)x0;
    const long long base = (long long)n * h4 * w4;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float v00 = view_ld(f4, base + y0 * w4 + x0, c), v01 = view_ld(f4, base + y0 * w4 + x1, c);
      const float v10 = view_ld(f4, base + y1 * w4 + x0, c), v11 = view_ld(f4, base + y1 * w4 + x1, c);
      const float top = v00 * (1.0f - wx) + v01 * wx;
      const float bot = v10 * (1.0f - wx) + v11 * wx;
      const float v = 20.0f * (top * (1.0f - wy) + bot * wy);
      offsets[((long long)n * 2 + c) * H * W + rem] = v * (c == 0 ? mulx : muly);
    }
  }
}

// merging.py:91-105: zeros for the reference frame, floor-mod for the others
__global__ void offsets_mod_kernel(const float* __restrict__ offsets, View out, int frames, float modulo) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int HW = out.h * out.w;
  const int rem = blockIdx.x * blockDim.x + threadIdx.x;
  if (rem >= HW) return;
  // 8-channel bf16 rows (the engine's layout): the two offsets and six zero channels are one 16-byte store
  const bool vec = out.dtype == DBSR_BF16 && out.c == 8 && out.c_off == 0 && out.c_pitch == 8 && ((uintptr_t)out.data % 16) == 0;
  for (int img = blockIdx.y; img < out.n; img += gridDim.y) {
    const int b = img / frames, f = img - b * frames;
    float v[2] = {0.0f, 0.0f};
    if (f > 0) {
      const long long p = (long long)b * (frames - 1) + (f - 1);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float a = __ldg(offsets + (p * 2 + c) * HW + rem);
        float r = a;
        if (modulo > 0.0f) {
          r = fmodf(a, modulo);               // torch.remainder: fmod, then fix the sign
          if (r != 0.0f && r < 0.0f) r += modulo;
        }
        v[c] = r;
      }
    }
    const long long i = (long long)img * HW + rem;
    if (vec) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out.data) + i * 8) =
          make_uint4(*reinterpret_cast<const uint32_t*>(&h), 0u, 0u, 0u);
    } else {
      for (int c = 0; c < out.c; ++c) view_st(out, i, c, c < 2 ? v[c] : 0.0f);
    }
  }
}

// merging.py:79-89: [base | diff] channels of the weight-predictor input
__global__ void build_wp_input_kernel(View proj, View wp_in, int frames) {
  const int C = proj.c;
  const int HW = proj.h * proj.w;
  const long long total = (long long)proj.n * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int img = (int)(pix / HW);
    const int rem = (int)(pix - (long long)img * HW);
    const int b = img / frames;
    const float base = view_ld(proj, (long long)b * frames * HW + rem, c);
    const float mine = view_ld(proj, pix, c);
    view_st(wp_in, pix, c, base);
    view_st(wp_in, pix, C + c, mine - base);
  }
}

// decoders.py:52,60: 1x1 conv to `cout` (<= 4) channels + ReLU, NCHW fp32 output.  One thread per pixel: the pixel's
// channels are read with 16-byte loads (they are contiguous in NHWC), the `cout` planes are written coalesced.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) predictor_kernel(View x, const float* __restrict__ w, const float* __restrict__ bias,
                                                        int cout, float* __restrict__ pred) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  extern __shared__ float ws[];  // [cout][C] + [cout]
  const int C = x.c;
  for (int i = threadIdx.x; i < cout * C; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) ws[cout * C + i] = bias ? bias[i] : 0.0f;
  __syncthreads();
  const int HW = x.h * x.w;
  const long long total = (long long)x.n * HW;
  const T* xb = reinterpret_cast<const T*>(x.data) + x.c_off;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const T* px = xb + i * x.c_pitch;
    if (VEC) {
      for (int c = 0; c < C; c += 8) {
        float v[8];
        if (sizeof(T) == 2) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(px + c));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
        } else {
          const float4 a = __ldg(reinterpret_cast<const float4*>(px + c));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(px + c) + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < cout) acc[o] = fmaf(v[k], ws[o * C + c + k], acc[o]);
      }
    } else {
      for (int c = 0; c < C; ++c) {
        const float v = view_ld(x, i, c);
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o < cout) acc[o] = fmaf(v, ws[o * C + c], acc[o]);
      }
    }
    const int n = (int)(i / HW);
    const int rem = (int)(i - (long long)n * HW);
#pragma unroll
    for (int o = 0; o < 4; ++o)
      if (o < cout) pred[((long long)n * cout + o) * HW + rem] = fmaxf(acc[o] + ws[cout * C + o], 0.0f);
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dbsr

using namespace dbsr;

extern "C" int dbsr_version(void) { return DBSR_B200_VERSION; }
extern "C" const char* dbsr_last_error(void) { return g_err; }

extern "C" int dbsr_device_check(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    set_error("device_check: cudaGetDeviceProperties(%d) failed: %s", device, cudaGetErrorString(e));
    return 2;
  }
  DBSR_REQUIRE(prop.major == 10, "device_check: device %d is sm_%d%d; libdbsr_b200 runs on sm_100 (B200) only", device,
               prop.major, prop.minor);
  return 0;
}

extern "C" int dbsr_nchw_to_nhwc(const float* src, const dbsr_nhwc_t* dst, void* stream) {
  DBSR_REQUIRE(src && view_ok(dst), "nchw_to_nhwc: bad arguments");
  dim3 grid(ceil_div((long long)dst->h * dst->w, 32), ceil_div(dst->c, 32), dst->n), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, make_view(dst));
  return check_launch("nchw_to_nhwc");
}

extern "C" int dbsr_nhwc_to_nchw(const dbsr_nhwc_t* src, float* dst, void* stream) {
  DBSR_REQUIRE(dst && view_ok(src), "nhwc_to_nchw: bad arguments");
  dim3 grid(ceil_div((long long)src->h * src->w, 32), ceil_div(src->c, 32), src->n), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(make_view(src), dst);
  return check_launch("nhwc_to_nchw");
}

extern "C" int dbsr_copy_channels(const dbsr_nhwc_t* src, const dbsr_nhwc_t* dst, int32_t group, int32_t src_group,
                                  int32_t src_first, void* stream) {
  DBSR_REQUIRE(view_ok(src) && view_ok(dst), "copy_channels: bad views");
  DBSR_REQUIRE(src->h == dst->h && src->w == dst->w && src->c == dst->c, "copy_channels: geometry mismatch");
  if (group > 0)
    DBSR_REQUIRE(dst->n % group == 0 && (dst->n / group - 1) * src_group + src_first < src->n,
                 "copy_channels: group mapping out of range");
  else
    DBSR_REQUIRE(src->n == dst->n, "copy_channels: image count mismatch");
  const long long total = (long long)dst->n * dst->h * dst->w * dst->c;
  auto al = [](const dbsr_nhwc_t* v) {
    return v->dtype == DBSR_BF16 && v->c_off % 8 == 0 && v->c_pitch % 8 == 0 && ((uintptr_t)v->data % 16) == 0;
  };
  if (al(src) && al(dst) && dst->c % 8 == 0 && (long long)dst->h * dst->w * (dst->c / 8) < (1ll << 31)) {
    const dim3 g((unsigned)ceil_div((long long)dst->h * dst->w * (dst->c / 8), 256), (unsigned)(dst->n < 65535 ? dst->n : 65535));
    launch_pdl(copy_channels_v8_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, make_view(src), make_view(dst), group, src_group,
               src_first);
    return check_launch("copy_channels");
  }
  launch_pdl(copy_channels_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, make_view(src), make_view(dst), group,
                                                                                src_group, src_first);
  return check_launch("copy_channels");
}

extern "C" int dbsr_space_to_depth2(const dbsr_nhwc_t* x, const dbsr_nhwc_t* y, void* stream) {
  DBSR_REQUIRE(view_ok(x) && view_ok(y) && y->n == x->n && y->h == (x->h + 1) / 2 && y->w == (x->w + 1) / 2 &&
                   y->c == 4 * x->c, "space_to_depth2: output must be [n, ceil(h/2), ceil(w/2), 4c]");
  const long long total = (long long)y->n * y->h * y->w * y->c;
  const bool v8 = x->dtype == DBSR_BF16 && y->dtype == DBSR_BF16 && x->c % 8 == 0 && x->c_off % 8 == 0 && x->c_pitch % 8 == 0 &&
                  y->c_off % 8 == 0 && y->c_pitch % 8 == 0 && ((uintptr_t)x->data % 16) == 0 && ((uintptr_t)y->data % 16) == 0;
  if (v8) launch_pdl(space_to_depth2_v8_kernel, dim3(grid_for(total / 8, 256)), dim3(256), 0, (cudaStream_t)stream, make_view(x), make_view(y));
  else launch_pdl(space_to_depth2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, make_view(x), make_view(y));
  return check_launch("space_to_depth2");
}

extern "C" int dbsr_prep_burst(const float* burst, int32_t frames, int32_t H, int32_t W, const dbsr_nhwc_t* enc_in,
                               const dbsr_nhwc_t* pwc_in, void* stream) {
  DBSR_REQUIRE(burst && view_ok(enc_in) && view_ok(pwc_in), "prep_burst: bad arguments");
  DBSR_REQUIRE(enc_in->n == frames && pwc_in->n == frames && enc_in->h == H && enc_in->w == W && enc_in->c >= 4 &&
                   pwc_in->c >= 3, "prep_burst: geometry mismatch");
  const long long total = (long long)frames * H * W + (long long)frames * pwc_in->h * pwc_in->w;
  launch_pdl(prep_burst_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, burst, H, W, make_view(enc_in),
                                                                             make_view(pwc_in));
  return check_launch("prep_burst");
}

// evaluation/burstsr/compute_score.py:110-111: (pred.clamp(0, 1) * 2 ** 14).short()  (float -> int16 truncates toward zero)
__global__ void quantize_q14_kernel(const float* __restrict__ src, short* __restrict__ dst, long long count) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (short)(fminf(fmaxf(__ldg(src + i), 0.0f), 1.0f) * 16384.0f);
}

extern "C" int dbsr_quantize_q14(const float* src, int16_t* dst, int64_t count, void* stream) {
  DBSR_REQUIRE(src && dst && count >= 0, "quantize_q14: bad arguments");
  if (count == 0) return 0;
  launch_pdl(quantize_q14_kernel, dim3(grid_for(count, 256)), dim3(256), 0, (cudaStream_t)stream, src, reinterpret_cast<short*>(dst),
             (long long)count);
  return check_launch("quantize_q14");
}

extern "C" int dbsr_prep_burst_s2d(const float* burst, int32_t frames, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                                   const dbsr_nhwc_t* enc_in, const dbsr_nhwc_t* pwc_s2d, void* stream) {
  DBSR_REQUIRE(burst && view_ok(enc_in) && view_ok(pwc_s2d), "prep_burst_s2d: bad arguments");
  DBSR_REQUIRE(enc_in->n == frames && enc_in->h == H && enc_in->w == W && enc_in->c >= 4, "prep_burst_s2d: enc_in geometry mismatch");
  DBSR_REQUIRE(Hp >= H && Wp >= W && Hp % 2 == 0 && Wp % 2 == 0 && pwc_s2d->n == frames && pwc_s2d->h == Hp / 2 &&
                   pwc_s2d->w == Wp / 2 && pwc_s2d->c == 12 && pwc_s2d->c_off == 0 && pwc_s2d->c_pitch == 16 &&
                   pwc_s2d->dtype == DBSR_BF16 && ((uintptr_t)pwc_s2d->data % 16) == 0,
               "prep_burst_s2d: pwc_s2d must be a dense bf16 [frames, Hp/2, Wp/2, 12 (+4 pad)] buffer");
  const long long per_frame = (long long)H * W + (long long)(Hp / 2) * (Wp / 2);
  DBSR_REQUIRE(per_frame < (1ll << 31), "prep_burst_s2d: frame too large");
  const dim3 g((unsigned)ceil_div(per_frame, 256), (unsigned)(frames < 65535 ? frames : 65535));
  launch_pdl(prep_burst_s2d_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, burst, H, W, Hp, Wp, make_view(enc_in),
             make_view(pwc_s2d));
  return check_launch("prep_burst_s2d");
}

extern "C" int dbsr_deconv4x4s2(const dbsr_nhwc_t* x, const float* w, const float* bias, const dbsr_nhwc_t* y,
                                const dbsr_nhwc_t* y2, void* stream) {
  DBSR_REQUIRE(view_ok(x) && view_ok(y) && w && bias, "deconv4x4s2: bad arguments");
  DBSR_REQUIRE(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w && y->c == 2, "deconv4x4s2: output geometry");
  const bool has2 = y2 && y2->data;
  if (has2) DBSR_REQUIRE(view_ok(y2) && y2->n == y->n && y2->h == y->h && y2->w == y->w && y2->c == 2,
                         "deconv4x4s2: second output geometry");
  const long long warps = (long long)x->n * y->h * y->w;
  const int block = 256;
  const long long blocks = (warps * 32 + block - 1) / block;
  deconv4x4s2_kernel<<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>(make_view(x), w, bias, make_view(y),
                                                                            make_view(has2 ? y2 : nullptr));
  return check_launch("deconv4x4s2");
}

extern "C" int dbsr_flow_from_taps(const dbsr_nhwc_t* ftaps, const float* bias, const dbsr_nhwc_t* y, void* stream) {
  DBSR_REQUIRE(view_ok(ftaps) && bias && view_ok(y) && ftaps->c == 18 && y->c == 2 && y->n == ftaps->n && y->h == ftaps->h &&
                   y->w == ftaps->w, "flow_from_taps: ftaps must be [n, h, w, 18], y [n, h, w, 2]");
  const long long total = (long long)y->n * y->h * y->w;
  launch_pdl(flow_from_taps_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, make_view(ftaps), bias,
             make_view(y));
  return check_launch("flow_from_taps");
}

extern "C" int dbsr_deconv_col2im(const dbsr_nhwc_t* taps, const float* bias_t, const dbsr_nhwc_t* y_t,
                                  const dbsr_nhwc_t* flow, const float* wf, const float* bias_f, const dbsr_nhwc_t* y_f,
                                  const dbsr_nhwc_t* y_f2, void* stream) {
  return dbsr_deconv_col2im_ftaps(taps, bias_t, y_t, flow, nullptr, nullptr, wf, bias_f, y_f, y_f2, stream);
}

extern "C" int dbsr_deconv_col2im_ftaps(const dbsr_nhwc_t* taps, const float* bias_t, const dbsr_nhwc_t* y_t,
                                        const dbsr_nhwc_t* flow, const dbsr_nhwc_t* ftaps, const float* bias6, const float* wf,
                                        const float* bias_f, const dbsr_nhwc_t* y_f, const dbsr_nhwc_t* y_f2, void* stream) {
  DBSR_REQUIRE(view_ok(taps) && bias_t && view_ok(y_t) && taps->c == 32 && y_t->c == 2 && y_t->n == taps->n &&
                   y_t->h == 2 * taps->h && y_t->w == 2 * taps->w, "deconv_col2im: bad tap / output geometry");
  const bool from_taps = ftaps && ftaps->data;
  if (from_taps) {
    DBSR_REQUIRE(!(flow && flow->data), "deconv_col2im: give the flow either as a map or as its 18 tap planes, not both");
    DBSR_REQUIRE(view_ok(ftaps) && bias6 && wf && bias_f && view_ok(y_f) && ftaps->c == 18 && ftaps->n == taps->n &&
                     ftaps->h == taps->h && ftaps->w == taps->w && y_f->c == 2 && y_f->n == y_t->n && y_f->h == y_t->h &&
                     y_f->w == y_t->w, "deconv_col2im: bad flow-tap geometry");
  }
  const bool has_flow = (flow && flow->data) || from_taps;
  if (has_flow && !from_taps)
    DBSR_REQUIRE(view_ok(flow) && wf && bias_f && view_ok(y_f) && flow->c == 2 && flow->n == taps->n && flow->h == taps->h &&
                     flow->w == taps->w && y_f->c == 2 && y_f->n == y_t->n && y_f->h == y_t->h && y_f->w == y_t->w,
                 "deconv_col2im: bad flow geometry");
  const bool has2 = has_flow && y_f2 && y_f2->data;
  if (has2) DBSR_REQUIRE(view_ok(y_f2) && y_f2->c == 2 && y_f2->n == y_t->n && y_f2->h == y_t->h && y_f2->w == y_t->w,
                         "deconv_col2im: second flow output geometry");
  const long long total = (long long)y_t->n * y_t->h * y_t->w;
  launch_pdl(deconv_col2im_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
      make_view(taps), bias_t, make_view(y_t), make_view(has_flow && !from_taps ? flow : nullptr), wf, bias_f,
      make_view(has_flow ? y_f : nullptr), make_view(has2 ? y_f2 : nullptr), make_view(from_taps ? ftaps : nullptr), bias6);
  return check_launch("deconv_col2im");
}

extern "C" int dbsr_flow_head(const dbsr_nhwc_t* flow4, float* offsets, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                              void* stream) {
  DBSR_REQUIRE(view_ok(flow4) && offsets && flow4->c == 2, "flow_head: bad arguments");
  const long long total = (long long)flow4->n * H * W;
  launch_pdl(flow_head_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, make_view(flow4), offsets, H, W,
                                                                            (float)W / (float)Wp, (float)H / (float)Hp);
  return check_launch("flow_head");
}

extern "C" int dbsr_offsets_mod(const float* offsets, const dbsr_nhwc_t* out, int32_t bursts, int32_t frames,
                                float modulo, void* stream) {
  DBSR_REQUIRE(offsets && view_ok(out) && out->n == bursts * frames && out->c >= 2 && frames >= 2,
               "offsets_mod: bad arguments");
  const dim3 g((unsigned)ceil_div((long long)out->h * out->w, 256), (unsigned)(out->n < 65535 ? out->n : 65535));
  launch_pdl(offsets_mod_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, offsets, make_view(out), frames, modulo);
  return check_launch("offsets_mod");
}

extern "C" int dbsr_build_wp_input(const dbsr_nhwc_t* proj, const dbsr_nhwc_t* wp_in, int32_t frames, void* stream) {
  DBSR_REQUIRE(view_ok(proj) && view_ok(wp_in) && proj->n == wp_in->n && proj->h == wp_in->h &&
                   proj->w == wp_in->w && wp_in->c >= 2 * proj->c && proj->n % frames == 0,
               "build_wp_input: bad arguments");
  const long long total = (long long)proj->n * proj->h * proj->w * proj->c;
  build_wp_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(make_view(proj), make_view(wp_in),
                                                                                 frames);
  return check_launch("build_wp_input");
}

extern "C" int dbsr_predictor(const dbsr_nhwc_t* x, const float* w, const float* bias, int32_t cout, float* pred,
                              void* stream) {
  DBSR_REQUIRE(view_ok(x) && w && pred && cout >= 1 && cout <= 4, "predictor: bad arguments");
  const long long total = (long long)x->n * x->h * x->w;
  const size_t smem = (size_t)(cout * x->c + cout) * sizeof(float);
  const size_t es = elem_size(x->dtype);
  const bool vec = x->c % 8 == 0 && (x->c_off * es) % 16 == 0 && (x->c_pitch * es) % 16 == 0 && ((uintptr_t)x->data % 16) == 0;
  const int g = grid_for(total, 256) * 4;   // grid_for caps at 16 blocks/SM worth; one thread per pixel here
  cudaStream_t st = (cudaStream_t)stream;
  if (x->dtype == DBSR_BF16) {
    if (vec) launch_pdl(predictor_kernel<__nv_bfloat16, true>, dim3(g), dim3(256), smem, st, make_view(x), w, bias, cout, pred);
    else launch_pdl(predictor_kernel<__nv_bfloat16, false>, dim3(g), dim3(256), smem, st, make_view(x), w, bias, cout, pred);
  } else {
    if (vec) launch_pdl(predictor_kernel<float, true>, dim3(g), dim3(256), smem, st, make_view(x), w, bias, cout, pred);
    else launch_pdl(predictor_kernel<float, false>, dim3(g), dim3(256), smem, st, make_view(x), w, bias, cout, pred);
  }
  return check_launch("predictor");
}
