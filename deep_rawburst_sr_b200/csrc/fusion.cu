// Flow-based backward warping of the frame embeddings and the fused warp + softmax + weighted sum.
//
//   dbsr_warp          : models/layers/warp.py:19-46 (grid_sample bilinear, zeros padding), NHWC, 16-byte
//                        vectorised along channels so the 4 taps are coalesced 128-channel rows.
//   dbsr_softmax_wsum  : models/dbsr/merging.py:117-124 (softmax over the burst + weighted sum) with the warp
//                        of encoders.py:80 recomputed on the fly, so neither `oth_feat` nor the normalised
//                        weights are materialised.  Per (burst, pixel, 4-channel group) one thread streams the N
//                        logits and the N (gathered) embeddings once, with an online softmax in fp32.
#include "common.cuh"
#include "tma.cuh"

#include <stdlib.h>
#include <string.h>

namespace dbsr {

struct Vec4 { float v[4]; };

template <typename T> __device__ __forceinline__ Vec4 ld4(const T* p);
template <> __device__ __forceinline__ Vec4 ld4<float>(const float* p) {
  const float4 q = __ldg(reinterpret_cast<const float4*>(p));
  return Vec4{{q.x, q.y, q.z, q.w}};
}
template <> __device__ __forceinline__ Vec4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
  return Vec4{{a.x, a.y, b.x, b.y}};
}
template <typename T> __device__ __forceinline__ void st4(T* p, const Vec4& v);
template <> __device__ __forceinline__ void st4<float>(float* p, const Vec4& v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const Vec4& v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.v[0], v.v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.v[2], v.v[3]);
  uint2 q;
  q.x = *reinterpret_cast<uint32_t*>(&a);
  q.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = q;
}

// bilinear gather of 4 channels at (u, v) from image `img` of a NHWC view; taps outside contribute zero
template <typename T>
__device__ __forceinline__ Vec4 gather4(const T* base, int c_pitch, int H, int W, long long img_pix0, float u, float v,
                                        int ch) {
  const float fu = floorf(u), fv = floorf(v);
  const float ax = u - fu, ay = v - fv;
  const int x0 = (int)fu, y0 = (int)fv;
  Vec4 r{{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
    const float wt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
    if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
      const Vec4 t = ld4<T>(base + (img_pix0 + (long long)yy * W + xx) * c_pitch + ch);
#pragma unroll
      for (int i = 0; i < 4; ++i) r.v[i] = fmaf(t.v[i], wt, r.v[i]);
    }
  }
  return r;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) warp_kernel(View feat, const float* __restrict__ offsets, View out, int frames) {
  const int H = out.h, W = out.w, C4 = out.c >> 2;
  const long long HW = (long long)H * W;
  const long long total = (long long)out.n * HW * C4;
  const TI* fbase = reinterpret_cast<const TI*>(feat.data) + feat.c_off;
  TO* obase = reinterpret_cast<TO*>(out.data) + out.c_off;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long pix = i / C4;
    const int f = (int)(pix / HW);
    const int rem = (int)(pix - (long long)f * HW);
    const int y = rem / W, x = rem - y * W;
    long long p = f;        // index into offsets
    bool identity = false;  // burst mode: frame 0 of every burst is the reference (copied)
    if (frames > 0) {
      const int b = f / frames, n = f - b * frames;
      identity = (n == 0);
      p = (long long)b * (frames - 1) + (n - 1);
    }
    Vec4 r;
    if (identity) {
      r = ld4<TI>(fbase + pix * feat.c_pitch + c4 * 4);
    } else {
      const float fx = __ldg(offsets + (p * 2 + 0) * HW + rem);
      const float fy = __ldg(offsets + (p * 2 + 1) * HW + rem);
      r = gather4<TI>(fbase, feat.c_pitch, H, W, (long long)f * HW, (float)x + fx, (float)y + fy, c4 * 4);
    }
    st4<TO>(obase + pix * out.c_pitch + c4 * 4, r);
  }
}

template <typename TF, typename TL, typename TO>
__global__ void __launch_bounds__(256)
softmax_wsum_kernel(View feat, View logits, const float* __restrict__ offsets, View fused, int frames) {
  const int H = fused.h, W = fused.w, C4 = fused.c >> 2;
  const long long HW = (long long)H * W;
  const long long total = (long long)fused.n * HW * C4;
  const TF* fbase = reinterpret_cast<const TF*>(feat.data) + feat.c_off;
  const TL* lbase = reinterpret_cast<const TL*>(logits.data) + logits.c_off;
  TO* obase = reinterpret_cast<TO*>(fused.data) + fused.c_off;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long pix = i / C4;
    const int b = (int)(pix / HW);
    const int rem = (int)(pix - (long long)b * HW);
    const int y = rem / W, x = rem - y * W;
    float m[4], s[4], acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = -INFINITY; s[k] = 0.0f; acc[k] = 0.0f; }
#pragma unroll 2
    for (int n = 0; n < frames; ++n) {
      const long long img = (long long)b * frames + n;
      const Vec4 l = ld4<TL>(lbase + (img * HW + rem) * logits.c_pitch + c4 * 4);
      Vec4 a;
      if (n == 0 || offsets == nullptr) {
        a = ld4<TF>(fbase + (img * HW + rem) * feat.c_pitch + c4 * 4);
      } else {
        const long long p = (long long)b * (frames - 1) + (n - 1);
        const float fx = __ldg(offsets + (p * 2 + 0) * HW + rem);
        const float fy = __ldg(offsets + (p * 2 + 1) * HW + rem);
        a = gather4<TF>(fbase, feat.c_pitch, H, W, img * HW, (float)x + fx, (float)y + fy, c4 * 4);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float mn = fmaxf(m[k], l.v[k]);
        const float sc = expf(m[k] - mn);   // exp(-inf) = 0 on the first frame
        const float e = expf(l.v[k] - mn);
        s[k] = fmaf(s[k], sc, e);
        acc[k] = fmaf(acc[k], sc, a.v[k] * e);
        m[k] = mn;
      }
    }
    Vec4 r;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.v[k] = acc[k] / s[k];
    st4<TO>(obase + pix * fused.c_pitch + c4 * 4, r);
  }
}

// optional materialisation of the reference's `fusion_weights` [B, N, C, H, W] fp32 (never read by any caller
// of the reference; produced on request only)
__global__ void fusion_weights_kernel(View logits, float* __restrict__ wout, int frames) {
  const int H = logits.h, W = logits.w, C = logits.c;
  const long long HW = (long long)H * W;
  const int bursts = logits.n / frames;
  const long long total = (long long)bursts * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int rem = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    float m = -INFINITY;
    for (int n = 0; n < frames; ++n) m = fmaxf(m, view_ld(logits, ((long long)b * frames + n) * HW + rem, c));
    float s = 0.0f;
    for (int n = 0; n < frames; ++n) s += expf(view_ld(logits, ((long long)b * frames + n) * HW + rem, c) - m);
    for (int n = 0; n < frames; ++n) {
      const float e = expf(view_ld(logits, ((long long)b * frames + n) * HW + rem, c) - m);
      wout[(((long long)b * frames + n) * C + c) * HW + rem] = e / s;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// 8-channel (16-byte for bf16, 2 x 16-byte for fp32) vector helpers for the bandwidth-bound kernels
// ---------------------------------------------------------------------------------------------------------
struct Vec8 { float v[8]; };

template <typename T> __device__ __forceinline__ Vec8 ld8(const T* p);
template <> __device__ __forceinline__ Vec8 ld8<float>(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  return Vec8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
template <> __device__ __forceinline__ Vec8 ld8<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
  Vec8 r;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    r.v[2 * k] = f.x; r.v[2 * k + 1] = f.y;
  }
  return r;
}
template <typename T> __device__ __forceinline__ void st8(T* p, const Vec8& v);
template <> __device__ __forceinline__ void st8<float>(float* p, const Vec8& v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const Vec8& v) {
  uint4 q;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v.v[0], v.v[1]), h1 = __floats2bfloat162_rn(v.v[2], v.v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v.v[4], v.v[5]), h3 = __floats2bfloat162_rn(v.v[6], v.v[7]);
  q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
  q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p) = q;
}

struct Taps {  // bilinear taps of one sample position (zeros outside the image)
  int off[4];   // pixel offset inside the image, -1 when out of bounds
  float w[4];
};
__device__ __forceinline__ Taps make_taps(float u, float v, int H, int W) {
  const float fu = floorf(u), fv = floorf(v);
  const float ax = u - fu, ay = v - fv;
  const int x0 = (int)fu, y0 = (int)fv;
  Taps t;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
    t.w[k] = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
    t.off[k] = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? yy * W + xx : -1;
  }
  return t;
}
template <typename T>
__device__ __forceinline__ Vec8 gather8(const T* img_base, int c_pitch, const Taps& t, int ch) {
  Vec8 r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.off[k] >= 0) {
      const Vec8 a = ld8<T>(img_base + (long long)t.off[k] * c_pitch + ch);
#pragma unroll
      for (int i = 0; i < 8; ++i) r.v[i] = fmaf(a.v[i], t.w[k], r.v[i]);
    }
  }
  return r;
}

// ---------------------------------------------------------------------------------------------------------
// warp_proj: the 1x1 projection (merging.py:75) commutes with the bilinear warp (both linear, the warp acts per
// channel), so the engine projects the UNWARPED embeddings first (512 -> 64 channels, tensor cores, no bias / act)
// and this kernel warps the 64-channel result instead of the 512-channel one:
//   p_n = relu(warp(q_n, flow_n) + bias)      (n = 0: no warp)       merging.py:72-75 + encoders.py:80
//   wp_in[:, 0:C] = p_0 ; wp_in[:, C:2C] = p_n - p_0                  merging.py:79-89
// ---------------------------------------------------------------------------------------------------------
// Grid: x covers (pixel, 8-channel group) of one frame in 32-bit arithmetic, y walks the frames (the first version
// decoded a 64-bit linear index with three 64-bit divisions per thread and was instruction bound at 2.1 TB/s).
// SPLIT: the weight predictor's first conv is split into a per-frame and a per-burst part (engine.merge): this kernel then
// writes wp_in[:, 0:C] = p_n for every frame and the base frame's p_0 once per burst into `p0out` [bursts, H, W, C].
template <typename TQ, typename TO, bool SPLIT>
__global__ void __launch_bounds__(256)
warp_proj_kernel(View q, const float* __restrict__ bias, const float* __restrict__ offsets, View wp_in, int frames, View p0out) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int H = q.h, W = q.w, C = q.c, C8 = C >> 3;
  const int HW = H * W;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)(HW * C8)) return;
  const int rem = (int)(idx / (unsigned)C8), c8 = (int)(idx - (unsigned)rem * (unsigned)C8);
  const int ch = c8 * 8;
  const TQ* qbase = reinterpret_cast<const TQ*>(q.data) + q.c_off;
  TO* obase = reinterpret_cast<TO*>(wp_in.data) + wp_in.c_off;
  const Vec8 bv = ld8<float>(bias + ch);
  const int y = rem / W, x = rem - y * W;
#pragma unroll 2
  for (int f = blockIdx.y; f < q.n; f += gridDim.y) {
    const int b = f / frames, n = f - b * frames;
    const long long pix = (long long)f * HW + rem;
    if (SPLIT) {
      Vec8 pn;
      if (n > 0 && offsets != nullptr) {
        const long long pr = (long long)b * (frames - 1) + (n - 1);
        const float fx = __ldg(offsets + (pr * 2 + 0) * HW + rem);
        const float fy = __ldg(offsets + (pr * 2 + 1) * HW + rem);
        const Taps t = make_taps((float)x + fx, (float)y + fy, H, W);
        pn = gather8<TQ>(qbase + (long long)f * HW * q.c_pitch, q.c_pitch, t, ch);
      } else {
        pn = ld8<TQ>(qbase + pix * q.c_pitch + ch);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) pn.v[k] = fmaxf(pn.v[k] + bv.v[k], 0.0f);
      st8<TO>(obase + pix * wp_in.c_pitch + ch, pn);
      if (n == 0) st8<TO>(reinterpret_cast<TO*>(p0out.data) + p0out.c_off + ((long long)b * HW + rem) * p0out.c_pitch + ch, pn);
      continue;
    }
    Vec8 p0 = ld8<TQ>(qbase + ((long long)b * frames * HW + rem) * q.c_pitch + ch);
#pragma unroll
    for (int k = 0; k < 8; ++k) p0.v[k] = fmaxf(p0.v[k] + bv.v[k], 0.0f);
    Vec8 d;
    if (n == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) d.v[k] = 0.0f;
    } else {
      Vec8 pn;
      if (offsets != nullptr) {
        const long long pr = (long long)b * (frames - 1) + (n - 1);
        const float fx = __ldg(offsets + (pr * 2 + 0) * HW + rem);
        const float fy = __ldg(offsets + (pr * 2 + 1) * HW + rem);
        const Taps t = make_taps((float)x + fx, (float)y + fy, H, W);
        pn = gather8<TQ>(qbase + (long long)f * HW * q.c_pitch, q.c_pitch, t, ch);
      } else {
        pn = ld8<TQ>(qbase + pix * q.c_pitch + ch);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) d.v[k] = fmaxf(pn.v[k] + bv.v[k], 0.0f) - p0.v[k];
    }
    st8<TO>(obase + pix * wp_in.c_pitch + ch, p0);
    st8<TO>(obase + pix * wp_in.c_pitch + C + ch, d);
  }
}

// ---------------------------------------------------------------------------------------------------------
// softmax over the burst + weighted sum, 8 channels per thread, warp of the embeddings recomputed on the fly
// ---------------------------------------------------------------------------------------------------------
// Work mapping: one block = an 8x4 pixel tile x a 64-channel slice of one burst (256 threads: 8 channel groups x 8 x 4
// pixels).  Neighbouring pixels share most of their bilinear taps, so keeping a compact 2-D footprint per block lets L1
// serve the 4-tap gather (L2->SM traffic ~1.4x the tensor instead of 4x); a warp reads 4 pixels x 128 contiguous bytes.
constexpr int WS_TW = 8, WS_TH = 4;
template <typename TF, typename TL, typename TO>
__global__ void __launch_bounds__(256)
softmax_wsum8_kernel(View feat, View logits, const float* __restrict__ offsets, View fused, int frames) {
  const int H = fused.h, W = fused.w;
  const int HW = H * W;
  const TF* fbase = reinterpret_cast<const TF*>(feat.data) + feat.c_off;
  const TL* lbase = reinterpret_cast<const TL*>(logits.data) + logits.c_off;
  TO* obase = reinterpret_cast<TO*>(fused.data) + fused.c_off;
  const int tiles_x = (W + WS_TW - 1) / WS_TW;
  const int g = threadIdx.x & 7, px = (threadIdx.x >> 3) & 7, py = threadIdx.x >> 6;
  const int b = blockIdx.z;
  const int ch_raw = blockIdx.y * 64 + g * 8;
  const int ch = min(ch_raw, fused.c - 8);   // clamped for the loads; the store is predicated on `live`
  __shared__ float2 offs_s[32][16];
  {
    const int y_raw = (blockIdx.x / tiles_x) * WS_TH + py, x_raw = (blockIdx.x % tiles_x) * WS_TW + px;
    const bool live = y_raw < H && x_raw < W && ch_raw < fused.c;   // whole warps stay alive for the __syncwarp below
    const int y = min(y_raw, H - 1), x = min(x_raw, W - 1);
    const int rem = y * W + x;
    const long long pix = (long long)b * HW + rem;
    // the flows of all frames at this pixel first: independent loads, shared by the 8 threads (channel groups) of the
    // pixel through smem -- one memory latency for the whole burst instead of one dependent load per frame
    const bool hoisted = offsets != nullptr && frames <= 17;
    if (hoisted) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n = g + 8 * q;
        if (n + 1 < frames) {
          const long long pr = (long long)b * (frames - 1) + n;
          offs_s[threadIdx.x >> 3][n] = make_float2(__ldg(offsets + (pr * 2 + 0) * HW + rem), __ldg(offsets + (pr * 2 + 1) * HW + rem));
        }
      }
      __syncwarp();
    }
    // online softmax with ONE exponential per element: d = l - m; x = exp(-|d|);
    //   d <= 0: (scale old, weight new) = (1, x)   else: (x, 1) and the running max moves to l
    float m[8], s[8], acc[8];
    {
      const long long img = (long long)b * frames;
      const Vec8 l = ld8<TL>(lbase + (img * HW + rem) * logits.c_pitch + ch);
      const Vec8 a = ld8<TF>(fbase + (img * HW + rem) * feat.c_pitch + ch);
#pragma unroll
      for (int k = 0; k < 8; ++k) { m[k] = l.v[k]; s[k] = 1.0f; acc[k] = a.v[k]; }
    }
#pragma unroll 2
    for (int n = 1; n < frames; ++n) {
      const long long img = (long long)b * frames + n;
      const Vec8 l = ld8<TL>(lbase + (img * HW + rem) * logits.c_pitch + ch);
      Vec8 a;
      if (offsets == nullptr) {
        a = ld8<TF>(fbase + (img * HW + rem) * feat.c_pitch + ch);
      } else {
        float fx, fy;
        if (hoisted) {
          const float2 o2 = offs_s[threadIdx.x >> 3][n - 1];
          fx = o2.x; fy = o2.y;
        } else {
          const long long pr = (long long)b * (frames - 1) + (n - 1);
          fx = __ldg(offsets + (pr * 2 + 0) * HW + rem);
          fy = __ldg(offsets + (pr * 2 + 1) * HW + rem);
        }
        const Taps t = make_taps((float)x + fx, (float)y + fy, H, W);
        a = gather8<TF>(fbase + img * HW * feat.c_pitch, feat.c_pitch, t, ch);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = l.v[k] - m[k];
        const float ex = (sizeof(TL) == 2) ? __expf(-fabsf(d)) : expf(-fabsf(d));
        const bool up = d > 0.0f;
        const float sc = up ? ex : 1.0f;
        const float e = up ? 1.0f : ex;
        s[k] = fmaf(s[k], sc, e);
        acc[k] = fmaf(acc[k], sc, a.v[k] * e);
        m[k] = up ? l.v[k] : m[k];
      }
    }
    Vec8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = acc[k] / s[k];
    if (live) st8<TO>(obase + pix * fused.c_pitch + ch, r);
  }
}


// ---------------------------------------------------------------------------------------------------------
// helpers of the asynchronously prefetched bf16 kernel below
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ Vec8 unpack_bf16x8(const uint4& q) {
  Vec8 r;
  r.v[0] = __uint_as_float(q.x << 16); r.v[1] = __uint_as_float(q.x & 0xFFFF0000u);
  r.v[2] = __uint_as_float(q.y << 16); r.v[3] = __uint_as_float(q.y & 0xFFFF0000u);
  r.v[4] = __uint_as_float(q.z << 16); r.v[5] = __uint_as_float(q.z & 0xFFFF0000u);
  r.v[6] = __uint_as_float(q.w << 16); r.v[7] = __uint_as_float(q.w & 0xFFFF0000u);
  return r;
}

// Packed fp32 pairs (Blackwell FFMA2: fma / mul / add .f32x2 operate on two IEEE fp32 values per instruction -- the same
// results as two scalar instructions, half the issue slots).  lo = element 2j, hi = element 2j + 1 of a Vec8.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
struct Vec8p { f32x2_t p[4]; };
// 8 bf16 -> 4 packed fp32 pairs (a bf16 is the high half of its fp32)
__device__ __forceinline__ Vec8p unpack_bf16x8_p(const uint4& q) {
  Vec8p r;
  r.p[0] = f2_pack(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xFFFF0000u));
  r.p[1] = f2_pack(__uint_as_float(q.y << 16), __uint_as_float(q.y & 0xFFFF0000u));
  r.p[2] = f2_pack(__uint_as_float(q.z << 16), __uint_as_float(q.z & 0xFFFF0000u));
  r.p[3] = f2_pack(__uint_as_float(q.w << 16), __uint_as_float(q.w & 0xFFFF0000u));
  return r;
}

// ---------------------------------------------------------------------------------------------------------
// Asynchronously prefetched, pair-pipelined kernel (bf16 embeddings and logits, warp on the fly).
// History (profiles/): the register-resident kernel above is latency bound (6 warps per scheduler, each waiting on 5
// loads per frame); a first cp.async version (every thread keeps 3 frames x 5 copies in flight in its own slots of a
// shared-memory ring, no block barrier in the loop) reached 3.0 TB/s and then turned out INSTRUCTION bound (ncu, B=32,
// 48^2: issue slots 88 % busy, IPC 3.5, DRAM 35 %): ~370 warp instructions per thread-frame, ~150 of them address
// arithmetic (64-bit image offsets, four tap offsets, bounds tests, ring-slot modulo) repeated by each of the 8
// channel-group threads of a pixel for every frame.  Here that work is done ONCE per (pixel, frame) before the loop:
// a 32-byte record {4 byte offsets of the clamped taps, 4 bilinear weights with out-of-image taps zeroed (= the zero
// padding of grid_sample)} in shared memory, so the per-frame issue path is 1 LDS.128 + pointer bumps + 5 cp.async
// and the consume path needs no masks.  Frames are processed in PAIRS (ring = 2 stages x 2 frames, a thread only
// reads slots it filled itself: cp.async.wait_group is the only synchronisation): the online softmax rescales the
// running sums once per pair (3 exponentials per 2 frames, no selects).  181 instructions per thread-frame,
// 4.1-4.2 TB/s (63 % of the measured HBM peak).
// ---------------------------------------------------------------------------------------------------------
struct __align__(16) WsRec { uint32_t o[4]; float w[4]; };
constexpr int WSP_MAX_OTHERS = 16;
constexpr int WSP_SMEM = 2 * 2 * 5 * 256 * 16;           // ring: [2 stages][2 frames][5 loads][256 threads] x 16 bytes
// TMA modes.  1 (default): the reference-frame tiles (embedding + logits of frame 0) are staged by TMA, ring as above.
// 2 (A/B, DBSR_WSUM_TMA=2): additionally all logit tiles: ring [2 stages][2 frames][4 taps][256 threads] x 16 B of gathered taps
// (cp.async) | per warp [2 stages][2 frames] 512 B logit boxes.  Behind the ring in both modes: per warp 2 x 512 B
// reference-frame boxes | per warp 3 mbarriers.
constexpr int WST_TAPS = 2 * 2 * 4 * 256 * 16;
constexpr int WST_LOGITS = 8 * 2 * 2 * 512;
constexpr int WST_REF = 8 * 2 * 512;
constexpr int WST_SMEM2 = WST_TAPS + WST_LOGITS + WST_REF + 8 * 3 * 8 + 128 /* alignment slack */;
constexpr int WST_SMEM1 = WSP_SMEM + WST_REF + 8 * 3 * 8 + 128;
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MODE >= 1 (north_star: "TMA staging of the reference-frame tiles"): the unwarped reference-frame tiles (embedding and logits
// of frame 0) are staged by the TMA engine: cp.async.bulk.tensor boxes {64 channels, 4 pixels, 1 row} = the 512 bytes one
// warp consumes, issued by one elected lane per warp before the gather records are computed, completion on a per-warp
// mbarrier (no block-wide synchronisation); out-of-image pixels / channels of ragged tiles are zero-filled by the tensor map.
// MODE == 2 stages ALL 14 logit tiles that way (15 of the 28 tile streams of a block; only the 13 x 4 flow-dependent bilinear
// taps stay per-thread cp.async gathers).  Measured on B200 (B = 32, 48^2, profiles/r02_wsum_tma_ncu_summary.txt): the kernel
// is bound by instruction issue, not by the copy mechanism, and the per-pair mbarrier wait / re-arm / box issue costs more
// warp instructions (467 M vs 418 M) than the two cp.async it removes: 554 us against 527 us.  So MODE 1 is the default.
template <typename TO, int MODE>
__global__ void __launch_bounds__(256, 2)
softmax_wsum8_pair_kernel(const __grid_constant__ CUtensorMap tmap_feat, const __grid_constant__ CUtensorMap tmap_logits,
                          View feat, View logits, const float* __restrict__ offsets, View fused, int frames) {
  extern __shared__ __align__(128) uint4 ring[];
  __shared__ WsRec rec_s[32][WSP_MAX_OTHERS];
  const int H = fused.h, W = fused.w;
  const int HW = H * W;
  const int tiles_x = (W + WS_TW - 1) / WS_TW;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = tid & 7, pix = tid >> 3, px = pix & 7, py = pix >> 3;
  const int b = blockIdx.z;
  const int ty0 = (blockIdx.x / tiles_x) * WS_TH, tx0 = (blockIdx.x % tiles_x) * WS_TW;
  constexpr bool TMA = MODE == 2;                         // logits through TMA boxes
  constexpr bool TMA_REF = MODE >= 1;                     // reference-frame tiles through TMA boxes
  constexpr int RING_BYTES = TMA ? WST_TAPS : WSP_SMEM;
  constexpr int LOGIT_BYTES = TMA ? WST_LOGITS : 0;
  constexpr int TAPS = TMA ? 4 : 5;                       // 16-byte ring slots per thread and frame
  constexpr uint32_t FRAME_BYTES = TAPS * 4096u, STAGE_BYTES = 2u * FRAME_BYTES;
  // per-warp TMA staging (TMA variant): 128-byte aligned boxes behind the tap ring
  const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t lg_s = ((ring_base + (uint32_t)RING_BYTES + 127u) & ~127u) + (uint32_t)warp * 2048u;      // [stage][frame] 512 B
  const uint32_t ref_s = ((ring_base + (uint32_t)RING_BYTES + 127u) & ~127u) + (uint32_t)LOGIT_BYTES + (uint32_t)warp * 1024u;
  const uint32_t bar_s = ((ring_base + (uint32_t)RING_BYTES + 127u) & ~127u) + (uint32_t)(LOGIT_BYTES + WST_REF) + (uint32_t)warp * 24u;
  // this warp's 4 pixels: tile row warp / 2, columns (warp & 1) * 4 .. + 3
  const int wy = ty0 + (warp >> 1), wx = tx0 + (warp & 1) * 4;
  if (TMA_REF) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s + 8u));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s + 16u));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  }
  griddep_launch_dependents_if_small();
  griddep_wait();
  const __nv_bfloat16* fbase = reinterpret_cast<const __nv_bfloat16*>(feat.data) + feat.c_off;
  const __nv_bfloat16* lbase = reinterpret_cast<const __nv_bfloat16*>(logits.data) + logits.c_off;
  TO* obase = reinterpret_cast<TO*>(fused.data) + fused.c_off;
  const int ch_raw = blockIdx.y * 64 + g * 8;
  const int ch = min(ch_raw, fused.c - 8);   // clamped for the loads; the store is predicated on `live`
  const int y_raw = ty0 + py, x_raw = tx0 + px;
  const bool live = y_raw < H && x_raw < W && ch_raw < fused.c;
  const int y = min(y_raw, H - 1), x = min(x_raw, W - 1);
  const int rem = y * W + x;
  const int others = frames - 1;
  auto tma_box = [&](const CUtensorMap* map, uint32_t bar, uint32_t dst, int img) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"((int)blockIdx.y * 64), "r"(wx), "r"(wy), "r"(img)
        : "memory");
  };
  if (TMA_REF && lane == 0) {     // reference frame of this burst: embedding tile + logit tile -> ref_s, completion on bar 2
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s + 16u), "r"(1024u) : "memory");
    tma_box(&tmap_feat, bar_s + 16u, ref_s, b * frames);
    tma_box(&tmap_logits, bar_s + 16u, ref_s + 512u, b * frames);
  }
  // ---- per (pixel, frame) gather records, computed by the 8 channel-group threads of the pixel (2 frames each)
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int k = g + 8 * q;
    if (k < others) {
      const long long pr = (long long)b * others + k;
      const float u = (float)x + __ldg(offsets + (pr * 2 + 0) * HW + rem);
      const float v = (float)y + __ldg(offsets + (pr * 2 + 1) * HW + rem);
      const float fu = floorf(u), fv = floorf(v);
      const float ax = u - fu, ay = v - fv;
      const int x0 = (int)fminf(fmaxf(fu, -2.0f), (float)W), y0 = (int)fminf(fmaxf(fv, -2.0f), (float)H);
      const bool okx0 = x0 >= 0 && x0 < W, okx1 = x0 + 1 >= 0 && x0 + 1 < W;
      const bool oky0 = y0 >= 0 && y0 < H, oky1 = y0 + 1 >= 0 && y0 + 1 < H;
      const int x0c = min(max(x0, 0), W - 1), x1c = min(max(x0 + 1, 0), W - 1);
      const int y0c = min(max(y0, 0), H - 1), y1c = min(max(y0 + 1, 0), H - 1);
      const uint32_t pb = (uint32_t)feat.c_pitch * 2u;
      WsRec r;
      r.o[0] = (uint32_t)(y0c * W + x0c) * pb; r.o[1] = (uint32_t)(y0c * W + x1c) * pb;
      r.o[2] = (uint32_t)(y1c * W + x0c) * pb; r.o[3] = (uint32_t)(y1c * W + x1c) * pb;
      r.w[0] = (okx0 && oky0) ? (1.0f - ax) * (1.0f - ay) : 0.0f; r.w[1] = (okx1 && oky0) ? ax * (1.0f - ay) : 0.0f;
      r.w[2] = (okx0 && oky1) ? (1.0f - ax) * ay : 0.0f;          r.w[3] = (okx1 && oky1) ? ax * ay : 0.0f;
      rec_s[pix][k] = r;
    }
  }
  __syncwarp();     // a pixel's 8 threads are in one warp
  const uint32_t ring_s = ring_base + (uint32_t)tid * 16u;
  const size_t lstride = (size_t)HW * logits.c_pitch * 2, fstride = (size_t)HW * feat.c_pitch * 2;
  const char* l0 = reinterpret_cast<const char*>(lbase + ((long long)b * frames * HW + rem) * logits.c_pitch + ch);
  const char* f0 = reinterpret_cast<const char*>(fbase + (long long)b * frames * HW * feat.c_pitch + ch);
  const char* lq = l0 + lstride;        // next frame to issue (frame 1)
  const char* fq = f0 + fstride;
  int nq = 1;
  auto issue_pair = [&](int stage) {   // the next two frames -> ring stage (skips frames past the burst)
    const uint32_t dst = ring_s + (uint32_t)stage * STAGE_BYTES;
    if (TMA) {
      // the warp has consumed this stage's logit boxes (values are in registers; __syncwarp by the caller): re-arm + refill
      if (lane == 0 && nq < frames) {
        const int nf = min(2, frames - nq);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s + (uint32_t)stage * 8u), "r"((uint32_t)nf * 512u) : "memory");
        tma_box(&tmap_logits, bar_s + (uint32_t)stage * 8u, lg_s + (uint32_t)stage * 1024u, b * frames + nq);
        if (nf == 2) tma_box(&tmap_logits, bar_s + (uint32_t)stage * 8u, lg_s + (uint32_t)stage * 1024u + 512u, b * frames + nq + 1);
      }
    }
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      if (nq < frames) {
        const uint32_t d = dst + (uint32_t)f * FRAME_BYTES;
        if (!TMA) cp_async_16(d, lq, 16u);
        const uint32_t t0 = TMA ? 0u : 4096u;
        const uint4 o = *reinterpret_cast<const uint4*>(rec_s[pix][nq - 1].o);
        cp_async_16(d + t0, fq + o.x, 16u);
        cp_async_16(d + t0 + 4096u, fq + o.y, 16u);
        cp_async_16(d + t0 + 8192u, fq + o.z, 16u);
        cp_async_16(d + t0 + 12288u, fq + o.w, 16u);
        lq += lstride; fq += fstride; ++nq;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto bar_wait = [&](uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
  };
  issue_pair(0);
  issue_pair(1);
  // reference frame (never warped).  The running maximum m is kept in log2 units.
  constexpr float LOG2E = 1.4426950408889634f;
  // the online-softmax state of the thread's 8 channels as 4 packed fp32 pairs (FFMA2: half the fp32 issue slots)
  const f32x2_t LOG2E2 = f2_pack(LOG2E, LOG2E), NEG1 = f2_pack(-1.0f, -1.0f);
  f32x2_t m[4], s[4], acc[4];
  {
    Vec8p l, a;
    if (TMA_REF) {
      bar_wait(bar_s + 16u, 0u);
      uint4 qa, ql;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qa.x), "=r"(qa.y), "=r"(qa.z), "=r"(qa.w) : "r"(ref_s + (uint32_t)lane * 16u));
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ql.x), "=r"(ql.y), "=r"(ql.z), "=r"(ql.w) : "r"(ref_s + 512u + (uint32_t)lane * 16u));
      a = unpack_bf16x8_p(qa); l = unpack_bf16x8_p(ql);
    } else {
      l = unpack_bf16x8_p(__ldg(reinterpret_cast<const uint4*>(l0)));
      a = unpack_bf16x8_p(__ldg(reinterpret_cast<const uint4*>(f0 + (size_t)rem * feat.c_pitch * 2)));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = f2_mul(l.p[k], LOG2E2); s[k] = f2_pack(1.0f, 1.0f); acc[k] = a.p[k]; }
  }
  const int npairs = (others + 1) >> 1;
  for (int pp = 0; pp < npairs; ++pp) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const int stg = pp & 1;
    const uint4* st = ring + stg * (2 * TAPS * 256) + tid;
    const int n0 = 2 * pp;                     // index of the pair's first frame among the others
    const bool has2 = n0 + 1 < others;
    Vec8p lA, lB, aA, aB;
    auto interp = [&](const uint4* sf, const WsRec& rc, Vec8p& a) {      // sf: the frame's first TAP slot of this thread
      const float4 w = *reinterpret_cast<const float4*>(rc.w);
      const f32x2_t w0 = f2_pack(w.x, w.x), w1 = f2_pack(w.y, w.y), w2 = f2_pack(w.z, w.z), w3 = f2_pack(w.w, w.w);
      const Vec8p t0 = unpack_bf16x8_p(sf[0]), t1 = unpack_bf16x8_p(sf[256]);
#pragma unroll
      for (int k = 0; k < 4; ++k) a.p[k] = f2_fma(t1.p[k], w1, f2_mul(t0.p[k], w0));
      const Vec8p t2 = unpack_bf16x8_p(sf[512]), t3 = unpack_bf16x8_p(sf[768]);
#pragma unroll
      for (int k = 0; k < 4; ++k) a.p[k] = f2_fma(t3.p[k], w3, f2_fma(t2.p[k], w2, a.p[k]));
    };
    constexpr int T0 = TMA ? 0 : 256;          // first tap slot behind the logit slot of the cp.async-only ring
    if (TMA) {
      bar_wait(bar_s + (uint32_t)stg * 8u, (uint32_t)(pp >> 1) & 1u);
      uint4 q;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(lg_s + (uint32_t)stg * 1024u + (uint32_t)lane * 16u));
      lA = unpack_bf16x8_p(q);
    } else {
      lA = unpack_bf16x8_p(st[0]);
    }
    interp(st + T0, rec_s[pix][n0], aA);
    if (has2) {
      if (TMA) {
        uint4 q;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(lg_s + (uint32_t)stg * 1024u + 512u + (uint32_t)lane * 16u));
        lB = unpack_bf16x8_p(q);
      } else {
        lB = unpack_bf16x8_p(st[TAPS * 256]);
      }
      interp(st + TAPS * 256 + T0, rec_s[pix][n0 + 1], aB);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) { lB.p[k] = f2_pack(-INFINITY, -INFINITY); aB.p[k] = f2_pack(0.0f, 0.0f); }
    }
    // this stage's slots are consumed (values are in registers): refill it with the pair after next
    if (TMA) __syncwarp();
    issue_pair(stg);
    // online softmax, one rescale per pair: 3 exponentials per 2 frames and element
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const f32x2_t la = f2_mul(lA.p[k], LOG2E2), lb = f2_mul(lB.p[k], LOG2E2);
      float m0, m1, a0, a1, b0, b1;
      f2_unpack(m[k], m0, m1); f2_unpack(la, a0, a1); f2_unpack(lb, b0, b1);
      const f32x2_t mn = f2_pack(fmaxf(m0, fmaxf(a0, b0)), fmaxf(m1, fmaxf(a1, b1)));
      // x - mn as fma(mn, -1, x): the product is exact, so this is the scalar subtraction bit for bit
      float d0, d1, e0, e1, g0, g1;
      f2_unpack(f2_fma(mn, NEG1, m[k]), d0, d1); f2_unpack(f2_fma(mn, NEG1, la), e0, e1); f2_unpack(f2_fma(mn, NEG1, lb), g0, g1);
      const f32x2_t sc = f2_pack(ex2f(d0), ex2f(d1)), ea = f2_pack(ex2f(e0), ex2f(e1)), eb = f2_pack(ex2f(g0), ex2f(g1));
      s[k] = f2_fma(s[k], sc, f2_add(ea, eb));
      acc[k] = f2_fma(acc[k], sc, f2_fma(aA.p[k], ea, f2_mul(aB.p[k], eb)));
      m[k] = mn;
    }
  }
  Vec8 r;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float a0, a1, s0, s1;
    f2_unpack(acc[k], a0, a1); f2_unpack(s[k], s0, s1);
    r.v[2 * k] = __fdividef(a0, s0); r.v[2 * k + 1] = __fdividef(a1, s1);
  }
  if (live) st8<TO>(obase + ((long long)b * HW + rem) * fused.c_pitch + ch, r);
}


// upsampling.py:59-65: per-channel 3x3 blur with zero padding, 8 channels per thread.
// Row-sliding variant: a thread owns (x, 8 channels) and walks BLUR_ROWS output rows; every input row is loaded once
// per thread (3 x-taps, the neighbours are the adjacent threads' lines -> L1) and feeds the three output rows it
// touches through rolling partial sums, instead of 9 loads per output (whose vertical neighbours miss L1 and made the
// kernel L2-bound).   out[y] = h0[y-1] + h1[y] + h2[y+1],  hk[r] = sum_dx K[k][dx] * in[r][x+dx-1]
constexpr int BLUR_ROWS = 16;
template <typename T>
__global__ void __launch_bounds__(256) blur3x3_rows_kernel(View x, View y, float k0, float k1, float k2, float k3, float k4,
                                                           float k5, float k6, float k7, float k8) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const float kk[9] = {k0, k1, k2, k3, k4, k5, k6, k7, k8};
  const int H = x.h, W = x.w, C8 = x.c >> 3;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;       // (pixel column, channel group)
  if (col >= W * C8) return;
  const int px = col / C8, c8 = col - px * C8;
  const int y0 = blockIdx.y * BLUR_ROWS, n = blockIdx.z;
  const T* xb = reinterpret_cast<const T*>(x.data) + x.c_off + c8 * 8;
  T* yb = reinterpret_cast<T*>(y.data) + y.c_off + c8 * 8;
  Vec8 a0, a1;          // a0: output row r-1 so far (h0[r-2] + h1[r-1]); a1: output row r so far (h0[r-1])
#pragma unroll
  for (int k = 0; k < 8; ++k) { a0.v[k] = 0.0f; a1.v[k] = 0.0f; }
  const int r_end = min(y0 + BLUR_ROWS, H);
#pragma unroll 3
  for (int r = y0 - 1; r <= r_end; ++r) {
    Vec8 h0, h1, h2;
#pragma unroll
    for (int k = 0; k < 8; ++k) { h0.v[k] = 0.0f; h1.v[k] = 0.0f; h2.v[k] = 0.0f; }
    if (r >= 0 && r < H) {
      const T* row = xb + ((long long)n * H + r) * W * x.c_pitch;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = px + dx - 1;
        if (xx >= 0 && xx < W) {
          const Vec8 a = ld8<T>(row + (long long)xx * x.c_pitch);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            h0.v[k] = fmaf(a.v[k], kk[dx], h0.v[k]);          // row r is the row ABOVE output r+1: kernel row 0
            h1.v[k] = fmaf(a.v[k], kk[3 + dx], h1.v[k]);      // ... the centre row of output r
            h2.v[k] = fmaf(a.v[k], kk[6 + dx], h2.v[k]);      // ... the row BELOW output r-1: kernel row 2
          }
        }
      }
    }
    // output row r-1 is complete once row r has contributed
    if (r - 1 >= y0 && r - 1 < r_end) {
      Vec8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = a0.v[k] + h2.v[k];
      st8<T>(yb + (((long long)n * H + (r - 1)) * W + px) * y.c_pitch, o);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { a0.v[k] = a1.v[k] + h1.v[k]; a1.v[k] = h0.v[k]; }
  }
}

// Separable form (the reference's blur IS separable: gauss_2d = outer product of two gauss_1d, filtering.py:28-40):
// hrow[r] = a0 in[r][x-1] + a1 in[r][x] + a2 in[r][x+1];  out[r-1] = b0 hrow[r-2] + b1 hrow[r-1] + b2 hrow[r]  -- 6 instead of
// 9 + 2 arithmetic instructions per element (ncu: the 9-tap kernel had 63 % of its issue slots busy at 2.6 TB/s), and 32
// rows per thread (6 % halo re-reads instead of 12 %).
constexpr int BLUR_SEP_ROWS = 32;
template <typename T>
__global__ void __launch_bounds__(256) blur3x3_sep_kernel(View x, View y, float a0, float a1, float a2, float b0, float b1, float b2, int rows) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int H = x.h, W = x.w, C8 = x.c >> 3;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;       // (pixel column, channel group)
  if (col >= W * C8) return;
  const int px = col / C8, c8 = col - px * C8;
  const int y0 = blockIdx.y * rows, n = blockIdx.z;
  const T* xb = reinterpret_cast<const T*>(x.data) + x.c_off + c8 * 8;
  T* yb = reinterpret_cast<T*>(y.data) + y.c_off + c8 * 8;
  const bool has_l = px > 0, has_r = px + 1 < W;
  const int r_end = min(y0 + rows, H);
  if constexpr (sizeof(T) == 2) {
    // bf16 maps: the same operations in the same order on packed fp32 pairs (FFMA2: half the arithmetic issue slots -- ncu
    // had this kernel at 54 % issue-slot utilisation and 42 % of the DRAM peak, i.e. instruction bound); results are
    // bit-identical to the scalar form below
    const f32x2_t A0 = f2_pack(a0, a0), A1 = f2_pack(a1, a1), A2 = f2_pack(a2, a2);
    const f32x2_t B0 = f2_pack(b0, b0), B1 = f2_pack(b1, b1), B2 = f2_pack(b2, b2), Z = f2_pack(0.0f, 0.0f);
    f32x2_t hm2[4] = {Z, Z, Z, Z}, hm1[4] = {Z, Z, Z, Z};
#pragma unroll 4
    for (int r = y0 - 1; r <= r_end; ++r) {
      f32x2_t h[4] = {Z, Z, Z, Z};
      if (r >= 0 && r < H) {
        const T* row = xb + (((long long)n * H + r) * W + px) * x.c_pitch;
        const Vec8p c = unpack_bf16x8_p(__ldg(reinterpret_cast<const uint4*>(row)));
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = f2_mul(c.p[k], A1);
        if (has_l) {
          const Vec8p l = unpack_bf16x8_p(__ldg(reinterpret_cast<const uint4*>(row - x.c_pitch)));
#pragma unroll
          for (int k = 0; k < 4; ++k) h[k] = f2_fma(l.p[k], A0, h[k]);
        }
        if (has_r) {
          const Vec8p rr = unpack_bf16x8_p(__ldg(reinterpret_cast<const uint4*>(row + x.c_pitch)));
#pragma unroll
          for (int k = 0; k < 4; ++k) h[k] = f2_fma(rr.p[k], A2, h[k]);
        }
      }
      if (r - 1 >= y0 && r - 1 < r_end) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float lo, hi;
          f2_unpack(f2_fma(h[k], B2, f2_fma(hm1[k], B1, f2_mul(hm2[k], B0))), lo, hi);
          const __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
          o[k] = *reinterpret_cast<const uint32_t*>(&pk);
        }
        *reinterpret_cast<uint4*>(yb + (((long long)n * H + (r - 1)) * W + px) * y.c_pitch) = make_uint4(o[0], o[1], o[2], o[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { hm2[k] = hm1[k]; hm1[k] = h[k]; }
    }
    return;
  }
  Vec8 hm2, hm1;       // horizontal sums of rows r-2 and r-1
#pragma unroll
  for (int k = 0; k < 8; ++k) { hm2.v[k] = 0.0f; hm1.v[k] = 0.0f; }
#pragma unroll 4
  for (int r = y0 - 1; r <= r_end; ++r) {
    Vec8 h;
#pragma unroll
    for (int k = 0; k < 8; ++k) h.v[k] = 0.0f;
    if (r >= 0 && r < H) {
      const T* row = xb + (((long long)n * H + r) * W + px) * x.c_pitch;
      const Vec8 c = ld8<T>(row);
#pragma unroll
      for (int k = 0; k < 8; ++k) h.v[k] = c.v[k] * a1;
      if (has_l) {
        const Vec8 l = ld8<T>(row - x.c_pitch);
#pragma unroll
        for (int k = 0; k < 8; ++k) h.v[k] = fmaf(l.v[k], a0, h.v[k]);
      }
      if (has_r) {
        const Vec8 rr = ld8<T>(row + x.c_pitch);
#pragma unroll
        for (int k = 0; k < 8; ++k) h.v[k] = fmaf(rr.v[k], a2, h.v[k]);
      }
    }
    if (r - 1 >= y0 && r - 1 < r_end) {
      Vec8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = fmaf(h.v[k], b2, fmaf(hm1.v[k], b1, hm2.v[k] * b0));
      st8<T>(yb + (((long long)n * H + (r - 1)) * W + px) * y.c_pitch, o);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { hm2.v[k] = hm1.v[k]; hm1.v[k] = h.v[k]; }
  }
}

static inline int grid_cap(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static bool vec4_ok(const dbsr_nhwc_t* v) {
  const size_t es = elem_size(v->dtype);
  const size_t al = 4 * es;  // 16 B for fp32, 8 B for bf16
  return v->c % 4 == 0 && (v->c_off * es) % al == 0 && (v->c_pitch * es) % al == 0 && ((uintptr_t)v->data % al) == 0;
}

static bool vec8_ok(const dbsr_nhwc_t* v) {
  const size_t es = elem_size(v->dtype);
  return v->c % 8 == 0 && (v->c_off * es) % 16 == 0 && (v->c_pitch * es) % 16 == 0 && ((uintptr_t)v->data % 16) == 0;
}

}  // namespace dbsr

using namespace dbsr;

extern "C" int dbsr_warp(const dbsr_nhwc_t* feat, const float* offsets, const dbsr_nhwc_t* out, int32_t frames,
                         void* stream) {
  DBSR_REQUIRE(view_ok(feat) && view_ok(out) && offsets, "warp: bad arguments");
  DBSR_REQUIRE(out->n == feat->n && out->h == feat->h && out->w == feat->w && out->c == feat->c, "warp: geometry");
  DBSR_REQUIRE(vec4_ok(feat) && vec4_ok(out), "warp: channel count/offset/pitch must be multiples of 4 and aligned");
  DBSR_REQUIRE(frames == 0 || (frames >= 2 && out->n % frames == 0), "warp: image count is not a multiple of frames");
  const long long total = (long long)out->n * out->h * out->w * (out->c / 4);
  const int g = grid_cap(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  View f = make_view(feat), o = make_view(out);
  if (feat->dtype == DBSR_F32 && out->dtype == DBSR_F32) warp_kernel<float, float><<<g, 256, 0, st>>>(f, offsets, o, frames);
  else if (feat->dtype == DBSR_BF16 && out->dtype == DBSR_BF16)
    warp_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>(f, offsets, o, frames);
  else if (feat->dtype == DBSR_F32 && out->dtype == DBSR_BF16)
    warp_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>(f, offsets, o, frames);
  else warp_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>(f, offsets, o, frames);
  return check_launch("warp");
}

extern "C" int dbsr_softmax_wsum(const dbsr_nhwc_t* feat, const dbsr_nhwc_t* logits, const float* offsets,
                                 const dbsr_nhwc_t* fused, float* weights_out, int32_t frames, void* stream) {
  DBSR_REQUIRE(view_ok(feat) && view_ok(logits) && view_ok(fused) && frames >= 1, "softmax_wsum: bad arguments");
  DBSR_REQUIRE(feat->n == fused->n * frames && logits->n == feat->n && feat->h == fused->h && feat->w == fused->w &&
                   logits->h == fused->h && logits->w == fused->w && feat->c == fused->c && logits->c == fused->c,
               "softmax_wsum: geometry mismatch");
  DBSR_REQUIRE(vec4_ok(feat) && vec4_ok(logits) && vec4_ok(fused),
               "softmax_wsum: channel count/offset/pitch must be multiples of 4 and aligned");
  cudaStream_t st = (cudaStream_t)stream;
  View f = make_view(feat), l = make_view(logits), o = make_view(fused);
  const int key = feat->dtype * 4 + logits->dtype * 2 + fused->dtype;
  const bool v8 = vec8_ok(feat) && vec8_ok(logits) && vec8_ok(fused) && (key == 0 || key == 7 || key == 5);
  const long long total = (long long)fused->n * fused->h * fused->w * (fused->c / 4);
  const int g = grid_cap(total, 256);
  if (v8) {
    dim3 grid8(((fused->w + WS_TW - 1) / WS_TW) * ((fused->h + WS_TH - 1) / WS_TH), (fused->c + 63) / 64, fused->n);
    // bf16 in / bf16 out with the warp folded in (the engine's path): the pair-pipelined cp.async kernel.  Its gather
    // records hold 32-bit byte offsets inside one image and at most WSP_MAX_OTHERS non-reference frames
    if (key == 7 && offsets != nullptr && frames >= 2 && frames - 1 <= WSP_MAX_OTHERS &&
        (long long)fused->h * fused->w * feat->c_pitch * 2 < (1ll << 32)) {
      // TMA mode (see the kernel): 1 = reference-frame tiles (default), 2 = + all logit tiles, 0 = none (A/B: DBSR_WSUM_TMA)
      static const int want_mode = getenv("DBSR_WSUM_TMA") ? atoi(getenv("DBSR_WSUM_TMA")) : 1;
      EncodeTiledFn encode = want_mode > 0 ? get_encode() : nullptr;
      alignas(64) CUtensorMap mf, ml;
      memset(&mf, 0, sizeof(mf)); memset(&ml, 0, sizeof(ml));
      int mode = encode != nullptr ? (want_mode >= 2 ? 2 : 1) : 0;
      if (mode) {
        const dbsr_nhwc_t* vs[2] = {feat, logits};
        CUtensorMap* ms[2] = {&mf, &ml};
        for (int i = 0; i < 2 && mode; ++i) {
          const dbsr_nhwc_t* v = vs[i];
          cuuint64_t dims[4] = {(cuuint64_t)v->c, (cuuint64_t)v->w, (cuuint64_t)v->h, (cuuint64_t)v->n};
          cuuint64_t strides[3] = {(cuuint64_t)v->c_pitch * 2, (cuuint64_t)v->w * v->c_pitch * 2, (cuuint64_t)v->h * v->w * v->c_pitch * 2};
          cuuint32_t box[4] = {64u, 4u, 1u, 1u};
          cuuint32_t es[4] = {1, 1, 1, 1};
          void* base = reinterpret_cast<__nv_bfloat16*>(v->data) + v->c_off;
          if (encode(ms[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            mode = 0;
        }
      }
      typedef void (*WsKernel)(const CUtensorMap, const CUtensorMap, View, View, const float*, View, int);
      const WsKernel kern = mode == 2 ? softmax_wsum8_pair_kernel<__nv_bfloat16, 2>
                            : (mode == 1 ? softmax_wsum8_pair_kernel<__nv_bfloat16, 1> : softmax_wsum8_pair_kernel<__nv_bfloat16, 0>);
      const int smem = mode == 2 ? WST_SMEM2 : (mode == 1 ? WST_SMEM1 : WSP_SMEM);
      static bool attr_set_dev[MAX_DEVICES][3] = {};
      bool& attr_set = attr_set_dev[current_device_slot()][mode];
      if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DBSR_REQUIRE(e == cudaSuccess, "softmax_wsum: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        attr_set = true;
      }
      launch_pdl(kern, dim3(grid8), dim3(256), (size_t)smem, st, mf, ml, f, l, offsets, o, frames);
    }
    else if (key == 0) softmax_wsum8_kernel<float, float, float><<<grid8, 256, 0, st>>>(f, l, offsets, o, frames);
    else if (key == 7) softmax_wsum8_kernel<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16><<<grid8, 256, 0, st>>>(f, l, offsets, o, frames);
    else softmax_wsum8_kernel<__nv_bfloat16, float, __nv_bfloat16><<<grid8, 256, 0, st>>>(f, l, offsets, o, frames);
  } else
  switch (key) {
    case 0: softmax_wsum_kernel<float, float, float><<<g, 256, 0, st>>>(f, l, offsets, o, frames); break;
    case 7: softmax_wsum_kernel<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>(f, l, offsets, o, frames); break;
    case 5: softmax_wsum_kernel<__nv_bfloat16, float, __nv_bfloat16><<<g, 256, 0, st>>>(f, l, offsets, o, frames); break;
    case 4: softmax_wsum_kernel<__nv_bfloat16, float, float><<<g, 256, 0, st>>>(f, l, offsets, o, frames); break;
    default:
      set_error("softmax_wsum: unsupported dtype combination feat=%d logits=%d fused=%d", feat->dtype, logits->dtype,
                fused->dtype);
      return 1;
  }
  int rc = check_launch("softmax_wsum");
  if (rc) return rc;
  if (weights_out) {
    const long long tw = (long long)fused->n * fused->c * fused->h * fused->w;
    fusion_weights_kernel<<<grid_cap(tw, 256), 256, 0, st>>>(l, weights_out, frames);
    rc = check_launch("fusion_weights");
  }
  return rc;
}

extern "C" int dbsr_warp_proj(const dbsr_nhwc_t* q, const float* bias, const float* offsets, const dbsr_nhwc_t* wp_in,
                              int32_t frames, void* stream) {
  DBSR_REQUIRE(view_ok(q) && view_ok(wp_in) && bias && frames >= 2, "warp_proj: bad arguments");
  DBSR_REQUIRE(q->n == wp_in->n && q->h == wp_in->h && q->w == wp_in->w && wp_in->c >= 2 * q->c && q->n % frames == 0,
               "warp_proj: geometry mismatch");
  DBSR_REQUIRE(vec8_ok(q) && vec8_ok(wp_in) && ((uintptr_t)bias % 16) == 0 && q->dtype == wp_in->dtype,
               "warp_proj: channels must be multiples of 8, 16-byte aligned, same dtype in and out");
  const long long per_frame = (long long)q->h * q->w * (q->c / 8);
  DBSR_REQUIRE(per_frame < (1ll << 31), "warp_proj: more than 2^31 (pixel, channel group) items per frame");
  const dim3 g((unsigned)ceil_div(per_frame, 256), (unsigned)(q->n < 65535 ? q->n : 65535));
  cudaStream_t st = (cudaStream_t)stream;
  const View none = make_view(nullptr);
  if (q->dtype == DBSR_F32) launch_pdl(warp_proj_kernel<float, float, false>, dim3(g), dim3(256), 0, st, make_view(q), bias, offsets, make_view(wp_in), frames, none);
  else launch_pdl(warp_proj_kernel<__nv_bfloat16, __nv_bfloat16, false>, dim3(g), dim3(256), 0, st, make_view(q), bias, offsets, make_view(wp_in), frames, none);
  return check_launch("warp_proj");
}

extern "C" int dbsr_warp_proj_split(const dbsr_nhwc_t* q, const float* bias, const float* offsets, const dbsr_nhwc_t* wp_in,
                                    const dbsr_nhwc_t* p0, int32_t frames, void* stream) {
  DBSR_REQUIRE(view_ok(q) && view_ok(wp_in) && view_ok(p0) && bias && frames >= 2, "warp_proj_split: bad arguments");
  DBSR_REQUIRE(q->n == wp_in->n && q->h == wp_in->h && q->w == wp_in->w && wp_in->c >= q->c && q->n % frames == 0 &&
                   p0->n == q->n / frames && p0->h == q->h && p0->w == q->w && p0->c == q->c, "warp_proj_split: geometry mismatch");
  DBSR_REQUIRE(vec8_ok(q) && vec8_ok(wp_in) && vec8_ok(p0) && ((uintptr_t)bias % 16) == 0 && q->dtype == wp_in->dtype &&
                   q->dtype == p0->dtype, "warp_proj_split: channels must be multiples of 8, 16-byte aligned, same dtype in and out");
  const long long per_frame = (long long)q->h * q->w * (q->c / 8);
  DBSR_REQUIRE(per_frame < (1ll << 31), "warp_proj_split: more than 2^31 (pixel, channel group) items per frame");
  // frames per CTA: a thread that handles one (pixel, 8 channels) of ONE frame lives for two dependent memory round trips;
  // walking several frames per thread amortises the index arithmetic and keeps the SMs' CTA slots turning over less often
  static const int fpb_env = getenv("DBSR_WARP_PROJ_FPB") ? atoi(getenv("DBSR_WARP_PROJ_FPB")) : 0;
  // (32 bursts of 48^2: 0.148 -> 0.115 ms at 4 frames per CTA, 0.120 at 8, 0.126 at 14)
  const int fpb = fpb_env > 0 ? fpb_env : ((long long)ceil_div(q->n, 4) * ceil_div(per_frame, 256) >= 4 * 148 ? 4 : 1);
  const int gy = ceil_div(q->n, fpb);
  const dim3 g((unsigned)ceil_div(per_frame, 256), (unsigned)(gy < 65535 ? gy : 65535));
  cudaStream_t st = (cudaStream_t)stream;
  if (q->dtype == DBSR_F32) launch_pdl(warp_proj_kernel<float, float, true>, dim3(g), dim3(256), 0, st, make_view(q), bias, offsets, make_view(wp_in), frames, make_view(p0));
  else launch_pdl(warp_proj_kernel<__nv_bfloat16, __nv_bfloat16, true>, dim3(g), dim3(256), 0, st, make_view(q), bias, offsets, make_view(wp_in), frames, make_view(p0));
  return check_launch("warp_proj_split");
}

extern "C" int dbsr_blur3x3(const dbsr_nhwc_t* x, const dbsr_nhwc_t* y, const float* k9, void* stream) {
  DBSR_REQUIRE(view_ok(x) && view_ok(y) && k9 && x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c &&
                   x->dtype == y->dtype, "blur3x3: bad arguments");
  DBSR_REQUIRE(vec8_ok(x) && vec8_ok(y), "blur3x3: channels must be multiples of 8 and 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  DBSR_REQUIRE(x->n <= 65535, "blur3x3: more than 65535 images");
  // rank-1 kernel k9[i][j] = b[i] * a[j] (every Gaussian): separable kernel
  {
    int pi = 0;
    for (int i = 1; i < 9; ++i) if (fabsf(k9[i]) > fabsf(k9[pi])) pi = i;
    const int pr = pi / 3, pc = pi % 3;
    float a[3], b[3], kmax = fabsf(k9[pi]), dev = 0.0f;
    if (kmax > 0.0f) {
      for (int j = 0; j < 3; ++j) a[j] = k9[pr * 3 + j];
      for (int i = 0; i < 3; ++i) b[i] = k9[i * 3 + pc] / k9[pi];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dev = fmaxf(dev, fabsf(k9[i * 3 + j] - b[i] * a[j]));
      if (dev <= 1e-7f * kmax) {
        // rows per thread: 32 (6 % halo re-reads) when that still gives every SM several CTAs, fewer for a few images
        // (2 x 384^2 x 32 channels at 32 rows is 144 CTAs of serial row loops: 35 us for 19 MB)
        static const int rows_env = getenv("DBSR_BLUR_ROWS") ? atoi(getenv("DBSR_BLUR_ROWS")) : 0;
        const long long cols = ceil_div((long long)x->w * (x->c / 8), 256);
        int rows = BLUR_SEP_ROWS;
        while (rows > 4 && cols * ceil_div(x->h, rows) * x->n < 148LL * 3) rows >>= 1;
        if (rows_env > 0) rows = rows_env;
        dim3 gs((unsigned)cols, (unsigned)ceil_div(x->h, rows), (unsigned)x->n);
        if (x->dtype == DBSR_F32)
          launch_pdl(blur3x3_sep_kernel<float>, dim3(gs), dim3(256), 0, st, make_view(x), make_view(y), a[0], a[1], a[2], b[0], b[1], b[2], rows);
        else
          launch_pdl(blur3x3_sep_kernel<__nv_bfloat16>, dim3(gs), dim3(256), 0, st, make_view(x), make_view(y), a[0], a[1], a[2], b[0], b[1], b[2], rows);
        return check_launch("blur3x3");
      }
    }
  }
  dim3 grid((unsigned)ceil_div((long long)x->w * (x->c / 8), 256), (unsigned)ceil_div(x->h, BLUR_ROWS), (unsigned)x->n);
  if (x->dtype == DBSR_F32)
    launch_pdl(blur3x3_rows_kernel<float>, dim3(grid), dim3(256), 0, st, make_view(x), make_view(y), k9[0], k9[1], k9[2], k9[3], k9[4], k9[5], k9[6], k9[7], k9[8]);
  else
    launch_pdl(blur3x3_rows_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, make_view(x), make_view(y), k9[0], k9[1], k9[2], k9[3], k9[4], k9[5], k9[6], k9[7], k9[8]);
  return check_launch("blur3x3");
}
