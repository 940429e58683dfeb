// Flow-based backward warping of the frame embeddings and the fused warp + softmax + weighted sum.
//
//   dbsr_warp          : models/layers/warp.py:19-46 (grid_sample bilinear, zeros padding), NHWC, 16-byte
//                        vectorised along channels so the 4 taps are coalesced 128-channel rows.
//   dbsr_softmax_wsum  : models/dbsr/merging.py:117-124 (softmax over the burst + weighted sum) with the warp
//                        of encoders.py:80 recomputed on the fly, so neither `oth_feat` nor the normalised
//                        weights are materialised.  Per (burst, pixel, 4-channel group) one thread streams the N
//                        logits and the N (gathered) embeddings once, with an online softmax in fp32.
#include "common.cuh"

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
  const long long total = (long long)fused->n * fused->h * fused->w * (fused->c / 4);
  const int g = grid_cap(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  View f = make_view(feat), l = make_view(logits), o = make_view(fused);
  const int key = feat->dtype * 4 + logits->dtype * 2 + fused->dtype;
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
