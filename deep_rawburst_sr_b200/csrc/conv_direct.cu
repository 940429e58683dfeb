// CUDA-core convolution (exact fp32 accumulate): smem-tiled implicit GEMM with on-the-fly im2col gather.
// Covers every conv shape on the path (any Cin/Cout, k in {1,3}, stride, dilation, channel-slice I/O,
// residual, activation, pixel-shuffle scatter).  This is the fp32 ("exact") path and the path of the
// PWC-Net long tail; the heavy DBSR convs go through conv_tc.cu (tcgen05) on the bf16 path.
// Replaces: nn.Conv2d + activation (+ residual) in reference models/layers/blocks.py:46-96 and
// models/alignment/pwcnet.py:49-204.
#include "common.cuh"

namespace dbsr {

struct ConvDirectParams {
  View x, y, r;
  const float* w;     // [taps][Cin][Cout]
  const float* bias;  // [Cout] or null
  int ksize, stride, dil, pad, act, shuffle_r;
  int Ho, Wo;
  long long M;
  int vecA;  // 1: input channel groups of 4 are 16B(f32)/8B(bf16) aligned and fully inside Cin
  int vecB;  // 1: Cout % 4 == 0
};

constexpr int BM = 128;
constexpr int BK = 16;

template <typename TI>
__device__ __forceinline__ void load_a8(const ConvDirectParams& p, const TI* base, bool inb, int ci, float (&a)[8]) {
  // 8 consecutive channels starting at ci of one input pixel (zero when out of bounds / beyond Cin)
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.0f;
  if (!inb) return;
  const int Cin = p.x.c;
  if (p.vecA && ci + 8 <= Cin) {
    if constexpr (sizeof(TI) == 4) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(base + ci));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + ci + 4));
      a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w; a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
    } else {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + ci));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(h[j]);
        a[2 * j] = f.x; a[2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (ci + j < Cin) a[j] = ld_as_float<TI>(base + ci + j);
  }
}

template <typename TI, int BN>
__global__ void __launch_bounds__(256) conv_direct_kernel(const ConvDirectParams p) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float A_s[2][BK][BM];
  __shared__ __align__(16) float B_s[2][BK][BN];

  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Cin = p.x.c, Cout = p.y.c * (p.shuffle_r > 1 ? p.shuffle_r * p.shuffle_r : 1);
  const int H = p.x.h, W = p.x.w;

  // --- loader role: one output pixel, 8 channels of the 16-wide K chunk
  const int lm = t & (BM - 1);
  const int lk = (t >> 7) * 8;
  const long long m_ld = m0 + lm;
  const bool m_ok = m_ld < p.M;
  int ln = 0, loy = 0, lox = 0;
  if (m_ok) {
    ln = (int)(m_ld / ((long long)p.Ho * p.Wo));
    int rem = (int)(m_ld - (long long)ln * p.Ho * p.Wo);
    loy = rem / p.Wo;
    lox = rem - loy * p.Wo;
  }
  const TI* xbase = reinterpret_cast<const TI*>(p.x.data) + p.x.c_off;

  // --- B loader role
  const int bk = t >> 4;
  const int bn = (t & 15) * TN;

  const int taps = p.ksize * p.ksize;
  const int cchunks = (Cin + BK - 1) / BK;
  const int nchunks = taps * cchunks;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  float a_pf[8];
  float b_pf[TN];

  auto fetch = [&](int chunk) {
    const int tap = chunk / cchunks;
    const int ci0 = (chunk - tap * cchunks) * BK;
    const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
    const int iy = loy * p.stride - p.pad + ky * p.dil;
    const int ix = lox * p.stride - p.pad + kx * p.dil;
    const bool inb = m_ok && iy >= 0 && iy < H && ix >= 0 && ix < W;
    const TI* px = xbase + ((long long)(ln * H + iy) * W + ix) * p.x.c_pitch;
    load_a8<TI>(p, px, inb, ci0 + lk, a_pf);
    const int ci = ci0 + bk;
    const float* wrow = p.w + ((long long)tap * Cin + ci) * Cout + n0 + bn;
    if (TN == 4 && p.vecB && ci < Cin && n0 + bn + 4 <= Cout) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(wrow));
      b_pf[0] = v.x;
      if (TN > 1) { b_pf[1 % TN] = v.y; b_pf[2 % TN] = v.z; b_pf[3 % TN] = v.w; }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) b_pf[j] = (ci < Cin && n0 + bn + j < Cout) ? __ldg(wrow + j) : 0.0f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) A_s[buf][lk + j][lm] = a_pf[j];
#pragma unroll
    for (int j = 0; j < TN; ++j) B_s[buf][bk][bn + j] = b_pf[j];
  };

  const int tx = t & 15, ty = t >> 4;

  fetch(0);
  stash(0);
  __syncthreads();
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int buf = chunk & 1;
    if (chunk + 1 < nchunks) fetch(chunk + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&A_s[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&A_s[buf][kk][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      if constexpr (TN == 4) {
        const float4 bv = *reinterpret_cast<const float4*>(&B_s[buf][kk][tx * 4]);
        b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = B_s[buf][kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (chunk + 1 < nchunks) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  // --- epilogue
  const int r = p.shuffle_r > 1 ? p.shuffle_r : 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    const int n = (int)(m / ((long long)p.Ho * p.Wo));
    const int rem = (int)(m - (long long)n * p.Ho * p.Wo);
    const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co >= Cout) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + co);
      long long opix;
      int och;
      if (r > 1) {
        const int c = co / (r * r);
        const int sub = co - c * r * r;
        const int si = sub / r, sj = sub - si * r;
        opix = ((long long)n * p.y.h + (oy * r + si)) * p.y.w + (ox * r + sj);
        och = c;
      } else {
        opix = m;
        och = co;
      }
      if (p.r.data) v += view_ld(p.r, opix, och);
      v = apply_act(v, p.act);
      view_st(p.y, opix, och, v);
    }
  }
}

}  // namespace dbsr

using namespace dbsr;

extern "C" int dbsr_conv2d_direct(const dbsr_conv_t* c, void* stream) {
  DBSR_REQUIRE(c != nullptr, "conv2d_direct: null descriptor");
  DBSR_REQUIRE(view_ok(&c->x) && view_ok(&c->y), "conv2d_direct: bad x/y view");
  DBSR_REQUIRE(c->w != nullptr, "conv2d_direct: null weights");
  DBSR_REQUIRE(c->residual_group <= 1, "conv2d_direct: residual_group is a feature of the tensor-core path");
  DBSR_REQUIRE(c->ksize == 1 || c->ksize == 3, "conv2d_direct: ksize %d unsupported", c->ksize);
  DBSR_REQUIRE(c->stride >= 1 && c->dilation >= 1, "conv2d_direct: bad stride/dilation");
  ConvDirectParams p;
  p.x = make_view(&c->x);
  p.y = make_view(&c->y);
  p.r = make_view(c->residual.data ? &c->residual : nullptr);
  p.w = reinterpret_cast<const float*>(c->w);
  p.bias = c->bias;
  p.ksize = c->ksize; p.stride = c->stride; p.dil = c->dilation;
  p.pad = c->dilation * (c->ksize - 1) / 2;
  p.act = c->act;
  p.shuffle_r = c->shuffle_r > 1 ? c->shuffle_r : 1;
  p.Ho = (c->x.h + 2 * p.pad - c->dilation * (c->ksize - 1) - 1) / c->stride + 1;
  p.Wo = (c->x.w + 2 * p.pad - c->dilation * (c->ksize - 1) - 1) / c->stride + 1;
  DBSR_REQUIRE(c->y.n == c->x.n && c->y.h == p.Ho * p.shuffle_r && c->y.w == p.Wo * p.shuffle_r,
               "conv2d_direct: output view %dx%dx%d does not match conv geometry %dx%dx%d (r=%d)", c->y.n, c->y.h,
               c->y.w, c->x.n, p.Ho, p.Wo, p.shuffle_r);
  if (p.r.data)
    DBSR_REQUIRE(c->residual.n == c->y.n && c->residual.h == c->y.h && c->residual.w == c->y.w &&
                     c->residual.c == c->y.c, "conv2d_direct: residual geometry mismatch");
  p.M = (long long)c->x.n * p.Ho * p.Wo;
  const int Cout = c->y.c * p.shuffle_r * p.shuffle_r;
  const size_t es = elem_size(c->x.dtype);
  p.vecA = ((c->x.c_off * es) % 16 == 0 && (c->x.c_pitch * es) % 16 == 0 && ((uintptr_t)c->x.data % 16) == 0 &&
            c->x.c % 8 == 0) ? 1 : 0;
  p.vecB = (Cout % 4 == 0 && ((uintptr_t)c->w % 16) == 0) ? 1 : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool narrow = Cout <= 16;
  dim3 grid(ceil_div(p.M, BM), ceil_div(Cout, narrow ? 16 : 64));
  if (c->x.dtype == DBSR_F32) {
    if (narrow) conv_direct_kernel<float, 16><<<grid, 256, 0, st>>>(p);
    else conv_direct_kernel<float, 64><<<grid, 256, 0, st>>>(p);
  } else {
    if (narrow) conv_direct_kernel<__nv_bfloat16, 16><<<grid, 256, 0, st>>>(p);
    else conv_direct_kernel<__nv_bfloat16, 64><<<grid, 256, 0, st>>>(p);
  }
  return check_launch("conv2d_direct");
}
