// PWC-Net cost volume (81 displacements, mean over channels), shared-memory tiled.
//
// Replaces the three cupy launches + two zero-padded NHWC temporaries of the reference
// (external/pwcnet/correlation/correlation.py:8-103, 280-330), the LeakyReLU that follows it
// (models/alignment/pwcnet.py:161,169) and -- when a flow is given -- `backwarp` of the second feature map
// (pwcnet.py:16-38), which is fused into the staging of the f2 halo tile so the warped map never touches HBM.
//
// Work decomposition: one CTA = one pair x one 8x16 output tile.  288 threads = 9 (dy) x 8 (rows) x 4 (strips
// of 4 pixels); a thread owns 4 pixels x 9 dx = 36 accumulators for its dy.  Channels are walked in chunks
// of 32 staged in smem as [pixel][32] fp32; inside a chunk each lane walks the 8 float4 groups in a rotated
// order ((j + lane) & 7) so that the 8 lanes of an LDS.128 phase hit 8 distinct 16-byte bank groups.
#include "common.cuh"

namespace dbsr {

constexpr int CT_H = 8, CT_W = 16, C_CH = 32;
constexpr int HALO_H = CT_H + 8, HALO_W = CT_W + 8;
constexpr int CORR_THREADS = 288;
// f2 halo tile + f1 tile (fp32, one 32-channel chunk) + one 32-byte gather record per halo pixel (fused backwarp)
struct __align__(16) CorrRec { uint32_t o[4]; float w[4]; };
constexpr int CORR_SMEM = (HALO_H * HALO_W + CT_H * CT_W) * C_CH * (int)sizeof(float) + HALO_H * HALO_W * (int)sizeof(CorrRec);

struct Vec8c { float v[8]; };
__device__ __forceinline__ Vec8c vec8_zero() {
  Vec8c r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0.0f;
  return r;
}
// 8 consecutive channels (16-byte aligned address); channels >= valid read as zero
template <typename T> __device__ __forceinline__ Vec8c vec8_ld(const T* p, int valid);
template <> __device__ __forceinline__ Vec8c vec8_ld<float>(const float* p, int valid) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  Vec8c r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  if (valid < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i >= valid) r.v[i] = 0.0f;
  }
  return r;
}
template <> __device__ __forceinline__ Vec8c vec8_ld<__nv_bfloat16>(const __nv_bfloat16* p, int valid) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  Vec8c r;
  r.v[0] = __uint_as_float(q.x << 16); r.v[1] = __uint_as_float(q.x & 0xFFFF0000u);
  r.v[2] = __uint_as_float(q.y << 16); r.v[3] = __uint_as_float(q.y & 0xFFFF0000u);
  r.v[4] = __uint_as_float(q.z << 16); r.v[5] = __uint_as_float(q.z & 0xFFFF0000u);
  r.v[6] = __uint_as_float(q.w << 16); r.v[7] = __uint_as_float(q.w & 0xFFFF0000u);
  if (valid < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i >= valid) r.v[i] = 0.0f;
  }
  return r;
}
__device__ __forceinline__ void vec8_sts(float* dst, const Vec8c& v) {
  *reinterpret_cast<float4*>(dst) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

struct CorrParams {
  View f1, f2, flow, out;
  float flow_scale;
  int group, act;
  int tiles_x;
  int vec_out;      // 1: the 81-channel volume is stored as 11 groups of 8 channels (the 7 pad channels are zeroed)
  View f1_copy;     // data != NULL: f1_copy[p] = f1[i1(p)] (the first map's slice of the decoder's concat buffer, pwcnet.py:173)
                    // written from the tile this kernel stages anyway (tensor-core and small-map kernels)
};
// 8 consecutive channels to global memory (16-byte aligned address, host-checked); channels >= valid are not written
template <typename T> __device__ __forceinline__ void vec8_stg(T* dst, const Vec8c& v, int valid);
template <> __device__ __forceinline__ void vec8_stg<float>(float* dst, const Vec8c& v, int valid) {
  if (valid >= 8) {
    *reinterpret_cast<float4*>(dst) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < valid) dst[i] = v.v[i];
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ void vec8_stg<__nv_bfloat16>(__nv_bfloat16* dst, const Vec8c& v, int valid) {
  if (valid >= 8) {     // the values came from bf16: the rounding is exact
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v.v[0], v.v[1]), pack_bf16x2(v.v[2], v.v[3]),
                                                pack_bf16x2(v.v[4], v.v[5]), pack_bf16x2(v.v[6], v.v[7]));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < valid) dst[i] = __float2bfloat16_rn(v.v[i]);
  }
}

template <bool VEC, typename T>
__global__ void __launch_bounds__(CORR_THREADS, 2) corr81_kernel(const CorrParams p) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  extern __shared__ __align__(16) float smem[];
  float* f2_s = smem;                                // [HALO_H*HALO_W][32]
  float* f1_s = smem + HALO_H * HALO_W * C_CH;       // [CT_H*CT_W][32]
  CorrRec* rec_s = reinterpret_cast<CorrRec*>(f1_s + CT_H * CT_W * C_CH);   // [HALO_H*HALO_W]

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int strip = t & 3, row = (t >> 2) & 7, dy = t >> 5;
  const int pair = blockIdx.y;
  const int ty0 = (blockIdx.x / p.tiles_x) * CT_H, tx0 = (blockIdx.x % p.tiles_x) * CT_W;
  const int H = p.f1.h, W = p.f1.w, C = p.f1.c;
  int i1 = pair, i2 = pair;
  if (p.group > 0) {
    const int b = pair / p.group;
    i1 = b * (p.group + 1);
    i2 = i1 + 1 + (pair - b * p.group);
  }
  const long long base1 = (long long)i1 * H * W, base2 = (long long)i2 * H * W, basef = (long long)pair * H * W;
  const bool warp2 = p.flow.data != nullptr;
  const float sxw = warp2 ? p.flow_scale * (float)W / (float)(W - 1) : 0.0f;  // pwcnet.py:28 + linspace grid :20
  const float syh = warp2 ? p.flow_scale * (float)H / (float)(H - 1) : 0.0f;

  float acc[4][9];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int d = 0; d < 9; ++d) acc[i][d] = 0.0f;
  // rows / columns of this tile that are inside the map: small pyramid levels (1x1 ... 8x8) use a fraction of the 8x16
  // tile, and all staging / accumulation work is bounded by it
  const int th = min(CT_H, H - ty0), tw = min(CT_W, W - tx0);
  const int hh = th + 8, hw = tw + 8;                  // halo extent actually needed
  const bool active = row < th && strip * 4 < tw;      // this thread owns at least one real output pixel

  if (VEC && warp2) {
    // ---- backwarp records, once per CTA (they do not depend on the channel chunk): per halo pixel the pixel offsets
    // of the 4 bilinear taps (clamped into the image) and their weights with out-of-image taps zeroed; pwcnet.py:34-38
    // zeroes the whole pixel unless the sampled ones-channel (= the sum of the in-image weights) exceeds 0.999.
    for (int pix = t; pix < hh * HALO_W; pix += CORR_THREADS) {
      const int y = ty0 - 4 + pix / HALO_W, x = tx0 - 4 + pix % HALO_W;
      CorrRec r;
      r.o[0] = r.o[1] = r.o[2] = r.o[3] = 0u;
      r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.0f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const long long fp = basef + (long long)y * W + x;
        const float u = (float)x + view_ld(p.flow, fp, 0) * sxw;
        const float w = (float)y + view_ld(p.flow, fp, 1) * syh;
        const float fu = floorf(u), fv = floorf(w);
        const float ax = u - fu, ay = w - fv;
        const int xa = (int)fminf(fmaxf(fu, -2.0f), (float)W), ya = (int)fminf(fmaxf(fv, -2.0f), (float)H);
        float m = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int xx = xa + (k & 1), yy = ya + (k >> 1);
          const float wt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
          if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
            r.o[k] = (uint32_t)(yy * W + xx);
            r.w[k] = wt;
            m += wt;
          }
        }
        if (!(m > 0.999f)) r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.0f;
      }
      rec_s[pix] = r;
    }
    __syncthreads();
  }

  for (int c0 = 0; c0 < C; c0 += C_CH) {
    if (VEC) {
      // ---- vectorised staging: one task = one pixel x 8 channels (16 / 32 bytes of global memory per load)
      const T* b1 = reinterpret_cast<const T*>(p.f1.data) + p.f1.c_off;
      const T* b2 = reinterpret_cast<const T*>(p.f2.data) + p.f2.c_off;
      for (int e = t; e < th * CT_W * (C_CH / 8); e += CORR_THREADS) {      // rows >= th are never read
        const int g = e & 3, pix = e >> 2;
        const int y = ty0 + pix / CT_W, x = tx0 + pix % CT_W;
        const int ch = c0 + g * 8;
        Vec8c v = vec8_zero();
        if (x < W && ch < C) v = vec8_ld<T>(b1 + (base1 + (long long)y * W + x) * p.f1.c_pitch + ch, C - ch);
        vec8_sts(&f1_s[pix * C_CH + g * 8], v);
      }
      for (int e = t; e < hh * HALO_W * (C_CH / 8); e += CORR_THREADS) {    // halo rows >= th + 8 are never read
        const int g = e & 3, pix = e >> 2;
        const int y = ty0 - 4 + pix / HALO_W, x = tx0 - 4 + pix % HALO_W;
        const int ch = c0 + g * 8;
        Vec8c v = vec8_zero();
        if (y >= 0 && y < H && x >= 0 && x < W && ch < C) {
          if (!warp2) {
            v = vec8_ld<T>(b2 + (base2 + (long long)y * W + x) * p.f2.c_pitch + ch, C - ch);
          } else {
            const uint4 ro = *reinterpret_cast<const uint4*>(rec_s[pix].o);
            const float4 rw = *reinterpret_cast<const float4*>(rec_s[pix].w);
            const uint32_t ok[4] = {ro.x, ro.y, ro.z, ro.w};
            const float wk[4] = {rw.x, rw.y, rw.z, rw.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (wk[k] != 0.0f) {
                const Vec8c a = vec8_ld<T>(b2 + (base2 + (long long)ok[k]) * p.f2.c_pitch + ch, C - ch);
#pragma unroll
                for (int i = 0; i < 8; ++i) v.v[i] = fmaf(a.v[i], wk[k], v.v[i]);
              }
            }
          }
        }
        vec8_sts(&f2_s[pix * C_CH + g * 8], v);
      }
    } else {
    // ---- stage f1 tile
      for (int e = t; e < th * CT_W * C_CH; e += CORR_THREADS) {      // rows >= th are never read
        const int ch = e & (C_CH - 1), pix = e >> 5;
        const int y = ty0 + pix / CT_W, x = tx0 + pix % CT_W;
        float v = 0.0f;
        if (y < H && x < W && c0 + ch < C) v = view_ld(p.f1, base1 + (long long)y * W + x, c0 + ch);
        f1_s[e] = v;
      }
      // ---- stage f2 halo tile (optionally backwarped)
      for (int e = t; e < hh * HALO_W * C_CH; e += CORR_THREADS) {       // halo rows >= th + 8 are never read
        const int ch = e & (C_CH - 1), pix = e >> 5;
        const int hx = pix % HALO_W;
        const int y = ty0 - 4 + pix / HALO_W, x = tx0 - 4 + hx;
        if (hx >= ((hw + 3) & ~3) + 4) { f2_s[e] = 0.0f; continue; }         // columns past the strips that exist
        float v = 0.0f;
        if (y >= 0 && y < H && x >= 0 && x < W && c0 + ch < C) {
          if (!warp2) {
            v = view_ld(p.f2, base2 + (long long)y * W + x, c0 + ch);
          } else {
            const long long fp = basef + (long long)y * W + x;
            const float u = (float)x + view_ld(p.flow, fp, 0) * sxw;
            const float w = (float)y + view_ld(p.flow, fp, 1) * syh;
            const float fu = floorf(u), fv = floorf(w);
            const float ax = u - fu, ay = w - fv;
            const int xa = (int)fu, ya = (int)fv;
            float s = 0.0f, m = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int xx = xa + (k & 1), yy = ya + (k >> 1);
              const float wt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
              if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
                s = fmaf(view_ld(p.f2, base2 + (long long)yy * W + xx, c0 + ch), wt, s);
                m += wt;
              }
            }
            v = (m > 0.999f) ? s : 0.0f;  // pwcnet.py:34-38
          }
        }
        f2_s[e] = v;
      }
    }
    __syncthreads();

    // ---- accumulate
    if (active)
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
      const int jj = ((j + lane) & 7) * 4;
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        a[i] = *reinterpret_cast<const float4*>(&f1_s[(row * CT_W + strip * 4 + i) * C_CH + jj]);
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const float4 b = *reinterpret_cast<const float4*>(&f2_s[((row + dy) * HALO_W + strip * 4 + k) * C_CH + jj]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int dx = k - i;
          if (dx >= 0 && dx <= 8) {
            float s = acc[i][dx];
            s = fmaf(a[i].x, b.x, s); s = fmaf(a[i].y, b.y, s); s = fmaf(a[i].z, b.z, s); s = fmaf(a[i].w, b.w, s);
            acc[i][dx] = s;
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue: scale, activation, transpose through smem.  out_s is displacement-major [81][OUT_P]: a thread's four
  // pixels of one displacement are one conflict-free STS.128 (a warp covers the 128 consecutive pixels of the tile).
  float* out_s = smem;
  constexpr int OUT_P = CT_H * CT_W + 4;
  const float invC = 1.0f / (float)C;
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    float4 v;
    v.x = apply_act(acc[0][d] * invC, p.act); v.y = apply_act(acc[1][d] * invC, p.act);
    v.z = apply_act(acc[2][d] * invC, p.act); v.w = apply_act(acc[3][d] * invC, p.act);
    *reinterpret_cast<float4*>(&out_s[(dy * 9 + d) * OUT_P + row * CT_W + strip * 4]) = v;
  }
  __syncthreads();
  if (p.vec_out) {
    // 16-byte (bf16) / 2 x 16-byte (fp32) stores of 8 displacements; the 7 channels after the 81st are the pad of the
    // volume's 8-aligned concat segment (host-checked: c_off + 88 <= c_pitch) and are written as zeros
    for (int e = t; e < 11 * CT_H * CT_W; e += CORR_THREADS) {
      const int pix = e & (CT_H * CT_W - 1), g = e >> 7;
      const int y = ty0 + pix / CT_W, x = tx0 + pix % CT_W;
      if (y < H && x < W) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (g * 8 + j < 81) ? out_s[(g * 8 + j) * OUT_P + pix] : 0.0f;
        const long long o = ((long long)pair * H * W + (long long)y * W + x) * p.out.c_pitch + p.out.c_off + g * 8;
        if (p.out.dtype == DBSR_BF16) {
          uint4 q;
          q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out.data) + o) = q;
        } else {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out.data) + o);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
  } else {
    for (int e = t; e < 81 * CT_H * CT_W; e += CORR_THREADS) {
      const int pix = e & (CT_H * CT_W - 1), ch = e >> 7;
      const int y = ty0 + pix / CT_W, x = tx0 + pix % CT_W;
      if (y < H && x < W) view_st(p.out, (long long)pair * H * W + (long long)y * W + x, ch, out_s[ch * OUT_P + pix]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-core cost volume for bf16 feature maps (the `precision='bf16'` path), C in {32, 64, 96, 128}.
// out[p, dy, dx] = sum_c f1[p, c] * f2[p + (dy, dx), c] is a BANDED product: for the 16 pixels of one tile row and the 24
// halo pixels of row (y + dy), G = F1 (16 x C) * F2^T (C x 24) holds the 16 x 9 wanted values on its diagonals
// (q - x = dx in 0..8), i.e. 37.5 % of G.  On CUDA cores the kernel above is instruction-bound (27.5 M warp instructions
// for the level-2 shape, 41 % FFMA); here one warp = one dy computes G with 3 (n-tiles) x C/16 warp-level
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate) per tile row and scatters its accumulator fragments -- whose (row, column)
// = (pixel, halo pixel) is known per lane -- straight into the displacement-major output staging tile; the vectorised
// store epilogue is the one of corr81_kernel.  (tcgen05 would need the band extracted from TMEM, where a warp reads the
// same columns for all 32 lanes: the register-fragment layout of mma.sync is what makes the diagonal scatter free.)
// Both maps of the tile are staged ONCE for all channels as bf16 [pixel][C + 8] (pitch = 4 banks mod 32: the fragment
// loads -- 8 pixels x 4 consecutive words per LDS.32 -- are conflict-free); the fused backwarp rounds the warped f2 tile
// to bf16 (the operand type of the tensor core; same order as the bf16 rounding of the volume it produces).
// ---------------------------------------------------------------------------------------------------------
constexpr int MMA_ROWS = 4;                          // tile rows per output phase (two phases per 8-row tile)
constexpr int OUT_PM = MMA_ROWS * CT_W + 4;          // displacement-major staging tile [81][OUT_PM] of one phase
template <int KS> struct CorrMma {
  static constexpr int PITCH = KS * 16 + 8;                                  // bf16 elements per staged pixel
  static constexpr int PW = PITCH / 2;                                       // 32-bit words per staged pixel
  static constexpr int F2_BYTES = HALO_H * HALO_W * PITCH * 2;
  static constexpr int F1_BYTES = CT_H * CT_W * PITCH * 2;
  static constexpr int REC_BYTES = HALO_H * HALO_W * (int)sizeof(CorrRec);
  static constexpr int OUT_BYTES = 81 * OUT_PM * (int)sizeof(float);
  // the output staging tile reuses the backwarp records' storage (the records are dead once the f2 tile is staged)
  static constexpr int SMEM = F2_BYTES + F1_BYTES + (REC_BYTES > OUT_BYTES ? REC_BYTES : OUT_BYTES);
};

__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <int KS>
__global__ void __launch_bounds__(CORR_THREADS, (CorrMma<KS>::SMEM <= 74 * 1024) ? 3 : ((CorrMma<KS>::SMEM <= 110 * 1024) ? 2 : 1)) corr81_mma_kernel(const CorrParams p) {
  using G = CorrMma<KS>;
  griddep_launch_dependents_if_small();
  griddep_wait();
  extern __shared__ __align__(16) unsigned char smem_b[];
  __nv_bfloat16* f2_s = reinterpret_cast<__nv_bfloat16*>(smem_b);                               // [HALO_H*HALO_W][PITCH]
  __nv_bfloat16* f1_s = reinterpret_cast<__nv_bfloat16*>(smem_b + G::F2_BYTES);                 // [CT_H*CT_W][PITCH]
  CorrRec* rec_s = reinterpret_cast<CorrRec*>(smem_b + G::F2_BYTES + G::F1_BYTES);              // [HALO_H*HALO_W]
  float* out_s = reinterpret_cast<float*>(rec_s);                                               // [81][OUT_PM], after staging

  const int t = threadIdx.x;
  const int lane = t & 31, dy = t >> 5;
  const int pair = blockIdx.y;
  const int ty0 = (blockIdx.x / p.tiles_x) * CT_H, tx0 = (blockIdx.x % p.tiles_x) * CT_W;
  const int H = p.f1.h, W = p.f1.w, C = p.f1.c;
  int i1 = pair, i2 = pair;
  if (p.group > 0) {
    const int b = pair / p.group;
    i1 = b * (p.group + 1);
    i2 = i1 + 1 + (pair - b * p.group);
  }
  const long long base1 = (long long)i1 * H * W, base2 = (long long)i2 * H * W, basef = (long long)pair * H * W;
  const bool warp2 = p.flow.data != nullptr;
  const float sxw = warp2 ? p.flow_scale * (float)W / (float)(W - 1) : 0.0f;
  const float syh = warp2 ? p.flow_scale * (float)H / (float)(H - 1) : 0.0f;
  const int th = min(CT_H, H - ty0);
  const int hh = th + 8;

  if (warp2) {      // backwarp records, as in corr81_kernel
    for (int pix = t; pix < hh * HALO_W; pix += CORR_THREADS) {
      const int y = ty0 - 4 + pix / HALO_W, x = tx0 - 4 + pix % HALO_W;
      CorrRec r;
      r.o[0] = r.o[1] = r.o[2] = r.o[3] = 0u;
      r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.0f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const long long fp = basef + (long long)y * W + x;
        const float u = (float)x + view_ld(p.flow, fp, 0) * sxw;
        const float w = (float)y + view_ld(p.flow, fp, 1) * syh;
        const float fu = floorf(u), fv = floorf(w);
        const float ax = u - fu, ay = w - fv;
        const int xa = (int)fminf(fmaxf(fu, -2.0f), (float)W), ya = (int)fminf(fmaxf(fv, -2.0f), (float)H);
        float m = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int xx = xa + (k & 1), yy = ya + (k >> 1);
          const float wt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
          if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
            r.o[k] = (uint32_t)(yy * W + xx) * (uint32_t)(p.f2.c_pitch * 2);
            r.w[k] = wt;
            m += wt;
          }
        }
        if (!(m > 0.999f)) r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.0f;
      }
      rec_s[pix] = r;
    }
    __syncthreads();
  }

  // ---- staging: one task = one pixel x 8 channels = one 16-byte global load and one 16-byte shared store
  constexpr int G8 = KS * 2;
  // (C is a multiple of 16 here: no channel tail to mask; per-image offsets fit 32 bits, host-checked)
  const __nv_bfloat16* b1 = reinterpret_cast<const __nv_bfloat16*>(p.f1.data) + p.f1.c_off + base1 * p.f1.c_pitch;
  const __nv_bfloat16* b2 = reinterpret_cast<const __nv_bfloat16*>(p.f2.data) + p.f2.c_off + base2 * p.f2.c_pitch;
  const unsigned char* b2b = reinterpret_cast<const unsigned char*>(b2);
  // plain copies go through cp.async (16 bytes, zero-filled when the pixel is outside the map: src-size 0), so that all
  // of a thread's loads are in flight at once instead of one load -> store round trip per task
  for (int e = t; e < th * CT_W * G8; e += CORR_THREADS) {
    const int g = e % G8, pix = e / G8;
    const int y = ty0 + pix / CT_W, x = tx0 + pix % CT_W;
    const bool in = x < W;
    cp_async16_zfill(f1_s + pix * G::PITCH + g * 8, b1 + (in ? (uint32_t)(y * W + x) * (uint32_t)p.f1.c_pitch + g * 8 : 0u), in);
  }
  if (!warp2) {
    for (int e = t; e < hh * HALO_W * G8; e += CORR_THREADS) {
      const int g = e % G8, pix = e / G8;
      const int y = ty0 - 4 + pix / HALO_W, x = tx0 - 4 + pix % HALO_W;
      const bool in = y >= 0 && y < H && x >= 0 && x < W;
      cp_async16_zfill(f2_s + pix * G::PITCH + g * 8, b2 + (in ? (uint32_t)(y * W + x) * (uint32_t)p.f2.c_pitch + g * 8 : 0u), in);
    }
  } else {
    // fused backwarp: the four taps are loaded unconditionally (a zero-weight tap points at pixel 0 of the image), so
    // the loads of a task -- and, unrolled by two, of the next one -- are issued back to back
#pragma unroll 2
    for (int e = t; e < hh * HALO_W * G8; e += CORR_THREADS) {
      const int g = e % G8, pix = e / G8;
      const uint4 ro = *reinterpret_cast<const uint4*>(rec_s[pix].o);
      const float4 rw = *reinterpret_cast<const float4*>(rec_s[pix].w);
      const unsigned char* src = b2b + g * 16;
      const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(src + ro.x)), q1 = __ldg(reinterpret_cast<const uint4*>(src + ro.y));
      const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(src + ro.z)), q3 = __ldg(reinterpret_cast<const uint4*>(src + ro.w));
      const uint32_t w0[4] = {q0.x, q0.y, q0.z, q0.w}, w1[4] = {q1.x, q1.y, q1.z, q1.w};
      const uint32_t w2[4] = {q2.x, q2.y, q2.z, q2.w}, w3[4] = {q3.x, q3.y, q3.z, q3.w};
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo = __uint_as_float(w0[i] << 16) * rw.x, hi = __uint_as_float(w0[i] & 0xFFFF0000u) * rw.x;
        lo = fmaf(__uint_as_float(w1[i] << 16), rw.y, lo); hi = fmaf(__uint_as_float(w1[i] & 0xFFFF0000u), rw.y, hi);
        lo = fmaf(__uint_as_float(w2[i] << 16), rw.z, lo); hi = fmaf(__uint_as_float(w2[i] & 0xFFFF0000u), rw.z, hi);
        lo = fmaf(__uint_as_float(w3[i] << 16), rw.w, lo); hi = fmaf(__uint_as_float(w3[i] & 0xFFFF0000u), rw.w, hi);
        o[i] = pack_bf16x2(lo, hi);
      }
      *reinterpret_cast<uint4*>(f2_s + pix * G::PITCH + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (p.f1_copy.data) {     // the staged f1 tile also goes to the concat slice (16-byte rows; alignment host-checked)
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.f1_copy.data) + p.f1_copy.c_off + basef * p.f1_copy.c_pitch;
    for (int e = t; e < th * CT_W * G8; e += CORR_THREADS) {
      const int g = e % G8, pix = e / G8;
      const int y = ty0 + pix / CT_W, x = tx0 + pix % CT_W;
      if (x < W)
        *reinterpret_cast<uint4*>(d + (long long)(y * W + x) * p.f1_copy.c_pitch + g * 8) =
            *reinterpret_cast<const uint4*>(f1_s + pix * G::PITCH + g * 8);
    }
  }

  // ---- banded product on the tensor cores: warp = dy, fragments straight from shared memory
  const uint32_t* f1_w = reinterpret_cast<const uint32_t*>(f1_s);
  const uint32_t* f2_w = reinterpret_cast<const uint32_t*>(f2_s);
  const int gq = lane >> 2, tq = lane & 3;
  // fragment element i of n-tile nt is (pixel x = gq + 8 (i >> 1), halo column q = 8 nt + 2 tq + (i & 1)): displacement
  // dx = q - x, kept when 0 <= dx <= 8.  Offsets and the keep mask depend on the lane only, not on the tile row.
  int sc_off[12];
  uint32_t sc_keep = 0u;
#pragma unroll
  for (int nt = 0; nt < 3; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = gq + 8 * (i >> 1);
      const int dx = nt * 8 + 2 * tq + (i & 1) - x;
      sc_off[nt * 4 + i] = (dy * 9 + dx) * OUT_PM + x;
      if (dx >= 0 && dx <= 8) sc_keep |= 1u << (nt * 4 + i);
    }
  // 1 / C and the activation are applied in the store loop, once per stored value: max(v, slope * v) with slope <= 1
  // (ReLU 0, LeakyReLU 0.1, none 1); the "+ 0" turns the -0 of 0 * negative into +0
  const float invC = 1.0f / (float)C;
  const float invCs = invC * (p.act == DBSR_ACT_LRELU ? 0.1f : (p.act == DBSR_ACT_RELU ? 0.0f : 1.0f));
  // two phases of MMA_ROWS tile rows: product + scatter into the staging tile, then the coalesced stores of those rows
  for (int r0 = 0; r0 < th; r0 += MMA_ROWS) {
    const int r1 = min(th, r0 + MMA_ROWS);
    for (int row = r0; row < r1; ++row) {
      uint32_t a[KS][4];
      const uint32_t* ap = f1_w + (row * CT_W + gq) * G::PW + tq;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        a[ks][0] = ap[ks * 8];
        a[ks][1] = ap[8 * G::PW + ks * 8];
        a[ks][2] = ap[ks * 8 + 4];
        a[ks][3] = ap[8 * G::PW + ks * 8 + 4];
      }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
        const uint32_t* bp = f2_w + ((row + dy) * HALO_W + nt * 8 + gq) * G::PW + tq;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t b0 = bp[ks * 8], b1r = bp[ks * 8 + 4];
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                       : "r"(a[ks][0]), "r"(a[ks][1]), "r"(a[ks][2]), "r"(a[ks][3]), "r"(b0), "r"(b1r));
        }
        const float cv[4] = {c0, c1, c2, c3};
        float* orow = out_s + (row - r0) * CT_W;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (sc_keep & (1u << (nt * 4 + i))) orow[sc_off[nt * 4 + i]] = cv[i];
      }
    }
    __syncthreads();
    if (p.vec_out) {
      for (int e = t; e < 11 * MMA_ROWS * CT_W; e += CORR_THREADS) {
        const int pix = e & (MMA_ROWS * CT_W - 1), g = e / (MMA_ROWS * CT_W);
        const int y = ty0 + r0 + pix / CT_W, x = tx0 + pix % CT_W;
        if (y < H && x < W) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float r = (g * 8 + j < 81) ? out_s[(g * 8 + j) * OUT_PM + pix] : 0.0f;
            v[j] = fmaxf(r * invC, fmaf(r, invCs, 0.0f));
          }
          const long long o = ((long long)pair * H * W + (long long)y * W + x) * p.out.c_pitch + p.out.c_off + g * 8;
          if (p.out.dtype == DBSR_BF16) {
            uint4 q;
            q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out.data) + o) = q;
          } else {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out.data) + o);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
    } else {
      for (int e = t; e < 81 * MMA_ROWS * CT_W; e += CORR_THREADS) {
        const int pix = e & (MMA_ROWS * CT_W - 1), ch = e / (MMA_ROWS * CT_W);
        const int y = ty0 + r0 + pix / CT_W, x = tx0 + pix % CT_W;
        if (y < H && x < W) {
          const float r = out_s[ch * OUT_PM + pix];
          view_st(p.out, (long long)pair * H * W + (long long)y * W + x, ch, fmaxf(r * invC, fmaf(r, invCs, 0.0f)));
        }
      }
    }
    __syncthreads();                                // the staging tile is rewritten by the next phase
  }
}

// ---------------------------------------------------------------------------------------------------------
// Small maps (H*W <= 64: pyramid levels 6..3 of the 48^2 .. 128^2 configs).  The tiled kernel above gives a 1x1 .. 8x8 map
// one 8x16 tile of which a handful of threads own real pixels (9 threads for a 1x1 map: a serial latency chain).  Here
// one CTA holds both whole maps of a pair in shared memory as [pixel][C] fp32 (the second one backwarped while it is
// staged) and spreads the H*W*81 outputs over the threads: thread -> (pixel p, displacement d), dot product over C.
// ---------------------------------------------------------------------------------------------------------
constexpr int CORR_SMALL_THREADS = 256;
template <typename T>
__global__ void __launch_bounds__(CORR_SMALL_THREADS) corr81_small_kernel(const CorrParams p) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  extern __shared__ __align__(16) float smem[];
  const int H = p.f1.h, W = p.f1.w, C = p.f1.c, HW = H * W;
  const int Cp = ((C + 7) & ~7) + 4;                 // row pitch in floats: 16-byte aligned, rows 4 banks apart
  float* f1_s = smem;
  float* f2_s = smem + HW * Cp;
  const int t = threadIdx.x, pair = blockIdx.x;
  int i1 = pair, i2 = pair;
  if (p.group > 0) {
    const int b = pair / p.group;
    i1 = b * (p.group + 1);
    i2 = i1 + 1 + (pair - b * p.group);
  }
  const long long base1 = (long long)i1 * HW, base2 = (long long)i2 * HW, basef = (long long)pair * HW;
  const bool warp2 = p.flow.data != nullptr;
  const float sxw = warp2 ? p.flow_scale * (float)W / (float)(W - 1) : 0.0f;
  const float syh = warp2 ? p.flow_scale * (float)H / (float)(H - 1) : 0.0f;
  const T* b1 = reinterpret_cast<const T*>(p.f1.data) + p.f1.c_off;
  const T* b2 = reinterpret_cast<const T*>(p.f2.data) + p.f2.c_off;
  const int G = (C + 7) >> 3;
  for (int e = t; e < HW * G; e += CORR_SMALL_THREADS) {
    const int g = e % G, pix = e / G;
    const int ch = g * 8;
    const Vec8c a1 = vec8_ld<T>(b1 + (base1 + pix) * p.f1.c_pitch + ch, C - ch);
    vec8_sts(&f1_s[pix * Cp + ch], a1);
    if (p.f1_copy.data)
      vec8_stg<T>(reinterpret_cast<T*>(p.f1_copy.data) + p.f1_copy.c_off + (basef + pix) * p.f1_copy.c_pitch + ch, a1, C - ch);
    Vec8c v = vec8_zero();
    if (!warp2) {
      v = vec8_ld<T>(b2 + (base2 + pix) * p.f2.c_pitch + ch, C - ch);
    } else {
      const int y = pix / W, x = pix - y * W;
      const float u = (float)x + view_ld(p.flow, basef + pix, 0) * sxw;
      const float w = (float)y + view_ld(p.flow, basef + pix, 1) * syh;
      const float fu = floorf(u), fv = floorf(w);
      const float ax = u - fu, ay = w - fv;
      const int xa = (int)fu, ya = (int)fv;
      float m = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xx = xa + (k & 1), yy = ya + (k >> 1);
        const float wt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
        if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
          const Vec8c a = vec8_ld<T>(b2 + (base2 + (long long)yy * W + xx) * p.f2.c_pitch + ch, C - ch);
#pragma unroll
          for (int i = 0; i < 8; ++i) v.v[i] = fmaf(a.v[i], wt, v.v[i]);
          m += wt;
        }
      }
      if (!(m > 0.999f)) v = vec8_zero();  // pwcnet.py:34-38
    }
    vec8_sts(&f2_s[pix * Cp + ch], v);
  }
  __syncthreads();
  const int C8 = G * 8;
  const float invC = 1.0f / (float)C;
  (void)invC;
  for (int e = t; e < HW * 81; e += CORR_SMALL_THREADS) {
    const int pix = e / 81, d = e - pix * 81;
    const int y = pix / W, x = pix - y * W;
    const int yy = y + d / 9 - 4, xx = x + d % 9 - 4;
    float acc = 0.0f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const float4* a = reinterpret_cast<const float4*>(&f1_s[pix * Cp]);
      const float4* b = reinterpret_cast<const float4*>(&f2_s[(yy * W + xx) * Cp]);
      float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
      for (int c4 = 0; c4 < C8 / 4; ++c4) {
        const float4 u = a[c4], v = b[c4];
        s0 = fmaf(u.x, v.x, s0); s1 = fmaf(u.y, v.y, s1); s2 = fmaf(u.z, v.z, s2); s3 = fmaf(u.w, v.w, s3);
      }
      acc = (s0 + s1) + (s2 + s3);
    }
    view_st(p.out, basef + pix, d, apply_act(acc / (float)C, p.act));
  }
}

}  // namespace dbsr

using namespace dbsr;

extern "C" int dbsr_corr81(const dbsr_nhwc_t* f1, const dbsr_nhwc_t* f2, const dbsr_nhwc_t* flow, float flow_scale,
                           const dbsr_nhwc_t* out, int32_t pairs, int32_t group, int32_t act, int32_t mode, void* stream) {
  return dbsr_corr81_copy(f1, f2, flow, flow_scale, out, nullptr, pairs, group, act, mode, stream);
}

extern "C" int dbsr_corr81_copy(const dbsr_nhwc_t* f1, const dbsr_nhwc_t* f2, const dbsr_nhwc_t* flow, float flow_scale,
                                const dbsr_nhwc_t* out, const dbsr_nhwc_t* f1_copy, int32_t pairs, int32_t group, int32_t act,
                                int32_t mode, void* stream) {
  DBSR_REQUIRE(view_ok(f1) && view_ok(f2) && view_ok(out), "corr81: bad views");
  DBSR_REQUIRE(f1->h == f2->h && f1->w == f2->w && f1->c == f2->c && out->h == f1->h && out->w == f1->w &&
                   out->c == 81 && out->n == pairs, "corr81: geometry mismatch");
  const bool has_flow = flow && flow->data;
  if (has_flow) {
    DBSR_REQUIRE(view_ok(flow) && flow->n == pairs && flow->h == f1->h && flow->w == f1->w && flow->c == 2,
                 "corr81: flow geometry mismatch");
    DBSR_REQUIRE(f1->h > 1 && f1->w > 1, "corr81: backwarp needs maps larger than 1x1 (reference divides by W-1)");
  }
  if (group > 0)
    DBSR_REQUIRE(pairs % group == 0 && f1->n >= (pairs / group) * (group + 1) && f2->n >= (pairs / group) * (group + 1),
                 "corr81: pair->image mapping out of range");
  else
    DBSR_REQUIRE(f1->n >= pairs && f2->n >= pairs, "corr81: not enough images");
  CorrParams p;
  p.f1 = make_view(f1); p.f2 = make_view(f2); p.flow = make_view(has_flow ? flow : nullptr); p.out = make_view(out);
  p.flow_scale = flow_scale; p.group = group; p.act = act;
  const bool want_copy = f1_copy && f1_copy->data;
  p.f1_copy = make_view(nullptr);
  if (want_copy) {
    DBSR_REQUIRE(view_ok(f1_copy) && f1_copy->n == pairs && f1_copy->h == f1->h && f1_copy->w == f1->w && f1_copy->c == f1->c &&
                     f1_copy->dtype == f1->dtype, "corr81: f1_copy must be a [pairs, h, w, C] view of the dtype of f1");
  }
  // the kernels that stage whole 8-channel groups write the copy themselves; every other path launches dbsr_copy_channels
  const int copy_es = f1->dtype == DBSR_BF16 ? 2 : 4;
  const bool copy_vec = want_copy && f1_copy->c_off % 8 == 0 && f1_copy->c_pitch % 8 == 0 &&
                        ((uintptr_t)f1_copy->data + (size_t)f1_copy->c_off * copy_es) % 16 == 0 &&
                        ((size_t)f1_copy->c_pitch * copy_es) % 16 == 0;
  auto copy_separately = [&]() -> int {
    return want_copy ? dbsr_copy_channels(f1, f1_copy, group, group > 0 ? group + 1 : 0, 0, stream) : 0;
  };
  p.tiles_x = ceil_div(f1->w, CT_W);
  const int out_es = out->dtype == DBSR_BF16 ? 2 : 4;
  p.vec_out = out->c_off % 8 == 0 && out->c_pitch % 8 == 0 && out->c_off + 88 <= out->c_pitch &&
              ((uintptr_t)out->data + (size_t)out->c_off * out_es) % 16 == 0;
  dim3 grid(p.tiles_x * ceil_div(f1->h, CT_H), pairs);
  // vectorised staging needs 8-channel groups on 16-byte (bf16) / 32-byte (fp32) boundaries in both feature maps
  auto vec_ok = [](const dbsr_nhwc_t* v) {
    return v->c_off % 8 == 0 && v->c_pitch % 8 == 0 && ((uintptr_t)v->data % 32) == 0;
  };
  const bool vec = f1->dtype == f2->dtype && vec_ok(f1) && vec_ok(f2);
  // small maps: both maps of a pair in shared memory, outputs spread over the threads
  const int cp_small = ((f1->c + 7) & ~7) + 4;
  const size_t small_smem = (size_t)2 * f1->h * f1->w * cp_small * sizeof(float);
  // bf16 maps with C in {32, 64, 96, 128}: banded product on the tensor cores (mma.sync), see corr81_mma_kernel
  const bool mma_ok = vec && mode != DBSR_CORR_CUDA_CORES && f1->dtype == DBSR_BF16 && f1->c % 16 == 0 && f1->c >= 32 && f1->c <= 128 &&
                      f1->c != 48 && f1->c != 80 && f1->c != 112 &&
                      (long long)f1->h * f1->w * (f1->c_pitch > f2->c_pitch ? f1->c_pitch : f2->c_pitch) * 2 < (1ll << 31);
  // 8x8 maps (level 3 of a 48^2 burst, level 4 of an 80^2 one) go to the tensor-core kernel when it covers them: the small-map
  // kernel is bound by its shared-memory loads there (2 LDS per FMA: 34 us for 416 pairs of 64 channels against ~14 us)
  if (vec && f1->h * f1->w <= 64 && small_smem <= 160 * 1024 && !(mma_ok && f1->h * f1->w >= 64)) {
    void (*ks)(const CorrParams) = f1->dtype == DBSR_BF16 ? corr81_small_kernel<__nv_bfloat16> : corr81_small_kernel<float>;
    static size_t configured_dev[MAX_DEVICES][2] = {};
    size_t* configured = configured_dev[current_device_slot()];
    const int si = f1->dtype == DBSR_BF16 ? 0 : 1;
    if (small_smem > 48 * 1024 && small_smem > configured[si]) {
      cudaError_t e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem);
      DBSR_REQUIRE(e == cudaSuccess, "corr81: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      configured[si] = small_smem;
    }
    if (copy_vec) p.f1_copy = make_view(f1_copy);
    else if (int rc = copy_separately()) return rc;
    launch_pdl(ks, dim3((unsigned)pairs), dim3(CORR_SMALL_THREADS), small_smem, (cudaStream_t)stream, p);
    return check_launch("corr81");
  }
  if (mma_ok) {
    void (*km)(const CorrParams) = nullptr;
    int smem = 0, ki = 0;
    switch (f1->c / 16) {
      case 2: km = corr81_mma_kernel<2>; smem = CorrMma<2>::SMEM; ki = 0; break;
      case 4: km = corr81_mma_kernel<4>; smem = CorrMma<4>::SMEM; ki = 1; break;
      case 6: km = corr81_mma_kernel<6>; smem = CorrMma<6>::SMEM; ki = 2; break;
      default: km = corr81_mma_kernel<8>; smem = CorrMma<8>::SMEM; ki = 3; break;
    }
    static bool mma_attr_dev[MAX_DEVICES][4] = {};
    bool* mma_attr = mma_attr_dev[current_device_slot()];
    if (!mma_attr[ki]) {
      cudaError_t e = cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      DBSR_REQUIRE(e == cudaSuccess, "corr81: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      e = cudaFuncSetAttribute(km, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      DBSR_REQUIRE(e == cudaSuccess, "corr81: carve-out attribute failed: %s", cudaGetErrorString(e));
      mma_attr[ki] = true;
    }
    if (copy_vec) p.f1_copy = make_view(f1_copy);
    else if (int rc = copy_separately()) return rc;
    launch_pdl(km, grid, dim3(CORR_THREADS), (size_t)smem, (cudaStream_t)stream, p);
    return check_launch("corr81");
  }
  void (*kern)(const CorrParams) = !vec ? corr81_kernel<false, float>
                                   : (f1->dtype == DBSR_BF16 ? corr81_kernel<true, __nv_bfloat16> : corr81_kernel<true, float>);
  static bool attr_set_dev[MAX_DEVICES][3] = {};
  bool* attr_set = attr_set_dev[current_device_slot()];
  const int ki = !vec ? 0 : (f1->dtype == DBSR_BF16 ? 1 : 2);
  if (!attr_set[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CORR_SMEM);
    DBSR_REQUIRE(e == cudaSuccess, "corr81: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    // ncu: with the default carve-out the driver sized shared memory for ONE 64 KB CTA per SM (9 warps, 69 % of the
    // cycles without an eligible warp); ask for the maximum so that two CTAs overlap staging and accumulation
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    DBSR_REQUIRE(e == cudaSuccess, "corr81: carve-out attribute failed: %s", cudaGetErrorString(e));
    attr_set[ki] = true;
  }
  if (int rc = copy_separately()) return rc;
  launch_pdl(kern, grid, dim3(CORR_THREADS), (size_t)CORR_SMEM, (cudaStream_t)stream, p);
  return check_launch("corr81");
}
