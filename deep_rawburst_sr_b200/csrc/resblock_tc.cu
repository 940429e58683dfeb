// Fused residual block of the decoder's high-resolution stage on tcgen05 / TMEM (sm_100a):
//     y = relu(x + conv2(relu(conv1(x) + b1)) + b2),   3x3 / 32 -> 32 channels, bf16 NHWC, fp32 accumulation
// replaces the two conv launches of `ResBlock.forward` (reference models/layers/blocks.py:84-96) for the four
// `post_res_layers` of `ResPixShuffleConv` (models/dbsr/decoders.py:45-50, 59), which run at 8H x 8W on 32 channels: there
// the unfused kernels are bound by HBM / the epilogue (each block moves the 302 MB map of 32 bursts four times through HBM).
// Here the intermediate map never leaves the SM, so a block reads its input once and writes its output once.
//
//   item      : 16 x 24 output pixels of one image.
//   X box     : ONE TMA box {32 ch, 28 px, 20 rows} = the tile with a 2-pixel halo, zero-filled outside the image (SWIZZLE_64B,
//               64 bytes per pixel).  Box position p = by * 28 + bx.
//   conv1     : "flat" implicit GEMM over box positions: M row m of tile t is position 29 + 128 t + m (4 tiles cover the 18 x 26
//               positions conv2 needs, plus the two wrap-around columns per row and a few rows past the end -- never read).
//               Tap (ky, kx) is the same box shifted by (ky - 1) * 28 + (kx - 1) positions: 9 taps x 2 K-steps x 4 tiles.
//   T buffer  : epilogue 1 (warps 0-3) reads the accumulators, adds b1, applies ReLU, ZEROES every position outside the image
//               (conv2's zero padding applies to the intermediate map, not to conv1 of the zero-padded input), rounds to bf16
//               and writes the rows straight into shared memory in the K-major SWIZZLE_64B layout conv2's MMA descriptors read.
//   conv2     : three 16 x 8-pixel patch tiles (8-row groups 28 positions apart) over the T buffer, 9 taps x 2 K-steps each,
//               plus the residual as 2 K-steps against an identity weight tile read from the X box (exact: fp32 accumulation).
//   epilogue 2: warps 4-7: + b2, ReLU, bf16 -> per-warp staging rows -> one TMA tensor store per (warp, tile); or the fused
//               1x1 predictor + ReLU (+ 14-bit quantisation) of the last block (decoders.py:52, 61), as in conv_tc.cu.
//   pipeline  : the single MMA-issuing thread alternates conv1(i + 1), conv2(i) so that the tensor pipe works on the next item
//               while epilogue 1 of item i fills the T buffer; X boxes 3 deep, T buffers and accumulators 2 deep.
//   weights   : both 3x3 kernels + the identity tile (19 x 2 KB) stay resident in shared memory.
#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"

#include <stdlib.h>
#include <string.h>

namespace dbsr {

constexpr int RB_OH = 16, RB_OW = 24;                   // output tile
constexpr int RB_BH = RB_OH + 4, RB_BW = RB_OW + 4;     // X box: 20 rows x 28 pixels
constexpr int RB_P0 = RB_BW + 1;                        // first conv1 position: box row 1, column 1
constexpr int RB_T1 = 4, RB_T2 = 3;                     // M tiles of conv1 (flat) / conv2 (16 x 8 patches)
constexpr int RB_ROW = 64;                              // bytes per position (32 bf16)
constexpr int RB_X_TX = RB_BH * RB_BW * RB_ROW;         // 35 840 bytes per box
constexpr int RB_X_BYTES = 36864;                       // 576 positions: conv1 reads up to position 569 (masked rows only)
constexpr int RB_T_BYTES = RB_T1 * 128 * RB_ROW;        // 32 768: positions 29 .. 540
constexpr int RB_W_TILE = 32 * RB_ROW;                  // one [32 x 32] bf16 weight tile
constexpr int RB_NW = 19;                               // 9 (conv1) + 9 (conv2) + identity
constexpr int RB_XS = 3, RB_TS = 2;
constexpr int RB_STG = 2048;                            // staging of one epilogue-2 warp: 32 pixels x 64 bytes
constexpr int RB_ACC_COLS = (RB_T1 + RB_T2) * 32;       // TMEM columns per accumulator stage: D1 128 | D2 96
constexpr int RB_THREADS = 384;
constexpr int RB_WARP_PROD = 10, RB_WARP_MMA = 11;
constexpr int RB_SMEM = RB_XS * RB_X_BYTES + RB_TS * RB_T_BYTES + RB_NW * RB_W_TILE + 4 * RB_STG + 1024 /* barriers, bias */ +
                        1024 /* alignment slack */;

struct RbParams {
  int n, H, W;
  int tiles_x, tiles_y;
  int total_items;
  int early_trigger;  // launch_dependents at the top: only when the grid leaves at least half of the SMs idle
  int static_w;     // DBSR_CONV_STATIC_WEIGHTS: weights and biases may be fetched before griddepcontrol.wait
  const float* b1;
  const float* b2;
  // fused 1x1 predictor (last block): y is NOT written; pred[n, k, y, x] = relu(pb[k] + sum_c pw[k][c] * block(x)[c])
  void* pred; int pred_c; int pred_q14;
  float pred_wb[4 * 32 + 4];
};

__device__ __forceinline__ void rb_decode(const RbParams& p, int item, int& img, int& y0, int& x0) {
  const int per = p.tiles_x * p.tiles_y;
  img = item / per;
  const int rem = item - img * per;
  const int ty = rem / p.tiles_x;
  y0 = ty * RB_OH;
  x0 = (rem - ty * p.tiles_x) * RB_OW;
}

__global__ void __launch_bounds__(RB_THREADS, 1)
resblock32_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                     const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_i,
                     const __grid_constant__ CUtensorMap tmap_y, const RbParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_x = smem;                                              // [3][36 864]
  uint8_t* smem_t = smem_x + RB_XS * RB_X_BYTES;                       // [2][32 768]
  uint8_t* smem_w = smem_t + RB_TS * RB_T_BYTES;                       // [19][2 048]
  uint8_t* smem_stg = smem_w + RB_NW * RB_W_TILE;                      // [4][2 048]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + 4 * RB_STG);
  uint64_t* x_full = bars;                 // [3]
  uint64_t* x_empty = x_full + RB_XS;      // [3]
  uint64_t* w_full = x_empty + RB_XS;      // [1]
  uint64_t* d1_full = w_full + 1;          // [2]
  uint64_t* d1_empty = d1_full + 2;        // [2]
  uint64_t* t_full = d1_empty + 2;         // [2]
  uint64_t* t_empty = t_full + 2;          // [2]
  uint64_t* d2_full = t_empty + 2;         // [2]
  uint64_t* d2_empty = d2_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 2);
  float* bias_tab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);      // [2][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RB_XS; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&d1_full[s], 1); mbar_init(&d1_empty[s], 4);
      mbar_init(&t_full[s], 4); mbar_init(&t_empty[s], 1);
      mbar_init(&d2_full[s], 1); mbar_init(&d2_empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_w1); tma_prefetch_desc(&tmap_w2); tma_prefetch_desc(&tmap_i);
    if (p.pred == nullptr) tma_prefetch_desc(&tmap_y);
  }
  if (warp == RB_WARP_MMA) tmem_alloc(tmem_slot, 512u);
  // programmatic dependent launch: everything above touches on-chip state only.  With static weights the 19 resident weight
  // tiles and the biases are fetched before the wait (while the preceding kernel drains); X loads and stores come after it.
  if (p.early_trigger) griddep_launch_dependents();
  if (!p.static_w) griddep_wait();
  if (threadIdx.x < 64) bias_tab[threadIdx.x] = __ldg((threadIdx.x < 32 ? p.b1 : p.b2 - 32) + threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.static_w && warp != RB_WARP_PROD && warp != RB_WARP_MMA) griddep_wait();
  const uint32_t tmem_base = *tmem_slot;
  const int n_local = (p.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == RB_WARP_PROD) {
    // ===================== producer: the resident weights once, then one X box per item =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, (uint32_t)(RB_NW * RB_W_TILE));
      for (int t = 0; t < 9; ++t) tma_load_2d(&tmap_w1, w_full, smem_w + t * RB_W_TILE, 0, t * 32);
      for (int t = 0; t < 9; ++t) tma_load_2d(&tmap_w2, w_full, smem_w + (9 + t) * RB_W_TILE, 0, t * 32);
      tma_load_2d(&tmap_i, w_full, smem_w + 18 * RB_W_TILE, 0, 0);
      if (p.static_w) griddep_wait();
      int item = blockIdx.x;
      for (int k = 0; k < n_local; ++k, item += gridDim.x) {
        int img, y0, x0;
        rb_decode(p, item, img, y0, x0);
        const int xs = k % RB_XS;
        mbar_wait(&x_empty[xs], (uint32_t)(((k / RB_XS) & 1) ^ 1), 100 + xs);
        mbar_arrive_expect_tx(&x_full[xs], (uint32_t)RB_X_TX);
        tma_load_4d(&tmap_x, &x_full[xs], smem_x + xs * RB_X_BYTES, 0, x0 - 2, y0 - 2, img);
      }
    }
  } else if (warp == RB_WARP_MMA) {
    // ===================== MMA issuer: conv1(k), then conv2(k - 1) =====================
    if (elect_one()) {
      // cute::UMMA::InstrDescriptor: c_format F32 | a, b BF16 | K-major | N = 32 | M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      // descriptor high words: SBO | version 1 | SWIZZLE_64B.  flat: dense 8-row groups; patch: 8-pixel groups one box row apart
      const uint32_t hi_flat = ((512u >> 4) & 0x3FFFu) | (1u << 14) | (4u << 29);
      const uint32_t hi_patch = (((uint32_t)(RB_BW * RB_ROW) >> 4) & 0x3FFFu) | (1u << 14) | (4u << 29);
      const uint32_t x_lo0 = ((smem_u32(smem_x) >> 4) & 0x3FFFu) | 0x10000u;
      const uint32_t t_lo0 = ((smem_u32(smem_t) >> 4) & 0x3FFFu) | 0x10000u;
      const uint32_t w_lo0 = ((smem_u32(smem_w) >> 4) & 0x3FFFu) | 0x10000u;
      uint32_t tap_off[9];                      // 16-byte units: (ky * 28 + kx) positions of 64 bytes
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)(((tap / 3) * RB_BW + (tap % 3)) * (RB_ROW / 16));
      mbar_wait(w_full, 0, 350);
      tc_fence_after();
      for (int k = 0; k <= n_local; ++k) {
        if (k < n_local) {
          const int s = k & 1, xs = k % RB_XS;
          mbar_wait(&x_full[xs], (uint32_t)((k / RB_XS) & 1), 300 + xs);
          mbar_wait(&d1_empty[s], (uint32_t)(((k >> 1) & 1) ^ 1), 200 + s);
          tc_fence_after();
          const uint32_t xl = x_lo0 + (uint32_t)(xs * (RB_X_BYTES >> 4));
          const uint32_t d1 = tmem_base + (uint32_t)(s * RB_ACC_COLS);
#pragma unroll 1
          for (int t = 0; t < RB_T1; ++t) {
            const uint32_t al = xl + (uint32_t)(t * 128 * (RB_ROW / 16));
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t bl = w_lo0 + (uint32_t)(tap * (RB_W_TILE >> 4));
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16)
                umma_bf16(d1 + (uint32_t)(t * 32), ((uint64_t)hi_flat << 32) | (uint64_t)(al + tap_off[tap] + 2u * k16),
                          ((uint64_t)hi_flat << 32) | (uint64_t)(bl + 2u * k16), idesc, (tap | k16) ? 1u : 0u);
            }
          }
          umma_commit(&d1_full[s]);
        }
        if (k >= 1) {
          const int j = k - 1, s = j & 1, xs = j % RB_XS;
          mbar_wait(&t_full[s], (uint32_t)((j >> 1) & 1), 320 + s);
          mbar_wait(&d2_empty[s], (uint32_t)(((j >> 1) & 1) ^ 1), 220 + s);
          tc_fence_after();
          const uint32_t tl = t_lo0 + (uint32_t)(s * (RB_T_BYTES >> 4));
          const uint32_t xl = x_lo0 + (uint32_t)(xs * (RB_X_BYTES >> 4));
          const uint32_t d2 = tmem_base + (uint32_t)(s * RB_ACC_COLS + RB_T1 * 32);
#pragma unroll 1
          for (int t = 0; t < RB_T2; ++t) {
            const uint32_t al = tl + (uint32_t)(t * 8 * (RB_ROW / 16));
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t bl = w_lo0 + (uint32_t)((9 + tap) * (RB_W_TILE >> 4));
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16)
                umma_bf16(d2 + (uint32_t)(t * 32), ((uint64_t)hi_patch << 32) | (uint64_t)(al + tap_off[tap] + 2u * k16),
                          ((uint64_t)hi_flat << 32) | (uint64_t)(bl + 2u * k16), idesc, (tap | k16) ? 1u : 0u);
            }
            // residual: D += X[centre of the tile] * I
            const uint32_t rl = xl + (uint32_t)((2 * RB_BW + 2 + t * 8) * (RB_ROW / 16));
            const uint32_t il = w_lo0 + (uint32_t)(18 * (RB_W_TILE >> 4));
#pragma unroll
            for (int k16 = 0; k16 < 2; ++k16)
              umma_bf16(d2 + (uint32_t)(t * 32), ((uint64_t)hi_patch << 32) | (uint64_t)(rl + 2u * k16),
                        ((uint64_t)hi_flat << 32) | (uint64_t)(il + 2u * k16), idesc, 1u);
          }
          umma_commit(&d2_full[s]);
          umma_commit(&x_empty[xs]);
          umma_commit(&t_empty[s]);
        }
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue 1: conv1 accumulators -> relu(. + b1), zero outside the image -> T buffer =====================
    const int q = warp;                                    // TMEM lane quarter
    const int row = q * 32 + lane;                         // M row inside a tile
    const uint32_t bias_s = smem_u32(bias_tab);
    const uint32_t swz = (uint32_t)((row >> 1) & 3);       // T stages are 1024-byte aligned: (address >> 7) & 3 of this row
    int item = blockIdx.x;
    for (int j = 0; j < n_local; ++j, item += gridDim.x) {
      int img, y0, x0;
      rb_decode(p, item, img, y0, x0);
      const int s = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      mbar_wait(&d1_full[s], ph, 400 + s);
      tc_fence_after();
      mbar_wait(&t_empty[s], ph ^ 1u, 420 + s);
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * RB_ACC_COLS);
      const uint32_t trow = smem_u32(smem_t) + (uint32_t)(s * RB_T_BYTES + row * RB_ROW);
#pragma unroll
      for (int t = 0; t < RB_T1; ++t) {
        uint32_t r[32];
        tmem_ld32(tb + (uint32_t)(t * 32), r);
        const int pos = RB_P0 + t * 128 + row;
        const int by = (pos * 2341) >> 16;                 // pos / 28, exact for pos < 1024 (checked on the host)
        const int bx = pos - by * RB_BW;
        const int gy = y0 - 2 + by, gx = x0 - 2 + bx;
        const bool inside = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 b4 = lds128f(bias_s + (uint32_t)c * 4u);
          v[c] = inside ? fmaxf(__uint_as_float(r[c]) + b4.x, 0.0f) : 0.0f;
          v[c + 1] = inside ? fmaxf(__uint_as_float(r[c + 1]) + b4.y, 0.0f) : 0.0f;
          v[c + 2] = inside ? fmaxf(__uint_as_float(r[c + 2]) + b4.z, 0.0f) : 0.0f;
          v[c + 3] = inside ? fmaxf(__uint_as_float(r[c + 3]) + b4.w, 0.0f) : 0.0f;
        }
        const uint32_t dst = trow + (uint32_t)(t * 128 * RB_ROW);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(dst + ((((uint32_t)c) ^ swz) << 4), pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                 pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
      }
      fence_async_smem();          // the rows are read by tcgen05.mma through the async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&t_full[s]); mbar_arrive(&d1_empty[s]); }
    }
  } else if (warp < 8) {
    // ===================== epilogue 2: conv2 (+ residual) accumulators -> relu(. + b2) -> y (TMA store) or predictor =====================
    const int q = warp - 4;
    const uint32_t bias_s = smem_u32(bias_tab) + 128u;
    const uint32_t stg_s = smem_u32(smem_stg) + (uint32_t)(q * RB_STG);
    const uint32_t row_s = stg_s + (uint32_t)(lane * RB_ROW);
    const uint32_t swz = ((stg_s >> 7) + (uint32_t)(lane >> 1)) & 3u;
    int item = blockIdx.x;
    for (int j = 0; j < n_local; ++j, item += gridDim.x) {
      int img, y0, x0;
      rb_decode(p, item, img, y0, x0);
      const int s = j & 1;
      mbar_wait(&d2_full[s], (uint32_t)((j >> 1) & 1), 440 + s);
      tc_fence_after();
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * RB_ACC_COLS + RB_T1 * 32);
#pragma unroll 1
      for (int t = 0; t < RB_T2; ++t) {
        uint32_t r[32];
        tmem_ld32(tb + (uint32_t)(t * 32), r);
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 b4 = lds128f(bias_s + (uint32_t)c * 4u);
          v[c] = fmaxf(__uint_as_float(r[c]) + b4.x, 0.0f); v[c + 1] = fmaxf(__uint_as_float(r[c + 1]) + b4.y, 0.0f);
          v[c + 2] = fmaxf(__uint_as_float(r[c + 2]) + b4.z, 0.0f); v[c + 3] = fmaxf(__uint_as_float(r[c + 3]) + b4.w, 0.0f);
        }
        if (p.pred != nullptr) {
          const int y = y0 + q * 4 + (lane >> 3), x = x0 + t * 8 + (lane & 7);
          if (y < p.H && x < p.W) {
            const long long plane = (long long)p.H * p.W;
            const long long o = (long long)img * p.pred_c * plane + (long long)y * p.W + x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < p.pred_c) {
                float s0 = p.pred_wb[128 + k], s1 = 0.0f;
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                  s0 = fmaf(v[c], p.pred_wb[k * 32 + c], s0);
                  s1 = fmaf(v[c + 1], p.pred_wb[k * 32 + c + 1], s1);
                }
                const float out = fmaxf(s0 + s1, 0.0f);
                if (p.pred_q14) reinterpret_cast<short*>(p.pred)[o + k * plane] = (short)(fminf(out, 1.0f) * 16384.0f);
                else reinterpret_cast<float*>(p.pred)[o + k * plane] = out;
              }
            }
          }
        } else {
          if (lane == 0) tma_store_wait_read();      // the previous tensor store has finished reading the staging rows
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts128(row_s + ((((uint32_t)c) ^ swz) << 4), pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                   pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
          fence_async_smem();
          __syncwarp();
          // the staging rows are the box {32 channels, 8 pixels, 4 rows} in the SWIZZLE_64B layout; clipped at the image border
          if (lane == 0 && x0 + t * 8 < p.W && y0 + q * 4 < p.H) tma_store_4d(&tmap_y, stg_s, 0, x0 + t * 8, y0 + q * 4, img);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d2_empty[s]);
    }
    if (p.pred == nullptr && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == RB_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace dbsr

using namespace dbsr;

static int rb_check(const dbsr_resblock_t* b, bool set_err) {
#define RB_REQ(cond, ...) do { if (!(cond)) { if (set_err) set_error(__VA_ARGS__); return 1; } } while (0)
  RB_REQ(b && view_ok(&b->x) && b->w1 && b->w2 && b->b1 && b->b2, "resblock32_tc: bad descriptor");
  RB_REQ(b->x.dtype == DBSR_BF16 && b->x.c == 32 && b->x.c_off % 8 == 0 && b->x.c_pitch % 8 == 0 && ((uintptr_t)b->x.data % 16) == 0,
         "resblock32_tc: x must be a 16-byte aligned bf16 view of 32 channels");
  RB_REQ(((uintptr_t)b->w1 % 16) == 0 && ((uintptr_t)b->w2 % 16) == 0 && ((uintptr_t)b->b1 % 16) == 0 && ((uintptr_t)b->b2 % 16) == 0,
         "resblock32_tc: weights / biases must be 16-byte aligned");
  if (b->pred) {
    RB_REQ(b->pred_w && b->pred_b && b->pred_c >= 1 && b->pred_c <= 4, "resblock32_tc: predictor needs 1..4 output channels and host weights");
  } else {
    RB_REQ(view_ok(&b->y) && b->y.dtype == DBSR_BF16 && b->y.c == 32 && b->y.c_off % 8 == 0 && b->y.c_pitch % 8 == 0 &&
               ((uintptr_t)b->y.data % 16) == 0 && b->y.n == b->x.n && b->y.h == b->x.h && b->y.w == b->x.w,
           "resblock32_tc: y must be a 16-byte aligned bf16 view with the geometry of x");
    RB_REQ(b->y.data != b->x.data, "resblock32_tc: y must not alias x (neighbouring tiles read x halos while y is written)");
  }
  RB_REQ((long long)b->x.n * ceil_div(b->x.h, RB_OH) * ceil_div(b->x.w, RB_OW) < (1LL << 31), "resblock32_tc: too many items");
  return 0;
#undef RB_REQ
}

extern "C" int dbsr_resblock32_tc_supported(const dbsr_resblock_t* b) { return rb_check(b, false) == 0 ? 1 : 0; }

extern "C" int dbsr_resblock32_tc(const dbsr_resblock_t* b, void* stream) {
  if (rb_check(b, true)) return 1;
  for (int pos = 0; pos < 1024; ++pos) DBSR_REQUIRE(((pos * 2341) >> 16) == pos / RB_BW, "resblock32_tc: internal: inexact division");
  EncodeTiledFn encode = get_encode();
  DBSR_REQUIRE(encode != nullptr, "resblock32_tc: cuTensorMapEncodeTiled entry point not available");
  cudaStream_t st = (cudaStream_t)stream;
  alignas(64) CUtensorMap mx, mw1, mw2, mi, my;
  memset(&my, 0, sizeof(my));
  {
    const dbsr_nhwc_t& v = b->x;
    cuuint64_t dims[4] = {32, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
    cuuint64_t strides[3] = {(cuuint64_t)v.c_pitch * 2, (cuuint64_t)v.w * v.c_pitch * 2, (cuuint64_t)v.h * v.w * v.c_pitch * 2};
    cuuint32_t box[4] = {32u, (cuuint32_t)RB_BW, (cuuint32_t)RB_BH, 1u};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* base = reinterpret_cast<__nv_bfloat16*>(v.data) + v.c_off;
    CUresult rc = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DBSR_REQUIRE(rc == CUDA_SUCCESS, "resblock32_tc: cuTensorMapEncodeTiled(x) failed with %d", (int)rc);
  }
  const void* eye = identity_weights(32, st);
  DBSR_REQUIRE(eye != nullptr, "resblock32_tc: could not create the identity weight tile (first call inside a stream capture? run once eagerly)");
  {
    const void* ptrs[3] = {b->w1, b->w2, eye};
    CUtensorMap* maps[3] = {&mw1, &mw2, &mi};
    for (int i = 0; i < 3; ++i) {
      cuuint64_t dims[2] = {32, (cuuint64_t)(i < 2 ? 9 * 32 : 32)};
      cuuint64_t strides[1] = {64};
      cuuint32_t box[2] = {32u, 32u};
      cuuint32_t es[2] = {1, 1};
      CUresult rc = encode(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptrs[i]), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      DBSR_REQUIRE(rc == CUDA_SUCCESS, "resblock32_tc: cuTensorMapEncodeTiled(w%d) failed with %d", i, (int)rc);
    }
  }
  if (!b->pred) {
    const dbsr_nhwc_t& v = b->y;
    cuuint64_t dims[4] = {32, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
    cuuint64_t strides[3] = {(cuuint64_t)v.c_pitch * 2, (cuuint64_t)v.w * v.c_pitch * 2, (cuuint64_t)v.h * v.w * v.c_pitch * 2};
    cuuint32_t box[4] = {32u, 8u, 4u, 1u};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* base = reinterpret_cast<__nv_bfloat16*>(v.data) + v.c_off;
    CUresult rc = encode(&my, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DBSR_REQUIRE(rc == CUDA_SUCCESS, "resblock32_tc: cuTensorMapEncodeTiled(y) failed with %d", (int)rc);
  }
  RbParams p;
  p.n = b->x.n; p.H = b->x.h; p.W = b->x.w;
  p.tiles_x = ceil_div(p.W, RB_OW); p.tiles_y = ceil_div(p.H, RB_OH);
  p.total_items = p.n * p.tiles_x * p.tiles_y;
  p.static_w = (b->flags & DBSR_CONV_STATIC_WEIGHTS) ? 1 : 0;
  p.b1 = b->b1; p.b2 = b->b2;
  p.pred = b->pred; p.pred_c = b->pred_c; p.pred_q14 = b->pred_q14 ? 1 : 0;
  memset(p.pred_wb, 0, sizeof(p.pred_wb));
  if (b->pred) {
    for (int k = 0; k < b->pred_c; ++k) {
      for (int c = 0; c < 32; ++c) p.pred_wb[k * 32 + c] = b->pred_w[k * 32 + c];      // HOST arrays
      p.pred_wb[128 + k] = b->pred_b[k];
    }
  }
  const int dev_slot = current_device_slot();
  static bool attr_dev[MAX_DEVICES] = {};
  static int sms_dev[MAX_DEVICES] = {};
  if (!attr_dev[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(resblock32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM);
    DBSR_REQUIRE(e == cudaSuccess, "resblock32_tc: cudaFuncSetAttribute(%d) failed: %s", RB_SMEM, cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms_dev[dev_slot], cudaDevAttrMultiProcessorCount, dev);
    attr_dev[dev_slot] = true;
  }
  int grid = p.total_items < sms_dev[dev_slot] ? p.total_items : sms_dev[dev_slot];
  if (b->grid_limit > 0 && grid > b->grid_limit) grid = b->grid_limit;
  p.early_trigger = grid * 2 <= sms_dev[dev_slot];
  launch_pdl(resblock32_tc_kernel, dim3((unsigned)grid), dim3(RB_THREADS), (size_t)RB_SMEM, st, mx, mw1, mw2, mi, my, p);
  return check_launch("resblock32_tc");
}
