// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
// Replaces the cuDNN fp32 convs + separate ReLU / LeakyReLU / residual-add / concat kernels of the DBSR encoder, fusion
// weight predictor, decoder and of PWC-Net (reference models/layers/blocks.py:46-96, models/dbsr/*.py,
// models/alignment/pwcnet.py:49-204).
//
// GEMM view:  D[pixels, Cout] = sum_{tap, cin} X[pixel + tap, cin] * W[tap, cout, cin]
//   work item: 16 x (8*mt) output pixels of one image (mt = 1 or 2 side-by-side 16x8 tiles, one M=128 accumulator
//              each, sharing every weight tile) x one N tile (n_tile = any multiple of 16 up to 128 = TMEM columns).
//   A operand: per K chunk (CK = 64 channels -> SWIZZLE_128B rows, 32 -> SWIZZLE_64B) ONE TMA box
//              {CK ch, 8*mt + 2*dil px, 16 + 2*dil rows} of the NHWC input, i.e. the tile plus its halo.  Coordinates
//              outside the image are zero-filled by TMA, which IS the conv zero padding.  All 9 taps read this box in
//              place: the UMMA descriptor of tap (ky, kx) starts (ky*dil*halo_w + kx*dil) pixel rows (128/64 B each)
//              into the box with SBO = halo_w pixel rows.  (Swizzling is a function of the absolute smem address, so
//              a 128 B-granular start inside a TMA-written box is consistent -- verified on B200.)
//   B operand: [n_tile x CK] weight tile per (chunk, tap).  Streamed through its own mbarrier ring, or -- when all
//              taps x chunks fit next to the A ring (e.g. 64->64: 72 KB) -- loaded once per CTA and kept resident.
//   roles    : warp 0 = A (activation) TMA producer, warp 2 = B (weight) TMA producer, warp 1 = TMEM allocator +
//              single-thread tcgen05.mma issuer, warps 3..6 = epilogue (tcgen05.ld -> +bias, +residual, activation ->
//              bf16 / fp32 stores; masked generic path for odd channel counts).  TMEM accumulators are double
//              buffered so the epilogue of item i overlaps the MMAs of item i+1.
//   persistent: grid = min(#items, #SMs); static round-robin over items, N tile fastest.
//   Pixel-shuffle (upsampling.py:57) is folded into the store addressing: the packed weight rows are permuted to
//   (i, j, c) order so an N tile of 128 channels is 4 adjacent HR pixels x 32 channels, contiguous.
#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

namespace dbsr {

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
constexpr int TILE_H = 16, TILE_W = 8;
constexpr int TC_THREADS = 384;   // warps 0-7: epilogue; warp 8 idle; warp 9: A producer; warp 10: B producer; warp 11: TMEM
                                  // allocator + tcgen05.mma issuer (highest warp id: the SM arbiter favours higher ids and
                                  // the issuer must not be starved by warps that are waiting on mbarriers)
constexpr int WARP_A = 9, WARP_B = 10, WARP_MMA = 11;

struct ConvTcParams {
  int n, H, W;          // input == output spatial size (stride 1, "same" padding)
  int ksize, dil;
  int nchunks;          // Kpad / CK
  int cout_pad;         // rows per tap in the packed weight matrix (= ntiles_n * n_tile)
  int cout;             // real output channels (stores are masked beyond it)
  int n_tile;           // UMMA N (multiple of 16, <= 128)
  int mt;               // 16x8 tiles per item (1 or 2, side by side along x)
  int tmem_cols;        // power of two >= max(32, acc_stages * mt * n_tile)
  int acc_stages;       // accumulator stages in TMEM (2 or 4): how many items the MMA issuer may run ahead of the epilogue
  int vec_ok;           // 1: aligned fast path (16-byte accesses, every 32-column chunk fully inside cout)
  int align_ok;         // 1: y / residual / bias allow 16-byte accesses (chunks that end inside cout use them per thread)
  int ntiles_n;         // cout_pad / n_tile
  int items_x, tiles_y;
  long long total_items;
  // exact division by multiply-high for item -> (N tile, image, tile row, tile column): q = umulhi(n, mul) + (n & one)
  // (host proves exactness for every n < total_items, else fastdiv = 0 and the kernel divides)
  unsigned md_nt, md_img, md_x, one_nt, one_img, one_x; int fastdiv;
  int bias_smem;        // the lean epilogue may run: its bias table (cout_pad floats) fits shared memory, or there is no bias at all
  int bias_fill;        // the table fits and is filled (zeros without a bias)
  int a_slots, b_stages, b_resident;
  int a_bytes, a_tx_bytes, b_bytes;
  int early_trigger;    // griddepcontrol.launch_dependents at the top (the launch leaves at least half of the SMs idle)
  int static_w;         // DBSR_CONV_STATIC_WEIGHTS: weights / bias may be fetched before griddepcontrol.wait
  int b_taps;           // streamed weights: taps of one (chunk, N tile) that share a ring stage, ONE TMA box and one barrier pair (1, 3 or 9)
  int b_stage_bytes;    // b_taps * b_bytes (resident: b_bytes)
  int halo_w;           // pixels per row of the A box
  int res_chunks;       // > 0: the residual is accumulated on the tensor core as res_chunks extra K chunks (identity weights)
  int r_tx_bytes;       // bytes of one residual box {CK, 8*mt px, 16 rows}
  int res_group;        // >= 1: output image i takes residual image i / res_group (a per-burst map broadcast over the frames)
  int flat;             // small-map mode: the A box holds flat_ni whole zero-bordered images, M rows = flat slots
  int flat_s, flat_ni;  // slots per image (H+2)*(W+2); images per M tile
  // second M tile of an item (mt == 2): offset of its A rows inside the box (16-byte units), image and column offset.
  //   wide maps : side by side in x        -> t1_step16 = 8 pixel rows, t1_dimg = 0,       t1_dx = 8
  //   narrow / flat maps: the NEXT image(s) -> t1_step16 = one image box,  t1_dimg = 1 / ni, t1_dx = 0
  int t1_step16, t1_dimg, t1_dx;
  int imgs_per_item;    // images one item (= one A box) covers: 1, 2 (narrow pair), flat_ni * mt
  // output
  void* y; int y_dtype; int y_pitch; int y_coff; int yH, yW;
  const void* res; int r_dtype; int r_pitch; int r_coff;   // residual with the geometry of y
  const float* bias;
  int act;
  int shuffle_r;        // 8: pixel shuffle addressing (y is the (8H, 8W, 32) map)
  // fused 1x1 predictor (dbsr_conv2d_tc_predictor): y is NOT written; every epilogue thread owns all n_tile = cout_pad
  // channels of its pixel and stores pred[n, k, y, x] = relu(pred_b[k] + sum_c pred_w[k][c] * act(conv)[c]), fp32 NCHW
  // The predictor weights travel IN the kernel parameters (constant bank): the 96 FMAs per pixel then take their weight
  // operand from the constant cache -- a first version that read them from shared memory added 1024 LDS wavefronts per
  // item to a kernel whose bottleneck is the shared-memory data pipe and was 100 us slower per launch.
  int tma_store;        // lean epilogue: the staged 32-pixel x GW-channel block of a warp leaves as ONE TMA tensor store
  void* pred; int pred_c;
  int pair_cta;         // 1: launched as clusters of two CTAs that execute every MMA together (tcgen05 cta_group::2, M = 256):
                        //    CTA rank r of pair q works on tile index 2 * (q / ntiles_n) + r of N tile q % ntiles_n; total_items
                        //    counts PAIR items; b_bytes is the half weight tile (n_tile / 2 rows) each CTA holds
  int pred_q14;         // 1: pred is int16 NCHW, value = (int16)(min(relu(.), 1) * 2^14)  (evaluation/*/compute_score.py:110)
  float pred_wb[4 * 32 + 4];   // [k][32] weights (zero beyond cout), then [k] biases
};

// epilogue of NC (32 or 16) accumulator columns held by one thread (= one output pixel).
// rq: residual values of this chunk prefetched by the caller (fast path only, nullptr = load here)
template <int NC>
__device__ __forceinline__ void epilogue_chunk(const ConvTcParams& p, const uint32_t (&r)[32], int co, long long off,
                                               long long roff, const uint4* rq_pref) {
  // co: first output channel of the chunk; off / roff: element offsets of that channel in y / residual
  float v[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) v[j] = __uint_as_float(r[j]);
  if (p.align_ok && co + NC <= p.cout) {
    uint4 rq[NC / 8];
    if (p.res) {   // bf16 on the fast path
      if (rq_pref) {
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) rq[j] = rq_pref[j];
      } else {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + roff;
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) rq[j] = __ldg(reinterpret_cast<const uint4*>(rp + 8 * j));
      }
    }
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < NC; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + co + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (p.res) {
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rq[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(h[k]);
          v[8 * j + 2 * k] += f.x; v[8 * j + 2 * k + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = apply_act(v[j], p.act);
    if (p.y_dtype == DBSR_BF16) {
      __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + off;
#pragma unroll
      for (int j = 0; j < NC; j += 8) {
        uint4 q;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
        q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
        q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(yp + j) = q;
      }
    } else {
      float* yp = reinterpret_cast<float*>(p.y) + off;
#pragma unroll
      for (int j = 0; j < NC; j += 4) *reinterpret_cast<float4*>(yp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  } else {
    // generic path: any channel count / alignment / residual dtype, masked at cout
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (co + j < p.cout) {
        float t = v[j];
        if (p.bias) t += __ldg(p.bias + co + j);
        if (p.res) {
          t += (p.r_dtype == DBSR_F32) ? __ldg(reinterpret_cast<const float*>(p.res) + roff + j)
                                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.res)[roff + j]);
        }
        t = apply_act(t, p.act);
        if (p.y_dtype == DBSR_BF16) reinterpret_cast<__nv_bfloat16*>(p.y)[off + j] = __float2bfloat16_rn(t);
        else reinterpret_cast<float*>(p.y)[off + j] = t;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Coalesced bf16 epilogue of one 16x8 tile for one warp (32 pixels = 4 tile rows x 8).
// The L1 data pipe is shared by the tensor core's smem operand reads and by LSU traffic, and a warp instruction
// whose lanes touch 32 different 128-byte lines costs 32 wavefronts.  So results (and residuals) are transposed
// through a padded per-warp smem staging area: thread-per-pixel rows on the TMEM side, 128-byte-line-per-8-lanes
// on the global side (4 wavefronts per instruction instead of 32).
// ---------------------------------------------------------------------------------------------------------
constexpr int STG_ROW = 144;                    // 128 data bytes + 16 pad: conflict-free for both access patterns
constexpr int STG_WARP_BYTES = 32 * STG_ROW;    // per epilogue warp
constexpr int BAR_BYTES = 2048;                 // mbarriers + TMEM slot (up to 64 resident weight stages)
constexpr int BIAS_TAB_BYTES = 2048;            // bias table of the lean epilogue (cout_pad <= 512 floats)

// geometry of one 16x8 tile for the coalesced epilogue: pointers to its pixel (0,0) at the first channel of the N
// tile, element strides between tile rows / columns (pixel-shuffle folds into these), and how many rows / columns of
// the tile are inside the image
struct TileGeo {
  __nv_bfloat16* y; const __nv_bfloat16* r;
  int y_row, y_col, r_row, r_col;   // element strides
  int rows_in, cols_in;
};

// one 32-column chunk: accumulator (+bias, +residual from the staging row) -> activation -> bf16 -> staging row
template <int NC>
__device__ __forceinline__ void finish_chunk(const ConvTcParams& p, const uint32_t (&r)[32], const float* bias, uint8_t* myrow_c,
                                             bool has_res) {
  float v[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) v[j] = __uint_as_float(r[j]);
  if (bias) {
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (has_res) {
#pragma unroll
    for (int j = 0; j < NC / 8; ++j) {
      const uint4 q = *reinterpret_cast<const uint4*>(myrow_c + j * 16);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        v[8 * j + 2 * k] += f.x; v[8 * j + 2 * k + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) v[j] = apply_act(v[j], p.act);
#pragma unroll
  for (int j = 0; j < NC / 8; ++j) {
    uint4 q;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
    q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
    q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(myrow_c + j * 16) = q;
  }
}

template <int GW>   // group width in channels: 64, 32 or 16  (GW * 2 bytes per pixel row)
__device__ __forceinline__ void epilogue_group_bf16(const ConvTcParams& p, uint32_t tbase, int g0, int co0, const TileGeo& tg,
                                                    int quarter, int lane, uint8_t* stg, bool res_inflight) {
  constexpr int LPP = GW / 8;        // lanes (16-byte chunks) per pixel
  constexpr int PPI = 32 / LPP;      // pixels per warp instruction
  const int sub = lane / LPP, chunk = lane % LPP;
  // both accumulator chunks of the group in flight before anything waits on them
  uint32_t r0[32], r1[32];
  if (GW >= 32) tmem_ld32_nowait(tbase + (uint32_t)g0, r0); else tmem_ld16_nowait(tbase + (uint32_t)g0, r0);
  if (GW == 64) tmem_ld32_nowait(tbase + (uint32_t)(g0 + 32), r1);
  const bool has_res = p.res != nullptr;
  // ---- residual: coalesced asynchronous global -> staging rows (LPP lanes cover one pixel's GW*2 contiguous bytes);
  //      the first group of a tile was already issued before the accumulator wait (res_inflight)
  if (has_res) {
    if (!res_inflight) {
#pragma unroll
      for (int it = 0; it < LPP; ++it) {
        const int ml = it * PPI + sub;
        const int row = quarter * 4 + (ml >> 3), col = ml & 7;
        const bool ok = row < tg.rows_in && col < tg.cols_in;
        cp_async16(stg + ml * STG_ROW + chunk * 16, ok ? (const void*)(tg.r + row * tg.r_row + col * tg.r_col + g0 + chunk * 8) : (const void*)tg.r,
                   ok ? 16u : 0u);
      }
    }
    cp_async_wait_all();
    __syncwarp();
  }
  // ---- accumulator -> bias / residual / activation -> bf16 -> own staging row
  uint8_t* myrow = stg + lane * STG_ROW;
  const float* bias = p.bias ? p.bias + co0 + g0 : nullptr;
  tmem_wait_ld();
  if (GW >= 32) finish_chunk<32>(p, r0, bias, myrow, has_res); else finish_chunk<16>(p, r0, bias, myrow, has_res);
  if (GW == 64) finish_chunk<32>(p, r1, bias ? bias + 32 : nullptr, myrow + 64, has_res);
  __syncwarp();
  // ---- staging rows -> coalesced global stores
#pragma unroll
  for (int it = 0; it < LPP; ++it) {
    const int ml = it * PPI + sub;
    const int row = quarter * 4 + (ml >> 3), col = ml & 7;
    if (row < tg.rows_in && col < tg.cols_in) {
      const uint4 q = *reinterpret_cast<const uint4*>(stg + ml * STG_ROW + chunk * 16);
      *(reinterpret_cast<uint4*>(tg.y + row * tg.y_row + col * tg.y_col + g0) + chunk) = q;
    }
  }
  __syncwarp();
}

struct ItemCoord { int nt, img, y0, x0; };
__device__ __forceinline__ unsigned fast_div(unsigned n, unsigned d, unsigned mul, unsigned one, int fast) {
  return fast ? __umulhi(n, mul) + (n & one) : n / d;
}
__device__ __forceinline__ ItemCoord decode_item(const ConvTcParams& p, long long item64, unsigned pair_rank = 0u) {
  ItemCoord c;
  const unsigned item = (unsigned)item64;            // item counts fit 32 bits (64-bit div/mod costs ~100 instr each)
  unsigned tm = fast_div(item, (unsigned)p.ntiles_n, p.md_nt, p.one_nt, p.fastdiv);
  c.nt = (int)(item - tm * (unsigned)p.ntiles_n);
  if (p.pair_cta) tm = 2u * tm + pair_rank;          // the two CTAs of a pair take neighbouring tiles of the SAME N tile (a tile index
                                                     // past the end decodes to an image past the batch: loads zero-fill, stores are skipped)
  if (p.flat) {   // item = imgs_per_item consecutive images
    c.img = (int)tm * p.imgs_per_item; c.y0 = 0; c.x0 = 0;
    return c;
  }
  const unsigned per_img = (unsigned)(p.items_x * p.tiles_y);
  const unsigned img = fast_div(tm, per_img, p.md_img, p.one_img, p.fastdiv);
  const unsigned rem = tm - img * per_img;
  const unsigned ry = fast_div(rem, (unsigned)p.items_x, p.md_x, p.one_x, p.fastdiv);
  c.img = (int)img * p.imgs_per_item;
  c.y0 = (int)ry * TILE_H;
  c.x0 = (int)(rem - ry * (unsigned)p.items_x) * (TILE_W * p.mt);
  return c;
}

// ---------------------------------------------------------------------------------------------------------
// Lean coalesced bf16 epilogue (no epilogue-side residual): the 8 epilogue warps are per-warp latency bound -- one warp
// walks accumulator -> bias -> activation -> bf16 -> staging row -> coalesced store for its 32 pixels -- so every
// instruction that is not one of those is hoisted out of the per-item path: shared-state-space ld/st on 32-bit addresses,
// bias from a shared-memory table, lane pointers advanced by precomputed strides, unpredicated stores for interior tiles.
// ---------------------------------------------------------------------------------------------------------
// 32 (or 16) accumulator columns of this thread's pixel -> + bias -> activation -> bf16 -> own staging row.
// The row's 16-byte slot c lives at physical slot (c ^ swz) (see lean_group); c0 = first slot of this chunk.
template <int NC>
__device__ __forceinline__ void lean_chunk(const uint32_t (&r)[32], uint32_t bias_s, bool has_bias, int act, uint32_t row_s,
                                           uint32_t c0, uint32_t swz, bool has_res) {
  float v[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) v[j] = __uint_as_float(r[j]);
  if (has_bias) {
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      const float4 b = lds128f(bias_s + j * 4);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (has_res) {   // residual of this pixel, prefetched into the same staging slots the result will overwrite
#pragma unroll
    for (int j = 0; j < NC; j += 8) {
      const uint4 q = lds128(row_s + (((c0 + (uint32_t)(j >> 3)) ^ swz) << 4));
      v[j] += __uint_as_float(q.x << 16); v[j + 1] += __uint_as_float(q.x & 0xFFFF0000u);
      v[j + 2] += __uint_as_float(q.y << 16); v[j + 3] += __uint_as_float(q.y & 0xFFFF0000u);
      v[j + 4] += __uint_as_float(q.z << 16); v[j + 5] += __uint_as_float(q.z & 0xFFFF0000u);
      v[j + 6] += __uint_as_float(q.w << 16); v[j + 7] += __uint_as_float(q.w & 0xFFFF0000u);
    }
  }
  if (act == DBSR_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (act == DBSR_ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = apply_act(v[j], DBSR_ACT_LRELU);
  }
#pragma unroll
  for (int j = 0; j < NC; j += 8)
    sts128(row_s + (((c0 + (uint32_t)(j >> 3)) ^ swz) << 4), pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
           pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
}

// one channel group (GW = 64 / 32 / 16 channels) of this warp's 32 pixels (4 tile rows x 8)
//   ywarp: element pointer of the warp's first pixel (tile row quarter*4, column 0) at the group's first channel
//   rows_left / cols_in: in-image extent relative to that pixel; full: all 4 x 8 pixels are inside the image
// Staging layout (per warp, dense rows of GW*2 bytes = LPP 16-byte slots): slot c of row r sits at physical slot
// c ^ ((r * LPP / 8) % LPP).  Both access patterns are then bank-conflict free: the transposing writes (8 consecutive
// rows, same c, per quarter-warp) and the coalescing reads (8 / LPP consecutive rows x LPP slots per quarter-warp) each
// touch 8 distinct 16-byte bank groups.
template <int GW>
__device__ __forceinline__ void lean_group(uint32_t taddr, uint32_t bias_s, bool has_bias, int act, uint32_t stg_s, int lane,
                                           __nv_bfloat16* ywarp, long long y_row, long long y_col, bool full, int rows_left,
                                           int cols_in, bool has_res, const CUtensorMap* tmap_y = nullptr, int sc = 0, int sx = 0,
                                           int sy = 0, int sn = 0, int si = -1) {
  constexpr int LPP = GW / 8;        // lanes (16-byte slots) per pixel
  constexpr int PPI = 32 / LPP;      // pixels per warp instruction
  uint32_t r0[32], r1[32];
  if (GW >= 32) tmem_ld32_nowait(taddr, r0); else tmem_ld16_nowait(taddr, r0);
  if (GW == 64) tmem_ld32_nowait(taddr + 32u, r1);
  const int sub = lane / LPP, chunk = lane % LPP;
  const int r_lane = (PPI > 8) ? (sub >> 3) : 0, c_lane = sub & 7;
  __nv_bfloat16* yl = ywarp + r_lane * y_row + c_lane * y_col + chunk * 8;
  const uint32_t row_s = stg_s + (uint32_t)(lane * GW * 2);
  // swizzle = 128-byte line index of the row (ABSOLUTE shared address, as the TMA swizzle modes define it) mod LPP
  const uint32_t base_sw = stg_s >> 7;
  const uint32_t wswz = (base_sw + (uint32_t)((lane * LPP) >> 3)) % LPP;
  if (has_res) { cp_async_wait_all(); __syncwarp(); }     // the prefetched residual rows have landed
  tmem_wait_ld();
  if (tmap_y != nullptr) {       // the previous TMA store of this warp has finished READING the staging rows
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  }
  if (GW >= 32) lean_chunk<32>(r0, bias_s, has_bias, act, row_s, 0u, wswz, has_res); else lean_chunk<16>(r0, bias_s, has_bias, act, row_s, 0u, wswz, has_res);
  if (GW == 64) lean_chunk<32>(r1, bias_s + 128u, has_bias, act, row_s, 4u, wswz, has_res);
  if (tmap_y != nullptr) {
    // The staging rows are exactly the box {GW channels, 8 pixels, 4 rows} in the SWIZZLE_(2*GW)B layout: one elected lane
    // hands them to the TMA engine, which clips at the image border; the warp does not read them back (ncu / A-B
    // experiments: the LDS + STG read-back was 22 % of the 32 -> 32 layers and 24 % of the 4 -> 64 layer).
    fence_async_smem();
    __syncwarp();
    if (lane == 0 && rows_left > 0) {
      if (si < 0) tma_store_4d(tmap_y, stg_s, sc, sx, sy, sn);
      else tma_store_5d(tmap_y, stg_s, sc, sx, si, sy, sn);       // pixel shuffle: {(j, c), x, i, y, n}
    }
    return;
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < LPP; ++it) {
    // staging row ml = it * PPI + sub  ->  tile row (ml >> 3), column (ml & 7)
    const int r_it = (it * PPI) >> 3, c_it = (it * PPI) & 7;
    const uint32_t rswz = (base_sw + (uint32_t)(it * 4 + ((sub * LPP) >> 3))) % LPP;     // (line index of row ml) % LPP
    const uint4 q = lds128(stg_s + (uint32_t)((it * PPI + sub) * GW * 2) + (((uint32_t)chunk ^ rswz) << 4));
    if (full || (r_lane + r_it < rows_left && c_lane + c_it < cols_in))
      *reinterpret_cast<uint4*>(yl + r_it * y_row + c_it * y_col) = q;
  }
  __syncwarp();
}

// residual rows of this warp's 32 pixels (one channel group of GW channels) -> staging, asynchronously, in the layout
// lean_group reads them from; issued before the wait for the accumulators so the latency hides behind the MMAs
template <int GW>
__device__ __forceinline__ void lean_prefetch_res(uint32_t stg_s, int lane, const __nv_bfloat16* rwarp, long long r_row,
                                                  long long r_col, int rows_left, int cols_in, const void* safe) {
  constexpr int LPP = GW / 8, PPI = 32 / LPP;
  const int sub = lane / LPP, chunk = lane % LPP;
  const int r_lane = (PPI > 8) ? (sub >> 3) : 0, c_lane = sub & 7;
  const __nv_bfloat16* rl = rwarp + r_lane * r_row + c_lane * r_col + chunk * 8;
#pragma unroll
  for (int it = 0; it < LPP; ++it) {
    const int r_it = (it * PPI) >> 3, c_it = (it * PPI) & 7;
    const uint32_t rswz = ((stg_s >> 7) + (uint32_t)(it * 4 + ((sub * LPP) >> 3))) % LPP;
    const bool ok = r_lane + r_it < rows_left && c_lane + c_it < cols_in;
    const uint32_t dst = stg_s + (uint32_t)((it * PPI + sub) * GW * 2) + (((uint32_t)chunk ^ rswz) << 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ok ? (const void*)(rl + r_it * r_row + c_it * r_col) : safe),
                 "r"(ok ? 16u : 0u) : "memory");
  }
}

template <int CK, bool RESIDENT, bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_i,
               const __grid_constant__ CUtensorMap tmap_y, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                         // [a_slots][a_bytes]
  uint8_t* smem_b = smem + (size_t)p.a_slots * p.a_bytes;         // [b_stages][b_bytes]
  uint8_t* smem_stg = smem_b + (size_t)p.b_stages * p.b_stage_bytes;    // [8 epilogue warps][STG_WARP_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + 8 * STG_WARP_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.a_slots;
  uint64_t* b_full = a_empty + p.a_slots;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* tfull_bar = b_empty + p.b_stages;
  uint64_t* tempty_bar = tfull_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 4);
  float* bias_tab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);   // [cout_pad] when p.bias_smem

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PAIR: cluster of two CTAs; rank 0 (the leader) issues every MMA for both, each CTA loads its own activation boxes and its
  // half of every weight tile and runs the epilogue of its own 128 accumulator lanes
  const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0u;
  const long long item_first = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long item_step = PAIR ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  constexpr uint32_t ROW_BYTES = CK * 2;                 // bytes per pixel of a K chunk
  constexpr uint32_t LAYOUT = (CK == 64) ? 2u : 4u;
  // cute::UMMA::InstrDescriptor: c_format F32 [4,6) | a,b format BF16 [7,10),[10,13) | K-major | N>>3 [17,23) | M>>4 [24,29)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
  const int NT = p.n_tile;
  const int taps = p.ksize * p.ksize;
  const int pad = (p.ksize == 3) ? p.dil : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_slots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    // PAIR: the leader's issuer waits for the epilogue warps of BOTH CTAs before it reuses an accumulator stage
    for (int a = 0; a < p.acc_stages; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if (p.res_chunks) { tma_prefetch_desc(&tmap_r); tma_prefetch_desc(&tmap_i); }
    if (p.tma_store) tma_prefetch_desc(&tmap_y);
  }
  if (warp == WARP_MMA) { if (PAIR) tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols); else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); }
  // Programmatic dependent launch: everything above touches only on-chip state and the kernel parameters, so it may run
  // while the preceding kernel of the stream is still draining; all global-memory traffic (activations, residual, y,
  // and -- because a caller may have produced them just before -- weights and bias) comes after this wait.
  // With static weights (constants of the caller) only the threads that read activations or write the output wait; the
  // weight producer starts right away, so the weight ring / the resident tiles fill while the preceding kernel still runs.
  if (p.early_trigger) griddep_launch_dependents();
  if (!p.static_w) griddep_wait();
  if (p.bias_fill && warp < 8)
    for (int i = threadIdx.x; i < p.cout_pad; i += 256) bias_tab[i] = p.bias ? __ldg(p.bias + i) : 0.0f;
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();          // both CTAs' barriers are initialised before any cross-CTA arrive / TMA completion
  tc_fence_after();
  if (p.static_w && warp != WARP_B && warp != WARP_MMA) griddep_wait();   // A producer, epilogue warps (residual loads, stores)
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WARP_A) {
    // ===================== A producer: one halo box per (item, K chunk), then the residual boxes =====================
    if (elect_one()) {
      int slot = 0; uint32_t phase = 0;
      for (long long item = item_first; item < p.total_items; item += item_step) {
        const ItemCoord c = decode_item(p, item, pair_rank);
        for (int ch = 0; ch < p.nchunks; ++ch) {
          mbar_wait(&a_empty[slot], phase ^ 1, 100 + slot);
          if (PAIR) {     // both boxes complete on the LEADER's barrier; only the leader arms it (for the bytes of both)
            if (pair_rank == 0) mbar_arrive_expect_tx(&a_full[slot], 2u * (uint32_t)p.a_tx_bytes);
            tma_load_4d_pair(&tmap_x, &a_full[slot], smem_a + (size_t)slot * p.a_bytes, ch * CK, c.x0 - pad, c.y0 - pad, c.img);
          } else {
            mbar_arrive_expect_tx(&a_full[slot], (uint32_t)p.a_tx_bytes);
            tma_load_4d(&tmap_x, &a_full[slot], smem_a + (size_t)slot * p.a_bytes, ch * CK, c.x0 - pad, c.y0 - pad, c.img);   // flat: (-1, -1, first image)
          }
          if (++slot == p.a_slots) { slot = 0; phase ^= 1; }
        }
        // residual of this N tile as extra K chunks: plain (no halo) box of the residual tensor
        for (int rc = 0; rc < p.res_chunks; ++rc) {
          mbar_wait(&a_empty[slot], phase ^ 1, 120 + slot);
          if (PAIR) {
            if (pair_rank == 0) mbar_arrive_expect_tx(&a_full[slot], 2u * (uint32_t)p.r_tx_bytes);
            tma_load_4d_pair(&tmap_r, &a_full[slot], smem_a + (size_t)slot * p.a_bytes, c.nt * NT + rc * CK, c.x0, c.y0, c.img / p.res_group);
          } else {
            mbar_arrive_expect_tx(&a_full[slot], (uint32_t)p.r_tx_bytes);
            tma_load_4d(&tmap_r, &a_full[slot], smem_a + (size_t)slot * p.a_bytes, c.nt * NT + rc * CK, c.x0, c.y0, c.img / p.res_group);
          }
          if (++slot == p.a_slots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == WARP_B) {
    // ===================== B producer: weight tiles =====================
    if (elect_one()) {
      if (RESIDENT) {
        // every item uses the same tiles (one N tile): load each (chunk, tap) tile once, they all signal b_full[0]
        const int n_main = p.nchunks * taps;
        // PAIR: each CTA holds rows [rank * NT / 2, (rank + 1) * NT / 2) of every weight tile (b_bytes is that half)
        const int row0 = PAIR ? (int)pair_rank * (NT >> 1) : 0;
        if (!PAIR || pair_rank == 0) mbar_arrive_expect_tx(&b_full[0], (uint32_t)((PAIR ? 2 : 1) * (n_main + p.res_chunks) * p.b_bytes));
        for (int t = 0; t < n_main; ++t) {
          const int ch = t / taps, tap = t - ch * taps;
          if (PAIR) tma_load_3d_pair(&tmap_w, &b_full[0], smem_b + (size_t)t * p.b_bytes, ch * CK, row0, tap);
          else tma_load_3d(&tmap_w, &b_full[0], smem_b + (size_t)t * p.b_bytes, ch * CK, 0, tap);
        }
        for (int rc = 0; rc < p.res_chunks; ++rc) {   // identity tiles of the residual chunks
          if (PAIR) tma_load_2d_pair(&tmap_i, &b_full[0], smem_b + (size_t)(n_main + rc) * p.b_bytes, rc * CK, row0);
          else tma_load_2d(&tmap_i, &b_full[0], smem_b + (size_t)(n_main + rc) * p.b_bytes, rc * CK, 0);
        }
      } else {
        int stage = 0; uint32_t phase = 0;
        const int row0 = PAIR ? (int)pair_rank * (NT >> 1) : 0;
        for (long long item = item_first; item < p.total_items; item += item_step) {
          const int nt = (int)((unsigned)item % (unsigned)p.ntiles_n);
          for (int ch = 0; ch < p.nchunks; ++ch) {
            for (int tap = 0; tap < taps; tap += p.b_taps) {      // one box {CK, N tile rows, b_taps taps} per stage
              mbar_wait(&b_empty[stage], phase ^ 1, 150 + stage);
              if (PAIR) {
                if (pair_rank == 0) mbar_arrive_expect_tx(&b_full[stage], 2u * (uint32_t)p.b_stage_bytes);
                tma_load_3d_pair(&tmap_w, &b_full[stage], smem_b + (size_t)stage * p.b_stage_bytes, ch * CK, nt * NT + row0, tap);
              } else {
                mbar_arrive_expect_tx(&b_full[stage], (uint32_t)p.b_stage_bytes);
                tma_load_3d(&tmap_w, &b_full[stage], smem_b + (size_t)stage * p.b_stage_bytes, ch * CK, nt * NT, tap);
              }
              if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
            }
          }
          for (int rc = 0; rc < p.res_chunks; ++rc) {
            mbar_wait(&b_empty[stage], phase ^ 1, 170 + stage);
            if (PAIR) {
              if (pair_rank == 0) mbar_arrive_expect_tx(&b_full[stage], 2u * (uint32_t)p.b_bytes);
              tma_load_2d_pair(&tmap_i, &b_full[stage], smem_b + (size_t)stage * p.b_stage_bytes, rc * CK, row0);
            } else {
              mbar_arrive_expect_tx(&b_full[stage], (uint32_t)p.b_bytes);
              tma_load_2d(&tmap_i, &b_full[stage], smem_b + (size_t)stage * p.b_stage_bytes, rc * CK, 0);
            }
            if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer =====================
    // One in-order instruction stream: everything that is not an MMA, a barrier wait or a commit is hoisted out of the
    // per-tap path (descriptor high words, tap offsets, weight-tile addresses are plain adds on the low word).
    if ((!PAIR || pair_rank == 0) && elect_one()) {
      int aslot = 0; uint32_t aphase = 0;
      int bstage = 0; uint32_t bphase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      // flat mode: M rows are consecutive slots of the box (dense 8-slot atoms); tap shift is ky*halo_w + kx slots
      const uint32_t a_sbo = (p.flat ? 8u : (uint32_t)p.halo_w) * ROW_BYTES;
      const uint32_t b_sbo = 8u * ROW_BYTES;
      const uint32_t r_sbo = (uint32_t)(TILE_W * p.mt) * ROW_BYTES;          // residual box: 8*mt pixels per row, no halo
      const uint32_t a_hi = ((a_sbo >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT << 29);
      const uint32_t b_hi = ((b_sbo >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT << 29);
      const uint32_t r_hi = ((r_sbo >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT << 29);
      uint32_t tap_off[9];                      // 16-byte units inside the A box
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - 3 * (tap / 3);
        tap_off[tap] = (p.ksize == 3) ? (uint32_t)((ky * p.dil) * p.halo_w + kx * p.dil) * (ROW_BYTES >> 4) : 0u;
      }
      const uint32_t tile_step = (uint32_t)p.t1_step16;                  // second M tile of the item
      const bool two_tiles = p.mt == 2;
      const uint32_t b_step = (uint32_t)p.b_bytes >> 4;
      const uint32_t b_stage_step = (uint32_t)p.b_stage_bytes >> 4;
      const int b_taps = p.b_taps;
      const uint32_t a_step = (uint32_t)p.a_bytes >> 4;
      const uint32_t a_base = ((smem_u32(smem_a) >> 4) & 0x3FFFu) | 0x10000u;
      const uint32_t b_base = ((smem_u32(smem_b) >> 4) & 0x3FFFu) | 0x10000u;
      if (RESIDENT) { mbar_wait(&b_full[0], 0, 350); tc_fence_after(); }   // all weight tiles landed, once per CTA

      auto mma_pair = [&](uint32_t a_lo, uint32_t a_hi_w, uint32_t b_lo, uint32_t d0, uint32_t first_acc) {
#pragma unroll
        for (int k16 = 0; k16 < CK / 16; ++k16)
          umma_bf16_t<PAIR>(d0, ((uint64_t)a_hi_w << 32) | (uint64_t)(a_lo + 2u * k16), ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + 2u * k16),
                            idesc, k16 == 0 ? first_acc : 1u);
        if (two_tiles) {
#pragma unroll
          for (int k16 = 0; k16 < CK / 16; ++k16)
            umma_bf16_t<PAIR>(d0 + (uint32_t)NT, ((uint64_t)a_hi_w << 32) | (uint64_t)(a_lo + tile_step + 2u * k16),
                              ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + 2u * k16), idesc, k16 == 0 ? first_acc : 1u);
        }
      };

      for (long long item = item_first; item < p.total_items; item += item_step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * p.mt * NT);
        uint32_t b_lo_res = b_base;               // resident: running tile address
        for (int ch = 0; ch < p.nchunks; ++ch) {
          mbar_wait(&a_full[aslot], aphase, 300 + aslot);
          tc_fence_after();
          const uint32_t a_lo0 = a_base + (uint32_t)aslot * a_step;
          if (taps == 9) {
            int tis = 0;                           // tap within the current weight stage
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              uint32_t b_lo;
              if (RESIDENT) { b_lo = b_lo_res; b_lo_res += b_step; }
              else {
                if (tis == 0) {
                  mbar_wait(&b_full[bstage], bphase, 350 + bstage);
                  tc_fence_after();
                }
                b_lo = b_base + (uint32_t)bstage * b_stage_step + (uint32_t)tis * b_step;
              }
              mma_pair(a_lo0 + tap_off[tap], a_hi, b_lo, d0, (tap == 0 && ch == 0) ? 0u : 1u);
              if (!RESIDENT) {
                if (++tis == b_taps) {
                  tis = 0;
                  umma_commit_t<PAIR>(&b_empty[bstage]);   // weight stage free once these MMAs retire
                  if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
                }
              }
            }
          } else {
            uint32_t b_lo;
            if (RESIDENT) { b_lo = b_lo_res; b_lo_res += b_step; }
            else {
              mbar_wait(&b_full[bstage], bphase, 350 + bstage);
              tc_fence_after();
              b_lo = b_base + (uint32_t)bstage * b_stage_step;
            }
            mma_pair(a_lo0, a_hi, b_lo, d0, ch == 0 ? 0u : 1u);
            if (!RESIDENT) {
              umma_commit_t<PAIR>(&b_empty[bstage]);
              if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
            }
          }
          umma_commit_t<PAIR>(&a_empty[aslot]);          // halo box free once every tap of this chunk retired
          if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
        }
        // residual: D += R * I  (exact: identity weights, fp32 accumulation) -- the epilogue never touches it
        for (int rc = 0; rc < p.res_chunks; ++rc) {
          mbar_wait(&a_full[aslot], aphase, 320 + aslot);
          tc_fence_after();
          uint32_t b_lo;
          if (RESIDENT) { b_lo = b_lo_res; b_lo_res += b_step; }
          else {
            mbar_wait(&b_full[bstage], bphase, 370 + bstage);
            tc_fence_after();
            b_lo = b_base + (uint32_t)bstage * b_stage_step;
          }
          mma_pair(a_base + (uint32_t)aslot * a_step, r_hi, b_lo, d0, 1u);
          if (!RESIDENT) {
            umma_commit_t<PAIR>(&b_empty[bstage]);
            if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
          }
          umma_commit_t<PAIR>(&a_empty[aslot]);
          if (++aslot == p.a_slots) { aslot = 0; aphase ^= 1; }
        }
        umma_commit_t<PAIR>(&tfull_bar[acc]);    // accumulators ready for the epilogue (PAIR: of both CTAs)
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue (warps 0..7) =====================
    // two warps per TMEM lane quarter: with mt == 2 group g owns tile g of every item, with mt == 1 the groups take
    // alternate channel groups.
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int group = warp >> 2;                   // 0 or 1
    const int m = quarter * 32 + lane;             // pixel within a 16x8 tile
    const int ty = m >> 3, tx = m & 7;
    uint8_t* stg = smem_stg + (size_t)warp * STG_WARP_BYTES;
    int acc = 0; uint32_t acc_phase = 0;
    if (p.pred == nullptr && !p.flat && p.bias_smem && p.vec_ok && p.y_dtype == DBSR_BF16 &&
        (p.res == nullptr || NT == 128 || NT == 64 || NT == 32 || NT == 16)) {
      // ---------- lean coalesced path (every bf16 layer of the encoder / fusion / decoder trunks) ----------
      const uint32_t stg_s = smem_u32(stg);
      const uint32_t bias_s0 = smem_u32(bias_tab);
      const bool has_bias = p.bias != nullptr;
      const int act = p.act;
      const int t = (p.mt == 2) ? group : 0;
      const bool shuffle = p.shuffle_r > 1;
      const long long y_row = shuffle ? (long long)p.shuffle_r * p.yW * p.y_pitch : (long long)p.yW * p.y_pitch;
      const long long y_col = shuffle ? (long long)p.shuffle_r * p.y_pitch : (long long)p.y_pitch;
      __nv_bfloat16* const ybase = reinterpret_cast<__nv_bfloat16*>(p.y) + p.y_coff + (long long)(quarter * 4) * y_row;
      const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * NT);
      const uint32_t acc_cols = (uint32_t)(p.mt * NT);
      // epilogue-side residual (single channel group per warp, N tile <= 64): prefetched into the staging rows
      const bool has_res = p.res != nullptr;
      const long long r_row = (long long)p.yW * p.r_pitch, r_col = (long long)p.r_pitch;
      const __nv_bfloat16* const rbase = reinterpret_cast<const __nv_bfloat16*>(p.res) + p.r_coff + (long long)(quarter * 4) * r_row;
      // NT <= 64 is one group: with mt == 1 warps 4..7 have no columns; NT == 128 with mt == 1: warps 4..7 own the second group
      const bool my_group = (p.mt == 2) || group == 0 || NT == 128;
      const int first_g0 = (p.mt == 2 || NT != 128) ? 0 : group * 64;
      for (long long item = item_first; item < p.total_items; item += item_step) {
        const ItemCoord c = decode_item(p, item, pair_rank);
        const int co0 = c.nt * NT;
        const int tx0 = c.x0 + t * p.t1_dx;
        const int img_t = c.img + t * p.t1_dimg;             // narrow maps: the second tile is the next image
        __nv_bfloat16* ywarp;
        if (shuffle) {
          // packed channel co' = i*256 + j*32 + c ; N tile = 128 -> HR row phase i = nt / 2, first HR column j0 = (nt & 1) * 4
          const int per_i = p.shuffle_r * 32;
          const int si = co0 / per_i, j0 = (co0 - si * per_i) >> 5;
          ywarp = ybase + ((long long)(img_t * p.yH + c.y0 * p.shuffle_r + si) * p.yW + (tx0 * p.shuffle_r + j0)) * p.y_pitch;
        } else {
          ywarp = ybase + ((long long)(img_t * p.yH + c.y0) * p.yW + tx0) * p.y_pitch + co0;
        }
        const int rows_left = (img_t < p.n) ? p.H - c.y0 - quarter * 4 : 0, cols_in = p.W - tx0;   // image past the batch: nothing
        const bool full = rows_left >= 4 && cols_in >= TILE_W;
        const uint32_t tbase = tq + (uint32_t)acc * acc_cols;
        const uint32_t bias_s = bias_s0 + (uint32_t)co0 * 4u;
        if (has_res && my_group) {
          if (p.tma_store) {      // the previous item's tensor store has finished reading the rows the prefetch overwrites
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
          const __nv_bfloat16* rwarp = rbase + ((long long)(img_t * p.yH + c.y0) * p.yW + tx0) * p.r_pitch + co0 + first_g0;
          if (NT >= 64) lean_prefetch_res<64>(stg_s, lane, rwarp, r_row, r_col, rows_left, cols_in, p.res);
          else if (NT == 32) lean_prefetch_res<32>(stg_s, lane, rwarp, r_row, r_col, rows_left, cols_in, p.res);
          else lean_prefetch_res<16>(stg_s, lane, rwarp, r_row, r_col, rows_left, cols_in, p.res);
        }
        mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
        tc_fence_after();
        int g0 = 0, gi = 0;
        while (g0 < NT) {
          const int rem = NT - g0;
          const int gw = rem >= 64 ? 64 : (rem >= 32 ? 32 : 16);
          if (p.mt == 2 || (gi & 1) == group) {
            if (has_res && g0 > first_g0) {     // second 64-channel group of an N = 128 tile: its residual rows, once the store
              if (p.tma_store) {                // of the first group has finished reading the staging rows
                if (lane == 0) tma_store_wait_read();
              }
              __syncwarp();
              const __nv_bfloat16* rwarp = rbase + ((long long)(img_t * p.yH + c.y0) * p.yW + tx0) * p.r_pitch + co0 + g0;
              lean_prefetch_res<64>(stg_s, lane, rwarp, r_row, r_col, rows_left, cols_in, p.res);
            }
            const CUtensorMap* ty_map = p.tma_store ? &tmap_y : nullptr;
            const int sy = c.y0 + quarter * 4;
            // pixel shuffle: packed channel co' = i*256 + j*32 + c -> store coordinates ((j, c) = co' % 256, i = co' / 256)
            const int sc = shuffle ? (co0 + g0) % (p.shuffle_r * 32) : co0 + g0;
            const int si = shuffle ? (co0 + g0) / (p.shuffle_r * 32) : -1;
            if (gw == 64) lean_group<64>(tbase + (uint32_t)g0, bias_s + (uint32_t)g0 * 4u, has_bias, act, stg_s, lane, ywarp + g0, y_row, y_col, full, rows_left, cols_in, has_res, ty_map, sc, tx0, sy, img_t, si);
            else if (gw == 32) lean_group<32>(tbase + (uint32_t)g0, bias_s + (uint32_t)g0 * 4u, has_bias, act, stg_s, lane, ywarp + g0, y_row, y_col, full, rows_left, cols_in, has_res, ty_map, sc, tx0, sy, img_t, si);
            else lean_group<16>(tbase + (uint32_t)g0, bias_s + (uint32_t)g0 * 4u, has_bias, act, stg_s, lane, ywarp + g0, y_row, y_col, full, rows_left, cols_in, has_res, ty_map, sc, tx0, sy, img_t, si);
          }
          g0 += gw; ++gi;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
      }
    } else
    for (long long item = item_first; item < p.total_items; item += item_step) {
      const ItemCoord c = decode_item(p, item, pair_rank);
      const int co0 = c.nt * NT;
      const int t = (p.mt == 2) ? group : 0;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * p.mt + t) * NT);
      if (p.pred != nullptr) {
        // ---------- fused 1x1 predictor (NT == cout_pad == 32): the thread holds all channels of its pixel ----------
        const int y = c.y0 + ty, x = c.x0 + t * p.t1_dx + tx;
        const long long img = c.img + t * p.t1_dimg;
        const bool valid = (y < p.H) && (x < p.W) && (img < p.n) && (p.mt == 2 || group == 0);
        mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
        tc_fence_after();
        if (p.mt == 2 || group == 0) {
          uint32_t r[32];
          tmem_ld32(tbase, r);
          if (valid) {
            float v[32];
            const uint32_t bias_s = smem_u32(bias_tab);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = lds128f(bias_s + (uint32_t)j * 4u);
              v[j] = apply_act(__uint_as_float(r[j]) + b4.x, p.act); v[j + 1] = apply_act(__uint_as_float(r[j + 1]) + b4.y, p.act);
              v[j + 2] = apply_act(__uint_as_float(r[j + 2]) + b4.z, p.act); v[j + 3] = apply_act(__uint_as_float(r[j + 3]) + b4.w, p.act);
            }
            const long long plane = (long long)p.H * p.W;
            const long long o = img * p.pred_c * plane + (long long)y * p.W + x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < p.pred_c) {
                float s0 = p.pred_wb[128 + k], s1 = 0.0f;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  s0 = fmaf(v[j], p.pred_wb[k * 32 + j], s0);
                  s1 = fmaf(v[j + 1], p.pred_wb[k * 32 + j + 1], s1);
                }
                const float out = fmaxf(s0 + s1, 0.0f);
                if (p.pred_q14) reinterpret_cast<short*>(p.pred)[o + k * plane] = (short)(fminf(out, 1.0f) * 16384.0f);
                else reinterpret_cast<float*>(p.pred)[o + k * plane] = out;
              }
            }
          }
        }
      } else if (!p.flat && p.vec_ok && p.y_dtype == DBSR_BF16) {
        // ---------- coalesced path: 64/32/16-channel groups through the per-warp staging rows ----------
        const int tx0 = c.x0 + t * p.t1_dx;
        const long long img_t = c.img + t * p.t1_dimg;
        TileGeo tg;
        if (p.shuffle_r > 1) {
          // packed channel co' = i*256 + j*32 + c ; N tile = 128 -> HR row phase i = nt / 2, first HR column j0 = (nt & 1) * 4
          const int per_i = p.shuffle_r * 32;
          const int si = co0 / per_i, j0 = (co0 - si * per_i) / 32;
          tg.y = reinterpret_cast<__nv_bfloat16*>(p.y) +
                 ((img_t * p.yH + (c.y0 * p.shuffle_r + si)) * p.yW + (tx0 * p.shuffle_r + j0)) * p.y_pitch + p.y_coff;
          tg.y_row = p.shuffle_r * p.yW * p.y_pitch; tg.y_col = p.shuffle_r * p.y_pitch;
        } else {
          tg.y = reinterpret_cast<__nv_bfloat16*>(p.y) + ((img_t * p.yH + c.y0) * p.yW + tx0) * p.y_pitch + p.y_coff + co0;
          tg.y_row = p.yW * p.y_pitch; tg.y_col = p.y_pitch;
        }
        tg.r = reinterpret_cast<const __nv_bfloat16*>(p.res) + ((img_t * p.yH + c.y0) * p.yW + tx0) * p.r_pitch + p.r_coff + co0;
        tg.r_row = p.yW * p.r_pitch; tg.r_col = p.r_pitch;
        tg.rows_in = (img_t < p.n) ? min(TILE_H, p.H - c.y0) : 0; tg.cols_in = min(TILE_W, p.W - tx0);
        mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
        tc_fence_after();
        int g0 = 0, gi = 0;
        while (g0 < NT) {
          const int rem = NT - g0;
          const int gw = rem >= 64 ? 64 : (rem >= 32 ? 32 : 16);
          if (p.mt == 2 || (gi & 1) == group) {
            if (gw == 64) epilogue_group_bf16<64>(p, tbase, g0, co0, tg, quarter, lane, stg, false);
            else if (gw == 32) epilogue_group_bf16<32>(p, tbase, g0, co0, tg, quarter, lane, stg, false);
            else epilogue_group_bf16<16>(p, tbase, g0, co0, tg, quarter, lane, stg, false);
          }
          g0 += gw; ++gi;
        }
      } else {
        // ---------- generic path (fp32 outputs, odd channel counts, flat mode): one thread stores its own pixel ----------
        int y = c.y0 + ty, x = c.x0 + t * p.t1_dx + tx;
        long long img = c.img + t * p.t1_dimg;
        bool valid = (y < p.H) && (x < p.W) && (img < p.n);
        if (p.flat) {   // M row m = flat slot: image j = m / S, padded row / column inside it
          const int j = m / p.flat_s, rem = m - j * p.flat_s;
          y = rem / p.halo_w; x = rem - y * p.halo_w;
          img += j;
          valid = (j < p.flat_ni) && (y < p.H) && (x < p.W) && (img < p.n);   // rows past the box belong to nobody
        }
        const long long off = ((img * p.yH + y) * p.yW + x) * p.y_pitch + p.y_coff + co0;
        const long long roff = ((img * p.yH + y) * p.yW + x) * p.r_pitch + p.r_coff + co0;
        mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c0 = ci * 32;
          if (c0 < NT && (p.mt == 2 || (ci & 1) == group)) {
            uint32_t r[32];
            if (c0 + 32 <= NT) {
              tmem_ld32(tbase + (uint32_t)c0, r);
              if (valid) epilogue_chunk<32>(p, r, co0 + c0, off + c0, roff + c0, nullptr);
            } else {   // 16-column tail (n_tile is a multiple of 16)
              tmem_ld16(tbase + (uint32_t)c0, r);
              if (valid) epilogue_chunk<16>(p, r, co0 + c0, off + c0, roff + c0, nullptr);
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld inside the loads): release the accumulators
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (p.tma_store && warp < 8 && lane == 0) tma_store_wait_all();     // this warp's tensor stores have been written
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();          // neither CTA leaves (or frees TMEM) while the peer may still signal its barriers
  if (warp == WARP_MMA) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Tiling rule shared with the Python weight packer through dbsr_conv2d_tc_geometry().
static void tc_geometry(int cin, int cout, int* ck, int* kpad, int* n_tile, int* cout_pad) {
  const int k32 = round_up(cin, 32), k64 = round_up(cin, 64);
  *ck = (k64 <= k32) ? 64 : 32;          // smallest padded K; ties go to the wider chunk
  *kpad = (*ck == 64) ? k64 : k32;
  int nt;
  if (cout % 128 == 0) nt = 128;
  else if (cout % 64 == 0) nt = 64;
  else {
    const int c16 = round_up(cout, 16);
    if (c16 <= 128) nt = c16;
    else {
      const int parts = (c16 + 127) / 128;
      nt = round_up((c16 + parts - 1) / parts, 16);
    }
  }
  *n_tile = nt;
  *cout_pad = round_up(cout, nt);
}

// [n][n] bf16 identity ("weights" of the residual K chunks), created once per (device, n) and kept for the process
// lifetime.  Creation is synchronous (cudaMalloc + blocking cudaMemcpy) so that the tile is complete before ANY stream of
// the device can read it; it therefore cannot happen inside a stream capture (the engine always runs a layer eagerly before
// capturing it).
const void* identity_weights(int n, cudaStream_t st) {
  static std::mutex mu;
  static void* cache[MAX_DEVICES][17] = {};
  const int slot = n / 16;
  if (slot < 1 || slot > 16) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  void*& entry = cache[current_device_slot()][slot];
  if (!entry) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
    std::vector<__nv_bfloat16> host((size_t)n * n, __float2bfloat16_rn(0.0f));
    for (int i = 0; i < n; ++i) host[(size_t)i * n + i] = __float2bfloat16_rn(1.0f);
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, (size_t)n * n * 2) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemcpy(ptr, host.data(), (size_t)n * n * 2, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); cudaFree(ptr); return nullptr; }
    entry = ptr;
  }
  return entry;
}

struct TcConfig {
  int flat, flat_s, flat_ni, res_chunks;
  int pair_img;    // narrow maps (one tile column): the two M tiles of an item are two consecutive images
  int n_tile, ck, nchunks, cout, cout_pad, mt, halo_w, rows, a_slots, b_stages, b_resident, a_bytes, a_tx_bytes, b_bytes,
      b_taps, b_stage_bytes, smem_bytes, tmem_cols, vec_ok, align_ok, acc_stages;
  int pair_cta;    // CTA pairs (cta_group::2): b_bytes is then the HALF weight tile one CTA holds
};

static int tc_plan(const dbsr_conv_t* c, TcConfig* cfg, bool set_err, bool allow_pair = true, int n_split = 1, bool allow_split = true) {
#define TC_REQ(cond, ...) do { if (!(cond)) { if (set_err) set_error(__VA_ARGS__); return 1; } } while (0)
  TC_REQ(c && view_ok(&c->x) && view_ok(&c->y) && c->w, "conv2d_tc: bad descriptor");
  TC_REQ(c->x.dtype == DBSR_BF16, "conv2d_tc: input must be bf16");
  TC_REQ(c->ksize == 1 || c->ksize == 3, "conv2d_tc: ksize must be 1 or 3");
  TC_REQ(c->stride == 1, "conv2d_tc: stride must be 1");
  TC_REQ(c->dilation >= 1 && c->dilation <= 16, "conv2d_tc: dilation out of range");
  const int r = c->shuffle_r > 1 ? c->shuffle_r : 1;
  const int cout = c->y.c * r * r;
  TC_REQ(c->y.n == c->x.n && c->y.h == c->x.h * r && c->y.w == c->x.w * r, "conv2d_tc: output geometry mismatch");
  TC_REQ((c->x.c_off % 8) == 0 && (c->x.c_pitch % 8) == 0 && ((uintptr_t)c->x.data % 16) == 0,
         "conv2d_tc: input view must be 16-byte aligned (c_off, c_pitch multiples of 8)");
  int ck, kpad, nt, cpad;
  tc_geometry(c->x.c, cout, &ck, &kpad, &nt, &cpad);
  // n_split > 1 (chosen at the end of this function for launches that would occupy a fraction of the SMs): the same packed
  // weights, N tiles n_split times narrower, so that n_split times as many CTAs share the work
  nt /= n_split;
  if (r > 1)
    TC_REQ(r == 8 && c->y.c == 32 && c->y.c_pitch == 32 && c->y.c_off == 0 && nt == 128,
           "conv2d_tc: pixel-shuffle mode needs r=8 and a dense 32-channel output map");
  const size_t yes = elem_size(c->y.dtype);
  // fast path: no channel padding (every accumulator chunk lies inside cout) and 16-byte aligned rows
  bool vec = ((c->y.c_off * yes) % 16) == 0 && ((c->y.c_pitch * yes) % 16) == 0 && ((uintptr_t)c->y.data % 16) == 0;
  if (c->residual.data) {
    const int rg = c->residual_group > 1 ? c->residual_group : 1;
    TC_REQ(r == 1 && view_ok(&c->residual) && (long long)c->residual.n * rg >= c->y.n && c->residual.h == c->y.h &&
               c->residual.w == c->y.w && c->residual.c == c->y.c,
           "conv2d_tc: residual must have the geometry of y (residual_group images of y per residual image)");
    vec = vec && c->residual.dtype == DBSR_BF16 && (c->residual.c_off % 8) == 0 && (c->residual.c_pitch % 8) == 0 &&
          ((uintptr_t)c->residual.data % 16) == 0;
  }
  if (c->bias) vec = vec && ((uintptr_t)c->bias % 16) == 0;
  cfg->align_ok = vec ? 1 : 0;
  vec = vec && cout == cpad;
  cfg->vec_ok = vec ? 1 : 0;
  cfg->n_tile = nt;
  cfg->ck = ck;
  cfg->nchunks = kpad / ck;
  cfg->cout = cout;
  cfg->cout_pad = cpad;
  const int tiles_x = ceil_div(c->x.w, TILE_W);
  const int pad = (c->ksize == 3) ? c->dilation : 0;
  const int budget = 227 * 1024 - (1024 + BAR_BYTES + BIAS_TAB_BYTES) - 8 * STG_WARP_BYTES;
  const int taps = c->ksize * c->ksize;
  // CTA pairs (tcgen05 cta_group::2): the large N = 64 / 128 layers run as clusters of two CTAs that execute every MMA together
  // (M = 256): each CTA fetches its own activations but only HALF of the weight rows, so the per-SM operand fetch of an MMA
  // drops from 4096 + 32 N to 4096 + 16 N bytes (tools/micro/mma_rate.cu: 48 -> 43 clk at N = 64) and the weight tiles take
  // half the shared memory.  Needs two-tile items of a map at least 16 pixels wide and enough of them for every pair.
  // Measured on B200 (profiles/r02_cta_pair_ab.txt): layers with >= 128 input channels (two K chunks, 144 MMAs per item) gain
  // 5-7 % (128->128 + residual 240 -> 226 us, 128->512 767 -> 716 us); 64-channel layers (72 MMAs per item) LOSE (64->64 80 -> 86 us,
  // 64->512 412 -> 441 us): the cross-CTA round trips per item are not amortised.  So pairs are used for K >= 128 only.
  // DBSR_TC_PAIR = 0 / 1 / 2: never / K >= 128 (default) / every eligible layer (A/B switch).
  static const int pair_mode = getenv("DBSR_TC_PAIR") ? atoi(getenv("DBSR_TC_PAIR")) : 1;
  const long long tm_items2 = (long long)c->x.n * ceil_div(c->x.w, 2 * TILE_W) * ceil_div(c->x.h, TILE_H);
  // (1x1 layers with K >= 128 -- the 512 -> 64 projection -- are HBM-bound passes: 187 us paired vs 183 us, left unpaired)
  const bool want_pair = allow_pair && (pair_mode >= 2 || (pair_mode == 1 && kpad / ck >= 2 && c->ksize == 3)) && (nt == 64 || nt == 128) &&
                         tiles_x >= 2 && r == 1 && tm_items2 >= 4 * 148 &&
                         !(c->ksize == 3 && c->dilation == 1 && c->x.h <= 8 && c->x.w <= 8);
  cfg->pair_cta = want_pair ? 1 : 0;
  cfg->b_bytes = (want_pair ? nt / 2 : nt) * ck * 2;
  TC_REQ(cfg->b_bytes % 1024 == 0, "conv2d_tc: internal: unaligned weight stage");
  cfg->rows = TILE_H + 2 * pad;
  // small maps: pack several whole images (with their zero borders) into one M = 128 tile -- "flat" mode
  cfg->flat = 0; cfg->flat_s = 0; cfg->flat_ni = 1;
  // (1x1 kernels need no border: the slots are the pixels themselves, 128 / (h*w) images per tile -- the tap GEMMs of PWC-Net's
  //  flow heads on 1x1 .. 8x8 maps use 100 % of their M rows instead of 1 .. 50 %)
  const int bord = c->ksize == 3 ? 2 : 0;
  static const bool flat_k1 = getenv("DBSR_TC_NO_FLAT_K1") == nullptr;      // A/B switch
  // (only for many images: with a few dozen, one image per tile keeps more CTAs busy -- 2 bursts: 1 503 vs 1 470 bursts/s)
  if (((c->ksize == 3 && c->dilation == 1) || (c->ksize == 1 && flat_k1 && c->x.n >= 2 * 148)) && r == 1 && c->x.h <= 8 && c->x.w <= 8) {
    const int S = (c->x.h + bord) * (c->x.w + bord);
    const int last = (c->x.h - 1) * (c->x.w + bord) + (c->x.w - 1);
    const int ni = (127 - last) / S + 1;
    if (ni >= 2) { cfg->flat = 1; cfg->flat_s = S; cfg->flat_ni = ni; }
  }
  bool found = false;
  cfg->pair_img = 0;
  if (cfg->flat) {
    // two M tiles (2 * ni images) per item share every weight tile when that still leaves work for half the SMs: the
    // CTAs of these tiny-M layers are bound by streaming the whole weight set from L2 once per item
    cfg->mt = (ceil_div(c->x.n, cfg->flat_ni) >= 100 && 2 * 2 * nt <= 512) ? 2 : 1;
    cfg->halo_w = c->x.w + bord; cfg->rows = c->x.h + bord;
    cfg->a_tx_bytes = cfg->mt * cfg->flat_ni * cfg->flat_s * ck * 2;
    // the MMA reads up to slot 127 + 2*halo_w + 2 of its tile (garbage rows beyond the box only feed masked outputs)
    const int a_read = ((cfg->mt - 1) * cfg->flat_ni * cfg->flat_s + 128 + (bord ? 2 * cfg->halo_w + 2 : 0)) * ck * 2;
    cfg->a_bytes = round_up(a_read > cfg->a_tx_bytes ? a_read : cfg->a_tx_bytes, 1024);
    cfg->a_slots = 4;
    while (cfg->a_slots > 1 && cfg->a_slots * cfg->a_bytes + 4 * cfg->b_bytes > budget) cfg->a_slots--;
    found = true;   // flat slots are not tile pixels: per-thread epilogue (16-byte accesses when vec_ok, else masked scalars)
  }
  // item width (mt tiles) and depth of the activation ring: prefer 2 tiles x 4 slots, shrink until the halo boxes
  // leave room for the weight stages (large dilations have large halos)
  // narrow maps (W <= 8, one tile column): pair two consecutive IMAGES in one item instead of two tile columns
  const bool can_pair = tiles_x == 1 && !cfg->flat && c->x.n >= 2 * 148 && !c->residual.data && r == 1 && 2 * 2 * nt <= 512;
  // small workloads (a few bursts per GPU, e.g. BASELINE configs[2] sharded over 8 GPUs): when two-tile items would fill
  // less than two waves of the persistent grid and one-tile items need strictly fewer half-item rounds, split the items
  // (16 x 8 pixels each) so that more SMs take part -- the decoder trunk of 2 bursts is 18 items of 16 x 16 otherwise
  bool split_items = false;
  if (tiles_x >= 2 && !can_pair && !cfg->flat) {
    const long long per = (long long)c->x.n * ceil_div(c->x.h, TILE_H) * (cpad / nt);
    const long long items2 = per * ceil_div(c->x.w, 2 * TILE_W), items1 = per * tiles_x;
    const int sms = 148;
    split_items = items2 < 2 * sms && ceil_div(items1, sms) < 2 * ceil_div(items2, sms);
  }
  for (int mt = (((tiles_x >= 2 && !split_items) || can_pair) ? 2 : 1); mt >= 1 && !found; --mt) {
    for (int slots = 4; slots >= 1 && !found; --slots) {
      const bool pair = can_pair && mt == 2;
      const int hw = TILE_W * (pair ? 1 : mt) + 2 * pad;
      const int ab = round_up(cfg->rows * hw * ck * 2 * (pair ? 2 : 1), 1024);
      const int b_all = ((kpad / ck) * taps + (c->residual.data ? (nt + ck - 1) / ck : 0)) * cfg->b_bytes;
      const bool want_resident = cpad == nt && (kpad / ck) * taps <= 64 && b_all <= budget / 2;
      const int b_need = want_resident ? b_all : 4 * cfg->b_bytes;
      if (slots * ab + b_need <= budget || (slots == 1 && ab + 2 * cfg->b_bytes <= budget)) {
        cfg->mt = mt; cfg->a_slots = slots; cfg->halo_w = hw; cfg->a_bytes = ab;
        cfg->a_tx_bytes = cfg->rows * hw * ck * 2 * (pair ? 2 : 1);
        cfg->pair_img = pair ? 1 : 0;
        found = true;
      }
    }
  }
  TC_REQ(found, "conv2d_tc: activation halo box does not fit in shared memory (dilation %d)", c->dilation);
  if (want_pair && !(cfg->mt == 2 && !cfg->pair_img && !cfg->flat)) return tc_plan(c, cfg, set_err, false, n_split, allow_split);   // pairs need two-tile items
  // accumulator stages: two (double buffering).  DBSR_TC_ACC_STAGES=4 uses four where they fit the 512 TMEM columns (N tile
  // <= 64 with two-tile items): measured on B200 it changes nothing -- neither single CTAs nor CTA pairs wait for a free
  // accumulator stage (profiles/r02_cta_pair_ab.txt) -- so it stays an A/B switch.
  static const int max_acc_stages = getenv("DBSR_TC_ACC_STAGES") ? atoi(getenv("DBSR_TC_ACC_STAGES")) : 2;
  cfg->acc_stages = (max_acc_stages >= 4 && 4 * cfg->mt * nt <= 512) ? 4 : 2;
  int tc = 32;
  while (tc < cfg->acc_stages * cfg->mt * nt) tc <<= 1;
  cfg->tmem_cols = tc;
  // residual on the tensor core: extra K chunks with identity weights (bf16 residual, aligned, same channel chunking)
  // DBSR_TC_RES128_EPI=1: N tile 128 also takes it in the lean epilogue (two 64-channel groups per warp; the second group's
  // rows are prefetched after the first group's store has drained the staging rows)
  static const bool res128_epilogue = getenv("DBSR_TC_RES128_EPI") != nullptr && atoi(getenv("DBSR_TC_RES128_EPI")) != 0;
  cfg->res_chunks = 0;
  if (c->residual.data && !cfg->flat && c->residual.dtype == DBSR_BF16 && (c->residual.c_off % 8) == 0 &&
      (c->residual.c_pitch % 8) == 0 && ((uintptr_t)c->residual.data % 16) == 0 && nt % ck == 0 &&
      !(vec && c->y.dtype == DBSR_BF16 && (nt == 64 || (nt == 128 && res128_epilogue && c->residual_group <= 1)) &&
        cpad * 4 <= BIAS_TAB_BYTES))
    cfg->res_chunks = nt / ck;       // N tile 64 takes the residual in the (prefetching) lean epilogue instead: there the extra
                                     // K chunk costs a whole halo slot of the A ring; at N = 32 the epilogue is the bottleneck and at
                                     // N = 128 the staging area holds half a pixel row, so both keep the tensor-core residual
  // a broadcast residual (one map per group of output images) exists on the tensor-core residual path only
  TC_REQ(!(c->residual.data && c->residual_group > 1) || cfg->res_chunks > 0,
         "conv2d_tc: residual_group needs the tensor-core residual (bf16, 16-byte aligned, N tile 32 or 128)");
  const int b_total = (cfg->nchunks * taps + cfg->res_chunks) * cfg->b_bytes;
  if (cpad == nt && cfg->a_slots * cfg->a_bytes + b_total <= budget && cfg->nchunks * taps + cfg->res_chunks <= 64) {
    cfg->b_resident = 1;
    cfg->b_taps = 1; cfg->b_stage_bytes = cfg->b_bytes;
    cfg->b_stages = cfg->nchunks * taps + cfg->res_chunks;
  } else {
    cfg->b_resident = 0;
    // depth of the weight ring: a (tap, chunk) step is one TMA round trip (~2 us from L2 when few CTAs run) divided by the
    // stages in flight; with 12 stages the small PWC-Net launches spent 0.35 us per step against 0.1 us of MMAs
    // and taps per stage: every ring step costs the single issuing thread a barrier wait + a commit and the producer thread
    // a wait + expect_tx + TMA issue (~0.3 us per step measured on the PWC-Net launches whatever the tile width), so the
    // taps of one (chunk, N tile) share a stage, one 3-D TMA box and one barrier pair while the stage stays <= 24 KB
    static const int stage_cap = getenv("DBSR_TC_BSTAGE_KB") ? atoi(getenv("DBSR_TC_BSTAGE_KB")) * 1024 : 24 * 1024;
    cfg->b_taps = 1;
    for (int t : {9, 3})
      if (taps % t == 0 && t * cfg->b_bytes <= stage_cap && (budget - cfg->a_slots * cfg->a_bytes) / (t * cfg->b_bytes) >= 3) { cfg->b_taps = t; break; }
    cfg->b_stage_bytes = cfg->b_taps * cfg->b_bytes;
    int st = (budget - cfg->a_slots * cfg->a_bytes) / cfg->b_stage_bytes;
    if (st > 32) st = 32;
    TC_REQ(st >= 2, "conv2d_tc: no room for the weight pipeline");
    cfg->b_stages = st;
  }
  int smem = cfg->a_slots * cfg->a_bytes + cfg->b_stages * cfg->b_stage_bytes + 8 * STG_WARP_BYTES + 1024 /*align slack*/ + BAR_BYTES + BIAS_TAB_BYTES;
  // a CTA that owns more than half of TMEM must be alone on its SM: make its smem footprint exclusive too
  if (cfg->tmem_cols > 256 && smem < 120 * 1024) smem = 120 * 1024;
  cfg->smem_bytes = smem;
  // Small grids: PWC-Net's pyramid levels 3..6 of a few bursts are a handful of items (26 pairs of 2x2 maps in flat mode: FOUR),
  // each streaming the layer's whole weight set (up to 1.5 MB) through one SM while 140 SMs idle -- 25-40 us per launch on the
  // critical path of the small-batch regime (profiles/r02_launches_b2.txt).  When the launch would fill less than half of the
  // SMs, the N tile is halved (same packed weights, narrower weight-tile boxes; >= 16 columns) until it does.
  static const bool split_enabled = getenv("DBSR_TC_NO_NSPLIT") == nullptr;      // A/B switch
  if (split_enabled && allow_split && r == 1 && !cfg->pair_cta && !(c->residual.data && c->residual_group > 1)) {
    const int imgs_per_item = cfg->flat ? cfg->flat_ni * cfg->mt : (cfg->pair_img ? 2 : 1);
    const long long tiles = cfg->flat ? 1 : (long long)(cfg->pair_img ? 1 : ceil_div(c->x.w, TILE_W * cfg->mt)) * ceil_div(c->x.h, TILE_H);
    const long long items = (long long)ceil_div(c->x.n, imgs_per_item) * tiles * (cpad / nt);
    if (items * 2 <= 148 && nt % 32 == 0 && n_split < 8) return tc_plan(c, cfg, set_err, allow_pair, n_split * 2, allow_split);
  }
  return 0;
#undef TC_REQ
}

// A 3x3 "same" convolution of a 1x1 map only ever sees its centre tap (the other eight read zero padding): run it as
// the 1x1 convolution with that tap's weight tile (PWC-Net level 6 at 48^2 / 64^2 inputs: 9 launches per forward).
// A 1x1 convolution does not care which pixels are neighbours, so n maps of 1x1 are then viewed as ONE image of
// (n / 8) x 8 pixels (they are contiguous in memory): full 16x8 M tiles instead of one pixel per 128-row tile.
static dbsr_conv_t centre_tap_form(const dbsr_conv_t* c) {
  dbsr_conv_t cc = *c;
  if (c->ksize == 3 && c->x.h == 1 && c->x.w == 1 && c->w && c->shuffle_r <= 1 && c->x.c > 0 && c->y.c > 0) {
    int ck, kpad, nt, cpad;
    tc_geometry(c->x.c, c->y.c, &ck, &kpad, &nt, &cpad);
    cc.ksize = 1; cc.dilation = 1;
    cc.w = reinterpret_cast<const __nv_bfloat16*>(c->w) + (size_t)4 * cpad * kpad;      // packed [tap][cout_pad][kpad]
  }
  if (cc.ksize == 1 && cc.x.h == 1 && cc.x.w == 1 && cc.x.n % 8 == 0 && cc.x.n >= 16 && cc.shuffle_r <= 1 &&
      cc.y.h == 1 && cc.y.w == 1 && cc.y.n == cc.x.n) {
    const int rows = cc.x.n / 8;
    cc.x.n = 1; cc.x.h = rows; cc.x.w = 8;
    cc.y.n = 1; cc.y.h = rows; cc.y.w = 8;
    if (cc.residual.data && cc.residual.h == 1 && cc.residual.w == 1) { cc.residual.n = 1; cc.residual.h = rows; cc.residual.w = 8; }
  }
  return cc;
}

template <int CK, bool RESIDENT, bool PAIR>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mr, const CUtensorMap& mi,
                     const CUtensorMap& my, const ConvTcParams& p, int smem, int grid_limit, cudaStream_t st) {
  const int dev_slot = current_device_slot();
  static int configured_smem_dev[MAX_DEVICES] = {};
  int& configured_smem = configured_smem_dev[dev_slot];
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<CK, RESIDENT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("conv2d_tc: cudaFuncSetAttribute(%d) failed: %s", smem, cudaGetErrorString(e));
      return 2;
    }
    configured_smem = smem;
  }
  static int num_sms_dev[MAX_DEVICES] = {};
  int& num_sms = num_sms_dev[dev_slot];
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int grid = (int)(p.total_items < num_sms ? p.total_items : num_sms);
  if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;
  if (PAIR) {      // clusters of two CTAs, one PAIR item at a time per cluster
    const int pairs_max = (grid_limit > 0 && grid_limit < num_sms ? grid_limit : num_sms) / 2;
    grid = 2 * (int)(p.total_items < pairs_max ? p.total_items : pairs_max);
  }
  // programmatic dependent launch: CTAs may be scheduled (barrier init, tensor-map prefetch, TMEM allocation) as soon as
  // the SMs of the preceding kernel drain; the kernel calls griddepcontrol.wait before its first global access
  static const bool pdl = getenv("DBSR_TC_NO_PDL") == nullptr;     // A/B switch: DBSR_TC_NO_PDL=1 launches normally
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(TC_THREADS); lc.dynamicSmemBytes = (size_t)smem; lc.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  lc.attrs = attr; lc.numAttrs = na;
  // early trigger of the dependent launch (see common.cuh): 0 never, 1 always, default: when half of the SMs stay idle
  static const int early_mode = getenv("DBSR_EARLY_TRIGGER") ? atoi(getenv("DBSR_EARLY_TRIGGER")) : 2;
  ConvTcParams pp = p;
  pp.early_trigger = early_mode == 1 || (early_mode == 2 && grid * 2 <= num_sms);
  cudaError_t le = cudaLaunchKernelEx(&lc, conv_tc_kernel<CK, RESIDENT, PAIR>, mx, mw, mr, mi, my, pp);
  if (le != cudaSuccess) { set_error("conv2d_tc: launch failed: %s", cudaGetErrorString(le)); return 2; }
  return check_launch("conv2d_tc");
}

}  // namespace dbsr

using namespace dbsr;

extern "C" int dbsr_conv2d_tc_geometry(int32_t cin, int32_t cout, int32_t* ck, int32_t* kpad, int32_t* n_tile,
                                       int32_t* cout_pad) {
  DBSR_REQUIRE(cin > 0 && cout > 0 && ck && kpad && n_tile && cout_pad, "conv2d_tc_geometry: bad arguments");
  int a, b, c, d;
  tc_geometry(cin, cout, &a, &b, &c, &d);
  *ck = a; *kpad = b; *n_tile = c; *cout_pad = d;
  return 0;
}

extern "C" int dbsr_conv2d_tc_supported(const dbsr_conv_t* c) {
  TcConfig cfg;
  if (!c) return 0;
  const dbsr_conv_t cc = centre_tap_form(c);
  return tc_plan(&cc, &cfg, false) == 0 ? 1 : 0;
}

static int conv2d_tc_impl(const dbsr_conv_t* c_in, void* stream, const float* pred_w, const float* pred_b, int pred_c,
                          void* pred, int pred_q14) {
  TcConfig cfg;
  DBSR_REQUIRE(c_in != nullptr, "conv2d_tc: null descriptor");
  const dbsr_conv_t cc = centre_tap_form(c_in);
  const dbsr_conv_t* c = &cc;
  if (tc_plan(c, &cfg, true, true, 1, /*allow_split=*/pred == nullptr)) return 1;      // the fused predictor needs all 32 channels in one N tile
  EncodeTiledFn encode = get_encode();
  DBSR_REQUIRE(encode != nullptr, "conv2d_tc: cuTensorMapEncodeTiled entry point not available");

  const int r = c->shuffle_r > 1 ? c->shuffle_r : 1;
  const CUtensorMapSwizzle swz = cfg.ck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

  alignas(64) CUtensorMap mx, mw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)c->x.c, (cuuint64_t)c->x.w, (cuuint64_t)c->x.h, (cuuint64_t)c->x.n};
    cuuint64_t strides[3] = {(cuuint64_t)c->x.c_pitch * 2, (cuuint64_t)c->x.w * c->x.c_pitch * 2,
                             (cuuint64_t)c->x.h * c->x.w * c->x.c_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)cfg.ck, (cuuint32_t)cfg.halo_w, (cuuint32_t)cfg.rows,
                         (cuuint32_t)(cfg.flat ? cfg.flat_ni * cfg.mt : (cfg.pair_img ? 2 : 1))};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* base = reinterpret_cast<__nv_bfloat16*>(c->x.data) + c->x.c_off;
    CUresult rc = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DBSR_REQUIRE(rc == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(x) failed with %d", (int)rc);
  }
  {
    const int taps = c->ksize * c->ksize;
    const int kpad = cfg.nchunks * cfg.ck;
    // packed weights [tap][cout_pad][kpad] as a 3-D tensor: one box = b_taps taps of one (K chunk, N tile)
    cuuint64_t dims[3] = {(cuuint64_t)kpad, (cuuint64_t)cfg.cout_pad, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)kpad * 2, (cuuint64_t)cfg.cout_pad * kpad * 2};
    cuuint32_t box[3] = {(cuuint32_t)cfg.ck, (cuuint32_t)(cfg.pair_cta ? cfg.n_tile / 2 : cfg.n_tile), (cuuint32_t)cfg.b_taps};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = encode(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(c->w), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DBSR_REQUIRE(rc == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(w) failed with %d", (int)rc);
  }

  alignas(64) CUtensorMap mr, mi;
  memset(&mr, 0, sizeof(mr)); memset(&mi, 0, sizeof(mi));
  if (cfg.res_chunks > 0) {
    {
      cuuint64_t dims[4] = {(cuuint64_t)c->residual.c, (cuuint64_t)c->residual.w, (cuuint64_t)c->residual.h, (cuuint64_t)c->residual.n};
      cuuint64_t strides[3] = {(cuuint64_t)c->residual.c_pitch * 2, (cuuint64_t)c->residual.w * c->residual.c_pitch * 2,
                               (cuuint64_t)c->residual.h * c->residual.w * c->residual.c_pitch * 2};
      cuuint32_t box[4] = {(cuuint32_t)cfg.ck, (cuuint32_t)(TILE_W * cfg.mt), (cuuint32_t)TILE_H, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      void* base = reinterpret_cast<__nv_bfloat16*>(c->residual.data) + c->residual.c_off;
      CUresult rc = encode(&mr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      DBSR_REQUIRE(rc == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(residual) failed with %d", (int)rc);
    }
    {
      const void* eye = identity_weights(cfg.n_tile, (cudaStream_t)stream);
      DBSR_REQUIRE(eye != nullptr, "conv2d_tc: could not create the identity weight tile (first call of this N tile inside a stream capture? run the layer once eagerly)");
      cuuint64_t dims[2] = {(cuuint64_t)cfg.n_tile, (cuuint64_t)cfg.n_tile};
      cuuint64_t strides[1] = {(cuuint64_t)cfg.n_tile * 2};
      cuuint32_t box[2] = {(cuuint32_t)cfg.ck, (cuuint32_t)(cfg.pair_cta ? cfg.n_tile / 2 : cfg.n_tile)};
      cuuint32_t es[2] = {1, 1};
      CUresult rc = encode(&mi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(eye), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      DBSR_REQUIRE(rc == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(identity) failed with %d", (int)rc);
    }
  }

  ConvTcParams p;
  p.n = c->x.n; p.H = c->x.h; p.W = c->x.w;
  p.ksize = c->ksize; p.dil = c->dilation;
  p.nchunks = cfg.nchunks; p.cout_pad = cfg.cout_pad; p.cout = cfg.cout; p.n_tile = cfg.n_tile; p.mt = cfg.mt;
  p.tmem_cols = cfg.tmem_cols; p.acc_stages = cfg.acc_stages; p.vec_ok = cfg.vec_ok; p.align_ok = cfg.align_ok;
  p.ntiles_n = cfg.cout_pad / cfg.n_tile;
  p.items_x = cfg.pair_img ? 1 : ceil_div(p.W, TILE_W * cfg.mt); p.tiles_y = ceil_div(p.H, TILE_H);
  p.flat = cfg.flat; p.flat_s = cfg.flat_s; p.flat_ni = cfg.flat_ni;
  p.imgs_per_item = cfg.flat ? cfg.flat_ni * cfg.mt : (cfg.pair_img ? 2 : 1);
  const int row16 = cfg.ck * 2 / 16;                                   // 16-byte units per pixel row of the A box
  if (cfg.flat) { p.t1_step16 = cfg.flat_ni * cfg.flat_s * row16; p.t1_dimg = cfg.flat_ni; p.t1_dx = 0; }
  else if (cfg.pair_img) { p.t1_step16 = cfg.rows * cfg.halo_w * row16; p.t1_dimg = 1; p.t1_dx = 0; }
  else { p.t1_step16 = TILE_W * row16; p.t1_dimg = 0; p.t1_dx = TILE_W; }
  const long long tm_items = (long long)ceil_div(p.n, p.imgs_per_item) * (cfg.flat ? 1 : p.items_x * p.tiles_y);      // per N tile
  p.pair_cta = cfg.pair_cta;
  p.total_items = (cfg.pair_cta ? (tm_items + 1) / 2 : tm_items) * p.ntiles_n;                  // pair_cta: PAIR items
  DBSR_REQUIRE(p.total_items < (1LL << 31) && (long long)c->y.n * c->y.h * c->y.w < (1LL << 31),
               "conv2d_tc: more than 2^31 work items / output pixels");
  {
    // q = umulhi(n, ceil(2^32 / d)) is floor(n / d) for every n with n * d < 2^32
    const unsigned d[3] = {(unsigned)p.ntiles_n, (unsigned)(p.items_x * p.tiles_y), (unsigned)p.items_x};
    unsigned mul[3], one[3];
    unsigned long long dmax = 1;
    for (int i = 0; i < 3; ++i) {
      if (d[i] <= 1) { mul[i] = 0u; one[i] = 0xFFFFFFFFu; }
      else { mul[i] = (unsigned)(((1ULL << 32) + d[i] - 1) / d[i]); one[i] = 0u; }
      if (d[i] > dmax) dmax = d[i];
    }
    p.md_nt = mul[0]; p.md_img = mul[1]; p.md_x = mul[2];
    p.one_nt = one[0]; p.one_img = one[1]; p.one_x = one[2];
    const unsigned long long nmax = (unsigned long long)(p.total_items > tm_items + 2 ? p.total_items : tm_items + 2);   // pairs decode tile indices up to tm_items + 1
    p.fastdiv = (nmax * dmax < (1ULL << 32)) ? 1 : 0;
  }
  // the lean (coalesced, TMA-store) epilogue reads its bias from a shared-memory table.  A layer WITHOUT bias needs no table
  // however wide it is: the decoder's 64 -> 2048 upsampling conv (no bias under ICNR init) fell back to the per-thread
  // store path in round 1 because its 8 KB table did not fit (146 us; ncu: predicated STG.E.128 at the top of the stall list).
  p.bias_fill = (cfg.cout_pad * 4 <= BIAS_TAB_BYTES) ? 1 : 0;
  p.bias_smem = (p.bias_fill || c->bias == nullptr) ? 1 : 0;
  p.a_slots = cfg.a_slots; p.b_stages = cfg.b_stages; p.b_resident = cfg.b_resident;
  p.a_bytes = cfg.a_bytes; p.a_tx_bytes = cfg.a_tx_bytes; p.static_w = (c->flags & DBSR_CONV_STATIC_WEIGHTS) ? 1 : 0;
  p.b_bytes = cfg.b_bytes; p.b_taps = cfg.b_taps; p.b_stage_bytes = cfg.b_stage_bytes; p.halo_w = cfg.halo_w;
  p.y = c->y.data; p.y_dtype = c->y.dtype; p.y_pitch = c->y.c_pitch; p.y_coff = c->y.c_off;
  p.yH = c->y.h; p.yW = c->y.w;
  p.res = c->residual.data; p.r_dtype = c->residual.dtype; p.r_pitch = c->residual.c_pitch; p.r_coff = c->residual.c_off;
  p.res_chunks = cfg.res_chunks;
  p.res_group = c->residual_group > 1 ? c->residual_group : 1;
  p.r_tx_bytes = TILE_H * TILE_W * cfg.mt * cfg.ck * 2;
  if (cfg.res_chunks > 0) p.res = nullptr;   // accumulated by the MMAs, nothing left for the epilogue
  p.bias = c->bias; p.act = c->act; p.shuffle_r = r;
  p.pred = pred; p.pred_c = pred_c; p.pred_q14 = pred_q14 ? 1 : 0;
  memset(p.pred_wb, 0, sizeof(p.pred_wb));
  if (pred != nullptr) {
    DBSR_REQUIRE(pred_w && pred_b && pred_c >= 1 && pred_c <= 4 && cfg.n_tile == 32 && cfg.cout_pad == 32 && !cfg.flat &&
                     r == 1 && p.bias_fill && (cfg.res_chunks > 0 || c->residual.data == nullptr),
                 "conv2d_tc_predictor: needs a 3x3 / 1x1 conv with <= 32 output channels on maps larger than 8x8, "
                 "residual (if any) accumulated on the tensor core, and 1..4 predictor channels");
    for (int k = 0; k < pred_c; ++k) {
      for (int j = 0; j < cfg.cout; ++j) p.pred_wb[k * 32 + j] = pred_w[k * cfg.cout + j];     // HOST arrays
      p.pred_wb[128 + k] = pred_b[k];
    }
  }
  // TMA tensor store of the lean epilogue: bf16 lean path without epilogue-side residual / pixel shuffle, one box shape
  // per launch (every channel group of the N tile has the same width), staging rows on 128-byte lines
  alignas(64) CUtensorMap my;
  memset(&my, 0, sizeof(my));
  p.tma_store = 0;
  {
    const int gw = (cfg.n_tile % 64 == 0) ? 64 : (cfg.n_tile == 32 ? 32 : (cfg.n_tile == 16 ? 16 : 0));
    const size_t stg_off = (size_t)cfg.a_slots * cfg.a_bytes + (size_t)cfg.b_stages * cfg.b_stage_bytes;
    static const bool enabled = getenv("DBSR_TC_NO_TMA_STORE") == nullptr;     // A/B switch: =1 keeps the LDS + STG read-back
    // (an epilogue-side residual -- N tile 64 -- is prefetched into the same staging rows and is compatible)
    const bool ok = enabled && gw && pred == nullptr && !cfg.flat && p.bias_smem && cfg.vec_ok && c->y.dtype == DBSR_BF16 &&
                    (p.res == nullptr || cfg.n_tile == 64 || cfg.n_tile == 32 || cfg.n_tile == 16) && stg_off % 128 == 0;
    const CUtensorMapSwizzle sw = gw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (gw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    void* base = reinterpret_cast<__nv_bfloat16*>(c->y.data) + c->y.c_off;
    if (ok && r == 1) {
      cuuint64_t dims[4] = {(cuuint64_t)c->y.c, (cuuint64_t)c->y.w, (cuuint64_t)c->y.h, (cuuint64_t)c->y.n};
      cuuint64_t strides[3] = {(cuuint64_t)c->y.c_pitch * 2, (cuuint64_t)c->y.w * c->y.c_pitch * 2,
                               (cuuint64_t)c->y.h * c->y.w * c->y.c_pitch * 2};
      cuuint32_t box[4] = {(cuuint32_t)gw, (cuuint32_t)TILE_W, 4u, 1u};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult rc = encode(&my, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      DBSR_REQUIRE(rc == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(y) failed with %d", (int)rc);
      p.tma_store = 1;
    } else if (ok && r == 8 && gw == 64) {
      // pixel shuffle: y is the dense [n, 8H, 8W, 32] map; element (n, 8y + i, 8x + j, c) is addressed as the 5-D tensor
      // {(j, c): 256, x: W, i: 8, y: H, n}; a warp's 32 LR pixels x 64 packed channels are the box {64, 8, 1, 4, 1}
      const cuuint64_t W = (cuuint64_t)c->x.w, H = (cuuint64_t)c->x.h;
      cuuint64_t dims[5] = {256, W, 8, H, (cuuint64_t)c->y.n};
      cuuint64_t strides[4] = {512, 8 * W * 64, 8 * 8 * W * 64, 8 * H * 8 * W * 64};
      cuuint32_t box[5] = {64u, (cuuint32_t)TILE_W, 1u, 4u, 1u};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult rc = encode(&my, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      DBSR_REQUIRE(rc == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(y, pixel shuffle) failed with %d", (int)rc);
      p.tma_store = 1;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int gl = c_in->grid_limit;
  if (cfg.pair_cta) {
    if (cfg.ck == 64) return cfg.b_resident ? launch_tc<64, true, true>(mx, mw, mr, mi, my, p, cfg.smem_bytes, gl, st)
                                            : launch_tc<64, false, true>(mx, mw, mr, mi, my, p, cfg.smem_bytes, gl, st);
    return cfg.b_resident ? launch_tc<32, true, true>(mx, mw, mr, mi, my, p, cfg.smem_bytes, gl, st)
                          : launch_tc<32, false, true>(mx, mw, mr, mi, my, p, cfg.smem_bytes, gl, st);
  }
  if (cfg.ck == 64) return cfg.b_resident ? launch_tc<64, true, false>(mx, mw, mr, mi, my, p, cfg.smem_bytes, gl, st)
                                          : launch_tc<64, false, false>(mx, mw, mr, mi, my, p, cfg.smem_bytes, gl, st);
  return cfg.b_resident ? launch_tc<32, true, false>(mx, mw, mr, mi, my, p, cfg.smem_bytes, gl, st)
                        : launch_tc<32, false, false>(mx, mw, mr, mi, my, p, cfg.smem_bytes, gl, st);
}

extern "C" int dbsr_conv2d_tc(const dbsr_conv_t* c, void* stream) {
  return conv2d_tc_impl(c, stream, nullptr, nullptr, 0, nullptr, 0);
}

extern "C" int dbsr_conv2d_tc_predictor(const dbsr_conv_t* c, const float* pred_w, const float* pred_b, int32_t pred_c,
                                        void* pred, int32_t pred_q14, void* stream) {
  DBSR_REQUIRE(pred != nullptr, "conv2d_tc_predictor: null output");
  return conv2d_tc_impl(c, stream, pred_w, pred_b, pred_c, pred, pred_q14);
}

