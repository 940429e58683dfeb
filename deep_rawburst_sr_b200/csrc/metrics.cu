// Image-quality metrics of the DBSR evaluation path as single-pass kernels:
//   SSIM / MS-SSIM  (reference models/loss/msssim.py:22-104, models/loss/image_quality_v2.py:104-136)
//   per-image MSE -> PSNR  (reference models/loss/image_quality_v2.py:47-101)
// Inputs are the NCHW fp32 tensors of the module seam (`pred`, `gt`); `boundary_ignore` is index arithmetic inside the
// kernels (no cropped copies); every reduction is two-stage with a fixed order (per-CTA partials -> one CTA per image),
// so the results are bit-identical from run to run and across batch compositions.
#include "common.cuh"

namespace dbsr {

constexpr int SS_TW = 32;             // output tile: 32 x 32 window positions per CTA
constexpr int SS_TH = 32;
constexpr int SS_MAXW = 11;           // msssim.py window_size (real_size = min(11, h, w) on small maps)
constexpr int SS_IN_H = SS_TH + SS_MAXW - 1;   // 42 input rows / cols per tile
constexpr int SS_IN_W = SS_TW + SS_MAXW - 1;
constexpr int SS_PITCH = 45;          // input-tile row pitch: 4 rows x 8 strips of one warp hit 32 distinct banks
constexpr int SS_THREADS = 256;
constexpr int RANGE_BLOCKS = 592;     // per-CTA (max, min) partials of the data-dependent value range (4 CTAs per SM)

struct SsimParams {
  const float* a;          // img1 [planes, H, W]
  const float* b;          // img2
  float* map;              // optional [planes, oh, ow] ssim map (spatial_out), may be null
  float* partial;          // [planes * tiles_y * tiles_x][2]  (sum ssim, sum cs) per CTA
  const float* range_ws;   // [RANGE_BLOCKS][2] (max, min) of img1, read when val_range <= 0
  const unsigned char* valid;   // optional [n, 1, H, W] mask: sums become (sum ssim * valid, sum valid), see dbsr_ssim
  int c;                   // channels per image (plane -> image for the mask)
  int H, W, crop, oh, ow;
  float val_range;
  float g[SS_MAXW];        // 1-D Gaussian (sigma 1.5, normalised), zero beyond the real window size
};

// msssim.py:24-35: L = (255 if max(img1) > 128 else 1) - (-1 if min(img1) < -0.5 else 0), over the CROPPED img1
__global__ void __launch_bounds__(256) value_range_kernel(const float* __restrict__ a, int planes, int H, int W, int crop,
                                                          float* __restrict__ range_ws) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int hc = H - 2 * crop, wc = W - 2 * crop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY, mn = INFINITY;
  const long long rows = (long long)planes * hc;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
    const int pl = (int)(r / hc), y = (int)(r - (long long)pl * hc) + crop;
    const float* row = a + ((long long)pl * H + y) * W + crop;
    for (int x = lane; x < wc; x += 32) {
      const float v = __ldg(row + x);
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
  }
  __shared__ float smx[8], smn[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) { smx[warp] = mx; smn[warp] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { mx = fmaxf(mx, smx[i]); mn = fminf(mn, smn[i]); }
    range_ws[2 * blockIdx.x] = mx;
    range_ws[2 * blockIdx.x + 1] = mn;
  }
}

// One CTA = one 32 x 32 tile of window positions of one (image, channel) plane.  The 42 x 42 input patch of both images
// is staged in shared memory once; the 11 x 11 Gaussian window is applied separably to the five moments
// (x, y, x^2, y^2, xy): horizontal pass (4 adjacent columns per thread, sliding registers) -> shared memory ->
// vertical pass (4 adjacent rows per thread), then the SSIM / contrast terms and the CTA's partial sums.
// Algorithmic traffic: 8 B per pixel read (both images once), 8 B per CTA written; ~2.5 k FMA per output pixel-column,
// i.e. the kernel sits above the FP32 ridge (27 FLOP/B) and is bound by the FMA pipe, not by HBM.
__global__ void __launch_bounds__(SS_THREADS) ssim_tile_kernel(const SsimParams p) {
  __shared__ float sa[SS_IN_H][SS_PITCH];
  __shared__ float sb[SS_IN_H][SS_PITCH];
  __shared__ __align__(16) float sh[5][SS_IN_H][SS_TW];
  __shared__ float red[2][SS_THREADS / 32];
  __shared__ float s_c[2];

  griddep_launch_dependents_if_small();
  griddep_wait();
  const int tid = threadIdx.x;
  const int plane = blockIdx.z;
  const int y0 = blockIdx.y * SS_TH, x0 = blockIdx.x * SS_TW;     // tile origin in window positions (cropped coordinates)
  const int hc = p.H - 2 * p.crop, wc = p.W - 2 * p.crop;
  const float* pa = p.a + ((long long)plane * p.H + p.crop) * p.W + p.crop;
  const float* pb = p.b + ((long long)plane * p.H + p.crop) * p.W + p.crop;

  // C1 / C2 (msssim.py:54-55): python doubles, rounded to fp32 when they meet the tensors
  float L = p.val_range;
  if (L <= 0.0f) {                                 // uniform branch: the whole CTA reduces the (max, min) partials
    float mx = -INFINITY, mn = INFINITY;
    for (int i = tid; i < RANGE_BLOCKS; i += SS_THREADS) {
      mx = fmaxf(mx, p.range_ws[2 * i]);
      mn = fminf(mn, p.range_ws[2 * i + 1]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = mx; red[1][tid >> 5] = mn; }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < SS_THREADS / 32; ++i) { mx = fmaxf(mx, red[0][i]); mn = fminf(mn, red[1][i]); }
      L = (mx > 128.0f ? 255.0f : 1.0f) - (mn < -0.5f ? -1.0f : 0.0f);
    }
  }
  if (tid == 0) {
    const double k1 = 0.01 * (double)L, k2 = 0.03 * (double)L;
    s_c[0] = (float)(k1 * k1);
    s_c[1] = (float)(k2 * k2);
  }
  __syncthreads();                                 // `red` is reused by the final reduction

  for (int i = tid; i < SS_IN_H * SS_IN_W; i += SS_THREADS) {
    const int r = i / SS_IN_W, c = i - r * SS_IN_W;
    const int y = y0 + r, x = x0 + c;
    float va = 0.0f, vb = 0.0f;
    if (y < hc && x < wc) {
      va = __ldg(pa + (long long)y * p.W + x);
      vb = __ldg(pb + (long long)y * p.W + x);
    }
    sa[r][c] = va;
    sb[r][c] = vb;
  }
  __syncthreads();

  // horizontal pass: item = (row r, strip s) -> columns 4s .. 4s+3, reads columns 4s .. 4s+13
  for (int it = tid; it < SS_IN_H * (SS_TW / 4); it += SS_THREADS) {
    const int r = it >> 3, s = it & 7;
    float acc[5][4];
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][j] = 0.0f;
#pragma unroll
    for (int k = 0; k < SS_MAXW + 3; ++k) {
      const float va = sa[r][4 * s + k], vb = sb[r][4 * s + k];
      const float aa = va * va, bb = vb * vb, ab = va * vb;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = k - j;                       // tap index of input column k for output column j
        if (t >= 0 && t < SS_MAXW) {
          const float g = p.g[t];
          acc[0][j] = fmaf(g, va, acc[0][j]);
          acc[1][j] = fmaf(g, vb, acc[1][j]);
          acc[2][j] = fmaf(g, aa, acc[2][j]);
          acc[3][j] = fmaf(g, bb, acc[3][j]);
          acc[4][j] = fmaf(g, ab, acc[4][j]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q)
      *reinterpret_cast<float4*>(&sh[q][r][4 * s]) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
  }
  __syncthreads();

  // vertical pass: thread = (column x, row strip j) -> rows 4j .. 4j+3
  const int cx = tid & 31, rs = tid >> 5;
  float out[5][4];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[q][j] = 0.0f;
#pragma unroll
    for (int k = 0; k < SS_MAXW + 3; ++k) {
      const float v = sh[q][4 * rs + k][cx];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = k - j;
        if (t >= 0 && t < SS_MAXW) out[q][j] = fmaf(p.g[t], v, out[q][j]);
      }
    }
  }
  const float C1 = s_c[0], C2 = s_c[1];
  float sum_ssim = 0.0f, sum_cs = 0.0f;
  const int ox = x0 + cx;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int oy = y0 + 4 * rs + j;
    if (oy < p.oh && ox < p.ow) {
      // msssim.py:47-63, one rounding per reference op (no contraction into FMAs)
      const float mu1 = out[0][j], mu2 = out[1][j];
      const float mu1_sq = __fmul_rn(mu1, mu1), mu2_sq = __fmul_rn(mu2, mu2), mu12 = __fmul_rn(mu1, mu2);
      const float s1 = __fsub_rn(out[2][j], mu1_sq), s2 = __fsub_rn(out[3][j], mu2_sq), s12 = __fsub_rn(out[4][j], mu12);
      const float v1 = __fadd_rn(__fmul_rn(2.0f, s12), C2);
      const float v2 = __fadd_rn(__fadd_rn(s1, s2), C2);
      const float cs = __fdiv_rn(v1, v2);
      const float num = __fmul_rn(__fadd_rn(__fmul_rn(2.0f, mu12), C1), v1);
      const float den = __fmul_rn(__fadd_rn(__fadd_rn(mu1_sq, mu2_sq), C1), v2);
      const float ss = __fdiv_rn(num, den);
      if (p.valid) {      // image_quality_v2.py:127-131: the mask is cropped like the images, then by 5 (window 11)
        const float m = p.valid[((long long)(plane / p.c) * p.H + p.crop + 5 + oy) * p.W + p.crop + 5 + ox] ? 1.0f : 0.0f;
        sum_ssim += ss * m;
        sum_cs += m;
      } else {
        sum_ssim += ss;
        sum_cs += cs;
      }
      if (p.map) p.map[((long long)plane * p.oh + oy) * p.ow + ox] = ss;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sum_ssim += __shfl_xor_sync(0xffffffffu, sum_ssim, o);
    sum_cs += __shfl_xor_sync(0xffffffffu, sum_cs, o);
  }
  if (cx == 0) { red[0][rs] = sum_ssim; red[1][rs] = sum_cs; }
  __syncthreads();
  if (tid == 0) {
    float a = 0.0f, c = 0.0f;
    for (int i = 0; i < SS_THREADS / 32; ++i) { a += red[0][i]; c += red[1][i]; }
    const long long cta = ((long long)plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    p.partial[2 * cta] = a;
    p.partial[2 * cta + 1] = c;
  }
}

// second stage of every metric reduction: out[image][v] = scale * sum_i partial[image][i][v], fixed order, fp64
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int per_image, int k, double scale,
                                                              float* __restrict__ out) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  __shared__ double sm[256];
  const float* base = partial + (long long)blockIdx.x * per_image * k;
  for (int v = 0; v < k; ++v) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < per_image; i += 256) acc += (double)base[(long long)i * k + v];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[(long long)blockIdx.x * k + v] = (float)(sm[0] * scale);
    __syncthreads();
  }
}

// F.avg_pool2d(img, (2, 2)) of both images between the MS-SSIM levels (msssim.py:88-89); floor output size
__global__ void __launch_bounds__(256) avgpool2_pair_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ oa,
                                                            float* __restrict__ ob, int planes, int H, int W) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int oh = H / 2, ow = W / 2;
  const long long total = (long long)planes * oh * ow;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % ow);
    const long long t = i / ow;
    const int y = (int)(t % oh);
    const long long pl = t / oh;
    const long long src = (pl * H + 2 * y) * W + 2 * x;
    const float2 a0 = make_float2(__ldg(a + src), __ldg(a + src + 1)), a1 = make_float2(__ldg(a + src + W), __ldg(a + src + W + 1));
    const float2 b0 = make_float2(__ldg(b + src), __ldg(b + src + 1)), b1 = make_float2(__ldg(b + src + W), __ldg(b + src + W + 1));
    oa[i] = (((a0.x + a0.y) + a1.x) + a1.y) * 0.25f;
    ob[i] = (((b0.x + b0.y) + b1.x) + b1.y) * 0.25f;
  }
}

// per-image sum of squared differences over the cropped planes (image_quality_v2.py:47-66 with metric 'l2', valid=None)
__global__ void __launch_bounds__(256) sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, const unsigned char* __restrict__ valid,
                                                     int c, int H, int W, int crop, float* __restrict__ partial) {
  griddep_launch_dependents_if_small();
  griddep_wait();
  const int hc = H - 2 * crop, wc = W - 2 * crop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int image = blockIdx.y;
  const int rows = c * hc;
  float acc = 0.0f, cnt = 0.0f;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const int pl = r / hc, y = r - pl * hc + crop;
    const long long off = (((long long)image * c + pl) * H + y) * W + crop;
    const unsigned char* vrow = valid ? valid + ((long long)image * H + y) * W + crop : nullptr;
    for (int x = lane; x < wc; x += 32) {
      const float d = __ldg(a + off + x) - __ldg(b + off + x);
      if (vrow) {
        const float m = vrow[x] ? 1.0f : 0.0f;
        acc = fmaf(d * d, m, acc);
        cnt += m;
      } else {
        acc = fmaf(d, d, acc);
      }
    }
  }
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { red[0][warp] = acc; red[1][warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f, n = 0.0f;
    for (int i = 0; i < 8; ++i) { s += red[0][i]; n += red[1][i]; }
    const long long slot = (long long)image * gridDim.x + blockIdx.x;
    if (valid) { partial[2 * slot] = s; partial[2 * slot + 1] = n; }
    else partial[slot] = s;
  }
}

}  // namespace dbsr

using namespace dbsr;

extern "C" int dbsr_ssim_workspace_floats(int32_t n, int32_t c, int32_t h, int32_t w, int32_t crop, int32_t window) {
  if (n <= 0 || c <= 0 || window < 1 || window > SS_MAXW || crop < 0) return -1;
  const int oh = h - 2 * crop - window + 1, ow = w - 2 * crop - window + 1;
  if (oh <= 0 || ow <= 0) return -1;
  const long long ctas = (long long)n * c * ceil_div(oh, SS_TH) * ceil_div(ow, SS_TW);
  const long long floats = 2 * ctas + 2 * RANGE_BLOCKS;
  return floats < (1ll << 31) ? (int)floats : -1;
}

extern "C" int dbsr_ssim(const float* img1, const float* img2, int32_t n, int32_t c, int32_t h, int32_t w, int32_t crop,
                         const float* window1d, int32_t window, float val_range, const uint8_t* valid, float* workspace, float* stats,
                         float* ssim_map, void* stream) {
  DBSR_REQUIRE(img1 && img2 && window1d && workspace && stats, "ssim: null argument");
  DBSR_REQUIRE(n > 0 && c > 0 && crop >= 0 && window >= 1 && window <= SS_MAXW, "ssim: bad geometry (window must be 1..11)");
  const int hc = h - 2 * crop, wc = w - 2 * crop;
  const int oh = hc - window + 1, ow = wc - window + 1;
  DBSR_REQUIRE(oh > 0 && ow > 0, "ssim: image %dx%d (crop %d) smaller than the %d-tap window", h, w, crop, window);
  DBSR_REQUIRE(!valid || window == SS_MAXW, "ssim: a valid mask assumes the 11-tap window (the reference crops it by 5)");
  const int tx = ceil_div(ow, SS_TW), ty = ceil_div(oh, SS_TH);
  DBSR_REQUIRE((long long)n * c <= 65535 && ty <= 65535, "ssim: too many planes / tiles for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  float* range_ws = workspace;
  float* partial = workspace + 2 * RANGE_BLOCKS;
  if (val_range <= 0.0f)
    launch_pdl(value_range_kernel, dim3(RANGE_BLOCKS), dim3(256), 0, st, img1, n * c, h, w, crop, range_ws);
  SsimParams p;
  p.a = img1; p.b = img2; p.map = ssim_map; p.partial = partial; p.range_ws = range_ws;
  p.H = h; p.W = w; p.crop = crop; p.oh = oh; p.ow = ow; p.val_range = val_range; p.valid = valid; p.c = c;
  for (int i = 0; i < SS_MAXW; ++i) p.g[i] = i < window ? window1d[i] : 0.0f;
  launch_pdl(ssim_tile_kernel, dim3(tx, ty, n * c), dim3(SS_THREADS), 0, st, p);
  launch_pdl(reduce_partials_kernel, dim3(n), dim3(256), 0, st, (const float*)partial, c * ty * tx, 2,
             valid ? 1.0 : 1.0 / ((double)c * oh * ow), stats);
  return check_launch("ssim");
}

extern "C" int dbsr_avgpool2_pair(const float* img1, const float* img2, float* out1, float* out2, int32_t planes, int32_t h, int32_t w,
                                  void* stream) {
  DBSR_REQUIRE(img1 && img2 && out1 && out2 && planes > 0 && h >= 2 && w >= 2, "avgpool2_pair: bad arguments");
  const long long total = (long long)planes * (h / 2) * (w / 2);
  long long blocks = (total + 255) / 256;
  const int grid = (int)(blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks));
  launch_pdl(avgpool2_pair_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, img1, img2, out1, out2, planes, h, w);
  return check_launch("avgpool2_pair");
}

extern "C" int dbsr_mse_workspace_floats(int32_t n) { return n > 0 ? n * 148 * 2 : -1; }

extern "C" int dbsr_mse_per_image(const float* pred, const float* gt, const uint8_t* valid, int32_t n, int32_t c, int32_t h, int32_t w,
                                  int32_t crop, float* workspace, float* mse, void* stream) {
  DBSR_REQUIRE(pred && gt && workspace && mse && n > 0 && n <= 65535 && c > 0 && crop >= 0, "mse_per_image: bad arguments");
  const int hc = h - 2 * crop, wc = w - 2 * crop;
  DBSR_REQUIRE(hc > 0 && wc > 0, "mse_per_image: boundary_ignore %d leaves nothing of a %dx%d image", crop, h, w);
  const int rows = c * hc;
  int per = ceil_div(rows, 8);
  if (per > 148) per = 148;
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl(sq_err_kernel, dim3(per, n), dim3(256), 0, st, pred, gt, (const unsigned char*)valid, c, h, w, crop, workspace);
  if (valid)      // mse[n][2] = (sum valid * err, sum valid over pixels and channels): the caller forms the reference's ratio
    launch_pdl(reduce_partials_kernel, dim3(n), dim3(256), 0, st, (const float*)workspace, per, 2, 1.0, mse);
  else
    launch_pdl(reduce_partials_kernel, dim3(n), dim3(256), 0, st, (const float*)workspace, per, 1, 1.0 / ((double)c * hc * wc), mse);
  return check_launch("mse_per_image");
}
