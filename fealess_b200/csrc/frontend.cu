// LINE-MOD modality extraction on sm_100a: colour-gradient quantisation, depth-normal quantisation, pyramids, and the
// fused spread -> response-map -> linear-memory kernel.  Integer/byte work, HBM-bound at best and launch-bound at VGA:
// every stage is one tiled kernel with a shared-memory halo so each input byte is read from HBM once.
//
// Reference semantics (paths relative to /root/reference):
//   colour : quantizedOrientations + hysteresisGradient      linemod/linemod.cpp:230-385
//   pyramid: ColorGradientPyramid::pyrDown (cv::pyrDown)      linemod.cpp:434-453
//   depth  : quantizedNormals (+ medianBlur 5)                linemod.cpp:567-685, table linemod/normal_lut.i
//   depth pyramid / masks: resize(INTER_NEAREST), copyTo(mask) linemod.cpp:448, 455-459, 721-745
//   spread / computeResponseMaps / linearize                  linemod.cpp:950-965, 979-1048, 1060-1088
// Compiled with -fmad=false: the fp32 expressions below must round exactly like the (uncontracted) reference.
#include "fl_internal.cuh"
#include <math.h>

// ------------------------------------------------------------------------------------------------
// tables: NORMAL_LUT depends only on (v2, v1) -> 400 bytes; response table packs the 8 per-label responses of one
// spread byte into a uint2 (labels 0-3 in .x, 4-7 in .y), i.e. SIMILARITY_LUT (linemod.cpp:970) re-indexed by pixel value.
// ------------------------------------------------------------------------------------------------
__constant__ uint8_t c_normal_plane[400];
__constant__ uint2 c_resp8[256];
__device__ uint2 g_resp8[256];          // same table in global memory: 256 threads copy it to shared memory with ONE coalesced load each
                                        // (the constant cache serialises a warp's 32 distinct addresses)

int fl_launch_tables_init() {
  uint8_t plane[400];
  for (int v2 = 0; v2 < 20; ++v2)
    for (int v1 = 0; v1 < 20; ++v1) {
      // one-hot 45-degree sector of atan2(v2-10, v1-10) + 22.5 deg: closed form of normal_lut.i:4 (sha256-pinned in tests)
      double a = atan2((double)(v2 - 10), (double)(v1 - 10)) * (180.0 / 3.14159265358979323846) + 22.5;
      a = fmod(a + 360.0, 360.0);
      plane[v2 * 20 + v1] = (uint8_t)(1u << (((int)floor(a / 45.0)) & 7));
    }
  uint2 resp[256];
  for (int v = 0; v < 256; ++v) {
    uint32_t w[2] = {0, 0};
    for (int i = 0; i < 8; ++i) {
      int best = 0;
      for (int j = 0; j < 8; ++j)
        if (v >> j & 1) {
          int d = i > j ? i - j : j - i;
          if (8 - d < d) d = 8 - d;
          int g = d == 0 ? 4 : (d == 1 ? 2 : (d == 2 ? 1 : 0));   // response per circular label distance
          if (g > best) best = g;
        }
      w[i >> 2] |= (uint32_t)best << (8 * (i & 3));
    }
    resp[v] = make_uint2(w[0], w[1]);
  }
  FL_CUDA(cudaMemcpyToSymbol(c_normal_plane, plane, sizeof plane));
  FL_CUDA(cudaMemcpyToSymbol(c_resp8, resp, sizeof resp));
  FL_CUDA(cudaMemcpyToSymbol(g_resp8, resp, sizeof resp));
  return FL_OK;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ------------------------------------------------------------------------------------------------
// K1 colour quantisation: blur7 -> sobel3 -> strongest channel -> 16-bin angle -> &7 -> 3x3 majority vote, one kernel.
// Tile = CQ_TW x CQ_TH output pixels; halo 5 = 3 (blur) + 1 (sobel) + 1 (vote).
// ------------------------------------------------------------------------------------------------
#define CQ_TW 32
#define CQ_TH 16
#define CQ_THREADS 256

// OpenCV fastAtan2 (cv::phase, angleInDegrees) in fp32, then convertTo(CV_8U, 16/360) = rint-half-even, saturate.
__device__ __forceinline__ int angle_q16(float x, float y) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float eps = 2.220446049250313e-16f;
  // the two octant branches of fastAtan2 evaluate the same polynomial on min/max: one division instead of a divergent pair
  const float ax = fabsf(x), ay = fabsf(y);
  const float c = __fdiv_rn(fminf(ax, ay), __fadd_rn(fmaxf(ax, ay), eps));
  const float c2 = __fmul_rn(c, c);
  float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  if (ax < ay) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  int r = __float2int_rn(__fmul_rn(a, (float)(16.0 / 360.0)));
  return min(max(r, 0), 255);
}

#define CQ_SMEM_BYTES ((CQ_TH + 10) * (CQ_TW + 10) * 3 + (CQ_TH + 10) * (CQ_TW + 4) * 3 * 2 + (CQ_TH + 4) * (CQ_TW + 4) * 3 + (CQ_TH + 2) * (CQ_TW + 2) + 16)

// one CQ_TW x CQ_TH tile; smem = CQ_SMEM_BYTES bytes, 16-byte aligned
__device__ __forceinline__ void dev_color_quantize(const uint8_t* __restrict__ bgr, int W, int H, float thr_sq, uint8_t* __restrict__ q,
                                                   int bx, int by, uint8_t* smem) {
  constexpr int SW = CQ_TW + 10, SH = CQ_TH + 10;   // source tile
  constexpr int BW = CQ_TW + 4, BH = CQ_TH + 4;     // blurred tile (positions x0-2 .. x0+TW+1)
  constexpr int QW = CQ_TW + 2, QH = CQ_TH + 2;     // unfiltered-bin tile (positions x0-1 .. x0+TW)
  constexpr int OFF_H = (SH * SW * 3 + 15) & ~15;
  uint8_t(*s_src)[SW * 3] = reinterpret_cast<uint8_t(*)[SW * 3]>(smem);
  uint16_t(*s_h)[BW * 3] = reinterpret_cast<uint16_t(*)[BW * 3]>(smem + OFF_H);
  uint8_t(*s_b)[BW * 3] = reinterpret_cast<uint8_t(*)[BW * 3]>(smem + OFF_H + SH * BW * 3 * 2);
  uint8_t(*s_q)[QW] = reinterpret_cast<uint8_t(*)[QW]>(smem + OFF_H + SH * BW * 3 * 2 + BH * BW * 3);   // bits 0-2 bin, bit 7 = magnitude above threshold
  const int x0 = bx * CQ_TW, y0 = by * CQ_TH;
  const int tid = threadIdx.x;

  // 1. source tile with BORDER_REPLICATE
  for (int i = tid; i < SH * SW; i += CQ_THREADS) {
    int r = i / SW, c = i - r * SW;
    int sy = clampi(y0 - 5 + r, 0, H - 1), sx = clampi(x0 - 5 + c, 0, W - 1);
    const uint8_t* p = bgr + ((size_t)sy * W + sx) * 3;
    s_src[r][c * 3 + 0] = p[0]; s_src[r][c * 3 + 1] = p[1]; s_src[r][c * 3 + 2] = p[2];
  }
  __syncthreads();
  // 2. horizontal 7-tap {8,28,56,72,56,28,8}: exact integers (<= 65280).  A blurred position outside the image stands
  //    for the replicated border pixel of the BLURRED image (Sobel's own BORDER_REPLICATE), so it is evaluated at the
  //    clamped position.
  for (int i = tid; i < SH * BW; i += CQ_THREADS) {
    int r = i / BW, bx = i - r * BW;
    int cpx = clampi(x0 - 2 + bx, 0, W - 1);
    int c0 = cpx - 3 - (x0 - 5);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const uint8_t* p = &s_src[r][c0 * 3 + ch];
      int s = 8 * (p[0] + p[18]) + 28 * (p[3] + p[15]) + 56 * (p[6] + p[12]) + 72 * p[9];
      s_h[r][bx * 3 + ch] = (uint16_t)s;
    }
  }
  __syncthreads();
  // 3. vertical 7-tap, single rounding (sum + 2^15) >> 16
  for (int i = tid; i < BH * BW * 3; i += CQ_THREADS) {
    int by = i / (BW * 3), k = i - by * (BW * 3);
    int cpy = clampi(y0 - 2 + by, 0, H - 1);
    int r0 = cpy - 3 - (y0 - 5);
    int s = 8 * (s_h[r0][k] + s_h[r0 + 6][k]) + 28 * (s_h[r0 + 1][k] + s_h[r0 + 5][k]) + 56 * (s_h[r0 + 2][k] + s_h[r0 + 4][k]) +
            72 * s_h[r0 + 3][k];
    s_b[by][k] = (uint8_t)((s + 32768) >> 16);
  }
  __syncthreads();
  // 4. Sobel 3x3 per channel on the blurred tile, strongest channel (ties: first, linemod.cpp:275-292), angle bin
  for (int i = tid; i < QH * QW; i += CQ_THREADS) {
    int qy = i / QW, qx = i - qy * QW;
    int sx = x0 - 1 + qx, sy = y0 - 1 + qy;
    uint8_t v = 0;
    if (sx > 0 && sx < W - 1 && sy > 0 && sy < H - 1) {   // the 1-px frame of quantized_unfiltered is zeroed (:318-325)
      int bx = qx + 1, by = qy + 1;                        // blurred-tile index of this position
      int best_m = -1, best_dx = 0, best_dy = 0;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        int a00 = s_b[by - 1][(bx - 1) * 3 + ch], a01 = s_b[by - 1][bx * 3 + ch], a02 = s_b[by - 1][(bx + 1) * 3 + ch];
        int a10 = s_b[by][(bx - 1) * 3 + ch], a12 = s_b[by][(bx + 1) * 3 + ch];
        int a20 = s_b[by + 1][(bx - 1) * 3 + ch], a21 = s_b[by + 1][bx * 3 + ch], a22 = s_b[by + 1][(bx + 1) * 3 + ch];
        int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
        int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
        int m = dx * dx + dy * dy;
        if (m > best_m) { best_m = m; best_dx = dx; best_dy = dy; }   // strict > keeps the earliest channel on ties
      }
      v = (uint8_t)(angle_q16((float)best_dx, (float)best_dy) & 7);
      if ((float)best_m > thr_sq) v |= 0x80;
    }
    s_q[qy][qx] = v;
  }
  __syncthreads();
  // 5. 3x3 histogram vote: majority (>= 5 of 9) of the bins, lowest bin wins ties (:346-381)
  for (int i = tid; i < CQ_TH * CQ_TW; i += CQ_THREADS) {
    int oy = i / CQ_TW, ox = i - oy * CQ_TW;
    int x = x0 + ox, y = y0 + oy;
    if (x >= W || y >= H) continue;
    uint8_t out = 0;
    if (x > 0 && x < W - 1 && y > 0 && y < H - 1 && (s_q[oy + 1][ox + 1] & 0x80)) {
      uint32_t hist = 0;   // eight 4-bit counters
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) hist += 1u << (4 * (s_q[oy + j][ox + k] & 7));
      int best = 0, idx = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        int c = (hist >> (4 * b)) & 15;
        if (c > best) { best = c; idx = b; }
      }
      if (best >= 5) out = (uint8_t)(1u << idx);
    }
    q[(size_t)y * W + x] = out;
  }
}

__global__ void __launch_bounds__(CQ_THREADS) k_color_quantize(const uint8_t* __restrict__ bgr, int W, int H, float thr_sq,
                                                               uint8_t* __restrict__ q) {
  __shared__ __align__(16) uint8_t smem[CQ_SMEM_BYTES];
  dev_color_quantize(bgr, W, H, thr_sq, q, blockIdx.x, blockIdx.y, smem);
}

void fl_launch_color_quantize(const uint8_t* bgr, int W, int H, float thr_sq, uint8_t* q, cudaStream_t s) {
  dim3 grid((W + CQ_TW - 1) / CQ_TW, (H + CQ_TH - 1) / CQ_TH);
  k_color_quantize<<<grid, CQ_THREADS, 0, s>>>(bgr, W, H, thr_sq, q);
}

#include "frontend_v2.cuh"

// ------------------------------------------------------------------------------------------------
// K2 cv::pyrDown of the BGR image: 5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

// one thread per output PIXEL (3 channels), CTA = 32 x 8 output pixels (the first version ran one thread per output byte
// and spent ~540 instructions on index arithmetic for each of them)
#define PD_TW 32
#define PD_TH 8
__device__ __forceinline__ int reflect101_near(int p, int n) {         // |overshoot| <= 2 and n >= 3: a single fold
  return p < 0 ? -p : (p >= n ? 2 * n - 2 - p : p);
}
// src_static: the source was complete before this launch (the caller's frame), so it may go through L1 (neighbouring threads
// share every line); a level produced by an earlier job of the same launch must be read from L2.
__device__ __forceinline__ void dev_pyrdown_bgr(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst, int bx, bool src_static) {
  const int dw = W / 2, dh = H / 2;
  const int gx = (dw + PD_TW - 1) / PD_TW;
  const int by = bx / gx; bx -= by * gx;
  const int x = bx * PD_TW + (threadIdx.x & 31), y = by * PD_TH + (threadIdx.x >> 5);
  if (x >= dw || y >= dh) return;
  int cx[5], ry[5];
  if (W >= 3 && H >= 3) {
#pragma unroll
    for (int t = 0; t < 5; ++t) { cx[t] = reflect101_near(2 * x + t - 2, W) * 3; ry[t] = reflect101_near(2 * y + t - 2, H); }
  } else {
#pragma unroll
    for (int t = 0; t < 5; ++t) { cx[t] = reflect101(2 * x + t - 2, W) * 3; ry[t] = reflect101(2 * y + t - 2, H); }
  }
  int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint8_t* row = src + (size_t)ry[j] * W * 3;
    const int kj = j == 0 || j == 4 ? 1 : (j == 2 ? 6 : 4);
    int r0 = 0, r1 = 0, r2 = 0;
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int kt = t == 0 || t == 4 ? 1 : (t == 2 ? 6 : 4);
      const uint8_t* p = row + cx[t];
      if (src_static) { r0 += kt * __ldg(p); r1 += kt * __ldg(p + 1); r2 += kt * __ldg(p + 2); }
      else { r0 += kt * __ldcg(p); r1 += kt * __ldcg(p + 1); r2 += kt * __ldcg(p + 2); }
    }
    s0 += kj * r0; s1 += kj * r1; s2 += kj * r2;
  }
  uint8_t* o = dst + ((size_t)y * dw + x) * 3;
  o[0] = (uint8_t)((s0 + 128) >> 8); o[1] = (uint8_t)((s1 + 128) >> 8); o[2] = (uint8_t)((s2 + 128) >> 8);
}

__global__ void __launch_bounds__(256) k_pyrdown_bgr(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst) {
  dev_pyrdown_bgr(src, W, H, dst, blockIdx.x, false);
}

static int pyrdown_ctas(int W, int H) { return ((W / 2 + PD_TW - 1) / PD_TW) * ((H / 2 + PD_TH - 1) / PD_TH); }
void fl_launch_pyrdown_bgr(const uint8_t* src, int W, int H, uint8_t* dst, cudaStream_t s) {
  if (pyrdown_ctas(W, H) > 0) k_pyrdown_bgr<<<pyrdown_ctas(W, H), 256, 0, s>>>(src, W, H, dst);
}

// ------------------------------------------------------------------------------------------------
// K3 depth normals: 8-tap bilateral least squares -> normal -> NORMAL_LUT, then 5x5 median, fused (halo 2 + 5).
// ------------------------------------------------------------------------------------------------
#define DQ_TW 32
#define DQ_TH 16
#define DQ_THREADS 256

#define DQ_SMEM_BYTES ((DQ_TH + 14) * (DQ_TW + 14) * 2 + 16 + (((DQ_TH + 4) * (DQ_TW + 4) + 15) & ~15) + DQ_TH * (DQ_TW + 4) * 8)

__device__ __forceinline__ void dev_depth_quantize(const uint16_t* __restrict__ depth, int W, int H, int dist_thr, int diff_thr,
                                                   uint8_t* __restrict__ q, int bx, int by, uint8_t* smem) {
  constexpr int RW = DQ_TW + 4, RH = DQ_TH + 4;      // raw label tile (median halo 2)
  constexpr int SW = RW + 10, SH = RH + 10;          // depth tile (tap radius 5)
  uint16_t(*s_d)[SW] = reinterpret_cast<uint16_t(*)[SW]>(smem);
  uint8_t(*s_r)[RW] = reinterpret_cast<uint8_t(*)[RW]>(smem + ((SH * SW * 2 + 15) & ~15));
  const int x0 = bx * DQ_TW, y0 = by * DQ_TH;
  const int tid = threadIdx.x;
  for (int i = tid; i < SH * SW; i += DQ_THREADS) {
    int r = i / SW, c = i - r * SW;
    int sy = y0 - 7 + r, sx = x0 - 7 + c;
    s_d[r][c] = (sx >= 0 && sx < W && sy >= 0 && sy < H) ? depth[(size_t)sy * W + sx] : (uint16_t)0;
  }
  __syncthreads();
  for (int i = tid; i < RH * RW; i += DQ_THREADS) {
    int ry = i / RW, rx = i - ry * RW;
    // medianBlur replicates the border of the label image; border labels are 0 (loop bounds :619, :624), so any
    // position outside [5, W-7] x [5, H-7] - inside or outside the image - contributes 0.
    int px = x0 - 2 + rx, py = y0 - 2 + ry;
    uint8_t v = 0;
    if (px >= 5 && px < W - 6 && py >= 5 && py < H - 6) {
      int cy = ry + 5, cx = rx + 5;
      int d = s_d[cy][cx];
      if (d < dist_thr) {
        int A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
#pragma unroll
        for (int j = -5; j <= 5; j += 5)
#pragma unroll
          for (int ii = -5; ii <= 5; ii += 5) {
            if (ii == 0 && j == 0) continue;
            int delta = (int)s_d[cy + j][cx + ii] - d;
            int f = abs(delta) < diff_thr ? 1 : 0;           // accumBilateral :567-579
            A0 += f * ii * ii; A1 += f * ii * j; A3 += f * j * j;
            b0 += f * ii * delta; b1 += f * j * delta;
          }
        int det = A0 * A3 - A1 * A1;                          // all fit int32
        int ddx = A3 * b0 - A1 * b1;
        int ddy = -A1 * b0 + A0 * b1;
        float nx = (float)(617 * ddx), ny = (float)(617 * ddy), nz = (float)(-det * d);
        float s = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
        if (s > 0.f) {
          float inv = __fdiv_rn(1.0f, s);
          nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv);
          int v1 = (int)__fadd_rn(__fmul_rn(nx, 10.f), 10.f);   // C truncation :665-667
          int v2 = (int)__fadd_rn(__fmul_rn(ny, 10.f), 10.f);
          v = c_normal_plane[clampi(v2, 0, 19) * 20 + clampi(v1, 0, 19)];   // table is v3-independent
        }
      }
    }
    s_r[ry][rx] = v;
  }
  __syncthreads();
  // 5x5 median (cv::medianBlur, :684).  The labels are one-hot bytes or 0 (NORMAL_LUT), i.e. 9 distinct values whose numeric
  // order is the order of idx = 0 (for 0) or bit position + 1, so the median is read off a 9-bin histogram: bins of 5 bits
  // packed in a u64, per-column histograms of 5 rows shared by the 5 pixels that overlap them.
  unsigned long long(*s_ch)[RW] = reinterpret_cast<unsigned long long(*)[RW]>(smem + ((SH * SW * 2 + 15) & ~15) + ((RH * RW + 15) & ~15));
  for (int i = tid; i < DQ_TH * RW; i += DQ_THREADS) {
    const int oy = i / RW, rx = i - oy * RW;
    unsigned long long hsum = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) hsum += 1ull << (5 * (32 - __clz((unsigned)s_r[oy + j][rx])));
    s_ch[oy][rx] = hsum;
  }
  __syncthreads();
  for (int i = tid; i < DQ_TH * DQ_TW; i += DQ_THREADS) {
    const int oy = i / DQ_TW, ox = i - oy * DQ_TW;
    const int x = x0 + ox, y = y0 + oy;
    if (x >= W || y >= H) continue;
    const unsigned long long hist = s_ch[oy][ox] + s_ch[oy][ox + 1] + s_ch[oy][ox + 2] + s_ch[oy][ox + 3] + s_ch[oy][ox + 4];
    int c = 0, med = 0;                                  // 13th smallest of 25
#pragma unroll
    for (int b = 0; b < 9; ++b) {
      c += (int)((hist >> (5 * b)) & 31);
      if (c < 13) med = b + 1;
    }
    q[(size_t)y * W + x] = med == 0 ? (uint8_t)0 : (uint8_t)(1u << (med - 1));
  }
}

__global__ void __launch_bounds__(DQ_THREADS) k_depth_quantize(const uint16_t* __restrict__ depth, int W, int H, int dist_thr,
                                                               int diff_thr, uint8_t* __restrict__ q) {
  __shared__ __align__(16) uint8_t smem[DQ_SMEM_BYTES];
  dev_depth_quantize(depth, W, H, dist_thr, diff_thr, q, blockIdx.x, blockIdx.y, smem);
}

void fl_launch_depth_quantize(const uint16_t* depth, int W, int H, int dist_thr, int diff_thr, uint8_t* q, cudaStream_t s) {
  dim3 grid((W + DQ_TW - 1) / DQ_TW, (H + DQ_TH - 1) / DQ_TH);
  k_depth_quantize<<<grid, DQ_THREADS, 0, s>>>(depth, W, H, dist_thr, diff_thr, q);
}

// ------------------------------------------------------------------------------------------------
// K4 nearest-neighbour x1/2 (cv::resize INTER_NEAREST to (W/2, H/2)) and mask application (copyTo(dst, mask))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dev_resize_nn_half(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst, int bx) {
  const int dw = W / 2, dh = H / 2;
  int i = bx * 256 + threadIdx.x;
  if (i >= dw * dh) return;
  int x = i % dw, y = i / dw;
  int sx = min((int)floor(x * ((double)W / dw)), W - 1);
  int sy = min((int)floor(y * ((double)H / dh)), H - 1);
  dst[i] = src[(size_t)sy * W + sx];
}
__global__ void __launch_bounds__(256) k_resize_nn_half(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst) {
  dev_resize_nn_half(src, W, H, dst, blockIdx.x);
}
void fl_launch_resize_nn_half(const uint8_t* src, int W, int H, uint8_t* dst, cudaStream_t s) {
  int n = (W / 2) * (H / 2);
  k_resize_nn_half<<<(n + 255) / 256, 256, 0, s>>>(src, W, H, dst);
}

__global__ void __launch_bounds__(256) k_apply_mask(const uint8_t* __restrict__ q, const uint8_t* __restrict__ mask, int n,
                                                    uint8_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = mask[i] ? q[i] : (uint8_t)0;
}
void fl_launch_apply_mask(const uint8_t* q, const uint8_t* mask, int n, uint8_t* out, cudaStream_t s) {
  k_apply_mask<<<(n + 255) / 256, 256, 0, s>>>(q, mask, n, out);
}

// ------------------------------------------------------------------------------------------------
// K5 spread + response maps + linearize, fused.  One CTA = one row of the decimated grid (T image rows) x SL_CW cells.
//   spread(y,x) = OR over the forward TxT window (clipped at the image end)                          :950-965
//   response_i  = table lookup of the spread byte (all 8 labels at once)                              :979-1048
//   LM_i[(y%T)*T + x%T][(y/T)*W' + x/T] = response_i(y,x)                                            :1060-1088
// HBM traffic = W*H read + 8*W*H written; each (label, grid-row) run of SL_CW output bytes is one full 32 B sector.
// ------------------------------------------------------------------------------------------------
#define SL_CW 32
#define SL_THREADS 256

// shared memory: response table (2 KB) + input tile + horizontal-OR tile (rows padded to whole 32-bit words + slack)
__host__ __device__ inline int spread_stride_in(int T) { return (SL_CW * T + T - 1 + 6) & ~3; }
__host__ __device__ inline int spread_stride_pw(int T) { return (SL_CW * T + 3) & ~3; }
__host__ __device__ inline size_t spread_smem_bytes(int T) {
  return 2048 + (size_t)(2 * T - 1) * spread_stride_in(T) + (size_t)(2 * T - 1) * spread_stride_pw(T);   // <= 35 KB at T = 16
}

// OR of the T bytes starting at every byte of a 32-bit word: words w[0..] hold the row from the word's own position on.
// TT > 0: T known at compile time (the loop unrolls to T funnel shifts); TT == 0: runtime T, words read on the fly.
template <int TT>
__device__ __forceinline__ uint32_t or_window(const uint32_t* __restrict__ w, int T) {
  uint32_t o = 0;
  if (TT > 0) {
    uint32_t v[(TT + 2) / 4 + 2];
#pragma unroll
    for (int i = 0; i < (TT + 2) / 4 + 2; ++i) v[i] = w[i];
#pragma unroll
    for (int k = 0; k < TT; ++k) o |= __funnelshift_r(v[k >> 2], v[(k >> 2) + 1], 8 * (k & 3));
  } else {
    for (int k = 0; k < T; ++k) o |= __funnelshift_r(w[k >> 2], w[(k >> 2) + 1], 8 * (k & 3));
  }
  return o;
}

template <int TT>
__device__ __forceinline__ void dev_spread_lm(const uint8_t* __restrict__ q, const fl_level_geom& g, uint8_t* __restrict__ lm,
                                              uint8_t* __restrict__ spread_out, int bx, int by, uint8_t* smem, int cw) {
  const int T = TT > 0 ? TT : g.T;
  const int W = g.W, H = g.H, Wd = g.Wd;
  const int gy = by, cx0 = bx * cw;                     // cw <= SL_CW cells per CTA (fewer on small levels: more CTAs in flight)
  const int ncell = min(cw, Wd - cx0);
  const int pw = ncell * T;                             // pixels owned by this CTA per row
  const int tw = pw + T - 1, th = 2 * T - 1;
  const int px0 = cx0 * T, py0 = gy * T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = SL_THREADS / 32;
  const int st_in = (tw + 6) & ~3;                      // row stride of the input tile: whole words, >= 3 bytes of zero slack
  const int nw = (pw + 3) >> 2, st_pw = nw << 2;        // words per row of the OR tiles
  uint2* s_resp = reinterpret_cast<uint2*>(smem);        // 256 x 8 B
  uint8_t* s_in = smem + 2048;                          // [th][st_in]   input labels (later reused for the final spread)
  uint8_t* s_hor = s_in + (size_t)(2 * T - 1) * spread_stride_in(T);   // [th][st_pw] horizontal OR
  for (int i = tid; i < 256; i += SL_THREADS) s_resp[i] = g_resp8[i];
  if (((reinterpret_cast<uintptr_t>(q) | (uintptr_t)W | (uintptr_t)px0) & 3) == 0) {
    // word-aligned tile: whole 32-bit words, every load of the tile in flight at once (fixed trip count, unrolled).  A word is
    // either entirely inside the image or entirely outside (W % 4 == 0); bytes past the CTA's own window are real pixels of
    // the neighbour and are never read by the OR passes below.
    const int nwi = st_in >> 2;
    const int n_items = th * nwi;
    uint32_t* s_in32 = reinterpret_cast<uint32_t*>(s_in);
    for (int i0 = 0; i0 < n_items; i0 += 4 * SL_THREADS) {
      uint32_t v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                     // 4 loads in flight per thread (the whole tile for T <= 8)
        const int i = i0 + tid + u * SL_THREADS;
        const int r = i / nwi, wi = i - r * nwi;
        const int y = py0 + r, x = px0 + 4 * wi;
        v[u] = (i < n_items && y < H && x < W) ? __ldcg(reinterpret_cast<const uint32_t*>(q + (size_t)y * W + x)) : 0u;   // clipped window == OR with zeros
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + tid + u * SL_THREADS; if (i < n_items) s_in32[i] = v[u]; }
    }
  } else {
    for (int r = warp; r < th; r += NWARP) {            // rows by warp, bytes by lane: coalesced, no index arithmetic
      const int y = py0 + r;
      const uint8_t* src = q + (size_t)y * W + px0;
      for (int c = lane; c < st_in; c += 32) s_in[r * st_in + c] = (c < tw && px0 + c < W && y < H) ? __ldcg(src + c) : (uint8_t)0;   // clipped window == OR with zeros
    }
  }
  __syncthreads();
  for (int r = warp; r < th; r += NWARP) {              // horizontal OR, four pixels per thread
    const uint32_t* row = reinterpret_cast<const uint32_t*>(s_in + r * st_in);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_hor + r * st_pw);
    for (int wi = lane; wi < nw; wi += 32) dst[wi] = or_window<TT>(row + wi, T);
  }
  __syncthreads();
  uint8_t* s_sp = s_in;                                 // [T][st_pw] final spread values (s_in is dead)
  for (int r = warp; r < T; r += NWARP) {               // vertical OR
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_sp + r * st_pw);
    for (int wi = lane; wi < nw; wi += 32) {
      uint32_t v = 0;
      if (TT > 0) {
#pragma unroll
        for (int k = 0; k < TT; ++k) v |= reinterpret_cast<const uint32_t*>(s_hor + (r + k) * st_pw)[wi];
      } else {
        for (int k = 0; k < T; ++k) v |= reinterpret_cast<const uint32_t*>(s_hor + (r + k) * st_pw)[wi];
      }
      dst[wi] = v;
      if (spread_out)
        for (int k = 0; k < 4; ++k) if (4 * wi + k < pw) spread_out[(size_t)(py0 + r) * W + px0 + 4 * wi + k] = (uint8_t)(v >> (8 * k));
    }
  }
  __syncthreads();
  // scatter: item = (residue row ry, residue column rx, group of 4 cells); 8 labels -> 8 word stores per item
  const int ngrp = (ncell + 3) >> 2;
  const int items = T * T * ngrp;
  const size_t cells = (size_t)g.cells;
  const uint32_t inv_ngrp = (65536u + ngrp - 1) / ngrp, inv_T = (65536u + T - 1) / T;   // exact for the small ranges used here
  for (int i = tid; i < items; i += SL_THREADS) {
    const int rr = (int)(((uint32_t)i * inv_ngrp) >> 16), c4 = i - rr * ngrp;
    const int ry = (int)(((uint32_t)rr * inv_T) >> 16), rx = rr - ry * T;
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c4 * 4 + k;
      const uint2 r8 = c < ncell ? s_resp[s_sp[ry * st_pw + c * T + rx]] : make_uint2(0, 0);
      lo[k] = r8.x; hi[k] = r8.y;
    }
    const size_t off = (size_t)rr * cells + (size_t)gy * Wd + cx0 + c4 * 4;
    const int nvalid = min(4, ncell - c4 * 4);
#pragma unroll
    for (int lab = 0; lab < 8; ++lab) {
      const uint32_t* src = lab < 4 ? lo : hi;
      // byte (lab & 3) of each of the four words -> one word
      const uint32_t sel = 0x0040u + (lab & 3) * 0x0011u;          // PRMT selector: {byte b of x, byte b of y}
      const uint32_t w01 = __byte_perm(src[0], src[1], sel), w23 = __byte_perm(src[2], src[3], sel);
      const uint32_t w = __byte_perm(w01, w23, 0x5410);
      uint8_t* dst = lm + (size_t)lab * g.label_stride + off;
      if (nvalid == 4 && (((uintptr_t)dst) & 3) == 0) *reinterpret_cast<uint32_t*>(dst) = w;
      else for (int k = 0; k < nvalid; ++k) dst[k] = (uint8_t)(w >> (8 * k));
    }
  }
}

__device__ __forceinline__ void dev_spread_lm_any(const uint8_t* __restrict__ q, const fl_level_geom& g, uint8_t* __restrict__ lm,
                                                  uint8_t* __restrict__ spread_out, int bx, int by, uint8_t* smem, int cw) {
  switch (g.T) {
    case 5: dev_spread_lm<5>(q, g, lm, spread_out, bx, by, smem, cw); break;
    case 8: dev_spread_lm<8>(q, g, lm, spread_out, bx, by, smem, cw); break;
    case 4: dev_spread_lm<4>(q, g, lm, spread_out, bx, by, smem, cw); break;
    default: dev_spread_lm<0>(q, g, lm, spread_out, bx, by, smem, cw); break;
  }
}

__global__ void __launch_bounds__(SL_THREADS) k_spread_lm(const uint8_t* __restrict__ q, fl_level_geom g, uint8_t* __restrict__ lm,
                                                          uint8_t* __restrict__ spread_out) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  dev_spread_lm_any(q, g, lm, spread_out, blockIdx.x, blockIdx.y, smem_dyn, SL_CW);
}

void fl_launch_spread_lm(const uint8_t* q, fl_level_geom g, uint8_t* lm_mod, uint8_t* spread_or_null, cudaStream_t s) {
  dim3 grid((g.Wd + SL_CW - 1) / SL_CW, g.Hd);
  k_spread_lm<<<grid, SL_THREADS, spread_smem_bytes(g.T), s>>>(q, g, lm_mod, spread_or_null);
}

// ------------------------------------------------------------------------------------------------
// Wave kernel: at VGA every stage above moves a few hundred KB, i.e. ~1 us of HBM time against ~5-8 us of launch +
// dependency latency, so the front end is launch bound.  All stages that are independent of each other at one point of
// the pyramid recursion are therefore issued as ONE launch: the grid is the concatenation of the jobs' grids.
//   wave 0      : colour quantise L0 | depth quantise L0 | pyrDown L0->L1
//   wave k >= 1 : colour quantise Lk | NN-downsample depth labels Lk | pyrDown Lk->Lk+1 | spread+LM of level k-1 (all modalities)
//   last wave   : spread+LM of the coarsest level
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 8) k_front_end_wave(fl_fe_wave w) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const int b = blockIdx.x;
  fl_grid_dep_wait();                                   // the previous kernel of the stream wrote this wave's inputs
  fl_grid_dep_launch();
  if (w.zero_me && b == 0 && threadIdx.x == 0) *w.zero_me = 0;
  int j = 0;
#pragma unroll 1
  while (j + 1 < w.n_jobs && b >= w.job[j + 1].cta_begin) ++j;
  const fl_fe_job& jb = w.job[j];
  const int local = b - jb.cta_begin;
  if (w.trace && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMin(w.trace + 2 * j, t); }
  if (jb.wait_row_base >= 0) {
    // row-grained wait of a spread CTA (see fl_fe_job): the producer's tile rows that hold this grid row's label rows
    if (threadIdx.x == 0) {
      const int gy = local / jb.gx, T = jb.g.T;
      const int y0 = gy * T, y1 = min(gy * T + 2 * T - 2, jb.g.H - 1);
      const int by0 = (y0 << jb.wait_row_shift) >> 4, by1 = min((y1 << jb.wait_row_shift) >> 4, jb.wait_row_count - 1);
      const long long t0 = clock64();
      for (int by = by0; by <= by1; ++by) {
        const unsigned* c = w.counters + jb.wait_row_base + by;
        unsigned v;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
          if ((int)(v - jb.wait_row_target) >= 0) break;
          if (clock64() - t0 > (1ll << 31)) { if (w.dep_error) *w.dep_error = 1 + j; __threadfence_system(); __trap(); }
          __nanosleep(64);
        }
      }
    }
    __syncthreads();
  } else if (jb.wait_slot >= 0) {
    // In-grid dependency: CTAs are dispatched in blockIdx order and a producing job always precedes its consumers in the grid,
    // so every producer CTA is resident (or done) before a consumer starts to wait here.  The wait is bounded: after ~1 s it gives
    // up, records the job in dep_error and traps (the host's next CUDA call fails with FL_ERR_CUDA) instead of hanging the device.
    if (threadIdx.x == 0) {
      const unsigned* c = w.counters + jb.wait_slot;
      const long long t0 = clock64();
      unsigned v;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
        if ((int)(v - jb.wait_target) >= 0) break;
        if (clock64() - t0 > (1ll << 31)) {           // never observed; fail loudly (the next CUDA call of the host reports it) instead of computing on stale data
          if (w.dep_error) *w.dep_error = 1 + j;
          __threadfence_system();
          __trap();
        }
        __nanosleep(64);
      }
    }
    __syncthreads();
  }
  switch (jb.kind) {
    case FL_JOB_COLOR: { const int tile = local + jb.p0; dev_color_quantize(jb.src, jb.W, jb.H, jb.thr_sq, jb.dst, tile % jb.gx, tile / jb.gx, smem_dyn); break; }
    case FL_JOB_DEPTH: dev_depth_quantize(reinterpret_cast<const uint16_t*>(jb.src), jb.W, jb.H, jb.p0, jb.p1, jb.dst, local % jb.gx, local / jb.gx, smem_dyn); break;
    case FL_JOB_PYRDOWN: dev_pyrdown_bgr(jb.src, jb.W, jb.H, jb.dst, local, jb.p0 != 0); break;
    case FL_JOB_RESIZE: dev_resize_nn_half(jb.src, jb.W, jb.H, jb.dst, local); break;
    case FL_JOB_SPREAD: dev_spread_lm_any(jb.src, jb.g, jb.dst, jb.dst2, local % jb.gx, local / jb.gx, smem_dyn, jb.p0); break;
    case FL_JOB_PREFETCH: {
      const size_t line = (size_t)local * 256 + threadIdx.x;
      if (line * 128 < (size_t)jb.W * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(jb.src + line * 128));
      break;
    }
    case FL_JOB_COLOR2: dev_color_quantize_v2(jb.src, jb.W, jb.H, jb.thr_sq, jb.dst, local % jb.gx, local / jb.gx, smem_dyn); break;
    case FL_JOB_DEPTH2:
      dev_depth_quantize_v2(reinterpret_cast<const uint16_t*>(jb.src), jb.W, jb.H, jb.p0, jb.p1, jb.dst, w, jb.gx < 0 ? -1 : (jb.thr_sq > 0.5f ? 1 : 0),
                            local % abs(jb.gx), local / abs(jb.gx), smem_dyn);
      break;
  }
  if (w.trace) { __syncthreads(); if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMax(w.trace + 2 * j + 1, t); } }
  if (jb.signal_slot >= 0) {
    __syncthreads();                                    // every thread's stores of this CTA ...
    if (threadIdx.x == 0) {                             // ... are visible before the counts
      __threadfence();
      if (jb.row_base >= 0) atomicAdd(w.counters + jb.row_base + local / abs(jb.gx), 1u);
      atomicAdd(w.counters + jb.signal_slot, 1u);
    }
  }
}

// part / n_parts: the tiles of one image may be spread over several waves (the job then covers tiles [first, first + count))
void fl_fe_add_color(fl_fe_wave* w, const uint8_t* bgr, int W, int H, float thr_sq, uint8_t* q, int part, int n_parts) {
  const int gx = (W + CQ_TW - 1) / CQ_TW, tiles = gx * ((H + CQ_TH - 1) / CQ_TH);
  const int first = (int)((long long)tiles * part / n_parts), last = (int)((long long)tiles * (part + 1) / n_parts);
  if (last <= first) return;
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_COLOR; j.src = bgr; j.dst = q; j.dst2 = nullptr; j.W = W; j.H = H; j.thr_sq = thr_sq;
  j.gx = gx; j.p0 = first; j.p1 = last - first; j.cta_begin = w->n_ctas; w->n_ctas += last - first;
  w->smem = w->smem > (size_t)CQ_SMEM_BYTES ? w->smem : (size_t)CQ_SMEM_BYTES;
}
void fl_fe_add_depth(fl_fe_wave* w, const uint16_t* depth, int W, int H, int dist_thr, int diff_thr, uint8_t* q) {
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_DEPTH; j.src = reinterpret_cast<const uint8_t*>(depth); j.dst = q; j.dst2 = nullptr; j.W = W; j.H = H; j.p0 = dist_thr; j.p1 = diff_thr;
  j.gx = (W + DQ_TW - 1) / DQ_TW; j.cta_begin = w->n_ctas; w->n_ctas += j.gx * ((H + DQ_TH - 1) / DQ_TH);
  w->smem = w->smem > (size_t)DQ_SMEM_BYTES ? w->smem : (size_t)DQ_SMEM_BYTES;
}
void fl_fe_add_color_v2(fl_fe_wave* w, const uint8_t* bgr, int W, int H, float thr_sq, uint8_t* q) {
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_COLOR2; j.src = bgr; j.dst = q; j.dst2 = nullptr; j.W = W; j.H = H; j.thr_sq = thr_sq;
  j.gx = (W + C2_TW - 1) / C2_TW; j.p0 = j.p1 = 0; j.cta_begin = w->n_ctas; w->n_ctas += j.gx * ((H + C2_TH - 1) / C2_TH);
  w->smem = w->smem > (size_t)C2_SMEM_BYTES ? w->smem : (size_t)C2_SMEM_BYTES;
}
// returns false (nothing added) when a pyramid is requested but the wave has no free pyramid slot
bool fl_fe_add_depth_v2(fl_fe_wave* w, const uint16_t* depth, int W, int H, int dist_thr, int diff_thr, uint8_t* q, const fl_depth_pyr* pyr) {
  int slot = -1;
  if (pyr && pyr->n > 0) {
    if (w->n_pyr >= FL_FE_MAX_DPYR) return false;
    slot = w->n_pyr++;
    w->pyr[slot] = *pyr;
  }
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_DEPTH2; j.src = reinterpret_cast<const uint8_t*>(depth); j.dst = q; j.dst2 = nullptr; j.W = W; j.H = H; j.p0 = dist_thr; j.p1 = diff_thr;
  const int gx = (W + D2_TW - 1) / D2_TW;
  j.gx = slot < 0 ? -gx : gx; j.thr_sq = (float)(slot < 0 ? 0 : slot);            // gx < 0: no pyramid; thr_sq carries the pyramid slot
  j.cta_begin = w->n_ctas; w->n_ctas += gx * ((H + D2_TH - 1) / D2_TH);
  w->smem = w->smem > (size_t)D2_SMEM_BYTES ? w->smem : (size_t)D2_SMEM_BYTES;
  return true;
}
void fl_fe_add_pyrdown(fl_fe_wave* w, const uint8_t* src, int W, int H, uint8_t* dst, bool src_static) {
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_PYRDOWN; j.src = src; j.dst = dst; j.dst2 = nullptr; j.W = W; j.H = H; j.gx = 1; j.p0 = src_static ? 1 : 0;
  j.cta_begin = w->n_ctas; w->n_ctas += pyrdown_ctas(W, H);
}
void fl_fe_add_prefetch(fl_fe_wave* w, const void* src, size_t bytes) {
  const int lines = (int)((bytes + 127) / 128);
  if (lines <= 0 || w->n_jobs >= FL_FE_MAX_JOBS) return;
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_PREFETCH; j.src = static_cast<const uint8_t*>(src); j.dst = nullptr; j.dst2 = nullptr; j.W = lines; j.H = 0; j.gx = 1;
  j.cta_begin = w->n_ctas; w->n_ctas += (lines + 255) / 256;
}
void fl_fe_add_resize(fl_fe_wave* w, const uint8_t* src, int W, int H, uint8_t* dst) {
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_RESIZE; j.src = src; j.dst = dst; j.dst2 = nullptr; j.W = W; j.H = H; j.gx = 1;
  j.cta_begin = w->n_ctas; w->n_ctas += ((W / 2) * (H / 2) + 255) / 256;
}
void fl_fe_add_spread(fl_fe_wave* w, const uint8_t* q, fl_level_geom g, uint8_t* lm_mod, uint8_t* spread_or_null) {
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_SPREAD; j.src = q; j.dst = lm_mod; j.dst2 = spread_or_null; j.g = g; j.W = g.W; j.H = g.H;
  int cw = SL_CW;                                       // small levels: narrower CTAs so that the job still fills the SMs
  while (cw > 8 && ((g.Wd + cw - 1) / cw) * g.Hd < 148) cw >>= 1;
  j.p0 = cw;
  j.gx = (g.Wd + cw - 1) / cw; j.cta_begin = w->n_ctas; w->n_ctas += j.gx * g.Hd;
  size_t sm = spread_smem_bytes(g.T);
  w->smem = w->smem > sm ? w->smem : sm;
}
void fl_launch_fe_wave(const fl_fe_wave& w, cudaStream_t s) {
  if (w.n_ctas > 0) fl_launch_pdl(k_front_end_wave, dim3(w.n_ctas), dim3(256), w.smem, s, w);
}

// every hot kernel asks for the maximum shared-memory carveout so that the SMs never re-partition L1/shared memory between
// the launches of one frame (the staged similarity kernel needs > 200 KB)
void fl_prefer_smem_carveout_frontend() {
  cudaFuncSetAttribute(k_front_end_wave, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(k_spread_lm, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(k_color_quantize, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(k_depth_quantize, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}
