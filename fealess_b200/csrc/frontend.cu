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
                                              uint8_t* __restrict__ spread_out, uint8_t* __restrict__ lm4, int bx, int by, uint8_t* smem, int cw) {
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
      if (lm4) {
        // the same four responses (each 0, 1, 2 or 4) as nibbles, cell j of the label in nibble j: what the staged similarity
        // kernel reads (half the bytes).  Only passed for grids of even width, so `off` is even and nvalid is 2 or 4.
        const uint32_t p2 = (w | (w >> 4)) & 0x00FF00FFu;
        const uint32_t p4 = (p2 | (p2 >> 8)) & 0xFFFFu;
        uint8_t* d4 = lm4 + (((size_t)lab * g.label_stride + off) >> 1);
        if (nvalid == 4) *reinterpret_cast<uint16_t*>(d4) = (uint16_t)p4;
        else *d4 = (uint8_t)p4;
      }
    }
  }
}

__device__ __forceinline__ void dev_spread_lm_any(const uint8_t* __restrict__ q, const fl_level_geom& g, uint8_t* __restrict__ lm,
                                                  uint8_t* __restrict__ spread_out, uint8_t* __restrict__ lm4, int bx, int by, uint8_t* smem, int cw) {
  switch (g.T) {
    case 5: dev_spread_lm<5>(q, g, lm, spread_out, lm4, bx, by, smem, cw); break;
    case 8: dev_spread_lm<8>(q, g, lm, spread_out, lm4, bx, by, smem, cw); break;
    case 4: dev_spread_lm<4>(q, g, lm, spread_out, lm4, bx, by, smem, cw); break;
    default: dev_spread_lm<0>(q, g, lm, spread_out, lm4, bx, by, smem, cw); break;
  }
}

__global__ void __launch_bounds__(SL_THREADS) k_spread_lm(const uint8_t* __restrict__ q, fl_level_geom g, uint8_t* __restrict__ lm,
                                                          uint8_t* __restrict__ spread_out, uint8_t* __restrict__ lm4) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  dev_spread_lm_any(q, g, lm, spread_out, lm4, blockIdx.x, blockIdx.y, smem_dyn, SL_CW);
}

void fl_launch_spread_lm(const uint8_t* q, fl_level_geom g, uint8_t* lm_mod, uint8_t* spread_or_null, uint8_t* lm4_mod_or_null, cudaStream_t s) {
  dim3 grid((g.Wd + SL_CW - 1) / SL_CW, g.Hd);
  k_spread_lm<<<grid, SL_THREADS, spread_smem_bytes(g.T), s>>>(q, g, lm_mod, spread_or_null, lm4_mod_or_null);
}

// ------------------------------------------------------------------------------------------------
// Wave kernel: at VGA every stage above moves a few hundred KB, i.e. ~1 us of HBM time against ~5-8 us of launch +
// dependency latency, so the front end is launch bound.  All stages that are independent of each other at one point of
// the pyramid recursion are therefore issued as ONE launch: the grid is the concatenation of the jobs' grids.
//   wave 0      : colour quantise L0 | depth quantise L0 | pyrDown L0->L1
//   wave k >= 1 : colour quantise Lk | NN-downsample depth labels Lk | pyrDown Lk->Lk+1 | spread+LM of level k-1 (all modalities)
//   last wave   : spread+LM of the coarsest level
// ------------------------------------------------------------------------------------------------
// bounded in-grid wait: true = stop waiting (this CTA timed out, or another one already did); see k_front_end_wave
__device__ __forceinline__ bool fe_dep_give_up(const fl_fe_wave& w, int job, long long t0) {
  if (w.dep_error && *reinterpret_cast<volatile int*>(w.dep_error) != 0) return true;
  if (clock64() - t0 <= (1ll << 31)) return false;
  if (w.dep_error) atomicCAS(w.dep_error, 0, 1 + job);
  if (w.dep_error_host) *reinterpret_cast<volatile int*>(w.dep_error_host) = 1 + job;
  __threadfence_system();
  return true;
}

__global__ void __launch_bounds__(256, 8) k_front_end_wave(fl_fe_wave w) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const int b = blockIdx.x;
  fl_grid_dep_wait();                                   // the previous kernel of the stream wrote this wave's inputs
  fl_grid_dep_launch();
  if (w.zero_me && b == 0 && threadIdx.x == 0) *w.zero_me = 0;
  // the job this CTA belongs to: last job whose first CTA is <= b.  Binary search over the (<= 16) jobs: every warp of the grid runs
  // this, and the linear walk it replaces was 10 % of the launch's instructions (ncu, round 2)
  int j = 0;
  {
    int hi = w.n_jobs - 1;
#pragma unroll 1
    while (j < hi) {
      const int mid = (j + hi + 1) >> 1;
      if (b >= w.job[mid].cta_begin) j = mid; else hi = mid - 1;
    }
  }
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
          if (fe_dep_give_up(w, j, t0)) break;
          __nanosleep(64);
        }
      }
    }
    __syncthreads();
  } else if (jb.wait_slot >= 0) {
    // In-grid dependency: a producing job always precedes its consumers in the grid and CTAs are dispatched in blockIdx order on
    // every configuration this was run on, so every producer CTA is resident (or done) before a consumer starts to wait here.
    // CUDA does not GUARANTEE that order (MPS time slicing, a debugger, compute-sanitizer, a future scheduler), so the wait is
    // bounded: after ~1 s the CTA records the job in dep_error (device word for the other waiters, mapped host word for the
    // library), every waiter of the grid gives up at once, the frame's result is discarded by the host, which re-runs the frame
    // as one launch per wave (no in-grid waits) and keeps this handle on that path (fe_dep_give_up, fl_match_wait).
    if (threadIdx.x == 0) {
      const unsigned* c = w.counters + jb.wait_slot;
      const long long t0 = clock64();
      unsigned v;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
        if ((int)(v - jb.wait_target) >= 0) break;
        if (fe_dep_give_up(w, j, t0)) break;
        __nanosleep(64);
      }
    }
    __syncthreads();
  }
  switch (jb.kind) {
    case FL_JOB_PYRDOWN: dev_pyrdown_bgr(jb.src, jb.W, jb.H, jb.dst, local, jb.p0 != 0); break;
    case FL_JOB_RESIZE: dev_resize_nn_half(jb.src, jb.W, jb.H, jb.dst, local); break;
    case FL_JOB_SPREAD: dev_spread_lm_any(jb.src, jb.g, jb.dst, jb.dst2, jb.dst3, local % jb.gx, local / jb.gx, smem_dyn, jb.p0); break;
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
void fl_fe_add_spread(fl_fe_wave* w, const uint8_t* q, fl_level_geom g, uint8_t* lm_mod, uint8_t* spread_or_null, uint8_t* lm4_mod_or_null) {
  fl_fe_job& j = w->job[w->n_jobs++];
  j.wait_slot = -1; j.signal_slot = -1; j.wait_target = 0; j.row_base = -1; j.wait_row_base = -1; j.wait_row_shift = 0; j.wait_row_count = 0; j.wait_row_target = 0;
  j.kind = FL_JOB_SPREAD; j.src = q; j.dst = lm_mod; j.dst2 = spread_or_null; j.dst3 = lm4_mod_or_null; j.g = g; j.W = g.W; j.H = g.H;
  int cw = SL_CW;                                       // small levels: narrower CTAs so that the job still fills the SMs
  while (cw > 8 && ((g.Wd + cw - 1) / cw) * g.Hd < 148) cw >>= 1;
  j.p0 = cw;
  j.gx = (g.Wd + cw - 1) / cw; j.cta_begin = w->n_ctas; w->n_ctas += j.gx * g.Hd;
  size_t sm = spread_smem_bytes(g.T);
  w->smem = w->smem > sm ? w->smem : sm;
}
cudaError_t fl_launch_fe_wave(const fl_fe_wave& w, cudaStream_t s) {
  return w.n_ctas > 0 ? fl_launch(k_front_end_wave, dim3(w.n_ctas), dim3(256), w.smem, s, w) : cudaSuccess;
}

