/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.  See fl_oracle.h for scope, parity status and the rules on
 * who may call this.  Plain C11, single-threaded unless n_threads > 1 is requested (OpenMP, templates only).
 * Build: gcc -O3 -msse4.2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: the fp32 index truncations of linemod.cpp:665-667 must not be fused.
 */
#include "fl_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LM_PAD 4096 /* zero bytes after each label's linear memory; defines the reference's over-read (see flat addressing note) */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int reflect101(int p, int n) { /* BORDER_REFLECT_101 */
  if (n == 1) return 0;
  while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
  return p;
}

/* ------------------------------------------------------------------------------------------ */
/* embedded tables                                                                            */
/* ------------------------------------------------------------------------------------------ */
void flo_normal_lut(uint8_t out[8000]) {
  /* NORMAL_LUT[v3][v2][v1] (normal_lut.i:4) does not depend on v3; entry = one-hot 45-degree sector of
   * atan2(v2-10, v1-10) + 22.5 deg.  Reproduces all 8000 reference bytes (sha256 pinned in tests). */
  uint8_t plane[400];
  for (int v2 = 0; v2 < 20; ++v2)
    for (int v1 = 0; v1 < 20; ++v1) {
      double a = atan2((double)(v2 - 10), (double)(v1 - 10)) * (180.0 / 3.14159265358979323846) + 22.5;
      a = fmod(a + 360.0, 360.0);
      plane[v2 * 20 + v1] = (uint8_t)(1u << (((int)floor(a / 45.0)) & 7));
    }
  for (int v3 = 0; v3 < 20; ++v3) memcpy(out + v3 * 400, plane, 400);
}

void flo_similarity_lut(uint8_t out[256]) {
  /* SIMILARITY_LUT (linemod.cpp:970): [32*i + 16*half + nibble] = max over set bits j of g(circular |i-j|),
   * g = 4,2,1,0,0. */
  static const uint8_t g[5] = {4, 2, 1, 0, 0};
  for (int i = 0; i < 8; ++i)
    for (int half = 0; half < 2; ++half)
      for (int nib = 0; nib < 16; ++nib) {
        int best = 0;
        for (int b = 0; b < 4; ++b)
          if (nib >> b & 1) {
            int j = b + 4 * half, d = abs(i - j);
            if (8 - d < d) d = 8 - d;
            if (g[d] > best) best = g[d];
          }
        out[32 * i + 16 * half + nib] = (uint8_t)best;
      }
}

static uint8_t g_normal_plane[400];
static uint8_t g_sim_lut[256];
static int g_tables_ready = 0;
static void ensure_tables(void) {
  if (g_tables_ready) return;
  uint8_t full[8000];
  flo_normal_lut(full);
  memcpy(g_normal_plane, full, 400);
  flo_similarity_lut(g_sim_lut);
  g_tables_ready = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* colour gradient modality                                                                   */
/* ------------------------------------------------------------------------------------------ */
void flo_gaussian7_bgr(const uint8_t* src, int W, int H, uint8_t* dst) {
  /* cv::GaussianBlur(Size(7,7), sigma 0) on 8U: OpenCV's fixed 7-tap table {8,28,56,72,56,28,8}/256, both passes in
   * exact integers, one rounding (sum + 2^15) >> 16, BORDER_REPLICATE (linemod.cpp:247). */
  static const int k[7] = {8, 28, 56, 72, 56, 28, 8};
  int* tmp = (int*)malloc(sizeof(int) * (size_t)W * H * 3);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < 3; ++c) {
        int s = 0;
        for (int i = 0; i < 7; ++i) s += k[i] * src[((size_t)y * W + clampi(x + i - 3, 0, W - 1)) * 3 + c];
        tmp[((size_t)y * W + x) * 3 + c] = s;
      }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < 3; ++c) {
        int s = 0;
        for (int i = 0; i < 7; ++i) s += k[i] * tmp[((size_t)clampi(y + i - 3, 0, H - 1) * W + x) * 3 + c];
        dst[((size_t)y * W + x) * 3 + c] = (uint8_t)((s + 32768) >> 16);
      }
  free(tmp);
}

void flo_sobel3_bgr(const uint8_t* src, int W, int H, int16_t* dx, int16_t* dy) {
  /* cv::Sobel(CV_16S, ksize 3, BORDER_REPLICATE) per channel (linemod.cpp:248-249) */
  for (int y = 0; y < H; ++y) {
    int ym = clampi(y - 1, 0, H - 1), yp = clampi(y + 1, 0, H - 1);
    for (int x = 0; x < W; ++x) {
      int xm = clampi(x - 1, 0, W - 1), xp = clampi(x + 1, 0, W - 1);
      for (int c = 0; c < 3; ++c) {
#define P(yy, xx) ((int)src[((size_t)(yy) * W + (xx)) * 3 + c])
        int gx = (P(ym, xp) + 2 * P(y, xp) + P(yp, xp)) - (P(ym, xm) + 2 * P(y, xm) + P(yp, xm));
        int gy = (P(yp, xm) + 2 * P(yp, x) + P(yp, xp)) - (P(ym, xm) + 2 * P(ym, x) + P(ym, xp));
#undef P
        dx[((size_t)y * W + x) * 3 + c] = (int16_t)gx;
        dy[((size_t)y * W + x) * 3 + c] = (int16_t)gy;
      }
    }
  }
}

/* cv::phase(.., angleInDegrees=true): OpenCV's fastAtan2 (core/src/mathfuncs_core), a degree-7 odd polynomial on
 * min/max ratio with octant fix-ups, evaluated in fp32.  OpenCV is an un-vendored dependency of the reference
 * (CMakeLists.txt:13-16, version unpinned); constants below are OpenCV's published ones, and the resulting BINS are
 * checked against cv2.phase over the whole reachable Sobel domain in tests/test_oracle_cv2.py. */
static inline float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

static inline uint8_t q16_of(float dx, float dy) {
  /* angle.convertTo(CV_8U, 16.0/360.0) (linemod.cpp:314): saturate_cast<uchar>(cvRound(a * (float)alpha)) */
  float v = fast_atan2_deg(dy, dx) * (float)(16.0 / 360.0);
  long r = lrintf(v); /* round-half-even under the default rounding mode == cvRound */
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

void flo_phase_q16(const float* dx, const float* dy, int n, uint8_t* q) {
  for (int i = 0; i < n; ++i) q[i] = q16_of(dx[i], dy[i]);
}

void flo_color_quantize(const uint8_t* bgr, int W, int H, float weak_thr, uint8_t* q, float* mag_out) {
  /* quantizedOrientations + hysteresisGradient, linemod.cpp:230-385 */
  size_t n = (size_t)W * H;
  uint8_t* sm = (uint8_t*)malloc(n * 3);
  int16_t* dx3 = (int16_t*)malloc(n * 3 * sizeof(int16_t));
  int16_t* dy3 = (int16_t*)malloc(n * 3 * sizeof(int16_t));
  float* mag = (float*)malloc(n * sizeof(float));
  uint8_t* qu = (uint8_t*)malloc(n);
  flo_gaussian7_bgr(bgr, W, H, sm);
  flo_sobel3_bgr(sm, W, H, dx3, dy3);
  for (size_t i = 0; i < n; ++i) {
    const int16_t* px = dx3 + i * 3; const int16_t* py = dy3 + i * 3;
    int m0 = px[0] * px[0] + py[0] * py[0];
    int m1 = px[1] * px[1] + py[1] * py[1];
    int m2 = px[2] * px[2] + py[2] * py[2];
    int c;                                                         /* :275-292 */
    if (m0 >= m1 && m0 >= m2) c = 0; else if (m1 >= m0 && m1 >= m2) c = 1; else c = 2;
    mag[i] = (float)(c == 0 ? m0 : (c == 1 ? m1 : m2));
    qu[i] = q16_of((float)px[c], (float)py[c]);                    /* :303, :314 */
  }
  /* :318-335 zero the 1-px frame, mask the interior to 3 bits */
  for (int x = 0; x < W; ++x) { qu[x] = 0; qu[(size_t)(H - 1) * W + x] = 0; }
  for (int y = 0; y < H; ++y) { qu[(size_t)y * W] = 0; qu[(size_t)y * W + W - 1] = 0; }
  for (int y = 1; y < H - 1; ++y) for (int x = 1; x < W - 1; ++x) qu[(size_t)y * W + x] &= 7;
  memset(q, 0, n);
  const float thr = weak_thr * weak_thr;                           /* :304 */
  for (int y = 1; y < H - 1; ++y)
    for (int x = 1; x < W - 1; ++x) {
      if (!(mag[(size_t)y * W + x] > thr)) continue;               /* :346 */
      int hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = -1; j <= 1; ++j) for (int i = -1; i <= 1; ++i) hist[qu[(size_t)(y + j) * W + x + i]]++;
      int best = 0, idx = -1;
      for (int b = 0; b < 8; ++b) if (best < hist[b]) { idx = b; best = hist[b]; }   /* :369-376 first max wins */
      if (best >= 5) q[(size_t)y * W + x] = (uint8_t)(1u << idx);  /* :380 */
    }
  if (mag_out) memcpy(mag_out, mag, n * sizeof(float));
  free(sm); free(dx3); free(dy3); free(mag); free(qu);
}

void flo_pyrdown_bgr(const uint8_t* src, int W, int H, uint8_t* dst) {
  /* cv::pyrDown to (W/2, H/2): 5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8 (linemod.cpp:441-444) */
  static const int k[5] = {1, 4, 6, 4, 1};
  int dw = W / 2, dh = H / 2;
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x)
      for (int c = 0; c < 3; ++c) {
        int s = 0;
        for (int j = 0; j < 5; ++j) {
          int sy = reflect101(2 * y + j - 2, H);
          int rs = 0;
          for (int i = 0; i < 5; ++i) rs += k[i] * src[((size_t)sy * W + reflect101(2 * x + i - 2, W)) * 3 + c];
          s += k[j] * rs;
        }
        dst[((size_t)y * dw + x) * 3 + c] = (uint8_t)((s + 128) >> 8);
      }
}

void flo_resize_nn_half_u8(const uint8_t* src, int W, int H, uint8_t* dst) {
  /* cv::resize(INTER_NEAREST) to (W/2, H/2): sx = min(floor(x * W/dw), W-1) (linemod.cpp:448, 731, 736) */
  int dw = W / 2, dh = H / 2;
  double fx = (double)W / dw, fy = (double)H / dh;
  for (int y = 0; y < dh; ++y) {
    int sy = (int)floor(y * fy); if (sy > H - 1) sy = H - 1;
    for (int x = 0; x < dw; ++x) {
      int sx = (int)floor(x * fx); if (sx > W - 1) sx = W - 1;
      dst[(size_t)y * dw + x] = src[(size_t)sy * W + sx];
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* depth normal modality                                                                      */
/* ------------------------------------------------------------------------------------------ */
void flo_median5_u8(const uint8_t* src, int W, int H, uint8_t* dst) {
  /* cv::medianBlur(ksize 5) on 8U: exact 13th of 25 with replicated borders (linemod.cpp:684).
   * Sliding 256-bin histogram with a 16-bin coarse level per row. */
  uint8_t* out = (src == dst) ? (uint8_t*)malloc((size_t)W * H) : dst;
  for (int y = 0; y < H; ++y) {
    int fine[256]; int coarse[16];
    memset(fine, 0, sizeof fine); memset(coarse, 0, sizeof coarse);
    const uint8_t* rows[5];
    for (int j = 0; j < 5; ++j) rows[j] = src + (size_t)clampi(y + j - 2, 0, H - 1) * W;
    for (int i = -2; i <= 2; ++i) { int sx = clampi(i, 0, W - 1); for (int j = 0; j < 5; ++j) { uint8_t v = rows[j][sx]; fine[v]++; coarse[v >> 4]++; } }
    for (int x = 0; x < W; ++x) {
      if (x > 0) {
        int xo = clampi(x - 3, 0, W - 1), xn = clampi(x + 2, 0, W - 1);
        for (int j = 0; j < 5; ++j) { uint8_t a = rows[j][xo], b = rows[j][xn]; fine[a]--; coarse[a >> 4]--; fine[b]++; coarse[b >> 4]++; }
      }
      int cum = 0, cb = 0;
      while (cum + coarse[cb] < 13) { cum += coarse[cb]; ++cb; }
      int v = cb << 4;
      while (cum + fine[v] < 13) { cum += fine[v]; ++v; }
      out[(size_t)y * W + x] = (uint8_t)v;
    }
  }
  if (out != dst) { memcpy(dst, out, (size_t)W * H); free(out); }
}

void flo_depth_quantize(const uint16_t* depth, int W, int H, int dist_thr, int diff_thr, uint8_t* out) {
  /* quantizedNormals, linemod.cpp:595-685 (accumBilateral :567-579 inlined) */
  ensure_tables();
  const int r = 5;
  uint8_t* raw = (uint8_t*)calloc((size_t)W * H, 1);
  static const int off[8][2] = {{-5, -5}, {0, -5}, {5, -5}, {-5, 0}, {5, 0}, {-5, 5}, {0, 5}, {5, 5}};   /* :633-640 (i = x, j = y) */
  for (int y = r; y < H - r - 1; ++y)
    for (int x = r; x < W - r - 1; ++x) {
      long d = depth[(size_t)y * W + x];
      uint8_t v = 0;
      if (d < dist_thr) {                                          /* :628 */
        long A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
        for (int k = 0; k < 8; ++k) {
          long i = off[k][0], j = off[k][1];
          long delta = (long)depth[(size_t)(y + j) * W + x + i] - d;
          long f = labs(delta) < diff_thr ? 1 : 0;                 /* :569 */
          long fi = f * i, fj = f * j;
          A0 += fi * i; A1 += fi * j; A3 += fj * j; b0 += fi * delta; b1 += fj * delta;
        }
        long det = A0 * A3 - A1 * A1;                              /* :643-645 */
        long ddx = A3 * b0 - A1 * b1;
        long ddy = -A1 * b0 + A0 * b1;
        float nx = (float)(617 * ddx), ny = (float)(617 * ddy), nz = (float)(-det * d);   /* :649-651 */
        float s = sqrtf(nx * nx + ny * ny + nz * nz);              /* :653 */
        if (s > 0) {
          float inv = 1.0f / s;                                    /* :657 */
          nx *= inv; ny *= inv; nz *= inv;
          int v1 = (int)(nx * 10 + 10), v2 = (int)(ny * 10 + 10), v3 = (int)(nz * 20 + 20);   /* :665-667 */
          (void)v3; /* table is v3-independent; v3 == 20 is the reference's out-of-bounds corner (SURVEY A.6 i) */
          v = g_normal_plane[clampi(v2, 0, 19) * 20 + clampi(v1, 0, 19)];
        }
      }
      raw[(size_t)y * W + x] = v;
    }
  flo_median5_u8(raw, W, H, out);                                  /* :684 */
  free(raw);
}

/* ------------------------------------------------------------------------------------------ */
/* spread / response maps / linear memories                                                   */
/* ------------------------------------------------------------------------------------------ */
void flo_spread(const uint8_t* q, int W, int H, int T, uint8_t* out) {
  /* spread + orUnaligned8u, linemod.cpp:882-965: T*T shifted full-image OR passes, like the reference */
  memset(out, 0, (size_t)W * H);
  for (int r = 0; r < T; ++r)
    for (int c = 0; c < T; ++c)
      for (int y = 0; y < H - r; ++y) {
        const uint8_t* __restrict s = q + (size_t)(y + r) * W + c;
        uint8_t* __restrict d = out + (size_t)y * W;
        int n = W - c;
        for (int x = 0; x < n; ++x) d[x] |= s[x];
      }
}

void flo_response_maps(const uint8_t* sp, int n, uint8_t* out8) {
  /* computeResponseMaps, linemod.cpp:979-1048 */
  ensure_tables();
  for (int ori = 0; ori < 8; ++ori) {
    const uint8_t* lo = g_sim_lut + 32 * ori; const uint8_t* hi = lo + 16;
    uint8_t* o = out8 + (size_t)ori * n;
    for (int i = 0; i < n; ++i) { uint8_t a = lo[sp[i] & 15], b = hi[sp[i] >> 4]; o[i] = a > b ? a : b; }
  }
}

void flo_linearize(const uint8_t* resp, int W, int H, int T, uint8_t* out) {
  /* linearize, linemod.cpp:1060-1088 */
  uint8_t* m = out;
  for (int r0 = 0; r0 < T; ++r0)
    for (int c0 = 0; c0 < T; ++c0)
      for (int r = r0; r < H; r += T) {
        const uint8_t* row = resp + (size_t)r * W;
        for (int c = c0; c < W; c += T) *m++ = row[c];
      }
}

/* ------------------------------------------------------------------------------------------ */
/* detector                                                                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int width, height, offset_x, offset_y, pyramid_level, feature_begin, feature_count; } tmpl_hdr;

struct flo_detector {
  int L, M;
  int T[8];
  int kind[4];
  float weak_thr; int dist_thr, diff_thr;
  int n_templates, n_features;
  tmpl_hdr* hdr; int32_t* feat; int32_t* class_of; int32_t* first_of_class; int n_classes;
  /* per frame */
  int W[8], H[8];
  uint8_t* quant[8][4];
  uint8_t* spr[8][4];
  uint8_t* lm[8][4];      /* 8 labels x (T*T*cells + LM_PAD), zero padded */
  size_t lm_stride[8];    /* bytes per label */
};

flo_detector* flo_detector_create(int L, const int* T, int M, const int* kind, float weak_thr, int dist_thr, int diff_thr) {
  if (L < 1 || L > 8 || M < 1 || M > 4) return NULL;
  flo_detector* d = (flo_detector*)calloc(1, sizeof *d);
  d->L = L; d->M = M;
  for (int l = 0; l < L; ++l) d->T[l] = T[l];
  for (int m = 0; m < M; ++m) d->kind[m] = kind[m];
  d->weak_thr = weak_thr; d->dist_thr = dist_thr; d->diff_thr = diff_thr;
  ensure_tables();
  return d;
}

static void free_frame(flo_detector* d) {
  for (int l = 0; l < 8; ++l) for (int m = 0; m < 4; ++m) {
    free(d->quant[l][m]); free(d->spr[l][m]); free(d->lm[l][m]);
    d->quant[l][m] = d->spr[l][m] = d->lm[l][m] = NULL;
  }
}

void flo_detector_destroy(flo_detector* d) {
  if (!d) return;
  free_frame(d);
  free(d->hdr); free(d->feat); free(d->class_of); free(d->first_of_class);
  free(d);
}

int flo_detector_set_templates(flo_detector* d, int n_templates, const int32_t* headers, const int32_t* features,
                               int n_features, const int32_t* class_of) {
  free(d->hdr); free(d->feat); free(d->class_of); free(d->first_of_class);
  int ne = n_templates * d->L * d->M;
  d->hdr = (tmpl_hdr*)malloc(sizeof(tmpl_hdr) * (size_t)(ne > 0 ? ne : 1));
  memcpy(d->hdr, headers, sizeof(tmpl_hdr) * (size_t)ne);
  d->feat = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)(n_features > 0 ? n_features : 1));
  memcpy(d->feat, features, sizeof(int32_t) * 3 * (size_t)n_features);
  d->class_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_templates > 0 ? n_templates : 1));
  memcpy(d->class_of, class_of, sizeof(int32_t) * (size_t)n_templates);
  d->n_templates = n_templates; d->n_features = n_features;
  int nc = 0;
  for (int t = 0; t < n_templates; ++t) { if (class_of[t] < 0) return -1; if (class_of[t] + 1 > nc) nc = class_of[t] + 1; }
  d->n_classes = nc;
  d->first_of_class = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nc > 0 ? nc : 1));
  for (int c = 0; c < nc; ++c) d->first_of_class[c] = -1;
  for (int t = 0; t < n_templates; ++t) {
    int c = class_of[t];
    if (d->first_of_class[c] < 0) d->first_of_class[c] = t;
    else if (class_of[t - 1] != c) return -1; /* classes must be contiguous */
  }
  for (int e = 0; e < ne; ++e) {
    const tmpl_hdr* h = &d->hdr[e];
    if (h->feature_begin < 0 || h->feature_count < 0 || h->feature_begin + h->feature_count > n_features) return -1;
    if (h->feature_count > 63) return -4;  /* CV_Assert(features.size() <= 63), linemod.cpp:1137, 1231 */
    for (int k = 0; k < h->feature_count; ++k) { int lab = d->feat[3 * (h->feature_begin + k) + 2]; if (lab < 0 || lab > 7) return -1; }
  }
  return 0;
}

int flo_detector_process(flo_detector* d, const uint8_t* bgr, const uint16_t* depth, int W, int H, const uint8_t* const* masks) {
  /* Detector::match front half, linemod.cpp:1369-1416 */
  free_frame(d);
  { int w = W, h = H;
    for (int l = 0; l < d->L; ++l) { int T = d->T[l]; if (w % T || h % T || (w * h) % 16) return -2; w /= 2; h /= 2; } }
  uint8_t* cur_bgr = NULL; /* colour pyramid source at the current level */
  uint8_t* cur_q[4] = {0, 0, 0, 0};
  uint8_t* cur_mask[4] = {0, 0, 0, 0};
  int w = W, h = H;
  for (int l = 0; l < d->L; ++l) {
    int T = d->T[l];
    if (l > 0) {
      int nw = w / 2, nh = h / 2;
      for (int m = 0; m < d->M; ++m) {
        if (d->kind[m] == 0) {                                     /* ColorGradientPyramid::pyrDown :434-453 */
          uint8_t* nb = (uint8_t*)malloc((size_t)nw * nh * 3);
          flo_pyrdown_bgr(cur_bgr ? cur_bgr : bgr, w, h, nb);
          free(cur_bgr); cur_bgr = nb;
          free(cur_q[m]); cur_q[m] = (uint8_t*)malloc((size_t)nw * nh);
          flo_color_quantize(cur_bgr, nw, nh, d->weak_thr, cur_q[m], NULL);
        } else {                                                   /* DepthNormalPyramid::pyrDown :721-739 */
          uint8_t* nq = (uint8_t*)malloc((size_t)nw * nh);
          flo_resize_nn_half_u8(cur_q[m], w, h, nq);
          free(cur_q[m]); cur_q[m] = nq;
        }
        if (cur_mask[m]) { uint8_t* nm = (uint8_t*)malloc((size_t)nw * nh); flo_resize_nn_half_u8(cur_mask[m], w, h, nm); free(cur_mask[m]); cur_mask[m] = nm; }
      }
      w = nw; h = nh;
    } else {
      for (int m = 0; m < d->M; ++m) {
        cur_q[m] = (uint8_t*)malloc((size_t)w * h);
        if (d->kind[m] == 0) flo_color_quantize(bgr, w, h, d->weak_thr, cur_q[m], NULL);
        else flo_depth_quantize(depth, w, h, d->dist_thr, d->diff_thr, cur_q[m]);
        if (masks && masks[m]) { cur_mask[m] = (uint8_t*)malloc((size_t)w * h); memcpy(cur_mask[m], masks[m], (size_t)w * h); }
      }
    }
    size_t n = (size_t)w * h;
    d->W[l] = w; d->H[l] = h;
    d->lm_stride[l] = n + LM_PAD;
    uint8_t* resp = (uint8_t*)malloc(8 * n);
    for (int m = 0; m < d->M; ++m) {
      uint8_t* q = (uint8_t*)malloc(n);                            /* quantize(): copyTo(dst, mask) :455-459, :741-745 */
      if (cur_mask[m]) { for (size_t i = 0; i < n; ++i) q[i] = cur_mask[m][i] ? cur_q[m][i] : 0; } else memcpy(q, cur_q[m], n);
      d->quant[l][m] = q;
      d->spr[l][m] = (uint8_t*)malloc(n);
      flo_spread(q, w, h, T, d->spr[l][m]);
      flo_response_maps(d->spr[l][m], (int)n, resp);
      d->lm[l][m] = (uint8_t*)calloc(8 * d->lm_stride[l], 1);
      for (int j = 0; j < 8; ++j) flo_linearize(resp + (size_t)j * n, w, h, T, d->lm[l][m] + (size_t)j * d->lm_stride[l]);
    }
    free(resp);
  }
  free(cur_bgr);
  for (int m = 0; m < 4; ++m) { free(cur_q[m]); free(cur_mask[m]); }
  return 0;
}

const uint8_t* flo_detector_quantized(const flo_detector* d, int l, int m, int* W, int* H) { if (W) *W = d->W[l]; if (H) *H = d->H[l]; return d->quant[l][m]; }
const uint8_t* flo_detector_spread(const flo_detector* d, int l, int m) { return d->spr[l][m]; }
const uint8_t* flo_detector_lm(const flo_detector* d, int l, int m, int label, int* rows, int* cols) {
  int T = d->T[l];
  if (rows) *rows = T * T;
  if (cols) *cols = (d->W[l] / T) * (d->H[l] / T);
  return d->lm[l][m] + (size_t)label * d->lm_stride[l];
}

/* accessLinearMemory, linemod.cpp:1094-1117.  The reference returns a raw pointer into a continuous [T*T][cells] Mat and
 * the callers read past the end of a row when a feature sits on the template's last row/column (x == width is legal,
 * cropTemplates :79-80): the read continues into the next row (defined) or, for the last row, out of the buffer
 * (undefined).  Flat addressing over a zero-padded buffer keeps the former and defines the latter as zeros. */
static inline const uint8_t* access_lm(const flo_detector* d, int l, int m, int x, int y, int label) {
  int T = d->T[l], Wd = d->W[l] / T, cells = Wd * (d->H[l] / T);
  return d->lm[l][m] + (size_t)label * d->lm_stride[l] + (size_t)((y % T) * T + (x % T)) * cells + (size_t)(y / T) * Wd + x / T;
}

/* similarity(), linemod.cpp:1130-1214: dst is a freshly zeroed H'xW' u8 map (allocated per call like the reference). */
static uint8_t* similarity_map(const flo_detector* d, int l, int m, const tmpl_hdr* h) {
  int T = d->T[l], Wd = d->W[l] / T, Hd = d->H[l] / T;
  uint8_t* dst = (uint8_t*)calloc((size_t)Wd * Hd, 1);                       /* :1160 */
  int wf = (h->width - 1) / T + 1, hf = (h->height - 1) / T + 1;             /* :1145-1146 */
  int span_x = Wd - wf, span_y = Hd - hf;
  int tp = span_y * Wd + span_x + 1;                                         /* :1155 */
  if (tp > Wd * Hd) tp = Wd * Hd;
  for (int i = 0; i < h->feature_count; ++i) {
    const int32_t* f = d->feat + 3 * (h->feature_begin + i);
    if (f[0] < 0 || f[0] >= d->W[l] || f[1] < 0 || f[1] >= d->H[l]) continue;   /* :1179 */
    const uint8_t* __restrict lmp = access_lm(d, l, m, f[0], f[1], f[2]);
    uint8_t* __restrict o = dst;
    for (int j = 0; j < tp; ++j) o[j] = (uint8_t)(o[j] + lmp[j]);            /* :1191-1212 (vectorised by the compiler: paddb) */
  }
  return dst;
}

/* similarityLocal(), linemod.cpp:1226-1300 */
static void similarity_local(const flo_detector* d, int l, int m, const tmpl_hdr* h, int cx, int cy, uint8_t dst[256]) {
  int T = d->T[l], Wd = d->W[l] / T;
  memset(dst, 0, 256);
  int ox = (cx / T - 8) * T, oy = (cy / T - 8) * T;                          /* :1240-1241 */
  for (int i = 0; i < h->feature_count; ++i) {
    const int32_t* f = d->feat + 3 * (h->feature_begin + i);
    int x = f[0] + ox, y = f[1] + oy;
    if (x < 0 || y < 0 || x >= d->W[l] || y >= d->H[l]) continue;            /* :1257 */
    const uint8_t* lmp = access_lm(d, l, m, x, y, f[2]);
    for (int row = 0; row < 16; ++row) {
      for (int col = 0; col < 16; ++col) dst[row * 16 + col] = (uint8_t)(dst[row * 16 + col] + lmp[col]);
      lmp += Wd;
    }
  }
}

int flo_detector_similarity(const flo_detector* d, int t, uint16_t* out) {
  int l = d->L - 1, T = d->T[l], n = (d->W[l] / T) * (d->H[l] / T);
  memset(out, 0, sizeof(uint16_t) * (size_t)n);
  for (int m = 0; m < d->M; ++m) {
    const tmpl_hdr* h = &d->hdr[(t * d->L + l) * d->M + m];
    uint8_t* s = similarity_map(d, l, m, h);
    for (int i = 0; i < n; ++i) out[i] = (uint16_t)(out[i] + s[i]);
    free(s);
  }
  return n;
}

typedef struct { flo_match_t* v; int n, cap; } mvec;
static void mv_push(mvec* a, flo_match_t m) {
  if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (flo_match_t*)realloc(a->v, sizeof(flo_match_t) * (size_t)a->cap); }
  a->v[a->n++] = m;
}

/* one iteration of the template loop of matchClass, linemod.cpp:1458-1576 */
static void match_one_template(const flo_detector* d, int t, float threshold, mvec* out) {
  const int L = d->L, M = d->M;
  const int cls = d->class_of[t], tid = t - d->first_of_class[cls];
  int l = L - 1, T = d->T[l], Wd = d->W[l] / T, Hd = d->H[l] / T, n = Wd * Hd;
  uint16_t* total = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)n);
  int nf = 0;
  for (int m = 0; m < M; ++m) {                                              /* :1471-1481 */
    const tmpl_hdr* h = &d->hdr[(t * L + l) * M + m];
    nf += h->feature_count;
    uint8_t* s = similarity_map(d, l, m, h);
    if (m == 0) for (int i = 0; i < n; ++i) total[i] = s[i]; else for (int i = 0; i < n; ++i) total[i] = (uint16_t)(total[i] + s[i]);
    free(s);
  }
  int raw_thr = (int)(2 * nf + (threshold / 100.f) * (2 * nf) + 0.5f);       /* :1487 */
  mvec c = {0, 0, 0};
  int off = T / 2 + (T % 2 - 1);                                             /* :1499 */
  for (int r = 0; r < Hd; ++r)
    for (int col = 0; col < Wd; ++col) {
      int raw = total[r * Wd + col];
      if (raw > raw_thr) {
        flo_match_t mm = {col * T + off, r * T + off, (raw * 100.f) / (4 * nf) + 0.5f, cls, tid};   /* :1500-1503 */
        mv_push(&c, mm);
      }
    }
  free(total);
  for (l = L - 2; l >= 0; --l) {                                             /* :1509-1573 */
    T = d->T[l];
    int border = 8 * T;
    off = T / 2 + (T % 2 - 1);
    const tmpl_hdr* h0 = &d->hdr[(t * L + l) * M];
    int max_x = d->W[l] - h0->width - border, max_y = d->H[l] - h0->height - border;
    for (int k = 0; k < c.n; ++k) {
      flo_match_t* mt = &c.v[k];
      int x = mt->x * 2 + 1, y = mt->y * 2 + 1;
      if (x < border) x = border;
      if (y < border) y = border;
      if (x > max_x) x = max_x;
      if (y > max_y) y = max_y;
      uint16_t tot[256]; memset(tot, 0, sizeof tot);
      int nf2 = 0;
      for (int m = 0; m < M; ++m) {
        const tmpl_hdr* h = &d->hdr[(t * L + l) * M + m];
        nf2 += h->feature_count;
        uint8_t loc[256];
        similarity_local(d, l, m, h, x, y, loc);
        for (int i = 0; i < 256; ++i) tot[i] = (uint16_t)(tot[i] + loc[i]);
      }
      int best = 0, br = -1, bc = -1;
      for (int r = 0; r < 16; ++r) for (int col = 0; col < 16; ++col) { int s = tot[r * 16 + col]; if (s > best) { best = s; br = r; bc = col; } }
      mt->x = (x / T - 8 + bc) * T + off;                                    /* :1564-1566 */
      mt->y = (y / T - 8 + br) * T + off;
      mt->similarity = (best * 100.f) / (4 * nf2);
    }
    int w = 0;
    for (int k = 0; k < c.n; ++k) if (!(c.v[k].similarity < threshold)) c.v[w++] = c.v[k];   /* :1570-1572 */
    c.n = w;
  }
  for (int k = 0; k < c.n; ++k) mv_push(out, c.v[k]);
  free(c.v);
}

static int canon_cmp(const void* pa, const void* pb) {
  const flo_match_t* a = (const flo_match_t*)pa; const flo_match_t* b = (const flo_match_t*)pb;
  if (a->similarity != b->similarity) return a->similarity > b->similarity ? -1 : 1;   /* linemod.hpp:262-269 */
  if (a->template_id != b->template_id) return a->template_id < b->template_id ? -1 : 1;
  if (a->class_idx != b->class_idx) return a->class_idx < b->class_idx ? -1 : 1;        /* canonical tie-breaks */
  if (a->y != b->y) return a->y < b->y ? -1 : 1;
  if (a->x != b->x) return a->x < b->x ? -1 : 1;
  return 0;
}

int flo_detector_match_templates(const flo_detector* d, float threshold, const int32_t* class_filter, int n_filter,
                                 int canonical, int n_threads, flo_match_t* out, int cap, int* n_total) {
  mvec all = {0, 0, 0};
  int nt = d->n_templates;
  uint8_t* want = (uint8_t*)malloc((size_t)(nt > 0 ? nt : 1));
  for (int t = 0; t < nt; ++t) {
    want[t] = 1;
    if (n_filter > 0) { want[t] = 0; for (int k = 0; k < n_filter; ++k) if (class_filter[k] == d->class_of[t]) want[t] = 1; }
  }
  if (n_threads <= 1) {
    for (int t = 0; t < nt; ++t) if (want[t]) match_one_template(d, t, threshold, &all);
  } else {
#ifdef _OPENMP
    mvec* per = (mvec*)calloc((size_t)(nt > 0 ? nt : 1), sizeof(mvec));
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads)
    for (int t = 0; t < nt; ++t) if (want[t]) match_one_template(d, t, threshold, &per[t]);
    for (int t = 0; t < nt; ++t) { for (int k = 0; k < per[t].n; ++k) mv_push(&all, per[t].v[k]); free(per[t].v); }
    free(per);
#else
    for (int t = 0; t < nt; ++t) if (want[t]) match_one_template(d, t, threshold, &all);
#endif
  }
  free(want);
  if (canonical && all.n > 0) {                                              /* :1437-1439 under the canonical total order */
    qsort(all.v, (size_t)all.n, sizeof(flo_match_t), canon_cmp);
    int w = 1;
    for (int k = 1; k < all.n; ++k) {
      const flo_match_t* a = &all.v[w - 1]; const flo_match_t* b = &all.v[k];
      if (a->x == b->x && a->y == b->y && a->similarity == b->similarity && a->class_idx == b->class_idx) continue;   /* linemod.hpp:271-274 */
      all.v[w++] = *b;
    }
    all.n = w;
  }
  if (n_total) *n_total = all.n;
  int nw = all.n < cap ? all.n : cap;
  if (nw > 0) memcpy(out, all.v, sizeof(flo_match_t) * (size_t)nw);
  free(all.v);
  return nw;
}

/* ------------------------------------------------------------------------------------------ */
/* ICP                                                                                        */
/* ------------------------------------------------------------------------------------------ */
void flo_depth_to_3d_mm(const uint16_t* depth, int W, int H, float fx, float fy, float cx, float cy, float* out3) {
  /* depthTo3dNoMask<float> (depth_to_3d.cpp:99-137) after rescaleDepth (244-260), then scale_mat_vec3f(.,1000)
   * (common.cpp:418-425; called at detection.cpp:39-40) */
  const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy;
  for (int y = 0; y < H; ++y) {
    float yc = ((float)y - cy) * inv_fy;
    for (int x = 0; x < W; ++x) {
      float xc = ((float)x - cx) * inv_fx;
      uint16_t dv = depth[(size_t)y * W + x];
      float z = dv == 0 ? NAN : (float)dv * (float)(1 / 1000.0);     /* convertTo scales in float */
      float* o = out3 + ((size_t)y * W + x) * 3;
      o[0] = (xc * z) * 1000.0f; o[1] = (yc * z) * 1000.0f; o[2] = z * 1000.0f;
    }
  }
}

static inline int pt_valid(const float* p) { return p[2] <= 900.0f; }   /* is_vec3f_valid, common.cpp:261-266 (NaN fails) */

int flo_pair_points(const float* ref3, const float* mod3, int W, int H, const int rr[4], const int rm[4], float* pts_ref, float* pts_mod) {
  /* crops (detection.cpp:43-44) + paired matToVec (common.cpp:382-405); iterates the ref crop, model crop in lockstep */
  (void)H;
  int n = 0;
  int total = rr[2] * rr[3], totm = rm[2] * rm[3];
  for (int i = 0; i < total && i < totm; ++i) {
    const float* pr = ref3 + ((size_t)(rr[1] + i / rr[2]) * W + rr[0] + i % rr[2]) * 3;
    const float* pm = mod3 + ((size_t)(rm[1] + i / rm[2]) * W + rm[0] + i % rm[2]) * 3;
    if (!pt_valid(pr) || !pt_valid(pm)) continue;
    memcpy(pts_ref + 3 * (size_t)n, pr, 12); memcpy(pts_mod + 3 * (size_t)n, pm, 12);
    ++n;
  }
  return n;
}

static void get_mean(const float* p, int n, float c[3]) {                /* getMean, ICP.cpp:8-25 */
  c[0] = c[1] = c[2] = 0.f;
  for (int i = 0; i < n; ++i) { c[0] += p[3 * i]; c[1] += p[3 * i + 1]; c[2] += p[3 * i + 2]; }
  if (n > 0) { c[0] /= (float)n; c[1] /= (float)n; c[2] /= (float)n; }
}

static inline void matvec(const float R[9], const float v[3], float o[3]) {   /* Matx33f * Vec3f: s = 0; s += a*b */
  for (int i = 0; i < 3; ++i) { float s = 0.f; s += R[3 * i] * v[0]; s += R[3 * i + 1] * v[1]; s += R[3 * i + 2] * v[2]; o[i] = s; }
}
static inline void matmul(const float A[9], const float B[9], float C[9]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { float s = 0.f; for (int k = 0; k < 3; ++k) s += A[3 * i + k] * B[3 * k + j]; C[3 * i + j] = s; }
}

static void transform_points(float* p, int n, const float R[9], const float T[3]) {   /* transformPoints in place, ICP.cpp:28-45 */
  for (int i = 0; i < n; ++i) {
    float* q = p + 3 * i;
    if (!pt_valid(q)) continue;
    float o[3]; matvec(R, q, o);
    q[0] = o[0] + T[0]; q[1] = o[1] + T[1]; q[2] = o[2] + T[2];
  }
}

static void copy_points(const float* src, int n, float* dst) {           /* copyPoints, ICP.cpp:48-65 */
  memset(dst, 0, sizeof(float) * 3 * (size_t)n);
  for (int i = 0; i < n; ++i) if (pt_valid(src + 3 * i)) memcpy(dst + 3 * i, src + 3 * i, 12);
}

static float l2_dist_clouds(const float* model, int n_model, const float* ref, float* dist_mean, float thr) {   /* ICP.cpp:68-111 */
  int nin = 0, counter = 0; float ratio = 0.f;
  *dist_mean = 0.f;
  for (int i = 0; i < n_model; ++i) {
    const float* r = ref + 3 * i; const float* m = model + 3 * i;
    if (!pt_valid(r) || !pt_valid(m)) continue;
    float dx = m[0] - r[0], dy = m[1] - r[1], dz = m[2] - r[2];
    float dist = (float)sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);   /* cv::norm(Vec3f): double accumulate */
    if (dist <= thr) { *dist_mean += dist; ++nin; }
    ++counter;
  }
  if (counter > 0) { *dist_mean /= (float)nin; ratio = (float)nin / (float)counter; }
  else *dist_mean = FLT_MAX;
  return ratio;
}

/* --- exact 1-NN: a KD-tree standing in for cvflann::KDTreeSingleIndex (ICP.cpp:658-659, 228).  FLANN is an un-vendored
 * dependency (inside OpenCV); its single KD-tree with SearchParams() defaults (eps = 0) returns the exact nearest
 * neighbour, and L2_Simple accumulates (a-b)^2 over x,y,z in fp32.  Any exact search with that distance formula gives the
 * same distances; indices can differ only on exact ties. --- */
typedef struct { int lo, hi, dim, left, right; float split; } kdnode;
typedef struct { kdnode* nodes; int n_nodes, cap; int* idx; const float* pts; } kdtree;

static int kd_build(kdtree* t, int lo, int hi) {
  if (t->n_nodes == t->cap) { t->cap *= 2; t->nodes = (kdnode*)realloc(t->nodes, sizeof(kdnode) * (size_t)t->cap); }
  int id = t->n_nodes++;
  kdnode nd; nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1; nd.dim = 0; nd.split = 0.f;
  if (hi - lo > 15) {                                               /* KDTreeSingleIndexParams(15): leaf_max_size */
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = lo; i < hi; ++i) for (int k = 0; k < 3; ++k) { float v = t->pts[3 * t->idx[i] + k]; if (v < mn[k]) mn[k] = v; if (v > mx[k]) mx[k] = v; }
    int dim = 0; for (int k = 1; k < 3; ++k) if (mx[k] - mn[k] > mx[dim] - mn[dim]) dim = k;
    if (mx[dim] > mn[dim]) {
      /* median split by nth_element-style quickselect on idx[lo..hi) */
      int mid = (lo + hi) / 2, a = lo, b = hi - 1;
      while (a < b) {
        float pv = t->pts[3 * t->idx[(a + b) / 2] + dim];
        int i = a, j = b;
        while (i <= j) {
          while (t->pts[3 * t->idx[i] + dim] < pv) ++i;
          while (t->pts[3 * t->idx[j] + dim] > pv) --j;
          if (i <= j) { int tmp = t->idx[i]; t->idx[i] = t->idx[j]; t->idx[j] = tmp; ++i; --j; }
        }
        if (mid <= j) b = j; else if (mid >= i) a = i; else break;
      }
      nd.dim = dim; nd.split = t->pts[3 * t->idx[mid] + dim];
      t->nodes[id] = nd;
      int left = kd_build(t, lo, mid);
      int right = kd_build(t, mid, hi);
      t->nodes[id].left = left; t->nodes[id].right = right;
      return id;
    }
  }
  t->nodes[id] = nd;
  return id;
}

static void kd_search(const kdtree* t, int id, const float q[3], float* best, int* best_i) {
  const kdnode* nd = &t->nodes[id];
  if (nd->left < 0) {
    for (int i = nd->lo; i < nd->hi; ++i) {
      const float* p = t->pts + 3 * t->idx[i];
      float d0 = q[0] - p[0], d1 = q[1] - p[1], d2 = q[2] - p[2];
      float dd = 0.f; dd += d0 * d0; dd += d1 * d1; dd += d2 * d2;   /* L2_Simple */
      if (dd < *best || (dd == *best && t->idx[i] < *best_i)) { *best = dd; *best_i = t->idx[i]; }
    }
    return;
  }
  float diff = q[nd->dim] - nd->split;
  int nearc = diff < 0 ? nd->left : nd->right, farc = diff < 0 ? nd->right : nd->left;
  kd_search(t, nearc, q, best, best_i);
  if (diff * diff <= *best) kd_search(t, farc, q, best, best_i);
}

/* cv::SVD::compute on a 3x3 float matrix: OpenCV's one-sided (Hestenes) Jacobi in fp32 with double dot products
 * (un-vendored dependency; algorithm restated).  Only R = V * U^T is used (ICP.cpp:744), which is insensitive to the
 * sign / order conventions of the factorisation. */
void flo_svd3_rot(const float cov[9], float R[9]) {
  /* work on At (rows = columns of A): after convergence rows of At are sigma_i * u_i, rows of Vt are v_i */
  float At[3][3], Vt[3][3];
  double Wd[3];
  for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) { At[i][k] = cov[3 * k + i]; Vt[i][k] = (i == k); }
  for (int i = 0; i < 3; ++i) { double sd = 0; for (int k = 0; k < 3; ++k) sd += (double)At[i][k] * At[i][k]; Wd[i] = sd; }
  const float eps = FLT_EPSILON * 2;
  for (int iter = 0; iter < 30; ++iter) {
    int changed = 0;
    for (int i = 0; i < 2; ++i)
      for (int j = i + 1; j < 3; ++j) {
        double a = Wd[i], p = 0, b = Wd[j];
        for (int k = 0; k < 3; ++k) p += (double)At[i][k] * At[j][k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        double beta = a - b, gamma = sqrt(p * p + beta * beta);   /* (OpenCV calls hypot; same to well below fp32 resolution, and sqrt is bit-reproducible on the GPU) */
        float c, s;
        if (beta < 0) { double delta = (gamma - beta) * 0.5; s = (float)sqrt(delta / gamma); c = (float)(p / (gamma * s * 2)); }
        else { c = (float)sqrt((gamma + beta) / (gamma * 2)); s = (float)(p / (gamma * c * 2)); }
        a = b = 0;
        for (int k = 0; k < 3; ++k) {
          float t0 = c * At[i][k] + s * At[j][k], t1 = -s * At[i][k] + c * At[j][k];
          At[i][k] = t0; At[j][k] = t1;
          a += (double)t0 * t0; b += (double)t1 * t1;
        }
        Wd[i] = a; Wd[j] = b;
        changed = 1;
        for (int k = 0; k < 3; ++k) {
          float t0 = c * Vt[i][k] + s * Vt[j][k], t1 = -s * Vt[i][k] + c * Vt[j][k];
          Vt[i][k] = t0; Vt[j][k] = t1;
        }
      }
    if (!changed) break;
  }
  /* U columns = At rows / sigma.  R = V * U^T = sum_i v_i u_i^T, R[r][c] = sum_i Vt[i][r] * U[c][i] */
  float U[3][3]; /* U[i] = u_i */
  for (int i = 0; i < 3; ++i) {
    double sd = 0; for (int k = 0; k < 3; ++k) sd += (double)At[i][k] * At[i][k];
    sd = sqrt(sd);
    float s = (float)(sd > DBL_MIN ? 1 / sd : 0.);
    for (int k = 0; k < 3; ++k) U[i][k] = At[i][k] * s;
  }
  /* Mat(vt.t() * u.t()) is a cv::gemm with both operands transposed: its kernel for CV_32F (GEMMSingleMul<float,double>)
   * accumulates the three products in double and rounds once (checked against oracle/_ref and cv2.gemm). */
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
    double s = 0.;
    for (int i = 0; i < 3; ++i) s += (double)Vt[i][r] * (double)U[i][c];
    R[3 * r + c] = (float)s;
  }
}

float flo_icp_cloud_to_cloud_ex(const float* pts_ref, int n_ref, const float* pts_model, int n_model, float R[9], float T[3],
                                float* inlier_ratio, int icp_it_thr, float dist_mean_thr, float dist_diff_thr, int* iterations,
                                float* trace) {
  /* icpCloudToCloud_Ex, ICP.cpp:617-809.  n_ref >= n_model is assumed by the reference's lock-step loops. */
  if (iterations) *iterations = 0;
  if (n_model < 3 || n_ref < 3) {                                           /* :633-638 (R,T stay value-initialised zeros) */
    memset(R, 0, 36); memset(T, 0, 12); if (inlier_ratio) *inlier_ratio = 0.f;
    return -1.f;
  }
  static const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(R, I3, 36); T[0] = T[1] = T[2] = 0.f;                              /* :644-645 */
  kdtree kt; kt.cap = 2 * n_ref / 8 + 16; kt.nodes = (kdnode*)malloc(sizeof(kdnode) * (size_t)kt.cap); kt.n_nodes = 0;
  kt.idx = (int*)malloc(sizeof(int) * (size_t)n_ref); kt.pts = pts_ref;
  for (int i = 0; i < n_ref; ++i) kt.idx[i] = i;
  kd_build(&kt, 0, n_ref);                                                  /* :650-659 */
  float* tmp = (float*)malloc(sizeof(float) * 3 * (size_t)n_model);
  float* cor_m = (float*)malloc(sizeof(float) * 3 * (size_t)(n_model > n_ref ? n_model : n_ref));
  float* cor_r = (float*)malloc(sizeof(float) * 3 * (size_t)(n_model > n_ref ? n_model : n_ref));
  copy_points(pts_model, n_model, tmp);                                     /* :667 */
  float dist_mean = 0.f;
  float ratio = l2_dist_clouds(tmp, n_model, pts_ref, &dist_mean, FLT_MAX);  /* :670 */
  float dist_diff = FLT_MAX;
  int iter = 0;
  while (dist_mean > dist_mean_thr && dist_diff > dist_diff_thr && iter < icp_it_thr) {   /* :684 */
    ++iter;
    int n_cm, n_cr;
    if (iter == 1) {                                                        /* :700-704 */
      copy_points(tmp, n_model, cor_m); n_cm = n_model;
      copy_points(pts_ref, n_ref, cor_r); n_cr = n_ref;
    } else {                                                                /* :708 -> :193-279 */
      float thr = 3 * dist_mean;
      n_cm = 0;
      for (int i = 0; i < n_model; ++i) {
        float best = FLT_MAX; int bi = -1;
        kd_search(&kt, 0, tmp + 3 * i, &best, &bi);
        if (best <= thr) { memcpy(cor_m + 3 * n_cm, tmp + 3 * i, 12); memcpy(cor_r + 3 * n_cm, pts_ref + 3 * bi, 12); ++n_cm; }   /* :266-273 */
      }
      n_cr = n_cm;
    }
    if (n_cr < 3 || n_cm < 3) { iter = icp_it_thr; continue; }             /* :711-715 */
    float mc[3], rc[3];
    get_mean(cor_m, n_cm, mc); get_mean(cor_r, n_cr, rc);                   /* :722-724 */
    float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};                             /* :731-735 (not centred) */
    for (int i = 0; i < n_cm; ++i) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cov[3 * a + b] += cor_m[3 * i + a] * cor_r[3 * i + b];
    float Ropt[9], Topt[3], rm[3];
    flo_svd3_rot(cov, Ropt);                                                /* :741-744 */
    matvec(Ropt, mc, rm);
    for (int k = 0; k < 3; ++k) Topt[k] = rc[k] - rm[k];                    /* :747 */
    int finite = 1;
    for (int k = 0; k < 9; ++k) if (!isfinite(Ropt[k])) finite = 0;
    for (int k = 0; k < 3; ++k) if (!isfinite(Topt[k])) finite = 0;
    if (!finite) continue;                                                  /* :748-749 */
    transform_points(tmp, n_model, Ropt, Topt);                             /* :756 */
    dist_diff = dist_mean;                                                  /* :778-780 */
    ratio = l2_dist_clouds(tmp, n_model, pts_ref, &dist_mean, 3 * dist_mean);
    dist_diff -= dist_mean;
    float nT[3], nR[9];
    matvec(Ropt, T, nT);                                                    /* :793-797 */
    for (int k = 0; k < 3; ++k) T[k] = nT[k] + Topt[k];
    matmul(Ropt, R, nR); memcpy(R, nR, 36);
    if (trace) { trace[3 * (iter - 1)] = dist_mean; trace[3 * (iter - 1) + 1] = dist_diff; trace[3 * (iter - 1) + 2] = (float)n_cm; }
  }
  if (iterations) *iterations = iter;
  if (inlier_ratio) *inlier_ratio = ratio;
  free(tmp); free(cor_m); free(cor_r); free(kt.nodes); free(kt.idx);
  return dist_mean;
}

int flo_detection(const uint16_t* model_depth, const uint16_t* ref_depth, int W, int H, const float K_ref[4],
                  const int rect_model[4], const int rect_ref[4], int icp_it_thr, float dist_mean_thr, float dist_diff_thr,
                  const float r_match[9], const float t_match[3], float d_match, float T_final[3], float R_final[9],
                  float* dist_mean_out, float* inlier_ratio, int* iterations, int* n_points) {
  /* detection(), ICP/detection.cpp:11-254, test_id == 2 (:147, :175-178) */
  (void)d_match;
  const int* rs[2] = {rect_model, rect_ref};
  for (int k = 0; k < 2; ++k) { const int* r = rs[k]; if (r[0] < 0 || r[1] < 0 || r[2] < 0 || r[3] < 0 || r[0] + r[2] > W || r[1] + r[3] > H) return -3; }
  size_t n = (size_t)W * H;
  float* ref3 = (float*)malloc(sizeof(float) * 3 * n);
  float* mod3 = (float*)malloc(sizeof(float) * 3 * n);
  flo_depth_to_3d_mm(ref_depth, W, H, K_ref[0], K_ref[1], K_ref[2], K_ref[3], ref3);        /* :31-32, :39 */
  flo_depth_to_3d_mm(model_depth, W, H, 608.f, 608.f, 320.f, 240.f, mod3);                  /* :35-36, :40; common.cpp:358 */
  size_t cap = (size_t)rect_ref[2] * rect_ref[3] + 1;
  float* pr = (float*)malloc(sizeof(float) * 3 * cap);
  float* pm = (float*)malloc(sizeof(float) * 3 * cap);
  int np = flo_pair_points(ref3, mod3, W, H, rect_ref, rect_model, pr, pm);                 /* :43-44, :114 */
  free(ref3); free(mod3);
  float mc[3], rc[3], t_tmp[3], t_init[3];
  get_mean(pm, np, mc); get_mean(pr, np, rc);                                               /* :165-166 */
  for (int k = 0; k < 3; ++k) { t_tmp[k] = rc[k] - mc[k]; t_init[k] = t_tmp[k] + t_match[k]; }   /* :177, :199 */
  static const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  transform_points(pm, np, I3, t_tmp);                                                      /* :206 */
  float R[9], T[3], ratio = 0.f; int it = 0;
  float dm = flo_icp_cloud_to_cloud_ex(pr, np, pm, np, R, T, &ratio, icp_it_thr, dist_mean_thr, dist_diff_thr, &it, NULL);   /* :228 */
  float rt[3]; matvec(R, t_init, rt);
  for (int k = 0; k < 3; ++k) T_final[k] = rt[k] + T[k];                                    /* :232-233 */
  matmul(R, r_match, R_final);                                                              /* :234 */
  if (dist_mean_out) *dist_mean_out = dm;
  if (inlier_ratio) *inlier_ratio = ratio;
  if (iterations) *iterations = it;
  if (n_points) *n_points = np;
  free(pr); free(pm);
  return 0;
}

int flo_nms(const float* t3, const int32_t* n_model_pts, const float* icp_dist, int n, float th, int32_t* out_idx) {
  /* nonMaximumSuppression, NMS.cpp:6-39 */
  uint8_t* done = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    if (done[i]) continue;
    int win = i;
    int size_th = (int)((float)n_model_pts[i] * 0.85);                       /* :17 (float * double literal) */
    for (int j = i + 1; j < n; ++j) {
      if (done[j]) continue;
      double s = 0;
      for (int k = 0; k < 3; ++k) { double dd = (double)(t3[3 * win + k] - t3[3 * j + k]); s += dd * dd; }   /* cv::norm(Mat,Mat): normL2Sqr<float,double>, v = double(a - b) with the difference taken in fp32 */
      if (sqrt(s) < th) {
        done[j] = 1;
        if (n_model_pts[j] > size_th && icp_dist[j] < icp_dist[win]) win = j;   /* :27-28 */
      }
    }
    out_idx[cnt++] = win;
  }
  free(done);
  return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* template training: Detector::addTemplate (linemod.cpp:1579-1615)                            */
/* ------------------------------------------------------------------------------------------ */
void flo_erode3_u8(const uint8_t* src, int W, int H, int iterations, uint8_t* dst) {
  /* cv::erode, default 3x3 rectangle, BORDER_REPLICATE (linemod.cpp:466, 753) */
  size_t n = (size_t)W * H;
  uint8_t* cur = (uint8_t*)malloc(n); uint8_t* nxt = (uint8_t*)malloc(n);
  memcpy(cur, src, n);
  for (int it = 0; it < iterations; ++it) {
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int m = 255;
        for (int j = -1; j <= 1; ++j) for (int i = -1; i <= 1; ++i) { int v = cur[(size_t)clampi(y + j, 0, H - 1) * W + clampi(x + i, 0, W - 1)]; if (v < m) m = v; }
        nxt[(size_t)y * W + x] = (uint8_t)m;
      }
    uint8_t* t = cur; cur = nxt; nxt = t;
  }
  memcpy(dst, cur, n);
  free(cur); free(nxt);
}

void flo_distance_c3(const uint8_t* src, int W, int H, float* dst) {
  /* cv::distanceTransform(DIST_C, 3) (linemod.cpp:765): chessboard distance to the nearest zero pixel, as the exact two-pass 3x3
   * chamfer with unit weights computes it; an image without a zero pixel gives 65535 (OpenCV 4.13's own code). */
  size_t n = (size_t)W * H;
  int any_zero = 0;
  for (size_t i = 0; i < n; ++i) if (!src[i]) { any_zero = 1; break; }
  if (!any_zero) { for (size_t i = 0; i < n; ++i) dst[i] = 65535.f; return; }
  const int INF = 1 << 28;
  int* t = (int*)malloc(sizeof(int) * n);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int v = INF;
      if (!src[(size_t)y * W + x]) v = 0;
      else {
        if (y > 0) { for (int i = -1; i <= 1; ++i) { int xx = x + i; if (xx >= 0 && xx < W) { int u = t[(size_t)(y - 1) * W + xx] + 1; if (u < v) v = u; } } }
        if (x > 0) { int u = t[(size_t)y * W + x - 1] + 1; if (u < v) v = u; }
      }
      t[(size_t)y * W + x] = v;
    }
  for (int y = H - 1; y >= 0; --y)
    for (int x = W - 1; x >= 0; --x) {
      int v = t[(size_t)y * W + x];
      if (y < H - 1) { for (int i = -1; i <= 1; ++i) { int xx = x + i; if (xx >= 0 && xx < W) { int u = t[(size_t)(y + 1) * W + xx] + 1; if (u < v) v = u; } } }
      if (x < W - 1) { int u = t[(size_t)y * W + x + 1] + 1; if (u < v) v = u; }
      t[(size_t)y * W + x] = v;
      dst[(size_t)y * W + x] = (float)v;
    }
  free(t);
}

typedef struct { int x, y, label; float score; int order; } train_cand;
static int cand_cmp(const void* a, const void* b) {
  /* Candidate::operator< (linemod.hpp:124-128): higher score first; std::stable_sort keeps the row-major order among ties */
  const train_cand* p = (const train_cand*)a; const train_cand* q = (const train_cand*)b;
  if (p->score > q->score) return -1;
  if (p->score < q->score) return 1;
  return p->order - q->order;
}
/* QuantizedPyramid::selectScatteredFeatures (linemod.cpp:134-163); out = x, y, label triples */
static void select_scattered(const train_cand* c, int n_cand, int32_t* out, int num_features, float distance) {
  float distance_sq = distance * distance;
  int n = 0, i = 0;
  while (n < num_features) {
    int keep = 1;
    for (int j = 0; j < n && keep; ++j) {
      int dx = c[i].x - out[3 * j], dy = c[i].y - out[3 * j + 1];
      keep = (float)(dx * dx + dy * dy) >= distance_sq;
    }
    if (keep) { out[3 * n] = c[i].x; out[3 * n + 1] = c[i].y; out[3 * n + 2] = c[i].label; ++n; }
    if (++i == n_cand) { i = 0; distance -= 1.0f; distance_sq = distance * distance; }
  }
}
static int label_of(int q) { int l = 0; while (l < 7 && !(q & (1 << l))) ++l; return l; }

int flo_add_template(const flo_detector* d, const uint8_t* bgr, const uint16_t* depth, const uint8_t* mask_or_null, int W, int H,
                     const int* num_features_per_modality, float strong_threshold, int extract_threshold,
                     int32_t* headers, int32_t* features, int feature_cap, int* n_features, int32_t bbox[4]) {
  const int L = d->L, M = d->M;
  *n_features = 0;
  int32_t* sel = (int32_t*)malloc(sizeof(int32_t) * 3 * 64 * (size_t)L * M);   /* [L*M][<= 63][3] */
  int* n_sel = (int*)calloc((size_t)L * M, sizeof(int));
  int rc = 0;
  for (int m = 0; m < M && rc == 0; ++m) {
    int w = W, h = H;
    uint8_t* cur_bgr = NULL; uint8_t* cur_q = NULL; uint8_t* cur_mask = NULL;
    if (mask_or_null) { cur_mask = (uint8_t*)malloc((size_t)W * H); memcpy(cur_mask, mask_or_null, (size_t)W * H); }
    size_t num_features = (size_t)num_features_per_modality[m];
    int extract_thr = extract_threshold;
    for (int l = 0; l < L && rc == 0; ++l) {
      if (l > 0) {                                                  /* qp->pyrDown() (:1595-1596; :427-453, :721-739) */
        int nw = w / 2, nh = h / 2;
        num_features /= 2; extract_thr /= 2;
        if (d->kind[m] == 0) { uint8_t* nb = (uint8_t*)malloc((size_t)nw * nh * 3); flo_pyrdown_bgr(cur_bgr ? cur_bgr : bgr, w, h, nb); free(cur_bgr); cur_bgr = nb; }
        else { uint8_t* nq = (uint8_t*)malloc((size_t)nw * nh); flo_resize_nn_half_u8(cur_q, w, h, nq); free(cur_q); cur_q = nq; }
        if (cur_mask) { uint8_t* nm = (uint8_t*)malloc((size_t)nw * nh); flo_resize_nn_half_u8(cur_mask, w, h, nm); free(cur_mask); cur_mask = nm; }
        w = nw; h = nh;
      }
      const size_t n = (size_t)w * h;
      train_cand* cand = (train_cand*)malloc(sizeof(train_cand) * n);
      int nc = 0;
      float distance = 0.f;
      if (d->kind[m] == 0) {                                        /* ColorGradientPyramid::extractTemplate :461-513 */
        uint8_t* q = (uint8_t*)malloc(n); float* mag = (float*)malloc(sizeof(float) * n);
        flo_color_quantize(l == 0 ? bgr : cur_bgr, w, h, d->weak_thr, q, mag);
        uint8_t* er = NULL;
        if (cur_mask) { er = (uint8_t*)malloc(n); flo_erode3_u8(cur_mask, w, h, 1, er); }
        const float thr_sq = strong_threshold * strong_threshold;
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
          size_t i = (size_t)y * w + x;
          if (cur_mask && !((int)cur_mask[i] - (int)er[i] > 0)) continue;      /* subtract(mask, erode(mask)) saturates */
          if (q[i] > 0 && mag[i] > thr_sq) { cand[nc].x = x; cand[nc].y = y; cand[nc].label = label_of(q[i]); cand[nc].score = mag[i]; cand[nc].order = nc; ++nc; }
        }
        free(q); free(mag); free(er);
        if ((size_t)nc < num_features) rc = -1;
        else if (num_features) distance = (float)((size_t)nc / num_features + 1);
      } else {                                                      /* DepthNormalPyramid::extractTemplate :747-825 */
        if (l == 0) { cur_q = (uint8_t*)malloc(n); flo_depth_quantize(depth, w, h, d->dist_thr, d->diff_thr, cur_q); }
        uint8_t* local = NULL;
        if (cur_mask) { local = (uint8_t*)malloc(n); flo_erode3_u8(cur_mask, w, h, 2, local); }
        float* dist8 = (float*)malloc(sizeof(float) * n * 8);
        uint8_t* plane = (uint8_t*)malloc(n);
        for (int lab = 0; lab < 8; ++lab) {
          for (size_t i = 0; i < n; ++i) plane[i] = (uint8_t)(((!local || local[i]) ? (1 << lab) : 0) & cur_q[i]);
          flo_distance_c3(plane, w, h, dist8 + (size_t)lab * n);
        }
        int counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
          size_t i = (size_t)y * w + x;
          if (local && !local[i]) continue;
          int q = cur_q[i];
          if (q == 0 || q == 255) continue;
          int lab = label_of(q);
          float sc = dist8[(size_t)lab * n + i];
          if (sc >= (float)extract_thr) { cand[nc].x = x; cand[nc].y = y; cand[nc].label = lab; cand[nc].score = sc; cand[nc].order = nc; ++nc; ++counts[lab]; }
        }
        if ((size_t)nc < num_features) rc = -1;
        else {
          for (int i = 0; i < nc; ++i) cand[i].score /= (float)counts[cand[i].label];
          size_t area = n;
          if (local) { area = 0; for (size_t i = 0; i < n; ++i) area += local[i] != 0; }
          if (num_features) distance = sqrtf((float)area) / sqrtf((float)num_features) + 1.5f;
        }
        free(local); free(dist8); free(plane);
      }
      if (rc == 0) {
        qsort(cand, (size_t)nc, sizeof(train_cand), cand_cmp);
        select_scattered(cand, nc, sel + (size_t)(l * M + m) * 3 * 64, (int)num_features, distance);
        n_sel[l * M + m] = (int)num_features;
      }
      free(cand);
    }
    free(cur_bgr); free(cur_q); free(cur_mask);
  }
  if (rc == 0) {                                                    /* cropTemplates :52-96 */
    int min_x = INT32_MAX, min_y = INT32_MAX, max_x = INT32_MIN, max_y = INT32_MIN;
    for (int l = 0; l < L; ++l) for (int m = 0; m < M; ++m) for (int k = 0; k < n_sel[l * M + m]; ++k) {
      const int32_t* f = sel + ((size_t)(l * M + m) * 64 + k) * 3;
      int x = f[0] << l, y = f[1] << l;
      if (x < min_x) min_x = x;
      if (y < min_y) min_y = y;
      if (x > max_x) max_x = x;
      if (y > max_y) max_y = y;
    }
    if (min_x % 2 == 1) --min_x;
    if (min_y % 2 == 1) --min_y;
    int nf = 0;
    for (int l = 0; l < L && rc == 0; ++l) for (int m = 0; m < M && rc == 0; ++m) {
      int32_t* hd = headers + (size_t)(l * M + m) * 7;
      hd[0] = (max_x - min_x) >> l; hd[1] = (max_y - min_y) >> l; hd[2] = min_x >> l; hd[3] = min_y >> l; hd[4] = l; hd[5] = nf; hd[6] = n_sel[l * M + m];
      for (int k = 0; k < n_sel[l * M + m]; ++k, ++nf) {
        if (nf >= feature_cap) { rc = -3; break; }
        const int32_t* f = sel + ((size_t)(l * M + m) * 64 + k) * 3;
        features[3 * nf] = f[0] - hd[2]; features[3 * nf + 1] = f[1] - hd[3]; features[3 * nf + 2] = f[2];
      }
    }
    if (rc == 0) { *n_features = nf; bbox[0] = min_x; bbox[1] = min_y; bbox[2] = max_x - min_x; bbox[3] = max_y - min_y; }
  }
  free(sel); free(n_sel);
  return rc;
}
