// Template training, per-pixel half: what Detector::addTemplate (reference linemod/linemod.cpp:1579-1615) computes
// over whole images before its sequential feature selection -
//   ColorGradientPyramid::extractTemplate (:461-513): gradient magnitude of the strongest channel (quantizedOrientations :247-292),
//     mask ring = mask - erode(mask), candidate = ring && quantised angle != 0 && magnitude > strong_threshold^2, score = magnitude;
//   DepthNormalPyramid::extractTemplate (:747-825): erode(mask, 2 iterations), per label the chessboard distance to the nearest pixel
//     that is not (inside the eroded mask and of that label) (cv::distanceTransform DIST_C 3x3), candidate = distance >= extract
//     threshold, score = distance, candidates counted per label.
// The quantised images themselves come from the front end (k_front_end_wave).  Everything here is offline work: the kernels are
// written for clarity and exactness, not for speed (a VGA view takes well under a millisecond of device time anyway).  The stable
// sort and the greedy scattered selection (selectScatteredFeatures :134-163) are sequential and run on the host (api.cu).
#include "fl_internal.cuh"

__device__ __forceinline__ int clampi_d(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cv::GaussianBlur 7x7 (sigma 0 -> {8,28,56,72,56,28,8}/256 per axis, exact integers, one rounding), BORDER_REPLICATE, one channel
__device__ __forceinline__ int blur7_at(const uint8_t* __restrict__ bgr, int W, int H, int x, int y, int ch) {
  const int k[7] = {8, 28, 56, 72, 56, 28, 8};
  int s = 0;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const uint8_t* row = bgr + (size_t)clampi_d(y + j - 3, 0, H - 1) * W * 3;
    int r = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) r += k[i] * row[clampi_d(x + i - 3, 0, W - 1) * 3 + ch];
    s += k[j] * r;
  }
  return (s + 32768) >> 16;
}

// magnitude image of quantizedOrientations: Sobel 3x3 (BORDER_REPLICATE) on the blurred image per channel, dx^2 + dy^2 in int,
// the strongest channel with the reference's tie rule (:275-292), stored as float
__global__ void k_train_magnitude(const uint8_t* __restrict__ bgr, int W, int H, float* __restrict__ mag) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  int m[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    int b[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) b[j][i] = blur7_at(bgr, W, H, clampi_d(x + i - 1, 0, W - 1), clampi_d(y + j - 1, 0, H - 1), ch);
    const int dx = (b[0][2] + 2 * b[1][2] + b[2][2]) - (b[0][0] + 2 * b[1][0] + b[2][0]);
    const int dy = (b[2][0] + 2 * b[2][1] + b[2][2]) - (b[0][0] + 2 * b[0][1] + b[0][2]);
    m[ch] = dx * dx + dy * dy;
  }
  int best;
  if (m[0] >= m[1] && m[0] >= m[2]) best = m[0];
  else if (m[1] >= m[0] && m[1] >= m[2]) best = m[1];
  else best = m[2];
  mag[(size_t)y * W + x] = (float)best;
}

// cv::erode with the default 3x3 rectangle, BORDER_REPLICATE, one iteration
__global__ void k_train_erode3(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  int m = 255;
#pragma unroll
  for (int j = -1; j <= 1; ++j)
#pragma unroll
    for (int i = -1; i <= 1; ++i) m = min(m, (int)src[(size_t)clampi_d(y + j, 0, H - 1) * W + clampi_d(x + i, 0, W - 1)]);
  dst[(size_t)y * W + x] = (uint8_t)m;
}

// colour candidates (:473-497): score = magnitude where the pixel is a candidate, -1 elsewhere.  mask / eroded may be NULL (no mask).
__global__ void k_train_color_score(const uint8_t* __restrict__ angle, const float* __restrict__ mag, const uint8_t* __restrict__ mask,
                                    const uint8_t* __restrict__ eroded, int W, int H, float threshold_sq, float* __restrict__ score) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t i = (size_t)y * W + x;
  bool in = true;
  if (mask) in = max((int)mask[i] - (int)eroded[i], 0) != 0;                 // cv::subtract(mask, erode(mask)) saturates
  const float s = mag[i];
  score[i] = (in && angle[i] > 0 && s > threshold_sq) ? s : -1.f;
}

// per (label, row): distance along the row to the nearest pixel that is NOT (inside the local mask and of this label); 0xFFFF = none
// in this row.  plane(l, x, y) != 0  <=>  (local_mask == NULL || local_mask != 0) && (normal & (1 << l))   (:759-763)
__global__ void k_train_rowdist(const uint8_t* __restrict__ normal, const uint8_t* __restrict__ local_mask, int W, int H,
                                uint16_t* __restrict__ hd, int* __restrict__ any_zero) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 8 * H) return;
  const int l = idx / H, y = idx - l * H;
  const uint8_t* nr = normal + (size_t)y * W;
  const uint8_t* mr = local_mask ? local_mask + (size_t)y * W : nullptr;
  uint16_t* out = hd + ((size_t)l * H + y) * W;
  int last = -1;
  bool zero_seen = false;
  for (int x = 0; x < W; ++x) {
    const bool nz = (!mr || mr[x]) && (nr[x] & (1 << l));
    if (!nz) { last = x; zero_seen = true; }
    out[x] = last < 0 ? 0xFFFF : (uint16_t)min(x - last, 0xFFFE);
  }
  last = -1;
  for (int x = W - 1; x >= 0; --x) {
    const bool nz = (!mr || mr[x]) && (nr[x] & (1 << l));
    if (!nz) last = x;
    if (last >= 0) out[x] = (uint16_t)min((int)out[x], min(last - x, 0xFFFE));
  }
  if (zero_seen) atomicOr(any_zero + l, 1);
}

// depth candidates (:772-800): chessboard distance of the pixel's own label plane = min over rows y' of max(|y - y'|, rowdist(y', x));
// a plane without a single zero pixel has distance 65535 everywhere (OpenCV 4.13's own code; its IPP path says FLT_MAX).  score = distance where candidate, -1 elsewhere;
// label_counts += 1 per candidate.
__global__ void k_train_depth_score(const uint8_t* __restrict__ normal, const uint8_t* __restrict__ local_mask, const uint16_t* __restrict__ hd,
                                    const int* __restrict__ any_zero, int W, int H, int extract_threshold, float* __restrict__ score,
                                    int* __restrict__ label_counts) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t i = (size_t)y * W + x;
  score[i] = -1.f;
  if (local_mask && !local_mask[i]) return;
  const int q = normal[i];
  if (q == 0 || q == 255) return;                                            // background and shadow
  const int l = __ffs(q) - 1;                                                // getLabel: the (single) set bit
  float dist;
  if (!any_zero[l]) dist = 65535.f;
  else {
    const uint16_t* h = hd + (size_t)l * H * W + x;
    int best = 0x7FFFFFFF;
    for (int dy = 0; dy < best; ++dy) {
      if (y - dy >= 0) { const int v = h[(size_t)(y - dy) * W]; if (v != 0xFFFF) best = min(best, max(dy, v)); }
      if (y + dy < H) { const int v = h[(size_t)(y + dy) * W]; if (v != 0xFFFF) best = min(best, max(dy, v)); }
      if (y - dy < 0 && y + dy >= H) break;
    }
    dist = (float)best;
  }
  if (dist >= (float)extract_threshold) { score[i] = dist; atomicAdd(label_counts + l, 1); }
}

void fl_launch_train_magnitude(const uint8_t* bgr, int W, int H, float* mag, cudaStream_t s) {
  k_train_magnitude<<<dim3((W + 127) / 128, H), 128, 0, s>>>(bgr, W, H, mag);
}
void fl_launch_train_erode3(const uint8_t* src, int W, int H, uint8_t* dst, cudaStream_t s) {
  k_train_erode3<<<dim3((W + 127) / 128, H), 128, 0, s>>>(src, W, H, dst);
}
void fl_launch_train_color_score(const uint8_t* angle, const float* mag, const uint8_t* mask, const uint8_t* eroded, int W, int H, float threshold_sq,
                                 float* score, cudaStream_t s) {
  k_train_color_score<<<dim3((W + 127) / 128, H), 128, 0, s>>>(angle, mag, mask, eroded, W, H, threshold_sq, score);
}
void fl_launch_train_depth_score(const uint8_t* normal, const uint8_t* local_mask, uint16_t* hd, int* any_zero8, int W, int H, int extract_threshold,
                                 float* score, int* label_counts8, cudaStream_t s) {
  cudaMemsetAsync(any_zero8, 0, 8 * sizeof(int), s);
  cudaMemsetAsync(label_counts8, 0, 8 * sizeof(int), s);
  k_train_rowdist<<<(8 * H + 127) / 128, 128, 0, s>>>(normal, local_mask, W, H, hd, any_zero8);
  k_train_depth_score<<<dim3((W + 127) / 128, H), 128, 0, s>>>(normal, local_mask, hd, any_zero8, W, H, extract_threshold, score, label_counts8);
}
