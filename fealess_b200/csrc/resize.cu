// resize.cu - cv::resize(..., INTER_LINEAR) for the two input formats of CObjRecoLmICP::PrepareInputData (reference
// CadReco/obj_reco_lmicp.cpp:38-45, 216-259: TImage2Mat(tRGB, 640, h, CV_8UC3, INTER_LINEAR) and the same for the 16UC1 depth).
// The arithmetic lives in OpenCV (un-vendored, imgproc/src/resize.cpp); it is restated here and in oracle/resize_oracle.py and
// pinned against cv2 4.13 (IPP off = OpenCV's own code) by tests/test_resize_oracle.py:
//   * index / weight tables per destination column and row are computed on the HOST exactly like OpenCV does
//     (double scale, float fractional part; columns clamp index AND weight at the borders, rows clamp only the index);
//   * 8U: 11-bit fixed-point weights, horizontal pass in int32, vertical pass ((b*(r>>4))>>16 summed, +2, >>2);
//   * 16U: float weights, plain fp32 multiply / add (no contraction), round-half-even, saturate;
//   * an exact 2x decimation takes OpenCV's INTER_AREA fast path for both types: (a+b+c+d+2)>>2.
#include "fl_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) k_resize_linear_8uc3(const uint8_t* __restrict__ src, int sW, int sH, uint8_t* __restrict__ dst, int dW, int dH,
                                                            const int* __restrict__ xofs, const short2* __restrict__ ialpha,
                                                            const int* __restrict__ yofs, const short2* __restrict__ ibeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;              // one thread per destination byte: consecutive lanes, consecutive bytes
  if (i >= dW * dH * 3) return;
  const int c = i % 3, dx = (i / 3) % dW, dy = i / (3 * dW);
  const int sx0 = xofs[dx], sx1 = min(sx0 + 1, sW - 1);
  const int y = yofs[dy], sy0 = min(max(y, 0), sH - 1), sy1 = min(max(y + 1, 0), sH - 1);
  const short2 a = ialpha[dx], b = ibeta[dy];
  const uint8_t* r0 = src + (size_t)sy0 * sW * 3;
  const uint8_t* r1 = src + (size_t)sy1 * sW * 3;
  const int h0 = (int)r0[sx0 * 3 + c] * a.x + (int)r0[sx1 * 3 + c] * a.y;      // HResizeLinear<uchar, int, short>
  const int h1 = (int)r1[sx0 * 3 + c] * a.x + (int)r1[sx1 * 3 + c] * a.y;
  dst[i] = (uint8_t)((((b.x * (h0 >> 4)) >> 16) + ((b.y * (h1 >> 4)) >> 16) + 2) >> 2);   // VResizeLinear<uchar, int, short, FixedPtCast>
}

__global__ void __launch_bounds__(256) k_resize_linear_16uc1(const uint16_t* __restrict__ src, int sW, int sH, uint16_t* __restrict__ dst, int dW, int dH,
                                                             const int* __restrict__ xofs, const float2* __restrict__ alpha,
                                                             const int* __restrict__ yofs, const float2* __restrict__ beta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dW * dH) return;
  const int dx = i % dW, dy = i / dW;
  const int sx0 = xofs[dx], sx1 = min(sx0 + 1, sW - 1);
  const int y = yofs[dy], sy0 = min(max(y, 0), sH - 1), sy1 = min(max(y + 1, 0), sH - 1);
  const float2 a = alpha[dx], b = beta[dy];
  const uint16_t* r0 = src + (size_t)sy0 * sW;
  const uint16_t* r1 = src + (size_t)sy1 * sW;
  const float h0 = __fadd_rn(__fmul_rn((float)r0[sx0], a.x), __fmul_rn((float)r0[sx1], a.y));   // HResizeLinear<ushort, float, float>
  const float h1 = __fadd_rn(__fmul_rn((float)r1[sx0], a.x), __fmul_rn((float)r1[sx1], a.y));
  const float v = __fadd_rn(__fmul_rn(h0, b.x), __fmul_rn(h1, b.y));                               // VResizeLinear<ushort, float, float, Cast>
  dst[i] = (uint16_t)min(max(__float2int_rn(v), 0), 65535);                                        // saturate_cast<ushort>(cvRound(v))
}

// exact 2x decimation: ResizeAreaFast, (a + b + c + d + 2) >> 2 per channel
template <typename T, int CN>
__global__ void __launch_bounds__(256) k_resize_half(const T* __restrict__ src, int sW, T* __restrict__ dst, int dW, int dH) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dW * dH * CN) return;
  const int c = i % CN, dx = (i / CN) % dW, dy = i / (CN * dW);
  const T* r0 = src + ((size_t)(2 * dy) * sW + 2 * dx) * CN + c;
  const T* r1 = r0 + (size_t)sW * CN;
  dst[i] = (T)(((int)r0[0] + (int)r0[CN] + (int)r1[0] + (int)r1[CN] + 2) >> 2);
}

}  // namespace

void fl_launch_resize_linear(const void* src, int sW, int sH, int type, void* dst, int dW, int dH, const fl_resize_tables& t, cudaStream_t s) {
  const int cn = type == FL_IMG_8UC3 ? 3 : 1;
  const int n = dW * dH * cn, grid = (n + 255) / 256;
  if (n <= 0) return;
  if (sW == 2 * dW && sH == 2 * dH) {
    if (type == FL_IMG_8UC3) k_resize_half<uint8_t, 3><<<grid, 256, 0, s>>>((const uint8_t*)src, sW, (uint8_t*)dst, dW, dH);
    else k_resize_half<uint16_t, 1><<<grid, 256, 0, s>>>((const uint16_t*)src, sW, (uint16_t*)dst, dW, dH);
  } else if (type == FL_IMG_8UC3) {
    k_resize_linear_8uc3<<<grid, 256, 0, s>>>((const uint8_t*)src, sW, sH, (uint8_t*)dst, dW, dH, t.xofs, t.ialpha, t.yofs, t.ibeta);
  } else {
    k_resize_linear_16uc1<<<grid, 256, 0, s>>>((const uint16_t*)src, sW, sH, (uint16_t*)dst, dW, dH, t.xofs, t.alpha, t.yofs, t.beta);
  }
}
