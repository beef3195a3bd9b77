// TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Bodies of oracle/ref_shim/cvshim.hpp (see its header comment).
//
// The image / maths primitives below restate OpenCV's published algorithms for exactly the type combinations the
// reference's hot path calls them with (call sites: linemod/linemod.cpp:247-249, 303, 314, 443, 448, 684, 731;
// ICP/ICP.cpp:228, 658-659, 742; ICP/depth_to_3d.cpp:257-259).  OpenCV is an un-vendored dependency of the reference
// (CMakeLists.txt:13-16, version not pinned).  Each primitive is compared with the real cv2 4.13 in
// tests/test_oracle_ref.py, and each can be swapped for a callback (cvshim::set_hook) at run time.
#include "cvshim.hpp"

#include <chrono>

namespace cvshim {
static hook_fn g_hooks[OP_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
void set_hook(int op, hook_fn f) { if (op >= 0 && op < OP_COUNT) g_hooks[op] = f; }
hook_fn get_hook(int op) { return (op >= 0 && op < OP_COUNT) ? g_hooks[op] : nullptr; }
}  // namespace cvshim

namespace cv {

int64 getTickCount() {
  return (int64)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
double getTickFrequency() { return 1e9; }
String format(const char* fmt, ...) {
  char buf[4096];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  return String(buf);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
  return p;
}
static Mat continuous(const Mat& m) { return m.isContinuous() ? m : m.clone(); }

// ---------------------------------------------------------------- Mat members
Mat Mat::eye(int r, int c, int type) {
  Mat m = zeros(r, c, type);
  for (int i = 0; i < std::min(r, c); ++i) {
    switch (m.depth()) {
      case CV_32F: m.at<float>(i, i) = 1.f; break;
      case CV_64F: m.at<double>(i, i) = 1.; break;
      default: CV_Error(Error::StsNotImplemented, "cvshim: Mat::eye depth");
    }
  }
  return m;
}
void Mat::copyTo(Mat& dst) const {
  if (empty()) { dst.release(); return; }
  if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.type() == type() && dst.step.v == step.v) return;
  dst.create(rows, cols, type());
  for (int r = 0; r < rows; ++r) std::memcpy(dst.ptr(r), ptr(r), (size_t)cols * elemSize());
}
void Mat::copyTo(OutputArray dst) const {
  if (empty()) { if (!dst.fixedSize()) dst.getMatRef().release(); return; }
  dst.create(rows, cols, type());
  Mat d = dst.getMat();
  if (d.data == data) return;
  for (int r = 0; r < rows; ++r) std::memcpy(d.ptr(r), ptr(r), (size_t)cols * elemSize());
}
void Mat::copyTo(OutputArray dst_, InputArray mask_) const {
  Mat mask = mask_.getMat();
  if (mask.empty()) { copyTo(dst_); return; }   // Mat::copyTo(dst, mask): an empty mask means "copy everything"
  CV_Assert(mask.depth() == CV_8U && mask.channels() == 1 && mask.rows == rows && mask.cols == cols);
  Mat d0 = dst_.getMat();
  bool fresh = d0.data == nullptr || d0.rows != rows || d0.cols != cols || d0.type() != type();
  dst_.create(rows, cols, type());
  Mat d = dst_.getMat();
  if (fresh) d.setTo(Scalar());   // a (re)allocated destination is zero-filled first
  size_t esz = elemSize();
  for (int r = 0; r < rows; ++r) {
    const uchar* m = mask.ptr(r); const uchar* s = ptr(r); uchar* o = d.ptr(r);
    for (int c = 0; c < cols; ++c) if (m[c]) std::memcpy(o + c * esz, s + c * esz, esz);
  }
}

template <typename S> static inline double load_as_double(const uchar* p) { return (double)*(const S*)p; }
template <typename D> static void store_rows(const Mat& src, Mat& dst, double alpha, double beta, bool noscale, bool use_float) {
  int n = src.cols * src.channels();
  for (int r = 0; r < src.rows; ++r) {
    const uchar* s = src.ptr(r); D* d = dst.ptr<D>(r);
    for (int i = 0; i < n; ++i) {
      switch (src.depth()) {
#define CASE(DEPTH, S)                                                                                          \
        case DEPTH: {                                                                                              \
          S v = ((const S*)s)[i];                                                                                  \
          if (noscale) d[i] = saturate_cast<D>(v);                                                                 \
          else if (use_float) d[i] = saturate_cast<D>((float)v * (float)alpha + (float)beta);   /* cvt_32f */     \
          else d[i] = saturate_cast<D>((double)v * alpha + beta);                               /* cvt_64f */     \
          break; }
        CASE(CV_8U, uchar) CASE(CV_8S, schar) CASE(CV_16U, ushort) CASE(CV_16S, short) CASE(CV_32S, int) CASE(CV_32F, float) CASE(CV_64F, double)
#undef CASE
        default: CV_Error(Error::StsNotImplemented, "cvshim: convertTo depth");
      }
    }
  }
}
void Mat::convert_into(Mat& dst, int rtype, double alpha, double beta) const {
  // Mat::convertTo (core/src/convert.cpp, convert_scale.cpp): without scaling dst = saturate_cast(src); with scaling the
  // work type is float unless a 64F (or 32S<->32S/64F) operand is involved, i.e. dst = saturate_cast(src*(float)a + (float)b)
  if (empty()) { dst.release(); return; }
  int ddepth = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
  bool noscale = alpha == 1.0 && beta == 0.0;
  bool use_float = !(depth() == CV_64F || ddepth == CV_64F || (depth() == CV_32S && ddepth == CV_32S));
  Mat src = *this;
  Mat out;
  Mat& d = (dst.data == data) ? out : dst;
  d.create(rows, cols, CV_MAKETYPE(ddepth, channels()));
  switch (ddepth) {
    case CV_8U: store_rows<uchar>(src, d, alpha, beta, noscale, use_float); break;
    case CV_16U: store_rows<ushort>(src, d, alpha, beta, noscale, use_float); break;
    case CV_16S: store_rows<short>(src, d, alpha, beta, noscale, use_float); break;
    case CV_32S: store_rows<int>(src, d, alpha, beta, noscale, use_float); break;
    case CV_32F: store_rows<float>(src, d, alpha, beta, noscale, use_float); break;
    case CV_64F: store_rows<double>(src, d, alpha, beta, noscale, use_float); break;
    default: CV_Error(Error::StsNotImplemented, "cvshim: convertTo destination depth");
  }
  if (&d != &dst) dst = d;
}
void Mat::convertTo(OutputArray dst_, int rtype, double alpha, double beta) const {
  if (empty()) { if (!dst_.fixedSize()) dst_.getMatRef().release(); return; }
  int ddepth = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
  Mat tmp;
  convert_into(tmp, ddepth, alpha, beta);
  dst_.create(rows, cols, tmp.type());
  Mat d = dst_.getMat();
  if (d.data != tmp.data) for (int r = 0; r < rows; ++r) std::memcpy(d.ptr(r), tmp.ptr(r), (size_t)cols * tmp.elemSize());
}
static void put_scalar(uchar* p, int depth, int cn, const Scalar& s) {
  for (int k = 0; k < cn; ++k) {
    double v = s.val[k < 4 ? k : 3];
    switch (depth) {
      case CV_8U: ((uchar*)p)[k] = saturate_cast<uchar>(v); break;
      case CV_16U: ((ushort*)p)[k] = saturate_cast<ushort>(v); break;
      case CV_16S: ((short*)p)[k] = saturate_cast<short>(v); break;
      case CV_32S: ((int*)p)[k] = saturate_cast<int>(v); break;
      case CV_32F: ((float*)p)[k] = (float)v; break;
      case CV_64F: ((double*)p)[k] = v; break;
      default: CV_Error(Error::StsNotImplemented, "cvshim: setTo depth");
    }
  }
}
Mat& Mat::setTo(const Scalar& s) {
  size_t esz = elemSize();
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) put_scalar(ptr(r) + c * esz, depth(), channels(), s);
  return *this;
}
Mat& Mat::setTo(const Scalar& s, InputArray mask_) {
  Mat mask = mask_.getMat();
  if (mask.empty()) return setTo(s);
  CV_Assert(mask.depth() == CV_8U && mask.rows == rows && mask.cols == cols);
  size_t esz = elemSize();
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) if (mask.at<uchar>(r, c)) put_scalar(ptr(r) + c * esz, depth(), channels(), s);
  return *this;
}
Mat Mat::t() const {
  Mat d(cols, rows, type());
  size_t esz = elemSize();
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) std::memcpy(d.ptr(c) + r * esz, ptr(r) + c * esz, esz);
  return d;
}
Mat Mat::reshape(int cn, int new_rows) const {
  Mat src = continuous(*this);
  if (cn == 0) cn = channels();
  size_t total_ch = (size_t)rows * cols * channels();
  int nr = new_rows ? new_rows : rows;
  CV_Assert(nr > 0 && total_ch % ((size_t)nr * cn) == 0);
  Mat d = src;
  d.flags = CV_MAKETYPE(depth(), cn);
  d.rows = nr; d.cols = (int)(total_ch / ((size_t)nr * cn));
  d.step = (size_t)d.cols * d.elemSize();
  return d;
}
void Mat::resize(size_t nrows) {
  if ((size_t)rows == nrows) return;
  if (nrows < (size_t)rows) { rows = (int)nrows; return; }
  Mat d((int)nrows, cols, type());
  for (int r = 0; r < rows; ++r) std::memcpy(d.ptr(r), ptr(r), (size_t)cols * elemSize());
  *this = d;
}

template <typename F> static Mat elementwise(const Mat& a, const Mat* b, F f) {
  CV_Assert(a.depth() == CV_32F || a.depth() == CV_64F);
  if (b) CV_Assert(b->rows == a.rows && b->cols == a.cols && b->type() == a.type());
  Mat d(a.rows, a.cols, a.type());
  int n = a.cols * a.channels();
  for (int r = 0; r < a.rows; ++r)
    for (int i = 0; i < n; ++i) {
      if (a.depth() == CV_32F) d.ptr<float>(r)[i] = (float)f((double)a.ptr<float>(r)[i], b ? (double)b->ptr<float>(r)[i] : 0.0);
      else d.ptr<double>(r)[i] = f(a.ptr<double>(r)[i], b ? b->ptr<double>(r)[i] : 0.0);
    }
  return d;
}
Mat operator+(const Mat& a, const Mat& b) { return elementwise(a, &b, [](double x, double y) { return x + y; }); }
Mat operator+(const Mat& a, double s) { return elementwise(a, nullptr, [s](double x, double) { return x + s; }); }
Mat operator-(const Mat& a, const Mat& b) { return elementwise(a, &b, [](double x, double y) { return x - y; }); }
Mat operator-(const Mat& a, double s) { return elementwise(a, nullptr, [s](double x, double) { return x - s; }); }
Mat operator-(const Mat& a) { return elementwise(a, nullptr, [](double x, double) { return -x; }); }
Mat operator*(const Mat& a, double s) { return elementwise(a, nullptr, [s](double x, double) { return x * s; }); }
Mat operator*(double s, const Mat& a) { return a * s; }
Mat operator/(const Mat& a, double s) { return elementwise(a, nullptr, [s](double x, double) { return x / s; }); }
Mat Mat::mul(const Mat& o, double scale) const { return elementwise(*this, &o, [scale](double x, double y) { return x * y * scale; }); }
Mat operator*(const Mat& a, const Mat& b) {
  // MatExpr product -> cv::gemm; its generic kernel (GEMMSingleMul<float,double>) accumulates float products in double
  CV_Assert(a.cols == b.rows && a.type() == b.type() && a.channels() == 1 && (a.depth() == CV_32F || a.depth() == CV_64F));
  Mat d(a.rows, b.cols, a.type());
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.cols; ++j) {
      double s = 0;
      for (int k = 0; k < a.cols; ++k)
        s += a.depth() == CV_32F ? (double)a.at<float>(i, k) * (double)b.at<float>(k, j) : a.at<double>(i, k) * b.at<double>(k, j);
      if (a.depth() == CV_32F) d.at<float>(i, j) = (float)s; else d.at<double>(i, j) = s;
    }
  return d;
}
Mat operator==(const Mat& a, double s) {
  CV_Assert(a.channels() == 1);
  Mat d(a.rows, a.cols, CV_8U);
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) {
      double v;
      switch (a.depth()) {
        case CV_8U: v = a.at<uchar>(r, c); break;
        case CV_16U: v = a.at<ushort>(r, c); break;
        case CV_16S: v = a.at<short>(r, c); break;
        case CV_32S: v = a.at<int>(r, c); break;
        case CV_32F: v = a.at<float>(r, c); break;
        default: v = a.at<double>(r, c); break;
      }
      d.at<uchar>(r, c) = v == s ? 255 : 0;
    }
  return d;
}
Mat operator|(const Mat& a, const Mat& b) {
  CV_Assert(a.type() == b.type() && a.rows == b.rows && a.cols == b.cols && a.depth() == CV_8U);
  Mat d(a.rows, a.cols, a.type());
  for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols * a.channels(); ++c) d.ptr(r)[c] = a.ptr(r)[c] | b.ptr(r)[c];
  return d;
}

// ---------------------------------------------------------------- core functions
double norm(InputArray a_, InputArray b_, int normType) {
  // cv::norm(a, b, NORM_L2) on CV_32F: squared differences accumulate in double (normDiffL2_32f), then sqrt
  CV_Assert(normType == NORM_L2);
  Mat a = a_.getMat(), b = b_.getMat();
  CV_Assert(a.type() == b.type() && a.rows == b.rows && a.cols == b.cols && (a.depth() == CV_32F || a.depth() == CV_64F));
  double s = 0;
  int n = a.cols * a.channels();
  for (int r = 0; r < a.rows; ++r)
    for (int i = 0; i < n; ++i) {
      double d = a.depth() == CV_32F ? (double)(a.ptr<float>(r)[i] - b.ptr<float>(r)[i]) : a.ptr<double>(r)[i] - b.ptr<double>(r)[i];
      s += d * d;
    }
  return std::sqrt(s);
}
template <typename OP> static void arith(InputArray a_, InputArray b_, OutputArray dst_, int dtype, OP op) {
  Mat a = a_.getMat(), b = b_.getMat();
  CV_Assert(a.rows == b.rows && a.cols == b.cols && a.channels() == b.channels());
  int ddepth = dtype >= 0 ? CV_MAT_DEPTH(dtype) : a.depth();
  Mat out(a.rows, a.cols, CV_MAKETYPE(ddepth, a.channels()));
  int n = a.cols * a.channels();
  auto load = [](const Mat& m, int r, int i) -> double {
    switch (m.depth()) {
      case CV_8U: return m.ptr<uchar>(r)[i];
      case CV_16U: return m.ptr<ushort>(r)[i];
      case CV_16S: return m.ptr<short>(r)[i];
      case CV_32S: return m.ptr<int>(r)[i];
      case CV_32F: return m.ptr<float>(r)[i];
      default: return m.ptr<double>(r)[i];
    }
  };
  for (int r = 0; r < a.rows; ++r)
    for (int i = 0; i < n; ++i) {
      double x = load(a, r, i), y = load(b, r, i);
      switch (ddepth) {
        case CV_8U: out.ptr<uchar>(r)[i] = saturate_cast<uchar>((int)op(x, y)); break;
        case CV_16U: out.ptr<ushort>(r)[i] = saturate_cast<ushort>((int)op(x, y)); break;
        case CV_16S: out.ptr<short>(r)[i] = saturate_cast<short>((int)op(x, y)); break;
        case CV_32S: out.ptr<int>(r)[i] = (int)op(x, y); break;
        case CV_32F: out.ptr<float>(r)[i] = op((float)x, (float)y); break;    // float operands: one rounding, as the fp32 add
        default: out.ptr<double>(r)[i] = op(x, y); break;
      }
    }
  dst_.create(out.rows, out.cols, out.type());
  Mat d = dst_.getMat();
  for (int r = 0; r < out.rows; ++r) std::memcpy(d.ptr(r), out.ptr(r), (size_t)out.cols * out.elemSize());
}
struct AddOp { double operator()(double x, double y) const { return x + y; } float operator()(float x, float y) const { return x + y; } };
struct SubOp { double operator()(double x, double y) const { return x - y; } float operator()(float x, float y) const { return x - y; } };
void add(InputArray a, InputArray b, OutputArray dst, InputArray mask, int dtype) { CV_Assert(mask.empty()); arith(a, b, dst, dtype, AddOp()); }
void subtract(InputArray a, InputArray b, OutputArray dst, InputArray mask, int dtype) { CV_Assert(mask.empty()); arith(a, b, dst, dtype, SubOp()); }
void bitwise_and(InputArray a_, InputArray b_, OutputArray dst_, InputArray mask) {
  CV_Assert(mask.empty());
  Mat a = a_.getMat(), b = b_.getMat();
  CV_Assert(a.type() == b.type() && a.rows == b.rows && a.cols == b.cols && a.depth() == CV_8U);
  Mat out(a.rows, a.cols, a.type());
  for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols * a.channels(); ++c) out.ptr(r)[c] = a.ptr(r)[c] & b.ptr(r)[c];
  dst_.create(out.rows, out.cols, out.type());
  Mat d = dst_.getMat();
  for (int r = 0; r < out.rows; ++r) std::memcpy(d.ptr(r), out.ptr(r), (size_t)out.cols * out.elemSize());
}
int countNonZero(InputArray a_) {
  Mat a = a_.getMat();
  CV_Assert(a.channels() == 1 && a.depth() == CV_8U);
  int n = 0;
  for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) n += a.at<uchar>(r, c) != 0;
  return n;
}
bool checkRange(InputArray a_, bool, Point*, double minVal, double maxVal) {
  // cv::checkRange: every element finite and in [minVal, maxVal)
  Mat a = a_.getMat();
  int n = a.cols * a.channels();
  for (int r = 0; r < a.rows; ++r)
    for (int i = 0; i < n; ++i) {
      double v = a.depth() == CV_32F ? (double)a.ptr<float>(r)[i] : a.depth() == CV_64F ? a.ptr<double>(r)[i] : 0.0;
      if (std::isnan(v) || std::isinf(v) || v < minVal || v >= maxVal) return false;
    }
  return true;
}
void merge(const std::vector<Mat>& mv, OutputArray dst_) {
  CV_Assert(!mv.empty());
  int cn = (int)mv.size();
  dst_.create(mv[0].rows, mv[0].cols, CV_MAKETYPE(mv[0].depth(), cn));
  Mat d = dst_.getMat();
  size_t e1 = mv[0].elemSize1();
  for (int k = 0; k < cn; ++k) {
    CV_Assert(mv[k].channels() == 1 && mv[k].rows == d.rows && mv[k].cols == d.cols && mv[k].depth() == d.depth());
    for (int r = 0; r < d.rows; ++r) for (int c = 0; c < d.cols; ++c) std::memcpy(d.ptr(r) + ((size_t)c * cn + k) * e1, mv[k].ptr(r) + c * e1, e1);
  }
}
void split(const Mat& src, std::vector<Mat>& mv) {
  int cn = src.channels();
  mv.resize(cn);
  size_t e1 = src.elemSize1();
  for (int k = 0; k < cn; ++k) {
    mv[k].create(src.rows, src.cols, src.depth());
    for (int r = 0; r < src.rows; ++r) for (int c = 0; c < src.cols; ++c) std::memcpy(mv[k].ptr(r) + c * e1, src.ptr(r) + ((size_t)c * cn + k) * e1, e1);
  }
}

// cv::phase(x, y, angle, angleInDegrees = true) on CV_32F: hal::fastAtan32f, a degree-7 odd polynomial on min/max with octant
// fix-ups, evaluated in fp32 (core/src/mathfuncs_core.simd.hpp).  The BINS it yields are compared with cv2.phase over the
// whole reachable Sobel domain in tests/test_oracle_cv2.py.
static inline float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) { c = ay / (ax + (float)DBL_EPSILON); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
  else { c = ax / (ay + (float)DBL_EPSILON); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}
void phase(InputArray x_, InputArray y_, OutputArray angle_, bool angleInDegrees) {
  Mat x = continuous(x_.getMat()), y = continuous(y_.getMat());
  CV_Assert(x.type() == CV_32FC1 && y.type() == CV_32FC1 && x.rows == y.rows && x.cols == y.cols && angleInDegrees);
  angle_.create(x.rows, x.cols, CV_32F);
  Mat a = angle_.getMat();
  if (cvshim::hook_fn h = cvshim::get_hook(cvshim::OP_PHASE_DEG)) {
    int dims[2] = {x.rows, x.cols};
    Mat out(x.rows, x.cols, CV_32F);
    h(cvshim::OP_PHASE_DEG, x.data, y.data, out.data, dims);
    out.copyTo(a);
    return;
  }
  for (int r = 0; r < x.rows; ++r) for (int c = 0; c < x.cols; ++c) a.at<float>(r, c) = fast_atan2_deg(y.at<float>(r, c), x.at<float>(r, c));
}

void Rodrigues(InputArray src_, OutputArray dst_) {
  // rotation vector -> matrix (calib3d); only the 3-vector direction is reachable (pose_result.h:115) and the hot path
  // never takes it (NMS passes 3x3 matrices)
  Mat s; src_.getMat().convert_into(s, CV_64F);
  CV_Assert(s.total() == 3);
  double rx = s.at<double>(0), ry = s.at<double>(1), rz = s.at<double>(2);
  double th = std::sqrt(rx * rx + ry * ry + rz * rz);
  Mat R = Mat::eye(3, 3, CV_64F);
  if (th > DBL_EPSILON) {
    double c = std::cos(th), sn = std::sin(th), c1 = 1 - c, itheta = 1 / th;
    rx *= itheta; ry *= itheta; rz *= itheta;
    double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    double rxm[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; ++k) R.at<double>(k / 3, k % 3) = c * (k % 4 == 0 ? 1.0 : 0.0) + c1 * rrt[k] + sn * rxm[k];
  }
  Mat out; R.convert_into(out, src_.getMat().depth() == CV_32F ? CV_32F : CV_64F);
  dst_.create(3, 3, out.type());
  out.copyTo(dst_.getMatRef());
}

// cv::SVD::compute for CV_32F without LAPACK: JacobiSVDImpl_<float> (core/src/lapack.cpp) - one-sided Jacobi on A^T, double
// dot products, float rotations, eps = 2*FLT_EPSILON, minval = FLT_MIN, at most max(m, 30) sweeps; singular values sorted
// in decreasing order; vanishing singular values get a pseudo-random left vector (RNG 0x12345678).
static void jacobi_svd32f(float* At, size_t astep, float* W_, float* Vt, size_t vstep, int m, int n, int n1) {
  const double minval = FLT_MIN; const float eps = FLT_EPSILON * 2;
  std::vector<double> W(n);
  int i, j, k, iter, max_iter = std::max(m, 30);
  float c, s; double sd;
  for (i = 0; i < n; i++) {
    for (k = 0, sd = 0; k < m; k++) { float t = At[i * astep + k]; sd += (double)t * t; }
    W[i] = sd;
    if (Vt) { for (k = 0; k < n; k++) Vt[i * vstep + k] = 0; Vt[i * vstep + i] = 1; }
  }
  for (iter = 0; iter < max_iter; iter++) {
    bool changed = false;
    for (i = 0; i < n - 1; i++)
      for (j = i + 1; j < n; j++) {
        float *Ai = At + i * astep, *Aj = At + j * astep;
        double a = W[i], p = 0, b = W[j];
        for (k = 0; k < m; k++) p += (double)Ai[k] * Aj[k];
        if (std::abs(p) <= eps * std::sqrt((double)a * b)) continue;
        p *= 2;
        double beta = a - b, gamma = hypot((double)p, beta);
        if (beta < 0) { double delta = (gamma - beta) * 0.5; s = (float)std::sqrt(delta / gamma); c = (float)(p / (gamma * s * 2)); }
        else { c = (float)std::sqrt((gamma + beta) / (gamma * 2)); s = (float)(p / (gamma * c * 2)); }
        a = b = 0;
        for (k = 0; k < m; k++) {
          float t0 = c * Ai[k] + s * Aj[k];
          float t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0; Aj[k] = t1;
          a += (double)t0 * t0; b += (double)t1 * t1;
        }
        W[i] = a; W[j] = b;
        changed = true;
        if (Vt) {
          float *Vi = Vt + i * vstep, *Vj = Vt + j * vstep;
          for (k = 0; k < n; k++) {
            float t0 = c * Vi[k] + s * Vj[k];
            float t1 = -s * Vi[k] + c * Vj[k];
            Vi[k] = t0; Vj[k] = t1;
          }
        }
      }
    if (!changed) break;
  }
  for (i = 0; i < n; i++) {
    for (k = 0, sd = 0; k < m; k++) { float t = At[i * astep + k]; sd += (double)t * t; }
    W[i] = std::sqrt(sd);
  }
  for (i = 0; i < n - 1; i++) {
    j = i;
    for (k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
    if (i != j) {
      std::swap(W[i], W[j]);
      if (Vt) {
        for (k = 0; k < m; k++) std::swap(At[i * astep + k], At[j * astep + k]);
        for (k = 0; k < n; k++) std::swap(Vt[i * vstep + k], Vt[j * vstep + k]);
      }
    }
  }
  for (i = 0; i < n; i++) W_[i] = (float)W[i];
  if (!Vt) return;
  uint64 rng = 0x12345678;
  for (i = 0; i < n1; i++) {
    sd = i < n ? W[i] : 0;
    for (int ii = 0; ii < 100 && sd <= minval; ii++) {
      const float val0 = (float)(1. / m);
      for (k = 0; k < m; k++) {
        rng = (uint64)(unsigned)rng * 4164903690U + (unsigned)(rng >> 32);
        At[i * astep + k] = ((unsigned)rng & 256) != 0 ? val0 : -val0;
      }
      for (iter = 0; iter < 2; iter++)
        for (j = 0; j < i; j++) {
          sd = 0;
          for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
          float asum = 0;
          for (k = 0; k < m; k++) { float t = (float)(At[i * astep + k] - sd * At[j * astep + k]); At[i * astep + k] = t; asum += std::abs(t); }
          asum = asum > eps * 100 ? 1 / asum : 0;
          for (k = 0; k < m; k++) At[i * astep + k] *= asum;
        }
      sd = 0;
      for (k = 0; k < m; k++) { float t = At[i * astep + k]; sd += (double)t * t; }
      sd = std::sqrt(sd);
    }
    s = (float)(sd > minval ? 1 / sd : 0.);
    for (k = 0; k < m; k++) At[i * astep + k] *= s;
  }
}
void SVD::compute(InputArray src_, OutputArray w_, OutputArray u_, OutputArray vt_, int flags) {
  Mat src = src_.getMat();
  CV_Assert(src.type() == CV_32FC1 && src.rows == src.cols && flags == 0);   // the only form the reference uses (3x3 float)
  int n = src.rows;
  Mat w(n, 1, CV_32F), u(n, n, CV_32F), vt(n, n, CV_32F);
  if (cvshim::hook_fn h = cvshim::get_hook(cvshim::OP_SVD3)) {
    Mat a = continuous(src);
    std::vector<float> out(n + 2 * n * n);
    int dims[1] = {n};
    h(cvshim::OP_SVD3, a.data, nullptr, out.data(), dims);
    std::memcpy(w.data, out.data(), sizeof(float) * n);
    std::memcpy(u.data, out.data() + n, sizeof(float) * n * n);
    std::memcpy(vt.data, out.data() + n + n * n, sizeof(float) * n * n);
  } else {
    Mat at = src.t();   // temp_a = A^T (m >= n branch of _SVDcompute)
    jacobi_svd32f(at.ptr<float>(), at.step1(), w.ptr<float>(), vt.ptr<float>(), vt.step1(), n, n, n);
    u = at.t();         // rows of temp_u are the left singular vectors -> transpose into U
  }
  w.copyTo(w_); u.copyTo(u_); vt.copyTo(vt_);
}

// ---------------------------------------------------------------- imgproc
void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
  // 8UC3, 7x7, sigma 0, BORDER_REPLICATE (linemod.cpp:247): OpenCV's fixed small-kernel table {8,28,56,72,56,28,8}/256, both
  // passes exact integers, one rounding (sum + 2^15) >> 16 (the 8-bit fixed-point separable filter)
  Mat src = continuous(src_.getMat());
  CV_Assert(src.type() == CV_8UC3 && ksize.width == 7 && ksize.height == 7 && sigmaX == 0 && sigmaY == 0 && borderType == BORDER_REPLICATE);
  int W = src.cols, H = src.rows;
  Mat out(H, W, CV_8UC3);
  if (cvshim::hook_fn h = cvshim::get_hook(cvshim::OP_GAUSSIAN7)) {
    int dims[3] = {H, W, 3};
    h(cvshim::OP_GAUSSIAN7, src.data, nullptr, out.data, dims);
  } else {
    static const int k[7] = {8, 28, 56, 72, 56, 28, 8};
    std::vector<int> tmp((size_t)W * H * 3);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < 3; ++c) {
          int s = 0;
          for (int i = 0; i < 7; ++i) s += k[i] * src.ptr(y)[clampi(x + i - 3, 0, W - 1) * 3 + c];
          tmp[((size_t)y * W + x) * 3 + c] = s;
        }
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < 3; ++c) {
          int s = 0;
          for (int i = 0; i < 7; ++i) s += k[i] * tmp[((size_t)clampi(y + i - 3, 0, H - 1) * W + x) * 3 + c];
          out.ptr(y)[x * 3 + c] = (uchar)((s + 32768) >> 16);
        }
  }
  out.copyTo(dst_);
}
void Sobel(InputArray src_, OutputArray dst_, int ddepth, int dx, int dy, int ksize, double scale, double delta, int borderType) {
  // 8UC3 -> 16SC3, 3x3, BORDER_REPLICATE (linemod.cpp:248-249): exact integers
  Mat src = continuous(src_.getMat());
  CV_Assert(src.depth() == CV_8U && ddepth == CV_16S && ksize == 3 && scale == 1.0 && delta == 0.0 && borderType == BORDER_REPLICATE &&
            ((dx == 1 && dy == 0) || (dx == 0 && dy == 1)));
  int W = src.cols, H = src.rows, cn = src.channels();
  Mat out(H, W, CV_MAKETYPE(CV_16S, cn));
  int op = dx == 1 ? cvshim::OP_SOBEL_DX : cvshim::OP_SOBEL_DY;
  if (cvshim::hook_fn h = cvshim::get_hook(op)) {
    int dims[3] = {H, W, cn};
    h(op, src.data, nullptr, out.data, dims);
  } else {
    for (int y = 0; y < H; ++y) {
      int ym = clampi(y - 1, 0, H - 1), yp = clampi(y + 1, 0, H - 1);
      for (int x = 0; x < W; ++x) {
        int xm = clampi(x - 1, 0, W - 1), xp = clampi(x + 1, 0, W - 1);
        for (int c = 0; c < cn; ++c) {
#define P(yy, xx) ((int)src.ptr(yy)[(xx) * cn + c])
          int g = dx == 1 ? (P(ym, xp) + 2 * P(y, xp) + P(yp, xp)) - (P(ym, xm) + 2 * P(y, xm) + P(yp, xm))
                          : (P(yp, xm) + 2 * P(yp, x) + P(yp, xp)) - (P(ym, xm) + 2 * P(ym, x) + P(ym, xp));
#undef P
          out.ptr<short>(y)[x * cn + c] = (short)g;
        }
      }
    }
  }
  out.copyTo(dst_);
}
void medianBlur(InputArray src_, OutputArray dst_, int ksize) {
  // 8UC1, ksize 5 (linemod.cpp:684): exact 13th of 25, replicated borders; in-place allowed
  Mat src = continuous(src_.getMat()).clone();
  CV_Assert(src.type() == CV_8UC1 && ksize == 5);
  int W = src.cols, H = src.rows;
  Mat out(H, W, CV_8U);
  if (cvshim::hook_fn h = cvshim::get_hook(cvshim::OP_MEDIAN5)) {
    int dims[2] = {H, W};
    h(cvshim::OP_MEDIAN5, src.data, nullptr, out.data, dims);
  } else {
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        uchar v[25]; int n = 0;
        for (int j = -2; j <= 2; ++j) for (int i = -2; i <= 2; ++i) v[n++] = src.ptr(clampi(y + j, 0, H - 1))[clampi(x + i, 0, W - 1)];
        std::nth_element(v, v + 12, v + 25);
        out.ptr(y)[x] = v[12];
      }
  }
  out.copyTo(dst_);
}
void pyrDown(InputArray src_, OutputArray dst_, const Size& dstsize, int borderType) {
  // 8UC3 -> (cols/2, rows/2), 5x5 [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, (sum + 128) >> 8 (linemod.cpp:441-444)
  Mat src = continuous(src_.getMat());
  CV_Assert(src.depth() == CV_8U && borderType == BORDER_DEFAULT);
  int W = src.cols, H = src.rows, cn = src.channels();
  int dw = dstsize.width > 0 ? dstsize.width : (W + 1) / 2, dh = dstsize.height > 0 ? dstsize.height : (H + 1) / 2;
  CV_Assert(std::abs(dw * 2 - W) <= 2 && std::abs(dh * 2 - H) <= 2);
  Mat out(dh, dw, src.type());
  if (cvshim::hook_fn h = cvshim::get_hook(cvshim::OP_PYRDOWN)) {
    int dims[5] = {H, W, cn, dh, dw};
    h(cvshim::OP_PYRDOWN, src.data, nullptr, out.data, dims);
  } else {
    static const int k[5] = {1, 4, 6, 4, 1};
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x)
        for (int c = 0; c < cn; ++c) {
          int s = 0;
          for (int j = 0; j < 5; ++j) {
            int sy = reflect101(2 * y + j - 2, H), rs = 0;
            for (int i = 0; i < 5; ++i) rs += k[i] * src.ptr(sy)[reflect101(2 * x + i - 2, W) * cn + c];
            s += k[j] * rs;
          }
          out.ptr(y)[x * cn + c] = (uchar)((s + 128) >> 8);
        }
  }
  out.copyTo(dst_);
}
void resize(InputArray src_, OutputArray dst_, Size dsize, double fx, double fy, int interpolation) {
  // INTER_NEAREST on 8UC1 (linemod.cpp:448, 731, 736): sx = min(floor(x * W / dw), W - 1)
  Mat src = continuous(src_.getMat());
  CV_Assert(interpolation == INTER_NEAREST && fx == 0 && fy == 0 && dsize.width > 0 && dsize.height > 0);
  int W = src.cols, H = src.rows, dw = dsize.width, dh = dsize.height;
  size_t esz = src.elemSize();
  Mat out(dh, dw, src.type());
  if (cvshim::hook_fn h = cvshim::get_hook(cvshim::OP_RESIZE_NN)) {
    int dims[5] = {H, W, (int)esz, dh, dw};
    h(cvshim::OP_RESIZE_NN, src.data, nullptr, out.data, dims);
  } else {
    double ifx = (double)W / dw, ify = (double)H / dh;
    for (int y = 0; y < dh; ++y) {
      int sy = std::min((int)std::floor(y * ify), H - 1);
      for (int x = 0; x < dw; ++x) {
        int sx = std::min((int)std::floor(x * ifx), W - 1);
        std::memcpy(out.ptr(y) + x * esz, src.ptr(sy) + sx * esz, esz);
      }
    }
  }
  out.copyTo(dst_);
}
void erode(InputArray src_, OutputArray dst_, InputArray kernel, Point, int iterations, int borderType) {
  // 3x3 rectangle (empty kernel), BORDER_REPLICATE, 8UC1 - template training only (linemod.cpp:466, 753)
  Mat cur = continuous(src_.getMat()).clone();
  CV_Assert(cur.type() == CV_8UC1 && kernel.empty() && borderType == BORDER_REPLICATE);
  int W = cur.cols, H = cur.rows;
  for (int it = 0; it < iterations; ++it) {
    Mat nxt(H, W, CV_8U);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        uchar m = 255;
        for (int j = -1; j <= 1; ++j) for (int i = -1; i <= 1; ++i) m = std::min(m, cur.ptr(clampi(y + j, 0, H - 1))[clampi(x + i, 0, W - 1)]);
        nxt.ptr(y)[x] = m;
      }
    cur = nxt;
  }
  cur.copyTo(dst_);
}
void distanceTransform(InputArray src_, OutputArray dst_, int distanceType, int maskSize, int dstType) {
  // DIST_C, 3x3 (linemod.cpp:765): OpenCV's two-pass 3x3 chamfer with a = b = 1 in 16.16 fixed point over a 1-px border of
  // INIT_DIST0 = INT_MAX >> 2; output float = min(dist, INIT_DIST0) / 65536 - template training only
  Mat src = continuous(src_.getMat());
  CV_Assert(src.type() == CV_8UC1 && distanceType == DIST_C && maskSize == 3 && dstType == CV_32F);
  int W = src.cols, H = src.rows;
  {   // an image without a single zero pixel: OpenCV 4.13's own code returns 65535 everywhere (its IPP path FLT_MAX); as for the resize,
      // the stand-in follows OpenCV's code (cv2.ipp.setUseIPP(False))
    bool any_zero = false;
    for (int y = 0; y < H && !any_zero; ++y) for (int x = 0; x < W; ++x) if (!src.ptr(y)[x]) { any_zero = true; break; }
    if (!any_zero) { Mat out(H, W, CV_32F); for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) out.ptr<float>(y)[x] = 65535.f; out.copyTo(dst_); return; }
  }
  const unsigned INIT = (unsigned)(INT_MAX >> 2), HV = 1u << 16, DIAG = 1u << 16;
  size_t step = (size_t)W + 2;
  std::vector<unsigned> tmp((size_t)(H + 2) * step, INIT);
  for (int y = 0; y < H; ++y) {
    unsigned* t = &tmp[(size_t)(y + 1) * step + 1];
    for (int x = 0; x < W; ++x) {
      if (!src.ptr(y)[x]) { t[x] = 0; continue; }
      unsigned t0 = t[x - (long)step - 1] + DIAG, v = t[x - (long)step] + HV;
      if (t0 > v) t0 = v;
      v = t[x - (long)step + 1] + DIAG; if (t0 > v) t0 = v;
      v = t[x - 1] + HV; if (t0 > v) t0 = v;
      t[x] = t0;
    }
  }
  Mat out(H, W, CV_32F);
  for (int y = H - 1; y >= 0; --y) {
    unsigned* t = &tmp[(size_t)(y + 1) * step + 1];
    for (int x = W - 1; x >= 0; --x) {
      unsigned t0 = t[x];
      if (t0 > HV) {
        unsigned v = t[x + step + 1] + DIAG; if (t0 > v) t0 = v;
        v = t[x + step] + HV; if (t0 > v) t0 = v;
        v = t[x + step - 1] + DIAG; if (t0 > v) t0 = v;
        v = t[x + 1] + HV; if (t0 > v) t0 = v;
        t[x] = t0;
      }
      t0 = t0 > INIT ? INIT : t0;
      out.ptr<float>(y)[x] = (float)(t0 * (1.f / 65536.f));
    }
  }
  out.copyTo(dst_);
}
void circle(Mat&, Point, int, const Scalar&, int) {}      // drawing helpers are viewers' business
void rectangle(Mat&, Rect, const Scalar&, int) {}

// ---------------------------------------------------------------- persistence: not part of the hot path; see ref_filestorage.cpp
}  // namespace cv

// ---------------------------------------------------------------- cvflann: exact 1-NN
namespace cvflann {
struct KdNode { int lo, hi, dim, left, right; float split; };
struct KdImpl { std::vector<KdNode> nodes; std::vector<int> idx; std::vector<float> pts; size_t n; int leaf; };
static int kd_rec(KdImpl* t, int lo, int hi) {
  int id = (int)t->nodes.size();
  KdNode nd; nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1; nd.dim = 0; nd.split = 0.f;
  t->nodes.push_back(nd);
  if (hi - lo > t->leaf) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = lo; i < hi; ++i) for (int k = 0; k < 3; ++k) { float v = t->pts[3 * (size_t)t->idx[i] + k]; mn[k] = std::min(mn[k], v); mx[k] = std::max(mx[k], v); }
    int dim = 0; for (int k = 1; k < 3; ++k) if (mx[k] - mn[k] > mx[dim] - mn[dim]) dim = k;
    if (mx[dim] > mn[dim]) {
      int mid = (lo + hi) / 2;
      std::nth_element(t->idx.begin() + lo, t->idx.begin() + mid, t->idx.begin() + hi,
                       [t, dim](int a, int b) { return t->pts[3 * (size_t)a + dim] < t->pts[3 * (size_t)b + dim]; });
      float split = t->pts[3 * (size_t)t->idx[mid] + dim];
      int l = kd_rec(t, lo, mid), r = kd_rec(t, mid, hi);
      t->nodes[id].dim = dim; t->nodes[id].split = split; t->nodes[id].left = l; t->nodes[id].right = r;
    }
  }
  return id;
}
KdImpl* kd_build(const float* pts, size_t n, size_t stride, int leaf) {
  KdImpl* t = new KdImpl; t->n = n; t->leaf = leaf > 0 ? leaf : 10;
  t->pts.resize(n * 3); t->idx.resize(n);
  for (size_t i = 0; i < n; ++i) { t->idx[i] = (int)i; for (int k = 0; k < 3; ++k) t->pts[3 * i + k] = pts[i * stride + k]; }
  if (n) kd_rec(t, 0, (int)n);
  return t;
}
void kd_free(KdImpl* t) { delete t; }
static void kd_query(const KdImpl* t, int id, const float* q, float& best, int& best_i) {
  const KdNode& nd = t->nodes[id];
  if (nd.left < 0) {
    for (int i = nd.lo; i < nd.hi; ++i) {
      const float* p = &t->pts[3 * (size_t)t->idx[i]];
      float d0 = q[0] - p[0], d1 = q[1] - p[1], d2 = q[2] - p[2];
      float dd = 0.f; dd += d0 * d0; dd += d1 * d1; dd += d2 * d2;   // L2_Simple: result += diff*diff in fp32, x then y then z
      if (dd < best || (dd == best && t->idx[i] < best_i)) { best = dd; best_i = t->idx[i]; }
    }
    return;
  }
  float diff = q[nd.dim] - nd.split;
  int nearc = diff < 0 ? nd.left : nd.right, farc = diff < 0 ? nd.right : nd.left;
  kd_query(t, nearc, q, best, best_i);
  if (diff * diff <= best) kd_query(t, farc, q, best, best_i);
}
void kd_knn1(const KdImpl* t, const float* q, size_t nq, size_t qstride, int* idx, size_t istride, float* dist, size_t dstride) {
  if (cvshim::hook_fn h = cvshim::get_hook(cvshim::OP_KNN1)) {
    std::vector<float> qq(nq * 3); std::vector<int> oi(nq); std::vector<float> od(nq);
    for (size_t i = 0; i < nq; ++i) for (int k = 0; k < 3; ++k) qq[3 * i + k] = q[i * qstride + k];
    int dims[2] = {(int)t->n, (int)nq};
    std::vector<char> out(nq * 8);
    h(cvshim::OP_KNN1, t->pts.data(), qq.data(), out.data(), dims);   // out = [nq int32 indices][nq float32 squared distances]
    std::memcpy(oi.data(), out.data(), nq * 4); std::memcpy(od.data(), out.data() + nq * 4, nq * 4);
    for (size_t i = 0; i < nq; ++i) { idx[i * istride] = oi[i]; dist[i * dstride] = od[i]; }
    return;
  }
  for (size_t i = 0; i < nq; ++i) {
    float best = FLT_MAX; int bi = -1;
    if (t->n) kd_query(t, 0, q + i * qstride, best, bi);
    idx[i * istride] = bi; dist[i * dstride] = best;
  }
}
}  // namespace cvflann
