// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// cvshim: a stand-in for the slice of the OpenCV C++ API that the reference's hot-path sources use, so that
// /root/reference/linemod/linemod.cpp, ICP/ICP.cpp, ICP/NMS.cpp, ICP/common.cpp, ICP/detection.cpp and
// ICP/depth_to_3d.cpp compile UNMODIFIED, from where they lie, into oracle/_ref/libfl_ref.so (recipe: oracle/build_ref.py).
// The image has no OpenCV C++ SDK (no headers, no libs), which is the only reason this file exists.
//
// What is the reference's and what is ours:
//   * every line of LINE-MOD / ICP / NMS logic executed by libfl_ref.so is the reference's own source;
//   * the containers (Mat, Matx, Vec, Ptr, InputArray ...) are ours and carry no arithmetic beyond what OpenCV documents
//     for them (Matx products accumulate `s = 0; s += a*b` in the element type, cv::norm accumulates in double, ...);
//   * the image/maths primitives OpenCV would supply (GaussianBlur, Sobel, phase, medianBlur, pyrDown, resize(NEAREST),
//     convertTo, SVD::compute, cvflann KD-tree 1-NN) are restated in cvshim_impl.cpp from OpenCV's published algorithms and
//     are each pinned against the real cv2 4.13 of this image by tests/test_oracle_ref.py; every one of them can also be
//     REPLACED at run time by a callback (cvshim::set_hook) so that the tests can run the reference's code on the real
//     OpenCV primitives (cv2 through ctypes callbacks) and show that nothing changes.
//
// Only what the reference's sources touch is implemented; anything else is absent on purpose.
#ifndef FL_REF_CVSHIM_HPP
#define FL_REF_CVSHIM_HPP

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <emmintrin.h>
#include <pmmintrin.h>
#include <tmmintrin.h>

// the reference's SSE paths are compiled in, as a stock x86-64 OpenCV 3.x build would have them
#define CV_SSE2 1
#define CV_SSE3 1
#define CV_SSSE3 1
#define CV_MAJOR_VERSION 3
#define CV_MINOR_VERSION 4

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef signed char schar;
typedef long long int64;
typedef unsigned long long uint64;

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAKE_TYPE CV_MAKETYPE
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16UC3 CV_MAKETYPE(CV_16U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16SC3 CV_MAKETYPE(CV_16S, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_DECL_ALIGNED(x) __attribute__((aligned(x)))
#define CV_WINDOW_AUTOSIZE 1
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)

#define CV_CPU_SSE2 2
#define CV_CPU_SSE3 3
#define CV_CPU_SSSE3 4

namespace cv {

typedef std::string String;

namespace Error { enum { StsBadArg = -5, StsAssert = -215, StsNotImplemented = -213 }; }

class Exception : public std::runtime_error {
 public:
  Exception(int code_, const std::string& m) : std::runtime_error(m), code(code_), msg(m) {}
  int code;
  std::string msg;
};
[[noreturn]] inline void error(int code, const std::string& msg, const char* file, int line) {
  std::ostringstream s; s << file << ":" << line << ": error (" << code << ") " << msg;
  throw Exception(code, s.str());
}
#define CV_Error(code, msg) ::cv::error((code), (msg), __FILE__, __LINE__)
#define CV_Assert(expr) do { if (!(expr)) ::cv::error(::cv::Error::StsAssert, #expr, __FILE__, __LINE__); } while (0)
#define CV_DbgAssert(expr) ((void)0)

enum { CPU_SSE2 = CV_CPU_SSE2, CPU_SSE3 = CV_CPU_SSE3, CPU_SSSE3 = CV_CPU_SSSE3 };
inline bool checkHardwareSupport(int) { return true; }   // built with -msse4.2; the GPU boxes and this container have it

int64 getTickCount();
double getTickFrequency();
String format(const char* fmt, ...);

inline int cvRound_(double v) { return (int)lrint(v); }              // round-half-even == OpenCV's cvRound (SSE2 cvtsd2si)
template <typename T> inline T saturate_cast(uchar v) { return (T)v; }
template <typename T> inline T saturate_cast(schar v) { return (T)v; }
template <typename T> inline T saturate_cast(ushort v) { return (T)v; }
template <typename T> inline T saturate_cast(short v) { return (T)v; }
template <typename T> inline T saturate_cast(int v) { return (T)v; }
template <typename T> inline T saturate_cast(unsigned v) { return (T)v; }
template <typename T> inline T saturate_cast(float v) { return (T)v; }
template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0); }
template <> inline uchar saturate_cast<uchar>(short v) { return saturate_cast<uchar>((int)v); }
template <> inline uchar saturate_cast<uchar>(ushort v) { return (uchar)std::min((unsigned)v, 255u); }
template <> inline uchar saturate_cast<uchar>(unsigned v) { return (uchar)std::min(v, 255u); }
template <> inline uchar saturate_cast<uchar>(float v) { return saturate_cast<uchar>(cvRound_(v)); }
template <> inline uchar saturate_cast<uchar>(double v) { return saturate_cast<uchar>(cvRound_(v)); }
template <> inline ushort saturate_cast<ushort>(int v) { return (ushort)((unsigned)v <= 65535u ? v : v > 0 ? 65535 : 0); }
template <> inline ushort saturate_cast<ushort>(short v) { return (ushort)std::max((int)v, 0); }
template <> inline ushort saturate_cast<ushort>(unsigned v) { return (ushort)std::min(v, 65535u); }
template <> inline ushort saturate_cast<ushort>(float v) { return saturate_cast<ushort>(cvRound_(v)); }
template <> inline ushort saturate_cast<ushort>(double v) { return saturate_cast<ushort>(cvRound_(v)); }
template <> inline short saturate_cast<short>(int v) { return (short)((unsigned)(v + 32768) <= 65535u ? v : v > 0 ? 32767 : -32768); }
template <> inline short saturate_cast<short>(ushort v) { return (short)std::min((int)v, 32767); }
template <> inline short saturate_cast<short>(float v) { return saturate_cast<short>(cvRound_(v)); }
template <> inline short saturate_cast<short>(double v) { return saturate_cast<short>(cvRound_(v)); }
template <> inline int saturate_cast<int>(float v) { return cvRound_(v); }
template <> inline int saturate_cast<int>(double v) { return cvRound_(v); }
}  // namespace cv
inline int cvRound(double v) { return cv::cvRound_(v); }
inline int cvIsNaN(double v) { return std::isnan(v) ? 1 : 0; }
inline int cvIsNaN(float v) { return std::isnan(v) ? 1 : 0; }

namespace cv {

// ---------------------------------------------------------------- small geometric types
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  T area() const { return width * height; }
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
};
typedef Rect_<int> Rect;

// ---------------------------------------------------------------- Matx / Vec (arithmetic as in opencv2/core/matx.hpp)
template <typename T, int m, int n> class Matx {
 public:
  enum { rows = m, cols = n, channels = m * n };
  T val[m * n];
  Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
  explicit Matx(T v0) { static_assert(m * n >= 1, ""); for (int i = 0; i < m * n; ++i) val[i] = T(0); val[0] = v0; }
  Matx(T v0, T v1) { static_assert(m * n >= 2, ""); for (int i = 0; i < m * n; ++i) val[i] = T(0); val[0] = v0; val[1] = v1; }
  Matx(T v0, T v1, T v2) { static_assert(m * n >= 3, ""); for (int i = 0; i < m * n; ++i) val[i] = T(0); val[0] = v0; val[1] = v1; val[2] = v2; }
  Matx(T v0, T v1, T v2, T v3) { static_assert(m * n >= 4, ""); for (int i = 0; i < m * n; ++i) val[i] = T(0); val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
  Matx(T v0, T v1, T v2, T v3, T v4, T v5, T v6, T v7, T v8) {
    static_assert(m * n >= 9, "");
    for (int i = 0; i < m * n; ++i) val[i] = T(0);
    val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; val[4] = v4; val[5] = v5; val[6] = v6; val[7] = v7; val[8] = v8;
  }
  static Matx eye() { Matx r; for (int i = 0; i < (m < n ? m : n); ++i) r.val[i * n + i] = T(1); return r; }
  static Matx zeros() { return Matx(); }
  T& operator()(int i, int j) { return val[i * n + j]; }
  const T& operator()(int i, int j) const { return val[i * n + j]; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  Matx<T, n, m> t() const { Matx<T, n, m> r; for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) r.val[j * m + i] = val[i * n + j]; return r; }
};
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<double, 3, 3> Matx33d;

template <typename T, int cn> class Vec : public Matx<T, cn, 1> {
 public:
  typedef T value_type;
  Vec() {}
  Vec(T v0, T v1) : Matx<T, cn, 1>(v0, v1) {}
  Vec(T v0, T v1, T v2) : Matx<T, cn, 1>(v0, v1, v2) {}
  Vec(T v0, T v1, T v2, T v3) : Matx<T, cn, 1>(v0, v1, v2, v3) {}
  Vec(const Matx<T, cn, 1>& a) : Matx<T, cn, 1>(a) {}
  T& operator[](int i) { return this->val[i]; }
  const T& operator[](int i) const { return this->val[i]; }
  T& operator()(int i) { return this->val[i]; }
  const T& operator()(int i) const { return this->val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 4> Scalar_d;

struct Scalar {
  double val[4];
  Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
  Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
  double operator[](int i) const { return val[i]; }
};

// Matx_MatMulOp: `_Tp s = 0; for k: s += a(i,k) * b(k,j)` in the element type (matx.hpp)
template <typename T, int m, int l, int n> inline Matx<T, m, n> operator*(const Matx<T, m, l>& a, const Matx<T, l, n>& b) {
  Matx<T, m, n> c;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      T s = 0;
      for (int k = 0; k < l; ++k) s += a(i, k) * b(k, j);
      c.val[i * n + j] = s;
    }
  return c;
}
template <typename T, int m, int n> inline Vec<T, m> operator*(const Matx<T, m, n>& a, const Vec<T, n>& b) {
  Matx<T, m, 1> c = a * static_cast<const Matx<T, n, 1>&>(b);
  return Vec<T, m>(c);
}
template <typename T, int m, int n> inline Matx<T, m, n> operator+(const Matx<T, m, n>& a, const Matx<T, m, n>& b) {
  Matx<T, m, n> c; for (int i = 0; i < m * n; ++i) c.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return c;
}
template <typename T, int m, int n> inline Matx<T, m, n> operator-(const Matx<T, m, n>& a, const Matx<T, m, n>& b) {
  Matx<T, m, n> c; for (int i = 0; i < m * n; ++i) c.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return c;
}
template <typename T, int m, int n> inline Matx<T, m, n> operator-(const Matx<T, m, n>& a) {
  Matx<T, m, n> c; for (int i = 0; i < m * n; ++i) c.val[i] = saturate_cast<T>(-a.val[i]); return c;   // Matx_ScaleOp with alpha = -1
}
template <typename T, int m, int n> inline Matx<T, m, n>& operator+=(Matx<T, m, n>& a, const Matx<T, m, n>& b) {
  for (int i = 0; i < m * n; ++i) a.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return a;
}
template <typename T, int m, int n> inline Matx<T, m, n>& operator-=(Matx<T, m, n>& a, const Matx<T, m, n>& b) {
  for (int i = 0; i < m * n; ++i) a.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return a;
}
template <typename T, int cn> inline Vec<T, cn> operator+(const Vec<T, cn>& a, const Vec<T, cn>& b) {
  Vec<T, cn> c; for (int i = 0; i < cn; ++i) c.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return c;
}
template <typename T, int cn> inline Vec<T, cn> operator-(const Vec<T, cn>& a, const Vec<T, cn>& b) {
  Vec<T, cn> c; for (int i = 0; i < cn; ++i) c.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return c;
}
template <typename T, int cn> inline Vec<T, cn> operator-(const Vec<T, cn>& a) {
  Vec<T, cn> c; for (int i = 0; i < cn; ++i) c.val[i] = saturate_cast<T>(-a.val[i]); return c;
}
template <typename T, int cn> inline Vec<T, cn>& operator+=(Vec<T, cn>& a, const Vec<T, cn>& b) {
  for (int i = 0; i < cn; ++i) a.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return a;
}
template <typename T, int cn> inline Vec<T, cn>& operator-=(Vec<T, cn>& a, const Vec<T, cn>& b) {
  for (int i = 0; i < cn; ++i) a.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return a;
}
template <typename T, int cn> inline Vec<T, cn>& operator*=(Vec<T, cn>& a, int alpha) {
  for (int i = 0; i < cn; ++i) a.val[i] = saturate_cast<T>(a.val[i] * alpha); return a;
}
template <typename T, int cn> inline Vec<T, cn>& operator*=(Vec<T, cn>& a, float alpha) {
  for (int i = 0; i < cn; ++i) a.val[i] = saturate_cast<T>(a.val[i] * alpha); return a;
}
template <typename T, int cn> inline Vec<T, cn>& operator*=(Vec<T, cn>& a, double alpha) {
  for (int i = 0; i < cn; ++i) a.val[i] = saturate_cast<T>(a.val[i] * alpha); return a;
}
// cv::norm(Matx): normL2Sqr<_Tp, double> then std::sqrt, i.e. the squares accumulate in double (matx.hpp / base.hpp)
template <typename T, int m, int n> inline double norm(const Matx<T, m, n>& M) {
  double s = 0; for (int i = 0; i < m * n; ++i) { double v = (double)M.val[i]; s += v * v; } return std::sqrt(s);
}

// ---------------------------------------------------------------- DataType traits
template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { depth = CV_8U, channels = 1, type = CV_8UC1 }; typedef uchar channel_type; };
template <> struct DataType<schar> { enum { depth = CV_8S, channels = 1, type = CV_MAKETYPE(CV_8S, 1) }; typedef schar channel_type; };
template <> struct DataType<ushort> { enum { depth = CV_16U, channels = 1, type = CV_16UC1 }; typedef ushort channel_type; };
template <> struct DataType<short> { enum { depth = CV_16S, channels = 1, type = CV_16SC1 }; typedef short channel_type; };
template <> struct DataType<int> { enum { depth = CV_32S, channels = 1, type = CV_32SC1 }; typedef int channel_type; };
template <> struct DataType<float> { enum { depth = CV_32F, channels = 1, type = CV_32FC1 }; typedef float channel_type; };
template <> struct DataType<double> { enum { depth = CV_64F, channels = 1, type = CV_64FC1 }; typedef double channel_type; };
template <typename T, int cn> struct DataType<Vec<T, cn> > {
  enum { depth = DataType<T>::depth, channels = cn, type = CV_MAKETYPE(DataType<T>::depth, cn) }; typedef T channel_type;
};

// ---------------------------------------------------------------- Ptr
template <typename T> class Ptr : public std::shared_ptr<T> {
 public:
  Ptr() {}
  Ptr(T* p) : std::shared_ptr<T>(p) {}
  Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
  template <typename Y> Ptr(const Ptr<Y>& o) : std::shared_ptr<T>(static_cast<const std::shared_ptr<Y>&>(o)) {}
  bool empty() const { return !this->get(); }
};
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return Ptr<T>(std::make_shared<T>(std::forward<A>(a)...)); }

// ---------------------------------------------------------------- Mat
class Mat;
class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef InputArray InputArrayOfArrays;
typedef OutputArray OutputArrayOfArrays;
typedef OutputArray InputOutputArray;

struct MatSize {
  const Mat* m;
  explicit MatSize(const Mat* m_) : m(m_) {}
  Size operator()() const;
  bool operator==(const MatSize& o) const { return (*this)() == o(); }
  bool operator!=(const MatSize& o) const { return !((*this)() == o()); }
};
struct MatStep {
  size_t v;
  MatStep() : v(0) {}
  operator size_t() const { return v; }
  MatStep& operator=(size_t s) { v = s; return *this; }
};

template <typename T> class Mat_;
template <typename T> class MatConstIterator_;
template <typename T> class MatIterator_;

class Mat {
 public:
  int flags;     // the type (depth + channels), nothing else
  int rows, cols;
  uchar* data;
  MatStep step;  // bytes per row
  MatSize size;

  Mat() : flags(0), rows(0), cols(0), data(nullptr), size(this) {}
  Mat(int r, int c, int type) : flags(0), rows(0), cols(0), data(nullptr), size(this) { create(r, c, type); }
  Mat(Size s, int type) : flags(0), rows(0), cols(0), data(nullptr), size(this) { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& s) : flags(0), rows(0), cols(0), data(nullptr), size(this) { create(r, c, type); setTo(s); }
  Mat(int r, int c, int type, void* user, size_t step_ = 0) : flags(type), rows(r), cols(c), data((uchar*)user), size(this) {
    step = step_ ? step_ : (size_t)c * elemSize();
  }
  Mat(const Mat& o) : flags(o.flags), rows(o.rows), cols(o.cols), data(o.data), step(o.step), size(this), owner_(o.owner_) {}
  Mat(const Mat& o, const Rect& roi) : flags(o.flags), rows(roi.height), cols(roi.width), data(nullptr), step(o.step), size(this), owner_(o.owner_) {
    // Mat::Mat(const Mat&, const Rect&): the ROI must lie inside the matrix (core/src/matrix.cpp CV_Assert)
    CV_Assert(0 <= roi.x && 0 <= roi.width && roi.x + roi.width <= o.cols && 0 <= roi.y && 0 <= roi.height && roi.y + roi.height <= o.rows);
    data = o.data + (size_t)roi.y * o.step + (size_t)roi.x * o.elemSize();
  }
  template <typename T, int m, int n> explicit Mat(const Matx<T, m, n>& M, bool copyData = true)
      : flags(0), rows(0), cols(0), data(nullptr), size(this) {
    (void)copyData;
    create(m, n, DataType<T>::type);
    std::memcpy(data, M.val, sizeof(T) * m * n);
  }
  template <typename T, int n> explicit Mat(const Vec<T, n>& v, bool copyData = true) : flags(0), rows(0), cols(0), data(nullptr), size(this) {
    (void)copyData;
    create(n, 1, DataType<T>::type);
    std::memcpy(data, v.val, sizeof(T) * n);
  }
  template <typename T> explicit Mat(const std::vector<T>& v, bool copyData = false) : flags(0), rows(0), cols(0), data(nullptr), size(this) {
    (void)copyData;
    if (!v.empty()) { create((int)v.size(), 1, DataType<T>::type); std::memcpy(data, v.data(), sizeof(T) * v.size()); }
  }
  Mat& operator=(const Mat& o) {
    if (this != &o) { flags = o.flags; rows = o.rows; cols = o.cols; data = o.data; step = o.step; owner_ = o.owner_; }
    return *this;
  }
  Mat& operator=(const Scalar& s) { setTo(s); return *this; }

  void create(int r, int c, int type) {
    type &= 0xfff;
    if (data && r == rows && c == cols && type == this->type()) return;
    flags = type; rows = r; cols = c; step = (size_t)c * elemSize();
    size_t bytes = (size_t)r * step + 64;   // slack: the reference's SSE loops never read past it, OpenCV pads too
    void* p = nullptr;
    if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) throw std::bad_alloc();
    owner_ = std::shared_ptr<uchar>((uchar*)p, free);
    data = (uchar*)p;
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  void release() { owner_.reset(); data = nullptr; rows = cols = 0; }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); for (int y = 0; y < r; ++y) std::memset(m.data + (size_t)y * m.step, 0, (size_t)c * m.elemSize()); return m; }
  static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
  static Mat eye(int r, int c, int type);

  int type() const { return flags & 0xfff; }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize1() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)sz[depth()]; }
  size_t elemSize() const { return elemSize1() * channels(); }
  size_t step1() const { return step.v / elemSize1(); }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return rows <= 1 || step.v == (size_t)cols * elemSize(); }

  uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.v); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.v); }
  template <typename T> T* ptr(int r, int c) { return (T*)(data + (size_t)r * step.v) + c; }
  template <typename T> const T* ptr(int r, int c) const { return (const T*)(data + (size_t)r * step.v) + c; }
  template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step.v))[c]; }
  template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step.v))[c]; }
  template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step.v); }   // 1-D access of a row or column vector
  template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step.v); }

  Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
  Mat clone() const { Mat m; copyTo(m); return m; }
  void copyTo(Mat& dst) const;
  void copyTo(OutputArray dst) const;
  void copyTo(OutputArray dst, InputArray mask) const;
  void convertTo(OutputArray dst, int rtype, double alpha = 1, double beta = 0) const;
  Mat& setTo(const Scalar& s);
  Mat& setTo(const Scalar& s, InputArray mask);
  Mat t() const;
  Mat mul(const Mat& o, double scale = 1) const;
  Mat reshape(int cn, int new_rows = 0) const;
  void resize(size_t nrows);

  template <typename T> MatIterator_<T> begin();
  template <typename T> MatIterator_<T> end();
  template <typename T> MatConstIterator_<T> begin() const;
  template <typename T> MatConstIterator_<T> end() const;

  template <typename T, int m, int n> operator Matx<T, m, n>() const {
    CV_Assert(rows == m && cols == n && channels() == 1);
    Matx<T, m, n> r;
    if (type() == DataType<T>::type) { for (int i = 0; i < m; ++i) std::memcpy(&r.val[i * n], ptr(i), sizeof(T) * n); }
    else { Mat tmp; convert_into(tmp, DataType<T>::type); for (int i = 0; i < m; ++i) std::memcpy(&r.val[i * n], tmp.ptr(i), sizeof(T) * n); }
    return r;
  }
  template <typename T, int n> operator Vec<T, n>() const {
    CV_Assert((rows == n && cols == 1) || (rows == 1 && cols == n));
    Vec<T, n> r; Mat tmp; convert_into(tmp, DataType<T>::type);
    for (int i = 0; i < n; ++i) r.val[i] = tmp.at<T>(i);
    return r;
  }
  void convert_into(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;   // the body of convertTo

 protected:
  std::shared_ptr<uchar> owner_;
};
inline Size MatSize::operator()() const { return Size(m->cols, m->rows); }

template <typename T> class MatConstIterator_ {
 public:
  const Mat* m; int r, c;
  MatConstIterator_() : m(nullptr), r(0), c(0) {}
  MatConstIterator_(const Mat* m_, int r_, int c_) : m(m_), r(r_), c(c_) {}
  const T& operator*() const { return m->at<T>(r, c); }
  const T* operator->() const { return &m->at<T>(r, c); }
  MatConstIterator_& operator++() { if (++c >= m->cols) { c = 0; ++r; } return *this; }
  MatConstIterator_ operator++(int) { MatConstIterator_ t = *this; ++*this; return t; }
  bool operator==(const MatConstIterator_& o) const { return r == o.r && c == o.c; }
  bool operator!=(const MatConstIterator_& o) const { return !(*this == o); }
};
template <typename T> class MatIterator_ : public MatConstIterator_<T> {
 public:
  MatIterator_() {}
  MatIterator_(Mat* m_, int r_, int c_) : MatConstIterator_<T>(m_, r_, c_) {}
  T& operator*() const { return const_cast<Mat*>(this->m)->template at<T>(this->r, this->c); }
  T* operator->() const { return &const_cast<Mat*>(this->m)->template at<T>(this->r, this->c); }
  MatIterator_& operator++() { MatConstIterator_<T>::operator++(); return *this; }
  MatIterator_ operator++(int) { MatIterator_ t = *this; ++*this; return t; }
};
template <typename T> inline MatIterator_<T> Mat::begin() { return MatIterator_<T>(this, 0, 0); }
template <typename T> inline MatIterator_<T> Mat::end() { return MatIterator_<T>(this, empty() ? 0 : rows, 0); }
template <typename T> inline MatConstIterator_<T> Mat::begin() const { return MatConstIterator_<T>(this, 0, 0); }
template <typename T> inline MatConstIterator_<T> Mat::end() const { return MatConstIterator_<T>(this, empty() ? 0 : rows, 0); }

template <typename T> class Mat_ : public Mat {
 public:
  typedef T value_type;
  typedef MatIterator_<T> iterator;
  typedef MatConstIterator_<T> const_iterator;
  Mat_() { flags = DataType<T>::type; }
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  Mat_(int r, int c, const T& v) : Mat(r, c, DataType<T>::type) { for (int y = 0; y < r; ++y) for (int x = 0; x < c; ++x) (*this)(y, x) = v; }
  explicit Mat_(Size s) : Mat(s.height, s.width, DataType<T>::type) {}
  Mat_(const Mat& m) { flags = DataType<T>::type; *this = m; }
  Mat_(const Mat_& m) : Mat(m) {}
  Mat_& operator=(const Mat_& m) { Mat::operator=(m); return *this; }
  Mat_& operator=(const Mat& m) {   // Mat_<_Tp>::operator=(const Mat&): share when the type matches, else convert
    if (m.empty()) { release(); flags = DataType<T>::type; return *this; }
    if (m.type() == DataType<T>::type) { Mat::operator=(m); return *this; }
    if (m.depth() == DataType<T>::depth) { Mat::operator=(m.reshape(DataType<T>::channels, m.rows)); return *this; }
    CV_Assert(DataType<T>::channels == m.channels());
    Mat tmp; m.convert_into(tmp, DataType<T>::type); Mat::operator=(tmp); return *this;
  }
  T& operator()(int r, int c) { return this->template at<T>(r, c); }
  const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
  T& operator()(int i) { return this->template at<T>(i); }
  const T& operator()(int i) const { return this->template at<T>(i); }
  T* operator[](int r) { return this->template ptr<T>(r); }
  const T* operator[](int r) const { return this->template ptr<T>(r); }
  Mat_ operator()(const Rect& roi) const { return Mat_(Mat(*this, roi)); }
  iterator begin() { return Mat::begin<T>(); }
  iterator end() { return Mat::end<T>(); }
  const_iterator begin() const { return Mat::begin<T>(); }
  const_iterator end() const { return Mat::end<T>(); }
  Mat_ clone() const { return Mat_(Mat::clone()); }
};

// element-wise helpers that OpenCV expresses through MatExpr; evaluated eagerly here (float / double matrices)
Mat operator+(const Mat& a, const Mat& b);
Mat operator+(const Mat& a, double s);
Mat operator-(const Mat& a, const Mat& b);
Mat operator-(const Mat& a, double s);
Mat operator-(const Mat& a);
Mat operator*(const Mat& a, double s);
Mat operator*(double s, const Mat& a);
Mat operator*(const Mat& a, const Mat& b);   // matrix product (gemm)
Mat operator/(const Mat& a, double s);
Mat operator==(const Mat& a, double s);      // 8U mask, 255 where equal
Mat operator|(const Mat& a, const Mat& b);

// ---------------------------------------------------------------- InputArray / OutputArray proxies
struct _NoArrayTag {};
class _InputArray {
 public:
  _InputArray() : vec_(nullptr) {}
  _InputArray(const _NoArrayTag&) : vec_(nullptr) {}
  _InputArray(const Mat& m) : mat_(m), vec_(nullptr) {}
  template <typename T> _InputArray(const Mat_<T>& m) : mat_(m), vec_(nullptr) {}
  template <typename T, int m, int n> _InputArray(const Matx<T, m, n>& M) : mat_(m, n, DataType<T>::type, (void*)M.val), vec_(nullptr) {}
  template <typename T, int n> _InputArray(const Vec<T, n>& v) : mat_(n, 1, DataType<T>::type, (void*)v.val), vec_(nullptr) {}
  _InputArray(const std::vector<Mat>& v) : vec_(&v) {}
  Mat getMat(int = -1) const { return mat_; }
  const std::vector<Mat>* getVec() const { return vec_; }
  bool empty() const { return vec_ ? vec_->empty() : mat_.empty(); }
 protected:
  Mat mat_;
  const std::vector<Mat>* vec_;
};
class _OutputArray {
 public:
  _OutputArray() : pm_(nullptr), pv_(nullptr), fixed_(false) {}
  _OutputArray(const _NoArrayTag&) : pm_(nullptr), pv_(nullptr), fixed_(false) {}
  _OutputArray(Mat& m) : pm_(&m), pv_(nullptr), fixed_(false) {}
  _OutputArray(const Mat& m) : hdr_(m), pm_(&hdr_), pv_(nullptr), fixed_(true) {}   // a header by value: data is written through
  template <typename T> _OutputArray(Mat_<T>& m) : pm_(&m), pv_(nullptr), fixed_(false) {}
  template <typename T, int m, int n> _OutputArray(Matx<T, m, n>& M) : hdr_(m, n, DataType<T>::type, (void*)M.val), pm_(&hdr_), pv_(nullptr), fixed_(true) {}
  template <typename T, int n> _OutputArray(Vec<T, n>& v) : hdr_(n, 1, DataType<T>::type, (void*)v.val), pm_(&hdr_), pv_(nullptr), fixed_(true) {}
  _OutputArray(std::vector<Mat>& v) : pm_(nullptr), pv_(&v), fixed_(false) {}
  _OutputArray(const _OutputArray& o) : hdr_(o.hdr_), pm_(o.fixed_ ? &hdr_ : o.pm_), pv_(o.pv_), fixed_(o.fixed_) {}
  bool needed() const { return pm_ || pv_; }
  bool fixedSize() const { return fixed_; }
  void create(int r, int c, int type) const {
    if (pv_) { pv_->resize((size_t)r * c); return; }   // create(1, n, CV_8U) on a vector<Mat>: n empty matrices (linemod.cpp:1361)
    CV_Assert(pm_ != nullptr);
    if (fixed_) { CV_Assert(pm_->rows == r && pm_->cols == c && pm_->type() == (type & 0xfff)); return; }
    pm_->create(r, c, type);
  }
  void create(Size s, int type) const { create(s.height, s.width, type); }
  Mat getMat(int = -1) const { return pm_ ? *pm_ : Mat(); }
  Mat& getMatRef(int i = -1) const { if (i < 0) { CV_Assert(pm_); return *pm_; } CV_Assert(pv_ && (size_t)i < pv_->size()); return (*pv_)[i]; }
  std::vector<Mat>* getVecPtr() const { return pv_; }
  bool empty() const { return pv_ ? pv_->empty() : (!pm_ || pm_->empty()); }
 protected:
  mutable Mat hdr_;
  Mat* pm_;
  std::vector<Mat>* pv_;
  bool fixed_;
};
inline _NoArrayTag noArray() { return _NoArrayTag(); }

// ---------------------------------------------------------------- persistence (FileStorage): a small reader / writer of the
// YAML dialect cv::FileStorage emits, enough for Detector::read / readClass / write / writeClass (linemod.cpp:1681-1786)
class FileNode;
class FileNodeIterator;
struct FsNode {
  enum { NONE = 0, SCALAR = 1, SEQ = 2, MAP = 3 };
  int kind; bool is_string; bool flow;
  std::string scalar;
  std::vector<std::shared_ptr<FsNode> > seq;
  std::vector<std::pair<std::string, std::shared_ptr<FsNode> > > map;
  FsNode() : kind(NONE), is_string(false), flow(false) {}
};
class FileNode {
 public:
  FileNode() {}
  explicit FileNode(const std::shared_ptr<FsNode>& n) : n_(n) {}
  FileNode operator[](const char* key) const;
  FileNode operator[](const String& key) const { return (*this)[key.c_str()]; }
  size_t size() const { return !n_ ? 0 : n_->kind == FsNode::SEQ ? n_->seq.size() : n_->kind == FsNode::MAP ? n_->map.size() : n_->kind == FsNode::SCALAR ? 1 : 0; }
  bool empty() const { return !n_ || n_->kind == FsNode::NONE; }
  FileNodeIterator begin() const;
  FileNodeIterator end() const;
  operator int() const;
  operator float() const;
  operator double() const;
  operator String() const;
  std::shared_ptr<FsNode> n_;
};
class FileNodeIterator {
 public:
  FileNodeIterator() : i_(0) {}
  FileNodeIterator(const std::shared_ptr<FsNode>& n, size_t i) : n_(n), i_(i) {}
  FileNode operator*() const;
  FileNodeIterator& operator++() { ++i_; return *this; }
  FileNodeIterator operator++(int) { FileNodeIterator t = *this; ++i_; return t; }
  bool operator==(const FileNodeIterator& o) const { return i_ == o.i_; }
  bool operator!=(const FileNodeIterator& o) const { return i_ != o.i_; }
  template <typename T> FileNodeIterator& operator>>(T& v) { v = (T)(**this); ++i_; return *this; }
  std::shared_ptr<FsNode> n_;
  size_t i_;
};
template <typename T> inline void operator>>(const FileNode& fn, std::vector<T>& v) {
  v.clear();
  for (FileNodeIterator it = fn.begin(), e = fn.end(); it != e; ++it) v.push_back((T)(*it));
}
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage(const String& filename, int mode);
  ~FileStorage();
  bool isOpened() const { return opened_; }
  FileNode root() const { return FileNode(root_); }
  FileNode operator[](const char* key) const { return root()[key]; }
  FileNode operator[](const String& key) const { return root()[key.c_str()]; }
  void release();
  // writer state machine (operator<< below)
  void put_string(const std::string& s);
  void put_scalar(const std::string& text, bool is_string);
 private:
  std::string filename_; int mode_; bool opened_;
  std::shared_ptr<FsNode> root_;
  std::vector<std::shared_ptr<FsNode> > stack_;
  std::string pending_key_; bool have_key_;
};
FileStorage& operator<<(FileStorage& fs, const char* s);
FileStorage& operator<<(FileStorage& fs, const String& s);
FileStorage& operator<<(FileStorage& fs, int v);
FileStorage& operator<<(FileStorage& fs, float v);
FileStorage& operator<<(FileStorage& fs, double v);
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const std::vector<T>& v) {
  fs << "[:";
  for (size_t i = 0; i < v.size(); ++i) fs << v[i];
  fs << "]";
  return fs;
}

// ---------------------------------------------------------------- core functions
enum { NORM_L2 = 4 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { DIST_C = 3, DIST_L2 = 2 };

double norm(InputArray a, InputArray b, int normType = NORM_L2);
void add(InputArray a, InputArray b, OutputArray dst, InputArray mask = noArray(), int dtype = -1);
void subtract(InputArray a, InputArray b, OutputArray dst, InputArray mask = noArray(), int dtype = -1);
void bitwise_and(InputArray a, InputArray b, OutputArray dst, InputArray mask = noArray());
int countNonZero(InputArray a);
bool checkRange(InputArray a, bool quiet = true, Point* pos = 0, double minVal = -DBL_MAX, double maxVal = DBL_MAX);
void merge(const std::vector<Mat>& mv, OutputArray dst);
void split(const Mat& src, std::vector<Mat>& mv);
void phase(InputArray x, InputArray y, OutputArray angle, bool angleInDegrees = false);
void Rodrigues(InputArray src, OutputArray dst);
class SVD {
 public:
  enum { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
  static void compute(InputArray src, OutputArray w, OutputArray u, OutputArray vt, int flags = 0);
};

// ---------------------------------------------------------------- imgproc
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void Sobel(InputArray src, OutputArray dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
void medianBlur(InputArray src, OutputArray dst, int ksize);
void pyrDown(InputArray src, OutputArray dst, const Size& dstsize = Size(), int borderType = BORDER_DEFAULT);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void erode(InputArray src, OutputArray dst, InputArray kernel, Point anchor = Point(-1, -1), int iterations = 1, int borderType = BORDER_CONSTANT);
void distanceTransform(InputArray src, OutputArray dst, int distanceType, int maskSize, int dstType = CV_32F);
void circle(Mat& img, Point center, int radius, const Scalar& color, int thickness = 1);
void rectangle(Mat& img, Rect r, const Scalar& color, int thickness = 1);

// ---------------------------------------------------------------- highgui (viewers: no-ops)
inline void namedWindow(const String&, int = 1) {}
inline void imshow(const String&, InputArray) {}
inline int waitKey(int = 0) { return -1; }

}  // namespace cv

// ---------------------------------------------------------------- cvflann (the single KD-tree 1-NN the ICP uses)
namespace cvflann {
template <typename T> class Matrix {
 public:
  size_t rows, cols, stride; T* data;
  Matrix() : rows(0), cols(0), stride(0), data(nullptr) {}
  Matrix(T* d, size_t r, size_t c, size_t s = 0) : rows(r), cols(c), stride(s ? s : c), data(d) {}
  T* operator[](size_t i) const { return data + i * stride; }
};
template <typename T> struct L2_Simple { typedef T ElementType; typedef float ResultType; };
template <typename T> struct L2 { typedef T ElementType; typedef float ResultType; };
struct KDTreeSingleIndexParams { int leaf_max_size; explicit KDTreeSingleIndexParams(int l = 10, bool = true, int = -1) : leaf_max_size(l) {} };
struct SearchParams { int checks; float eps; bool sorted; explicit SearchParams(int c = 32, float e = 0, bool s = true) : checks(c), eps(e), sorted(s) {} };
struct KdImpl;
KdImpl* kd_build(const float* pts, size_t n, size_t stride, int leaf);
void kd_free(KdImpl*);
void kd_knn1(const KdImpl*, const float* q, size_t nq, size_t qstride, int* idx, size_t istride, float* dist, size_t dstride);
template <typename Distance> class Index {
 public:
  Index(const Matrix<float>& pts, const KDTreeSingleIndexParams& p) : pts_(pts), leaf_(p.leaf_max_size), impl_(nullptr) {}
  ~Index() { if (impl_) kd_free(impl_); }
  void buildIndex() { if (impl_) kd_free(impl_); impl_ = kd_build(pts_.data, pts_.rows, pts_.stride, leaf_); }
  void knnSearch(const Matrix<float>& q, Matrix<int>& indices, Matrix<float>& dists, int knn, const SearchParams&) {
    if (knn != 1 || !impl_) throw cv::Exception(cv::Error::StsNotImplemented, "cvshim: only 1-NN on a built index");
    kd_knn1(impl_, q.data, q.rows, q.stride, indices.data, indices.stride, dists.data, dists.stride);
  }
 private:
  Index(const Index&);
  Index& operator=(const Index&);
  Matrix<float> pts_; int leaf_; KdImpl* impl_;
};
}  // namespace cvflann

// ---------------------------------------------------------------- primitive hooks (tests plug the real cv2 in here)
namespace cvshim {
enum Op { OP_GAUSSIAN7 = 0, OP_SOBEL_DX = 1, OP_SOBEL_DY = 2, OP_PHASE_DEG = 3, OP_MEDIAN5 = 4, OP_PYRDOWN = 5, OP_RESIZE_NN = 6,
          OP_SVD3 = 7, OP_KNN1 = 8, OP_COUNT = 9 };
// a, b: inputs; out: output; dims: op-specific (rows, cols, channels, out rows, out cols, ...)
typedef void (*hook_fn)(int op, const void* a, const void* b, void* out, const int* dims);
void set_hook(int op, hook_fn f);
hook_fn get_hook(int op);
}  // namespace cvshim

#endif  // FL_REF_CVSHIM_HPP
