// fealess_b200/cv_min.hpp - the handful of OpenCV core types that appear in the reference's call surface for this path
// (linemod/linemod.hpp, linemod/linemod_if.h, ICP/detection.h, ICP/ICP.h, ICP/NMS.h, ICP/depth_to_3d.h), for hosts that
// build WITHOUT the OpenCV C++ SDK (this repository's image has none, so the C++ mirror is compiled and tested against
// this stand-in).  A FEALESS build defines FEALESS_B200_WITH_OPENCV and gets the real <opencv2/core.hpp> instead; the
// mirror headers only use the members declared here, which are source compatible with OpenCV 3.x / 4.x.
#ifndef FEALESS_B200_CV_MIN_HPP
#define FEALESS_B200_CV_MIN_HPP

#ifdef FEALESS_B200_WITH_OPENCV
#include <opencv2/core.hpp>
#else

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)

namespace cv {

typedef std::string String;
template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& what_arg, int code_ = -215) : std::runtime_error(what_arg), code(code_) {}
  int code;   // -215 = cv::Error::StsAssert
};

template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
};
typedef Rect_<int> Rect;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Point { int x, y; Point() : x(0), y(0) {} Point(int x_, int y_) : x(x_), y(y_) {} };

template <typename T, int N> struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
  Vec(T a, T b, T c) { static_assert(N == 3, "3-element constructor"); val[0] = a; val[1] = b; val[2] = c; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 3> Vec3f;

template <typename T, int R, int C> struct Matx {
  T val[R * C];
  Matx() { for (int i = 0; i < R * C; ++i) val[i] = T(0); }
  static Matx eye() { Matx m; for (int i = 0; i < (R < C ? R : C); ++i) m.val[i * C + i] = T(1); return m; }
  T& operator()(int r, int c) { return val[r * C + c]; }
  const T& operator()(int r, int c) const { return val[r * C + c]; }
};
typedef Matx<float, 3, 3> Matx33f;

// dense 2-D matrix header with shared ownership of its buffer (or none, when wrapping user memory)
class Mat {
 public:
  int rows, cols;
  uchar* data;
  size_t step;   // bytes per row
  Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
  Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0), type_(0) { create(r, c, type); }
  Mat(int r, int c, int type, void* user, size_t step_ = 0) : rows(r), cols(c), data(static_cast<uchar*>(user)), step(step_), type_(type) {
    if (!step) step = (size_t)c * elemSize();
  }
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && data && owner_) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
    owner_ = std::make_shared<std::vector<uchar>>((size_t)r * step);
    data = owner_->data();
  }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); if (m.data) std::memset(m.data, 0, (size_t)r * m.step); return m; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
  size_t elemSize() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)sz[depth()] * channels(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step == (size_t)cols * elemSize(); }
  Size size() const { return Size(cols, rows); }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  Mat clone() const {
    Mat m; if (empty()) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
    return m;
  }
 private:
  int type_;
  std::shared_ptr<std::vector<uchar>> owner_;
};

// the proxy types of the reference's signatures, reduced to what this path passes through them
struct _NoArray {};
inline _NoArray noArray() { return _NoArray(); }
class _InputArray {
 public:
  _InputArray() : mat_(nullptr) {}
  _InputArray(const _NoArray&) : mat_(nullptr) {}
  _InputArray(const Mat& m) : mat_(&m) {}
  template <typename T, int R, int C> _InputArray(const Matx<T, R, C>& m) : mat_(nullptr), matx_(R, C, CV_MAKETYPE(CV_32F, 1), const_cast<T*>(m.val)) { mat_ = &matx_; }
  bool empty() const { return !mat_ || mat_->empty(); }
  const Mat& getMat() const { static const Mat none; return mat_ ? *mat_ : none; }
 private:
  const Mat* mat_;
  Mat matx_;
};
class _OutputArray {
 public:
  _OutputArray() : mat_(nullptr), vec_(nullptr) {}
  _OutputArray(const _NoArray&) : mat_(nullptr), vec_(nullptr) {}
  _OutputArray(Mat& m) : mat_(&m), vec_(nullptr) {}
  _OutputArray(std::vector<Mat>& v) : mat_(nullptr), vec_(&v) {}
  bool needed() const { return mat_ || vec_; }
  Mat* mat() const { return mat_; }
  std::vector<Mat>* vec() const { return vec_; }
 private:
  Mat* mat_;
  std::vector<Mat>* vec_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& OutputArrayOfArrays;

}  // namespace cv

#endif  // FEALESS_B200_WITH_OPENCV
#endif  // FEALESS_B200_CV_MIN_HPP
