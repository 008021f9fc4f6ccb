// Minimal stand-in for OpenCV's core types (TEST INFRASTRUCTURE, NOT PRODUCT; our own code, not OpenCV's).
//
// OpenCV's C++ headers are absent from this image, so the reference cannot be compiled as a whole.  Two of its translation units,
// however, touch OpenCV only through a handful of value types and five image primitives:
//   * the vendored DBoW2 (cslam/thirdparty/DBoW2): cv::Mat as a 1 x 32 byte container, cv::FileStorage in YAML members never called;
//   * cslam/src/ORBextractor.cpp: cv::Mat views (ROI / rowRange / colRange sharing one buffer), KeyPoint / Point / Size / Rect,
//     and FAST / resize / GaussianBlur / copyMakeBorder / fastAtan2 (declared in opencv2/opencv.hpp next to this file and
//     implemented in oracle/ref_orb_wrap.cpp on top of the oracle's OpenCV-primitive restatements, which tests/test_oracle_orb.py pins
//     to cv2 4.13).
// This header provides exactly that surface so that oracle/Makefile can compile those reference sources where they lie under
// /root/reference into oracle/_ref/*.so.  Semantics that the reference relies on and that are easy to get wrong are kept:
//   * a Mat is a (shared buffer, data pointer, rows, cols, step) view; operator()(Rect), rowRange, colRange alias the parent;
//   * assigning Mat::zeros(r, c, t) to a Mat that already has that shape fills it IN PLACE (OpenCV's MatExpr assignment re-uses the
//     destination), which is how computeDescriptors writes through the row view of the output matrix (ORBextractor.cpp:1208);
//   * create() on a Mat that already has the requested shape keeps its storage (resize / copyMakeBorder into an existing view);
//   * cvRound rounds half to even (lrint under the default rounding mode).
// FileStorage is inert but reports "opened" and yields zeros (or what CCM_REF_STUB_CONF names): cslam/config.h reads its parameters through it during static
// initialisation and would exit(-1) otherwise; none of those parameters is used by the code compiled here.
#ifndef CCM_ORACLE_REF_STUB_OPENCV_CORE_HPP
#define CCM_ORACLE_REF_STUB_OPENCV_CORE_HPP
// the real header pulls these in, and the reference leans on that (pow / log, stringstream, ...)
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <unistd.h>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

template <class T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <class T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;

template <class T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <class T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct MatZeros { int rows, cols, type; };

#ifndef CV_MAJOR_VERSION
#define CV_MAJOR_VERSION 2   /* the reference targets OpenCV 2.4 / 3.x; only compared against 3 by shim/ORBextractor_shim.cpp */
#endif

class Mat {
 public:
  int rows, cols;
  size_t step;
  uchar* data;
  Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) {}
  Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) { create(r, c, type); }
  Mat(Size s, int type) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) { create(s.height, s.width, type); }
  Mat(const MatZeros& z) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) { *this = z; }
  static int esz(int type) { return type == CV_32F ? 4 : 1; }
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;   // same shape: the storage (possibly a view) is kept
    rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
    buf_ = std::make_shared<std::vector<uchar> >((size_t)r * step + 64, (uchar)0);
    data = buf_->data();
  }
  static Mat eye(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < (r < c ? r : c); i++) { if (type == CV_32F) m.at<float>(i, i) = 1.0f; else m.at<uchar>(i, i) = 1; }
    return m;
  }
  static MatZeros zeros(int r, int c, int type) { MatZeros z; z.rows = r; z.cols = c; z.type = type; return z; }
  Mat& operator=(const MatZeros& z) {
    create(z.rows, z.cols, z.type);
    for (int r = 0; r < rows; r++) memset(data + (size_t)r * step, 0, (size_t)cols * esz(type_));
    return *this;
  }
  Mat clone() const {
    Mat m;
    if (!data) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; r++) memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz(type_));
    return m;
  }
  void copyTo(Mat& dst) const {   // OpenCV: dst.create(size, type) then a row-wise copy
    dst.create(rows, cols, type_);
    for (int r = 0; r < rows; r++) memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * esz(type_));
  }
  void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  size_t step1() const { return step / esz(type_); }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return step == (size_t)cols * esz(type_); }
  Mat operator()(const Rect& r) const { Mat m(*this); m.data = data + (size_t)r.y * step + (size_t)r.x * esz(type_); m.rows = r.height; m.cols = r.width; return m; }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  uchar* ptr(int row = 0) { return data + (size_t)row * step; }
  const uchar* ptr(int row = 0) const { return data + (size_t)row * step; }
  template <class T> T* ptr(int row = 0) { return reinterpret_cast<T*>(data + (size_t)row * step); }
  template <class T> const T* ptr(int row = 0) const { return reinterpret_cast<const T*>(data + (size_t)row * step); }
  template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
  template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }
  // InputArray / OutputArray surface
  Mat getMat() const { return *this; }
  // small dense float algebra (cslam/src/ORBmatcher.cpp composes poses with it).  Evaluated eagerly; a product with inner dimension
  // 2..4 rounds as cv::gemm's small-matrix path does (f32, left to right), larger ones accumulate in double.  The tests that run the
  // reference's matchers (tests/test_oracle_vs_reference_matchers.py) use poses and points whose arithmetic is exact in float anyway.
  Mat row(int r) const { return (*this)(Rect(0, r, cols, 1)); }
  Mat col(int c) const { return (*this)(Rect(c, 0, 1, rows)); }
  template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <class T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  Mat t() const {
    Mat m(cols, rows, type_);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(c, r) = at<float>(r, c);
    return m;
  }
  double dot(const Mat& o) const {
    double s = 0;
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) s += (double)at<float>(r, c) * (double)o.at<float>(r, c);
    return s;
  }

 private:
  int type_;
  std::shared_ptr<std::vector<uchar> > buf_;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;

inline Mat operator*(const Mat& a, const Mat& b) {
  assert(a.cols == b.rows && a.type() == CV_32F && b.type() == CV_32F);
  Mat m(a.rows, b.cols, CV_32F);
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < b.cols; c++) {
      if (a.cols >= 2 && a.cols <= 4) {   // cv::gemm's small-matrix path: f32 products summed left to right in f32 (checked against cv2 4.13,
        float s = a.at<float>(r, 0) * b.at<float>(0, c);                                  // tests/test_map_update.py::test_gemm_rounding_is_f32_left_to_right)
        for (int k = 1; k < a.cols; k++) s = s + a.at<float>(r, k) * b.at<float>(k, c);
        m.at<float>(r, c) = s;
        continue;
      }
      double s = 0;
      for (int k = 0; k < a.cols; k++) s += (double)a.at<float>(r, k) * (double)b.at<float>(k, c);
      m.at<float>(r, c) = (float)s;
    }
  return m;
}
template <class F> inline Mat mat_map2(const Mat& a, const Mat& b, F f) {
  assert(a.rows == b.rows && a.cols == b.cols);
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = f(a.at<float>(r, c), b.at<float>(r, c));
  return m;
}
template <class F> inline Mat mat_map1(const Mat& a, F f) {
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = f(a.at<float>(r, c));
  return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { return mat_map2(a, b, [](float x, float y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return mat_map2(a, b, [](float x, float y) { return x - y; }); }
inline Mat operator-(const Mat& a) { return mat_map1(a, [](float x) { return -x; }); }
inline Mat operator*(const Mat& a, double s) { return mat_map1(a, [s](float x) { return (float)((double)x * s); }); }
inline Mat operator*(double s, const Mat& a) { return a * s; }
inline Mat operator/(const Mat& a, double s) { return mat_map1(a, [s](float x) { return (float)((double)x / s); }); }
inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }

class FileNode {
 public:
  FileNode() : v_(0.0) {}
  explicit FileNode(double v) : v_(v) {}
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  operator int() const { return (int)v_; }
  operator double() const { return v_; }
  operator std::string() const { return std::string(); }
 private:
  double v_;
};

class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return true; }
  // a test can supply parameter values through the environment: CCM_REF_STUB_CONF="Opt.EssGraphMinFeats=100,Other.Key=2.5"
  FileNode operator[](const std::string& key) const {
    const char* e = getenv("CCM_REF_STUB_CONF");
    if (!e) return FileNode();
    const std::string all(e), pat = key + "=";
    size_t at = 0;
    while ((at = all.find(pat, at)) != std::string::npos) {
      if (at == 0 || all[at - 1] == ',') return FileNode(atof(all.c_str() + at + pat.size()));
      at += pat.size();
    }
    return FileNode();
  }
  FileNode operator[](const char* key) const { return (*this)[std::string(key)]; }
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
#endif
