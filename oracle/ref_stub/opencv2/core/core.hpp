// Minimal stand-in for <opencv2/core/core.hpp> (TEST INFRASTRUCTURE, NOT PRODUCT; our own code, not OpenCV's).
//
// OpenCV's C++ headers are absent from this image, so the reference cannot be compiled as a whole — but its vendored DBoW2
// (cslam/thirdparty/DBoW2) only touches OpenCV through cv::Mat as a 1 x 32 byte container and through cv::FileStorage in its
// YAML save/load members.  This header provides exactly that surface so that oracle/Makefile can compile the reference's OWN
// DBoW2 sources, where they lie under /root/reference, into oracle/_ref/libdbow2_ref.so: the text-file loader, the tree descent,
// FORB::distance and the BowVector / FeatureVector containers that check oracle/bow_oracle.cpp are then the reference's code.
// cv::Mat here is a dense row-major byte buffer (create / zeros / clone / release / ptr<T> / rows / cols); the FileStorage family
// is inert (the YAML path is never exercised: the reference itself loads its vocabulary with loadFromTextFile).
#ifndef CCM_ORACLE_REF_STUB_OPENCV_CORE_HPP
#define CCM_ORACLE_REF_STUB_OPENCV_CORE_HPP
// the real header pulls these in, and DBoW2 leans on that (pow / log, stringstream, ...)
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

class Mat {
 public:
  int rows, cols;
  Mat() : rows(0), cols(0), type_(CV_8U) {}
  Mat(int r, int c, int type) : rows(0), cols(0), type_(CV_8U) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    data_.assign((size_t)r * c * (type == CV_32F ? 4 : 1), 0);
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  Mat clone() const { return *this; }
  void release() { rows = cols = 0; data_.clear(); }
  bool empty() const { return data_.empty(); }
  template <class T> T* ptr(int row = 0) { return reinterpret_cast<T*>(data_.data() + (size_t)row * cols * (type_ == CV_32F ? 4 : 1)); }
  template <class T> const T* ptr(int row = 0) const {
    return reinterpret_cast<const T*>(data_.data() + (size_t)row * cols * (type_ == CV_32F ? 4 : 1));
  }

 private:
  int type_;
  std::vector<unsigned char> data_;
};

class FileNode {
 public:
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  operator int() const { return 0; }
  operator double() const { return 0.0; }
  operator std::string() const { return std::string(); }
};

class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
#endif
