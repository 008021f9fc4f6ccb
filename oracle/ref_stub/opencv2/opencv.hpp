// Stand-in for <opencv2/opencv.hpp> (TEST INFRASTRUCTURE, NOT PRODUCT): core types + the image primitives ORBextractor.cpp calls.
// The primitives are DECLARED here and DEFINED in oracle/ref_orb_wrap.cpp on top of the oracle's restatements (orc_fast,
// orc_resize_linear_u8, orc_gaussian_blur7, orc_fast_atan2), themselves pinned to cv2 4.13 by tests/test_oracle_orb.py.
#ifndef CCM_ORACLE_REF_STUB_OPENCV_HPP
#define CCM_ORACLE_REF_STUB_OPENCV_HPP
#include "core/core.hpp"

namespace cv {
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16, INTER_LINEAR = 1 };
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType);
float fastAtan2(float y, float x);
struct KeyPointsFilter {   // only ComputeKeyPointsOld (never called, ORBextractor.cpp:1230) uses it
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};
}  // namespace cv
#endif
