// ref_extractor_shim_wrap.cpp — runs shim/ORBextractor_shim.cpp through the reference's own cslam::ORBextractor interface
// (TEST INFRASTRUCTURE, NOT PRODUCT): constructor, operator(), the getters and the public pyramid.
#include <cslam/ORBextractor.h>

#include <cstdint>
#include <cstring>

using namespace cslam;

extern "C" {
struct xs_keypoint { float x, y, size, angle, response; int32_t octave; };

/* returns the number of keypoints; tables = 4 * nlevels floats (scale factor, inverse, sigma2, inverse sigma2);
 * pyramid = the levels back to back, level_wh = 2 per level */
int xshim_extract(const uint8_t* img, int w, int h, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, xs_keypoint* kps,
                  int max_kp, uint8_t* desc, float* tables, uint8_t* pyramid, int32_t* level_wh, int32_t* getters /*2: GetLevels, n*/) {
  try {
    ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
    cv::Mat im(h, w, CV_8U);
    for (int r = 0; r < h; r++) std::memcpy(im.ptr<uchar>(r), img + (size_t)r * w, w);
    std::vector<cv::KeyPoint> k;
    cv::Mat d, mask;
    ex(im, mask, k, d);
    ex(im, mask, k, d);                               // a second call re-uses the cached handle and must give the same
    const int n = (int)k.size() < max_kp ? (int)k.size() : max_kp;
    for (int i = 0; i < n; i++) kps[i] = xs_keypoint{k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave};
    if (n) std::memcpy(desc, d.ptr<uchar>(0), 32 * (size_t)n);
    const std::vector<float> sf = ex.GetScaleFactors(), isf = ex.GetInverseScaleFactors(), s2 = ex.GetScaleSigmaSquares(), is2 = ex.GetInverseScaleSigmaSquares();
    for (int l = 0; l < nlevels; l++) { tables[4 * l] = sf[l]; tables[4 * l + 1] = isf[l]; tables[4 * l + 2] = s2[l]; tables[4 * l + 3] = is2[l]; }
    size_t pos = 0;
    for (int l = 0; l < nlevels; l++) {
      const cv::Mat& P = ex.mvImagePyramid[l];
      level_wh[2 * l] = P.cols; level_wh[2 * l + 1] = P.rows;
      for (int r = 0; r < P.rows; r++) { std::memcpy(pyramid + pos, P.ptr<uchar>(r), P.cols); pos += P.cols; }
    }
    getters[0] = ex.GetLevels(); getters[1] = (int)k.size();
    return n;
  } catch (...) { return -1; }
}
}
