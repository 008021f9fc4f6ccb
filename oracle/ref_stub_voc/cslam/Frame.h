// Stand-ins for cslam::Frame / KeyFrame as shim/ORBVocabulary_shim.cpp sees them (TEST INFRASTRUCTURE): the members ComputeBoW touches,
// with the reference's names and types (cslam/include/cslam/Frame.h:77,106,134-135; KeyFrame.h:145,330-331).  cslam/ORBVocabulary.h and
// the DBoW2 headers it includes are the reference's own.
#ifndef CCM_REF_STUB_VOC_CSLAM_H
#define CCM_REF_STUB_VOC_CSLAM_H
#include <opencv2/core/core.hpp>

#include <cslam/ORBVocabulary.h>
#include <cslam/config.h>
#include <cslam/estd.h>
#include <thirdparty/DBoW2/DBoW2/BowVector.h>
#include <thirdparty/DBoW2/DBoW2/FeatureVector.h>

namespace cslam {
class Frame {
 public:
  void ComputeBoW();
  vocptr mpORBvocabulary;
  cv::Mat mDescriptors;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
};
class KeyFrame {
 public:
  void ComputeBoW();
  vocptr mpORBvocabulary;
  cv::Mat mDescriptors;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
};
}  // namespace cslam
#endif
