// stand-in for the generated ROS message ccmslam_msgs/CvKeyPoint (cslam_msgs/msg/CvKeyPoint.msg): the fields cslam/Converter.h touches
#ifndef CCM_REF_STUB_CVKEYPOINT_MSG
#define CCM_REF_STUB_CVKEYPOINT_MSG
#include <cstdint>
namespace ccmslam_msgs {
struct CvKeyPoint { float fPoint2f_x, fPoint2f_y, angle; int8_t octave; uint8_t response, size; };
struct KF; struct KFred; struct MP; struct MPred; struct Map;
}
#endif
