// stand-in: cslam/Converter.h names this header; nothing of it is used
#include <opencv2/core/core.hpp>
