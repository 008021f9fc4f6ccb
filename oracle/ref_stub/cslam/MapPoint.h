// stand-in (TEST INFRASTRUCTURE): see Frame.h next to this file
#include <cslam/Frame.h>
