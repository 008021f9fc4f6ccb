// stand-in (TEST INFRASTRUCTURE): cslam/estd.h includes PCL but the code compiled here uses none of it
#pragma once
