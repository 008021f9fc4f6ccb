// stand-in (TEST INFRASTRUCTURE): boost::shared_ptr as std::shared_ptr
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; using std::enable_shared_from_this; using std::static_pointer_cast; using std::dynamic_pointer_cast; }
