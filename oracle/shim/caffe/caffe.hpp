// oracle/shim/caffe/caffe.hpp -- compile-only stand-in (BVLC Caffe is not in this image).
// TEST INFRASTRUCTURE (see oracle/__init__.py).
#pragma once
#include <memory>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
namespace caffe { template <class T> class Net {}; }
