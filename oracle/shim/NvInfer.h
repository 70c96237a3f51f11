// oracle/shim/NvInfer.h -- compile-only stand-in for TensorRT 5 (not in this image).
// TEST INFRASTRUCTURE (see oracle/__init__.py).  Declares only the names the reference's
// tensorrt/trtnetbase.h and trtretinafacenet.h mention.
#pragma once
#include <cstddef>
namespace nvinfer1 {
class IRuntime; class ICudaEngine; class IExecutionContext; class IHostMemory;
class ILogger { public: enum class Severity { kINTERNAL_ERROR, kERROR, kWARNING, kINFO };
  virtual void log(Severity, const char *) = 0; virtual ~ILogger() {} };
class IProfiler { public: virtual void reportLayerTime(const char *, float) = 0; virtual ~IProfiler() {} };
class DimsCHW { int d_[3];
 public: DimsCHW() : d_{0, 0, 0} {} DimsCHW(int c, int h, int w) : d_{c, h, w} {}
  int c() const { return d_[0]; } int h() const { return d_[1]; } int w() const { return d_[2]; } };
}  // namespace nvinfer1
