// oracle/shim/cuda_runtime_api.h -- compile-only stand-in so oracle/_ref needs no CUDA.
// TEST INFRASTRUCTURE (see oracle/__init__.py).  Reached only from the reference's
// detect()/detectBatchImages(), which oracle/_ref never calls.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
typedef struct CUstream_st *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
[[noreturn]] inline void rf_shim_cuda(const char *w) { std::fprintf(stderr, "oracle/shim: %s is a stub\n", w); std::abort(); }
inline cudaError_t cudaMemcpy(void *, const void *, size_t, cudaMemcpyKind) { rf_shim_cuda("cudaMemcpy"); }
inline cudaError_t cudaMemset(void *, int, size_t) { rf_shim_cuda("cudaMemset"); }
inline cudaError_t cudaMalloc(void **, size_t) { rf_shim_cuda("cudaMalloc"); }
template <class T> inline cudaError_t cudaMalloc(T **, size_t) { rf_shim_cuda("cudaMalloc"); }
inline cudaError_t cudaDeviceSynchronize() { rf_shim_cuda("cudaDeviceSynchronize"); }
