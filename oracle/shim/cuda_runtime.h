// TEST INFRASTRUCTURE (oracle). Stub of the CUDA runtime surface named by the reference's
// src/std_cuda_tensor.hpp:15-70.  The oracle is a CPU build: "device" memory is host memory, so the
// constructor's cudaMalloc (src/post-process.h:149-150) works without a GPU.  This directory must
// only ever be on the include path of the oracle's host compiler, never nvcc's.
#pragma once
#include <cstdlib>
#include <cstring>
typedef int cudaError_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
inline cudaError_t cudaMalloc(void **p, std::size_t n) { *p = std::malloc(n); return 0; }
template <typename T> inline cudaError_t cudaMalloc(T **p, std::size_t n) { *p = (T *)std::malloc(n); return 0; }
inline cudaError_t cudaFree(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, std::size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
