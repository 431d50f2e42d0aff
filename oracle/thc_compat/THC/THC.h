// oracle/thc_compat -- THC was removed from PyTorch (>= 1.11).  The reference's CUDA sources use four of its names; they are
// mapped onto their c10 successors here so that the sources compile UNMODIFIED (oracle/ref_cuda/*.cu, test/baseline only).
#pragma once
#include <c10/cuda/CUDACachingAllocator.h>
#include <c10/cuda/CUDAException.h>
struct THCState;
#define THCudaCheck(x) C10_CUDA_CHECK(x)
static inline void* THCudaMalloc(THCState*, size_t n) { return c10::cuda::CUDACachingAllocator::raw_alloc(n); }
static inline void THCudaFree(THCState*, void* p) { c10::cuda::CUDACachingAllocator::raw_delete(p); }
