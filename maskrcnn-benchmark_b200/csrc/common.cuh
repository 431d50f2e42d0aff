// common.cuh -- shared helpers for the sm_100a kernels of libmrb_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mrb_b200.h"

#define MRB_API extern "C" __attribute__((visibility("default")))

#define MRB_CUDA_TRY(expr)                      \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

#define MRB_LAUNCH_CHECK()                      \
  do {                                          \
    cudaError_t _e = cudaGetLastError();        \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

namespace mrb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Grid size for a grid-stride elementwise kernel: enough CTAs for `work` items but no more
// than `waves` full waves of the chip at `ctas_per_sm` residency.
static inline int grid_for(int64_t work, int block, int ctas_per_sm, int waves = 4) {
  int64_t need = (work + block - 1) / block;
  int64_t cap = (int64_t)kNumSMs * ctas_per_sm * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace mrb
