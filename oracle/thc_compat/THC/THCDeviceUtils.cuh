// oracle/thc_compat: THCCeilDiv (THCDeviceUtils.cuh)
#pragma once
template <typename T>
__host__ __device__ __forceinline__ T THCCeilDiv(T a, T b) { return (a + b - 1) / b; }
