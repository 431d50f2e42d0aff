// oracle/thc_compat: the atomicAdd overloads of THCAtomics.cuh live on in ATen/cuda/Atomic.cuh
#pragma once
#include <ATen/cuda/Atomic.cuh>
