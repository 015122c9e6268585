// Internal prototypes: one host launcher per op, implemented next to its kernels.
#pragma once
#include "common.cuh"

namespace scb {
int gemm(const scb_gemm_args& a, cudaStream_t stream);
}  // namespace scb
