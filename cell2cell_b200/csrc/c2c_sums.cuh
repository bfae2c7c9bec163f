// Streaming recon_sums (c2c_sums.cu): launcher visible to the C ABI translation unit.
#pragma once
#include "c2c_common.cuh"

// a.sums receives [grid][5] partial sums; *grid_out = CTAs launched.  cudaErrorNotSupported: shape or
// rank outside this kernel (masked, row length not a multiple of 4, more float4 positions per plane than threads, R > 32).
cudaError_t c2c_launch_sums_stream(const ContractArgs& a, int grid_max, cudaStream_t st, int* grid_out);
