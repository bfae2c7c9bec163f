// CorrIndex of every pair of decompositions (c2c_metrics.cu): launcher visible to the C ABI translation unit.
#pragma once
#include "c2c_common.cuh"

size_t c2c_corrindex_smem_bytes(int R);
// method: 0 stacked / single block, 1 max over the blocks, 2 min, 3 mean.  *status is set to 1 when a column norm is zero.
cudaError_t c2c_launch_corrindex_pairs(const double* F, int n, long long S, int R, const long long* seg_off, int nseg, int method,
                                       double tol, double* out, int* status, cudaStream_t st, int* launches);
