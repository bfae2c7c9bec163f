// ADMM mode update of constrained_parafac (c2c_admm.cu): launcher visible to the C ABI translation unit.
#pragma once
#include "c2c_common.cuh"

// constraint: 0 none, 1 non_negative, 2 l1_reg (param), 3 l2_square_reg (param).  out: device double [2].
cudaError_t c2c_launch_admm_update(float* A, double* X64, double* dual, double* aux, const float* M, const float* G, long long I,
                                   int R, int ldr, int constraint, double param, int n_inner, double tol, double* out,
                                   cudaStream_t st);
