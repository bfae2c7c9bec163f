// One-CTA-per-run factorization of small (L2-resident) tensors (c2c_multi.cu): launcher visible to the C ABI translation unit.
#pragma once
#include "c2c_common.cuh"

struct MultiRun {                 // one per run, device array
    float* A[C2C_MAX_DIMS];       // the run's factors [I_n x ldr], initial values in, result out
    float* T;                     // [P x ldr] scratch: plane contraction, then the Khatri-Rao rows of the leading factors
    double* errors;               // [n_iter_max]
    int* state;                   // [2] = {iterations done, converged}
    double* sums;                 // [5] directly accumulated sums of the final factors (may be null)
    int rank, ldr;
};

struct MultiArgs {
    const float* X;
    int ndim; int shape[C2C_MAX_DIMS];
    long long P; int E, I2, I3;
    const MultiRun* runs; int nruns;
    int n_iter_max; double tol; int cvg;
    const double* normX2;
    float eps;
};

size_t c2c_multi_smem_bytes(int ndim, int ldr, long long E, int I2, int I3);
cudaError_t c2c_launch_ncp_multi(const MultiArgs& a, size_t smem, cudaStream_t st);
