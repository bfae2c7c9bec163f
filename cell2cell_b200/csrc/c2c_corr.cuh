// Missing-entry corrections of the masked factorization (c2c_corr.cu): launchers visible to the C ABI translation unit.
#pragma once
#include "c2c_common.cuh"

#define C2C_CORR_SMEM_MAX (200 * 1024)
#define C2C_CORR_MAX_PARTS 320           // CTAs (= partials) of fiber_corr

// *flag |= 1 if some entry with mask == 0 holds a non-zero value (the dense + correction form needs zeros there)
cudaError_t c2c_launch_masked_nonzero(const float* X, const uint8_t* mask, long long n, int* flag, int nsm, cudaStream_t st);
// shared memory fiber_corr needs for this in-plane size / factor pitch (eligible when <= C2C_CORR_SMEM_MAX)
size_t c2c_fiber_corr_smem(long long E, int ldr);
// Tw[p, :] = T0[p, :] + sum_{e: mask[p, e] == 0} <W[p, :], B[e, :]> B[e, :]
cudaError_t c2c_launch_plane_corr(const uint8_t* mask, const LeadInfo& lead, const float* Bmat, const float* T0, float* Tw,
                                  long long P, long long E, int R, int ldr, const int* done, int nsm, cudaStream_t st);
// Uw[e, :] = U0[e, :] + sum_{p: mask[p, e] == 0} <W[p, :], B[e, :]> W[p, :];  part: [max_parts][E * ldr] scratch
cudaError_t c2c_launch_fiber_corr(const uint8_t* mask, const LeadInfo& lead, const float* Bmat, const float* U0, float* Uw,
                                  float* part, int max_parts, long long P, long long E, int R, int ldr, const int* done,
                                  cudaStream_t st);
