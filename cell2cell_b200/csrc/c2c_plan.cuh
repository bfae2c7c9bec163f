// Mask plan of the masked factorization (c2c_plan.cu): built once per (tensor, mask), used by every iteration of every
// factorization of that tensor.  Launchers visible to the C ABI translation unit.
//
// tensorly re-imputes the missing entries with the current reconstruction before every mode's MTTKRP (mirrored in the
// reference at cell2cell/tensor/coupled_factorization.py:342-345).  With zeros at the missing entries the MTTKRP is the
// contraction of the tensor as stored (the unmasked streaming passes) plus the contraction of the reconstruction over the
// MISSING set.  A how='outer' tensor (cell2cell/tensor/tensor.py:1083-1145) misses whole (context, LR pair) planes -- a
// partner gene absent from the context -- and whole rows / columns of every plane of a context -- a cell type absent from
// it.  Its mask is separable:
//
//     observed[p, s, v] = g[p] * hs[c(p), s] * hv[c(p), v]            c(p) = index of the first mode of plane p
//
// and over such a set the sum of reconstructions collapses to products of small (R x R) restricted Grams: O((P + C0 E) R^2)
// flops per mode instead of O(N_missing R), and no pass over the mask at all.  The plan holds the tightest separable model
// of ANY mask (g / hs / hv = "some entry observed": it never calls an observed entry missing) plus the RESIDUAL -- entries
// the model calls observed but the mask does not (empty for cell2cell's own masks, everything for an unstructured mask) --
// as two lists of the same form: by plane (in-plane positions, for the leading modes) and by position within chunks of
// C2C_PLAN_CHUNK planes (plane numbers within the chunk, for the trailing modes).  Both contractions are then the SAME
// row-wise sum  out[row, :] = sum over the row's entries j of <own[row], tab[j]> tab[j, :]  -- own = W[p], tab = B for a plane
// row; own = B[e], tab = the chunk's W rows for a position row -- with `tab` staged in shared memory (corr_rows_kernel).
#pragma once
#include "c2c_common.cuh"

#define C2C_PLAN_CHUNK 1024           // planes per chunk of the position-major residual lists (their W rows fit shared memory)
#define C2C_PLAN_MAX_E 65535          // in-plane positions are stored as uint16
#define C2C_PLAN_NCOUNTERS 8

struct PlanView {
    // header (sized from the shape alone)
    unsigned long long* counters;     // [0] missing entries, [1] residual entries, [2] non-zero value under a zero mask byte,
                                      // [3] slots of the position-major lists, [4] model has missing entries, [5] overflow,
                                      // [6] residual entries without the class padding of the by-plane list ([1] counts its slots)
    uint8_t* g;                       // [P]        plane has an observed entry
    int* hs;                          // [C0 x I2]  row s of the planes of context c has an observed entry
    int* hv;                          // [C0 x I3]  column v ...
    int* row_off;                     // [P + 1]    list slots by plane (8 x the longest position class of the plane)
    int* pos_off;                     // [nchunks * E + 1]  list slots by (chunk, position) (8 x the longest plane class)
    // lists (sized by the scan)
    uint16_t* csr_ent;                // [slots]    in-plane position, plane by plane; slot i holds a position of class i mod 8
                                      //            (increasing within a class), 0xFFFF = padding
    uint16_t* pos_ent;                // [slots]    plane number within its chunk, (chunk, position) by (chunk, position);
                                      //            slot i holds a plane of class i mod 8, 0xFFFF = padding
    long long P; long long ppc0;      // planes; planes per index of the first mode
    int E, I2, I3, C0, nchunks;
};

size_t c2c_plan_header_bytes(long long P, int E, int I2, int I3, int C0);
size_t c2c_plan_lists_bytes(long long n_resid, long long pos_slots);
// carve the caller's buffers (lists may be null before the scan)
void c2c_plan_view(PlanView& v, long long P, int E, int I2, int I3, int C0, void* header, void* lists, long long n_resid);

// phase 1: separable model, residual counts and list offsets into the header (X may be null: no zero-fill check)
cudaError_t c2c_plan_scan(const float* X, const uint8_t* mask, const PlanView& v, int nsm, cudaStream_t st);
// phase 2: the two residual lists
cudaError_t c2c_plan_fill(const uint8_t* mask, const PlanView& v, int nsm, cudaStream_t st);

// ---- corrections over the model-missing set (Gram form) ---------------------------------------------------------------
// Dmat[(C0 + 1)][ldr * ldr]: D_c = sum over the (s, v) pairs the model misses in context c of B[s,v] (x) B[s,v]; row C0 =
// the same sum over every pair (for planes with g = 0)
cudaError_t c2c_launch_trail_ctx_gram(const PlanView& v, const float* A2, const float* A3, int R, int ldr, float* Dmat,
                                      const int* done, cudaStream_t st, int pdl);
// Tw[p, :] = T0[p, :] + W[p, :] * (g[p] ? D_c(p) : D_all)
cudaError_t c2c_launch_plane_corr_gram(const PlanView& v, const LeadInfo& lead, const float* Dmat, const float* T0, float* Tw,
                                       int R, int ldr, const int* done, int nsm, cudaStream_t st, int pdl);
// Ymat[(C0 + 1)][ldr * ldr]: Y_c = sum over the planes of context c with g = 1 of W[p] (x) W[p]; row C0 = the same sum over
// every plane with g = 0.  Ypart: scratch [C0][parts][2][ldr * ldr], parts = c2c_lead_ctx_parts(C0, nsm)
int c2c_lead_ctx_parts(int C0, int nsm);
cudaError_t c2c_launch_lead_ctx_gram(const PlanView& v, const LeadInfo& lead, int R, int ldr, float* Ypart, float* Ymat,
                                     const int* done, int nsm, cudaStream_t st, int pdl);
// Uw[e, :] = U0[e, :] + B[e, :] * (Y_missing + sum over the contexts whose row s or column v the model misses of Y_c)
cudaError_t c2c_launch_fiber_corr_gram(const PlanView& v, const float* A2, const float* A3, const float* Ymat, const float* U0,
                                       float* Uw, int R, int ldr, const int* done, cudaStream_t st, int pdl);

// ---- corrections over the residual lists ----------------------------------------------------------------------------------
// Wmat[p, :] = Khatri-Rao row of the leading factors of plane p ([P][ldr]; recomputed whenever a leading factor changed)
cudaError_t c2c_launch_wmat(const LeadInfo& lead, long long P, int ldr, float* Wmat, const int* done, cudaStream_t st);
// Tout[p, :] = Tin[p, :] + sum over the residual entries e of plane p of <W[p], B[e]> B[e, :]   (Tin == Tout: in place)
cudaError_t c2c_launch_plane_corr_list(const PlanView& v, const float* Wmat, const float* Bmat, const float* Tin, float* Tout,
                                       int ldr, const int* done, int nsm, cudaStream_t st, int pdl);
// Uout[e, :] = Uin[e, :] + sum over the residual planes p at position e of <W[p], B[e]> W[p, :];  part: scratch
// [max_parts][E * ldr]
cudaError_t c2c_launch_fiber_corr_list(const PlanView& v, const float* Wmat, const float* Bmat, const float* Uin, float* Uout,
                                       float* part, int max_parts, int ldr, const int* done, int nsm, cudaStream_t st, int pdl);
