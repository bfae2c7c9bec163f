/* c2c_b200.h -- C ABI of libc2c_b200.so: the B200 (sm_100a) hot path of Tensor-cell2cell.
 *
 * The reference (earmingol/cell2cell v0.8.1) is pure Python and has no FFI of its own; the boundary this library
 * replaces is the set of Python/tensorly calls named beside each entry point (file:line relative to the reference
 * repository).  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its comment says "host"; buffers are caller-owned, nothing is retained
 *     after return, the library allocates no persistent device memory: scratch comes in as (ws, ws_bytes), sized by the
 *     matching *_workspace_bytes() query;
 *   - tensors are C-ordered fp32; the mask is uint8 (1 = observed, 0 = missing); factor matrices are [I_n x ldr]
 *     row-major fp32 with ldr = rank rounded up to a multiple of 4 and the padding columns zero;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, and unless stated otherwise the call
 *     returns without synchronising;
 *   - return value: 0 = OK, < 0 = argument error, > 0 = cudaError_t; c2c_last_error() describes the last failure of
 *     the calling thread.  No entry point falls back to the CPU.
 */
#ifndef C2C_B200_H
#define C2C_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2C_MAX_NDIM 8
#define C2C_MAX_RANK 128

/* communication scores: cell2cell/core/communication_scores.py:195-203 */
#define C2C_SCORE_PRODUCT 0
#define C2C_SCORE_MEAN 1
#define C2C_SCORE_GMEAN 2
/* complex aggregation: cell2cell/preprocessing/rnaseq.py:185-192; group aggregation: communication_scores.py:235-242 */
#define C2C_AGG_MIN 0
#define C2C_AGG_MEAN 1
#define C2C_AGG_GMEAN 2
#define C2C_AGG_SUM 3
/* factorization algorithms: cell2cell/tensor/factorization.py:91-104 (MU) and :106-115 (HALS) */
#define C2C_ALGO_MU 0
#define C2C_ALGO_HALS 1
/* tensorly cvg_criterion */
#define C2C_CVG_ABS_REC_ERROR 0
#define C2C_CVG_REC_ERROR 1

int c2c_version(void);
const char* c2c_last_error(void);
/* kernels this library has enqueued so far in the process (graph replays count every kernel they run) */
long long c2c_launch_count(void);
/* A factorization run that is never polled (check_every <= 0) returns once its iterations are ENQUEUED on `stream`: results
 * (factors, errors, state) are ordered after it on that stream, as for any asynchronous CUDA call.  The CUDA-graph objects of
 * such runs are parked and released by later calls; this waits for and releases those of the current device (optional). */
int c2c_release_pending(void);
/* Smallest rank whose two streaming passes run on the tensor pipe (TMA + tcgen05.mma kind::tf32, three-term split for
 * fp32 accuracy) instead of the FFMA kernels; below it the contraction is cheap enough for the CUDA cores to hide under
 * the HBM stream.  Returns the previous value; a value > 32 turns the tensor-pipe path off.  No counterpart in the
 * reference (its arithmetic is tensorly's numpy / torch dot, cell2cell/tensor/factorization.py:91-104). */
int c2c_set_mma_min_rank(int rank);

/* ---- K1: context-wise ligand-receptor score build --------------------------------------------------------------
 * Replaces add_complexes_to_expression (cell2cell/preprocessing/rnaseq.py:176-196), the generate_ccc_tensor loop
 * (cell2cell/tensor/tensor.py:1195-1202), compute_ccc_matrix (cell2cell/core/communication_scores.py:195-203) and the
 * mask / nan_to_num step (cell2cell/tensor/tensor.py:1141-1145).
 *
 * expr      all contexts' expression matrices, concatenated; context c is [G_c x expr_ld[c]] row-major at float offset
 *           expr_off[c]
 * cell_col  [C x K]: column of tensor cell type k in context c's matrix, or -1 when that context lacks it (-> NaN)
 * For side A (ligand) and side B (receptor), per (context c, LR pair l), entry c*L + l:
 *   count >= 1 : the partner is the aggregate (agg_kind) of rows sub_rows[first .. first+count) of context c
 *                (count = 1: a plain gene; count > 1: a complex whose subunits are all present in the context)
 *   count == 0 : partner absent from the context -> NaN -> X = 0, mask = 0
 *   count == -1: complex with a subunit missing from the context -> a row of zeros (rnaseq.py:195)
 * X         out [C, L, K, K] fp32 (NaN -> 0, +-inf -> +-FLT_MAX, as numpy.nan_to_num)
 * mask      out [C, L, K, K] uint8, 1 where the score is not NaN; may be NULL (how='inner')
 */
int c2c_build_ccc(const float* expr, const int64_t* expr_off, const int32_t* expr_ld, const int32_t* cell_col,
                  const int32_t* a_first, const int32_t* a_count, const int32_t* b_first, const int32_t* b_count,
                  const int32_t* sub_rows, int C, int L, int K, int score_kind, int agg_kind,
                  float* X, uint8_t* mask, void* stream);

/* ---- K1b: aggregate LR pairs into groups ------------------------------------------------------------------------
 * Replaces aggregate_ccc_tensor (cell2cell/tensor/tensor.py:1245-1252) -> aggregate_ccc_matrices
 * (cell2cell/core/communication_scores.py:235-242).  X_in/mask_in are [C, L, E]; group g aggregates LR rows
 * grp_rows[grp_off[g] .. grp_off[g+1]) ; outputs are [C, n_groups, E].  mask_in NULL = every entry valid.
 * method: C2C_AGG_GMEAN (exp(mean(log x)): zero -> 0, NaN / negative -> NaN), C2C_AGG_SUM (nansum), C2C_AGG_MEAN (nanmean).
 */
int c2c_group_aggregate(const float* X_in, const uint8_t* mask_in, const int32_t* grp_off, const int32_t* grp_rows,
                        int C, int L, int n_groups, int64_t E, int method, float* X_out, uint8_t* mask_out,
                        void* stream);

/* ---- workspace query for every factorization entry point below -------------------------------------------------- */
size_t c2c_ncp_workspace_bytes(int ndim, const int64_t* shape /*host*/, int rank, int masked);

/* ---- ||X||^2 (fp64 accumulation): tensorly tl.norm(tensor, 2) in non_negative_parafac ---------------------------- */
int c2c_sumsq(const float* X, int64_t n, double* out /*device, 1 double*/, void* ws, size_t ws_bytes, void* stream);

/* ---- K2: fused (masked) MTTKRP for one mode ---------------------------------------------------------------------
 * Replaces tensorly unfolding_dot_khatri_rao + the masked imputation `tensor*mask + cp_to_tensor(.., mask=1-mask)`
 * (mirrored in-repo at cell2cell/tensor/coupled_factorization.py:342-345).  One pass over X (+mask):
 *   M[i, r] = sum_{idx: idx[mode] = i} Xh[idx] * prod_{k != mode} A_k[idx[k], r]
 * A: host array of ndim device pointers.  M: [shape[mode] x ldr].
 */
int c2c_mttkrp(const float* X, const uint8_t* mask, int ndim, const int64_t* shape /*host*/,
               const float* const* A /*host array*/, int ldr, int rank, int mode, float* M,
               void* ws, size_t ws_bytes, void* stream);

/* The two streaming passes of K2 on their own (what c2c_mttkrp and c2c_ncp_run are built from):
 *   plane contraction  T[p, r]    = sum_{s,v} Xh[p, s, v] * A_{N-2}[s, r] * A_{N-1}[v, r]        T: [P x ldr], P = prod of leading dims
 *   fiber contraction  U[s, v, r] = sum_p     Xh[p, s, v] * prod_{k < N-2} A_k[i_k(p), r]        U: [I_{N-2} * I_{N-1} x ldr]
 * T serves the MTTKRP of every leading mode, U that of the two trailing modes. */
int c2c_plane_contract(const float* X, const uint8_t* mask, int ndim, const int64_t* shape /*host*/,
                       const float* const* A /*host array*/, int ldr, int rank, float* T,
                       void* ws, size_t ws_bytes, void* stream);
int c2c_fiber_contract(const float* X, const uint8_t* mask, int ndim, const int64_t* shape /*host*/,
                       const float* const* A /*host array*/, int ldr, int rank, float* U,
                       void* ws, size_t ws_bytes, void* stream);

/* Second half of c2c_mttkrp on its own: M of `mode` from T (mode < ndim - 2) or from U (the two trailing modes).  A sharded
 * factorization (shards along a leading mode) all-reduces the partial M / U between the halves. */
int c2c_mttkrp_from_contraction(const float* TU, int ndim, const int64_t* shape /*host*/, const float* const* A /*host array*/,
                                int ldr, int rank, int mode, float* M, void* stream);

/* ---- K3: Gram / Hadamard ------------------------------------------------------------------------------------------
 * G = prod_{k != skip_mode} A_k^T A_k  (fp32 [ldr x ldr]); rec_norm2 (device double, may be NULL) receives
 * sum_{r,s} prod_{all k} (A_k^T A_k)[r,s] = cp_norm^2.  skip_mode = -1 skips nothing.
 * Mirrors cell2cell/tensor/coupled_factorization.py:336-340 (accum) and tensorly cp_norm.
 */
int c2c_gram_hadamard(const float* const* A /*host array*/, int ndim, const int64_t* shape /*host*/, int ldr, int rank,
                      int skip_mode, float* G, double* rec_norm2, void* ws, size_t ws_bytes, void* stream);
/* Same result when only factor `changed_mode` was modified since the last c2c_gram_hadamard / _update call made with this
 * workspace: its Gram is recomputed, the other Grams are the ones that call left in `ws` (2 launches instead of 2 * ndim;
 * the Gauss-Seidel order of the update loop changes one factor at a time). */
int c2c_gram_hadamard_update(const float* const* A /*host array*/, int ndim, const int64_t* shape /*host*/, int ldr, int rank,
                             int changed_mode, int skip_mode, float* G, double* rec_norm2, void* ws, size_t ws_bytes,
                             void* stream);

/* ---- K4: multiplicative update  A <- A * max(M, eps) / max(A G, eps)  (coupled_factorization.py:346-348) ---------- */
int c2c_mu_update(float* A_mode, const float* M, const float* G, int64_t I, int rank, int ldr, float eps,
                  void* ws, size_t ws_bytes, void* stream);

/* ---- K4c: shared-mode multiplicative update of the coupled factorization (coupled_factorization.py:432-458) ---------
 * new = A1 * max((M1 + M2) / 2, eps) / max((A1 G1 + A2 G2) / 2, eps), written to A1_mode and A2_mode (the two tensors'
 * copies of the shared factor, distinct buffers).  iprod (device double[2], may be NULL) receives sum(M1 * new) and
 * sum(M2 * new) -- the inner products <X_k, cp_k> of tensorly's error_calc (coupled_factorization.py:473-485) when
 * this is the last update of the iteration. */
int c2c_mu_update_coupled(float* A1_mode, float* A2_mode, const float* M1, const float* M2, const float* G1, const float* G2,
                          int64_t I, int rank, int ldr, float eps, double* iprod, void* stream);

/* ---- Coupled factorization of two tensors sharing modes (cell2cell/tensor/coupled_factorization.py:123-482) --------
 * shared_pairs: host int[2 * npairs] = (tensor-1 mode, tensor-2 mode) per shared pair, in the order of
 * mode_mapping['shared'].  Per iteration: tensor-1-only modes, tensor-2-only modes (each a plain multiplicative update),
 * then the shared pairs in reverse order with averaged numerators / denominators (K4c); A1 / A2 hold separate copies of
 * a shared factor and both receive the update.  An unmasked tensor is read twice per iteration whatever the order (its
 * plane / fiber contractions are reused until a factor they depend on changes); a masked one once per mode.
 * errors: device double[2 * n_iter_max], errors[2 it + k] = Gram-form error of tensor k after iteration it
 * (tensorly error_calc, :420-431; for a masked tensor the direct form on the re-imputed tensor: the Gram form plus the
 * squared norm of the imputed entries); the stop rule (:437-458) runs on (w1 e1 + w2 e2) / (w1 + w2).  A masked tensor must hold
 * zeros at its missing entries.
 * state: device int[2] = {iterations done, converged}.  normX2: device double[2] (||X1||^2, ||X2||^2) or NULL.
 * ws1 / ws2: one workspace per tensor (c2c_ncp_workspace_bytes(shape_k, rank, mask_k != NULL)).  rank <= 32. */
int c2c_coupled_ncp_run(const float* X1, const uint8_t* mask1, int ndim1, const int64_t* shape1 /*host*/, float* const* A1 /*host array*/,
                        const float* X2, const uint8_t* mask2, int ndim2, const int64_t* shape2 /*host*/, float* const* A2 /*host array*/,
                        int ldr, int rank, const int* shared_pairs /*host*/, int npairs, int n_iter_max, double tol,
                        int cvg_criterion, double w1, double w2, const double* normX2, double* errors, int* state,
                        int check_every, void* ws1, size_t ws1_bytes, void* ws2, size_t ws2_bytes, void* stream);

/* ---- K5: HALS column sweeps (tensorly hals_nnls reached from cell2cell/tensor/factorization.py:106-115) ------------
 * sweeps k = 0..R-1: A[:,k] <- max(A[:,k] + (M[:,k] - A G[:,k]) / G[k,k], 0); stops when the squared change of a sweep
 * falls below tol * (first sweep's) or after min(max_sweeps, 2 + 0.5 * (1 + 2 I R / (R R + R))) sweeps. */
int c2c_hals_update(float* A_mode, const float* M, const float* G, int64_t I, int rank, int ldr, int max_sweeps,
                    float tol, void* stream);

/* ---- K6: reconstruction error sums (cell2cell/tensor/factorization.py:408-431, 458; tensor.py:659-688) -------------
 * out[0..4] (device doubles) = sum m (x - rec), sum m (x - rec)^2, sum m x, sum m x^2, sum m  with m = mask (1 if NULL). */
int c2c_recon_sums(const float* X, const uint8_t* mask, int ndim, const int64_t* shape /*host*/,
                   const float* const* A /*host array*/, int ldr, int rank, double* out,
                   void* ws, size_t ws_bytes, void* stream);

/* out (device double) = ||(1 - mask) * reconstruction||^2: the squared norm of the entries tensorly's imputation
 * `tensor*mask + cp_to_tensor(cp, mask=1-mask)` writes (coupled_factorization.py:342-343).  The coupled factorization's error is
 * error_calc WITHOUT an MTTKRP (coupled_factorization.py:420-431), i.e. the direct norm of (imputed tensor - reconstruction):
 * ||imputed||^2 = ||mask * X||^2 + this term. */
int c2c_missing_rec_sumsq(const float* X, const uint8_t* mask, int ndim, const int64_t* shape /*host*/,
                          const float* const* A /*host array*/, int ldr, int rank, double* out, void* ws, size_t ws_bytes,
                          void* stream);

/* ---- K8: cp_normalize (cell2cell/tensor/tensor.py:327): unit-norm columns in place, norms multiplied into
 * weights [rank] (device, fp32; input value = current weights, folded into factor 0 first). */
int c2c_cp_normalize(float* const* A /*host array*/, int ndim, const int64_t* shape /*host*/, int ldr, int rank,
                     float* weights, void* stream);

/* ---- one mode's update of tensorly's constrained_parafac (AO-ADMM; tensorly/solvers/admm.py `admm`), reached from cell2cell
 * with tf_type='constrained_parafac' (cell2cell/tensor/factorization.py:128-136).  M: this mode's MTTKRP [I x ldr] fp32;
 * G: Hadamard product of the other modes' Grams [ldr x ldr] fp32 (c2c_gram_hadamard); X64 / dual / aux: the factor, the dual
 * variable and the split variable [I x rank] fp64, kept by the caller between outer iterations; A_mode [I x ldr] fp32 receives
 * the new factor for the streaming passes.  constraint: 0 none, 1 non_negative, 2 l1_reg(param), 3 l2_square_reg(param).
 * out: device double[2] = {||x - aux|| / ||x|| (this mode's term of tensorly's constraint error), inner iterations made}. */
int c2c_admm_update(float* A_mode, double* X64, double* dual, double* aux, const float* M, const float* G, int64_t I,
                    int rank, int ldr, int constraint, double param, int n_inner, double tol, double* out, void* stream);

/* ---- CorrIndex of every pair of n decompositions (cell2cell/tensor/metrics.py:12-179: correlation_index,
 * _compute_correlation_index, pairwise_correlation_index; called per rank from cell2cell/tensor/factorization.py:360-369).
 * F: device double [n][rows][rank], the factor matrices of each decomposition stacked along the rows; seg_off: device
 * int64 [nseg + 1] row offsets of the blocks scored separately ({0, rows} for method 'stacked', the mode offsets
 * otherwise); method: 0 stacked, 1 max_score, 2 min_score, 3 avg_score; out: device double [n][n] (symmetric, zero
 * diagonal); *status (device int, zeroed by the caller) becomes 1 when a column norm is zero (the reference raises). */
int c2c_corrindex_pairs(const double* F, int n, int64_t rows, int rank, const int64_t* seg_off, int nseg, int method,
                        double tol, double* out, int* status, void* stream);

/* Measurement aid (no counterpart in the reference): average duration, in ms, of the streaming kernel of one unmasked pass ALONE
 * -- plane_stream / plane_mma (which = 0) or fiber_stream / fiber_mma (which = 1) -- over `reps` back-to-back launches, timed
 * with CUDA events on `stream` (c2c_plane_contract / c2c_fiber_contract add a memset, a tail kernel, the reduction of the
 * partials and a copy).  *kind (optional): 0 = FFMA kernel, 1 = tensor-pipe kernel.  -60: the shape takes the register-staged
 * kernels.  Synchronises the stream. */
int c2c_time_pass_kernel(const float* X, int ndim, const int64_t* shape /*host*/, const float* const* A /*host array*/, int ldr,
                         int rank, int which, int reps, float* ms_out /*host*/, int* kind /*host*/, void* ws, size_t ws_bytes,
                         void* stream);

/* ---- whole factorization (tensorly non_negative_parafac / non_negative_parafac_hals semantics) --------------------
 * Factors A hold the initial values on entry (K7: host-seeded random or svd init is done by the caller) and the result
 * on exit.  errors: device double[n_iter_max] (tensorly's rec_errors); state: device int[2] = {iterations done,
 * converged flag}.  The convergence test runs on the device every iteration; the host polls the flag every
 * `check_every` iterations (0 = never: run exactly n_iter_max iterations unless the device flag turns the remaining
 * launches into no-ops).  Synchronises the stream before returning only when it polls.
 * normX2: device double holding ||X||^2, or NULL to have it computed here.
 */
int c2c_ncp_run(const float* X, const uint8_t* mask, int ndim, const int64_t* shape /*host*/,
                float* const* A /*host array*/, int ldr, int rank, int algo, int n_iter_max, double tol,
                int cvg_criterion, const double* normX2, double* errors, int* state, int check_every,
                void* ws, size_t ws_bytes, void* stream);

/* `nruns` independent multiplicative-update factorizations of the same tensor side by side in one set of factor matrices
 * ([I_n x ldr]): run b has rank run_ranks[b] and owns the next run_ranks[b] columns.  The streaming passes serve every run
 * with ONE read of the tensor, the updates treat the runs as independent (no cross-run Gram terms), every run has its own
 * error history errors[b * n_iter_max + it] and stops on its own (its columns are frozen from then on).  This is the double
 * loop of _multiple_runs_elbow_analysis (cell2cell/tensor/factorization.py:335-352): runs of any ranks can share a pass as
 * long as their ranks add up to <= 32.  state: int32[2 + 2 * nruns] = {iterations enqueued, all runs stopped,
 * stopped[nruns], iterations[nruns]}.  Unmasked tensors only. */
int c2c_ncp_run_batched(const float* X, int ndim, const int64_t* shape, float* const* A, int ldr, const int* run_ranks, int nruns,
                        int n_iter_max, double tol, int cvg_criterion, const double* normX2, double* errors, int* state,
                        int check_every, void* workspace, size_t workspace_bytes, void* stream);

/* ---- small tensors: a grid of independent factorizations, one CTA each ---------------------------------------------------
 * The double loop of _multiple_runs_elbow_analysis (cell2cell/tensor/factorization.py:335-352) and the best-of-runs loop of
 * compute_tensor_factorization (cell2cell/tensor/tensor.py:294-320) on tensors small enough to stay in L2 (BALF-sized: a few
 * MB): every run is one CTA that runs ALL its iterations -- both contractions, the multiplicative updates of every mode, the
 * Gram-form error and the stop rule (semantics of c2c_ncp_run, unmasked MU) -- with no launches or grid syncs in between; the
 * grid is the list of runs (order = launch order: put the most expensive first).
 * A: host array of nruns * ndim device pointers, A[b * ndim + k] = factor k of run b, [I_k x ldr_b] with ldr_b = ranks[b]
 * rounded up to 4.  errors: device double[nruns * n_iter_max]; state: device int[2 * nruns] = {iterations, converged} per run;
 * sums: device double[5 * nruns] (or NULL) = the c2c_recon_sums values of each run's final factors.  normX2: device double.
 * Returns -61 when the shape is outside this path (trailing modes > 128, rank > 32, shared memory). */
size_t c2c_ncp_multi_workspace_bytes(int ndim, const int64_t* shape /*host*/, const int* ranks /*host*/, int nruns);
int c2c_ncp_run_multi(const float* X, int ndim, const int64_t* shape /*host*/, float* const* A /*host array*/,
                      const int* ranks /*host*/, int nruns, int n_iter_max, double tol, int cvg_criterion, const double* normX2,
                      double* errors, int* state, double* sums, void* ws, size_t ws_bytes, void* stream);

/* ---- one factorization sharded along the first (context) mode over the GPUs of a box --------------------------------------
 * BASELINE config 4 / SURVEY section 8(e) row 3; the reference has no multi-process factorization (its loop is tensorly's,
 * cell2cell/tensor/factorization.py:91-104), the arithmetic is c2c_ncp_run's.  Every rank (one process per GPU) calls
 * c2c_ncp_run_sharded with ITS slab X_local[c_lo:c_hi, ...] (local_shape[0] = c_hi - c_lo), A[0] = the local rows of the
 * context factor, A[1..] = full copies of the other factors (identical on every rank), normX2 = ||X||^2 of the WHOLE tensor.
 * Per iteration each rank makes the two streaming passes over its slab; the sums that cross slabs (the context factor's Gram,
 * the partial M of the other leading modes, the partial U) are exchanged inside the update kernels through peer memory:
 * comm[g] = rank g's communication buffer as mapped in this process (comm[me] local; the others opened from CUDA IPC handles
 * after c2c_enable_peer_access), each c2c_shard_comm_bytes(...) bytes, zero-filled once when allocated, 256-byte aligned.
 * epoch_base: any value larger than the last epoch_base + n_iter_max used with these buffers (same on every rank).  Every rank
 * must enqueue the call at about the same time (a wait on a peer gives up after ~5 s and sets the int at offset 0 of the local
 * buffer; the results are then invalid).  Replicated factors, errors and the stop decision come out bit-identical on every
 * rank.  Unmasked multiplicative updates, ndim >= 4, rank <= 32, world <= 8; -71 when the shape is outside this path. */
size_t c2c_shard_comm_bytes(int ndim, const int64_t* shape /*host; mode 0 unused*/, int rank, int world);
int c2c_enable_peer_access(int peer_device);
/* Communication buffers owned by the library (the one exception to "caller-owned buffers": CUDA IPC exports cudaMalloc
 * allocations): c2c_comm_alloc returns a zero-filled device buffer and its 64-byte cudaIpcMemHandle_t; a peer process opens it
 * with ITS device current (c2c_comm_open), closes it with c2c_comm_close; the owner frees it with c2c_comm_free.
 * c2c_comm_probe runs one exchange over the mapped buffers (every rank must call it; out[0] = value * world (world + 1) / 2). */
int c2c_comm_alloc(size_t bytes, void** ptr_out /*host*/, unsigned char* handle_out /*host, 64 bytes*/);
int c2c_comm_open(const unsigned char* handle /*host, 64 bytes*/, void** ptr_out /*host*/);
int c2c_comm_close(void* ptr);
int c2c_comm_free(void* ptr);
int c2c_comm_probe(int world, int me, void* const* comm /*host array*/, size_t comm_bytes, unsigned int epoch, float value,
                   float* out, void* stream);
int c2c_ncp_run_sharded(const float* X_local, int ndim, const int64_t* local_shape /*host*/, float* const* A /*host array*/, int ldr,
                        int rank, int n_iter_max, double tol, int cvg_criterion, const double* normX2, double* errors, int* state,
                        int check_every, void* ws, size_t ws_bytes, int world, int me, void* const* comm /*host array*/,
                        size_t comm_bytes, unsigned int epoch_base, void* stream);

/* ---- mask plan: the masked factorization without a pass over the mask ---------------------------------------------------
 * tensorly re-imputes the missing entries with the current reconstruction before every mode's MTTKRP
 * (`tensor*mask + cp_to_tensor(cp, mask=1-mask)`, mirrored at cell2cell/tensor/coupled_factorization.py:342-345).  With zeros
 * stored at the missing entries (what cell2cell holds: tensor.py:842-847, 933-946) that MTTKRP is the unmasked contraction of
 * the tensor plus the contraction of the reconstruction over the MISSING set.  The mask of a how='outer' tensor
 * (cell2cell/tensor/tensor.py:1083-1145) is separable -- observed[p, s, v] = g[p] * hs[c(p), s] * hv[c(p), v]: a gene pair
 * absent from a context removes a plane, a cell type absent from it removes a row and a column of every plane of the context --
 * and over such a set the sum collapses to products of R x R restricted Grams.  The plan, built once per (tensor, mask), holds
 * the tightest separable model of any mask plus the residual entries (missing, but observed by the model) as lists.
 *
 * header: device buffer of c2c_mask_plan_header_bytes(shape) bytes (256-byte aligned).
 * c2c_mask_plan_scan fills it and returns (synchronising the stream) info[8] (host):
 *   [0] missing entries  [1] slots of the by-plane residual list (entries + class padding; 0 = no residual entries)
 *   [2] 1 if a non-zero value sits under a zero mask byte (plan not usable)
 *   [3] slots of the by-position residual lists (per 1024-plane chunk)          [4] 1 if the separable model misses anything
 *   [5] 1 on overflow (plan not usable)
 *   [6] bytes of the `lists` device buffer c2c_mask_plan_fill needs          [7] residual entries
 * X may be NULL (no zero-fill check).  c2c_mask_plan_fill writes the residual lists (nothing to do when info[1] == 0).
 */
size_t c2c_mask_plan_header_bytes(int ndim, const int64_t* shape /*host*/);
int c2c_mask_plan_scan(const float* X, const uint8_t* mask, int ndim, const int64_t* shape /*host*/, void* header,
                       size_t header_bytes, int64_t* info /*host, 8*/, void* stream);
int c2c_mask_plan_fill(const uint8_t* mask, int ndim, const int64_t* shape /*host*/, const void* header, size_t header_bytes,
                       const int64_t* info /*host*/, void* lists, size_t lists_bytes, void* stream);
/* c2c_ncp_run (multiplicative updates) for a masked tensor with its plan: per iteration the tensor is read twice, the mask
 * never.  Returns -50 when the plan path does not apply (rank > 32, trailing modes too large for the fused updates, info[2] or
 * info[5] set): call c2c_ncp_run instead.  ws: c2c_ncp_workspace_bytes(.., masked = 1). */
int c2c_ncp_run_planned(const float* X, const uint8_t* mask, int ndim, const int64_t* shape /*host*/, float* const* A /*host array*/,
                        int ldr, int rank, int n_iter_max, double tol, int cvg_criterion, const double* normX2, double* errors,
                        int* state, int check_every, void* ws, size_t ws_bytes, const void* header, size_t header_bytes,
                        const void* lists, size_t lists_bytes, const int64_t* info /*host*/, void* stream);
/* c2c_plane_contract (which = 0, out = T [P x ldr]) / c2c_fiber_contract (which = 1, out = U [E x ldr]) of the imputed tensor:
 * the unmasked pass over X plus the plan's corrections for the given factors. */
int c2c_masked_contract_planned(const float* X, int ndim, const int64_t* shape /*host*/, const float* const* A /*host array*/,
                                int ldr, int rank, int which, float* out, void* ws, size_t ws_bytes, const void* header,
                                size_t header_bytes, const void* lists, size_t lists_bytes, const int64_t* info /*host*/,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* C2C_B200_H */
