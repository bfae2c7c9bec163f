// One mode's update of tensorly's constrained_parafac (AO-ADMM, Huang / Sidiropoulos / Liavas 2016) as a single-CTA kernel:
// reached from cell2cell with tf_type='constrained_parafac' (cell2cell/tensor/factorization.py:128-136).  tensorly's inner
// solver (tensorly/solvers/admm.py `admm`, called from decomposition/_constrained_cp.py) restated:
//
//     rho = trace(G) / R                                   G = Hadamard product of the other modes' Grams (R x R)
//     repeat n_inner times:
//         x_split = solve((G + rho I)^T, (M + rho (x + dual))^T)          M = this mode's MTTKRP (I x R)
//         x       = prox(x_split^T - dual)                                non_negative | l1_reg | l2_square_reg | none
//         dual   += x - x_split^T
//         stop when ||x - x_split^T|| < tol ||x||  and  ||x - x_old|| < tol ||dual||
//
// The R x R system is the same for every row: it is factorized once (Cholesky: G is a Hadamard product of Grams, so
// G + rho I is symmetric positive definite) and every thread then carries whole rows through the two triangular solves,
// the proximal step and the dual update; the four norms of the stop rule are reduced over the CTA in a fixed order.  The
// factor is kept in fp64 (`X64`) between outer iterations, with the fp32 copy the streaming passes read written alongside.
#include "c2c_admm.cuh"

#define ADMM_THREADS 1024
#define ADMM_MAX_R 128

__device__ __forceinline__ double admm_prox(double v, int constraint, double param)
{
    switch (constraint) {
    case 1: return v > 0.0 ? v : 0.0;                                           // non_negative: clip(v, 0, max)
    case 2: { const double a = fabs(v) - param; return a > 0.0 ? (v > 0.0 ? a : -a) : 0.0; }   // l1_reg: soft thresholding
    case 3: return v / (1.0 + 2.0 * param);                                     // l2_square_reg
    default: return v;
    }
}

__global__ void __launch_bounds__(ADMM_THREADS)
admm_update_kernel(float* __restrict__ A, double* __restrict__ X64, double* __restrict__ dual, double* __restrict__ aux,
                   const float* __restrict__ M, const float* __restrict__ G, long long I, int R, int ldr, int constraint,
                   double param, int n_inner, double tol, double* __restrict__ out /* [2]: ||x - aux|| / ||x||, inner iterations */)
{
    extern __shared__ __align__(16) unsigned char admm_smem[];
    double* L = reinterpret_cast<double*>(admm_smem);                 // [R][R] lower Cholesky factor of G + rho I
    double* red = L + (size_t)R * R;                                  // [4][32] warp partials
    __shared__ double s_rho, s_tot[4];
    __shared__ int s_stop;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    for (int e = tid; e < R * R; e += blockDim.x) L[e] = (double)G[(size_t)(e / R) * ldr + (e % R)];
    __syncthreads();
    if (tid == 0) {
        double tr = 0.0;
        for (int q = 0; q < R; ++q) tr += L[(size_t)q * R + q];
        s_rho = tr / (double)R;
        s_stop = 0;
    }
    __syncthreads();
    const double rho = s_rho;
    // Cholesky (right-looking, one column at a time; the CTA shares the trailing update)
    for (int k = 0; k < R; ++k) {
        if (tid == 0) L[(size_t)k * R + k] = sqrt(L[(size_t)k * R + k] + rho);
        __syncthreads();
        const double d = L[(size_t)k * R + k];
        for (int i = k + 1 + tid; i < R; i += blockDim.x) L[(size_t)i * R + k] = 0.5 * (L[(size_t)i * R + k] + L[(size_t)k * R + i]) / d;
        __syncthreads();
        const int m = R - k - 1;
        for (int e = tid; e < m * m; e += blockDim.x) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) {
                const double v = L[(size_t)i * R + j] - L[(size_t)i * R + k] * L[(size_t)j * R + k];
                L[(size_t)i * R + j] = v;
                if (j < i) L[(size_t)j * R + i] = v;                // keeps the untouched part symmetric for the next column
            }
        }
        __syncthreads();                                              // (every diagonal entry gets its rho in its own sqrt)
    }
    int it_done = 0;
    double last_dr = 0.0, last_xn = 0.0;
    for (int it = 0; it < n_inner; ++it) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};                          // ||x - y||^2, ||x||^2, ||x - x_old||^2, ||dual||^2
        for (long long i = tid; i < I; i += blockDim.x) {
            double y[ADMM_MAX_R];
            double* xr = X64 + (size_t)i * R;
            double* dr = dual + (size_t)i * R;
            for (int q = 0; q < R; ++q) y[q] = (double)M[(size_t)i * ldr + q] + rho * (xr[q] + dr[q]);
            for (int q = 0; q < R; ++q) {                              // L z = rhs
                double s = y[q];
                for (int p = 0; p < q; ++p) s -= L[(size_t)q * R + p] * y[p];
                y[q] = s / L[(size_t)q * R + q];
            }
            for (int q = R - 1; q >= 0; --q) {                         // L^T y = z
                double s = y[q];
                for (int p = q + 1; p < R; ++p) s -= L[(size_t)p * R + q] * y[p];
                y[q] = s / L[(size_t)q * R + q];
            }
            for (int q = 0; q < R; ++q) {
                const double xo = xr[q], du = dr[q];
                const double xn = admm_prox(y[q] - du, constraint, param);
                const double dn = du + xn - y[q];
                aux[(size_t)i * R + q] = y[q];
                xr[q] = xn;
                dr[q] = dn;
                A[(size_t)i * ldr + q] = (float)xn;
                acc[0] += (xn - y[q]) * (xn - y[q]);
                acc[1] += xn * xn;
                acc[2] += (xn - xo) * (xn - xo);
                acc[3] += dn * dn;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double v = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[j * 32 + wid] = v;
        }
        __syncthreads();
        if (tid < 4) {
            double v = 0.0;
            for (int w = 0; w < nw; ++w) v += red[tid * 32 + w];
            s_tot[tid] = v;
        }
        __syncthreads();
        it_done = it + 1;
        last_dr = s_tot[0];
        last_xn = s_tot[1];
        if (tid == 0 && sqrt(s_tot[0]) < tol * sqrt(s_tot[1]) && sqrt(s_tot[2]) < tol * sqrt(s_tot[3])) s_stop = 1;
        __syncthreads();
        if (s_stop) break;
    }
    if (tid == 0) {
        out[0] = sqrt(last_dr) / sqrt(last_xn);
        out[1] = (double)it_done;
    }
}

cudaError_t c2c_launch_admm_update(float* A, double* X64, double* dual, double* aux, const float* M, const float* G, long long I,
                                   int R, int ldr, int constraint, double param, int n_inner, double tol, double* out,
                                   cudaStream_t st)
{
    if (R > ADMM_MAX_R) return cudaErrorInvalidValue;
    const size_t smem = ((size_t)R * R + 4 * 32) * sizeof(double);
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(admm_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int threads = ADMM_THREADS;
    while (threads > 64 && threads / 2 >= I) threads /= 2;
    admm_update_kernel<<<1, threads, smem, st>>>(A, X64, dual, aux, M, G, I, R, ldr, constraint, param, n_inner, tol, out);
    return cudaGetLastError();
}
