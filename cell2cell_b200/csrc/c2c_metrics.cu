// CorrIndex between every pair of a set of decompositions (Sobhani et al. 2022), the similarity metric of the elbow analysis:
// cell2cell/tensor/metrics.py:12-179 (correlation_index / _compute_correlation_index / pairwise_correlation_index), called
// per rank over the runs' factor sets from cell2cell/tensor/factorization.py:360-369.  The reference loops over the
// n (n - 1) / 2 pairs in Python (normalise, stack, |X1^T X2|, row / column maxima); here one launch scores all pairs:
//
//   F[d][s][r]   the n decompositions' factor matrices stacked along the rows (S = sum of the mode sizes), fp64, dense
//   seg_off      row offsets of the `nseg` blocks scored separately: {0, S} for method 'stacked', the mode offsets otherwise
//   CTA (i < j)  per block: column norms of both operands, C = |X_i^T X_j| / (norm_i norm_j^T) accumulated in fp64 in a fixed
//                order, score = (sum_q |max_r C[q,r] - 1| + sum_r |max_q C[q,r] - 1|) / (2 R), below `tol` -> 0
//                blocks combined by `method` (stacked / max / min / mean)
//
// The inputs are a few thousand rows by <= 128 columns: the kernel is latency-sized, not bandwidth-sized; what it replaces is
// the O(n^2) host loop of a rank's runs.
#include "c2c_metrics.cuh"

#define CI_THREADS 256
#define CI_TILE 32                      // rows staged per step

__global__ void __launch_bounds__(CI_THREADS)
corrindex_pairs_kernel(const double* __restrict__ F, int n, long long S, int R, const long long* __restrict__ seg_off, int nseg,
                       int method, double tol, double* __restrict__ out, int* __restrict__ status)
{
    extern __shared__ __align__(16) unsigned char ci_smem[];
    double* t1 = reinterpret_cast<double*>(ci_smem);                 // [CI_TILE][R]
    double* t2 = t1 + (size_t)CI_TILE * R;                           // [CI_TILE][R]
    double* C = t2 + (size_t)CI_TILE * R;                            // [R][R]
    double* n1 = C + (size_t)R * R;                                  // [R]
    double* n2 = n1 + R;                                             // [R]
    double* red = n2 + R;                                            // [2 R] row / column terms
    // pair index -> (i, j), i < j
    int i = 0, rem = blockIdx.x;
    while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
    const int j = i + 1 + rem;
    const double* X1 = F + (size_t)i * S * R;
    const double* X2 = F + (size_t)j * S * R;
    const int tid = threadIdx.x;
    const int nent = R * R;
    double best = 0.0, sum = 0.0;
    for (int g = 0; g < nseg; ++g) {
        const long long s0 = seg_off[g], s1 = seg_off[g + 1];
        // entries of C owned by this thread: e = tid, tid + 256, ... (<= 64 at R = 128); norms by the first 2 R threads
        double acc[64];
#pragma unroll
        for (int u = 0; u < 64; ++u) acc[u] = 0.0;
        double nacc = 0.0;
        for (long long r0 = s0; r0 < s1; r0 += CI_TILE) {
            const int rows = (int)(s1 - r0 < CI_TILE ? s1 - r0 : CI_TILE);
            __syncthreads();
            for (int t = tid; t < rows * R; t += CI_THREADS) {
                t1[t] = X1[(size_t)r0 * R + t];
                t2[t] = X2[(size_t)r0 * R + t];
            }
            __syncthreads();
            // two-level sums (a tile's 32 rows first, then across tiles): rounding grows with 32 + S / 32, not with S
            if (tid < 2 * R) {
                const double* tt = tid < R ? t1 + tid : t2 + (tid - R);
                double a = 0.0;
                for (int s = 0; s < rows; ++s) a += tt[(size_t)s * R] * tt[(size_t)s * R];
                nacc += a;
            }
#pragma unroll
            for (int u = 0; u < 64; ++u) {
                const int e = tid + u * CI_THREADS;
                if (e < nent) {
                    const int q = e / R, r = e - q * R;
                    double a = 0.0;
                    for (int s = 0; s < rows; ++s) a += t1[(size_t)s * R + q] * t2[(size_t)s * R + r];
                    acc[u] += a;
                }
            }
        }
        __syncthreads();
        if (tid < R) n1[tid] = sqrt(nacc);
        else if (tid < 2 * R) n2[tid - R] = sqrt(nacc);
        __syncthreads();
        if (tid < 2 * R && (tid < R ? n1[tid] : n2[tid - R]) == 0.0) atomicExch(status, 1);   // 'Column norms must be non-zero'
#pragma unroll
        for (int u = 0; u < 64; ++u) {
            const int e = tid + u * CI_THREADS;
            if (e < nent) {
                const int q = e / R, r = e - q * R;
                C[e] = fabs(acc[u] / (n1[q] * n2[r]));
            }
        }
        __syncthreads();
        if (tid < R) {                                               // row maxima
            double m = C[(size_t)tid * R];
            for (int r = 1; r < R; ++r) m = fmax(m, C[(size_t)tid * R + r]);
            red[tid] = fabs(m - 1.0);
        } else if (tid < 2 * R) {                                    // column maxima
            const int r = tid - R;
            double m = C[r];
            for (int q = 1; q < R; ++q) m = fmax(m, C[(size_t)q * R + r]);
            red[tid] = fabs(m - 1.0);
        }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int q = 0; q < R; ++q) a += red[q];
            for (int r = 0; r < R; ++r) b += red[R + r];
            double score = (a + b) / (double)(2 * R);
            if (score < tol) score = 0.0;
            if (g == 0) best = score;
            else if (method == 1) best = fmax(best, score);
            else if (method == 2) best = fmin(best, score);
            sum += score;
        }
    }
    if (tid == 0) {
        const double v = (method == 3) ? sum / (double)nseg : best;
        out[(size_t)i * n + j] = v;
        out[(size_t)j * n + i] = v;
    }
}

__global__ void corrindex_diag_kernel(int n, double* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(size_t)i * n + i] = 0.0;
}

size_t c2c_corrindex_smem_bytes(int R)
{
    return ((size_t)2 * CI_TILE * R + (size_t)R * R + 4 * (size_t)R) * sizeof(double);
}

cudaError_t c2c_launch_corrindex_pairs(const double* F, int n, long long S, int R, const long long* seg_off, int nseg, int method,
                                       double tol, double* out, int* status, cudaStream_t st, int* launches)
{
    *launches = 0;
    corrindex_diag_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ++*launches;
    if (n < 2) return cudaSuccess;
    const size_t smem = c2c_corrindex_smem_bytes(R);
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(corrindex_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int npairs = n * (n - 1) / 2;
    corrindex_pairs_kernel<<<npairs, CI_THREADS, smem, st>>>(F, n, S, R, seg_off, nseg, method, tol, out, status);
    e = cudaGetLastError();
    if (e == cudaSuccess) ++*launches;
    return e;
}
