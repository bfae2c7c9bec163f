// Fused factor updates of the multiplicative-update NCP (everything between two streaming passes).
//
//   lead_update   one launch per leading mode k:  M_k rows from T (the plane contraction), the multiplicative update of
//                 those rows, the Gram of the new factor.  A CTA owns a few complete rows of the factor, so the update is
//                 row-local; per-CTA Gram partials are summed in a fixed order by the last CTA to finish.
//   trail_update  one single-CTA launch for BOTH trailing modes: U (the fiber contraction) is staged in shared memory
//                 once, then M_{N-2} -> update -> Gram -> M_{N-1} (with the new A_{N-2}) -> update -> Gram -> the
//                 reconstruction-error scalar, the stop rule and the iteration counter.
// Replaces, per iteration, 2 x (lead_mttkrp, mu_update, gram_finalize) + 2 x (trail_mttkrp, mu_update, gram_finalize):
// 12 dependent launches become 3.  Semantics are those of tensorly's non_negative_parafac update mirrored at
// cell2cell/tensor/coupled_factorization.py:336-348 (accum = Hadamard of the other modes' Grams, clip at eps, A*num/den)
// and its error / stop rule (:437-458).
#pragma once
#include "c2c_common.cuh"

#define C2C_LU_THREADS 256
#define C2C_TU_THREADS 1024

// The Grams S_j = A_j^T A_j live in two banks of N matrices: iteration t accumulates the new Grams into bank t & 1 (zero on
// entry) and reads the not-yet-updated modes' Grams from the other bank; trail_update clears the bank the next iteration
// will accumulate into.  `state[0]` is the device-side iteration counter.
struct LeadUpdArgs {
    const float* T;            // [P x ldr]
    LeadInfo lead;
    int k;                     // leading mode updated (also its global mode index)
    long long P;
    int R, ldr, N;
    float* A;                  // [I_k x ldr] factor of mode k, updated in place
    double* S;                 // [2][N][ldr*ldr]
    const int* state;
    int rows_per_cta;
    float eps;
    const int* done;
    // batched runs: the R (<= 32) columns belong to independent factorizations, column c to run run_of[c] (all zero: one
    // run).  Columns of different runs do not interact (their Gram-Hadamard entries are taken as zero); a run whose flag is
    // set is frozen.
    unsigned char run_of[32];
    const int* run_flags;      // [number of runs] or null
};

// G[q][r] = prod_{j < k} Snew_j[q][r] * prod_{j > k} Sold_j[q][r] into shared memory (float), R x R; zero across runs
__device__ __forceinline__ void hadamard_to_smem(const double* Snew, const double* Sold, int N, int k, int R, int ldr, const unsigned char* run_of, float* sG)
{
    for (int t = threadIdx.x; t < R * R; t += blockDim.x) {
        const int q = t / R, r = t - q * R;
        double g = 1.0;
        for (int j = 0; j < N; ++j)
            if (j != k) g *= (j < k ? Snew : Sold)[(size_t)j * ldr * ldr + q * ldr + r];
        sG[t] = (run_of[q & 31] == run_of[r & 31]) ? (float)g : 0.0f;
    }
}

// `tpr` threads cooperate on one row (32: a warp per row, for modes whose rows each touch few planes; blockDim: the whole
// CTA per row, for short modes whose rows touch thousands of planes); blockDim / tpr rows are in flight per CTA.
__global__ void __launch_bounds__(1024)
lead_update_kernel(const LeadUpdArgs a, const int tpr)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(a.done)) return;
    extern __shared__ __align__(16) unsigned char lu_smem[];
    const int R = a.R, ldr = a.ldr, q4 = ldr >> 2, LL = ldr * ldr;
    const int nthr = blockDim.x;
    const int nrc = nthr / tpr;                              // rows in flight
    const int njr = tpr / q4;                                // j-lanes per row
    float* sG = reinterpret_cast<float*>(lu_smem);           // [R*R]
    float* sNew = sG + R * R;                                // [nrc][ldr] updated rows
    double* sRed = reinterpret_cast<double*>(lu_smem + (((size_t)(R * R + nrc * ldr) * sizeof(float) + 7) & ~(size_t)7));   // [nrc][njr][ldr]
    const int par = a.state[0] & 1;
    double* Snew = a.S + (size_t)par * a.N * LL;
    const double* Sold = a.S + (size_t)(par ^ 1) * a.N * LL;

    hadamard_to_smem(Snew, Sold, a.N, a.k, R, ldr, a.run_of, sG);

    const int dk = a.lead.dim[a.k];
    int inner = 1;
    for (int m = a.k + 1; m < a.lead.nl; ++m) inner *= a.lead.dim[m];
    const int cnt = (int)(a.P / dk);                         // planes carrying one index of mode k
    const int rslot = threadIdx.x / tpr, lt = threadIdx.x - rslot * tpr;
    const int quad = lt % q4, jl = lt / q4;
    const int i_begin = blockIdx.x * a.rows_per_cta;
    int i_end = i_begin + a.rows_per_cta; if (i_end > dk) i_end = dk;
    const int nl = a.lead.nl;
    // 4-way tensors (two leading modes): plane of (i, j) and the other mode's factor row are affine in j
    const float4* oth = nl == 2 ? reinterpret_cast<const float4*>(a.lead.fac[1 - a.k]) + quad : nullptr;
    const long long sI = (nl == 2 && a.k == 0) ? a.lead.dim[1] : 1, sJ = (nl == 2 && a.k == 1) ? a.lead.dim[1] : 1;

    double gacc[4] = {0.0, 0.0, 0.0, 0.0};                   // Gram entries tid, tid + nthr, ..  (R * R <= 4 * nthr)
    __syncthreads();

    for (int ib = i_begin; ib < i_end; ib += nrc) {
        const int i = ib + rslot;
        const bool row_on = i < i_end;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_on && jl < njr) {
            if (nl <= 2) {
                const float4* tb = reinterpret_cast<const float4*>(a.T) + quad;
                int j = jl;
                for (; j + 3 * njr < cnt; j += 4 * njr) {    // four independent (T, factor) row pairs in flight
                    float4 t[4], f[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const long long p = nl == 2 ? i * sI + (long long)(j + u * njr) * sJ : i;
                        t[u] = __ldg(tb + (size_t)p * q4);
                        f[u] = nl == 2 ? __ldg(oth + (size_t)(j + u * njr) * q4) : make_float4(1.f, 1.f, 1.f, 1.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc.x = fmaf(t[u].x, f[u].x, acc.x); acc.y = fmaf(t[u].y, f[u].y, acc.y);
                        acc.z = fmaf(t[u].z, f[u].z, acc.z); acc.w = fmaf(t[u].w, f[u].w, acc.w);
                    }
                }
                for (; j < cnt; j += njr) {
                    const long long p = nl == 2 ? i * sI + (long long)j * sJ : i;
                    const float4 t = __ldg(tb + (size_t)p * q4);
                    const float4 f = nl == 2 ? __ldg(oth + (size_t)j * q4) : make_float4(1.f, 1.f, 1.f, 1.f);
                    acc.x = fmaf(t.x, f.x, acc.x); acc.y = fmaf(t.y, f.y, acc.y);
                    acc.z = fmaf(t.z, f.z, acc.z); acc.w = fmaf(t.w, f.w, acc.w);
                }
            } else {
                for (int j = jl; j < cnt; j += njr) {
                    const int o = j / inner, n = j - o * inner;
                    const long long p = ((long long)o * dk + i) * inner + n;
                    const float4 t = __ldg(reinterpret_cast<const float4*>(a.T + (size_t)p * ldr) + quad);
                    // weights of the other leading modes: modes above k from n, modes below k from o
                    float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
                    int rem = n;
                    for (int m = nl - 1; m > a.k; --m) {
                        const int d = a.lead.dim[m]; const int qq = rem / d; const int idx = rem - qq * d; rem = qq;
                        const float4 f = __ldg(reinterpret_cast<const float4*>(a.lead.fac[m] + (size_t)idx * ldr) + quad);
                        w.x *= f.x; w.y *= f.y; w.z *= f.z; w.w *= f.w;
                    }
                    rem = o;
                    for (int m = a.k - 1; m >= 0; --m) {
                        const int d = a.lead.dim[m]; const int qq = rem / d; const int idx = rem - qq * d; rem = qq;
                        const float4 f = __ldg(reinterpret_cast<const float4*>(a.lead.fac[m] + (size_t)idx * ldr) + quad);
                        w.x *= f.x; w.y *= f.y; w.z *= f.z; w.w *= f.w;
                    }
                    acc.x = fmaf(t.x, w.x, acc.x); acc.y = fmaf(t.y, w.y, acc.y);
                    acc.z = fmaf(t.z, w.z, acc.z); acc.w = fmaf(t.w, w.w, acc.w);
                }
            }
        }
        if (jl < njr) {
            double* dst = sRed + ((size_t)rslot * njr + jl) * ldr + quad * 4;
            dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w;
        }
        __syncthreads();
        if (lt < R) {                                        // M[i, r] and the update of row i (zero row when off)
            const int r = lt;
            float nv = 0.0f;
            if (row_on) {
                double m = 0.0;
                for (int g = 0; g < njr; ++g) m += sRed[((size_t)rslot * njr + g) * ldr + r];
                const float* arow = a.A + (size_t)i * ldr;
                float den = 0.0f;
                for (int q = 0; q < R; ++q) den = fmaf(arow[q], sG[q * R + r], den);
                nv = arow[r] * fmaxf((float)m, a.eps) / fmaxf(den, a.eps);
                if (a.run_flags != nullptr && a.run_flags[a.run_of[r & 31]]) nv = arow[r];  // this run has stopped
            }
            sNew[rslot * ldr + r] = nv;
        }
        __syncthreads();
        if (lt < R && row_on) a.A[(size_t)i * ldr + lt] = sNew[rslot * ldr + lt];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = threadIdx.x + u * nthr;
            if (e < R * R) {
                const int er = e / R, es = e - er * R;
                double g = 0.0;
                for (int rs = 0; rs < nrc; ++rs) g += (double)sNew[rs * ldr + er] * (double)sNew[rs * ldr + es];
                gacc[u] += g;
            }
        }
        // (sNew / sRed are rewritten only after the next batch's first barrier)
    }
    // this CTA's share of the new Gram (fp64 atomics into the zeroed bank)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = threadIdx.x + u * nthr;
        if (e < R * R) atomicAdd(Snew + (size_t)a.k * LL + (e / R) * ldr + (e % R), gacc[u]);
    }
}

struct TrailUpdArgs {
    const float* U;            // [I2*I3 x ldr]
    float* A2; float* A3;      // factors of the two trailing modes, updated in place
    int I2, I3, R, ldr, N;
    double* S;                 // [2][N][ldr*ldr]
    float eps;
    int u_in_smem;             // 1: U is staged in shared memory
    int which_begin, which_end; // trailing modes updated by this launch: [0, 2) = both (unmasked), [0, 1) / [1, 2) = one (masked)
    const double* normX2; double* errors; int* state; int n_iter_max; double tol; int cvg;
    const int* done;
    // batched runs (see LeadUpdArgs): errors is [nruns][n_iter_max], run_flags / run_iters are [nruns]
    unsigned char run_of[32];
    int nruns;
    int* run_flags; int* run_iters;
};

__device__ __forceinline__ double block_sum_1024(double v, double* red /*[32]*/)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(C2C_TU_THREADS, 1)
trail_update_kernel(const TrailUpdArgs a)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(a.done)) return;
    extern __shared__ __align__(16) unsigned char tu_smem[];
    const int R = a.R, ldr = a.ldr, I2 = a.I2, I3 = a.I3, N = a.N, LL = ldr * ldr, RR = R * R;
    const int E = I2 * I3, Imax = I2 > I3 ? I2 : I3;
    double* sS = reinterpret_cast<double*>(tu_smem);               // [N][R*R] current Grams (new where updated, else old)
    double* red = sS + (size_t)N * RR;                             // [32]
    float* sG = reinterpret_cast<float*>(red + 32);                // [R*R]
    float* sA2 = sG + RR;                                          // [I2*R]
    float* sA3 = sA2 + (size_t)I2 * R;                             // [I3*R]
    float* sM = sA3 + (size_t)I3 * R;                              // [Imax*R]
    float* sU = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sM + (size_t)Imax * R) + 15) & ~(uintptr_t)15);
    const int par = a.state[0] & 1;
    double* Snew = a.S + (size_t)par * N * LL;
    double* Sold = a.S + (size_t)(par ^ 1) * N * LL;
    // everything global this launch needs, in one round trip
    for (int t = threadIdx.x; t < N * RR; t += blockDim.x) {
        const int j = t / RR, e = t - j * RR;
        const bool is_new = j < N - 2 + a.which_begin;             // modes already updated in this iteration
        sS[t] = (is_new ? Snew : Sold)[(size_t)j * LL + (e / R) * ldr + (e % R)];
    }
    const float* Usrc = a.U;
    if (a.u_in_smem) {
        const float4* src = reinterpret_cast<const float4*>(a.U);
        float4* dst = reinterpret_cast<float4*>(sU);
        for (int t = threadIdx.x; t < E * (ldr >> 2); t += blockDim.x) dst[t] = __ldg(src + t);
        Usrc = sU;
    }
    for (int t = threadIdx.x; t < I2 * R; t += blockDim.x) sA2[t] = a.A2[(size_t)(t / R) * ldr + (t % R)];
    for (int t = threadIdx.x; t < I3 * R; t += blockDim.x) sA3[t] = a.A3[(size_t)(t / R) * ldr + (t % R)];
    double iprod = 0.0;
    __shared__ double s_ip[32], s_rn[32];                      // per-run <M, A> and ||rec||^2 (batched runs)
    if (threadIdx.x < 32) { s_ip[threadIdx.x] = 0.0; s_rn[threadIdx.x] = 0.0; }
    __syncthreads();

    for (int which = a.which_begin; which < a.which_end; ++which) {
        const int mode = N - 2 + which;
        const int I = which == 0 ? I2 : I3, J = which == 0 ? I3 : I2;
        float* sA = which == 0 ? sA2 : sA3;                        // factor being updated
        const float* sO = which == 0 ? sA3 : sA2;                  // the other trailing factor (current values)
        float* gA = which == 0 ? a.A2 : a.A3;
        for (int t = threadIdx.x; t < RR; t += blockDim.x) {
            double g = 1.0;
            for (int j = 0; j < N; ++j) if (j != mode) g *= sS[(size_t)j * RR + t];
            sG[t] = (a.run_of[(t / R) & 31] == a.run_of[(t % R) & 31]) ? (float)g : 0.0f;
        }
        // M[i, r] = sum_j U[e(i, j), r] * O[j, r]
        for (int t = threadIdx.x; t < I * R; t += blockDim.x) {
            const int i = t / R, r = t - i * R;
            float acc0 = 0.0f, acc1 = 0.0f;
            int j = 0;
            for (; j + 1 < J; j += 2) {
                const int e0 = which == 0 ? i * I3 + j : j * I3 + i;
                const int e1 = which == 0 ? e0 + 1 : e0 + I3;
                acc0 = fmaf(Usrc[(size_t)e0 * ldr + r], sO[j * R + r], acc0);
                acc1 = fmaf(Usrc[(size_t)e1 * ldr + r], sO[(j + 1) * R + r], acc1);
            }
            if (j < J) {
                const int e0 = which == 0 ? i * I3 + j : j * I3 + i;
                acc0 = fmaf(Usrc[(size_t)e0 * ldr + r], sO[j * R + r], acc0);
            }
            sM[t] = acc0 + acc1;
        }
        __syncthreads();
        // multiplicative update (row-local): new = old * max(M, eps) / max(old G, eps)
        double ip = 0.0;
        float nv_keep[2] = {0.0f, 0.0f};                           // I * R <= 2 * blockDim (checked by the host)
        int nk = 0;
        for (int t = threadIdx.x; t < I * R; t += blockDim.x, ++nk) {
            const int i = t / R, r = t - i * R;
            float den = 0.0f;
            for (int q = 0; q < R; ++q) den = fmaf(sA[i * R + q], sG[q * R + r], den);
            const float m = sM[t];
            float nv = sA[t] * fmaxf(m, a.eps) / fmaxf(den, a.eps);
            if (a.run_flags != nullptr && a.run_flags[a.run_of[r & 31]]) nv = sA[t];          // this run has stopped
            ip += (double)m * (double)nv;
            gA[(size_t)i * ldr + r] = nv;
            nv_keep[nk & 1] = nv;
        }
        __syncthreads();                                           // every thread is done reading the old factor
        nk = 0;
        for (int t = threadIdx.x; t < I * R; t += blockDim.x, ++nk) sA[t] = nv_keep[nk & 1];
        if (which == 1) iprod = ip;
        __syncthreads();
        if (a.nruns > 1 && which == 1) {
            // batched runs: <M, A_new> per run, in a fixed order (was: one shared-memory fp64 atomicAdd -- a CAS loop -- per
            // factor entry onto a handful of addresses: 86 us of a 28-column bin's iteration).  Column sums by warp w for the
            // columns w, w + nwarps, ... (lanes stride the rows, fixed shuffle tree), then one thread per run adds its columns.
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
            for (int r = warp; r < R; r += nwarps) {
                double v = 0.0;
                for (int i = lane; i < I; i += 32) v += (double)sM[i * R + r] * (double)sA[i * R + r];
                v = warp_sum(v);
                if (lane == 0) red[r] = v;
            }
            __syncthreads();
            if (threadIdx.x < a.nruns) {
                double sum = 0.0;
                for (int r = 0; r < R; ++r) if (a.run_of[r & 31] == threadIdx.x) sum += red[r];
                s_ip[threadIdx.x] = sum;
            }
            __syncthreads();
        }
        // Gram of the new factor (fp64), to shared memory and to the new bank
        for (int t = threadIdx.x; t < RR; t += blockDim.x) {
            const int q = t / R, r = t - q * R;
            double g0 = 0.0, g1 = 0.0;
            int i = 0;
            for (; i + 1 < I; i += 2) {
                g0 += (double)sA[i * R + q] * (double)sA[i * R + r];
                g1 += (double)sA[(i + 1) * R + q] * (double)sA[(i + 1) * R + r];
            }
            if (i < I) g0 += (double)sA[i * R + q] * (double)sA[i * R + r];
            sS[(size_t)mode * RR + t] = g0 + g1;
            Snew[(size_t)mode * LL + q * ldr + r] = g0 + g1;
        }
        __syncthreads();
    }
    if (a.which_end < 2) return;
    // the bank the next iteration accumulates into
    for (int t = threadIdx.x; t < N * LL; t += blockDim.x) Sold[t] = 0.0;
    // reconstruction error (Gram form), stop rule, iteration counter
    if (a.nruns > 1) {
        // batched runs: one error / stop decision per run
        // ||rec||^2 per run = sum over the run's own R_b x R_b block of the Hadamard product of the Grams, in a fixed order:
        // thread q sums row q over the columns of its run, then one thread per run adds its rows
        if (threadIdx.x < R) {
            const int q = threadIdx.x;
            double rowsum = 0.0;
            for (int r = 0; r < R; ++r) {
                if (a.run_of[q & 31] == a.run_of[r & 31]) {
                    double g = 1.0;
                    for (int j = 0; j < N; ++j) g *= sS[(size_t)j * RR + q * R + r];
                    rowsum += g;
                }
            }
            red[q] = rowsum;
        }
        __syncthreads();
        if (threadIdx.x < a.nruns) {
            double sum = 0.0;
            for (int q = 0; q < R; ++q) if (a.run_of[q & 31] == threadIdx.x) sum += red[q];
            s_rn[threadIdx.x] = sum;
        }
        __syncthreads();
        const int it = a.state[0];
        if (threadIdx.x < a.nruns && a.errors && !a.run_flags[threadIdx.x]) {
            const int b = threadIdx.x;
            const double nx2 = *a.normX2;
            const double err = sqrt(fabs(nx2 + s_rn[b] - 2.0 * s_ip[b])) / sqrt(nx2);
            double* eb = a.errors + (size_t)b * a.n_iter_max;
            if (it < a.n_iter_max) eb[it] = err;
            a.run_iters[b] = it + 1;
            if (it >= 1) {
                const double dec = eb[it - 1] - err;
                if (a.cvg == 0 ? (fabs(dec) < a.tol) : (dec < a.tol)) a.run_flags[b] = 1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int all = 1;
            for (int b = 0; b < a.nruns; ++b) all &= (a.run_flags[b] != 0);
            if (all) a.state[1] = 1;
            a.state[0] = it + 1;
        }
        return;
    }
    iprod = block_sum_1024(iprod, red);
    double t = 0.0;
    for (int pr = threadIdx.x; pr < RR; pr += blockDim.x) {
        double g = 1.0;
        for (int j = 0; j < N; ++j) g *= sS[(size_t)j * RR + pr];
        t += g;
    }
    const double rn2 = block_sum_1024(t, red);
    if (threadIdx.x == 0 && a.errors) {
        const double nx2 = *a.normX2;
        const double err = sqrt(fabs(nx2 + rn2 - 2.0 * iprod)) / sqrt(nx2);
        const int it = a.state[0];
        if (it < a.n_iter_max) a.errors[it] = err;
        if (it >= 1) {
            const double dec = a.errors[it - 1] - err;
            const bool stop = a.cvg == 0 ? (fabs(dec) < a.tol) : (dec < a.tol);
            if (stop) a.state[1] = 1;
        }
        a.state[0] = it + 1;
    }
}
