// Small tensors: ONE CTA per factorization, all iterations inside the kernel, a grid of independent (rank, seed) runs.
//
// The elbow analysis of the reference is a double loop of independent factorizations of the same tensor
// (cell2cell/tensor/factorization.py:335-352: `for r in range(1, upper_rank + 1): for run in range(runs): ...`), and
// compute_tensor_factorization(runs > 1) another (cell2cell/tensor/tensor.py:294-320).  On the tensors real Tensor-cell2cell
// analyses use (BALF 12 x 1054 x 6 x 6 = 1.8 MB, the toy 4 x 300 x 6 x 6) a whole-GPU kernel per MTTKRP is launch- and
// latency-bound: the tensor is L2-resident and one pass takes microseconds.  Here each run gets one CTA that keeps the run's
// Grams, the trailing factors, B = A_{N-2} (*) A_{N-1} and U in shared memory, streams the tensor from L2 twice per iteration
// (plane contraction, fiber contraction -- the same two-pass restructuring as the large-tensor path, c2c_api.cu) and does
// every mode's multiplicative update, the Gram-form error and tensorly's stop rule itself: no launches, no grid syncs, and
// the 148 SMs work on (up to) 296 runs at a time.  Semantics per run are those of c2c_ncp_run (tensorly non_negative_parafac,
// mirrored at cell2cell/tensor/coupled_factorization.py:336-348, 437-458).
#include "c2c_multi.cuh"

#define MT 256                                     // threads per CTA

struct MultiSmem {                                 // offsets (bytes) into dynamic shared memory, computed on the host
    int sS, sG, sBU, sA2, sA3, sNew, sM, sPart, sRed, total;
};

__host__ __device__ inline MultiSmem multi_layout(int N, int NQ, int E, int I2, int I3)
{
    MultiSmem m;
    int off = 0;
    auto take = [&](int bytes) { const int o = off; off += (bytes + 15) / 16 * 16; return o; };
    const int Imax = I2 > I3 ? I2 : I3;
    m.sS = take(N * NQ * NQ * 8);                  // double [N][NQ*NQ]   Grams
    m.sRed = take(64 * 8);                         // double [64]
    m.sG = take(NQ * NQ * 4);                      // float  [NQ*NQ]      Hadamard of the other modes' Grams
    m.sBU = take(E * NQ * 4);                      // float  [E][NQ]      B (plane phase) / U (fiber and trailing phases)
    m.sA2 = take(I2 * NQ * 4);
    m.sA3 = take(I3 * NQ * 4);
    m.sNew = take(Imax * NQ * 4);
    m.sM = take(128 * NQ * 4);                     // float  [<=128][NQ]  M of a short mode
    m.sPart = take(MT * (NQ > 16 ? NQ : 16) * 4);  // float  [MT][max(NQ, 16)]  slice partials / Gram tiles
    m.total = off;
    return m;
}

size_t c2c_multi_smem_bytes(int ndim, int ldr, long long E, int I2, int I3)
{
    return (size_t)multi_layout(ndim, ldr, (int)E, I2, I3).total;
}

template <int NQ>
struct Run {
    const MultiArgs& a; const MultiRun& r; unsigned char* sm; MultiSmem L;
    double* sS; double* sRed; float* sG; float* sBU; float* sA2; float* sA3; float* sNew; float* sM; float* sPart;
    int N, nl, R;
    __device__ Run(const MultiArgs& a_, const MultiRun& r_, unsigned char* sm_) : a(a_), r(r_), sm(sm_)
    {
        N = a.ndim; nl = N - 2; R = r.rank;
        L = multi_layout(N, NQ, a.E, a.I2, a.I3);
        sS = reinterpret_cast<double*>(sm + L.sS); sRed = reinterpret_cast<double*>(sm + L.sRed);
        sG = reinterpret_cast<float*>(sm + L.sG); sBU = reinterpret_cast<float*>(sm + L.sBU);
        sA2 = reinterpret_cast<float*>(sm + L.sA2); sA3 = reinterpret_cast<float*>(sm + L.sA3);
        sNew = reinterpret_cast<float*>(sm + L.sNew); sM = reinterpret_cast<float*>(sm + L.sM);
        sPart = reinterpret_cast<float*>(sm + L.sPart);
    }

    // Gram of a factor [I x NQ] (global or shared) -> sS[k]: 4 x 4 tiles x row slices in fp32, slices summed in fp64
    __device__ void gram(const float* A, int I, int k)
    {
        constexpr int q4 = NQ / 4, ntile = q4 * q4, nsl = MT / ntile;
        const int t = threadIdx.x, sl = t / ntile, tile = t - sl * ntile, tq = tile / q4, tr = tile - tq * q4;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
        if (sl < nsl) {
            for (int i = sl; i < I; i += nsl) {
                const float4 x = *(reinterpret_cast<const float4*>(A + (size_t)i * NQ) + tq);
                const float4 y = *(reinterpret_cast<const float4*>(A + (size_t)i * NQ) + tr);
                const float xv[4] = {x.x, x.y, x.z, x.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int w = 0; w < 4; ++w) acc[u * 4 + w] = fmaf(xv[u], yv[w], acc[u * 4 + w]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) sPart[(size_t)t * 16 + i] = acc[i];
        }
        __syncthreads();
        for (int e = t; e < NQ * NQ; e += MT) {
            const int q = e / NQ, rr = e - q * NQ;
            const int tl = (q >> 2) * q4 + (rr >> 2), ix = (q & 3) * 4 + (rr & 3);
            double s = 0.0;
            for (int k2 = 0; k2 < nsl; ++k2) s += (double)sPart[((size_t)k2 * ntile + tl) * 16 + ix];
            sS[(size_t)k * NQ * NQ + e] = s;
        }
        __syncthreads();
    }

    // sG = Hadamard product of every mode's Gram except `k`
    __device__ void hadamard(int k)
    {
        for (int e = threadIdx.x; e < NQ * NQ; e += MT) {
            double g = 1.0;
            for (int j = 0; j < N; ++j) if (j != k) g *= sS[(size_t)j * NQ * NQ + e];
            sG[e] = (float)g;
        }
        __syncthreads();
    }

    // Khatri-Rao row of plane p over the leading modes except `skip` (-1: none)
    __device__ __forceinline__ void lead_row(long long p, int skip, float (&w)[NQ]) const
    {
#pragma unroll
        for (int q = 0; q < NQ; ++q) w[q] = 1.0f;
        long long rem = p;
        for (int k = nl - 1; k >= 0; --k) {
            const int d = a.shape[k];
            const long long qq = rem / d;
            const int i = (int)(rem - qq * d);
            rem = qq;
            if (k == skip) continue;
            const float4* row = reinterpret_cast<const float4*>(r.A[k] + (size_t)i * NQ);
#pragma unroll
            for (int c = 0; c < NQ / 4; ++c) {
                const float4 f = row[c];
                w[4 * c] *= f.x; w[4 * c + 1] *= f.y; w[4 * c + 2] *= f.z; w[4 * c + 3] *= f.w;
            }
        }
    }

    // new row = old * max(m, eps) / max(old G, eps)   (coupled_factorization.py:346-348)
    __device__ __forceinline__ void mu_row(const float (&old)[NQ], const float (&m)[NQ], float (&nw)[NQ]) const
    {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float den = 0.0f;
#pragma unroll
            for (int s = 0; s < NQ; ++s) den = fmaf(old[s], sG[s * NQ + q], den);
            nw[q] = old[q] * fmaxf(m[q], a.eps) / fmaxf(den, a.eps);
        }
    }

    // T[p, :] = sum_e X[p, e] B[e, :]   with B = A_{N-2} (*) A_{N-1} in shared memory; a thread per plane
    __device__ void plane_phase()
    {
        const int E = a.E, I3 = a.I3;
        for (int t = threadIdx.x; t < E * NQ; t += MT) {
            const int e = t / NQ, q = t - e * NQ, s = e / I3, v = e - s * I3;
            sBU[t] = sA2[s * NQ + q] * sA3[v * NQ + q];
        }
        __syncthreads();
        const bool vec = (E & 3) == 0;
        for (long long p = threadIdx.x; p < a.P; p += MT) {
            float acc[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[q] = 0.0f;
            const float* xp = a.X + (size_t)p * E;
            if (vec) {
                const float4* x4 = reinterpret_cast<const float4*>(xp);
#pragma unroll 2
                for (int e4 = 0; e4 < (E >> 2); ++e4) {
                    const float4 xv = __ldg(x4 + e4);
                    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4* b4 = reinterpret_cast<const float4*>(sBU + (size_t)(e4 * 4 + u) * NQ);
#pragma unroll
                        for (int c = 0; c < NQ / 4; ++c) {
                            const float4 b = b4[c];
                            acc[4 * c] = fmaf(xs[u], b.x, acc[4 * c]); acc[4 * c + 1] = fmaf(xs[u], b.y, acc[4 * c + 1]);
                            acc[4 * c + 2] = fmaf(xs[u], b.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(xs[u], b.w, acc[4 * c + 3]);
                        }
                    }
                }
            } else {
#pragma unroll 4
                for (int e = 0; e < E; ++e) {
                    const float x = __ldg(xp + e);
                    const float4* b4 = reinterpret_cast<const float4*>(sBU + (size_t)e * NQ);
#pragma unroll
                    for (int c = 0; c < NQ / 4; ++c) {
                        const float4 b = b4[c];
                        acc[4 * c] = fmaf(x, b.x, acc[4 * c]); acc[4 * c + 1] = fmaf(x, b.y, acc[4 * c + 1]);
                        acc[4 * c + 2] = fmaf(x, b.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(x, b.w, acc[4 * c + 3]);
                    }
                }
            }
            float4* out = reinterpret_cast<float4*>(r.T + (size_t)p * NQ);
#pragma unroll
            for (int c = 0; c < NQ / 4; ++c) out[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        }
        __syncthreads();
    }

    // M_k rows from T, multiplicative update of A_k, its Gram
    __device__ void lead_update(int k)
    {
        hadamard(k);
        const int dk = a.shape[k];
        long long inner = 1;
        for (int m = k + 1; m < nl; ++m) inner *= a.shape[m];
        const long long cnt = a.P / dk;
        float* Ak = r.A[k];
        auto plane_of = [&](int i, long long j) { const long long o = j / inner, n = j - o * inner; return (o * dk + i) * inner + n; };
        if (dk > 128) {
            // long mode: a thread per row
            for (int i = threadIdx.x; i < dk; i += MT) {
                float m[NQ], w[NQ], old[NQ], nw[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) m[q] = 0.0f;
                for (long long j = 0; j < cnt; ++j) {
                    const long long p = plane_of(i, j);
                    lead_row(p, k, w);
                    const float4* t4 = reinterpret_cast<const float4*>(r.T + (size_t)p * NQ);
#pragma unroll
                    for (int c = 0; c < NQ / 4; ++c) {
                        const float4 tv = t4[c];
                        m[4 * c] = fmaf(tv.x, w[4 * c], m[4 * c]); m[4 * c + 1] = fmaf(tv.y, w[4 * c + 1], m[4 * c + 1]);
                        m[4 * c + 2] = fmaf(tv.z, w[4 * c + 2], m[4 * c + 2]); m[4 * c + 3] = fmaf(tv.w, w[4 * c + 3], m[4 * c + 3]);
                    }
                }
                float4* row = reinterpret_cast<float4*>(Ak + (size_t)i * NQ);
#pragma unroll
                for (int c = 0; c < NQ / 4; ++c) { const float4 f = row[c]; old[4 * c] = f.x; old[4 * c + 1] = f.y; old[4 * c + 2] = f.z; old[4 * c + 3] = f.w; }
                mu_row(old, m, nw);
#pragma unroll
                for (int c = 0; c < NQ / 4; ++c) row[c] = make_float4(nw[4 * c], nw[4 * c + 1], nw[4 * c + 2], nw[4 * c + 3]);
            }
        } else {
            // short mode: the planes of a row are split over MT / dk slices, partial rows summed through shared memory
            const int nsl = MT / dk;
            const int i = threadIdx.x % dk, sl = threadIdx.x / dk;
            if (sl < nsl) {
                float m[NQ], w[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) m[q] = 0.0f;
                for (long long j = sl; j < cnt; j += nsl) {
                    const long long p = plane_of(i, j);
                    lead_row(p, k, w);
                    const float4* t4 = reinterpret_cast<const float4*>(r.T + (size_t)p * NQ);
#pragma unroll
                    for (int c = 0; c < NQ / 4; ++c) {
                        const float4 tv = t4[c];
                        m[4 * c] = fmaf(tv.x, w[4 * c], m[4 * c]); m[4 * c + 1] = fmaf(tv.y, w[4 * c + 1], m[4 * c + 1]);
                        m[4 * c + 2] = fmaf(tv.z, w[4 * c + 2], m[4 * c + 2]); m[4 * c + 3] = fmaf(tv.w, w[4 * c + 3], m[4 * c + 3]);
                    }
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) sPart[((size_t)sl * dk + i) * NQ + q] = m[q];
            }
            __syncthreads();
            for (int t = threadIdx.x; t < dk * NQ; t += MT) {
                float s = 0.0f;
                for (int k2 = 0; k2 < nsl; ++k2) s += sPart[(size_t)k2 * dk * NQ + t];
                sM[t] = s;
            }
            __syncthreads();
            if (threadIdx.x < dk) {
                const int row_i = threadIdx.x;
                float m[NQ], old[NQ], nw[NQ];
                float4* row = reinterpret_cast<float4*>(Ak + (size_t)row_i * NQ);
#pragma unroll
                for (int c = 0; c < NQ / 4; ++c) { const float4 f = row[c]; old[4 * c] = f.x; old[4 * c + 1] = f.y; old[4 * c + 2] = f.z; old[4 * c + 3] = f.w; }
#pragma unroll
                for (int q = 0; q < NQ; ++q) m[q] = sM[row_i * NQ + q];
                mu_row(old, m, nw);
#pragma unroll
                for (int c = 0; c < NQ / 4; ++c) row[c] = make_float4(nw[4 * c], nw[4 * c + 1], nw[4 * c + 2], nw[4 * c + 3]);
            }
        }
        __syncthreads();
        gram(Ak, dk, k);
    }

    // W[p, :] (Khatri-Rao rows of the updated leading factors) overwrite T, then U[e, :] = sum_p X[p, e] W[p, :] -> sBU
    __device__ void fiber_phase()
    {
        const int E = a.E;
        for (long long p = threadIdx.x; p < a.P; p += MT) {
            float w[NQ];
            lead_row(p, -1, w);
            float4* out = reinterpret_cast<float4*>(r.T + (size_t)p * NQ);
#pragma unroll
            for (int c = 0; c < NQ / 4; ++c) out[c] = make_float4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
        __syncthreads();
        const int nsl = E <= 128 ? MT / E : 1;                        // plane slices per position
        for (int e0 = 0; e0 < E; e0 += (nsl > 1 ? E : MT)) {
            const int e = nsl > 1 ? (int)(threadIdx.x % E) : e0 + (int)threadIdx.x;
            const int sl = nsl > 1 ? (int)(threadIdx.x / E) : 0;
            float acc[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[q] = 0.0f;
            if (e < E && sl < nsl) {
                const float* xe = a.X + e;
                long long p = sl;
                for (; p + 3 * nsl < a.P; p += 4 * nsl) {             // four tensor loads in flight
                    float x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) x[u] = __ldg(xe + (size_t)(p + (long long)u * nsl) * E);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4* w4 = reinterpret_cast<const float4*>(r.T + (size_t)(p + (long long)u * nsl) * NQ);
#pragma unroll
                        for (int c = 0; c < NQ / 4; ++c) {
                            const float4 w = w4[c];
                            acc[4 * c] = fmaf(x[u], w.x, acc[4 * c]); acc[4 * c + 1] = fmaf(x[u], w.y, acc[4 * c + 1]);
                            acc[4 * c + 2] = fmaf(x[u], w.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(x[u], w.w, acc[4 * c + 3]);
                        }
                    }
                }
                for (; p < a.P; p += nsl) {
                    const float x = __ldg(xe + (size_t)p * E);
                    const float4* w4 = reinterpret_cast<const float4*>(r.T + (size_t)p * NQ);
#pragma unroll
                    for (int c = 0; c < NQ / 4; ++c) {
                        const float4 w = w4[c];
                        acc[4 * c] = fmaf(x, w.x, acc[4 * c]); acc[4 * c + 1] = fmaf(x, w.y, acc[4 * c + 1]);
                        acc[4 * c + 2] = fmaf(x, w.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(x, w.w, acc[4 * c + 3]);
                    }
                }
                if (nsl > 1) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) sPart[((size_t)sl * E + e) * NQ + q] = acc[q];
                } else {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) sBU[(size_t)e * NQ + q] = acc[q];
                }
            }
        }
        __syncthreads();
        if (nsl > 1) {
            for (int t = threadIdx.x; t < E * NQ; t += MT) {
                float s = 0.0f;
                for (int k2 = 0; k2 < nsl; ++k2) s += sPart[(size_t)k2 * E * NQ + t];
                sBU[t] = s;
            }
            __syncthreads();
        }
    }

    // trailing mode `which` (0: N-2, 1: N-1) from U in shared memory; returns sum(M * new) to every thread
    __device__ double trail_update(int which)
    {
        const int mode = N - 2 + which, I2 = a.I2, I3 = a.I3;
        const int I = which == 0 ? I2 : I3, J = which == 0 ? I3 : I2;
        float* sA = which == 0 ? sA2 : sA3;
        const float* sO = which == 0 ? sA3 : sA2;
        hadamard(mode);
        for (int t = threadIdx.x; t < I * NQ; t += MT) {
            const int i = t / NQ, q = t - i * NQ;
            float acc = 0.0f;
            for (int j = 0; j < J; ++j) {
                const int e = which == 0 ? i * I3 + j : j * I3 + i;
                acc = fmaf(sBU[(size_t)e * NQ + q], sO[j * NQ + q], acc);
            }
            sM[t] = acc;
        }
        __syncthreads();
        double ip = 0.0;
        for (int t = threadIdx.x; t < I * NQ; t += MT) {
            const int i = t / NQ, q = t - i * NQ;
            float den = 0.0f;
            for (int s = 0; s < NQ; ++s) den = fmaf(sA[i * NQ + s], sG[s * NQ + q], den);
            const float m = sM[t];
            const float nv = sA[t] * fmaxf(m, a.eps) / fmaxf(den, a.eps);
            sNew[t] = nv;
            ip += (double)m * (double)nv;
        }
        __syncthreads();
        float* gA = r.A[mode];
        for (int t = threadIdx.x; t < I * NQ; t += MT) { sA[t] = sNew[t]; gA[t] = sNew[t]; }
        __syncthreads();
        gram(sA, I, mode);
        return block_sum(ip);
    }

    // [sum (x - rec), sum (x - rec)^2, sum x, sum x^2, n] of the final factors, accumulated directly (the kept error of
    // cell2cell, tensor.py:311-315, and the sums of explained_variance, tensor.py:659-688): T still holds the Khatri-Rao rows of
    // the final leading factors, B is rebuilt from the final trailing factors
    __device__ void final_sums()
    {
        const int E = a.E, I3 = a.I3;
        for (int t = threadIdx.x; t < E * NQ; t += MT) {
            const int e = t / NQ, q = t - e * NQ, s = e / I3, v = e - s * I3;
            sBU[t] = sA2[s * NQ + q] * sA3[v * NQ + q];
        }
        __syncthreads();
        double sd = 0.0, sd2 = 0.0, sx = 0.0, sx2 = 0.0;
        for (long long p = threadIdx.x; p < a.P; p += MT) {
            float w[NQ];
            const float4* w4 = reinterpret_cast<const float4*>(r.T + (size_t)p * NQ);
#pragma unroll
            for (int c = 0; c < NQ / 4; ++c) { const float4 f = w4[c]; w[4 * c] = f.x; w[4 * c + 1] = f.y; w[4 * c + 2] = f.z; w[4 * c + 3] = f.w; }
            const float* xp = a.X + (size_t)p * E;
            float fd = 0.0f, fd2 = 0.0f, fx = 0.0f, fx2 = 0.0f;
            for (int e = 0; e < E; ++e) {
                const float x = __ldg(xp + e);
                const float4* b4 = reinterpret_cast<const float4*>(sBU + (size_t)e * NQ);
                float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
                for (int c = 0; c < NQ / 4; ++c) {
                    const float4 b = b4[c];
                    r0 = fmaf(w[4 * c], b.x, r0); r1 = fmaf(w[4 * c + 1], b.y, r1);
                    r0 = fmaf(w[4 * c + 2], b.z, r0); r1 = fmaf(w[4 * c + 3], b.w, r1);
                }
                const float d = x - (r0 + r1);
                fd += d; fd2 = fmaf(d, d, fd2); fx += x; fx2 = fmaf(x, x, fx2);
            }
            sd += (double)fd; sd2 += (double)fd2; sx += (double)fx; sx2 += (double)fx2;
        }
        sd = block_sum(sd); sd2 = block_sum(sd2); sx = block_sum(sx); sx2 = block_sum(sx2);
        if (threadIdx.x == 0) {
            r.sums[0] = sd; r.sums[1] = sd2; r.sums[2] = sx; r.sums[3] = sx2; r.sums[4] = (double)a.P * (double)E;
        }
    }

    __device__ double block_sum(double v)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = 0.0;
        for (int i = 0; i < MT / 32; ++i) s += sRed[i];
        __syncthreads();
        return s;
    }

    __device__ void run()
    {
        const int I2 = a.I2, I3 = a.I3;
        for (int t = threadIdx.x; t < I2 * NQ; t += MT) sA2[t] = r.A[N - 2][t];
        for (int t = threadIdx.x; t < I3 * NQ; t += MT) sA3[t] = r.A[N - 1][t];
        __syncthreads();
        for (int k = 0; k < nl; ++k) gram(r.A[k], a.shape[k], k);
        gram(sA2, I2, N - 2);
        gram(sA3, I3, N - 1);
        const double nx2 = *a.normX2;
        double prev = 0.0;
        int it = 0, conv = 0;
        for (; it < a.n_iter_max; ++it) {
            plane_phase();
            for (int k = 0; k < nl; ++k) lead_update(k);
            fiber_phase();
            trail_update(0);
            const double ip = trail_update(1);
            double t = 0.0;
            for (int e = threadIdx.x; e < NQ * NQ; e += MT) {
                double g = 1.0;
                for (int j = 0; j < N; ++j) g *= sS[(size_t)j * NQ * NQ + e];
                t += g;
            }
            const double rn2 = block_sum(t);
            const double err = sqrt(fabs(nx2 + rn2 - 2.0 * ip)) / sqrt(nx2);
            if (threadIdx.x == 0) r.errors[it] = err;
            if (it >= 1) {
                const double dec = prev - err;
                if (a.cvg == 0 ? (fabs(dec) < a.tol) : (dec < a.tol)) { conv = 1; ++it; break; }
            }
            prev = err;
        }
        if (threadIdx.x == 0) { r.state[0] = it; r.state[1] = conv; }
        if (r.sums != nullptr && it > 0) final_sums();
    }
};

__global__ void __launch_bounds__(MT, 2)
ncp_multi_kernel(const MultiArgs a)
{
    extern __shared__ __align__(16) unsigned char multi_smem[];
    __shared__ MultiRun r;
    if (threadIdx.x == 0) r = a.runs[blockIdx.x];
    __syncthreads();
    switch (r.ldr) {
        case 4:  { Run<4> x(a, r, multi_smem); x.run(); break; }
        case 8:  { Run<8> x(a, r, multi_smem); x.run(); break; }
        case 12: { Run<12> x(a, r, multi_smem); x.run(); break; }
        case 16: { Run<16> x(a, r, multi_smem); x.run(); break; }
        case 20: { Run<20> x(a, r, multi_smem); x.run(); break; }
        case 24: { Run<24> x(a, r, multi_smem); x.run(); break; }
        case 28: { Run<28> x(a, r, multi_smem); x.run(); break; }
        case 32: { Run<32> x(a, r, multi_smem); x.run(); break; }
        default: break;
    }
}

cudaError_t c2c_launch_ncp_multi(const MultiArgs& a, size_t smem, cudaStream_t st)
{
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(ncp_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    ncp_multi_kernel<<<a.nruns, MT, smem, st>>>(a);
    return cudaGetLastError();
}
