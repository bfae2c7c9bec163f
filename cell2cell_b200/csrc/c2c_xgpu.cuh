// One factorization sharded along the first (context) mode over the GPUs of an NVSwitch box (SURVEY.md section 8e, row 3;
// BASELINE config 4: "NCP rank 20 sharded along context mode at 1/2/4/8 GPUs").  Nothing in the reference does this (its
// loop is single-process tensorly, cell2cell/tensor/factorization.py:91-104); the arithmetic is that of the single-GPU
// iteration (c2c_api.cu, c2c_update.cuh): every rank streams ITS slab of the tensor and the three small sums that cross
// slabs are reduced INSIDE the update kernels over peer memory -- no NCCL call, no extra launch:
//
//   mode 0 (the sharded mode)    rows of the local contexts are complete locally: the single-GPU lead_update kernel.  Its
//                                Gram  S_0 = sum_g S_0^g  is needed by every later mode.
//   modes 1 .. N-3 (replicated)  lead_update_xg:  every CTA forms its rows of the PARTIAL M_k from the local T, stores them
//                                into every peer's buffer (NVLink stores, then one release flag per peer), waits for the
//                                peers' flags of the same CTA, sums the world's partial rows in rank order and updates its
//                                rows -- identically on every rank.  CTA 0 also carries the local S_0^g (k = 1).
//   trailing modes (replicated)  fiber_reduce_xg: the fixed-order sum of the local fiber partials becomes the rank's partial
//                                U, exchanged the same way; trail_update then runs unchanged on the summed U.
//
// Replicated quantities stay bit-identical on every rank (so every rank takes the same stop decision and nobody waits for a
// peer that has stopped): a cross-rank sum is always taken in rank order from values that have ONE producer, and the Grams
// of the replicated leading modes (accumulated with fp64 atomics, hence order-dependent) are taken from rank 0, which sends
// them along with the next exchange.
//
// Buffers: one communication buffer per rank (allocated by the caller, mapped into every peer by CUDA IPC), two slots per
// exchange (iteration parity), flags that only grow (epoch_base + iteration + 1): a slot is rewritten two iterations later,
// after its readers have signalled the exchange in between.  A wait that does not complete within ~5 s sets `err` in the
// local buffer and goes on (the host then reports the failure) instead of hanging the GPU.
#pragma once
#include "c2c_common.cuh"
#include "c2c_update.cuh"

#define C2C_XG_MAX_WORLD 8
#define C2C_XG_MAX_CTAS 512
#define C2C_XG_SPIN_CLOCKS (10000000000LL)

struct XgComm {
    int world, me;
    unsigned int epoch_base;
    char* peer[C2C_XG_MAX_WORLD];          // base of every rank's buffer as mapped here; peer[me] is the local one
    long long off_flags_m;                 // u32 [N-3][world][MAX_CTAS]      exchange of replicated leading mode k -> [k - 1]
    long long off_flags_u;                 // u32 [world][MAX_CTAS]
    long long off_s0;                      // double [2][world][ldr*ldr]      per-rank Gram partials of the sharded mode
    long long off_srep;                    // double [2][MAX_LEAD][ldr*ldr]   rank 0's Grams of the replicated leading modes
    long long off_m;                       // float [2][world][m_block]       partial M rows, mode k at m_mode_off[k]
    long long off_u;                       // float [2][world][E*ldr]         partial U
    long long off_err;                     // int
    long long m_mode_off[C2C_MAX_LEAD];
    long long m_block;
    long long total;
};

static inline void xg_layout(XgComm& x, int nl, const int64_t* shape /* full or local: mode 0 unused */, long long E, int ldr, int world)
{
    long long off = 0;
    auto take = [&](long long bytes) { const long long o = off; off += (bytes + 255) / 256 * 256; return o; };
    const long long LL = (long long)ldr * ldr;
    x.world = world;
    x.off_err = take(256);
    x.off_flags_m = take((long long)C2C_MAX_LEAD * world * C2C_XG_MAX_CTAS * 4);
    x.off_flags_u = take((long long)world * C2C_XG_MAX_CTAS * 4);
    x.off_s0 = take(2 * world * LL * 8);
    x.off_srep = take(2LL * C2C_MAX_LEAD * LL * 8);
    long long mb = 0;
    for (int k = 0; k < C2C_MAX_LEAD; ++k) x.m_mode_off[k] = 0;
    for (int k = 1; k < nl; ++k) { x.m_mode_off[k] = mb; mb += shape[k] * ldr; }
    x.m_block = mb;
    x.off_m = take(2 * world * mb * 4);
    x.off_u = take(2 * world * E * ldr * 4);
    x.total = off;
}

__device__ __forceinline__ void xg_store_flag(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int xg_load_flag(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// every thread of the CTA has stored its share into the peers' buffers: publish `epoch` in flag `idx` of every peer
__device__ __forceinline__ void xg_signal(const XgComm& x, long long off_flags, int idx, unsigned int epoch)
{
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < x.world && (int)threadIdx.x != x.me) {
        unsigned int* f = reinterpret_cast<unsigned int*>(x.peer[threadIdx.x] + off_flags) + (size_t)x.me * C2C_XG_MAX_CTAS + idx;
        xg_store_flag(f, epoch);
    }
}

// wait until every peer has published `epoch` in flags `idx` (and `idx2` if >= 0) of the local buffer
__device__ __forceinline__ void xg_wait(const XgComm& x, long long off_flags, int idx, int idx2, unsigned int epoch)
{
    if ((int)threadIdx.x < x.world && (int)threadIdx.x != x.me) {
        const unsigned int* base = reinterpret_cast<const unsigned int*>(x.peer[x.me] + off_flags) + (size_t)threadIdx.x * C2C_XG_MAX_CTAS;
        volatile int* err = reinterpret_cast<volatile int*>(x.peer[x.me] + x.off_err);
        const long long t0 = clock64();
        for (int pass = 0; pass < 2; ++pass) {
            const int i = pass == 0 ? idx : idx2;
            if (i < 0) continue;
            while ((int)(xg_load_flag(base + i) - epoch) < 0) {
                if (*err != 0) break;                                     // an earlier wait gave up: do not wait again
                if (clock64() - t0 > C2C_XG_SPIN_CLOCKS) { *err = 1; break; }
                __nanosleep(40);
            }
        }
    }
    __syncthreads();
}

// Replicated leading mode k >= 1 of a context-sharded run (see the file comment).  Same row ownership and thread roles as
// lead_update_kernel (c2c_update.cuh): `tpr` threads per row, blockDim / tpr rows in flight.
__global__ void __launch_bounds__(1024)
lead_update_xg_kernel(const LeadUpdArgs a, const int tpr, const XgComm x)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(a.done)) return;
    extern __shared__ __align__(16) unsigned char lux_smem[];
    const int R = a.R, ldr = a.ldr, q4 = ldr >> 2, LL = ldr * ldr;
    const int nthr = blockDim.x;
    const int nrc = nthr / tpr;
    const int njr = tpr / q4;
    float* sG = reinterpret_cast<float*>(lux_smem);
    float* sNew = sG + R * R;
    double* sRed = reinterpret_cast<double*>(lux_smem + (((size_t)(R * R + nrc * ldr) * sizeof(float) + 7) & ~(size_t)7));
    const int it = a.state[0];
    const int par = it & 1;
    const unsigned int epoch = x.epoch_base + (unsigned int)it + 1u;
    double* Snew = a.S + (size_t)par * a.N * LL;
    const double* Sold = a.S + (size_t)(par ^ 1) * a.N * LL;
    const int k = a.k, nl = a.lead.nl;
    const int dk = a.lead.dim[k];
    int inner = 1;
    for (int m = k + 1; m < nl; ++m) inner *= a.lead.dim[m];
    const int cnt = (int)(a.P / dk);
    const int rslot = threadIdx.x / tpr, lt = threadIdx.x - rslot * tpr;
    const int quad = lt % q4, jl = lt / q4;
    const int i_begin = blockIdx.x * a.rows_per_cta;
    int i_end = i_begin + a.rows_per_cta; if (i_end > dk) i_end = dk;
    const float4* oth = nl == 2 ? reinterpret_cast<const float4*>(a.lead.fac[1 - k]) + quad : nullptr;
    const long long sI = (nl == 2 && k == 0) ? a.lead.dim[1] : 1, sJ = (nl == 2 && k == 1) ? a.lead.dim[1] : 1;
    const size_t m_slot = ((size_t)par * x.world + x.me) * x.m_block + x.m_mode_off[k];

    // ---- phase 1: this rank's partial rows of M_k, stored into every rank's buffer ----
    for (int ib = i_begin; ib < i_end; ib += nrc) {
        const int i = ib + rslot;
        const bool row_on = i < i_end;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_on && jl < njr) {
            if (nl == 2) {
                const float4* tb = reinterpret_cast<const float4*>(a.T) + quad;
                int j = jl;
                for (; j + 3 * njr < cnt; j += 4 * njr) {
                    float4 t[4], f[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const long long p = i * sI + (long long)(j + u * njr) * sJ;
                        t[u] = __ldg(tb + (size_t)p * q4);
                        f[u] = __ldg(oth + (size_t)(j + u * njr) * q4);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc.x = fmaf(t[u].x, f[u].x, acc.x); acc.y = fmaf(t[u].y, f[u].y, acc.y);
                        acc.z = fmaf(t[u].z, f[u].z, acc.z); acc.w = fmaf(t[u].w, f[u].w, acc.w);
                    }
                }
                for (; j < cnt; j += njr) {
                    const long long p = i * sI + (long long)j * sJ;
                    const float4 t = __ldg(tb + (size_t)p * q4);
                    const float4 f = __ldg(oth + (size_t)j * q4);
                    acc.x = fmaf(t.x, f.x, acc.x); acc.y = fmaf(t.y, f.y, acc.y);
                    acc.z = fmaf(t.z, f.z, acc.z); acc.w = fmaf(t.w, f.w, acc.w);
                }
            } else {
                for (int j = jl; j < cnt; j += njr) {
                    const int o = j / inner, n = j - o * inner;
                    const long long p = ((long long)o * dk + i) * inner + n;
                    const float4 t = __ldg(reinterpret_cast<const float4*>(a.T + (size_t)p * ldr) + quad);
                    float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
                    int rem = n;
                    for (int m = nl - 1; m > k; --m) {
                        const int d = a.lead.dim[m]; const int qq = rem / d; const int idx = rem - qq * d; rem = qq;
                        const float4 f = __ldg(reinterpret_cast<const float4*>(a.lead.fac[m] + (size_t)idx * ldr) + quad);
                        w.x *= f.x; w.y *= f.y; w.z *= f.z; w.w *= f.w;
                    }
                    rem = o;
                    for (int m = k - 1; m >= 0; --m) {
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
        if (lt < ldr && row_on) {
            double m = 0.0;
            if (lt < R) for (int g = 0; g < njr; ++g) m += sRed[((size_t)rslot * njr + g) * ldr + lt];
            const float mf = (float)m;
            for (int g = 0; g < x.world; ++g)
                reinterpret_cast<float*>(x.peer[g] + x.off_m)[m_slot + (size_t)i * ldr + lt] = mf;
        }
        __syncthreads();
    }
    // CTA 0 carries the Grams: every rank its share of S_0 (first replicated mode), rank 0 the Gram of the mode before this one
    if (blockIdx.x == 0) {
        if (k == 1) {
            for (int e = threadIdx.x; e < LL; e += nthr) {
                const double v = Snew[e];
                for (int g = 0; g < x.world; ++g)
                    reinterpret_cast<double*>(x.peer[g] + x.off_s0)[((size_t)par * x.world + x.me) * LL + e] = v;
            }
        } else if (x.me == 0) {
            for (int e = threadIdx.x; e < LL; e += nthr) {
                const double v = Snew[(size_t)(k - 1) * LL + e];
                for (int g = 1; g < x.world; ++g)
                    reinterpret_cast<double*>(x.peer[g] + x.off_srep)[((size_t)par * C2C_MAX_LEAD + (k - 1)) * LL + e] = v;
            }
        }
    }
    const long long off_f = x.off_flags_m + (long long)(k - 1) * x.world * C2C_XG_MAX_CTAS * 4;
    xg_signal(x, off_f, blockIdx.x, epoch);
    xg_wait(x, off_f, blockIdx.x, blockIdx.x == 0 ? -1 : 0, epoch);

    // ---- phase 2: world sums in rank order, the update, the Gram share ----
    // Hadamard of the other modes' Grams: S_0 summed over the ranks (k = 1) / rank 0's Gram of mode k - 1 (k >= 2)
    const double* s0 = reinterpret_cast<const double*>(x.peer[x.me] + x.off_s0) + (size_t)par * x.world * LL;
    const double* sp = reinterpret_cast<const double*>(x.peer[x.me] + x.off_srep) + ((size_t)par * C2C_MAX_LEAD + (k - 1)) * LL;
    for (int t = threadIdx.x; t < R * R; t += nthr) {
        const int q = t / R, r = t - q * R, e = q * ldr + r;
        double g = 1.0;
        for (int j = 0; j < a.N; ++j) {
            if (j == k) continue;
            double s;
            if (j == 0 && k == 1) {
                s = 0.0;
                for (int w = 0; w < x.world; ++w) s += __ldcg(s0 + (size_t)w * LL + e);
                if (blockIdx.x == 0) Snew[e] = s;                        // the global Gram, for the later modes and the error
            } else if (j == k - 1 && k >= 2 && x.me != 0) {
                s = __ldcg(sp + e);
                if (blockIdx.x == 0) Snew[(size_t)j * LL + e] = s;
            } else s = (j < k ? Snew : Sold)[(size_t)j * LL + e];
            g *= s;
        }
        sG[t] = (float)g;
    }
    double gacc[4] = {0.0, 0.0, 0.0, 0.0};
    __syncthreads();
    const float* mb = reinterpret_cast<const float*>(x.peer[x.me] + x.off_m) + (size_t)par * x.world * x.m_block + x.m_mode_off[k];
    for (int ib = i_begin; ib < i_end; ib += nrc) {
        const int i = ib + rslot;
        const bool row_on = i < i_end;
        if (lt < R) {
            float nv = 0.0f;
            if (row_on) {
                double m = 0.0;
                for (int w = 0; w < x.world; ++w) m += (double)__ldcg(mb + (size_t)w * x.m_block + (size_t)i * ldr + lt);
                const float* arow = a.A + (size_t)i * ldr;
                float den = 0.0f;
                for (int q = 0; q < R; ++q) den = fmaf(arow[q], sG[q * R + lt], den);
                nv = arow[lt] * fmaxf((float)m, a.eps) / fmaxf(den, a.eps);
            }
            sNew[rslot * ldr + lt] = nv;
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
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = threadIdx.x + u * nthr;
        if (e < R * R) atomicAdd(Snew + (size_t)k * LL + (e / R) * ldr + (e % R), gacc[u]);
    }
}

// Fixed-order sum of the local fiber partials (as fiber_reduce2_kernel), exchanged over peer memory and summed over the
// ranks in rank order into U.  CTA 0 of rank 0 sends the Gram of the last replicated leading mode along.
struct FiberXgArgs {
    const float* part; int nparts; long long n4; int R, ldr, N, nl;
    float* U; double* S; const int* state; const int* done;
};

__global__ void __launch_bounds__(256)
fiber_reduce_xg_kernel(const FiberXgArgs a, const XgComm x)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(a.done)) return;
    __shared__ double red[8][32][4];
    const int lane = threadIdx.x & 31, pg = threadIdx.x >> 5;
    const int it = a.state[0], par = it & 1, LL = a.ldr * a.ldr;
    const unsigned int epoch = x.epoch_base + (unsigned int)it + 1u;
    const size_t u_slot4 = ((size_t)par * x.world + x.me) * a.n4;
    for (long long c0 = (long long)blockIdx.x * 32; c0 < a.n4; c0 += (long long)gridDim.x * 32) {
        const long long i = c0 + lane;
        double s[4] = {0, 0, 0, 0};
        if (i < a.n4) {
            const float4* src = reinterpret_cast<const float4*>(a.part) + i;
#pragma unroll 4
            for (int b = pg; b < a.nparts; b += 8) {
                const float4 v = __ldg(src + (size_t)b * a.n4);
                s[0] += (double)v.x; s[1] += (double)v.y; s[2] += (double)v.z; s[3] += (double)v.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) red[pg][lane][j] = s[j];
        __syncthreads();
        if (i < a.n4) {
            // the eight warps share the stores: warp pg writes to ranks pg, pg + 8, ...
            double t[4] = {0, 0, 0, 0};
            for (int g = 0; g < 8; ++g)
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] += red[g][lane][j];
            const int r0 = (int)(i % (a.ldr >> 2)) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (r0 + j >= a.R) t[j] = 0.0;
            const float4 v = make_float4((float)t[0], (float)t[1], (float)t[2], (float)t[3]);
            for (int g = pg; g < x.world; g += 8) reinterpret_cast<float4*>(x.peer[g] + x.off_u)[u_slot4 + i] = v;
        }
        __syncthreads();
    }
    const int kl = a.nl - 1;                                             // last replicated leading mode (>= 1)
    if (blockIdx.x == 0 && x.me == 0)
        for (int e = threadIdx.x; e < LL; e += blockDim.x) {
            const double v = a.S[((size_t)par * a.N + kl) * LL + e];
            for (int g = 1; g < x.world; ++g)
                reinterpret_cast<double*>(x.peer[g] + x.off_srep)[((size_t)par * C2C_MAX_LEAD + kl) * LL + e] = v;
        }
    xg_signal(x, x.off_flags_u, blockIdx.x, epoch);
    xg_wait(x, x.off_flags_u, blockIdx.x, -1, epoch);
    if (blockIdx.x == 0 && x.me != 0) {
        const double* sp = reinterpret_cast<const double*>(x.peer[x.me] + x.off_srep) + ((size_t)par * C2C_MAX_LEAD + kl) * LL;
        for (int e = threadIdx.x; e < LL; e += blockDim.x) a.S[((size_t)par * a.N + kl) * LL + e] = __ldcg(sp + e);
    }
    const float4* ub = reinterpret_cast<const float4*>(x.peer[x.me] + x.off_u) + (size_t)par * x.world * a.n4;
    // the chunks this CTA exchanged (the flags cover exactly those), one warp per chunk
    for (long long c0 = ((long long)blockIdx.x + (long long)pg * gridDim.x) * 32; c0 < a.n4; c0 += 8LL * gridDim.x * 32) {
        const long long q = c0 + lane;
        if (q < a.n4) {
            double t[4] = {0, 0, 0, 0};
            for (int w = 0; w < x.world; ++w) {
                const float4 v = __ldcg(ub + (size_t)w * a.n4 + q);
                t[0] += (double)v.x; t[1] += (double)v.y; t[2] += (double)v.z; t[3] += (double)v.w;
            }
            reinterpret_cast<float4*>(a.U)[q] = make_float4((float)t[0], (float)t[1], (float)t[2], (float)t[3]);
        }
    }
}
