// Masked MTTKRP as "dense pass + missing-entry correction" (tensorly's imputation, mirrored at
// cell2cell/tensor/coupled_factorization.py:342-345: tensor = tensor*mask + cp_to_tensor(cp, mask=1-mask), then
// unfolding_dot_khatri_rao).  The contraction is linear in the tensor, so with X zero at its missing entries
//
//     T[p, :]  = sum_e Xz[p, e] B[e, :]        +  sum_{e missing in plane p}  rec[p, e] B[e, :]        (leading modes)
//     U[e, :]  = sum_p Xz[p, e] W[p, :]        +  sum_{p missing at e}        rec[p, e] W[p, :]        (trailing modes)
//
// with B[e, :] = A2[s, :] * A3[v, :], W[p, :] the Khatri-Rao row of the leading factors and rec[p, e] = <W[p, :], B[e, :]>.
// The first terms are the UNMASKED streaming passes (4 bytes per entry at HBM speed; they do not depend on the factors of
// their own side, so one plane pass serves every leading mode of an iteration and one fiber pass both trailing modes); the
// second terms are the kernels below: they stream the mask only (1 byte per entry) and do 2R FMAs per MISSING entry.
//   plane_corr<NQ>   a warp per plane: lanes stride the plane's mask bytes, B rows come from the L1-resident Bmat,
//                    R partial sums per lane, shuffle tree, Tw[p, :] = T0[p, :] + correction (a plain copy for complete planes)
//   fiber_corr<NQ>   a CTA per contiguous plane range: thread t owns in-plane positions t, t + 512, ... and their rows of
//                    the CTA's U correction in shared memory (exclusive: no atomics); the Khatri-Rao rows of 16 planes at a
//                    time are staged in shared memory and read as broadcasts; one partial per CTA, summed in a fixed order
//                    onto U0 by corr_reduce
// Both are exact restatements of the imputation (no approximation): the same products in a different association.
#include "c2c_common.cuh"
#include "c2c_corr.cuh"

// Khatri-Rao weight of plane p, column r, with 32-bit index arithmetic (the corrections run with P < 2^31)
__device__ __forceinline__ float lead_weight32(unsigned p, const LeadInfo& lead, int ldr, int r)
{
    float w = 1.0f;
    unsigned rem = p;
    for (int k = lead.nl - 1; k >= 0; --k) {
        const unsigned d = (unsigned)lead.dim[k];
        const unsigned q = rem / d;
        const unsigned i = rem - q * d;
        rem = q;
        w *= __ldg(lead.fac[k] + (size_t)i * ldr + r);
    }
    return w;
}

#define C2C_CORR_G 16            // planes per group (fiber_corr) / mask bytes in flight per lane
#define C2C_CORR_THREADS 512

__global__ void __launch_bounds__(256)
masked_nonzero_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask, long long n, int* __restrict__ flag)
{
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        bad |= (mask[i] == 0 && X[i] != 0.0f) ? 1 : 0;
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

template <int NQ>
__global__ void __launch_bounds__(256)
plane_corr_kernel(const uint8_t* __restrict__ mask, const LeadInfo lead, const float* __restrict__ Bmat,
                  const float* __restrict__ T0, float* __restrict__ Tw, long long P, int E, int R, const int* done)
{
    if (run_done(done)) return;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < P; p += nwarps) {
        const float wl = lane < R ? lead_weight32((unsigned)p, lead, NQ, lane) : 0.0f;
        float w[NQ], acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { w[q] = __shfl_sync(0xffffffffu, wl, q); acc[q] = 0.0f; }
        const uint8_t* mp = mask + (size_t)p * E;
        for (int e0 = 0; e0 < E; e0 += 32 * C2C_CORR_G) {
            unsigned char m[C2C_CORR_G];
#pragma unroll
            for (int i = 0; i < C2C_CORR_G; ++i) {
                const int e = e0 + i * 32 + lane;
                m[i] = e < E ? __ldg(mp + e) : (unsigned char)1;
            }
#pragma unroll
            for (int i = 0; i < C2C_CORR_G; ++i) {
                if (m[i] == 0) {
                    const int e = e0 + i * 32 + lane;
                    const float4* b4 = reinterpret_cast<const float4*>(Bmat + (size_t)e * NQ);
                    float b[NQ];
#pragma unroll
                    for (int q4 = 0; q4 < NQ / 4; ++q4) {
                        const float4 t = __ldg(b4 + q4);
                        b[4 * q4] = t.x; b[4 * q4 + 1] = t.y; b[4 * q4 + 2] = t.z; b[4 * q4 + 3] = t.w;
                    }
                    float rec = 0.0f;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) rec = fmaf(w[q], b[q], rec);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) acc[q] = fmaf(rec, b[q], acc[q]);
                }
            }
        }
        float mine = 0.0f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float s = warp_sum(acc[q]);
            if (lane == q) mine = s;
        }
        if (lane < NQ) Tw[(size_t)p * NQ + lane] = T0[(size_t)p * NQ + lane] + mine;
    }
}


// bytes that are zero in a 32-bit word -> bit 7 of that byte set, exactly (no borrow between bytes, unlike the
// (v - 0x01010101) & ~v test, which also flags a 0x01 byte above a zero byte)
__device__ __forceinline__ unsigned zero_bytes(unsigned v)
{
    const unsigned t = (v & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | v | 0x7f7f7f7fu);
}

// in-plane size a multiple of 16 (and < 65536): a lane reads 16 mask bytes per load (the whole plane's mask is in flight per
// warp); the positions of the zero bytes are compacted into a per-warp queue in shared memory (ballot-free: popcount +
// warp prefix sum), then visited 32 at a time -- every lane busy whatever the pattern of the missing entries.
// (Prefetching the next plane's mask words into a second register set was measured: slower, 348 vs 287 us -- fewer warps.)
template <int NQ>
__global__ void __launch_bounds__(256)
plane_corr16_kernel(const uint8_t* __restrict__ mask, const LeadInfo lead, const float* __restrict__ Bmat,
                    const float* __restrict__ T0, float* __restrict__ Tw, long long P, int E, int R, const int* done)
{
    if (run_done(done)) return;
    extern __shared__ __align__(16) unsigned short queue_all[];      // [warps per CTA][E]
    const int lane = threadIdx.x & 31;
    unsigned short* queue = queue_all + (size_t)(threadIdx.x >> 5) * E;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int n16 = E >> 4;
    const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    for (long long p = warp; p < P; p += nwarps) {
        const uint4* mp = reinterpret_cast<const uint4*>(mask + (size_t)p * E);
        uint4 mm[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = i * 32 + lane;
            mm[i] = c < n16 ? __ldg(mp + c) : ones;
        }
        const float wl = lane < R ? lead_weight32((unsigned)p, lead, NQ, lane) : 0.0f;
        int count = 0;
        for (int c0 = 0; c0 < n16; c0 += 128) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned z0 = zero_bytes(mm[i].x), z1 = zero_bytes(mm[i].y), z2 = zero_bytes(mm[i].z), z3 = zero_bytes(mm[i].w);
                // bit j of zm <=> byte j of the lane's 16 bytes is zero
                auto nib = [](unsigned z) { return ((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u); };
                unsigned zm = nib(z0) | (nib(z1) << 4) | (nib(z2) << 8) | (nib(z3) << 12);
                const int n = __popc(zm);
                int incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total) {
                    int off = count + incl - n;
                    const int ebase = (c0 + i * 32 + lane) << 4;
                    while (zm) {
                        queue[off++] = (unsigned short)(ebase + __ffs(zm) - 1);
                        zm &= zm - 1;
                    }
                    count += total;
                }
            }
            if (c0 + 128 < n16) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = c0 + 128 + i * 32 + lane;
                    mm[i] = c < n16 ? __ldg(mp + c) : ones;
                }
            }
        }
        float mine = 0.0f;
        if (count) {                                                 // warp-uniform
            float w[NQ], acc[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) { w[q] = __shfl_sync(0xffffffffu, wl, q); acc[q] = 0.0f; }
            __syncwarp();
            for (int i = lane; i < count; i += 32) {
                const int e = queue[i];
                const float4* b4 = reinterpret_cast<const float4*>(Bmat + (size_t)e * NQ);
                float b[NQ];
#pragma unroll
                for (int q4 = 0; q4 < NQ / 4; ++q4) {
                    const float4 t = __ldg(b4 + q4);
                    b[4 * q4] = t.x; b[4 * q4 + 1] = t.y; b[4 * q4 + 2] = t.z; b[4 * q4 + 3] = t.w;
                }
                float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
                for (int q = 0; q < NQ; q += 2) { r0 = fmaf(w[q], b[q], r0); r1 = fmaf(w[q + 1], b[q + 1], r1); }
                const float rec = r0 + r1;
#pragma unroll
                for (int q = 0; q < NQ; ++q) acc[q] = fmaf(rec, b[q], acc[q]);
            }
            __syncwarp();                                            // the queue is free for the warp's next plane
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float s = warp_sum(acc[q]);
                if (lane == q) mine = s;
            }
        }
        if (lane < NQ) Tw[(size_t)p * NQ + lane] = T0[(size_t)p * NQ + lane] + mine;
    }
}

// in-plane size a multiple of 4: thread t owns the four positions 4t .. 4t+3 (and their rows of Us); it reads their mask bytes
// of 32 planes as 32 independent 32-bit loads, then visits the (position, plane) pairs that are missing
#define C2C_CORR_G4 32
// THREADS = 256: two CTAs per SM (one CTA's barriers / staging run under the other's loads) when two corrections fit shared
// memory; THREADS = 512: one CTA per SM
template <int NQ, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 1)
fiber_corr4_kernel(const uint8_t* __restrict__ mask, const LeadInfo lead, const float* __restrict__ Bmat,
                   float* __restrict__ part, long long P, int E, int R, int pitch, long long ppc, const int* done)
{
    if (run_done(done)) return;
    extern __shared__ __align__(16) float smem_f[];
    float* Us = smem_f;                                              // [E][pitch]
    float* sW = smem_f + (size_t)E * pitch;                          // [G4][NQ]
    const int t = threadIdx.x;
    for (int i = t; i < E * pitch; i += THREADS) Us[i] = 0.0f;
    const long long p_begin = (long long)blockIdx.x * ppc;
    const long long p_end = p_begin + ppc < P ? p_begin + ppc : P;
    const int n4 = E >> 2;
    for (long long pg = p_begin; pg < p_end; pg += C2C_CORR_G4) {
        __syncthreads();
        for (int i = t; i < C2C_CORR_G4 * NQ; i += THREADS) {
            const int gq = i / NQ, q = i - gq * NQ;
            const long long p = pg + gq;
            sW[i] = (p < p_end && q < R) ? lead_weight32((unsigned)p, lead, NQ, q) : 0.0f;
        }
        __syncthreads();
        for (int c = t; c < n4; c += THREADS) {
            unsigned zm[4] = {0u, 0u, 0u, 0u};                       // per owned position: bit g set <=> missing in plane pg + g
            {
                // the scan is the bulk of the instructions when few entries are missing (ncu: 46 per mask word before): one
                // pointer per thread, whole groups without per-plane bound checks, complete words (0x01010101) skipped at once
                unsigned m[C2C_CORR_G4];
                const unsigned* mrow = reinterpret_cast<const unsigned*>(mask + (size_t)pg * E) + c;
                const size_t wstride = (size_t)(E >> 2);
                const int ng = (int)(p_end - pg < C2C_CORR_G4 ? p_end - pg : C2C_CORR_G4);
                if (ng == C2C_CORR_G4) {
#pragma unroll
                    for (int gq = 0; gq < C2C_CORR_G4; ++gq) m[gq] = __ldg(mrow + gq * wstride);
                } else {
#pragma unroll
                    for (int gq = 0; gq < C2C_CORR_G4; ++gq) m[gq] = gq < ng ? __ldg(mrow + gq * wstride) : 0x01010101u;
                }
#pragma unroll
                for (int gq = 0; gq < C2C_CORR_G4; ++gq) {
                    if (m[gq] == 0x01010101u) continue;
                    const unsigned z = zero_bytes(m[gq]);
                    zm[0] |= ((z >> 7) & 1u) << gq; zm[1] |= ((z >> 15) & 1u) << gq;
                    zm[2] |= ((z >> 23) & 1u) << gq; zm[3] |= ((z >> 31) & 1u) << gq;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                unsigned z = zm[j];
                if (z == 0u) continue;
                const int e = 4 * c + j;
                float b[NQ], a[NQ];
                const float4* b4 = reinterpret_cast<const float4*>(Bmat + (size_t)e * NQ);
                float4* u4 = reinterpret_cast<float4*>(Us + (size_t)e * pitch);
#pragma unroll
                for (int q4 = 0; q4 < NQ / 4; ++q4) {
                    const float4 tb = __ldg(b4 + q4);
                    b[4 * q4] = tb.x; b[4 * q4 + 1] = tb.y; b[4 * q4 + 2] = tb.z; b[4 * q4 + 3] = tb.w;
                    const float4 ta = u4[q4];
                    a[4 * q4] = ta.x; a[4 * q4 + 1] = ta.y; a[4 * q4 + 2] = ta.z; a[4 * q4 + 3] = ta.w;
                }
                while (z) {
                    const int gq = __ffs(z) - 1;
                    z &= z - 1;
                    const float4* w4 = reinterpret_cast<const float4*>(sW + gq * NQ);
                    float w[NQ];
#pragma unroll
                    for (int q4 = 0; q4 < NQ / 4; ++q4) {
                        const float4 tw = w4[q4];
                        w[4 * q4] = tw.x; w[4 * q4 + 1] = tw.y; w[4 * q4 + 2] = tw.z; w[4 * q4 + 3] = tw.w;
                    }
                    float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
                    for (int q = 0; q < NQ; q += 2) { r0 = fmaf(w[q], b[q], r0); r1 = fmaf(w[q + 1], b[q + 1], r1); }
                    const float rec = r0 + r1;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) a[q] = fmaf(rec, w[q], a[q]);
                }
#pragma unroll
                for (int q4 = 0; q4 < NQ / 4; ++q4) u4[q4] = make_float4(a[4 * q4], a[4 * q4 + 1], a[4 * q4 + 2], a[4 * q4 + 3]);
            }
        }
    }
    __syncthreads();
    float* out = part + (size_t)blockIdx.x * E * NQ;
    for (int i = t; i < E * NQ; i += THREADS) {
        const int e = i / NQ, q = i - e * NQ;
        out[i] = Us[(size_t)e * pitch + q];
    }
}

template <int NQ>
__global__ void __launch_bounds__(C2C_CORR_THREADS, 1)
fiber_corr_kernel(const uint8_t* __restrict__ mask, const LeadInfo lead, const float* __restrict__ Bmat,
                  float* __restrict__ part, long long P, int E, int R, int pitch, long long ppc, const int* done)
{
    if (run_done(done)) return;
    extern __shared__ __align__(16) float smem_f[];
    float* Us = smem_f;                                              // [E][pitch]
    float* sW = smem_f + (size_t)E * pitch;                          // [G][NQ]
    const int t = threadIdx.x;
    for (int i = t; i < E * pitch; i += C2C_CORR_THREADS) Us[i] = 0.0f;
    const long long p_begin = (long long)blockIdx.x * ppc;
    const long long p_end = p_begin + ppc < P ? p_begin + ppc : P;
    for (long long pg = p_begin; pg < p_end; pg += C2C_CORR_G) {
        __syncthreads();                                             // the previous group's rows are no longer read (Us cleared)
        for (int i = t; i < C2C_CORR_G * NQ; i += C2C_CORR_THREADS) {
            const int gq = i / NQ, q = i - gq * NQ;
            const long long p = pg + gq;
            sW[i] = (p < p_end && q < R) ? lead_weight32((unsigned)p, lead, NQ, q) : 0.0f;
        }
        __syncthreads();
        for (int e = t; e < E; e += C2C_CORR_THREADS) {
            unsigned char m[C2C_CORR_G];
            bool any = false;
#pragma unroll
            for (int gq = 0; gq < C2C_CORR_G; ++gq) {
                m[gq] = (pg + gq < p_end) ? __ldg(mask + (size_t)(pg + gq) * E + e) : (unsigned char)1;
                any = any || (m[gq] == 0);
            }
            if (any) {
                float b[NQ], a[NQ];
                const float4* b4 = reinterpret_cast<const float4*>(Bmat + (size_t)e * NQ);
                float4* u4 = reinterpret_cast<float4*>(Us + (size_t)e * pitch);
#pragma unroll
                for (int q4 = 0; q4 < NQ / 4; ++q4) {
                    const float4 tb = __ldg(b4 + q4);
                    b[4 * q4] = tb.x; b[4 * q4 + 1] = tb.y; b[4 * q4 + 2] = tb.z; b[4 * q4 + 3] = tb.w;
                    const float4 ta = u4[q4];
                    a[4 * q4] = ta.x; a[4 * q4 + 1] = ta.y; a[4 * q4 + 2] = ta.z; a[4 * q4 + 3] = ta.w;
                }
#pragma unroll
                for (int gq = 0; gq < C2C_CORR_G; ++gq) {
                    if (m[gq] == 0) {
                        const float4* w4 = reinterpret_cast<const float4*>(sW + gq * NQ);
                        float w[NQ];
#pragma unroll
                        for (int q4 = 0; q4 < NQ / 4; ++q4) {
                            const float4 tw = w4[q4];
                            w[4 * q4] = tw.x; w[4 * q4 + 1] = tw.y; w[4 * q4 + 2] = tw.z; w[4 * q4 + 3] = tw.w;
                        }
                        float rec = 0.0f;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) rec = fmaf(w[q], b[q], rec);
#pragma unroll
                        for (int q = 0; q < NQ; ++q) a[q] = fmaf(rec, w[q], a[q]);
                    }
                }
#pragma unroll
                for (int q4 = 0; q4 < NQ / 4; ++q4) u4[q4] = make_float4(a[4 * q4], a[4 * q4 + 1], a[4 * q4 + 2], a[4 * q4 + 3]);
            }
        }
    }
    __syncthreads();
    float* out = part + (size_t)blockIdx.x * E * NQ;
    for (int i = t; i < E * NQ; i += C2C_CORR_THREADS) {
        const int e = i / NQ, q = i - e * NQ;
        out[i] = Us[(size_t)e * pitch + q];
    }
}

// Uw = U0 + sum of the CTA partials, in a fixed order
__global__ void __launch_bounds__(256)
corr_reduce_kernel(const float* __restrict__ part, int nparts, long long n, const float* __restrict__ U0, float* __restrict__ Uw,
                   const int* done)
{
    if (run_done(done)) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.0f;
    for (int c = 0; c < nparts; ++c) s += part[(size_t)c * n + i];
    Uw[i] = U0[i] + s;
}

static inline int corr_pitch(int ldr) { return ((ldr / 4) & 1) ? ldr : ldr + 4; }   // odd number of 16-byte units per row

size_t c2c_fiber_corr_smem(long long E, int ldr)
{
    return ((size_t)E * corr_pitch(ldr) + (size_t)C2C_CORR_G4 * ldr) * sizeof(float);
}

cudaError_t c2c_launch_masked_nonzero(const float* X, const uint8_t* mask, long long n, int* flag, int nsm, cudaStream_t st)
{
    masked_nonzero_kernel<<<nsm * 8, 256, 0, st>>>(X, mask, n, flag);
    return cudaGetLastError();
}

#define C2C_CORR_DISPATCH(CALL)                                                                                          \
    switch (ldr) {                                                                                                       \
        case 4: CALL(4); break; case 8: CALL(8); break; case 12: CALL(12); break; case 16: CALL(16); break;              \
        case 20: CALL(20); break; case 24: CALL(24); break; case 28: CALL(28); break; case 32: CALL(32); break;          \
        default: return cudaErrorNotSupported;                                                                           \
    }

cudaError_t c2c_launch_plane_corr(const uint8_t* mask, const LeadInfo& lead, const float* Bmat, const float* T0, float* Tw,
                                  long long P, long long E, int R, int ldr, const int* done, int nsm, cudaStream_t st)
{
    if (E >= (1LL << 31)) return cudaErrorNotSupported;
    long long grid = (P + 7) / 8;
    if (grid > (long long)nsm * 8) grid = (long long)nsm * 8;
    if ((E & 15) == 0 && ((uintptr_t)mask & 15) == 0 && E <= 8192) {
        const size_t qsmem = (size_t)8 * E * sizeof(unsigned short);                      // <= 128 KB; > 48 KB needs the opt-in
#define PC16_CALL(N)                                                                                                     \
    do {                                                                                                                 \
        if (qsmem > 48 * 1024) {                                                                                         \
            const cudaError_t ea = cudaFuncSetAttribute(plane_corr16_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                        128 * 1024);                                                     \
            if (ea != cudaSuccess) return ea;                                                                            \
        }                                                                                                                \
        plane_corr16_kernel<N><<<(unsigned)grid, 256, qsmem, st>>>(mask, lead, Bmat, T0, Tw, P, (int)E, R, done);        \
    } while (0)
        C2C_CORR_DISPATCH(PC16_CALL)
#undef PC16_CALL
        return cudaGetLastError();
    }
#define PC_CALL(N) plane_corr_kernel<N><<<(unsigned)grid, 256, 0, st>>>(mask, lead, Bmat, T0, Tw, P, (int)E, R, done)
    C2C_CORR_DISPATCH(PC_CALL)
#undef PC_CALL
    return cudaGetLastError();
}

cudaError_t c2c_launch_fiber_corr(const uint8_t* mask, const LeadInfo& lead, const float* Bmat, const float* U0, float* Uw,
                                  float* part, int max_parts, long long P, long long E, int R, int ldr, const int* done,
                                  cudaStream_t st)
{
    if (E >= (1LL << 31)) return cudaErrorNotSupported;
    const size_t smem = c2c_fiber_corr_smem(E, ldr);
    if (smem > C2C_CORR_SMEM_MAX) return cudaErrorNotSupported;
    const bool two_per_sm = (E & 3) == 0 && ((uintptr_t)mask & 3) == 0 && smem <= 100 * 1024;
    long long grid = two_per_sm ? max_parts : max_parts / 2;
    if (grid < 1) grid = 1;
    if (grid > P) grid = P;
    const long long ppc = (P + grid - 1) / grid;
    grid = (P + ppc - 1) / ppc;
    const int pitch = corr_pitch(ldr);
#define FC_CALL(N)                                                                                                       \
    do {                                                                                                                 \
        if (smem > 48 * 1024) {                                                                                          \
            const cudaError_t ea = cudaFuncSetAttribute(fiber_corr_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                        (int)C2C_CORR_SMEM_MAX);                                         \
            if (ea != cudaSuccess) return ea;                                                                            \
        }                                                                                                                \
        fiber_corr_kernel<N><<<(unsigned)grid, C2C_CORR_THREADS, smem, st>>>(mask, lead, Bmat, part, P, (int)E, R, pitch, ppc, done); \
    } while (0)
#define FC4_CALL_T(N, T)                                                                                                 \
    do {                                                                                                                 \
        if (smem > 48 * 1024) {                                                                                          \
            const cudaError_t ea = cudaFuncSetAttribute(fiber_corr4_kernel<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                        (int)C2C_CORR_SMEM_MAX);                                         \
            if (ea != cudaSuccess) return ea;                                                                            \
        }                                                                                                                \
        fiber_corr4_kernel<N, T><<<(unsigned)grid, T, smem, st>>>(mask, lead, Bmat, part, P, (int)E, R, pitch, ppc, done); \
    } while (0)
#define FC4_CALL(N) do { if (two_per_sm) FC4_CALL_T(N, 256); else FC4_CALL_T(N, 512); } while (0)
    if ((E & 3) == 0 && ((uintptr_t)mask & 3) == 0) {
        C2C_CORR_DISPATCH(FC4_CALL)
    } else {
        C2C_CORR_DISPATCH(FC_CALL)
    }
#undef FC4_CALL
#undef FC4_CALL_T
#undef FC_CALL
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const long long n = E * ldr;
    corr_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, (int)grid, n, U0, Uw, done);
    return cudaGetLastError();
}
