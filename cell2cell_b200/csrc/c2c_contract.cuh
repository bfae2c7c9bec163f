// The three streaming kernels that read the whole tensor (one HBM pass each).  Everything else in the factorization
// works on data that is O(sum I_n * R) or O(P * R) and lives in L2.
//
//   plane_contract<VEC,NP2,MASKED>  T[p, r]  = sum_{s,v} Xh[p,s,v] * A2[s,r] * A3[v,r]          (serves the leading modes)
//   fiber_contract<VEC,NP2,MASKED>  U[s,v,r] = sum_p     Xh[p,s,v] * w[p,r],  w = prod_lead A_k   (serves the trailing modes)
//   recon_sums<VEC,NP2>             sum m(x-rec), sum m(x-rec)^2, sum m x, sum m x^2, sum m       (final masked loss)
//
// Xh = X where mask = 1 and the current reconstruction sum_q w[p,q] A2[s,q] A3[v,q] where mask = 0 -- tensorly's
// `tensor*mask + cp_to_tensor(..., mask=1-mask)` (mirrored at cell2cell/tensor/coupled_factorization.py:342-343) fused
// into the same read instead of materialising an N_X x R Khatri-Rao product.
// They replace tensorly's unfolding_dot_khatri_rao (R chains of mode-vector products, each re-reading X; mirrored at
// cell2cell/tensor/coupled_factorization.py:345), reached from cell2cell/tensor/factorization.py:92-101.
#pragma once
#include "c2c_common.cuh"

template <int VEC> struct XVec;
template <> struct XVec<4> { typedef float4 T; typedef uchar4 M; };
template <> struct XVec<2> { typedef float2 T; typedef uchar2 M; };
template <> struct XVec<1> { typedef float  T; typedef unsigned char M; };

template <int VEC> __device__ __forceinline__ void ld_x(float (&x)[VEC], const float* p);
template <> __device__ __forceinline__ void ld_x<4>(float (&x)[4], const float* p) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p)); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
template <> __device__ __forceinline__ void ld_x<2>(float (&x)[2], const float* p) {
    const float2 v = __ldcs(reinterpret_cast<const float2*>(p)); x[0] = v.x; x[1] = v.y;
}
template <> __device__ __forceinline__ void ld_x<1>(float (&x)[1], const float* p) { x[0] = __ldcs(p); }

template <int VEC> __device__ __forceinline__ unsigned ld_m(const uint8_t* p);   // returns VEC mask bytes packed, byte j = element j
template <> __device__ __forceinline__ unsigned ld_m<4>(const uint8_t* p) { return __ldcs(reinterpret_cast<const unsigned*>(p)); }
template <> __device__ __forceinline__ unsigned ld_m<2>(const uint8_t* p) { return __ldcs(reinterpret_cast<const unsigned short*>(p)); }
template <> __device__ __forceinline__ unsigned ld_m<1>(const uint8_t* p) { return __ldcs(p); }

#ifndef C2C_PC_THREADS
#define C2C_PC_THREADS 256      // plane_contract / recon_sums block size
#endif
#ifndef C2C_PC_UNROLL
#define C2C_PC_UNROLL 4         // rows of a plane loaded ahead per thread
#endif
#ifndef C2C_PC_MINB
#define C2C_PC_MINB 1           // resident CTAs per SM the register allocation must allow
#endif

// --------------------------------------------------------------------------------------------------------------------
// plane_contract: a group of G = min(I3/VEC, 32) lanes owns one plane; lane `lig` owns the VEC-wide column strips
// lig, lig+G, ...; it walks the I2 rows of the strip (independent 16-byte loads, C2C_PC_UNROLL in flight), doing
// acc[j][r] += x[s, v+j] * A2[s, r] with A2 rows broadcast from shared memory (stored as (a,a) pairs so FFMA2 can pair
// two columns j), then folds A3[v+j, r] in once per strip and reduces over the group's lanes with shuffles.
// Shared memory: A2 duplicated [I2][4*NP2] + A3 [I3][2*NP2].
// --------------------------------------------------------------------------------------------------------------------
template <int VEC, int NP2, bool MASKED, bool SUMS>
__global__ void __launch_bounds__(C2C_PC_THREADS, C2C_PC_MINB)
plane_contract_kernel(const ContractArgs a)
{
    if (run_done(a.done)) return;
    constexpr int RC = 2 * NP2;                 // rank columns handled (padded to even)
    extern __shared__ __align__(16) float smem[];
    float* sA2 = smem;                          // [I2][2*RC]  (a_r, a_r) pairs
    float* sA3 = smem + (size_t)a.I2 * 2 * RC;  // [I3][RC]
    const int I2 = a.I2, I3 = a.I3, ldr = a.ldr, r0 = a.r0, R = a.R;
    for (int i = threadIdx.x; i < I2 * RC; i += blockDim.x) {
        const int s = i / RC, r = i - s * RC;
        const float v = (r0 + r < R) ? a.A2[(size_t)s * ldr + r0 + r] : 0.0f;
        sA2[2 * i] = v; sA2[2 * i + 1] = v;
    }
    for (int i = threadIdx.x; i < I3 * RC; i += blockDim.x) {
        const int s = i / RC, r = i - s * RC;
        sA3[i] = (r0 + r < R) ? a.A3[(size_t)s * ldr + r0 + r] : 0.0f;
    }
    __syncthreads();

    const int nstrips = I3 / VEC;
    const int G = nstrips < 32 ? nstrips : 32;
    const int ppw = 32 / G;
    const int lane = threadIdx.x & 31;
    const int grp = lane / G, lig = lane - grp * G;
    const bool lane_on = grp < ppw;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t E = (size_t)I2 * I3;

    double s_d = 0, s_dd = 0, s_x = 0, s_xx = 0, s_n = 0;   // SUMS only

    for (long long base = a.p0 + wg * ppw; base < a.P; base += nw * ppw) {
        const long long p = base + grp;
        const bool on = lane_on && p < a.P;
        float out[RC];
#pragma unroll
        for (int r = 0; r < RC; ++r) out[r] = 0.0f;
        float w[RC];
        if (MASKED || SUMS) {
            if (on) lead_weights<RC>(w, p, a.lead, ldr, r0, R);
        }
        if (on) {
            for (int strip = lig; strip < nstrips; strip += G) {
                const int v0 = strip * VEC;
                float2 acc[(VEC + 1) / 2][RC];
#pragma unroll
                for (int j = 0; j < (VEC + 1) / 2; ++j)
#pragma unroll
                    for (int r = 0; r < RC; ++r) acc[j][r] = make_float2(0.0f, 0.0f);
                float wa[(MASKED || SUMS) ? VEC : 1][RC];
                if (MASKED || SUMS) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
#pragma unroll
                        for (int r = 0; r < RC; ++r) wa[j][r] = w[r] * sA3[(v0 + j) * RC + r];
                }
                const float* xp = a.X + (size_t)p * E + v0;
                const uint8_t* mp = (MASKED || SUMS) && a.mask ? a.mask + (size_t)p * E + v0 : nullptr;
                for (int s0 = 0; s0 < I2; s0 += C2C_PC_UNROLL) {
                    float x[C2C_PC_UNROLL][VEC];
                    unsigned m[C2C_PC_UNROLL];
#pragma unroll
                    for (int u = 0; u < C2C_PC_UNROLL; ++u) {
                        if (s0 + u < I2) {
                            ld_x<VEC>(x[u], xp + (size_t)(s0 + u) * I3);
                            m[u] = mp ? ld_m<VEC>(mp + (size_t)(s0 + u) * I3) : 0x01010101u;
                        } else {
#pragma unroll
                            for (int j = 0; j < VEC; ++j) x[u][j] = 0.0f;
                            m[u] = 0x01010101u;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < C2C_PC_UNROLL; ++u) {
                        const int s = s0 + u;
                        if (s < I2) {
                            const float4* a2 = reinterpret_cast<const float4*>(sA2 + (size_t)s * 2 * RC);
                            if (MASKED || SUMS) {
                                constexpr unsigned full = VEC == 4 ? 0x01010101u : (VEC == 2 ? 0x0101u : 0x01u);
                                if (SUMS || (m[u] & full) != full) {
                                    // reconstruction of the VEC entries of this row: sum_q wa[j][q] * A2[s][q]
                                    float rec[VEC];
#pragma unroll
                                    for (int j = 0; j < VEC; ++j) rec[j] = 0.0f;
#pragma unroll
                                    for (int q = 0; q < NP2; ++q) {
                                        const float4 t = a2[q];
#pragma unroll
                                        for (int j = 0; j < VEC; ++j) {
                                            rec[j] = fmaf(wa[j][2 * q], t.x, rec[j]);
                                            rec[j] = fmaf(wa[j][2 * q + 1], t.z, rec[j]);
                                        }
                                    }
#pragma unroll
                                    for (int j = 0; j < VEC; ++j) {
                                        const bool obs = (m[u] >> (8 * j)) & 1u;
                                        if (SUMS) {
                                            if (obs != (a.sums_missing != 0)) {
                                                const double xv = a.sums_missing ? 0.0 : (double)x[u][j], d = xv - (double)rec[j];
                                                s_d += d; s_dd += d * d; s_x += xv; s_xx += xv * xv; s_n += 1.0;
                                            }
                                        } else if (!obs) {
                                            x[u][j] = rec[j];
                                        }
                                    }
                                }
                            }
                            if (!SUMS) {
                                if (VEC >= 2) {
#pragma unroll
                                    for (int q = 0; q < NP2; ++q) {
                                        const float4 t = a2[q];
#pragma unroll
                                        for (int j = 0; j < VEC / 2; ++j) {
                                            const float2 xx = make_float2(x[u][2 * j], x[u][2 * j + 1]);
                                            acc[j][2 * q]     = ffma2(xx, make_float2(t.x, t.y), acc[j][2 * q]);
                                            acc[j][2 * q + 1] = ffma2(xx, make_float2(t.z, t.w), acc[j][2 * q + 1]);
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int q = 0; q < NP2; ++q) {
                                        const float4 t = a2[q];
                                        acc[0][2 * q].x     = fmaf(x[u][0], t.x, acc[0][2 * q].x);
                                        acc[0][2 * q + 1].x = fmaf(x[u][0], t.z, acc[0][2 * q + 1].x);
                                    }
                                }
                            }
                        }
                    }
                }
                if (!SUMS) {
                    // fold the receiver factor: out[r] += sum_j acc_j[r] * A3[v0+j, r]
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        const float* a3 = sA3 + (v0 + j) * RC;
#pragma unroll
                        for (int r = 0; r < RC; ++r) {
                            const float av = (j & 1) ? acc[j / 2][r].y : acc[j / 2][r].x;
                            out[r] = fmaf(av, a3[r], out[r]);
                        }
                    }
                }
            }
        }
        if (!SUMS) {
            // segmented reduction over the G lanes of the group (all 32 lanes take part in the shuffles)
#pragma unroll
            for (int r = 0; r < RC; ++r) {
                float v = out[r];
                int width = G;                     // lanes [0, width) of the group still hold live partial sums
                for (int o = 16; o > 0; o >>= 1) {
                    const float t = __shfl_down_sync(0xffffffffu, v, o);
                    if (o < width) {
                        if (lig + o < width) v += t;
                        width = o;
                    }
                }
                out[r] = v;
            }
            if (on && lig == 0) {
                float* t = a.out + (size_t)p * ldr + r0;
#pragma unroll
                for (int r = 0; r < RC; ++r)
                    if (r0 + r < R) t[r] = out[r];
            }
        }
    }
    if (SUMS) {
        // block reduction of the five sums -> a.sums[blockIdx.x * 5 + k] (summed again, in fixed order, by one CTA)
        __shared__ double red[5][C2C_PC_THREADS / 32];
        double v[5] = {s_d, s_dd, s_x, s_xx, s_n};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double t = warp_sum(v[k]);
            if (lane == 0) red[k][threadIdx.x >> 5] = t;
        }
        __syncthreads();
        if (threadIdx.x < 5) {
            double t = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
            a.sums[(size_t)blockIdx.x * 5 + threadIdx.x] = t;
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// fiber_contract: thread (slot, pos) owns the VEC consecutive in-plane positions pos*VEC.. of every plane that its slot
// visits; the CTA owns a contiguous range of planes.  Per plane: one 16-byte load + NP2 broadcast LDS.128 of the plane's
// lead weights (stored as (w,w) pairs) + VEC*NP2 FFMA2.  Accumulators stay in registers for the whole range and are
// written once to the partials buffer [gridDim.x * nslots][E][ldr]; fiber_reduce sums the partials in a fixed order
// (deterministic, no atomics).
// --------------------------------------------------------------------------------------------------------------------
#ifndef C2C_FC_TILE
#define C2C_FC_TILE 32          // planes per shared-memory tile of lead weights
#endif
#ifndef C2C_FC_UNROLL
#define C2C_FC_UNROLL 4
#endif
#ifndef C2C_FC_THREADS_LO
#define C2C_FC_THREADS_LO 512   // block size for NP2 <= 8
#endif
#ifndef C2C_FC_MINB
#define C2C_FC_MINB 1
#endif
#define C2C_FC_MAXTHREADS(NP2) ((NP2) > 8 ? 256 : C2C_FC_THREADS_LO)

template <int VEC, int NP2, bool MASKED>
__global__ void __launch_bounds__(C2C_FC_MAXTHREADS(NP2), C2C_FC_MINB)
fiber_contract_kernel(const ContractArgs a, const int nvec, const int nvec_total, const int nslots,
                      const long long planes_per_cta)
{
    if (run_done(a.done)) return;
    constexpr int RC = 2 * NP2;
    __shared__ __align__(16) float sW[2][C2C_FC_TILE][2 * RC];      // double-buffered (w, w) pairs
    const int ldr = a.ldr, r0 = a.r0, R = a.R;
    const size_t E = (size_t)a.I2 * a.I3;
    // nvec = positions per (block, slot); blockIdx.y walks position chunks when a plane has more than blockDim.x of them
    const int slot = threadIdx.x / nvec;
    const int pos = blockIdx.y * nvec + (threadIdx.x - slot * nvec);
    const bool on = slot < nslots && pos < nvec_total;
    const long long pb = a.p0 + (long long)blockIdx.x * planes_per_cta;
    long long pe = pb + planes_per_cta; if (pe > a.P) pe = a.P;

    float2 acc[(VEC + 1) / 2][RC];
#pragma unroll
    for (int j = 0; j < (VEC + 1) / 2; ++j)
#pragma unroll
        for (int r = 0; r < RC; ++r) acc[j][r] = make_float2(0.0f, 0.0f);

    const float* xb = a.X + (size_t)pos * VEC;
    const uint8_t* mb = (MASKED && a.mask) ? a.mask + (size_t)pos * VEC : nullptr;
    const float* brow = MASKED ? a.Bmat + (size_t)pos * VEC * ldr + r0 : nullptr;

    int buf = 0;
    for (long long t0 = pb; t0 < pe; t0 += C2C_FC_TILE, buf ^= 1) {
        const int tn = (int)((pe - t0) < C2C_FC_TILE ? (pe - t0) : C2C_FC_TILE);
        // lead weights of the tile's planes: thread i -> (plane i / RC, column i % RC)
        for (int i = threadIdx.x; i < tn * RC; i += blockDim.x) {
            const int pl = i / RC, r = i - pl * RC;
            const float w = (r0 + r < R) ? lead_weight_col(t0 + pl, a.lead, ldr, r0 + r, -1) : 0.0f;
            sW[buf][pl][2 * r] = w; sW[buf][pl][2 * r + 1] = w;
        }
        __syncthreads();   // one barrier per tile: the other buffer is only rewritten after the next barrier
        if (on) {
            for (int q0 = slot; q0 < tn; q0 += nslots * C2C_FC_UNROLL) {
                float x[C2C_FC_UNROLL][VEC];
                unsigned m[C2C_FC_UNROLL];
#pragma unroll
                for (int u = 0; u < C2C_FC_UNROLL; ++u) {
                    const int pl = q0 + u * nslots;
                    if (pl < tn) {
                        ld_x<VEC>(x[u], xb + (size_t)(t0 + pl) * E);
                        m[u] = mb ? ld_m<VEC>(mb + (size_t)(t0 + pl) * E) : 0x01010101u;
                    }
                }
#pragma unroll
                for (int u = 0; u < C2C_FC_UNROLL; ++u) {
                    const int pl = q0 + u * nslots;
                    if (pl < tn) {
                        const float4* wv = reinterpret_cast<const float4*>(&sW[buf][pl][0]);
                        if (MASKED) {
                            constexpr unsigned full = VEC == 4 ? 0x01010101u : (VEC == 2 ? 0x0101u : 0x01u);
                            if ((m[u] & full) != full) {
#pragma unroll
                                for (int j = 0; j < VEC; ++j) {
                                    if (!((m[u] >> (8 * j)) & 1u)) {
                                        const float* b = brow + (size_t)j * ldr;
                                        float rec = 0.0f;
#pragma unroll
                                        for (int q = 0; q < NP2; ++q) {
                                            const float4 t = wv[q];
                                            const float2 bb = __ldg(reinterpret_cast<const float2*>(b + 2 * q));
                                            rec = fmaf(t.x, bb.x, rec);
                                            rec = fmaf(t.z, bb.y, rec);
                                        }
                                        x[u][j] = rec;
                                    }
                                }
                            }
                        }
                        if (VEC >= 2) {
#pragma unroll
                            for (int q = 0; q < NP2; ++q) {
                                const float4 t = wv[q];
#pragma unroll
                                for (int j = 0; j < VEC / 2; ++j) {
                                    const float2 xx = make_float2(x[u][2 * j], x[u][2 * j + 1]);
                                    acc[j][2 * q]     = ffma2(xx, make_float2(t.x, t.y), acc[j][2 * q]);
                                    acc[j][2 * q + 1] = ffma2(xx, make_float2(t.z, t.w), acc[j][2 * q + 1]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < NP2; ++q) {
                                const float4 t = wv[q];
                                acc[0][2 * q].x     = fmaf(x[u][0], t.x, acc[0][2 * q].x);
                                acc[0][2 * q + 1].x = fmaf(x[u][0], t.z, acc[0][2 * q + 1].x);
                            }
                        }
                    }
                }
            }
        }
    }
    if (on) {
        float* o = a.out + ((size_t)blockIdx.x * nslots + slot) * E * ldr + (size_t)pos * VEC * ldr + r0;
#pragma unroll
        for (int j = 0; j < VEC; ++j)
#pragma unroll
            for (int r = 0; r < RC; ++r)
                if (r0 + r < R) o[(size_t)j * ldr + r] = (j & 1) ? acc[j / 2][r].y : acc[j / 2][r].x;
    }
}

// host-side launchers, one translation unit per (VEC, MASKED) so the 16 NP2 instances compile in parallel
typedef cudaError_t (*plane_launch_fn)(int np2, const ContractArgs& a, int grid, size_t smem, cudaStream_t st);
typedef cudaError_t (*fiber_launch_fn)(int np2, const ContractArgs& a, dim3 grid, int block, int nvec, int nvec_total,
                                       int nslots, long long planes_per_cta, cudaStream_t st);

#define C2C_DECL_LAUNCHERS(V, M)                                                                                   \
    cudaError_t c2c_launch_plane_v##V##_m##M(int np2, const ContractArgs& a, int grid, size_t smem, cudaStream_t st); \
    cudaError_t c2c_launch_fiber_v##V##_m##M(int np2, const ContractArgs& a, dim3 grid, int block, int nvec,       \
                                             int nvec_total, int nslots, long long planes_per_cta, cudaStream_t st);
C2C_DECL_LAUNCHERS(4, 0) C2C_DECL_LAUNCHERS(4, 1) C2C_DECL_LAUNCHERS(2, 0) C2C_DECL_LAUNCHERS(2, 1)
C2C_DECL_LAUNCHERS(1, 0) C2C_DECL_LAUNCHERS(1, 1)
cudaError_t c2c_launch_sums_v4(int np2, const ContractArgs& a, int grid, size_t smem, cudaStream_t st);
cudaError_t c2c_launch_sums_v2(int np2, const ContractArgs& a, int grid, size_t smem, cudaStream_t st);
cudaError_t c2c_launch_sums_v1(int np2, const ContractArgs& a, int grid, size_t smem, cudaStream_t st);
