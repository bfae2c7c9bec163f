// Shared-memory-staged versions of the two streaming passes (sm_100a): the tensor is moved HBM -> shared memory by the
// bulk-copy engine (cp.async.bulk, completion on an mbarrier -- SASS UBLKCP), several stages ahead of the warps that
// consume it, so the bytes in flight per SM no longer depend on how many registers the accumulators leave for loads.
//
//   plane_stream<VEC,NP2>   T[p, r]    = sum_{s,v} X[p,s,v] * A2[s,r] * A3[v,r]
//       one persistent CTA per SM; every consumer warp owns a private ring of stages (a stage = `spw` whole planes,
//       contiguous in X) that its lane 0 re-arms as soon as the warp has finished reading it: no CTA-wide barrier in
//       the steady state.  Inside a stage the mapping is the register-blocked one of plane_contract: G lanes per
//       plane, a lane owns VEC-wide column strips and walks the rows, A2 rows are broadcast from shared memory.
//   fiber_stream<VEC,NP2>   U[s,v,r]   = sum_p X[p,s,v] * W[p,r],   W = Khatri-Rao rows of the leading factors
//       one persistent CTA per SM owning a contiguous range of planes; a producer warp streams stages of `tp` planes
//       of X together with their `tp` rows of W; thread `pos` of the consumer warps owns VEC consecutive in-plane
//       positions and keeps their R accumulators in registers for the whole range (one partial per CTA, summed in a
//       fixed order by fiber_reduce).
// Accumulators are rank pairs (float2 over ranks 2q, 2q+1) fed to the packed FFMA2: acc[j][q] += (x_j, x_j) * (b_2q, b_2q+1).
#pragma once
#include "c2c_common.cuh"

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; `bytes` multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int VEC> __device__ __forceinline__ void lds_x(float (&x)[VEC], const float* p);
template <> __device__ __forceinline__ void lds_x<4>(float (&x)[4], const float* p) {
    const float4 v = *reinterpret_cast<const float4*>(p); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
template <> __device__ __forceinline__ void lds_x<2>(float (&x)[2], const float* p) {
    const float2 v = *reinterpret_cast<const float2*>(p); x[0] = v.x; x[1] = v.y;
}
template <> __device__ __forceinline__ void lds_x<1>(float (&x)[1], const float* p) { x[0] = *p; }

// NP2 rank pairs of a row whose pitch is a multiple of 4 floats (16-byte aligned): LDS.128 x NP2/2 (+ LDS.64)
template <int NP2> __device__ __forceinline__ void lds_pairs(float2 (&b)[NP2], const float* row) {
#pragma unroll
    for (int q = 0; q + 1 < NP2; q += 2) {
        const float4 t = *reinterpret_cast<const float4*>(row + 2 * q);
        b[q] = make_float2(t.x, t.y); b[q + 1] = make_float2(t.z, t.w);
    }
    if (NP2 & 1) b[NP2 - 1] = *reinterpret_cast<const float2*>(row + 2 * (NP2 - 1));
}

#define C2C_ST_MAXWARPS 8
#define C2C_ST_MAXRING 4             // deepest ring per consumer warp (plane_stream)
#ifndef C2C_FS_STAGES
#define C2C_FS_STAGES 3              // ring depth per CTA (fiber_stream)
#endif

struct StreamArgs {
    ContractArgs c;
    const float* W;            // [P x ldr] Khatri-Rao rows of the leading factors (fiber_stream; masked plane_stream)
    int spw;                   // planes per stage
    int rs;                    // plane_stream: row splits per plane (a stage holds I2 / rs rows of `spw` planes); 1 = whole planes
    int nring;                 // plane_stream: stages in a warp's ring
    long long nstage;          // plane_stream: number of full plane groups (of spw planes); fiber_stream: full stages.
                               // The tail planes go to the register-staged kernels.
    int rp;                    // pitch (floats, multiple of 4) of the A2 / A3 copies in shared memory
    int pitch;                 // plane_stream: pitch (floats) of one plane's rows inside a stage (>= (I2 / rs) * I3)
    int nwarps;                // consumer warps per CTA
    long long stages_per_cta;  // fiber_stream: contiguous stages per CTA
};

// --------------------------------------------------------------------------------------------------------------------
// plane_stream.  A warp walks plane groups g = first, first + stride, ...; a group is `spw` planes and is fetched as `rs`
// stages.  rs == 1: a stage is `spw` whole planes, contiguous in X (one bulk copy), consumed ppw planes at a time.
// rs > 1 (large planes): spw == ppw and stage h holds rows [h * I2/rs, (h+1) * I2/rs) of each plane (ppw bulk copies,
// each landing at a pitch chosen so the lanes of consecutive planes fall in consecutive banks); the accumulators stay in
// registers across the rs stages of a group.
// Register blocking: a lane owns NS column strips of VEC columns (strip sets lig, lig + nsets) -> every broadcast A2
// value read from shared memory feeds NS * VEC FMAs.  (A broadcast LDS costs one shared-memory wavefront per 32-bit word
// whatever its width, so the wavefront count per FMA is what bounds this kernel once HBM is hidden.)
template <int VEC, int NS, int NP2>
__global__ void __launch_bounds__(32 * C2C_ST_MAXWARPS, 1)
plane_stream_kernel(const StreamArgs sa)
{
    const ContractArgs& a = sa.c;
    if (run_done(a.done)) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int I2 = a.I2, I3 = a.I3, ldr = a.ldr, R = a.R, rp = sa.rp;
    const int S = sa.nring, RS = sa.rs, rps = I2 / RS;
    const size_t E = (size_t)I2 * I3;
    float* sA2 = reinterpret_cast<float*>(smem_raw);                 // [I2][rp]
    float* sA3 = sA2 + (size_t)I2 * rp;                              // [I3][rp]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sA3 + (size_t)I3 * rp);   // [nwarps][S]
    const size_t chunk_floats = (size_t)rps * I3;                    // one plane's share of a stage
    const size_t chunk_pitch = (size_t)sa.pitch;                     // its pitch in shared memory (>= chunk_floats)
    const size_t stage_floats = (size_t)sa.spw * chunk_pitch;
    const size_t off = ((size_t)(I2 + I3) * rp * sizeof(float) + (size_t)sa.nwarps * S * 8 + 127) / 128 * 128;
    float* stage0 = reinterpret_cast<float*>(smem_raw + off);
    constexpr int PITCH2 = NP2 | 1;                                  // odd pitch (in float2): conflict-free scratch rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* red = reinterpret_cast<float2*>(stage0 + (size_t)sa.nwarps * S * stage_floats) + warp * 32 * PITCH2;
    unsigned long long* mybar = bars + warp * S;
    float* mystage = stage0 + (size_t)warp * S * stage_floats;
    const unsigned stage_bytes = (unsigned)((size_t)sa.spw * chunk_floats * sizeof(float));
    const long long first = (long long)blockIdx.x * sa.nwarps + warp;       // this warp's first plane group
    const long long stride = (long long)gridDim.x * sa.nwarps;
    const long long ngroups_mine = first < sa.nstage ? (sa.nstage - first + stride - 1) / stride : 0;
    const long long nstages_mine = ngroups_mine * RS;

    // The warp is its own producer.  Its stage sequence is (group, row split) = (first, 0), (first, 1), ..,
    // (first + stride, 0), ..; `pg / ph` walk that sequence for the producer side, S stages ahead of the consumer.
    const unsigned chunk_bytes = (unsigned)(chunk_floats * sizeof(float));
    long long pg = first; int ph = 0, pslot = 0;                     // next stage to request and the ring slot it goes to
    auto issue = [&]() {
        float* dst = mystage + (size_t)pslot * stage_floats;
        if (lane == 0) mbar_expect_tx(mybar + pslot, stage_bytes);
        __syncwarp();
        if (RS == 1) {
            if (lane == 0) bulk_g2s(dst, a.X + (size_t)pg * sa.spw * E, stage_bytes, mybar + pslot);
        } else if (lane < sa.spw) {                                  // one row slab per plane, one lane each
            bulk_g2s(dst + (size_t)lane * chunk_pitch, a.X + ((size_t)pg * sa.spw + lane) * E + (size_t)ph * chunk_floats,
                     chunk_bytes, mybar + pslot);
        }
        if (++ph == RS) { ph = 0; pg += stride; }
        if (++pslot == S) pslot = 0;
    };
    if (lane == 0) {
        for (int j = 0; j < S; ++j) mbar_init(mybar + j, 1);
        mbar_init_fence();
    }
    __syncwarp();
    // (filling the ring before this wait -- the tensor does not depend on the previous kernel -- was measured and is not
    // faster here: 0.1487 vs 0.1465 ms per pass, C2C_PS_PREFETCH_FIRST)
#ifdef C2C_PS_PREFETCH_FIRST
    for (int j = 0; j < S && pg < sa.nstage; ++j) issue();
#endif
    pdl_wait();
    for (int i = threadIdx.x; i < I2 * rp; i += blockDim.x) {
        const int s = i / rp, r = i - s * rp;
        sA2[i] = r < R ? a.A2[(size_t)s * ldr + r] : 0.0f;
    }
    for (int i = threadIdx.x; i < I3 * rp; i += blockDim.x) {
        const int s = i / rp, r = i - s * rp;
        sA3[i] = r < R ? a.A3[(size_t)s * ldr + r] : 0.0f;
    }
#ifndef C2C_PS_PREFETCH_FIRST
    for (int j = 0; j < S && pg < sa.nstage; ++j) issue();
#endif
    __syncthreads();                     // A2 / A3 copies visible; the only CTA-wide barrier

    const int nsets = I3 / (VEC * NS);   // strip sets per plane row
    const int G = nsets < 32 ? nsets : 32;
    const int ppw = 32 / G;
    const int grp = lane / G, lig = lane - grp * G;
    const bool lane_on = grp < ppw;

    const int i3 = I3, pitch_i = (int)chunk_pitch, soff = nsets * VEC;
    float2 acc[NS][VEC][NP2];
    int cslot = 0; unsigned cphase = 0;
    for (long long g = first; g < sa.nstage; g += stride) {
        for (int h = 0; h < RS; ++h) {
            mbar_wait(mybar + cslot, cphase);
            const float* xs_stage = mystage + (size_t)cslot * stage_floats;
            const float* a2h = sA2 + h * rps * rp;
            for (int pl0 = 0; pl0 < sa.spw; pl0 += ppw) {
                const int pl = pl0 + grp;
                const bool on = lane_on && pl < sa.spw;
                float2 out[NP2];
#pragma unroll
                for (int q = 0; q < NP2; ++q) out[q] = make_float2(0.0f, 0.0f);
                if (on) {
                    for (int ss = lig; ss < nsets; ss += G) {
                        const int v0 = ss * VEC;
                        if (h == 0) {
#pragma unroll
                            for (int n = 0; n < NS; ++n)
#pragma unroll
                                for (int jj = 0; jj < VEC; ++jj)
#pragma unroll
                                    for (int q = 0; q < NP2; ++q) acc[n][jj][q] = make_float2(0.0f, 0.0f);
                        }
                        const float* xp = xs_stage + pl * pitch_i + v0;
                        const float* bp = a2h;
                        auto row = [&](const float* xr, const float* br) {
                            float x[NS][VEC];
#pragma unroll
                            for (int n = 0; n < NS; ++n) lds_x<VEC>(x[n], xr + n * soff);
                            float2 b[NP2];
                            lds_pairs<NP2>(b, br);
#pragma unroll
                            for (int n = 0; n < NS; ++n)
#pragma unroll
                                for (int jj = 0; jj < VEC; ++jj) {
                                    const float2 xx = make_float2(x[n][jj], x[n][jj]);
#pragma unroll
                                    for (int q = 0; q < NP2; ++q) acc[n][jj][q] = ffma2(xx, b[q], acc[n][jj][q]);
                                }
                        };
                        int s = 0;
                        for (; s + 2 <= rps; s += 2) {               // two rows per trip, no remainder code in the hot loop
                            row(xp, bp);
                            row(xp + i3, bp + rp);
                            xp += 2 * i3; bp += 2 * rp;
                        }
                        if (s < rps) row(xp, bp);
                        if (h == RS - 1) {
#pragma unroll
                            for (int n = 0; n < NS; ++n)
#pragma unroll
                                for (int jj = 0; jj < VEC; ++jj) {
                                    float2 b[NP2];
                                    lds_pairs<NP2>(b, sA3 + (v0 + n * soff + jj) * rp);
#pragma unroll
                                    for (int q = 0; q < NP2; ++q) out[q] = ffma2(acc[n][jj][q], b[q], out[q]);
                                }
                        }
                    }
                }
                if (h == RS - 1) {
                    // sum over the G lanes of a plane through the warp's scratch rows: lane writes its NP2 pairs, then lane
                    // (plane, pair) adds the G partial pairs of its plane and stores them (G need not be a power of two)
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < NP2; ++q) red[lane * PITCH2 + q] = out[q];
                    __syncwarp();
                    for (int j = lane; j < ppw * NP2; j += 32) {
                        const int pj = j / NP2, q = j - pj * NP2;
                        if (pl0 + pj < sa.spw) {
                            const float2* src = red + (pj * G) * PITCH2 + q;
                            float2 sum = src[0];
                            for (int i = 1; i < G; ++i) { const float2 v = src[i * PITCH2]; sum.x += v.x; sum.y += v.y; }
                            float* t = a.out + ((size_t)g * sa.spw + pl0 + pj) * ldr;
                            if (2 * q + 1 < R) *reinterpret_cast<float2*>(t + 2 * q) = sum;
                            else if (2 * q < R) t[2 * q] = sum.x;
                        }
                    }
                }
            }
            __syncwarp();
            if (pg < sa.nstage) issue();                             // re-arm the slot just drained (pslot == cslot here)
            if (++cslot == S) { cslot = 0; cphase ^= 1u; }
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// fiber_stream: blockDim = 32 * (nwarps + 1); the last warp is the producer.  Consumer thread (slot, pos): `nslots`
// thread groups take alternate planes of a stage; a thread owns NS float4 / float2 / float column sets pos, pos + nthr.
#define C2C_FS_MAXTHREADS(NS, NP2) (((NS) == 2 ? (NP2) <= 5 : (NP2) <= 12) ? 512 : 256)
template <int VEC, int NS, int NP2>
__global__ void __launch_bounds__(C2C_FS_MAXTHREADS(NS, NP2), 1)
fiber_stream_kernel(const StreamArgs sa, const int nthr /* threads per slot */, const int nslots, const int merge,
                    const int w_all /* 1: the CTA's Khatri-Rao rows are built once, up front, into shared memory */)
{
    const ContractArgs& a = sa.c;
    if (run_done(a.done)) return;
    constexpr int S = C2C_FS_STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ldr = a.ldr, R = a.R;
    const size_t E = (size_t)a.I2 * a.I3;
    const int tp = sa.spw;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw);      // [S]
    unsigned long long* empty = full + S;                                            // [S]
    const size_t x_floats = (size_t)tp * E, w_floats = w_all ? 0 : (size_t)tp * ldr;
    const size_t stage_floats = (x_floats + w_floats + 31) / 32 * 32;
    float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
    float* wall = stage0 + (size_t)S * stage_floats;                 // [stages_per_cta * tp][ldr] when w_all
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = sa.nwarps;
    long long sb = (long long)blockIdx.x * sa.stages_per_cta;
    long long se = sb + sa.stages_per_cta; if (se > sa.nstage) se = sa.nstage;
    if (sb > se) sb = se;
    if (threadIdx.x == 0) {
        for (int j = 0; j < S; ++j) { mbar_init(full + j, w_all ? 1 : 2); mbar_init(empty + j, ncw); }
        mbar_init_fence();
    }
    __syncthreads();

    if (warp == ncw) {
        // ---- producer: bulk copy of the stage's planes, then its Khatri-Rao rows W[p, :] = prod_k A_k[i_k(p), :] built
        // in place by the 32 lanes ("full" takes two arrivals: the copy's expect_tx and the one that publishes W) ----
        const unsigned xb = (unsigned)(x_floats * sizeof(float));
        const int q4 = ldr >> 2, nitems = tp * q4;
        int slot = 0; unsigned phase = 0; bool wrapped = false;
        if (!w_all) pdl_wait();          // the per-stage Khatri-Rao rows read the factors the previous kernel wrote
        for (long long g = sb; g < se; ++g) {
            if (wrapped) mbar_wait(empty + slot, phase ^ 1u);
            float* dst = stage0 + (size_t)slot * stage_floats;
            if (lane == 0) {
                mbar_expect_tx(full + slot, xb);
                bulk_g2s(dst, a.X + (size_t)g * x_floats, xb, full + slot);
            }
            for (int it = lane; it < (w_all ? 0 : nitems); it += 32) {
                const int pl = it / q4, quad = it - pl * q4;
                long long rem = g * tp + pl;
                float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
                for (int k = a.lead.nl - 1; k >= 0; --k) {
                    const int d = a.lead.dim[k];
                    const long long qq = rem / d;
                    const int i = (int)(rem - qq * d);
                    rem = qq;
                    const float4 f = __ldg(reinterpret_cast<const float4*>(a.lead.fac[k] + (size_t)i * ldr) + quad);
                    w.x *= f.x; w.y *= f.y; w.z *= f.z; w.w *= f.w;
                }
                *reinterpret_cast<float4*>(dst + x_floats + (size_t)pl * ldr + quad * 4) = w;
            }
            __syncwarp();
            if (lane == 0 && !w_all) mbar_arrive(full + slot);
            if (++slot == S) { slot = 0; phase ^= 1u; wrapped = true; }
        }
        return;
    }
    // ---- consumers ----
    const int slot_id = threadIdx.x / nthr, pos = threadIdx.x - slot_id * nthr;
    const bool on = slot_id < nslots;
    float2 acc[NS][VEC][NP2];
#pragma unroll
    for (int n = 0; n < NS; ++n)
#pragma unroll
        for (int j = 0; j < VEC; ++j)
#pragma unroll
            for (int q = 0; q < NP2; ++q) acc[n][j][q] = make_float2(0.0f, 0.0f);
    const int e_i = (int)E, step = nslots * e_i, wstep = nslots * ldr, noff = nthr * VEC;
    pdl_wait();                          // (the producer warp is already streaming the tensor)
    if (w_all) {
        // Khatri-Rao rows of this CTA's planes, built by the consumer warps while the first stages are in flight
        const int q4 = ldr >> 2;
        const long long nitems = (se - sb) * tp * q4;
        for (long long it = threadIdx.x; it < nitems; it += (long long)ncw * 32) {
            const long long pl = it / q4; const int quad = (int)(it - pl * q4);
            long long rem = sb * tp + pl;
            float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
            for (int k = a.lead.nl - 1; k >= 0; --k) {
                const int d = a.lead.dim[k];
                const long long qq = rem / d;
                const int i = (int)(rem - qq * d);
                rem = qq;
                const float4 f = __ldg(reinterpret_cast<const float4*>(a.lead.fac[k] + (size_t)i * ldr) + quad);
                w.x *= f.x; w.y *= f.y; w.z *= f.z; w.w *= f.w;
            }
            *reinterpret_cast<float4*>(wall + (size_t)pl * ldr + quad * 4) = w;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(ncw * 32) : "memory");
    }
    int cslot = 0; unsigned cphase = 0;
    const float* wrow_cta = wall + slot_id * ldr;                    // w_all: advances tp rows per stage
    for (long long g = sb; g < se; ++g) {
        mbar_wait(full + cslot, cphase);
        const float* xs = stage0 + (size_t)cslot * stage_floats;
        if (on) {
            const float* xp = xs + slot_id * e_i + pos * VEC;
            const float* wp = w_all ? wrow_cta : xs + x_floats + slot_id * ldr;
            auto plane = [&](const float* xr, const float* wr) {
                float x[NS][VEC];
#pragma unroll
                for (int n = 0; n < NS; ++n) lds_x<VEC>(x[n], xr + n * noff);
                float2 b[NP2];
                lds_pairs<NP2>(b, wr);
#pragma unroll
                for (int n = 0; n < NS; ++n)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        const float2 xx = make_float2(x[n][j], x[n][j]);
#pragma unroll
                        for (int q = 0; q < NP2; ++q) acc[n][j][q] = ffma2(xx, b[q], acc[n][j][q]);
                    }
            };
            // tp is a multiple of 2 * nslots (host): two planes per trip, no remainder code
            for (int pl = 0; pl < tp; pl += 2 * nslots) {
                plane(xp, wp);
                plane(xp + step, wp + wstep);
                xp += 2 * step; wp += 2 * wstep;
            }
        }
        wrow_cta += (size_t)tp * ldr;
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + cslot);
        if (++cslot == S) { cslot = 0; cphase ^= 1u; }
    }
    // the slots' accumulators are combined through shared memory (the ring is drained), slot 0 writes the CTA's partial
    const int nout = merge ? 1 : nslots;
    if (merge && nslots > 1) {
        constexpr int PER = NS * VEC * NP2;                          // float2 per thread
        float2* sacc = reinterpret_cast<float2*>(stage0);
        asm volatile("bar.sync 1, %0;" ::"r"(ncw * 32) : "memory");
        if (on && slot_id > 0) {
            float2* dst = sacc + ((size_t)(slot_id - 1) * nthr + pos) * PER;
#pragma unroll
            for (int n = 0; n < NS; ++n)
#pragma unroll
                for (int j = 0; j < VEC; ++j)
#pragma unroll
                    for (int q = 0; q < NP2; ++q) dst[(n * VEC + j) * NP2 + q] = acc[n][j][q];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(ncw * 32) : "memory");
        if (on && slot_id == 0) {
            for (int sl = 1; sl < nslots; ++sl) {
                const float2* src = sacc + ((size_t)(sl - 1) * nthr + pos) * PER;
#pragma unroll
                for (int n = 0; n < NS; ++n)
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
#pragma unroll
                        for (int q = 0; q < NP2; ++q) {
                            const float2 t = src[(n * VEC + j) * NP2 + q];
                            acc[n][j][q].x += t.x; acc[n][j][q].y += t.y;
                        }
            }
        }
    }
    if (on && (slot_id == 0 || !merge)) {
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            float* o = a.out + (((size_t)blockIdx.x * nout + (merge ? 0 : slot_id)) * E + ((size_t)n * nthr + pos) * VEC) * ldr;
#pragma unroll
            for (int j = 0; j < VEC; ++j)
#pragma unroll
                for (int q = 0; q < NP2; ++q) {
                    if (2 * q + 1 < R) *reinterpret_cast<float2*>(o + (size_t)j * ldr + 2 * q) = acc[n][j][q];
                    else if (2 * q < R) o[(size_t)j * ldr + 2 * q] = acc[n][j][q].x;
                }
        }
    }
}

// launchers: (VEC, NS) in {(4,1), (4,2), (2,1), (1,1)}; NS = 2 instances exist for NP2 <= C2C_NS2_MAXNP2 only
#define C2C_NS2_MAXNP2 10
#define C2C_DECL_STREAM(V, N)                                                                                       \
    cudaError_t c2c_launch_plane_stream_v##V##_s##N(int np2, const StreamArgs& a, int grid, int block, size_t smem, cudaStream_t st, int pdl); \
    cudaError_t c2c_launch_fiber_stream_v##V##_s##N(int np2, const StreamArgs& a, int nthr, int nslots, int merge, int w_all, int grid, int block, size_t smem, cudaStream_t st, int pdl);
C2C_DECL_STREAM(4, 1) C2C_DECL_STREAM(4, 2) C2C_DECL_STREAM(2, 1) C2C_DECL_STREAM(1, 1)
