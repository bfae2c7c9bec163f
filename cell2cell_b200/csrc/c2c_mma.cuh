// Tensor-core versions of the two streaming passes (sm_100a: TMA + tcgen05.mma kind::tf32 + TMEM accumulators).
//
// Why: the contraction is 2 R flop per tensor entry.  The FFMA kernels of c2c_stream.cuh keep up with HBM only up to
// R ~ 10 (measured: fma pipe 48 % busy at R = 10, 0.226 ms per pass at R = 20 against a 0.110 ms read floor); on the
// tensor pipe the same contraction costs a few per cent of the pass, for every R <= 64.
//
//   plane_mma   T[p, r] = sum_k X[p, k] * KR[k, r],  KR[(s,v), r] = A2[s, r] * A3[v, r]        (k = in-plane position)
//               GEMM M = planes (tiles of 128), K = E, N = rank.  A operand = a 128 x 32 tile of X, K-major, written by
//               TMA with the 128-byte swizzle; B operand = a 32-wide slab of KR^T, generated in shared memory.
//   fiber_mma   U[e, r] = sum_p X[p, e] * W[p, r],   W = Khatri-Rao rows of the leading factors
//               GEMM M = in-plane positions (tiles of 128), K = planes (8 per MMA), N = rank.  A operand = X^T: eight
//               whole planes, MN-major, written by one 3-D TMA copy in the 32-byte-atom 128-byte swizzle (the only
//               MN-major layout kind::tf32 takes); B operand = the 8 rows of W, K-major core matrices.
//
// fp32 accuracy on a tf32 pipe: kind::tf32 truncates its operands to 10 mantissa bits (measured, scripts/umma_probe.cu),
// so every product is issued as three MMAs   hi(x) hi(b) + lo(x) hi(b) + hi(x) lo(b),   lo(v) = v - trunc_tf32(v)
// (exact in fp32).  hi() costs nothing (the hardware truncation of the untouched tile); the lo tiles are written by the
// "prep" warps between the TMA arrival and the MMA issue.  What is dropped is lo*lo and the truncation of the lo terms:
// ~2^-21 relative per product, accumulated in fp32 in TMEM.
//
// Warp roles (576 threads, one CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) and TMEM
// owner, warps 2-5 = epilogue (TMEM -> registers), warps 6-17 = prep, three groups of four warps working on three
// consecutive items at once (a group's warp w serves TMEM lanes 32 (w % 4) ..).
// Shared-memory bandwidth (128 B / clk / SM) is the budget that shapes the kernels: the TMA write, the prep read and every
// MMA operand read of a tile all go through it.  So lo(X) never touches shared memory -- the prep warps write it to
// TMEM (tcgen05.st) and the MMA takes it as a TMEM A operand -- and hi(x) hi(b) + hi(x) lo(b) are ONE MMA over a B tile
// that stacks hi(b) and lo(b) along N (the lo products land in their own accumulator columns).
#pragma once
#include <cuda.h>
#include "c2c_stream.cuh"

#define C2C_MMA_THREADS 576        // 18 warps
#define C2C_MMA_PREP_WARP0 6
#define C2C_MMA_PREP_GROUPS 3      // prep groups of 4 warps (one per TMEM lane quarter); group q takes items t = q, q + 3, ..
#ifndef C2C_MMA_TRUNC_LOSS
#define C2C_MMA_TRUNC_LOSS 5.9604645e-8f   // 2^-24: mean relative loss of one truncating accumulation into TMEM
#endif
#define C2C_MMA_RING2 4            // prep -> MMA ring: B tile (shared memory) + lo(X) tile (TMEM) + ready / empty barriers

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// one lane of the (converged) warp; the warp-uniform control flow around it stays on the uniform datapath
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// warp-wide
__device__ __forceinline__ void tmem_alloc(unsigned* slot, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned addr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// shared-memory matrix descriptor (sm_100 format: version 1 at bit 46)
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes, unsigned layout) {
    return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((unsigned long long)(layout & 7u) << 61);
}
#define C2C_UMMA_LAYOUT_NONE 0u
#define C2C_UMMA_LAYOUT_SW128_32B 1u       // 128-byte swizzle of 32-byte atoms: the MN-major layout of 32-bit operands
#define C2C_UMMA_LAYOUT_SW128 2u
// instruction descriptor: fp32 accumulate, tf32 x tf32, M = 128
__host__ __device__ constexpr unsigned umma_idesc_tf32(unsigned n, unsigned a_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major << 15) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive TMEM columns of this thread's lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// store 8 / 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_st8(unsigned taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float4 tf32_lo4(float4 v) { return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)); }

__host__ __device__ constexpr unsigned c2c_tmem_cols(unsigned need) { return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512; }

// --------------------------------------------------------------------------------------------------------------------
// Accumulation note.  The tensor pipe adds into TMEM with truncation: every MMA that accumulates into D costs ~0.5 ulp(D),
// always downwards (measured: 600 chained MMAs -> 3.4e-5 relative bias).  Chains are therefore kept short -- an
// accumulator takes a few MMAs of full-size terms (the lo terms go to columns of their own, where the loss is relative to
// a 2^-10 times smaller sum), then the epilogue warps add it, round-to-nearest, into fp32 sums of their own.
// --------------------------------------------------------------------------------------------------------------------
// plane_mma
// --------------------------------------------------------------------------------------------------------------------
#define C2C_PM_CHUNK 2          // k blocks per accumulator chain (8 MMAs into the hi columns)
// C2C_PM_ALL_TS (default): hi(X) goes through TMEM too (both MMAs take their A operand from TMEM): the X tile in shared
// memory is read once (by the prep warps) instead of twice, and its ring slot is released by them instead of by the MMA
// commit.  Measured against the shared-memory A operand (-DC2C_PM_SMEM_A): 0.160 -> 0.135 ms per pass at N = 16,
// 0.178 -> 0.164 at N = 32 -- shared-memory bandwidth (TMA write + prep read + MMA read) is what bounds these kernels.
#ifndef C2C_PM_SMEM_A
#define C2C_PM_ALL_TS 1
#endif
#ifdef C2C_PM_ALL_TS
#define C2C_PM_EMPTYX_COUNT 4
#define C2C_PM_A_COLS 64        // TMEM columns per prep slot: hi | lo
#else
#define C2C_PM_EMPTYX_COUNT 1
#define C2C_PM_A_COLS 32
#endif

struct PlaneMmaArgs {
    long long P;               // planes
    int E, I2, I3, R, ldr;
    const float* A2; const float* A3;     // trailing factors [I2 x ldr], [I3 x ldr]
    float* T;                  // [P x ldr], zero on entry: every (tile, k range) piece is added atomically (at most two per tile
                               // -> the sum does not depend on arrival order)
    const int* done;
    int nkb;                   // k blocks of 32 in-plane positions: ceil(E / 32)
    long long ntiles;          // tiles of 128 planes: ceil(P / 128)
    int nx;                    // X ring depth
    int dbg;                   // timing experiments only: 1 skip lo(X), 2 skip B generation, 4 skip MMAs, 8 skip epilogue reads
};

// N = 24 (ranks 17-24): an M = 128 MMA takes N in multiples of 16, so the stacked MMA runs at 2N = 48 and the lo(x) hi(b) MMA at
// N = 32 over B rows 0..31 -- its columns 24..31 (lo(x) x the first eight lo(b) rows) land in the 8 spare columns of a 64-column
// accumulator slot and are never read: 80 tensor columns per k-step instead of the 96 of N = 32.
template <int N> struct PlaneMmaSmem {
    static constexpr int X_BYTES = 128 * 128;              // 128 planes x 32 floats
    static constexpr int B_BYTES = 2 * N * 128;            // rows 0..N-1: hi(KR^T), rows N..2N-1: lo(KR^T); 32 floats each
    static constexpr int ACC = N == 24 ? 64 : 2 * N;       // TMEM columns of one accumulator slot: hi | lo (| spare)
    static constexpr int N_LO = N == 24 ? 32 : N;          // N of the lo(x) hi(b) MMA
    static constexpr int NR = 256 / ACC;                   // accumulator ring
    static constexpr int ALO_COL = 256;                    // lo(X) ring: C2C_MMA_RING2 x 32 columns
    static constexpr int TM_COLS = 512;
};

template <int N>
__global__ void __launch_bounds__(C2C_MMA_THREADS, 1)
plane_mma_kernel(const __grid_constant__ CUtensorMap mapX, const PlaneMmaArgs a)
{
    if (run_done(a.done)) return;
    using L = PlaneMmaSmem<N>;
    constexpr int NR = L::NR, D2 = C2C_MMA_RING2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int DX = a.nx, I2 = a.I2, I3 = a.I3, R = a.R, ldr = a.ldr, nkb = a.nkb;
    unsigned char* xring = base;
    unsigned char* bring = base + (size_t)DX * L::X_BYTES;
    float* A2s = reinterpret_cast<float*>(bring + (size_t)D2 * L::B_BYTES);            // [I2][N]
    float* A3s = A2s + (size_t)I2 * N;                                                // [I3][N]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(A3s + (size_t)I3 * N);
    unsigned long long *fullx = bars, *emptyx = bars + DX, *ready = bars + 2 * DX, *empty2 = ready + D2, *tfull = empty2 + D2, *tempty = tfull + NR;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + NR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < DX; ++s) { mbar_init(fullx + s, 1); mbar_init(emptyx + s, C2C_PM_EMPTYX_COUNT); }
        for (int s = 0; s < D2; ++s) { mbar_init(ready + s, 4); mbar_init(empty2 + s, 1); }
        for (int s = 0; s < NR; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc(tmem_slot, L::TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    const long long G = a.ntiles * nkb;                          // work items: (tile, k block), tile-major
    const long long g0 = (long long)blockIdx.x * G / gridDim.x, g1 = (long long)(blockIdx.x + 1) * G / gridDim.x;

    if (warp == 0) {
        // ---- TMA producer ----
        if (lane == 0) {
            int sx = 0; unsigned phx = 0;
            int tile = (int)(g0 / nkb), kb = (int)(g0 - (long long)tile * nkb);
            for (long long g = g0; g < g1; ++g) {
                mbar_wait(emptyx + sx, phx ^ 1u);
                mbar_expect_tx(fullx + sx, L::X_BYTES);
                tma_load_2d(xring + (size_t)sx * L::X_BYTES, &mapX, kb * 32, tile * 128, fullx + sx);
                if (++sx == DX) { sx = 0; phx ^= 1u; }
                if (++kb == nkb) { kb = 0; ++tile; }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: chains of up to C2C_PM_CHUNK k blocks, each into the next accumulator of the ring.  The whole
        // warp runs the (uniform) control flow, one elected lane issues; descriptors advance by adding to their address field.
        {
            constexpr unsigned IDESC2 = umma_idesc_tf32(2 * N, 0), IDESC1 = umma_idesc_tf32(L::N_LO, 0);
            const unsigned long long DK = umma_desc(0, 16, 1024, C2C_UMMA_LAYOUT_SW128);
            const unsigned xa0 = smem_u32(xring) >> 4, ba0 = smem_u32(bring) >> 4;
            int sx = 0; unsigned phx = 0; int s2 = 0; unsigned ph2 = 0; int r = 0; unsigned phr = 0;
            long long g = g0;
            long long tile_end = (g0 / nkb + 1) * nkb;
            while (g < g1) {
                const long long piece_end = tile_end < g1 ? tile_end : g1;
                while (g < piece_end) {
                    const long long chunk_end = g + C2C_PM_CHUNK < piece_end ? g + C2C_PM_CHUNK : piece_end;
                    mbar_wait(tempty + r, phr ^ 1u);
                    tc_fence_after();
                    const unsigned d = tmem + (unsigned)(r * L::ACC);
                    for (long long gg = g; gg < chunk_end; ++gg) {
#ifndef C2C_PM_ALL_TS
                        mbar_wait(fullx + sx, phx);              // (all-TS: the issuer never touches the X ring -- its slots may be
#endif                                                           //  a phase ahead by now, a wait here would alias)
                        mbar_wait(ready + s2, ph2);
                        tc_fence_after();
                        if (elect_one()) {
                            const unsigned long long dx = DK + (xa0 + (unsigned)sx * (L::X_BYTES >> 4));
                            const unsigned long long db = DK + (ba0 + (unsigned)s2 * (L::B_BYTES >> 4));
                            const unsigned alo = tmem + (unsigned)(L::ALO_COL + s2 * C2C_PM_A_COLS + (C2C_PM_A_COLS - 32));
                            if (!(a.dbg & 4)) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
#ifdef C2C_PM_ALL_TS
                                    umma_tf32_ts(d, alo - 32 + ks * 8, db + 2 * ks, IDESC2, (gg > g || ks > 0) ? 1u : 0u);
#else
                                    umma_tf32(d, dx + 2 * ks, db + 2 * ks, IDESC2, (gg > g || ks > 0) ? 1u : 0u);   // [hi(x) hi(b) | hi(x) lo(b)]
#endif
                                    umma_tf32_ts(d + N, alo + ks * 8, db + 2 * ks, IDESC1, 1u);                    // lo(x) hi(b) -> the lo columns
                                }
                                if (a.dbg & 16) {                   // timing experiment: the same MMAs once more (results are wrong)
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks) {
                                        umma_tf32(d, dx + 2 * ks, db + 2 * ks, IDESC2, 1u);
                                        umma_tf32_ts(d + N, alo + ks * 8, db + 2 * ks, IDESC1, 1u);
                                    }
                                }
                            }
#ifndef C2C_PM_ALL_TS
                            umma_commit(emptyx + sx);
#else
                            (void)dx;
#endif
                            umma_commit(empty2 + s2);
                            if (gg + 1 == chunk_end) umma_commit(tfull + r);
                        }
                        __syncwarp();
                        if (++sx == DX) { sx = 0; phx ^= 1u; }
                        if (++s2 == D2) { s2 = 0; ph2 ^= 1u; }
                    }
                    if (++r == NR) { r = 0; phr ^= 1u; }
                    g = chunk_end;
                }
                tile_end += nkb;
            }
        }
    } else if (warp < C2C_MMA_PREP_WARP0) {
        // ---- epilogue: TMEM lane = plane of the tile; chains are summed in registers, pieces are added into T ----
        const int lg = warp & 3;
        int r = 0; unsigned phr = 0;
        long long g = g0;
        pdl_wait();
        while (g < g1) {
            const long long tile = g / nkb;
            const long long tile_end = (tile + 1) * nkb;
            const long long piece_end = tile_end < g1 ? tile_end : g1;
            float acc[N];
#pragma unroll
            for (int j = 0; j < N; ++j) acc[j] = 0.0f;
            while (g < piece_end) {
                const long long chunk_end = g + C2C_PM_CHUNK < piece_end ? g + C2C_PM_CHUNK : piece_end;
                mbar_wait(tfull + r, phr);
                tc_fence_after();
                // the chain's 4 MMAs per k block each truncated the running sum: ~2^-24 of it per MMA on average (measured)
                const float unbias = 1.0f + C2C_MMA_TRUNC_LOSS * (float)(4 * (int)(chunk_end - g));
                if (!(a.dbg & 8)) {
#pragma unroll
                    for (int c0 = 0; c0 < N; c0 += 16) {
                        float vh[16], vl[16];
                        tmem_ld16(tmem + ((unsigned)(lg * 32) << 16) + (unsigned)(r * L::ACC + c0), vh);
                        tmem_ld16(tmem + ((unsigned)(lg * 32) << 16) + (unsigned)(r * L::ACC + N + c0), vl);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < N) acc[c0 + j] += fmaf(vh[j], unbias, vl[j]);   // N = 24: the last load reads 8 columns too many
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + r);
                if (++r == NR) { r = 0; phr ^= 1u; }
                g = chunk_end;
            }
            const long long p = tile * 128 + lg * 32 + lane;
            if (p < a.P) {
                float* trow = a.T + (size_t)p * ldr;
#pragma unroll
                for (int j = 0; j < N; ++j)
                    if (j < ldr) atomicAdd(trow + j, acc[j]);
            }
        }
    } else {
        // ---- prep: group q takes items q, q + 3, ..: the KR slab (hi | lo) into the B ring, lo(X) into TMEM ----
        const int q = (warp - C2C_MMA_PREP_WARP0) >> 2, lq = warp & 3;
        const int tg = ((warp - C2C_MMA_PREP_WARP0) & 3) * 32 + lane;          // thread index inside the group, 0..127
        const long long nitem = g1 - g0;
        // the trailing factors come from the previous kernel (the TMA producer is already filling the X ring)
        pdl_wait();
        {
            const int tp = tid - 32 * C2C_MMA_PREP_WARP0, np = 128 * C2C_MMA_PREP_GROUPS;
            for (int i = tp; i < I2 * N; i += np) { const int s = i / N, n = i - s * N; A2s[i] = n < R ? a.A2[(size_t)s * ldr + n] : 0.0f; }
            for (int i = tp; i < I3 * N; i += np) { const int s = i / N, n = i - s * N; A3s[i] = n < R ? a.A3[(size_t)s * ldr + n] : 0.0f; }
            asm volatile("bar.sync 2, %0;" ::"n"(128 * C2C_MMA_PREP_GROUPS) : "memory");
        }
        for (long long t = q; t < nitem; t += C2C_MMA_PREP_GROUPS) {
            const long long g = g0 + t;
            const long long tile = g / nkb; const int kb = (int)(g - tile * nkb);
            const int sx = (int)(t % DX), s2 = (int)(t % D2);
            const unsigned phx = (unsigned)((t / DX) & 1), ph2 = (unsigned)((t / D2) & 1);
            unsigned char* bt = bring + (size_t)s2 * L::B_BYTES;
            mbar_wait(empty2 + s2, ph2 ^ 1u);                    // the MMAs of the slot's previous round are done
            // KR slab (does not need the X tile): row n, 16-byte chunk c = k positions kb*32 + 4c .. +3
            // (computing the values before the wait was measured and is not faster here)
            for (int it = tg; it < ((a.dbg & 2) ? 0 : N * 8); it += 128) {
                const int n = it % N, c = it / N;
                const int k0 = kb * 32 + c * 4;
                float h[4], l[4];
                int sI = k0 / I3, vI = k0 - sI * I3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float val = 0.0f;
                    if (k0 + j < a.E) val = A2s[sI * N + n] * A3s[vI * N + n];
                    h[j] = val; l[j] = tf32_lo(val);
                    if (++vI == I3) { vI = 0; ++sI; }
                }
                *reinterpret_cast<float4*>(bt + (unsigned)n * 128u + (unsigned)((c ^ (n & 7)) << 4)) = make_float4(h[0], h[1], h[2], h[3]);
                const int nl = n + N;
                *reinterpret_cast<float4*>(bt + (unsigned)nl * 128u + (unsigned)((c ^ (nl & 7)) << 4)) = make_float4(l[0], l[1], l[2], l[3]);
            }
            mbar_wait(fullx + sx, phx);
            if (!(a.dbg & 1)) {
                // lane = plane row of the tile: its 32 floats, lo part, to the TMEM columns of this slot
                const int row = lq * 32 + lane;
                const unsigned char* xr = xring + (size_t)sx * L::X_BYTES + (size_t)row * 128;
                const unsigned taddr = tmem + ((unsigned)(lq * 32) << 16) + (unsigned)(L::ALO_COL + s2 * C2C_PM_A_COLS + (C2C_PM_A_COLS - 32));
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    const float4 v0 = *reinterpret_cast<const float4*>(xr + ((c ^ (row & 7)) << 4));
                    const float4 v1 = *reinterpret_cast<const float4*>(xr + (((c + 1) ^ (row & 7)) << 4));
                    const float lo8[8] = {tf32_lo(v0.x), tf32_lo(v0.y), tf32_lo(v0.z), tf32_lo(v0.w),
                                          tf32_lo(v1.x), tf32_lo(v1.y), tf32_lo(v1.z), tf32_lo(v1.w)};
                    tmem_st8(taddr + c * 4, lo8);
#ifdef C2C_PM_ALL_TS
                    const float hi8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};      // the hardware truncates: hi(x)
                    tmem_st8(taddr - 32 + c * 4, hi8);
#endif
                }
                tmem_st_wait();
            }
#ifdef C2C_PM_ALL_TS
            __syncwarp();
            if (lane == 0) mbar_arrive(emptyx + sx);             // this warp is done with the X tile
#endif
            proxy_fence_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ready + s2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, L::TM_COLS);
}

// --------------------------------------------------------------------------------------------------------------------
// fiber_mma
// --------------------------------------------------------------------------------------------------------------------
#define C2C_FM_MAXTILE 8
// C2C_FM_ALL_TS (default): as in plane_mma, hi(X) and lo(X) both go through TMEM (16 columns per tile and prep slot), every MMA
// takes its A operand from TMEM, the X ring slot is released by the prep warps.  The TMEM budget then decides the depth of
// the prep ring (`ring2`): (512 - accumulator columns) / (16 columns x tiles), 2..4.  -DC2C_FM_SMEM_A restores the
// shared-memory A operand (MN-major descriptors over the TMA image).
// It is a template parameter (ALLTS): validated for N = 16 (prep ring of 3).  The N = 32 instance only fits TMEM with a prep
// ring of 2 and gave wrong sums / did not finish: three prep groups take the stages round-robin, so with a 2-deep ring the
// group that has finished stage t + 1 moves on to stage t + 4 -- the slot of stage t, two laps later -- while the group of
// stage t may still be waiting for that slot; both wait on the same mbarrier parity, the later one passes on the completion
// of the EARLIER phase and the two groups fill the slot together.  A ring at least as deep as the number of prep groups makes
// every slot-lap wait unambiguous (a group is at most one lap ahead); the launcher refuses shallower rings (c2c_api.cu), and
// N = 32 keeps the shared-memory A operand (ring of 4).

struct FiberMmaArgs {
    long long P; int E, R, ldr;
    LeadInfo lead;
    float* part;               // [nj][E][ldr] partial sums, one per plane range
    const int* done;
    int es;                    // the in-plane positions (na_total 32-wide atoms) are split into `es` parts; part h starts at atom
                               // floor(h * na_total / es) and owns the atoms up to the next part's start
    int na_total;
    int na;                    // atoms fetched per part: ceil(na_total / es) >= 4 (a part may own one atom less)
    int ntile;                 // 128-position tiles per part: ceil(na / 4) <= C2C_FM_MAXTILE; the last one is shifted back to stay inside the box
    int nx;                    // X ring depth
    int dbg;                   // timing experiments only (see PlaneMmaArgs)
    int epoch;                 // stages per accumulator chain
    long long nst_total;       // 8-plane stages: ceil(P / 8)
    int ring2;                 // depth of the prep -> MMA ring (B tile in shared memory + hi/lo(X) columns in TMEM), <= C2C_MMA_RING2
    int sums_in_smem;          // 1: the CTA's running sums live in shared memory ([na * 32][ldr] floats after the B ring) and are
                               // copied out at the end; 0: they are kept in the (L2-resident) global partial itself
    int boxes2d;               // 1: the tensor map is 2-D {E, P} and a stage is `na` boxes of {32, 8} (in-plane size not a multiple of 32
                               // floats: the ragged last atom is zero-filled by the TMA unit); 0: one 3-D box per stage
    float4* zero_ptr;          // optional: buffer this pass clears for the next plane_mma (T, no longer read once the leading
    long long zero_n4;         // modes are updated), in float4 units; the epilogue warps do it while the stream runs
};

// STACK: hi(x) hi(w) and hi(x) lo(w) in one MMA over the stacked B tile (needs 2N accumulator columns per tile)
template <int N, bool STACK, bool ALLTS>
__global__ void __launch_bounds__(C2C_MMA_THREADS, 1)
fiber_mma_kernel(const __grid_constant__ CUtensorMap mapX, const FiberMmaArgs a)
{
    if (run_done(a.done)) return;
    const int D2 = a.ring2;
    constexpr int ACC_COLS = STACK ? 2 * N : N;                  // accumulator columns per tile
    constexpr int A_COLS = ALLTS ? 16 : 8;                       // TMEM columns per tile and prep slot: hi | lo, or lo only
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int DX = a.nx, na = a.na, ntile = a.ntile, R = a.R, ldr = a.ldr, F = a.epoch;
    const int x_bytes = na * 1024;                               // 8 planes x na atoms x 128 B
    constexpr int B_BYTES = 2 * N * 32;                          // rows 0..N-1 hi(W^T), N..2N-1 lo(W^T); 8 planes each, K-major core matrices
    unsigned char* xring = base;
    unsigned char* bring = base + (size_t)DX * x_bytes;
    float* spart = reinterpret_cast<float*>(bring + (size_t)C2C_MMA_RING2 * B_BYTES);  // [na * 32][ldr] when sums_in_smem
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(spart + (a.sums_in_smem ? (size_t)na * 32 * ldr : 0));
    unsigned long long *fullx = bars, *emptyx = bars + DX, *ready = bars + 2 * DX, *empty2 = ready + C2C_MMA_RING2, *tfull = empty2 + C2C_MMA_RING2, *tempty = tfull + C2C_FM_MAXTILE;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + C2C_FM_MAXTILE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned alo_col = (unsigned)(ntile * ACC_COLS);       // hi/lo(X) ring: D2 x ntile x A_COLS columns
    const unsigned tm_cols = c2c_tmem_cols(alo_col + (unsigned)(D2 * ntile * A_COLS));

    if (tid == 0) {
        for (int s = 0; s < DX; ++s) { mbar_init(fullx + s, 1); mbar_init(emptyx + s, ALLTS ? 4 : 1); }
        for (int s = 0; s < D2; ++s) { mbar_init(ready + s, 4); mbar_init(empty2 + s, 1); }
        for (int s = 0; s < C2C_FM_MAXTILE; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tm_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    const int h = blockIdx.x % a.es, j = blockIdx.x / a.es, nj = gridDim.x / a.es;
    const int at0 = (int)((long long)h * a.na_total / a.es);                 // first atom of this part
    const int na_own = (int)((long long)(h + 1) * a.na_total / a.es) - at0;    // atoms it owns (<= na)
    const long long sb = (long long)j * a.nst_total / nj, se = (long long)(j + 1) * a.nst_total / nj;

    if (warp == 0) {
        if (a.boxes2d) {
            int sx = 0; unsigned phx = 0;
            for (long long g = sb; g < se; ++g) {
                mbar_wait(emptyx + sx, phx ^ 1u);
                if (lane == 0) mbar_expect_tx(fullx + sx, (unsigned)x_bytes);
                __syncwarp();
                for (int at = lane; at < na; at += 32)               // one {32 positions x 8 planes} box per atom and lane
                    tma_load_2d(xring + (size_t)sx * x_bytes + (size_t)at * 1024, &mapX, (at0 + at) * 32, (int)(g * 8), fullx + sx);
                if (++sx == DX) { sx = 0; phx ^= 1u; }
            }
        } else if (lane == 0) {
            int sx = 0; unsigned phx = 0;
            for (long long g = sb; g < se; ++g) {
                mbar_wait(emptyx + sx, phx ^ 1u);
                mbar_expect_tx(fullx + sx, (unsigned)x_bytes);
                tma_load_3d(xring + (size_t)sx * x_bytes, &mapX, 0, (int)(g * 8), at0, fullx + sx);
                if (++sx == DX) { sx = 0; phx ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: one accumulator per tile; a chain is an epoch of F stages, drained tile by tile ----
        {
            constexpr unsigned IDESC2 = umma_idesc_tf32(2 * N, 1), IDESC1 = umma_idesc_tf32(N, 1);   // X^T from shared memory: MN-major A
            constexpr unsigned IDESC1T = umma_idesc_tf32(N, 0);                                       // lo(X) from TMEM
            const unsigned long long DKA = umma_desc(0, 1024, 512, C2C_UMMA_LAYOUT_SW128_32B);
            const unsigned long long DKB = umma_desc(0, 128, 256, C2C_UMMA_LAYOUT_NONE);
            const unsigned xa0 = smem_u32(xring) >> 4, ba0 = smem_u32(bring) >> 4;
            const int last_at = (na & 3) == 0 ? 4 * (ntile - 1) : na - 4;
            int sx = 0; unsigned phx = 0; int s2 = 0; unsigned ph2 = 0; unsigned phe = 0;
            long long g = sb;
            while (g < se) {
                const long long ep_end = g + F < se ? g + F : se;
                for (long long gg = g; gg < ep_end; ++gg) {
                    if (!ALLTS) mbar_wait(fullx + sx, phx);      // (all-TS: the issuer never touches the X ring)
                    mbar_wait(ready + s2, ph2);
                    if (gg == g) {
                        for (int i = 0; i < ntile; ++i) mbar_wait(tempty + i, phe ^ 1u);      // the previous chains are drained
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const unsigned long long dxa = DKA + (xa0 + (unsigned)sx * (unsigned)(x_bytes >> 4));
                        const unsigned long long dbh = DKB + (ba0 + (unsigned)s2 * (B_BYTES >> 4));      // rows 0 ..
                        const unsigned long long dbl = dbh + (N / 8) * (256 >> 4);                     // rows N ..
                        const unsigned acc = gg > g ? 1u : 0u;
                        const unsigned alo = tmem + alo_col + (unsigned)(s2 * ntile * A_COLS);
                        const bool last = gg + 1 == ep_end;
#pragma unroll
                        for (int i = 0; i < C2C_FM_MAXTILE; ++i) {
                            if (i < ((a.dbg & 4) ? 0 : ntile)) {
                                const int at = i + 1 < ntile ? 4 * i : last_at;
                                const unsigned long long dxh = dxa + (unsigned)(at * 64);
                                const unsigned d = tmem + (unsigned)(i * ACC_COLS);
                                if constexpr (ALLTS) {
                                    constexpr unsigned IDESC2T = umma_idesc_tf32(2 * N, 0);
                                    const unsigned ahi = alo + i * 16, alw = ahi + 8;
                                    if (STACK) {
                                        umma_tf32_ts(d, ahi, dbh, IDESC2T, acc);                       // [hi(x) hi(w) | hi(x) lo(w)]
                                        umma_tf32_ts(d + N, alw, dbh, IDESC1T, 1u);                    // lo(x) hi(w) -> the lo columns
                                    } else {
                                        umma_tf32_ts(d, alw, dbh, IDESC1T, acc);                       // small terms first
                                        umma_tf32_ts(d, ahi, dbl, IDESC1T, 1u);
                                        umma_tf32_ts(d, ahi, dbh, IDESC1T, 1u);
                                    }
                                } else if (STACK) {
                                    umma_tf32(d, dxh, dbh, IDESC2, acc);                               // [hi(x) hi(w) | hi(x) lo(w)]
                                    umma_tf32_ts(d + N, alo + i * 8, dbh, IDESC1T, 1u);                // lo(x) hi(w) -> the lo columns
                                } else {
                                    umma_tf32_ts(d, alo + i * 8, dbh, IDESC1T, acc);                   // small terms first
                                    umma_tf32(d, dxh, dbl, IDESC1, 1u);
                                    umma_tf32(d, dxh, dbh, IDESC1, 1u);
                                }
                                if (last) umma_commit(tfull + i);
                            }
                        }
                        if ((a.dbg & 4) && last) { for (int i = 0; i < ntile; ++i) umma_commit(tfull + i); }
                        if (!ALLTS) umma_commit(emptyx + sx);
                        umma_commit(empty2 + s2);
                    }
                    __syncwarp();
                    if (++sx == DX) { sx = 0; phx ^= 1u; }
                    if (++s2 == D2) { s2 = 0; ph2 ^= 1u; }
                }
                phe ^= 1u;
                g = ep_end;
            }
        }
    } else if (warp < C2C_MMA_PREP_WARP0) {
        // ---- epilogue: TMEM lane = in-plane position of the tile; every chain is added, round-to-nearest, into the CTA's
        // running sums -- in shared memory when they leave room for >= 5 X slots, else in the (L2-resident) global partial
        // itself; a row is only ever touched by its own thread, in a fixed order ----
        const int lg = warp & 3;
        unsigned phe = 0;
        bool first = true;
        pdl_wait();
        if (a.zero_ptr != nullptr) {
            const long long per = (a.zero_n4 + gridDim.x - 1) / gridDim.x;
            const long long z0 = (long long)blockIdx.x * per, z1 = z0 + per < a.zero_n4 ? z0 + per : a.zero_n4;
            for (long long i = z0 + (warp - 2) * 32 + lane; i < z1; i += 128) a.zero_ptr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (se <= sb) {                                          // empty plane range: the partial still has to be defined
            const int q4 = ldr >> 2, t128 = (warp - 2) * 32 + lane;
            for (int it = t128; it < na_own * 32 * q4; it += 128) {
                const long long e = (long long)at0 * 32 + it / q4;
                if (e < a.E) *reinterpret_cast<float4*>(a.part + ((size_t)j * a.E + (size_t)e) * ldr + (it % q4) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (long long g = sb; g < se; ) {
            const long long ep_end = g + F < se ? g + F : se;
            for (int i = 0; i < ntile; ++i) {
                const bool shifted = !(i + 1 < ntile || (na & 3) == 0);
                const int el = (shifted ? na - 4 : 4 * i) * 32 + lg * 32 + lane;   // position inside the part
                const long long e = (long long)at0 * 32 + el;
                // rows the previous tile already covers, rows of the next part (fetched, not owned) and rows past the end are skipped
                const bool on = !(shifted && el < 128 * (ntile - 1)) && el < na_own * 32 && e < a.E;
                float* row = a.sums_in_smem ? spart + (size_t)el * ldr : a.part + ((size_t)j * a.E + (size_t)(on ? e : 0)) * ldr;
                // the running sums of this row are fetched before the chain is waited for (they do not depend on it)
                float4 old[N / 4];
                if (on && !first) {
#pragma unroll
                    for (int qq = 0; qq < N / 4; ++qq)
                        if (4 * qq < ldr) old[qq] = *reinterpret_cast<const float4*>(row + 4 * qq);
                }
                mbar_wait(tfull + i, phe);
                tc_fence_after();
                // every MMA into the full-size columns truncated the running sum: ~2^-24 of it per MMA on average (measured)
                const float unbias = 1.0f + C2C_MMA_TRUNC_LOSS * (float)((STACK ? 1 : 3) * (int)(ep_end - g));
                if (!(a.dbg & 8)) {
#pragma unroll
                    for (int c0 = 0; c0 < N; c0 += 16) {
                        float v[16];
                        tmem_ld16(tmem + ((unsigned)(lg * 32) << 16) + (unsigned)(i * ACC_COLS + c0), v);
                        if (STACK) {
                            float vl[16];
                            tmem_ld16(tmem + ((unsigned)(lg * 32) << 16) + (unsigned)(i * ACC_COLS + N + c0), vl);
#pragma unroll
                            for (int qq = 0; qq < 16; ++qq) v[qq] = fmaf(v[qq], unbias, vl[qq]);
                        } else {
#pragma unroll
                            for (int qq = 0; qq < 16; ++qq) v[qq] *= unbias;
                        }
                        if (on) {
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq) {
                                if (c0 + 4 * qq < ldr) {
                                    float4 o = make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]);
                                    if (!first) { const float4 od = old[c0 / 4 + qq]; o.x += od.x; o.y += od.y; o.z += od.z; o.w += od.w; }
                                    *reinterpret_cast<float4*>(row + c0 + 4 * qq) = o;
                                }
                            }
                        }
                    }
                } else if (on && first) {
                    for (int c = 0; c < ldr; ++c) row[c] = 0.0f;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + i);
            }
            phe ^= 1u;
            first = false;
            g = ep_end;
        }
        if (a.sums_in_smem && se > sb) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int q4 = ldr >> 2, t128 = (warp - 2) * 32 + lane;
            for (int it = t128; it < na_own * 32 * q4; it += 128) {
                const int el = it / q4, c = it - el * q4;
                const long long e = (long long)at0 * 32 + el;
                if (e < a.E) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!(a.dbg & 8)) v = *reinterpret_cast<const float4*>(spart + (size_t)el * ldr + c * 4);
                    *reinterpret_cast<float4*>(a.part + ((size_t)j * a.E + (size_t)e) * ldr + c * 4) = v;
                }
            }
        }
    } else {
        // ---- prep: group q takes stages q, q + 3, ..: the 8 Khatri-Rao rows (hi | lo) into the B ring, lo(X) into TMEM ----
        const int q = (warp - C2C_MMA_PREP_WARP0) >> 2, lq = warp & 3;
        const int tg = ((warp - C2C_MMA_PREP_WARP0) & 3) * 32 + lane;
        const long long nitem = se - sb;
        pdl_wait();                          // the leading factors come from the previous kernel
        for (long long t = q; t < nitem; t += C2C_MMA_PREP_GROUPS) {
            const long long g = sb + t;
            const int sx = (int)(t % DX), s2 = (int)(t % D2);
            const unsigned phx = (unsigned)((t / DX) & 1), ph2 = (unsigned)((t / D2) & 1);
            unsigned char* bt = bring + (size_t)s2 * B_BYTES;
            // the stage's 8 Khatri-Rao rows: the factor loads are issued before the slot is waited for (N * 8 <= 2 * 128)
            float wv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int it = tg + u * 128;
                wv[u] = 0.0f;
                if (it < ((a.dbg & 2) ? 0 : N * 8)) {
                    const int n = it % N, k = it / N;
                    long long rem = g * 8 + k;
                    float w = (rem < a.P && n < R) ? 1.0f : 0.0f;
                    if (w != 0.0f) {
                        for (int m = a.lead.nl - 1; m >= 0; --m) {
                            const int dm = a.lead.dim[m];
                            const long long qq = rem / dm;
                            const int idx = (int)(rem - qq * dm);
                            rem = qq;
                            w *= __ldg(a.lead.fac[m] + (size_t)idx * ldr + n);
                        }
                    }
                    wv[u] = w;
                }
            }
            mbar_wait(empty2 + s2, ph2 ^ 1u);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int it = tg + u * 128;
                if (it < ((a.dbg & 2) ? 0 : N * 8)) {
                    const int n = it % N, k = it / N, nl = n + N;
                    *reinterpret_cast<float*>(bt + (unsigned)(n >> 3) * 256u + (unsigned)(k >> 2) * 128u + (unsigned)(n & 7) * 16u + (unsigned)(k & 3) * 4u) = wv[u];
                    *reinterpret_cast<float*>(bt + (unsigned)(nl >> 3) * 256u + (unsigned)(k >> 2) * 128u + (unsigned)(nl & 7) * 16u + (unsigned)(k & 3) * 4u) = tf32_lo(wv[u]);
                }
            }
            mbar_wait(fullx + sx, phx);
            if (!(a.dbg & 1)) {
                // lane = in-plane position of the tile: its 8 planes, lo part, to the tile's 8 TMEM columns of this slot
                const unsigned char* xst = xring + (size_t)sx * x_bytes;
                for (int i = 0; i < ntile; ++i) {
                    const int at = (i + 1 < ntile || (na & 3) == 0) ? 4 * i : na - 4;
                    const int ei = lane;                                            // position inside the atom
                    const unsigned char* xa = xst + (size_t)(at + lq) * 1024;
                    float lo8[8], hi8[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float v = *reinterpret_cast<const float*>(xa + k * 128 + ((((ei >> 3) ^ (k & 3))) << 5) + (ei & 7) * 4);
                        lo8[k] = tf32_lo(v);
                        hi8[k] = v;                                                 // the hardware truncates: hi(x)
                    }
                    const unsigned ta = tmem + ((unsigned)(lq * 32) << 16) + alo_col + (unsigned)(s2 * ntile * A_COLS + i * A_COLS);
                    if (ALLTS) { tmem_st8(ta, hi8); tmem_st8(ta + 8, lo8); }
                    else tmem_st8(ta, lo8);
                }
                tmem_st_wait();
            }
            if (ALLTS) {
                __syncwarp();
                if (lane == 0) mbar_arrive(emptyx + sx);         // this warp is done with the stage's planes
            }
            proxy_fence_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ready + s2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, tm_cols);
}

// launchers (c2c_mma.cu); n = MMA N in {16, 32}
cudaError_t c2c_launch_plane_mma(int n, const CUtensorMap& map, const PlaneMmaArgs& a, int grid, size_t smem, cudaStream_t st, int pdl);
cudaError_t c2c_launch_fiber_mma(int n, int stack, int allts, const CUtensorMap& map, const FiberMmaArgs& a, int grid, size_t smem, cudaStream_t st, int pdl);
size_t c2c_plane_mma_smem(int n, int I2, int I3, int nx);
size_t c2c_fiber_mma_smem(int n, int na, int ldr, int nx);       // ldr = 0: running sums not in shared memory
// tensor maps over X: [P][E] rows of 32 floats (128-byte swizzle) / {32, P, E/32} (32-byte-atom 128-byte swizzle)
int c2c_make_plane_map(CUtensorMap* map, const float* X, long long P, long long E);
int c2c_make_fiber_map(CUtensorMap* map, const float* X, long long P, long long E, int na);
int c2c_make_fiber_map2d(CUtensorMap* map, const float* X, long long P, long long E);
int c2c_mma_available();
