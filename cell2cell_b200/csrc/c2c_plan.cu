// Mask plan of the masked factorization: see c2c_plan.cuh for the model.  Three groups of kernels:
//   plan_*            built once per (tensor, mask): separable model g / hs / hv, residual counts, offsets, the two lists
//   *_ctx_gram, *_corr_gram   the contraction of the reconstruction over the model-missing set, as restricted Grams
//   wmat, corr_rows   the same contraction over the residual lists (entries the separable model cannot express)
// All of them are exact restatements of tensorly's imputation `tensor*mask + cp_to_tensor(cp, mask=1-mask)` followed by
// unfolding_dot_khatri_rao (cell2cell/tensor/coupled_factorization.py:342-345): the same products, summed in another order.
#include "c2c_plan.cuh"
#include <cstdlib>

static inline size_t plan_align(size_t x) { return (x + 255) / 256 * 256; }

size_t c2c_plan_header_bytes(long long P, int E, int I2, int I3, int C0)
{
    const long long nchunks = (P + C2C_PLAN_CHUNK - 1) / C2C_PLAN_CHUNK;
    return plan_align(C2C_PLAN_NCOUNTERS * sizeof(unsigned long long)) + plan_align((size_t)P) +
           plan_align((size_t)C0 * I2 * sizeof(int)) + plan_align((size_t)C0 * I3 * sizeof(int)) +
           plan_align((size_t)(P + 1) * sizeof(int)) + plan_align((size_t)(nchunks * E + 1) * sizeof(int));
}

size_t c2c_plan_lists_bytes(long long n_resid, long long pos_slots)
{
    return plan_align((size_t)(n_resid > 0 ? n_resid : 1) * sizeof(uint16_t)) +
           plan_align((size_t)(pos_slots > 0 ? pos_slots : 1) * sizeof(uint16_t));
}

void c2c_plan_view(PlanView& v, long long P, int E, int I2, int I3, int C0, void* header, void* lists, long long n_resid)
{
    v.P = P; v.E = E; v.I2 = I2; v.I3 = I3; v.C0 = C0; v.ppc0 = P / C0;
    v.nchunks = (int)((P + C2C_PLAN_CHUNK - 1) / C2C_PLAN_CHUNK);
    char* h = (char*)header;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = h ? h + off : nullptr; off += plan_align(bytes); return p; };
    v.counters = (unsigned long long*)take(C2C_PLAN_NCOUNTERS * sizeof(unsigned long long));
    v.g = (uint8_t*)take((size_t)P);
    v.hs = (int*)take((size_t)C0 * I2 * sizeof(int));
    v.hv = (int*)take((size_t)C0 * I3 * sizeof(int));
    v.row_off = (int*)take((size_t)(P + 1) * sizeof(int));
    v.pos_off = (int*)take((size_t)((long long)v.nchunks * E + 1) * sizeof(int));
    char* l = (char*)lists;
    v.csr_ent = (uint16_t*)l;
    v.pos_ent = l ? (uint16_t*)(l + plan_align((size_t)(n_resid > 0 ? n_resid : 1) * sizeof(uint16_t))) : nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------------------------------------------------
// a warp per plane: g[p], the row / column flags of its context (idempotent stores of 1), the missing count, the zero-fill check
__global__ void __launch_bounds__(256)
plan_fit_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask, const PlanView v)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long nmiss = 0;
    int bad = 0;
    for (long long p = warp; p < v.P; p += nwarps) {
        const int c = (int)(p / v.ppc0);
        const uint8_t* mp = mask + (size_t)p * v.E;
        const float* xp = X ? X + (size_t)p * v.E : nullptr;
        bool any = false;
        for (int v0 = 0; v0 < v.I3; v0 += 32) {
            const int vv = v0 + lane;
            bool colany = false;
            for (int s = 0; s < v.I2; ++s) {
                bool obs = false;
                if (vv < v.I3) {
                    const int e = s * v.I3 + vv;
                    obs = __ldg(mp + e) != 0;
                    if (!obs) {
                        ++nmiss;
                        if (xp && __ldg(xp + e) != 0.0f) bad = 1;
                    }
                }
                const unsigned rowany = __ballot_sync(0xffffffffu, obs);
                if (rowany && lane == 0 && v.hs[c * v.I2 + s] == 0) v.hs[c * v.I2 + s] = 1;
                colany |= obs;
            }
            if (colany && v.hv[c * v.I3 + vv] == 0) v.hv[c * v.I3 + vv] = 1;
            any |= colany;
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) v.g[p] = any ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmiss += __shfl_xor_sync(0xffffffffu, nmiss, o);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        if (nmiss) atomicAdd(v.counters + 0, nmiss);
        if (bad) v.counters[2] = 1ull;
    }
}

// residual = missing in the mask, observed by the model (the caller has checked g[p])
__device__ __forceinline__ bool plan_resid_sv(const PlanView& v, const uint8_t* mp, int c, int s, int vv)
{
    return __ldg(mp + s * v.I3 + vv) == 0 && v.hs[c * v.I2 + s] != 0 && v.hv[c * v.I3 + vv] != 0;
}

// a warp per plane: residual entries of the plane -> row_off[p + 1]
__global__ void __launch_bounds__(256)
plan_count_rows_kernel(const uint8_t* __restrict__ mask, const PlanView v)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < v.P; p += nwarps) {
        int n = 0;
        if (v.g[p]) {
            const int c = (int)(p / v.ppc0);
            const uint8_t* mp = mask + (size_t)p * v.E;
            for (int e = lane; e < v.E; e += 32) {
                const int s = e / v.I3;
                n += plan_resid_sv(v, mp, c, s, e - s * v.I3) ? 1 : 0;
            }
        }
        // The list of a plane is stored by CLASS c = position mod 8, slot i of the list holding an entry of class i mod 8
        // (0xFFFF where a class has run out): the eight lanes of a quarter-warp then gather B rows 48 B apart modulo 8 rows,
        // i.e. from eight different 16-byte bank groups -- every LDS.128 of plane_corr_list is conflict-free.  Lane l scans
        // the positions l, l + 32, ...: all of class l mod 8.  Length = 8 x the longest class.
        n += __shfl_xor_sync(0xffffffffu, n, 8);
        n += __shfl_xor_sync(0xffffffffu, n, 16);                      // every lane: the count of its class
        int tot = n;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
            n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
        }
        if (lane == 0) {
            v.row_off[p + 1] = 8 * n;
            if (tot) atomicAdd(v.counters + 6, (unsigned long long)tot);   // residual entries, without the padding
        }
    }
}

// a warp per (chunk of planes, group of 32 positions), lane = position: the list of (chunk, position) holds the chunk's planes
// that are residual at the position, stored by CLASS = plane number mod 8 exactly as the by-plane lists are (so that the
// gathers of W rows from shared memory are conflict-free too); slots = 8 x the longest class -> pos_off[k E + e + 1]
__global__ void __launch_bounds__(256)
plan_count_pos_kernel(const uint8_t* __restrict__ mask, const PlanView v)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int ngroups = (v.E + 31) / 32;
    const long long nblocks = (long long)v.nchunks * ngroups;
    for (long long b = warp; b < nblocks; b += nwarps) {
        const int k = (int)(b / ngroups), gq = (int)(b - (long long)k * ngroups);
        const int e = gq * 32 + lane;
        if (e >= v.E) continue;
        const int s = e / v.I3, vv = e - s * v.I3;
        const long long p0 = (long long)k * C2C_PLAN_CHUNK;                    // a multiple of 8: class of plane p0 + j is j mod 8
        const long long p1 = p0 + C2C_PLAN_CHUNK < v.P ? p0 + C2C_PLAN_CHUNK : v.P;
        int cnt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cnt[j] = 0;
        for (long long pb = p0; pb < p1; pb += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long p = pb + j;
                if (p < p1 && v.g[p]) cnt[j] += plan_resid_sv(v, mask + (size_t)p * v.E, (int)(p / v.ppc0), s, vv) ? 1 : 0;
            }
        }
        int m = cnt[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) m = max(m, cnt[j]);
        v.pos_off[(size_t)k * v.E + e + 1] = 8 * m;
    }
}

// in-place inclusive scan of a[1 .. n] (a[0] = 0) by one CTA of 1024 threads; returns the total to every thread
__device__ unsigned long long plan_block_scan(int* a, long long n, unsigned long long* sh /*[1024]*/)
{
    const int t = threadIdx.x;
    const long long seg = (n + 1023) / 1024;
    const long long b = 1 + (long long)t * seg, e = (b + seg < n + 1) ? b + seg : n + 1;
    unsigned long long s = 0;
    for (long long i = b; i < e; ++i) s += (unsigned long long)a[i];
    sh[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long x = t >= o ? sh[t - o] : 0ull;
        __syncthreads();
        sh[t] += x;
        __syncthreads();
    }
    unsigned long long run = sh[t] - s;
    for (long long i = b; i < e; ++i) { run += (unsigned long long)a[i]; a[i] = (int)run; }
    const unsigned long long total = sh[1023];
    __syncthreads();
    if (t == 0) a[0] = 0;
    return total;
}

__global__ void __launch_bounds__(1024)
plan_scan_kernel(const PlanView v)
{
    __shared__ unsigned long long sh[1024];
    const unsigned long long n_resid = plan_block_scan(v.row_off, v.P, sh);
    const unsigned long long pos_slots = plan_block_scan(v.pos_off, (long long)v.nchunks * v.E, sh);
    int missing = 0;
    for (long long i = threadIdx.x; i < v.P; i += 1024) missing |= v.g[i] == 0;
    for (int i = threadIdx.x; i < v.C0 * v.I2; i += 1024) missing |= v.hs[i] == 0;
    for (int i = threadIdx.x; i < v.C0 * v.I3; i += 1024) missing |= v.hv[i] == 0;
    missing = __syncthreads_or(missing);
    if (threadIdx.x == 0) {
        v.counters[1] = n_resid;
        v.counters[3] = pos_slots;
        v.counters[4] = missing ? 1ull : 0ull;
        v.counters[5] = (n_resid >= (1ull << 31) || pos_slots >= (1ull << 31)) ? 1ull : 0ull;
    }
}

cudaError_t c2c_plan_scan(const float* X, const uint8_t* mask, const PlanView& v, int nsm, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(v.counters, 0, c2c_plan_header_bytes(v.P, v.E, v.I2, v.I3, v.C0), st);
    if (e != cudaSuccess) return e;
    long long grid = (v.P + 7) / 8;
    if (grid > (long long)nsm * 16) grid = (long long)nsm * 16;
    plan_fit_kernel<<<(unsigned)grid, 256, 0, st>>>(X, mask, v);
    plan_count_rows_kernel<<<(unsigned)grid, 256, 0, st>>>(mask, v);
    long long g2 = ((long long)v.nchunks * ((v.E + 31) / 32) + 7) / 8;
    if (g2 > (long long)nsm * 16) g2 = (long long)nsm * 16;
    plan_count_pos_kernel<<<(unsigned)g2, 256, 0, st>>>(mask, v);
    plan_scan_kernel<<<1, 1024, 0, st>>>(v);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// phase 2
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
plan_fill_csr_kernel(const uint8_t* __restrict__ mask, const PlanView v)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < v.P; p += nwarps) {
        const int off = v.row_off[p];
        const int len = v.row_off[p + 1] - off;
        if (len == 0) continue;
        const int c = (int)(p / v.ppc0);
        const uint8_t* mp = mask + (size_t)p * v.E;
        const int cls = lane & 7;                                      // class of every position this lane scans
        const unsigned cmask = 0x01010101u << cls;                     // the four lanes of the class
        int cnt = 0;                                                   // entries of the class placed so far
        for (int e0 = 0; e0 < v.E; e0 += 32) {
            const int e = e0 + lane;
            bool r = false;
            if (e < v.E) { const int s = e / v.I3; r = plan_resid_sv(v, mp, c, s, e - s * v.I3); }
            const unsigned bal = __ballot_sync(0xffffffffu, r) & cmask;
            if (r) v.csr_ent[off + 8 * (cnt + __popc(bal & ((1u << lane) - 1u))) + cls] = (uint16_t)e;
            cnt += __popc(bal);
        }
        for (int k = cnt + (lane >> 3); k < len / 8; k += 4) v.csr_ent[off + 8 * k + cls] = (uint16_t)0xFFFFu;
    }
}

__global__ void __launch_bounds__(256)
plan_fill_pos_kernel(const uint8_t* __restrict__ mask, const PlanView v)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int ngroups = (v.E + 31) / 32;
    const long long nblocks = (long long)v.nchunks * ngroups;
    for (long long b = warp; b < nblocks; b += nwarps) {
        const int k = (int)(b / ngroups), gq = (int)(b - (long long)k * ngroups);
        const int e = gq * 32 + lane;
        if (e >= v.E) continue;
        const int off = v.pos_off[(size_t)k * v.E + e];
        const int len8 = (v.pos_off[(size_t)k * v.E + e + 1] - off) / 8;
        if (len8 == 0) continue;
        const int s = e / v.I3, vv = e - s * v.I3;
        const long long p0 = (long long)k * C2C_PLAN_CHUNK;
        const long long p1 = p0 + C2C_PLAN_CHUNK < v.P ? p0 + C2C_PLAN_CHUNK : v.P;
        int cnt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cnt[j] = 0;
        for (long long pb = p0; pb < p1; pb += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long p = pb + j;
                if (p < p1 && v.g[p] && plan_resid_sv(v, mask + (size_t)p * v.E, (int)(p / v.ppc0), s, vv)) {
                    v.pos_ent[(size_t)off + 8 * cnt[j] + j] = (uint16_t)(p - p0);
                    ++cnt[j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            for (int c = cnt[j]; c < len8; ++c) v.pos_ent[(size_t)off + 8 * c + j] = (uint16_t)0xFFFFu;
    }
}

cudaError_t c2c_plan_fill(const uint8_t* mask, const PlanView& v, int nsm, cudaStream_t st)
{
    long long grid = (v.P + 7) / 8;
    if (grid > (long long)nsm * 16) grid = (long long)nsm * 16;
    plan_fill_csr_kernel<<<(unsigned)grid, 256, 0, st>>>(mask, v);
    long long g2 = ((long long)v.nchunks * ((v.E + 31) / 32) + 7) / 8;
    if (g2 > (long long)nsm * 16) g2 = (long long)nsm * 16;
    plan_fill_pos_kernel<<<(unsigned)g2, 256, 0, st>>>(mask, v);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Gram form of the correction over the model-missing set
// ---------------------------------------------------------------------------------------------------------------------
// float4 of the Khatri-Rao row of plane p: columns 4 * quad .. 4 * quad + 3
__device__ __forceinline__ float4 lead_quad(long long p, const LeadInfo& lead, int ldr, int quad)
{
    float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
    long long rem = p;
    for (int k = lead.nl - 1; k >= 0; --k) {
        const int d = lead.dim[k];
        const long long q = rem / d;
        const int i = (int)(rem - q * d);
        rem = q;
        const float4 f = *(reinterpret_cast<const float4*>(lead.fac[k] + (size_t)i * ldr) + quad);
        w.x *= f.x; w.y *= f.y; w.z *= f.z; w.w *= f.w;
    }
    return w;
}

// CTA c < C0: D_c = G2_all o G3_miss + G2_miss o G3_obs  (= G2_all o G3_all - G2_obs o G3_obs, without the cancellation);
// CTA C0: D = G2_all o G3_all.   G2_obs[q, r] = sum over the rows s with hs[c, s] = 1 of A2[s, q] A2[s, r], etc.
// The two factors and the context's flags are staged in shared memory (the sums are latency chains otherwise).
__global__ void __launch_bounds__(256)
trail_ctx_gram_kernel(const PlanView v, const float* __restrict__ A2, const float* __restrict__ A3, int R, int ldr,
                      float* __restrict__ Dmat, const int* done)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(done)) return;
    extern __shared__ __align__(16) float tcg_smem[];
    float* s2 = tcg_smem;                                   // [I2][ldr]
    float* s3 = s2 + (size_t)v.I2 * ldr;                    // [I3][ldr]
    int* f2 = reinterpret_cast<int*>(s3 + (size_t)v.I3 * ldr);   // [I2] row observed in this context
    int* f3 = f2 + v.I2;                                    // [I3]
    const int c = blockIdx.x;
    for (int t = threadIdx.x; t < v.I2 * ldr; t += blockDim.x) s2[t] = A2[t];
    for (int t = threadIdx.x; t < v.I3 * ldr; t += blockDim.x) s3[t] = A3[t];
    for (int t = threadIdx.x; t < v.I2; t += blockDim.x) f2[t] = c < v.C0 ? v.hs[c * v.I2 + t] : 0;
    for (int t = threadIdx.x; t < v.I3; t += blockDim.x) f3[t] = c < v.C0 ? v.hv[c * v.I3 + t] : 0;
    __syncthreads();
    float* D = Dmat + (size_t)c * ldr * ldr;
    for (int t = threadIdx.x; t < ldr * ldr; t += blockDim.x) {
        const int q = t / ldr, r = t - q * ldr;
        float d = 0.0f;
        if (q < R && r < R) {
            float g2o = 0.f, g2m = 0.f, g3o = 0.f, g3m = 0.f;
            for (int s = 0; s < v.I2; ++s) {
                const float x = s2[s * ldr + q] * s2[s * ldr + r];
                if (f2[s]) g2o += x; else g2m += x;
            }
            for (int s = 0; s < v.I3; ++s) {
                const float x = s3[s * ldr + q] * s3[s * ldr + r];
                if (f3[s]) g3o += x; else g3m += x;
            }
            d = (g2o + g2m) * g3m + g2m * g3o;
        }
        D[t] = d;
    }
}

cudaError_t c2c_launch_trail_ctx_gram(const PlanView& v, const float* A2, const float* A3, int R, int ldr, float* Dmat,
                                      const int* done, cudaStream_t st, int pdl)
{
    const size_t smem = (size_t)(v.I2 + v.I3) * ldr * sizeof(float) + (size_t)(v.I2 + v.I3) * sizeof(int);
    if (smem > 48 * 1024) {
        const cudaError_t ea = cudaFuncSetAttribute(trail_ctx_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (ea != cudaSuccess) return ea;
    }
    return c2c_launch_ex(trail_ctx_gram_kernel, dim3(v.C0 + 1), dim3(256), smem, st, pdl, v, A2, A3, R, ldr, Dmat, done);
}

// CTA = (context c, batch): blockDim / (ldr / 4) consecutive planes of ONE context, so D_c and D_all sit in shared memory.
// thread = (plane slot, column quad): the slot's Khatri-Rao row goes through shared memory, then 4 R FMAs.
__global__ void __launch_bounds__(256)
plane_corr_gram_kernel(const PlanView v, const LeadInfo lead, const float* __restrict__ Dmat, const float* __restrict__ T0,
                       float* __restrict__ Tw, int R, int ldr, int parts, const int* done)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(done)) return;
    extern __shared__ __align__(16) float pcg_smem[];
    const int q4 = ldr >> 2, LL = ldr * ldr;
    const int npp = blockDim.x / q4;                                   // plane slots
    float* sD = pcg_smem;                                              // [ldr][ldr] D_c
    float* sF = sD + LL;                                               // [ldr][ldr] D_all
    float* sW = sF + LL;                                               // [npp][ldr]
    const int c = blockIdx.x / parts, part = blockIdx.x - c * parts;
    const int slot = threadIdx.x / q4, quad = threadIdx.x - slot * q4;
    const long long p = (long long)c * v.ppc0 + (long long)part * npp + slot;
    const bool on = slot < npp && p < (long long)(c + 1) * v.ppc0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int gp = 0;
    if (on) {
        acc = __ldg(reinterpret_cast<const float4*>(T0 + (size_t)p * ldr) + quad);
        gp = v.g[p];
        *(reinterpret_cast<float4*>(sW + (size_t)slot * ldr) + quad) = lead_quad(p, lead, ldr, quad);
    }
    for (int t = threadIdx.x; t < LL; t += blockDim.x) {
        sD[t] = Dmat[(size_t)c * LL + t];
        sF[t] = Dmat[(size_t)v.C0 * LL + t];
    }
    __syncthreads();
    if (on) {
        const float* D = gp ? sD : sF;
        const float* w = sW + (size_t)slot * ldr;
        for (int q = 0; q < R; ++q) {
            const float wq = w[q];
            const float4 d = *(reinterpret_cast<const float4*>(D + (size_t)q * ldr) + quad);
            acc.x = fmaf(wq, d.x, acc.x); acc.y = fmaf(wq, d.y, acc.y);
            acc.z = fmaf(wq, d.z, acc.z); acc.w = fmaf(wq, d.w, acc.w);
        }
        *(reinterpret_cast<float4*>(Tw + (size_t)p * ldr) + quad) = acc;
    }
}

cudaError_t c2c_launch_plane_corr_gram(const PlanView& v, const LeadInfo& lead, const float* Dmat, const float* T0, float* Tw,
                                       int R, int ldr, const int* done, int nsm, cudaStream_t st, int pdl)
{
    (void)nsm;
    const int q4 = ldr / 4, npp = 256 / q4;
    const long long parts = (v.ppc0 + npp - 1) / npp;
    const size_t smem = ((size_t)2 * ldr * ldr + (size_t)npp * ldr) * sizeof(float);
    return c2c_launch_ex(plane_corr_gram_kernel, dim3((unsigned)(v.C0 * parts)), dim3(256), smem, st, pdl, v, lead, Dmat, T0, Tw, R,
                         ldr, (int)parts, done);
}

// CTA = (context c, part): Yo = sum over its planes with g = 1 of W[p] (x) W[p], Ym = the same over g = 0.
// thread = (plane subgroup, 4 x 4 tile of the R x R matrix); the W rows of a batch of planes (normally the CTA's whole share)
// go through shared memory.
__global__ void __launch_bounds__(256)
lead_ctx_gram_kernel(const PlanView v, const LeadInfo lead, int R, int ldr, int parts, int batch, float* __restrict__ Ypart,
                     const int* done)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(done)) return;
    extern __shared__ __align__(16) float lcg_smem[];
    const int q4 = ldr >> 2, ntile = q4 * q4, LL = ldr * ldr;
    const int nsub = blockDim.x / ntile;                               // >= 4 (ntile <= 64)
    float* sW = lcg_smem;                                              // [batch][ldr]
    int* sg = reinterpret_cast<int*>(sW + (size_t)batch * ldr);        // [batch]
    float* sRed = reinterpret_cast<float*>(sg + batch);                // [nsub][2][LL]
    const int c = blockIdx.x / parts, part = blockIdx.x - c * parts;
    const long long per = (v.ppc0 + parts - 1) / parts;
    const long long p_begin = (long long)c * v.ppc0 + (long long)part * per;
    long long p_end = p_begin + per;
    if (p_end > (long long)(c + 1) * v.ppc0) p_end = (long long)(c + 1) * v.ppc0;
    const int sub = threadIdx.x / ntile, tile = threadIdx.x - sub * ntile;
    const int tq = tile / q4, tr = tile - tq * q4;
    float ao[16], am[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { ao[i] = 0.0f; am[i] = 0.0f; }
    for (long long pb = p_begin; pb < p_end; pb += batch) {
        const int nb = (int)(p_end - pb < batch ? p_end - pb : batch);
        __syncthreads();
        for (int t = threadIdx.x; t < nb * q4; t += blockDim.x) {
            const int pl = t / q4, quad = t - pl * q4;
            *(reinterpret_cast<float4*>(sW + (size_t)pl * ldr) + quad) = lead_quad(pb + pl, lead, ldr, quad);
            if (quad == 0) sg[pl] = v.g[pb + pl];
        }
        __syncthreads();
        if (sub < nsub) {
            for (int pl = sub; pl < nb; pl += nsub) {
                const float4 a = *(reinterpret_cast<const float4*>(sW + (size_t)pl * ldr) + tq);
                const float4 b = *(reinterpret_cast<const float4*>(sW + (size_t)pl * ldr) + tr);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                if (sg[pl]) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) ao[i * 4 + j] = fmaf(av[i], bv[j], ao[i * 4 + j]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) am[i * 4 + j] = fmaf(av[i], bv[j], am[i * 4 + j]);
                }
            }
        }
    }
    __syncthreads();
    if (sub < nsub) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = (tq * 4 + i) * ldr + tr * 4 + j;
                sRed[((size_t)sub * 2 + 0) * LL + idx] = ao[i * 4 + j];
                sRed[((size_t)sub * 2 + 1) * LL + idx] = am[i * 4 + j];
            }
    }
    __syncthreads();
    float* out = Ypart + ((size_t)c * parts + part) * 2 * LL;
    for (int t = threadIdx.x; t < 2 * LL; t += blockDim.x) {
        float s = 0.0f;
        for (int k = 0; k < nsub; ++k) s += sRed[(size_t)k * 2 * LL + t];
        const int idx = t % LL, q = idx / ldr, r = idx - q * ldr;
        out[t] = (q < R && r < R) ? s : 0.0f;
    }
}

// Ymat[c] = sum of the parts' Yo of context c (blocks 0 .. C0 - 1);  Ymat[C0] = sum over all contexts and parts of Ym (the
// remaining blocks: 8 matrix entries each, a warp per entry, lanes stride the partials, fixed shuffle tree)
__global__ void __launch_bounds__(256)
ctx_gram_finalize_kernel(const float* __restrict__ Ypart, int C0, int parts, int LL, float* __restrict__ Ymat, const int* done)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(done)) return;
    const int c = blockIdx.x;
    if (c < C0) {
        for (int t = threadIdx.x; t < LL; t += blockDim.x) {
            float s = 0.0f;
            for (int k = 0; k < parts; ++k) s += Ypart[(((size_t)c * parts + k) * 2 + 0) * LL + t];
            Ymat[(size_t)c * LL + t] = s;
        }
        return;
    }
    const int lane = threadIdx.x & 31;
    const int t = (c - C0) * 8 + (threadIdx.x >> 5);
    if (t >= LL) return;
    float s = 0.0f;
    for (int k = lane; k < C0 * parts; k += 32) s += Ypart[((size_t)k * 2 + 1) * LL + t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) Ymat[(size_t)C0 * LL + t] = s;
}

static int lead_ctx_batch(int ldr) { const int b = 16384 / ldr; return b > 512 ? 512 : b; }

int c2c_lead_ctx_parts(int C0, int nsm)
{
    int parts = (2 * nsm + C0 - 1) / C0;
    return parts < 1 ? 1 : parts;
}

cudaError_t c2c_launch_lead_ctx_gram(const PlanView& v, const LeadInfo& lead, int R, int ldr, float* Ypart, float* Ymat,
                                     const int* done, int nsm, cudaStream_t st, int pdl)
{
    const int parts = c2c_lead_ctx_parts(v.C0, nsm);
    const int q4 = ldr / 4, ntile = q4 * q4, nsub = 256 / ntile, LL = ldr * ldr;
    const long long per = (v.ppc0 + parts - 1) / parts;
    int batch = lead_ctx_batch(ldr);
    if (batch > per) batch = (int)(per < 1 ? 1 : per);
    const size_t smem = (size_t)batch * ldr * sizeof(float) + (size_t)batch * sizeof(int) + (size_t)nsub * 2 * LL * sizeof(float);
    if (smem > 48 * 1024) {
        const cudaError_t ea = cudaFuncSetAttribute(lead_ctx_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        if (ea != cudaSuccess) return ea;
    }
    cudaError_t e = c2c_launch_ex(lead_ctx_gram_kernel, dim3((unsigned)(v.C0 * parts)), dim3(256), smem, st, pdl, v, lead, R, ldr, parts,
                                  batch, Ypart, done);
    if (e != cudaSuccess) return e;
    return c2c_launch_ex(ctx_gram_finalize_kernel, dim3((unsigned)(v.C0 + (LL + 7) / 8)), dim3(256), 0, st, pdl, (const float*)Ypart,
                         v.C0, parts, LL, Ymat, done);
}

#define C2C_PLAN_DISPATCH(CALL)                                                                                          \
    switch (ldr) {                                                                                                       \
        case 4: CALL(4); break; case 8: CALL(8); break; case 12: CALL(12); break; case 16: CALL(16); break;              \
        case 20: CALL(20); break; case 24: CALL(24); break; case 28: CALL(28); break; case 32: CALL(32); break;          \
        default: return cudaErrorNotSupported;                                                                           \
    }

// a warp per (position e, column quad); lanes take the contexts (and the "missing planes" matrix, index C0):
// Uw[e, quad] = U0[e, quad] + sum_q B[e, q] * Z_e[q, quad],  Z_e = Y_missing + sum over the contexts c whose row s or column
// v the model misses of Y_c.  Lane sums are combined by a fixed shuffle tree.
template <int NQ>
__global__ void __launch_bounds__(256)
fiber_corr_gram_kernel(const PlanView v, const float* __restrict__ A2, const float* __restrict__ A3, const float* __restrict__ Ymat,
                       const float* __restrict__ U0, float* __restrict__ Uw, int R, const int* done)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(done)) return;
    constexpr int q4 = NQ / 4, LL = NQ * NQ;
    const int lane = threadIdx.x & 31;
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= (long long)v.E * q4) return;
    const int e = (int)(t / q4), quad = (int)(t - (long long)e * q4);
    const int s = e / v.I3, vv = e - s * v.I3;
    float b[NQ];
#pragma unroll
    for (int k = 0; k < q4; ++k) {
        const float4 x = *(reinterpret_cast<const float4*>(A2 + (size_t)s * NQ) + k);
        const float4 y = *(reinterpret_cast<const float4*>(A3 + (size_t)vv * NQ) + k);
        b[4 * k] = x.x * y.x; b[4 * k + 1] = x.y * y.y; b[4 * k + 2] = x.z * y.z; b[4 * k + 3] = x.w * y.w;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = lane; c <= v.C0; c += 32) {
        if (c < v.C0 && v.hs[c * v.I2 + s] && v.hv[c * v.I3 + vv]) continue;
        const float4* Y = reinterpret_cast<const float4*>(Ymat + (size_t)c * LL) + quad;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (q < R) {
                const float4 y = Y[q * q4];
                acc.x = fmaf(b[q], y.x, acc.x); acc.y = fmaf(b[q], y.y, acc.y);
                acc.z = fmaf(b[q], y.z, acc.z); acc.w = fmaf(b[q], y.w, acc.w);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (lane == 0) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(U0 + (size_t)e * NQ) + quad);
        *(reinterpret_cast<float4*>(Uw + (size_t)e * NQ) + quad) = make_float4(u.x + acc.x, u.y + acc.y, u.z + acc.z, u.w + acc.w);
    }
}

cudaError_t c2c_launch_fiber_corr_gram(const PlanView& v, const float* A2, const float* A3, const float* Ymat, const float* U0,
                                       float* Uw, int R, int ldr, const int* done, cudaStream_t st, int pdl)
{
    const long long n = (long long)v.E * (ldr / 4);                    // warps
    const dim3 grid((unsigned)((n + 7) / 8));
#define FCG_CALL(N) do { const cudaError_t e_ = c2c_launch_ex(fiber_corr_gram_kernel<N>, grid, dim3(256), 0, st, pdl, v, A2, A3, Ymat, U0, Uw, R, done); \
                         if (e_ != cudaSuccess) return e_; } while (0)
    C2C_PLAN_DISPATCH(FCG_CALL)
#undef FCG_CALL
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------------
// the residual lists
// ---------------------------------------------------------------------------------------------------------------------
// Wmat[p, 4 quad .. 4 quad + 3]: a thread per (plane, column quad); the padding columns of packed factors are zero, so are W's
__global__ void __launch_bounds__(256)
wmat_kernel(const LeadInfo lead, long long P, int ldr, float* __restrict__ Wmat, const int* done)
{
    if (run_done(done)) return;
    const int nq = ldr / 4;
    const long long n = P * nq;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long p = t / nq;
        const int quad = (int)(t - p * nq);
        reinterpret_cast<float4*>(Wmat)[t] = lead_quad(p, lead, ldr, quad);
    }
}

cudaError_t c2c_launch_wmat(const LeadInfo& lead, long long P, int ldr, float* Wmat, const int* done, cudaStream_t st)
{
    const long long n = P * (ldr / 4);
    long long grid = (n + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    return c2c_launch_ex(wmat_kernel, dim3((unsigned)grid), dim3(256), 0, st, 0, lead, P, ldr, Wmat, done);
}

constexpr __host__ __device__ int plan_b_pitch(int nq) { return ((nq / 4) & 1) ? nq : nq + 4; }
constexpr __host__ __device__ int plan_pow2(int n) { return n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }
constexpr __host__ __device__ int plan_lpr(int nq) { return nq <= 16 ? 16 : 32; }     // lanes per list row

// Every lane of a group of LPR lanes holds NP <= LPR values (NP a power of two); returns the sum over the group of value `col`,
// where `col` is fixed by the lane number: at the step of width o the lanes with that bit set keep the upper half of what they
// hold and send the lower half (and vice versa), so the count halves per step; the widths that are left over when one value
// remains are plain butterflies.  Lanes that differ only in those leftover low bits end with the same column.  Fixed tree.
template <int NP, int LPR>
__device__ __forceinline__ float plan_halving_sum(float (&a)[NP], int sl, int& col)
{
    static_assert(NP <= LPR, "one value per lane at the end");
    col = 0;
    int o = LPR / 2;
#pragma unroll
    for (int n = NP; n > 1; n >>= 1) {
        const bool up = (sl & o) != 0;
#pragma unroll
        for (int j = 0; j < n / 2; ++j) {
            const float send = up ? a[j] : a[j + n / 2];
            const float keep = up ? a[j + n / 2] : a[j];
            a[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        if (up) col += n / 2;
        o >>= 1;
    }
    float v = a[0];
    for (; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#define C2C_PLAN_PRELOAD 8
// One list row per group of LPR lanes (two rows per warp at NQ <= 16): sum over the row's entries j of <own, tab[j]> tab[j, :].
// The lanes stride the row's class-ordered list, so the eight lanes of a quarter-warp gather rows of eight different classes
// (row number mod 8): with an odd number of 16-byte quads per table row every LDS.128 is conflict-free.  `n` = this group's
// slots (0: nothing to do), `nmax` = the longest of the warp's rows (the loop is warp-uniform).  Returns the group's sum of
// column `col`.
template <int NQ, int LPR, int PB>
__device__ __forceinline__ float plan_row_sum(const float* tab, const uint16_t* __restrict__ ent, int n, int nmax,
                                              const float* __restrict__ own, int sl, int& col)
{
    float2 w[NQ / 2], acc[NQ / 2];
#pragma unroll
    for (int q = 0; q < NQ / 2; ++q) { w[q] = make_float2(0.0f, 0.0f); acc[q] = make_float2(0.0f, 0.0f); }
    if (n > 0) {
#pragma unroll
        for (int k = 0; k < NQ / 4; ++k) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(own) + k);
            w[2 * k] = make_float2(x.x, x.y); w[2 * k + 1] = make_float2(x.z, x.w);
        }
    }
    for (int i0 = 0; i0 < nmax; i0 += LPR * C2C_PLAN_PRELOAD) {
        int ev[C2C_PLAN_PRELOAD];                                       // the list reads of a batch are in flight together
#pragma unroll
        for (int u = 0; u < C2C_PLAN_PRELOAD; ++u) {
            const int i = i0 + u * LPR + sl;
            ev[u] = i < n ? (int)ent[i] : 0xFFFF;
        }
#pragma unroll
        for (int u = 0; u < C2C_PLAN_PRELOAD; ++u) {
            if (ev[u] == 0xFFFF) continue;                              // padding of a class that has run out / past the row
            const float4* b4 = reinterpret_cast<const float4*>(tab + (size_t)ev[u] * PB);
            float2 b[NQ / 2];
#pragma unroll
            for (int k = 0; k < NQ / 4; ++k) {
                const float4 x = b4[k];
                b[2 * k] = make_float2(x.x, x.y); b[2 * k + 1] = make_float2(x.z, x.w);
            }
            float2 r2 = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int q = 0; q < NQ / 2; ++q) r2 = ffma2(w[q], b[q], r2);
            const float rec = r2.x + r2.y;
            const float2 rr = make_float2(rec, rec);
#pragma unroll
            for (int q = 0; q < NQ / 2; ++q) acc[q] = ffma2(rr, b[q], acc[q]);
        }
    }
    // column sums over the group's lanes by recursive halving (about NQ shuffles instead of 5 NQ)
    constexpr int NP = plan_pow2(NQ);
    float vals[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) vals[q] = 0.0f;
#pragma unroll
    for (int q = 0; q < NQ / 2; ++q) { vals[2 * q] = acc[q].x; vals[2 * q + 1] = acc[q].y; }
    return plan_halving_sum<NP, LPR>(vals, sl, col);
}

// MODE 0 -- rows = planes: own = Wmat[p], table = Bmat (every in-plane position; in shared memory when SMEM_TAB, else read
//           through L1), entries = in-plane positions; Tout[p] = Tin[p] + sum.  1-D grid, the warps stride the planes.
// MODE 1 -- rows = positions of a chunk of planes: own = Bmat[e], table = the chunk's rows of Wmat (always in shared memory),
//           entries = plane numbers within the chunk; CTA (i, j) takes the positions [i epi, (i + 1) epi) and the chunks
//           [j cpj, (j + 1) cpj), restaging the table per chunk; out = part[j][e] (first chunk: store, then accumulate --
//           the same lane every time, so no race and a fixed order).
template <int NQ, int MODE, bool SMEM_TAB>
__global__ void __launch_bounds__(NQ <= 16 ? 512 : 256)
corr_rows_kernel(const PlanView v, const float* __restrict__ Wmat, const float* __restrict__ Bmat, const float* Tin,
                 float* out, int epi, int cpj, const int* done)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(done)) return;
    extern __shared__ __align__(16) float crk_smem[];
    constexpr int LPR = plan_lpr(NQ), NROW = 32 / LPR;
    constexpr int PB = SMEM_TAB ? plan_b_pitch(NQ) : NQ;
    constexpr int NP = plan_pow2(NQ);
    const int lane = threadIdx.x & 31, sl = lane & (LPR - 1), sub = lane / LPR;
    const bool writer = (sl & (LPR / NP - 1)) == 0;
    if (MODE == 0) {
        if (SMEM_TAB) {
            const float4* src = reinterpret_cast<const float4*>(Bmat);
            float4* dst = reinterpret_cast<float4*>(crk_smem);
            for (int i = threadIdx.x; i < v.E * (NQ / 4); i += blockDim.x) {
                const int e = i / (NQ / 4), k = i - e * (NQ / 4);
                dst[e * (PB / 4) + k] = src[i];
            }
            __syncthreads();
        }
        const float* tab = SMEM_TAB ? crk_smem : Bmat;
        const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
        const bool copy = Tin != out;
        for (long long g0 = warp * NROW; g0 < v.P; g0 += nwarps * NROW) {
            const long long p = g0 + sub;
            int b0 = 0, n = 0;
            if (p < v.P) { b0 = v.row_off[p]; n = v.row_off[p + 1] - b0; }
            int nmax = n;
            if (NROW == 2) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, n, 16));
            if (nmax == 0) {
                if (copy && p < v.P && sl < NQ) out[(size_t)p * NQ + sl] = Tin[(size_t)p * NQ + sl];
                continue;
            }
            int col;
            const float mine = plan_row_sum<NQ, LPR, PB>(tab, v.csr_ent + b0, n, nmax, Wmat + (size_t)(p < v.P ? p : 0) * NQ, sl, col);
            if (p < v.P && col < NQ && writer) out[(size_t)p * NQ + col] = Tin[(size_t)p * NQ + col] + mine;
        }
    } else {
        const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const int e_begin = blockIdx.x * epi;
        const int e_end = e_begin + epi < v.E ? e_begin + epi : v.E;
        const int k_begin = blockIdx.y * cpj;
        const int k_end = k_begin + cpj < v.nchunks ? k_begin + cpj : v.nchunks;
        float* part = out + (size_t)blockIdx.y * v.E * NQ;
        for (int k = k_begin; k < k_end; ++k) {
            const long long p0 = (long long)k * C2C_PLAN_CHUNK;
            const int np = (int)(v.P - p0 < C2C_PLAN_CHUNK ? v.P - p0 : C2C_PLAN_CHUNK);
            __syncthreads();
            {
                const float4* src = reinterpret_cast<const float4*>(Wmat + (size_t)p0 * NQ);
                float4* dst = reinterpret_cast<float4*>(crk_smem);
                for (int i = threadIdx.x; i < np * (NQ / 4); i += blockDim.x) {
                    const int pl = i / (NQ / 4), q = i - pl * (NQ / 4);
                    dst[pl * (PB / 4) + q] = src[i];
                }
            }
            __syncthreads();
            const int* off = v.pos_off + (size_t)k * v.E;
            for (int g0 = e_begin + wid * NROW; g0 < e_end; g0 += nw * NROW) {
                const int e = g0 + sub;
                const bool live = e < e_end;
                int b0 = 0, n = 0;
                if (live) { b0 = off[e]; n = off[e + 1] - b0; }
                int nmax = n;
                if (NROW == 2) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, n, 16));
                int col = 0;
                float mine = 0.0f;
                if (nmax > 0) mine = plan_row_sum<NQ, LPR, PB>(crk_smem, v.pos_ent + b0, n, nmax, Bmat + (size_t)(live ? e : 0) * NQ, sl, col);
                else {                                                     // the column this lane would end with (plan_halving_sum)
                    int o = LPR / 2;
                    for (int m = NP; m > 1; m >>= 1) { if (sl & o) col += m / 2; o >>= 1; }
                }
                if (live && col < NQ && writer) {
                    float* dstp = part + (size_t)e * NQ + col;
                    *dstp = k == k_begin ? mine : *dstp + mine;
                }
            }
        }
    }
}

cudaError_t c2c_launch_plane_corr_list(const PlanView& v, const float* Wmat, const float* Bmat, const float* Tin, float* Tout,
                                       int ldr, const int* done, int nsm, cudaStream_t st, int pdl)
{
    const size_t bbytes = (size_t)v.E * plan_b_pitch(ldr) * sizeof(float);
    const bool smem_b = bbytes <= 100 * 1024;
    const int block = ldr <= 16 ? 512 : 256;                           // <= 64 registers per thread at ldr <= 16: 32 warps per SM
    const int nrow = 32 / plan_lpr(ldr);
    long long grid = ((v.P + nrow - 1) / nrow * 32 + block - 1) / block;
    if (grid > (long long)nsm * (smem_b ? 2 : 8)) grid = (long long)nsm * (smem_b ? 2 : 8);
#define PCL_CALL(N)                                                                                                      \
    do {                                                                                                                 \
        cudaError_t e_;                                                                                                  \
        if (smem_b) {                                                                                                    \
            if (bbytes > 48 * 1024) {                                                                                    \
                e_ = cudaFuncSetAttribute(corr_rows_kernel<N, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); \
                if (e_ != cudaSuccess) return e_;                                                                        \
            }                                                                                                            \
            e_ = c2c_launch_ex(corr_rows_kernel<N, 0, true>, dim3((unsigned)grid), dim3(block), bbytes, st, pdl, v, Wmat, Bmat, Tin, Tout, 0, 0, done); \
        } else                                                                                                           \
            e_ = c2c_launch_ex(corr_rows_kernel<N, 0, false>, dim3((unsigned)grid), dim3(block), 0, st, pdl, v, Wmat, Bmat, Tin, Tout, 0, 0, done); \
        if (e_ != cudaSuccess) return e_;                                                                                \
    } while (0)
    C2C_PLAN_DISPATCH(PCL_CALL)
#undef PCL_CALL
    return cudaSuccess;
}

// Uout = Uin + sum of the chunk-range partials, in a fixed order: thread (g, x) sums the partials g, g + 4, ... of float4 x
// (four short chains instead of one long one: the kernel is a latency chain of L2 reads), the four sums are added g = 0..3
__global__ void __launch_bounds__(256)
plan_reduce_kernel(const float* __restrict__ part, int nparts, long long n4, const float* Uin, float* Uout,
                   const int* done)
{
    pdl_trigger();
    pdl_wait();
    if (run_done(done)) return;
    __shared__ float4 sh[3][64];
    const int x = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const long long i = (long long)blockIdx.x * 64 + x;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
        const float4* p4 = reinterpret_cast<const float4*>(part) + i;
        int c = grp;
        for (; c + 12 < nparts; c += 16) {                             // four loads in flight
            const float4 a = p4[(size_t)c * n4], b = p4[(size_t)(c + 4) * n4], d = p4[(size_t)(c + 8) * n4], f = p4[(size_t)(c + 12) * n4];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
            s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
            s.x += f.x; s.y += f.y; s.z += f.z; s.w += f.w;
        }
        for (; c < nparts; c += 4) {
            const float4 a = p4[(size_t)c * n4];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    if (grp > 0) sh[grp - 1][x] = s;
    __syncthreads();
    if (grp == 0 && i < n4) {
#pragma unroll
        for (int g = 0; g < 3; ++g) { const float4 a = sh[g][x]; s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
        const float4 u = reinterpret_cast<const float4*>(Uin)[i];
        reinterpret_cast<float4*>(Uout)[i] = make_float4(u.x + s.x, u.y + s.y, u.z + s.z, u.w + s.w);
    }
}

cudaError_t c2c_launch_fiber_corr_list(const PlanView& v, const float* Wmat, const float* Bmat, const float* Uin, float* Uout,
                                       float* part, int max_parts, int ldr, const int* done, int nsm, cudaStream_t st, int pdl)
{
    const size_t tbytes = (size_t)C2C_PLAN_CHUNK * plan_b_pitch(ldr) * sizeof(float);
    const int block = ldr <= 16 ? 512 : 256;
    // CTAs resident per SM: 64 registers x 512 threads (or ~128 x 256) -> 2 by registers; the staged chunk of W rows by smem
    int per_sm = (int)((220 * 1024) / (tbytes + 1024));
    if (per_sm > 2) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    const int resident = nsm * per_sm;
    {   // C2C_B200_PLAN_MAX_PARTS: fewer chunk ranges than chunks, so that a CTA walks several chunks (tests of that path;
        // it is otherwise only taken beyond ~300 k planes)
        const char* e = getenv("C2C_B200_PLAN_MAX_PARTS");
        if (e && atoi(e) > 0 && atoi(e) < max_parts) max_parts = atoi(e);
    }
    int nparts = v.nchunks < max_parts ? v.nchunks : max_parts;       // chunk ranges (= partials)
    if (nparts > resident) nparts = resident;
    if (nparts < 1) nparts = 1;
    const int cpj = (v.nchunks + nparts - 1) / nparts;
    nparts = (v.nchunks + cpj - 1) / cpj;
    // position ranges: one wave of CTAs in all, or two when that fills the machine better (118 chunks: 2 x 118 of 296 places
    // against 5 x 118 of 2 x 296)
    int ni = resident / nparts;
    {
        const int ni2 = 2 * resident / nparts;
        if (ni < 1 || (double)ni2 * nparts / (2.0 * resident) > (double)ni * nparts / resident + 0.1) ni = ni2;
    }
    const int nrow = 32 / plan_lpr(ldr), rows_per_trip = (block / 32) * nrow;
    const int max_ni = (v.E + rows_per_trip - 1) / rows_per_trip;
    if (ni > max_ni) ni = max_ni;
    if (ni < 1) ni = 1;
    const int epi = (v.E + ni - 1) / ni;
    ni = (v.E + epi - 1) / epi;
    const dim3 grid((unsigned)ni, (unsigned)nparts);
#define FCL_CALL(N)                                                                                                      \
    do {                                                                                                                 \
        cudaError_t e_;                                                                                                  \
        if (tbytes > 48 * 1024) {                                                                                        \
            e_ = cudaFuncSetAttribute(corr_rows_kernel<N, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tbytes); \
            if (e_ != cudaSuccess) return e_;                                                                            \
        }                                                                                                                \
        e_ = c2c_launch_ex(corr_rows_kernel<N, 1, true>, grid, dim3(block), tbytes, st, pdl, v, Wmat, Bmat, (const float*)nullptr, part, epi, cpj, done); \
        if (e_ != cudaSuccess) return e_;                                                                                \
    } while (0)
    C2C_PLAN_DISPATCH(FCL_CALL)
#undef FCL_CALL
    const long long n4 = (long long)v.E * ldr / 4;
    return c2c_launch_ex(plan_reduce_kernel, dim3((unsigned)((n4 + 63) / 64)), dim3(256), 0, st, pdl, (const float*)part, nparts, n4, Uin,
                         Uout, done);
}
