// recon_sums, streaming version (unmasked, in-plane row length a multiple of 4):
//     sum (x - rec), sum (x - rec)^2, sum x, sum x^2        rec[p, s, v] = sum_q W[p, q] A2[s, q] A3[v, q]
// -- the directly accumulated ||X - rec|| the factorization keeps as its final error (tensor.py:311-315) and the sums of
// explained_variance (tensor.py:659-688).  One read of X at streaming speed:
//   * a thread owns ONE float4 position (row s, columns v0..v0+3) of the plane for the whole kernel, so its A2 row and its four
//     A3 rows live in registers (5 * NQ floats) and a reconstruction costs NQ multiplies + 4 * NQ FMAs per 4 elements with
//     no shared-memory traffic besides the NQ / 4 broadcast reads of the plane's Khatri-Rao row W[p, :];
//   * `nslots` thread groups take alternate planes; every thread keeps U = 6..8 float4 loads in flight (48-64 KB per SM with 512
//     threads): a register quad is re-loaded for the next batch as soon as it has been consumed; the W rows of a batch of
//     nslots * U planes are staged in shared memory (double-buffered) one batch ahead;
//   * fp32 sums over the 4 * U elements of a batch, fp64 across batches, fixed-order block / grid reduction.
#include "c2c_common.cuh"
#include "c2c_sums.cuh"


template <int NQ, int THREADS, int U>
__global__ void __launch_bounds__(THREADS, 1)
recon_sums_stream_kernel(const float* __restrict__ X, const LeadInfo lead, const float* __restrict__ A2,
                         const float* __restrict__ A3, long long P, int I2, int I3, int R, int ldr, int nquads, int nslots,
                         int nq_total, double* __restrict__ sums_part)
{
    extern __shared__ __align__(16) float sW[];                      // [2][nslots * U][NQ]
    __shared__ double red[4][THREADS / 32];
    const int t = threadIdx.x;
    // blockIdx.y = part of the plane's float4 positions this CTA owns (planes with more positions than a CTA has threads are
    // split over gridDim.y CTAs, each reading only its share of every plane): positions [blockIdx.y * nquads, ... + nquads)
    const int slot = t / nquads, pos = blockIdx.y * nquads + (t - slot * nquads);
    const bool on = slot < nslots && pos < nq_total;
    const int k4 = I3 >> 2;
    const int s = pos / k4, v0 = (pos - s * k4) << 2;
    const size_t E = (size_t)I2 * I3;
    float a2[NQ];
    float2 a3lo[NQ], a3hi[NQ];                                       // (A3[v0], A3[v0+1]) and (A3[v0+2], A3[v0+3]) per rank
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const bool in = on && q < R;
        a2[q] = in ? __ldg(A2 + (size_t)s * ldr + q) : 0.0f;
        a3lo[q] = in ? make_float2(__ldg(A3 + (size_t)(v0 + 0) * ldr + q), __ldg(A3 + (size_t)(v0 + 1) * ldr + q)) : make_float2(0.f, 0.f);
        a3hi[q] = in ? make_float2(__ldg(A3 + (size_t)(v0 + 2) * ldr + q), __ldg(A3 + (size_t)(v0 + 3) * ldr + q)) : make_float2(0.f, 0.f);
    }
    const int batch = nslots * U;
    const long long stride = (long long)gridDim.x * batch;
    long long p0 = (long long)blockIdx.x * batch;
    auto stage = [&](float* dst, long long pb) {                     // Khatri-Rao rows of the planes of one batch
        for (int i = t; i < batch * NQ; i += THREADS) {
            const int pl = i / NQ, q = i - pl * NQ;
            const long long p = pb + pl;
            dst[i] = (p < P && q < R) ? lead_weight_col(p, lead, ldr, q, -1) : 0.0f;
        }
    };
    auto loadx = [&](int u, long long pb) -> float4 {                // planes past the end: x = 0 and W = 0 -> no contribution
        const long long p = pb + (long long)u * nslots + slot;
        return (on && p < P) ? __ldcs(reinterpret_cast<const float4*>(X + (size_t)p * E) + pos) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) x[u] = loadx(u, p0);
    if (p0 < P) stage(sW, p0);
    double acc_d = 0.0, acc_dd = 0.0, acc_x = 0.0, acc_xx = 0.0;
    int cur = 0;
    // steady state: the loads of the NEXT batch are issued as soon as a register quad has been consumed, and its W rows are
    // staged (other buffer) after the compute, so both latencies run under the loads in flight; one barrier per batch
    for (; p0 < P; p0 += stride) {
        __syncthreads();
        const long long pn = p0 + stride;
        if (on) {
            const float* wbase = sW + (size_t)cur * batch * NQ;
            float fd = 0.f, fdd = 0.f, fx = 0.f, fxx = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float4* w4 = reinterpret_cast<const float4*>(wbase + (size_t)(u * nslots + slot) * NQ);
                float2 r01 = make_float2(0.f, 0.f), r23 = make_float2(0.f, 0.f);
#pragma unroll
                for (int q4 = 0; q4 < NQ / 4; ++q4) {
                    const float4 w = w4[q4];
                    const float wq[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int q = q4 * 4 + i;
                        const float wa = wq[i] * a2[q];
                        const float2 ww = make_float2(wa, wa);
                        r01 = ffma2(ww, a3lo[q], r01);
                        r23 = ffma2(ww, a3hi[q], r23);
                    }
                }
                const float4 xv = x[u];
                x[u] = loadx(u, pn);
                const float d0 = xv.x - r01.x, d1 = xv.y - r01.y, d2 = xv.z - r23.x, d3 = xv.w - r23.y;
                fd += (d0 + d1) + (d2 + d3);
                fdd = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, fdd))));
                fx += (xv.x + xv.y) + (xv.z + xv.w);
                fxx = fmaf(xv.x, xv.x, fmaf(xv.y, xv.y, fmaf(xv.z, xv.z, fmaf(xv.w, xv.w, fxx))));
            }
            acc_d += (double)fd; acc_dd += (double)fdd; acc_x += (double)fx; acc_xx += (double)fxx;
        }
        if (pn < P) stage(sW + (size_t)(cur ^ 1) * batch * NQ, pn);
        cur ^= 1;
    }
    // fixed-order block reduction: shuffle tree per warp, then the per-warp values in order
    double v[4] = {acc_d, acc_dd, acc_x, acc_xx};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double wsum = warp_sum(v[k]);
        if ((t & 31) == 0) red[k][t >> 5] = wsum;
    }
    __syncthreads();
    const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    if (t < 4) {
        double tot = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) tot += red[t][w];
        sums_part[blk * 5 + t] = tot;
    }
    if (t == 4) sums_part[blk * 5 + 4] = blk == 0 ? (double)P * (double)E : 0.0;
}

template <int NQ, int THREADS, int U>
static cudaError_t launch_one(const ContractArgs& a, int nq_total, int grid_max, cudaStream_t st, int* grid_out)
{
    const int nparts = (nq_total + THREADS - 1) / THREADS;           // CTAs sharing a plane
    const int nquads = (nq_total + nparts - 1) / nparts;             // positions per CTA
    grid_max /= nparts;
    if (grid_max < 1) return cudaErrorNotSupported;
    int nslots_max = THREADS / nquads;
    const int smem_cap = (int)((48 * 1024) / (2 * U * NQ * sizeof(float)));     // both W buffers within 48 KB
    if (nslots_max > smem_cap) nslots_max = smem_cap;
    // no more slots than give every CTA of a full grid at least one batch
    long long nslots = nslots_max;
    const long long per_cta = (a.P + grid_max - 1) / grid_max;
    if (nslots * U > per_cta) nslots = (per_cta + U - 1) / U;
    if (nslots < 1) nslots = 1;
    const long long batch = nslots * U;
    long long grid = (a.P + batch - 1) / batch;
    if (grid > grid_max) grid = grid_max;
    const size_t smem = 2 * (size_t)batch * NQ * sizeof(float);
    if (smem > 48 * 1024) return cudaErrorNotSupported;
    recon_sums_stream_kernel<NQ, THREADS, U><<<dim3((unsigned)grid, (unsigned)nparts), THREADS, smem, st>>>(
        a.X, a.lead, a.A2, a.A3, a.P, a.I2, a.I3, a.R, a.ldr, nquads, (int)nslots, nq_total, a.sums);
    *grid_out = (int)grid * nparts;
    return cudaGetLastError();
}

// Returns cudaErrorNotSupported when the shape / rank is outside this kernel (the caller then takes the general kernel).
cudaError_t c2c_launch_sums_stream(const ContractArgs& a, int grid_max, cudaStream_t st, int* grid_out)
{
    if (a.mask != nullptr || a.r0 != 0 || a.p0 != 0 || (a.I3 & 3) != 0 || a.R > 32 || a.R < 1) return cudaErrorNotSupported;
    if (((uintptr_t)a.X & 15) != 0) return cudaErrorNotSupported;
    const long long nq = (long long)a.I2 * (a.I3 >> 2);
    const int nq4 = (a.R + 3) / 4;
    if (nq > 8 * 512) return cudaErrorNotSupported;                   // at most 8 CTAs per plane (in-plane size <= 128 x 128)
    if (nq4 <= 3) {
        switch (nq4) {
            case 1: return launch_one<4, 512, 8>(a, (int)nq, grid_max, st, grid_out);
            case 2: return launch_one<8, 512, 8>(a, (int)nq, grid_max, st, grid_out);
            default: return launch_one<12, 512, 6>(a, (int)nq, grid_max, st, grid_out);
        }
    }
    switch (nq4) {
        case 4: return launch_one<16, 256, 8>(a, (int)nq, grid_max, st, grid_out);
        case 5: return launch_one<20, 256, 8>(a, (int)nq, grid_max, st, grid_out);
        case 6: return launch_one<24, 256, 8>(a, (int)nq, grid_max, st, grid_out);
        case 7: return launch_one<28, 256, 8>(a, (int)nq, grid_max, st, grid_out);
        default: return launch_one<32, 256, 8>(a, (int)nq, grid_max, st, grid_out);
    }
}
