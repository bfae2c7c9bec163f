// Read-bandwidth ceilings on B200 for the access patterns the streaming kernels use (not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/membench scripts/membench.cu && build/membench
// Patterns, all reading the same 768 MB fp32 buffer (60 x 2000 x 40 x 40) once per launch:
//   ldg        grid-stride LDG.128 (ld.global.cs), 8 loads in flight per thread, sum
//   warpring   one persistent CTA per SM, every warp owns a ring of `depth` stages of `stage` bytes filled by
//              cp.async.bulk (mbarrier complete_tx); touch = 0: lane 0 reads one word per stage, touch = 1: the warp reads
//              the whole stage with LDS.128 and sums it
//   ctaring    one persistent CTA per SM, a producer warp fills a CTA-wide ring; consumers read it whole (touch 1) or not
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

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
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(256) ldg_kernel(const float4* __restrict__ x, long long n4, float* out)
{
    float s = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcs(x + i + k * stride);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    for (; i < n4; i += stride) { const float4 v = __ldcs(x + i); s += v.x + v.y + v.z + v.w; }
    if (s == 123.456f) out[0] = s;
}

// contiguous = 1: a CTA owns a contiguous range of stages (its warps interleave inside it); 0: stages are dealt round robin
__global__ void __launch_bounds__(256, 1) warpring_kernel(const float* __restrict__ x, long long nstage, int stage_floats,
                                                           int depth, int touch, int contiguous, float* out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
    float* st0 = reinterpret_cast<float*>(smem + 1024) + (size_t)warp * depth * stage_floats;
    unsigned long long* mybar = bars + warp * depth;
    long long first, stride, end;
    if (contiguous) {
        const long long per = (nstage + gridDim.x - 1) / gridDim.x;
        first = blockIdx.x * per + warp; stride = nw; end = min(nstage, (blockIdx.x + 1) * per);
    } else { first = (long long)blockIdx.x * nw + warp; stride = (long long)gridDim.x * nw; end = nstage; }
    if (lane == 0) { for (int j = 0; j < depth; ++j) mbar_init(mybar + j, 1); mbar_init_fence(); }
    __syncwarp();
    const unsigned bytes = stage_floats * 4;
    long long pg = first; int pslot = 0;
    auto issue = [&]() {
        if (lane == 0) { mbar_expect_tx(mybar + pslot, bytes); bulk_g2s(st0 + (size_t)pslot * stage_floats, x + pg * stage_floats, bytes, mybar + pslot); }
        pg += stride; if (++pslot == depth) pslot = 0;
    };
    for (int j = 0; j < depth && pg < end; ++j) issue();
    float s = 0.f; int cslot = 0; unsigned ph = 0;
    for (long long g = first; g < end; g += stride) {
        mbar_wait(mybar + cslot, ph);
        const float* p = st0 + (size_t)cslot * stage_floats;
        if (touch) { for (int i = lane * 4; i < stage_floats; i += 128) { const float4 v = *reinterpret_cast<const float4*>(p + i); s += v.x + v.y + v.z + v.w; } }
        else if (lane == 0) s += p[0];
        __syncwarp();
        if (pg < end) issue();
        if (++cslot == depth) { cslot = 0; ph ^= 1u; }
    }
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(288, 1) ctaring_kernel(const float* __restrict__ x, long long nstage, int stage_floats,
                                                          int depth, int touch, int chunks, float* out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, ncw = (blockDim.x >> 5) - 1;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem);
    unsigned long long* empty = full + depth;
    float* st0 = reinterpret_cast<float*>(smem + 1024);
    const long long per = (nstage + gridDim.x - 1) / gridDim.x;
    const long long sb = blockIdx.x * per, se = min(nstage, sb + per);
    if (threadIdx.x == 0) { for (int j = 0; j < depth; ++j) { mbar_init(full + j, 1); mbar_init(empty + j, ncw); } mbar_init_fence(); }
    __syncthreads();
    const unsigned bytes = stage_floats * 4;
    if (warp == ncw) {
        int slot = 0; unsigned ph = 0; bool wrapped = false;
        const int cf = stage_floats / chunks;
        for (long long g = sb; g < se; ++g) {
            if (wrapped) mbar_wait(empty + slot, ph ^ 1u);
            if (lane == 0) mbar_expect_tx(full + slot, bytes);
            __syncwarp();
            if (lane < chunks) bulk_g2s(st0 + (size_t)slot * stage_floats + lane * cf, x + g * stage_floats + lane * cf, cf * 4, full + slot);
            if (++slot == depth) { slot = 0; ph ^= 1u; wrapped = true; }
        }
        return;
    }
    float s = 0.f; int cslot = 0; unsigned ph = 0;
    for (long long g = sb; g < se; ++g) {
        mbar_wait(full + cslot, ph);
        const float* p = st0 + (size_t)cslot * stage_floats;
        if (touch) { for (int i = threadIdx.x * 4; i < stage_floats; i += ncw * 128) { const float4 v = *reinterpret_cast<const float4*>(p + i); s += v.x + v.y + v.z + v.w; } }
        else if (threadIdx.x == 0) s += p[0];
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + cslot);
        if (++cslot == depth) { cslot = 0; ph ^= 1u; }
    }
    if (s == 123.456f) out[0] = s;
}

template <class F> static float time_ms(F f, int reps)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main()
{
    const long long n = 60LL * 2000 * 40 * 40;
    float* x; float* out;
    CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(x, 0, n * 4));
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    const double gb = n * 4 / 1e9;
    for (int mult : {4, 8, 16, 32}) {
        const float ms = time_ms([&] { ldg_kernel<<<nsm * mult, 256>>>((const float4*)x, n / 4, out); }, 20);
        printf("{\"pattern\": \"ldg\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n", mult, ms, gb / ms * 1e3);
    }
    CK(cudaFuncSetAttribute(warpring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(ctaring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (int touch : {0, 1})
        for (int contiguous : {0, 1})
            for (int nw : {4, 8})
                for (int stage_b : {4800, 9600, 19200, 25600})
                    for (int depth : {2, 3, 4, 6}) {
                        const size_t smem = 1024 + (size_t)nw * depth * stage_b;
                        if (smem > 216 * 1024) continue;
                        const long long nstage = n * 4 / stage_b;
                        const float ms = time_ms([&] { warpring_kernel<<<nsm, nw * 32, smem>>>(x, nstage, stage_b / 4, depth, touch, contiguous, out); }, 10);
                        printf("{\"pattern\": \"warpring\", \"touch\": %d, \"contiguous\": %d, \"warps\": %d, \"stage\": %d, \"depth\": %d, \"inflight_kb\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n",
                               touch, contiguous, nw, stage_b, depth, (int)(nw * depth * stage_b / 1024), ms, gb / ms * 1e3);
                    }
    for (int touch : {0, 1})
        for (int stage_b : {25600, 51200, 64000})
            for (int depth : {2, 3, 4})
                for (int chunks : {1, 4, 8}) {
                    const size_t smem = 1024 + (size_t)depth * stage_b;
                    if (smem > 216 * 1024) continue;
                    const long long nstage = n * 4 / stage_b;
                    const float ms = time_ms([&] { ctaring_kernel<<<nsm, 288, smem>>>(x, nstage, stage_b / 4, depth, touch, chunks, out); }, 10);
                    printf("{\"pattern\": \"ctaring\", \"touch\": %d, \"stage\": %d, \"depth\": %d, \"chunks\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n",
                           touch, stage_b, depth, chunks, ms, gb / ms * 1e3);
                }
    return 0;
}
