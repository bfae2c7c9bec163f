// Issue-rate probes on B200 (not part of the product library):
//   mma_tf32   legacy warp-level mma.sync.m16n8k8 tf32 (fp32 accumulate), NACC independent accumulators per warp
//   ffma2      packed fp32x2 FMA, 32 independent accumulator pairs per thread
//   ffma       scalar FFMA, 64 independent accumulators per thread
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mmabench scripts/mmabench.cu && build/mmabench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) mma_tf32_kernel(int iters, float* out, unsigned a_seed)
{
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a0 = a_seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) ffma2_kernel(int iters, float* out, float seed)
{
    float2 acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = make_float2(0.f, 0.f);
    float2 x = make_float2(seed + threadIdx.x, seed), b = make_float2(seed * 0.5f, seed * 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = __ffma2_rn(x, b, acc[i]);
        x.x += 1e-9f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i].x + acc[i].y;
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) ffma_kernel(int iters, float* out, float seed)
{
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    float x = seed + threadIdx.x, b = seed * 0.5f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = fmaf(x, b, acc[i]);
        x += 1e-9f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

template <class F> static float time_ms(F f)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main()
{
    float* out; CK(cudaMalloc(&out, 4));
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        {
            const float ms = time_ms([&] { mma_tf32_kernel<8><<<nsm, warps * 32>>>(iters, out, 1u); });
            const double flop = 2.0 * 16 * 8 * 8 * 8 * (double)iters * warps * nsm;
            printf("{\"probe\": \"mma_tf32\", \"warps_per_sm\": %d, \"nacc\": 8, \"ms\": %.3f, \"tflops\": %.1f, \"mma_per_clk_per_sm_at_1.9GHz\": %.3f}\n",
                   warps, ms, flop / ms / 1e9, 8.0 * iters * warps / (ms * 1e-3 * 1.9e9));
        }
        {
            const float ms = time_ms([&] { mma_tf32_kernel<12><<<nsm, warps * 32>>>(iters, out, 1u); });
            const double flop = 2.0 * 16 * 8 * 8 * 12 * (double)iters * warps * nsm;
            printf("{\"probe\": \"mma_tf32\", \"warps_per_sm\": %d, \"nacc\": 12, \"ms\": %.3f, \"tflops\": %.1f}\n", warps, ms, flop / ms / 1e9);
        }
        {
            const float ms = time_ms([&] { ffma2_kernel<<<nsm, warps * 32>>>(iters, out, 1.f); });
            const double flop = 2.0 * 2 * 32 * 32 * (double)iters * warps * nsm;
            printf("{\"probe\": \"ffma2\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.1f, \"ffma2_per_clk_per_smsp_at_1.9GHz\": %.3f}\n",
                   warps, ms, flop / ms / 1e9, 32.0 * iters * warps / 4 / (ms * 1e-3 * 1.9e9));
        }
        {
            const float ms = time_ms([&] { ffma_kernel<<<nsm, warps * 32>>>(iters, out, 1.f); });
            const double flop = 2.0 * 64 * 32 * (double)iters * warps * nsm;
            printf("{\"probe\": \"ffma\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.1f, \"ffma_per_clk_per_smsp_at_1.9GHz\": %.3f}\n",
                   warps, ms, flop / ms / 1e9, 64.0 * iters * warps / 4 / (ms * 1e-3 * 1.9e9));
        }
    }
    return 0;
}
