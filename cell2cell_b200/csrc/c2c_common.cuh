// Shared definitions for the cell2cell_b200 CUDA kernels (sm_100a only).
//
// Vocabulary: the tensor X[I_0, ..., I_{N-1}] (C order, fp32) is seen as P "planes" of
// E = I_{N-2} x I_{N-1} floats: the N-2 *leading* modes (contexts, ligand-receptor pairs) index the plane, the two
// *trailing* modes (sender, receiver cells) index inside it.  R = rank, ldr = factor row pitch (R rounded up to 4,
// padding columns are zero), NP2 = ceil(R / 2) = number of rank pairs a kernel instance is compiled for.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define C2C_MAX_DIMS 8
#define C2C_MAX_LEAD (C2C_MAX_DIMS - 2)
#define C2C_MAX_NP2 16            // fast kernels cover R <= 32 per rank chunk

struct LeadInfo {
    int nl;                                   // number of leading modes (N - 2 >= 1)
    int dim[C2C_MAX_LEAD];                    // their sizes, outermost first
    const float* fac[C2C_MAX_LEAD];           // their factor matrices [dim x ldr]
};

struct ContractArgs {
    const float* X;            // [P, I2, I3]
    const uint8_t* mask;       // same shape, 1 = observed, 0 = missing (imputed with the current reconstruction); may be null
    long long P;               // number of planes
    long long p0;              // first plane this launch handles (the register-staged kernels finish the tail of a streamed pass)
    int I2, I3;                // trailing mode sizes
    int R;                     // rank
    int ldr;                   // factor pitch (floats), multiple of 4, >= R
    int r0;                    // first rank column handled by this launch (rank chunking for R > 32)
    const float* A2;           // [I2 x ldr]
    const float* A3;           // [I3 x ldr]
    const float* Bmat;         // [I2*I3 x ldr]: A2[s,:] * A3[v,:]  (masked fiber contraction only)
    LeadInfo lead;
    float* out;                // plane_contract: T [P x ldr];  fiber_contract: partials [nparts x E x ldr]
    double* sums;              // recon_sums: per-CTA partial sums
    int sums_missing;          // recon_sums: sum over the MISSING entries (mask == 0) with x taken as 0 -> sums[1] = ||(1 - m) rec||^2
    const int* done;           // device flag: non-zero => the run has converged, kernels return immediately (may be null)
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    // packed fp32x2 FMA (sm_100 FFMA2): half the issue slots of two FFMAs
    return __ffma2_rn(a, b, c);
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor in the stream is still running; pdl_wait() blocks until that predecessor has completed and its
// writes are visible (a no-op without the attribute), pdl_trigger() lets the successor's CTAs be scheduled from now on.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool run_done(const int* done) {
    return done != nullptr && *reinterpret_cast<const volatile int*>(done) != 0;
}

// Decompose plane index p into the leading-mode indices and multiply the corresponding factor rows:
// w[r] = prod_k A_k[i_k(p), r0 + r]  for r in [0, n)   (n <= 32)
template <int N>
__device__ __forceinline__ void lead_weights(float (&w)[N], long long p, const LeadInfo& lead, int ldr, int r0, int R) {
#pragma unroll
    for (int r = 0; r < N; ++r) w[r] = 1.0f;
    long long rem = p;
#pragma unroll 1
    for (int k = lead.nl - 1; k >= 0; --k) {
        const int d = lead.dim[k];
        const long long q = rem / d;
        const int i = (int)(rem - q * d);
        rem = q;
        const float* row = lead.fac[k] + (size_t)i * ldr + r0;
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const float a = (r0 + r < R) ? __ldg(row + r) : 0.0f;
            w[r] *= a;
        }
    }
}

// scalar version used by the small kernels: w_r for one column
__device__ __forceinline__ float lead_weight_col(long long p, const LeadInfo& lead, int ldr, int r, int skip) {
    float w = 1.0f;
    long long rem = p;
    for (int k = lead.nl - 1; k >= 0; --k) {
        const int d = lead.dim[k];
        const long long q = rem / d;
        const int i = (int)(rem - q * d);
        rem = q;
        if (k != skip) w *= __ldg(lead.fac[k] + (size_t)i * ldr + r);
    }
    return w;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// host: launch with (pdl != 0) or without the programmatic-stream-serialization attribute
template <typename... KArgs, typename... Args>
static inline cudaError_t c2c_launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
