// tcgen05 / TMA layout probe for sm_100a (not part of the product library): checks, on a real B200, every layout and
// descriptor assumption the tensor-core streaming kernels rely on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_probe scripts/umma_probe.cu && build/umma_probe
// Tests
//   tma2d   2D tensor map over G[rows][cols] fp32, box {32, 128}, SWIZZLE_128B  -> smem image == K-major SW128 formula
//   tma3d   3D tensor map over X[P][E] seen as {32 (e inner), P, E/32 (e outer)}, box {32, 8, na}, SWIZZLE_128B
//           -> smem image == MN-major SW128 atoms [e outer][plane][32 e]
//   mmaK    D[128 x 32] = A[128 x 32] * B[32 x 32]^T, A and B K-major SW128 tiles, 4 k-steps (descriptor start + 32 B)
//   mmaMN   D[128 x 32] = A[128 x 8] * B[32 x 8]^T, A MN-major SW128 (4 atoms), B K-major no-swizzle core matrices
//   trunc   what kind::tf32 does with the low 13 mantissa bits of an fp32 operand (truncate / round)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
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
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---- TMA image dumps ----
__global__ void tma2d_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, float* out, int nfloats)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, nfloats * 4); tma_2d(smem, &map, c0, c1, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
__global__ void tma3d_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, float* out, int nfloats)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, nfloats * 4); tma_3d(smem, &map, c0, c1, c2, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

// ---- MMA probe ----
struct MmaArgs {
    const float* A;      // a_mode 0: [128][32] (m, k);  a_mode 1: [128][8] (m, k)
    const float* B;      // b_mode 0: [32][32] (n, k);   b_mode 1: [32][8] (n, k)
    float* D;            // [128][32]
    int a_mode, b_mode;
    unsigned a_lbo, a_sbo, b_lbo, b_sbo;       // bytes
    unsigned a_layout, b_layout;                // descriptor layout_type
    unsigned a_major, b_major;                  // 0 = K, 1 = MN
    int ksteps; unsigned a_step, b_step;        // descriptor start-address advance per k-step, bytes
};

__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo, unsigned sbo, unsigned layout)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;                                   // version = 1 (Blackwell)
    d |= (unsigned long long)(layout & 7) << 61;
    return d;
}

__global__ void __launch_bounds__(128) mma_kernel(const MmaArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    unsigned char* sA = smem;                 // 16 KB region
    unsigned char* sB = smem + 16384;         // 4 KB region
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + 4096) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    __syncthreads();
    if (a.a_mode == 0) {
        for (int i = tid; i < 128 * 32; i += blockDim.x) {
            const int r = i >> 5, k = i & 31;
            const unsigned off = r * 128 + ((((k >> 2) ^ (r & 7))) << 4) + (k & 3) * 4;
            *reinterpret_cast<float*>(sA + off) = a.A[i];
        }
    } else {
        for (int i = tid; i < 128 * 8; i += blockDim.x) {
            const int m = i >> 3, k = i & 7;
            const int at = m >> 5, mi = m & 31;
            const unsigned off = at * 1024 + k * 128 + ((((mi >> 3) ^ (k & 3))) << 5) + (mi & 7) * 4;   // SW128 with 32 B atoms
            *reinterpret_cast<float*>(sA + off) = a.A[i];
        }
    }
    if (a.b_mode == 0) {
        for (int i = tid; i < 32 * 32; i += blockDim.x) {
            const int r = i >> 5, k = i & 31;
            const unsigned off = r * 128 + ((((k >> 2) ^ (r & 7))) << 4) + (k & 3) * 4;
            *reinterpret_cast<float*>(sB + off) = a.B[i];
        }
    } else {
        for (int i = tid; i < 32 * 8; i += blockDim.x) {
            const int n = i >> 3, k = i & 7;
            const unsigned off = (n >> 3) * 256 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4;
            *reinterpret_cast<float*>(sB + off) = a.B[i];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_base_s;
    if (tid == 0) {
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a.a_major << 15) | (a.b_major << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < a.ksteps; ++ks) {
            const unsigned long long da = make_desc(smem_u32(sA) + ks * a.a_step, a.a_lbo, a.a_sbo, a.a_layout);
            const unsigned long long db = make_desc(smem_u32(sB) + ks * a.b_step, a.b_lbo, a.b_sbo, a.b_layout);
            const unsigned acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned v[32];
    const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) a.D[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u));
}

// ---- A operand from TMEM: lane = row, column = k (32-bit each); written with tcgen05.st by the thread that owns the lane ----
__global__ void __launch_bounds__(128) mma_ts_kernel(const float* A /*[128][8]*/, const float* B /*[32][32], k < 8 used*/, float* D)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    unsigned char* sB = smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 4096 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < 32 * 32; i += blockDim.x) {
        const int r = i >> 5, k = i & 31;
        const unsigned off = r * 128 + ((((k >> 2) ^ (r & 7))) << 4) + (k & 3) * 4;
        *reinterpret_cast<float*>(sB + off) = B[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_base_s;
    {
        unsigned v[8];
        for (int k = 0; k < 8; ++k) v[k] = __float_as_uint(A[tid * 8 + k]);
        const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + 32u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const unsigned long long db = make_desc(smem_u32(sB), 16, 1024, 2);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                     ::"r"(tmem), "r"(tmem + 32u), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned v[32];
    const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode()
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
    return (EncodeFn)fn;
}

static int run_mma(const char* name, const std::vector<float>& A, int a_ld, const std::vector<float>& B, int b_ld, int K, MmaArgs a, int verbose)
{
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * 32 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * 32 * 4));
    a.A = dA; a.B = dB; a.D = dD;
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    mma_kernel<<<1, 128, 16384 + 4096 + 1024>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"test\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<float> D(128 * 32);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * a_ld + k] * (double)B[n * b_ld + k];
            const double err = fabs(ref - D[m * 32 + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3) { if (bad < verbose) printf("  %s mismatch m=%d n=%d got %g want %g\n", name, m, n, D[m * 32 + n], ref); ++bad; }
        }
    printf("{\"test\": \"%s\", \"bad\": %d, \"maxerr\": %g, \"a_lbo\": %u, \"a_sbo\": %u, \"b_lbo\": %u, \"b_sbo\": %u}\n", name, bad, maxerr,
           a.a_lbo, a.a_sbo, a.b_lbo, a.b_sbo);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad;
}

int main()
{
    EncodeFn encode = get_encode();
    // ---------------- tma2d ----------------
    {
        const int rows = 256, cols = 64;
        std::vector<float> G(rows * cols);
        for (int i = 0; i < rows * cols; ++i) G[i] = (float)i;
        float* dG; float* dO; CK(cudaMalloc(&dG, G.size() * 4)); CK(cudaMalloc(&dO, 128 * 32 * 4));
        CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
        cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dG, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("{\"test\": \"tma2d\", \"encode_error\": %d}\n", (int)r); }
        else {
            CK(cudaFuncSetAttribute(tma2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
            tma2d_kernel<<<1, 128, 16384 + 1024>>>(map, 32, 128, dO, 128 * 32);
            CK(cudaDeviceSynchronize());
            std::vector<float> O(128 * 32);
            CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int k = 0; k < 32; ++k) {
                    const int off = (m * 128 + ((((k >> 2) ^ (m & 7))) << 4) + (k & 3) * 4) / 4;
                    if (O[off] != G[(128 + m) * cols + 32 + k]) { if (bad < 5) printf("  tma2d m=%d k=%d got %g want %g\n", m, k, O[off], G[(128 + m) * cols + 32 + k]); ++bad; }
                }
            printf("{\"test\": \"tma2d\", \"bad\": %d}\n", bad);
        }
    }
    // ---------------- tma3d ----------------
    {
        const int P = 16, E = 1600, na = 25;
        std::vector<float> X(P * E);
        for (int i = 0; i < P * E; ++i) X[i] = (float)i;
        float* dX; float* dO; CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dO, na * 256 * 4));
        CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        CUtensorMap map;
        cuuint64_t dims[3] = {32, (cuuint64_t)P, (cuuint64_t)(E / 32)}; cuuint64_t strides[2] = {(cuuint64_t)E * 4, 128};
        cuuint32_t box[3] = {32, 8, (cuuint32_t)na}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dX, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("{\"test\": \"tma3d\", \"encode_error\": %d}\n", (int)r); }
        else {
            CK(cudaFuncSetAttribute(tma3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            tma3d_kernel<<<1, 128, na * 1024 + 1024>>>(map, 0, 8, 25, dO, na * 256);
            CK(cudaDeviceSynchronize());
            std::vector<float> O(na * 256);
            CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int at = 0; at < na; ++at)
                for (int k = 0; k < 8; ++k)
                    for (int mi = 0; mi < 32; ++mi) {
                        const int off = (at * 1024 + k * 128 + ((((mi >> 3) ^ (k & 3))) << 5) + (mi & 7) * 4) / 4;
                        const float want = X[(8 + k) * E + (25 + at) * 32 + mi];
                        if (O[off] != want) { if (bad < 5) printf("  tma3d at=%d k=%d mi=%d got %g want %g\n", at, k, mi, O[off], want); ++bad; }
                    }
            printf("{\"test\": \"tma3d\", \"bad\": %d}\n", bad);
        }
    }
    // ---------------- mmaK ----------------
    srand(1);
    {
        std::vector<float> A(128 * 32), B(32 * 32);
        for (auto& v : A) v = (float)(rand() % 17 - 8);
        for (auto& v : B) v = (float)(rand() % 17 - 8);
        MmaArgs a; memset(&a, 0, sizeof(a));
        a.a_mode = 0; a.b_mode = 0; a.a_layout = 2; a.b_layout = 2; a.a_major = 0; a.b_major = 0; a.ksteps = 4; a.a_step = 32; a.b_step = 32;
        a.a_lbo = 16; a.a_sbo = 1024; a.b_lbo = 16; a.b_sbo = 1024;
        run_mma("mmaK_sw128", A, 32, B, 32, 32, a, 4);
    }
    // ---------------- isolate: validated SW128 K-major on one side, the candidate layout on the other ----------------
    {
        std::vector<float> A(128 * 32, 0.f), A8(128 * 8), B(32 * 32, 0.f), B8(32 * 8);
        for (int m = 0; m < 128; ++m) for (int k = 0; k < 8; ++k) { A8[m * 8 + k] = (float)(rand() % 17 - 8); A[m * 32 + k] = A8[m * 8 + k]; }
        for (int n = 0; n < 32; ++n) for (int k = 0; k < 8; ++k) { B8[n * 8 + k] = (float)(rand() % 17 - 8); B[n * 32 + k] = B8[n * 8 + k]; }
        {   // T1: A K-major SW128 (validated), B K-major no-swizzle core matrices
            const unsigned cand[][2] = {{128, 256}, {256, 128}};
            for (auto& c : cand) {
                MmaArgs a; memset(&a, 0, sizeof(a));
                a.a_mode = 0; a.b_mode = 1; a.a_layout = 2; a.b_layout = 0; a.ksteps = 1;
                a.a_lbo = 16; a.a_sbo = 1024; a.b_lbo = c[0]; a.b_sbo = c[1];
                // the fill always uses (lbo = 128, sbo = 256) geometry: pass the same to the fill through b_lbo / b_sbo only when equal
                run_mma("T1_aK128_bKnone", A, 32, B8, 8, 8, a, 2);
            }
        }
        {   // T2: A MN-major SW128 atoms, B K-major SW128 (validated), one k-step
            const unsigned cand[][2] = {{1024, 512}, {512, 1024}, {1024, 1024}};
            for (auto& c : cand) {
                MmaArgs a; memset(&a, 0, sizeof(a));
                a.a_mode = 1; a.b_mode = 0; a.a_layout = 1; a.b_layout = 2; a.a_major = 1; a.ksteps = 1;
                a.a_lbo = c[0]; a.a_sbo = c[1]; a.b_lbo = 16; a.b_sbo = 1024;
                run_mma("T2_aMN128_bK128", A8, 8, B, 32, 8, a, 2);
            }
        }
        {   // T3: readout of how the hardware interprets the MN-major A tile: B = identity on k -> D[m][n] = A[m][k = n]
            std::vector<float> Ar(128 * 8), Bi(32 * 32, 0.f);
            for (int m = 0; m < 128; ++m) for (int k = 0; k < 8; ++k) Ar[m * 8 + k] = (float)(m * 8 + k + 1);
            for (int n = 0; n < 8; ++n) Bi[n * 32 + n] = 1.f;
            MmaArgs a; memset(&a, 0, sizeof(a));
            a.a_mode = 1; a.b_mode = 0; a.a_layout = 1; a.b_layout = 2; a.a_major = 1; a.ksteps = 1;
            a.a_lbo = 1024; a.a_sbo = 512; a.b_lbo = 16; a.b_sbo = 1024;
            float *dA, *dB, *dD;
            CK(cudaMalloc(&dA, Ar.size() * 4)); CK(cudaMalloc(&dB, Bi.size() * 4)); CK(cudaMalloc(&dD, 128 * 32 * 4));
            CK(cudaMemcpy(dA, Ar.data(), Ar.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, Bi.data(), Bi.size() * 4, cudaMemcpyHostToDevice));
            a.A = dA; a.B = dB; a.D = dD;
            mma_kernel<<<1, 128, 16384 + 4096 + 1024>>>(a);
            CK(cudaDeviceSynchronize());
            std::vector<float> D(128 * 32);
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            for (int m : {0, 1, 2, 3, 4, 5, 8, 31, 32, 33, 64, 127}) {
                printf("  T3 row %3d:", m);
                for (int n = 0; n < 8; ++n) printf(" %6g", D[m * 32 + n]);
                printf("   (want %d..%d)\n", m * 8 + 1, m * 8 + 8);
            }
        }
    }
    // ---------------- A from TMEM ----------------
    {
        std::vector<float> A8(128 * 8), B(32 * 32, 0.f);
        for (auto& v : A8) v = (float)(rand() % 17 - 8);
        for (int n = 0; n < 32; ++n) for (int k = 0; k < 8; ++k) B[n * 32 + k] = (float)(rand() % 17 - 8);
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, A8.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * 32 * 4));
        CK(cudaMemcpy(dA, A8.data(), A8.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xff, 128 * 32 * 4));
        mma_ts_kernel<<<1, 128, 4096 + 1024>>>(dA, dB, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"test\": \"mma_ts\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 2; }
        std::vector<float> D(128 * 32);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) {
                double ref = 0;
                for (int k = 0; k < 8; ++k) ref += (double)A8[m * 8 + k] * B[n * 32 + k];
                if (fabs(ref - D[m * 32 + n]) > 1e-3) { if (bad < 4) printf("  mma_ts mismatch m=%d n=%d got %g want %g\n", m, n, D[m * 32 + n], ref); ++bad; }
            }
        printf("{\"test\": \"mma_ts_a_from_tmem\", \"bad\": %d}\n", bad);
    }
    // ---------------- trunc ----------------
    {
        std::vector<float> A(128 * 32, 0.f), B(32 * 32, 0.f);
        const float x = 1.0f + 1.0f / 2048 + 1.0f / 4096;      // 1 + 2^-11 + 2^-12: truncation -> 1, rounding -> 1 + 2^-10
        for (int m = 0; m < 128; ++m) A[m * 32] = x;
        for (int n = 0; n < 32; ++n) B[n * 32] = 1.0f;
        B[1 * 32] = x;                                          // column 1: both operands carry low bits
        MmaArgs a; memset(&a, 0, sizeof(a));
        a.a_mode = 0; a.b_mode = 0; a.a_layout = 2; a.b_layout = 2; a.ksteps = 4; a.a_step = 32; a.b_step = 32;
        a.a_lbo = 16; a.a_sbo = 1024; a.b_lbo = 16; a.b_sbo = 1024;
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * 32 * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        a.A = dA; a.B = dB; a.D = dD;
        mma_kernel<<<1, 128, 16384 + 4096 + 1024>>>(a);
        CK(cudaDeviceSynchronize());
        float D[64];
        CK(cudaMemcpy(D, dD, sizeof(D), cudaMemcpyDeviceToHost));
        printf("{\"test\": \"trunc\", \"x\": %.10f, \"d00\": %.10f, \"d01\": %.10f, \"trunc_expect\": 1.0, \"round_expect\": %.10f}\n", x, D[0], D[1],
               1.0f + 1.0f / 1024);
    }
    return 0;
}
