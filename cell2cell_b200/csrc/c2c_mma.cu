// Launchers and tensor maps of the tensor-core streaming passes (c2c_mma.cuh).
#include "c2c_mma.cuh"

template <int N>
static cudaError_t launch_plane_n(const CUtensorMap& map, const PlaneMmaArgs& a, int grid, size_t smem, cudaStream_t st, int pdl)
{
    cudaError_t e = cudaFuncSetAttribute(plane_mma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return c2c_launch_ex(plane_mma_kernel<N>, dim3(grid), dim3(C2C_MMA_THREADS), smem, st, pdl, map, a);
}

template <int N, bool STACK, bool ALLTS>
static cudaError_t launch_fiber_n(const CUtensorMap& map, const FiberMmaArgs& a, int grid, size_t smem, cudaStream_t st, int pdl)
{
    cudaError_t e = cudaFuncSetAttribute(fiber_mma_kernel<N, STACK, ALLTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return c2c_launch_ex(fiber_mma_kernel<N, STACK, ALLTS>, dim3(grid), dim3(C2C_MMA_THREADS), smem, st, pdl, map, a);
}

cudaError_t c2c_launch_plane_mma(int n, const CUtensorMap& map, const PlaneMmaArgs& a, int grid, size_t smem, cudaStream_t st, int pdl)
{
    switch (n) {
        case 16: return launch_plane_n<16>(map, a, grid, smem, st, pdl);
        case 24: return launch_plane_n<24>(map, a, grid, smem, st, pdl);
        case 32: return launch_plane_n<32>(map, a, grid, smem, st, pdl);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t c2c_launch_fiber_mma(int n, int stack, int allts, const CUtensorMap& map, const FiberMmaArgs& a, int grid, size_t smem, cudaStream_t st, int pdl)
{
    switch (n) {
        case 16: return allts ? (stack ? launch_fiber_n<16, true, true>(map, a, grid, smem, st, pdl) : launch_fiber_n<16, false, true>(map, a, grid, smem, st, pdl))
                              : (stack ? launch_fiber_n<16, true, false>(map, a, grid, smem, st, pdl) : launch_fiber_n<16, false, false>(map, a, grid, smem, st, pdl));
        case 32: return stack ? launch_fiber_n<32, true, false>(map, a, grid, smem, st, pdl) : launch_fiber_n<32, false, false>(map, a, grid, smem, st, pdl);
        default: return cudaErrorInvalidValue;
    }
}

size_t c2c_plane_mma_smem(int n, int I2, int I3, int nx)
{
    const size_t xb = 128 * 128, bb = 2 * (size_t)n * 128;
    const int nr = 256 / (n == 24 ? 64 : 2 * n);
    return 1024 + (size_t)nx * xb + (size_t)C2C_MMA_RING2 * bb + (size_t)(I2 + I3) * n * sizeof(float) +
           (size_t)(2 * nx + 2 * C2C_MMA_RING2 + 2 * nr) * 8 + 16;
}

size_t c2c_fiber_mma_smem(int n, int na, int ldr, int nx)
{
    const size_t xb = (size_t)na * 1024, bb = 2 * (size_t)n * 32;
    return 1024 + (size_t)nx * xb + (size_t)C2C_MMA_RING2 * bb + (size_t)na * 32 * ldr * sizeof(float) +
           (size_t)(2 * nx + 2 * C2C_MMA_RING2 + 2 * C2C_FM_MAXTILE) * 8 + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the driver entry point is resolved through the runtime, so the library does not link against libcuda directly
static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int c2c_make_plane_map(CUtensorMap* map, const float* X, long long P, long long E)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)E, (cuuint64_t)P};
    cuuint64_t strides[1] = {(cuuint64_t)E * sizeof(float)};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int c2c_make_fiber_map(CUtensorMap* map, const float* X, long long P, long long E, int na)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -1;
    cuuint64_t dims[3] = {32, (cuuint64_t)P, (cuuint64_t)((E + 31) / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)E * sizeof(float), 128};
    cuuint32_t box[3] = {32, 8, (cuuint32_t)na};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(X), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int c2c_make_fiber_map2d(CUtensorMap* map, const float* X, long long P, long long E)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)E, (cuuint64_t)P};
    cuuint64_t strides[1] = {(cuuint64_t)E * sizeof(float)};
    cuuint32_t box[2] = {32, 8};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// 1 when the driver exposes cuTensorMapEncodeTiled (the tensor-pipe passes need TMA tensor maps)
int c2c_mma_available() { return encode_fn() != nullptr ? 1 : 0; }
