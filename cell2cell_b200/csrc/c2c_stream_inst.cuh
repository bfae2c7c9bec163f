// Explicit instantiation of the shared-memory-staged streaming kernels for one (VEC, NS); included by c2c_stream_v*_s*.cu.
#include "c2c_stream.cuh"

#define C2C_SCAT_(a, b, c, d) a##b##c##d
#define C2C_SCAT(a, b, c, d) C2C_SCAT_(a, b, c, d)

#define C2C_PS_CASE(N)                                                                                              \
    case N:                                                                                                         \
        e = cudaFuncSetAttribute(plane_stream_kernel<C2C_INST_VEC, C2C_INST_NS, N>,                                 \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                           \
        if (e != cudaSuccess) return e;                                                                             \
        e = c2c_launch_ex(plane_stream_kernel<C2C_INST_VEC, C2C_INST_NS, N>, dim3(grid), dim3(block), smem, st, pdl, a); \
        if (e != cudaSuccess) return e;                                                                             \
        break;

#define C2C_FS_CASE(N)                                                                                              \
    case N:                                                                                                         \
        e = cudaFuncSetAttribute(fiber_stream_kernel<C2C_INST_VEC, C2C_INST_NS, N>,                                 \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                           \
        if (e != cudaSuccess) return e;                                                                             \
        e = c2c_launch_ex(fiber_stream_kernel<C2C_INST_VEC, C2C_INST_NS, N>, dim3(grid), dim3(block), smem, st, pdl, a, nthr, nslots, merge, w_all); \
        if (e != cudaSuccess) return e;                                                                             \
        break;

#if C2C_INST_NS == 2
#define C2C_ALL_CASES(M) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10)
#else
#define C2C_ALL_CASES(M) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16)
#endif

cudaError_t C2C_SCAT(c2c_launch_plane_stream_v, C2C_INST_VEC, _s, C2C_INST_NS)(int np2, const StreamArgs& a, int grid, int block,
                                                                               size_t smem, cudaStream_t st, int pdl)
{
    cudaError_t e = cudaSuccess;
    switch (np2) {
        C2C_ALL_CASES(C2C_PS_CASE)
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t C2C_SCAT(c2c_launch_fiber_stream_v, C2C_INST_VEC, _s, C2C_INST_NS)(int np2, const StreamArgs& a, int nthr, int nslots,
                                                                               int merge, int w_all, int grid, int block, size_t smem, cudaStream_t st, int pdl)
{
    cudaError_t e = cudaSuccess;
    switch (np2) {
        C2C_ALL_CASES(C2C_FS_CASE)
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
