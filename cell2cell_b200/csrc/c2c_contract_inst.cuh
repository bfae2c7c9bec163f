// Explicit instantiation of the streaming kernels for one (VEC, MASKED) pair; included by c2c_contract_v*_m*.cu with
// C2C_INST_VEC / C2C_INST_MASKED defined, so the 16 rank-pair variants of each pair compile as separate jobs.
#include "c2c_contract.cuh"

#define C2C_CAT_(a, b, c, d) a##b##c##d
#define C2C_CAT(a, b, c, d) C2C_CAT_(a, b, c, d)

#define C2C_PLANE_CASE(N)                                                                                          \
    case N:                                                                                                        \
        if (smem > 48 * 1024) {                                                                                    \
            e = cudaFuncSetAttribute(plane_contract_kernel<C2C_INST_VEC, N, C2C_INST_MASKED != 0, false>,          \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
            if (e != cudaSuccess) return e;                                                                        \
        }                                                                                                          \
        plane_contract_kernel<C2C_INST_VEC, N, C2C_INST_MASKED != 0, false><<<grid, C2C_PC_THREADS, smem, st>>>(a); \
        break;

cudaError_t C2C_CAT(c2c_launch_plane_v, C2C_INST_VEC, _m, C2C_INST_MASKED)(int np2, const ContractArgs& a, int grid,
                                                                           size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaSuccess;
    switch (np2) {
        C2C_PLANE_CASE(1) C2C_PLANE_CASE(2) C2C_PLANE_CASE(3) C2C_PLANE_CASE(4) C2C_PLANE_CASE(5) C2C_PLANE_CASE(6)
        C2C_PLANE_CASE(7) C2C_PLANE_CASE(8) C2C_PLANE_CASE(9) C2C_PLANE_CASE(10) C2C_PLANE_CASE(11) C2C_PLANE_CASE(12)
        C2C_PLANE_CASE(13) C2C_PLANE_CASE(14) C2C_PLANE_CASE(15) C2C_PLANE_CASE(16)
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

#define C2C_FIBER_CASE(N)                                                                                          \
    case N:                                                                                                        \
        fiber_contract_kernel<C2C_INST_VEC, N, C2C_INST_MASKED != 0><<<grid, block, 0, st>>>(                      \
            a, nvec, nvec_total, nslots, planes_per_cta);                                                          \
        break;

cudaError_t C2C_CAT(c2c_launch_fiber_v, C2C_INST_VEC, _m, C2C_INST_MASKED)(int np2, const ContractArgs& a, dim3 grid,
                                                                           int block, int nvec, int nvec_total,
                                                                           int nslots, long long planes_per_cta,
                                                                           cudaStream_t st)
{
    switch (np2) {
        C2C_FIBER_CASE(1) C2C_FIBER_CASE(2) C2C_FIBER_CASE(3) C2C_FIBER_CASE(4) C2C_FIBER_CASE(5) C2C_FIBER_CASE(6)
        C2C_FIBER_CASE(7) C2C_FIBER_CASE(8) C2C_FIBER_CASE(9) C2C_FIBER_CASE(10) C2C_FIBER_CASE(11) C2C_FIBER_CASE(12)
        C2C_FIBER_CASE(13) C2C_FIBER_CASE(14) C2C_FIBER_CASE(15) C2C_FIBER_CASE(16)
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

#if C2C_INST_MASKED == 1
// the final-loss sums kernel shares the masked plane mapping; one instance set per VEC
#define C2C_SUMS_CASE(N)                                                                                           \
    case N:                                                                                                        \
        if (smem > 48 * 1024) {                                                                                    \
            e = cudaFuncSetAttribute(plane_contract_kernel<C2C_INST_VEC, N, true, true>,                           \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
            if (e != cudaSuccess) return e;                                                                        \
        }                                                                                                          \
        plane_contract_kernel<C2C_INST_VEC, N, true, true><<<grid, C2C_PC_THREADS, smem, st>>>(a);                 \
        break;

cudaError_t C2C_CAT(c2c_launch_sums_v, C2C_INST_VEC, , )(int np2, const ContractArgs& a, int grid, size_t smem,
                                                         cudaStream_t st)
{
    cudaError_t e = cudaSuccess;
    switch (np2) {
        C2C_SUMS_CASE(1) C2C_SUMS_CASE(2) C2C_SUMS_CASE(3) C2C_SUMS_CASE(4) C2C_SUMS_CASE(5) C2C_SUMS_CASE(6)
        C2C_SUMS_CASE(7) C2C_SUMS_CASE(8) C2C_SUMS_CASE(9) C2C_SUMS_CASE(10) C2C_SUMS_CASE(11) C2C_SUMS_CASE(12)
        C2C_SUMS_CASE(13) C2C_SUMS_CASE(14) C2C_SUMS_CASE(15) C2C_SUMS_CASE(16)
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
#endif
