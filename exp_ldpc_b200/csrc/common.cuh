// common.cuh — shared device helpers: error plumbing, Philox4x32-10, bit utilities.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace qd {

// ---------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry point returns a code and leaves a message here
// ---------------------------------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define QD_CUDA(expr)                                                                            \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ::qd::fail(-3 /*QD_ECUDA*/, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10; the definition shared with oracle/qldpc_oracle.c (orc_philox4x32_10)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ __forceinline__ int words_for(int bits) { return (bits + 31) >> 5; }

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace qd
