// Shared device helpers of the R-NaD hot-path kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rnad_b200.h"

namespace rnad {

// ---------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();

#define RNAD_REQUIRE(cond, ...)                \
    do {                                       \
        if (!(cond)) {                         \
            rnad::set_error(__VA_ARGS__);      \
            return RNAD_EINVAL;                \
        }                                      \
    } while (0)

#define RNAD_CHECK_LAUNCH(what)                                      \
    do {                                                             \
        int _rc = rnad::check_cuda(cudaGetLastError(), what);        \
        if (_rc) return _rc;                                         \
    } while (0)

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ constexpr int ev_stride_of(int A) { return round_up(A * A + 1, 4); }
__host__ __device__ constexpr int tr_stride_of(int C) { return round_up(3 * C, 4); }

// ------------------------------------------------------------- device side

// Philox4x32-10 (Salmon et al., SC'11).  Same stream definition as
// oracle/rnad_oracle.py::philox_uniforms: counter = (game_lo, game_hi, t, 0),
// key = (seed_lo, seed_hi); word 0 -> action uniform, word 1 -> chance uniform.
struct Uniforms2 {
    float action, chance;
};

__device__ __forceinline__ Uniforms2 philox_uniforms(uint64_t seed, uint32_t t, uint64_t game) {
    uint32_t c0 = (uint32_t)game, c1 = (uint32_t)(game >> 32), c2 = t, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Uniforms2 u;
    u.action = (float)(c0 >> 8) * 5.9604644775390625e-8f;   // 2^-24
    u.chance = (float)(c1 >> 8) * 5.9604644775390625e-8f;
    return u;
}

// Inverse-CDF categorical draw (oracle sample_icdf): first k with p[k] > 0 and
// u < p[0] + ... + p[k] (fp32, left to right); if rounding leaves u >= total,
// the last k with p[k] > 0.
template <int N>
__device__ __forceinline__ int sample_icdf(const float (&p)[N], int n, float u) {
    float acc = 0.f;
    int choice = 0;
    bool done = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (k < n) {
            acc = __fadd_rn(acc, p[k]);
            const bool positive = p[k] > 0.f;
            if (!done && positive) choice = k;
            done = done || (positive && (u < acc));
        }
    }
    return choice;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming (write-once) stores: keep the trajectory out of L1
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(int64_t* p, int64_t v) { __stcs((long long*)p, (long long)v); }

}  // namespace rnad
