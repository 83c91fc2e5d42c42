// finrl_b200 — shared device/host helpers for the sm_100a env-step kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "finrl_b200.h"

namespace frl {

// ---- error plumbing (host) -------------------------------------------------------------------
void set_error(const char *fmt, ...);
int32_t check_launch(const char *what);

#define FRL_REQUIRE(cond, ...)                                                                     \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ::frl::set_error(__VA_ARGS__);                                                         \
            return FRL_E_INVALID;                                                                  \
        }                                                                                          \
    } while (0)

constexpr int kWarp = 32;

// ---- exact IEEE helpers (device) ---------------------------------------------------------------
// The reference's arithmetic is one IEEE rounding per Python/numpy operation.  Every product and
// sum on the bit-exact list goes through these so that ptxas can never contract a*b+c into an FMA
// (SURVEY.md H2).
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// numpy's float floor-division (npy_divmod) returns, for every finite a and b > 0 with a moderate
// quotient, the exact mathematical floor(a/b) of the two floating-point values (its fmod is exact
// and its final snap repairs the one rounding of (a-mod)/b).  We get the same integer without
// fmod's long loop: round-to-nearest quotient, floor, then one exact FMA remainder to repair the
// (at most one-off) error.  b == 0 gives +-inf like numpy (a != 0).
__device__ __forceinline__ double floor_div_f64(double a, double b)
{
    double q = floor(__ddiv_rn(a, b));
    double r = __fma_rn(-q, b, a);  // exact: |r| < 2b and r is a multiple of ulp(b)
    if (r < 0.0)
        q -= 1.0;
    else if (r >= b)
        q += 1.0;
    return q;
}

__device__ __forceinline__ float floor_div_f32(float a, float b)
{
    float q = floorf(__fdiv_rn(a, b));
    float r = __fmaf_rn(-q, b, a);
    if (r < 0.0f)
        q -= 1.0f;
    else if (r >= b)
        q += 1.0f;
    return q;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace frl
