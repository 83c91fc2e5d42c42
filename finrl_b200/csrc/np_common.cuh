// Shared pieces of the numpy-env kernels (nptrading.cu: D <= 32, state in registers; np_wide.cu: D <= 128,
// state streamed): the NEP-50 scalar with its kind, the action cast and the amount slot of the observation.
#pragma once

#include "common.cuh"

namespace frl {

struct NV {  // a numpy / Python scalar: value + NEP-50 kind
    double v;
    int k;
};
__device__ __forceinline__ NV nv(double v, int k) { return NV{v, k}; }
// NEP-50 binary op: the result kind is max(kind_x, kind_y) with PY(0) < F32(1) < F64(2) — two Python floats
// stay a Python float, a weak Python float adopts float32, float64 wins — and the arithmetic is float32 only
// when the result is.  Written select-style (both candidates are cheap) so that the unrolled trade loops stay
// branch-light.
#ifndef FRL_NV_SELECT
#define FRL_NV_SELECT 1
#endif
#if FRL_NV_SELECT
__device__ __forceinline__ NV nv_add(NV x, NV y)
{
    const int k = max(x.k, y.k);
    return nv(k == FRL_KIND_F32 ? (double)fadd((float)x.v, (float)y.v) : dadd(x.v, y.v), k);
}
__device__ __forceinline__ NV nv_sub(NV x, NV y)
{
    const int k = max(x.k, y.k);
    return nv(k == FRL_KIND_F32 ? (double)fsub((float)x.v, (float)y.v) : dsub(x.v, y.v), k);
}
__device__ __forceinline__ NV nv_mul(NV x, NV y)
{
    const int k = max(x.k, y.k);
    return nv(k == FRL_KIND_F32 ? (double)fmul((float)x.v, (float)y.v) : dmul(x.v, y.v), k);
}
#else
__device__ __forceinline__ NV nv_add(NV x, NV y)
{
    if (x.k == FRL_KIND_PY && y.k == FRL_KIND_PY) return nv(dadd(x.v, y.v), FRL_KIND_PY);
    if (x.k == FRL_KIND_F64 || y.k == FRL_KIND_F64) return nv(dadd(x.v, y.v), FRL_KIND_F64);
    return nv((double)fadd((float)x.v, (float)y.v), FRL_KIND_F32);  // f32 (a weak Python float adopts it)
}
__device__ __forceinline__ NV nv_sub(NV x, NV y)
{
    if (x.k == FRL_KIND_PY && y.k == FRL_KIND_PY) return nv(dsub(x.v, y.v), FRL_KIND_PY);
    if (x.k == FRL_KIND_F64 || y.k == FRL_KIND_F64) return nv(dsub(x.v, y.v), FRL_KIND_F64);
    return nv((double)fsub((float)x.v, (float)y.v), FRL_KIND_F32);
}
__device__ __forceinline__ NV nv_mul(NV x, NV y)
{
    if (x.k == FRL_KIND_PY && y.k == FRL_KIND_PY) return nv(dmul(x.v, y.v), FRL_KIND_PY);
    if (x.k == FRL_KIND_F64 || y.k == FRL_KIND_F64) return nv(dmul(x.v, y.v), FRL_KIND_F64);
    return nv((double)fmul((float)x.v, (float)y.v), FRL_KIND_F32);
}
#endif

// The same three operations when one operand is known to be np.float64 (A64, the kernels' steady-state
// instantiation): the result is float64 whatever the other kind is, so the kind arithmetic and the float32
// candidate disappear at compile time.
template <bool A64>
__device__ __forceinline__ NV nv_add_t(NV x, NV y)
{
    if (A64) return nv(dadd(x.v, y.v), FRL_KIND_F64);
    return nv_add(x, y);
}
template <bool A64>
__device__ __forceinline__ NV nv_sub_t(NV x, NV y)
{
    if (A64) return nv(dsub(x.v, y.v), FRL_KIND_F64);
    return nv_sub(x, y);
}
template <bool A64>
__device__ __forceinline__ NV nv_mul_t(NV x, NV y)
{
    if (A64) return nv(dmul(x.v, y.v), FRL_KIND_F64);
    return nv_mul(x, y);
}

// Exact widening conversions WITHOUT the conversion unit (XU, 16 lanes/clk/SM — the busiest pipe of these
// kernels, which convert between float32, float64 and int64 on every trade): float -> double by re-biasing
// the exponent with integer ops (zero handled inline; denormals / inf / nan fall back to the converter), and a
// non-negative int -> double with the 2**52 magic-number trick (one DADD).  Both are exact, like the casts.
#ifndef FRL_NP_ALU_CONVERT
#define FRL_NP_ALU_CONVERT 0  // A/B on B200: 1 grows the unrolled trade loops past the instruction cache (0.44 -> 0.76 ms)
#endif
__device__ __forceinline__ double np_f2d(float f)
{
#if FRL_NP_ALU_CONVERT
    const unsigned b = __float_as_uint(f);
    const unsigned e = b & 0x7f800000u;
    if (e - 0x00800000u < 0x7f000000u)  // normal number
        return __hiloint2double((int)((b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u)), (int)(b << 29));
    if ((b & 0x7fffffffu) == 0u) return __hiloint2double((int)b, 0);  // +-0
#endif
    return (double)f;
}
__device__ __forceinline__ double np_u2d(int x)  // x >= 0
{
#if FRL_NP_ALU_CONVERT
    return __hiloint2double(0x43300000, x) - 4503599627370496.0;
#else
    return (double)x;
#endif
}

// The cash-limited buy needs numpy's floor division, a ~40-instruction sequence that the unrolled trade loops
// would inline once per stock slot; it is rare at run time, so it lives out of line and the hot path of the
// loops stays compact (instruction-cache footprint is what limits these kernels' issue rate).
#ifndef FRL_NP_OUTLINE_DIV
#define FRL_NP_OUTLINE_DIV 1
#endif
#if FRL_NP_OUTLINE_DIV
static __device__ __noinline__ double np_floor_div(double a, double b, int f64)
#else
__device__ __forceinline__ double np_floor_div(double a, double b, int f64)
#endif
{
    return f64 ? floor_div_f64(a, b) : (double)floor_div_f32((float)a, (float)b);
}

template <typename ActT>
__device__ __forceinline__ int np_action_to_shares(ActT a, double max_stock);
template <>
__device__ __forceinline__ int np_action_to_shares<float>(float a, double max_stock)
{
    return __float2int_rz(fmul(a, (float)max_stock));  // f32 array * Python float -> f32; astype(int)
}
template <>
__device__ __forceinline__ int np_action_to_shares<double>(double a, double max_stock)
{
    return __double2int_rz(dmul(a, max_stock));
}

// np.array(self.amount * 2**-12, dtype=np.float32): the power-of-two scale commutes with the cast
// (StockEnvNAS100 shows max(amount, 1e4): Python's max returns the float floor only when it is larger)
__device__ __forceinline__ float np_amount_obs(NV amount, double floor_)
{
    const double a = floor_ > amount.v ? floor_ : amount.v;
    return fmul((float)a, 0.000244140625f);
}

// np_wide.cu
void launch_np_wide(const frl_np_params &p, const void *actions, int actions_f64, long long sstride, long long estride, int n_steps,
                    double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st);
void launch_np_observe_wide(const frl_np_params &p, float *obs, cudaStream_t st);
extern int g_npw_bulk;  // -1: from FRL_NPW_BULK (default on); frl_set_option("np_wide_bulk", v)

}  // namespace frl
