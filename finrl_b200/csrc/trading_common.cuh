// Shared by the two StockTradingEnv kernels (trading.cu: thread per env; trading_small.cu: 8 lanes per env).
#pragma once

#include "common.cuh"

namespace frl {

constexpr int kMaxAbsAction = (1 << 26) - 1;  // |int(action*hmax)| is clamped to this (key packing)

template <typename ActT>
__device__ __forceinline__ int action_to_shares(ActT a, double hmax);
template <>
__device__ __forceinline__ int action_to_shares<float>(float a, double hmax)
{
    // float32 array * python int -> float32 product, then astype(int) truncates toward zero.
    // cvt.rzi.s32.f32 saturates, so |v| >= 2^31 lands on the clamp like the int64 cast would.
    const int t = __float2int_rz(fmul(a, (float)hmax));
    return max(-kMaxAbsAction, min(kMaxAbsAction, t));
}
template <>
__device__ __forceinline__ int action_to_shares<double>(double a, double hmax)
{
    const int t = __double2int_rz(dmul(a, hmax));
    return max(-kMaxAbsAction, min(kMaxAbsAction, t));
}

__device__ __forceinline__ int state_day(int sday) { return sday < 0 ? -sday - 1 : sday; }

// host: the low-latency kernel for small batches (defined in trading_small.cu)
void launch_trading_small(const frl_trading_params &p, const void *actions, int actions_f64, long long act_step_stride,
                          long long act_env_stride, int n_steps, double *rewards, uint8_t *flags, float *obs, int obs_mode,
                          int auto_reset, double *stats, cudaStream_t st);

// host: the thread-per-env kernel for 33..128 stocks at large batches (defined in trading_wide.cu)
int32_t launch_trading_wide(const frl_trading_params &p, const void *actions, int actions_f64, long long act_step_stride,
                            long long act_env_stride, int n_steps, double *rewards, uint8_t *flags, float *obs, int obs_mode,
                            int auto_reset, double *stats, cudaStream_t st);

}  // namespace frl
