// Sibling env — CryptoEnv (reference: finrl/meta/env_cryptocurrency_trading/env_multiple_crypto.py).
//
// The numpy env's structure (one thread per env, ascending-index sells then buys, registers for the
// per-asset state, warp-cooperative observation rows) without turbulence / cool-down and with FRACTIONAL
// float32 positions against float64 prices and cash: every product with the price promotes to float64,
// `min()` keeps the dtype of the side that wins, the position array stays float32.
#include "common.cuh"

namespace frl {
namespace {

constexpr int kPitch = 33;

template <int SLOTS, typename ActT>
struct alignas(16) CryptoWarpSmem {
    union {
        ActT act[32 * SLOTS];      // staged actions, flat [32 envs][D]
        float sc[SLOTS * kPitch];  // stocks[j][lane] for the observation writer
    };
    float cashf[32];
    int time[32];
};

// cash + (stocks * price).sum(): float32 x float64 -> float64 products, numpy pairwise sum
template <int SLOTS>
__device__ __forceinline__ double crypto_total(double cash, const float (&stv)[SLOTS], const double *__restrict__ prow, int D)
{
    double x[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) x[j] = (j < D) ? dmul((double)stv[j], __ldg(prow + j)) : 0.0;
    double res;
    if (D < 8) {
        res = 0.0;
#pragma unroll
        for (int j = 0; j < 8 && j < SLOTS; ++j)
            if (j < D) res = dadd(res, x[j]);
    } else {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = x[j];
        const int nb = D >> 3;
#pragma unroll
        for (int b = 1; b < SLOTS / 8; ++b) {
            if (b < nb) {
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = dadd(r[j], x[8 * b + j]);
            }
        }
        res = dadd(dadd(dadd(r[0], r[1]), dadd(r[2], r[3])), dadd(dadd(r[4], r[5]), dadd(r[6], r[7])));
#pragma unroll
        for (int j = 8; j < SLOTS; ++j)
            if (j >= 8 * nb && j < D) res = dadd(res, x[j]);
    }
    return dadd(cash, res);
}

template <typename SM>
__device__ __forceinline__ void crypto_write_obs_tile(const frl_crypto_params &p, SM &sm, float *__restrict__ obs,
                                                      long long env0, int nvalid, int lane)
{
    // row = [cash * 2^-18, stocks * 2^-3 x D, tech rows * 2^-15]  (float32)
    const int O = p.obs_dim, D = p.stock_dim;
    for (int r = 0; r < nvalid; ++r) {
        const float *trow = p.obs_tmpl + (size_t)sm.time[r] * O;
        float *orow = obs + (size_t)(env0 + r) * O;
        for (int pos = lane; pos < O; pos += 32) {
            float v;
            if (pos == 0)
                v = sm.cashf[r];
            else if (pos <= D)
                v = fmul(sm.sc[(pos - 1) * kPitch + r], 0.125f);
            else
                v = __ldg(trow + pos);
            orow[pos] = v;
        }
    }
}

template <int SLOTS, typename ActT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 512 / (WARPS * 32))
crypto_rollout_kernel(const frl_crypto_params p, const ActT *__restrict__ actions, long long act_step_stride,
                      long long act_env_stride, int n_steps, double *__restrict__ rewards, uint8_t *__restrict__ flags_out,
                      float *__restrict__ obs, int obs_mode, int auto_reset, double *__restrict__ stats)
{
    stats_exchange_previous(stats);
    using SM = CryptoWarpSmem<SLOTS, ActT>;
    __shared__ SM smem[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SM &sm = smem[warp];
    const int N = p.n_envs, D = p.stock_dim, T = p.n_days, ld = p.env_stride;
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;
    const int max_step = T - p.lookback - 1;

    double cash = p.cash[n], total = p.total[n], gret = p.gamma_return[n];
    int time = p.time[n];
    float stv[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) stv[j] = (j < D) ? __ldcs(p.stocks + n + j * ld) : 0.0f;
    const double one_minus_sc = dsub(1.0, p.sell_cost_pct), one_plus_bc = dadd(1.0, p.buy_cost_pct);
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        stage_actions_flat<SLOTS, ActT>(sm.act, abase, env0, act_env_stride, D, nvalid, lane);
        __syncwarp();

        int flags = 0;
        double reward = 0.0;
        if (time >= T - 1) {
            flags = FRL_FLAG_DONE;  // past the data (the reference would raise IndexError): inert
        } else {
            time += 1;
            const double *prow = p.price + (size_t)time * 32;
            const ActT *arow = sm.act + lane * D;
            // actions[i] * norm_vector_i in the action dtype (:62-64); sells in ascending index (:66-70)
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
                if (j < D) {
                    const double nj = __ldg(p.act_norm + j);
                    const double a = sizeof(ActT) == 4 ? (double)fmul((float)arow[j], (float)nj) : dmul((double)arow[j], nj);
                    const double pj = __ldg(prow + j);
                    if (a < 0.0 && pj > 0.0) {
                        double nsh;
                        if (-a < (double)stv[j]) {  // min(stocks, -action) -> the action's dtype
                            nsh = -a;
                            stv[j] = sizeof(ActT) == 4 ? fsub(stv[j], (float)nsh) : (float)dsub((double)stv[j], nsh);
                        } else {
                            nsh = (double)stv[j];
                            stv[j] = fsub(stv[j], stv[j]);
                        }
                        cash = dadd(cash, dmul(dmul(pj, nsh), one_minus_sc));
                    }
                }
            }
            // buys in ascending index, limited by cash // price (no cost term, like the numpy env) (:72-76)
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
                if (j < D) {
                    const double nj = __ldg(p.act_norm + j);
                    const double a = sizeof(ActT) == 4 ? (double)fmul((float)arow[j], (float)nj) : dmul((double)arow[j], nj);
                    const double pj = __ldg(prow + j);
                    if (a > 0.0 && pj > 0.0) {
                        const double avail = floor_div_f64(cash, pj);
                        double nsh;
                        if (a < avail) {  // min(avail, action) -> the action
                            nsh = a;
                            stv[j] = sizeof(ActT) == 4 ? fadd(stv[j], (float)nsh) : (float)dadd((double)stv[j], nsh);
                        } else {
                            nsh = avail;
                            stv[j] = (float)dadd((double)stv[j], nsh);
                        }
                        cash = dsub(cash, dmul(dmul(pj, nsh), one_plus_bc));
                    }
                }
            }
            const double next_total = crypto_total<SLOTS>(cash, stv, prow, D);
            reward = dmul(dsub(next_total, total), 1.52587890625e-05);  // 2 ** -16
            total = next_total;
            gret = dadd(dmul(gret, p.gamma), reward);
            if (time == max_step) {
                flags = FRL_FLAG_DONE;
                reward = gret;
                if (valid) {
                    p.episode_return[n] = __ddiv_rn(total, p.initial_capital);
                    st_done += 1.0;
                    st_epi += total;
                }
            }
        }
        if (valid) {
            if (rewards) rewards[(size_t)k * N + n] = reward;
            if (flags_out) flags_out[(size_t)k * N + n] = (uint8_t)flags;
            st_r += reward;
            st_r2 += reward * reward;
        }
        if ((flags & FRL_FLAG_DONE) && auto_reset) {  // reset (:47-57)
            time = p.lookback - 1;
            cash = p.initial_capital;
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) stv[j] = 0.0f;
            total = crypto_total<SLOTS>(cash, stv, p.price + (size_t)time * 32, D);
        }
        if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < SLOTS; ++j)
                if (j < D) sm.sc[j * kPitch + lane] = stv[j];
            sm.cashf[lane] = (float)dmul(cash, 3.814697265625e-06);  // cash * 2 ** -18
            sm.time[lane] = time;
            __syncwarp();
            float *o = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * p.obs_dim : (size_t)0);
            crypto_write_obs_tile(p, sm, o, env0, nvalid, lane);
        }
    }
    if (valid) {
        p.cash[n] = cash;
        p.total[n] = total;
        p.gamma_return[n] = gret;
        p.time[n] = time;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
            if (j < D) p.stocks[n + j * ld] = stv[j];
    }
    if (stats) {
        double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, valid ? total : 0.0, 0.0, valid ? (double)n_steps : 0.0, 0.0};
        reduce_stats8(v, lane, stats);
    }
}

__global__ void crypto_reset_kernel(const frl_crypto_params p, const uint8_t *__restrict__ mask)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.n_envs) return;
    if (mask && !mask[n]) return;
    for (int j = 0; j < p.stock_dim; ++j) p.stocks[n + (size_t)j * p.env_stride] = 0.0f;
    p.time[n] = p.lookback - 1;
    p.cash[n] = p.initial_capital;
    p.total[n] = dadd(p.initial_capital, 0.0);  // cash + (0 * price).sum()
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) crypto_observe_kernel(const frl_crypto_params p, float *__restrict__ obs)
{
    using SM = CryptoWarpSmem<32, float>;
    __shared__ SM smem[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SM &sm = smem[warp];
    const int N = p.n_envs, D = p.stock_dim;
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const long long n = lane < nvalid ? env0 + lane : (long long)N - 1;
    for (int j = 0; j < D; ++j) sm.sc[j * kPitch + lane] = p.stocks[n + (size_t)j * p.env_stride];
    sm.cashf[lane] = (float)dmul(p.cash[n], 3.814697265625e-06);
    sm.time[lane] = p.time[n];
    __syncwarp();
    crypto_write_obs_tile(p, sm, obs, env0, nvalid, lane);
}

int32_t crypto_validate(const frl_crypto_params *p)
{
    FRL_REQUIRE(p != nullptr, "crypto: params is NULL");
    FRL_REQUIRE(p->n_envs >= 1, "crypto: n_envs must be >= 1 (got %d)", p->n_envs);
    FRL_REQUIRE(p->stock_dim >= 1 && p->stock_dim <= 32, "crypto: stock_dim must be in 1..32 (got %d)", p->stock_dim);
    FRL_REQUIRE(p->lookback >= 1 && p->n_days >= p->lookback + 1, "crypto: bad lookback/n_days (%d, %d)", p->lookback, p->n_days);
    FRL_REQUIRE(p->obs_dim == 1 + p->stock_dim + p->tech_dim * p->lookback, "crypto: obs_dim %d != 1 + D + tech_dim*lookback = %d",
                p->obs_dim, 1 + p->stock_dim + p->tech_dim * p->lookback);
    FRL_REQUIRE(p->env_stride >= p->n_envs && (long long)p->env_stride * 32 < (1LL << 31), "crypto: bad env_stride %d", p->env_stride);
    FRL_REQUIRE(p->price && p->act_norm && p->obs_tmpl, "crypto: table pointer is NULL");
    FRL_REQUIRE(p->cash && p->stocks && p->time && p->total && p->gamma_return && p->episode_return, "crypto: state pointer is NULL");
    return FRL_OK;
}

template <int SLOTS, typename ActT, int WARPS>
void crypto_launch(const frl_crypto_params &p, const void *actions, long long sstride, long long estride, int n_steps,
                   double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    crypto_rollout_kernel<SLOTS, ActT, WARPS><<<(unsigned)((tiles + WARPS - 1) / WARPS), WARPS * 32, 0, st>>>(
        p, (const ActT *)actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats);
}

}  // namespace
}  // namespace frl

using namespace frl;

extern "C" int32_t frl_crypto_observe(const frl_crypto_params *p, float *obs, void *stream)
{
    if (int32_t rc = crypto_validate(p)) return rc;
    FRL_REQUIRE(obs != nullptr, "crypto_observe: obs is NULL");
    constexpr int W = 4;
    const long long tiles = ((long long)p->n_envs + 31) / 32;
    crypto_observe_kernel<W><<<(unsigned)((tiles + W - 1) / W), W * 32, 0, (cudaStream_t)stream>>>(*p, obs);
    return check_launch("crypto_observe");
}

extern "C" int32_t frl_crypto_reset(const frl_crypto_params *p, const uint8_t *mask, float *obs, void *stream)
{
    if (int32_t rc = crypto_validate(p)) return rc;
    crypto_reset_kernel<<<(p->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, mask);
    if (int32_t rc = check_launch("crypto_reset")) return rc;
    if (obs) return frl_crypto_observe(p, obs, stream);
    return FRL_OK;
}

extern "C" int32_t frl_crypto_rollout(const frl_crypto_params *p, const void *actions, int32_t actions_f64,
                                      int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps, double *rewards,
                                      uint8_t *flags, float *obs, int32_t obs_mode, int32_t auto_reset, double *stats,
                                      void *stream)
{
    if (int32_t rc = crypto_validate(p)) return rc;
    FRL_REQUIRE(actions != nullptr, "crypto_rollout: actions is NULL");
    FRL_REQUIRE(n_steps >= 1, "crypto_rollout: n_steps must be >= 1 (got %d)", n_steps);
    FRL_REQUIRE(act_env_stride >= p->stock_dim, "crypto_rollout: act_env_stride %lld < stock_dim", (long long)act_env_stride);
    FRL_REQUIRE(obs_mode >= FRL_OBS_NONE && obs_mode <= FRL_OBS_ALL, "crypto_rollout: bad obs_mode %d", obs_mode);
    FRL_REQUIRE(obs_mode == FRL_OBS_NONE || obs != nullptr, "crypto_rollout: obs is NULL but obs_mode=%d", obs_mode);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = p->stock_dim;
#define FRL_GO(SLOTS)                                                                                             \
    do {                                                                                                          \
        if (actions_f64)                                                                                          \
            crypto_launch<SLOTS, double, 2>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags, \
                                            obs, obs_mode, auto_reset, stats, st);                                \
        else                                                                                                      \
            crypto_launch<SLOTS, float, 4>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags,  \
                                           obs, obs_mode, auto_reset, stats, st);                                 \
    } while (0)
    if (D <= 8)
        FRL_GO(8);
    else if (D <= 16)
        FRL_GO(16);
    else
        FRL_GO(32);
#undef FRL_GO
    return check_launch("crypto_rollout");
}

extern "C" int32_t frl_crypto_step(const frl_crypto_params *p, const void *actions, int32_t actions_f64, double *rewards,
                                   uint8_t *flags, float *obs, int32_t auto_reset, double *stats, void *stream)
{
    if (p == nullptr) {
        set_error("crypto_step: params is NULL");
        return FRL_E_INVALID;
    }
    return frl_crypto_rollout(p, actions, actions_f64, (int64_t)p->n_envs * p->stock_dim, p->stock_dim, 1, rewards, flags,
                              obs, obs ? FRL_OBS_LAST : FRL_OBS_NONE, auto_reset, stats, stream);
}
