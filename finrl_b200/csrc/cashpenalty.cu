// A4 — StockTradingEnvCashpenalty (reference: finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py).
//
// This env is vector arithmetic over D (up to NASDAQ-100) assets with three dot products and NO
// sequential per-stock dependency, and its state is dominated by D fractional fp64 holdings — so the
// mapping is ONE WARP PER ENV: lane l owns assets l, l+32, l+64, l+96 in registers, the three np.dot
// reductions are xor-shuffle butterflies (every lane ends with the same bits), and every global
// access (action row, holdings row, 601-float observation row) is a coalesced row access with no
// shared-memory staging.  Warps walk the env range with a grid stride.
#include "common.cuh"

namespace frl {
namespace {

constexpr int kPerLane = 4;  // D <= 128

__device__ __forceinline__ double cp_reward(const frl_cashpenalty_params &p, double assets, double cash, int current_step)
{
    // get_reward (:246-256)
    if (current_step == 0) return 0.0;
    double pen = dsub(dmul(assets, p.cash_penalty_proportion), cash);
    if (!(pen > 0.0)) pen = 0.0;
    assets = dsub(assets, pen);
    double r = dsub(__ddiv_rn(assets, p.initial_amount), 1.0);
    return __ddiv_rn(r, (double)current_step);
}

__device__ __forceinline__ long long floordiv_ll(long long a, long long b)
{
    long long q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q -= 1;
    return q;
}

// four independent butterflies interleaved (ILP): every lane ends with the same bits for each sum
__device__ __forceinline__ void warp_sum4(double &a, double &b, double &c, double &d)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
        const double tc = __shfl_xor_sync(0xffffffffu, c, o), td = __shfl_xor_sync(0xffffffffu, d, o);
        a += ta;
        b += tb;
        c += tc;
        d += td;
    }
}

constexpr int kTmplChunks = 20;  // register-cached template covers O <= 640 (NASDAQ-100 x 5 columns: 601)

// one observation row: [coh, holdings x D, daily information of date di] as float32
__device__ __forceinline__ void cp_write_obs_row(const frl_cashpenalty_params &p, float *__restrict__ orow, double cash,
                                                 const double (&h)[kPerLane], int di, int lane, float (&t)[kTmplChunks],
                                                 int &cached_di)
{
    const int O = p.obs_dim, D = p.stock_dim;
    const float *trow = p.obs_tmpl + (size_t)di * O;
    const bool use_cache = O <= kTmplChunks * 32;
    if (use_cache && di != cached_di) {  // warp-uniform: consecutive envs of a warp usually share the date
#pragma unroll
        for (int c = 0; c < kTmplChunks; ++c) {
            const int pos = lane + 32 * c;
            t[c] = pos < O ? __ldg(trow + pos) : 0.0f;
        }
        cached_di = di;
    }
    const int src = (lane - 1) & 31;
    float hf[kPerLane];
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) hf[i] = (float)h[i];
    // positions 1..D hold holdings[pos-1]; asset j lives in lane j%32, slot j/32
#pragma unroll
    for (int c = 0; c <= kPerLane; ++c) {
        const int pos = lane + 32 * c;
        if (32 * c > D) break;  // warp-uniform
        const float same = __shfl_sync(0xffffffffu, c < kPerLane ? hf[c] : 0.0f, src);  // lanes >= 1: slot c
        const float prev = __shfl_sync(0xffffffffu, c > 0 ? hf[c - 1] : 0.0f, src);     // lane 0: slot c-1
        if (pos < O) {
            float v;
            if (pos == 0)
                v = (float)cash;
            else if (pos <= D)
                v = lane == 0 ? prev : same;
            else
                v = use_cache ? t[c] : __ldg(trow + pos);
            orow[pos] = v;
        }
    }
    const int first_c = (D >> 5) + 1;  // first chunk entirely past the holdings
    if (use_cache) {
#pragma unroll
        for (int c = 1; c < kTmplChunks; ++c)
            if (c >= first_c && lane + 32 * c < O) orow[lane + 32 * c] = t[c];
    } else {
        for (int pos = (first_c << 5) + lane; pos < O; pos += 32) orow[pos] = __ldg(trow + pos);
    }
}

template <typename ActT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
cashpenalty_rollout_kernel(const frl_cashpenalty_params p, const ActT *__restrict__ actions, long long act_step_stride,
                           long long act_env_stride, int n_steps, double *__restrict__ rewards,
                           uint8_t *__restrict__ flags_out, float *__restrict__ obs, int obs_mode, int auto_reset,
                           double *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * WARPS;
    const int N = p.n_envs, D = p.stock_dim, T = p.n_days, O = p.obs_dim;
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0, st_asset = 0.0, st_liq = 0.0, st_steps = 0.0, st_short = 0.0;
    float tmpl[kTmplChunks];
    int tmpl_di = -1;

    for (long long n = warp0; n < N; n += nwarps) {
        // ---- load state (scalars replicated in every lane; holdings: 4 assets per lane) ----
        double cash = p.cash[n], last_cash = p.last_cash[n], last_total = p.last_total[n], sum_trades = p.sum_trades[n];
        int di = p.date_index[n], start = p.start[n];
        bool fresh = p.fresh[n] != 0;
        double h[kPerLane];
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            const int j = lane + 32 * i;
            h[i] = j < D ? p.hold[(size_t)n * D + j] : 0.0;
        }
        for (int k = 0; k < n_steps; ++k) {
            const ActT *arow = actions + (size_t)k * act_step_stride + (size_t)n * act_env_stride;
            ActT a[kPerLane];
            double asum = 0.0;
#pragma unroll
            for (int i = 0; i < kPerLane; ++i) {
                const int j = lane + 32 * i;
                a[i] = j < D ? arow[j] : ActT(0);
                asum += fabs((double)a[i]);
            }
            int flags = 0;
            double reward;
            const int current_step = di - start;
            bool reset_now = false;
            if (di == T - 1) {
                // last date (:308-310): reward from the previously logged (assets, cash); state unchanged
                sum_trades += warp_sum(asum);  // self.sum_trades += np.sum(np.abs(actions)) (:302), logging only
                flags = FRL_FLAG_DONE;
                reward = cp_reward(p, last_total, last_cash, current_step);
                reset_now = auto_reset != 0;
            } else {
                const double *crow = p.close + (size_t)di * D;
                double c[kPerLane], part = 0.0;
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) {
                    const int j = lane + 32 * i;
                    c[i] = j < D ? __ldg(crow + j) : 0.0;
                    part += h[i] * c[i];
                }
                const double begin_cash = cash;

                // ---- get_transactions (:258-298) ----
                const double turbulence = fresh ? 0.0 : __ldg(p.turb + di);
                const bool liq = p.use_turbulence && turbulence >= p.turbulence_threshold;
                double tx[kPerLane], psell = 0.0, pbuy = 0.0;
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) {
                    const int j = lane + 32 * i;
                    double v;  // actions * hmax in the input dtype
                    if (sizeof(ActT) == 4)
                        v = (double)fmul((float)a[i], (float)p.hmax);
                    else
                        v = dmul((double)a[i], p.hmax);
                    if (!(c[i] > 0.0)) v = 0.0;  // np.where(closings > 0, actions, 0)
                    if (j < D) {
                        if (p.discrete_actions) {
                            long long q = (long long)floor_div_f64(v, c[i]);  // actions // closings, astype(int)
                            const long long inc = p.shares_increment;
                            q = (q >= 0) ? floordiv_ll(q, inc) * inc : floordiv_ll(q + inc, inc) * inc;
                            v = (double)q;
                        } else {
                            v = __ddiv_rn(v, c[i]);
                        }
                        v = (v > -h[i]) ? v : -h[i];  // np.maximum(actions, -holdings)
                        if (liq) v = -h[i];           // turbulence: clear out all positions
                    } else {
                        v = 0.0;
                    }
                    tx[i] = v;
                    psell += (v < 0.0 ? -v : 0.0) * c[i];
                    pbuy += (v > 0.0 ? v : 0.0) * c[i];
                }
                if (liq) flags |= FRL_FLAG_LIQUIDATE;
                // np.sum(|actions|), np.dot(holdings, closings) (:319), np.dot(sells, closings), np.dot(buys,
                // closings) (:334,:339): one fused reduction
                warp_sum4(asum, part, psell, pbuy);
                sum_trades += asum;
                const double asset_value = part;
                last_cash = begin_cash;
                last_total = dadd(begin_cash, asset_value);
                reward = cp_reward(p, last_total, last_cash, current_step);  // computed BEFORE trading (:326)
                const double proceeds = psell;
                double spend = pbuy;
                double costs = dmul(proceeds, p.sell_cost_pct);
                double coh = dadd(begin_cash, proceeds);
                costs = dadd(costs, dmul(spend, p.buy_cost_pct));
                bool terminate = false;
                if (dadd(spend, costs) > coh) {
                    flags |= FRL_FLAG_SHORTAGE;
                    if (p.patient) {  // no buys until there is cash again; sell costs are dropped too (Q9)
#pragma unroll
                        for (int i = 0; i < kPerLane; ++i)
                            if (tx[i] > 0.0) tx[i] = 0.0;
                        spend = 0.0;
                        costs = 0.0;
                    } else {
                        terminate = true;  // CASH SHORTAGE (:349-353): state unchanged, current reward, done
                    }
                }
                if (terminate) {
                    flags |= FRL_FLAG_DONE;
                    reset_now = auto_reset != 0;
                } else {
                    cash = dsub(dsub(coh, spend), costs);
#pragma unroll
                    for (int i = 0; i < kPerLane; ++i) h[i] = dadd(h[i], tx[i]);
                    di += 1;
                    if (p.use_turbulence) fresh = false;  // self.turbulence is refreshed only with a threshold
                }
            }
            if (lane == 0) {
                if (rewards) rewards[(size_t)k * N + n] = reward;
                if (flags_out) flags_out[(size_t)k * N + n] = (uint8_t)flags;
                st_r += reward;
                st_r2 += reward * reward;
                st_steps += 1.0;
                if (flags & FRL_FLAG_DONE) {
                    st_done += 1.0;
                    st_epi += last_total;
                }
                if (flags & FRL_FLAG_LIQUIDATE) st_liq += 1.0;
                if (flags & FRL_FLAG_SHORTAGE) st_short += 1.0;
            }
            if (reset_now) {  // DummyVecEnv.step_wait -> reset (:132-158), random_start=False
                cash = p.initial_amount;
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) h[i] = 0.0;
                di = 0;
                start = 0;
                fresh = true;
                sum_trades = 0.0;
                last_cash = 0.0;
                last_total = 0.0;
            }
            if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
                float *orow = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * O : (size_t)0) + (size_t)n * O;
                cp_write_obs_row(p, orow, cash, h, di, lane, tmpl, tmpl_di);
            }
        }
        // ---- store state ----
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            const int j = lane + 32 * i;
            if (j < D) p.hold[(size_t)n * D + j] = h[i];
        }
        if (lane == 0) {
            p.cash[n] = cash;
            p.date_index[n] = di;
            p.start[n] = start;
            p.fresh[n] = fresh ? 1 : 0;
            p.last_cash[n] = last_cash;
            p.last_total[n] = last_total;
            p.sum_trades[n] = sum_trades;
            st_asset += last_total;
        }
    }
    if (stats && lane == 0) {
        const double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, st_asset, st_liq, st_steps, st_short};
#pragma unroll
        for (int i = 0; i < FRL_N_STATS; ++i)
            if (v[i] != 0.0) atomicAdd(stats + i, v[i]);
    }
}

__global__ void cashpenalty_reset_kernel(const frl_cashpenalty_params p, const uint8_t *__restrict__ mask,
                                         const int32_t *__restrict__ start_points)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= p.n_envs) return;
    if (mask && !mask[w]) return;
    for (int j = lane; j < p.stock_dim; j += 32) p.hold[(size_t)w * p.stock_dim + j] = 0.0;
    if (lane == 0) {
        const int sp = start_points ? start_points[w] : 0;
        p.cash[w] = p.initial_amount;
        p.date_index[w] = sp;
        p.start[w] = sp;
        p.fresh[w] = 1;
        p.last_cash[w] = 0.0;
        p.last_total[w] = 0.0;
        p.sum_trades[w] = 0.0;
    }
}

__global__ void cashpenalty_observe_kernel(const frl_cashpenalty_params p, float *__restrict__ obs)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= p.n_envs) return;
    double h[kPerLane];
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        const int j = lane + 32 * i;
        h[i] = j < p.stock_dim ? p.hold[(size_t)w * p.stock_dim + j] : 0.0;
    }
    float tmpl[kTmplChunks];
    int tmpl_di = -1;
    cp_write_obs_row(p, obs + (size_t)w * p.obs_dim, p.cash[w], h, p.date_index[w], lane, tmpl, tmpl_di);
}

int32_t cp_validate(const frl_cashpenalty_params *p)
{
    FRL_REQUIRE(p != nullptr, "cashpenalty: params is NULL");
    FRL_REQUIRE(p->n_envs >= 1, "cashpenalty: n_envs must be >= 1 (got %d)", p->n_envs);
    FRL_REQUIRE(p->stock_dim >= 1 && p->stock_dim <= 32 * kPerLane, "cashpenalty: stock_dim must be in 1..128 (got %d)",
                p->stock_dim);
    FRL_REQUIRE(p->n_cols >= 0 && p->n_days >= 1, "cashpenalty: bad n_cols/n_days (%d, %d)", p->n_cols, p->n_days);
    FRL_REQUIRE(p->obs_dim == 1 + p->stock_dim + p->stock_dim * p->n_cols, "cashpenalty: obs_dim %d != 1 + D + D*C = %d",
                p->obs_dim, 1 + p->stock_dim + p->stock_dim * p->n_cols);
    FRL_REQUIRE(!p->discrete_actions || p->shares_increment >= 1, "cashpenalty: shares_increment must be >= 1");
    FRL_REQUIRE(p->close && p->obs_tmpl && (!p->use_turbulence || p->turb), "cashpenalty: table pointer is NULL");
    FRL_REQUIRE(p->cash && p->hold && p->date_index && p->start && p->fresh && p->last_cash && p->last_total && p->sum_trades,
                "cashpenalty: state pointer is NULL");
    return FRL_OK;
}

}  // namespace
}  // namespace frl

using namespace frl;

extern "C" int32_t frl_cashpenalty_observe(const frl_cashpenalty_params *p, float *obs, void *stream)
{
    if (int32_t rc = cp_validate(p)) return rc;
    FRL_REQUIRE(obs != nullptr, "cashpenalty_observe: obs is NULL");
    const long long threads = (long long)p->n_envs * 32;
    cashpenalty_observe_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, obs);
    return check_launch("cashpenalty_observe");
}

extern "C" int32_t frl_cashpenalty_reset(const frl_cashpenalty_params *p, const uint8_t *mask, const int32_t *start_points,
                                         float *obs, void *stream)
{
    if (int32_t rc = cp_validate(p)) return rc;
    const long long threads = (long long)p->n_envs * 32;
    cashpenalty_reset_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, mask, start_points);
    if (int32_t rc = check_launch("cashpenalty_reset")) return rc;
    if (obs) return frl_cashpenalty_observe(p, obs, stream);
    return FRL_OK;
}

extern "C" int32_t frl_cashpenalty_rollout(const frl_cashpenalty_params *p, const void *actions, int32_t actions_f64,
                                           int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps,
                                           double *rewards, uint8_t *flags, float *obs, int32_t obs_mode,
                                           int32_t auto_reset, double *stats, void *stream)
{
    if (int32_t rc = cp_validate(p)) return rc;
    FRL_REQUIRE(actions != nullptr, "cashpenalty_rollout: actions is NULL");
    FRL_REQUIRE(n_steps >= 1, "cashpenalty_rollout: n_steps must be >= 1 (got %d)", n_steps);
    FRL_REQUIRE(act_env_stride >= p->stock_dim, "cashpenalty_rollout: act_env_stride %lld < stock_dim", (long long)act_env_stride);
    FRL_REQUIRE(obs_mode >= FRL_OBS_NONE && obs_mode <= FRL_OBS_ALL, "cashpenalty_rollout: bad obs_mode %d", obs_mode);
    FRL_REQUIRE(obs_mode == FRL_OBS_NONE || obs != nullptr, "cashpenalty_rollout: obs is NULL but obs_mode=%d", obs_mode);
    constexpr int W = 8;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent-style grid: a multiple of the SM count, warps stride over the env range
    const long long need = ((long long)p->n_envs + W - 1) / W;
    const unsigned grid = (unsigned)(need < (long long)sms * 8 ? need : (long long)sms * 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (actions_f64)
        cashpenalty_rollout_kernel<double, W><<<grid, W * 32, 0, st>>>(*p, (const double *)actions, act_step_stride,
                                                                       act_env_stride, n_steps, rewards, flags, obs,
                                                                       obs_mode, auto_reset, stats);
    else
        cashpenalty_rollout_kernel<float, W><<<grid, W * 32, 0, st>>>(*p, (const float *)actions, act_step_stride,
                                                                      act_env_stride, n_steps, rewards, flags, obs,
                                                                      obs_mode, auto_reset, stats);
    return check_launch("cashpenalty_rollout");
}

extern "C" int32_t frl_cashpenalty_step(const frl_cashpenalty_params *p, const void *actions, int32_t actions_f64,
                                        double *rewards, uint8_t *flags, float *obs, int32_t auto_reset, double *stats,
                                        void *stream)
{
    if (p == nullptr) {
        set_error("cashpenalty_step: params is NULL");
        return FRL_E_INVALID;
    }
    return frl_cashpenalty_rollout(p, actions, actions_f64, (int64_t)p->n_envs * p->stock_dim, p->stock_dim, 1, rewards,
                                   flags, obs, obs ? FRL_OBS_LAST : FRL_OBS_NONE, auto_reset, stats, stream);
}
