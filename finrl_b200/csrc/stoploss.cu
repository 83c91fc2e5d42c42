// Sibling env — StockTradingEnvStopLoss (reference: finrl/meta/env_stock_trading/env_stocktrading_stoploss.py).
//
// The cash-penalty env plus per-asset average-buy-price tracking, stop-loss liquidation and three extra
// dot products in the reward.  One thread per env over the D <= 128 assets (the six
// per-asset state arrays are stock-major, so every access is a coalesced warp access), in ONE streaming pass
// over double-buffered state:
//   pass    reward terms of the PREVIOUS step's penalty arrays, np.dot(holdings, closings), the transaction
//           of every asset, proceeds and spend, and the tentative new state (holdings / previous holdings /
//           average buy price / buy counts / both diffs) written to the env's OTHER buffer; the float32
//           holdings stay in the staging rows for the observation writer
//   decide  cash shortage -> terminate (the buffer does not flip; only closing_diff_avg_buy changes, as in
//           the reference) / patient (the buy slots are put back) / flip the buffer.
// np.dot's order is BLAS-specific (1e-9 tolerance); the sums here are sequential in asset order.
#include "common.cuh"
#include "cp_obs.cuh"

namespace frl {
namespace {

__device__ __forceinline__ long long sl_floordiv_ll(long long a, long long b)
{
    long long q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q -= 1;
    return q;
}

struct SlTx {
    double v;      // transaction (shares, signed)
    double cdiff;  // closing_diff_avg_buy of this step
};

// the action pipeline of step() for one asset (:321-361)
template <typename ActT, bool HVEC>
__device__ __forceinline__ SlTx sl_transaction(const frl_stoploss_params &p, ActT a, double hmax, double c, double h, double avg, bool liq,
                                               bool stop_on)
{
    double v;
    if (sizeof(ActT) == 4 && !(HVEC && !p.hmax_vec_f32))
        v = (double)fmul((float)a, (float)hmax);  // scalar (weak Python float) or float32 array: float32 product
    else
        v = dmul((double)a, hmax);
    const bool pos = c > 0.0;
    if (!pos) v = 0.0;                // np.where(closings > 0, actions, 0)
    if (liq) v = -dmul(h, c);         // -(holdings * closings): currency, divided by the price again below
    if (p.discrete_actions) {
        long long q = pos ? (long long)floor_div_f64(v, c) : 0;
        const long long inc = p.shares_increment;
        q = (q >= 0) ? sl_floordiv_ll(q, inc) * inc : sl_floordiv_ll(q + inc, inc) * inc;
        v = (double)q;
    } else {
        v = pos ? __ddiv_rn(v, c) : 0.0;
    }
    v = (v > -h) ? v : -h;            // np.maximum(actions, -holdings)
    SlTx r;
    r.cdiff = dsub(c, dmul(p.stoploss_penalty, avg));  // closings - stoploss_penalty * avg_buy_price
    if (stop_on && r.cdiff < 0.0) v = -h;              // stop-loss: clear the position
    r.v = v;
    return r;
}

// get_reward (:255-290) from the logged (total_assets, cash) and the three penalty dot products
__device__ __forceinline__ double sl_reward(const frl_stoploss_params &p, double total_assets, double cash, int current_step,
                                            double dot_prev_negc, double dot_hold_negp, double dot_hold_posp)
{
    if (current_step == 0) return 0.0;
    double cash_penalty = dsub(dmul(total_assets, p.cash_penalty_proportion), cash);
    if (!(cash_penalty > 0.0)) cash_penalty = 0.0;
    const double stop_loss_penalty = current_step > 1 ? -dot_prev_negc : 0.0;
    const double low_profit_penalty = -dot_hold_negp;
    const double total_penalty = dadd(dadd(cash_penalty, stop_loss_penalty), low_profit_penalty);
    double r = dsub(__ddiv_rn(dadd(dsub(total_assets, total_penalty), dot_hold_posp), p.initial_amount), 1.0);
    return __ddiv_rn(r, (double)current_step);
}

// The six per-asset arrays live in ONE allocation, assets[2][6][D][env_stride]: two buffers of (holdings,
// previous holdings, average buy price, buy counts, closing diff, profit-sell diff).  Bit 1 of an env's `fresh`
// byte names its current buffer.  A step reads the current buffer and writes the other one, so the whole
// update is ONE streaming pass (each array read once and written once); the buffer flips only when the step
// goes through (no cash-shortage termination).
enum { SL_HOLD = 0, SL_PREV = 1, SL_AVG = 2, SL_NB = 3, SL_CD = 4, SL_PD = 5, SL_ARRAYS = 6 };

// Address of (array 0, asset 0) of env n in buffer `which`; array a, asset j then sits at [a * arr + j * ld].
// (A tile-major layout, one contiguous block per 32-env tile, was measured too: 0.794 ms vs 0.781 ms for this
// stock-major one — DRAM-page / TLB locality is not what limits the pass.)
__device__ __forceinline__ double *sl_env(const frl_stoploss_params &p, int which, long long n)
{
    return p.assets + (size_t)which * SL_ARRAYS * p.stock_dim * p.env_stride + n;
}
__device__ __forceinline__ size_t sl_arr(const frl_stoploss_params &p) { return (size_t)p.stock_dim * p.env_stride; }
__device__ __forceinline__ int sl_ld(const frl_stoploss_params &p) { return p.env_stride; }

#ifndef FRL_SL_U
#define FRL_SL_U 2  // assets per software-pipelined batch of the pass (6 state arrays each)
#endif
constexpr int SL_U = FRL_SL_U;

__device__ __forceinline__ void sl_cp_async(float *dst, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void sl_cp_async(double *dst, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

#ifndef FRL_SL_MIN_BLOCKS
#define FRL_SL_MIN_BLOCKS 3  // 128-thread blocks per SM the register allocator must allow
#endif

template <typename ActT, int WARPS, bool HVEC>
__global__ void __launch_bounds__(WARPS * 32, FRL_SL_MIN_BLOCKS * 128 / (WARPS * 32))
stoploss_rollout_kernel(const frl_stoploss_params p, const ActT *__restrict__ actions, long long act_step_stride,
                        long long act_env_stride, int n_steps, double *__restrict__ rewards, uint8_t *__restrict__ flags_out,
                        float *__restrict__ obs, int obs_mode, int auto_reset, double *__restrict__ stats)
{
    stats_exchange_previous(stats);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.n_envs, D = p.stock_dim, T = p.n_days, ld = sl_ld(p);
    const size_t arr = sl_arr(p);  // elements between two arrays of one env
    const int P = D | 1;
    const size_t warp_bytes = (size_t)32 * P * sizeof(ActT) + 32 * sizeof(float) + 32 * sizeof(int);
    unsigned char *base = smem_raw + warp * ((warp_bytes + 15) & ~(size_t)15);
    ActT *stage = reinterpret_cast<ActT *>(base);
    float *cashf = reinterpret_cast<float *>(base + (size_t)32 * P * sizeof(ActT));
    int *di_s = reinterpret_cast<int *>(cashf + 32);
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;

    double cash = p.cash[n], last_cash = p.last_cash[n], last_total = p.last_total[n], sum_trades = p.sum_trades[n];
    int di = p.date_index[n], start = p.start[n];
    const int fresh0 = p.fresh[n];
    bool fresh = (fresh0 & 1) != 0;
    int cur = (fresh0 >> 1) & 1;
    ActT *myrow = stage + (size_t)lane * P;
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0, st_liq = 0.0, st_short = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        if (act_env_stride == D) {
            const ActT *tile = abase + (size_t)env0 * D;
            const int cnt = nvalid * D;
            int row = 0, col = lane;
            while (col >= D) { col -= D; ++row; }
            // every element goes global -> shared with cp.async: the whole tile in flight at once, waited for
            // after the first state batch has been requested
            for (int e = lane; e < 32 * D; e += 32) {
                ActT *dst = stage + row * P + col;
                if (e < cnt)
                    sl_cp_async(dst, tile + e);
                else
                    *dst = ActT(0);
                col += 32;
                while (col >= D) { col -= D; ++row; }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            for (int r = 0; r < 32; ++r)
                for (int j = lane; j < D; j += 32)
                    stage[r * P + j] = r < nvalid ? abase[(size_t)(env0 + r) * act_env_stride + j] : ActT(0);
        }
        const double *rd = sl_env(p, cur, n);   // array a, asset j at rd[a * arr + j * ld]
        double *wr = sl_env(p, cur ^ 1, n);
        const double *crow = p.close + (size_t)di * D;
        // first batch of the pass, requested before the staged actions are waited for
        constexpr int U = SL_U;
        double sb[SL_ARRAYS][U], cb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t o = (size_t)(u < D ? u : D - 1) * ld;
#pragma unroll
            for (int a = 0; a < SL_ARRAYS; ++a) sb[a][u] = __ldcs(rd + a * arr + o);
            cb[u] = __ldg(crow + (u < D ? u : D - 1));
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();

        int flags = 0;
        double reward;
        const int current_step = di - start;
        bool reset_now = false, moved = false;
        // dot products of get_reward over the arrays as they stand at the start of the step
        double d_prev_negc = 0.0, d_hold_negp = 0.0, d_hold_posp = 0.0, asum = 0.0;
        if (di == T - 1) {
            for (int j = 0; j < D; ++j) {
                const size_t o = (size_t)j * ld;
                const double h = rd[SL_HOLD * arr + o], pv = rd[SL_PREV * arr + o], cd = rd[SL_CD * arr + o],
                             pd = rd[SL_PD * arr + o];
                d_prev_negc = dadd(d_prev_negc, dmul(pv, cd < 0.0 ? cd : 0.0));
                d_hold_negp = dadd(d_hold_negp, dmul(h, pd < 0.0 ? pd : 0.0));
                d_hold_posp = dadd(d_hold_posp, dmul(h, pd > 0.0 ? pd : 0.0));
                asum += fabs((double)myrow[j]);
            }
            sum_trades += asum;
            flags = FRL_FLAG_DONE;
            reward = sl_reward(p, last_total, last_cash, current_step, d_prev_negc, d_hold_negp, d_hold_posp);
            reset_now = auto_reset != 0;
        } else {
            const double turbulence = fresh ? 0.0 : __ldg(p.turb + di);
            const bool liq = p.use_turbulence && turbulence >= p.turbulence_threshold;
            if (liq) flags |= FRL_FLAG_LIQUIDATE;
            const double begin_cash = cash;
            const bool stop_on = begin_cash >= dmul(p.stoploss_penalty, p.initial_amount);
            // ---- the pass: sums of the reward and of the trade, tentative new state into the other buffer ----
            double asset_value = 0.0, proceeds = 0.0, spend = 0.0, d_prev_negc_new = 0.0;
            for (int j0 = 0; j0 < D; j0 += U) {
                // software pipeline: the next batch's 6 x U loads are in flight while this one is traded
                double sn[SL_ARRAYS][U], cn[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + U + u < D ? j0 + U + u : D - 1;
                    const size_t o = (size_t)j * ld;
#pragma unroll
                    for (int a = 0; a < SL_ARRAYS; ++a) sn[a][u] = __ldcs(rd + a * arr + o);
                    cn[u] = __ldg(crow + j);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u;
                    if (j < D) {
                        const double c = cb[u], h = sb[SL_HOLD][u], pv = sb[SL_PREV][u], avg = sb[SL_AVG][u],
                                     cd = sb[SL_CD][u], pd = sb[SL_PD][u];
                        double nb = sb[SL_NB][u];
                        const ActT a = myrow[j];
                        const SlTx t = sl_transaction<ActT, HVEC>(p, a, HVEC ? __ldg(p.hmax_vec + j) : p.hmax, c, h, avg, liq, stop_on);
                        asum += fabs((double)a);
                        asset_value = dadd(asset_value, dmul(h, c));
                        d_prev_negc = dadd(d_prev_negc, dmul(pv, cd < 0.0 ? cd : 0.0));
                        d_hold_negp = dadd(d_hold_negp, dmul(h, pd < 0.0 ? pd : 0.0));
                        d_hold_posp = dadd(d_hold_posp, dmul(h, pd > 0.0 ? pd : 0.0));
                        d_prev_negc_new = dadd(d_prev_negc_new, dmul(pv, t.cdiff < 0.0 ? t.cdiff : 0.0));
                        const double sell = -(t.v < 0.0 ? t.v : 0.0), buy = t.v > 0.0 ? t.v : 0.0;
                        proceeds = dadd(proceeds, dmul(sell, c));
                        spend = dadd(spend, dmul(buy, c));
                        // tentative update, as if the step goes through with its buys
                        const double scp = sell > 0.0 ? c : 0.0;  // profitable sells (:391-404)
                        const double pdn = (dsub(scp, avg) > 0.0) ? dsub(c, dmul(p.min_profit_penalty, avg)) : 0.0;
                        const double hn = dadd(h, t.v);
                        double avgn = avg;  // incremental average buy price (:416-428)
                        if (buy > 0.0) {
                            nb = dadd(nb, 1.0);
                            avgn = dadd(avg, __ddiv_rn(dsub(c, avg), nb));
                        }
                        if (!(hn > 0.0)) {
                            nb = 0.0;
                            avgn = 0.0;
                        }
                        if (valid) {
                            const size_t o = (size_t)j * ld;
                            wr[SL_HOLD * arr + o] = hn;
                            wr[SL_PREV * arr + o] = h;
                            wr[SL_AVG * arr + o] = avgn;
                            wr[SL_NB * arr + o] = nb;
                            wr[SL_CD * arr + o] = t.cdiff;
                            wr[SL_PD * arr + o] = pdn;
                        }
                        *reinterpret_cast<float *>(myrow + j) = (float)hn;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int a = 0; a < SL_ARRAYS; ++a) sb[a][u] = sn[a][u];
                    cb[u] = cn[u];
                }
            }
            sum_trades += asum;
            // reward from the PREVIOUS log entry (:313), then this step's entry is logged (:315-319)
            reward = sl_reward(p, last_total, last_cash, current_step, d_prev_negc, d_hold_negp, d_hold_posp);
            last_cash = begin_cash;
            last_total = dadd(begin_cash, asset_value);
            double costs = dmul(proceeds, p.sell_cost_pct);
            double coh = dadd(begin_cash, proceeds);
            costs = dadd(costs, dmul(spend, p.buy_cost_pct));
            bool terminate = false, no_buys = false;
            if (dadd(spend, costs) > coh) {
                flags |= FRL_FLAG_SHORTAGE;
                if (p.patient) {
                    no_buys = true;
                    spend = 0.0;
                    costs = 0.0;
                } else {
                    terminate = true;
                }
            }
            if (terminate) {
                // return_terminal(reward=self.get_reward()) (:383-386): this step's log entry and closing diff,
                // the previous profit diffs and holdings; closing_diff_avg_buy is the only array that changed —
                // it is written into the CURRENT buffer, the tentative one is dropped
                flags |= FRL_FLAG_DONE;
                reward = sl_reward(p, last_total, last_cash, current_step, d_prev_negc_new, d_hold_negp, d_hold_posp);
                if (valid && !auto_reset) {
                    double *cw = sl_env(p, cur, n);
                    for (int j = 0; j < D; ++j) {
                        const size_t o = (size_t)j * ld;
                        cw[SL_CD * arr + o] = dsub(__ldg(crow + j), dmul(p.stoploss_penalty, cw[SL_AVG * arr + o]));
                    }
                }
                reset_now = auto_reset != 0;
            } else {
                cash = dsub(dsub(coh, spend), costs);
                if (no_buys) {
                    // patient: the buys do not happen (:376-381).  Buy counts, average price and the profit diffs
                    // were driven by the PRE-patient vectors (like the reference), so only the holdings of the buy
                    // slots go back — and with them the "position closed" reset of count and average
                    for (int j = 0; j < D; ++j) {
                        const size_t o = (size_t)j * ld;
                        const double h = rd[SL_HOLD * arr + o], hn = wr[SL_HOLD * arr + o];
                        if (hn > h) {
                            const double hk = dadd(h, 0.0);
                            if (valid) {
                                wr[SL_HOLD * arr + o] = hk;
                                if (!(hk > 0.0)) {
                                    wr[SL_AVG * arr + o] = 0.0;
                                    wr[SL_NB * arr + o] = 0.0;
                                }
                            }
                            *reinterpret_cast<float *>(myrow + j) = (float)hk;
                        }
                    }
                }
                cur ^= 1;
                moved = true;
                di += 1;
                if (p.use_turbulence) fresh = false;
            }
        }
        if (valid) {
            if (rewards) rewards[(size_t)k * N + n] = reward;
            if (flags_out) flags_out[(size_t)k * N + n] = (uint8_t)flags;
            st_r += reward;
            st_r2 += reward * reward;
            if (flags & FRL_FLAG_DONE) {
                st_done += 1.0;
                st_epi += last_total;
            }
            if (flags & FRL_FLAG_LIQUIDATE) st_liq += 1.0;
            if (flags & FRL_FLAG_SHORTAGE) st_short += 1.0;
        }
        if (reset_now) {  // reset (:134-165); starting point 0 or, with p.random_start, randint(0, int(T * 0.5))
            cash = p.initial_amount;
            double *cw = sl_env(p, cur, n);
            for (int j = 0; j < D; ++j) {
                const size_t o = (size_t)j * ld;
                if (valid) {
#pragma unroll
                    for (int a = 0; a < SL_ARRAYS; ++a) cw[a * arr + o] = 0.0;
                }
                *reinterpret_cast<float *>(myrow + j) = 0.0f;
            }
            moved = true;
            start = p.random_start ? reset_randint(reset_bits(p.reset_seed, n, k, 0), (int)(p.n_days * 0.5)) : 0;
            di = start;
            fresh = true;
            sum_trades = 0.0;
            last_cash = 0.0;
            last_total = 0.0;
        }
        if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
            if (!moved) {
                const double *cr = sl_env(p, cur, n);
                for (int j = 0; j < D; ++j) *reinterpret_cast<float *>(myrow + j) = (float)cr[SL_HOLD * arr + (size_t)j * ld];
            }
            cashf[lane] = (float)cash;
            di_s[lane] = di;
            __syncwarp();
            float *o = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * p.obs_dim : (size_t)0);
            cp_write_obs_tile<ActT>(p, stage, P, cashf, di_s, o, env0, nvalid, lane, D);
        }
    }
    if (valid) {
        p.cash[n] = cash;
        p.date_index[n] = di;
        p.start[n] = start;
        p.fresh[n] = (uint8_t)((fresh ? 1 : 0) | (cur << 1));
        p.last_cash[n] = last_cash;
        p.last_total[n] = last_total;
        p.sum_trades[n] = sum_trades;
    }
    if (stats) {
        double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, valid ? last_total : 0.0, st_liq,
                                 valid ? (double)n_steps : 0.0, st_short};
        reduce_stats8(v, lane, stats);
    }
}

__global__ void stoploss_reset_kernel(const frl_stoploss_params p, const uint8_t *__restrict__ mask,
                                      const int32_t *__restrict__ start_points)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.n_envs) return;
    if (mask && !mask[n]) return;
    const size_t arr = sl_arr(p);
    const int ld = sl_ld(p);
    double *b0 = sl_env(p, 0, n);
    for (int j = 0; j < p.stock_dim; ++j)
        for (int a = 0; a < SL_ARRAYS; ++a) b0[a * arr + (size_t)j * ld] = 0.0;
    const int sp = start_points ? start_points[n] : 0;
    p.cash[n] = p.initial_amount;
    p.date_index[n] = sp;
    p.start[n] = sp;
    p.fresh[n] = 1;  // fresh, current buffer 0
    p.last_cash[n] = 0.0;
    p.last_total[n] = 0.0;
    p.sum_trades[n] = 0.0;
}

__global__ void stoploss_observe_kernel(const frl_stoploss_params p, float *__restrict__ obs)
{
    const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= p.n_envs) return;
    const int O = p.obs_dim, D = p.stock_dim;
    const float *trow = p.obs_tmpl + (size_t)p.date_index[n] * O;
    const double *hold = sl_env(p, (p.fresh[n] >> 1) & 1, n);  // SL_HOLD is array 0
    const int ld = sl_ld(p);
    float *orow = obs + (size_t)n * O;
    for (int pos = lane; pos < O; pos += 32) {
        float v;
        if (pos == 0)
            v = (float)p.cash[n];
        else if (pos <= D)
            v = (float)hold[(size_t)(pos - 1) * ld];
        else
            v = __ldg(trow + pos);
        orow[pos] = v;
    }
}

int32_t sl_validate(const frl_stoploss_params *p)
{
    FRL_REQUIRE(p != nullptr, "stoploss: params is NULL");
    FRL_REQUIRE(p->n_envs >= 1, "stoploss: n_envs must be >= 1 (got %d)", p->n_envs);
    FRL_REQUIRE(p->stock_dim >= 1 && p->stock_dim <= 128, "stoploss: stock_dim must be in 1..128 (got %d)", p->stock_dim);
    FRL_REQUIRE(p->n_cols >= 0 && p->n_days >= 1, "stoploss: bad n_cols/n_days (%d, %d)", p->n_cols, p->n_days);
    FRL_REQUIRE(p->obs_dim == 1 + p->stock_dim + p->stock_dim * p->n_cols, "stoploss: obs_dim %d != 1 + D + D*C = %d",
                p->obs_dim, 1 + p->stock_dim + p->stock_dim * p->n_cols);
    FRL_REQUIRE(p->env_stride >= p->n_envs, "stoploss: env_stride %d < n_envs %d", p->env_stride, p->n_envs);
    FRL_REQUIRE(!p->discrete_actions || p->shares_increment >= 1, "stoploss: shares_increment must be >= 1");
    FRL_REQUIRE(p->close && p->obs_tmpl && (!p->use_turbulence || p->turb), "stoploss: table pointer is NULL");
    FRL_REQUIRE(p->cash && p->assets && p->date_index && p->start && p->fresh && p->last_cash && p->last_total && p->sum_trades,
                "stoploss: state pointer is NULL");
    return FRL_OK;
}

template <typename ActT, int WARPS>
int32_t sl_launch(const frl_stoploss_params &p, const void *actions, long long sstride, long long estride, int n_steps,
                  double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    const int P = p.stock_dim | 1;
    const size_t warp_bytes = (size_t)32 * P * sizeof(ActT) + 32 * sizeof(float) + 32 * sizeof(int);
    const size_t smem = WARPS * ((warp_bytes + 15) & ~(size_t)15);
    auto kern = p.hmax_vec ? stoploss_rollout_kernel<ActT, WARPS, true> : stoploss_rollout_kernel<ActT, WARPS, false>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("stoploss_rollout: cannot reserve %zu B of shared memory (%s)", smem, cudaGetErrorString(e));
            return FRL_E_CUDA;
        }
    }
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    kern<<<(unsigned)((tiles + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p, (const ActT *)actions, sstride, estride, n_steps,
                                                                           rewards, flags, obs, obs_mode, auto_reset, stats);
    return check_launch("stoploss_rollout");
}

}  // namespace
}  // namespace frl

using namespace frl;

extern "C" int32_t frl_stoploss_observe(const frl_stoploss_params *p, float *obs, void *stream)
{
    if (int32_t rc = sl_validate(p)) return rc;
    FRL_REQUIRE(obs != nullptr, "stoploss_observe: obs is NULL");
    const long long threads = (long long)p->n_envs * 32;
    stoploss_observe_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, obs);
    return check_launch("stoploss_observe");
}

extern "C" int32_t frl_stoploss_reset(const frl_stoploss_params *p, const uint8_t *mask, const int32_t *start_points,
                                      float *obs, void *stream)
{
    if (int32_t rc = sl_validate(p)) return rc;
    stoploss_reset_kernel<<<(p->n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*p, mask, start_points);
    if (int32_t rc = check_launch("stoploss_reset")) return rc;
    if (obs) return frl_stoploss_observe(p, obs, stream);
    return FRL_OK;
}

extern "C" int32_t frl_stoploss_rollout(const frl_stoploss_params *p, const void *actions, int32_t actions_f64,
                                        int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps, double *rewards,
                                        uint8_t *flags, float *obs, int32_t obs_mode, int32_t auto_reset, double *stats,
                                        void *stream)
{
    if (int32_t rc = sl_validate(p)) return rc;
    FRL_REQUIRE(actions != nullptr, "stoploss_rollout: actions is NULL");
    FRL_REQUIRE(n_steps >= 1, "stoploss_rollout: n_steps must be >= 1 (got %d)", n_steps);
    FRL_REQUIRE(act_env_stride >= p->stock_dim, "stoploss_rollout: act_env_stride %lld < stock_dim", (long long)act_env_stride);
    FRL_REQUIRE(obs_mode >= FRL_OBS_NONE && obs_mode <= FRL_OBS_ALL, "stoploss_rollout: bad obs_mode %d", obs_mode);
    FRL_REQUIRE(obs_mode == FRL_OBS_NONE || obs != nullptr, "stoploss_rollout: obs is NULL but obs_mode=%d", obs_mode);
    cudaStream_t st = (cudaStream_t)stream;
    if (actions_f64)
        return sl_launch<double, 2>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags, obs, obs_mode,
                                    auto_reset, stats, st);
    return sl_launch<float, 4>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags, obs, obs_mode, auto_reset,
                               stats, st);
}

extern "C" int32_t frl_stoploss_step(const frl_stoploss_params *p, const void *actions, int32_t actions_f64, double *rewards,
                                     uint8_t *flags, float *obs, int32_t auto_reset, double *stats, void *stream)
{
    if (p == nullptr) {
        set_error("stoploss_step: params is NULL");
        return FRL_E_INVALID;
    }
    return frl_stoploss_rollout(p, actions, actions_f64, (int64_t)p->n_envs * p->stock_dim, p->stock_dim, 1, rewards, flags, obs,
                                obs ? FRL_OBS_LAST : FRL_OBS_NONE, auto_reset, stats, stream);
}
