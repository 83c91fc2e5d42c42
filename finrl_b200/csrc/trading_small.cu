// A1 — StockTradingEnv, LOW-LATENCY variant for small batches (BASELINE config 2: 4096 envs, K = 64).
//
// trading.cu maps one thread to one env: ideal when there are enough envs to fill the machine, but a
// fused K-step rollout of a few thousand envs is bound by the per-step latency of that single thread
// (~7.7 us: a 240-compare-exchange network, three 30-term fp64 chains and the trade loops, all serial).
// Here EIGHT LANES share one env (a warp carries 4 envs) and everything that does not depend on the
// running cash is done in parallel across them:
//   * the argsort network runs distributed: slot s lives in lane s/R, register s%R (R = slots/8); the
//     stages whose partner is in another lane exchange packed keys with __shfl_xor_sync, with the same
//     strict-greater / no-swap-on-tie rule (bit-identical order);
//   * per sorted slot, one lane forms the sell terms (independent of cash) and, for buys, the price, the
//     "enough cash" threshold (a+1)*unit and the full-size spend / cost — all into shared memory;
//   * the group's leader lane then runs only the irreducible serial part in the reference's order:
//     cash/cost += sell terms, the cash-limited buys (one compare + one subtract each unless cash is short),
//     and the sequential total-asset sum over products the other lanes prepared.
// Same arithmetic, same order => same bits as trading.cu and the reference (the whole parity suite runs
// under FRL_TRADING_KERNEL=small as well).  It issues ~4x more warp-instructions per env-step, so it only
// pays while the machine is under-filled.  Measured on B200, fused K=64 rollout (ms per launch):
//     n_envs     1024    2048    4096    8192    16384
//     tile       0.493   0.499   0.516   0.519   0.521
//     small      0.288   0.294   0.356   0.478   0.999
// => the host uses this kernel for n_envs <= 8192 (frl_trading_rollout).
#include "trading_common.cuh"

#ifndef FRL_SMALL_WARPS
#define FRL_SMALL_WARPS 2
#endif

#ifndef FRL_SMALL_DCT30
#define FRL_SMALL_DCT30 1  // DOW-30 instantiation of the 8-lanes-per-env kernel (A/B switch)
#endif

namespace frl {
namespace {

constexpr int kGroup = 8;  // lanes per env

template <int S>
struct alignas(16) OctEnv {  // per-env scratch, S = 8R slots
    double sell_cash[S], sell_cost[S];  // by sorted position (0.0 where nothing is sold)
    // enabled buys, COMPACTED in execution order (largest action first): entry 0 is processed first
    double buy_thr[S], buy_spend[S], buy_cost[S], buy_p[S];
    double prod[S];  // by stock index: price * holding for the sequential asset sum
    int buy_aj[S];   // (a << IB) | j
    int buy_h[S];    // holding before the buy
    int hold[S];     // by stock index
};

// packed keys (a << IB) | index with IB index bits (5 up to 32 slots, 7 up to 128)
template <int R>
struct KeyBits {
    static constexpr int IB = (8 * R <= 32) ? 5 : 7;
    static constexpr int MASK = (1 << IB) - 1;
    static constexpr int AMAX = (1 << (31 - IB)) - 1;  // |int(action*hmax)| clamp that keeps the key in int32
                                                       // (IB = 5: 2^26 - 1, the same clamp as trading.cu)
};

// strict-greater compare-exchange on packed keys: ties keep network order
template <int MASK>
__device__ __forceinline__ void cex_local(int &lo, int &hi)
{
    const bool sw = lo > (hi | MASK);
    const int t = sw ? hi : lo;
    hi = sw ? lo : hi;
    lo = t;
}
template <int MASK>
__device__ __forceinline__ int cex_remote(int mine, int other, bool i_am_lo)
{
    const bool sw = i_am_lo ? (mine > (other | MASK)) : (other > (mine | MASK));
    return sw ? other : mine;
}

// The bitonic network of np.argsort (flip stage + half-cleaners per block size) on 8*R slots spread over
// the 8 lanes of a group: slot s = l*R + r.
template <int R>
__device__ __forceinline__ void distributed_network(int (&key)[R], int l, unsigned gmask)
{
    constexpr int SLOTS = kGroup * R;
    constexpr int MASK = KeyBits<R>::MASK;
#pragma unroll
    for (int blk = 2; blk <= SLOTS; blk <<= 1) {
        // ---- flip stage: pairs (b+i, b+blk-1-i) ----
        if (blk <= R) {
#pragma unroll
            for (int b = 0; b < R; b += blk)
#pragma unroll
                for (int i = 0; i < blk / 2; ++i) cex_local<MASK>(key[b + i], key[b + blk - 1 - i]);
        } else {
            const int lanes = blk / R;                   // lanes spanned by one block
            const int partner_xor = lanes - 1;           // mirrored lane inside the block
            const bool i_am_lo = (l & (lanes >> 1)) == 0;  // lower half of the block holds the lo slots
            int other[R];
#pragma unroll
            for (int r = 0; r < R; ++r) other[r] = __shfl_xor_sync(gmask, key[R - 1 - r], partner_xor);
#pragma unroll
            for (int r = 0; r < R; ++r) key[r] = cex_remote<MASK>(key[r], other[r], i_am_lo);
        }
        // ---- half-cleaners: pairs (x, x+d), d = blk/4 ... 1 ----
#pragma unroll
        for (int d = blk / 4; d >= 1; d >>= 1) {
            if (d < R) {
#pragma unroll
                for (int b = 0; b < R; b += 2 * d)
#pragma unroll
                    for (int i = 0; i < d; ++i) cex_local<MASK>(key[b + i], key[b + i + d]);
            } else {
                const int lx = d / R;
                const bool i_am_lo = (l & lx) == 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int other = __shfl_xor_sync(gmask, key[r], lx);
                    key[r] = cex_remote<MASK>(key[r], other, i_am_lo);
                }
            }
        }
    }
}

// sequential Python sum() of price*holding (lanes prepare the products, the leader adds them in order)
template <int R>
__device__ __forceinline__ double group_total_asset(OctEnv<8 * R> &e, double cash, const double *__restrict__ prow, int D, int l,
                                                    unsigned gmask)
{
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = l * R + r;
        if (j < D) e.prod[j] = dmul(__ldg(prow + j), (double)e.hold[j]);
    }
    __syncwarp(gmask);
    double acc = 0.0;
    if (l == 0) {
#pragma unroll 6
        for (int j = 0; j < D; ++j) acc = dadd(acc, e.prod[j]);
        acc = dadd(cash, acc);
    }
    __syncwarp(gmask);
    return acc;  // valid in the leader lane
}

// DCT > 0: the stock count compiled in (DOW-30: the two pad slots and every `j < D` guard fold away)
template <int R, typename ActT, int WARPS, int DCT>
__global__ void __launch_bounds__(WARPS * 32)
trading_small_kernel(const frl_trading_params p, const ActT *__restrict__ actions, long long act_step_stride,
                     long long act_env_stride, int n_steps, double *__restrict__ rewards, uint8_t *__restrict__ flags_out,
                     float *__restrict__ obs, int obs_mode, int auto_reset, double *__restrict__ stats)
{
    stats_exchange_previous(stats);
    using Env = OctEnv<8 * R>;
    constexpr int IB = KeyBits<R>::IB, IMASK = KeyBits<R>::MASK, AMAX = KeyBits<R>::AMAX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Env *smem = reinterpret_cast<Env *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, l = lane & 7;
    const unsigned gmask = 0xffu << (8 * g);
    Env &e = smem[warp * 4 + g];
    const int N = p.n_envs, D = DCT > 0 ? DCT : p.stock_dim, T = p.n_days, O = p.obs_dim;
    const long long n = ((long long)blockIdx.x * WARPS + warp) * 4 + g;
    if (n >= N) return;  // group-uniform: nothing below synchronises wider than the group

    // ---- state: scalars in every lane (only the leader's copy is authoritative), holdings in smem ----
    double cash = p.cash[n], cost = p.cost[n], last_reward = p.reward[n];
    int day = p.day[n], sday = p.sday[n], trades = p.trades[n];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = l * R + r;
        e.hold[j] = j < D ? p.hold[(size_t)j * p.env_stride + n] : 0;
    }
    __syncwarp(gmask);
    const double one_minus_sc = dsub(1.0, p.sell_cost_pct), one_plus_bc = dadd(1.0, p.buy_cost_pct);
    const int hmax_i = (int)max(-(double)AMAX, min((double)AMAX, p.hmax));
    double asset = 0.0;
    bool asset_ok = false;
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0, st_liq = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        const ActT *arow = actions + (size_t)k * act_step_stride + (size_t)n * act_env_stride;
        ActT av[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = l * R + r;
            av[r] = j < D ? arow[j] : ActT(0);
        }
        int flags = 0;
        double reward = 0.0;
        if (day >= T - 1) {
            // ---- terminal branch (:221-301): no state change, previous scaled reward again (Q3) ----
            flags = FRL_FLAG_DONE;
            reward = last_reward;
            if (!asset_ok) {
                asset = __shfl_sync(gmask, group_total_asset<R>(e, cash, p.close + (size_t)state_day(sday) * p.close_pitch, D, l, gmask), 8 * g);
                asset_ok = true;
            }
            st_done += 1.0;
            st_epi += asset;
            if (auto_reset) {  // DummyVecEnv.step_wait -> reset (:359-393), stale-day quirk Q1
                cash = p.initial_amount;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int j = l * R + r;
                    if (j < D) e.hold[j] = p.init_hold ? __ldg(p.init_hold + j) : 0;
                }
                __syncwarp(gmask);
                sday = -day - 1;
                day = 0;
                cost = 0.0;
                trades = 0;
                if (l == 0) p.episode[n] += 1;
                asset_ok = false;
            }
        } else {
            const int sd = state_day(sday);
            const double turb = sday < 0 ? 0.0 : __ldg(p.risk + sd);
            const bool liq = p.use_turbulence && (turb >= p.turbulence_threshold);
            const double *prow = p.close + (size_t)sd * p.close_pitch;
            if (!asset_ok) asset = __shfl_sync(gmask, group_total_asset<R>(e, cash, prow, D, l, gmask), 8 * g);
            const double begin = asset;

            // ---- keys in slot order; pads sort to the end ----
            int key[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = l * R + r;
                const int a = liq ? -hmax_i : max(-AMAX, min(AMAX, action_to_shares<ActT>(av[r], p.hmax)));
                key[r] = j < D ? (a << IB) + j : 0x7fffffff;
            }
            if (liq)
                flags = FRL_FLAG_LIQUIDATE;  // all keys tie: the network never swaps, order = index order
            else
                distributed_network<R>(key, l, gmask);
            const uint32_t *dis_row = p.disable_mask + (size_t)sd * (p.close_pitch >> 5);
            const bool use_dis = !liq && p.disable_mask != nullptr;

            // ---- everything that does not depend on the running cash, one sorted slot per (lane, r) ----
            int my_trades = 0, my_buys = 0;
            double b_thr[R], b_spend[R], b_cost[R], b_p[R];
            int b_aj[R], b_h[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int pos = l * R + r;
                double s_cash = 0.0, s_cost = 0.0;
                b_aj[r] = 0;
                b_thr[r] = b_spend[r] = b_cost[r] = b_p[r] = 0.0;
                b_h[r] = 0;
                if (pos < D) {
                    const int kk = key[r];
                    const int a = kk >> IB, j = kk & IMASK;
                    const double pj = __ldg(prow + j);
                    const int h = e.hold[j];
                    const bool disabled = use_dis && ((__ldg(dis_row + (j >> 5)) >> (j & 31)) & 1u);
                    if (a < 0) {
                        // _sell_stock (:102-169): liquidation checks price > 0, normal mode the disable flag
                        const bool ok = liq ? (pj > 0.0 && h > 0) : (!disabled && h > 0);
                        if (ok) {
                            const int m = liq ? h : min(-a, h);
                            const double pv = dmul(pj, (double)m);
                            s_cash = dmul(pv, one_minus_sc);
                            s_cost = dmul(pv, p.sell_cost_pct);
                            e.hold[j] = h - m;  // distinct stock per slot: no conflict
                            my_trades += 1;
                        }
                    } else if (a > 0 && !liq && !disabled) {
                        // _buy_stock (:171-201): the part that is independent of cash
                        const double unit = dmul(pj, one_plus_bc);
                        const double pv = dmul(pj, (double)a);
                        b_thr[r] = dmul((double)a + 1.0, unit);
                        b_spend[r] = dmul(pv, one_plus_bc);
                        b_cost[r] = dmul(pv, p.buy_cost_pct);
                        b_p[r] = pj;
                        b_aj[r] = kk;
                        b_h[r] = h;
                        my_buys += 1;
                    }
                }
                e.sell_cash[pos] = s_cash;
                e.sell_cost[pos] = s_cost;
            }
            // compaction: buys execute from the highest sorted position down, so an entry's slot in the list is
            // the number of enabled buys at higher positions (higher lanes, then higher registers of this lane)
            int above = 0;  // enabled buys in lanes > l
            {
                int incl = my_buys;  // inclusive suffix sum over the group's lanes
#pragma unroll
                for (int o = 1; o < kGroup; o <<= 1) {
                    const int t = __shfl_down_sync(gmask, incl, o, kGroup);
                    if (l + o < kGroup) incl += t;
                }
                above = incl - my_buys;
                const int total_buys = __shfl_sync(gmask, incl, 8 * g);
                int idx = above;
#pragma unroll
                for (int r = R - 1; r >= 0; --r) {
                    if (b_aj[r] != 0) {
                        e.buy_thr[idx] = b_thr[r];
                        e.buy_spend[idx] = b_spend[r];
                        e.buy_cost[idx] = b_cost[r];
                        e.buy_p[idx] = b_p[r];
                        e.buy_aj[idx] = b_aj[r];
                        e.buy_h[idx] = b_h[r];
                        ++idx;
                    }
                }
                my_trades += my_buys;  // every enabled buy attempt counts, even when 0 shares are bought (Q5)
                my_trades += __shfl_xor_sync(gmask, my_trades, 1);
                my_trades += __shfl_xor_sync(gmask, my_trades, 2);
                my_trades += __shfl_xor_sync(gmask, my_trades, 4);
                trades += my_trades;
                __syncwarp(gmask);

                // ---- the irreducible serial part, leader lane, reference order ----
                if (l == 0) {
#pragma unroll 6
                    for (int pos = 0; pos < D; ++pos) {  // sells, most negative first (adding 0.0 is the identity)
                        cash = dadd(cash, e.sell_cash[pos]);
                        cost = dadd(cost, e.sell_cost[pos]);
                    }
                    // buys, largest first, each limited by the cash left; the next entry is fetched while the
                    // current one resolves (the cash chain is the only true dependency)
                    double thr = e.buy_thr[0], spend = e.buy_spend[0], bc = e.buy_cost[0], pj = e.buy_p[0];
                    int kk = e.buy_aj[0], bh = e.buy_h[0];
                    for (int i = 0; i < total_buys; ++i) {
                        const int nx = min(i + 1, 8 * R - 1);
                        const double thr_n = e.buy_thr[nx], spend_n = e.buy_spend[nx], bc_n = e.buy_cost[nx], pj_n = e.buy_p[nx];
                        const int kk_n = e.buy_aj[nx], bh_n = e.buy_h[nx];
                        const int a = kk >> IB, j = kk & IMASK;
                        if (cash >= thr) {
                            cash = dsub(cash, spend);
                            cost = dadd(cost, bc);
                            e.hold[j] = bh + a;
                        } else {
                            const double unit = dmul(pj, one_plus_bc);
                            // not even one share affordable: cash // unit == 0 exactly and buying 0 shares is a
                            // no-op (the common state of a cash-starved env) — skip the division
                            if (!(cash >= 0.0 && cash < unit)) {
                                const double avail = floor_div_f64(cash, unit);
                                double nsh = (double)a;
                                nsh = (nsh < avail) ? nsh : avail;
                                const double pv = dmul(pj, nsh);
                                cash = dsub(cash, dmul(pv, one_plus_bc));
                                cost = dadd(cost, dmul(pv, p.buy_cost_pct));
                                e.hold[j] = bh + (int)nsh;
                            }
                        }
                        thr = thr_n; spend = spend_n; bc = bc_n; pj = pj_n; kk = kk_n; bh = bh_n;
                    }
                }
            }
            __syncwarp(gmask);
            cash = __shfl_sync(gmask, cash, 8 * g);
            cost = __shfl_sync(gmask, cost, 8 * g);

            // ---- state: s -> s+1 (:335-352) ----
            day += 1;
            sday = day;
            asset = __shfl_sync(gmask, group_total_asset<R>(e, cash, p.close + (size_t)day * p.close_pitch, D, l, gmask), 8 * g);
            asset_ok = true;
            reward = dmul(dsub(asset, begin), p.reward_scaling);
            last_reward = reward;
            if (liq) st_liq += 1.0;
        }
        if (l == 0) {
            if (rewards) rewards[(size_t)k * N + n] = reward;
            if (flags_out) flags_out[(size_t)k * N + n] = (uint8_t)flags;
            st_r += reward;
            st_r2 += reward * reward;
        }
        if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
            float *orow = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * O : (size_t)0) + (size_t)n * O;
            const float *trow = p.obs_tmpl + (size_t)state_day(sday) * O;
            for (int pos = l; pos < O; pos += kGroup) {
                float v = __ldg(trow + pos);
                if (pos == 0)
                    v = (float)cash;
                else if (pos > D && pos <= 2 * D)
                    v = (float)e.hold[pos - 1 - D];
                orow[pos] = v;
            }
        }
    }

    // ---- store state ----
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = l * R + r;
        if (j < D) p.hold[(size_t)j * p.env_stride + n] = e.hold[j];
    }
    if (p.asset_out || stats) {
        if (!asset_ok)
            asset = __shfl_sync(gmask, group_total_asset<R>(e, cash, p.close + (size_t)state_day(sday) * p.close_pitch, D, l, gmask), 8 * g);
    }
    if (l == 0) {
        p.cash[n] = cash;
        p.cost[n] = cost;
        p.reward[n] = last_reward;
        p.day[n] = day;
        p.sday[n] = sday;
        p.trades[n] = trades;
        if (p.asset_out) p.asset_out[n] = asset;
        if (stats) {
            // st_done / st_epi / st_liq were accumulated identically in every lane; only the leader reports
            const double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, asset, st_liq, (double)n_steps, (double)trades};
#pragma unroll
            for (int i = 0; i < FRL_N_STATS; ++i)
                if (v[i] != 0.0) atomicAdd(stats + i, v[i]);
        }
    }
}

template <int R, typename ActT, int DCT = 0>
void launch_small_r(const frl_trading_params &p, const void *actions, long long sstride, long long estride, int n_steps,
                    double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    constexpr int W = (R <= 4) ? FRL_SMALL_WARPS : 1;  // 4 envs x 2.2 KB (R = 4) ... 8.7 KB (R = 16) of scratch per warp
    const size_t smem = sizeof(OctEnv<8 * R>) * W * 4;
    static_assert(sizeof(OctEnv<8 * R>) * W * 4 <= 48 * 1024, "scratch must fit the default shared-memory limit");
    const long long groups = p.n_envs;
    const unsigned grid = (unsigned)((groups + W * 4 - 1) / (W * 4));
    trading_small_kernel<R, ActT, W, DCT><<<grid, W * 32, smem, st>>>(p, (const ActT *)actions, sstride, estride, n_steps, rewards,
                                                                flags, obs, obs_mode, auto_reset, stats);
}

}  // namespace

void launch_trading_small(const frl_trading_params &p, const void *actions, int actions_f64, long long sstride,
                          long long estride, int n_steps, double *rewards, uint8_t *flags, float *obs, int obs_mode,
                          int auto_reset, double *stats, cudaStream_t st)
{
    const int D = p.stock_dim;
#define FRL_GO(R)                                                                                                  \
    do {                                                                                                           \
        if (actions_f64)                                                                                           \
            launch_small_r<R, double>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st); \
        else                                                                                                       \
            launch_small_r<R, float>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st); \
    } while (0)
    if (D <= 8)
        FRL_GO(1);  // numpy pads to next_pow2(max(D, 8)) slots
    else if (D <= 16)
        FRL_GO(2);
    else if (D == 30 && !actions_f64 && FRL_SMALL_DCT30)  // DOW-30, float32 actions: stock count compiled in
        launch_small_r<4, float, 30>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st);
    else if (D <= 32)
        FRL_GO(4);
    else if (D <= 64)
        FRL_GO(8);
    else
        FRL_GO(16);
#undef FRL_GO
}

}  // namespace frl
