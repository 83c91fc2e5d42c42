// A4 — StockTradingEnvCashpenalty (reference: finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py).
//
// Mapping: ONE THREAD PER ENV, one warp per 32-env tile (the layout of trading.cu).  A first version ran
// one warp per env (lanes = assets, shuffle-reduced dot products): every scalar decision was replicated
// in 32 lanes and it issued ~650-800 warp-instructions per env-step (43-47 % of the HBM roofline).  Per
// thread the same work is ONE streaming pass over the D assets in the common case:
//   pass    read holding (stock-major, coalesced), closing price (warp-uniform, L1), staged action; form
//           the transaction; accumulate np.dot(holdings, closings), proceeds, spend, sum|a|; write the
//           TENTATIVE new holding into the env's other holdings buffer (ping-pong) and its float32 image
//           into the lane's own staging row, which the observation writer later reads row-wise — the
//           action staging buffer doubles as the transposition buffer.
//   decide  (scalar, per lane) normal: flip the env's current-buffer bit.  CASH SHORTAGE + terminate: do
//           not flip — the state is unchanged by construction.  CASH SHORTAGE + patient: a cheap fix-up
//           pass undoes the buys (new > old => keep old), then flip.
// Holdings therefore live in two stock-major buffers and bit 1 of the per-env `fresh` byte says which one
// is current; reads + writes per step are the same 2 x 8D bytes as an in-place update.
// np.dot's order is BLAS-specific (tolerance 1e-9 in the tests); the sums here are sequential in asset
// order, which happens to be the CPU oracle's order as well.
#include "common.cuh"
#include "cp_obs.cuh"

#ifndef FRL_CP_MIN_BLOCKS
#define FRL_CP_MIN_BLOCKS 4  // 128-thread blocks per SM the register allocator must allow
#endif

#ifndef FRL_CP_DCT100
#define FRL_CP_DCT100 1  // instantiation with the stock count compiled in for D = 100 (A/B switch)
#endif

namespace frl {
namespace {

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

// get_transactions (:258-298) for one asset: actions*hmax in the input dtype, zero where the price is
// not positive, shares = currency / price (or the discretised floor), never sell more than held,
// turbulence clears the position.
// HVEC: hmax is the per-asset array (its own instantiation, so the scalar path carries no checks)
template <typename ActT, bool HVEC>
__device__ __forceinline__ double cp_transaction(const frl_cashpenalty_params &p, ActT a, double hmax, double c, double h,
                                                 bool liq)
{
    double v;
    if (sizeof(ActT) == 4 && !(HVEC && !p.hmax_vec_f32))
        v = (double)fmul((float)a, (float)hmax);  // scalar (weak Python float) or float32 array: float32 product
    else
        v = dmul((double)a, hmax);
    if (!(c > 0.0)) v = 0.0;  // np.where(closings > 0, actions, 0)
    if (p.discrete_actions) {
        long long q = (long long)floor_div_f64(v, c);  // actions // closings, astype(int)
        const long long inc = p.shares_increment;
        q = (q >= 0) ? floordiv_ll(q, inc) * inc : floordiv_ll(q + inc, inc) * inc;
        v = (double)q;
    } else {
        v = __ddiv_rn(v, c);
    }
    v = (v > -h) ? v : -h;  // np.maximum(actions, -holdings)
    if (liq) v = -h;        // turbulence: clear out all positions
    return v;
}

#ifndef FRL_CP_ASYNC_STAGE
#define FRL_CP_ASYNC_STAGE 1  // stage the actions with cp.async instead of load + store batches
#endif
#ifndef FRL_CP_PF
#define FRL_CP_PF 2  // bulk-staged variant: batches of four holdings in flight ahead of the one being traded
#endif
#ifndef FRL_CP_BULK_STAGE
#define FRL_CP_BULK_STAGE 1  // A/B switch for the bulk-staged variant
#endif
#ifndef FRL_CP_U
#define FRL_CP_U 4  // assets per software-pipelined batch of the pass (A/B on B200: 2 -> 0.290 ms, 3 -> 0.255, 4..6 -> 0.240, 8 -> 0.250)
#endif
constexpr int CP_U = FRL_CP_U;

__device__ __forceinline__ void cp_async_elem(float *dst, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(double *dst, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// BULK (decided at launch): float actions in the default [N][D] layout, D a multiple of four, 16-byte aligned tiles,
// continuous shares and the (close, 1/close) table.  The tile's 32 x D actions are ONE contiguous run that the copy
// engine (TMA) stages as flat [32][D] rows with a single instruction — instead of D cp.async per lane and their index
// arithmetic (20 % of the kernel's instructions) — the pass handles four assets per 128-bit shared-memory word, and
// `actions / closings` is the three-instruction exact division by the tabulated reciprocal.
template <typename ActT, int WARPS, bool HVEC, bool BULK, int DCT = 0>
__global__ void __launch_bounds__(WARPS * 32, FRL_CP_MIN_BLOCKS * 128 / (WARPS * 32))
cashpenalty_rollout_kernel(const frl_cashpenalty_params p, const ActT *__restrict__ actions, long long act_step_stride,
                           long long act_env_stride, int n_steps, double *__restrict__ rewards,
                           uint8_t *__restrict__ flags_out, float *__restrict__ obs, int obs_mode, int auto_reset,
                           double *__restrict__ stats)
{
    stats_exchange_previous(stats);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.n_envs, D = DCT > 0 ? DCT : p.stock_dim, T = p.n_days, ld = p.env_stride;  // DCT: stock count compiled in
    // row pitch: odd (conflict-free per-lane row walks with scalar accesses), or D itself in the bulk-staged variant
    // (flat rows; its 128-bit accesses are conflict-free for any pitch that is a multiple of four)
    const int P = BULK ? D : (D | 1);
    const size_t warp_bytes = (size_t)32 * (D | 1) * sizeof(ActT) + 32 * sizeof(float) + 32 * sizeof(int) + 16;
    unsigned char *base = smem_raw + warp * ((warp_bytes + 15) & ~(size_t)15);
    ActT *stage = reinterpret_cast<ActT *>(base);                       // [32 envs][P]
    float *cashf = reinterpret_cast<float *>(base + (size_t)32 * (D | 1) * sizeof(ActT));
    int *di_s = reinterpret_cast<int *>(cashf + 32);
    const unsigned mbar = smem_u32(di_s + 32);  // staging mbarrier (BULK); 8-byte aligned: all pieces before it are 128 B multiples

    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;

    double cash = p.cash[n], last_cash = p.last_cash[n], last_total = p.last_total[n], sum_trades = p.sum_trades[n];
    int di = p.date_index[n], start = p.start[n];
    const int bits0 = p.fresh[n];
    bool fresh = (bits0 & 1) != 0;
    bool cur = (bits0 & 2) != 0;  // which holdings buffer is current: hold (0) or hold_alt (1)
    ActT *myrow = stage + (size_t)lane * P;
    static_assert(!BULK || (sizeof(ActT) == 4 && CP_U == 4), "the bulk-staged variant reads float4 batches");
    unsigned stage_phase = 0;
    if (BULK && lane == 0) mbar_init(mbar, 1);
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0, st_liq = 0.0, st_short = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        // ---- stage this step's actions (coalesced) into [env][P] rows ----
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        const bool tma = BULK && nvalid == 32;  // warp-uniform; the ragged last tile is copied element-wise
        if (tma) {
            fence_proxy_async_smem();  // the rows were read / written through the generic proxy during the last step
            __syncwarp();
            if (lane == 0) {
                const unsigned bytes = 32u * (unsigned)D * (unsigned)sizeof(ActT);
                mbar_expect_tx(mbar, bytes);
                bulk_copy_g2s(smem_u32(stage), abase + (size_t)env0 * D, bytes, mbar);
            }
        } else if (BULK) {
            const ActT *tile = abase + (size_t)env0 * D;
            for (int e = lane; e < 32 * D; e += 32) stage[e] = e < nvalid * D ? tile[e] : ActT(0);
        } else if (act_env_stride == D) {
            const ActT *tile = abase + (size_t)env0 * D;
            const int cnt = nvalid * D;
            int row = 0, col = lane;
            while (col >= D) { col -= D; ++row; }
#if FRL_CP_ASYNC_STAGE
            // every element goes global -> shared with cp.async (no registers, the whole tile in flight at once);
            // the wait sits after the first holdings batch has been requested
            for (int e = lane; e < 32 * D; e += 32) {
                ActT *dst = stage + row * P + col;
                if (e < cnt)
                    cp_async_elem(dst, tile + e);
                else
                    *dst = ActT(0);
                col += 32;
                while (col >= D) { col -= D; ++row; }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
#else
            // batches of 8 independent coalesced loads in flight per lane, then parked row-wise
            for (int e0 = lane; e0 < 32 * D; e0 += 32 * 8) {
                ActT v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = e0 + 32 * u;
                    v[u] = e < cnt ? __ldcs(tile + e) : ActT(0);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (e0 + 32 * u < 32 * D) stage[row * P + col] = v[u];
                    col += 32;
                    while (col >= D) { col -= D; ++row; }
                }
            }
#endif
        } else {
            for (int r = 0; r < 32; ++r)
                for (int j = lane; j < D; j += 32)
                    stage[r * P + j] = r < nvalid ? abase[(size_t)(env0 + r) * act_env_stride + j] : ActT(0);
        }
        // first batch of holdings / closes of the pass, requested before the staged actions are waited for
        const double *hq = (cur ? p.hold_alt : p.hold) + n;  // hold[j][n] at hq[j * ld]
        const double *crow = p.close + (size_t)di * D;
        double hb[CP_U], cb[CP_U];
#pragma unroll
        for (int u = 0; u < CP_U; ++u) {
            hb[u] = u < D ? __ldcg(hq) : 0.0;
            cb[u] = u < D ? __ldg(crow + u) : 1.0;
            hq += ld;
        }
        // bulk-staged variant: FRL_CP_PF more batches right away, each in its OWN registers.  The holdings stream is
        // the DRAM-latency critical path of the pass (51 % of the stall samples were long-scoreboard waits with one
        // batch in flight).  The pass below is unrolled over the FRL_CP_PF + 1 register sets so that a set is simply
        // reloaded after its batch has been traded — rotating the sets with moves would wait for the newest loads.
        double hset[BULK ? FRL_CP_PF : 1][4];
        if constexpr (BULK) {
#pragma unroll
            for (int b = 0; b < FRL_CP_PF; ++b) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    hset[b][u] = 4 * (b + 1) + u < D ? __ldcg(hq) : 0.0;
                    hq += ld;
                }
            }
        }
        if (tma) {
            mbar_wait(mbar, stage_phase);
            stage_phase ^= 1u;
        } else if (!BULK) {
#if FRL_CP_ASYNC_STAGE
            asm volatile("cp.async.wait_all;" ::: "memory");
#endif
        }
        __syncwarp();

        int flags = 0;
        double reward;
        const int current_step = di - start;
        bool reset_now = false, moved = false;
        if (di == T - 1) {
            // last date (:308-310): reward from the previously logged (assets, cash); state unchanged
            double asum = 0.0;
            for (int j = 0; j < D; ++j) asum += fabs((double)myrow[j]);
            sum_trades += asum;
            flags = FRL_FLAG_DONE;
            reward = cp_reward(p, last_total, last_cash, current_step);
            reset_now = auto_reset != 0;
        } else {
            const double turbulence = fresh ? 0.0 : __ldg(p.turb + di);
            const bool liq = p.use_turbulence && turbulence >= p.turbulence_threshold;
            if (liq) flags |= FRL_FLAG_LIQUIDATE;
            // ---- the pass: np.sum(|actions|), np.dot(holdings, closings), proceeds, spend, tentative holdings ----
            double asum = 0.0, asset_value = 0.0, proceeds = 0.0, spend = 0.0;
            double *hw = (cur ? p.hold : p.hold_alt) + n;
            if constexpr (BULK) {
                const double2 *crc = reinterpret_cast<const double2 *>(p.close_rc) + (size_t)di * D;
                // one batch of four assets out of register set h4; afterwards the set is reloaded with the batch
                // FRL_CP_PF + 1 further on (hq already points there)
                auto trade4 = [&](double (&h4)[4], int j0) {
                    const float4 a4 = *reinterpret_cast<const float4 *>(myrow + j0);
                    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                    float hf[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double2 cr = __ldg(crc + j0 + u);  // (close, RN(1 / close)): warp-uniform, L1-resident
                        const double c = cr.x, h = h4[u];
                        double v = HVEC ? (p.hmax_vec_f32 ? (double)fmul(av[u], (float)__ldg(p.hmax_vec + j0 + u))
                                                          : dmul((double)av[u], __ldg(p.hmax_vec + j0 + u)))
                                        : (double)fmul(av[u], (float)p.hmax);
                        if (!(c > 0.0)) v = 0.0;  // np.where(closings > 0, actions, 0)
                        {  // v / c, correctly rounded (Markstein); 0 * (1/0 = inf) = nan like numpy's 0 / 0
                            const double q = dmul(v, cr.y);
                            const double e = __fma_rn(-c, q, v);
                            v = __fma_rn(e, cr.y, q);
                        }
                        v = (v > -h) ? v : -h;  // np.maximum(actions, -holdings)
                        if (liq) v = -h;        // turbulence: clear out all positions
                        asum += fabs((double)av[u]);
                        asset_value = dadd(asset_value, dmul(h, c));
                        proceeds = dadd(proceeds, dmul(v < 0.0 ? -v : 0.0, c));
                        spend = dadd(spend, dmul(v > 0.0 ? v : 0.0, c));
                        const double hn = dadd(h, v);  // holdings_updated = holdings + transactions (:361)
                        if (valid) *hw = hn;
                        hf[u] = (float)hn;
                        hw += ld;
                    }
                    *reinterpret_cast<float4 *>(myrow + j0) = make_float4(hf[0], hf[1], hf[2], hf[3]);  // observation image
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        h4[u] = j0 + 4 * (FRL_CP_PF + 1) + u < D ? __ldcg(hq) : 0.0;
                        hq += ld;
                    }
                };
                for (int j0 = 0; j0 < D; j0 += 4 * (FRL_CP_PF + 1)) {  // D is a multiple of four
                    trade4(hb, j0);
#pragma unroll
                    for (int b = 0; b < FRL_CP_PF; ++b)
                        if (j0 + 4 * (b + 1) < D) trade4(hset[b], j0 + 4 * (b + 1));
                }
            } else
            for (int j0 = 0; j0 < D; j0 += CP_U) {
                // software pipeline: the next batch of CP_U independent holding loads is in flight while this
                // one is traded (the holdings stream is the DRAM-latency critical path of the pass)
                double hn_[CP_U], cn_[CP_U];
#pragma unroll
                for (int u = 0; u < CP_U; ++u) {
                    const int j = j0 + CP_U + u;
                    hn_[u] = j < D ? __ldcg(hq) : 0.0;
                    cn_[u] = j < D ? __ldg(crow + j) : 1.0;
                    hq += ld;
                }
#pragma unroll
                for (int u = 0; u < CP_U; ++u) {
                    const int j = j0 + u;
                    if (j < D) {
                        const ActT a = myrow[j];
                        const double v = cp_transaction<ActT, HVEC>(p, a, HVEC ? __ldg(p.hmax_vec + j) : p.hmax, cb[u], hb[u], liq);
                        asum += fabs((double)a);
                        asset_value = dadd(asset_value, dmul(hb[u], cb[u]));
                        proceeds = dadd(proceeds, dmul(v < 0.0 ? -v : 0.0, cb[u]));
                        spend = dadd(spend, dmul(v > 0.0 ? v : 0.0, cb[u]));
                        const double hn = dadd(hb[u], v);  // holdings_updated = holdings + transactions (:361)
                        if (valid) *hw = hn;
                        *reinterpret_cast<float *>(myrow + j) = (float)hn;  // observation image, lane-private slot
                    }
                    hw += ld;
                }
#pragma unroll
                for (int u = 0; u < CP_U; ++u) {
                    hb[u] = hn_[u];
                    cb[u] = cn_[u];
                }
            }
            sum_trades += asum;  // (:302), logging only
            const double begin_cash = cash;
            last_cash = begin_cash;
            last_total = dadd(begin_cash, asset_value);
            reward = cp_reward(p, last_total, last_cash, current_step);  // computed BEFORE trading (:326)
            double costs = dmul(proceeds, p.sell_cost_pct);
            double coh = dadd(begin_cash, proceeds);
            costs = dadd(costs, dmul(spend, p.buy_cost_pct));
            bool terminate = false, no_buys = false;
            if (dadd(spend, costs) > coh) {
                flags |= FRL_FLAG_SHORTAGE;
                if (p.patient) {  // no buys until there is cash again; the sell costs are dropped too (Q9)
                    no_buys = true;
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
                if (no_buys) {
                    // patient: transactions = where(transactions > 0, 0, transactions) (:346) — undo the buys
                    // of the tentative update (a buy is exactly a slot whose new holding exceeds the old one)
                    const double *ho = (cur ? p.hold_alt : p.hold) + n;
                    double *hn_p = (cur ? p.hold : p.hold_alt) + n;
                    for (int j = 0; j < D; ++j) {
                        const double h = __ldcg(ho + (size_t)j * ld), hn = __ldcg(hn_p + (size_t)j * ld);
                        if (hn > h) {
                            if (valid) hn_p[(size_t)j * ld] = h;
                            *reinterpret_cast<float *>(myrow + j) = (float)h;
                        }
                    }
                }
                cur = !cur;  // the tentative buffer becomes the state
                moved = true;
                di += 1;
                if (p.use_turbulence) fresh = false;  // self.turbulence is refreshed only with a threshold
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
        if (reset_now) {  // DummyVecEnv.step_wait -> reset (:132-158); starting point 0 or, with p.random_start, randint(0, int(T * 0.5))
            cash = p.initial_amount;
            double *hz = (cur ? p.hold_alt : p.hold) + n;
            for (int j = 0; j < D; ++j) {
                if (valid) hz[(size_t)j * ld] = 0.0;
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
            if (!moved) {  // terminal / terminated without reset: image of the unchanged holdings
                const double *hc = (cur ? p.hold_alt : p.hold) + n;
                for (int j = 0; j < D; ++j) *reinterpret_cast<float *>(myrow + j) = (float)__ldcg(hc + (size_t)j * ld);
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
        p.fresh[n] = (uint8_t)((fresh ? 1 : 0) | (cur ? 2 : 0));
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

__global__ void cashpenalty_reset_kernel(const frl_cashpenalty_params p, const uint8_t *__restrict__ mask,
                                         const int32_t *__restrict__ start_points)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.n_envs) return;
    if (mask && !mask[n]) return;
    for (int j = 0; j < p.stock_dim; ++j) p.hold[(size_t)j * p.env_stride + n] = 0.0;  // buffer 0 becomes current
    const int sp = start_points ? start_points[n] : 0;
    p.cash[n] = p.initial_amount;
    p.date_index[n] = sp;
    p.start[n] = sp;
    p.fresh[n] = 1;
    p.last_cash[n] = 0.0;
    p.last_total[n] = 0.0;
    p.sum_trades[n] = 0.0;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) cashpenalty_observe_kernel(const frl_cashpenalty_params p, float *__restrict__ obs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.n_envs, D = p.stock_dim;
    const int P = D | 1;
    const size_t warp_bytes = (size_t)32 * P * sizeof(float) + 32 * sizeof(float) + 32 * sizeof(int);
    unsigned char *base = smem_raw + warp * ((warp_bytes + 15) & ~(size_t)15);
    float *stage = reinterpret_cast<float *>(base);
    float *cashf = stage + (size_t)32 * P;
    int *di_s = reinterpret_cast<int *>(cashf + 32);
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const long long n = lane < nvalid ? env0 + lane : (long long)N - 1;
    const double *hc = ((p.fresh[n] & 2) ? p.hold_alt : p.hold) + n;
    for (int j = 0; j < D; ++j) stage[(size_t)lane * P + j] = (float)hc[(size_t)j * p.env_stride];
    cashf[lane] = (float)p.cash[n];
    di_s[lane] = p.date_index[n];
    __syncwarp();
    cp_write_obs_tile<float>(p, stage, P, cashf, di_s, obs, env0, nvalid, lane, D);
}

int32_t cp_validate(const frl_cashpenalty_params *p)
{
    FRL_REQUIRE(p != nullptr, "cashpenalty: params is NULL");
    FRL_REQUIRE(p->n_envs >= 1, "cashpenalty: n_envs must be >= 1 (got %d)", p->n_envs);
    FRL_REQUIRE(p->stock_dim >= 1 && p->stock_dim <= 128, "cashpenalty: stock_dim must be in 1..128 (got %d)", p->stock_dim);
    FRL_REQUIRE(p->n_cols >= 0 && p->n_days >= 1, "cashpenalty: bad n_cols/n_days (%d, %d)", p->n_cols, p->n_days);
    FRL_REQUIRE(p->obs_dim == 1 + p->stock_dim + p->stock_dim * p->n_cols, "cashpenalty: obs_dim %d != 1 + D + D*C = %d",
                p->obs_dim, 1 + p->stock_dim + p->stock_dim * p->n_cols);
    FRL_REQUIRE(p->env_stride >= p->n_envs, "cashpenalty: env_stride %d < n_envs %d", p->env_stride, p->n_envs);
    FRL_REQUIRE(!p->discrete_actions || p->shares_increment >= 1, "cashpenalty: shares_increment must be >= 1");
    FRL_REQUIRE(p->close && p->obs_tmpl && (!p->use_turbulence || p->turb), "cashpenalty: table pointer is NULL");
    FRL_REQUIRE(p->cash && p->hold && p->hold_alt && p->date_index && p->start && p->fresh && p->last_cash && p->last_total &&
                    p->sum_trades,
                "cashpenalty: state pointer is NULL");
    return FRL_OK;
}

size_t cp_smem_bytes(int D, size_t elem, int warps)
{
    const int P = D | 1;
    const size_t warp_bytes = (size_t)32 * P * elem + 32 * sizeof(float) + 32 * sizeof(int) + 16;  // + the staging mbarrier
    return warps * ((warp_bytes + 15) & ~(size_t)15);
}

template <typename ActT, int WARPS>
int32_t cp_launch(const frl_cashpenalty_params &p, const void *actions, long long sstride, long long estride, int n_steps,
                  double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    const size_t smem = cp_smem_bytes(p.stock_dim, sizeof(ActT), WARPS);
    // bulk-staged variant: float actions, default layout, D % 4 == 0, every tile's first action 16-byte aligned,
    // continuous shares, reciprocal table present
    constexpr bool kCanBulk = FRL_CP_BULK_STAGE && sizeof(ActT) == 4 && CP_U == 4;
    const bool bulk = kCanBulk && estride == p.stock_dim && (p.stock_dim & 3) == 0 && (sstride & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(actions) & 15) == 0 && p.close_rc != nullptr && !p.discrete_actions;
    auto kern = bulk ? (p.hmax_vec ? cashpenalty_rollout_kernel<ActT, WARPS, true, kCanBulk> : cashpenalty_rollout_kernel<ActT, WARPS, false, kCanBulk>)
                     : (p.hmax_vec ? cashpenalty_rollout_kernel<ActT, WARPS, true, false> : cashpenalty_rollout_kernel<ActT, WARPS, false, false>);
    if constexpr (kCanBulk && FRL_CP_DCT100) {  // NASDAQ-100: stock count compiled in (scalar hmax)
        if (bulk && !p.hmax_vec && p.stock_dim == 100) kern = cashpenalty_rollout_kernel<ActT, WARPS, false, kCanBulk, 100>;
    }
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cashpenalty_rollout: cannot reserve %zu B of shared memory (%s)", smem, cudaGetErrorString(e));
            return FRL_E_CUDA;
        }
    }
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    const unsigned grid = (unsigned)((tiles + WARPS - 1) / WARPS);
    kern<<<grid, WARPS * 32, smem, st>>>(p, (const ActT *)actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode,
                                         auto_reset, stats);
    return check_launch("cashpenalty_rollout");
}

}  // namespace
}  // namespace frl

using namespace frl;

extern "C" int32_t frl_cashpenalty_observe(const frl_cashpenalty_params *p, float *obs, void *stream)
{
    if (int32_t rc = cp_validate(p)) return rc;
    FRL_REQUIRE(obs != nullptr, "cashpenalty_observe: obs is NULL");
    constexpr int W = 2;
    const long long tiles = ((long long)p->n_envs + 31) / 32;
    const size_t smem = cp_smem_bytes(p->stock_dim, sizeof(float), W);
    cashpenalty_observe_kernel<W><<<(unsigned)((tiles + W - 1) / W), W * 32, smem, (cudaStream_t)stream>>>(*p, obs);
    return check_launch("cashpenalty_observe");
}

extern "C" int32_t frl_cashpenalty_reset(const frl_cashpenalty_params *p, const uint8_t *mask, const int32_t *start_points,
                                         float *obs, void *stream)
{
    if (int32_t rc = cp_validate(p)) return rc;
    cashpenalty_reset_kernel<<<(p->n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*p, mask, start_points);
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
    cudaStream_t st = (cudaStream_t)stream;
    if (actions_f64)
        return cp_launch<double, 2>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags, obs, obs_mode,
                                    auto_reset, stats, st);
    return cp_launch<float, 4>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags, obs, obs_mode,
                               auto_reset, stats, st);
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
