// A2 — numpy / ElegantRL StockTradingEnv (reference: finrl/meta/env_stock_trading/env_stocktrading_np.py).
//
// Same mapping as trading.cu: one thread per env for the arithmetic, one warp per 32-env tile for
// coalesced I/O.  There is no sort here (sells then buys run in ascending stock index, :112,:120), so
// stocks / cool-down counters live in shared memory with static indices and registers stay light.
//
// Bit-exactness: the reference silently mixes Python floats, np.float32 and np.float64 (NEP 50); the
// dtype of `amount`, `total_asset`, `gamma_reward` and `reward` depends on which `min()` branch ran
// (SURVEY.md H3).  Each of those is carried as (double value, 2-bit kind) and every operation is done
// in the precision numpy would pick, with explicit round-to-nearest intrinsics (no FMA contraction).
// (stocks*price).sum() is numpy's pairwise float32 summation.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "np_common.cuh"

#ifndef FRL_NP_MIN_BLOCKS
#define FRL_NP_MIN_BLOCKS 8  // 64-thread units per SM the register allocator must allow (8 -> 128 regs)
#endif
#ifndef FRL_NP_WARPS
#define FRL_NP_WARPS 4
#endif
#ifndef FRL_NP_A64
#define FRL_NP_A64 1  // steady-state instantiation of the step body with every NEP-50 kind fixed to float64 (A/B switch)
#endif

namespace frl {
namespace {

constexpr int kPitch = 33;

template <int SLOTS, typename ActT>
struct alignas(16) NpWarpSmem {
    // the two big buffers are never live together: `act` holds the step's staged actions until the trade
    // loops have consumed them, `sc` is the transposition buffer of the observation writer afterwards
    union {
        ActT act[32 * SLOTS];          // staged actions, flat [32 envs][D]
        float sc[2 * SLOTS * kPitch];  // rows 0..D-1: stocks[j][lane]; rows D..2D-1: cool-down[j][lane]
    };
    float amountf[32];
    int day[32];
};

// (self.stocks * price).sum(): float32 products, numpy pairwise summation (n < 8 sequential; else 8
// accumulators over the full blocks of 8, tree-combined, then the tail sequentially).
template <int SLOTS>
__device__ __forceinline__ float np_asset_f32(const float (&stv)[SLOTS], const float *__restrict__ prow, int D)
{
    float x[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) x[j] = (j < D) ? fmul(stv[j], __ldg(prow + j)) : 0.0f;
    if (D < 8) {
        float res = 0.0f;
#pragma unroll
        for (int j = 0; j < 8 && j < SLOTS; ++j)
            if (j < D) res = fadd(res, x[j]);
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = x[j];
    const int nb = D >> 3;
#pragma unroll
    for (int b = 1; b < SLOTS / 8; ++b) {
        if (b < nb) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = fadd(r[j], x[8 * b + j]);
        }
    }
    float res = fadd(fadd(fadd(r[0], r[1]), fadd(r[2], r[3])), fadd(fadd(r[4], r[5]), fadd(r[6], r[7])));
#pragma unroll
    for (int j = 8; j < SLOTS; ++j)
        if (j >= 8 * nb && j < D) res = fadd(res, x[j]);
    return res;
}

// ---- observation rows: [amount, turb, turb_bool, price*2^-6 x D, stocks*2^-6 x D, cool x D, tech] ----
template <int NCH, int DCT, typename SM>
__device__ __forceinline__ void np_write_obs_rows_uniform(const frl_np_params &p, SM &sm, float *__restrict__ obs,
                                                          long long env0, int nvalid, int lane, int day0)
{
    const int O = p.obs_dim, D = DCT > 0 ? DCT : p.stock_dim;
    const int s_beg = 3 + D, c_beg = 3 + 2 * D, sp_end = 3 + 3 * D;
    float t[NCH];
    const float *trow = p.obs_tmpl + (size_t)day0 * O + lane;
#pragma unroll
    for (int c = 0; c < NCH; ++c) t[c] = (c < NCH - 1 || lane + 32 * c < O) ? __ldg(trow + 32 * c) : 0.0f;
    constexpr int NSP = NCH < 4 ? NCH : 4;  // 3 + 3D <= 99 -> at most 4 chunks touch per-env slots
    int soff[NSP];
    float mul[NSP];
#pragma unroll
    for (int c = 0; c < NSP; ++c) {
        const int pos = lane + 32 * c;
        soff[c] = -1;
        mul[c] = 1.0f;
        if (pos >= s_beg && pos < c_beg) {
            soff[c] = (pos - s_beg) * kPitch;
            mul[c] = 0.015625f;  // 2**-6
        } else if (pos >= c_beg && pos < sp_end) {
            soff[c] = (D + pos - c_beg) * kPitch;
        }
    }
    const bool tail_ok = lane + 32 * (NCH - 1) < O;
    float *orow = obs + (size_t)env0 * O + lane;
#pragma unroll 2
    for (int r = 0; r < nvalid; ++r) {
        float v[NSP];
#pragma unroll
        for (int c = 0; c < NSP; ++c) v[c] = (soff[c] >= 0) ? fmul(sm.sc[soff[c] + r], mul[c]) : t[c];
        const float am = sm.amountf[r];
        if (lane == 0) v[0] = am;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float x = c < NSP ? v[c] : t[c];
            if (c < NCH - 1 || tail_ok) obs_store<kStoreCG>(orow + 32 * c, x);
        }
        orow += O;
    }
}

// Compile-time stock count (the DOW-30 instantiation): the env-specific slots of a row, positions 3+D .. 3+3D-1,
// fall into the 32-float chunks CLO..CHI.  The kernel leaves exactly those chunks FINISHED in the transposition
// buffer — stocks already scaled, the few template positions that share a chunk filled in per env — as
// sc[pos - 32*CLO][env], so a row costs one shared-memory load per env chunk (static offsets, no select, no
// multiply), one for the amount, and its NCH stores.
template <int DCT>
struct NpObsCT {
    static constexpr int CLO = (3 + DCT) / 32, CHI = (3 + 3 * DCT - 1) / 32, NE = CHI - CLO + 1;
    static_assert(DCT <= 0 || CLO >= 1, "position 0 (amount) must not share a chunk with the per-env slots");
};

template <int NCH, int DCT, typename SM>
__device__ __forceinline__ void np_write_obs_rows_ct(const frl_np_params &p, SM &sm, float *__restrict__ obs,
                                                     long long env0, int nvalid, int lane, int day0)
{
    using L = NpObsCT<DCT>;
    const int O = p.obs_dim;
    float t[NCH];
    const float *trow = p.obs_tmpl + (size_t)day0 * O + lane;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
        t[c] = (c >= L::CLO && c <= L::CHI) ? 0.0f : ((c < NCH - 1 || lane + 32 * c < O) ? __ldg(trow + 32 * c) : 0.0f);
    const bool tail_ok = lane + 32 * (NCH - 1) < O;
    const float *col = sm.sc + lane * kPitch;
    float *orow = obs + (size_t)env0 * O + lane;
#pragma unroll 4
    for (int r = 0; r < nvalid; ++r) {
        const float am = sm.amountf[r];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            float x = t[c];
            if (c >= L::CLO && c <= L::CHI) x = col[(c - L::CLO) * 32 * kPitch + r];
            if (c == 0 && lane == 0) x = am;
            if (c < NCH - 1 || tail_ok) obs_store<kStoreCG>(orow + 32 * c, x);
        }
        orow += O;
    }
}

template <int DCT, typename SM>
__device__ __forceinline__ void np_write_obs_tile(const frl_np_params &p, SM &sm, float *__restrict__ obs,
                                                  long long env0, int nvalid, int lane)
{
    const int O = p.obs_dim, D = DCT > 0 ? DCT : p.stock_dim;
    const int day0 = sm.day[0];
    bool uniform = true;
    if (lane < nvalid) uniform = (sm.day[lane] == day0);
    uniform = __all_sync(0xffffffffu, uniform);
    const int nch = (O + 31) >> 5;
    if (uniform && nch <= 12) {
        switch (nch) {
#define FRL_CASE(N)                                                                                \
    case N:                                                                                        \
        if (DCT > 0)                                                                               \
            np_write_obs_rows_ct<N, (DCT > 0 ? DCT : 30)>(p, sm, obs, env0, nvalid, lane, day0);          \
        else                                                                                       \
            np_write_obs_rows_uniform<N, DCT>(p, sm, obs, env0, nvalid, lane, day0);                    \
        break;
            FRL_CASE(1) FRL_CASE(2) FRL_CASE(3) FRL_CASE(4) FRL_CASE(5) FRL_CASE(6)
            FRL_CASE(7) FRL_CASE(8) FRL_CASE(9) FRL_CASE(10) FRL_CASE(11) FRL_CASE(12)
#undef FRL_CASE
        }
    } else {
        const int s_beg = 3 + D, c_beg = 3 + 2 * D, sp_end = 3 + 3 * D;
        for (int r = 0; r < nvalid; ++r) {
            float *orow = obs + (size_t)(env0 + r) * O;
            const float *trow = p.obs_tmpl + (size_t)sm.day[r] * O;
            for (int pos = lane; pos < O; pos += 32) {
                float v = __ldg(trow + pos);
                if (pos == 0)
                    v = sm.amountf[r];
                else if (DCT > 0) {  // finished chunks, see np_write_obs_rows_ct
                    if (pos >= s_beg && pos < sp_end) v = sm.sc[(pos - 32 * NpObsCT<DCT>::CLO) * kPitch + r];
                } else if (pos >= s_beg && pos < c_beg)
                    v = fmul(sm.sc[(pos - s_beg) * kPitch + r], 0.015625f);
                else if (pos >= c_beg && pos < sp_end)
                    v = sm.sc[(D + pos - c_beg) * kPitch + r];
                orow[pos] = v;
            }
        }
    }
}

// reset (:80-101) for the calling thread's env (state in registers): the deterministic branch, or with
// p.train_reset the if_train branch — stocks = initial_stocks + randint(0, 64, D), amount = initial_capital *
// uniform(0.95, 1.05) - (stocks * price).sum() [python float - float32 -> float32] — with counter-based draws
template <int SLOTS>
__device__ __forceinline__ void np_reset_regs(const frl_np_params &p, float (&stv)[SLOTS], float (&clv)[SLOTS], int D,
                                              NV &amount, NV &total, NV &gr, double &init_total, int &day, long long env,
                                              int step)
{
    uint64_t bits = 0;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        float s0 = (j < D && p.init_stocks) ? __ldg(p.init_stocks + j) : 0.0f;
        if (p.train_reset && j < D) {
            if (j % 10 == 0) bits = reset_bits(p.reset_seed, env, step, 1 + j / 10);  // ten 6-bit draws per word
            s0 = fadd(s0, (float)(int)(bits & 63));
            bits >>= 6;
        }
        stv[j] = s0;
        clv[j] = 0.0f;
    }
    day = 0;
    const float asset0 = np_asset_f32<SLOTS>(stv, p.price, D);
    if (p.train_reset) {
        const double factor = 0.95 + (1.05 - 0.95) * reset_uniform01(reset_bits(p.reset_seed, env, step, 0));
        amount = nv_sub(nv(dmul(p.initial_capital, factor), FRL_KIND_PY), nv((double)asset0, FRL_KIND_F32));
    } else {
        amount = nv(p.initial_capital, FRL_KIND_PY);
    }
    total = nv_add(amount, nv((double)asset0, FRL_KIND_F32));
    init_total = total.v;
    gr = nv(0.0, FRL_KIND_PY);
}

template <int SLOTS, int DCT, typename ActT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, FRL_NP_MIN_BLOCKS * 64 / (WARPS * 32))
np_rollout_kernel(const frl_np_params p, const ActT *__restrict__ actions, long long act_step_stride,
                  long long act_env_stride, int n_steps, double *__restrict__ rewards, uint8_t *__restrict__ flags_out,
                  float *__restrict__ obs, int obs_mode, int auto_reset, double *__restrict__ stats)
{
    stats_exchange_previous(stats);
    using SM = NpWarpSmem<SLOTS, ActT>;
    __shared__ SM smem[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SM &sm = smem[warp];
    const int N = p.n_envs, D = DCT > 0 ? DCT : p.stock_dim, T = p.n_days, ld = p.env_stride;
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;

    // ---- load state ----
    const int kinds = p.kinds[n];
    NV amount = nv(p.amount[n], kinds & 3);
    NV total = nv(p.total[n], (kinds >> 2) & 3);
    NV gr = nv(p.gamma_reward[n], (kinds >> 4) & 3);
    int day = p.day[n];
    double init_total = 0.0;
    bool init_total_loaded = false;
    // stocks / cool-down stay in REGISTERS for the arithmetic (all indices are static here); shared
    // memory is only the transposition buffer of the observation writer
    float stv[SLOTS], clv[SLOTS];
    {
        const float *sp = p.stocks + n, *cp = p.cool + n;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            stv[j] = (j < D) ? __ldcs(sp + j * ld) : 0.0f;
            clv[j] = (j < D) ? __ldcs(cp + j * ld) : 0.0f;
        }
    }
    const NV one_minus_sc = nv(dsub(1.0, p.sell_cost_pct), FRL_KIND_PY);
    const NV one_plus_bc = nv(dadd(1.0, p.buy_cost_pct), FRL_KIND_PY);
    const int min_action = (int)dmul(p.max_stock, p.min_stock_rate);  // int(max_stock * min_stock_rate) (:111)
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0, st_liq = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        // ---- stage this step's actions (coalesced, flat) ----
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        stage_actions_flat<SLOTS, ActT>(sm.act, abase, env0, act_env_stride, D, nvalid, lane);
        __syncwarp();

        int flags = 0;
        NV reward = nv(0.0, FRL_KIND_PY);
        // Steady state: `amount` becomes np.float64 at the first trade that the action (not the holding / the cash)
        // caps, and `total_asset` / `gamma_reward` follow one step later; from then on every NEP-50 kind of the
        // step is float64.  When that holds for the whole tile the body runs with the kinds as compile-time
        // constants (A64): no float32 candidates, no kind arithmetic — same operations, same bits.
        const bool all64 = FRL_NP_A64 && __all_sync(0xffffffffu, amount.k == FRL_KIND_F64 && total.k == FRL_KIND_F64 &&
                                                       gr.k == FRL_KIND_F64);
        auto step_body = [&](auto a64_tag) {
            constexpr bool A64 = decltype(a64_tag)::value;
            if (day >= T - 1) {
                // already past the last day (the reference would raise IndexError): inert, done again
                flags = FRL_FLAG_DONE;
            } else {
                // (actions * max_stock).astype(int) is recomputed from the staged row where needed (two uses per
                // stock) instead of holding D more registers
                const ActT *arow = sm.act + lane * D;
    #define FRL_A(j) np_action_to_shares<ActT>(arow[j], p.max_stock)

                day += 1;  // trades happen at the NEW day's prices (:106-107)
                const float *prow = p.price + (size_t)day * p.price_pitch;
    #pragma unroll
                for (int j = 0; j < SLOTS; ++j) clv[j] = fadd(clv[j], 1.0f);  // cool_down += 1

                if (__ldg(p.turb_bool + day) == 0.0f) {
                    // ---- sells in ascending index (:112-119) ----
    #pragma unroll
                    for (int j = 0; j < SLOTS; ++j) {
                        const int aj = (j < D) ? FRL_A(j) : 0;
                        const float pj = (j < D) ? __ldg(prow + j) : 0.0f;  // unconditional: issued ahead of the chain
                        if (j < D && aj < -min_action) {
                            if (pj > 0.0f) {
                                float st = stv[j];
                                NV x;
                                const double nsh = np_u2d(-aj), std_ = np_f2d(st);
                                if (nsh < std_) {  // min(stocks, -action) -> the int64
                                    st = (float)dsub(std_, nsh);
                                    x = nv_mul(nv(dmul(np_f2d(pj), nsh), FRL_KIND_F64), one_minus_sc);
                                } else {  // -> the float32 holding
                                    x = nv_mul(nv((double)fmul(pj, st), FRL_KIND_F32), one_minus_sc);
                                    st = fsub(st, st);
                                }
                                // (common code after the divergent branches: executed once per warp, not once per branch)
                                stv[j] = st;
                                amount = nv_add_t<A64>(amount, x);
                                clv[j] = 0.0f;
                            }
                        }
                    }
                    // ---- buys in ascending index; the divisor has NO cost term (:120-129, quirk Q6) ----
    #pragma unroll
                    for (int j = 0; j < SLOTS; ++j) {
                        const int aj = (j < D) ? FRL_A(j) : 0;
                        const float pj = (j < D) ? __ldg(prow + j) : 0.0f;
                        if (j < D && aj > min_action) {
                            if (pj > 0.0f) {
                                float st = stv[j];
                                NV x;
                                // avail = amount // price is an exact floor, so `action < avail` <=> amount >=
                                // (action+1)*price, which is exact in fp64 (24-bit price x small int): the
                                // division only runs for the cash-limited buys
                                // amount // price is a float32 operation unless amount is float64; a float32 amount is
                                // already float32-valued, only a (weak) Python float has to be rounded first
                                const double am = (!A64 && amount.k == FRL_KIND_PY) ? (double)(float)amount.v : amount.v;
                                const double pjd = np_f2d(pj);
                                const bool plenty = am >= dmul(np_u2d(aj + 1), pjd);
                                double avail = 0.0;
                                // 0 <= amount < price: the quotient is exactly 0 (the usual state of a cash-starved
                                // env) — only the division is skipped, the zero-share update still runs (it can
                                // change the numpy kind of `amount` and it resets the cool-down counter)
                                if (!plenty && !(am >= 0.0 && am < pjd))
                                    avail = np_floor_div(am, pjd, A64 || amount.k == FRL_KIND_F64);
                                if (plenty) {  // min(avail, action) -> the int64
                                    const double nsh = np_u2d(aj);
                                    st = (float)dadd(np_f2d(st), nsh);
                                    x = nv_mul(nv(dmul(pjd, nsh), FRL_KIND_F64), one_plus_bc);
                                } else if (A64 || amount.k == FRL_KIND_F64) {
                                    st = (float)dadd(np_f2d(st), avail);
                                    x = nv_mul(nv(dmul(pjd, avail), FRL_KIND_F64), one_plus_bc);
                                } else {
                                    const float nsh = (float)avail;
                                    st = fadd(st, nsh);
                                    x = nv_mul(nv((double)fmul(pj, nsh), FRL_KIND_F32), one_plus_bc);
                                }
                                stv[j] = st;
                                amount = nv_sub_t<A64>(amount, x);
                                clv[j] = 0.0f;
                            }
                        }
                    }
                } else {
                    // ---- sell everything when turbulence (:131-134) ----
                    flags |= FRL_FLAG_LIQUIDATE;
                    const NV x = nv_mul(nv((double)np_asset_f32<SLOTS>(stv, prow, D), FRL_KIND_F32), one_minus_sc);
                    amount = nv_add_t<A64>(amount, x);
    #pragma unroll
                    for (int j = 0; j < SLOTS; ++j) {
                        stv[j] = 0.0f;
                        clv[j] = 0.0f;
                    }
                    if (valid) st_liq += 1.0;
                }
    #undef FRL_A
                // ---- reward bookkeeping (:136-145) ----
                const NV tot = nv_add_t<A64>(amount, nv((double)np_asset_f32<SLOTS>(stv, prow, D), FRL_KIND_F32));
                reward = nv_mul_t<A64>(nv_sub_t<A64>(tot, total), nv(p.reward_scaling, FRL_KIND_PY));
                total = tot;
                gr = nv_add_t<A64>(nv_mul_t<A64>(gr, nv(p.gamma, FRL_KIND_PY)), reward);
                if (day == T - 1) {
                    flags |= FRL_FLAG_DONE;
                    reward = gr;
                    if (!init_total_loaded) {
                        init_total = p.init_total[n];
                        init_total_loaded = true;
                    }
                    const double er = (A64 || total.k == FRL_KIND_F64) ? __ddiv_rn(total.v, init_total)
                                                                 : (double)__fdiv_rn((float)total.v, (float)init_total);
                    if (valid) {
                        p.episode_return[n] = er;
                        st_done += 1.0;
                        st_epi += total.v;
                    }
                }
            }
        };
        if (all64)
            step_body(std::true_type{});
        else
            step_body(std::false_type{});
        if (valid) {
            if (rewards) rewards[(size_t)k * N + n] = reward.v;
            if (flags_out) flags_out[(size_t)k * N + n] = (uint8_t)(flags | (reward.k << FRL_NP_REWARD_KIND_SHIFT));
            st_r += reward.v;
            st_r2 += reward.v * reward.v;
        }
        if ((flags & FRL_FLAG_DONE) && auto_reset) {
            np_reset_regs<SLOTS>(p, stv, clv, D, amount, total, gr, init_total, day, n, k);
            init_total_loaded = true;
            if (valid) p.init_total[n] = init_total;
        }
        if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
            __syncwarp();  // the previous step's rows have been consumed
            if (DCT > 0) {
                // finished chunks CLO..CHI of the row (np_write_obs_rows_ct): scaled stocks, cool-down counters and
                // the template positions sharing those chunks (this env's day)
                using L = NpObsCT<(DCT > 0 ? DCT : 30)>;
                static_assert(DCT <= 0 || 32 * L::NE <= 2 * SLOTS, "transposition buffer too small");
                constexpr int P0 = 32 * L::CLO, S0 = 3 + DCT, E0 = 3 + 3 * DCT, P1 = 32 * (L::CHI + 1);
                const float *trow = p.obs_tmpl + (size_t)day * p.obs_dim;
#pragma unroll
                for (int pos = P0; pos < S0; ++pos) sm.sc[(pos - P0) * kPitch + lane] = __ldg(trow + pos);
#pragma unroll
                for (int pos = E0; pos < P1; ++pos)
                    sm.sc[(pos - P0) * kPitch + lane] = pos < p.obs_dim ? __ldg(trow + pos) : 0.0f;
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) {
                    if (j < D) {
                        sm.sc[(S0 + j - P0) * kPitch + lane] = fmul(stv[j], 0.015625f);  // 2**-6
                        sm.sc[(S0 + DCT + j - P0) * kPitch + lane] = clv[j];
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < SLOTS; ++j) {
                    if (j < D) {
                        sm.sc[j * kPitch + lane] = stv[j];
                        sm.sc[(D + j) * kPitch + lane] = clv[j];
                    }
                }
            }
            sm.amountf[lane] = np_amount_obs(amount, p.obs_amount_floor);
            sm.day[lane] = day;
            __syncwarp();
            float *o = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * p.obs_dim : (size_t)0);
            np_write_obs_tile<DCT>(p, sm, o, env0, nvalid, lane);
        }
    }

    // ---- store state ----
    if (valid) {
        p.amount[n] = amount.v;
        p.total[n] = total.v;
        p.gamma_reward[n] = gr.v;
        p.kinds[n] = (uint8_t)(amount.k | (total.k << 2) | (gr.k << 4));
        p.day[n] = day;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            if (j < D) {
                __stcs(p.stocks + n + j * ld, stv[j]);  // streaming: nothing on this SM reads them back (+0.4 %)
                __stcs(p.cool + n + j * ld, clv[j]);
            }
        }
    }
    if (stats) {
        double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, valid ? total.v : 0.0, st_liq,
                                 valid ? (double)n_steps : 0.0, 0.0};
        reduce_stats8(v, lane, stats);
    }
}

// ---- reset / observe ---------------------------------------------------------------------------
__global__ void np_reset_kernel(const frl_np_params p, const uint8_t *__restrict__ mask,
                                const float *__restrict__ stocks0, const double *__restrict__ factor)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.n_envs) return;
    if (mask && !mask[n]) return;
    const int D = p.stock_dim, ld = p.env_stride;
    // (stocks * price[0]).sum() with numpy's pairwise order, from a thread-private copy
    float x[128];
    for (int j = 0; j < 128; ++j) x[j] = 0.0f;
    for (int j = 0; j < D; ++j) {
        const float st = (stocks0 && factor) ? stocks0[n + (size_t)j * ld] : (p.init_stocks ? p.init_stocks[j] : 0.0f);
        p.stocks[n + (size_t)j * ld] = st;
        p.cool[n + (size_t)j * ld] = 0.0f;
        x[j] = fmul(st, p.price[j]);
    }
    float asset;
    if (D < 8) {
        asset = 0.0f;
        for (int j = 0; j < D; ++j) asset = fadd(asset, x[j]);
    } else {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = x[j];
        const int nb = D >> 3;
        for (int b = 1; b < nb; ++b)
            for (int j = 0; j < 8; ++j) r[j] = fadd(r[j], x[8 * b + j]);
        asset = fadd(fadd(fadd(r[0], r[1]), fadd(r[2], r[3])), fadd(fadd(r[4], r[5]), fadd(r[6], r[7])));
        for (int j = 8 * nb; j < D; ++j) asset = fadd(asset, x[j]);
    }
    NV amount;
    if (stocks0 && factor)  // initial_capital * rd.uniform() [py] - (stocks*price).sum() [f32] -> f32
        amount = nv_sub(nv(dmul(p.initial_capital, factor[n]), FRL_KIND_PY), nv((double)asset, FRL_KIND_F32));
    else
        amount = nv(p.initial_capital, FRL_KIND_PY);
    const NV total = nv_add(amount, nv((double)asset, FRL_KIND_F32));
    p.day[n] = 0;
    p.amount[n] = amount.v;
    p.total[n] = total.v;
    p.init_total[n] = total.v;
    p.gamma_reward[n] = 0.0;
    p.kinds[n] = (uint8_t)(amount.k | (total.k << 2) | (FRL_KIND_PY << 4));
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) np_observe_kernel(const frl_np_params p, float *__restrict__ obs)
{
    using SM = NpWarpSmem<32, float>;
    __shared__ SM smem[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SM &sm = smem[warp];
    const int N = p.n_envs, D = p.stock_dim;
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const long long n = lane < nvalid ? env0 + lane : (long long)N - 1;
    for (int j = 0; j < D; ++j) {
        sm.sc[j * kPitch + lane] = p.stocks[n + (size_t)j * p.env_stride];
        sm.sc[(D + j) * kPitch + lane] = p.cool[n + (size_t)j * p.env_stride];
    }
    sm.amountf[lane] = np_amount_obs(nv(p.amount[n], p.kinds[n] & 3), p.obs_amount_floor);
    sm.day[lane] = p.day[n];
    __syncwarp();
    np_write_obs_tile<0>(p, sm, obs, env0, nvalid, lane);
}

int32_t np_validate(const frl_np_params *p)
{
    FRL_REQUIRE(p != nullptr, "np: params is NULL");
    FRL_REQUIRE(p->n_envs >= 1, "np: n_envs must be >= 1 (got %d)", p->n_envs);
    FRL_REQUIRE(p->stock_dim >= 1 && p->stock_dim <= 128, "np: stock_dim must be in 1..128 (got %d)", p->stock_dim);
    FRL_REQUIRE((p->price_pitch == 32 || p->price_pitch == 128) && p->price_pitch >= p->stock_dim,
                "np: price_pitch must be 32 or 128 and >= stock_dim (got %d for D=%d)", p->price_pitch, p->stock_dim);
    FRL_REQUIRE(p->tech_dim >= 0 && p->n_days >= 2, "np: bad tech_dim/n_days (%d, %d)", p->tech_dim, p->n_days);
    FRL_REQUIRE(p->obs_dim == 3 + 3 * p->stock_dim + p->tech_dim, "np: obs_dim %d != 1 + 2 + 3D + tech_dim = %d",
                p->obs_dim, 3 + 3 * p->stock_dim + p->tech_dim);
    FRL_REQUIRE(p->env_stride >= p->n_envs, "np: env_stride %d < n_envs %d", p->env_stride, p->n_envs);
    FRL_REQUIRE((long long)p->env_stride * p->price_pitch < (1LL << 31), "np: env_stride %d too large", p->env_stride);
    FRL_REQUIRE(p->price && p->turb_bool && p->obs_tmpl, "np: table pointer is NULL");
    FRL_REQUIRE(p->amount && p->kinds && p->stocks && p->cool && p->day && p->total && p->gamma_reward &&
                    p->init_total && p->episode_return,
                "np: state pointer is NULL");
    return FRL_OK;
}

// stock counts from which frl_np_* use the streaming kernel of np_wide.cu (always above 32);
// frl_set_option("np_wide_min_d", v) / FRL_NP_WIDE_MIN_D lower it so that tests can run every shape through it
int g_np_wide_min_d = -1;
int np_wide_min_d()
{
    if (g_np_wide_min_d < 0) {
        const char *m = getenv("FRL_NP_WIDE_MIN_D");
        g_np_wide_min_d = m ? atoi(m) : 33;
        if (g_np_wide_min_d < 1 || g_np_wide_min_d > 33) g_np_wide_min_d = 33;
    }
    return g_np_wide_min_d;
}

template <int SLOTS, int DCT, typename ActT, int WARPS>
void np_launch(const frl_np_params &p, const void *actions, long long sstride, long long estride, int n_steps,
               double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    const unsigned grid = (unsigned)((tiles + WARPS - 1) / WARPS);
    np_rollout_kernel<SLOTS, DCT, ActT, WARPS><<<grid, WARPS * 32, 0, st>>>(p, (const ActT *)actions, sstride, estride, n_steps,
                                                                       rewards, flags, obs, obs_mode, auto_reset, stats);
}

}  // namespace

int32_t np_set_option(const char *name, int64_t value)
{
    if (!strcmp(name, "np_wide_bulk")) {  // 0: the streaming kernel keeps its generic action staging for every shape
        g_npw_bulk = value != 0;
        return FRL_OK;
    }
    if (strcmp(name, "np_wide_min_d")) return 1;  // not ours
    g_np_wide_min_d = value < 1 ? 1 : (value > 33 ? 33 : (int)value);
    return FRL_OK;
}

}  // namespace frl

using namespace frl;

extern "C" int32_t frl_np_observe(const frl_np_params *p, float *obs, void *stream)
{
    if (int32_t rc = np_validate(p)) return rc;
    FRL_REQUIRE(obs != nullptr, "np_observe: obs is NULL");
    if (p->stock_dim >= np_wide_min_d()) {
        launch_np_observe_wide(*p, obs, (cudaStream_t)stream);
        return check_launch("np_observe(wide)");
    }
    constexpr int W = 2;
    const long long tiles = ((long long)p->n_envs + 31) / 32;
    np_observe_kernel<W><<<(unsigned)((tiles + W - 1) / W), W * 32, 0, (cudaStream_t)stream>>>(*p, obs);
    return check_launch("np_observe");
}

extern "C" int32_t frl_np_reset(const frl_np_params *p, const uint8_t *mask, const float *stocks0, const double *factor,
                                float *obs, void *stream)
{
    if (int32_t rc = np_validate(p)) return rc;
    FRL_REQUIRE((stocks0 == nullptr) == (factor == nullptr), "np_reset: stocks0 and factor must be given together");
    np_reset_kernel<<<(p->n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*p, mask, stocks0, factor);
    if (int32_t rc = check_launch("np_reset")) return rc;
    if (obs) return frl_np_observe(p, obs, stream);
    return FRL_OK;
}

extern "C" int32_t frl_np_rollout(const frl_np_params *p, const void *actions, int32_t actions_f64, int64_t act_step_stride,
                                  int64_t act_env_stride, int32_t n_steps, double *rewards, uint8_t *flags, float *obs,
                                  int32_t obs_mode, int32_t auto_reset, double *stats, void *stream)
{
    if (int32_t rc = np_validate(p)) return rc;
    FRL_REQUIRE(actions != nullptr, "np_rollout: actions is NULL");
    FRL_REQUIRE(n_steps >= 1, "np_rollout: n_steps must be >= 1 (got %d)", n_steps);
    FRL_REQUIRE(act_env_stride >= p->stock_dim, "np_rollout: act_env_stride %lld < stock_dim", (long long)act_env_stride);
    FRL_REQUIRE(obs_mode >= FRL_OBS_NONE && obs_mode <= FRL_OBS_ALL, "np_rollout: bad obs_mode %d", obs_mode);
    FRL_REQUIRE(obs_mode == FRL_OBS_NONE || obs != nullptr, "np_rollout: obs is NULL but obs_mode=%d", obs_mode);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = p->stock_dim;
    if (D >= np_wide_min_d()) {
        launch_np_wide(*p, actions, actions_f64, act_step_stride, act_env_stride, n_steps, rewards, flags, obs, obs_mode,
                       auto_reset, stats, st);
        return check_launch("np_rollout(wide)");
    }
#define FRL_GO(SLOTS, DCT)                                                                                        \
    do {                                                                                                          \
        if (actions_f64)                                                                                          \
            np_launch<SLOTS, DCT, double, 2>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards,      \
                                             flags, obs, obs_mode, auto_reset, stats, st);                        \
        else                                                                                                      \
            np_launch<SLOTS, DCT, float, FRL_NP_WARPS>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards,       \
                                            flags, obs, obs_mode, auto_reset, stats, st);                         \
    } while (0)
    if (D <= 8)
        FRL_GO(8, 0);
    else if (D <= 16)
        FRL_GO(16, 0);
    else if (D == 30)
        FRL_GO(32, 30);  // DOW-30: stock count compiled in
    else
        FRL_GO(32, 0);
#undef FRL_GO
    return check_launch("np_rollout");
}

extern "C" int32_t frl_np_step(const frl_np_params *p, const void *actions, int32_t actions_f64, double *rewards,
                               uint8_t *flags, float *obs, int32_t auto_reset, double *stats, void *stream)
{
    if (p == nullptr) {
        set_error("np_step: params is NULL");
        return FRL_E_INVALID;
    }
    return frl_np_rollout(p, actions, actions_f64, (int64_t)p->n_envs * p->stock_dim, p->stock_dim, 1, rewards, flags, obs,
                          obs ? FRL_OBS_LAST : FRL_OBS_NONE, auto_reset, stats, stream);
}
