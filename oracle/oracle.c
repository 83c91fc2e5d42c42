/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the FinRL env step path.  See oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC   (NO fused multiply-add, NO fast-math:
 * the reference's arithmetic is plain IEEE double/float, one rounding per operation).
 * Parity: pinned against the unmodified reference by the npz fixtures under tests/golden.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* thread count of the env-parallel loops (bench.py's CPU baseline uses every host thread; torchrun
 * exports OMP_NUM_THREADS=1, so the environment variable alone is not enough) */
int ora_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ======================================================================================= */
/* building blocks                                                                         */
/* ======================================================================================= */

/* numpy/core/src/npymath/npy_math_internal.h.src: npy_divmod / npy_floor_divide (numpy 2.3).
 * Reached from env_stocktrading.py:178-180 (`self.state[0] // (price * (1 + buy_cost_pct))`) and
 * env_stocktrading_np.py:124 (`self.amount // price[index]`). */
double ora_floor_divide_f64(double a, double b)
{
    if (b == 0.0) return a / b;
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0) != (mod < 0)) {
            mod += b;
            div -= 1.0;
        }
    }
    double fl;
    if (div != 0.0) {
        fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
    } else {
        fl = copysign(0.0, a / b);
    }
    return fl;
}

float ora_floor_divide_f32(float a, float b)
{
    if (b == 0.0f) return a / b;
    float mod = fmodf(a, b);
    float div = (a - mod) / b;
    if (mod != 0.0f) {
        if ((b < 0) != (mod < 0)) {
            mod += b;
            div -= 1.0f;
        }
    }
    float fl;
    if (div != 0.0f) {
        fl = floorf(div);
        if (div - fl > 0.5f) fl += 1.0f;
    } else {
        fl = copysignf(0.0f, a / b);
    }
    return fl;
}

/* np.argsort(int64) default kind on AVX2/AVX-512 hosts (x86-simd-sort) for n <= 256: an
 * ascending bitonic network on next_pow2(max(n,8)) slots padded with INT64_MAX, whose
 * compare-exchange swaps (key,index) only on strict `>` (so ties keep network order, which is
 * NOT the stable order).  SURVEY.md H1; call site env_stocktrading.py:317. */
void ora_argsort_i64(const int64_t *keys, int n, int32_t *order_out)
{
    int slots = 8;
    while (slots < n) slots <<= 1;
    int64_t *k = (int64_t *)malloc(sizeof(int64_t) * (size_t)slots);
    int32_t *ix = (int32_t *)malloc(sizeof(int32_t) * (size_t)slots);
    for (int j = 0; j < slots; ++j) {
        k[j] = j < n ? keys[j] : INT64_MAX;
        ix[j] = j;
    }
#define CEX(lo, hi)                                                                                \
    do {                                                                                           \
        if (k[lo] > k[hi]) {                                                                       \
            int64_t tk = k[lo]; k[lo] = k[hi]; k[hi] = tk;                                         \
            int32_t ti = ix[lo]; ix[lo] = ix[hi]; ix[hi] = ti;                                     \
        }                                                                                          \
    } while (0)
    for (int blk = 2; blk <= slots; blk <<= 1) {
        /* "flip" stage: mirror pairs inside each block */
        for (int b = 0; b < slots; b += blk)
            for (int i = 0; i < blk / 2; ++i) CEX(b + i, b + blk - 1 - i);
        /* half-cleaners */
        for (int d = blk / 4; d >= 1; d >>= 1)
            for (int b = 0; b < slots; b += 2 * d)
                for (int i = 0; i < d; ++i) CEX(b + i, b + i + d);
    }
#undef CEX
    for (int j = 0; j < n; ++j) order_out[j] = ix[j];
    free(k);
    free(ix);
}

/* numpy/core/src/umath/loops_utils.h.src: pairwise_sum (ndarray.sum on a contiguous vector). */
float ora_pairwise_sum_f32(const float *a, int n)
{
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        int i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return ora_pairwise_sum_f32(a, n2) + ora_pairwise_sum_f32(a + n2, n - n2);
    }
}

double ora_pairwise_sum_f64(const double *a, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return ora_pairwise_sum_f64(a, n2) + ora_pairwise_sum_f64(a + n2, n - n2);
    }
}

/* ======================================================================================= */
/* A1: StockTradingEnv — finrl/meta/env_stock_trading/env_stocktrading.py                  */
/* ======================================================================================= */

#define MAXD 256

static inline int32_t trd_state_day(int32_t sday) { return sday < 0 ? -sday - 1 : sday; }

/* _initiate_state / _update_state (env_stocktrading.py:398-478) followed by the float32 cast
 * DummyVecEnv applies when it copies the list into its observation buffer. */
static void trd_obs_one(const ora_trading_cfg *c, double cash, const int32_t *hold, int32_t sday,
                        float *obs)
{
    const int D = c->stock_dim, K = c->n_tech, T = c->n_days;
    const int sd = trd_state_day(sday);
    obs[0] = (float)cash;
    for (int i = 0; i < D; ++i) obs[1 + i] = (float)c->close[(size_t)sd * D + i];
    for (int i = 0; i < D; ++i) obs[1 + D + i] = (float)(double)hold[i];
    for (int k = 0; k < K; ++k)
        for (int i = 0; i < D; ++i)
            obs[1 + 2 * D + k * D + i] = (float)c->tech[((size_t)k * T + sd) * D + i];
}

void ora_trading_obs(const ora_trading_cfg *c, const ora_trading_state *s, float *obs)
{
    const int D = c->stock_dim, O = 1 + 2 * D + c->n_tech * D;
    for (int n = 0; n < c->n_envs; ++n)
        trd_obs_one(c, s->cash[n], s->hold + (size_t)n * D, s->sday[n], obs + (size_t)n * O);
}

/* __init__ (env_stocktrading.py:48-100) */
void ora_trading_init(const ora_trading_cfg *c, ora_trading_state *s, int32_t day0)
{
    const int D = c->stock_dim;
    for (int n = 0; n < c->n_envs; ++n) {
        s->cash[n] = c->initial_amount;
        for (int i = 0; i < D; ++i) s->hold[(size_t)n * D + i] = c->init_hold ? c->init_hold[i] : 0;
        s->day[n] = day0;
        s->sday[n] = -day0 - 1;
        s->cost[n] = 0.0;
        s->trades[n] = 0;
        s->reward[n] = 0.0;
        s->episode[n] = 0;
    }
}

/* reset (env_stocktrading.py:359-393): the state list is rebuilt from whatever day's rows are
 * still loaded (line 361) BEFORE day is set to 0 (line 380) — quirk Q1. */
static void trd_reset_one(const ora_trading_cfg *c, ora_trading_state *s, int n)
{
    const int D = c->stock_dim;
    s->cash[n] = c->initial_amount;
    for (int i = 0; i < D; ++i) s->hold[(size_t)n * D + i] = c->init_hold ? c->init_hold[i] : 0;
    s->sday[n] = -s->day[n] - 1;
    s->day[n] = 0;
    s->cost[n] = 0.0;
    s->trades[n] = 0;
    s->episode[n] += 1;
}

void ora_trading_reset(const ora_trading_cfg *c, ora_trading_state *s, const uint8_t *mask)
{
    for (int n = 0; n < c->n_envs; ++n)
        if (!mask || mask[n]) trd_reset_one(c, s, n);
}

static void trd_step_one(const ora_trading_cfg *c, ora_trading_state *s, int n, const void *actions,
                         int actions_f64, double *reward_out, uint8_t *flags_out, float *obs,
                         int auto_reset)
{
    const int D = c->stock_dim, T = c->n_days, O = 1 + 2 * D + c->n_tech * D;
    int32_t *hold = s->hold + (size_t)n * D;
    uint8_t flags = 0;

    /* step(): terminal branch (env_stocktrading.py:221-301): no state change, previous scaled
     * reward is returned again (quirk Q3). */
    if (s->day[n] >= T - 1) {
        flags |= ORA_FLAG_DONE;
        if (reward_out) reward_out[n] = s->reward[n];
        if (auto_reset) trd_reset_one(c, s, n); /* DummyVecEnv.step_wait: obs = env.reset() */
        if (obs) trd_obs_one(c, s->cash[n], hold, s->sday[n], obs + (size_t)n * O);
        if (flags_out) flags_out[n] = flags;
        return;
    }

    const int sd = trd_state_day(s->sday[n]);
    const double turb = (s->sday[n] < 0) ? 0.0 : c->risk[sd];
    const int liq = c->use_turbulence && (turb >= c->turbulence_threshold);
    if (liq) flags |= ORA_FLAG_LIQUIDATE;
    const double *price = c->close + (size_t)sd * D;
    const double *tech0 = c->n_tech > 0 ? c->tech + (size_t)sd * D : NULL; /* first indicator */

    /* actions = (actions * hmax).astype(int) in the INPUT dtype (lines 304-307) */
    int64_t a[MAXD];
    for (int i = 0; i < D; ++i) {
        if (actions_f64) {
            double v = ((const double *)actions)[(size_t)n * D + i] * c->hmax;
            a[i] = (int64_t)v;
        } else {
            float v = ((const float *)actions)[(size_t)n * D + i] * (float)c->hmax;
            a[i] = (int64_t)v;
        }
    }
    if (liq)
        for (int i = 0; i < D; ++i) a[i] = (int64_t)(-c->hmax); /* lines 308-310 */

    /* begin_total_asset: Python sum(), sequential (lines 311-314) */
    double acc = 0.0;
    for (int i = 0; i < D; ++i) acc = acc + price[i] * (double)hold[i];
    const double begin = s->cash[n] + acc;

    int32_t order[MAXD];
    ora_argsort_i64(a, D, order); /* line 317 */
    int nsell = 0, nbuy = 0;
    for (int i = 0; i < D; ++i) {
        nsell += a[i] < 0;
        nbuy += a[i] > 0;
    }
    double cash = s->cash[n], cost = s->cost[n];
    int32_t trades = s->trades[n];
    const double one_minus_sc = 1 - c->sell_cost_pct, one_plus_bc = 1 + c->buy_cost_pct;

    /* _sell_stock (lines 102-169) */
    for (int k = 0; k < nsell; ++k) {
        const int i = order[k];
        if (liq) {
            if (price[i] > 0 && hold[i] > 0) {
                const double nsh = (double)hold[i];
                cash += price[i] * nsh * one_minus_sc;
                hold[i] = 0;
                cost += price[i] * nsh * c->sell_cost_pct;
                trades += 1;
            }
        } else {
            const int disabled = tech0 && (tech0[i] == 1.0); /* `state[...] != True` (Q2) */
            if (!disabled && hold[i] > 0) {
                int64_t m = a[i] < 0 ? -a[i] : a[i];
                if ((int64_t)hold[i] < m) m = hold[i];
                const double nsh = (double)m;
                cash += price[i] * nsh * one_minus_sc;
                hold[i] -= (int32_t)m;
                cost += price[i] * nsh * c->sell_cost_pct;
                trades += 1;
            }
        }
    }
    /* _buy_stock (lines 171-213); buy_index = argsort[::-1][:nbuy] (line 319) */
    for (int k = 0; k < nbuy; ++k) {
        const int i = order[D - 1 - k];
        if (liq) continue; /* turbulence >= threshold: no buys (lines 204-211) */
        const int disabled = tech0 && (tech0[i] == 1.0);
        if (disabled) continue;
        const double avail = ora_floor_divide_f64(cash, price[i] * one_plus_bc);
        /* Python min(avail, action): returns action only if action < avail */
        const double nsh = ((double)a[i] < avail) ? (double)a[i] : avail;
        cash -= price[i] * nsh * one_plus_bc;
        hold[i] += (int32_t)nsh;
        cost += price[i] * nsh * c->buy_cost_pct;
        trades += 1; /* even when nsh == 0 (Q5) */
    }

    /* state: s -> s+1 (lines 335-342) */
    s->day[n] += 1;
    s->sday[n] = s->day[n];
    const double *pnew = c->close + (size_t)s->day[n] * D;
    acc = 0.0;
    for (int i = 0; i < D; ++i) acc = acc + pnew[i] * (double)hold[i];
    const double end = cash + acc;
    const double reward = (end - begin) * c->reward_scaling; /* lines 350-352 */

    s->cash[n] = cash;
    s->cost[n] = cost;
    s->trades[n] = trades;
    s->reward[n] = reward;
    if (reward_out) reward_out[n] = reward;
    if (flags_out) flags_out[n] = flags;
    if (obs) trd_obs_one(c, cash, hold, s->sday[n], obs + (size_t)n * O);
}

void ora_trading_step(const ora_trading_cfg *c, ora_trading_state *s, const void *actions,
                      int actions_f64, double *reward_out, uint8_t *flags_out, float *obs,
                      int auto_reset)
{
#pragma omp parallel for schedule(static)
    for (int n = 0; n < c->n_envs; ++n)
        trd_step_one(c, s, n, actions, actions_f64, reward_out, flags_out, obs, auto_reset);
}

/* ======================================================================================= */
/* A2: numpy / ElegantRL StockTradingEnv — finrl/meta/env_stock_trading/env_stocktrading_np.py */
/* ======================================================================================= */
/* The reference mixes Python floats, np.float32 and np.float64 scalars; under NEP 50 the dtype of
 * `amount`, `total_asset`, `gamma_reward` and `reward` is data dependent.  A value is
 * carried as (double v, kind); kind PY behaves as f64 in Python-only arithmetic and is "weak"
 * (adopts the other operand's dtype) when it meets a numpy scalar. */

typedef struct { double v; int k; } nv; /* numpy/Python scalar */

static inline nv nv_make(double v, int k) { nv r = {v, k}; return r; }

/* x (kind kx) <op> y where y is a strong numpy scalar of kind ky (F32 or F64) */
static inline nv nv_add(nv x, nv y)
{
    if (x.k == ORA_KIND_PY && y.k == ORA_KIND_PY) return nv_make(x.v + y.v, ORA_KIND_PY);
    if (x.k == ORA_KIND_F64 || y.k == ORA_KIND_F64) return nv_make(x.v + y.v, ORA_KIND_F64);
    return nv_make((double)((float)x.v + (float)y.v), ORA_KIND_F32); /* f32 (+ weak py) */
}
static inline nv nv_sub(nv x, nv y)
{
    if (x.k == ORA_KIND_PY && y.k == ORA_KIND_PY) return nv_make(x.v - y.v, ORA_KIND_PY);
    if (x.k == ORA_KIND_F64 || y.k == ORA_KIND_F64) return nv_make(x.v - y.v, ORA_KIND_F64);
    return nv_make((double)((float)x.v - (float)y.v), ORA_KIND_F32);
}
static inline nv nv_mul(nv x, nv y)
{
    if (x.k == ORA_KIND_PY && y.k == ORA_KIND_PY) return nv_make(x.v * y.v, ORA_KIND_PY);
    if (x.k == ORA_KIND_F64 || y.k == ORA_KIND_F64) return nv_make(x.v * y.v, ORA_KIND_F64);
    return nv_make((double)((float)x.v * (float)y.v), ORA_KIND_F32);
}

static float np_asset_f32(const ora_np_cfg *c, const float *stocks, const float *price)
{
    float prod[MAXD];
    for (int i = 0; i < c->stock_dim; ++i) prod[i] = stocks[i] * price[i];
    return ora_pairwise_sum_f32(prod, c->stock_dim); /* (self.stocks * price).sum() */
}

/* get_state (:149-162) */
static void np_obs_one(const ora_np_cfg *c, const ora_np_state *s, int n, float *obs)
{
    const int D = c->stock_dim, day = s->day[n];
    const float *price = c->price + (size_t)day * D;
    const float scale = 0.015625f; /* 2**-6 */
    /* np.array(self.amount * 2**-12, dtype=np.float32): a power-of-two scale commutes with the cast */
    double a = s->amount[n];
    int ak = s->amount_kind[n];
    if (c->obs_amount_floor > a) { /* Python max(self.amount, 1e4) returns the float 1e4 only if it is larger */
        a = c->obs_amount_floor;
        ak = ORA_KIND_PY;
    }
    obs[0] = (ak == ORA_KIND_F32) ? (float)a * 0.000244140625f : (float)(a * 0.000244140625);
    obs[1] = c->turb_ary[day];
    obs[2] = c->turb_bool[day];
    for (int i = 0; i < D; ++i) obs[3 + i] = price[i] * scale;
    for (int i = 0; i < D; ++i) obs[3 + D + i] = s->stocks[(size_t)n * D + i] * scale;
    for (int i = 0; i < D; ++i) obs[3 + 2 * D + i] = s->cool[(size_t)n * D + i];
    for (int i = 0; i < c->tech_dim; ++i) obs[3 + 3 * D + i] = c->tech[(size_t)day * c->tech_dim + i];
}

void ora_np_obs(const ora_np_cfg *c, const ora_np_state *s, float *obs)
{
    const int O = 3 + 3 * c->stock_dim + c->tech_dim;
    for (int n = 0; n < c->n_envs; ++n) np_obs_one(c, s, n, obs + (size_t)n * O);
}

/* reset (:80-101) */
void ora_np_reset(const ora_np_cfg *c, ora_np_state *s, const uint8_t *mask, const float *stocks0,
                  const double *factor)
{
    const int D = c->stock_dim;
    for (int n = 0; n < c->n_envs; ++n) {
        if (mask && !mask[n]) continue;
        float *st = s->stocks + (size_t)n * D;
        s->day[n] = 0;
        const float *price = c->price;
        nv amount;
        if (stocks0 && factor) { /* if_train: random initial position supplied by the caller */
            for (int i = 0; i < D; ++i) st[i] = stocks0[(size_t)n * D + i];
            /* initial_capital * rd.uniform(..)  [py*py]  -  (stocks*price).sum() [f32]  ->  f32 */
            amount = nv_sub(nv_make(c->initial_capital * factor[n], ORA_KIND_PY),
                            nv_make(np_asset_f32(c, st, price), ORA_KIND_F32));
        } else {
            for (int i = 0; i < D; ++i) st[i] = c->init_stocks ? c->init_stocks[i] : 0.0f;
            amount = nv_make(c->initial_capital, ORA_KIND_PY);
        }
        for (int i = 0; i < D; ++i) s->cool[(size_t)n * D + i] = 0.0f;
        nv total = nv_add(amount, nv_make(np_asset_f32(c, st, price), ORA_KIND_F32));
        s->amount[n] = amount.v; s->amount_kind[n] = (uint8_t)amount.k;
        s->total[n] = total.v;   s->total_kind[n] = (uint8_t)total.k;
        s->init_total[n] = total.v;
        s->gamma_reward[n] = 0.0; s->gr_kind[n] = ORA_KIND_PY;
    }
}

static void np_step_one(const ora_np_cfg *c, ora_np_state *s, int n, const float *actions, double *reward_out,
                        uint8_t *reward_kind_out, uint8_t *flags_out, float *obs)
{
    const int D = c->stock_dim, O = 3 + 3 * D + c->tech_dim;
    float *stocks = s->stocks + (size_t)n * D, *cool = s->cool + (size_t)n * D;
    const nv one_minus_sc = nv_make(1 - c->sell_cost_pct, ORA_KIND_PY);
    const nv one_plus_bc = nv_make(1 + c->buy_cost_pct, ORA_KIND_PY);
    uint8_t flags = 0;

    /* actions = (actions * self.max_stock).astype(int)  — f32 array * Python float -> f32 (:104) */
    int64_t a[MAXD];
    for (int i = 0; i < D; ++i) a[i] = (int64_t)(actions[(size_t)n * D + i] * (float)c->max_stock);

    s->day[n] += 1; /* trades happen at the NEW day's price (:106-107) */
    const int day = s->day[n];
    const float *price = c->price + (size_t)day * D;
    for (int i = 0; i < D; ++i) cool[i] += 1.0f;
    nv amount = nv_make(s->amount[n], s->amount_kind[n]);

    if (c->turb_bool[day] == 0.0f) {
        const int64_t min_action = (int64_t)(c->max_stock * c->min_stock_rate); /* int(...) (:111) */
        for (int i = 0; i < D; ++i) { /* sells, ascending index (:112-119) */
            if (a[i] < -min_action && price[i] > 0) {
                /* min(self.stocks[index], -actions[index]) -> the int64 iff it is SMALLER */
                nv x;
                if ((double)(-a[i]) < (double)stocks[i]) {
                    const double nsh = (double)(-a[i]); /* int64: f32 * int64 -> f64 */
                    stocks[i] = (float)((double)stocks[i] - nsh);
                    x = nv_mul(nv_make((double)price[i] * nsh, ORA_KIND_F64), one_minus_sc);
                } else {
                    const float nsh = stocks[i]; /* f32 */
                    stocks[i] = stocks[i] - nsh;
                    x = nv_mul(nv_make((double)(price[i] * nsh), ORA_KIND_F32), one_minus_sc);
                }
                amount = nv_add(amount, x);
                cool[i] = 0.0f;
            }
        }
        for (int i = 0; i < D; ++i) { /* buys, ascending index (:120-129) */
            if (a[i] > min_action && price[i] > 0) {
                /* self.amount // price[index]: f32 unless amount is already f64 */
                nv avail;
                if (amount.k == ORA_KIND_F64)
                    avail = nv_make(ora_floor_divide_f64(amount.v, (double)price[i]), ORA_KIND_F64);
                else
                    avail = nv_make((double)ora_floor_divide_f32((float)amount.v, price[i]), ORA_KIND_F32);
                nv x;
                if ((double)a[i] < avail.v) { /* min(avail, action) -> the int64 */
                    const double nsh = (double)a[i];
                    stocks[i] = (float)((double)stocks[i] + nsh);
                    x = nv_mul(nv_make((double)price[i] * nsh, ORA_KIND_F64), one_plus_bc);
                } else if (avail.k == ORA_KIND_F64) {
                    stocks[i] = (float)((double)stocks[i] + avail.v);
                    x = nv_mul(nv_make((double)price[i] * avail.v, ORA_KIND_F64), one_plus_bc);
                } else {
                    const float nsh = (float)avail.v;
                    stocks[i] = stocks[i] + nsh;
                    x = nv_mul(nv_make((double)(price[i] * nsh), ORA_KIND_F32), one_plus_bc);
                }
                amount = nv_sub(amount, x);
                cool[i] = 0.0f;
            }
        }
    } else { /* sell everything when turbulence (:131-134) */
        flags |= ORA_FLAG_LIQUIDATE;
        nv x = nv_mul(nv_make((double)np_asset_f32(c, stocks, price), ORA_KIND_F32), one_minus_sc);
        amount = nv_add(amount, x);
        for (int i = 0; i < D; ++i) { stocks[i] = 0.0f; cool[i] = 0.0f; }
    }
    s->amount[n] = amount.v; s->amount_kind[n] = (uint8_t)amount.k;
    if (obs) np_obs_one(c, s, n, obs + (size_t)n * O);

    nv total = nv_add(amount, nv_make((double)np_asset_f32(c, stocks, price), ORA_KIND_F32));
    nv reward = nv_mul(nv_sub(total, nv_make(s->total[n], s->total_kind[n])), nv_make(c->reward_scaling, ORA_KIND_PY));
    s->total[n] = total.v; s->total_kind[n] = (uint8_t)total.k;
    nv gr = nv_add(nv_mul(nv_make(s->gamma_reward[n], s->gr_kind[n]), nv_make(c->gamma, ORA_KIND_PY)), reward);
    s->gamma_reward[n] = gr.v; s->gr_kind[n] = (uint8_t)gr.k;
    if (day == c->n_days - 1) { /* done = self.day == self.max_step (:142-145) */
        flags |= ORA_FLAG_DONE;
        reward = gr;
        /* total_asset / initial_total_asset: initial is np.float32 (reset), so f32 unless total is f64 */
        s->episode_return[n] = (total.k == ORA_KIND_F64) ? total.v / s->init_total[n]
                                                         : (double)((float)total.v / (float)s->init_total[n]);
    }
    if (reward_out) reward_out[n] = reward.v;
    if (reward_kind_out) reward_kind_out[n] = (uint8_t)reward.k;
    if (flags_out) flags_out[n] = flags;
}

void ora_np_step(const ora_np_cfg *c, ora_np_state *s, const float *actions, double *reward_out,
                 uint8_t *reward_kind_out, uint8_t *flags_out, float *obs)
{
#pragma omp parallel for schedule(static)
    for (int n = 0; n < c->n_envs; ++n) np_step_one(c, s, n, actions, reward_out, reward_kind_out, flags_out, obs);
}

/* ======================================================================================= */
/* A3: StockPortfolioEnv — finrl/meta/env_portfolio_allocation/env_portfolio.py              */
/* ======================================================================================= */

void ora_portfolio_reset(const ora_portfolio_cfg *c, ora_portfolio_state *s, const uint8_t *mask)
{
    for (int n = 0; n < c->n_envs; ++n) { /* reset (:202-220): no stale-day quirk here */
        if (mask && !mask[n]) continue;
        s->pv[n] = c->initial_amount;
        s->day[n] = 0;
    }
}

void ora_portfolio_obs(const ora_portfolio_cfg *c, const ora_portfolio_state *s, double *obs)
{
    const int D = c->stock_dim, K = c->n_tech, T = c->n_days;
    const size_t O = (size_t)(D + K) * D;
    for (int n = 0; n < c->n_envs; ++n) {
        const int day = s->day[n];
        double *o = obs + (size_t)n * O;
        memcpy(o, c->cov + (size_t)day * D * D, sizeof(double) * (size_t)D * D);
        for (int k = 0; k < K; ++k)
            memcpy(o + (size_t)(D + k) * D, c->tech + ((size_t)k * T + day) * D, sizeof(double) * (size_t)D);
    }
}

static void pf_step_one(const ora_portfolio_cfg *c, ora_portfolio_state *s, int n, const void *actions,
                        int actions_f64, double *reward_out, uint8_t *flags_out, double *weights_out,
                        double *pret_out, int auto_reset)
{
    const int D = c->stock_dim, T = c->n_days;
    if (s->day[n] >= T - 1) { /* terminal branch (:127-156): previous reward again */
        if (reward_out) reward_out[n] = s->reward[n];
        if (flags_out) flags_out[n] = ORA_FLAG_DONE;
        if (pret_out) pret_out[n] = 0.0;
        if (auto_reset) {
            s->pv[n] = c->initial_amount;
            s->day[n] = 0;
        }
        return;
    }
    /* softmax_normalization (:225-229): exp in the input dtype, np.sum pairwise, no max-subtraction (Q8) */
    double w[MAXD];
    if (actions_f64) {
        const double *a = (const double *)actions + (size_t)n * D;
        double e[MAXD];
        for (int i = 0; i < D; ++i) e[i] = exp(a[i]);
        const double den = ora_pairwise_sum_f64(e, D);
        for (int i = 0; i < D; ++i) w[i] = e[i] / den;
    } else {
        const float *a = (const float *)actions + (size_t)n * D;
        float e[MAXD];
        for (int i = 0; i < D; ++i) e[i] = expf(a[i]);
        const float den = ora_pairwise_sum_f32(e, D);
        for (int i = 0; i < D; ++i) w[i] = (double)(e[i] / den);
    }
    const double *c0 = c->close + (size_t)s->day[n] * D;
    s->day[n] += 1;
    const double *c1 = c->close + (size_t)s->day[n] * D;
    /* portfolio_return = sum(((close_new / close_old) - 1) * weights): Python sum, sequential (:183-185) */
    double pr = 0.0;
    for (int i = 0; i < D; ++i) pr = pr + ((c1[i] / c0[i]) - 1.0) * w[i];
    const double pv = s->pv[n] * (1.0 + pr);
    s->pv[n] = pv;
    s->reward[n] = pv; /* reward = new portfolio value, unscaled (:196) */
    if (reward_out) reward_out[n] = pv;
    if (flags_out) flags_out[n] = 0;
    if (pret_out) pret_out[n] = pr;
    if (weights_out)
        for (int i = 0; i < D; ++i) weights_out[(size_t)n * D + i] = w[i];
}

void ora_portfolio_step(const ora_portfolio_cfg *c, ora_portfolio_state *s, const void *actions, int actions_f64,
                        double *reward_out, uint8_t *flags_out, double *weights_out, double *pret_out, int auto_reset)
{
#pragma omp parallel for schedule(static)
    for (int n = 0; n < c->n_envs; ++n)
        pf_step_one(c, s, n, actions, actions_f64, reward_out, flags_out, weights_out, pret_out, auto_reset);
}

/* ======================================================================================= */
/* A4: StockTradingEnvCashpenalty — finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py */
/* ======================================================================================= */

static void cp_reset_one(const ora_cp_cfg *c, ora_cp_state *s, int n, int32_t start)
{
    /* reset (:132-158) with random_start=False or a caller-supplied starting point */
    s->cash[n] = c->initial_amount;
    for (int i = 0; i < c->stock_dim; ++i) s->hold[(size_t)n * c->stock_dim + i] = 0.0;
    s->date_index[n] = start;
    s->start[n] = start;
    s->fresh[n] = 1;
    s->sum_trades[n] = 0.0;
    /* account_information is emptied; last_* are only read after a step has logged them */
    s->last_cash[n] = 0.0;
    s->last_total[n] = 0.0;
}

void ora_cp_reset(const ora_cp_cfg *c, ora_cp_state *s, const uint8_t *mask, const int32_t *start_points)
{
    for (int n = 0; n < c->n_envs; ++n)
        if (!mask || mask[n]) cp_reset_one(c, s, n, start_points ? start_points[n] : 0);
}

void ora_cp_obs(const ora_cp_cfg *c, const ora_cp_state *s, double *obs)
{
    const int D = c->stock_dim, DC = D * c->n_cols, O = 1 + D + DC;
    for (int n = 0; n < c->n_envs; ++n) {
        double *o = obs + (size_t)n * O;
        o[0] = s->cash[n];
        for (int i = 0; i < D; ++i) o[1 + i] = s->hold[(size_t)n * D + i];
        memcpy(o + 1 + D, c->info + (size_t)s->date_index[n] * DC, sizeof(double) * (size_t)DC);
    }
}

/* get_reward (:246-256) from the last logged (total_assets, cash) pair */
static double cp_reward(const ora_cp_cfg *c, double assets, double cash, int current_step)
{
    if (current_step == 0) return 0.0;
    double pen = assets * c->cash_penalty_proportion - cash;
    if (!(pen > 0.0)) pen = 0.0; /* max(0, x) */
    assets -= pen;
    double r = (assets / c->initial_amount) - 1;
    r /= current_step;
    return r;
}

static int64_t cp_floordiv_i64(int64_t a, int64_t b)
{
    int64_t q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q -= 1;
    return q;
}

static void cp_step_one(const ora_cp_cfg *c, ora_cp_state *s, int n, const void *actions, int actions_f64,
                        double *reward_out, uint8_t *flags_out, int auto_reset)
{
    const int D = c->stock_dim, T = c->n_days;
    double *hold = s->hold + (size_t)n * D;
    uint8_t flags = 0;
    /* self.sum_trades += np.sum(np.abs(actions)) (:302): summed in the action dtype (pairwise), then added */
    {
        if (actions_f64) {
            double tmp[MAXD];
            for (int i = 0; i < D; ++i) tmp[i] = fabs(((const double *)actions)[(size_t)n * D + i]);
            s->sum_trades[n] += ora_pairwise_sum_f64(tmp, D);
        } else {
            float tmp[MAXD];
            for (int i = 0; i < D; ++i) tmp[i] = fabsf(((const float *)actions)[(size_t)n * D + i]);
            s->sum_trades[n] += (double)ora_pairwise_sum_f32(tmp, D);
        }
    }
    const int di = s->date_index[n];
    const int current_step = di - s->start[n];
    if (di == T - 1) { /* last date (:308-310): reward from the previously logged pair, state unchanged */
        flags = ORA_FLAG_DONE;
        if (reward_out) reward_out[n] = cp_reward(c, s->last_total[n], s->last_cash[n], current_step);
        if (flags_out) flags_out[n] = flags;
        if (auto_reset) cp_reset_one(c, s, n, 0);
        return;
    }
    const double *close = c->close + (size_t)di * D;
    const double begin_cash = s->cash[n];
    double asset_value = 0.0; /* np.dot(holdings, closings) (:319) */
    for (int i = 0; i < D; ++i) asset_value += hold[i] * close[i];
    s->last_cash[n] = begin_cash;
    s->last_total[n] = begin_cash + asset_value;
    const double reward = cp_reward(c, s->last_total[n], s->last_cash[n], current_step); /* before trading (:326) */

    /* get_transactions (:258-298) */
    double tx[MAXD];
    const double turbulence = s->fresh[n] ? 0.0 : c->turb[di];
    const int liq = c->use_turbulence && turbulence >= c->turbulence_threshold;
    for (int i = 0; i < D; ++i) {
        double a; /* actions * hmax in the input dtype */
        /* scalar hmax: weak Python float, the product stays in the action dtype; array hmax: numpy array-array
           promotion (f32 * f64 -> f64, f32 * f32 -> f32) */
        if (actions_f64)
            a = ((const double *)actions)[(size_t)n * D + i] * (c->hmax_vec ? c->hmax_vec[i] : c->hmax);
        else if (c->hmax_vec && !c->hmax_vec_f32)
            a = (double)((const float *)actions)[(size_t)n * D + i] * c->hmax_vec[i];
        else
            a = (double)(((const float *)actions)[(size_t)n * D + i] * (float)(c->hmax_vec ? c->hmax_vec[i] : c->hmax));
        if (!(close[i] > 0)) a = 0.0; /* np.where(closings > 0, actions, 0) */
        if (c->discrete_actions) {
            /* actions // closings (float floor-div), astype(int), then toward-zero multiples of the increment */
            int64_t q = (int64_t)ora_floor_divide_f64(a, close[i]);
            const int64_t inc = c->shares_increment;
            q = (q >= 0) ? cp_floordiv_i64(q, inc) * inc : cp_floordiv_i64(q + inc, inc) * inc;
            a = (double)q;
        } else {
            a = a / close[i];
        }
        tx[i] = (a > -hold[i]) ? a : -hold[i]; /* np.maximum(actions, -holdings) */
    }
    if (liq) {
        flags |= ORA_FLAG_LIQUIDATE;
        for (int i = 0; i < D; ++i) tx[i] = -hold[i];
    }
    /* proceeds / spend (:333-340) */
    double proceeds = 0.0, spend = 0.0;
    for (int i = 0; i < D; ++i) {
        const double sell = -(tx[i] < 0 ? tx[i] : 0.0);
        proceeds += sell * close[i];
    }
    double costs = proceeds * c->sell_cost_pct;
    double coh = begin_cash + proceeds;
    for (int i = 0; i < D; ++i) {
        const double buy = tx[i] > 0 ? tx[i] : 0.0;
        spend += buy * close[i];
    }
    costs += spend * c->buy_cost_pct;
    if ((spend + costs) > coh) {
        flags |= ORA_FLAG_SHORTAGE;
        if (c->patient) { /* don't buy; NOTE the sell costs are dropped too (quirk Q9) */
            for (int i = 0; i < D; ++i)
                if (tx[i] > 0) tx[i] = 0.0;
            spend = 0.0;
            costs = 0.0;
        } else { /* CASH SHORTAGE termination (:349-353): unchanged state, current reward */
            flags |= ORA_FLAG_DONE;
            if (reward_out) reward_out[n] = reward;
            if (flags_out) flags_out[n] = flags;
            if (auto_reset) cp_reset_one(c, s, n, 0);
            return;
        }
    }
    coh = coh - spend - costs;
    for (int i = 0; i < D; ++i) hold[i] = hold[i] + tx[i];
    s->cash[n] = coh;
    s->date_index[n] = di + 1;
    s->fresh[n] = c->use_turbulence ? 0 : 1; /* self.turbulence is only refreshed when a threshold is set */
    if (reward_out) reward_out[n] = reward;
    if (flags_out) flags_out[n] = flags;
}

void ora_cp_step(const ora_cp_cfg *c, ora_cp_state *s, const void *actions, int actions_f64, double *reward_out,
                 uint8_t *flags_out, int auto_reset)
{
#pragma omp parallel for schedule(static)
    for (int n = 0; n < c->n_envs; ++n) cp_step_one(c, s, n, actions, actions_f64, reward_out, flags_out, auto_reset);
}

/* ======================================================================================= */
/* sibling: CryptoEnv — finrl/meta/env_cryptocurrency_trading/env_multiple_crypto.py         */
/* ======================================================================================= */
/* Fractional positions in a float32 `stocks` array against float64 prices and a float64 cash:
 * every product with the price promotes to float64; min() keeps the dtype of whichever side wins. */

static double crypto_total(const ora_crypto_cfg *c, const float *stocks, const double *price, double cash)
{
    double prod[MAXD];
    for (int i = 0; i < c->stock_dim; ++i) prod[i] = (double)stocks[i] * price[i];
    return cash + ora_pairwise_sum_f64(prod, c->stock_dim); /* cash + (stocks * price).sum() */
}

void ora_crypto_reset(const ora_crypto_cfg *c, ora_crypto_state *s, const uint8_t *mask)
{
    for (int n = 0; n < c->n_envs; ++n) { /* reset (:47-57); gamma_return is NOT reset by the reference */
        if (mask && !mask[n]) continue;
        s->time[n] = c->lookback - 1;
        s->cash[n] = c->initial_capital;
        for (int i = 0; i < c->stock_dim; ++i) s->stocks[(size_t)n * c->stock_dim + i] = 0.0f;
        s->total[n] = crypto_total(c, s->stocks + (size_t)n * c->stock_dim, c->price + (size_t)s->time[n] * c->stock_dim,
                                   s->cash[n]);
    }
}

static void crypto_obs_one(const ora_crypto_cfg *c, const ora_crypto_state *s, int n, float *obs)
{
    /* get_state (:93-99): hstack((cash*2**-18, stocks*2**-3)) is float64, each tech row * 2**-15 is appended
     * and the whole vector is cast to float32 */
    const int D = c->stock_dim, TD = c->tech_dim;
    obs[0] = (float)(s->cash[n] * 3.814697265625e-06);
    for (int i = 0; i < D; ++i) obs[1 + i] = (float)((double)(s->stocks[(size_t)n * D + i] * 0.125f));
    for (int l = 0; l < c->lookback; ++l) {
        const double *row = c->tech + (size_t)(s->time[n] - l) * TD;
        for (int i = 0; i < TD; ++i) obs[1 + D + l * TD + i] = (float)(row[i] * 3.0517578125e-05);
    }
}

void ora_crypto_obs(const ora_crypto_cfg *c, const ora_crypto_state *s, float *obs)
{
    const int O = 1 + c->stock_dim + c->tech_dim * c->lookback;
    for (int n = 0; n < c->n_envs; ++n) crypto_obs_one(c, s, n, obs + (size_t)n * O);
}

static void crypto_step_one(const ora_crypto_cfg *c, ora_crypto_state *s, int n, const void *actions, int actions_f64,
                            double *reward_out, uint8_t *flags_out, float *obs)
{
    const int D = c->stock_dim, O = 1 + D + c->tech_dim * c->lookback;
    float *stocks = s->stocks + (size_t)n * D;
    const int max_step = c->n_days - c->lookback - 1;
    if (s->time[n] >= c->n_days - 1) { /* past the data (the reference would raise IndexError): inert */
        if (reward_out) reward_out[n] = 0.0;
        if (flags_out) flags_out[n] = ORA_FLAG_DONE;
        if (obs) crypto_obs_one(c, s, n, obs + (size_t)n * O);
        return;
    }
    s->time[n] += 1;
    const double *price = c->price + (size_t)s->time[n] * D;
    /* actions[i] = actions[i] * norm_vector_i, in place, in the action dtype (:62-64) */
    double a[MAXD];
    for (int i = 0; i < D; ++i) {
        if (actions_f64)
            a[i] = ((const double *)actions)[(size_t)n * D + i] * c->act_norm[i];
        else
            a[i] = (double)(((const float *)actions)[(size_t)n * D + i] * (float)c->act_norm[i]);
    }
    double cash = s->cash[n];
    for (int i = 0; i < D; ++i) { /* sells (:66-70) */
        if (a[i] < 0 && price[i] > 0) {
            double nsh; /* min(stocks, -action) -> -action iff it is smaller */
            if (-a[i] < (double)stocks[i]) {
                nsh = -a[i]; /* dtype of the action: f32 - f32 stays f32, f32 - f64 is f64 then cast back */
                stocks[i] = actions_f64 ? (float)((double)stocks[i] - nsh) : (float)(stocks[i] - (float)nsh);
            } else {
                nsh = (double)stocks[i];
                stocks[i] = stocks[i] - stocks[i];
            }
            cash += price[i] * nsh * (1 - c->sell_cost_pct);
        }
    }
    for (int i = 0; i < D; ++i) { /* buys (:72-76) */
        if (a[i] > 0 && price[i] > 0) {
            const double avail = ora_floor_divide_f64(cash, price[i]);
            double nsh;
            if (a[i] < avail) { /* min(avail, action) -> the action */
                nsh = a[i];
                stocks[i] = actions_f64 ? (float)((double)stocks[i] + nsh) : (float)(stocks[i] + (float)nsh);
            } else {
                nsh = avail;
                stocks[i] = (float)((double)stocks[i] + nsh);
            }
            cash -= price[i] * nsh * (1 + c->buy_cost_pct);
        }
    }
    s->cash[n] = cash;
    const int done = s->time[n] == max_step;
    if (obs) crypto_obs_one(c, s, n, obs + (size_t)n * O);
    const double next_total = crypto_total(c, stocks, price, cash);
    double reward = (next_total - s->total[n]) * 1.52587890625e-05; /* 2 ** -16 */
    s->total[n] = next_total;
    s->gamma_return[n] = s->gamma_return[n] * c->gamma + reward;
    if (done) {
        reward = s->gamma_return[n];
        s->episode_return[n] = s->total[n] / c->initial_capital;
    }
    if (reward_out) reward_out[n] = reward;
    if (flags_out) flags_out[n] = done ? ORA_FLAG_DONE : 0;
}

void ora_crypto_step(const ora_crypto_cfg *c, ora_crypto_state *s, const void *actions, int actions_f64,
                     double *reward_out, uint8_t *flags_out, float *obs)
{
#pragma omp parallel for schedule(static)
    for (int n = 0; n < c->n_envs; ++n) crypto_step_one(c, s, n, actions, actions_f64, reward_out, flags_out, obs);
}

/* ======================================================================================= */
/* sibling: StockTradingEnvStopLoss — finrl/meta/env_stock_trading/env_stocktrading_stoploss.py */
/* ======================================================================================= */

static void sl_reset_one(const ora_sl_cfg *c, ora_sl_state *s, int n, int32_t start)
{
    const int D = c->stock_dim; /* reset (:134-165) */
    s->cash[n] = c->initial_amount;
    for (int i = 0; i < D; ++i) {
        const size_t k = (size_t)n * D + i;
        s->hold[k] = s->prev_hold[k] = s->avg_buy[k] = s->n_buys[k] = s->cdiff[k] = s->pdiff[k] = 0.0;
    }
    s->date_index[n] = start;
    s->start[n] = start;
    s->fresh[n] = 1;
    s->sum_trades[n] = 0.0;
    s->last_cash[n] = 0.0;
    s->last_total[n] = 0.0;
}

void ora_sl_reset(const ora_sl_cfg *c, ora_sl_state *s, const uint8_t *mask, const int32_t *start_points)
{
    for (int n = 0; n < c->n_envs; ++n)
        if (!mask || mask[n]) sl_reset_one(c, s, n, start_points ? start_points[n] : 0);
}

void ora_sl_obs(const ora_sl_cfg *c, const ora_sl_state *s, double *obs)
{
    const int D = c->stock_dim, DC = D * c->n_cols, O = 1 + D + DC;
    for (int n = 0; n < c->n_envs; ++n) {
        double *o = obs + (size_t)n * O;
        o[0] = s->cash[n];
        for (int i = 0; i < D; ++i) o[1 + i] = s->hold[(size_t)n * D + i];
        memcpy(o + 1 + D, c->info + (size_t)s->date_index[n] * DC, sizeof(double) * (size_t)DC);
    }
}

/* get_reward (:255-290) with the arrays as they stand when it is called */
static double sl_reward(const ora_sl_cfg *c, const ora_sl_state *s, int n, int current_step, const double *cdiff)
{
    if (current_step == 0) return 0.0;
    const int D = c->stock_dim;
    const double *hold = s->hold + (size_t)n * D, *prev = s->prev_hold + (size_t)n * D, *pdiff = s->pdiff + (size_t)n * D;
    const double total_assets = s->last_total[n], cash = s->last_cash[n];
    double cash_penalty = total_assets * c->cash_penalty_proportion - cash;
    if (!(cash_penalty > 0.0)) cash_penalty = 0.0;
    double stop_loss_penalty = 0.0;
    if (current_step > 1) {
        double d = 0.0;
        for (int i = 0; i < D; ++i) d += prev[i] * (cdiff[i] < 0 ? cdiff[i] : 0.0);
        stop_loss_penalty = -1 * d;
    }
    double lp = 0.0, ar = 0.0;
    for (int i = 0; i < D; ++i) {
        lp += hold[i] * (pdiff[i] < 0 ? pdiff[i] : 0.0);
        ar += hold[i] * (pdiff[i] > 0 ? pdiff[i] : 0.0);
    }
    const double low_profit_penalty = -1 * lp;
    const double total_penalty = cash_penalty + stop_loss_penalty + low_profit_penalty;
    double reward = ((total_assets - total_penalty + ar) / c->initial_amount) - 1;
    reward /= current_step;
    return reward;
}

static void sl_step_one(const ora_sl_cfg *c, ora_sl_state *s, int n, const void *actions, int actions_f64,
                        double *reward_out, uint8_t *flags_out, int auto_reset)
{
    const int D = c->stock_dim, T = c->n_days;
    const size_t base = (size_t)n * D;
    double *hold = s->hold + base, *prev = s->prev_hold + base, *avg = s->avg_buy + base, *nb = s->n_buys + base;
    double *cdiff = s->cdiff + base, *pdiff = s->pdiff + base;
    uint8_t flags = 0;
    {   /* self.sum_trades += np.sum(np.abs(actions)) (:294) */
        if (actions_f64) {
            double tmp[MAXD];
            for (int i = 0; i < D; ++i) tmp[i] = fabs(((const double *)actions)[base + i]);
            s->sum_trades[n] += ora_pairwise_sum_f64(tmp, D);
        } else {
            float tmp[MAXD];
            for (int i = 0; i < D; ++i) tmp[i] = fabsf(((const float *)actions)[base + i]);
            s->sum_trades[n] += (double)ora_pairwise_sum_f32(tmp, D);
        }
    }
    const int di = s->date_index[n];
    const int current_step = di - s->start[n];
    if (di == T - 1) { /* last date (:302-304) */
        if (reward_out) reward_out[n] = sl_reward(c, s, n, current_step, cdiff);
        if (flags_out) flags_out[n] = ORA_FLAG_DONE;
        if (auto_reset) sl_reset_one(c, s, n, 0);
        return;
    }
    const double *close = c->close + (size_t)di * D;
    const double begin_cash = s->cash[n];
    double asset_value = 0.0;
    for (int i = 0; i < D; ++i) asset_value += hold[i] * close[i];
    /* reward from the PREVIOUS log entry and the previous step's penalty arrays, then log (:313-319) */
    const double reward = sl_reward(c, s, n, current_step, cdiff);
    s->last_cash[n] = begin_cash;
    s->last_total[n] = begin_cash + asset_value;

    double tx[MAXD], new_cdiff[MAXD];
    const double turbulence = s->fresh[n] ? 0.0 : c->turb[di];
    const int liq = c->use_turbulence && turbulence >= c->turbulence_threshold;
    if (liq) flags |= ORA_FLAG_LIQUIDATE;
    const int stop_on = begin_cash >= c->stoploss_penalty * c->initial_amount;
    for (int i = 0; i < D; ++i) {
        double a;
        if (actions_f64)
            a = ((const double *)actions)[base + i] * (c->hmax_vec ? c->hmax_vec[i] : c->hmax);
        else if (c->hmax_vec && !c->hmax_vec_f32)
            a = (double)((const float *)actions)[base + i] * c->hmax_vec[i];
        else
            a = (double)(((const float *)actions)[base + i] * (float)(c->hmax_vec ? c->hmax_vec[i] : c->hmax));
        if (!(close[i] > 0)) a = 0.0;
        if (liq) a = -(hold[i] * close[i]); /* currency, divided by the price again below (:331-334) */
        if (c->discrete_actions) {
            int64_t q = (close[i] > 0) ? (int64_t)ora_floor_divide_f64(a, close[i]) : 0;
            const int64_t inc = c->shares_increment;
            q = (q >= 0) ? cp_floordiv_i64(q, inc) * inc : cp_floordiv_i64(q + inc, inc) * inc;
            a = (double)q;
        } else {
            a = (close[i] > 0) ? a / close[i] : 0.0;
        }
        a = (a > -hold[i]) ? a : -hold[i];
        new_cdiff[i] = close[i] - (c->stoploss_penalty * avg[i]); /* closing_diff_avg_buy (:354-356) */
        if (stop_on && new_cdiff[i] < 0) a = -hold[i];            /* stop-loss: clear the position (:357-361) */
        tx[i] = a;
    }
    for (int i = 0; i < D; ++i) cdiff[i] = new_cdiff[i];
    double proceeds = 0.0, spend = 0.0;
    double sells[MAXD], buys[MAXD];
    for (int i = 0; i < D; ++i) {
        sells[i] = -(tx[i] < 0 ? tx[i] : 0.0);
        buys[i] = tx[i] > 0 ? tx[i] : 0.0;
    }
    for (int i = 0; i < D; ++i) proceeds += sells[i] * close[i];
    double costs = proceeds * c->sell_cost_pct;
    double coh = begin_cash + proceeds;
    for (int i = 0; i < D; ++i) spend += buys[i] * close[i];
    costs += spend * c->buy_cost_pct;
    if ((spend + costs) > coh) {
        flags |= ORA_FLAG_SHORTAGE;
        if (c->patient) {
            for (int i = 0; i < D; ++i)
                if (tx[i] > 0) tx[i] = 0.0;
            spend = 0.0;
            costs = 0.0;
        } else { /* terminate (:383-386): get_reward() again, now with this step's log entry and closing diff */
            flags |= ORA_FLAG_DONE;
            if (reward_out) reward_out[n] = sl_reward(c, s, n, current_step, cdiff);
            if (flags_out) flags_out[n] = flags;
            if (auto_reset) sl_reset_one(c, s, n, 0);
            return;
        }
    }
    /* profitable sells (:391-404); `sells` and `buys` are the PRE-patient vectors */
    for (int i = 0; i < D; ++i) {
        const double scp = sells[i] > 0 ? close[i] : 0.0;
        const int profit = (scp - avg[i]) > 0;
        pdiff[i] = profit ? close[i] - ((1 + c->profit_loss_ratio * (1 - c->stoploss_penalty)) * avg[i]) : 0.0;
    }
    coh = coh - spend - costs;
    for (int i = 0; i < D; ++i) {
        const double hn = hold[i] + tx[i];
        const double sg = buys[i] > 0 ? 1.0 : 0.0; /* np.sign(buys) */
        nb[i] += sg;
        if (sg > 0) avg[i] = avg[i] + ((close[i] - avg[i]) / nb[i]);
        if (!(hn > 0)) {
            nb[i] = 0.0;
            avg[i] = 0.0;
        }
        prev[i] = hold[i];
        hold[i] = hn;
    }
    s->cash[n] = coh;
    s->date_index[n] = di + 1;
    s->fresh[n] = c->use_turbulence ? 0 : 1;
    if (reward_out) reward_out[n] = reward;
    if (flags_out) flags_out[n] = flags;
}

void ora_sl_step(const ora_sl_cfg *c, ora_sl_state *s, const void *actions, int actions_f64, double *reward_out,
                 uint8_t *flags_out, int auto_reset)
{
#pragma omp parallel for schedule(static)
    for (int n = 0; n < c->n_envs; ++n) sl_step_one(c, s, n, actions, actions_f64, reward_out, flags_out, auto_reset);
}
