/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the FinRL env step path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (finrl_b200/) never links, imports or calls it.
 *
 * Every function restates an algorithm of the reference (superyuri/FinRL) and cites the
 * file:line it follows (paths relative to /root/reference).  Parity is PINNED: the oracle is
 * checked against golden vectors produced by executing the unmodified reference
 * (tests/golden/make_golden.py, tests/test_oracle_golden.py).
 *
 * Layout is the natural row-major one ([env][stock]); it deliberately differs from the
 * stock-major device layout of the CUDA engine so that a layout bug cannot cancel out.
 */
#ifndef FINRL_ORACLE_H
#define FINRL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_FLAG_DONE 1u
#define ORA_FLAG_LIQUIDATE 2u
#define ORA_FLAG_SHORTAGE 4u

/* OpenMP threads used by the *_step loops over envs; returns the count in effect. */
int ora_set_threads(int n);

/* ---- building blocks ------------------------------------------------------------------ */

/* numpy's float floor-division (npy_divmod), used by `cash // (price*(1+cost))`
 * (finrl/meta/env_stock_trading/env_stocktrading.py:178-180). */
double ora_floor_divide_f64(double a, double b);
float ora_floor_divide_f32(float a, float b);

/* numpy's default (unstable, SIMD bitonic-network) argsort tie order for n <= 256 int64 keys,
 * as used by `np.argsort(actions)` (env_stocktrading.py:317).  SURVEY.md H1. */
void ora_argsort_i64(const int64_t *keys, int n, int32_t *order_out);

/* numpy pairwise float32 summation of x[0..n) (ndarray.sum on a contiguous f32 vector). */
float ora_pairwise_sum_f32(const float *x, int n);
double ora_pairwise_sum_f64(const double *x, int n);

/* ---- A1: StockTradingEnv (env_stocktrading.py) ----------------------------------------- */

typedef struct {
    int32_t n_envs, stock_dim, n_tech, n_days;
    double hmax;
    double initial_amount;
    double buy_cost_pct, sell_cost_pct;
    double reward_scaling;
    int32_t use_turbulence; /* turbulence_threshold is not None */
    double turbulence_threshold;
    const double *close;     /* [T][D] */
    const double *tech;      /* [K][T][D] */
    const double *risk;      /* [T] (risk_indicator_col) */
    const int32_t *init_hold; /* [D] num_stock_shares (or previous_state holdings) */
} ora_trading_cfg;

typedef struct {
    double *cash;     /* [N] */
    int32_t *hold;    /* [N][D] */
    int32_t *day;     /* [N] */
    int32_t *sday;    /* [N] day whose prices/tech sit in the state list; < 0 means "fresh":
                         prices of day (-sday-1), turbulence == 0 (after ctor / reset) */
    double *cost;     /* [N] */
    int32_t *trades;  /* [N] */
    double *reward;   /* [N] last scaled reward (returned again by the terminal no-op step) */
    int32_t *episode; /* [N] */
} ora_trading_state;

/* ctor semantics (env_stocktrading.py:48-100): day=day0, state from day0, fresh. */
void ora_trading_init(const ora_trading_cfg *c, ora_trading_state *s, int32_t day0);
/* reset semantics incl. the stale-day quirk (env_stocktrading.py:359-393). mask may be NULL. */
void ora_trading_reset(const ora_trading_cfg *c, ora_trading_state *s, const uint8_t *mask);
/* observation = float32 cast of the reference's state list (what DummyVecEnv hands an agent). */
void ora_trading_obs(const ora_trading_cfg *c, const ora_trading_state *s, float *obs /*[N][O]*/);
/* one step() for every env (env_stocktrading.py:220-357).  actions [N][D] f32 or f64.
 * reward_out/flags_out/obs may be NULL.  auto_reset applies DummyVecEnv's reset-on-done. */
void ora_trading_step(const ora_trading_cfg *c, ora_trading_state *s, const void *actions,
                      int actions_f64, double *reward_out, uint8_t *flags_out, float *obs,
                      int auto_reset);

/* ---- A2: numpy / ElegantRL StockTradingEnv (env_stocktrading_np.py) ---------------------- */

/* numpy scalar "kind" a Python-level variable carries under NEP 50 (numpy >= 2):
 * 0 = Python float (weak), 1 = np.float32, 2 = np.float64.  Values are stored in doubles. */
#define ORA_KIND_PY 0
#define ORA_KIND_F32 1
#define ORA_KIND_F64 2

typedef struct {
    int32_t n_envs, stock_dim, tech_dim /* columns of tech_ary = D*K */, n_days;
    double gamma, max_stock, min_stock_rate, buy_cost_pct, sell_cost_pct, reward_scaling;
    double initial_capital;
    double obs_amount_floor; /* get_state shows max(amount, floor): 1e4 in StockEnvNAS100
                                (env_nas100_wrds.py:157); -inf for env_stocktrading_np */
    const float *price;      /* [T][D]  price_ary  (:27) */
    const float *tech;       /* [T][tech_dim] tech_ary = f32(tech_array) * 2^-7 (:28,31) */
    const float *turb_bool;  /* [T] (:32) */
    const float *turb_ary;   /* [T] (:33-35) */
    const float *init_stocks; /* [D] initial_stocks */
} ora_np_cfg;

typedef struct {
    double *amount;      uint8_t *amount_kind;   /* [N] self.amount */
    float *stocks;       /* [N][D] */
    float *cool;         /* [N][D] stocks_cool_down */
    int32_t *day;        /* [N] */
    double *total;       uint8_t *total_kind;    /* [N] self.total_asset */
    double *gamma_reward; uint8_t *gr_kind;      /* [N] */
    double *init_total;  /* [N] initial_total_asset */
    double *episode_return; /* [N] */
} ora_np_state;

/* reset() (:80-101).  stocks0 [N][D] / factor [N] non-NULL = the if_train branch with the random
 * draws supplied by the caller (rd.randint(0,64,D) already added to initial_stocks; rd.uniform). */
void ora_np_reset(const ora_np_cfg *c, ora_np_state *s, const uint8_t *mask, const float *stocks0,
                  const double *factor);
void ora_np_obs(const ora_np_cfg *c, const ora_np_state *s, float *obs /*[N][O]*/);
/* step() (:103-147): reward_out f64 value + kind, flags ORA_FLAG_DONE|ORA_FLAG_LIQUIDATE. */
void ora_np_step(const ora_np_cfg *c, ora_np_state *s, const float *actions, double *reward_out,
                 uint8_t *reward_kind_out, uint8_t *flags_out, float *obs);

/* ---- A3: StockPortfolioEnv (env_portfolio_allocation/env_portfolio.py) --------------------- */
typedef struct {
    int32_t n_envs, stock_dim, n_tech, n_days;
    double initial_amount;
    const double *close; /* [T][D] */
    const double *cov;   /* [T][D][D] cov_list of each day */
    const double *tech;  /* [K][T][D] */
} ora_portfolio_cfg;

typedef struct {
    double *pv;     /* [N] portfolio_value */
    int32_t *day;   /* [N] */
    double *reward; /* [N] self.reward (returned again by the terminal step) */
} ora_portfolio_state;

void ora_portfolio_reset(const ora_portfolio_cfg *c, ora_portfolio_state *s, const uint8_t *mask);
/* state = np.append(cov, tech rows, axis=0): obs[N][(D+K)*D] f64 */
void ora_portfolio_obs(const ora_portfolio_cfg *c, const ora_portfolio_state *s, double *obs);
/* step (:125-200). actions [N][D] f32 or f64 (np.exp runs in that dtype).  weights_out [N][D] f64
 * (nullable) receives the softmax weights, pret_out [N] the portfolio return. */
void ora_portfolio_step(const ora_portfolio_cfg *c, ora_portfolio_state *s, const void *actions, int actions_f64,
                        double *reward_out, uint8_t *flags_out, double *weights_out, double *pret_out, int auto_reset);

/* ---- A4: StockTradingEnvCashpenalty (env_stocktrading_cashpenalty.py) --------------------- */
typedef struct {
    int32_t n_envs, stock_dim, n_cols /* len(daily_information_cols) */, n_days;
    double buy_cost_pct, sell_cost_pct, hmax;
    int32_t discrete_actions, shares_increment;
    int32_t use_turbulence;
    double turbulence_threshold;
    double initial_amount, cash_penalty_proportion;
    int32_t patient;
    const double *close; /* [T][D] */
    const double *turb;  /* [T] "turbulence" column (only read when use_turbulence) */
    const double *info;  /* [T][D*C] get_date_vector(date): asset-major daily information */
    const double *hmax_vec; /* [D] or NULL: per-asset hmax array (`actions * self.hmax` broadcasts, :268) */
    int32_t hmax_vec_f32;   /* the array is float32: the product with float32 actions stays float32 */
} ora_cp_cfg;

typedef struct {
    double *cash;        /* [N] state_memory[-1][0] */
    double *hold;        /* [N][D] */
    int32_t *date_index; /* [N] */
    int32_t *start;      /* [N] starting_point */
    uint8_t *fresh;      /* [N] 1 after reset: self.turbulence == 0 until the first completed step */
    double *last_cash;   /* [N] account_information["cash"][-1] */
    double *last_total;  /* [N] account_information["total_assets"][-1] */
    double *sum_trades;  /* [N] */
} ora_cp_state;

void ora_cp_reset(const ora_cp_cfg *c, ora_cp_state *s, const uint8_t *mask, const int32_t *start_points);
void ora_cp_obs(const ora_cp_cfg *c, const ora_cp_state *s, double *obs /*[N][1+D+D*C]*/);
void ora_cp_step(const ora_cp_cfg *c, ora_cp_state *s, const void *actions, int actions_f64, double *reward_out,
                 uint8_t *flags_out, int auto_reset);

/* ---- sibling: StockTradingEnvStopLoss (env_stocktrading_stoploss.py) ----------------------- */
typedef struct {
    int32_t n_envs, stock_dim, n_cols, n_days;
    double buy_cost_pct, sell_cost_pct, hmax;
    int32_t discrete_actions, shares_increment;
    double stoploss_penalty, profit_loss_ratio;
    int32_t use_turbulence;
    double turbulence_threshold;
    double initial_amount, cash_penalty_proportion;
    int32_t patient;
    const double *close; /* [T][D] */
    const double *turb;  /* [T] */
    const double *info;  /* [T][D*C] */
    const double *hmax_vec; /* [D] or NULL, as in ora_cp_cfg */
    int32_t hmax_vec_f32;
} ora_sl_cfg;

typedef struct {
    double *cash;        /* [N] */
    double *hold;        /* [N][D] state_memory[-1] holdings */
    double *prev_hold;   /* [N][D] state_memory[-2] holdings */
    double *avg_buy;     /* [N][D] avg_buy_price */
    double *n_buys;      /* [N][D] */
    double *cdiff;       /* [N][D] closing_diff_avg_buy */
    double *pdiff;       /* [N][D] profit_sell_diff_avg_buy */
    int32_t *date_index; /* [N] */
    int32_t *start;      /* [N] */
    uint8_t *fresh;      /* [N] */
    double *last_cash;   /* [N] account_information["cash"][-1] */
    double *last_total;  /* [N] account_information["total_assets"][-1] */
    double *sum_trades;  /* [N] */
} ora_sl_state;

void ora_sl_reset(const ora_sl_cfg *c, ora_sl_state *s, const uint8_t *mask, const int32_t *start_points);
void ora_sl_obs(const ora_sl_cfg *c, const ora_sl_state *s, double *obs /*[N][1+D+D*C]*/);
void ora_sl_step(const ora_sl_cfg *c, ora_sl_state *s, const void *actions, int actions_f64, double *reward_out,
                 uint8_t *flags_out, int auto_reset);

/* ---- sibling: CryptoEnv (env_cryptocurrency_trading/env_multiple_crypto.py) ---------------- */
typedef struct {
    int32_t n_envs, stock_dim, tech_dim, n_days, lookback;
    double initial_capital, buy_cost_pct, sell_cost_pct, gamma;
    const double *price;     /* [T][D] price_array (float64) */
    const double *tech;      /* [T][tech_dim] tech_array */
    const double *act_norm;  /* [D] action_norm_vector (:103-111) */
} ora_crypto_cfg;

typedef struct {
    double *cash;         /* [N] */
    float *stocks;        /* [N][D] */
    int32_t *time;        /* [N] */
    double *total;        /* [N] total_asset */
    double *gamma_return; /* [N] */
    double *episode_return; /* [N] */
} ora_crypto_state;

void ora_crypto_reset(const ora_crypto_cfg *c, ora_crypto_state *s, const uint8_t *mask);
void ora_crypto_obs(const ora_crypto_cfg *c, const ora_crypto_state *s, float *obs /*[N][1+D+tech_dim*lookback]*/);
void ora_crypto_step(const ora_crypto_cfg *c, ora_crypto_state *s, const void *actions, int actions_f64,
                     double *reward_out, uint8_t *flags_out, float *obs);

#ifdef __cplusplus
}
#endif
#endif
