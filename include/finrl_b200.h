/* finrl_b200 — C-ABI of the B200-native batched trading-environment engine.
 *
 * Drop-in boundary for the env step path of superyuri/FinRL (reference paths below are relative
 * to the reference checkout).  The reference has no FFI layer of its own — its boundary is the
 * gym protocol of four Python classes — so every entry point here names the Python method whose
 * body it replaces.  Host code (finrl_b200/*.py, or any other language with a C FFI) owns all
 * memory; this library allocates nothing (except the 256-byte peer-mapped statistics blocks of
 * frl_exchange_*, which must come from cudaMalloc to be exportable), keeps no global state except the
 * last error string, and only enqueues kernels on the CUDA stream it is handed.
 *
 * Conventions
 *   - every pointer inside a *_params struct is a DEVICE pointer (CUDA, sm_100a);
 *     the struct itself is passed by pointer from HOST memory and copied at call time;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; NULL = legacy default stream;
 *   - return value: 0 = success, negative = FRL_E_* (see frl_last_error());
 *   - N envs, D stocks, K technical indicators, T days, O observation length;
 *   - "stock-major" arrays are laid out [D][env_stride] (env index fastest) so that one thread
 *     per env reads and writes them coalesced.
 */
#ifndef FINRL_B200_H
#define FINRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRL_ABI_VERSION 4

#if defined(__GNUC__)
#define FRL_API __attribute__((visibility("default")))
#else
#define FRL_API
#endif

#define FRL_OK 0
#define FRL_E_INVALID (-1) /* bad argument (shape, null pointer, unsupported size) */
#define FRL_E_CUDA (-2)    /* CUDA runtime error at launch */

/* per env-step flag byte written by every step/rollout entry point */
#define FRL_FLAG_DONE 1u      /* the step returned done=True */
#define FRL_FLAG_LIQUIDATE 2u /* turbulence liquidation branch taken in this step */
#define FRL_FLAG_SHORTAGE 4u  /* cash-penalty env: CASH SHORTAGE branch taken */

/* observation emission of a rollout */
#define FRL_OBS_NONE 0
#define FRL_OBS_LAST 1 /* obs[N][O] after the final step */
#define FRL_OBS_ALL 2  /* obs[n_steps][N][O] after every step */

/* slots of the f64 statistics vector a rollout accumulates into (atomicAdd; exchanged between
 * GPUs as described under frl_stats_block — there is no reference counterpart) */
#define FRL_STAT_REWARD_SUM 0    /* sum of returned rewards over all env-steps */
#define FRL_STAT_REWARD_SQSUM 1  /* sum of squares of the same */
#define FRL_STAT_DONE_COUNT 2    /* number of done flags */
#define FRL_STAT_EPISODE_ASSET 3 /* sum of end-of-episode total assets over those dones */
#define FRL_STAT_ASSET_SUM 4     /* sum of total assets of all envs after the final step */
#define FRL_STAT_LIQ_COUNT 5     /* number of liquidation env-steps */
#define FRL_STAT_ENV_STEPS 6     /* env-steps processed */
#define FRL_STAT_TRADES 7        /* sum of trade counters after the final step */
#define FRL_N_STATS 8

/* The `stats` argument of every step/rollout entry point is one of the two accumulators, sum[0] or sum[1], of
 * an frl_stats_block (device memory, 384 bytes, 128-byte aligned, zero-initialised by the caller; NULL = no
 * statistics).  Kernels atomicAdd their partial sums into the accumulator they were handed.  With
 * n_peers == 0 that is all (callers then simply always pass sum[0]).
 * With n_peers > 0 (multi-GPU) the caller ALTERNATES between sum[0] and sum[1] on consecutive launches of a
 * stream, and the first thread block of every launch moves the OTHER accumulator — the previous launch's
 * sums, complete by stream order — out: it adds them to total[] of every rank, its own included, through the
 * peer-mapped pointers in peer_total[] (fp64 atomics over NVLink / NVSwitch) and clears them.  One-sided and
 * free of any synchronisation: no completion counter, no fence, no collective launch, and no rank ever waits
 * for a peer inside its step loop.  frl_exchange_flush moves what is still in both accumulators (the last
 * launch) before totals are read.  An NCCL all-reduce of sum[0] is the portable fallback (finrl_b200/dist.py).
 * One block serves one stream at a time. */
#define FRL_MAX_PEERS 8
#define FRL_STATS_BLOCK_BYTES 384
typedef struct frl_stats_block {
    double sum[2][FRL_N_STATS];         /* the two accumulators (`stats` = sum[0] or sum[1]) */
    double total[FRL_N_STATS];          /* exchange target: the launch sums of all ranks */
    double *peer_total[FRL_MAX_PEERS];  /* [n_peers] device pointers to every rank's total[], own included */
    uint32_t n_peers;                   /* 0 = no exchange */
    uint32_t reserved[31];
} frl_stats_block;
/* Peer-mapped statistics blocks for the exchange above (plain CUDA IPC; the processes of one node).
 * alloc: cudaMalloc + zero a block on the current device.  export: 64-byte handle another process opens.
 * open: map a peer's block into this process (enables peer access from the current device); close undoes it.
 * bind: write peer_total[] / n_peers of `block` from the blocks of all ranks (own included, any order;
 * n == 0 unbinds).  flush: one tiny kernel that moves both accumulators of `block` to the peers (call it before
 * reading total[]; a no-op for n_peers == 0).  free: release a block from alloc. */
FRL_API int32_t frl_exchange_alloc(void **block);
FRL_API int32_t frl_exchange_free(void *block);
FRL_API int32_t frl_exchange_export(const void *block, uint8_t handle[64]);
FRL_API int32_t frl_exchange_open(const uint8_t handle[64], void **peer_block);
FRL_API int32_t frl_exchange_close(void *peer_block);
FRL_API int32_t frl_exchange_bind(void *block, void *const *blocks, int32_t n, void *stream);
FRL_API int32_t frl_exchange_flush(void *block, void *stream);

FRL_API int32_t frl_abi_version(void);
/* Tuning knobs.  "trading_small_max": frl_trading_step/rollout use the low-latency 8-lanes-per-env kernel
 * for n_envs <= value and the thread-per-env kernel above (default 8192, the measured crossover; 0 = never).
 * "trading_wide_min_envs": for stock_dim > 32, batches above this many envs run in the thread-per-env kernel with
 * keys / holdings in shared memory (trading_wide.cu), smaller ones in the 8-lanes-per-env kernel (default 3072,
 * the measured crossover at D = 100).
 * "np_wide_min_d": frl_np_* stream stocks / cool-down from global memory (np_wide.cu) for stock_dim >= value
 * and keep them in registers below (1..33, default 33: the register kernel holds at most 32 stocks).
 * "np_wide_bulk" (default 1): 0 keeps the streaming kernel on its generic action staging even where the bulk-staged
 * variant applies (float32 actions, default layout, 16-byte-aligned rows, stock_dim a multiple of four).
 * "trading_wide_regs" (default 1): 0 keeps stock_dim == 100 on the generic wide kernel instead of the instantiation
 * with the stock count compiled in.  Both exist so that tests can compare the variants bit for bit. */
FRL_API int32_t frl_set_option(const char *name, int64_t value);
/* thread-local, never NULL; valid until the next failing call on this thread */
FRL_API const char *frl_last_error(void);

/* =========================================================================================
 * A1  StockTradingEnv — finrl/meta/env_stock_trading/env_stocktrading.py
 * ========================================================================================= */
typedef struct frl_trading_params {
    int32_t n_envs;     /* N >= 1 */
    int32_t stock_dim;  /* D, 1..128 (np.argsort's network rule is pinned for n <= 256).  D <= 32
                           runs in the thread-per-env kernel (or the 8-lanes-per-env one for small batches),
                           33..128 in the 8-lanes-per-env kernel for small batches and in the thread-per-env
                           kernel with shared-memory keys (trading_wide.cu) for large ones */
    int32_t n_tech;     /* K >= 0 */
    int32_t n_days;     /* T >= 1 = len(df.index.unique()) */
    int32_t obs_dim;    /* O = 1 + 2D + K*D (state_space) */
    int32_t env_stride; /* leading dimension of `hold` (>= N) */
    double hmax;
    double initial_amount;
    double buy_cost_pct, sell_cost_pct;
    double reward_scaling;
    int32_t use_turbulence; /* turbulence_threshold is not None */
    int32_t close_pitch;    /* row pitch of `close` in doubles: 32 for D <= 32, 128 for D <= 128 */
    double turbulence_threshold;
    /* ---- tables (read-only, replicated per GPU) ---- */
    const double *close;          /* [T][close_pitch] close price, rows zero-padded */
    const uint32_t *disable_mask; /* [T][close_pitch/32] bit i of the row set <=> first tech indicator of stock
                                     i == 1.0 (the "disable" flag of env_stocktrading.py:105,174) */
    const double *risk;           /* [T] risk_indicator_col (turbulence / vix) */
    const float *obs_tmpl;        /* [T][O] float32 image of the state list of day t with the
                                     cash and holdings slots zeroed */
    const int32_t *init_hold;     /* [D] num_stock_shares (or previous_state holdings) */
    /* ---- per-env state (read-write) ---- */
    double *cash;     /* [N] state[0] */
    int32_t *hold;    /* [D][env_stride] state[D+1 : 2D+1], integer-valued */
    int32_t *day;     /* [N] self.day */
    int32_t *sday;    /* [N] day whose prices/indicators sit in the state list.  >= 0: equals day,
                         turbulence = risk[sday].  < 0: "fresh" (after __init__/reset): prices of
                         day (-sday-1), turbulence = 0 (the stale-reset quirk, :359-393) */
    double *cost;     /* [N] self.cost */
    int32_t *trades;  /* [N] self.trades */
    double *reward;   /* [N] self.reward (last scaled reward; returned again by the terminal step) */
    int32_t *episode; /* [N] self.episode */
    /* ---- optional per-env output ---- */
    double *asset_out; /* [N] or NULL: total asset (cash + sum price*holding, the reference's
                          end_total_asset / asset_memory entry) of the state after the last step */
    /* ---- optional table ---- */
    const float *obs_tmpl4; /* [T][4*O] or NULL: each obs_tmpl row repeated four times (16*O bytes, 16-byte
                               aligned).  When present, a tile whose 32 envs sit on the same day gets its
                               observation rows from this image through the bulk-copy engine (TMA): load the
                               4-row image to shared memory, patch cash / holdings, one bulk store per 4 rows */
} frl_trading_params;

/* StockTradingEnv.__init__ (:24-100): day = day0, state built from day0, fresh. */
FRL_API int32_t frl_trading_init(const frl_trading_params *p, int32_t day0, void *stream);

/* StockTradingEnv.reset (:359-393) for envs with mask[n] != 0 (mask NULL = all).
 * obs (nullable) receives the [N][O] float32 observation of ALL envs afterwards. */
FRL_API int32_t frl_trading_reset(const frl_trading_params *p, const uint8_t *mask, float *obs, void *stream);

/* StockTradingEnv.render / _update_state (:395-396, :453-478): obs[N][O] float32 of the
 * current state list (the float32 cast SB3's DummyVecEnv applies). */
FRL_API int32_t frl_trading_observe(const frl_trading_params *p, float *obs, void *stream);

/* The same observation in FACTORED form for host-resident callers (137 instead of 1213 bytes per env over
 * PCIe for DOW-30): env_part[N][1+D] float32 = the env-specific slots of the state list (cash, holdings;
 * the float32 values frl_trading_observe writes) and state_day[N] = the day whose per-day row (obs_tmpl, a
 * table the host holds once) fills every other slot — day T-1 after a stale reset (quirk Q1).
 * obs[n] == obs_tmpl[state_day[n]] with slots 0 and 1+D..2D taken from env_part[n]. */
FRL_API int32_t frl_trading_observe_factored(const frl_trading_params *p, float *env_part, int32_t *state_day,
                                             void *stream);
/* HOST helper for the above (no GPU involved): expand n factored rows into dense out[n][O] float32 with
 * n_threads host threads (<= 0: hardware concurrency).  tmpl is the host copy of obs_tmpl [T][O]. */
FRL_API int32_t frl_expand_obs_host(const float *tmpl, int32_t n_days, int32_t obs_dim, int32_t stock_dim,
                                    const float *env_part, const int32_t *state_day, int64_t n, float *out,
                                    int32_t n_threads);
/* The same for rows [chunk_start[c], chunk_start[c] + chunk_count[c]) of n_chunks chunks, in order; events[c] (a
 * cudaEvent_t, or NULL; `events` itself may be NULL) is waited for before chunk c is touched, so the expansion of a
 * chunk overlaps the device -> host transfers of the following ones.  Rows leave with non-temporal stores. */
FRL_API int32_t frl_expand_obs_host_chunks(const float *tmpl, int32_t n_days, int32_t obs_dim, int32_t stock_dim,
                                           const float *env_part, const int32_t *state_day, float *out, int32_t n_chunks,
                                           const int64_t *chunk_start, const int64_t *chunk_count, void *const *events,
                                           int32_t n_threads);

/* n_steps fused calls of StockTradingEnv.step (:220-357) for every env.
 *   actions      element (k, n, j) at actions[k*act_step_stride + n*act_env_stride + j];
 *                float32 (actions_f64 == 0) or float64; the `* hmax` product and the truncating
 *                int cast are done in that dtype, as numpy does (:304-307)
 *   rewards      [n_steps][N] f64, nullable (p->reward always receives the last one)
 *   flags        [n_steps][N] u8 FRL_FLAG_*, nullable
 *   obs          per obs_mode; float32
 *   auto_reset   apply DummyVecEnv.step_wait's reset-on-done after a terminal step
 *   stats        [FRL_N_STATS] f64 accumulators, nullable
 */
FRL_API int32_t frl_trading_rollout(const frl_trading_params *p, const void *actions, int32_t actions_f64,
                            int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps,
                            double *rewards, uint8_t *flags, float *obs, int32_t obs_mode,
                            int32_t auto_reset, double *stats, void *stream);

/* One StockTradingEnv.step: frl_trading_rollout with n_steps = 1, contiguous [N][D] actions,
 * FRL_OBS_LAST when obs != NULL. */
FRL_API int32_t frl_trading_step(const frl_trading_params *p, const void *actions, int32_t actions_f64,
                         double *rewards, uint8_t *flags, float *obs, int32_t auto_reset,
                         double *stats, void *stream);

/* =========================================================================================
 * A2  numpy / ElegantRL StockTradingEnv — finrl/meta/env_stock_trading/env_stocktrading_np.py
 * ========================================================================================= */
/* numpy scalar kinds the reference's Python-level variables carry under NEP 50 */
#define FRL_KIND_PY 0  /* Python float (weak) */
#define FRL_KIND_F32 1 /* np.float32 */
#define FRL_KIND_F64 2 /* np.float64 */
/* flag byte of this env: bits 0-1 FRL_FLAG_DONE / FRL_FLAG_LIQUIDATE, bits 4-5 kind of the reward */
#define FRL_NP_REWARD_KIND_SHIFT 4

typedef struct frl_np_params {
    int32_t n_envs;     /* N */
    int32_t stock_dim;  /* D, 1..128: D <= 32 keeps stocks / cool-down in registers (nptrading.cu), 33..128 streams
                           them (np_wide.cu; NASDAQ-100) */
    int32_t tech_dim;   /* columns of tech_array (= D*K, stock-major) */
    int32_t n_days;     /* T; max_step = T-1 */
    int32_t obs_dim;    /* O = state_dim = 1 + 2 + 3D + tech_dim (:63) */
    int32_t env_stride; /* leading dimension of stocks / cool (>= N) */
    double gamma, max_stock, min_stock_rate;
    double buy_cost_pct, sell_cost_pct, reward_scaling;
    double initial_capital;
    double obs_amount_floor; /* get_state shows max(amount, floor): 1e4 for the sibling StockEnvNAS100
                                (env_nas100_wrds.py:157), -inf for env_stocktrading_np */
    /* ---- tables ---- */
    const float *price;       /* [T][price_pitch] price_ary = f32(price_array), rows zero-padded (:27) */
    const float *turb_bool;   /* [T] f32(turbulence_array > thresh) (:32) */
    const float *obs_tmpl;    /* [T][O] get_state() row of day t with amount/stocks/cool-down zeroed:
                                 [0, turbulence_ary[t], turbulence_bool[t], price*2^-6, 0.., 0.., tech_ary[t]] */
    const float *init_stocks; /* [D] initial_stocks */
    /* ---- per-env state ---- */
    double *amount;         /* [N] self.amount (value) */
    uint8_t *kinds;         /* [N] bits 0-1 kind of amount, 2-3 of total_asset, 4-5 of gamma_reward */
    float *stocks;          /* [D][env_stride] self.stocks */
    float *cool;            /* [D][env_stride] self.stocks_cool_down */
    int32_t *day;           /* [N] */
    double *total;          /* [N] self.total_asset */
    double *gamma_reward;   /* [N] */
    double *init_total;     /* [N] self.initial_total_asset */
    double *episode_return; /* [N] self.episode_return (written when done) */
    int32_t price_pitch;    /* row pitch of `price` in floats: 32 for D <= 32, 128 for D <= 128 */
    int32_t train_reset;    /* != 0: the in-kernel auto-reset takes the if_train branch (:85-92) with draws from the
                               counter-based generator below instead of the deterministic branch */
    uint64_t reset_seed;    /* stream id of this LAUNCH for those draws; the caller changes it on every launch.  Draw
                               i of env n at rollout step k is splitmix64(reset_seed, n, k, i): stocks =
                               initial_stocks + randint(0, 64) per stock, amount = initial_capital *
                               uniform(0.95, 1.05) - (stocks * price).sum() — the reference's distributions (it draws
                               from numpy's global RandomState, whose order over a batch is undefined anyway) */
} frl_np_params;

/* StockTradingEnv.reset (:80-101) for envs with mask[n] != 0 (NULL = all).  stocks0 ([D][env_stride]
 * f32) and factor ([N] f64), both non-NULL, select the if_train branch with the caller's random
 * draws (initial_stocks + randint(0,64) and uniform(0.95,1.05)); otherwise the deterministic branch.
 * obs (nullable) receives [N][O]. */
FRL_API int32_t frl_np_reset(const frl_np_params *p, const uint8_t *mask, const float *stocks0,
                             const double *factor, float *obs, void *stream);
/* get_state (:149-162) of every env. */
FRL_API int32_t frl_np_observe(const frl_np_params *p, float *obs, void *stream);
/* n_steps fused StockTradingEnv.step (:103-147).  Same conventions as frl_trading_rollout; rewards
 * are the f64 values (their numpy kind is in the flag byte).  auto_reset applies the reset right after a
 * done step (the returned obs is then the reset obs): the deterministic branch, or with train_reset the
 * if_train branch with in-kernel random draws (reset_seed). */
FRL_API int32_t frl_np_rollout(const frl_np_params *p, const void *actions, int32_t actions_f64,
                               int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps,
                               double *rewards, uint8_t *flags, float *obs, int32_t obs_mode,
                               int32_t auto_reset, double *stats, void *stream);
FRL_API int32_t frl_np_step(const frl_np_params *p, const void *actions, int32_t actions_f64, double *rewards,
                            uint8_t *flags, float *obs, int32_t auto_reset, double *stats, void *stream);

/* =========================================================================================
 * A3  StockPortfolioEnv — finrl/meta/env_portfolio_allocation/env_portfolio.py
 * ========================================================================================= */
typedef struct frl_portfolio_params {
    int32_t n_envs;    /* N */
    int32_t stock_dim; /* D, 1..128 (D <= 32: exp values in registers; above: two sweeps over the staged row) */
    int32_t n_tech;    /* K */
    int32_t n_days;    /* T */
    int32_t obs_dim;   /* (D + K) * D : np.append(cov (D x D), tech rows (K x D), axis=0) flattened */
    int32_t _pad0;
    double initial_amount;
    /* ---- tables ---- */
    const double *ret;      /* [T][ret_pitch] ret[t][j] = close[t][j] / close[t-1][j] - 1 (row 0 unused), the
                               env-independent factor of the weighted return (:183-185) */
    const float *obs_table; /* [T][obs_dim] float32 image of the day's state matrix */
    /* ---- per-env state ---- */
    double *pv;     /* [N] self.portfolio_value */
    int32_t *day;   /* [N] self.day */
    double *reward; /* [N] self.reward (returned again by the terminal step) */
    /* ---- optional per-env outputs of the last step (NULL = skip) ---- */
    double *ret_out;     /* [N] portfolio_return (portfolio_return_memory entry) */
    double *weights_out; /* [N][D] softmax weights (actions_memory entry), float64 */
    int32_t ret_pitch; /* row pitch of `ret` in doubles: 32 for D <= 32, 128 for D <= 128 */
    int32_t reserved_;
} frl_portfolio_params;

/* StockPortfolioEnv.reset (:202-220) for envs with mask[n] != 0 (NULL = all). obs nullable [N][obs_dim]. */
FRL_API int32_t frl_portfolio_reset(const frl_portfolio_params *p, const uint8_t *mask, float *obs, void *stream);
/* materialise the state matrix of every env: obs[N][obs_dim] float32 (it depends on the day only, so
 * callers may instead index obs_table with p->day and skip this traffic entirely). */
FRL_API int32_t frl_portfolio_observe(const frl_portfolio_params *p, float *obs, void *stream);
/* n_steps fused StockPortfolioEnv.step (:125-200): softmax weights (no max-subtraction), weighted
 * one-day return, portfolio value update; reward = new portfolio value.  Conventions as
 * frl_trading_rollout.  Floating point: np.exp is reproduced to 1 ulp, not bit for bit. */
FRL_API int32_t frl_portfolio_rollout(const frl_portfolio_params *p, const void *actions, int32_t actions_f64,
                                      int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps,
                                      double *rewards, uint8_t *flags, float *obs, int32_t obs_mode,
                                      int32_t auto_reset, double *stats, void *stream);
FRL_API int32_t frl_portfolio_step(const frl_portfolio_params *p, const void *actions, int32_t actions_f64,
                                   double *rewards, uint8_t *flags, float *obs, int32_t auto_reset, double *stats,
                                   void *stream);

/* =========================================================================================
 * A4  StockTradingEnvCashpenalty — finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py
 * ========================================================================================= */
typedef struct frl_cashpenalty_params {
    int32_t n_envs;    /* N */
    int32_t stock_dim; /* D = len(assets), 1..128 */
    int32_t n_cols;    /* C = len(daily_information_cols) */
    int32_t n_days;    /* T = len(dates) */
    int32_t obs_dim;   /* O = state_space = 1 + D + D*C (:88-90) */
    int32_t discrete_actions;
    int32_t shares_increment;
    int32_t use_turbulence; /* turbulence_threshold is not None */
    int32_t patient;
    int32_t env_stride; /* leading dimension of `hold` (>= N) */
    double buy_cost_pct, sell_cost_pct;
    double hmax; /* currency per trade (scalar) */
    double turbulence_threshold;
    double initial_amount;
    double cash_penalty_proportion;
    /* ---- tables ---- */
    const double *close;   /* [T][D] closings per date */
    const double *turb;    /* [T] "turbulence" column (read only when use_turbulence) */
    const float *obs_tmpl; /* [T][O] float32: [0, 0 x D, get_date_vector(t) asset-major (:160-173)] */
    /* ---- per-env state ---- */
    double *cash;        /* [N] cash_on_hand */
    double *hold;        /* [D][env_stride] holdings, stock-major (fractional unless discrete_actions) */
    double *hold_alt;    /* [D][env_stride] second holdings buffer: a step writes the new holdings into the
                            buffer that is NOT current and flips the env's bit, so a CASH SHORTAGE termination
                            leaves the state untouched without a second pass */
    int32_t *date_index; /* [N] */
    int32_t *start;      /* [N] starting_point */
    uint8_t *fresh;      /* [N] bit 0: self.turbulence is still the 0 set by reset; bit 1: hold_alt is current */
    double *last_cash;   /* [N] account_information["cash"][-1] */
    double *last_total;  /* [N] account_information["total_assets"][-1] */
    double *sum_trades;  /* [N] */
    /* ---- optional ---- */
    const double *hmax_vec; /* [D] or NULL: per-asset hmax array (`actions * self.hmax` broadcasts, :268); numpy
                               array-array promotion then applies: float32 actions * float64 hmax -> float64 */
    int32_t hmax_vec_f32;   /* the caller's array was float32: the product with float32 actions stays float32 */
    int32_t random_start;   /* != 0: the in-kernel auto-reset draws starting_point = randint(0, int(T * 0.5)) like
                               reset() with random_start=True (:135-137) instead of 0 */
    uint64_t reset_seed;    /* stream id of this LAUNCH for those draws (see frl_np_params.reset_seed) */
    const double *close_rc; /* [T][D][2] or NULL: (close, RN(1 / close)) pairs — the reciprocal by IEEE division on the
                               host (1/0 = inf, sign kept).  With it `actions / closings` (:286) costs three
                               instructions instead of a division routine and stays correctly rounded: q = RN(v * r),
                               e = fma(-c, q, v) exact, RN(q + e * r) = RN(v / c) (Markstein's theorem; the correction
                               step CUDA's own division ends with).  Required by the bulk-staged kernel variant. */
} frl_cashpenalty_params;

/* reset (:132-158) for envs with mask[n] != 0 (NULL = all); start_points [N] (NULL = 0, i.e.
 * random_start=False). obs nullable [N][O]. */
FRL_API int32_t frl_cashpenalty_reset(const frl_cashpenalty_params *p, const uint8_t *mask, const int32_t *start_points,
                                      float *obs, void *stream);
FRL_API int32_t frl_cashpenalty_observe(const frl_cashpenalty_params *p, float *obs, void *stream);
/* n_steps fused step (:300-372) incl. get_transactions (:258-298) and get_reward (:246-256).
 * Conventions as frl_trading_rollout; flags add FRL_FLAG_SHORTAGE.  np.dot's summation order is
 * BLAS-specific, so fp64 results agree with the reference to 1e-9 relative, not bit for bit. */
FRL_API int32_t frl_cashpenalty_rollout(const frl_cashpenalty_params *p, const void *actions, int32_t actions_f64,
                                        int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps,
                                        double *rewards, uint8_t *flags, float *obs, int32_t obs_mode,
                                        int32_t auto_reset, double *stats, void *stream);
FRL_API int32_t frl_cashpenalty_step(const frl_cashpenalty_params *p, const void *actions, int32_t actions_f64,
                                     double *rewards, uint8_t *flags, float *obs, int32_t auto_reset, double *stats,
                                     void *stream);

/* =========================================================================================
 * Sibling  StockTradingEnvStopLoss — finrl/meta/env_stock_trading/env_stocktrading_stoploss.py
 * (the cash-penalty env plus average-buy-price tracking, stop-loss liquidation and stop-loss /
 * low-profit penalties in the reward)
 * ========================================================================================= */
typedef struct frl_stoploss_params {
    int32_t n_envs;    /* N */
    int32_t stock_dim; /* D, 1..128 */
    int32_t n_cols;    /* C */
    int32_t n_days;    /* T */
    int32_t obs_dim;   /* O = 1 + D + D*C */
    int32_t discrete_actions;
    int32_t shares_increment;
    int32_t use_turbulence;
    int32_t patient;
    int32_t env_stride; /* leading dimension of the stock-major arrays (>= N) */
    double buy_cost_pct, sell_cost_pct;
    double hmax;
    double turbulence_threshold;
    double initial_amount;
    double cash_penalty_proportion;
    double stoploss_penalty;
    double min_profit_penalty; /* 1 + profit_loss_ratio * (1 - stoploss_penalty) (:101) */
    /* ---- tables ---- */
    const double *close;   /* [T][D] */
    const double *turb;    /* [T] */
    const float *obs_tmpl; /* [T][O] */
    /* ---- per-env state ---- */
    double *cash;        /* [N] */
    double *assets;      /* [2][6][D][env_stride]: two buffers of the six stock-major per-asset arrays, in the
                            order holdings (state_memory[-1]), previous holdings (state_memory[-2]),
                            avg_buy_price, n_buys, closing_diff_avg_buy, profit_sell_diff_avg_buy.  A step
                            reads the env's current buffer and writes the other one (single streaming pass) */
    int32_t *date_index; /* [N] */
    int32_t *start;      /* [N] */
    uint8_t *fresh;      /* [N] bit 0: self.turbulence is still the 0 set by reset; bit 1: current buffer */
    double *last_cash;   /* [N] */
    double *last_total;  /* [N] */
    double *sum_trades;  /* [N] */
    /* ---- optional ---- */
    const double *hmax_vec; /* [D] or NULL: per-asset hmax array, as in frl_cashpenalty_params */
    int32_t hmax_vec_f32;
    int32_t random_start;   /* as in frl_cashpenalty_params */
    uint64_t reset_seed;
} frl_stoploss_params;

FRL_API int32_t frl_stoploss_reset(const frl_stoploss_params *p, const uint8_t *mask, const int32_t *start_points,
                                   float *obs, void *stream);
FRL_API int32_t frl_stoploss_observe(const frl_stoploss_params *p, float *obs, void *stream);
/* n_steps fused step (:292-443) incl. get_reward (:255-290); conventions as frl_cashpenalty_rollout. */
FRL_API int32_t frl_stoploss_rollout(const frl_stoploss_params *p, const void *actions, int32_t actions_f64,
                                     int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps, double *rewards,
                                     uint8_t *flags, float *obs, int32_t obs_mode, int32_t auto_reset, double *stats,
                                     void *stream);
FRL_API int32_t frl_stoploss_step(const frl_stoploss_params *p, const void *actions, int32_t actions_f64, double *rewards,
                                  uint8_t *flags, float *obs, int32_t auto_reset, double *stats, void *stream);

/* =========================================================================================
 * Sibling  CryptoEnv — finrl/meta/env_cryptocurrency_trading/env_multiple_crypto.py
 * ========================================================================================= */
typedef struct frl_crypto_params {
    int32_t n_envs;     /* N */
    int32_t stock_dim;  /* D = crypto_num, 1..32 */
    int32_t tech_dim;   /* columns of tech_array */
    int32_t n_days;     /* T; max_step = T - lookback - 1 */
    int32_t lookback;
    int32_t obs_dim;    /* O = 1 + D + tech_dim * lookback (what get_state really returns, :93-99) */
    int32_t env_stride; /* leading dimension of stocks (>= N) */
    int32_t _pad0;
    double initial_capital, buy_cost_pct, sell_cost_pct, gamma;
    /* ---- tables ---- */
    const double *price;    /* [T][32] price_array (float64), rows zero-padded */
    const double *act_norm; /* [32] action_norm_vector (:103-111) */
    const float *obs_tmpl;  /* [T][O] get_state row of time t with the cash / stocks slots zeroed */
    /* ---- per-env state ---- */
    double *cash;           /* [N] */
    float *stocks;          /* [D][env_stride] fractional positions (float32 like the reference) */
    int32_t *time;          /* [N] */
    double *total;          /* [N] total_asset */
    double *gamma_return;   /* [N] (not cleared by reset, like the reference) */
    double *episode_return; /* [N] written when done */
} frl_crypto_params;

FRL_API int32_t frl_crypto_reset(const frl_crypto_params *p, const uint8_t *mask, float *obs, void *stream);
FRL_API int32_t frl_crypto_observe(const frl_crypto_params *p, float *obs, void *stream);
/* n_steps fused CryptoEnv.step (:59-91); conventions as frl_trading_rollout (auto_reset = reset after done). */
FRL_API int32_t frl_crypto_rollout(const frl_crypto_params *p, const void *actions, int32_t actions_f64,
                                   int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps, double *rewards,
                                   uint8_t *flags, float *obs, int32_t obs_mode, int32_t auto_reset, double *stats,
                                   void *stream);
FRL_API int32_t frl_crypto_step(const frl_crypto_params *p, const void *actions, int32_t actions_f64, double *rewards,
                                uint8_t *flags, float *obs, int32_t auto_reset, double *stats, void *stream);

/* =========================================================================================
 * Table precompute: the pandas loops that feed the envs
 * ========================================================================================= */
/* Rolling sample covariance (ddof = 1) and mean of daily returns.
 *   ret      [T][D] f64 daily returns (pct_change of close; row 0 is NaN and never read)
 *   window w (0 <= w < n_out) covers rows [first_row + w, first_row + w + n_rows)
 *   cov_out  [n_out][D][D] f64, mean_out [n_out][D] f64 (nullable)
 * With first_row = 1, n_rows = lookback this is the tutorial's `cov_list`
 * (tutorials/2-Advance/FinRL_PortfolioAllocation_Explainable_DRL.py:157-174: 253 closes -> 252 returns
 * -> .cov()); with first_row = 1 .. and n_rows = 252 shifted one day back it is the history block of
 * FeatureEngineer.calculate_turbulence (finrl/meta/preprocessor/preprocessors.py:215-267). */
FRL_API int32_t frl_rolling_cov(const double *ret, int32_t n_days, int32_t stock_dim, int32_t first_row,
                                int32_t n_rows, int32_t n_out, double *cov_out, double *mean_out, void *stream);

/* FeatureEngineer.calculate_turbulence (finrl/meta/preprocessor/preprocessors.py:215-267) for a complete
 * (NaN-free) return table ret[T][D] (row 0 = the NaN row of pct_change, never read): for every day
 * i >= start, temp = x^T pinv(cov_i) x with x = ret[i] - mean_i, where cov[i - start] / mean[i - start] are the
 * window statistics from frl_rolling_cov; np.linalg.pinv's cut-off (singular values <= rcond * largest, 1e-15
 * by default) is applied to the eigenvalues of the symmetric matrix (batched Jacobi eigen-solve, one thread
 * block per day).  out[T]: 0 for the first `start` days, then temp where positive except the first two positive
 * values (:250-262).  scratch: [T - start] doubles. */
FRL_API int32_t frl_turbulence(const double *ret, int32_t n_days, int32_t stock_dim, int32_t start, const double *cov,
                               const double *mean, double rcond, double *scratch, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FINRL_B200_H */
