// A3 — StockPortfolioEnv (reference: finrl/meta/env_portfolio_allocation/env_portfolio.py).
//
// One thread per env: softmax weights exp(a)/sum(exp(a)) in the action dtype with numpy's pairwise
// sum and NO max-subtraction (quirk Q8), the sequential weighted return over the D stocks, and
// portfolio_value *= 1 + return.  The (D+K) x D observation depends only on the day, so it is either
// not written at all (callers index obs_table with the day vector) or broadcast from a register-
// cached table row by the whole warp.  The env-independent ratio close[t]/close[t-1]-1 is a per-day
// table row built on the host with the reference's own numpy expression.
#include "common.cuh"

namespace frl {
namespace {

template <int SLOTS, typename ActT>
struct alignas(16) PfWarpSmem {
    ActT act[32 * SLOTS];
    int day[32];
    unsigned long long mbar;  // completion barrier of the observation image load
};

#ifndef FRL_PF_IMG_ROWS
#define FRL_PF_IMG_ROWS 1  // observation rows per shared-memory image (= rows per bulk store)
#endif

template <typename T>
__device__ __forceinline__ T pf_exp(T x);
template <>
__device__ __forceinline__ float pf_exp<float>(float x) { return expf(x); }
template <>
__device__ __forceinline__ double pf_exp<double>(double x) { return exp(x); }
template <typename T>
__device__ __forceinline__ T pf_add(T a, T b);
template <>
__device__ __forceinline__ float pf_add<float>(float a, float b) { return fadd(a, b); }
template <>
__device__ __forceinline__ double pf_add<double>(double a, double b) { return dadd(a, b); }
template <typename T>
__device__ __forceinline__ T pf_div(T a, T b);
template <>
__device__ __forceinline__ float pf_div<float>(float a, float b) { return __fdiv_rn(a, b); }
template <>
__device__ __forceinline__ double pf_div<double>(double a, double b) { return __ddiv_rn(a, b); }

// np.sum over a contiguous vector: numpy's pairwise summation (n < 8 sequential, else 8 accumulators)
template <int SLOTS, typename T>
__device__ __forceinline__ T pf_pairwise_sum(const T (&x)[SLOTS], int D)
{
    if (D < 8) {
        T res = T(0);
#pragma unroll
        for (int j = 0; j < 8 && j < SLOTS; ++j)
            if (j < D) res = pf_add(res, x[j]);
        return res;
    }
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = x[j];
    const int nb = D >> 3;
#pragma unroll
    for (int b = 1; b < SLOTS / 8; ++b) {
        if (b < nb) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = pf_add(r[j], x[8 * b + j]);
        }
    }
    T res = pf_add(pf_add(pf_add(r[0], r[1]), pf_add(r[2], r[3])), pf_add(pf_add(r[4], r[5]), pf_add(r[6], r[7])));
#pragma unroll
    for (int j = 8; j < SLOTS; ++j)
        if (j >= 8 * nb && j < D) res = pf_add(res, x[j]);
    return res;
}

// The observation of an env depends only on its day, so a tile whose 32 envs sit on one day writes 32 copies
// of one table row: the copy engine (TMA) loads the row into shared memory (`rows` copies of it, back to
// back) and stores it 32/rows times — no LSU store instructions, nothing to patch, all stores in flight at
// once.  otile = first row of the tile; rows and tile are 16-byte aligned (checked by the caller).
__device__ __forceinline__ void pf_write_obs_tile_bulk(const frl_portfolio_params &p, float *img, unsigned mbar, unsigned &phase,
                                                       float *__restrict__ otile, int d0, int rows, int lane)
{
    const unsigned row_bytes = 4u * (unsigned)p.obs_dim;
    __syncwarp();
    if (lane == 0) {
        const float *src = p.obs_table + (size_t)d0 * p.obs_dim;
        mbar_expect_tx(mbar, row_bytes * rows);
        for (int r = 0; r < rows; ++r) bulk_copy_g2s(smem_u32(img) + r * row_bytes, src, row_bytes, mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    if (lane == 0) {
        const unsigned chunk = row_bytes * rows;
        char *dst = reinterpret_cast<char *>(otile);
        for (int r = 0; r < 32; r += rows) bulk_store_s2g(dst + (size_t)r * row_bytes, smem_u32(img), chunk);
        bulk_wait_read<0>();  // the image is reloaded by the next step / the block may retire
    }
    __syncwarp();
}

// broadcast the day's observation row to the tile's envs (obs[N][O] float32)
__device__ __forceinline__ void pf_write_obs_tile(const frl_portfolio_params &p, const int *day_s,
                                                  float *__restrict__ obs, long long env0, int nvalid, int lane)
{
    const int O = p.obs_dim;
    const int d0 = day_s[0];
    bool uniform = true;
    if (lane < nvalid) uniform = (day_s[lane] == d0);
    uniform = __all_sync(0xffffffffu, uniform);
    if (uniform && (O & 3) == 0 && ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(p.obs_table)) & 15) == 0) {
        // rows are 16-B aligned: 128-bit stores, 8 x 512 B of the row per pass cached in registers
        const int O4 = O >> 2;
        const float4 *trow = reinterpret_cast<const float4 *>(p.obs_table + (size_t)d0 * O);
        for (int base = 0; base < O4; base += 256) {
            float4 t[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int pos = base + 32 * c + lane;
                t[c] = pos < O4 ? __ldg(trow + pos) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 *orow = reinterpret_cast<float4 *>(obs + (size_t)env0 * O) + base + lane;
            const int full = min(8, (O4 - base) >> 5);           // chunks entirely inside the row
            const bool tail = (base + 32 * full + lane) < O4;     // the partial chunk after them
            for (int r = 0; r < nvalid; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c < full)
                        __stcs(orow + 32 * c, t[c]);
                    else if (c == full && tail)
                        __stcs(orow + 32 * c, t[c]);
                }
                orow += O4;
            }
        }
    } else if (uniform) {
        const float *trow = p.obs_table + (size_t)d0 * O;
        for (int base = 0; base < O; base += 256) {  // 8 x 128 B of the row per pass, cached in registers
            float t[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int pos = base + 32 * c + lane;
                t[c] = pos < O ? __ldg(trow + pos) : 0.0f;
            }
            float *orow = obs + (size_t)env0 * O + base + lane;
            for (int r = 0; r < nvalid; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (base + 32 * c + lane < O) orow[32 * c] = t[c];
                orow += O;
            }
        }
    } else {
        for (int r = 0; r < nvalid; ++r) {
            const float *trow = p.obs_table + (size_t)day_s[r] * O;
            float *orow = obs + (size_t)(env0 + r) * O;
            for (int pos = lane; pos < O; pos += 32) orow[pos] = __ldg(trow + pos);
        }
    }
}

template <int SLOTS, typename ActT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
portfolio_rollout_kernel(const frl_portfolio_params p, const ActT *__restrict__ actions, long long act_step_stride,
                         long long act_env_stride, int n_steps, double *__restrict__ rewards,
                         uint8_t *__restrict__ flags_out, float *__restrict__ obs, int obs_mode, int auto_reset,
                         double *__restrict__ stats, int img_rows)
{
    stats_exchange_previous(stats);
    using SM = PfWarpSmem<SLOTS, ActT>;
    __shared__ SM smem[WARPS];
    extern __shared__ __align__(16) float pf_img[];  // [WARPS][img_rows * O] when the bulk writer is enabled
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SM &sm = smem[warp];
    const int N = p.n_envs, D = p.stock_dim, T = p.n_days;
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;
    const bool bulk_ok = img_rows > 0 && nvalid == 32;
    float *img = pf_img + (size_t)warp * img_rows * p.obs_dim;
    unsigned img_phase = 0;
    if (bulk_ok && lane == 0) mbar_init(smem_u32(&sm.mbar), 1);

    double pv = p.pv[n], last_reward = p.reward[n];
    int day = p.day[n];
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        stage_actions_flat<SLOTS, ActT>(sm.act, abase, env0, act_env_stride, D, nvalid, lane);
        __syncwarp();

        uint8_t flags = 0;
        double reward;
        if (day >= T - 1) {
            // terminal branch (:127-156): no state change, the previous reward again
            flags = FRL_FLAG_DONE;
            reward = last_reward;
            if (valid) {
                st_done += 1.0;
                st_epi += pv;
            }
            if (auto_reset) {  // DummyVecEnv.step_wait -> reset (:202-220)
                pv = p.initial_amount;
                day = 0;
            }
        } else {
            // softmax_normalization (:225-229)
            ActT e[SLOTS];
            const ActT *arow = sm.act + lane * D;
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) e[j] = (j < D) ? pf_exp<ActT>(arow[j]) : ActT(0);
            const ActT den = pf_pairwise_sum<SLOTS, ActT>(e, D);
            day += 1;
            const double *rrow = p.ret + (size_t)day * p.ret_pitch;
            // sum(((close_new / close_old) - 1) * weights): Python sum, sequential (:183-185)
            double pr = 0.0;
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
                if (j < D) {
                    const double w = (double)pf_div<ActT>(e[j], den);
                    pr = dadd(pr, dmul(__ldg(rrow + j), w));
                    if (p.weights_out && valid) p.weights_out[(size_t)n * D + j] = w;
                }
            }
            if (p.ret_out && valid) p.ret_out[n] = pr;
            pv = dmul(pv, dadd(1.0, pr));
            reward = pv;  // reward = new portfolio value, unscaled (:196)
            last_reward = reward;
        }
        if (valid) {
            if (rewards) rewards[(size_t)k * N + n] = reward;
            if (flags_out) flags_out[(size_t)k * N + n] = flags;
            st_r += reward;
            st_r2 += reward * reward;
        }
        if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
            sm.day[lane] = day;
            float *o = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * p.obs_dim : (size_t)0);
            float *otile = o + (size_t)env0 * p.obs_dim;
            const int d0 = __shfl_sync(0xffffffffu, day, 0);
            if (bulk_ok && (reinterpret_cast<uintptr_t>(otile) & 15) == 0 && __all_sync(0xffffffffu, day == d0)) {
                pf_write_obs_tile_bulk(p, img, smem_u32(&sm.mbar), img_phase, otile, d0, img_rows, lane);
            } else {
                __syncwarp();
                pf_write_obs_tile(p, sm.day, o, env0, nvalid, lane);
            }
        }
    }
    if (valid) {
        p.pv[n] = pv;
        p.day[n] = day;
        p.reward[n] = last_reward;
    }
    if (stats) {
        double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, valid ? pv : 0.0, 0.0, valid ? (double)n_steps : 0.0, 0.0};
        reduce_stats8(v, lane, stats);
    }
}

// ---- D = 33..128 (e.g. a NASDAQ-100 portfolio): the exp values do not fit in registers, so the step makes two
// sweeps over the lane's own staged action row — exp in place + numpy's pairwise sum fed in blocks of 8, then
// weights and the sequential weighted return.  Same arithmetic and order as the register kernel.
template <typename T>
struct PfPairwise {
    T r[8];
    T res;
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = T(0);
        res = T(0);
    }
    __device__ __forceinline__ void block(const T (&x)[8], int j0, int D)
    {
        const int nb = D >> 3, b = j0 >> 3;
        if (b < nb) {
            if (b == 0) {
#pragma unroll
                for (int u = 0; u < 8; ++u) r[u] = x[u];
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) r[u] = pf_add(r[u], x[u]);
            }
            if (b == nb - 1)
                res = pf_add(pf_add(pf_add(r[0], r[1]), pf_add(r[2], r[3])), pf_add(pf_add(r[4], r[5]), pf_add(r[6], r[7])));
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (j0 + u < D) res = pf_add(res, x[u]);
        }
    }
};

__device__ __forceinline__ void pfw_cp_async(float *dst, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void pfw_cp_async(double *dst, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <typename ActT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
portfolio_wide_kernel(const frl_portfolio_params p, const ActT *__restrict__ actions, long long act_step_stride,
                      long long act_env_stride, int n_steps, double *__restrict__ rewards, uint8_t *__restrict__ flags_out,
                      float *__restrict__ obs, int obs_mode, int auto_reset, double *__restrict__ stats)
{
    stats_exchange_previous(stats);
    extern __shared__ __align__(16) unsigned char pfw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.n_envs, D = p.stock_dim, T = p.n_days;
    const int P = D | 1;  // odd row pitch: conflict-free per-lane row walks
    const size_t warp_bytes = (((size_t)32 * P * sizeof(ActT) + 32 * sizeof(int)) + 15) & ~(size_t)15;
    ActT *stage = reinterpret_cast<ActT *>(pfw_smem + warp * warp_bytes);
    int *day_s = reinterpret_cast<int *>(stage + (size_t)32 * P);
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;
    ActT *myrow = stage + (size_t)lane * P;

    double pv = p.pv[n], last_reward = p.reward[n];
    int day = p.day[n];
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        if (act_env_stride == D) {
            const ActT *tile = abase + (size_t)env0 * D;
            const int cnt = nvalid * D;
            int row = 0, col = lane;
            while (col >= D) { col -= D; ++row; }
            for (int e = lane; e < 32 * D; e += 32) {  // cp.async: the whole tile in flight at once
                ActT *dst = stage + row * P + col;
                if (e < cnt)
                    pfw_cp_async(dst, tile + e);
                else
                    *dst = ActT(0);
                col += 32;
                while (col >= D) { col -= D; ++row; }
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_all;" ::: "memory");
        } else {
            for (int r = 0; r < 32; ++r)
                for (int j = lane; j < D; j += 32)
                    stage[r * P + j] = r < nvalid ? abase[(size_t)(env0 + r) * act_env_stride + j] : ActT(0);
        }
        __syncwarp();

        uint8_t flags = 0;
        double reward;
        if (day >= T - 1) {
            flags = FRL_FLAG_DONE;
            reward = last_reward;
            if (valid) {
                st_done += 1.0;
                st_epi += pv;
            }
            if (auto_reset) {
                pv = p.initial_amount;
                day = 0;
            }
        } else {
            PfPairwise<ActT> acc;
            acc.init();
            for (int j0 = 0; j0 < D; j0 += 8) {
                ActT e8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    e8[u] = ActT(0);
                    if (j0 + u < D) {
                        e8[u] = pf_exp<ActT>(myrow[j0 + u]);
                        myrow[j0 + u] = e8[u];
                    }
                }
                acc.block(e8, j0, D);
            }
            const ActT den = acc.res;
            day += 1;
            const double *rrow = p.ret + (size_t)day * p.ret_pitch;
            double pr = 0.0;
            for (int j = 0; j < D; ++j) {
                const double w = (double)pf_div<ActT>(myrow[j], den);
                pr = dadd(pr, dmul(__ldg(rrow + j), w));
                if (p.weights_out && valid) p.weights_out[(size_t)n * D + j] = w;
            }
            if (p.ret_out && valid) p.ret_out[n] = pr;
            pv = dmul(pv, dadd(1.0, pr));
            reward = pv;
            last_reward = reward;
        }
        if (valid) {
            if (rewards) rewards[(size_t)k * N + n] = reward;
            if (flags_out) flags_out[(size_t)k * N + n] = flags;
            st_r += reward;
            st_r2 += reward * reward;
        }
        if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
            day_s[lane] = day;
            __syncwarp();
            float *o = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * p.obs_dim : (size_t)0);
            pf_write_obs_tile(p, day_s, o, env0, nvalid, lane);
        }
    }
    if (valid) {
        p.pv[n] = pv;
        p.day[n] = day;
        p.reward[n] = last_reward;
    }
    if (stats) {
        double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, valid ? pv : 0.0, 0.0, valid ? (double)n_steps : 0.0, 0.0};
        reduce_stats8(v, lane, stats);
    }
}

template <typename ActT, int WARPS>
void pf_launch_wide(const frl_portfolio_params &p, const void *actions, long long sstride, long long estride, int n_steps,
                    double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    const int P = p.stock_dim | 1;
    const size_t smem = WARPS * ((((size_t)32 * P * sizeof(ActT) + 32 * sizeof(int)) + 15) & ~(size_t)15);
    auto kern = portfolio_wide_kernel<ActT, WARPS>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    kern<<<(unsigned)((tiles + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p, (const ActT *)actions, sstride, estride, n_steps,
                                                                           rewards, flags, obs, obs_mode, auto_reset, stats);
}

__global__ void portfolio_reset_kernel(const frl_portfolio_params p, const uint8_t *__restrict__ mask)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.n_envs) return;
    if (mask && !mask[n]) return;
    p.pv[n] = p.initial_amount;
    p.day[n] = 0;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) portfolio_observe_kernel(const frl_portfolio_params p, float *__restrict__ obs)
{
    __shared__ int day_s[WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.n_envs;
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    day_s[warp][lane] = p.day[lane < nvalid ? env0 + lane : (long long)N - 1];
    __syncwarp();
    pf_write_obs_tile(p, day_s[warp], obs, env0, nvalid, lane);
}

int32_t pf_validate(const frl_portfolio_params *p)
{
    FRL_REQUIRE(p != nullptr, "portfolio: params is NULL");
    FRL_REQUIRE(p->n_envs >= 1, "portfolio: n_envs must be >= 1 (got %d)", p->n_envs);
    FRL_REQUIRE(p->stock_dim >= 1 && p->stock_dim <= 128, "portfolio: stock_dim must be in 1..128 (got %d)", p->stock_dim);
    FRL_REQUIRE((p->ret_pitch == 32 || p->ret_pitch == 128) && p->ret_pitch >= p->stock_dim,
                "portfolio: ret_pitch must be 32 or 128 and >= stock_dim (got %d for D=%d)", p->ret_pitch, p->stock_dim);
    FRL_REQUIRE(p->n_tech >= 0 && p->n_days >= 1, "portfolio: bad n_tech/n_days (%d, %d)", p->n_tech, p->n_days);
    FRL_REQUIRE(p->obs_dim == (p->stock_dim + p->n_tech) * p->stock_dim, "portfolio: obs_dim %d != (D+K)*D = %d",
                p->obs_dim, (p->stock_dim + p->n_tech) * p->stock_dim);
    FRL_REQUIRE(p->ret && p->obs_table, "portfolio: table pointer is NULL");
    FRL_REQUIRE(p->pv && p->day && p->reward, "portfolio: state pointer is NULL");
    return FRL_OK;
}

template <int SLOTS, typename ActT, int WARPS>
void pf_launch(const frl_portfolio_params &p, const void *actions, long long sstride, long long estride, int n_steps,
               double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    const unsigned grid = (unsigned)((tiles + WARPS - 1) / WARPS);
    auto kern = portfolio_rollout_kernel<SLOTS, ActT, WARPS>;
    // bulk observation writer: rows must be 16-byte multiples on a 16-byte-aligned table, and the images of
    // the block's warps must leave room for several blocks per SM
    int img_rows = 0;
    if (obs_mode != FRL_OBS_NONE && (p.obs_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(p.obs_table) & 15) == 0 &&
        (size_t)FRL_PF_IMG_ROWS * p.obs_dim * 4 * WARPS <= 40 * 1024)
        img_rows = FRL_PF_IMG_ROWS;
    const size_t dyn = (size_t)img_rows * p.obs_dim * 4 * WARPS;
    // static + dynamic shared memory above 48 KB needs the opt-in (per device, so it is not cached here)
    if (dyn + WARPS * sizeof(PfWarpSmem<SLOTS, ActT>) > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    kern<<<grid, WARPS * 32, dyn, st>>>(p, (const ActT *)actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode,
                                        auto_reset, stats, img_rows);
}

}  // namespace
}  // namespace frl

using namespace frl;

extern "C" int32_t frl_portfolio_observe(const frl_portfolio_params *p, float *obs, void *stream)
{
    if (int32_t rc = pf_validate(p)) return rc;
    FRL_REQUIRE(obs != nullptr, "portfolio_observe: obs is NULL");
    constexpr int W = 4;
    const long long tiles = ((long long)p->n_envs + 31) / 32;
    portfolio_observe_kernel<W><<<(unsigned)((tiles + W - 1) / W), W * 32, 0, (cudaStream_t)stream>>>(*p, obs);
    return check_launch("portfolio_observe");
}

extern "C" int32_t frl_portfolio_reset(const frl_portfolio_params *p, const uint8_t *mask, float *obs, void *stream)
{
    if (int32_t rc = pf_validate(p)) return rc;
    portfolio_reset_kernel<<<(p->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, mask);
    if (int32_t rc = check_launch("portfolio_reset")) return rc;
    if (obs) return frl_portfolio_observe(p, obs, stream);
    return FRL_OK;
}

extern "C" int32_t frl_portfolio_rollout(const frl_portfolio_params *p, const void *actions, int32_t actions_f64,
                                         int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps,
                                         double *rewards, uint8_t *flags, float *obs, int32_t obs_mode,
                                         int32_t auto_reset, double *stats, void *stream)
{
    if (int32_t rc = pf_validate(p)) return rc;
    FRL_REQUIRE(actions != nullptr, "portfolio_rollout: actions is NULL");
    FRL_REQUIRE(n_steps >= 1, "portfolio_rollout: n_steps must be >= 1 (got %d)", n_steps);
    FRL_REQUIRE(act_env_stride >= p->stock_dim, "portfolio_rollout: act_env_stride %lld < stock_dim", (long long)act_env_stride);
    FRL_REQUIRE(obs_mode >= FRL_OBS_NONE && obs_mode <= FRL_OBS_ALL, "portfolio_rollout: bad obs_mode %d", obs_mode);
    FRL_REQUIRE(obs_mode == FRL_OBS_NONE || obs != nullptr, "portfolio_rollout: obs is NULL but obs_mode=%d", obs_mode);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = p->stock_dim;
#define FRL_GO(SLOTS)                                                                                             \
    do {                                                                                                          \
        if (actions_f64)                                                                                          \
            pf_launch<SLOTS, double, 2>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags,    \
                                        obs, obs_mode, auto_reset, stats, st);                                    \
        else                                                                                                      \
            pf_launch<SLOTS, float, 4>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags,     \
                                       obs, obs_mode, auto_reset, stats, st);                                     \
    } while (0)
    if (D > 32) {
        if (actions_f64)
            pf_launch_wide<double, 2>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags, obs, obs_mode,
                                      auto_reset, stats, st);
        else
            pf_launch_wide<float, 4>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, flags, obs, obs_mode,
                                     auto_reset, stats, st);
    } else if (D <= 8)
        FRL_GO(8);
    else if (D <= 16)
        FRL_GO(16);
    else
        FRL_GO(32);
#undef FRL_GO
    return check_launch("portfolio_rollout");
}

extern "C" int32_t frl_portfolio_step(const frl_portfolio_params *p, const void *actions, int32_t actions_f64,
                                      double *rewards, uint8_t *flags, float *obs, int32_t auto_reset, double *stats,
                                      void *stream)
{
    if (p == nullptr) {
        set_error("portfolio_step: params is NULL");
        return FRL_E_INVALID;
    }
    return frl_portfolio_rollout(p, actions, actions_f64, (int64_t)p->n_envs * p->stock_dim, p->stock_dim, 1, rewards,
                                 flags, obs, obs ? FRL_OBS_LAST : FRL_OBS_NONE, auto_reset, stats, stream);
}
