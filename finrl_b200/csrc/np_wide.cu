// A2 wide — numpy / ElegantRL StockTradingEnv for 33..128 stocks (NASDAQ-100: StockEnvNAS100 at its natural
// size; reference: finrl/meta/env_stock_trading/env_stocktrading_np.py, env_nas100_wrds.py).
//
// nptrading.cu keeps stocks / cool-down counters of <= 32 stocks in registers.  Here they stay in the
// stock-major global arrays and are STREAMED, one thread per env, in the order the reference walks them —
// sells and buys both run in ascending stock index (:112,:120), so the step is two passes of blocks of 8
// stocks (loads of a block issued together), the second of which also accumulates numpy's pairwise
// float32 sum of stocks*price (8 accumulators over the full blocks, then the tail) and leaves the float32
// stocks image in the lane's own action-staging row for the observation writer.  Same NEP-50 scalar
// arithmetic (np_common.cuh) as the register kernel: bit-identical results (the whole numpy-env parity suite
// runs under both kernels).
#include <stdlib.h>

#include <type_traits>

#include "np_common.cuh"

namespace frl {
namespace {

#ifndef FRL_NPW_MIN_BLOCKS
#define FRL_NPW_MIN_BLOCKS 4
#endif
#ifndef FRL_NPW_A64
#define FRL_NPW_A64 1  // all-float64 instantiation of the step body (A/B switch)
#endif

__device__ __forceinline__ void npw_cp_async(float *dst, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void npw_cp_async(double *dst, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// numpy pairwise float32 sum for n <= 128, fed one block of 8 products at a time in index order
struct Pairwise8 {
    float r[8];
    float res;
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = 0.0f;
        res = 0.0f;
    }
    // x: the products of stocks j0..j0+7 (entries >= D ignored); nb = D >> 3 full blocks
    __device__ __forceinline__ void block(const float (&x)[8], int j0, int D)
    {
        const int nb = D >> 3;
        const int b = j0 >> 3;
        if (b < nb) {
            if (b == 0) {
#pragma unroll
                for (int u = 0; u < 8; ++u) r[u] = x[u];
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) r[u] = fadd(r[u], x[u]);
            }
            if (b == nb - 1) res = fadd(fadd(fadd(r[0], r[1]), fadd(r[2], r[3])), fadd(fadd(r[4], r[5]), fadd(r[6], r[7])));
        } else {  // the tail (or everything when D < 8): sequential
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (j0 + u < D) res = fadd(res, x[u]);
        }
    }
};

template <typename ActT>
__device__ __forceinline__ void npw_write_obs_range(const frl_np_params &p, const ActT *stage, int P, const float *amountf,
                                                    const int *day_s, float *__restrict__ obs, long long env0, int nvalid, int lane,
                                                    int beg, int end, int img_beg, float img_mul)
{
    // positions [beg, end) of every row; [img_beg, img_beg + D) come from the staged image (times img_mul)
    const int O = p.obs_dim, D = p.stock_dim;
    constexpr int step = sizeof(ActT) / sizeof(float);
    for (int r = 0; r < nvalid; ++r) {
        const float *irow = reinterpret_cast<const float *>(stage + (size_t)r * P);
        const float *trow = p.obs_tmpl + (size_t)day_s[r] * O;
        float *orow = obs + (size_t)(env0 + r) * O;
        for (int pos = beg + lane; pos < end; pos += 32) {
            float v;
            if (pos == 0)
                v = amountf[r];
            else if (pos >= img_beg && pos < img_beg + D)
                v = fmul(irow[(pos - img_beg) * step], img_mul);
            else
                v = __ldg(trow + pos);
            orow[pos] = v;
        }
    }
}

// Same range, all rows on one day and at most 16 chunks wide: the template values are loaded once into
// registers and every row costs one (patched) store per chunk.  The per-chunk facts (written at all? from the image?)
// are plain booleans: with a compile-time stock count (the NASDAQ-100 instantiation) every argument but `end` of the
// second phase is a constant after inlining and the facts of all chunks that lie wholly inside or outside the image
// fold away (`end_lo` is a compile-time lower bound of `end`).
template <typename ActT, int MAXC>
__device__ __forceinline__ void npw_write_obs_range_uniform(const frl_np_params &p, const ActT *stage, int P, const float *amountf,
                                                            int day0, float *__restrict__ obs, long long env0, int nvalid, int lane,
                                                            int beg, int end, int end_lo, int img_beg, float img_mul, int D)
{
    const int O = p.obs_dim;
    constexpr int step = sizeof(ActT) / sizeof(float);
    const float *trow = p.obs_tmpl + (size_t)day0 * O;
    float t[MAXC];
    bool ok[MAXC], img[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int pos = beg + lane + 32 * c;
        ok[c] = pos < end_lo || pos < end;
        img[c] = pos >= img_beg && pos < img_beg + D;
        t[c] = (ok[c] && !img[c]) ? __ldg(trow + pos) : 0.0f;
    }
    const bool first = (beg + lane) == 0;
    float *orow = obs + (size_t)env0 * O + beg + lane;
    const float *irow = reinterpret_cast<const float *>(stage) + (beg + lane - img_beg) * step;
#pragma unroll 2
    for (int r = 0; r < nvalid; ++r) {
        const float am = beg == 0 ? amountf[r] : 0.0f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (ok[c]) {
                float v = t[c];
                if (img[c]) v = fmul(irow[32 * step * c], img_mul);
                if (c == 0 && first) v = am;
                obs_store<kStoreCS>(orow + 32 * c, v);
            }
        }
        orow += O;
        irow += (size_t)P * step;
    }
}

template <typename ActT>
__device__ __forceinline__ void npw_write_obs(const frl_np_params &p, const ActT *stage, int P, const float *amountf, const int *day_s,
                                              float *__restrict__ obs, long long env0, int nvalid, int lane, int beg, int end,
                                              int end_lo, int img_beg, float img_mul, int D)
{
    const int day0 = day_s[0];
    bool uniform = true;
    if (lane < nvalid) uniform = day_s[lane] == day0;
    uniform = __all_sync(0xffffffffu, uniform);
    const int nch = (end - beg + 31) >> 5;  // chunk count compiled in (rounded up to 4 / 8 / 12 / 16)
#define FRL_NPW_RANGE(MAXC) \
    npw_write_obs_range_uniform<ActT, MAXC>(p, stage, P, amountf, day0, obs, env0, nvalid, lane, beg, end, end_lo, img_beg, img_mul, D)
    if (uniform && nch <= 4)
        FRL_NPW_RANGE(4);
    else if (uniform && nch <= 8)
        FRL_NPW_RANGE(8);
    else if (uniform && nch <= 12)
        FRL_NPW_RANGE(12);
    else if (uniform && nch <= 16)
        FRL_NPW_RANGE(16);
    else
        npw_write_obs_range<ActT>(p, stage, P, amountf, day_s, obs, env0, nvalid, lane, beg, end, img_beg, img_mul);
#undef FRL_NPW_RANGE
}

// Row pitch of the staging tile.  Generic variant: odd (conflict-free per-lane row walks with scalar accesses).
// Bulk-staged variant (float32 actions, D a multiple of four): rows are read and rewritten as 128-bit words of four
// stocks, which is conflict-free within each quarter-warp when the pitch is an ODD multiple of four floats.
__host__ __device__ inline int npw_pitch(int D, bool bulk) { return bulk ? (((D >> 2) & 1) ? D : D + 4) : (D | 1); }
__host__ __device__ inline size_t npw_warp_bytes(int D, bool bulk, size_t act_size)
{
    return (((size_t)32 * npw_pitch(D, bulk) * act_size + 32 * sizeof(float) + 32 * sizeof(int) + 16) + 15) & ~(size_t)15;
}

// BULK: the tile's 32 action rows are contiguous, 16-byte-aligned runs — every lane hands ITS row to the bulk-copy
// engine (TMA; one instruction instead of D cp.async with their row / column arithmetic) and the passes read four
// actions per 128-bit shared-memory word.
// DCT > 0: the stock count compiled in (NASDAQ-100, D = 100): loop bounds, pitch and the observation writer's chunk
// facts become constants.
template <typename ActT, int WARPS, bool BULK, int DCT>
__global__ void __launch_bounds__(WARPS * 32, FRL_NPW_MIN_BLOCKS * 128 / (WARPS * 32))
np_wide_kernel(const frl_np_params p, const ActT *__restrict__ actions, long long act_step_stride, long long act_env_stride,
               int n_steps, double *__restrict__ rewards, uint8_t *__restrict__ flags_out, float *__restrict__ obs, int obs_mode,
               int auto_reset, double *__restrict__ stats)
{
    static_assert(!BULK || sizeof(ActT) == 4, "the bulk-staged variant reads float4 batches");
    stats_exchange_previous(stats);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.n_envs, D = DCT > 0 ? DCT : p.stock_dim, T = p.n_days;
    const size_t ld = (size_t)p.env_stride;
    const int P = npw_pitch(D, BULK);
    unsigned char *base = smem_raw + warp * npw_warp_bytes(D, BULK, sizeof(ActT));
    ActT *stage = reinterpret_cast<ActT *>(base);  // [32 envs][P]
    float *amountf = reinterpret_cast<float *>(base + (size_t)32 * P * sizeof(ActT));
    int *day_s = reinterpret_cast<int *>(amountf + 32);
    const unsigned mbar = smem_u32(day_s + 32);  // staging mbarrier (BULK); 8-byte aligned
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;

    const int kinds = p.kinds[n];
    NV amount = nv(p.amount[n], kinds & 3);
    NV total = nv(p.total[n], (kinds >> 2) & 3);
    NV gr = nv(p.gamma_reward[n], (kinds >> 4) & 3);
    int day = p.day[n];
    double init_total = 0.0;
    bool init_total_loaded = false;
    float *sp = p.stocks + n, *cp = p.cool + n;  // stock j at [j * ld]
    ActT *myrow = stage + (size_t)lane * P;
    const NV one_minus_sc = nv(dsub(1.0, p.sell_cost_pct), FRL_KIND_PY);
    const NV one_plus_bc = nv(dadd(1.0, p.buy_cost_pct), FRL_KIND_PY);
    const int min_action = (int)dmul(p.max_stock, p.min_stock_rate);  // int(max_stock * min_stock_rate) (:111)
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0, st_liq = 0.0;
    unsigned stage_phase = 0;
    if (BULK && lane == 0) mbar_init(mbar, 1);

    for (int k = 0; k < n_steps; ++k) {
        // ---- stage this step's actions into [env][P] rows (cp.async: the whole tile in flight at once) ----
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        if constexpr (BULK) {
            fence_proxy_async_smem();  // the rows were read / written through the generic proxy during the last step
            __syncwarp();
            if (lane == 0) mbar_expect_tx(mbar, (unsigned)nvalid * (unsigned)D * 4u);
            __syncwarp();
            if (valid) {
                bulk_copy_g2s(smem_u32(stage + (size_t)lane * P), abase + (size_t)(env0 + lane) * D, (unsigned)D * 4u, mbar);
            } else {
                for (int j = 0; j < D; ++j) stage[(size_t)lane * P + j] = ActT(0);
            }
            mbar_wait(mbar, stage_phase);
            stage_phase ^= 1u;
        } else if (act_env_stride == D) {
            const ActT *tile = abase + (size_t)env0 * D;
            const int cnt = nvalid * D;
            int row = 0, col = lane;
            while (col >= D) { col -= D; ++row; }
            for (int e = lane; e < 32 * D; e += 32) {
                ActT *dst = stage + row * P + col;
                if (e < cnt)
                    npw_cp_async(dst, tile + e);
                else
                    *dst = ActT(0);
                col += 32;
                while (col >= D) { col -= D; ++row; }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else {
            for (int r = 0; r < 32; ++r)
                for (int j = lane; j < D; j += 32)
                    stage[r * P + j] = r < nvalid ? abase[(size_t)(env0 + r) * act_env_stride + j] : ActT(0);
        }
        __syncwarp();

        int flags = 0;
        NV reward = nv(0.0, FRL_KIND_PY);
        bool image_ok = false;  // the staging row holds the float32 stocks of the current state
        // steady state of the NEP-50 kinds (see nptrading.cu): all float64 for the whole tile -> the step body runs with
        // the kinds as compile-time constants (same operations, same bits)
        const bool all64 = FRL_NPW_A64 && __all_sync(0xffffffffu, amount.k == FRL_KIND_F64 && total.k == FRL_KIND_F64 &&
                                                                      gr.k == FRL_KIND_F64);
        auto step_body = [&](auto a64_tag) {
            constexpr bool A64 = decltype(a64_tag)::value;
            if (day >= T - 1) {
                flags = FRL_FLAG_DONE;  // past the last day: inert, done again
            } else {
                day += 1;  // trades happen at the NEW day's prices (:106-107)
                const float *prow = p.price + (size_t)day * p.price_pitch;
                Pairwise8 acc;
                acc.init();
                if (__ldg(p.turb_bool + day) == 0.0f) {
                    // ---- pass 1: cool_down += 1, sells in ascending index (:108-119) ----
                    // Both passes walk the stock-major arrays with pointer cursors in blocks of 8 stocks out of two
                    // STATIC register sets (A, B): a set is reloaded with the block 16 stocks further on as soon as its
                    // block has been traded, so two blocks of loads are always in flight and no register is moved.
                    const float *spq = sp, *cpq = cp;  // read cursors
                    float *spw = sp, *cpw = cp;        // write cursors
                    auto load_sc = [&](float (&st8)[8], float (&cl8)[8], int j0) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const bool in = j0 + u < D;
                            st8[u] = in ? __ldcg(spq) : 0.0f;
                            cl8[u] = in ? __ldcg(cpq) : 0.0f;
                            spq += ld;
                            cpq += ld;
                        }
                    };
                    auto load_actions = [&](float (&av)[8], int j0) {  // a float32 action is used as it is, a float64 one
                                                                       // through np_action_to_shares<double> below
                        if constexpr (BULK) {
                            const float4 a0 = *reinterpret_cast<const float4 *>(myrow + j0);
                            const float4 a1 = *reinterpret_cast<const float4 *>(myrow + j0 + 4);  // (beyond D: never used)
                            av[0] = a0.x, av[1] = a0.y, av[2] = a0.z, av[3] = a0.w;
                            av[4] = a1.x, av[5] = a1.y, av[6] = a1.z, av[7] = a1.w;
                        }
                    };
                    auto load_prices = [&](float (&pv)[8], int j0) {  // rows are 128 floats: always in bounds, 16-byte aligned
                        const float4 p0 = __ldg(reinterpret_cast<const float4 *>(prow + j0));
                        const float4 p1 = __ldg(reinterpret_cast<const float4 *>(prow + j0 + 4));
                        pv[0] = p0.x, pv[1] = p0.y, pv[2] = p0.z, pv[3] = p0.w;
                        pv[4] = p1.x, pv[5] = p1.y, pv[6] = p1.z, pv[7] = p1.w;
                    };
                    float stA[8], clA[8], stB[8], clB[8];
                    load_sc(stA, clA, 0);
                    load_sc(stB, clB, 8);
                    auto sell8 = [&](float (&st8)[8], float (&cl8)[8], int j0) {
                        float av[8], pv[8];
                        load_actions(av, j0);
                        load_prices(pv, j0);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = j0 + u;
                            if (j < D) {
                                int aj;
                                if constexpr (BULK)
                                    aj = np_action_to_shares<float>(av[u], p.max_stock);
                                else
                                    aj = np_action_to_shares<ActT>(myrow[j], p.max_stock);
                                float c = fadd(cl8[u], 1.0f);
                                if (aj < -min_action && pv[u] > 0.0f) {
                                    float s = st8[u];
                                    NV x;
                                    if ((double)(-aj) < (double)s) {  // min(stocks, -action) -> the int64
                                        const double nsh = (double)(-aj);
                                        s = (float)dsub((double)s, nsh);
                                        x = nv_mul(nv(dmul((double)pv[u], nsh), FRL_KIND_F64), one_minus_sc);
                                    } else {  // -> the float32 holding
                                        x = nv_mul(nv((double)fmul(pv[u], s), FRL_KIND_F32), one_minus_sc);
                                        s = fsub(s, s);
                                    }
                                    amount = nv_add_t<A64>(amount, x);
                                    c = 0.0f;
                                    if (valid) *spw = s;
                                }
                                if (valid) *cpw = c;
                            }
                            spw += ld;
                            cpw += ld;
                        }
                        load_sc(st8, cl8, j0 + 16);
                    };
#pragma unroll 1
                    for (int j0 = 0; j0 < D; j0 += 16) {
                        sell8(stA, clA, j0);
                        if (j0 + 8 < D) sell8(stB, clB, j0 + 8);
                    }
                    // ---- pass 2: buys in ascending index (:120-129, quirk Q6), asset sum, stocks image ----
                    // (re-reads the stocks pass 1 left behind — this thread's own stores, L2-resident)
                    spq = sp;
                    spw = sp;
                    cpw = cp;
                    auto load_s = [&](float (&st8)[8], int j0) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            st8[u] = j0 + u < D ? __ldcg(spq) : 0.0f;
                            spq += ld;
                        }
                    };
                    load_s(stA, 0);
                    load_s(stB, 8);
                    auto buy8 = [&](float (&st8)[8], int j0) {
                        float av[8], pv[8], x8[8], img[8];
                        load_actions(av, j0);
                        load_prices(pv, j0);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = j0 + u;
                            x8[u] = 0.0f;
                            img[u] = 0.0f;
                            if (j < D) {
                                int aj;
                                if constexpr (BULK)
                                    aj = np_action_to_shares<float>(av[u], p.max_stock);
                                else
                                    aj = np_action_to_shares<ActT>(myrow[j], p.max_stock);
                                const float pj = pv[u];
                                float s = st8[u];
                                if (aj > min_action && pj > 0.0f) {
                                    NV x;
                                    const double am = (A64 || amount.k == FRL_KIND_F64) ? amount.v : (double)(float)amount.v;
                                    const bool plenty = am >= dmul((double)(aj + 1), (double)pj);
                                    double avail = 0.0;
                                    if (!plenty && !(am >= 0.0 && am < (double)pj))
                                        avail = np_floor_div(am, (double)pj, A64 || amount.k == FRL_KIND_F64);
                                    if (plenty) {  // min(avail, action) -> the int64
                                        const double nsh = (double)aj;
                                        s = (float)dadd((double)s, nsh);
                                        x = nv_mul(nv(dmul((double)pj, nsh), FRL_KIND_F64), one_plus_bc);
                                    } else if (A64 || amount.k == FRL_KIND_F64) {
                                        s = (float)dadd((double)s, avail);
                                        x = nv_mul(nv(dmul((double)pj, avail), FRL_KIND_F64), one_plus_bc);
                                    } else {
                                        const float nsh = (float)avail;
                                        s = fadd(s, nsh);
                                        x = nv_mul(nv((double)fmul(pj, nsh), FRL_KIND_F32), one_plus_bc);
                                    }
                                    amount = nv_sub_t<A64>(amount, x);
                                    if (valid) {
                                        *spw = s;
                                        *cpw = 0.0f;
                                    }
                                }
                                x8[u] = fmul(s, pj);
                                img[u] = s;
                                if constexpr (!BULK) *reinterpret_cast<float *>(myrow + j) = s;  // action j is consumed: the slot takes the image
                            }
                            spw += ld;
                            cpw += ld;
                        }
                        if constexpr (BULK) {  // the four-action words are consumed: they take the float32 stocks image
                            *reinterpret_cast<float4 *>(myrow + j0) = make_float4(img[0], img[1], img[2], img[3]);
                            if (j0 + 4 < D) *reinterpret_cast<float4 *>(myrow + j0 + 4) = make_float4(img[4], img[5], img[6], img[7]);
                        }
                        acc.block(x8, j0, D);
                        load_s(st8, j0 + 16);
                    };
#pragma unroll 1
                    for (int j0 = 0; j0 < D; j0 += 16) {
                        buy8(stA, j0);
                        if (j0 + 8 < D) buy8(stB, j0 + 8);
                    }
                } else {
                    // ---- sell everything when turbulence (:131-134) ----
                    flags |= FRL_FLAG_LIQUIDATE;
                    for (int j0 = 0; j0 < D; j0 += 8) {
                        float x8[8];
    #pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = j0 + u;
                            x8[u] = j < D ? fmul(sp[(size_t)j * ld], __ldg(prow + j)) : 0.0f;
                        }
                        acc.block(x8, j0, D);
                    }
                    amount = nv_add_t<A64>(amount, nv_mul(nv((double)acc.res, FRL_KIND_F32), one_minus_sc));
                    for (int j = 0; j < D; ++j) {
                        if (valid) {
                            sp[(size_t)j * ld] = 0.0f;
                            cp[(size_t)j * ld] = 0.0f;
                        }
                        *reinterpret_cast<float *>(myrow + j) = 0.0f;
                    }
                    acc.init();  // (stocks * price).sum() of the emptied book: +0.0f
                    if (valid) st_liq += 1.0;
                }
                image_ok = true;
                // ---- reward bookkeeping (:136-145) ----
                const NV tot = nv_add_t<A64>(amount, nv((double)acc.res, FRL_KIND_F32));
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
            // reset (:80-101): deterministic branch, or the if_train branch with counter-based draws (train_reset)
            Pairwise8 acc;
            acc.init();
            for (int j0 = 0; j0 < D; j0 += 8) {
                float x8[8];
                uint64_t bits = p.train_reset ? reset_bits(p.reset_seed, n, k, 1 + j0 / 8) : 0;  // eight 6-bit draws
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + u;
                    x8[u] = 0.0f;
                    if (j < D) {
                        float s0 = p.init_stocks ? __ldg(p.init_stocks + j) : 0.0f;
                        if (p.train_reset) {
                            s0 = fadd(s0, (float)(int)(bits & 63));
                            bits >>= 6;
                        }
                        if (valid) {
                            sp[(size_t)j * ld] = s0;
                            cp[(size_t)j * ld] = 0.0f;
                        }
                        *reinterpret_cast<float *>(myrow + j) = s0;
                        x8[u] = fmul(s0, __ldg(p.price + j));
                    }
                }
                acc.block(x8, j0, D);
            }
            day = 0;
            if (p.train_reset) {  // initial_capital * uniform(0.95, 1.05) [py] - (stocks * price).sum() [f32] -> f32
                const double factor = 0.95 + (1.05 - 0.95) * reset_uniform01(reset_bits(p.reset_seed, n, k, 0));
                amount = nv_sub(nv(dmul(p.initial_capital, factor), FRL_KIND_PY), nv((double)acc.res, FRL_KIND_F32));
            } else {
                amount = nv(p.initial_capital, FRL_KIND_PY);
            }
            total = nv_add(amount, nv((double)acc.res, FRL_KIND_F32));
            init_total = total.v;
            init_total_loaded = true;
            gr = nv(0.0, FRL_KIND_PY);
            if (valid) p.init_total[n] = init_total;
            image_ok = true;
        }
        if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
            if (!image_ok) {
                for (int j = 0; j < D; ++j) npw_cp_async(reinterpret_cast<float *>(myrow + j), sp + (size_t)j * ld);
                asm volatile("cp.async.commit_group;\ncp.async.wait_all;" ::: "memory");
            }
            amountf[lane] = np_amount_obs(amount, p.obs_amount_floor);
            day_s[lane] = day;
            __syncwarp();
            float *o = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * p.obs_dim : (size_t)0);
            const int s_beg = 3 + D, c_beg = 3 + 2 * D;
            // [amount, turb, turb_bool, price, stocks * 2**-6 | cool-down, tech]: two phases through the one image
            npw_write_obs<ActT>(p, stage, P, amountf, day_s, o, env0, nvalid, lane, 0, c_beg, c_beg, s_beg, 0.015625f, D);
            __syncwarp();
            // (cp.async: all D cool-down lines of the tile in flight at once instead of D dependent load+store pairs)
            for (int j = 0; j < D; ++j) npw_cp_async(reinterpret_cast<float *>(myrow + j), cp + (size_t)j * ld);
            asm volatile("cp.async.commit_group;\ncp.async.wait_all;" ::: "memory");
            __syncwarp();
            npw_write_obs<ActT>(p, stage, P, amountf, day_s, o, env0, nvalid, lane, c_beg, p.obs_dim, 3 + 3 * D, c_beg, 1.0f, D);
        }
    }
    if (valid) {
        p.amount[n] = amount.v;
        p.total[n] = total.v;
        p.gamma_reward[n] = gr.v;
        p.kinds[n] = (uint8_t)(amount.k | (total.k << 2) | (gr.k << 4));
        p.day[n] = day;
    }
    if (stats) {
        double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, valid ? total.v : 0.0, st_liq,
                                 valid ? (double)n_steps : 0.0, 0.0};
        reduce_stats8(v, lane, stats);
    }
}

// one warp per env, stocks / cool-down straight from the stock-major arrays (observe / reset only)
__global__ void np_observe_wide_kernel(const frl_np_params p, float *__restrict__ obs)
{
    const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= p.n_envs) return;
    const int O = p.obs_dim, D = p.stock_dim;
    const int s_beg = 3 + D, c_beg = 3 + 2 * D, sp_end = 3 + 3 * D;
    const float *trow = p.obs_tmpl + (size_t)p.day[n] * O;
    float *orow = obs + (size_t)n * O;
    for (int pos = lane; pos < O; pos += 32) {
        float v;
        if (pos == 0)
            v = np_amount_obs(nv(p.amount[n], p.kinds[n] & 3), p.obs_amount_floor);
        else if (pos >= s_beg && pos < c_beg)
            v = fmul(p.stocks[n + (size_t)(pos - s_beg) * p.env_stride], 0.015625f);
        else if (pos >= c_beg && pos < sp_end)
            v = p.cool[n + (size_t)(pos - c_beg) * p.env_stride];
        else
            v = __ldg(trow + pos);
        orow[pos] = v;
    }
}

template <typename ActT, int WARPS, bool BULK, int DCT = 0>
void npw_launch(const frl_np_params &p, const void *actions, long long sstride, long long estride, int n_steps, double *rewards,
                uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    const size_t smem = WARPS * npw_warp_bytes(p.stock_dim, BULK, sizeof(ActT));
    auto kern = np_wide_kernel<ActT, WARPS, BULK, DCT>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("np_rollout(wide): cannot reserve %zu B of shared memory", smem);
            return;
        }
    }
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    kern<<<(unsigned)((tiles + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p, (const ActT *)actions, sstride, estride, n_steps,
                                                                           rewards, flags, obs, obs_mode, auto_reset, stats);
}

}  // namespace

// frl_set_option("np_wide_bulk", 0) / FRL_NPW_BULK=0 keeps every shape on the generic staging path (tests run both)
int g_npw_bulk = -1;
int npw_bulk_enabled()
{
    if (g_npw_bulk < 0) {
        const char *m = getenv("FRL_NPW_BULK");
        g_npw_bulk = m ? (atoi(m) != 0) : 1;
    }
    return g_npw_bulk;
}

void launch_np_wide(const frl_np_params &p, const void *actions, int actions_f64, long long sstride, long long estride, int n_steps,
                    double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    // bulk-staged variant: float32 actions in the default layout, every row a 16-byte-aligned multiple of 16 bytes
    const bool bulk = npw_bulk_enabled() && !actions_f64 && (p.stock_dim & 3) == 0 && estride == p.stock_dim &&
                      (sstride & 3) == 0 && (reinterpret_cast<uintptr_t>(actions) & 15) == 0;
    if (actions_f64)
        npw_launch<double, 2, false>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st);
    else if (bulk && p.stock_dim == 100)  // NASDAQ-100: stock count compiled in
        npw_launch<float, 4, true, 100>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st);
    else if (bulk)
        npw_launch<float, 4, true>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st);
    else
        npw_launch<float, 4, false>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st);
}

void launch_np_observe_wide(const frl_np_params &p, float *obs, cudaStream_t st)
{
    const long long threads = (long long)p.n_envs * 32;
    np_observe_wide_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p, obs);
}

}  // namespace frl
