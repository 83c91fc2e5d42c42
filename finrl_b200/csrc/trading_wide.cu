// A1 wide — StockTradingEnv for 33..128 stocks (NASDAQ-100) at LARGE batch sizes.
//
// trading.cu keeps the 32 sort keys of a DOW-30 env in registers; trading_small.cu spends 8 lanes per env and
// is built for latency at small batches.  Here the thread-per-env mapping of trading.cu is kept (32 envs of a
// warp in lock-step, sequential fp64 chains per thread) and the per-env arrays that no longer fit in registers
// — the 64/128 packed sort keys and the holdings — live in shared memory as [slot][33]-pitch columns (lane =
// env: conflict-free for the per-lane walks of the network and the trade loops AND for the per-row reads of the
// observation writer).  np.argsort's network runs as loops over the key column.  Same arithmetic
// and order as the other two kernels => bit-identical (the D > 32 goldens and fuzz cases run under both).
#include <stdlib.h>

#include "common.cuh"
#include "sort_network.inc"
#include "trading_common.cuh"

#ifndef FRL_TW_REGNET
#define FRL_TW_REGNET 0  // 1: the D = 100 instantiation sorts on registers (straight-line code; measured slower, see profiles/)
#endif

namespace frl {
namespace {

constexpr int kPitchW = 33;
constexpr int IBW = 7;                              // index bits of a packed key (D <= 128)
constexpr int AMAXW = (1 << (31 - IBW)) - 1;        // same clamp as the 8-lanes-per-env kernel for D > 32

__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_all;" ::: "memory"); }

// keys are (a << 7) | index; swap only on STRICT a[lo] > a[hi] — ties keep network order
#define FRL_CEXW(lo, hi)                                                                           \
    {                                                                                              \
        const bool sw_ = (lo) > ((hi) | ((1 << IBW) - 1));                                         \
        const int t_ = sw_ ? (hi) : (lo);                                                          \
        (hi) = sw_ ? (lo) : (hi);                                                                  \
        (lo) = t_;                                                                                 \
    }

// ascending bitonic network on `slots` (64 or 128) keys of this lane's column: for block size 2, 4, ..., slots a
// flip stage (mirror pairs inside each block) followed by half-cleaners of distance blk/4 ... 1.  Every stage of
// distance <= 16 only touches aligned groups of 32 slots, so those run on 32 keys held in registers: the block sizes
// 2..32 are exactly the 32-slot network per group, and each later level ends with the half-cleaners 16..1 per group.
// Only the flip stages of block size 64 / 128 and the half-cleaner of distance 32 exchange through shared memory —
// 24 of the 28 stages of a 128-slot network stay in registers.
//
// Pad slots (index >= D, key INT_MAX) never move: every compare-exchange is ascending (lo < hi, swap on strict
// greater), a pad is never smaller than anything, and the pads start on top — so no real key ever enters a slot
// >= D and any exchange whose upper slot is a pad is a no-op.  The column therefore only holds D keys, the
// shared-memory stages run over exactly the pairs whose upper slot is live, and a last group with at most 16 live
// keys runs the 16-slot forms (its upper half is all pads).
// (MAKE turns what the column holds into the packed key of slot j: the identity, or — in the network's first pass,
// when the column still holds the staged float32 actions — the (actions * hmax).astype(int) cast and the packing)
struct KeyIdentity {
    __device__ __forceinline__ int operator()(int v, int) const { return v; }
};
template <bool FULL, typename MAKE = KeyIdentity>
__device__ __forceinline__ void load_keys32(int (&k)[32], const int *col, int g, int D, MAKE make = MAKE())
{
#pragma unroll
    for (int i = 0; i < 32; ++i) k[i] = (FULL || g + i < D) ? make(col[(g + i) * kPitchW], g + i) : 0x7fffffff;
}
template <bool FULL>
__device__ __forceinline__ void store_keys32(const int (&k)[32], int *col, int g, int D)
{
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (FULL || g + i < D) col[(g + i) * kPitchW] = k[i];
}
template <typename MAKE = KeyIdentity>
__device__ __forceinline__ void load_keys16(int (&k)[16], const int *col, int g, int D, MAKE make = MAKE())
{
#pragma unroll
    for (int i = 0; i < 16; ++i) k[i] = g + i < D ? make(col[(g + i) * kPitchW], g + i) : 0x7fffffff;
}
__device__ __forceinline__ void store_keys16(const int (&k)[16], int *col, int g, int D)
{
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (g + i < D) col[(g + i) * kPitchW] = k[i];
}

template <typename MAKE = KeyIdentity>
__device__ __forceinline__ void network_w(int *col, int slots, int D, MAKE make = MAKE())
{
    const int d32 = D & ~31;  // groups below are full: no pad checks
    // ---- block sizes 2..32 ----
#pragma unroll 1
    for (int g = 0; g < d32; g += 32) {
        int k[32];
        load_keys32<true>(k, col, g, D, make);
        FRL_SORT_NETWORK_32(FRL_CEXW, k)
        store_keys32<true>(k, col, g, D);
    }
    if (D - d32 > 16) {
        int k[32];
        load_keys32<false>(k, col, d32, D, make);
        FRL_SORT_NETWORK_32(FRL_CEXW, k)
        store_keys32<false>(k, col, d32, D);
    } else if (D > d32) {
        int k[16];
        load_keys16(k, col, d32, D, make);
        FRL_SORT_NETWORK_16(FRL_CEXW, k)
        store_keys16(k, col, d32, D);
    }
    // ---- block sizes 64 and 128 ----
    for (int blk = 64; blk <= slots; blk <<= 1) {
        const int half = blk >> 1;
        for (int b = 0; b < D; b += blk) {  // flip stage: (b + i, b + blk - 1 - i), upper slot live <=> i >= b + blk - D
            int *plo = col + (b + max(0, b + blk - D)) * kPitchW;
            int *phi = col + (b + blk - 1 - max(0, b + blk - D)) * kPitchW;
            for (int i = max(0, b + blk - D); i < half; ++i) {
                const int x = *plo, y = *phi;
                if (x > (y | ((1 << IBW) - 1))) {
                    *plo = y;
                    *phi = x;
                }
                plo += kPitchW;
                phi -= kPitchW;
            }
        }
        for (int d = blk >> 2; d >= 32; d >>= 1) {  // half-cleaners of distance >= 32: (b + i, b + i + d), i < D - b - d
            for (int b = 0; b + d < D; b += 2 * d) {
                int *plo = col + b * kPitchW;
                const int cnt = min(d, D - b - d);
                for (int i = 0; i < cnt; ++i) {
                    const int x = *plo, y = plo[d * kPitchW];
                    if (x > (y | ((1 << IBW) - 1))) {
                        *plo = y;
                        plo[d * kPitchW] = x;
                    }
                    plo += kPitchW;
                }
            }
        }
        // half-cleaners 16..1 on registers, per group of 32
#pragma unroll 1
        for (int g = 0; g < d32; g += 32) {
            int k[32];
            load_keys32<true>(k, col, g, D);
            FRL_MERGE_HALF_32(FRL_CEXW, k)
            store_keys32<true>(k, col, g, D);
        }
        if (D - d32 > 16) {
            int k[32];
            load_keys32<false>(k, col, d32, D);
            FRL_MERGE_HALF_32(FRL_CEXW, k)
            store_keys32<false>(k, col, d32, D);
        } else if (D > d32) {  // the distance-16 partners are all pads
            int k[16];
            load_keys16(k, col, d32, D);
#pragma unroll
            for (int d = 8; d >= 1; d >>= 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if ((i & d) == 0) FRL_CEXW(k[i], k[i + d])
            }
            store_keys16(k, col, d32, D);
        }
    }
}

// Stock count compiled in (NASDAQ-100): the whole 128-slot network runs on REGISTERS — the block of slots 0..63 in
// k0, slots 64..127 in k1 — as straight-line code: 32-slot networks on the four quarters, the block-size-64 level on
// each half, the block-size-128 flip stage between the halves and its half-cleaners on each half (the decomposition is
// checked against the full network in gen_sort_network.py's terms by tests/test_host_cpu.py).  The pad slots are
// compile-time INT_MAX constants here, so every compare-exchange that touches one folds away: what is left are the
// ~1400 exchanges between live keys, with no shared-memory traffic in between.
template <int DCT>
__device__ __forceinline__ void network_regs_w(int (&k0)[64], int (&k1)[64])
{
    static_assert(DCT > 64 && DCT <= 128, "two 64-slot register blocks");
    FRL_SORT_NETWORK_32(FRL_CEXW, k0)
    FRL_SORT_NETWORK_32(FRL_CEXW, (k0 + 32))
    FRL_MERGE_FLIP_64(FRL_CEXW, k0)
    FRL_SORT_NETWORK_32(FRL_CEXW, k1)
    FRL_SORT_NETWORK_32(FRL_CEXW, (k1 + 32))
    FRL_MERGE_FLIP_64(FRL_CEXW, k1)
#pragma unroll
    for (int i = 0; i < 64; ++i) FRL_CEXW(k0[i], k1[63 - i])
    FRL_MERGE_HALF_64(FRL_CEXW, k0)
    FRL_MERGE_HALF_64(FRL_CEXW, k1)
}

__device__ __forceinline__ double total_asset_w(double cash, const double *__restrict__ prow, const int *hcol, int D)
{
    double acc = 0.0;
    for (int j = 0; j < D; ++j) acc = dadd(acc, dmul(__ldg(prow + j), (double)hcol[j * kPitchW]));
    return dadd(cash, acc);
}

// observation rows [cash, close x D, holdings x D, tech]: per-day template cached in registers (up to 32
// chunks of 32 positions), holdings patched in from the [slot][33] column array
template <int NCH>
__device__ __forceinline__ void write_obs_rows_w(const frl_trading_params &p, const int *hold, const float *cashf, float *__restrict__ obs,
                                                 long long env0, int nvalid, int lane, int sd0, int D)
{
    // (D is a compile-time constant in the NASDAQ-100 instantiation: the per-chunk facts of every chunk that lies wholly
    // inside or outside the holdings fold away)
    const int O = p.obs_dim;
    float t[NCH];
    bool ok[NCH], img[NCH];
    const float *trow = p.obs_tmpl + (size_t)sd0 * O + lane;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int pos = lane + 32 * c;
        ok[c] = pos < 1 + 2 * D || pos < O;  // O >= 1 + 2D
        img[c] = pos > D && pos <= 2 * D;
        t[c] = (ok[c] && !img[c]) ? __ldg(trow + 32 * c) : 0.0f;
    }
    const int *hsrc = hold + (lane - 1 - D) * kPitchW;  // holdings of stock (pos - 1 - D), chunk c adds 32 rows
    float *orow = obs + (size_t)env0 * O + lane;
#pragma unroll 2
    for (int r = 0; r < nvalid; ++r) {
        const float cf = cashf[r];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (ok[c]) {
                float x = t[c];
                if (img[c]) x = (float)hsrc[32 * c * kPitchW + r];
                if (c == 0 && lane == 0) x = cf;
                obs_store<kStoreCS>(orow + 32 * c, x);
            }
        }
        orow += O;
    }
}

__device__ __forceinline__ void write_obs_tile_w(const frl_trading_params &p, const int *hold, const float *cashf, const int *sd_s,
                                                 float *__restrict__ obs, long long env0, int nvalid, int lane, int D)
{
    const int O = p.obs_dim;
    const int sd0 = sd_s[0];
    bool uniform = true;
    if (lane < nvalid) uniform = sd_s[lane] == sd0;
    uniform = __all_sync(0xffffffffu, uniform);
    const int nch = (O + 31) >> 5;
    if (uniform && nch <= 32) {
        if (nch <= 8)
            write_obs_rows_w<8>(p, hold, cashf, obs, env0, nvalid, lane, sd0, D);
        else if (nch <= 16)
            write_obs_rows_w<16>(p, hold, cashf, obs, env0, nvalid, lane, sd0, D);
        else if (nch <= 24)
            write_obs_rows_w<24>(p, hold, cashf, obs, env0, nvalid, lane, sd0, D);
        else
            write_obs_rows_w<32>(p, hold, cashf, obs, env0, nvalid, lane, sd0, D);
    } else {
        for (int r = 0; r < nvalid; ++r) {
            float *orow = obs + (size_t)(env0 + r) * O;
            const float *trow = p.obs_tmpl + (size_t)sd_s[r] * O;
            for (int pos = lane; pos < O; pos += 32) {
                float v = __ldg(trow + pos);
                if (pos == 0)
                    v = cashf[r];
                else if (pos > D && pos <= 2 * D)
                    v = (float)hold[(pos - 1 - D) * kPitchW + r];
                orow[pos] = v;
            }
        }
    }
}

template <typename ActT, int WARPS, int DCT>
__global__ void __launch_bounds__(WARPS * 32)
trading_wide_kernel(const frl_trading_params p, const ActT *__restrict__ actions, long long act_step_stride, long long act_env_stride,
                    int n_steps, double *__restrict__ rewards, uint8_t *__restrict__ flags_out, float *__restrict__ obs, int obs_mode,
                    int auto_reset, double *__restrict__ stats)
{
    stats_exchange_previous(stats);
    extern __shared__ __align__(16) unsigned char tw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.n_envs, D = DCT > 0 ? DCT : p.stock_dim, T = p.n_days, ld = p.env_stride;
    const int slots = D <= 64 ? 64 : 128;
    const size_t warp_ints = (size_t)(2 * D) * kPitchW + 64;
    int *key = reinterpret_cast<int *>(tw_smem) + warp * warp_ints;  // [D][33] (pad slots are never stored)
    int *hold = key + (size_t)D * kPitchW;                           // [D][33]
    float *cashf = reinterpret_cast<float *>(hold + (size_t)D * kPitchW);
    int *sd_s = reinterpret_cast<int *>(cashf + 32);
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;
    int *kcol = key + lane, *hcol = hold + lane;

    double cash = p.cash[n], cost = p.cost[n], last_reward = p.reward[n];
    int day = p.day[n], sday = p.sday[n], trades = p.trades[n];
    // holdings: global -> shared with cp.async (all D lines of the tile in flight at once; waited for together
    // with the first step's actions)
    for (int j = 0; j < D; ++j) cp_async4(hcol + j * kPitchW, p.hold + n + (size_t)j * ld);

    const double one_minus_sc = dsub(1.0, p.sell_cost_pct), one_plus_bc = dadd(1.0, p.buy_cost_pct);
    const int hmax_i = (int)max(-(double)AMAXW, min((double)AMAXW, p.hmax));
    const int mask_words = p.close_pitch >> 5;
    double asset = 0.0;
    bool asset_ok = false;
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0, st_liq = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        // ---- stage this step's actions, transposed, as raw bits into the key columns (coalesced reads) ----
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        ActT *acol = nullptr;  // f64 actions do not fit a 4-byte slot: they are read straight from global below
        if (sizeof(ActT) == 4) {
            if (act_env_stride == D) {
                // row by row, 32 stocks (128 contiguous bytes) per instruction: no row / column arithmetic
                const float *src = reinterpret_cast<const float *>(abase) + (size_t)env0 * D + lane;
                int *dst = key + lane * kPitchW;
                for (int r = 0; r < 32; ++r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (lane + 32 * c < D) {
                            if (r < nvalid)
                                cp_async4(dst + 32 * c * kPitchW, src + 32 * c);
                            else
                                dst[32 * c * kPitchW] = 0;
                        }
                    }
                    src += D;
                    dst += 1;
                }
            } else {
                for (int r = 0; r < 32; ++r)
                    for (int j = lane; j < D; j += 32)
                        key[j * kPitchW + r] =
                            r < nvalid ? __float_as_int((float)abase[(size_t)(env0 + r) * act_env_stride + j]) : 0;
            }
        }
        cp_async_wait_all();
        __syncwarp();
        (void)acol;

        uint8_t flags = 0;
        double reward;
        if (day >= T - 1) {
            // ---- terminal branch (:221-301): no state change, previous scaled reward again (Q3) ----
            flags = FRL_FLAG_DONE;
            reward = last_reward;
            if (valid) {
                const double *prow = p.close + (size_t)state_day(sday) * p.close_pitch;
                if (!asset_ok) asset = total_asset_w(cash, prow, hcol, D);
                st_done += 1.0;
                st_epi += asset;
            }
            if (auto_reset) {  // DummyVecEnv.step_wait -> reset (:359-393), stale-day quirk Q1
                cash = p.initial_amount;
                for (int j = 0; j < D; ++j) hcol[j * kPitchW] = p.init_hold ? __ldg(p.init_hold + j) : 0;
                sday = -day - 1;
                day = 0;
                cost = 0.0;
                trades = 0;
                if (valid) p.episode[n] += 1;
                asset_ok = false;
            }
        } else {
            const int sd = state_day(sday);
            const double turb = sday < 0 ? 0.0 : __ldg(p.risk + sd);
            const bool liq = p.use_turbulence && (turb >= p.turbulence_threshold);
            const double *prow = p.close + (size_t)sd * p.close_pitch;
            const double begin = asset_ok ? asset : total_asset_w(cash, prow, hcol, D);

            if (liq) {
                // actions = [-hmax]*D (:308-310): all keys tie, order = index order; price > 0 check (:138-163)
                flags = FRL_FLAG_LIQUIDATE;
                if (hmax_i > 0) {
                    for (int j = 0; j < D; ++j) {
                        const double pj = __ldg(prow + j);
                        const int h = hcol[j * kPitchW];
                        if (pj > 0.0 && h > 0) {
                            const double pv = dmul(pj, (double)h);
                            cash = dadd(cash, dmul(pv, one_minus_sc));
                            hcol[j * kPitchW] = 0;
                            cost = dadd(cost, dmul(pv, p.sell_cost_pct));
                            trades += 1;
                        }
                    }
                }
            } else {
                // ---- (actions * hmax).astype(int), packed sort keys, np.argsort order ----
                if constexpr (FRL_TW_REGNET && sizeof(ActT) == 4 && DCT > 64) {
                    int k0[64], k1[64];
#pragma unroll
                    for (int j = 0; j < 128; ++j) {
                        int key_j = 0x7fffffff;  // pad
                        if (j < DCT) {
                            const float a = __int_as_float(kcol[j * kPitchW]);
                            key_j = (max(-AMAXW, min(AMAXW, action_to_shares<float>(a, p.hmax))) << IBW) + j;
                        }
                        if (j < 64)
                            k0[j] = key_j;
                        else
                            k1[j - 64] = key_j;
                    }
                    network_regs_w<(DCT > 64 ? DCT : 128)>(k0, k1);
#pragma unroll
                    for (int j = 0; j < DCT; ++j) kcol[j * kPitchW] = j < 64 ? k0[j < 64 ? j : 0] : k1[j < 64 ? 0 : j - 64];
                } else if (sizeof(ActT) == 4) {
                    // the cast and the packing happen in the network's first register load (below)
                } else {
                    const ActT *arow = abase + (size_t)n * act_env_stride;
                    for (int j = 0; j < D; ++j) {
                        const int sh = max(-AMAXW, min(AMAXW, action_to_shares<ActT>(arow[j], p.hmax)));
                        kcol[j * kPitchW] = (sh << IBW) + j;
                    }
                }
                if constexpr (!(FRL_TW_REGNET && sizeof(ActT) == 4 && DCT > 64)) {
                    if constexpr (sizeof(ActT) == 4) {
                        const double hmax = p.hmax;
                        network_w(kcol, slots, D, [hmax](int bits, int j) {
                            return (max(-AMAXW, min(AMAXW, action_to_shares<float>(__int_as_float(bits), hmax))) << IBW) + j;
                        });
                    } else {
                        network_w(kcol, slots, D);
                    }
                }
                const uint32_t *dis_row = p.disable_mask ? p.disable_mask + (size_t)sd * mask_words : nullptr;
                if (dis_row) {  // days without any disabled stock (nearly all) skip the per-trade lookup
                    uint32_t any = 0;
                    for (int w = 0; w < mask_words; ++w) any |= __ldg(dis_row + w);
                    if (any == 0) dis_row = nullptr;
                }

                // Both loops are software-pipelined: the next order entry with its price, holding and disable bit is
                // fetched before the current trade's dependent fp64 chain.
                const int IM = (1 << IBW) - 1;
                // ---- sells, most negative first (:321-324, _sell_stock :102-135) ----
                {
                    int kk = kcol[0];
                    int j = kk & IM;
                    double pj = __ldg(prow + j);
                    int h = hcol[j * kPitchW];
                    bool dis = dis_row && ((__ldg(dis_row + (j >> 5)) >> (j & 31)) & 1u);
                    for (int s = 0; s < D; ++s) {
                        if (kk >= 0) break;
                        const int kn = kcol[min(s + 1, D - 1) * kPitchW];
                        const int jn = kn & IM;
                        const double pn = __ldg(prow + jn);
                        const int hn = hcol[jn * kPitchW];  // another stock: not touched below
                        const bool dn = dis_row && ((__ldg(dis_row + (jn >> 5)) >> (jn & 31)) & 1u);
                        const int a = kk >> IBW;
                        if (!dis && h > 0) {
                            const int m = min(-a, h);
                            const double pv = dmul(pj, (double)m);
                            cash = dadd(cash, dmul(pv, one_minus_sc));
                            hcol[j * kPitchW] = h - m;
                            cost = dadd(cost, dmul(pv, p.sell_cost_pct));
                            trades += 1;
                        }
                        kk = kn;
                        j = jn;
                        pj = pn;
                        h = hn;
                        dis = dn;
                    }
                }
                // ---- buys, largest first, each limited by the cash left (:328-330, _buy_stock :171-201) ----
                {
                    int kk = kcol[(D - 1) * kPitchW];
                    int j = kk & IM;
                    double pj = __ldg(prow + j);
                    double unit = dmul(pj, one_plus_bc);
                    bool dis = dis_row && ((__ldg(dis_row + (j >> 5)) >> (j & 31)) & 1u);
                    for (int s = D - 1; s >= 0; --s) {
                        if (kk < (1 << IBW)) break;
                        const int kn = kcol[max(s - 1, 0) * kPitchW];
                        const int jn = kn & IM;
                        const double pn = __ldg(prow + jn);
                        const double un = dmul(pn, one_plus_bc);
                        const bool dn = dis_row && ((__ldg(dis_row + (jn >> 5)) >> (jn & 31)) & 1u);
                        const int a = kk >> IBW;
                        if (!dis) {
                            double nsh = (double)a;
                            trades += 1;  // even when 0 shares end up bought (Q5)
                            bool buy = true;
                            if (!(cash >= dmul(nsh + 1.0, unit))) {
                                if (cash >= 0.0 && cash < unit) {
                                    buy = false;  // 0 shares: nothing changes
                                } else {
                                    const double avail = floor_div_f64(cash, unit);
                                    nsh = (nsh < avail) ? nsh : avail;
                                }
                            }
                            if (buy) {
                                const double pv = dmul(pj, nsh);
                                cash = dsub(cash, dmul(pv, one_plus_bc));
                                hcol[j * kPitchW] += (int)nsh;
                                cost = dadd(cost, dmul(pv, p.buy_cost_pct));
                            }
                        }
                        kk = kn;
                        j = jn;
                        pj = pn;
                        unit = un;
                        dis = dn;
                    }
                }
            }
            // ---- state: s -> s+1 (:335-352) ----
            day += 1;
            sday = day;
            asset = total_asset_w(cash, p.close + (size_t)day * p.close_pitch, hcol, D);
            asset_ok = true;
            reward = dmul(dsub(asset, begin), p.reward_scaling);
            last_reward = reward;
            if (liq && valid) st_liq += 1.0;
        }
        if (valid) {
            if (rewards) rewards[(size_t)k * N + n] = reward;
            if (flags_out) flags_out[(size_t)k * N + n] = flags;
            st_r += reward;
            st_r2 += reward * reward;
        }
        if (obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1)) {
            cashf[lane] = (float)cash;
            sd_s[lane] = state_day(sday);
            __syncwarp();
            float *o = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * p.obs_dim : (size_t)0);
            write_obs_tile_w(p, hold, cashf, sd_s, o, env0, nvalid, lane, D);
        }
    }

    if (valid) {
        p.cash[n] = cash;
        p.cost[n] = cost;
        p.reward[n] = last_reward;
        p.day[n] = day;
        p.sday[n] = sday;
        p.trades[n] = trades;
        for (int j = 0; j < D; ++j) p.hold[n + (size_t)j * ld] = hcol[j * kPitchW];
    }
    if (p.asset_out || stats) {
        if (!asset_ok) asset = total_asset_w(cash, p.close + (size_t)state_day(sday) * p.close_pitch, hcol, D);
        if (p.asset_out && valid) p.asset_out[n] = asset;
    }
    if (stats) {
        double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, valid ? asset : 0.0, st_liq, valid ? (double)n_steps : 0.0,
                                 valid ? (double)trades : 0.0};
        reduce_stats8(v, lane, stats);
    }
}

template <typename ActT, int WARPS, int DCT = 0>
int32_t tw_launch(const frl_trading_params &p, const void *actions, long long sstride, long long estride, int n_steps, double *rewards,
                  uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats, cudaStream_t st)
{
    const size_t smem = (size_t)WARPS * ((size_t)(2 * p.stock_dim) * kPitchW + 64) * sizeof(int);
    auto kern = trading_wide_kernel<ActT, WARPS, DCT>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("trading_rollout(wide): cannot reserve %zu B of shared memory (%s)", smem, cudaGetErrorString(e));
            return FRL_E_CUDA;
        }
    }
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    kern<<<(unsigned)((tiles + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p, (const ActT *)actions, sstride, estride, n_steps,
                                                                           rewards, flags, obs, obs_mode, auto_reset, stats);
    return FRL_OK;
}

}  // namespace

// frl_set_option("trading_wide_regs", 0) / FRL_TW_REGS=0 keeps D = 100 on the generic (runtime stock count) kernel
int g_tw_regs = -1;
int tw_regs_enabled()
{
    if (g_tw_regs < 0) {
        const char *m = getenv("FRL_TW_REGS");
        g_tw_regs = m ? (atoi(m) != 0) : 1;
    }
    return g_tw_regs;
}

int32_t launch_trading_wide(const frl_trading_params &p, const void *actions, int actions_f64, long long sstride, long long estride,
                            int n_steps, double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats,
                            cudaStream_t st)
{
    if (actions_f64)
        return tw_launch<double, 2>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st);
    if (p.stock_dim == 100 && tw_regs_enabled())  // NASDAQ-100: stock count compiled in, np.argsort's network in registers
        return tw_launch<float, 2, 100>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st);
    return tw_launch<float, 2>(p, actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats, st);
}

}  // namespace frl
