// A1 — StockTradingEnv (reference: finrl/meta/env_stock_trading/env_stocktrading.py).
//
// Mapping: ONE THREAD PER ENV for the arithmetic — the reference's sums are
// sequential and the buys form a serial cash chain, so lanes of a warp run 32 independent envs in
// lock-step instead of 32 lanes idling behind one chain — and ONE WARP PER 32-ENV TILE for the
// memory traffic: actions are staged through shared memory with coalesced loads, the stock-major
// state arrays are read/written coalesced straight from registers, and the observation rows (74 %
// of the bytes of a step) leave through the bulk-copy engine (TMA): a 4-row template image per day is
// bulk-loaded into shared memory, patched with the envs' cash / holdings and bulk-stored, four rows per
// cp.async.bulk (tiles that are partial, unaligned or not on one day fall back to 128-B warp stores).
//
// Bit-exactness: every product/sum on the cash path is an explicit __dmul_rn/__dadd_rn in the
// reference's order; np.argsort's tie order is reproduced by running the same bitonic network
//  on register-resident packed keys.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <thread>
#include <vector>

#include "common.cuh"
#include "sort_network.inc"
#include "trading_common.cuh"

// tuning knobs (A/B-tested on B200, see profiles/)
#ifndef FRL_LD_STREAM
#define FRL_LD_STREAM 1  // ld.global.cs for the read-once state/action streams
#endif
#ifndef FRL_ST_STREAM
#define FRL_ST_STREAM 0  // st.global.cs for the state write-back
#endif
#ifndef FRL_LOOP_PIPE
#define FRL_LOOP_PIPE 1  // software-pipelined sell / buy loops
#endif
#ifndef FRL_TRADING_MIN_BLOCKS
#define FRL_TRADING_MIN_BLOCKS 5  // 128-thread blocks per SM the allocator must allow: 5 -> 96 regs, no spills.  A/B with the copy-engine obs writer: 3 -> 0.354 ms, 4 -> 0.297, 5 -> 0.290, 6 -> 0.301 (before it, 4 was best: the register template of the row writer spilled at 96)
#endif

namespace frl {
namespace {

template <typename T>
__device__ __forceinline__ T ld_stream(const T *p)
{
#if FRL_LD_STREAM
    return __ldcs(p);
#else
    return *p;
#endif
}
template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v)
{
#if FRL_ST_STREAM
    __stcs(p, v);
#else
    *p = v;
#endif
}

// keys are (a << 5) | index; swap only on STRICT a[lo] > a[hi] — ties keep network order.
#define FRL_CEX(lo, hi)                                                                            \
    {                                                                                              \
        const bool sw_ = (lo) > ((hi) | 31);                                                       \
        const int t_ = sw_ ? (hi) : (lo);                                                          \
        (hi) = sw_ ? (lo) : (hi);                                                                  \
        (lo) = t_;                                                                                 \
    }

// np.argsort(int64) tie order for n <= 32: ascending bitonic network (flip stage + half-cleaners),
// emitted as straight-line code by gen_sort_network.py.
template <int SLOTS>
__device__ __forceinline__ void bitonic_network(int (&key)[SLOTS])
{
    static_assert(SLOTS == 8 || SLOTS == 16 || SLOTS == 32, "unsupported slot count");
    if constexpr (SLOTS == 8) {
        FRL_SORT_NETWORK_8(FRL_CEX, key)
    } else if constexpr (SLOTS == 16) {
        FRL_SORT_NETWORK_16(FRL_CEX, key)
    } else {
        FRL_SORT_NETWORK_32(FRL_CEX, key)
    }
}

constexpr int kHoldPitch = 33;  // hold_s[j][lane]: conflict-free both per-lane (compute) and per-row (obs)

#ifndef FRL_OBS_BULK
#define FRL_OBS_BULK 1  // one-day tiles: 4-row template image loaded by the copy engine (TMA); rows leave as
                        // 1 = bulk stores (A/B on B200: 0.306 ms), 2 = aligned 16-byte vector stores from the
                        // patched image (0.320), 0 = off, per-row 4-byte stores (0.320)
#endif
#ifndef FRL_OBS_PARTS
#define FRL_OBS_PARTS 1  // bulk stores per image: 1, 2 or 4
#endif
constexpr int kObsImgBytes = 4864;  // >= 16 * O for DOW-30 with 8 indicators (O = 301 -> 4816 B)

template <int SLOTS, typename ActT, int WARPS>
struct alignas(16) WarpSmem {
    static constexpr int kActBytes = 32 * SLOTS * (int)sizeof(ActT);
    static constexpr int kImgBytes = kActBytes > kObsImgBytes ? kActBytes : kObsImgBytes;
    // staged actions, flat [32 envs][D] exactly as they lie in global memory.  Once a lane has turned
    // ITS row into sort keys, the row is reused for that lane's sorted order (lane-private, so no
    // cross-lane hazard and no barrier).  After the trades the region is dead and becomes the 4-row
    // observation image of the bulk-copy writer.
    union alignas(16) {
        ActT act[32 * SLOTS];
        float img[kImgBytes / 4];
    };
    int hold[SLOTS * kHoldPitch];
    float cashf[32];
    int sd[32];
    unsigned long long mbar;  // completion barrier of the image load
};

// ---- bulk-copy (TMA) observation writer --------------------------------------------------------
// The 32 observation rows of a tile are ONE contiguous, 16-byte-aligned block of 32*O floats, and all but
// the cash / holdings slots of a row are the day's template.  So: bulk-load the day's 4-row template image
// (16*O bytes, obs_tmpl4) into shared memory once, and per 4 rows patch the 4*(D+1) env-specific floats
// and hand the image to the copy engine with one cp.async.bulk store — ~20 instructions per lane per four
// rows instead of ~60 stores and selects.

// all 32 envs of the tile valid and on day sd0: start the image load (the action region must be dead)
template <typename SM>
__device__ __forceinline__ void obs_image_load(const frl_trading_params &p, SM &sm, int lane, int sd0)
{
    fence_proxy_async_smem();
    __syncwarp();  // every lane is done with its action / order row; cashf and hold are visible
    if (lane == 0) {
        const unsigned bytes = 16u * (unsigned)p.obs_dim, mbar = smem_u32(&sm.mbar);
        const float *src = p.obs_tmpl4 + (size_t)sd0 * 4 * p.obs_dim;
        mbar_expect_tx(mbar, bytes);
        bulk_copy_g2s(smem_u32(sm.img), src, bytes, mbar);
    }
}

// ... and patch + store the 8 x 4 rows; otile = first row of the tile, 16-byte aligned
template <int DCT, typename SM>
__device__ __forceinline__ void write_obs_tile_bulk(const frl_trading_params &p, SM &sm, float *__restrict__ otile, int lane,
                                                    int sd0, unsigned &phase)
{
    const int O = p.obs_dim, D = DCT > 0 ? DCT : p.stock_dim;
    const unsigned bytes = 16u * (unsigned)O;
    const unsigned mbar = smem_u32(&sm.mbar);
    // the env-specific floats of the first four rows, fetched while the image is in flight
    const int *hcol = sm.hold + lane * kHoldPitch;
    float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f, cq = 0.f;
    if (lane < D) {
        h0 = (float)hcol[0];
        h1 = (float)hcol[1];
        h2 = (float)hcol[2];
        h3 = (float)hcol[3];
    }
    if (lane < 4) cq = sm.cashf[lane];
    mbar_wait(mbar, phase);
    phase ^= 1u;
    float *pimg = sm.img + D + 1 + lane;
#if FRL_OBS_BULK == 1
    // Copy engine: the image leaves in FRL_OBS_PARTS pieces (whole rows each, cut at 16-byte boundaries just
    // below a row start), one bulk store per piece, so a piece can be patched for the next four rows as soon
    // as ITS last store has been read while the stores of the other pieces are still in flight.
    constexpr int PARTS = FRL_OBS_PARTS, R = 4 / PARTS;
    const char *img_b = reinterpret_cast<const char *>(sm.img);
    char *out_b = reinterpret_cast<char *>(otile);
#pragma unroll 1
    for (int s = 0; s < 8; ++s) {
        float hv[4] = {h0, h1, h2, h3};
#pragma unroll
        for (int q = 0; q < PARTS; ++q) {
            if (s > 0) {
                if (lane == 0) bulk_wait_read<PARTS - 1>();  // this piece's previous store has read the image
                __syncwarp();
            }
            if (lane < D) {
#pragma unroll
                for (int r = q * R; r < (q + 1) * R; ++r) pimg[r * O] = hv[r];
            }
            if (lane >= q * R && lane < (q + 1) * R) sm.img[lane * O] = cq;
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                const unsigned off = (unsigned)(q * R * O * 4) & ~15u;
                const unsigned end = q + 1 == PARTS ? bytes : ((unsigned)((q + 1) * R * O * 4) & ~15u);
                bulk_store_s2g(out_b + (size_t)s * bytes + off, smem_u32(img_b + off), end - off);
            }
        }
        if (s < 7) {  // next four rows' values while the engine reads the image
            if (lane < D) {
                h0 = (float)hcol[4 * s + 4];
                h1 = (float)hcol[4 * s + 5];
                h2 = (float)hcol[4 * s + 6];
                h3 = (float)hcol[4 * s + 7];
            }
            if (lane < 4) cq = sm.cashf[4 * s + 4 + lane];
        }
    }
    if (lane == 0) bulk_wait_read<0>();  // the action region is restaged next
    __syncwarp();
#else
#pragma unroll 1
    for (int s = 0; s < 8; ++s) {
        if (lane < D) {
            pimg[0] = h0;
            pimg[O] = h1;
            pimg[2 * O] = h2;
            pimg[3 * O] = h3;
        }
        if (lane < 4) sm.img[lane * O] = cq;
        // the four rows are O aligned 16-byte vectors: LDS.128 -> STG.128, nothing to wait for
        __syncwarp();
        {
            const float4 *iv = reinterpret_cast<const float4 *>(sm.img) + lane;
            float4 *ov = reinterpret_cast<float4 *>(otile + (size_t)s * 4 * O) + lane;
            constexpr int NV = DCT == 30 ? 10 : 16;  // ceil(O / 32) vectors per lane
#pragma unroll
            for (int c0 = 0; c0 < NV; c0 += 5) {
                float4 v[5];
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    if (c0 + c < NV && lane + 32 * (c0 + c) < O) v[c] = iv[32 * (c0 + c)];
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    if (c0 + c < NV && lane + 32 * (c0 + c) < O) ov[32 * (c0 + c)] = v[c];
            }
        }
        if (s < 7) {
            if (lane < D) {
                h0 = (float)hcol[4 * s + 4];
                h1 = (float)hcol[4 * s + 5];
                h2 = (float)hcol[4 * s + 6];
                h3 = (float)hcol[4 * s + 7];
            }
            if (lane < 4) cq = sm.cashf[4 * s + 4 + lane];
        }
        __syncwarp();
    }
#endif
}


// sequential Python sum() of price*holding over the D stocks, then cash + that (:311-314, :344-347)
template <int SLOTS>
__device__ __forceinline__ double total_asset(double cash, const double *__restrict__ prow, const int *hold_s,
                                              int lane, int D)
{
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        if (j < D) acc = dadd(acc, dmul(__ldg(prow + j), (double)hold_s[j * kHoldPitch + lane]));
    }
    return dadd(cash, acc);
}

// Warp-cooperative write of the 32 observation rows of a tile: row r = [cash, close[sd] x D,
// holdings x D, tech[.][sd] x K*D] as float32.  Everything except the cash and holdings slots comes
// from the per-day template row, which is identical for every env of the tile when they are in
// lock-step (the common case): then it is loaded once into registers and only stored 32 times —
// NCH store instructions of 128 contiguous bytes per row, plus the cash/holdings patch-up of the
// first chunks.
template <int NCH, int DCT, typename SM>
__device__ __forceinline__ void write_obs_rows_uniform(const frl_trading_params &p, SM &sm, float *__restrict__ obs,
                                                       long long env0, int nvalid, int lane, int sd0)
{
    const int O = p.obs_dim, D = DCT > 0 ? DCT : p.stock_dim;
    const int special_end = 2 * D + 1;  // positions [0, special_end) hold cash / close / holdings
    float t[NCH];
    const float *trow = p.obs_tmpl + (size_t)sd0 * O + lane;
#pragma unroll
    for (int c = 0; c < NCH; ++c) t[c] = (c < NCH - 1 || lane + 32 * c < O) ? __ldg(trow + 32 * c) : 0.0f;
    // offsets into sm.hold for the (at most 3) chunks that overlap the holdings slots
    constexpr int NSP = NCH < 3 ? NCH : 3;
    int hoff[NSP];
#pragma unroll
    for (int c = 0; c < NSP; ++c) {
        const int pos = lane + 32 * c;
        hoff[c] = (pos > D && pos < special_end) ? (pos - 1 - D) * kHoldPitch : -1;
    }
    const bool third = special_end > 64;  // only D == 32 reaches the third chunk
    const bool tail_ok = lane + 32 * (NCH - 1) < O;
    float *orow = obs + (size_t)env0 * O + lane;
#pragma unroll 2
    for (int r = 0; r < nvalid; ++r) {
        float v[NSP];
#pragma unroll
        for (int c = 0; c < NSP; ++c) v[c] = t[c];
        const float cashf = sm.cashf[r];
        if (hoff[0] >= 0) v[0] = (float)sm.hold[hoff[0] + r];
        if (lane == 0) v[0] = cashf;
        if (NSP > 1 && hoff[1] >= 0) v[1] = (float)sm.hold[hoff[1] + r];
        if (NSP > 2 && third && hoff[2] >= 0) v[2] = (float)sm.hold[hoff[2] + r];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float x = c < NSP ? v[c] : t[c];
            if (c < NCH - 1 || tail_ok) orow[32 * c] = x;
        }
        orow += O;
    }
}

template <int DCT, typename SM>
__device__ __forceinline__ void write_obs_tile(const frl_trading_params &p, SM &sm, float *__restrict__ obs,
                                               long long env0, int nvalid, int lane)
{
    const int O = p.obs_dim, D = DCT > 0 ? DCT : p.stock_dim;
    const int sd0 = sm.sd[0];
    bool uniform = true;
    if (lane < nvalid) uniform = (sm.sd[lane] == sd0);
    uniform = __all_sync(0xffffffffu, uniform);
    const int nch = (O + 31) >> 5;
    if (uniform && nch <= 12) {
        switch (nch) {
#define FRL_CASE(N)                                                                                \
    case N:                                                                                        \
        write_obs_rows_uniform<N, DCT>(p, sm, obs, env0, nvalid, lane, sd0);                            \
        break;
            FRL_CASE(1) FRL_CASE(2) FRL_CASE(3) FRL_CASE(4) FRL_CASE(5) FRL_CASE(6)
            FRL_CASE(7) FRL_CASE(8) FRL_CASE(9) FRL_CASE(10) FRL_CASE(11) FRL_CASE(12)
#undef FRL_CASE
        }
    } else {
        const int special_end = 2 * D + 1;
        for (int r = 0; r < nvalid; ++r) {
            float *orow = obs + (size_t)(env0 + r) * O;
            const float *trow = p.obs_tmpl + (size_t)sm.sd[r] * O;
            const float cashf = sm.cashf[r];
            for (int pos = lane; pos < O; pos += 32) {
                float v = __ldg(trow + pos);
                if (pos == 0)
                    v = cashf;
                else if (pos > D && pos < special_end)
                    v = (float)sm.hold[(pos - 1 - D) * kHoldPitch + r];
                orow[pos] = v;
            }
        }
    }
}

// DCT > 0 compiles the stock count in (DOW-30 fast path: every `j < D` guard and the sort-network
// pads fold away); DCT == 0 reads it from the params.
template <int SLOTS, int DCT, typename ActT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, FRL_TRADING_MIN_BLOCKS * 128 / (WARPS * 32))
trading_rollout_kernel(const frl_trading_params p, const ActT *__restrict__ actions, long long act_step_stride,
                       long long act_env_stride, int n_steps, double *__restrict__ rewards,
                       uint8_t *__restrict__ flags_out, float *__restrict__ obs, int obs_mode, int auto_reset,
                       double *__restrict__ stats)
{
    stats_exchange_previous(stats);
    using SM = WarpSmem<SLOTS, ActT, WARPS>;
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SM &sm = reinterpret_cast<SM *>(smem_dyn)[warp];
    const int N = p.n_envs, D = DCT > 0 ? DCT : p.stock_dim, T = p.n_days;
    const int ld = p.env_stride;  // 32-bit index math: SLOTS * env_stride < 2^31 is validated on the host
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;  // whole warp out of range (no block-level barriers are used)
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const bool valid = lane < nvalid;
    const long long n = valid ? env0 + lane : (long long)N - 1;
    unsigned img_phase = 0;
#if FRL_OBS_BULK
    const bool bulk_ok = p.obs_tmpl4 != nullptr && nvalid == 32 && 16 * p.obs_dim <= SM::kImgBytes && p.obs_dim <= 512 &&
                         p.n_tech * D >= 4;  // a piece boundary never cuts through a row's cash / holdings slots
    if (bulk_ok && lane == 0) mbar_init(smem_u32(&sm.mbar), 1);
#else
    const bool bulk_ok = false;
#endif

    // ---- load state: one thread per env, stock-major holdings => coalesced ----
    double cash = p.cash[n], cost = p.cost[n], last_reward = p.reward[n];
    int day = p.day[n], sday = p.sday[n], trades = p.trades[n];
    const bool act_flat = act_env_stride == D;
    {
        // all D holdings loads and the first step's action loads are issued back to back (independent, one
        // DRAM round trip), then parked in shared memory
        int hv[SLOTS];
        const int *hp = p.hold + n;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) hv[j] = (j < D) ? ld_stream(hp + j * ld) : 0;
        if (act_flat) {
            ActT av[SLOTS];
            stage_actions_load<SLOTS, ActT>(av, actions, env0, D, nvalid, lane);
            stage_actions_store<SLOTS, ActT>(sm.act, av, D, lane);
        }
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) sm.hold[j * kHoldPitch + lane] = hv[j];
    }

    const double one_minus_sc = dsub(1.0, p.sell_cost_pct), one_plus_bc = dadd(1.0, p.buy_cost_pct);
    const int hmax_i = (int)max(-(double)kMaxAbsAction, min((double)kMaxAbsAction, p.hmax));

    double asset = 0.0;        // total asset of the current state when asset_ok
    bool asset_ok = false;     // carried from the previous step of this launch (bitwise equal to a recompute)
    double st_r = 0.0, st_r2 = 0.0, st_done = 0.0, st_epi = 0.0, st_liq = 0.0;

    for (int k = 0; k < n_steps; ++k) {
        // ---- stage this step's actions for the tile (coalesced) ----
        const ActT *abase = actions + (size_t)k * act_step_stride;
        __syncwarp();
        if (k > 0 || !act_flat) stage_actions_flat<SLOTS, ActT>(sm.act, abase, env0, act_env_stride, D, nvalid, lane);
        __syncwarp();

        uint8_t flags = 0;
        double reward;
        double begin = 0.0;
        bool liq = false, stepped = false;
        if (day >= T - 1) {
            // ---- terminal branch (:221-301): no state change, previous scaled reward again (Q3) ----
            flags = FRL_FLAG_DONE;
            reward = last_reward;
            if (valid) {
                const double *prow = p.close + (size_t)state_day(sday) * p.close_pitch;
                if (!asset_ok) asset = total_asset<SLOTS>(cash, prow, sm.hold, lane, D);
                st_done += 1.0;
                st_epi += asset;
            }
            if (auto_reset) {
                // DummyVecEnv.step_wait -> env.reset() (:359-393): state list rebuilt from the rows
                // of the day still loaded, THEN day = 0 (stale-day quirk Q1)
                cash = p.initial_amount;
#pragma unroll
                for (int j = 0; j < SLOTS; ++j)
                    if (j < D) sm.hold[j * kHoldPitch + lane] = p.init_hold ? __ldg(p.init_hold + j) : 0;
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
            liq = p.use_turbulence && (turb >= p.turbulence_threshold);
            const double *prow = p.close + (size_t)sd * p.close_pitch;
            begin = asset_ok ? asset : total_asset<SLOTS>(cash, prow, sm.hold, lane, D);

            if (liq) {
                // actions = [-hmax]*D (:308-310): all keys tie, the network never swaps, so the sell
                // order is the index order; liquidation checks price > 0, not the disable flag (:138-163)
                flags = FRL_FLAG_LIQUIDATE;
                if (hmax_i > 0) {
#pragma unroll
                    for (int j = 0; j < SLOTS; ++j) {
                        if (j < D) {
                            const double pj = __ldg(prow + j);
                            const int h = sm.hold[j * kHoldPitch + lane];
                            if (pj > 0.0 && h > 0) {
                                const double pv = dmul(pj, (double)h);
                                cash = dadd(cash, dmul(pv, one_minus_sc));
                                sm.hold[j * kHoldPitch + lane] = 0;
                                cost = dadd(cost, dmul(pv, p.sell_cost_pct));
                                trades += 1;
                            }
                        }
                    }
                }
            } else {
                // ---- (actions * hmax).astype(int), packed sort keys, np.argsort order ----
                int key[SLOTS];
                const ActT *arow = sm.act + lane * D;
#pragma unroll
                for (int j = 0; j < SLOTS; ++j)
                    key[j] = (j < D) ? action_to_shares<ActT>(arow[j], p.hmax) * 32 + j : 0x7fffffff;
                bitonic_network<SLOTS>(key);
                // a < 0 <=> key < 0 and a > 0 <=> key >= 32, so the sorted list itself delimits the
                // sell prefix (argsort[:n_neg]) and the buy suffix (argsort[::-1][:n_pos])
                int *ord = reinterpret_cast<int *>(sm.act) + lane * (D * (int)(sizeof(ActT) / sizeof(int)));
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
                    if (s < D) ord[s] = key[s];
                const uint32_t dis = p.disable_mask ? __ldg(p.disable_mask + sd) : 0u;

#if FRL_LOOP_PIPE
                // Both loops are software-pipelined: the next order entry, its price (and holding) are fetched
                // before the current trade's dependent fp64 chain, so the chain never waits for a load.
                // ---- sells, most negative first (:321-324, _sell_stock :102-135) ----
                {
                    int kk = ord[0];
                    double pj = __ldg(prow + (kk & 31));
                    int h = sm.hold[(kk & 31) * kHoldPitch + lane];
                    for (int s = 0; s < D; ++s) {
                        if (kk >= 0) break;
                        const int kn = ord[min(s + 1, D - 1)];
                        const double pn = __ldg(prow + (kn & 31));
                        const int hn = sm.hold[(kn & 31) * kHoldPitch + lane];  // another stock: not touched below
                        const int a = kk >> 5, j = kk & 31;
                        if (!((dis >> j) & 1u) && h > 0) {
                            const int m = min(-a, h);
                            const double pv = dmul(pj, (double)m);
                            cash = dadd(cash, dmul(pv, one_minus_sc));
                            sm.hold[j * kHoldPitch + lane] = h - m;
                            cost = dadd(cost, dmul(pv, p.sell_cost_pct));
                            trades += 1;
                        }
                        kk = kn;
                        pj = pn;
                        h = hn;
                    }
                }
                // ---- buys, largest first, each limited by the cash left (:328-330, _buy_stock :171-201) ----
                {
                    int kk = ord[D - 1];
                    double pj = __ldg(prow + (kk & 31));
                    double unit = dmul(pj, one_plus_bc);
                    for (int s = D - 1; s >= 0; --s) {
                        if (kk < 32) break;
                        const int kn = ord[max(s - 1, 0)];
                        const double pn = __ldg(prow + (kn & 31));
                        const double un = dmul(pn, one_plus_bc);
                        const int a = kk >> 5, j = kk & 31;
                        if (!((dis >> j) & 1u)) {
                            double nsh = (double)a;
                            trades += 1;  // even when 0 shares end up bought (Q5)
                            bool buy = true;
                            // min(cash // unit, a): the quotient only matters when cash < (a+1)*unit ...
                            if (!(cash >= dmul(nsh + 1.0, unit))) {
                                // ... and when not even one share is affordable it is exactly 0: buying 0 shares
                                // changes nothing (x - 0.0 and x + 0.0 are identities) — the common state of a
                                // cash-starved env, which then skips the division and the whole update
                                if (cash >= 0.0 && cash < unit) {
                                    buy = false;
                                } else {
                                    const double avail = floor_div_f64(cash, unit);
                                    nsh = (nsh < avail) ? nsh : avail;
                                }
                            }
                            if (buy) {
                                const double pv = dmul(pj, nsh);
                                cash = dsub(cash, dmul(pv, one_plus_bc));
                                sm.hold[j * kHoldPitch + lane] += (int)nsh;
                                cost = dadd(cost, dmul(pv, p.buy_cost_pct));
                            }
                        }
                        kk = kn;
                        pj = pn;
                        unit = un;
                    }
                }
#else
                // ---- sells, most negative first (:321-324, _sell_stock :102-135) ----
                for (int s = 0; s < D; ++s) {
                    const int kk = ord[s];
                    if (kk >= 0) break;
                    const int a = kk >> 5, j = kk & 31;
                    const int h = sm.hold[j * kHoldPitch + lane];
                    if (!((dis >> j) & 1u) && h > 0) {
                        const int m = min(-a, h);
                        const double pv = dmul(__ldg(prow + j), (double)m);
                        cash = dadd(cash, dmul(pv, one_minus_sc));
                        sm.hold[j * kHoldPitch + lane] = h - m;
                        cost = dadd(cost, dmul(pv, p.sell_cost_pct));
                        trades += 1;
                    }
                }
                // ---- buys, largest first, each limited by the cash left (:328-330, _buy_stock :171-201) ----
                for (int s = D - 1; s >= 0; --s) {
                    const int kk = ord[s];
                    if (kk < 32) break;
                    const int a = kk >> 5, j = kk & 31;
                    if (!((dis >> j) & 1u)) {
                        const double pj = __ldg(prow + j);
                        const double unit = dmul(pj, one_plus_bc);
                        double nsh = (double)a;
                        trades += 1;  // even when 0 shares end up bought (Q5)
                        // min(cash // unit, a): the quotient only matters when cash < (a+1)*unit ...
                        if (!(cash >= dmul(nsh + 1.0, unit))) {
                            // ... and when not even one share is affordable it is exactly 0: buying 0 shares
                            // changes nothing (x - 0.0 and x + 0.0 are identities) — the common state of a
                            // cash-starved env, which then skips the division and the whole update
                            if (cash >= 0.0 && cash < unit) continue;
                            const double avail = floor_div_f64(cash, unit);
                            nsh = (nsh < avail) ? nsh : avail;
                        }
                        const double pv = dmul(pj, nsh);
                        cash = dsub(cash, dmul(pv, one_plus_bc));
                        sm.hold[j * kHoldPitch + lane] += (int)nsh;
                        cost = dadd(cost, dmul(pv, p.buy_cost_pct));
                    }
                }
#endif
            }
            // ---- state: s -> s+1 (:335-352) ----
            day += 1;
            sday = day;
            stepped = true;
        }
        // the action region is dead from here on: start the observation image load, then value the portfolio
        const bool want_obs = obs_mode == FRL_OBS_ALL || (obs_mode == FRL_OBS_LAST && k == n_steps - 1);
        float *otile = nullptr;
        bool bulk = false;
        int sd0 = 0;
        if (want_obs) {
            const int sdn = state_day(sday);
            sm.cashf[lane] = (float)cash;
            sm.sd[lane] = sdn;
            otile = obs + (obs_mode == FRL_OBS_ALL ? (size_t)k * N * p.obs_dim : (size_t)0) + (size_t)env0 * p.obs_dim;
            sd0 = __shfl_sync(0xffffffffu, sdn, 0);
            bulk = bulk_ok && (reinterpret_cast<uintptr_t>(otile) & 15) == 0 && __all_sync(0xffffffffu, sdn == sd0);
            if (bulk) obs_image_load(p, sm, lane, sd0);
        }
        if (stepped) {
            asset = total_asset<SLOTS>(cash, p.close + (size_t)day * p.close_pitch, sm.hold, lane, D);
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
        if (want_obs) {
            if (bulk) {
                write_obs_tile_bulk<DCT>(p, sm, otile, lane, sd0, img_phase);
            } else {
                __syncwarp();
                write_obs_tile<DCT>(p, sm, otile - (size_t)env0 * p.obs_dim, env0, nvalid, lane);
            }
        }
    }

    // ---- store state ----
    if (valid) {
        p.cash[n] = cash;
        p.cost[n] = cost;
        p.reward[n] = last_reward;
        p.day[n] = day;
        p.sday[n] = sday;
        p.trades[n] = trades;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
            if (j < D) st_stream(p.hold + n + j * ld, sm.hold[j * kHoldPitch + lane]);
    }
    if (p.asset_out) {
        if (!asset_ok) {
            asset = total_asset<SLOTS>(cash, p.close + (size_t)state_day(sday) * p.close_pitch, sm.hold, lane, D);
            asset_ok = true;
        }
        if (valid) p.asset_out[n] = asset;
    }
    if (stats) {
        double fin_asset = 0.0, fin_trades = 0.0, steps = 0.0;
        if (valid) {
            if (!asset_ok)
                asset = total_asset<SLOTS>(cash, p.close + (size_t)state_day(sday) * p.close_pitch, sm.hold, lane, D);
            fin_asset = asset;
            fin_trades = (double)trades;
            steps = (double)n_steps;
        }
        double v[FRL_N_STATS] = {st_r, st_r2, st_done, st_epi, fin_asset, st_liq, steps, fin_trades};
        reduce_stats8(v, lane, stats);
    }
}

// ---- init / reset / observe ------------------------------------------------------------------
__global__ void trading_init_kernel(const frl_trading_params p, int day0)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.n_envs) return;
    p.cash[n] = p.initial_amount;
    for (int j = 0; j < p.stock_dim; ++j) p.hold[(size_t)j * p.env_stride + n] = p.init_hold ? p.init_hold[j] : 0;
    p.day[n] = day0;
    p.sday[n] = -day0 - 1;
    p.cost[n] = 0.0;
    p.trades[n] = 0;
    p.reward[n] = 0.0;
    p.episode[n] = 0;
}

__global__ void trading_reset_kernel(const frl_trading_params p, const uint8_t *__restrict__ mask)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.n_envs) return;
    if (mask && !mask[n]) return;
    p.cash[n] = p.initial_amount;
    for (int j = 0; j < p.stock_dim; ++j) p.hold[(size_t)j * p.env_stride + n] = p.init_hold ? p.init_hold[j] : 0;
    p.sday[n] = -p.day[n] - 1;  // state list built from the rows still loaded (Q1) ...
    p.day[n] = 0;               // ... then day = 0
    p.cost[n] = 0.0;
    p.trades[n] = 0;
    p.episode[n] += 1;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) trading_observe_kernel(const frl_trading_params p, float *__restrict__ obs)
{
    struct SM {  // what write_obs_tile reads
        int hold[32 * kHoldPitch];
        float cashf[32];
        int sd[32];
    };
    __shared__ SM smem[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SM &sm = smem[warp];
    const int N = p.n_envs, D = p.stock_dim;
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const long long n = lane < nvalid ? env0 + lane : (long long)N - 1;
    for (int j = 0; j < D; ++j) sm.hold[j * kHoldPitch + lane] = p.hold[(size_t)j * p.env_stride + n];
    sm.cashf[lane] = (float)p.cash[n];
    sm.sd[lane] = state_day(p.sday[n]);
    __syncwarp();
    write_obs_tile<0>(p, sm, obs, env0, nvalid, lane);
}

// Factored observation: per 32-env tile the env-specific slots [cash, holdings x D] as float32 rows (one
// contiguous 32*(1+D)-float run per tile) and the state-list day.  Stock-major loads and row-major stores are
// both coalesced through a [D+1][33] shared-memory tile.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
trading_observe_factored_kernel(const frl_trading_params p, float *__restrict__ env_part, int32_t *__restrict__ sday_out)
{
    extern __shared__ float fsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.n_envs, D = p.stock_dim, W = D + 1;
    float *tile = fsm + (size_t)warp * W * 33;
    const long long env0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (env0 >= N) return;
    const int nvalid = (int)min((long long)32, (long long)N - env0);
    const long long n = lane < nvalid ? env0 + lane : (long long)N - 1;
    tile[lane] = (float)p.cash[n];
    for (int j = 0; j < D; ++j) tile[(j + 1) * 33 + lane] = (float)p.hold[(size_t)j * p.env_stride + n];
    if (lane < nvalid) sday_out[n] = state_day(p.sday[n]);
    __syncwarp();
    float *out = env_part + (size_t)env0 * W;
    for (int i = lane; i < nvalid * W; i += 32) {
        const int r = i / W, c = i - r * W;
        out[i] = tile[c * 33 + r];
    }
}

// D > 32: one warp per env, holdings straight from the stock-major array (only used by observe/reset)
__global__ void trading_observe_wide_kernel(const frl_trading_params p, float *__restrict__ obs)
{
    const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= p.n_envs) return;
    const int O = p.obs_dim, D = p.stock_dim;
    const float *trow = p.obs_tmpl + (size_t)state_day(p.sday[n]) * O;
    float *orow = obs + (size_t)n * O;
    for (int pos = lane; pos < O; pos += 32) {
        float v = __ldg(trow + pos);
        if (pos == 0)
            v = (float)p.cash[n];
        else if (pos > D && pos <= 2 * D)
            v = (float)p.hold[(size_t)(pos - 1 - D) * p.env_stride + n];
        orow[pos] = v;
    }
}

int32_t validate(const frl_trading_params *p)
{
    FRL_REQUIRE(p != nullptr, "trading: params is NULL");
    FRL_REQUIRE(p->n_envs >= 1, "trading: n_envs must be >= 1 (got %d)", p->n_envs);
    FRL_REQUIRE(p->stock_dim >= 1 && p->stock_dim <= 128, "trading: stock_dim must be in 1..128 (got %d)", p->stock_dim);
    FRL_REQUIRE((p->close_pitch == 32 || p->close_pitch == 128) && p->close_pitch >= p->stock_dim,
                "trading: close_pitch must be 32 or 128 and >= stock_dim (got %d for D=%d)", p->close_pitch, p->stock_dim);
    FRL_REQUIRE(p->n_tech >= 0 && p->n_days >= 1, "trading: bad n_tech/n_days (%d, %d)", p->n_tech, p->n_days);
    FRL_REQUIRE(p->obs_dim == 1 + 2 * p->stock_dim + p->n_tech * p->stock_dim,
                "trading: obs_dim %d != 1 + 2D + K*D = %d", p->obs_dim, 1 + 2 * p->stock_dim + p->n_tech * p->stock_dim);
    FRL_REQUIRE(p->env_stride >= p->n_envs, "trading: env_stride %d < n_envs %d", p->env_stride, p->n_envs);
    FRL_REQUIRE((long long)p->env_stride * p->close_pitch < (1LL << 31), "trading: env_stride %d too large (pitch*stride must be < 2^31)",
                p->env_stride);
    FRL_REQUIRE(p->close && p->risk && p->obs_tmpl, "trading: table pointer is NULL");
    FRL_REQUIRE(p->cash && p->hold && p->day && p->sday && p->cost && p->trades && p->reward && p->episode,
                "trading: state pointer is NULL");
    return FRL_OK;
}

// n_envs at or below which frl_trading_rollout uses the 8-lanes-per-env kernel; FRL_TRADING_KERNEL=tile|small
// or FRL_TRADING_SMALL_MAX set the default, frl_set_option("trading_small_max", v) changes it at run time.
int g_small_max = -1;
int trading_small_max()
{
    if (g_small_max < 0) {
        const char *f = getenv("FRL_TRADING_KERNEL");
        const char *m = getenv("FRL_TRADING_SMALL_MAX");
        if (f && !strcmp(f, "tile"))
            g_small_max = 0;
        else if (f && !strcmp(f, "small"))
            g_small_max = 0x7fffffff;
        else
            g_small_max = m ? atoi(m) : 8192;
    }
    return g_small_max;
}

// D > 32: batches above this many envs run in the thread-per-env kernel of trading_wide.cu (keys and holdings in
// shared memory), smaller ones in the 8-lanes-per-env kernel; frl_set_option("trading_wide_min_envs", v)
int g_wide_min_envs = -1;
int trading_wide_min_envs()
{
    if (g_wide_min_envs < 0) {
        const char *m = getenv("FRL_TRADING_WIDE_MIN_ENVS");
        g_wide_min_envs = m ? atoi(m) : 3072;  // measured crossover at D=100: 2048 -> 0.053 vs 0.070 ms, 4096 -> 0.108 vs 0.070
    }
    return g_wide_min_envs;
}

template <int SLOTS, int DCT, typename ActT, int WARPS>
void launch_rollout(const frl_trading_params &p, const void *actions, long long sstride, long long estride, int n_steps,
                    double *rewards, uint8_t *flags, float *obs, int obs_mode, int auto_reset, double *stats,
                    cudaStream_t st)
{
    const long long tiles = ((long long)p.n_envs + 31) / 32;
    const unsigned grid = (unsigned)((tiles + WARPS - 1) / WARPS);
    auto kern = trading_rollout_kernel<SLOTS, DCT, ActT, WARPS>;
    constexpr int smem = WARPS * (int)sizeof(WarpSmem<SLOTS, ActT, WARPS>);
    static_assert(smem <= 48 * 1024, "the tile kernel's shared memory must stay within the default limit");
    kern<<<grid, WARPS * 32, smem, st>>>(
        p, (const ActT *)actions, sstride, estride, n_steps, rewards, flags, obs, obs_mode, auto_reset, stats);
}

}  // namespace
}  // namespace frl

namespace frl {
int32_t np_set_option(const char *name, int64_t value);  // nptrading.cu
extern int g_tw_regs;                                           // trading_wide.cu
}

using namespace frl;

extern "C" int32_t frl_set_option(const char *name, int64_t value)
{
    FRL_REQUIRE(name != nullptr, "set_option: name is NULL");
    if (!strcmp(name, "trading_wide_min_envs")) {
        FRL_REQUIRE(value >= 0, "set_option: trading_wide_min_envs must be >= 0");
        g_wide_min_envs = value > 0x7fffffff ? 0x7fffffff : (int)value;
        return FRL_OK;
    }
    if (!strcmp(name, "trading_small_max")) {
        FRL_REQUIRE(value >= 0, "set_option: trading_small_max must be >= 0");
        g_small_max = value > 0x7fffffff ? 0x7fffffff : (int)value;
        return FRL_OK;
    }
    if (!strcmp(name, "trading_wide_regs")) {  // 0: D = 100 stays on the generic (runtime stock count) wide kernel
        g_tw_regs = value != 0;
        return FRL_OK;
    }
    if (np_set_option(name, value) == FRL_OK) return FRL_OK;
    set_error("set_option: unknown option '%s'", name);
    return FRL_E_INVALID;
}

extern "C" int32_t frl_trading_init(const frl_trading_params *p, int32_t day0, void *stream)
{
    if (int32_t rc = validate(p)) return rc;
    FRL_REQUIRE(day0 >= 0 && day0 < p->n_days, "trading_init: day0 %d outside [0, %d)", day0, p->n_days);
    trading_init_kernel<<<(p->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, day0);
    return check_launch("trading_init");
}

extern "C" int32_t frl_trading_observe(const frl_trading_params *p, float *obs, void *stream)
{
    if (int32_t rc = validate(p)) return rc;
    FRL_REQUIRE(obs != nullptr, "trading_observe: obs is NULL");
    constexpr int W = 4;
    if (p->stock_dim > 32) {
        const long long threads = (long long)p->n_envs * 32;
        trading_observe_wide_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, obs);
        return check_launch("trading_observe");
    }
    const long long tiles = ((long long)p->n_envs + 31) / 32;
    trading_observe_kernel<W><<<(unsigned)((tiles + W - 1) / W), W * 32, 0, (cudaStream_t)stream>>>(*p, obs);
    return check_launch("trading_observe");
}

extern "C" int32_t frl_trading_observe_factored(const frl_trading_params *p, float *env_part, int32_t *state_day_out,
                                                void *stream)
{
    if (int32_t rc = validate(p)) return rc;
    FRL_REQUIRE(env_part != nullptr && state_day_out != nullptr, "trading_observe_factored: output pointer is NULL");
    constexpr int W = 4;
    const long long tiles = ((long long)p->n_envs + 31) / 32;
    const size_t smem = (size_t)W * (p->stock_dim + 1) * 33 * sizeof(float);  // <= 68 KB at D = 128
    auto kern = trading_observe_factored_kernel<W>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)((tiles + W - 1) / W), W * 32, smem, (cudaStream_t)stream>>>(*p, env_part, state_day_out);
    return check_launch("trading_observe_factored");
}

namespace {
// Rows [lo, hi) of the dense observation from their factored form.  Rows are built in a small thread-local block
// (L1/L2-resident) and leave with NON-TEMPORAL stores: a plain memcpy into the 1.2 GB destination would first read every
// line it is about to overwrite (read-for-ownership), doubling the memory traffic of what is a pure write stream.
// Returns false when a state_day lies outside the table.
bool expand_rows(const float *tmpl, int n_days, int O, int D, const float *env_part, const int32_t *sday, float *out,
                 int64_t lo, int64_t hi, std::vector<float> &scratch)
{
    constexpr int kBlockRows = 32;
    const int W = D + 1;
    bool ok = true;
    scratch.resize((size_t)kBlockRows * O + 4);
    for (int64_t i0 = lo; i0 < hi; i0 += kBlockRows) {
        const int rows = (int)std::min<int64_t>(kBlockRows, hi - i0);
        for (int r = 0; r < rows; ++r) {
            const int64_t i = i0 + r;
            const int32_t d = sday[i];
            float *row = scratch.data() + (size_t)r * O;
            if (d < 0 || d >= n_days) {
                ok = false;
                memset(row, 0, sizeof(float) * O);
                continue;
            }
            memcpy(row, tmpl + (size_t)d * O, sizeof(float) * O);
            const float *e = env_part + (size_t)i * W;
            row[0] = e[0];
            memcpy(row + 1 + D, e + 1, sizeof(float) * D);
        }
        float *dst = out + (size_t)i0 * O;
        const float *src = scratch.data();
        size_t cnt = (size_t)rows * O;
#if defined(__SSE2__)
        while (cnt && (reinterpret_cast<uintptr_t>(dst) & 15)) {  // up to three floats before the first 16-byte boundary
            *dst++ = *src++;
            --cnt;
        }
        for (; cnt >= 4; cnt -= 4, dst += 4, src += 4) _mm_stream_ps(dst, _mm_loadu_ps(src));
#endif
        for (; cnt; --cnt) *dst++ = *src++;
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
    return ok;
}
}  // namespace

extern "C" int32_t frl_expand_obs_host_chunks(const float *tmpl, int32_t n_days, int32_t obs_dim, int32_t stock_dim,
                                              const float *env_part, const int32_t *sday, float *out, int32_t n_chunks,
                                              const int64_t *chunk_start, const int64_t *chunk_count, void *const *events,
                                              int32_t n_threads)
{
    FRL_REQUIRE(tmpl && env_part && sday && out, "expand_obs_host: NULL argument");
    FRL_REQUIRE(n_days >= 1 && stock_dim >= 1 && obs_dim >= 1 + 2 * stock_dim, "expand_obs_host: bad sizes (T=%d, O=%d, D=%d)",
                n_days, obs_dim, stock_dim);
    FRL_REQUIRE(n_chunks >= 0 && (n_chunks == 0 || (chunk_start && chunk_count)), "expand_obs_host: bad chunk list");
    int64_t total = 0;
    for (int c = 0; c < n_chunks; ++c) {
        FRL_REQUIRE(chunk_start[c] >= 0 && chunk_count[c] >= 0, "expand_obs_host: negative chunk bounds");
        total += chunk_count[c];
    }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (total + 4095) / 4096));
    std::atomic<int> bad(0), cuda_err(0);
    // every thread walks the chunks in order, waits for the chunk's event (its factored data has landed in host
    // memory) and expands pieces of the chunk: the expansion of chunk c overlaps the transfers of chunks c+1..
    // Pieces of kPiece rows are handed out by a per-chunk counter rather than cut into one fixed slice per thread: on a
    // shared host a thread that loses its CPU for a moment would otherwise hold up the whole step.
#ifndef FRL_EXPAND_PIECE
#define FRL_EXPAND_PIECE 2048
#endif
    constexpr int64_t kPiece = FRL_EXPAND_PIECE;
    std::vector<std::atomic<int64_t>> next(n_chunks > 0 ? n_chunks : 1);
    for (auto &a : next) a.store(0);
    auto work = [&](int) {
        std::vector<float> scratch;
        for (int c = 0; c < n_chunks; ++c) {
            if (events && events[c]) {
                const cudaError_t e = cudaEventSynchronize((cudaEvent_t)events[c]);
                if (e != cudaSuccess) cuda_err.store((int)e);
            }
            for (;;) {
                const int64_t off = next[c].fetch_add(kPiece);
                if (off >= chunk_count[c]) break;
                const int64_t lo = chunk_start[c] + off, hi = chunk_start[c] + std::min(chunk_count[c], off + kPiece);
                if (!expand_rows(tmpl, n_days, obs_dim, stock_dim, env_part, sday, out, lo, hi, scratch)) bad.store(1);
            }
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    if (cuda_err.load()) {
        set_error("expand_obs_host: cudaEventSynchronize failed (%s)", cudaGetErrorString((cudaError_t)cuda_err.load()));
        return FRL_E_CUDA;
    }
    FRL_REQUIRE(bad.load() == 0, "expand_obs_host: a state_day lies outside [0, %d)", n_days);
    return FRL_OK;
}

extern "C" int32_t frl_expand_obs_host(const float *tmpl, int32_t n_days, int32_t obs_dim, int32_t stock_dim,
                                       const float *env_part, const int32_t *sday, int64_t n, float *out, int32_t n_threads)
{
    FRL_REQUIRE(n >= 0, "expand_obs_host: n must be >= 0 (got %lld)", (long long)n);
    const int64_t start = 0;
    return frl_expand_obs_host_chunks(tmpl, n_days, obs_dim, stock_dim, env_part, sday, out, 1, &start, &n, nullptr, n_threads);
}

extern "C" int32_t frl_trading_reset(const frl_trading_params *p, const uint8_t *mask, float *obs, void *stream)
{
    if (int32_t rc = validate(p)) return rc;
    trading_reset_kernel<<<(p->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, mask);
    if (int32_t rc = check_launch("trading_reset")) return rc;
    if (obs) return frl_trading_observe(p, obs, stream);
    return FRL_OK;
}

extern "C" int32_t frl_trading_rollout(const frl_trading_params *p, const void *actions, int32_t actions_f64,
                                       int64_t act_step_stride, int64_t act_env_stride, int32_t n_steps,
                                       double *rewards, uint8_t *flags, float *obs, int32_t obs_mode,
                                       int32_t auto_reset, double *stats, void *stream)
{
    if (int32_t rc = validate(p)) return rc;
    FRL_REQUIRE(actions != nullptr, "trading_rollout: actions is NULL");
    FRL_REQUIRE(n_steps >= 1, "trading_rollout: n_steps must be >= 1 (got %d)", n_steps);
    FRL_REQUIRE(act_env_stride >= p->stock_dim, "trading_rollout: act_env_stride %lld < stock_dim", (long long)act_env_stride);
    FRL_REQUIRE(obs_mode >= FRL_OBS_NONE && obs_mode <= FRL_OBS_ALL, "trading_rollout: bad obs_mode %d", obs_mode);
    FRL_REQUIRE(obs_mode == FRL_OBS_NONE || obs != nullptr, "trading_rollout: obs is NULL but obs_mode=%d", obs_mode);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = p->stock_dim;
    // Small batches are latency-bound in the thread-per-env kernel; below the measured crossover (~8K envs)
    // the 8-lanes-per-env kernel of trading_small.cu is faster (table in its header).
    // FRL_TRADING_KERNEL=tile|small forces one of them (tests run the whole parity suite under both).
    if (p->stock_dim > 32 && p->n_envs > trading_wide_min_envs()) {
        if (int32_t rc = launch_trading_wide(*p, actions, actions_f64, act_step_stride, act_env_stride, n_steps, rewards, flags,
                                             obs, obs_mode, auto_reset, stats, st))
            return rc;
        return check_launch("trading_rollout(wide)");
    }
    if (p->n_envs <= trading_small_max() || p->stock_dim > 32) {
        launch_trading_small(*p, actions, actions_f64, act_step_stride, act_env_stride, n_steps, rewards, flags, obs,
                             obs_mode, auto_reset, stats, st);
        return check_launch("trading_rollout(small)");
    }
#define FRL_GO(SLOTS, DCT)                                                                                        \
    do {                                                                                                          \
        if (actions_f64)                                                                                          \
            launch_rollout<SLOTS, DCT, double, 2>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards, \
                                                  flags, obs, obs_mode, auto_reset, stats, st);                   \
        else                                                                                                      \
            launch_rollout<SLOTS, DCT, float, 4>(*p, actions, act_step_stride, act_env_stride, n_steps, rewards,  \
                                                 flags, obs, obs_mode, auto_reset, stats, st);                    \
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
    return check_launch("trading_rollout");
}

extern "C" int32_t frl_trading_step(const frl_trading_params *p, const void *actions, int32_t actions_f64,
                                    double *rewards, uint8_t *flags, float *obs, int32_t auto_reset, double *stats,
                                    void *stream)
{
    if (p == nullptr) {
        set_error("trading_step: params is NULL");
        return FRL_E_INVALID;
    }
    return frl_trading_rollout(p, actions, actions_f64, (int64_t)p->n_envs * p->stock_dim, p->stock_dim, 1, rewards,
                               flags, obs, obs ? FRL_OBS_LAST : FRL_OBS_NONE, auto_reset, stats, stream);
}
