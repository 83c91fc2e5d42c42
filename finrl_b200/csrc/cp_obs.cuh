// Observation writer shared by the cash-penalty and stop-loss kernels (same layout: [cash, holdings x D,
// daily information]; frl_cashpenalty_params and frl_stoploss_params both carry obs_dim, stock_dim, obs_tmpl).
#pragma once

#include "common.cuh"

namespace frl {

// Warp-cooperative write of the tile's observation rows: [coh, holdings x D, daily information].
// Env r's staging row holds its float32 holdings after pass 2.  NCH = ceil(O/32) is compiled in so the
// row loop is NCH plain stores; only the first (D/32)+1 chunks need the cash / holdings patch-up.
template <int NCH, typename ActT, typename Params>
__device__ __forceinline__ void cp_write_obs_rows_uniform(const Params &p, const ActT *stage, int P,
                                                          const float *cashf, float *__restrict__ obs, long long env0,
                                                          int nvalid, int lane, int d0, int D)
{
    // (D is a compile-time constant in the NASDAQ-100 instantiation of the cash-penalty kernel: the per-chunk facts of
    // the chunks that lie wholly inside or outside the holdings then fold away)
    const int O = p.obs_dim;
    constexpr int step = sizeof(ActT) / sizeof(float);  // the float image sits in the low word of each slot
    constexpr int NSP = NCH < 5 ? NCH : 5;              // D <= 128: holdings end inside chunk 4
    float t[NCH];
    bool img[NSP];
    const float *trow = p.obs_tmpl + (size_t)d0 * O + lane;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int pos = lane + 32 * c;
        const bool im = c < NSP && pos >= 1 && pos <= D;
        if (c < NSP) img[c] = im;
        t[c] = (!im && (c < NCH - 1 || pos < O)) ? __ldg(trow + 32 * c) : 0.0f;
    }
    const bool tail_ok = lane + 32 * (NCH - 1) < O;
    const float *hrow = reinterpret_cast<const float *>(stage) + (lane - 1) * step;  // holding of stock (pos - 1)
    float *orow = obs + (size_t)env0 * O + lane;
    const int pitch = P * step;
#pragma unroll 2
    for (int r = 0; r < nvalid; ++r) {
        float v[NSP];
#pragma unroll
        for (int c = 0; c < NSP; ++c) v[c] = img[c] ? hrow[32 * c * step] : t[c];
        const float cf = cashf[r];
        if (lane == 0) v[0] = cf;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float x = c < NSP ? v[c] : t[c];
            if (c < NCH - 1 || tail_ok) obs_store<kStoreCG>(orow + 32 * c, x);
        }
        orow += O;
        hrow += pitch;
    }
}

template <typename ActT, typename Params>
__device__ __forceinline__ void cp_write_obs_tile(const Params &p, const ActT *stage, int P,
                                                  const float *cashf, const int *di_s, float *__restrict__ obs,
                                                  long long env0, int nvalid, int lane, int D)
{
    const int O = p.obs_dim;
    const int d0 = di_s[0];
    bool uniform = true;
    if (lane < nvalid) uniform = (di_s[lane] == d0);
    uniform = __all_sync(0xffffffffu, uniform);
    const int nch = (O + 31) >> 5;
    if (uniform && nch <= 24) {
        switch (nch) {
#define FRL_CASE(N)                                                                                \
    case N:                                                                                        \
        cp_write_obs_rows_uniform<N, ActT, Params>(p, stage, P, cashf, obs, env0, nvalid, lane, d0, D);    \
        break;
            FRL_CASE(1) FRL_CASE(2) FRL_CASE(3) FRL_CASE(4) FRL_CASE(5) FRL_CASE(6) FRL_CASE(7) FRL_CASE(8)
            FRL_CASE(9) FRL_CASE(10) FRL_CASE(11) FRL_CASE(12) FRL_CASE(13) FRL_CASE(14) FRL_CASE(15) FRL_CASE(16)
            FRL_CASE(17) FRL_CASE(18) FRL_CASE(19) FRL_CASE(20) FRL_CASE(21) FRL_CASE(22) FRL_CASE(23) FRL_CASE(24)
#undef FRL_CASE
        }
    } else {
        for (int r = 0; r < nvalid; ++r) {
            const float *hrow = reinterpret_cast<const float *>(stage + (size_t)r * P);
            constexpr int step = sizeof(ActT) / sizeof(float);
            const float *trow = p.obs_tmpl + (size_t)di_s[r] * O;
            float *orow = obs + (size_t)(env0 + r) * O;
            for (int pos = lane; pos < O; pos += 32) {
                float v;
                if (pos == 0)
                    v = cashf[r];
                else if (pos <= D)
                    v = hrow[(pos - 1) * step];
                else
                    v = __ldg(trow + pos);
                orow[pos] = v;
            }
        }
    }
}

}  // namespace frl
