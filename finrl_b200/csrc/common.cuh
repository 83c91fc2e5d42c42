// finrl_b200 — shared device/host helpers for the sm_100a env-step kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "finrl_b200.h"

namespace frl {

// ---- error plumbing (host) -------------------------------------------------------------------
void set_error(const char *fmt, ...);
int32_t check_launch(const char *what);

#define FRL_REQUIRE(cond, ...)                                                                     \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ::frl::set_error(__VA_ARGS__);                                                         \
            return FRL_E_INVALID;                                                                  \
        }                                                                                          \
    } while (0)

constexpr int kWarp = 32;

// ---- exact IEEE helpers (device) ---------------------------------------------------------------
// The reference's arithmetic is one IEEE rounding per Python/numpy operation.  Every product and
// sum on the bit-exact list goes through these so that ptxas can never contract a*b+c into an FMA
// (SURVEY.md H2).
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// numpy's float floor-division (npy_divmod) returns, for every finite a and b > 0 with a moderate
// quotient, the exact mathematical floor(a/b) of the two floating-point values (its fmod is exact
// and its final snap repairs the one rounding of (a-mod)/b).  We get the same integer without
// fmod's long loop: round-to-nearest quotient, floor, then one exact FMA remainder to repair the
// (at most one-off) error.  b == 0 gives +-inf like numpy (a != 0).
__device__ __forceinline__ double floor_div_f64(double a, double b)
{
    double q = floor(__ddiv_rn(a, b));
    double r = __fma_rn(-q, b, a);  // exact: |r| < 2b and r is a multiple of ulp(b)
    if (r < 0.0)
        q -= 1.0;
    else if (r >= b)
        q += 1.0;
    return q;
}

__device__ __forceinline__ float floor_div_f32(float a, float b)
{
    float q = floorf(__fdiv_rn(a, b));
    float r = __fmaf_rn(-q, b, a);
    if (r < 0.0f)
        q -= 1.0f;
    else if (r >= b)
        q += 1.0f;
    return q;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- counter-based random draws for the in-kernel auto-reset -------------------------------------------------
// splitmix64 of a counter built from (launch stream id, env index, rollout step, draw index): no generator
// state to store or race on, every (env, step) has its own stream, and a rarely taken branch needs no registers
// outside it.  The reference draws from numpy's / Python's GLOBAL generators, whose order over a batch is
// undefined; what is reproduced is the distribution (randint / uniform / random.choice).
__device__ __forceinline__ uint64_t reset_bits(uint64_t seed, long long env, int step, int idx)
{
    uint64_t z = seed + 0x9e3779b97f4a7c15ull * (uint64_t)(env + 1) + 0xd1342543de82ef95ull * (uint64_t)(step + 1) +
                 0xaf251af3b0f025b5ull * (uint64_t)(idx + 1);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// uniform double in [0, 1) with 53 random bits, like numpy's random_sample
__device__ __forceinline__ double reset_uniform01(uint64_t bits) { return (double)(bits >> 11) * 0x1.0p-53; }
// uniform integer in [0, n), n < 2^31 (multiply-shift on 32 fresh bits; bias < n / 2^32)
__device__ __forceinline__ int reset_randint(uint64_t bits, int n) { return (int)(((bits >> 32) * (uint64_t)n) >> 32); }

// ---- the 8-slot statistics vector -----------------------------------------------------------------
// 8 values x 32 lanes: three halving exchanges leave lane l with the partial sum of value (l & 7) over 4
// lanes, two more butterflies finish it — 9 shuffles instead of 40 — then one atomicAdd per slot and warp.
__device__ __forceinline__ void reduce_stats8(double (&v)[FRL_N_STATS], int lane, double *__restrict__ stats)
{
#pragma unroll
    for (int w = 4; w >= 1; w >>= 1) {
        const bool up = (lane & w) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const double keep = up ? v[i + w] : v[i];
            const double send = up ? v[i] : v[i + w];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    double s = v[0];
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (lane < 8 && s != 0.0) atomicAdd(stats + lane, s);
}

// One-sided multi-GPU exchange fused into the step kernels (frl_stats_block in the header).  `stats` is one of
// the block's two accumulators; the caller alternates them launch by launch, so the OTHER one holds the previous
// launch's sums, complete by stream order.  The first thread of the grid adds those to total[] of every rank over
// NVLink (fp64 atomics on peer-mapped memory, fire and forget) and clears them.  No counter, no fence, no
// rendezvous: measured cost on the 1M-env StockTradingEnv step: none (a completion ticket per warp or per block,
// the textbook last-block pattern, cost 4-8 % there — the kernel is latency-bound and every fence + atomic round
// trip sits on a warp's critical path; tools/ab_stats_tail.sh, profiles/r02_ab_stats_exchange.txt).
static __device__ __noinline__ void stats_push_accumulator(frl_stats_block *b, int which)
{
    const unsigned n_peers = b->n_peers;
    if (n_peers == 0) return;
#pragma unroll 1
    for (int s = 0; s < FRL_N_STATS; ++s) {
        const double v = __longlong_as_double(
            (long long)atomicExch(reinterpret_cast<unsigned long long *>(&b->sum[which][s]), 0ull));
        if (v != 0.0) {
#pragma unroll 1
            for (unsigned r = 0; r < n_peers; ++r)  // RED.E.ADD.F64.RN.STRONG.SYS on peer-mapped global memory
                asm volatile("red.global.sys.add.f64 [%0], %1;" ::"l"(b->peer_total[r] + s), "d"(v) : "memory");
        }
    }
}

// Called by every thread at the top of a step / rollout kernel (before any early return).
__device__ __forceinline__ void stats_exchange_previous(double *__restrict__ stats)
{
    if (stats == nullptr || blockIdx.x != 0 || threadIdx.x != 0) return;
    const uintptr_t a = reinterpret_cast<uintptr_t>(stats);
    const int which = (int)((a >> 6) & 1);  // blocks are 128-byte aligned: sum[0] at +0, sum[1] at +64
    stats_push_accumulator(reinterpret_cast<frl_stats_block *>(a - 64 * which), which ^ 1);
}

// ---- action staging -------------------------------------------------------------------------------
// Stage one step's actions of a 32-env tile into shared memory, flat [32 envs][D] exactly as they lie in
// global memory.  With the default layout (act_env_stride == D) the tile is one contiguous run: D fully
// coalesced, independent loads per lane are issued back to back, then parked.  Rows beyond the valid envs
// are zero-filled.  Callers bracket this with __syncwarp().
template <int SLOTS, typename ActT>
__device__ __forceinline__ void stage_actions_flat(ActT *__restrict__ dst, const ActT *__restrict__ abase, long long env0,
                                                   long long act_env_stride, int D, int nvalid, int lane)
{
    if (act_env_stride == D) {
        const ActT *tile = abase + (size_t)env0 * D + lane;
        const int cnt = nvalid * D - lane;
        ActT av[SLOTS];
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) av[i] = (i < D && 32 * i < cnt) ? __ldcs(tile + 32 * i) : ActT(0);
#pragma unroll
        for (int i = 0; i < SLOTS; ++i)
            if (i < D) dst[lane + 32 * i] = av[i];
    } else {
        for (int r = 0; r < 32; ++r)
            if (lane < D) dst[r * D + lane] = r < nvalid ? abase[(size_t)(env0 + r) * act_env_stride + lane] : ActT(0);
    }
}

// The two halves of the contiguous-layout fast path above, for kernels that want the action loads in flight
// together with their state loads (one DRAM round trip instead of two).
template <int SLOTS, typename ActT>
__device__ __forceinline__ void stage_actions_load(ActT (&av)[SLOTS], const ActT *__restrict__ abase, long long env0, int D,
                                                   int nvalid, int lane)
{
    const ActT *tile = abase + (size_t)env0 * D + lane;
    const int cnt = nvalid * D - lane;
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) av[i] = (i < D && 32 * i < cnt) ? __ldcs(tile + 32 * i) : ActT(0);
}
template <int SLOTS, typename ActT>
__device__ __forceinline__ void stage_actions_store(ActT *__restrict__ dst, const ActT (&av)[SLOTS], int D, int lane)
{
#pragma unroll
    for (int i = 0; i < SLOTS; ++i)
        if (i < D) dst[lane + 32 * i] = av[i];
}

// Observation rows are written once and never read back by the kernel: the store's cache operator decides whether
// they pass through (and wash out) the little L1 that is left beside the shared-memory carve-out.  Measured per
// kernel (profiles/r02_ab_obs_writer.txt): st.global.cg (L2 only) for the cash-penalty / stop-loss and numpy-env
// writers, st.global.cs (streaming) for the two wide kernels.
enum ObsStore { kStorePlain = 0, kStoreCG = 1, kStoreCS = 2 };
template <int MODE>
__device__ __forceinline__ void obs_store(float *p, float v)
{
    if (MODE == kStoreCG)
        __stcg(p, v);
    else if (MODE == kStoreCS)
        __stcs(p, v);
    else
        *p = v;
}
// ---- bulk-copy engine (TMA, non-tensor form) ------------------------------------------------------
// cp.async.bulk moves a contiguous, 16-byte-aligned run between global and shared memory without touching
// the LSU/L1 path.  Loads complete on an mbarrier (expect_tx bytes), stores are tracked in bulk groups.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(mbar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_store_s2g(void *dst, unsigned src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}

}  // namespace frl
