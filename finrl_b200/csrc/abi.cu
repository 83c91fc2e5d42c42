// finrl_b200 — C-ABI plumbing shared by every env kind: version + thread-local error string.
#include <stdarg.h>
#include <stddef.h>
#include <string.h>

#include "common.cuh"

namespace frl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t check_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return FRL_E_CUDA;
    }
    return FRL_OK;
}

}  // namespace frl

extern "C" int32_t frl_abi_version(void) { return FRL_ABI_VERSION; }
extern "C" const char *frl_last_error(void) { return frl::g_err; }

// ---- peer-mapped statistics blocks (frl_stats_block) -------------------------------------------------
static_assert(sizeof(frl_stats_block) == FRL_STATS_BLOCK_BYTES, "frl_stats_block must be 384 bytes");
static_assert(offsetof(frl_stats_block, total) == 128 && offsetof(frl_stats_block, n_peers) == 256, "frl_stats_block layout");
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");

#define FRL_CUDA(call, what)                                                                        \
    do {                                                                                            \
        const cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                                    \
            (void)cudaGetLastError();                                                               \
            frl::set_error("%s: CUDA error %d (%s)", what, (int)e_, cudaGetErrorString(e_));       \
            return FRL_E_CUDA;                                                                      \
        }                                                                                           \
    } while (0)

extern "C" int32_t frl_exchange_alloc(void **block)
{
    FRL_REQUIRE(block != nullptr, "exchange_alloc: block is NULL");
    void *p = nullptr;
    FRL_CUDA(cudaMalloc(&p, FRL_STATS_BLOCK_BYTES), "exchange_alloc");
    FRL_CUDA(cudaMemset(p, 0, FRL_STATS_BLOCK_BYTES), "exchange_alloc(memset)");
    *block = p;
    return FRL_OK;
}

extern "C" int32_t frl_exchange_free(void *block)
{
    if (block) FRL_CUDA(cudaFree(block), "exchange_free");
    return FRL_OK;
}

extern "C" int32_t frl_exchange_export(const void *block, uint8_t handle[64])
{
    FRL_REQUIRE(block != nullptr && handle != nullptr, "exchange_export: NULL argument");
    cudaIpcMemHandle_t h;
    FRL_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(block)), "exchange_export");
    memcpy(handle, &h, sizeof(h));
    return FRL_OK;
}

extern "C" int32_t frl_exchange_open(const uint8_t handle[64], void **peer_block)
{
    FRL_REQUIRE(handle != nullptr && peer_block != nullptr, "exchange_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    FRL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "exchange_open");
    *peer_block = p;
    return FRL_OK;
}

extern "C" int32_t frl_exchange_close(void *peer_block)
{
    if (peer_block) FRL_CUDA(cudaIpcCloseMemHandle(peer_block), "exchange_close");
    return FRL_OK;
}

extern "C" int32_t frl_exchange_bind(void *block, void *const *blocks, int32_t n, void *stream)
{
    FRL_REQUIRE(block != nullptr, "exchange_bind: block is NULL");
    FRL_REQUIRE(n >= 0 && n <= FRL_MAX_PEERS, "exchange_bind: n must be in 0..%d (got %d)", FRL_MAX_PEERS, n);
    FRL_REQUIRE(n == 0 || blocks != nullptr, "exchange_bind: blocks is NULL");
    FRL_REQUIRE((reinterpret_cast<uintptr_t>(block) & 127) == 0, "exchange_bind: block must be 128-byte aligned");
    struct Tail {
        double *peer_total[FRL_MAX_PEERS];
        uint32_t n_peers, reserved;
    } t = {};
    for (int i = 0; i < n; ++i) {
        FRL_REQUIRE(blocks[i] != nullptr, "exchange_bind: blocks[%d] is NULL", i);
        t.peer_total[i] = reinterpret_cast<frl_stats_block *>(blocks[i])->total;  // address arithmetic only
    }
    t.n_peers = (uint32_t)n;
    // pageable source: the runtime stages it before returning, so the stack copy may go out of scope
    FRL_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(block) + offsetof(frl_stats_block, peer_total), &t, sizeof(t),
                             cudaMemcpyHostToDevice, (cudaStream_t)stream),
             "exchange_bind");
    return FRL_OK;
}

namespace {
__global__ void exchange_flush_kernel(frl_stats_block *b)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        frl::stats_push_accumulator(b, 0);
        frl::stats_push_accumulator(b, 1);
    }
}
}  // namespace

extern "C" int32_t frl_exchange_flush(void *block, void *stream)
{
    FRL_REQUIRE(block != nullptr, "exchange_flush: block is NULL");
    exchange_flush_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<frl_stats_block *>(block));
    return frl::check_launch("exchange_flush");
}
