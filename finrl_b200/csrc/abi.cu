// finrl_b200 — C-ABI plumbing shared by every env kind: version + thread-local error string.
#include <stdarg.h>

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
