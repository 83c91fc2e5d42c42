// Table precompute on the GPU: the rolling 252-day covariance that is the portfolio
// env's observation and the history statistics of the turbulence index.  The reference does both with
// per-day pandas pivots (tutorial :157-174, preprocessors.py:215-267), the slowest step feeding the envs.
//
// One block per window.  The window's rows (n_rows x D doubles, 60 KB for 252 x 30) are L2-resident
// (the whole return table is < 1 MB), so each thread owns a few (a, b) pairs and streams the rows:
// two passes (mean, then centred cross products) in fp64.
#include "common.cuh"

namespace frl {
namespace {

__global__ void __launch_bounds__(256)
rolling_cov_kernel(const double *__restrict__ ret, int D, int first_row, int n_rows, double *__restrict__ cov_out,
                   double *__restrict__ mean_out)
{
    extern __shared__ double mean_s[];  // [D]
    const int w = blockIdx.x;
    const double *rows = ret + (size_t)(first_row + w) * D;
    for (int a = threadIdx.x; a < D; a += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < n_rows; ++t) s += rows[(size_t)t * D + a];
        const double m = s / (double)n_rows;
        mean_s[a] = m;
        if (mean_out) mean_out[(size_t)w * D + a] = m;
    }
    __syncthreads();
    const double inv = 1.0 / (double)(n_rows - 1);
    for (int pair = threadIdx.x; pair < D * D; pair += blockDim.x) {
        const int a = pair / D, b = pair - a * D;
        if (b < a) continue;  // symmetric: compute the upper triangle, mirror it
        const double ma = mean_s[a], mb = mean_s[b];
        double acc = 0.0;
        for (int t = 0; t < n_rows; ++t) acc += (rows[(size_t)t * D + a] - ma) * (rows[(size_t)t * D + b] - mb);
        const double c = acc * inv;
        cov_out[(size_t)w * D * D + (size_t)a * D + b] = c;
        cov_out[(size_t)w * D * D + (size_t)b * D + a] = c;
    }
}

}  // namespace
}  // namespace frl

using namespace frl;

extern "C" int32_t frl_rolling_cov(const double *ret, int32_t n_days, int32_t stock_dim, int32_t first_row, int32_t n_rows,
                                   int32_t n_out, double *cov_out, double *mean_out, void *stream)
{
    FRL_REQUIRE(ret != nullptr && cov_out != nullptr, "rolling_cov: NULL pointer");
    FRL_REQUIRE(stock_dim >= 1 && stock_dim <= 4096, "rolling_cov: stock_dim must be in 1..4096 (got %d)", stock_dim);
    FRL_REQUIRE(n_rows >= 2 && n_out >= 1 && first_row >= 0, "rolling_cov: bad window (first_row %d, n_rows %d, n_out %d)",
                first_row, n_rows, n_out);
    FRL_REQUIRE((long long)first_row + n_out - 1 + n_rows <= n_days, "rolling_cov: last window ends at row %lld > n_days %d",
                (long long)first_row + n_out - 1 + n_rows, n_days);
    rolling_cov_kernel<<<(unsigned)n_out, 256, stock_dim * sizeof(double), (cudaStream_t)stream>>>(ret, stock_dim, first_row,
                                                                                                 n_rows, cov_out, mean_out);
    return check_launch("rolling_cov");
}
