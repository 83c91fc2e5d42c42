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

// ---- turbulence index (preprocessors.py:215-267) -------------------------------------------------------------
// temp_i = x^T pinv(C_i) x with x = (returns of day i) - (mean of the window) and C_i the window's covariance.
// np.linalg.pinv is an SVD with the singular values <= rcond * max cut off; for the symmetric C that equals the
// eigen-decomposition C = V diag(lambda) V^T with |lambda| <= rcond * max|lambda| dropped:
//     temp = sum_k (v_k^T x)^2 / lambda_k.
// One thread block per day runs a cyclic two-sided Jacobi eigen-solve of C in shared memory (fp64; n/2 disjoint
// rotations per step in the round-robin order, n-1 steps per sweep) and applies every rotation to x as well, so V
// is never formed: at convergence the diagonal holds lambda and x holds V^T x.
constexpr int kTurbThreads = 256;
constexpr int kTurbMaxSweeps = 40;

__global__ void __launch_bounds__(kTurbThreads)
turbulence_kernel(const double *__restrict__ ret, int D, int start, const double *__restrict__ cov,
                  const double *__restrict__ mean, double rcond, double *__restrict__ temp_out)
{
    extern __shared__ double tsm[];
    const int n = (D + 1) & ~1;  // even size: an odd D gets one zero row / column (eigenvalue 0, cut off)
    const int ld = n + 1;
    double *A = tsm;                        // [n][ld]
    double *x = A + (size_t)n * ld;         // [n]
    double *cs = x + n;                     // [n/2][2]
    int *pq = reinterpret_cast<int *>(cs + n);  // [n/2][2]
    __shared__ double red[kTurbThreads / 32];
    __shared__ double off_s, diag_s;
    const int w = blockIdx.x, tid = threadIdx.x;
    const double *C = cov + (size_t)w * D * D;
    for (int i = tid; i < n * n; i += kTurbThreads) {
        const int r = i / n, c = i - r * n;
        A[r * ld + c] = (r < D && c < D) ? C[(size_t)r * D + c] : 0.0;
    }
    for (int i = tid; i < n; i += kTurbThreads)
        x[i] = i < D ? ret[(size_t)(start + w) * D + i] - mean[(size_t)w * D + i] : 0.0;
    __syncthreads();

    const int half = n / 2;
    for (int sweep = 0; sweep < kTurbMaxSweeps; ++sweep) {
        // convergence: off-diagonal mass against the diagonal's
        double off = 0.0, dg = 0.0;
        for (int i = tid; i < n * n; i += kTurbThreads) {
            const int r = i / n, c = i - r * n;
            const double v = A[r * ld + c];
            if (r == c)
                dg += v * v;
            else
                off += v * v;
        }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if ((tid & 31) == 0) red[tid >> 5] = off;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < kTurbThreads / 32; ++i) t += red[i];
            off_s = t;
        }
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = dg;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < kTurbThreads / 32; ++i) t += red[i];
            diag_s = t;
        }
        __syncthreads();
        if (off_s <= 1e-30 * diag_s || off_s == 0.0) break;  // uniform over the block

        for (int step = 0; step < n - 1; ++step) {
            // round-robin pairing of the n indices: n-1 stays, the others rotate
            if (tid < half) {
                int p, q;
                if (tid == 0) {
                    p = n - 1;
                    q = step;
                } else {
                    p = (step + tid) % (n - 1);
                    q = (step - tid + (n - 1)) % (n - 1);
                }
                if (p > q) {
                    const int t = p;
                    p = q;
                    q = t;
                }
                const double apq = A[p * ld + q];
                double c = 1.0, sn = 0.0;
                if (apq != 0.0) {
                    const double theta = (A[q * ld + q] - A[p * ld + p]) / (2.0 * apq);
                    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    c = 1.0 / sqrt(t * t + 1.0);
                    sn = t * c;
                }
                cs[2 * tid] = c;
                cs[2 * tid + 1] = sn;
                pq[2 * tid] = p;
                pq[2 * tid + 1] = q;
            }
            __syncthreads();
            // A <- A J : columns p, q of every row
            for (int i = tid; i < half * n; i += kTurbThreads) {
                const int k = i / n, r = i - k * n;
                const double c = cs[2 * k], sn = cs[2 * k + 1];
                const int p = pq[2 * k], q = pq[2 * k + 1];
                const double ap = A[r * ld + p], aq = A[r * ld + q];
                A[r * ld + p] = c * ap - sn * aq;
                A[r * ld + q] = sn * ap + c * aq;
            }
            __syncthreads();
            // A <- J^T A : rows p, q of every column; x <- J^T x
            for (int i = tid; i < half * (n + 1); i += kTurbThreads) {
                const int k = i / (n + 1), j = i - k * (n + 1);
                const double c = cs[2 * k], sn = cs[2 * k + 1];
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (j < n) {
                    const double ap = A[p * ld + j], aq = A[q * ld + j];
                    A[p * ld + j] = c * ap - sn * aq;
                    A[q * ld + j] = sn * ap + c * aq;
                } else {
                    const double xp = x[p], xq = x[q];
                    x[p] = c * xp - sn * xq;
                    x[q] = sn * xp + c * xq;
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        double lmax = 0.0;
        for (int k = 0; k < n; ++k) lmax = fmax(lmax, fabs(A[k * ld + k]));
        const double cut = rcond * lmax;
        double t = 0.0;
        for (int k = 0; k < n; ++k) {
            const double l = A[k * ld + k];
            if (fabs(l) > cut) t += x[k] * x[k] / l;
        }
        temp_out[w] = t;
    }
}

// the reference's bookkeeping of the raw values (:250-262): zero unless positive, and the first two positive
// values are suppressed ("avoid large outlier because of the calculation just begins")
__global__ void turbulence_finish_kernel(const double *__restrict__ temp, int n, int start, double *__restrict__ out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int i = 0; i < start; ++i) out[i] = 0.0;
    int count = 0;
    for (int k = 0; k < n; ++k) {
        const double t = temp[k];
        double v = 0.0;
        if (t > 0.0) {
            count += 1;
            if (count > 2) v = t;
        }
        out[start + k] = v;
    }
}

}  // namespace
}  // namespace frl

using namespace frl;

extern "C" int32_t frl_turbulence(const double *ret, int32_t n_days, int32_t stock_dim, int32_t start, const double *cov,
                                  const double *mean, double rcond, double *scratch, double *out, void *stream)
{
    FRL_REQUIRE(ret && cov && mean && scratch && out, "turbulence: NULL pointer");
    FRL_REQUIRE(stock_dim >= 1 && stock_dim <= 160, "turbulence: stock_dim must be in 1..160 (got %d)", stock_dim);
    FRL_REQUIRE(start >= 1 && start < n_days, "turbulence: start %d outside [1, n_days = %d)", start, n_days);
    const int n = (stock_dim + 1) & ~1, n_out = n_days - start;
    const size_t smem = ((size_t)n * (n + 1) + n + n) * sizeof(double) + (size_t)n * sizeof(int);
    if (smem > 48 * 1024)
        if (cudaFuncSetAttribute(turbulence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch("turbulence(smem)");
    turbulence_kernel<<<(unsigned)n_out, kTurbThreads, smem, (cudaStream_t)stream>>>(ret, stock_dim, start, cov, mean, rcond, scratch);
    if (int32_t rc = check_launch("turbulence")) return rc;
    turbulence_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scratch, n_out, start, out);
    return check_launch("turbulence(finish)");
}

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
