// Batched pseudo-inverse + rank of small symmetric positive semi-definite matrices with the
// reference's rule (inv_rank, src/normalisr/association.py:66-80: singular values below
// tol * largest are dropped).  normvar needs one per gene (norm.py:159-160) and de(single=1) one
// per grouping (association.py:348-350); the reference calls scipy.linalg.svd in a Python loop.
//
// One warp per matrix, cyclic Jacobi in shared memory (for symmetric PSD matrices the singular
// values are the eigenvalues; Jacobi computes them to high relative accuracy).
#include "nsr_common.cuh"

namespace {

constexpr int kPMax = 16;              // largest matrix order
constexpr int kPWarps = 4;             // matrices per CTA
constexpr int kPSweeps = 30;

__global__ void __launch_bounds__(32 * kPWarps)
sym_pinv_kernel(const double* __restrict__ G, int64_t batch, int n, double tol, double* __restrict__ out,
                int32_t* __restrict__ rank_out) {
    __shared__ double s_a[kPWarps][kPMax][kPMax + 1];
    __shared__ double s_v[kPWarps][kPMax][kPMax + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * kPWarps + warp;
    if (m >= batch) return;                               // whole warps leave together; no CTA-wide sync below
    double (*a)[kPMax + 1] = s_a[warp];
    double (*v)[kPMax + 1] = s_v[warp];
    const double* g = G + m * n * n;
    for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx / n, j = idx % n;
        a[i][j] = 0.5 * (g[i * n + j] + g[j * n + i]);    // symmetrise
        v[i][j] = i == j ? 1.0 : 0.0;
    }
    __syncwarp();
    for (int sweep = 0; sweep < kPSweeps; ++sweep) {
        int rotations = 0;                                // identical on every lane (shared-memory values)
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p][q], app = a[p][p], aqq = a[q][q];
                // classical Jacobi stopping rule: |a_pq| negligible against sqrt(a_pp a_qq)
                if (apq != 0.0 && fabs(apq) > 1e-17 * sqrt(fabs(app * aqq))) {
                    ++rotations;
                    const double theta = (aqq - app) / (2.0 * apq);
                    const double t = fabs(theta) > 1e150 ? 0.5 / theta
                                                         : (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                    const double c = 1.0 / sqrt(fma(t, t, 1.0)), s = t * c;
                    __syncwarp();
                    if (lane < n) {                       // columns p, q of A and V
                        const double akp = a[lane][p], akq = a[lane][q];
                        a[lane][p] = c * akp - s * akq;
                        a[lane][q] = s * akp + c * akq;
                        const double vkp = v[lane][p], vkq = v[lane][q];
                        v[lane][p] = c * vkp - s * vkq;
                        v[lane][q] = s * vkp + c * vkq;
                    }
                    __syncwarp();
                    if (lane < n) {                       // rows p, q of A
                        const double apk = a[p][lane], aqk = a[q][lane];
                        a[p][lane] = c * apk - s * aqk;
                        a[q][lane] = s * apk + c * aqk;
                    }
                    __syncwarp();
                }
            }
        }
        if (rotations == 0) break;
    }
    __syncwarp();
    double lmax = 0.0;
    for (int i = 0; i < n; ++i) lmax = fmax(lmax, a[i][i]);
    int rank = 0;
    for (int i = 0; i < n; ++i) rank += (a[i][i] >= tol * lmax && a[i][i] > 0.0) ? 1 : 0;
    double* o = out + m * n * n;
    for (int idx = lane; idx < n * n; idx += 32) {
        const int r = idx / n, cidx = idx % n;
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            const double lam = a[i][i];
            if (lam >= tol * lmax && lam > 0.0) acc = fma(v[r][i] / lam, v[cidx][i], acc);
        }
        o[idx] = acc;
    }
    if (lane == 0) rank_out[m] = rank;
}

}  // namespace

extern "C" int nsr_sym_pinv(nsr_ctx* ctx, uintptr_t stream, const double* G, int64_t batch, int n, double tol,
                            double* pinv, int32_t* rank) {
    NSR_REQUIRE(ctx && G && pinv && rank, "nsr_sym_pinv: null argument");
    NSR_REQUIRE(batch >= 1 && n >= 1 && n <= kPMax && tol > 0.0, "nsr_sym_pinv: bad shape batch=%lld n=%d (n <= %d)",
                (long long)batch, n, kPMax);
    NSR_CHECK(cudaSetDevice(ctx->device));
    sym_pinv_kernel<<<(unsigned)((batch + kPWarps - 1) / kPWarps), 32 * kPWarps, 0, (cudaStream_t)stream>>>(
        G, batch, n, tol, pinv, rank);
    NSR_CHECK(cudaGetLastError());
    return 0;
}
