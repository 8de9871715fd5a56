// Strand smoothing (Utils/Utils.py:1148-1198, smnooth_strand / smooth_strands; called by HairGrow.py:914 and :950, :975).
//
// Reference, per strand of n points and per axis: least squares of  A x = b  with  A = [lap * L ; pos * I]  (2n x n),
// L = second-difference operator with first-difference end rows (rows (1,-1), (-1,2,-1) ..., (-1,1)),
// b = [0 ; pos * s]; solved through the normal equations  (A^T A) x = A^T b  with scipy's sparse LU in float64, the
// result stored back into the strand's float32 array.  A^T A = lap^2 L^T L + pos^2 I is symmetric positive definite
// and pentadiagonal, A^T b = pos * fl32(pos * s).
//
// Here: one thread per strand builds the three bands of A^T A by accumulating the rows' outer products (so every n >= 2
// follows the reference's matrix exactly), factors them with a banded Cholesky in float64 and solves the three axes;
// results are rounded to float32 like the reference's store.  The two solvers differ by a few float64 ulps (condition
// number ~65 at lap = 4, pos = 2), i.e. the float32 results agree except where a value sits within ~1e-14 of a
// rounding boundary.  Bound: latency (dependent recurrences of length n); work is tiny (O(30 n) flops per strand).
#include "mh_common.cuh"

namespace {

// scratch per point: {d, e, f, x0, x1, x2} doubles at [6 * (offset + k)]
__global__ void __launch_bounds__(128)
smooth_kernel(const float* __restrict__ pts, const int64_t* __restrict__ offsets, const int* __restrict__ lengths,
              int64_t n_strands, double lap, double pos, double* __restrict__ scratch, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_strands) return;
    const int n = lengths[i];
    const int64_t off = offsets[i];
    const float* p = pts + 3 * off;
    float* o = out + 3 * off;
    if (n < 2) {                                         // nothing to smooth (the reference needs n >= 2)
        for (int k = 0; k < 3 * n; ++k) o[k] = p[k];
        return;
    }
    double* w = scratch + 6 * off;
    auto D = [&](int k) -> double& { return w[6 * k]; };      // diagonal            -> Cholesky diagonal
    auto E = [&](int k) -> double& { return w[6 * k + 1]; };  // M[k+1][k]           -> Lc[k+1][k]
    auto F = [&](int k) -> double& { return w[6 * k + 2]; };  // M[k+2][k]           -> Lc[k+2][k]
    auto X = [&](int k, int a) -> double& { return w[6 * k + 3 + a]; };
    const double l2 = lap * lap, p2 = pos * pos;
    for (int k = 0; k < n; ++k) { D(k) = p2; E(k) = 0.0; F(k) = 0.0; }
    // rows of L: (0: +1, 1: -1), (k-1: -1, k: +2, k+1: -1) for k = 1..n-2, (n-2: -1, n-1: +1)
    D(0) += l2; D(1) += l2; E(0) += -l2;
    for (int k = 1; k <= n - 2; ++k) {
        D(k - 1) += l2; D(k) += 4.0 * l2; D(k + 1) += l2;
        E(k - 1) += -2.0 * l2; E(k) += -2.0 * l2; F(k - 1) += l2;
    }
    D(n - 2) += l2; D(n - 1) += l2; E(n - 2) += -l2;
    // banded Cholesky, in place
    for (int j = 0; j < n; ++j) {
        double s = D(j);
        if (j >= 1) s -= E(j - 1) * E(j - 1);
        if (j >= 2) s -= F(j - 2) * F(j - 2);
        const double ljj = sqrt(s);
        D(j) = ljj;
        if (j + 1 < n) {
            double t = E(j);
            if (j >= 1) t -= F(j - 1) * E(j - 1);
            E(j) = t / ljj;
        }
        if (j + 2 < n) F(j) = F(j) / ljj;
    }
    // forward  Lc y = pos^2 s,  backward  Lc^T x = y, three axes together
    for (int j = 0; j < n; ++j) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            // b = strand * pos in float32 (numpy: float32 array times a Python scalar), then A^T b = pos * b in float64
            double y = pos * (double)(p[3 * j + a] * (float)pos);
            if (j >= 1) y -= E(j - 1) * X(j - 1, a);
            if (j >= 2) y -= F(j - 2) * X(j - 2, a);
            X(j, a) = y / D(j);
        }
    }
    for (int j = n - 1; j >= 0; --j) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double x = X(j, a);
            if (j + 1 < n) x -= E(j) * X(j + 1, a);
            if (j + 2 < n) x -= F(j) * X(j + 2, a);
            x = x / D(j);
            X(j, a) = x;
            o[3 * j + a] = (float)x;
        }
    }
}

}  // namespace

extern "C" int64_t mh_smooth_strands_workspace_bytes(int64_t total_points) { return 8 * 6 * (total_points + 1); }

extern "C" int mh_smooth_strands(void* stream, const float* points, const int64_t* offsets, const int32_t* lengths,
                                 int64_t n_strands, double lap_constraint, double pos_constraint, float* points_out,
                                 void* workspace, int64_t workspace_bytes, int64_t total_points) {
    if (n_strands == 0) return 0;
    MH_CHECK_ARG(points && offsets && lengths && points_out && workspace && n_strands > 0 && total_points >= 0, "bad arguments");
    MH_CHECK_ARG(workspace_bytes >= mh_smooth_strands_workspace_bytes(total_points), "workspace too small");
    MH_CHECK_ARG(pos_constraint != 0.0, "pos_constraint must be non-zero (singular system)");
    smooth_kernel<<<(unsigned)((n_strands + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        points, offsets, lengths, n_strands, lap_constraint, pos_constraint, reinterpret_cast<double*>(workspace), points_out);
    MH_COUNT_LAUNCH();
    MH_CHECK_LAUNCH();
    return 0;
}
