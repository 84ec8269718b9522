// newton_batched.cu -- C ABI "Level 1.5": the NLSolver level of the drop-in boundary (include/b200ida.h).
//
// `Newton<T, D>::solve` of the reference (crates/nonlinear/src/newton.rs:51-167) over a batch of independent nonlinear
// problems, one CUDA thread per problem: residual -> [Jacobian + LU] -> { negate, lsolve, y += delta, ctest, residual }*
// with one retry on a fresh Jacobian when the iteration fails to converge on a stale one.  The NLProblem callbacks
// (crates/nonlinear/src/traits.rs:5-127: sys / setup / solve / ctest) cannot be host closures inside a kernel, so -- as
// for the DAE problems of Level 2 -- they are device functors selected by a problem kind with per-instance data:
//   B200IDA_NLP_QUAD3   the reference's own known-answer problem (newton.rs:192-343):
//                       f1 = x^2+y^2+z^2-1, f2 = 2x^2+y^2-4z, f3 = 3x^2-4y+z^2, analytic J, Dense<U3> LU,
//                       ctest = (wrms(del, w) <= tol)
// A new kind is one more functor with sys / jac (see INTEGRATION.md "Adding a problem").  Arithmetic: un-fused IEEE in
// the reference's order (-fmad=false), the LU / triangular solves are dense_get_rf / dense_get_rs (dense.rs:86-206)
// and the norm is NormWRMS (norm_wrms.rs:24-33), so results are bit-identical to the oracle's Newton.
// Control-flow deviation shared with the oracle: "ConvergenceRecover while the Jacobian is current"
// returns the recoverable error (SUNDIALS semantics) where the reference's loop would not terminate (newton.rs:146-153).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200ida.h"
#include "host_util.h"

namespace b200ida {

// dense_get_rf on a register-resident N x N matrix (crates/linear/src/dense.rs:86-158); returns 0 or the 1-based column
template <int N>
__device__ __forceinline__ int getrf_reg(double (&A)[N][N], int (&piv)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        int l = k;
        double best = fabs(A[k][k]);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double v = fabs(A[i][k]);
            if (v > best) {
                best = v;
                l = i;
            }
        }
        piv[k] = l;
        if (best == 0.0) return k + 1;
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            if (l == i) {
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    const double t = A[k][c];
                    A[k][c] = A[i][c];
                    A[i][c] = t;
                }
            }
        }
        const double mult = 1.0 / A[k][k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) A[i][k] *= mult;
#pragma unroll
        for (int j = k + 1; j < N; ++j) {
            const double a_kj = A[k][j];
            if (a_kj != 0.0) {
#pragma unroll
                for (int i = k + 1; i < N; ++i) A[i][j] -= a_kj * A[i][k];
            }
        }
    }
    return 0;
}

// dense_get_rs (dense.rs:165-206)
template <int N>
__device__ __forceinline__ void getrs_reg(const double (&A)[N][N], const int (&piv)[N], double (&b)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            if (piv[k] == i) {
                const double t = b[k];
                b[k] = b[i];
                b[i] = t;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
        const double bk = b[k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) b[i] -= A[i][k] * bk;
    }
#pragma unroll
    for (int k = N - 1; k >= 1; --k) {
        b[k] = b[k] / A[k][k];
        const double bk = b[k];
#pragma unroll
        for (int i = 0; i < k; ++i) b[i] -= A[i][k] * bk;
    }
    b[0] = b[0] / A[0][0];
}

// NormWRMS::norm_wrms (crates/nonlinear/src/norm_wrms.rs:24-33): ordered fold from 0
template <int N>
__device__ __forceinline__ double wrms_reg(const double (&x)[N], const double (&w)[N]) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double p = x[i] * w[i];
        acc = acc + p * p;
    }
    return sqrt(acc / (double)N);
}

// the reference's TestProblem (newton.rs:192-343)
struct NlpQuad3 {
    static constexpr int N = 3;
    __device__ __forceinline__ void sys(const double (&v)[3], double (&f)[3]) const {
        const double x = v[0], y = v[1], z = v[2];
        f[0] = x * x + y * y + z * z - 1.0;
        f[1] = 2.0 * (x * x) + y * y - 4.0 * z;
        f[2] = 3.0 * (x * x) - 4.0 * y + z * z;
    }
    // J[i][j] = d f_i / d v_j (newton.rs:205-222 fills it column by column)
    __device__ __forceinline__ void jac(const double (&v)[3], double (&J)[3][3]) const {
        const double x = v[0], y = v[1], z = v[2];
        J[0][0] = 2.0 * x; J[1][0] = 4.0 * x; J[2][0] = 6.0 * x;
        J[0][1] = 2.0 * y; J[1][1] = 2.0 * y; J[2][1] = -4.0;
        J[0][2] = 2.0 * z; J[1][2] = -4.0;    J[2][2] = 2.0 * z;
    }
    // ctest (newton.rs:289-303): converged iff wrms(del, w) <= tol
    __device__ __forceinline__ bool ctest(const double (&del)[3], const double (&w)[3], double tol) const {
        return wrms_reg<3>(del, w) <= tol;
    }
};

template <class NLP>
__global__ void __launch_bounds__(128)
newton_solve_kernel(const double* __restrict__ y0, const double* __restrict__ w, double tol, int maxiters, int call_lsetup_in,
                    double* __restrict__ yout, int32_t* __restrict__ status, int32_t* __restrict__ niters_out,
                    int32_t* __restrict__ nconvfails_out, int64_t B) {
    constexpr int N = NLP::N;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    NLP prob;
    double v0[N], wt[N], y[N], delta[N], J[N][N];
    int piv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        v0[i] = y0[(int64_t)i * B + b];
        wt[i] = w[(int64_t)i * B + b];
        y[i] = 0.0;
        piv[i] = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) J[i][j] = 0.0;
    }
    bool call_lsetup = call_lsetup_in != 0, jcur = false;
    int niters = 0, nconvfails = 0, ret = 0;
#pragma unroll 1
    for (;;) {  // newton.rs:74 'outer
        prob.sys(v0, delta);
        if (call_lsetup) {
            prob.jac(v0, J);
            const int info = getrf_reg<N>(J, piv);
            if (info != 0) {  // Error::LinearSetupFailed { source: LUFactFail { col } }
                ret = -info;
                break;
            }
            jcur = true;
        }
        int curiter = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = v0[i];
        ret = 0;
#pragma unroll 1
        for (;;) {  // 'inner
            niters += 1;
#pragma unroll
            for (int i = 0; i < N; ++i) delta[i] = -delta[i];
            getrs_reg<N>(J, piv, delta);
#pragma unroll
            for (int i = 0; i < N; ++i) y[i] += delta[i];
            if (prob.ctest(delta, wt, tol)) {
                jcur = false;
                ret = 0;
                break;
            }
            curiter += 1;
            if (curiter >= maxiters) {
                ret = 1;  // Error::ConvergenceRecover
                break;
            }
            prob.sys(y, delta);
        }
        if (ret == 1 && !jcur) {  // newton.rs:146-152: retry once on a fresh Jacobian
            nconvfails += 1;
            call_lsetup = true;
            continue;
        }
        break;
    }
    if (ret != 0) nconvfails += 1;  // newton.rs:162
#pragma unroll
    for (int i = 0; i < N; ++i) yout[(int64_t)i * B + b] = y[i];
    if (status) status[b] = ret;
    if (niters_out) niters_out[b] = niters;
    if (nconvfails_out) nconvfails_out[b] = nconvfails;
}

}  // namespace b200ida

using namespace b200ida;

extern "C" int b200ida_newton_batched(int nlp_kind, int n, int64_t batch, const double* y0, const double* w, double tol,
                                      int maxiters, int call_lsetup, double* y, int32_t* status, int32_t* niters,
                                      int32_t* nconvfails, int mem, void* stream) {
    if (!y0 || !w || !y || batch < 1 || maxiters < 1 || (mem != B200IDA_MEM_DEVICE && mem != B200IDA_MEM_HOST))
        return set_error(B200IDA_ERR_ARG, "b200ida_newton_batched: bad argument");
    if (nlp_kind != B200IDA_NLP_QUAD3) return set_error(B200IDA_ERR_UNSUPPORTED, "b200ida_newton_batched: problem kind %d has no functor", nlp_kind);
    if (n != NlpQuad3::N) return set_error(B200IDA_ERR_ARG, "b200ida_newton_batched: kind QUAD3 has n = 3 (got %d)", n);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return set_error(B200IDA_ERR_NO_DEVICE, "b200ida_newton_batched: no CUDA device (there is no CPU fallback)");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((batch + 127) / 128);
    if (mem == B200IDA_MEM_DEVICE) {
        newton_solve_kernel<NlpQuad3><<<grid, 128, 0, st>>>(y0, w, tol, maxiters, call_lsetup, y, status, niters, nconvfails, batch);
        count_launch();
        B200_CUDA_TRY(cudaGetLastError());
        return B200IDA_OK;
    }
    const size_t vb = (size_t)n * batch * sizeof(double), ib = (size_t)batch * sizeof(int32_t);
    DevBuf dy0, dw, dy, dst, dni, dnc;
    B200_CUDA_TRY(dy0.alloc(vb));
    B200_CUDA_TRY(dw.alloc(vb));
    B200_CUDA_TRY(dy.alloc(vb));
    B200_CUDA_TRY(dst.alloc(ib));
    B200_CUDA_TRY(dni.alloc(ib));
    B200_CUDA_TRY(dnc.alloc(ib));
    B200_CUDA_TRY(cudaMemcpyAsync(dy0.p, y0, vb, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(dw.p, w, vb, cudaMemcpyHostToDevice, st));
    newton_solve_kernel<NlpQuad3><<<grid, 128, 0, st>>>(dy0.as<double>(), dw.as<double>(), tol, maxiters, call_lsetup, dy.as<double>(),
                                                       dst.as<int32_t>(), dni.as<int32_t>(), dnc.as<int32_t>(), batch);
    count_launch();
    B200_CUDA_TRY(cudaGetLastError());
    B200_CUDA_TRY(cudaMemcpyAsync(y, dy.p, vb, cudaMemcpyDeviceToHost, st));
    if (status) B200_CUDA_TRY(cudaMemcpyAsync(status, dst.p, ib, cudaMemcpyDeviceToHost, st));
    if (niters) B200_CUDA_TRY(cudaMemcpyAsync(niters, dni.p, ib, cudaMemcpyDeviceToHost, st));
    if (nconvfails) B200_CUDA_TRY(cudaMemcpyAsync(nconvfails, dnc.p, ib, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    return B200IDA_OK;
}
