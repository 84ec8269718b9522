// problems.cuh -- device-side implementations of the reference's problem traits
// (ModelSpec / Residual / Jacobian, src/traits.rs:6-69) for the thread-per-instance kernels.
//
// A GPU step loop cannot call back into host Rust generics once per Newton iteration, so problems are
// device functors selected by B200IDA_PROBLEM_* with per-instance parameter rows (SoA [NP][batch]).
// Arithmetic is written operation-for-operation like the reference (left to right, no FMA), so that
// -fmad=false builds reproduce it bit-for-bit.
#pragma once
#include "ida_common.cuh"

namespace b200ida {

// Robertson kinetics, src/sample_problems/roberts.rs ; (k1,k2,k3) generalise (0.04, 1e4, 3e7).
struct Roberts {
    static constexpr int N = 3;
    static constexpr int NP = 3;
    static constexpr int NR = 2;  // Root::num_roots, roberts.rs:99-113
    static constexpr bool ROOT_USES_YP = false;
    __device__ __forceinline__ void root(double, const double (&y)[N], const double (&)[N], double (&g)[NR]) const {
        g[0] = y[0] - 0.0001;
        g[1] = y[2] - 0.01;
    }
    double k1, k2, k3;
    __device__ __forceinline__ void load(const double* __restrict__ params, int64_t B, int64_t b) {
        k1 = params[b];
        k2 = params[B + b];
        k3 = params[2 * B + b];
    }
    // Residual::res -- roberts.rs:47-62 (arithmetic :58-61)
    __device__ __forceinline__ void res(double, const double (&yy)[N], const double (&yp)[N], double (&r)[N]) const {
        r[0] = -k1 * yy[0] + k2 * yy[1] * yy[2];
        r[1] = -r[0] - k3 * yy[1] * yy[1] - yp[1];
        r[0] -= yp[0];
        r[2] = yy[0] + yy[1] + yy[2] - 1.0;
    }
    // Jacobian::jac -- roberts.rs:66-91 (arithmetic :80-90); J[i][j] = dF_i/dy_j + cj dF_i/dy'_j
    __device__ __forceinline__ void jac(double, double cj, const double (&yy)[N], const double (&)[N],
                                        const double (&)[N], double (&J)[N][N]) const {
        J[0][0] = -k1 - cj;
        J[0][1] = k2 * yy[2];
        J[0][2] = k2 * yy[1];
        J[1][0] = k1;
        J[1][1] = -k2 * yy[2] - (2.0 * k3) * yy[1] - cj;
        J[1][2] = -k2 * yy[1];
        J[2][0] = 1.0;
        J[2][1] = 1.0;
        J[2][2] = 1.0;
    }
};

// Lorenz-63 posed as the DAE F = y' - f(y); RHS f from tests/lorenz63.rs:47-52, params (p, r, b) :17-25.
struct Lorenz63 {
    static constexpr int N = 3;
    static constexpr int NP = 3;
    static constexpr int NR = 0;  // no root functions
    static constexpr bool ROOT_USES_YP = false;
    double p, r, b;
    __device__ __forceinline__ void load(const double* __restrict__ params, int64_t B, int64_t i) {
        p = params[i];
        r = params[B + i];
        b = params[2 * B + i];
    }
    __device__ __forceinline__ void res(double, const double (&yy)[N], const double (&yp)[N], double (&f)[N]) const {
        f[0] = yp[0] - p * (yy[1] - yy[0]);
        f[1] = yp[1] - (yy[0] * (r - yy[2]) - yy[1]);
        f[2] = yp[2] - (yy[0] * yy[1] - b * yy[2]);
    }
    __device__ __forceinline__ void jac(double, double cj, const double (&yy)[N], const double (&)[N],
                                        const double (&)[N], double (&J)[N][N]) const {
        J[0][0] = cj + p;
        J[0][1] = -p;
        J[0][2] = 0.0;
        J[1][0] = yy[2] - r;
        J[1][1] = cj + 1.0;
        J[1][2] = yy[0];
        J[2][0] = -yy[1];
        J[2][1] = -yy[0];
        J[2][2] = cj + b;
    }
};

}  // namespace b200ida
