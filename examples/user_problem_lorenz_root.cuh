// examples/user_problem_lorenz_root.cuh -- a user-supplied problem for the B200 engine (B200IDA_PROBLEM_USER).
//
// The reference's problem traits (src/traits.rs:6-94: ModelSpec, Residual, Jacobian, Root) written as a device functor:
// Lorenz-63 posed as the DAE F = y' - f(y) (tests/lorenz63.rs:17-52) WITH a root function, g0 = y0 (the x = 0
// crossings between the two lobes of the attractor), which the built-in Lorenz63 problem does not have.
//
//     make -C rust_ida_b200/csrc user PROBLEM_HDR=$PWD/examples/user_problem_lorenz_root.cuh
//     B200IDA_LIB=rust_ida_b200/csrc/libb200ida_user.so python ...      (rust_ida_b200.problems.UserProblem)
//
// Arithmetic is IEEE un-fused, left to right as written (the engine is built with -fmad=false).
#pragma once
#include <stdint.h>

namespace b200ida {

struct UserProblem {
    static constexpr int N = 3;    // ModelSpec::model_size
    static constexpr int NP = 3;   // per-instance parameter rows (p, r, b)
    static constexpr int NR = 1;   // Root::num_roots
    static constexpr bool ROOT_USES_YP = false;
    double p, r, b;
    __device__ __forceinline__ void load(const double* __restrict__ params, int64_t B, int64_t i) {
        p = params[i];
        r = params[B + i];
        b = params[2 * B + i];
    }
    // Residual::res (src/traits.rs:28-38)
    __device__ __forceinline__ void res(double, const double (&yy)[N], const double (&yp)[N], double (&f)[N]) const {
        f[0] = yp[0] - p * (yy[1] - yy[0]);
        f[1] = yp[1] - (yy[0] * (r - yy[2]) - yy[1]);
        f[2] = yp[2] - (yy[0] * yy[1] - b * yy[2]);
    }
    // Jacobian::jac (src/traits.rs:57-69): J[i][j] = dF_i/dy_j + cj dF_i/dy'_j
    __device__ __forceinline__ void jac(double, double cj, const double (&yy)[N], const double (&)[N], const double (&)[N],
                                        double (&J)[N][N]) const {
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
    // Root::root (src/traits.rs:72-89)
    __device__ __forceinline__ void root(double, const double (&y)[N], const double (&)[N], double (&g)[NR]) const { g[0] = y[0]; }
};

}  // namespace b200ida
