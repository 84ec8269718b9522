// b200ida.hpp -- header-only C++ mirror of the reference's plugin interfaces over the C ABI (b200ida.h).
//
// The reference is compiled (Rust) code and there is no Rust toolchain in this image, so this is the compiled
// host side above the ABI: same names, argument meaning and error behaviour as
//   linear::LSolver / linear::Dense      crates/linear/src/traits.rs:27-91, crates/linear/src/dense.rs:13-63
//   Ida::new / Ida::solve / getters      src/lib.rs:278-283, src/impl_solve.rs:69-74, src/ida_io.rs
// but acting on a whole ensemble per call.  Buffers are HOST memory here (B200IDA_MEM_HOST): the calls copy in,
// run the CUDA kernels, copy out.  See examples/roberts_ensemble.cpp.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200ida.h"

namespace b200ida_host {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != B200IDA_OK) throw Error(rc, b200ida_last_error());
}

enum class LSolverType { Direct, Iterative, MatrixIterative };  // crates/linear/src/lib.rs:15-20

// linear::Error::LUFactFail { col } (1-based), crates/linear/src/lib.rs:9-13
struct LUFactFail : std::runtime_error {
    int col;
    int64_t index;
    LUFactFail(int c, int64_t i)
        : std::runtime_error("A singular matrix was encountered during a LU factorization (col " + std::to_string(c) + ")"),
          col(c), index(i) {}
};

// `impl LSolver<f64, D> for Dense<D>` over a batch.  Matrices SoA [n][n][batch], vectors SoA [n][batch].
class Dense {
  public:
    Dense(int n, int64_t batch) : n_(n), batch_(batch), pivots_((size_t)n * batch), info_((size_t)batch) {}
    static Dense new_(int n, int64_t batch) { return Dense(n, batch); }           // LSolver::new, traits.rs:32
    LSolverType get_type() const { return LSolverType::Direct; }                 // traits.rs:34
    void setup(double* mat_a) {                                                  // traits.rs:52-54
        check(b200ida_getrf_batched(mat_a, pivots_.data(), info_.data(), n_, batch_, B200IDA_LAYOUT_SOA, B200IDA_MEM_HOST, nullptr));
        for (int64_t b = 0; b < batch_; ++b)
            if (info_[(size_t)b] != 0) throw LUFactFail(info_[(size_t)b], b);
    }
    void solve(const double* mat_a, double* x, const double* b, double /*tol*/) const {  // traits.rs:70-80
        for (size_t i = 0; i < (size_t)n_ * batch_; ++i) x[i] = b[i];                    // dense.rs:58-59
        check(b200ida_getrs_batched(mat_a, pivots_.data(), x, n_, batch_, B200IDA_LAYOUT_SOA, B200IDA_MEM_HOST, nullptr));
    }
    int num_iters() const { return 0; }      // traits.rs:83
    double res_norm() const { return 0.0; }  // traits.rs:88
    const std::vector<int64_t>& pivots() const { return pivots_; }

  private:
    int n_;
    int64_t batch_;
    std::vector<int64_t> pivots_;
    std::vector<int32_t> info_;
};

// `impl NLSolver<f64, D> for Newton<f64, D>` over a batch (crates/nonlinear/src/traits.rs:129-209, newton.rs:14-167).
// The NLProblem is a device functor selected by kind (B200IDA_NLP_*).  Vectors SoA [n][batch].
class Newton {
  public:
    explicit Newton(int maxiters) : maxiters_(maxiters) {}
    static Newton new_(int maxiters) { return Newton(maxiters); }  // NLSolver::new, traits.rs:136
    // NLSolver::solve, traits.rs:176-193; returns the per-system status (0 Ok, 1 ConvergenceRecover, -col LinearSetupFailed)
    std::vector<int32_t> solve(int nlp_kind, int n, int64_t batch, const double* y0, double* y, const double* w, double tol,
                               bool call_lsetup) {
        std::vector<int32_t> st((size_t)batch), ni((size_t)batch), nc((size_t)batch);
        check(b200ida_newton_batched(nlp_kind, n, batch, y0, w, tol, maxiters_, call_lsetup ? 1 : 0, y, st.data(), ni.data(),
                                     nc.data(), B200IDA_MEM_HOST, nullptr));
        if (niters_.size() != (size_t)batch) {
            niters_.assign((size_t)batch, 0);
            nconvfails_.assign((size_t)batch, 0);
        }
        for (size_t b = 0; b < (size_t)batch; ++b) {
            niters_[b] += ni[b];
            nconvfails_[b] += nc[b];
        }
        curiter_ = ni;
        return st;
    }
    const std::vector<int64_t>& get_num_iters() const { return niters_; }           // traits.rs:196
    const std::vector<int32_t>& get_cur_iter() const { return curiter_; }           // traits.rs:202
    const std::vector<int64_t>& get_num_conv_fails() const { return nconvfails_; }  // traits.rs:205

  private:
    int maxiters_;
    std::vector<int64_t> niters_, nconvfails_;
    std::vector<int32_t> curiter_;
};

enum class IdaTask { Normal = B200IDA_TASK_NORMAL };  // src/lib.rs:51-55

// Ida<P, Dense, Newton, TolControlSV> for an ensemble.
class EnsembleIda {
  public:
    // Ida::new(problem, yy0, yp0, TolControlSV::new(rtol, atol)), src/lib.rs:278-283
    EnsembleIda(int problem, int n, int64_t batch, int nparams, const double* params, const double* yy0, const double* yp0,
                double rtol, const double* atol, int device = 0)
        : n_(n), batch_(batch) {
        check(b200ida_create(&h_, problem, n, batch, nparams, params, yy0, yp0, rtol, atol, B200IDA_MEM_HOST, device, nullptr));
    }
    ~EnsembleIda() { b200ida_destroy(h_); }
    EnsembleIda(const EnsembleIda&) = delete;
    EnsembleIda& operator=(const EnsembleIda&) = delete;

    // Ida::solve(tout, &mut tret, itask), src/impl_solve.rs:69-74; returns per-instance B200IDA_ST_* codes
    std::vector<int32_t> solve(double tout, IdaTask itask = IdaTask::Normal) {
        std::vector<int32_t> st((size_t)batch_);
        check(b200ida_solve(h_, tout, (int)itask, st.data(), B200IDA_MEM_HOST));
        return st;
    }
    // every output time of `touts` in ONE launch (each still a separate solve() call per instance: mxstep is per call,
    // h0 comes from the first one); y_traj, if given, receives y(tout_k) as [ntout][n][batch]
    std::vector<int32_t> solve_ladder(const std::vector<double>& touts, std::vector<double>* y_traj = nullptr,
                                      IdaTask itask = IdaTask::Normal) {
        std::vector<int32_t> st(touts.size() * (size_t)batch_);
        if (y_traj) y_traj->assign(touts.size() * (size_t)n_ * batch_, 0.0);
        check(b200ida_solve_ladder(h_, touts.data(), (int)touts.size(), (int)itask, y_traj ? y_traj->data() : nullptr, st.data(),
                                   B200IDA_MEM_HOST));
        return st;
    }
    // Ida::get_dky(t, k, &mut dky), src/lib.rs:424-529: dky [n][batch]; the Result of each instance in `status`
    // (0, B200IDA_ST_BAD_K, B200IDA_ST_BAD_T); columns of failed instances keep what dky held
    std::vector<int32_t> get_dky(double t, int k, std::vector<double>& dky) const {
        dky.resize((size_t)n_ * batch_);
        std::vector<int32_t> st((size_t)batch_);
        check(b200ida_get_dky(h_, t, k, dky.data(), st.data(), B200IDA_MEM_HOST));
        return st;
    }
    std::vector<double> get_yy() const {  // src/ida_io.rs:11
        std::vector<double> y((size_t)n_ * batch_);
        check(b200ida_get_state(h_, y.data(), nullptr, nullptr, B200IDA_MEM_HOST));
        return y;
    }
    std::vector<double> get_yp() const {  // src/ida_io.rs:16
        std::vector<double> y((size_t)n_ * batch_);
        check(b200ida_get_state(h_, nullptr, y.data(), nullptr, B200IDA_MEM_HOST));
        return y;
    }
    std::vector<int64_t> get_num_steps() const { return counter(0); }                   // :58
    std::vector<int64_t> get_num_res_evals() const { return counter(1); }               // :64
    std::vector<int64_t> get_num_jac_evals() const { return counter(2); }               // :79
    std::vector<int64_t> get_num_nonlin_solv_iters() const { return counter(3); }       // :84
    std::vector<int64_t> get_num_err_test_fails() const { return counter(4); }          // :74
    std::vector<int64_t> get_num_nonlin_solv_conv_fails() const { return counter(5); }  // :104
    std::vector<int64_t> get_num_lin_solv_setups() const { return counter(6); }         // :69
    std::vector<int32_t> get_last_order() const { return step_i(0); }                   // :22
    std::vector<int32_t> get_current_order() const { return step_i(1); }                // :27
    std::vector<double> get_last_step() const { return step_d(2); }                     // :38
    std::vector<double> get_current_step() const { return step_d(3); }                  // :43 (`get_current_setp` there)
    std::vector<double> get_current_time() const { return step_d(4); }                  // :48
    std::vector<double> get_actual_init_step() const { return step_d(5); }              // :32

  private:
    std::vector<int64_t> counter(int which) const {
        std::vector<int64_t> c((size_t)batch_);
        int64_t* a[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        a[which] = c.data();
        check(b200ida_get_counters(h_, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], B200IDA_MEM_HOST));
        return c;
    }
    std::vector<int32_t> step_i(int which) const {
        std::vector<int32_t> v((size_t)batch_);
        check(b200ida_get_step_info(h_, which == 0 ? v.data() : nullptr, which == 1 ? v.data() : nullptr, nullptr, nullptr, nullptr,
                                    nullptr, B200IDA_MEM_HOST));
        return v;
    }
    std::vector<double> step_d(int which) const {
        std::vector<double> v((size_t)batch_);
        check(b200ida_get_step_info(h_, nullptr, nullptr, which == 2 ? v.data() : nullptr, which == 3 ? v.data() : nullptr,
                                    which == 4 ? v.data() : nullptr, which == 5 ? v.data() : nullptr, B200IDA_MEM_HOST));
        return v;
    }
    b200ida_ensemble* h_ = nullptr;
    int n_;
    int64_t batch_;
};

}  // namespace b200ida_host
