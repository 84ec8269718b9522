// ida_oracle.cpp -- CPU restatement of rust-ida's BDF/Newton hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
// (rust_ida_b200/csrc) never links, includes or calls anything in oracle/.
//
// What it is: a scalar, single-instance, line-by-line C++17 restatement of the arithmetic
// of jondo2010/rust-ida (a pure-Rust port of SUNDIALS IDA), written from the reference's
// behaviour; every function cites the reference file:line it follows.  The Rust reference
// cannot be built in this environment (no rustc/cargo; the tree is mid-refactor, see
// SURVEY.md section 8c), so this file is a PORT, pinned against every golden vector the
// reference's own tests hold (tests/golden/*.json, extracted by tools/extract_golden.py):
//   bit-exact: LU getrf/getrs known answers, complete_step x3 (incl. rr through pow), norms
//   16 digits: set_coeffs x2, predict, restore, test_error x2, get_solution
//   end-to-end: Roberts to t=4e10 (check_ans < 1; 377 step attempts as printed by the
//   reference author's trace notebook).
//
// Arithmetic policy: compile with  g++ -O2 -ffp-contract=off -fno-fast-math  so that no
// a*b+c is fused (Rust never contracts), evaluation is left-to-right as written in the
// reference, x.recip() is 1.0/x, powi(2) is x*x, powf is the platform libm pow (glibc here).
// The pow implementation is injectable (oracle_set_pow) so tests can swap in the
// bit-reproducible glibc-pow port that the GPU engine uses and measure the difference.
//
// Deliberate deviations from the reference (hard part 4), all flagged
// where they occur:
//   * Newton "ConvergenceRecover with a current Jacobian": the new-API Newton re-loops
//     forever (crates/nonlinear/src/newton.rs:146-153) and the old-API handle_n_flag treats it
//     as fatal (src/lib.rs:1133-1140).  We follow SUNDIALS C / the evident intent of
//     src/lib.rs:1201-1237: return the recoverable error, ncfn++, rr=0.25, retry <= MXNCF.
//   * singular Jacobian: the reference panics (src/ida_ls.rs:287 unwrap); we return status
//     ORC_LSETUP_FAIL for that instance.
// Bug-compatible on purpose: reset() scales ALL of phi (src/lib.rs:1251); jac gets tt=0
// (src/ida_ls.rs:260-261).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---- constants: src/constants.rs:4-31,56-112 ; RATEMAX src/ida_nls.rs:15 -------------------
constexpr int MXORDP1 = 6;
constexpr int MAXORD_DEFAULT = 5;
constexpr long MXSTEP_DEFAULT = 500;
constexpr int MXNCF = 10;
constexpr int MXNEF = 10;
constexpr double EPCON = 0.33;
constexpr double XRATE = 0.25;
constexpr int MAXNLSIT = 4;
constexpr double RATEMAX = 0.9;
constexpr double HMAX_INV_DEFAULT = 0.0;
constexpr double DBL_EPS = 2.220446049250313e-16;  // f64::EPSILON

// ---- algorithmic operation counter (-DORACLE_COUNT_OPS build only: oracle/libida_oracle_ops.so) -----------------
// One CNT per arithmetic block, placed where the block is: FP64 additions / subtractions, multiplications, divisions,
// square roots and pow calls of the REFERENCE algorithm (negation, fabs, min / max, comparisons and integer -> double
// conversions are not counted).  bench.py turns the per-instance totals into the work unit of its FP64 roofline
// (ops per step attempt); the production build of the oracle compiles CNT away.
#ifdef ORACLE_COUNT_OPS
struct OpCount { long add = 0, mul = 0, div = 0, sqrt_ = 0, pow_ = 0; };
thread_local OpCount g_ops;
#define CNT(a, m, d, s, p) do { g_ops.add += (a); g_ops.mul += (m); g_ops.div += (d); g_ops.sqrt_ += (s); g_ops.pow_ += (p); } while (0)
#else
#define CNT(a, m, d, s, p) do { } while (0)
#endif

typedef double (*pow_fn_t)(double, double);
pow_fn_t g_pow = static_cast<double (*)(double, double)>(std::pow);

// status codes shared with the engine (include/b200ida.h B200IDA_ST_*)
enum Status : int {
    ST_SUCCESS = 0,
    ST_ROOT = 2,             // IdaSolveStatus::Root (src/lib.rs:57-63)
    ST_CLOSE_ROOTS = -9,     // IdaError::CloseRoots (src/impl_r_check.rs:199-203)
    ST_TOO_MUCH_WORK = -1,   // mxstep steps taken (IdaError::IllegalInput in the reference)
    ST_TOO_MUCH_ACC = -2,    // IdaError::TooMuchAccuracy
    ST_ERR_FAIL = -3,        // "IDA_ERR_FAIL"
    ST_CONV_FAIL = -4,       // IdaError::ConvergenceFail
    ST_LSETUP_FAIL = -5,     // singular Jacobian (reference: panic)
    ST_BAD_EWT = -6,         // some ewt component <= 0
    ST_ILL_INPUT = -7,       // tout too close to t0
    ST_BAD_K = -10,          // IdaError::BadK from get_dky (src/lib.rs:443-447)
    ST_BAD_T = -8,           // get_solution BadTimeValue (propagated from stop_test1)
};

enum ProblemKind : int { PK_ROBERTS = 0, PK_LORENZ63 = 1, PK_HEAT2D = 2 };

// ---- problems ------------------------------------------------------------------------------
struct Problem {
    int kind;
    int n;
    const double* p;  // per-instance parameters

    // Root::num_roots / Root::root -- src/traits.rs:72-89 ; Roberts: src/sample_problems/roberts.rs:99-113
    // Lorenz: g0 = y0, the x = 0 crossings (not in the reference; the oracle side of the user-problem example,
    // examples/user_problem_lorenz_root.cuh -- rootfinding is only active when a caller enables it)
    int num_roots() const { return kind == PK_ROBERTS ? 2 : (kind == PK_LORENZ63 ? 1 : 0); }
    void root(double /*t*/, const double* y, const double* /*yp*/, double* gout) const {
        if (kind == PK_ROBERTS) {
            gout[0] = y[0] - 0.0001;
            gout[1] = y[2] - 0.01;
        } else if (kind == PK_LORENZ63) {
            gout[0] = y[0];
        }
    }

    // Residual::res -- src/traits.rs:28-38
    void res(double /*tt*/, const double* yy, const double* yp, double* r) const {
        switch (kind) {
        case PK_ROBERTS: {
            // src/sample_problems/roberts.rs:58-61 with (0.04, 1e4, 3e7) -> (k1,k2,k3)
            const double k1 = p[0], k2 = p[1], k3 = p[2];
            r[0] = -k1 * yy[0] + k2 * yy[1] * yy[2];
            r[1] = -r[0] - k3 * yy[1] * yy[1] - yp[1];
            r[0] -= yp[0];
            r[2] = yy[0] + yy[1] + yy[2] - 1.0;
            CNT(7, 5, 0, 0, 0);
            break;
        }
        case PK_LORENZ63: {
            // tests/lorenz63.rs:47-52 (commented RHS f), posed as F = y' - f(y)
            const double s = p[0], rr = p[1], b = p[2];
            r[0] = yp[0] - s * (yy[1] - yy[0]);
            r[1] = yp[1] - (yy[0] * (rr - yy[2]) - yy[1]);
            r[2] = yp[2] - (yy[0] * yy[1] - b * yy[2]);
            CNT(7, 4, 0, 0, 0);
            break;
        }
        case PK_HEAT2D: {
            // not in the reference; SUNDIALS idaHeat2D: M x M grid, boundary rows algebraic
            const int M = (int)std::lround(std::sqrt((double)n));
            const double coeff = p[0];
            for (int i = 0; i < n; ++i) r[i] = yy[i];
            for (int j = 1; j < M - 1; ++j)
                for (int i = 1; i < M - 1; ++i) {
                    const int loc = j * M + i;
                    const double dif1 = yy[loc - 1] + yy[loc + 1] - 2.0 * yy[loc];
                    const double dif2 = yy[loc - M] + yy[loc + M] - 2.0 * yy[loc];
                    r[loc] = yp[loc] - coeff * (dif1 + dif2);
                    CNT(6, 3, 0, 0, 0);
                }
            break;
        }
        }
    }

    // Jacobian::jac -- src/traits.rs:57-69 ; J(i,j) = dF_i/dy_j + cj dF_i/dy'_j, column-major
    void jac(double /*tt*/, double cj, const double* yy, const double* /*yp*/, const double* /*r*/,
             double* J) const {
        auto A = [&](int i, int j) -> double& { return J[j * n + i]; };
        switch (kind) {
        case PK_ROBERTS: {
            // src/sample_problems/roberts.rs:80-90 ; 6.0e7 == 2*k3 exactly
            const double k1 = p[0], k2 = p[1], k3 = p[2];
            A(0, 0) = -k1 - cj;
            A(0, 1) = k2 * yy[2];
            A(0, 2) = k2 * yy[1];
            A(1, 0) = k1;
            A(1, 1) = -k2 * yy[2] - (2.0 * k3) * yy[1] - cj;
            A(1, 2) = -k2 * yy[1];
            A(2, 0) = 1.0;
            A(2, 1) = 1.0;
            A(2, 2) = 1.0;
            CNT(3, 6, 0, 0, 0);
            break;
        }
        case PK_LORENZ63: {
            const double s = p[0], rr = p[1], b = p[2];
            A(0, 0) = cj + s;
            A(0, 1) = -s;
            A(0, 2) = 0.0;
            A(1, 0) = yy[2] - rr;
            A(1, 1) = cj + 1.0;
            A(1, 2) = yy[0];
            A(2, 0) = -yy[1];
            A(2, 1) = -yy[0];
            A(2, 2) = cj + b;
            CNT(4, 0, 0, 0, 0);
            break;
        }
        case PK_HEAT2D: {
            const int M = (int)std::lround(std::sqrt((double)n));
            const double coeff = p[0];
            for (int i = 0; i < n; ++i) A(i, i) = 1.0;
            for (int j = 1; j < M - 1; ++j)
                for (int i = 1; i < M - 1; ++i) {
                    const int loc = j * M + i;
                    CNT(1, 1, 0, 0, 0);
                    A(loc, loc) = cj + 4.0 * coeff;
                    A(loc, loc - 1) = -coeff;
                    A(loc, loc + 1) = -coeff;
                    A(loc, loc - M) = -coeff;
                    A(loc, loc + M) = -coeff;
                }
            break;
        }
        }
    }
};

// ---- dense LU: crates/linear/src/dense.rs -----------------------------------------------------
// dense_get_rf: dense.rs:86-158.  A column-major n x n, returns 0 or the 1-based column of the
// zero pivot.  pivot = first row with strictly largest |a_ik| (:111-116); whole-row swap
// (:126-130); mult = 1/a_kk then multiply (:134-137); column-oriented update skipped when
// a_kj == 0 (:142-154).
int dense_get_rf(double* a, int64_t* pivot, int n) {
    auto A = [&](int i, int j) -> double& { return a[(size_t)j * n + i]; };
    for (int k = 0; k < n; ++k) {
        int l = k;
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(A(i, k)) > std::fabs(A(l, k))) l = i;
        pivot[k] = l;
        if (A(l, k) == 0.0) return k + 1;
        if (l != k)
            for (int i = 0; i < n; ++i) std::swap(A(k, i), A(l, i));
        const double mult = 1.0 / A(k, k);
        for (int i = k + 1; i < n; ++i) A(i, k) *= mult;
        CNT(0, n - k - 1, 1, 0, 0);
        for (int j = k + 1; j < n; ++j) {
            const double a_kj = A(k, j);
            if (a_kj != 0.0) {
                CNT(n - k - 1, n - k - 1, 0, 0, 0);
                for (int i = k + 1; i < n; ++i) {
                    const double a_ik = A(i, k);
                    A(i, j) -= a_kj * a_ik;
                }
            }
        }
    }
    return 0;
}

// dense_get_rs: dense.rs:165-206.
void dense_get_rs(const double* a, const int64_t* pivot, double* b, int n) {
    auto A = [&](int i, int j) -> double { return a[(size_t)j * n + i]; };
    for (int k = 0; k < n; ++k) {
        const int64_t pk = pivot[k];
        if (pk != k) std::swap(b[k], b[pk]);
    }
    for (int k = 0; k < n - 1; ++k) {
        const double bk = b[k];
        for (int i = k + 1; i < n; ++i) b[i] -= A(i, k) * bk;
    }
    for (int k = n - 1; k >= 1; --k) {
        b[k] /= A(k, k);
        const double bk = b[k];
        for (int i = 0; i < k; ++i) b[i] -= A(i, k) * bk;
    }
    b[0] /= A(0, 0);
    CNT(n * (n - 1), n * (n - 1), n, 0, 0);
}

// ---- WRMS norm: src/norm_rms.rs:31-38 ; crates/nonlinear/src/norm_wrms.rs:24-33 -------------
double norm_wrms(const double* x, const double* w, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        const double p = x[i] * w[i];
        acc = acc + p * p;
    }
    CNT(n, 2 * n, 1, 1, 0);
    return std::sqrt(acc / (double)n);
}

// ---- masked WRMS norm: src/norm_rms.rs:49-57 (mask = id ? 1 : 0; ((x*w)*mask)^2 folded from 0, / len, sqrt) ----
double norm_wrms_masked(const double* x, const double* w, const unsigned char* id, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        const double p = (x[i] * w[i]) * (id[i] ? 1.0 : 0.0);
        acc = acc + p * p;
    }
    CNT(n, 3 * n, 1, 1, 0);
    return std::sqrt(acc / (double)n);
}

// ---- Newton: crates/nonlinear/src/newton.rs:14-32 (state) and :51-167 (solve) -------------------
// Generic over an NLProblem-like object P providing (crates/nonlinear/src/traits.rs:5-127):
//   void sys(const double* y, double* f); int setup(const double* y, const double* f, bool jbad)
//   [0 ok, else fatal]; void lsolve(const double* y, double* b);
//   int ctest(int curiter, const double* y, const double* del, double tol, const double* w)
//   [1 converged, 0 continue, -1 ConvergenceRecover].
// C semantics for the "recover while jcur" arm (see file header).
struct Newton {
    std::vector<double> delta;
    bool jcur = false;
    int curiter = 0;
    int maxiters = MAXNLSIT;
    long niters = 0;
    long nconvfails = 0;

    // returns 0 ok, 1 recoverable convergence failure, <0 fatal from setup
    template <class P>
    int solve(P& problem, int n, const double* y0, double* y, const double* w, double tol, bool call_lsetup) {
        bool jbad = false;
        int ret = 0;
        for (;;) {
            problem.sys(y0, delta.data());
            if (call_lsetup) {
                const int s = problem.setup(y0, delta.data(), jbad);
                if (s != 0) { ret = s; break; }
                jcur = true;
            }
            curiter = 0;
            for (int i = 0; i < n; ++i) y[i] = y0[i];
            for (;;) {
                niters += 1;
                for (int i = 0; i < n; ++i) delta[i] = -delta[i];
                problem.lsolve(y, delta.data());
                for (int i = 0; i < n; ++i) y[i] += delta[i];
                CNT(n, 0, 0, 0, 0);
                const int c = problem.ctest(curiter, y, delta.data(), tol, w);
                if (c == 1) {
                    jcur = false;
                    return 0;
                }
                if (c == -1) { ret = 1; break; }
                curiter += 1;
                if (curiter >= maxiters) { ret = 1; break; }
                problem.sys(y, delta.data());
            }
            if (ret == 1 && !jcur) {  // newton.rs:146-152
                nconvfails += 1;
                call_lsetup = true;
                jbad = true;
                continue;
            }
            break;
        }
        nconvfails += 1;  // newton.rs:162
        return ret;
    }
};

// ---- Ida state: src/lib.rs:89-244, src/ida_nls.rs:20-60, src/ida_ls.rs:15-106,
//      crates/nonlinear/src/newton.rs:14-32, crates/linear/src/dense.rs:13-20 -----------------
struct Ida {
    Problem prob;
    int n;
    double rtol;
    std::vector<double> atol;  // TolControlSV src/tol_control.rs:51-83

    bool setup_done = false;
    // ida_id (true = differential component) and ida_suppressalg, src/lib.rs:102,122: when suppressalg is set, the
    // local-error norms leave the algebraic components out (wrms_norm, src/lib.rs:1353-1370).  The reference has no
    // setter (both stay false, :318); the oracle and the engine expose one (b200ida_set_id).
    std::vector<unsigned char> id;
    bool suppressalg = false;
    // wrms_norm -- src/lib.rs:1353-1370: mask = false from the nonlinear solver, = suppressalg everywhere else
    double wrms_norm(const double* x, const double* w, bool mask) const {
        if (mask) return norm_wrms_masked(x, w, id.data(), n);
        return norm_wrms(x, w, n);
    }
    std::vector<double> phi;  // [6][n]
    double psi[MXORDP1], alpha[MXORDP1], beta[MXORDP1], sigma[MXORDP1], gamma[MXORDP1];
    std::vector<double> delta, ee, ewt, yy, yp, yypredict, yppredict, savres, x, tmp;
    std::vector<double> J;
    std::vector<int64_t> pivots;

    int kk = 0, kused = 0, knew = 0, phase = 0, ns = 0;
    double hin = 0, h0u = 0, hh = 0, hused = 0, rr = 0, tretlast = 0, cjlast = 0;
    double eps_newt = 0, epcon = EPCON, hmax_inv = HMAX_INV_DEFAULT, tolsf = 1.0;
    int maxncf = MXNCF, maxnef = MXNEF, maxord = MAXORD_DEFAULT;
    long mxstep = MXSTEP_DEFAULT;
    // IdaNLProblem / IdaLProblem scalars
    double tn = 0, ss = 0, oldnrm = 0, toldel = 0, cj = 0, cjold = 0, cjratio = 0;
    Newton nls;
    // counters: src/lib.rs:70-84 ; src/ida_nls.rs:51,54 ; src/ida_ls.rs:45 ; newton.rs:29-31
    long nst = 0, nre = 0, nje = 0, nsetups = 0, netf = 0, ncfn = 0;
    long nattempts = 0;  // one per pass of the step loop (= data_trace frames, src/lib.rs:635-639)

    double* PHI(int j) { return &phi[(size_t)j * n]; }

    // ---- per-attempt state trace in the engine's field order (include/b200ida.h, b200ida_set_trace) --------------
    double* trace_d = nullptr;
    int32_t* trace_i = nullptr;
    long trace_max = 0, trace_count = 0;
    double trace_tret = 0.0;  // the caller's tret (a local of Ida::solve's caller in the reference)
    // ---- per-attempt event log (lane-divergence analysis, tests/analysis/sim_lanes.py): one word per attempt --------
    // bits 0-2 kk at the top, 3 coefficients recomputed (kk+1 >= ns), 4-6 ns after its update, 7-8 Jacobian setups,
    // 9-12 Newton iterations, 13-14 wrms calls of test_error (0 = not run), 15 converged, 16 pow in complete_step /
    // handle_n_flag, 17 err_kp1 norm formed, 18-20 kk after, 21 phase after, 22 output time reached by this step
    uint32_t* ev = nullptr;
    float* ev_h = nullptr;  // |hh| after the attempt (sort-key experiments)
    float* ev_c = nullptr;  // cj the next attempt will use / cjold, i.e. the next attempt's cjratio (optional)
    float ev_cjr() const {
        double a = 0.0;
        for (int j = 1; j <= kk; ++j) a += 1.0 / (double)j;
        return (float)((a / hh) / cjold);
    }
    long ev_max = 0, ev_count = 0;
    uint32_t ev_word = 0;
    void dump_frame() {
        const long c = trace_count++;
        if (c >= trace_max) return;
        const int nd = 10 * n + 45 + n * n, ni = 16 + n;
        double* d = trace_d + (size_t)c * nd;
        int32_t* q = trace_i + (size_t)c * ni;
        int o = 0;
        for (int j = 0; j < MXORDP1; ++j)
            for (int i = 0; i < n; ++i) d[o++] = PHI(j)[i];
        for (const double* a : {psi, alpha, beta, sigma, gamma})
            for (int j = 0; j < MXORDP1; ++j) d[o++] = a[j];
        for (const std::vector<double>* v : {&ee, &ewt, &yy, &yp})
            for (int i = 0; i < n; ++i) d[o++] = (*v)[i];
        const double sc[15] = {hh, hused, rr, cj, cjold, cjlast, cjratio, ss, toldel, eps_newt, tn, tretlast, h0u, tolsf, trace_tret};
        for (double v : sc) d[o++] = v;
        for (int e = 0; e < n * n; ++e) d[o++] = J[e];  // column-major like the engine
        const long iv[16] = {kk, kused, knew, phase, ns, setup_done ? 1 : 0, nls.jcur ? 1 : 0, nst, nre, nje, nls.niters, netf, ncfn,
                             nsetups, nls.nconvfails, nattempts};
        for (int k = 0; k < 16; ++k) q[k] = (int32_t)iv[k];
        for (int i = 0; i < n; ++i) q[16 + i] = (int32_t)pivots[i];
    }

    // Ida::new -- src/lib.rs:278-405
    void init(const Problem& p, const double* yy0, const double* yp0, double rtol_, const double* atol_) {
        prob = p;
        n = p.n;
        rtol = rtol_;
        atol.assign(atol_, atol_ + n);
        if ((int)id.size() != n) id.assign(n, 0);
        phi.assign((size_t)MXORDP1 * n, 0.0);
        std::copy(yy0, yy0 + n, PHI(0));
        std::copy(yp0, yp0 + n, PHI(1));
        for (int i = 0; i < MXORDP1; ++i) psi[i] = alpha[i] = beta[i] = sigma[i] = gamma[i] = 0.0;
        for (auto* v : {&delta, &ee, &ewt, &yypredict, &yppredict, &savres, &x, &tmp, &nls.delta})
            v->assign(n, 0.0);
        yy.assign(yy0, yy0 + n);  // src/ida_nls.rs:82-83
        yp.assign(yp0, yp0 + n);
        J.assign((size_t)n * n, 0.0);
        pivots.assign(n, 0);
    }

    // TolControlSV::ewt_set -- src/tol_control.rs:71-82
    void ewt_set(const double* ycur) {
        for (int i = 0; i < n; ++i) ewt[i] = 1.0 / (rtol * std::fabs(ycur[i]) + atol[i]);
        CNT(n, n, n, 0, 0);
    }

    // set_coeffs -- src/lib.rs:722-782
    double set_coeffs() {
        if (hh != hused || kk != kused) ns = 0;
        ns = std::min(ns + 1, kused + 2);
        if (kk + 1 >= ns) {
            beta[0] = 1.0;
            alpha[0] = 1.0;
            double temp1 = hh;
            gamma[0] = 0.0;
            sigma[0] = 1.0;
            for (int i = 1; i <= kk; ++i) {
                const double scalar_i = (double)i;
                const double temp2 = psi[i - 1];
                psi[i - 1] = temp1;
                beta[i] = beta[i - 1] * psi[i - 1] / temp2;
                temp1 = temp2 + hh;
                alpha[i] = hh / temp1;
                sigma[i] = scalar_i * sigma[i - 1] * alpha[i];
                gamma[i] = gamma[i - 1] + alpha[i - 1] / hh;
                CNT(2, 3, 3, 0, 0);
            }
            psi[kk] = temp1;
        }
        double alphas = 0.0, alpha0 = 0.0;
        for (int i = 0; i < kk; ++i) {
            const double scalar_i = (double)(i + 1);
            alphas -= 1.0 / scalar_i;
            alpha0 -= alpha[i];
            CNT(2, 0, 1, 0, 0);
        }
        cjlast = cj;
        cj = -alphas / hh;
        double ck = std::fabs(alpha[kk] + alphas - alpha0);
        CNT(2, 0, 1, 0, 0);
        ck = std::fmax(ck, alpha[kk]);
        if (ns <= kk)
            for (int j = ns; j <= kk; ++j) {
                double* ph = PHI(j);
                for (int i = 0; i < n; ++i) ph[i] *= beta[j];
                CNT(0, n, 0, 0, 0);
            }
        return ck;
    }

    // predict -- src/lib.rs:894-959
    void predict() {
        for (int i = 0; i < n; ++i) yypredict[i] = 0.0;
        for (int j = 0; j <= kk; ++j) {
            const double* ph = PHI(j);
            for (int i = 0; i < n; ++i) yypredict[i] += ph[i];
        }
        for (int i = 0; i < n; ++i) yppredict[i] = 0.0;
        for (int j = 1; j <= kk; ++j) {
            const double* ph = PHI(j);
            for (int i = 0; i < n; ++i) yppredict[i] += gamma[j] * ph[i];
        }
        CNT((kk + 1) * n + kk * n, kk * n, 0, 0, 0);
    }

    // IdaNLProblem::sys -- src/ida_nls.rs:118-153
    void nls_sys(const double* ycor, double* res) {
        for (int i = 0; i < n; ++i) yy[i] = yypredict[i] + ycor[i];
        for (int i = 0; i < n; ++i) yp[i] = yppredict[i];
        for (int i = 0; i < n; ++i) yp[i] += cj * ycor[i];
        CNT(2 * n, n, 0, 0, 0);
        prob.res(tn, yy.data(), yp.data(), res);
        nre += 1;
        for (int i = 0; i < n; ++i) savres[i] = res[i];
    }

    // IdaNLProblem::setup -- src/ida_nls.rs:156-187 -> IdaLProblem::setup src/ida_ls.rs:232-290
    // returns 0 or the 1-based singular column
    int nls_setup(const double* res) {
        nsetups += 1;
        nje += 1;
        std::fill(J.begin(), J.end(), 0.0);
        prob.jac(0.0, cj, yy.data(), yp.data(), res, J.data());  // tt = 0: src/ida_ls.rs:260-261
        const int info = dense_get_rf(J.data(), pivots.data(), n);  // Dense::setup dense.rs:38-44
        if (info != 0) return info;  // reference: unwrap() panic, src/ida_ls.rs:287
        cjold = cj;
        cjratio = 1.0;
        ss = 20.0;
        return 0;
    }

    // IdaNLProblem::solve -- src/ida_nls.rs:190-215 -> IdaLProblem::solve src/ida_ls.rs:298-455
    void nls_lsolve(double* b) {
        for (int i = 0; i < n; ++i) x[i] = 0.0;
        for (int i = 0; i < n; ++i) x[i] = b[i];  // Dense::solve copies b into x, dense.rs:58-59
        dense_get_rs(J.data(), pivots.data(), x.data(), n);
        for (int i = 0; i < n; ++i) b[i] = x[i];
        if (cjratio != 1.0) {  // src/ida_ls.rs:406-410
            const double f = 2.0 / (1.0 + cjratio);
            for (int i = 0; i < n; ++i) b[i] *= f;
            CNT(1, n, 1, 0, 0);
        }
    }

    // IdaNLProblem::ctest -- src/ida_nls.rs:218-266 ; 1 converged, 0 continue, -1 recover
    int nls_ctest(int m, const double* del, double tol, const double* w) {
        const double delnrm = norm_wrms(del, w, n);
        if (m == 0) {
            oldnrm = delnrm;
            CNT(0, 1, 0, 0, 0);
            if (delnrm <= 0.0001 * toldel) return 1;
        } else {
            const double base = delnrm / oldnrm;
            const double arg = 1.0 / (double)m;
            const double rate = g_pow(base, arg);
            CNT(0, 0, 2, 0, 1);
            if (rate > RATEMAX) return -1;
            ss = rate / (1.0 - rate);
            CNT(1, 0, 1, 0, 0);
        }
        CNT(0, 1, 0, 0, 0);
        if (ss * delnrm <= tol) return 1;
        return 0;
    }

    // adapter giving Newton::solve the NLProblem view of this Ida (src/ida_nls.rs:104-266)
    struct NlpView {
        Ida& s;
        void sys(const double* y, double* f) { s.nls_sys(y, f); }
        int setup(const double*, const double* f, bool) { return s.nls_setup(f) != 0 ? (int)ST_LSETUP_FAIL : 0; }
        void lsolve(const double*, double* b) { s.nls_lsolve(b); }
        int ctest(int curiter, const double*, const double* del, double tol, const double* w) {
            return s.nls_ctest(curiter, del, tol, w);
        }
    };

    // nonlinear_solve -- src/lib.rs:787-890
    int nonlinear_solve() {
        bool call_lsetup = false;
        if (nst == 0) {
            cjold = cj;
            ss = 20.0;
            call_lsetup = true;
        }
        cjratio = cj / cjold;
        const double temp1 = (1.0 - XRATE) / (1.0 + XRATE);
        const double temp2 = 1.0 / temp1;
        CNT(0, 0, 1, 0, 0);  // cjratio only: temp1 / temp2 are compile-time constants in any optimised build
        if (cjratio < temp1 || cjratio > temp2) call_lsetup = true;
        if (cj != cjlast) ss = 100.0;
        for (int i = 0; i < n; ++i) delta[i] = 0.0;
        // w = ewt.clone()
        NlpView view{*this};
        const int ret = nls.solve(view, n, delta.data(), ee.data(), ewt.data(), eps_newt, call_lsetup);
        for (int i = 0; i < n; ++i) yy[i] = yypredict[i] + ee[i];
        for (int i = 0; i < n; ++i) yp[i] = yppredict[i];
        for (int i = 0; i < n; ++i) yp[i] += cj * ee[i];
        CNT(2 * n, n, 0, 0, 0);
        return ret;
    }

    // test_error -- src/lib.rs:967-1039 ; returns converged
    bool test_error(double ck, double& err_k, double& err_km1) {
        const double scalar_kk = (double)kk;
        const double enorm_k = wrms_norm(ee.data(), ewt.data(), suppressalg);
        err_k = sigma[kk] * enorm_k;
        const double terr_k = err_k * (scalar_kk + 1.0);
        CNT(1, 3, 0, 0, 0);  // incl. ck * enorm_k of the acceptance test
        err_km1 = 0.0;
        int knew_ = kk;
        if (kk > 1) {
            CNT(n, 2, 0, 0, 0);
            const double* phk = PHI(kk);
            for (int i = 0; i < n; ++i) delta[i] = phk[i] + ee[i];
            const double enorm_km1 = wrms_norm(delta.data(), ewt.data(), suppressalg);
            err_km1 = sigma[kk - 1] * enorm_km1;
            const double terr_km1 = scalar_kk * err_km1;
            if (kk > 2) {
                CNT(n + 1, 2, 0, 0, 0);
                const double* phkm1 = PHI(kk - 1);
                for (int i = 0; i < n; ++i) delta[i] += phkm1[i];
                const double enorm_km2 = wrms_norm(delta.data(), ewt.data(), suppressalg);
                const double err_km2 = sigma[kk - 2] * enorm_km2;
                const double terr_km2 = (scalar_kk - 1.0) * err_km2;
                if (std::fmax(terr_km1, terr_km2) <= terr_k) knew_ = kk - 1;
            } else {
                CNT(0, 1, 0, 0, 0);
                if (terr_km1 <= terr_k * 0.5) knew_ = kk - 1;
            }
        }
        knew = knew_;
        return (ck * enorm_k) <= 1.0;
    }

    // restore -- src/lib.rs:1044-1083
    void restore(double saved_t) {
        tn = saved_t;
        for (int j = 1; j <= kk; ++j) psi[j - 1] = psi[j] - hh;
        CNT(kk, 0, 0, 0, 0);
        if (ns <= kk)
            for (int j = ns; j <= kk; ++j) {
                CNT(0, n, 1, 0, 0);
                const double c = 1.0 / beta[j];
                double* ph = PHI(j);
                for (int i = 0; i < n; ++i) ph[i] *= c;
            }
    }

    // handle_n_flag -- src/lib.rs:1120-1244 ; nflag: 2 = error test fail, 1 = recoverable conv
    // fail.  returns 0 (predict again) or a fatal Status.
    int handle_n_flag(int nflag, double err_k, double err_km1, long& ncf, long& nef) {
        phase = 1;
        if (nflag == 2) {
            nef += 1;
            netf += 1;
            if (nef == 1) {
                const double err_knew = (kk == knew) ? err_k : err_km1;
                kk = knew;
                const double base = 2.0 * err_knew + 0.0001;
                const double arg = 1.0 / (double)(kk + 1);
                rr = 0.9 * g_pow(base, -arg);
                rr = std::fmax(0.25, std::fmin(0.9, rr));
                hh *= rr;
                CNT(1, 3, 1, 0, 1);
                return 0;
            } else if (nef == 2) {
                kk = knew;
                rr = 0.25;
                hh *= rr;
                CNT(0, 1, 0, 0, 0);
                return 0;
            } else if (nef < maxnef) {
                kk = 1;
                rr = 0.25;
                hh *= rr;
                CNT(0, 1, 0, 0, 0);
                return 0;
            }
            return ST_ERR_FAIL;
        }
        // recoverable convergence failure (C semantics; src/lib.rs:1201-1237)
        ncf += 1;
        ncfn += 1;
        rr = 0.25;
        hh *= rr;
        CNT(0, 1, 0, 0, 0);
        if (ncf < maxncf) return 0;
        return ST_CONV_FAIL;
    }

    // reset -- src/lib.rs:1249-1252 (scales ALL of phi: bug-compatible)
    void reset() {
        psi[0] = hh;
        for (auto& v : phi) v *= rr;
        CNT(0, MXORDP1 * n, 0, 0, 0);
    }

    // complete_step -- src/impl_complete_step.rs:22-177
    void complete_step(double err_k, double err_km1) {
        nst += 1;
        const int kdiff = kk - kused;
        kused = kk;
        hused = hh;
        if (knew == kk - 1 || kk == maxord) phase = 1;
        if (phase == 0) {
            if (nst > 1) {
                kk += 1;
                double hnew = 2.0 * hh;
                const double t = std::fabs(hnew) * hmax_inv;
                if (t > 1.0) hnew /= t;
                hh = hnew;
                CNT(0, 2, 0, 0, 0);
            }
        } else {
            enum { LOWER, MAINTAIN, RAISE } action;
            double err_kp1 = 0.0;
            if (knew == kk - 1) {
                action = LOWER;
            } else if (kk == maxord) {
                action = MAINTAIN;
            } else if (kk + 1 >= ns || kdiff == 1) {
                action = MAINTAIN;
            } else {
                const double* php1 = PHI(kk + 1);
                for (int i = 0; i < n; ++i) tmp[i] = ee[i] - php1[i];
                const double enorm = wrms_norm(tmp.data(), ewt.data(), suppressalg);
                err_kp1 = enorm / (double)(kk + 2);
                const double terr_k = (double)(kk + 1) * err_k;
                const double terr_kp1 = (double)(kk + 2) * err_kp1;
                CNT(n, 3, 1, 0, 0);  // incl. 0.5 * terr_k (kk == 1) or kk * err_km1 (kk > 1)
                if (kk == 1) {
                    action = (terr_kp1 >= 0.5 * terr_k) ? MAINTAIN : RAISE;
                } else {
                    const double terr_km1 = (double)kk * err_km1;
                    if (terr_km1 <= std::fmin(terr_k, terr_kp1)) action = LOWER;
                    else if (terr_kp1 >= terr_k) action = MAINTAIN;
                    else action = RAISE;
                }
            }
            double err_knew;
            if (action == RAISE) { kk += 1; err_knew = err_kp1; }
            else if (action == LOWER) { kk -= 1; err_knew = err_km1; }
            else err_knew = err_k;

            double hnew = hh;
            {
                const double base = 2.0 * err_knew + 0.0001;
                const double arg = -(1.0 / (double)(kk + 1));
                rr = g_pow(base, arg);
                CNT(1, 1, 1, 0, 1);
            }
            if (rr >= 2.0) {
                hnew = 2.0 * hh;
                const double t = std::fabs(hnew) * hmax_inv;
                if (t > 1.0) hnew /= t;
                CNT(0, 2, 0, 0, 0);
            } else if (rr <= 1.0) {
                rr = std::fmax(0.5, std::fmin(rr, 0.9));
                hnew = hh * rr;
                CNT(0, 1, 0, 0, 0);
            }
            hh = hnew;
        }
        if (kused < maxord) std::copy(ee.begin(), ee.end(), PHI(kused + 1));
        for (int i = 0; i < n; ++i) tmp[i] = ee[i];
        CNT((kused + 1) * n, 0, 0, 0, 0);
        for (int j = kused; j >= 0; --j) {
            double* ph = PHI(j);
            for (int i = 0; i < n; ++i) {
                tmp[i] += ph[i];
                ph[i] = tmp[i];
            }
        }
    }

    // step -- src/lib.rs:613-711 ; returns 0 or fatal Status
    int step() {
        const double saved_t = tn;
        if (nst == 0) {
            kk = 1;
            kused = 0;
            hused = 0.0;
            psi[0] = hh;
            cj = 1.0 / hh;
            phase = 0;
            ns = 0;
        }
        long ncf = 0, nef = 0;
        double ck, err_k = 0.0, err_km1 = 0.0;
        for (;;) {
            if (trace_d) dump_frame();  // data_trace frame: state at the top of the attempt (src/lib.rs:635-639)
            nattempts += 1;
            uint32_t evw = 0;
            const long ev_nje0 = nje, ev_nni0 = nls.niters;
            if (ev) evw = (uint32_t)kk;
            ck = set_coeffs();
            if (ev) evw |= (uint32_t)(kk + 1 >= ns) << 3 | (uint32_t)std::min(ns, 7) << 4;
            tn += hh;
            CNT(1, 0, 0, 0, 0);
            predict();
            int nflag = nonlinear_solve();
            bool converged = false;
            err_k = 0.0;
            err_km1 = 0.0;
            if (nflag == 0) {
                converged = test_error(ck, err_k, err_km1);
                if (ev) evw |= (uint32_t)(1 + (kk > 1) + (kk > 2)) << 13;
                if (!converged) nflag = 2;
            }
            if (ev) evw |= (uint32_t)std::min(nje - ev_nje0, 3L) << 7 | (uint32_t)std::min(nls.niters - ev_nni0, 15L) << 9;
            if (converged) {
                ev_word = evw | 1u << 15;
                break;
            }
            restore(saved_t);
            if (nflag < 0) {  // singular Jacobian: fatal (reference panics)
                ncfn += 1;
                return nflag;
            }
            const int kflag = handle_n_flag(nflag, err_k, err_km1, ncf, nef);
            if (ev) {
                evw |= (uint32_t)(nflag == 2 && nef == 1) << 16 | (uint32_t)kk << 18 | (uint32_t)phase << 21;
                if (ev_count < ev_max) {
                    ev[ev_count] = evw;
                    ev_h[ev_count] = (float)std::fabs(hh);
                    if (ev_c) ev_c[ev_count] = ev_cjr();
                }
                ev_count += 1;
            }
            if (kflag != 0) return kflag;
            if (nst == 0) reset();
        }
        const int ev_ph0 = phase, ev_kk0 = kk, ev_ns = ns, ev_kdiff = kk - kused;
        complete_step(err_k, err_km1);
        for (int i = 0; i < n; ++i) ee[i] *= ck;
        CNT(0, n, 0, 0, 0);
        if (ev) {
            // which branches complete_step took (restated from its conditions on the state before the call)
            const bool ph1 = ev_ph0 == 1 || knew == ev_kk0 - 1 || ev_kk0 == maxord;
            const bool kp1 = ph1 && !(knew == ev_kk0 - 1) && ev_kk0 != maxord && !(ev_kk0 + 1 >= ev_ns || ev_kdiff == 1);
            ev_word |= (uint32_t)ph1 << 16 | (uint32_t)kp1 << 17 | (uint32_t)kk << 18 | (uint32_t)phase << 21;
            if (ev_count < ev_max) {
                ev[ev_count] = ev_word;
                ev_h[ev_count] = (float)std::fabs(hh);
                if (ev_c) ev_c[ev_count] = ev_cjr();
            }
            ev_count += 1;
        }
        return 0;
    }

    // ---- rootfinding: src/impl_r_check.rs:32-576, state src/lib.rs:198-229 (enabled by enable_roots) -------------
    int nrtfn = 0;
    std::vector<double> glo, ghi, grout, iroots;  // iroots is a Scalar array in the reference (src/lib.rs:200)
    std::vector<char> gactive;                     // initialised FALSE like the reference (src/lib.rs:373; C: TRUE)
    double tlo = 0, thi = 0, trout = 0, ttol = 0, toutc = 0;
    bool irfnd = false;
    long nge = 0;
    void enable_roots() {
        nrtfn = prob.num_roots();
        glo.assign(nrtfn, 0.0);
        ghi.assign(nrtfn, 0.0);
        grout.assign(nrtfn, 0.0);
        iroots.assign(nrtfn, 0.0);
        gactive.assign(nrtfn, 0);
    }
    static double signum(double x) { return (x != x) ? x : (std::signbit(x) ? -1.0 : 1.0); }  // f64::signum

    // IDARcheck1 -- src/impl_r_check.rs:32-105
    void r_check1() {
        for (int i = 0; i < nrtfn; ++i) iroots[i] = 0.0;
        tlo = tn;
        ttol = (std::fabs(tn) + std::fabs(hh)) * DBL_EPS * 100.0;
        prob.root(tlo, PHI(0), PHI(1), glo.data());
        nge = 1;
        bool zroot = false;
        for (int i = 0; i < nrtfn; ++i) {
            if (std::fabs(glo[i]) == 0.0) {
                gactive[i] = 0;
                zroot = true;
            }
        }
        if (zroot) {
            const double hratio = std::max(ttol / std::fabs(hh), 0.1);
            const double smallh = hratio * hh;
            const double tplus = tlo + smallh;
            for (int i = 0; i < n; ++i) yy[i] = PHI(0)[i];
            for (int i = 0; i < n; ++i) yy[i] += smallh * PHI(1)[i];  // scaled_add
            prob.root(tplus, yy.data(), PHI(1), ghi.data());
            nge += 1;
            for (int i = 0; i < nrtfn; ++i) {
                if (!gactive[i] && std::fabs(ghi[i]) != 0.0) {
                    gactive[i] = 1;
                    glo[i] = ghi[i];
                }
            }
        }
    }

    // IDARcheck2 -- src/impl_r_check.rs:123-209 ; returns 0 continue, 1 root found, ST_CLOSE_ROOTS
    int r_check2() {
        if (!irfnd) return 0;
        get_solution(tlo);
        prob.root(tlo, yy.data(), yp.data(), glo.data());
        nge += 1;
        for (int i = 0; i < nrtfn; ++i) iroots[i] = 0.0;
        bool zroot = false;
        for (int i = 0; i < nrtfn; ++i) {
            if (gactive[i] && std::fabs(glo[i]) == 0.0) {
                zroot = true;
                iroots[i] = 1.0;
            }
        }
        if (zroot) {
            ttol = (std::fabs(tn) + std::fabs(hh)) * DBL_EPS * 100.0;
            const double smallh = ttol * signum(hh);
            const double tplus = tlo + smallh;
            if ((tplus - tn) * hh >= 0.0) {
                const double hratio = smallh / hh;
                for (int i = 0; i < n; ++i) yy[i] += hratio * PHI(1)[i];
            } else {
                get_solution(tplus);
            }
            prob.root(tplus, yy.data(), yp.data(), ghi.data());
            nge += 1;
            bool z2 = false;
            for (int i = 0; i < nrtfn; ++i) {
                if (gactive[i]) {
                    if (std::fabs(ghi[i]) == 0.0) {
                        if (iroots[i] > 0.0) return ST_CLOSE_ROOTS;  // FoldWhile::Done
                        z2 = true;
                        iroots[i] = 1.0;
                    } else {
                        if (iroots[i] > 0.0) glo[i] = ghi[i];
                    }
                }
            }
            if (z2) return 1;
        }
        return 0;
    }

    // one pass over the g_i of root_find: src/impl_r_check.rs:332-369 and :457-492
    void scan_roots(const std::vector<double>& gx, bool first, int imax_in, bool& zroot, bool& sgnchg, int& imax) const {
        zroot = false;
        sgnchg = false;
        double maxfrac = 0.0;
        imax = first ? 0 : imax_in;
        for (int i = 0; i < nrtfn; ++i) {
            if (!gactive[i]) continue;
            const bool rootdir_glo_neg = 0.0 * glo[i] <= 0.0;  // rootdir is all zeros (src/lib.rs:368): 0*glo <= 0
            if (first) {
                if (std::fabs(gx[i]) == 0.0) {
                    if (rootdir_glo_neg) zroot = true;
                } else if (glo[i] * gx[i] < 0.0 && rootdir_glo_neg) {
                    const double gfrac = std::fabs(gx[i] / (gx[i] - glo[i]));
                    if (gfrac > maxfrac) {
                        sgnchg = true;
                        maxfrac = gfrac;
                        imax = i;
                    }
                }
            } else {
                if (std::fabs(gx[i]) == 0.0 && rootdir_glo_neg) {
                    zroot = true;
                } else if (glo[i] * gx[i] < 0.0 && rootdir_glo_neg) {
                    const double gfrac = std::fabs(gx[i] / (gx[i] - glo[i]));
                    if (gfrac > maxfrac) {
                        sgnchg = true;
                        maxfrac = gfrac;
                        imax = i;
                    }
                }
            }
        }
    }

    // IDARootfind (Illinois algorithm) -- src/impl_r_check.rs:330-576 ; returns 0 continue, 1 root found
    int root_find() {
        bool zroot, sgnchg;
        int imax;
        scan_roots(ghi, true, 0, zroot, sgnchg, imax);
        if (!sgnchg) {
            trout = thi;
            grout = ghi;
            if (!zroot) return 0;
            for (int i = 0; i < nrtfn; ++i) {
                iroots[i] = 0.0;
                if (gactive[i]) {
                    const bool rootdir_glo_neg = 0.0 * glo[i] <= 0.0;
                    if (std::fabs(ghi[i]) == 0.0 && rootdir_glo_neg) iroots[i] = signum(glo[i]);
                }
            }
            return 1;
        }
        double alph = 1.0;
        int side = 0, sideprev = -1;
        for (;;) {
            if (std::fabs(thi - tlo) <= ttol) break;
            alph = (sideprev == side) ? ((side == 2) ? alph * 2.0 : alph * 0.5) : 1.0;
            double tmid = thi - (thi - tlo) * ghi[imax] / (ghi[imax] - alph * glo[imax]);
            if (std::fabs(tmid - tlo) < 0.5 * ttol) {
                const double fracint = std::fabs(thi - tlo) / ttol;
                const double fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
                tmid = tlo + fracsub * (thi - tlo);
            }
            if (std::fabs(thi - tmid) < 0.5 * ttol) {
                const double fracint = std::fabs(thi - tlo) / ttol;
                const double fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
                tmid = thi - fracsub * (thi - tlo);
            }
            get_solution(tmid);
            prob.root(tmid, yy.data(), yp.data(), grout.data());
            nge += 1;
            sideprev = side;
            int imax2;
            scan_roots(grout, false, imax, zroot, sgnchg, imax2);
            imax = imax2;
            if (sgnchg) {
                thi = tmid;
                ghi = grout;
                side = 1;
                if (std::fabs(thi - tlo) <= ttol) break;
                continue;
            }
            if (zroot) {
                thi = tmid;
                ghi = grout;
                break;
            }
            tlo = tmid;
            glo = grout;
            side = 2;
            if (std::fabs(thi - tlo) <= ttol) break;
        }
        trout = thi;
        grout = ghi;
        for (int i = 0; i < nrtfn; ++i) {
            iroots[i] = 0.0;
            if (gactive[i]) {
                const bool rootdir_glo_neg = 0.0 * glo[i] <= 0.0;
                if (rootdir_glo_neg && (std::fabs(ghi[i]) == 0.0 || glo[i] * ghi[i] < 0.0)) iroots[i] = signum(glo[i]);
            }
        }
        return 1;
    }

    // IDARcheck3 -- src/impl_r_check.rs:211-267 ; returns 0 continue, 1 root found
    int r_check3() {
        thi = ((toutc - tn) * hh >= 0.0) ? tn : toutc;  // IdaTask::Normal
        get_solution(thi);
        prob.root(thi, yy.data(), yp.data(), ghi.data());
        nge += 1;
        ttol = (std::fabs(tn) + std::fabs(hh)) * DBL_EPS * 100.0;
        const int ier = root_find();
        for (int i = 0; i < nrtfn; ++i)
            if (!gactive[i] && grout[i] != 0.0) gactive[i] = 1;
        tlo = trout;
        glo = grout;
        if (ier == 1) get_solution(trout);
        return ier;
    }

    // get_solution -- src/lib.rs:1274-1343 ; returns 0 or ST_BAD_T
    int get_solution(double t) {
        const double sgn = (hh > 0.0 || (hh == 0.0 && !std::signbit(hh))) ? 1.0 : -1.0;  // f64::signum
        const double tfuzz = 100.0 * DBL_EPS * (std::fabs(tn) + std::fabs(hh)) * sgn;
        const double tp = tn - hused - tfuzz;
        CNT(5, 4, 0, 0, 0);  // tfuzz, tp, (t - tp) * hh, delt
        if ((t - tp) * hh < 0.0) return ST_BAD_T;
        const int kord = (kused == 0) ? 1 : kused;
        const double delt = t - tn;
        double c = 1.0, d = 0.0;
        double gam = delt / psi[0];
        CNT(2 * kord + (2 * kord + 1) * n, 2 * kord + (2 * kord + 1) * n, 1 + 2 * kord, 0, 0);  // recurrence + both sums
        double cvals[MXORDP1], dvals[MXORDP1];
        cvals[0] = c;
        for (int j = 1; j <= kord; ++j) {
            d = d * gam + c / psi[j - 1];
            c = c * gam;
            gam = (delt + psi[j - 1]) / psi[j];
            cvals[j] = c;
            dvals[j - 1] = d;
        }
        for (int i = 0; i < n; ++i) yy[i] = 0.0;
        for (int j = 0; j <= kord; ++j) {
            const double* ph = PHI(j);
            for (int i = 0; i < n; ++i) yy[i] += cvals[j] * ph[i];
        }
        for (int i = 0; i < n; ++i) yp[i] = 0.0;
        for (int j = 1; j <= kord; ++j) {
            const double* ph = PHI(j);
            for (int i = 0; i < n; ++i) yp[i] += dvals[j - 1] * ph[i];
        }
        return 0;
    }

    // get_dky -- src/lib.rs:424-529 ; returns 0, ST_BAD_K or ST_BAD_T.  k-th derivative of the interpolating polynomial
    // at t, from phi / psi of the last step.  Followed as written, including where it departs from SUNDIALS' IDAGetDky:
    // the inner recurrences run to j = kused - k + 1 for EVERY i (SUNDIALS: kused - k + i), so for k >= 2 the higher
    // c_j^(k) are never formed and stay 0 (src/lib.rs:496-506); psij_1 is not reset between passes.  The one place
    // the reference cannot be followed: k = 0 with kused = 5 indexes cjk[6] of a 6-array and panics (:497); here the
    // recurrence stops at j = 5 -- entry 6 would not be used by the sum (:512-526) anyway.
    // Sum: ndarray sum_axis(Axis(0)) over rows j = k..=kused of phi_j * c_j, folded from zeros in row order.
    int get_dky(double t, int k, double* dky) const {
        if (k < 0 || k > kused) return ST_BAD_K;
        const double sgn = (hh > 0.0 || (hh == 0.0 && !std::signbit(hh))) ? 1.0 : -1.0;  // f64::signum
        const double tfuzz = 100.0 * DBL_EPS * (std::fabs(tn) + std::fabs(hh)) * sgn;
        const double tp = tn - hused - tfuzz;
        if ((t - tp) * hh < 0.0) return ST_BAD_T;
        double cjk[MXORDP1] = {0, 0, 0, 0, 0, 0}, cjk_1[MXORDP1] = {0, 0, 0, 0, 0, 0};
        const double delt = t - tn;
        double psij_1 = 0.0;
        const int jtop = std::min(kused - k + 1, MXORDP1 - 1);
        for (int i = 0; i <= k; ++i) {
            const double si = (double)i;
            if (i == 0) {
                cjk[i] = 1.0;
            } else {
                cjk[i] = cjk[i - 1] * si / psi[i - 1];
                psij_1 = psi[i - 1];
            }
            for (int j = i + 1; j <= jtop; ++j) {
                cjk[j] = (si * cjk_1[j - 1] + cjk[j - 1] * (delt + psij_1)) / psi[j - 1];
                psij_1 = psi[j - 1];
            }
            for (int j = i + 1; j <= jtop; ++j) cjk_1[j] = cjk[j];
        }
        for (int i = 0; i < n; ++i) dky[i] = 0.0;
        for (int j = k; j <= kused; ++j) {
            const double* ph = &phi[(size_t)j * n];
            for (int i = 0; i < n; ++i) dky[i] = dky[i] + ph[i] * cjk[j];
        }
        return 0;
    }

    // solve (IdaTask::Normal, optional rootfinding, no tstop) -- src/impl_solve.rs:69-376 with
    // stop_test1/2 NORMAL branches src/impl_stop_test.rs:53-65,152-161
    int solve(double tout, double& tret) {
        toutc = tout;  // src/impl_solve.rs:78-82 (IdaTask::Normal)
        if (nst == 0) {
            if (!setup_done) {
                ewt_set(PHI(0));  // initial_setup src/lib.rs:537-545
                setup_done = true;
            }
            const double tdist = std::fabs(tout - tn);
            if (tdist == 0.0) return ST_ILL_INPUT;
            const double troundoff = 2.0 * DBL_EPS * (std::fabs(tn) + std::fabs(tout));
            if (tdist < troundoff) return ST_ILL_INPUT;
            hh = hin;
            if (hh != 0.0 && (tout - tn) * hh < 0.0) return ST_ILL_INPUT;
            if (hh == 0.0) {
                hh = 0.001 * tdist;
                const double ypnorm = wrms_norm(PHI(1), ewt.data(), suppressalg);
                CNT(2, 5 + n, 1, 0, 0);  // tdist, troundoff, 0.001 tdist, 2/hh, rh, phi[1] *= hh, toldel
                if (ypnorm > 2.0 / hh) {
                    hh = 0.5 / ypnorm;
                    CNT(0, 0, 1, 0, 0);
                }
                if (tout < tn) hh = -hh;
            }
            const double rh = std::fabs(hh) * hmax_inv;
            if (rh > 1.0) hh /= rh;
            h0u = hh;
            kk = 0;
            kused = 0;
            if (nrtfn > 0) r_check1();  // src/impl_solve.rs:161-164, before phi[1] is scaled
            {
                double* p1 = PHI(1);
                for (int i = 0; i < n; ++i) p1[i] *= hh;
            }
            eps_newt = epcon;
            toldel = 0.0001 * eps_newt;
        }
        long nstloc = 0;
        if (nst > 0) {
            // root stop checks: src/impl_solve.rs:186-227
            if (nrtfn > 0) {
                const int ier2 = r_check2();
                if (ier2 < 0) return ier2;
                if (ier2 == 1) {
                    tretlast = tlo;
                    tret = tlo;
                    return ST_ROOT;
                }
                const double troundoff = (std::fabs(tn) + std::fabs(hh)) * DBL_EPS * 100.0;
                if (std::fabs(tn - tretlast) > troundoff) {
                    if (r_check3() == 1) {
                        irfnd = true;
                        tretlast = tlo;
                        tret = tlo;
                        return ST_ROOT;
                    }
                    irfnd = false;
                }
            }
            // stop_test1, NORMAL
            if (tout == tretlast) {
                tretlast = tout;
                tret = tout;
                return ST_SUCCESS;
            }
            if ((tn - tout) * hh >= 0.0) {
                const int ier = get_solution(tout);
                if (ier != 0) return ier;
                tretlast = tout;
                tret = tout;
                return ST_SUCCESS;
            }
        }
        for (;;) {
            if (mxstep > 0 && nstloc >= mxstep) {
                tret = tn;
                tretlast = tn;
                return ST_TOO_MUCH_WORK;
            }
            if (nst > 0) {
                ewt_set(PHI(0));
                bool bad = false;
                for (int i = 0; i < n; ++i) bad = bad || (ewt[i] <= 0.0);
                if (bad) {
                    get_solution(tn);
                    tret = tn;
                    tretlast = tn;
                    return ST_BAD_EWT;
                }
            }
            const double nrm = wrms_norm(PHI(0), ewt.data(), suppressalg);
            tolsf = DBL_EPS * nrm;
            CNT(1, 2, 0, 0, 0);  // tolsf and the stop test (tn - tout) * hh after the step
            if (tolsf > 1.0) {
                tolsf *= 10.0;
                tret = tn;
                tretlast = tn;
                if (nst > 0) get_solution(tn);
                return ST_TOO_MUCH_ACC;
            }
            const int sflag = step();
            if (sflag != 0) {
                if (get_solution(tn) == 0) {
                    tret = tn;
                    tretlast = tn;
                }
                return sflag;
            }
            nstloc += 1;
            // root in the last step taken: src/impl_solve.rs:335-344
            if (nrtfn > 0 && r_check3() == 1) {
                irfnd = true;
                tretlast = tlo;
                tret = tlo;
                return ST_ROOT;
            }
            // stop_test2, NORMAL
            if ((tn - tout) * hh >= 0.0) {
                tret = tout;
                tretlast = tout;
                get_solution(tout);
                if (ev && ev_count >= 1 && ev_count <= ev_max) ev[ev_count - 1] |= 1u << 22;
                return ST_SUCCESS;
            }
        }
    }
};

}  // namespace

// =================================================================================================
// C interface for ctypes (tests, bench cpu_baseline).  Everything here is test infrastructure.
// =================================================================================================
// ida_id / ida_suppressalg applied by the ensemble entry points to every instance they integrate (process-global, like
// oracle_set_pow; oracle_set_id(nullptr, 0, 0) switches it off again)
static std::vector<unsigned char> g_id;
static bool g_suppressalg = false;
static void apply_id(Ida& ida) {
    if (!g_id.empty()) ida.id = g_id;
    ida.suppressalg = g_suppressalg;
}

extern "C" {
void oracle_set_id(const unsigned char* id, int n, int suppressalg) {
    g_id.assign(id ? id : nullptr, id ? id + n : nullptr);
    g_suppressalg = suppressalg != 0;
}

void oracle_set_pow(pow_fn_t f) { g_pow = f ? f : static_cast<double (*)(double, double)>(std::pow); }
double oracle_pow(double x, double y) { return g_pow(x, y); }

int oracle_getrf(double* a_colmajor, int64_t* pivot, int n) { return dense_get_rf(a_colmajor, pivot, n); }
void oracle_getrs(const double* a_colmajor, const int64_t* pivot, double* b, int n) {
    dense_get_rs(a_colmajor, pivot, b, n);
}
double oracle_wrms(const double* x, const double* w, int n) { return norm_wrms(x, w, n); }
double oracle_wrms_masked(const double* x, const double* w, const unsigned char* id, int n) { return norm_wrms_masked(x, w, id, n); }

void oracle_res(int kind, int n, const double* params, double t, const double* yy, const double* yp, double* r) {
    Problem p{kind, n, params};
    p.res(t, yy, yp, r);
}
void oracle_jac(int kind, int n, const double* params, double cj, const double* yy, const double* yp,
                double* J_colmajor) {
    Problem p{kind, n, params};
    std::fill(J_colmajor, J_colmajor + (size_t)n * n, 0.0);
    p.jac(0.0, cj, yy, yp, nullptr, J_colmajor);
}

// ---- unit hooks used by the golden-vector tests (state poked like src/tests/*.rs do) ----------
struct OracleUnit {
    Ida ida;
};

static double kZeros[512];
void* oracle_unit_new(int n) {
    auto* u = new OracleUnit();
    static const double par[3] = {0.04, 1.0e4, 3.0e7};
    Problem p{PK_ROBERTS, n, par};
    std::vector<double> atol(n, 1e-4);
    u->ida.init(p, kZeros, kZeros, 1e-4, atol.data());
    return u;
}
void oracle_unit_free(void* h) { delete static_cast<OracleUnit*>(h); }

// field access by name: doubles
static double* unit_dptr(Ida& s, const char* f, int* len) {
    const int n = s.n;
#define FV(name, vec) if (!std::strcmp(f, name)) { *len = (int)(vec).size(); return (vec).data(); }
#define FA(name, arr) if (!std::strcmp(f, name)) { *len = MXORDP1; return (arr); }
#define FS(name, sc) if (!std::strcmp(f, name)) { *len = 1; return &(sc); }
    FV("phi", s.phi) FA("psi", s.psi) FA("alpha", s.alpha) FA("beta", s.beta) FA("sigma", s.sigma)
    FA("gamma", s.gamma) FV("ee", s.ee) FV("ewt", s.ewt) FV("delta", s.delta) FV("yy", s.yy) FV("yp", s.yp)
    FV("yypredict", s.yypredict) FV("yppredict", s.yppredict)
    FS("hh", s.hh) FS("hused", s.hused) FS("rr", s.rr) FS("hmax_inv", s.hmax_inv) FS("cj", s.cj)
    FS("cjlast", s.cjlast) FS("cjold", s.cjold) FS("cjratio", s.cjratio) FS("ss", s.ss) FS("tn", s.tn)
    FS("eps_newt", s.eps_newt) FS("toldel", s.toldel)
#undef FV
#undef FA
#undef FS
    (void)n;
    return nullptr;
}
int oracle_unit_set_d(void* h, const char* field, const double* v, int len) {
    int l = 0;
    double* p = unit_dptr(static_cast<OracleUnit*>(h)->ida, field, &l);
    if (!p || l != len) return -1;
    std::copy(v, v + len, p);
    return 0;
}
int oracle_unit_get_d(void* h, const char* field, double* v, int len) {
    int l = 0;
    double* p = unit_dptr(static_cast<OracleUnit*>(h)->ida, field, &l);
    if (!p || l != len) return -1;
    std::copy(p, p + len, v);
    return 0;
}
static long* unit_iptr(Ida& s, const char* f, int* which) {
    *which = 0;
#define FI(name, v) if (!std::strcmp(f, name)) { *which = 1; static long t; t = (long)(v); return &t; }
    FI("kk", s.kk) FI("kused", s.kused) FI("knew", s.knew) FI("phase", s.phase) FI("ns", s.ns)
    FI("nst", s.nst) FI("maxord", s.maxord)
#undef FI
    return nullptr;
}
int oracle_unit_set_i(void* h, const char* f, long v) {
    Ida& s = static_cast<OracleUnit*>(h)->ida;
    if (!std::strcmp(f, "kk")) s.kk = (int)v;
    else if (!std::strcmp(f, "kused")) s.kused = (int)v;
    else if (!std::strcmp(f, "knew")) s.knew = (int)v;
    else if (!std::strcmp(f, "phase")) s.phase = (int)v;
    else if (!std::strcmp(f, "ns")) s.ns = (int)v;
    else if (!std::strcmp(f, "nst")) s.nst = v;
    else if (!std::strcmp(f, "maxord")) s.maxord = (int)v;
    else return -1;
    return 0;
}
long oracle_unit_get_i(void* h, const char* f) {
    int w;
    long* p = unit_iptr(static_cast<OracleUnit*>(h)->ida, f, &w);
    return p ? *p : -999999;
}
double oracle_unit_set_coeffs(void* h) { return static_cast<OracleUnit*>(h)->ida.set_coeffs(); }
void oracle_unit_predict(void* h) { static_cast<OracleUnit*>(h)->ida.predict(); }
void oracle_unit_restore(void* h, double saved_t) { static_cast<OracleUnit*>(h)->ida.restore(saved_t); }
int oracle_unit_test_error(void* h, double ck, double* err_k, double* err_km1) {
    return static_cast<OracleUnit*>(h)->ida.test_error(ck, *err_k, *err_km1) ? 1 : 0;
}
void oracle_unit_complete_step(void* h, double err_k, double err_km1) {
    static_cast<OracleUnit*>(h)->ida.complete_step(err_k, err_km1);
}
int oracle_unit_get_solution(void* h, double t) { return static_cast<OracleUnit*>(h)->ida.get_solution(t); }
int oracle_unit_get_dky(void* h, double t, int k, double* dky) { return static_cast<OracleUnit*>(h)->ida.get_dky(t, k, dky); }

// ---- Newton known-answer problem: the data of crates/nonlinear/src/newton.rs:192-343 -----------
// f1 = x^2+y^2+z^2-1, f2 = 2x^2+y^2-4z, f3 = 3x^2-4y+z^2 ; J analytic; ctest = (wrms(del) <= tol)
struct Quad3 {
    double a[9];
    int64_t piv[3];
    void sys(const double* v, double* f) {
        const double x = v[0], y = v[1], z = v[2];
        f[0] = x * x + y * y + z * z - 1.0;
        f[1] = 2.0 * (x * x) + y * y - 4.0 * z;
        f[2] = 3.0 * (x * x) - 4.0 * y + z * z;
    }
    int setup(const double* v, const double*, bool) {
        const double x = v[0], y = v[1], z = v[2];
        const double col0[3] = {2.0 * x, 4.0 * x, 6.0 * x}, col1[3] = {2.0 * y, 2.0 * y, -4.0},
                     col2[3] = {2.0 * z, -4.0, 2.0 * z};
        for (int i = 0; i < 3; ++i) { a[i] = col0[i]; a[3 + i] = col1[i]; a[6 + i] = col2[i]; }
        return -dense_get_rf(a, piv, 3);  // 0, or -(1-based singular column): LinearSetupFailed { LUFactFail { col } }
    }
    void lsolve(const double*, double* b) { dense_get_rs(a, piv, b, 3); }
    int ctest(int, const double*, const double* del, double tol, const double* w) {
        return norm_wrms(del, w, 3) <= tol ? 1 : 0;
    }
};
int oracle_newton_quad3(const double* y0, const double* w, double tol, int maxiters, double* y, long* niters,
                        long* nconvfails, int call_lsetup) {
    Newton nls;
    nls.maxiters = maxiters;
    nls.delta.assign(3, 0.0);
    Quad3 p;
    for (int i = 0; i < 9; ++i) p.a[i] = 0.0;  // Matrix3::zeros() (newton.rs:316)
    for (int i = 0; i < 3; ++i) p.piv[i] = 0;
    const int r = nls.solve(p, 3, y0, y, w, tol, call_lsetup != 0);
    *niters = nls.niters;
    *nconvfails = nls.nconvfails;
    return r;
}

// ---- ensemble driver: one Ida per instance, instances spread over nthreads --------------------
// Layouts are SoA: params [P][B], yy0/yp0/yy/yp [n][B].  touts: the ladder of output times, each
// one a separate solve() call (h0 depends on the first tout; mxstep is per call).
// yy_out/yp_out: final state [n][B] (may be null); yy_traj: [ntout][n][B] (may be null).
// counters [9][B]: nst, nre, nje, nni, netf, ncfn, nsetups, nls_nconvfails, nattempts.
// stepinfo_i [2][B]: kused, kk ; stepinfo_d [4][B]: hused, hh, tn, tret.
// status [ntout][B] per-call status (may be null); returns 0.
int oracle_ensemble_solve(int kind, int n, int64_t B, int nparams, const double* params, const double* yy0,
                          const double* yp0, double rtol, const double* atol, const double* touts, int ntout,
                          int nthreads, double* yy_out, double* yp_out, double* yy_traj, int64_t* counters,
                          int64_t* stepinfo_i, double* stepinfo_d, int32_t* status) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<double> par(nparams), y0(n), yp0v(n);
        for (;;) {
            const int64_t chunk = 16;
            const int64_t b0 = next.fetch_add(chunk);
            if (b0 >= B) break;
            const int64_t b1 = std::min(B, b0 + chunk);
            for (int64_t b = b0; b < b1; ++b) {
                for (int k = 0; k < nparams; ++k) par[k] = params[(size_t)k * B + b];
                for (int i = 0; i < n; ++i) {
                    y0[i] = yy0[(size_t)i * B + b];
                    yp0v[i] = yp0[(size_t)i * B + b];
                }
                Ida ida;
                Problem p{kind, n, par.data()};
                apply_id(ida);
                ida.init(p, y0.data(), yp0v.data(), rtol, atol);
                double tret = 0.0;
                for (int it = 0; it < ntout; ++it) {
                    const int st = ida.solve(touts[it], tret);
                    if (status) status[(size_t)it * B + b] = st;
                    if (yy_traj)
                        for (int i = 0; i < n; ++i) yy_traj[((size_t)it * n + i) * B + b] = ida.yy[i];
                }
                for (int i = 0; i < n; ++i) {
                    if (yy_out) yy_out[(size_t)i * B + b] = ida.yy[i];
                    if (yp_out) yp_out[(size_t)i * B + b] = ida.yp[i];
                }
                if (counters) {
                    const long c[9] = {ida.nst, ida.nre, ida.nje, ida.nls.niters, ida.netf,
                                       ida.ncfn, ida.nsetups, ida.nls.nconvfails, ida.nattempts};
                    for (int k = 0; k < 9; ++k) counters[(size_t)k * B + b] = c[k];
                }
                if (stepinfo_i) {
                    stepinfo_i[b] = ida.kused;
                    stepinfo_i[(size_t)B + b] = ida.kk;
                }
                if (stepinfo_d) {
                    stepinfo_d[b] = ida.hused;
                    stepinfo_d[(size_t)B + b] = ida.hh;
                    stepinfo_d[(size_t)2 * B + b] = ida.tn;
                    stepinfo_d[(size_t)3 * B + b] = tret;
                }
            }
        }
    };
    if (nthreads == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    return 0;
}

// Integrate every instance over the ladder, then evaluate get_dky(tq[q], kq[q]) for each query q on the final state:
// dky [nq][n][B] (rows of a failed query are left untouched), st [nq][B].
int oracle_ensemble_dky(int kind, int n, int64_t B, int nparams, const double* params, const double* yy0, const double* yp0,
                        double rtol, const double* atol, const double* touts, int ntout, int nthreads, int nq,
                        const double* tq, const int32_t* kq, double* dky, int32_t* st) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<double> par(nparams), y0(n), yp0v(n), out(n);
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= B) break;
            for (int k = 0; k < nparams; ++k) par[k] = params[(size_t)k * B + b];
            for (int i = 0; i < n; ++i) {
                y0[i] = yy0[(size_t)i * B + b];
                yp0v[i] = yp0[(size_t)i * B + b];
            }
            Ida ida;
            Problem p{kind, n, par.data()};
            apply_id(ida);
            ida.init(p, y0.data(), yp0v.data(), rtol, atol);
            double tret = 0.0;
            for (int it = 0; it < ntout; ++it) ida.solve(touts[it], tret);
            for (int q = 0; q < nq; ++q) {
                const int r = ida.get_dky(tq[q], kq[q], out.data());
                st[(size_t)q * B + b] = r;
                if (r == 0)
                    for (int i = 0; i < n; ++i) dky[((size_t)q * n + i) * B + b] = out[i];
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return 0;
}

// per-attempt event words of every instance (see Ida::ev): ev [B][max_att] uint32, ev_h [B][max_att] float, natt [B].
int oracle_ensemble_events2(int kind, int n, int64_t B, int nparams, const double* params, const double* yy0, const double* yp0,
                            double rtol, const double* atol, const double* touts, int ntout, int nthreads, long max_att,
                            uint32_t* ev, float* ev_h, int32_t* natt, float* ev_c);
int oracle_ensemble_events(int kind, int n, int64_t B, int nparams, const double* params, const double* yy0, const double* yp0,
                           double rtol, const double* atol, const double* touts, int ntout, int nthreads, long max_att,
                           uint32_t* ev, float* ev_h, int32_t* natt) {
    return oracle_ensemble_events2(kind, n, B, nparams, params, yy0, yp0, rtol, atol, touts, ntout, nthreads, max_att, ev, ev_h, natt,
                                   nullptr);
}
// the same with ev_c [B][max_att]: the cjratio the NEXT attempt will see (optional, may be null)
int oracle_ensemble_events2(int kind, int n, int64_t B, int nparams, const double* params, const double* yy0, const double* yp0,
                            double rtol, const double* atol, const double* touts, int ntout, int nthreads, long max_att,
                            uint32_t* ev, float* ev_h, int32_t* natt, float* ev_c) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<double> par(nparams), y0(n), yp0v(n);
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= B) break;
            for (int k = 0; k < nparams; ++k) par[k] = params[(size_t)k * B + b];
            for (int i = 0; i < n; ++i) {
                y0[i] = yy0[(size_t)i * B + b];
                yp0v[i] = yp0[(size_t)i * B + b];
            }
            Ida ida;
            Problem p{kind, n, par.data()};
            apply_id(ida);
            ida.init(p, y0.data(), yp0v.data(), rtol, atol);
            ida.ev = ev + (size_t)b * max_att;
            ida.ev_h = ev_h + (size_t)b * max_att;
            if (ev_c) ida.ev_c = ev_c + (size_t)b * max_att;
            ida.ev_max = max_att;
            double tret = 0.0;
            for (int it = 0; it < ntout; ++it) ida.solve(touts[it], tret);
            natt[b] = (int32_t)ida.ev_count;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return 0;
}

#ifdef ORACLE_COUNT_OPS
// algorithmic operation counts of every instance integrated over the ladder: ops [B][5] = add, mul, div, sqrt, pow;
// natt [B] step attempts
int oracle_ensemble_ops(int kind, int n, int64_t B, int nparams, const double* params, const double* yy0, const double* yp0,
                        double rtol, const double* atol, const double* touts, int ntout, int nthreads, int64_t* ops, int64_t* natt) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<double> par(nparams), y0(n), yp0v(n);
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= B) break;
            for (int k = 0; k < nparams; ++k) par[k] = params[(size_t)k * B + b];
            for (int i = 0; i < n; ++i) {
                y0[i] = yy0[(size_t)i * B + b];
                yp0v[i] = yp0[(size_t)i * B + b];
            }
            g_ops = OpCount();
            Ida ida;
            Problem p{kind, n, par.data()};
            apply_id(ida);
            ida.init(p, y0.data(), yp0v.data(), rtol, atol);
            double tret = 0.0;
            for (int it = 0; it < ntout; ++it) ida.solve(touts[it], tret);
            const long v[5] = {g_ops.add, g_ops.mul, g_ops.div, g_ops.sqrt_, g_ops.pow_};
            for (int k = 0; k < 5; ++k) ops[b * 5 + k] = v[k];
            natt[b] = ida.nattempts;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return 0;
}
#endif

// single-instance trace: records (nst, kused, hused, y[0..n)) after every solve() call, used to
// reproduce the table of examples/roberts.rs:87-113
int oracle_single_trace(int kind, int n, const double* params, const double* yy0, const double* yp0, double rtol,
                        const double* atol, const double* touts, int ntout, double* rows /*[ntout][3+n]*/,
                        int64_t* counters /*[9]*/) {
    Ida ida;
    Problem p{kind, n, params};
    apply_id(ida);
    ida.init(p, yy0, yp0, rtol, atol);
    double tret = 0.0;
    int last = 0;
    for (int it = 0; it < ntout; ++it) {
        last = ida.solve(touts[it], tret);
        double* r = rows + (size_t)it * (3 + n);
        r[0] = (double)ida.nst;
        r[1] = (double)ida.kused;
        r[2] = ida.hused;
        for (int i = 0; i < n; ++i) r[3 + i] = ida.yy[i];
        if (last != 0) break;
    }
    const long c[9] = {ida.nst, ida.nre, ida.nje, ida.nls.niters, ida.netf, ida.ncfn, ida.nsetups,
                       ida.nls.nconvfails, ida.nattempts};
    for (int k = 0; k < 9; ++k) counters[k] = c[k];
    return last;
}

// attempt trace of one instance over a ladder of output times: rows as documented at b200ida_set_trace.
// Returns the number of attempts seen (may exceed max_rows; only the first max_rows are stored).
long oracle_attempt_trace(int kind, int n, const double* params, const double* yy0, const double* yp0, double rtol,
                          const double* atol, const double* touts, int ntout, double* dtrace, int32_t* itrace, long max_rows,
                          int32_t* status /*[ntout]*/) {
    Ida ida;
    Problem p{kind, n, params};
    apply_id(ida);
    ida.init(p, yy0, yp0, rtol, atol);
    ida.trace_d = dtrace;
    ida.trace_i = itrace;
    ida.trace_max = max_rows;
    double tret = 0.0;
    for (int it = 0; it < ntout; ++it) {
        ida.trace_tret = tret;
        const int st = ida.solve(touts[it], tret);
        if (status) status[it] = st;
    }
    return ida.trace_count;
}

// One instance with rootfinding, driven like examples/roberts.rs:117-140: for every tout, call solve() again after each
// Root return.  Events: ev_t[k], ev_iroots[k][nr] (get_root_info), ev_y[k][n]; returns the number of events (may
// exceed max_ev; only the first max_ev are stored).  out: yy[n], counters[10] = the 9 of oracle_single_trace + nge.
long oracle_roots_run(int kind, int n, const double* params, const double* yy0, const double* yp0, double rtol, const double* atol,
                      const double* touts, int ntout, long max_ev, double* ev_t, double* ev_iroots, double* ev_y, double* yy_out,
                      int64_t* counters, int32_t* status) {
    Ida ida;
    Problem p{kind, n, params};
    apply_id(ida);
    ida.init(p, yy0, yp0, rtol, atol);
    ida.enable_roots();
    const int nr = ida.nrtfn;
    long nev = 0;
    double tret = 0.0;
    for (int it = 0; it < ntout; ++it) {
        int st;
        for (;;) {
            st = ida.solve(touts[it], tret);
            if (st != ST_ROOT) break;
            if (nev < max_ev) {
                ev_t[nev] = tret;
                for (int i = 0; i < nr; ++i) ev_iroots[nev * nr + i] = ida.iroots[i];
                for (int i = 0; i < n; ++i) ev_y[nev * n + i] = ida.yy[i];
            }
            nev += 1;
        }
        status[it] = st;
        if (st != 0) break;
    }
    for (int i = 0; i < n; ++i) yy_out[i] = ida.yy[i];
    const long c[10] = {ida.nst, ida.nre, ida.nje, ida.nls.niters, ida.netf, ida.ncfn, ida.nsetups, ida.nls.nconvfails, ida.nattempts,
                        ida.nge};
    for (int k = 0; k < 10; ++k) counters[k] = c[k];
    return nev;
}

}  // extern "C"
