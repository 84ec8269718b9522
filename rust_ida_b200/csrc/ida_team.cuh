// ida_team.cuh -- team-per-instance ensemble IDA for run-time n: one WARP per small system (n <= 32) or
// one CTA per medium system (n <= 256, e.g. the n = 100 heat problem).
//
// A team of T threads owns one DAE instance at a time.  The instance's working set -- phi[6][n], the
// n-vectors, the n x n iteration matrix / LU factors, pivots and all scalars -- is staged from the SoA
// arrays in HBM into the team's SHARED-memory arena once, the whole `Ida::solve` ladder runs against that
// arena, and the state is written back once.  In particular the LU factors stay resident in shared
// memory across all the Newton iterations and steps that reuse them (60 factorisations vs 537 solves for
// Roberts), instead of being re-read from HBM per solve.
//
// Work split inside a team:
//   * scalar control (set_coeffs, the order/step-size decisions of test_error / complete_step /
//     handle_n_flag, stop tests) is executed by the team LEADER against scalars in the arena; the other
//     lanes/warps wait at the next team barrier, so a CTA team issues that code from one warp only;
//   * vector work (predict, residual, Jacobian fill, LU, triangular solves, phi updates) is spread over the
//     team, component i -> thread i mod T;
//   * WRMS norms are summed by the leader in index order (src/norm_rms.rs:32-37 is a sequential fold; a
//     tree reduction would round differently and break bit-parity of the step-size sequence).
// All control flow is team-uniform, so there is no divergence inside a team; teams are independent.
//
// Arithmetic: operation-for-operation like the reference, -fmad=false, pow = detpow.h.
#pragma once
#include "detpow.h"
#include "fastdiv.cuh"
#include "ida_common.cuh"
#include "lu_team.cuh"
#include "lu_blocked.cuh"

#ifndef IDA_TEAM_LUB_POL
#define IDA_TEAM_LUB_POL 0  // lub_factor panel policy (lu_blocked.cuh), tools/build_variants.sh experiments
#endif

namespace b200ida {

// ---- team policies ------------------------------------------------------------------------------------
#ifndef IDA_WARP_MINB
#define IDA_WARP_MINB 3  // the warp-team kernel is compiled for 3 resident CTAs (of 4 teams) per SM = 168 registers: 1-2 (232) lose a third of the warps, 4 (128) spills
#endif
struct WarpTeam {
    static constexpr int T = 32;
    static constexpr int MAXT = 128, MINB = IDA_WARP_MINB;
    __device__ __forceinline__ static int tid() { return threadIdx.x & 31; }
    __device__ __forceinline__ static void sync() { __syncwarp(); }
    __device__ __forceinline__ static int teams_per_cta() { return blockDim.x >> 5; }
    __device__ __forceinline__ static int team_in_cta() { return threadIdx.x >> 5; }
};
struct CtaTeam {
    static constexpr int T = 0;  // run-time: blockDim.x
    static constexpr int MAXT = 256, MINB = 1;  // up to 255 registers
    __device__ __forceinline__ static int tid() { return threadIdx.x; }
    __device__ __forceinline__ static void sync() { __syncthreads(); }
    __device__ __forceinline__ static int teams_per_cta() { return 1; }
    __device__ __forceinline__ static int team_in_cta() { return 0; }
};
template <class Team>
__device__ __forceinline__ int team_size() {
    if constexpr (Team::T > 0) return Team::T;
    else return blockDim.x;
}

// ---- run-time problems (device side of ModelSpec/Residual/Jacobian, src/traits.rs:6-69) ---------------
struct TeamProblem {
    int kind;
    int n;
    int M;           // heat: grid side
    double p[4];     // per-instance parameters

    // F_i(t, yy, yp)
    __device__ __forceinline__ double res_i(int i, const double* yy, const double* yp) const {
        switch (kind) {
        case B200IDA_PROBLEM_ROBERTS: {  // src/sample_problems/roberts.rs:58-61
            const double r0 = -p[0] * yy[0] + p[1] * yy[1] * yy[2];
            if (i == 0) return r0 - yp[0];
            if (i == 1) return -r0 - p[2] * yy[1] * yy[1] - yp[1];
            return yy[0] + yy[1] + yy[2] - 1.0;
        }
        case B200IDA_PROBLEM_LORENZ63: {  // tests/lorenz63.rs:47-52 as F = y' - f(y)
            if (i == 0) return yp[0] - p[0] * (yy[1] - yy[0]);
            if (i == 1) return yp[1] - (yy[0] * (p[1] - yy[2]) - yy[1]);
            return yp[2] - (yy[0] * yy[1] - p[2] * yy[2]);
        }
        default: {  // B200IDA_PROBLEM_HEAT2D (SUNDIALS idaHeat2D): boundary rows algebraic u = 0
            const int r = i / M, c = i % M;
            if (r == 0 || r == M - 1 || c == 0 || c == M - 1) return yy[i];
            const double dif1 = yy[i - 1] + yy[i + 1] - 2.0 * yy[i];
            const double dif2 = yy[i - M] + yy[i + M] - 2.0 * yy[i];
            return yp[i] - p[0] * (dif1 + dif2);
        }
        }
    }
    // J(i, j) = dF_i/dy_j + cj dF_i/dy'_j for the whole row i (J already zeroed), column-major, leading dimension ld
    __device__ __forceinline__ void jac_row(int i, double cj, const double* yy, double* J, int ld) const {
        auto A = [&](int r, int c) -> double& { return J[(size_t)c * ld + r]; };
        switch (kind) {
        case B200IDA_PROBLEM_ROBERTS:  // src/sample_problems/roberts.rs:80-90
            if (i == 0) {
                A(0, 0) = -p[0] - cj;
                A(0, 1) = p[1] * yy[2];
                A(0, 2) = p[1] * yy[1];
            } else if (i == 1) {
                A(1, 0) = p[0];
                A(1, 1) = -p[1] * yy[2] - (2.0 * p[2]) * yy[1] - cj;
                A(1, 2) = -p[1] * yy[1];
            } else {
                A(2, 0) = 1.0;
                A(2, 1) = 1.0;
                A(2, 2) = 1.0;
            }
            break;
        case B200IDA_PROBLEM_LORENZ63:
            if (i == 0) {
                A(0, 0) = cj + p[0];
                A(0, 1) = -p[0];
                A(0, 2) = 0.0;
            } else if (i == 1) {
                A(1, 0) = yy[2] - p[1];
                A(1, 1) = cj + 1.0;
                A(1, 2) = yy[0];
            } else {
                A(2, 0) = -yy[1];
                A(2, 1) = -yy[0];
                A(2, 2) = cj + p[2];
            }
            break;
        default: {
            const int r = i / M, c = i % M;
            if (r == 0 || r == M - 1 || c == 0 || c == M - 1) {
                A(i, i) = 1.0;
            } else {
                A(i, i) = cj + 4.0 * p[0];
                A(i, i - 1) = -p[0];
                A(i, i + 1) = -p[0];
                A(i, i - M) = -p[0];
                A(i, i + M) = -p[0];
            }
        }
        }
    }
};

// ---- the team's shared-memory arena ----------------------------------------------------------------------
struct TeamScalars {
    double psi[MXORDP1], alpha[MXORDP1], beta[MXORDP1], sigma[MXORDP1], gamma[MXORDP1];
    double hh, hused, rr, cj, cjold, cjlast, cjratio, ss, toldel, eps_newt, tn, tretlast, h0u, tolsf, tret;
    double oldnrm, ck, err_k, err_km1, saved_t, tout, nrm, scale;
    double en1, en2;  // test_error: the norms of the order k-1 / k-2 error estimates (summed next to enorm_k)
    int kk, kused, knew, phase, ns, setup_done, jcur;
    int nst, nre, nje, nni, netf, ncfn, nsetups, nlscf, natt;
    int flag, info, nstloc, ncf, nef, it, call_status, curiter;
    LuShared lu;
};

// leading dimension of the iteration matrix in the arena: odd, so that row accesses (pivot-row moves) are
// shared-memory bank-conflict free as well as column accesses
__host__ __device__ inline int team_ldj(int n) { return n | 1; }
__host__ __device__ inline size_t team_arena_bytes(int n) {
    // phi[6][n] + 11 n-vectors + J[ldj][n] (+1 pad) as doubles, the blocked-LU scratch (lu_blocked.cuh), pivots as ints,
    // scalars; 16-byte aligned pieces
    size_t d = (size_t)17 * n + (size_t)team_ldj(n) * n + 1;
    size_t bytes = (d * sizeof(double) + 15) / 16 * 16 + (lub_ub_bytes(n) + 15) / 16 * 16 + 2 * (((size_t)n * sizeof(int) + 15) / 16 * 16) +
                   (sizeof(LubCtl) + 15) / 16 * 16 + (sizeof(TeamScalars) + 15) / 16 * 16;
    return (bytes + 15) / 16 * 16;
}

template <class Team>
struct IdaTeam {
    int n, T, t, ldj;
    // arena views
    double *phi, *ee, *ewt, *yy, *yp, *yypred, *yppred, *delta, *nd, *tmp, *prod, *rdiag, *J, *Ub;
    int *piv, *perm;
    LubCtl* lub;
    TeamScalars* S;
    TeamProblem prob;
    const double* atol;  // global, [n]
    const double* idmask = nullptr;  // global, [n]: ida_id as 1.0 / 0.0 when ida_suppressalg is set, else null
    double rtol;
    TraceParams tr = {nullptr, nullptr, nullptr, 0, -1};  // b200ida_set_trace

    __device__ __forceinline__ void bind(unsigned char* arena, int n_) {
        n = n_;
        T = team_size<Team>();
        t = Team::tid();
        double* d = reinterpret_cast<double*>(arena);
        phi = d;            d += 6 * n;
        ee = d;             d += n;
        ewt = d;            d += n;
        yy = d;             d += n;
        yp = d;             d += n;
        yypred = d;         d += n;
        yppred = d;         d += n;
        delta = d;          d += n;
        nd = d;             d += n;
        tmp = d;            d += n;
        prod = d;           d += n;
        rdiag = d;          d += n;
        ldj = team_ldj(n);
        J = d;              d += (size_t)ldj * n + 1;
        unsigned char* q = arena + ((size_t)(d - reinterpret_cast<double*>(arena)) * sizeof(double) + 15) / 16 * 16;
        Ub = reinterpret_cast<double*>(q);
        q += (lub_ub_bytes(n) + 15) / 16 * 16;
        piv = reinterpret_cast<int*>(q);
        q += ((size_t)n * sizeof(int) + 15) / 16 * 16;
        perm = reinterpret_cast<int*>(q);
        q += ((size_t)n * sizeof(int) + 15) / 16 * 16;
        lub = reinterpret_cast<LubCtl*>(q);
        q += (sizeof(LubCtl) + 15) / 16 * 16;
        S = reinterpret_cast<TeamScalars*>(q);
    }
    __device__ __forceinline__ bool leader() const { return t == 0; }
    __device__ __forceinline__ double* PHI(int j) const { return phi + (size_t)j * n; }

    // ---- HBM <-> arena -------------------------------------------------------------------------------------
    __device__ void load(const double* __restrict__ ds, const int* __restrict__ is, int64_t B, int64_t b) {
        const Layout lay(n);
        for (int i = t; i < n; i += T) {
            for (int j = 0; j < MXORDP1; ++j) phi[j * n + i] = ds[(int64_t)lay.phi(j, i) * B + b];
            ee[i] = ds[(int64_t)lay.ee(i) * B + b];
            ewt[i] = ds[(int64_t)lay.ewt(i) * B + b];
            yy[i] = ds[(int64_t)lay.yy(i) * B + b];
            yp[i] = ds[(int64_t)lay.yp(i) * B + b];
            piv[i] = is[(int64_t)lay.piv(i) * B + b];
        }
        for (int e = t; e < n * n; e += T) J[(e / n) * ldj + e % n] = ds[(int64_t)lay.J(e % n, e / n) * B + b];
        Team::sync();
        for (int k = t; k < n; k += T) rdiag[k] = 1.0 / J[(size_t)k * ldj + k];  // not persisted; from the stored factors
        if (t == T - 1) lu_build_perm(n, piv, perm);                              // likewise
        if (leader()) {
            for (int j = 0; j < MXORDP1; ++j) {
                S->psi[j] = ds[(int64_t)lay.psi(j) * B + b];
                S->alpha[j] = ds[(int64_t)lay.alpha(j) * B + b];
                S->beta[j] = ds[(int64_t)lay.beta(j) * B + b];
                S->sigma[j] = ds[(int64_t)lay.sigma(j) * B + b];
                S->gamma[j] = ds[(int64_t)lay.gamma(j) * B + b];
            }
#define LD_S(f, F) S->f = ds[(int64_t)lay.scal(Layout::F) * B + b];
            LD_S(hh, HH) LD_S(hused, HUSED) LD_S(rr, RR) LD_S(cj, CJ) LD_S(cjold, CJOLD) LD_S(cjlast, CJLAST)
            LD_S(cjratio, CJRATIO) LD_S(ss, SS) LD_S(toldel, TOLDEL) LD_S(eps_newt, EPS_NEWT) LD_S(tn, TN)
            LD_S(tretlast, TRETLAST) LD_S(h0u, H0U) LD_S(tolsf, TOLSF) LD_S(tret, TRET)
#undef LD_S
#define LD_I(f, F) S->f = is[(int64_t)Layout::F * B + b];
            LD_I(kk, KK) LD_I(kused, KUSED) LD_I(knew, KNEW) LD_I(phase, PHASE) LD_I(ns, NS) LD_I(setup_done, SETUP_DONE)
            LD_I(jcur, JCUR) LD_I(nst, NST) LD_I(nre, NRE) LD_I(nje, NJE) LD_I(nni, NNI) LD_I(netf, NETF)
            LD_I(ncfn, NCFN) LD_I(nsetups, NSETUPS) LD_I(nlscf, NLSCF) LD_I(natt, NATT)
#undef LD_I
        }
        Team::sync();
    }
    __device__ void store(double* __restrict__ ds, int* __restrict__ is, int64_t B, int64_t b) {
        Team::sync();
        const Layout lay(n);
        for (int i = t; i < n; i += T) {
            for (int j = 0; j < MXORDP1; ++j) ds[(int64_t)lay.phi(j, i) * B + b] = phi[j * n + i];
            ds[(int64_t)lay.ee(i) * B + b] = ee[i];
            ds[(int64_t)lay.ewt(i) * B + b] = ewt[i];
            ds[(int64_t)lay.yy(i) * B + b] = yy[i];
            ds[(int64_t)lay.yp(i) * B + b] = yp[i];
            is[(int64_t)lay.piv(i) * B + b] = piv[i];
        }
        for (int e = t; e < n * n; e += T) ds[(int64_t)lay.J(e % n, e / n) * B + b] = J[(e / n) * ldj + e % n];
        if (leader()) {
            for (int j = 0; j < MXORDP1; ++j) {
                ds[(int64_t)lay.psi(j) * B + b] = S->psi[j];
                ds[(int64_t)lay.alpha(j) * B + b] = S->alpha[j];
                ds[(int64_t)lay.beta(j) * B + b] = S->beta[j];
                ds[(int64_t)lay.sigma(j) * B + b] = S->sigma[j];
                ds[(int64_t)lay.gamma(j) * B + b] = S->gamma[j];
            }
#define ST_S(f, F) ds[(int64_t)lay.scal(Layout::F) * B + b] = S->f;
            ST_S(hh, HH) ST_S(hused, HUSED) ST_S(rr, RR) ST_S(cj, CJ) ST_S(cjold, CJOLD) ST_S(cjlast, CJLAST)
            ST_S(cjratio, CJRATIO) ST_S(ss, SS) ST_S(toldel, TOLDEL) ST_S(eps_newt, EPS_NEWT) ST_S(tn, TN)
            ST_S(tretlast, TRETLAST) ST_S(h0u, H0U) ST_S(tolsf, TOLSF) ST_S(tret, TRET)
#undef ST_S
#define ST_I(f, F) is[(int64_t)Layout::F * B + b] = S->f;
            ST_I(kk, KK) ST_I(kused, KUSED) ST_I(knew, KNEW) ST_I(phase, PHASE) ST_I(ns, NS) ST_I(setup_done, SETUP_DONE)
            ST_I(jcur, JCUR) ST_I(nst, NST) ST_I(nre, NRE) ST_I(nje, NJE) ST_I(nni, NNI) ST_I(netf, NETF)
            ST_I(ncfn, NCFN) ST_I(nsetups, NSETUPS) ST_I(nlscf, NLSCF) ST_I(natt, NATT)
#undef ST_I
        }
        Team::sync();
    }

    // ---- WRMS norm (src/norm_rms.rs:31-38): products by the team, ordered fold by the leader -> S->nrm --------
    // Callers must have a barrier between writing x and calling this; it ends with a barrier.
    // masked = the call is a local-error norm (wrms_norm with mask = suppressalg, src/lib.rs:1353-1370): with
    // ida_suppressalg set the algebraic components are left out like norm_wrms_masked does (src/norm_rms.rs:49-57)
    __device__ void wrms_to(const double* x, double* out, bool masked = true) {
        const bool m = masked && idmask != nullptr;
        for (int i = t; i < n; i += T) {
            double p = x[i] * ewt[i];
            if (m) p = p * idmask[i];
            prod[i] = p * p;
        }
        Team::sync();
        if (leader()) {
            double acc = 0.0;
            for (int i = 0; i < n; ++i) acc = acc + prod[i];
            *out = sqrt(acc / (double)n);
        }
        Team::sync();
    }

    // ---- TolControlSV::ewt_set (src/tol_control.rs:71-82); flags S->flag if some ewt <= 0 --------------------
    __device__ void ewt_set(bool check) {
        if (check && leader()) S->flag = 0;
        Team::sync();
        for (int i = t; i < n; i += T) {
            const double w = 1.0 / (rtol * fabs(phi[i]) + atol[i]);
            ewt[i] = w;
            if (check && w <= 0.0) S->flag = 1;
        }
        Team::sync();
    }

    // ---- set_coeffs (src/lib.rs:722-782): leader does the scalar recurrence, team scales phi ------------------
    __device__ void set_coeffs() {
        if (leader()) {
            TeamScalars& s = *S;
            if (s.hh != s.hused || s.kk != s.kused) s.ns = 0;
            s.ns = min(s.ns + 1, s.kused + 2);
            if (s.kk + 1 >= s.ns) {
                // The recurrence of src/lib.rs:738-755 with its three divisions per order taken out of the serial loop
                // where the operands allow it: alpha_i = hh / (psi_old[i-1] + hh) and alpha_{i-1} / hh do not depend on the
                // previous order's results, so the (at most 5 + 5) quotients run as two batches of interleaved chains
                // (fastdiv.cuh: one leader lane, bound by dependent-issue latency); beta_i = beta_{i-1} psi_{i-1} / psi_old[i-1]
                // stays a chain.  Every quotient has the operands it had in the loop.
                s.beta[0] = 1.0;
                s.alpha[0] = 1.0;
                s.gamma[0] = 0.0;
                s.sigma[0] = 1.0;
                const int kk = s.kk;
                const double hh = s.hh;
                double t2[MXORDP1], tp[MXORDP1];  // psi_old[i-1] ; temp1 after order i (tp[0] = hh)
                tp[0] = hh;
#pragma unroll
                for (int i = 1; i < MXORDP1; ++i) {
                    t2[i] = (i <= kk) ? s.psi[i - 1] : 1.0;
                    tp[i] = t2[i] + hh;
                }
                double num[5], den[5], al[5], gq[5];
#pragma unroll
                for (int i = 1; i <= 5; ++i) {
                    num[i - 1] = (i <= kk) ? hh : 1.0;
                    den[i - 1] = (i <= kk) ? tp[i] : 1.0;
                }
                div_batch<5>(num, den, al);  // alpha[i] = hh / temp1
#pragma unroll
                for (int i = 1; i <= 5; ++i) {
                    num[i - 1] = (i == 1) ? 1.0 : ((i <= kk) ? al[i - 2] : 1.0);  // alpha[i-1]
                    den[i - 1] = (i <= kk) ? hh : 1.0;
                }
                div_batch<5>(num, den, gq);  // alpha[i-1] / hh
#pragma unroll
                for (int i = 1; i <= 5; ++i) {
                    if (i <= kk) {
                        s.psi[i - 1] = tp[i - 1];
                        s.beta[i] = s.beta[i - 1] * tp[i - 1] / t2[i];
                        s.alpha[i] = al[i - 1];
                        s.sigma[i] = (double)i * s.sigma[i - 1] * al[i - 1];
                        s.gamma[i] = s.gamma[i - 1] + gq[i - 1];
                    }
                }
                s.psi[kk] = pick6(tp, kk);
            }
            // alphas = -sum 1/(i+1): the quotients are compile-time constants (correctly rounded, as at run time)
            constexpr double RCPI[5] = {1.0 / 1.0, 1.0 / 2.0, 1.0 / 3.0, 1.0 / 4.0, 1.0 / 5.0};
            double alphas = 0.0, alpha0 = 0.0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (i < s.kk) {
                    alphas -= RCPI[i];
                    alpha0 -= s.alpha[i];
                }
            }
            s.cjlast = s.cj;
            s.cj = -alphas / s.hh;
            double ck = fabs(s.alpha[s.kk] + alphas - alpha0);
            s.ck = fmax(ck, s.alpha[s.kk]);
        }
        Team::sync();
        const int ns = S->ns, kk = S->kk;
        for (int j = ns; j <= kk; ++j) {
            const double bj = S->beta[j];
            for (int i = t; i < n; i += T) phi[j * n + i] *= bj;
        }
    }

    // ---- predict (src/lib.rs:894-959) ----------------------------------------------------------------------
    __device__ void predict() {
        const int kk = S->kk;
        for (int i = t; i < n; i += T) {
            double a = 0.0, b = 0.0;
            for (int j = 0; j <= kk; ++j) a += phi[j * n + i];
            for (int j = 1; j <= kk; ++j) b += S->gamma[j] * phi[j * n + i];
            yypred[i] = a;
            yppred[i] = b;
        }
    }

    // ---- IdaNLProblem::sys (src/ida_nls.rs:118-153): res -> nd ------------------------------------------------
    __device__ void nls_sys(const double* ycor) {
        const double cj = S->cj;
        for (int i = t; i < n; i += T) {
            yy[i] = yypred[i] + ycor[i];
            yp[i] = yppred[i] + cj * ycor[i];
        }
        Team::sync();
        for (int i = t; i < n; i += T) nd[i] = prob.res_i(i, yy, yp);
        if (leader()) S->nre += 1;
        Team::sync();
    }

    // ---- dense_get_rf on the arena's J (crates/linear/src/dense.rs:86-158); S->info = 0 or the 1-based singular
    //      column.  CTA teams with n > 32 use the blocked, look-ahead factorisation of lu_blocked.cuh (bit-identical to
    //      the unblocked one), warp teams and small systems the group LU of lu_team.cuh -------------------------------
    __device__ void getrf() {
        Team::sync();
        if (Team::T == 0 && n > 32) {
            lub_factor<5>(J, ldj, n, lub, piv, Ub, IDA_TEAM_LUB_POL);
            if (leader()) S->info = lub->info;
        } else {
            lu_factor<Team::T == 32>(J, ldj, n, &S->lu, piv);
            if (leader()) S->info = S->lu.info;
        }
        for (int k = t; k < n; k += T) rdiag[k] = 1.0 / J[(size_t)k * ldj + k];
        if (t == T - 1) lu_build_perm(n, piv, perm);
        Team::sync();
    }

    // ---- dense_get_rs on nd (crates/linear/src/dense.rs:165-206) + cjratio rescale (src/ida_ls.rs:406-410) ------
    __device__ void lsolve() {
        if (Team::T == 0 && n > 32) {
            Team::sync();
            lu_solve_cta_blocked(J, ldj, n, perm, nd, tmp, rdiag);
        } else {
            lu_solve<Team::T == 32>(J, ldj, n, piv, nd, rdiag, perm);
        }
        const double cjratio = S->cjratio;
        if (cjratio != 1.0) {
            const double f = 2.0 / (1.0 + cjratio);
            for (int i = t; i < n; i += T) nd[i] *= f;
            Team::sync();
        }
    }

    // ---- IdaNLProblem::setup (src/ida_nls.rs:156-187, src/ida_ls.rs:232-290) ---------------------------------
    __device__ void nls_setup() {
        for (int e = t; e < ldj * n; e += T) J[e] = 0.0;
        Team::sync();
        const double cj = S->cj;
        for (int i = t; i < n; i += T) prob.jac_row(i, cj, yy, J, ldj);
        Team::sync();
        getrf();
        if (leader()) {
            S->nsetups += 1;
            S->nje += 1;
            if (S->info == 0) {
                S->cjold = S->cj;
                S->cjratio = 1.0;
                S->ss = 20.0;
            }
        }
        Team::sync();
    }

    // ---- nonlinear_solve + Newton::solve (src/lib.rs:787-890, newton.rs:51-167); result in S->flag ---------------
    // 0 ok, 1 recoverable convergence failure, B200IDA_ST_LSETUP_FAIL singular Jacobian
    __device__ void nonlinear_solve() {
        if (leader()) {
            TeamScalars& s = *S;
            bool call_lsetup = false;
            if (s.nst == 0) {
                s.cjold = s.cj;
                s.ss = 20.0;
                call_lsetup = true;
            }
            s.cjratio = s.cj / s.cjold;
            const double temp1 = (1.0 - XRATE) / (1.0 + XRATE);
            const double temp2 = 1.0 / temp1;
            if (s.cjratio < temp1 || s.cjratio > temp2) call_lsetup = true;
            if (s.cj != s.cjlast) s.ss = 100.0;
            s.info = call_lsetup ? 1 : 0;  // hand the decision to the team
        }
        Team::sync();
        bool call_lsetup = S->info != 0;
        int ret = 0;
        for (;;) {
            for (int i = t; i < n; i += T) delta[i] = 0.0;
            Team::sync();
            nls_sys(delta);
            if (call_lsetup) {
                nls_setup();
                if (S->info != 0) {
                    ret = B200IDA_ST_LSETUP_FAIL;
                    break;
                }
                if (leader()) S->jcur = 1;
            }
            if (leader()) S->curiter = 0;
            for (int i = t; i < n; i += T) ee[i] = 0.0;
            Team::sync();
            for (;;) {
                if (leader()) S->nni += 1;
                for (int i = t; i < n; i += T) nd[i] = -nd[i];
                Team::sync();
                lsolve();
                for (int i = t; i < n; i += T) ee[i] += nd[i];
                Team::sync();
                wrms_to(nd, &S->nrm, false);  // the nonlinear solver's norm is never masked
                if (leader()) {  // ctest: src/ida_nls.rs:218-266
                    TeamScalars& s = *S;
                    const double delnrm = s.nrm;
                    int c = 0;
                    bool decided = false;
                    if (s.curiter == 0) {
                        s.oldnrm = delnrm;
                        if (delnrm <= 0.0001 * s.toldel) {
                            c = 1;
                            decided = true;
                        }
                    } else {
                        const double base = delnrm / s.oldnrm;  // pow(x, 1.0) == x bit-for-bit (tests/test_detpow.py)
                        const double rate = (s.curiter == 1) ? base : detpow::det_pow(base, 1.0 / (double)s.curiter);
                        if (rate > RATEMAX) {
                            c = -1;
                            decided = true;
                        } else {
                            s.ss = rate / (1.0 - rate);
                        }
                    }
                    if (!decided) c = (s.ss * delnrm <= s.eps_newt) ? 1 : 0;
                    if (c == 1) s.jcur = 0;
                    if (c == 0) {
                        s.curiter += 1;
                        if (s.curiter >= MAXNLSIT) c = -1;
                    }
                    s.flag = c;
                }
                Team::sync();
                const int c = S->flag;
                if (c == 1) {
                    ret = 0;
                    break;
                }
                if (c == -1) {
                    ret = 1;
                    break;
                }
                nls_sys(ee);
            }
            if (ret == 1 && !S->jcur) {
                Team::sync();
                if (leader()) S->nlscf += 1;
                call_lsetup = true;
                continue;
            }
            break;
        }
        Team::sync();
        if (leader()) {
            if (ret != 0) S->nlscf += 1;
            S->flag = ret;
        }
        const double cj = S->cj;
        for (int i = t; i < n; i += T) {
            yy[i] = yypred[i] + ee[i];
            yp[i] = yppred[i] + cj * ee[i];
        }
        Team::sync();
    }

    // ---- test_error (src/lib.rs:967-1039): S->flag = converged ? 0 : 2 ; S->err_k, S->err_km1, S->knew ----------
    __device__ void test_error() {
        const int kk = S->kk;
        // The up to three norms (ee; ee + phi[kk]; ee + phi[kk] + phi[kk-1]) are independent: their weighted squares are formed
        // in ONE pass (delta gets the operands and the order of the reference's two assignments) and the three index-ordered
        // sums -- each a chain of n dependent adds run by one thread -- are taken by three different threads (three warps of a
        // CTA team, three lanes of a warp team) at the same time: one chain's latency and three barriers instead of three
        // and ten.  nd and tmp are free between two Newton solves and hold the second and third set of squares.
        const bool m = idmask != nullptr;
        const int nnorm = 1 + (kk > 1 ? 1 : 0) + (kk > 2 ? 1 : 0);
        for (int i = t; i < n; i += T) {
            const double w = ewt[i], e = ee[i];
            double q = e * w;
            if (m) q = q * idmask[i];
            prod[i] = q * q;
            if (kk > 1) {
                double d = phi[kk * n + i] + e;
                q = d * w;
                if (m) q = q * idmask[i];
                nd[i] = q * q;
                if (kk > 2) {
                    d += phi[(kk - 1) * n + i];
                    q = d * w;
                    if (m) q = q * idmask[i];
                    tmp[i] = q * q;
                }
                delta[i] = d;
            }
        }
        Team::sync();
        {
            const int step = (T >= 96) ? 32 : 1;  // CTA teams: one summing thread per warp; warp teams: lanes 0, 1, 2
            const int which = (t == 0) ? 0 : ((t == step) ? 1 : ((t == 2 * step) ? 2 : 3));
            if (which < nnorm) {
                const double* pp = (which == 0) ? prod : ((which == 1) ? nd : tmp);
                double acc = 0.0;
                for (int i = 0; i < n; ++i) acc = acc + pp[i];
                const double nrm = sqrt(acc / (double)n);
                if (which == 0) S->nrm = nrm;
                else if (which == 1) S->en1 = nrm;
                else S->en2 = nrm;
            }
        }
        Team::sync();
        const double enorm_km1 = (kk > 1) ? S->en1 : 0.0, enorm_km2 = (kk > 2) ? S->en2 : 0.0;
        if (leader()) {
            TeamScalars& s = *S;
            const double scalar_kk = (double)kk;
            const double enorm_k = s.nrm;
            s.err_k = s.sigma[kk] * enorm_k;
            const double terr_k = s.err_k * (scalar_kk + 1.0);
            s.err_km1 = 0.0;
            int knew = kk;
            if (kk > 1) {
                s.err_km1 = s.sigma[kk - 1] * enorm_km1;
                const double terr_km1 = scalar_kk * s.err_km1;
                if (kk > 2) {
                    const double err_km2 = s.sigma[kk - 2] * enorm_km2;
                    const double terr_km2 = (scalar_kk - 1.0) * err_km2;
                    if (fmax(terr_km1, terr_km2) <= terr_k) knew = kk - 1;
                } else {
                    if (terr_km1 <= terr_k * 0.5) knew = kk - 1;
                }
            }
            s.knew = knew;
            s.flag = ((s.ck * enorm_k) <= 1.0) ? 0 : 2;
        }
        Team::sync();
    }

    // ---- restore (src/lib.rs:1044-1083) ---------------------------------------------------------------------------
    __device__ void restore() {
        const int kk = S->kk, ns = S->ns;
        for (int j = ns; j <= kk; ++j) {
            const double c = 1.0 / S->beta[j];
            for (int i = t; i < n; i += T) phi[j * n + i] *= c;
        }
        Team::sync();
        if (leader()) {
            TeamScalars& s = *S;
            s.tn = s.saved_t;
            for (int j = 1; j <= kk; ++j) s.psi[j - 1] = s.psi[j] - s.hh;
        }
        Team::sync();
    }

    // ---- handle_n_flag (src/lib.rs:1120-1244), leader only; returns kflag in S->flag --------------------------------
    __device__ void handle_n_flag_leader(int nflag) {
        TeamScalars& s = *S;
        s.phase = 1;
        if (nflag == 2) {
            s.nef += 1;
            s.netf += 1;
            if (s.nef == 1) {
                const double err_knew = (s.kk == s.knew) ? s.err_k : s.err_km1;
                s.kk = s.knew;
                const double base = 2.0 * err_knew + 0.0001;
                const double arg = 1.0 / (double)(s.kk + 1);
                s.rr = 0.9 * detpow::det_pow(base, -arg);
                s.rr = fmax(0.25, fmin(0.9, s.rr));
                s.hh *= s.rr;
                s.flag = 0;
            } else if (s.nef == 2) {
                s.kk = s.knew;
                s.rr = 0.25;
                s.hh *= s.rr;
                s.flag = 0;
            } else if (s.nef < MXNEF) {
                s.kk = 1;
                s.rr = 0.25;
                s.hh *= s.rr;
                s.flag = 0;
            } else {
                s.flag = B200IDA_ST_ERR_FAIL;
            }
            return;
        }
        s.ncf += 1;
        s.ncfn += 1;
        s.rr = 0.25;
        s.hh *= s.rr;
        s.flag = (s.ncf < MXNCF) ? 0 : B200IDA_ST_CONV_FAIL;
    }

    // ---- complete_step (src/impl_complete_step.rs:22-177) -----------------------------------------------------------
    __device__ void complete_step() {
        // leader: bookkeeping up to the point where the order-(k+1) error estimate may be needed
        if (leader()) {
            TeamScalars& s = *S;
            s.nst += 1;
            s.info = s.kk - s.kused;  // kdiff
            s.kused = s.kk;
            s.hused = s.hh;
            if (s.knew == s.kk - 1 || s.kk == MAXORD) s.phase = 1;
            int need = 0;
            if (s.phase != 0 && !(s.knew == s.kk - 1) && !(s.kk == MAXORD) && !(s.kk + 1 >= s.ns || s.info == 1)) need = 1;
            s.flag = need;
        }
        Team::sync();
        if (S->flag) {
            const int kk = S->kk;
            for (int i = t; i < n; i += T) tmp[i] = ee[i] - phi[(kk + 1) * n + i];
            Team::sync();
            wrms_to(tmp, &S->scale);
        }
        if (leader()) {
            TeamScalars& s = *S;
            const int kdiff = s.info;
            if (s.phase == 0) {
                if (s.nst > 1) {
                    s.kk += 1;
                    double hnew = 2.0 * s.hh;
                    const double tt = fabs(hnew) * HMAX_INV;
                    if (tt > 1.0) hnew /= tt;
                    s.hh = hnew;
                }
            } else {
                int action;
                double err_kp1 = 0.0;
                if (s.knew == s.kk - 1) {
                    action = 0;
                } else if (s.kk == MAXORD) {
                    action = 1;
                } else if (s.kk + 1 >= s.ns || kdiff == 1) {
                    action = 1;
                } else {
                    const double enorm = s.scale;
                    err_kp1 = enorm / (double)(s.kk + 2);
                    const double terr_k = (double)(s.kk + 1) * s.err_k;
                    const double terr_kp1 = (double)(s.kk + 2) * err_kp1;
                    if (s.kk == 1) {
                        action = (terr_kp1 >= 0.5 * terr_k) ? 1 : 2;
                    } else {
                        const double terr_km1 = (double)s.kk * s.err_km1;
                        if (terr_km1 <= fmin(terr_k, terr_kp1)) action = 0;
                        else if (terr_kp1 >= terr_k) action = 1;
                        else action = 2;
                    }
                }
                double err_knew;
                if (action == 2) {
                    s.kk += 1;
                    err_knew = err_kp1;
                } else if (action == 0) {
                    s.kk -= 1;
                    err_knew = s.err_km1;
                } else {
                    err_knew = s.err_k;
                }
                double hnew = s.hh;
                s.rr = detpow::det_pow(2.0 * err_knew + 0.0001, -(1.0 / (double)(s.kk + 1)));
                if (s.rr >= 2.0) {
                    hnew = 2.0 * s.hh;
                    const double tt = fabs(hnew) * HMAX_INV;
                    if (tt > 1.0) hnew /= tt;
                } else if (s.rr <= 1.0) {
                    s.rr = fmax(0.5, fmin(s.rr, 0.9));
                    hnew = s.hh * s.rr;
                }
                s.hh = hnew;
            }
        }
        Team::sync();
        const int kused = S->kused;
        const double ck = S->ck;
        for (int i = t; i < n; i += T) {
            const double e = ee[i];
            if (kused < MAXORD) phi[(kused + 1) * n + i] = e;
            double acc = e;
            for (int j = kused; j >= 0; --j) {
                acc += phi[j * n + i];
                phi[j * n + i] = acc;
            }
            ee[i] = e * ck;  // src/lib.rs:708
        }
        Team::sync();
    }

    // ---- get_solution (src/lib.rs:1274-1343); S->flag = 0 or B200IDA_ST_BAD_T ----------------------------------------
    __device__ void get_solution(double tq) {
        Team::sync();
        if (leader()) {
            TeamScalars& s = *S;
            const double sgn = (s.hh > 0.0 || (s.hh == 0.0 && !signbit(s.hh))) ? 1.0 : -1.0;
            const double tfuzz = 100.0 * DBL_EPS * (fabs(s.tn) + fabs(s.hh)) * sgn;
            const double tp = s.tn - s.hused - tfuzz;
            s.flag = ((tq - tp) * s.hh < 0.0) ? B200IDA_ST_BAD_T : 0;
        }
        Team::sync();
        if (S->flag != 0) return;
        const int kord = (S->kused == 0) ? 1 : S->kused;
        const double delt = tq - S->tn;
        // every thread recomputes the (cheap, scalar) cvals/dvals recurrence for its components
        for (int i = t; i < n; i += T) {
            double c = 1.0, d = 0.0;
            double gam = delt / S->psi[0];
            double y = 0.0 + c * phi[i], ypv = 0.0;
            for (int j = 1; j <= kord; ++j) {
                d = d * gam + c / S->psi[j - 1];
                c = c * gam;
                gam = (delt + S->psi[j - 1]) / S->psi[j];
                const double p = phi[j * n + i];
                y += c * p;
                ypv += d * p;
            }
            yy[i] = y;
            yp[i] = ypv;
        }
        Team::sync();
    }

    // ---- first-call block of solve() (src/impl_solve.rs:84-173); S->flag = 0 or B200IDA_ST_ILL_INPUT -------------------
    __device__ void first_call(double tout) {
        if (!S->setup_done) ewt_set(false);
        wrms_to(PHI(1), &S->nrm);
        if (leader()) {
            TeamScalars& s = *S;
            s.setup_done = 1;
            s.flag = 0;
            const double tdist = fabs(tout - s.tn);
            const double troundoff = 2.0 * DBL_EPS * (fabs(s.tn) + fabs(tout));
            if (tdist == 0.0 || tdist < troundoff) {
                s.flag = B200IDA_ST_ILL_INPUT;
            } else {
                s.hh = 0.001 * tdist;
                const double ypnorm = s.nrm;
                if (ypnorm > 2.0 / s.hh) s.hh = 0.5 / ypnorm;
                if (tout < s.tn) s.hh = -s.hh;
                const double rh = fabs(s.hh) * HMAX_INV;
                if (rh > 1.0) s.hh /= rh;
                s.h0u = s.hh;
                s.kk = 0;
                s.kused = 0;
                s.eps_newt = EPCON;
                s.toldel = 0.0001 * s.eps_newt;
            }
        }
        Team::sync();
        if (S->flag == 0) {
            const double hh = S->hh;
            for (int i = t; i < n; i += T) phi[n + i] *= hh;
            Team::sync();
        }
    }

    // ---- Ida::solve for a ladder of output times (src/impl_solve.rs:69-376, NORMAL stop tests) ------------------------------
    __device__ void solve_ladder(const LadderParams& lad, double* __restrict__ yy_traj, int32_t* __restrict__ status, int64_t B,
                                 int64_t b) {
        for (int it = 0; it < lad.ntout; ++it) {
            const double tout = lad.touts[it];
            int call_status = 0;
            bool call_done = false;
            Team::sync();
            if (S->nst == 0) {
                first_call(tout);
                if (S->flag != 0) {
                    call_status = S->flag;
                    call_done = true;
                }
            } else {
                // stop_test1, NORMAL (src/impl_stop_test.rs:53-65)
                if (tout == S->tretlast) {
                    Team::sync();
                    if (leader()) S->tret = tout;
                    call_done = true;
                } else if ((S->tn - tout) * S->hh >= 0.0) {
                    get_solution(tout);
                    call_status = S->flag;
                    Team::sync();
                    if (leader() && call_status == 0) {
                        S->tretlast = tout;
                        S->tret = tout;
                    }
                    call_done = true;
                }
            }
            Team::sync();
            if (leader()) S->nstloc = 0;
            Team::sync();
            while (!call_done) {
                // ---- top of the step loop: src/impl_solve.rs:246-311 ----
                if (S->nstloc >= MXSTEP) {
                    Team::sync();
                    if (leader()) {
                        S->tret = S->tn;
                        S->tretlast = S->tn;
                    }
                    call_status = B200IDA_ST_TOO_MUCH_WORK;
                    break;
                }
                if (S->nst > 0) {
                    ewt_set(true);
                    if (S->flag) {
                        get_solution(S->tn);
                        Team::sync();
                        if (leader()) {
                            S->tret = S->tn;
                            S->tretlast = S->tn;
                        }
                        call_status = B200IDA_ST_BAD_EWT;
                        break;
                    }
                }
                wrms_to(PHI(0), &S->nrm);
                if (leader()) {
                    S->tolsf = DBL_EPS * S->nrm;
                    S->flag = (S->tolsf > 1.0) ? 1 : 0;
                    if (S->flag) S->tolsf *= 10.0;
                }
                Team::sync();
                if (S->flag) {
                    const bool had_steps = S->nst > 0;
                    Team::sync();
                    if (leader()) {
                        S->tret = S->tn;
                        S->tretlast = S->tn;
                    }
                    if (had_steps) get_solution(S->tn);
                    call_status = B200IDA_ST_TOO_MUCH_ACC;
                    break;
                }
                // ---- step(): src/lib.rs:613-711 ----
                if (leader()) {
                    TeamScalars& s = *S;
                    s.saved_t = s.tn;
                    if (s.nst == 0) {
                        s.kk = 1;
                        s.kused = 0;
                        s.hused = 0.0;
                        s.psi[0] = s.hh;
                        s.cj = 1.0 / s.hh;
                        s.phase = 0;
                        s.ns = 0;
                    }
                    s.ncf = 0;
                    s.nef = 0;
                }
                Team::sync();
                int kflag = 0;
                for (;;) {  // attempts
                    if (tr.id == (long long)b) {  // team-uniform: data_trace frame of the traced instance (src/lib.rs:635-639)
                        const Layout lay(n);
                        const int c = *reinterpret_cast<volatile int*>(tr.count);
                        if (c < tr.max) store(tr.d + (size_t)c * lay.ndouble(), tr.i + (size_t)c * lay.nint(), 1, 0);
                        Team::sync();
                        if (leader()) *tr.count = c + 1;
                        Team::sync();
                    }
                    if (leader()) S->natt += 1;
                    set_coeffs();
                    Team::sync();
                    if (leader()) S->tn += S->hh;
                    predict();
                    Team::sync();
                    nonlinear_solve();
                    int nflag = S->flag;
                    Team::sync();
                    if (nflag == 0) {
                        test_error();
                        nflag = S->flag;
                        Team::sync();
                    } else if (leader()) {
                        S->err_k = 0.0;
                        S->err_km1 = 0.0;
                    }
                    if (nflag == 0) break;  // converged and passed the error test
                    restore();
                    if (leader()) {
                        if (nflag < 0) {
                            S->ncfn += 1;
                            S->flag = nflag;
                        } else {
                            handle_n_flag_leader(nflag);
                        }
                    }
                    Team::sync();
                    kflag = S->flag;
                    Team::sync();
                    if (kflag != 0) break;
                    if (S->nst == 0) {  // reset(): src/lib.rs:1249-1252 (scales ALL of phi)
                        const double rr = S->rr;
                        for (int e = t; e < MXORDP1 * n; e += T) phi[e] *= rr;
                        if (leader()) S->psi[0] = S->hh;
                        Team::sync();
                    }
                }
                if (kflag != 0) {
                    // failed step: src/impl_solve.rs:314-327
                    get_solution(S->tn);
                    const int ok = S->flag == 0;
                    Team::sync();
                    if (leader() && ok) {
                        S->tret = S->tn;
                        S->tretlast = S->tn;
                    }
                    call_status = kflag;
                    break;
                }
                complete_step();
                if (leader()) S->nstloc += 1;
                Team::sync();
                // stop_test2, NORMAL (src/impl_stop_test.rs:152-161)
                if ((S->tn - tout) * S->hh >= 0.0) {
                    Team::sync();
                    if (leader()) {
                        S->tret = tout;
                        S->tretlast = tout;
                    }
                    get_solution(tout);
                    call_status = B200IDA_ST_SUCCESS;
                    break;
                }
            }
            Team::sync();
            if (yy_traj)
                for (int i = t; i < n; i += T) yy_traj[((int64_t)it * n + i) * B + b] = yy[i];
            if (status && leader()) status[(int64_t)it * B + b] = call_status;
        }
        Team::sync();
    }
};

}  // namespace b200ida
