// ida_thread.cuh -- thread-per-instance ensemble IDA for small compile-time n (Roberts, Lorenz: n = 3).
//
// One CUDA thread integrates one DAE instance; the whole `Ida::solve` loop of the reference
// (src/impl_solve.rs:69-376) runs on the device.  Per-instance state is loaded once from the SoA arrays
// in HBM, integrated, and stored back once: HBM traffic is ~1.5 KB per instance per launch, three orders
// of magnitude below what the FP64 pipe needs, so the kernel is FP64-issue-bound by design.
//
// Where the state lives while integrating (v2; profiles/r1a_* show why v1's all-in-registers layout with
// fully unrolled predicated loops lost: 248 registers -> 12 % occupancy, a 267 KB kernel body -> the
// dominant stall was instruction fetch):
//   * the divided-difference array phi[6][n] and the five 6-entry coefficient arrays psi/alpha/beta/
//     sigma/gamma are indexed by the (per-instance, run-time) order kk, so they sit in SHARED memory,
//     one column per thread ([field][thread], 8-byte words -> conflict-free), and the loops over the
//     order are real loops (small code, divergence only in trip count);
//   * n-vectors (ee, ewt, yy, yp, predictors), the LU factors, pivots and ~20 scalars are registers;
//   * pow, the WRMS norm and the interpolation are out-of-line functions shared by all call sites.
//
// Divergence: lanes differ in step count, retries and Newton iterations.  The solve loop is FLATTENED:
// one trip of the single `while` = one step ATTEMPT of every live lane, whatever output time (ladder
// rung) or retry that lane is on, so a warp's time is max-over-lanes of total attempts, not a sum of
// per-step or per-rung maxima.
//
// Arithmetic is written operation-for-operation like the reference (file:line on each routine), built
// with -fmad=false: + - * / sqrt are IEEE correctly rounded and un-fused, pow is detpow.h.
#pragma once
#include "detpow.h"
#include "fastdiv.cuh"
#include "ida_common.cuh"

namespace b200ida {
// -DROOT_PROF: clock64 phase profile of solve_ladder (tools/prof_roots.py); compiled out otherwise
#ifdef ROOT_PROF
__device__ unsigned long long root_prof[8];
#define RP_LEAD() ((__ffs(__activemask()) - 1) == (int)(threadIdx.x & 31))
#define RP_T(v) long long v = clock64()
#define RP_ACC(slot, v) do { if (RP_LEAD()) atomicAdd(&root_prof[slot], (unsigned long long)(clock64() - (v))); } while (0)
#define RP_COUNT(slot) do { if (RP_LEAD()) atomicAdd(&root_prof[slot], 1ull); } while (0)
#else
#define RP_T(v)
#define RP_ACC(slot, v)
#define RP_COUNT(slot)
#endif

#ifndef IDA_THREAD_BLOCK
#define IDA_THREAD_BLOCK 64  // 6 CTAs of 2 warps per SM: same 12 warps as 3 x 128, finer tail granularity (measured -1.5 %)
#endif
#ifndef IDA_WRMS_INLINE
#define IDA_WRMS_INLINE 1
#endif
#ifndef IDA_POW_INLINE
#define IDA_POW_INLINE 0
#endif
constexpr int THREAD_BLOCK = IDA_THREAD_BLOCK;  // threads per CTA of the thread-per-instance kernels

template <int N>
struct TolN {
    double rtol;
    double atol[N];
    double idmask[N];  // ida_id as 1.0 (differential) / 0.0 (algebraic); read by the MASK instantiations only
};

#ifndef IDA_DIV_NOINLINE
#define IDA_DIV_NOINLINE 0
#endif
// IDA_FASTDIV: divisions through fastdiv.cuh (straight-line fast paths, independent quotients batched under one range
// test) instead of nvcc's inline expansion (one taken branch per quotient, chains not interleaved); bit-identical
#ifndef IDA_FASTDIV
#define IDA_FASTDIV 1
#endif
#if IDA_DIV_NOINLINE
static __device__ __noinline__ double ddiv(double a, double b) { return a / b; }
#elif IDA_FASTDIV
__device__ __forceinline__ double ddiv(double a, double b) { return div1(a, b); }
#else
__device__ __forceinline__ double ddiv(double a, double b) { return a / b; }
#endif
#if IDA_FASTDIV
__device__ __forceinline__ double drcp(double b) { return rcp1(b); }
#else
__device__ __forceinline__ double drcp(double b) { return 1.0 / b; }
#endif
// det_pow's main path only (FAST kernels): NaN when the arguments leave it -- the caller poisons the lane, the generic
// pow fallback (several hundred instructions) is not in the kernel
static __device__ __noinline__ double pow_main_noinline(double x, double y) {
    double r;
    if (!detpow::pow_main(x, y, &r)) r = __longlong_as_double(0x7ff8000000000000LL);
    return r;
}
#if IDA_POW_INLINE
static __device__ __forceinline__ double pow_noinline(double x, double y) { return detpow::det_pow(x, y); }
#else
static __device__ __noinline__ double pow_noinline(double x, double y) { return detpow::det_pow(x, y); }
#endif

// x / D for a small integer constant D, bit-identical to the IEEE division the reference performs.
// Markstein's sequence: y = RN(1/D); q = RN(x*y); r = x - D*q (exact in one FMA); q' = RN(q + r*y) is the
// correctly rounded quotient whenever nothing over/underflows (checked against `/` on 3e8 random x per D in
// [2,7]; tools/ notes in DESIGN.md).  3 FP64 instructions instead of the ~25 of a generic division.  Outside
// 2^-900 <= |x| <= 2^900 (incl. 0, inf, nan) it falls back to the real division.
template <int D>
__device__ __forceinline__ double div_const(double x) {
    constexpr double y = 1.0 / (double)D;
    const unsigned hi = (unsigned)__double2hiint(x) & 0x7ff00000u;
    if (IDA_UNSAFE_NOFALLBACK || hi - 0x07b00000u <= 0x70800000u) {  // biased exponent in [123, 1923]
        const double q = __dmul_rn(x, y);
        const double r = __fma_rn(-(double)D, q, x);
        return __fma_rn(r, y, q);
    }
    return x / (double)D;
}

// ---- WRMS norm: src/norm_rms.rs:31-38 (ordered fold from 0) -------------------------------------------
// sqrt(acc / n) is the expensive, shareable part (~35 instructions with the division and the square root); the fold
// itself is 8 FP64 instructions on values the caller already holds in registers.  So the fold is inlined at every call
// site and only the tail is an out-of-line function -- one double in, one out -- instead of marshalling six doubles
// through the call (the round-1 profile: 9 % of all executed instructions were register moves around wrms3).
template <int N>
__device__ __forceinline__ double wrms_acc(const double (&x)[N], const double (&w)[N]) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double p = x[i] * w[i];
        acc = acc + p * p;
    }
    return acc;
}
template <int N>
__device__ __forceinline__ double wrms_tail_inline(double acc) {
#if IDA_UNSAFE_NOFALLBACK
    bool ok;
    return sqrt_fast(div_const<N>(acc), ok);
#else
    if constexpr (N >= 2 && N <= 7) return sqrt(div_const<N>(acc));
    else return sqrt(acc / (double)N);
#endif
}
template <int N>
__device__ __noinline__ double wrms_tail(double acc) { return wrms_tail_inline<N>(acc); }
// the three norms of test_error (src/lib.rs:980-1012) are independent: one call, three interleaved chains
template <int N>
__device__ __noinline__ void wrms_tail3(double a0, double a1, double a2, double& r0, double& r1, double& r2) {
    r0 = wrms_tail_inline<N>(a0);
    r1 = wrms_tail_inline<N>(a1);
    r2 = wrms_tail_inline<N>(a2);
}
template <int N>
__device__ __forceinline__ double wrms_inline(const double (&x)[N], const double (&w)[N]) {
    return wrms_tail_inline<N>(wrms_acc<N>(x, w));
}
template <int N>
__device__ __forceinline__ double wrms(const double (&x)[N], const double (&w)[N]) {
    if constexpr (IDA_WRMS_INLINE) return wrms_inline<N>(x, w);
    else return wrms_tail<N>(wrms_acc<N>(x, w));
}

// ---- get_solution: src/lib.rs:1274-1343 ; returns 0 or B200IDA_ST_BAD_T.  Out of line (called ~12 times per
// instance); reads phi / psi from the thread's shared-memory column, writes yret to out[0..N), ypret to out[N..2N).
template <int N, int STRIDE = THREAD_BLOCK>
__device__ __noinline__ int get_solution_fn(const double* sm, double hh, double hused, double tn, int kused, double t,
                                            double* out) {
    const double sgn = (hh > 0.0 || (hh == 0.0 && !signbit(hh))) ? 1.0 : -1.0;
    const double tfuzz = 100.0 * DBL_EPS * (fabs(tn) + fabs(hh)) * sgn;
    const double tp = tn - hused - tfuzz;
    if ((t - tp) * hh < 0.0) return B200IDA_ST_BAD_T;
    const int kord = (kused == 0) ? 1 : kused;
    const double delt = t - tn;
    double c = 1.0, d = 0.0;
    // The reference computes gam = delt / psi[0] before the loop and gam = (delt + psi[j-1]) / psi[j] at the end of
    // iteration j, the last of which is never read.  Here iteration j computes the gam it uses at its top (same
    // operands, same operations, one division fewer per call and one inlined division sequence less code).
    double num = delt;
    // yret = sum_j cvals[j] phi[j], ypret = sum_j dvals[j-1] phi[j], accumulated from 0 in j order
    double y[N], yp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        y[i] = 0.0 + c * sm[i * STRIDE];
        yp[i] = 0.0;
    }
#pragma unroll 1
    for (int j = 1; j <= kord; ++j) {
        const double psi_jm1 = sm[(6 * N + j - 1) * STRIDE];
        const double gam = num / psi_jm1;
        d = d * gam + c / psi_jm1;
        c = c * gam;
        num = delt + psi_jm1;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double p = sm[(j * N + i) * STRIDE];
            y[i] += c * p;
            yp[i] += d * p;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        out[i] = y[i];
        out[N + i] = yp[i];
    }
    return 0;
}

__host__ __device__ constexpr double alphas_seq(int kk) {
    double a = 0.0;
    for (int i = 0; i < kk; ++i) a -= 1.0 / (double)(i + 1);
    return a;
}
__device__ __forceinline__ double alphas_of(int kk) {
    constexpr double t0 = alphas_seq(0), t1 = alphas_seq(1), t2 = alphas_seq(2), t3 = alphas_seq(3), t4 = alphas_seq(4),
                     t5 = alphas_seq(5);
    double v = t0;
    v = (kk == 1) ? t1 : v;
    v = (kk == 2) ? t2 : v;
    v = (kk == 3) ? t3 : v;
    v = (kk == 4) ? t4 : v;
    v = (kk == 5) ? t5 : v;
    return v;
}

// 1.0 / (double)k for k = 1..6, the exponents of the step-size formulas (src/impl_complete_step.rs:128-132,
// src/lib.rs:1155-1160): correctly rounded constants, selected instead of divided
__device__ __forceinline__ double inv_small(int k) {
    double v = 1.0;
    v = (k == 2) ? 1.0 / 2.0 : v;
    v = (k == 3) ? 1.0 / 3.0 : v;
    v = (k == 4) ? 1.0 / 4.0 : v;
    v = (k == 5) ? 1.0 / 5.0 : v;
    v = (k == 6) ? 1.0 / 6.0 : v;
    return v;
}

// The Illinois iteration of IDARootfind (src/impl_r_check.rs:402-576) on the root state in the thread's shared-memory
// column: entered with a sign change of g_imax in (tlo, thi), leaves trout / grout / iroots set and returns the number
// of g evaluations it made.  Out of line on purpose (see IdaThread::root_find).
template <class P>
__device__ __noinline__ int root_illinois(double* sm, P prob, unsigned gmask, int imax, double hh, double hused, double tn,
                                          int kused) {
    constexpr int N = P::N, NR = P::NR > 0 ? P::NR : 1, SD = 6 * N + 30;
    auto rsm = [&](int k) -> double& { return sm[(SD + k) * THREAD_BLOCK]; };
    auto signum = [](double x) { return (x != x) ? x : (signbit(x) ? -1.0 : 1.0); };
    int evals = 0;
    double alph = 1.0;
    int side = 0, sideprev = -1;
#pragma unroll 1
    for (;;) {
        if (fabs(rsm(4 * NR + 1) - rsm(4 * NR + 0)) <= rsm(4 * NR + 3)) break;
        alph = (sideprev == side) ? ((side == 2) ? alph * 2.0 : alph * 0.5) : 1.0;
        const double gh = rsm(1 * NR + imax), gl = rsm(0 * NR + imax);
        double tmid = rsm(4 * NR + 1) - (rsm(4 * NR + 1) - rsm(4 * NR + 0)) * gh / (gh - alph * gl);
        if (fabs(tmid - rsm(4 * NR + 0)) < 0.5 * rsm(4 * NR + 3)) {
            const double fracint = fabs(rsm(4 * NR + 1) - rsm(4 * NR + 0)) / rsm(4 * NR + 3);
            const double fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
            tmid = rsm(4 * NR + 0) + fracsub * (rsm(4 * NR + 1) - rsm(4 * NR + 0));
        }
        if (fabs(rsm(4 * NR + 1) - tmid) < 0.5 * rsm(4 * NR + 3)) {
            const double fracint = fabs(rsm(4 * NR + 1) - rsm(4 * NR + 0)) / rsm(4 * NR + 3);
            const double fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
            tmid = rsm(4 * NR + 1) - fracsub * (rsm(4 * NR + 1) - rsm(4 * NR + 0));
        }
        double out[2 * N], y[N], yp[N], g[NR];
        get_solution_fn<N>(sm, hh, hused, tn, kused, tmid, out);  // t inside the last step: cannot fail
#pragma unroll
        for (int i = 0; i < N; ++i) {
            y[i] = out[i];
            yp[i] = out[N + i];
        }
        if constexpr (P::NR > 0) prob.root(tmid, y, yp, g);
#pragma unroll
        for (int i = 0; i < NR; ++i) rsm(2 * NR + i) = g[i];
        evals += 1;
        sideprev = side;
        // in which subinterval does g change sign: src/impl_r_check.rs:457-492
        bool zroot = false, sgnchg = false;
        double maxfrac = 0.0;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (!((gmask >> i) & 1u)) continue;
            const double glo_i = rsm(0 * NR + i), gx = rsm(2 * NR + i);
            const bool rootdir_glo_neg = 0.0 * glo_i <= 0.0;
            if (fabs(gx) == 0.0 && rootdir_glo_neg) {
                zroot = true;
            } else if (glo_i * gx < 0.0 && rootdir_glo_neg) {
                const double gfrac = fabs(gx / (gx - glo_i));
                if (gfrac > maxfrac) {
                    sgnchg = true;
                    maxfrac = gfrac;
                    imax = i;
                }
            }
        }
        if (sgnchg) {
            rsm(4 * NR + 1) = tmid;
#pragma unroll
            for (int i = 0; i < NR; ++i) rsm(1 * NR + i) = rsm(2 * NR + i);
            side = 1;
            if (fabs(rsm(4 * NR + 1) - rsm(4 * NR + 0)) <= rsm(4 * NR + 3)) break;
            continue;
        }
        if (zroot) {
            rsm(4 * NR + 1) = tmid;
#pragma unroll
            for (int i = 0; i < NR; ++i) rsm(1 * NR + i) = rsm(2 * NR + i);
            break;
        }
        rsm(4 * NR + 0) = tmid;
#pragma unroll
        for (int i = 0; i < NR; ++i) rsm(0 * NR + i) = rsm(2 * NR + i);
        side = 2;
        if (fabs(rsm(4 * NR + 1) - rsm(4 * NR + 0)) <= rsm(4 * NR + 3)) break;
    }
    rsm(4 * NR + 2) = rsm(4 * NR + 1);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        rsm(2 * NR + i) = rsm(1 * NR + i);
        rsm(3 * NR + i) = 0.0;
        const double glo_i = rsm(0 * NR + i), ghi_i = rsm(1 * NR + i);
        if (((gmask >> i) & 1u) && 0.0 * glo_i <= 0.0 && (fabs(ghi_i) == 0.0 || glo_i * ghi_i < 0.0)) rsm(3 * NR + i) = signum(glo_i);
    }
    return evals;
}

// IDARcheck3 + the first pass of IDARootfind (src/impl_r_check.rs:211-267, 330-400) as ONE out-of-line function on the
// root state in the thread's shared-memory column, so that the step loop of the ROOTS kernel only carries a call: with
// these ~400 instructions inline the loop outgrew the instruction cache (ncu: 5.0 no_instruction stalls per issue, 2.1x
// the run time of the plain kernel).  Returns ier | gactive mask << 1 | g evaluations << 8; when ier == 1 the caller
// still has to interpolate y(trout) (get_solution), everything else is done here.
template <class P>
__device__ __noinline__ unsigned root_check3(double* sm, P prob, unsigned gmask, double tn, double hh, double hused, int kused) {
    constexpr int N = P::N, NR = P::NR > 0 ? P::NR : 1, SD = 6 * N + 30;
    auto rsm = [&](int k) -> double& { return sm[(SD + k) * THREAD_BLOCK]; };
    auto signum = [](double x) { return (x != x) ? x : (signbit(x) ? -1.0 : 1.0); };
    unsigned evals = 1;
    const bool at_tn = (rsm(4 * NR + 4) - tn) * hh >= 0.0;
    const double thi = at_tn ? tn : rsm(4 * NR + 4);  // IdaTask::Normal
    rsm(4 * NR + 1) = thi;
    {
        double y[N], yp[N], g[NR];
        if (at_tn && !P::ROOT_USES_YP) {
            // get_solution(tn) is the interpolant at delt = 0: c_0 = 1 and every other c_j is 0 * (finite) = 0, so it
            // returns phi[0] (src/lib.rs:1274-1343) -- evaluated here without its ~10 divisions, once per step.  yy / yp
            // themselves are not refreshed; every path on which the reference's values would be observed refreshes
            // them (root return, output time, failure returns, and the mxstep return patched in solve_ladder).
#pragma unroll
            for (int i = 0; i < N; ++i) {
                y[i] = 0.0 + sm[i * THREAD_BLOCK];
                yp[i] = 0.0;
            }
        } else {
            double out[2 * N];
            get_solution_fn<N>(sm, hh, hused, tn, kused, thi, out);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                y[i] = out[i];
                yp[i] = out[N + i];
            }
        }
        if constexpr (P::NR > 0) prob.root(thi, y, yp, g);
#pragma unroll
        for (int i = 0; i < NR; ++i) rsm(1 * NR + i) = g[i];
    }
    rsm(4 * NR + 3) = (fabs(tn) + fabs(hh)) * DBL_EPS * 100.0;
    // first pass of IDARootfind: sign change in, or zero of, g(thi)
    bool zroot = false, sgnchg = false;
    double maxfrac = 0.0;
    int imax = 0;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        if (!((gmask >> i) & 1u)) continue;
        const double glo_i = rsm(0 * NR + i), ghi_i = rsm(1 * NR + i);
        const bool rootdir_glo_neg = 0.0 * glo_i <= 0.0;  // rootdir is all zeros (src/lib.rs:368)
        if (fabs(ghi_i) == 0.0) {
            if (rootdir_glo_neg) zroot = true;
        } else if (glo_i * ghi_i < 0.0 && rootdir_glo_neg) {
            const double gfrac = fabs(ghi_i / (ghi_i - glo_i));
            if (gfrac > maxfrac) {
                sgnchg = true;
                maxfrac = gfrac;
                imax = i;
            }
        }
    }
    unsigned ier = 0;
    if (!sgnchg) {
        rsm(4 * NR + 2) = thi;
#pragma unroll
        for (int i = 0; i < NR; ++i) rsm(2 * NR + i) = rsm(1 * NR + i);
        if (zroot) {
            ier = 1;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const double glo_i = rsm(0 * NR + i);
                rsm(3 * NR + i) = 0.0;
                if (((gmask >> i) & 1u) && fabs(rsm(1 * NR + i)) == 0.0 && 0.0 * glo_i <= 0.0) rsm(3 * NR + i) = signum(glo_i);
            }
        }
    } else {
        evals += (unsigned)root_illinois<P>(sm, prob, gmask, imax, hh, hused, tn, kused);
        ier = 1;
    }
    // src/impl_r_check.rs:247-258: inactive components that moved away from zero become active; tlo, glo <- trout, grout
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const double gr = rsm(2 * NR + i);
        if (!((gmask >> i) & 1u) && gr != 0.0) gmask |= 1u << i;
        rsm(0 * NR + i) = gr;
    }
    rsm(4 * NR + 0) = rsm(4 * NR + 2);
    return ier | (gmask << 1) | (evals << 8);
}

// MASK: ida_suppressalg is set -- the local-error norms (wrms_norm with mask = suppressalg, src/lib.rs:1353-1370: ypnorm of
// the first call, the tolsf norm, test_error, complete_step) leave the algebraic components out, as norm_wrms_masked
// (src/norm_rms.rs:49-57) does: ((x_i w_i) m_i)^2 folded in index order, divided by n.  A separate instantiation, so the
// production kernel carries none of it.
//
// FAST: the production instantiation.  Every division, reciprocal, square root and pow runs its fast path ONLY -- no
// range test branches to a fallback, no slow-path code is in the kernel at all -- and a failed range test just sets the
// lane's `poison` flag; so do the rare control paths (fatal statuses, mxstep, an output time already passed).  A
// poisoned lane leaves the loop at once and its results are discarded: the launcher re-integrates exactly those
// instances from the untouched input state with the exact (FAST = false) instantiation of the same code, which has
// every fallback.  On a healthy ensemble nothing is poisoned and the results are, bit for bit, the exact kernel's --
// a fast path that passes its range test returns the correctly rounded value by construction (fastdiv.cuh).  Why it
// pays: the cold fallback blocks were interleaved with the hot code (19 % of the static instructions), pushing the
// touched part of the step loop past the 32 KB instruction cache; without them the same loop runs 27 % faster.
template <class P, bool MASK = false, bool FAST = false>
struct IdaThread {
    static constexpr int N = P::N;
#ifndef IDA_POISON_REASON
#define IDA_POISON_REASON 0  // 1 (diagnostic builds): xp.poisoned[] holds a bit mask of WHICH test failed
#endif
#if IDA_POISON_REASON
    unsigned poison = 0;
#define IDA_POISON(bit, cond) poison |= ((cond) ? (bit) : 0u)
#else
    bool poison = false;  // FAST only: some fast path left its range or a rare path was reached -> redo exactly
#define IDA_POISON(bit, cond) poison = poison || (cond)
#endif

    // ---- arithmetic through one door, so that FAST can drop the fallbacks ----------------------------------------
    __device__ __forceinline__ double D(double a, double b) {  // a / b
        if constexpr (FAST) {
            bool ok;
            const double q = div_fast(a, b, ok);
            IDA_POISON(1u, !ok);
            return q;
        } else {
            return ddiv(a, b);
        }
    }
    // a / b where a may be exactly zero (a norm that vanished): FAST handles that in line instead of poisoning the lane
    __device__ __forceinline__ double DZ(double a, double b) {
        if constexpr (FAST) {
            bool ok;
            const double q = div_fast_z(a, b, ok);
            IDA_POISON(1u, !ok);
            return q;
        } else {
            return ddiv(a, b);
        }
    }
    __device__ __forceinline__ double R(double b) {  // 1.0 / b
        if constexpr (FAST) {
            bool ok;
            const double r = rcp_fast(b, ok);
            IDA_POISON(2u, !ok);
            return r;
        } else {
            return drcp(b);
        }
    }
    template <int K>
    __device__ __forceinline__ void DK(const double (&a)[K], const double (&b)[K], double (&q)[K]) {
        if constexpr (FAST) {
            bool all = true;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                bool ok;
                q[k] = div_fast(a[k], b[k], ok);
                all = all && ok;
            }
            IDA_POISON(4u, !all);
        } else {
            div_batch<K>(a, b, q);
        }
    }
    template <int K>
    __device__ __forceinline__ void RK(const double (&b)[K], double (&r)[K]) {
        if constexpr (FAST) {
            bool all = true;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                bool ok;
                r[k] = rcp_fast(b[k], ok);
                all = all && ok;
            }
            IDA_POISON(8u, !all);
        } else {
            rcp_batch<K>(b, r);
        }
    }
    // sqrt(acc / N): the tail of every WRMS norm
    __device__ __forceinline__ double norm_tail(double acc) {
        if constexpr (FAST && N >= 2 && N <= 7) {
            constexpr double y = 1.0 / (double)N;
            const unsigned hi = (unsigned)__double2hiint(acc) & 0x7ff00000u;
            const double q = __dmul_rn(acc, y);  // Markstein's exact x / N (div_const), its range test folded into poison
            const double r = __fma_rn(-(double)N, q, acc);
            bool ok;
            const double v = sqrt_fast(__fma_rn(r, y, q), ok);
            const bool zero = acc == 0.0;  // an exactly vanishing vector: sqrt(0 / N) = 0 (acc is a sum of squares: +0)
            IDA_POISON(16u, !zero && !(hi - 0x07b00000u <= 0x70800000u));
            IDA_POISON(32u, !zero && !ok);
            return zero ? 0.0 : v;
        } else if constexpr (IDA_WRMS_INLINE) {
            return wrms_tail_inline<N>(acc);
        } else {
            return wrms_tail<N>(acc);
        }
    }
    __device__ __forceinline__ double PW(double x, double y) {  // pow(x, y)
        if constexpr (FAST) {
            const double r = pow_main_noinline(x, y);  // NaN when (x, y) is off det_pow's main path
            IDA_POISON(64u, r != r);
            return r;
        } else {
            return pow_noinline(x, y);
        }
    }
    static constexpr int SMEM_DOUBLES = 6 * N + 30;  // phi + psi, alpha, beta, sigma, gamma
    using L = Layout;

    // ---- shared-memory resident arrays (run-time indexed by the order) ------------------------------
    double* sm;  // this thread's column: element f at sm[f * THREAD_BLOCK]
    __device__ __forceinline__ double& phi(int j, int i) { return sm[(j * N + i) * THREAD_BLOCK]; }
    __device__ __forceinline__ double& psi(int j) { return sm[(6 * N + j) * THREAD_BLOCK]; }
    __device__ __forceinline__ double& alpha(int j) { return sm[(6 * N + 6 + j) * THREAD_BLOCK]; }
    __device__ __forceinline__ double& beta(int j) { return sm[(6 * N + 12 + j) * THREAD_BLOCK]; }
    __device__ __forceinline__ double& sigma(int j) { return sm[(6 * N + 18 + j) * THREAD_BLOCK]; }
    __device__ __forceinline__ double& gamma(int j) { return sm[(6 * N + 24 + j) * THREAD_BLOCK]; }

    // ---- register state (src/lib.rs:89-244, src/ida_nls.rs:20-60, src/ida_ls.rs:15-106,
    //      newton.rs:14-32, dense.rs:13-20) ---------------------------------------------------------------
    double ee[N], ewt[N], yy[N], yp[N];
    double J[N][N];
    int piv[N];
    double hh, hused, rr, cj, cjold, cjlast, cjratio, ss, toldel, eps_newt, tn, tretlast, h0u, tolsf, tret;
    int kk, kused, knew, phase, ns, setup_done, jcur;
    int nst, nre, nje, nni, netf, ncfn, nsetups, nlscf, natt;
    // ---- scratch ------------------------------------------------------------------------------------
    double yypredict[N], yppredict[N], delta[N], oldnrm;
    P prob;
    double idm[N];  // TolN::idmask (MASK instantiations)

    // wrms_norm(x, ewt, suppressalg): src/lib.rs:1353-1370
    __device__ __forceinline__ double err_acc(const double (&x)[N]) const {
        if constexpr (MASK) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double p = (x[i] * ewt[i]) * idm[i];
                acc = acc + p * p;
            }
            return acc;
        } else {
            return wrms_acc<N>(x, ewt);
        }
    }
    __device__ __forceinline__ double err_norm(const double (&x)[N]) {
        return norm_tail(err_acc(x));
    }

    // ---- HBM <-> registers / shared memory -----------------------------------------------------------
    __device__ __forceinline__ void load(const double* __restrict__ ds, const int* __restrict__ is, int64_t B, int64_t b) {
        constexpr Layout lay(N);
#pragma unroll 1
        for (int f = 0; f < SMEM_DOUBLES; ++f) sm[f * THREAD_BLOCK] = ds[(int64_t)f * B + b];  // phi..gamma are fields 0..6N+29
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ee[i] = ds[(int64_t)lay.ee(i) * B + b];
            ewt[i] = ds[(int64_t)lay.ewt(i) * B + b];
            yy[i] = ds[(int64_t)lay.yy(i) * B + b];
            yp[i] = ds[(int64_t)lay.yp(i) * B + b];
            piv[i] = is[(int64_t)lay.piv(i) * B + b];
#pragma unroll
            for (int j = 0; j < N; ++j) J[i][j] = ds[(int64_t)lay.J(i, j) * B + b];
        }
#define LD_S(f, F) f = ds[(int64_t)lay.scal(L::F) * B + b];
        LD_S(hh, HH) LD_S(hused, HUSED) LD_S(rr, RR) LD_S(cj, CJ) LD_S(cjold, CJOLD) LD_S(cjlast, CJLAST)
        LD_S(cjratio, CJRATIO) LD_S(ss, SS) LD_S(toldel, TOLDEL) LD_S(eps_newt, EPS_NEWT) LD_S(tn, TN)
        LD_S(tretlast, TRETLAST) LD_S(h0u, H0U) LD_S(tolsf, TOLSF) LD_S(tret, TRET)
#undef LD_S
#define LD_I(f, F) f = is[(int64_t)L::F * B + b];
        LD_I(kk, KK) LD_I(kused, KUSED) LD_I(knew, KNEW) LD_I(phase, PHASE) LD_I(ns, NS) LD_I(setup_done, SETUP_DONE)
        LD_I(jcur, JCUR) LD_I(nst, NST) LD_I(nre, NRE) LD_I(nje, NJE) LD_I(nni, NNI) LD_I(netf, NETF)
        LD_I(ncfn, NCFN) LD_I(nsetups, NSETUPS) LD_I(nlscf, NLSCF) LD_I(natt, NATT)
#undef LD_I
    }

    __device__ __forceinline__ void store(double* __restrict__ ds, int* __restrict__ is, int64_t B, int64_t b) {
        constexpr Layout lay(N);
#pragma unroll 1
        for (int f = 0; f < SMEM_DOUBLES; ++f) ds[(int64_t)f * B + b] = sm[f * THREAD_BLOCK];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ds[(int64_t)lay.ee(i) * B + b] = ee[i];
            ds[(int64_t)lay.ewt(i) * B + b] = ewt[i];
            ds[(int64_t)lay.yy(i) * B + b] = yy[i];
            ds[(int64_t)lay.yp(i) * B + b] = yp[i];
            is[(int64_t)lay.piv(i) * B + b] = piv[i];
#pragma unroll
            for (int j = 0; j < N; ++j) ds[(int64_t)lay.J(i, j) * B + b] = J[i][j];
        }
#define ST_S(f, F) ds[(int64_t)lay.scal(L::F) * B + b] = f;
        ST_S(hh, HH) ST_S(hused, HUSED) ST_S(rr, RR) ST_S(cj, CJ) ST_S(cjold, CJOLD) ST_S(cjlast, CJLAST)
        ST_S(cjratio, CJRATIO) ST_S(ss, SS) ST_S(toldel, TOLDEL) ST_S(eps_newt, EPS_NEWT) ST_S(tn, TN)
        ST_S(tretlast, TRETLAST) ST_S(h0u, H0U) ST_S(tolsf, TOLSF) ST_S(tret, TRET)
#undef ST_S
#define ST_I(f, F) is[(int64_t)L::F * B + b] = f;
        ST_I(kk, KK) ST_I(kused, KUSED) ST_I(knew, KNEW) ST_I(phase, PHASE) ST_I(ns, NS) ST_I(setup_done, SETUP_DONE)
        ST_I(jcur, JCUR) ST_I(nst, NST) ST_I(nre, NRE) ST_I(nje, NJE) ST_I(nni, NNI) ST_I(netf, NETF)
        ST_I(ncfn, NCFN) ST_I(nsetups, NSETUPS) ST_I(nlscf, NLSCF) ST_I(natt, NATT)
#undef ST_I
    }

    // ---- TolControlSV::ewt_set: src/tol_control.rs:71-82 -------------------------------------------------
    __device__ __forceinline__ void ewt_set(const TolN<N>& tol) {
#if IDA_FASTDIV
        double den[N];
#pragma unroll
        for (int i = 0; i < N; ++i) den[i] = tol.rtol * fabs(phi(0, i)) + tol.atol[i];
        RK<N>(den, ewt);
#else
#pragma unroll
        for (int i = 0; i < N; ++i) ewt[i] = R(tol.rtol * fabs(phi(0, i)) + tol.atol[i]);
#endif
    }

    // ---- set_coeffs: src/lib.rs:722-782 ----------------------------------------------------------------
    // (the scaling of phi[ns..=kk] by beta, :772-779, is done by predict() below in the same pass that sums phi)
    __device__ __forceinline__ double set_coeffs() {
        if (hh != hused || kk != kused) ns = 0;
        ns = min(ns + 1, kused + 2);
        double alpha0 = 0.0;  // -(alpha[0] + ... + alpha[kk-1]) accumulated in index order, src/lib.rs:757-763
        if (kk + 1 >= ns) {
            beta(0) = 1.0;
            alpha(0) = 1.0;
            double temp1 = hh;
            gamma(0) = 0.0;
            sigma(0) = 1.0;
            double beta_prev = 1.0, alpha_prev = 1.0, sigma_prev = 1.0, gamma_prev = 0.0;
#pragma unroll 1
            for (int i = 1; i <= kk; ++i) {
                const double temp2 = psi(i - 1);
                psi(i - 1) = temp1;
                alpha0 -= alpha_prev;  // the alpha[i-1] this pass has in a register
#if IDA_FASTDIV
                // the three quotients of one pass are independent of each other: one basic block, one range test
                const double num[3] = {beta_prev * temp1, hh, alpha_prev};
                temp1 = temp2 + hh;
                const double den[3] = {temp2, temp1, hh};
                double quo[3];
                DK<3>(num, den, quo);
                const double b_i = quo[0], a_i = quo[1];
                const double s_i = (double)i * sigma_prev * a_i;
                const double g_i = gamma_prev + quo[2];
#else
                const double b_i = ddiv(beta_prev * temp1, temp2);
                temp1 = temp2 + hh;
                const double a_i = ddiv(hh, temp1);
                const double s_i = (double)i * sigma_prev * a_i;
                const double g_i = gamma_prev + ddiv(alpha_prev, hh);
#endif
                beta(i) = b_i;
                alpha(i) = a_i;
                sigma(i) = s_i;
                gamma(i) = g_i;
                beta_prev = b_i;
                alpha_prev = a_i;
                sigma_prev = s_i;
                gamma_prev = g_i;
            }
            psi(kk) = temp1;
        } else {
#pragma unroll 1
            for (int i = 0; i < kk; ++i) alpha0 -= alpha(i);
        }
        // alphas = -(1 + 1/2 + ... + 1/kk) accumulated in that order (src/lib.rs:757-763): a constant per kk,
        // so it comes from a table built with the same sequential IEEE operations at compile time.
        const double alphas = alphas_of(kk);
        cjlast = cj;
        cj = D(-alphas, hh);
        const double akk = alpha(kk);
        double ck = fabs(akk + alphas - alpha0);
        ck = fmax(ck, akk);
        return ck;
    }

    // ---- predict: src/lib.rs:894-959, fused with the tail of set_coeffs (src/lib.rs:772-779): phi[j] *= beta[j] for
    // j = ns..kk happens in the pass that reads phi[j] anyway (phi[0] is never scaled: ns >= 1) ---------------------------
    __device__ __forceinline__ void predict() {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            yypredict[i] = 0.0 + phi(0, i);
            yppredict[i] = 0.0;
        }
#pragma unroll 1
        for (int j = 1; j <= kk; ++j) {
            const double gj = gamma(j), bj = beta(j);
            const bool scale = j >= ns;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double p = phi(j, i);
                if (scale) {
                    p *= bj;
                    phi(j, i) = p;
                }
                yypredict[i] += p;
                yppredict[i] += gj * p;
            }
        }
    }

    // ---- dense_get_rf: crates/linear/src/dense.rs:86-158 ; returns 0 or 1-based zero-pivot column -------
    __device__ __forceinline__ int getrf() {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            int l = k;
            double best = fabs(J[k][k]);
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const double v = fabs(J[i][k]);
                if (v > best) {
                    best = v;
                    l = i;
                }
            }
            piv[k] = l;
            if (best == 0.0) return k + 1;  // |a_lk| == 0  <=>  a_lk == 0 (NaN never wins the > test)
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                if (l == i) {
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        const double t = J[k][c];
                        J[k][c] = J[i][c];
                        J[i][c] = t;
                    }
                }
            }
            const double mult = R(J[k][k]);
#pragma unroll
            for (int i = k + 1; i < N; ++i) J[i][k] *= mult;
#pragma unroll
            for (int j = k + 1; j < N; ++j) {
                const double a_kj = J[k][j];
                if (a_kj != 0.0) {
#pragma unroll
                    for (int i = k + 1; i < N; ++i) J[i][j] -= a_kj * J[i][k];
                }
            }
        }
        return 0;
    }

    // ---- dense_get_rs: crates/linear/src/dense.rs:165-206 ---------------------------------------------
    __device__ __forceinline__ void getrs(double (&b)[N]) {
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
            for (int i = k + 1; i < N; ++i) b[i] -= J[i][k] * bk;
        }
#if IDA_FASTDIV
        // back substitution with the N quotients on their fast paths in one basic block: the reciprocal refinement of
        // every u_kk is independent of b and is scheduled ahead of the chain; one range test for all of them, and a
        // cold redo from the saved right-hand side
        double save[N];
#pragma unroll
        for (int i = 0; i < N; ++i) save[i] = b[i];
        bool all = true;
#pragma unroll
        for (int k = N - 1; k >= 1; --k) {
            bool ok;
            b[k] = div_fast_z(b[k], J[k][k], ok);
            all = all && ok;
            const double bk = b[k];
#pragma unroll
            for (int i = 0; i < k; ++i) b[i] -= J[i][k] * bk;
        }
        {
            bool ok;
            b[0] = div_fast_z(b[0], J[0][0], ok);
            all = all && ok;
        }
        if constexpr (FAST) IDA_POISON(256u, !all);
        if (!FAST && !IDA_UNSAFE_NOFALLBACK && __builtin_expect(!all, 0)) {
#pragma unroll
            for (int i = 0; i < N; ++i) b[i] = save[i];
#pragma unroll
            for (int k = N - 1; k >= 1; --k) {
                b[k] = div_slow(b[k], J[k][k]);
                const double bk = b[k];
#pragma unroll
                for (int i = 0; i < k; ++i) b[i] -= J[i][k] * bk;
            }
            b[0] = div_slow(b[0], J[0][0]);
        }
#else
#pragma unroll
        for (int k = N - 1; k >= 1; --k) {
            b[k] = ddiv(b[k], J[k][k]);
            const double bk = b[k];
#pragma unroll
            for (int i = 0; i < k; ++i) b[i] -= J[i][k] * bk;
        }
        b[0] = ddiv(b[0], J[0][0]);
#endif
    }

    // ---- IdaNLProblem::sys: src/ida_nls.rs:118-153 ------------------------------------------------------
    __device__ __forceinline__ void nls_sys(const double (&ycor)[N], double (&res)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            yy[i] = yypredict[i] + ycor[i];
            yp[i] = yppredict[i] + cj * ycor[i];
        }
        prob.res(tn, yy, yp, res);
        nre += 1;
    }

    // ---- IdaNLProblem::setup -> IdaLProblem::setup: src/ida_nls.rs:156-187, src/ida_ls.rs:232-290 --------
    __device__ __forceinline__ int nls_setup(const double (&res)[N]) {
        nsetups += 1;
        nje += 1;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) J[i][j] = 0.0;
        prob.jac(0.0, cj, yy, yp, res, J);  // tt = 0 as in src/ida_ls.rs:260-261
        const int info = getrf();
        if (info != 0) return info;
        cjold = cj;
        cjratio = 1.0;
        ss = 20.0;
        return 0;
    }

    // ---- IdaNLProblem::solve -> IdaLProblem::solve: src/ida_nls.rs:190-215, src/ida_ls.rs:298-455 --------
    __device__ __forceinline__ void nls_lsolve(double (&b)[N]) {
        getrs(b);
        if (cjratio != 1.0) {
            const double f = D(2.0, 1.0 + cjratio);
#pragma unroll
            for (int i = 0; i < N; ++i) b[i] *= f;
        }
    }

    // ---- IdaNLProblem::ctest: src/ida_nls.rs:218-266 ; 1 converged, 0 continue, -1 recover ----------------
    __device__ __forceinline__ int nls_ctest(int m, const double (&del)[N], double tol) {
        const double delnrm = norm_tail(wrms_acc<N>(del, ewt));
        if (m == 0) {
            oldnrm = delnrm;
            if (delnrm <= 0.0001 * toldel) return 1;
        } else {
            const double base = DZ(delnrm, oldnrm);
            const double arg = inv_small(m);  // 1.0 / m, m = 1..3 (MAXNLSIT = 4)
            // m == 1 (the second Newton iteration, by far the most frequent case): pow(x, 1.0) returns x bit-for-bit
            // in the reference's libm (tests/test_detpow.py::test_pow_with_exponent_one_is_the_identity), so the
            // ~60-instruction pow is skipped for the lanes that do not need it
            const double rate = (m == 1) ? base : PW(base, arg);
            if (rate > RATEMAX) return -1;
            ss = DZ(rate, 1.0 - rate);
        }
        if (ss * delnrm <= tol) return 1;
        return 0;
    }

    // ---- nonlinear_solve + Newton::solve: src/lib.rs:787-890, crates/nonlinear/src/newton.rs:51-167 -------
    // returns 0 ok, 1 recoverable convergence failure, B200IDA_ST_LSETUP_FAIL on a singular Jacobian.
    // (C semantics for "recover while the Jacobian is current": fail the attempt; DESIGN.md "Deviations".)
    __device__ __forceinline__ int nonlinear_solve() {
        bool call_lsetup = false;
        if (nst == 0) {
            cjold = cj;
            ss = 20.0;
            call_lsetup = true;
        }
        cjratio = D(cj, cjold);
        const double temp1 = (1.0 - XRATE) / (1.0 + XRATE);
        const double temp2 = 1.0 / temp1;
        if (cjratio < temp1 || cjratio > temp2) call_lsetup = true;
        if (cj != cjlast) ss = 100.0;

        double nd[N];  // Newton.delta
        int ret = 0;
#pragma unroll 1
        for (;;) {  // at most two passes: the second with a fresh Jacobian
#pragma unroll
            for (int i = 0; i < N; ++i) delta[i] = 0.0;
            nls_sys(delta, nd);
            if (call_lsetup) {
                if (nls_setup(nd) != 0) {
                    ret = B200IDA_ST_LSETUP_FAIL;
                    break;
                }
                jcur = 1;
            }
            int curiter = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) ee[i] = 0.0;  // y = y0 = 0
#pragma unroll 1
            for (;;) {
                nni += 1;
#pragma unroll
                for (int i = 0; i < N; ++i) nd[i] = -nd[i];
                nls_lsolve(nd);
#pragma unroll
                for (int i = 0; i < N; ++i) ee[i] += nd[i];
                const int c = nls_ctest(curiter, nd, eps_newt);
                if (c == 1) {
                    jcur = 0;
                    ret = 0;
                    break;
                }
                if (c == -1) {
                    ret = 1;
                    break;
                }
                curiter += 1;
                if (curiter >= MAXNLSIT) {
                    ret = 1;
                    break;
                }
                nls_sys(ee, nd);
            }
            if (ret == 1 && !jcur) {
                nlscf += 1;
                call_lsetup = true;
                continue;
            }
            break;
        }
        if (ret != 0) nlscf += 1;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            yy[i] = yypredict[i] + ee[i];
            yp[i] = yppredict[i] + cj * ee[i];
        }
        return ret;
    }

    // ---- test_error: src/lib.rs:967-1039 -------------------------------------------------------------------
    // The up-to-three norms (of ee, ee + phi[kk], ee + phi[kk] + phi[kk-1]) are formed for every lane and finished by ONE
    // out-of-line call whose three division / square-root chains interleave; a lane whose order does not use the
    // second or third norm ignores it (same operands and operations wherever a value is used: bit-identical).
    __device__ __forceinline__ bool test_error(double ck, double& err_k, double& err_km1) {
        const double scalar_kk = (double)kk;
        const int km1 = max(kk - 1, 0), km2 = max(kk - 2, 0);
        double d2[N];
#pragma unroll
        for (int i = 0; i < N; ++i) delta[i] = phi(kk, i) + ee[i];
#pragma unroll
        for (int i = 0; i < N; ++i) d2[i] = delta[i] + phi(km1, i);
        double enorm_k, enorm_km1, enorm_km2;
        if constexpr (IDA_WRMS_INLINE || FAST) {
            enorm_k = err_norm(ee);
            enorm_km1 = err_norm(delta);
            enorm_km2 = err_norm(d2);
        } else {
            wrms_tail3<N>(err_acc(ee), err_acc(delta), err_acc(d2), enorm_k, enorm_km1, enorm_km2);
        }
        err_k = sigma(kk) * enorm_k;
        const double terr_k = err_k * (scalar_kk + 1.0);
        err_km1 = 0.0;
        int knew_ = kk;
        if (kk > 1) {
            err_km1 = sigma(km1) * enorm_km1;
            const double terr_km1 = scalar_kk * err_km1;
            if (kk > 2) {
                const double err_km2 = sigma(km2) * enorm_km2;
                const double terr_km2 = (scalar_kk - 1.0) * err_km2;
                if (fmax(terr_km1, terr_km2) <= terr_k) knew_ = kk - 1;
            } else {
                if (terr_km1 <= terr_k * 0.5) knew_ = kk - 1;
            }
        }
        knew = knew_;
        return (ck * enorm_k) <= 1.0;
    }

    // ---- restore: src/lib.rs:1044-1083 -----------------------------------------------------------------------
    __device__ __forceinline__ void restore(double saved_t) {
        tn = saved_t;
#pragma unroll 1
        for (int j = 1; j <= kk; ++j) psi(j - 1) = psi(j) - hh;
#pragma unroll 1
        for (int j = ns; j <= kk; ++j) {
            const double c = R(beta(j));
#pragma unroll
            for (int i = 0; i < N; ++i) phi(j, i) *= c;
        }
    }

    // ---- handle_n_flag: src/lib.rs:1120-1244 ; nflag 2 = error-test failure, 1 = convergence failure ----------
    __device__ __forceinline__ int handle_n_flag(int nflag, double err_k, double err_km1, int& ncf, int& nef) {
        phase = 1;
        if (nflag == 2) {
            nef += 1;
            netf += 1;
            if (nef == 1) {
                const double err_knew = (kk == knew) ? err_k : err_km1;
                kk = knew;
                const double base = 2.0 * err_knew + 0.0001;
                const double arg = inv_small(kk + 1);
                rr = 0.9 * PW(base, -arg);
                rr = fmax(0.25, fmin(0.9, rr));
                hh *= rr;
                return 0;
            } else if (nef == 2) {
                kk = knew;
                rr = 0.25;
                hh *= rr;
                return 0;
            } else if (nef < MXNEF) {
                kk = 1;
                rr = 0.25;
                hh *= rr;
                return 0;
            }
            return B200IDA_ST_ERR_FAIL;
        }
        ncf += 1;
        ncfn += 1;
        rr = 0.25;
        hh *= rr;
        if (ncf < MXNCF) return 0;
        return B200IDA_ST_CONV_FAIL;
    }

    // ---- reset: src/lib.rs:1249-1252 (scales ALL of phi, like the reference) ---------------------------------
    __device__ __forceinline__ void reset() {
        psi(0) = hh;
#pragma unroll 1
        for (int f = 0; f < MXORDP1 * N; ++f) sm[f * THREAD_BLOCK] *= rr;
    }

    // ---- complete_step: src/impl_complete_step.rs:22-177 -------------------------------------------------------
    __device__ __forceinline__ void complete_step(double err_k, double err_km1) {
        nst += 1;
        const int kdiff = kk - kused;
        kused = kk;
        hused = hh;
        if (knew == kk - 1 || kk == MAXORD) phase = 1;
        if (phase == 0) {
            if (nst > 1) {
                kk += 1;
                double hnew = 2.0 * hh;
                if constexpr (HMAX_INV != 0.0) {  // no hmax setter exists: |h| * 0 > 1 is never true, and nvcc would
                    const double t = fabs(hnew) * HMAX_INV;  // otherwise evaluate the division speculatively (by zero:
                    if (t > 1.0) hnew /= t;                  // the slow path, 4 % of the round-1 kernel's samples)
                }
                hh = hnew;
            }
        } else {
            int action;  // 0 lower, 1 maintain, 2 raise
            double err_kp1 = 0.0;
            if (knew == kk - 1) {
                action = 0;
            } else if (kk == MAXORD) {
                action = 1;
            } else if (kk + 1 >= ns || kdiff == 1) {
                action = 1;
            } else {
                double t[N];
#pragma unroll
                for (int i = 0; i < N; ++i) t[i] = ee[i] - phi(kk + 1, i);
                const double enorm = err_norm(t);
                err_kp1 = D(enorm, (double)(kk + 2));
                const double terr_k = (double)(kk + 1) * err_k;
                const double terr_kp1 = (double)(kk + 2) * err_kp1;
                if (kk == 1) {
                    action = (terr_kp1 >= 0.5 * terr_k) ? 1 : 2;
                } else {
                    const double terr_km1 = (double)kk * err_km1;
                    if (terr_km1 <= fmin(terr_k, terr_kp1)) action = 0;
                    else if (terr_kp1 >= terr_k) action = 1;
                    else action = 2;
                }
            }
            double err_knew;
            if (action == 2) {
                kk += 1;
                err_knew = err_kp1;
            } else if (action == 0) {
                kk -= 1;
                err_knew = err_km1;
            } else {
                err_knew = err_k;
            }
            double hnew = hh;
            {
                const double base = 2.0 * err_knew + 0.0001;
                const double arg = -inv_small(kk + 1);
                rr = PW(base, arg);
            }
            if (rr >= 2.0) {
                hnew = 2.0 * hh;
                if constexpr (HMAX_INV != 0.0) {
                    const double t = fabs(hnew) * HMAX_INV;
                    if (t > 1.0) hnew /= t;
                }
            } else if (rr <= 1.0) {
                rr = fmax(0.5, fmin(rr, 0.9));
                hnew = hh * rr;
            }
            hh = hnew;
        }
        if (kused < MAXORD) {
#pragma unroll
            for (int i = 0; i < N; ++i) phi(kused + 1, i) = ee[i];
        }
        double tmp[N];
#pragma unroll
        for (int i = 0; i < N; ++i) tmp[i] = ee[i];
#pragma unroll 1
        for (int j = kused; j >= 0; --j) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                tmp[i] += phi(j, i);
                phi(j, i) = tmp[i];
            }
        }
    }

    // ---- get_solution: src/lib.rs:1274-1343 (out-of-line body: get_solution_fn below the class) -----------------
    __device__ __forceinline__ int get_solution(double t) {
        double out[2 * N];
        const int r = get_solution_fn<N>(sm, hh, hused, tn, kused, t, out);
        if (r == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                yy[i] = out[i];
                yp[i] = out[N + i];
            }
        }
        return r;
    }

    // ---- rootfinding: src/impl_r_check.rs:32-576 (state src/lib.rs:198-229).  Only instantiated by the ROOTS kernels. --
    // The 4 NR + 5 doubles of root state live in the thread's shared-memory column behind the solver's arrays (the
    // ROOTS kernel is launched with SMEM_DOUBLES + ROOT_DOUBLES words per thread): in registers they pushed the kernel
    // over its 168-register budget and the spills cost 2.5x.
    static constexpr int NR = P::NR > 0 ? P::NR : 1;
    static constexpr int ROOT_DOUBLES = 4 * NR + 5;
    __device__ __forceinline__ double& rsm(int k) { return sm[(SMEM_DOUBLES + k) * THREAD_BLOCK]; }
    __device__ __forceinline__ const double& rsm(int k) const { return sm[(SMEM_DOUBLES + k) * THREAD_BLOCK]; }
    bool gactive[NR];  // FALSE at start like the reference (src/lib.rs:373)
    bool irfnd;
    int nge;
    int evc;  // events logged so far (rp.ev_count): kept in a register and written back with the root state, so that a
              // lane the FAST pass abandons leaves no trace for the exact pass that redoes it
    __device__ __forceinline__ static double signum(double x) { return (x != x) ? x : (signbit(x) ? -1.0 : 1.0); }
    enum { GLO = 0, GHI = 1, GROUT = 2 };  // which array of the root state
    __device__ __forceinline__ void eval_root(double t, int which) {  // g(t, yy, yp) -> array `which`
        if constexpr (P::NR > 0) {
            double g[NR];
            prob.root(t, yy, yp, g);
#pragma unroll
            for (int i = 0; i < NR; ++i) rsm(which * NR + i) = g[i];
        }
    }
    __device__ __forceinline__ void load_roots(const RootParams& rp, int64_t B, int64_t b) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            rsm(0 * NR + (i)) = rp.rd[(int64_t)i * B + b];
            rsm(1 * NR + (i)) = rp.rd[(int64_t)(NR + i) * B + b];
            rsm(2 * NR + (i)) = rp.rd[(int64_t)(2 * NR + i) * B + b];
            rsm(3 * NR + (i)) = rp.rd[(int64_t)(3 * NR + i) * B + b];
            gactive[i] = rp.ri[(int64_t)i * B + b] != 0;
        }
        rsm(4 * NR + 0) = rp.rd[(int64_t)(4 * NR) * B + b];
        rsm(4 * NR + 1) = rp.rd[(int64_t)(4 * NR + 1) * B + b];
        rsm(4 * NR + 2) = rp.rd[(int64_t)(4 * NR + 2) * B + b];
        rsm(4 * NR + 3) = rp.rd[(int64_t)(4 * NR + 3) * B + b];
        rsm(4 * NR + 4) = rp.rd[(int64_t)(4 * NR + 4) * B + b];
        irfnd = rp.ri[(int64_t)NR * B + b] != 0;
        nge = rp.ri[(int64_t)(NR + 1) * B + b];
    }
    __device__ __forceinline__ void load_event_count(const RootParams& rp, int64_t ob) { evc = rp.ev_count[ob]; }
    __device__ __forceinline__ void store_event_count(const RootParams& rp, int64_t ob) const { rp.ev_count[ob] = evc; }
    __device__ __forceinline__ void store_roots(const RootParams& rp, int64_t B, int64_t b) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            rp.rd[(int64_t)i * B + b] = rsm(0 * NR + (i));
            rp.rd[(int64_t)(NR + i) * B + b] = rsm(1 * NR + (i));
            rp.rd[(int64_t)(2 * NR + i) * B + b] = rsm(2 * NR + (i));
            rp.rd[(int64_t)(3 * NR + i) * B + b] = rsm(3 * NR + (i));
            rp.ri[(int64_t)i * B + b] = gactive[i] ? 1 : 0;
        }
        rp.rd[(int64_t)(4 * NR) * B + b] = rsm(4 * NR + 0);
        rp.rd[(int64_t)(4 * NR + 1) * B + b] = rsm(4 * NR + 1);
        rp.rd[(int64_t)(4 * NR + 2) * B + b] = rsm(4 * NR + 2);
        rp.rd[(int64_t)(4 * NR + 3) * B + b] = rsm(4 * NR + 3);
        rp.rd[(int64_t)(4 * NR + 4) * B + b] = rsm(4 * NR + 4);
        rp.ri[(int64_t)NR * B + b] = irfnd ? 1 : 0;
        rp.ri[(int64_t)(NR + 1) * B + b] = nge;
    }

    // IDARcheck1 -- src/impl_r_check.rs:32-105 (called before phi[1] is scaled by hh)
    __device__ __forceinline__ void r_check1() {
#pragma unroll
        for (int i = 0; i < NR; ++i) rsm(3 * NR + (i)) = 0.0;
        rsm(4 * NR + 0) = tn;
        rsm(4 * NR + 3) = (fabs(tn) + fabs(hh)) * DBL_EPS * 100.0;
        double y0[N], y1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            y0[i] = phi(0, i);
            y1[i] = phi(1, i);
        }
        if constexpr (P::NR > 0) {
            double g[NR];
            prob.root(rsm(4 * NR + 0), y0, y1, g);
#pragma unroll
            for (int i = 0; i < NR; ++i) rsm(GLO * NR + i) = g[i];
        }
        nge = 1;
        bool zroot = false;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (fabs(rsm(0 * NR + (i))) == 0.0) {
                gactive[i] = false;
                zroot = true;
            }
        }
        if (zroot) {
            const double hratio = fmax(rsm(4 * NR + 3) / fabs(hh), 0.1);
            const double smallh = hratio * hh;
            const double tplus = rsm(4 * NR + 0) + smallh;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                yy[i] = y0[i];
                yy[i] += smallh * y1[i];
            }
            if constexpr (P::NR > 0) {
                double g[NR];
                prob.root(tplus, yy, y1, g);
#pragma unroll
                for (int i = 0; i < NR; ++i) rsm(GHI * NR + i) = g[i];
            }
            nge += 1;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                if (!gactive[i] && fabs(rsm(1 * NR + (i))) != 0.0) {
                    gactive[i] = true;
                    rsm(0 * NR + (i)) = rsm(1 * NR + (i));
                }
            }
        }
    }

    // IDARcheck2 -- src/impl_r_check.rs:123-209 ; 0 continue, 1 root found, B200IDA_ST_CLOSE_ROOTS
    __device__ __forceinline__ int r_check2() {
        if (!irfnd) return 0;
        get_solution(rsm(4 * NR + 0));
        eval_root(rsm(4 * NR + 0), GLO);
        nge += 1;
        bool zroot = false;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            rsm(3 * NR + (i)) = 0.0;
            if (gactive[i] && fabs(rsm(0 * NR + (i))) == 0.0) {
                zroot = true;
                rsm(3 * NR + (i)) = 1.0;
            }
        }
        if (zroot) {
            rsm(4 * NR + 3) = (fabs(tn) + fabs(hh)) * DBL_EPS * 100.0;
            const double sgn = (hh > 0.0 || (hh == 0.0 && !signbit(hh))) ? 1.0 : -1.0;
            const double smallh = rsm(4 * NR + 3) * sgn;
            const double tplus = rsm(4 * NR + 0) + smallh;
            if ((tplus - tn) * hh >= 0.0) {
                const double hratio = smallh / hh;
#pragma unroll
                for (int i = 0; i < N; ++i) yy[i] += hratio * phi(1, i);
            } else {
                get_solution(tplus);
            }
            eval_root(tplus, GHI);
            nge += 1;
            bool z2 = false;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                if (gactive[i]) {
                    if (fabs(rsm(1 * NR + (i))) == 0.0) {
                        if (rsm(3 * NR + (i)) > 0.0) return B200IDA_ST_CLOSE_ROOTS;
                        z2 = true;
                        rsm(3 * NR + (i)) = 1.0;
                    } else if (rsm(3 * NR + (i)) > 0.0) {
                        rsm(0 * NR + (i)) = rsm(1 * NR + (i));
                    }
                }
            }
            if (z2) return 1;
        }
        return 0;
    }

    // IDARcheck3 -- src/impl_r_check.rs:211-267 ; 0 continue, 1 root found (body: root_check3 above, out of line)
    __device__ __forceinline__ int r_check3() {
        // Fast path, inline: the check after an ordinary step (thi = t_n, g(t_n) from phi[0], see root_check3) when no
        // active g_i changes sign or vanishes -- all but ~2 of the ~370 checks of a Roberts instance.  It performs
        // exactly the state updates of root_check3's no-root path; anything else (sign change, zero, NaN, output time
        // inside the step) takes the out-of-line function, which starts over from the unmodified state.
        if constexpr (!P::ROOT_USES_YP) {
            if ((rsm(4 * NR + 4) - tn) * hh >= 0.0) {
                double y[N], ypz[N], g[NR];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    y[i] = 0.0 + phi(0, i);
                    ypz[i] = 0.0;
                }
                if constexpr (P::NR > 0) prob.root(tn, y, ypz, g);
                bool quiet = true;
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const double glo_i = rsm(0 * NR + i);
                    const bool live = gactive[i] && (0.0 * glo_i <= 0.0);
                    if (live && (fabs(g[i]) == 0.0 || glo_i * g[i] < 0.0)) quiet = false;
                }
                if (quiet) {
                    rsm(4 * NR + 1) = tn;                                    // thi
                    rsm(4 * NR + 3) = (fabs(tn) + fabs(hh)) * DBL_EPS * 100.0;  // ttol
                    rsm(4 * NR + 2) = tn;                                    // trout
                    rsm(4 * NR + 0) = tn;                                    // tlo
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        rsm(1 * NR + i) = g[i];  // ghi
                        rsm(2 * NR + i) = g[i];  // grout
                        rsm(0 * NR + i) = g[i];  // glo
                        if (!gactive[i] && g[i] != 0.0) gactive[i] = true;
                    }
                    nge += 1;
                    return 0;
                }
            }
        }
        unsigned gmask = 0;
#pragma unroll
        for (int i = 0; i < NR; ++i) gmask |= gactive[i] ? (1u << i) : 0u;
        const unsigned r = root_check3<P>(sm, prob, gmask, tn, hh, hused, kused);
#pragma unroll
        for (int i = 0; i < NR; ++i) gactive[i] = ((r >> (1 + i)) & 1u) != 0;
        nge += (int)(r >> 8);
        if (r & 1u) get_solution(rsm(4 * NR + 2));  // y(trout), y'(trout)
        return (int)(r & 1u);
    }

    // ---- first-call block of solve(): src/impl_solve.rs:84-173 ; returns 0 or B200IDA_ST_ILL_INPUT ---------------
    template <bool ROOTS>
    __device__ __forceinline__ int first_call(double tout, const TolN<N>& tol) {
        if (!setup_done) {
            ewt_set(tol);  // initial_setup, src/lib.rs:537-545
            setup_done = 1;
        }
        const double tdist = fabs(tout - tn);
        if (tdist == 0.0) return B200IDA_ST_ILL_INPUT;
        const double troundoff = 2.0 * DBL_EPS * (fabs(tn) + fabs(tout));
        if (tdist < troundoff) return B200IDA_ST_ILL_INPUT;
        hh = 0.001 * tdist;  // hin = 0: no setter exists in the reference
        double p1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) p1[i] = phi(1, i);
        const double ypnorm = err_norm(p1);
        if (ypnorm > D(2.0, hh)) hh = D(0.5, ypnorm);
        if (tout < tn) hh = -hh;
        if constexpr (HMAX_INV != 0.0) {
            const double rh = fabs(hh) * HMAX_INV;
            if (rh > 1.0) hh /= rh;
        }
        h0u = hh;
        kk = 0;
        kused = 0;
        if constexpr (ROOTS) r_check1();  // src/impl_solve.rs:161-164, before phi[1] is scaled
#pragma unroll
        for (int i = 0; i < N; ++i) phi(1, i) = p1[i] * hh;
        eps_newt = EPCON;
        toldel = 0.0001 * eps_newt;
        return 0;
    }

    // ---- the flattened solve loop over a ladder of output times ---------------------------------------------------
    // Equivalent to calling Ida::solve(touts[k], .., Normal) for k = 0..ntout-1 (src/impl_solve.rs:69-376 with the
    // NORMAL branches of stop_test1/2, src/impl_stop_test.rs:53-65,152-161).  After call k, yy is written to
    // yy_traj[k][.][b] and the call's status to status[k][b] when those pointers are non-null.
    // DEFER: y(tout) of a call that returns through stop_test2 is not interpolated here but snapshotted for
    // ida_interp_kernel (DeferParams in ida_common.cuh); every other return path interpolates in line.
    template <bool TRACE, bool ROOTS = false, bool DEFER = false>
    __device__ __forceinline__ void solve_ladder(const LadderParams& lad, const TolN<N>& tol, double* __restrict__ yy_traj,
                                                 int32_t* __restrict__ status, int64_t B, int64_t b, const TraceParams& tr,
                                                 const RootParams& rp = RootParams(), const DeferParams& df = DeferParams(), int64_t slot = 0,
                                                 const SegParams& sg = SegParams()) {
        int it = 0;
        bool deferred = false;
        bool need_prologue = true;
        bool in_step = false;
        bool call_done = false;
        int call_status = 0;
        int nstloc = 0;
        double saved_t = 0.0, tout = 0.0;
        int ncf = 0, nef = 0;
        int trips_left = sg.max_trips < 0 ? 0x7fffffff : sg.max_trips;
        if (sg.max_trips >= 0 && sg.resume && sg.li) {  // continue the API call where the previous segment's launch parked it
            const int w0 = sg.li[slot];
            it = w0 & 0xff;
            need_prologue = (w0 >> 8) & 1;
            in_step = (w0 >> 9) & 1;
            ncf = (w0 >> 10) & 0xff;
            nef = (w0 >> 18) & 0xff;
            nstloc = sg.li[sg.stride + slot];
            saved_t = sg.ld[slot];
            if (it < lad.ntout) tout = lad.touts[it];
        }

#pragma unroll 1
        while (it < lad.ntout && !(FAST && poison) && trips_left > 0) {
            trips_left -= 1;
            RP_T(ti0);
            if (need_prologue) {
                need_prologue = false;
                tout = lad.touts[it];
                nstloc = 0;
                in_step = false;
                call_done = false;
                call_status = 0;
                deferred = false;
                if constexpr (ROOTS) rsm(4 * NR + 4) = tout;  // src/impl_solve.rs:78-82
                if (nst == 0) {
                    const int r = first_call<ROOTS>(tout, tol);
                    if (r != 0) {
                        if constexpr (FAST) IDA_POISON(128u, true);  // rare: the exact instantiation reports it
                        call_done = true;
                        call_status = r;
                    }
                } else {
                    if constexpr (ROOTS) {  // root stop checks: src/impl_solve.rs:186-227
                        RP_T(tp0);
                        const int ier2 = r_check2();
                        if (ier2 != 0) {
                            if (ier2 == 1) {
                                tretlast = rsm(4 * NR + 0);
                                tret = rsm(4 * NR + 0);
                            }
                            call_status = (ier2 == 1) ? B200IDA_ST_ROOT : ier2;
                            call_done = true;
                        } else {
                            const double troundoff = (fabs(tn) + fabs(hh)) * DBL_EPS * 100.0;
                            if (fabs(tn - tretlast) > troundoff) {
                                if (r_check3() == 1) {
                                    irfnd = true;
                                    tretlast = rsm(4 * NR + 0);
                                    tret = rsm(4 * NR + 0);
                                    call_status = B200IDA_ST_ROOT;
                                    call_done = true;
                                } else {
                                    irfnd = false;
                                }
                            }
                        }
                        RP_ACC(0, tp0);
                    }
                    // stop_test1, NORMAL: src/impl_stop_test.rs:53-65
                    if (call_done) {
                    } else if (tout == tretlast) {
                        tretlast = tout;
                        tret = tout;
                        call_done = true;
                    } else if ((tn - tout) * hh >= 0.0) {
                        if constexpr (FAST) {
                            IDA_POISON(128u, true);  // output time already passed: interpolated in line by the exact instantiation
                            call_done = true;
                        } else {
                            const int ier = get_solution(tout);
                            if (ier == 0) {
                                tretlast = tout;
                                tret = tout;
                            }
                            call_status = ier;
                            call_done = true;
                        }
                    }
                }
            }
            if (!call_done && !in_step) {
                // top of the step loop: src/impl_solve.rs:246-311
                if (nstloc >= MXSTEP) {
                    if constexpr (FAST) IDA_POISON(128u, true);
                    if constexpr (ROOTS) get_solution(tn);  // what the reference's r_check3 left in yy / yp (see r_check3)
                    tret = tn;
                    tretlast = tn;
                    call_status = B200IDA_ST_TOO_MUCH_WORK;
                    call_done = true;
                } else {
                    bool bad = false;
                    if (nst > 0) {
                        ewt_set(tol);
#pragma unroll
                        for (int i = 0; i < N; ++i) bad = bad || (ewt[i] <= 0.0);
                    }
                    if (bad) {
                        if constexpr (FAST) IDA_POISON(128u, true);
                        else get_solution(tn);
                        tret = tn;
                        tretlast = tn;
                        call_status = B200IDA_ST_BAD_EWT;
                        call_done = true;
                    } else {
                        double p0[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) p0[i] = phi(0, i);
                        const double nrm = err_norm(p0);
                        tolsf = DBL_EPS * nrm;
                        if (tolsf > 1.0) {
                            tolsf *= 10.0;
                            tret = tn;
                            tretlast = tn;
                            if constexpr (FAST) IDA_POISON(128u, true);
                            else if (nst > 0) get_solution(tn);
                            call_status = B200IDA_ST_TOO_MUCH_ACC;
                            call_done = true;
                        } else {
                            // step() entry: src/lib.rs:617-631
                            saved_t = tn;
                            if (nst == 0) {
                                kk = 1;
                                kused = 0;
                                hused = 0.0;
                                psi(0) = hh;
                                cj = R(hh);
                                phase = 0;
                                ns = 0;
                            }
                            ncf = 0;
                            nef = 0;
                            in_step = true;
                        }
                    }
                }
            }
            if (!call_done) {
                // one step attempt: src/lib.rs:634-694
                if constexpr (TRACE) {  // data_trace frame: the whole state at the top of the attempt (src/lib.rs:635-639)
                    if (b == (int64_t)tr.id) {
                        constexpr Layout lay(N);
                        const int c = *tr.count;
                        if (c < tr.max) store(tr.d + (size_t)c * lay.ndouble(), tr.i + (size_t)c * lay.nint(), 1, 0);
                        *tr.count = c + 1;
                    }
                }
                natt += 1;
                const double ck = set_coeffs();
                tn += hh;
                predict();
                int nflag = nonlinear_solve();
                double err_k = 0.0, err_km1 = 0.0;
                bool converged = false;
                if (nflag == 0) {
                    converged = test_error(ck, err_k, err_km1);
                    if (!converged) nflag = 2;
                }
                if (converged) {
                    complete_step(err_k, err_km1);
#pragma unroll
                    for (int i = 0; i < N; ++i) ee[i] *= ck;  // src/lib.rs:708
                    in_step = false;
                    nstloc += 1;
                    bool root_ret = false;
                    if constexpr (ROOTS) {  // root in the last step taken: src/impl_solve.rs:335-344
                        RP_T(td0);
                        const int rc3 = r_check3();
                        RP_ACC(rc3 == 1 ? 2 : 1, td0);
                        if (rc3 == 1) {
                            irfnd = true;
                            tretlast = rsm(4 * NR + 0);
                            tret = rsm(4 * NR + 0);
                            call_status = B200IDA_ST_ROOT;
                            call_done = true;
                            root_ret = true;
                        }
                    }
                    // stop_test2, NORMAL: src/impl_stop_test.rs:152-161
                    if (!root_ret && (tn - tout) * hh >= 0.0) {
                        tret = tout;
                        tretlast = tout;
                        if constexpr (DEFER) {
                            // what get_solution(tout) reads, to row (instance id, it) of the snapshot buffer
                            double* sp = df.snap + ((size_t)b * lad.ntout + it) * snap_doubles(N);
                            const int kord = (kused == 0) ? 1 : kused;
#pragma unroll 1
                            for (int j = 0; j <= kord; ++j) {
#pragma unroll
                                for (int i = 0; i < N; ++i) sp[j * N + i] = phi(j, i);
                            }
#pragma unroll 1
                            for (int j = 0; j < kord; ++j) sp[6 * N + j] = psi(j);
                            sp[6 * N + 6] = (double)kused;
                            sp[6 * N + 7] = tn;
                            sp[6 * N + 8] = hh;
                            sp[6 * N + 9] = hused;
                            deferred = true;
                        } else {
                            get_solution(tout);  // TRACE / ROOTS / MASK: in line (exact, out-of-line body)
                        }
                        call_status = B200IDA_ST_SUCCESS;
                        call_done = true;
                    }
                } else {
                    restore(saved_t);
                    int kflag;
                    if (nflag < 0) {
                        ncfn += 1;
                        kflag = nflag;
                    } else {
                        kflag = handle_n_flag(nflag, err_k, err_km1, ncf, nef);
                    }
                    if (kflag != 0) {
                        // failed step: src/impl_solve.rs:314-327
                        if constexpr (FAST) {
                            IDA_POISON(512u, true);
                        } else if (get_solution(tn) == 0) {
                            tret = tn;
                            tretlast = tn;
                        }
                        call_status = kflag;
                        call_done = true;
                        in_step = false;
                    } else if (nst == 0) {
                        reset();
                    }
                }
            }
            if constexpr (ROOTS) {
                if (call_done && call_status == B200IDA_ST_ROOT) {
                    // the caller of examples/roberts.rs:117-140: log (t, get_root_info) and call solve() again
                    const int c = evc;
                    if (c < rp.max_ev) {
                        rp.ev_t[(int64_t)c * B + b] = tret;
#pragma unroll
                        for (int i = 0; i < NR; ++i) rp.ev_iroots[((int64_t)c * NR + i) * B + b] = rsm(3 * NR + i);
                    }
                    evc = c + 1;
                    call_done = false;  // falls through to the loop end (no `continue`: a second back edge splits the warp)
                    need_prologue = true;
                }
            }
            if (call_done) {
                if constexpr (DEFER) df.valid[b * lad.ntout + it] = deferred ? 1 : 0;
                if (yy_traj && !(DEFER && deferred)) {
#pragma unroll
                    for (int i = 0; i < N; ++i) yy_traj[((int64_t)it * N + i) * B + b] = yy[i];
                }
                if (status) status[(int64_t)it * B + b] = call_status;
                it += 1;
                need_prologue = true;
            }
            RP_ACC(3, ti0);
            RP_COUNT(4);
        }
        if (sg.max_trips >= 0 && sg.li_out) {  // park the loop state for the next segment (call_done is always consumed by here)
            sg.li_out[slot] = it | ((int)need_prologue << 8) | ((int)in_step << 9) | (ncf << 10) | (nef << 18);
            sg.li_out[sg.stride + slot] = nstloc;
            sg.ld_out[slot] = saved_t;
        }
    }
};

}  // namespace b200ida
