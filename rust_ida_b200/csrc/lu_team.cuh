// lu_team.cuh -- dense partial-pivot LU and triangular solves executed by a GROUP of threads (one warp or one
// whole CTA) on one column-major matrix that sits in shared memory (or, for matrices too big for it, in global
// memory).  Shared by the Level-1 batched kernels (linalg_batched.cu) and the team IDA kernels (ida_team.cuh).
//
// Semantics are bit-for-bit dense_get_rf / dense_get_rs (crates/linear/src/dense.rs:86-206):
//   pivot = first row with strictly largest |a_ik| (a NaN never wins; a NaN in row k keeps l = k); whole-row swap;
//   l_ik = a_ik * (1/a_kk); a_ij -= a_kj * l_ik with column j skipped when a_kj == 0; zero pivot -> k+1 and stop.
// Within elimination step k every a_ij update is one mul and one sub that depend only on step k-1, so spreading
// (i, j) over threads changes no rounding; only the k loop is sequential.  Built with -fmad=false.
//
// Schedule per step k (2 group barriers):
//   [A] swap rows k,l in every column j != k and rewrite column k as (pivot, multipliers)   -- threads over j / i
//   barrier
//   [B] trailing update, column j -> warp (j-k-1) mod nwarps, rows -> lanes, multipliers cached in registers;
//       warp 0 owns column k+1, so as soon as it has updated it, it runs the pivot search of step k+1
//       (look-ahead) while the other warps finish their columns
//   barrier
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "exact_div.cuh"

namespace b200ida {

struct LuShared {  // per-group scratch in shared memory
    double pv;     // pivot value a_lk
    double mult;   // 1 / pv
    double akk;    // a_kk before the swap
    int l;         // pivot row
    int info;      // 0 or 1-based column of the zero pivot
};

template <bool WARP>
struct LuGroup {
    __device__ __forceinline__ static int nthreads() { return WARP ? 32 : (int)blockDim.x; }
    __device__ __forceinline__ static int tid() { return WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }
    __device__ __forceinline__ static int nwarps() { return WARP ? 1 : (int)(blockDim.x >> 5); }
    __device__ __forceinline__ static int warp() { return WARP ? 0 : (int)(threadIdx.x >> 5); }
    __device__ __forceinline__ static void sync() {
        if (WARP) __syncwarp();
        else __syncthreads();
    }
};

// pivot search of column k by ONE warp (all 32 lanes call it); publishes l / pv / akk / info and piv[k]
template <class PivT>
__device__ __forceinline__ void lu_pivot_search(const double* M, int ld, int n, int k, LuShared* sh, PivT* piv) {
    const int lane = threadIdx.x & 31;
    const double* col = M + (size_t)k * ld;
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = k + lane; i < n; i += 32) {
        const double v = fabs(col[i]);
        if (v > best) {  // NaN fails '>' ; ascending i keeps the lowest index on ties
            best = v;
            bi = i;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) {
            best = ov;
            bi = oi;
        }
    }
    if (lane == 0) {
        const double akk = col[k];
        if (akk != akk || bi >= n) bi = k;  // NaN in row k (or nothing comparable): the reference keeps l = k
        const double pv = col[bi];
        sh->l = bi;
        sh->pv = pv;
        sh->mult = 1.0 / pv;
        sh->akk = akk;
        if (pv == 0.0) sh->info = k + 1;
        piv[k] = (PivT)bi;
    }
}

// trailing update of elimination step k for the columns of one warp, R = ceil((n-k-1)/32) row slots per lane
// (compile-time, so no predicated-off slots are issued); two columns per trip with all loads ahead of the stores.
// Warp 0 owns column k+1 and runs the pivot search of step k+1 as soon as that column is final (look-ahead).
template <int R, class PivT>
__device__ __forceinline__ void lu_update_cols(double* M, int ld, int n, int k, int w, int nw, int lane, LuShared* sh, PivT* piv) {
    const double* ck = M + (size_t)k * ld;
    const int i0 = k + 1 + lane;
    double lik[R];
#pragma unroll
    for (int r = 0; r < R; ++r) lik[r] = (i0 + 32 * r < n) ? ck[i0 + 32 * r] : 0.0;
    // column ownership: with several warps, warp 0 takes ONLY column k+1 and then does the look-ahead pivot search
    // (a serial chain: reduction, a division), while warps 1.. share columns k+2.. round-robin -- so the step's
    // critical path is max(search, update) instead of their sum.  A single-warp group takes everything.
    const bool split = nw > 1;
    const int jbeg = !split ? k + 1 : (w == 0 ? k + 1 : k + 1 + w);
    const int jstep = !split ? 1 : nw - 1;
    const int jend = (split && w == 0) ? k + 2 : n;
    for (int j0 = jbeg; j0 < jend && j0 < n; j0 += 2 * jstep) {
        const int j1 = j0 + jstep;
        const bool has1 = j1 < jend && j1 < n;
        double* c0 = M + (size_t)j0 * ld;
        double* c1 = M + (size_t)(has1 ? j1 : j0) * ld;
        const double a0 = c0[k];
        const double a1 = has1 ? c1[k] : 0.0;
        double v0[R], v1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool in = (r + 1 < R) || (i0 + 32 * r < n);  // only the last slot can be ragged
            v0[r] = in ? c0[i0 + 32 * r] : 0.0;
            v1[r] = in ? c1[i0 + 32 * r] : 0.0;
        }
        if (a0 != 0.0) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if ((r + 1 < R) || (i0 + 32 * r < n)) c0[i0 + 32 * r] = v0[r] - a0 * lik[r];
        }
        if (has1 && a1 != 0.0) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if ((r + 1 < R) || (i0 + 32 * r < n)) c1[i0 + 32 * r] = v1[r] - a1 * lik[r];
        }
        if (w == 0 && j0 == k + 1) {
            __syncwarp();
            lu_pivot_search(M, ld, n, k + 1, sh, piv);
        }
    }
}

// In-place LU of the n x n column-major matrix M (leading dimension ld).  Every thread of the group calls it.
// On return (after the final barrier) sh->info holds 0 or the 1-based singular column, piv[0..n) the pivot rows.
template <bool WARP, int RMAX, class PivT>
__device__ __forceinline__ void lu_factor_r(double* M, int ld, int n, LuShared* sh, PivT* piv) {
    using G = LuGroup<WARP>;  // RMAX = ceil(n / 32): upper bound of the row slots per lane
    const int tid = G::tid(), nt = G::nthreads(), nw = G::nwarps(), w = G::warp(), lane = tid & 31;
    if (tid == 0) sh->info = 0;
    G::sync();
    if (w == 0) lu_pivot_search(M, ld, n, 0, sh, piv);
    G::sync();
    for (int k = 0; k < n; ++k) {
        if (sh->info != 0) break;  // uniform: written before the barrier that precedes this read
        const int l = sh->l;
        const double pv = sh->pv, akk = sh->akk, mult = sh->mult;
        // [A] row swap outside column k, then column k := (pivot, multipliers)
        if (l != k) {
            for (int j = tid; j < n; j += nt) {
                if (j != k) {
                    double* c = M + (size_t)j * ld;
                    const double t = c[k];
                    c[k] = c[l];
                    c[l] = t;
                }
            }
        }
        {
            double* ck = M + (size_t)k * ld;
            for (int i = k + tid; i < n; i += nt) {
                if (i == k) {
                    ck[k] = pv;
                } else {
                    const double v = (i == l) ? akk : ck[i];
                    ck[i] = v * mult;
                }
            }
        }
        G::sync();
        // [B] trailing update with look-ahead pivot search
        const int R = (n - k - 1 + 31) >> 5;  // row slots needed at this step (uniform)
        if (RMAX >= 8 && R > 4) {
            if (R == 8) lu_update_cols<8>(M, ld, n, k, w, nw, lane, sh, piv);
            else if (R == 7) lu_update_cols<7>(M, ld, n, k, w, nw, lane, sh, piv);
            else if (R == 6) lu_update_cols<6>(M, ld, n, k, w, nw, lane, sh, piv);
            else lu_update_cols<5>(M, ld, n, k, w, nw, lane, sh, piv);
        } else if (RMAX >= 4 && R > 2) {
            if (R == 4) lu_update_cols<4>(M, ld, n, k, w, nw, lane, sh, piv);
            else lu_update_cols<3>(M, ld, n, k, w, nw, lane, sh, piv);
        } else if (RMAX >= 2 && R == 2) {
            lu_update_cols<2>(M, ld, n, k, w, nw, lane, sh, piv);
        } else if (R >= 1) {
            lu_update_cols<1>(M, ld, n, k, w, nw, lane, sh, piv);
        }
        G::sync();
    }
}

// ---- one warp, n <= N (N = 16 or 32), matrix in shared memory, factorisation in REGISTERS --------------------------------
// Lane i loads row i of the matrix once and keeps it in registers for the whole factorisation; rows are never moved: each
// lane tracks the LOGICAL position the reference's whole-row swaps would have given its row and writes the row there at
// the end.  Per step: arg-max of |a_ik| by three integer redux.sync on the IEEE bit pattern (NaN below every candidate, ties
// to the lowest logical position, the key carries the lane), exact reciprocal of the pivot in every lane, the pivot row
// broadcast through the matrix's own storage (the owner lane writes u_kj to its final place M[j][k], everybody reads it
// back as a broadcast), n-1-k independent mul / sub per lane.  Lanes >= n hold no row ("done" from the start).  The
// shared-memory group LU above pays two warp barriers, a 10-shuffle butterfly and a shared-memory round trip of the whole
// trailing matrix per step instead.  Same algorithm as the Level-1 getrf_warp_regs (linalg_batched.cu), any n <= N.
// full arg-max key of one pivot column over the lanes of `mask` (out of line: ties, inf and NaN only): bits of |x| (NaN below
// every candidate, lanes that are done below that), ties to the lowest logical position (pp = position << 8 | lane); nothing
// comparable, or a NaN in the row at logical position K (atk), keeps l = K like the reference
static __device__ __noinline__ unsigned lu_argmax_full(unsigned mask, double x, bool done, unsigned pp, bool atk) {
    const bool isnan_ = x != x;
    const int hi = done ? -2 : (isnan_ ? -1 : (__double2hiint(x) & 0x7fffffff));
    const unsigned lo = (done || isnan_) ? 0u : (unsigned)__double2loint(x);
    const int mhi = __reduce_max_sync(mask, hi);
    const unsigned mlo = __reduce_max_sync(mask, (hi == mhi) ? lo : 0u);
    unsigned mpp = __reduce_min_sync(mask, (!done && hi == mhi && lo == mlo) ? pp : 0xffffffffu);
    if (mhi < 0 || __any_sync(mask, atk && isnan_)) mpp = __reduce_min_sync(mask, atk ? pp : 0xffffffffu);
    return mpp;
}

template <int N, int K>
__device__ __forceinline__ void lu_wregs_step(double (&row)[N], double* M, int ld, int n, int lane, int& pos, bool& done, int& inf,
                                              int& mypiv) {
    if (K < n && inf == 0) {  // uniform
        const double x = row[K];
        const unsigned pp = ((unsigned)pos << 8) | (unsigned)lane;
        const bool atk = !done && pos == K;  // the lane that holds logical row K (positions < K are final pivot rows)
        // fast path: a unique, finite maximum among the high words of |x| (one reduction + a ballot); ties, inf and NaN
        // (whose high word is the maximum) take the full key -- see getrf_warp_regs
        const int key = done ? -1 : (__double2hiint(x) & 0x7fffffff);
        const int mkey = __reduce_max_sync(0xffffffffu, key);
        const unsigned wm = __ballot_sync(0xffffffffu, key == mkey);
        unsigned mpp;
        if (__popc(wm) == 1 && mkey >= 0 && mkey < 0x7ff00000) {  // uniform
            mpp = __shfl_sync(0xffffffffu, pp, __ffs(wm) - 1);
        } else {
            mpp = lu_argmax_full(0xffffffffu, x, done, pp, atk);
        }
        const int l = (int)(mpp >> 8), lw = (int)(mpp & 0xffu);
        const double pv = __shfl_sync(0xffffffffu, x, lw);
        if (lane == K) mypiv = l;
        if (pv == 0.0) {
            inf = K + 1;
        } else {
            const double mult = 1.0 / pv;
            const bool owner = lane == lw;
            if (owner) {
                pos = K;
                done = true;
            } else if (atk) {
                pos = l;
            }
            const double lik = x * mult;
            if (!done) row[K] = lik;
            if (owner) {
#pragma unroll
                for (int j = K + 1; j < N; ++j)
                    if (j < n) M[(size_t)j * ld + K] = row[j];
            }
            __syncwarp();
#pragma unroll
            for (int j = K + 1; j < N; ++j) {
                if (j < n) {  // uniform
                    const double a = M[(size_t)j * ld + K];
                    if (!done && a != 0.0) row[j] -= a * lik;
                }
            }
        }
    }
    if constexpr (K + 1 < N) lu_wregs_step<N, K + 1>(row, M, ld, n, lane, pos, done, inf, mypiv);
}

template <int N, class PivT>
__device__ __noinline__ void lu_factor_warp_regs(double* M, int ld, int n, LuShared* sh, PivT* piv) {
    const int lane = threadIdx.x & 31;
    const bool mine = lane < n;
    const int lc = mine ? lane : n - 1;
    double row[N];
#pragma unroll
    for (int j = 0; j < N; ++j) row[j] = (j < n) ? M[(size_t)j * ld + lc] : 0.0;
    __syncwarp();
    int pos = lane;
    bool done = !mine;
    int inf = 0, mypiv = 0;
    lu_wregs_step<N, 0>(row, M, ld, n, lane, pos, done, inf, mypiv);
    __syncwarp();
    if (mine) {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (j < n) M[(size_t)j * ld + pos] = row[j];
        if (inf == 0 || lane < inf) piv[lane] = (PivT)mypiv;
    }
    if (lane == 0) sh->info = inf;
    __syncwarp();
}

// dispatch on ceil(n/32) so that small systems do not issue predicated-off row slots
template <bool WARP, class PivT>
__device__ __forceinline__ void lu_factor(double* M, int ld, int n, LuShared* sh, PivT* piv) {
    if (WARP && n <= 16) lu_factor_warp_regs<16>(M, ld, n, sh, piv);
    else if (WARP && n <= 32) lu_factor_warp_regs<32>(M, ld, n, sh, piv);
    else if (n <= 32) lu_factor_r<WARP, 1>(M, ld, n, sh, piv);
    else if (n <= 64) lu_factor_r<WARP, 2>(M, ld, n, sh, piv);
    else if (n <= 128) lu_factor_r<WARP, 4>(M, ld, n, sh, piv);
    else lu_factor_r<WARP, 8>(M, ld, n, sh, piv);
}

// Solve with the factors: b := A^{-1} b, all in shared (or global) memory; every thread of the group calls it.
// The substitutions are a chain of 2n dependent steps with O(n) work each, so they run on ONE warp and, in a CTA
// group, the other warps wait at the final barrier.  The right-hand side lives in REGISTERS (lane owns rows lane,
// lane+32, ...), the pivot-row value of each step is broadcast with a shuffle, and the column of L/U is read from
// shared memory with loads that do not depend on the chain -- so a step costs one shuffle + mul + sub of latency
// instead of a shared-memory store->load round trip per row slot.
// rdiag (optional): rdiag[k] = RN(1 / u_kk), computed off the critical path; the division of the back substitution
// then costs 3 dependent instructions instead of ~10 (exact_div.cuh) and stays bit-identical.
// forward / backward sweeps over the 32 pivots owned by register slot R0 (compile-time), so the owner value is
// x[R0] without a select chain and only the slots that can still change are touched
template <int RMAX, int R0>
__device__ __forceinline__ void lu_fwd_block(const double* M, int ld, int n, int lane, double (&x)[RMAX]) {
    const int kend = (32 * R0 + 32 < n - 1) ? 32 * R0 + 32 : n - 1;
    for (int k = 32 * R0; k < kend; ++k) {
        const double* ck = M + (size_t)k * ld;
        const int kk = k & 31;
        double c[RMAX];
#pragma unroll
        for (int r = R0; r < RMAX; ++r) {
            const int i = lane + 32 * r;
            c[r] = (i < n) ? ck[i] : 0.0;
        }
        const double bk = __shfl_sync(0xffffffffu, x[R0], kk);
        if (lane > kk) x[R0] -= c[R0] * bk;
#pragma unroll
        for (int r = R0 + 1; r < RMAX; ++r)
            if (lane + 32 * r < n) x[r] -= c[r] * bk;
    }
    if constexpr (R0 + 1 < RMAX) {
        if (32 * (R0 + 1) < n - 1) lu_fwd_block<RMAX, R0 + 1>(M, ld, n, lane, x);
    }
}
template <int RMAX, int R0>
__device__ __forceinline__ void lu_bwd_block(const double* M, int ld, int n, int lane, double (&x)[RMAX], const double* rdiag) {
    if (32 * R0 < n) {
        const int kbeg = (32 * R0 + 31 < n - 1) ? 32 * R0 + 31 : n - 1;
        for (int k = kbeg; k >= 32 * R0; --k) {
            const double* ck = M + (size_t)k * ld;
            const int kk = k & 31;
            const double ukk = ck[k];
            const double y = rdiag ? rdiag[k] : 0.0;
            double c[RMAX];
#pragma unroll
            for (int r = 0; r < R0; ++r) c[r] = ck[lane + 32 * r];  // rows < 32*R0 <= k: always in range
            c[R0] = (lane < kk) ? ck[lane + 32 * R0] : 0.0;
            double xv = x[R0];
            if (lane == kk) xv = rdiag ? div_by_recip(xv, ukk, y) : xv / ukk;
            const double bk = __shfl_sync(0xffffffffu, xv, kk);
            if (lane == kk) x[R0] = bk;
            else if (lane < kk) x[R0] -= c[R0] * bk;
#pragma unroll
            for (int r = 0; r < R0; ++r) x[r] -= c[r] * bk;
        }
    }
    if constexpr (R0 > 0) lu_bwd_block<RMAX, R0 - 1>(M, ld, n, lane, x, rdiag);
}

template <int RMAX, class PivT>
__device__ __forceinline__ void lu_solve_warp_r(const double* M, int ld, int n, const PivT* piv, double* b, const double* rdiag) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        for (int k = 0; k < n; ++k) {
            const int pk = (int)piv[k];
            if (pk != k) {
                const double t = b[k];
                b[k] = b[pk];
                b[pk] = t;
            }
        }
    }
    __syncwarp();
    double x[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        const int i = lane + 32 * r;
        x[r] = (i < n) ? b[i] : 0.0;
    }
    lu_fwd_block<RMAX, 0>(M, ld, n, lane, x);
    lu_bwd_block<RMAX, RMAX - 1>(M, ld, n, lane, x, rdiag);
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        const int i = lane + 32 * r;
        if (i < n) b[i] = x[r];
    }
}

// n <= 32 with the row permutation (lu_build_perm) and the reciprocals of the diagonal at hand: one warp, lane = row, the
// right-hand side gathered straight into registers (no serial interchange loop), the backward sweep's division done by
// every lane in line (exact-reciprocal sequence) with the pivot lanes' range tests voted once at the end -- a failed
// vote redoes the backward sweep with the guarded division.  Same operands and per-row order as dense_get_rs.
__device__ __forceinline__ void lu_solve_warp32(const double* M, int ld, int n, const int* perm, double* b, const double* rdiag) {
    const int lane = threadIdx.x & 31;
    const int ic = lane < n ? lane : n - 1;  // lanes beyond the matrix mirror row n-1 and store nothing
    double x = b[perm[ic]];
    __syncwarp();
#pragma unroll 4
    for (int k = 0; k < n - 1; ++k) {
        const double c = M[(size_t)k * ld + ic];
        const double bk = __shfl_sync(0xffffffffu, x, k);
        const double t = x - c * bk;
        x = (lane > k) ? t : x;
    }
    const double x_in = x;
    double xp = 1.0;
#pragma unroll 4
    for (int k = n - 1; k >= 0; --k) {
        const double c = M[(size_t)k * ld + ic];
        const double y = rdiag[k];
        const double q0 = __dmul_rn(x, y);
        const double rr = __fma_rn(-c, q0, x);
        const double q1 = __fma_rn(rr, y, q0);
        const bool piv = lane == k;
        const double q = (x == 0.0) ? q0 : q1;  // +-0: the zero with the sign of x * y
        xp = piv ? x : xp;
        const double bk = __shfl_sync(0xffffffffu, piv ? q : x, k);
        const double t = x - c * bk;
        x = piv ? bk : ((lane < k) ? t : x);
    }
    const double dg = M[(size_t)ic * ld + ic];
    const unsigned ex = ((unsigned)__double2hiint(xp) >> 20) & 0x7ffu, ed = ((unsigned)__double2hiint(dg) >> 20) & 0x7ffu;
    const bool ok = (xp == 0.0) ? (ed - 1u < 0x7feu) : ((ex - 543u) <= 960u && (ed - 543u) <= 960u);
    if (!__all_sync(0xffffffffu, ok)) {
        x = x_in;
        for (int k = n - 1; k >= 0; --k) {
            const double c = M[(size_t)k * ld + ic];
            double xv = x;
            if (lane == k) xv = div_by_recip(x, c, rdiag[k]);
            const double bk = __shfl_sync(0xffffffffu, xv, k);
            const double t = x - c * bk;
            x = (lane == k) ? bk : ((lane < k) ? t : x);
        }
    }
    if (lane < n) b[lane] = x;
}

template <bool WARP, class PivT>
__device__ __forceinline__ void lu_solve(const double* M, int ld, int n, const PivT* piv, double* b,
                                         const double* rdiag = nullptr, const int* perm = nullptr) {
    using G = LuGroup<WARP>;
    G::sync();
    if (G::warp() == 0) {
        if (n <= 32 && perm != nullptr && rdiag != nullptr) lu_solve_warp32(M, ld, n, perm, b, rdiag);
        else if (n <= 32) lu_solve_warp_r<1>(M, ld, n, piv, b, rdiag);
        else if (n <= 64) lu_solve_warp_r<2>(M, ld, n, piv, b, rdiag);
        else if (n <= 128) lu_solve_warp_r<4>(M, ld, n, piv, b, rdiag);
        else lu_solve_warp_r<8>(M, ld, n, piv, b, rdiag);
    }
    G::sync();
}

// ---- blocked triangular solves for a whole CTA (n > 32) -----------------------------------------------------------------
// The 2n-step substitution chain is split into 32 x 32 diagonal blocks (one warp, right-hand side in registers, one
// shuffle + mul + sub per step) and the off-diagonal panels below / above them, which are independent per row: row i
// subtracts l_ik x_k for the 32 pivots of the block in the reference's order (ascending k in the forward sweep,
// descending in the backward sweep), a chain of subtractions with no shuffles, and the rows are spread over all
// warps.  Per element the operations and their order are those of dense_get_rs (dense.rs:165-206), so the result is
// bit-identical to the single-warp chain; the critical path drops from 2n full-height steps to 2n single-slot steps.
// perm (optional): perm[i] = source row of row i after the row interchanges (built once per factorisation by
// lu_build_perm), so that the interchanges are one parallel gather instead of a serial loop per solve.
template <class PivT>
__device__ __forceinline__ void lu_build_perm(int n, const PivT* piv, int* perm) {  // one thread
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        const int pk = (int)piv[k];
        if (pk != k) {
            const int t = perm[k];
            perm[k] = perm[pk];
            perm[pk] = t;
        }
    }
}

__device__ __forceinline__ void lu_solve_cta_blocked(const double* M, int ld, int n, const int* perm, double* b, double* tmp,
                                                     const double* rdiag) {
    const int tid = threadIdx.x, nt = blockDim.x, w = tid >> 5, nw = nt >> 5, lane = tid & 31;
    const int nblk = (n + 31) >> 5;
    // row interchanges: b := P b
    for (int i = tid; i < n; i += nt) tmp[i] = b[perm[i]];
    __syncthreads();
    for (int i = tid; i < n; i += nt) b[i] = tmp[i];
    __syncthreads();
    // forward sweep, unit lower triangle
    for (int R0 = 0; R0 < nblk; ++R0) {
        const int k0 = 32 * R0;
        const int kend = (k0 + 32 < n) ? k0 + 32 : n;
        if (w == 0) {
            const int i = k0 + lane;
            const int ic = i < n ? i : n - 1;
            double x = b[ic];
#pragma unroll 4
            for (int k = k0; k < kend - 1; ++k) {
                const double c = M[(size_t)k * ld + ic];
                const double bk = __shfl_sync(0xffffffffu, x, k - k0);
                const double t = x - c * bk;
                x = (i > k) ? t : x;
            }
            if (i < n) b[i] = x;
        }
        __syncthreads();
        // rows below the block: i = 32 (R0 + 1) + ..; each warp takes whole 32-row slots
        for (int r = R0 + 1 + w; r < nblk; r += nw) {
            const int i = 32 * r + lane;
            const int ic = i < n ? i : n - 1;
            double acc = b[ic];
#pragma unroll 8
            for (int k = k0; k < kend; ++k) acc -= M[(size_t)k * ld + ic] * b[k];
            if (i < n) b[i] = acc;
        }
        __syncthreads();
    }
    // backward sweep, upper triangle with the divisions by u_kk
    for (int R0 = nblk - 1; R0 >= 0; --R0) {
        const int k0 = 32 * R0;
        const int ktop = (k0 + 31 < n - 1) ? k0 + 31 : n - 1;
        if (w == 0) {
            const int i = k0 + lane;
            const int ic = i < n ? i : n - 1;
            double x = b[ic];
            bool done = false;
            if (rdiag) {
                // Straight-line chain: EVERY lane forms the exact-reciprocal quotient of its own (x, c) each step (three
                // FP64 instructions, exact_div.cuh) and only the pivot lane keeps it -- the one-lane divergent region with
                // its range tests and fallbacks was most of this chain's length (round-2 profile: the backward diagonal
                // chains were a third of the heat kernel's run time, 4.4 x the forward ones).  The pivot lane's range test
                // (and the exact +-0 numerator, which boundary rows produce all the time) is carried in a flag and voted
                // on ONCE after the block; a block that fails is redone from its saved right-hand side by the guarded loop.
                const double x_in = x;
                double xp = 1.0;  // this lane's numerator at the step where it was the pivot lane (range-tested after the loop)
#pragma unroll 4
                for (int k = ktop; k >= k0; --k) {
                    const double c = M[(size_t)k * ld + ic];  // lane == k - k0 reads u_kk
                    const double y = rdiag[k];
                    const double q0 = __dmul_rn(x, y);
                    const double rr = __fma_rn(-c, q0, x);
                    const double q1 = __fma_rn(rr, y, q0);
                    const bool piv = i == k;
                    // x = +-0: the quotient is the zero with the sign of x * y, which q0 already is (q1 can lose -0)
                    const double q = (x == 0.0) ? q0 : q1;
                    xp = piv ? x : xp;
                    const double bk = __shfl_sync(0xffffffffu, piv ? q : x, k - k0);
                    const double t = x - c * bk;
                    x = piv ? bk : ((i < k) ? t : x);
                }
                // the pivot lanes' range tests, once per block: divisor u_ii and numerator within 2^-480 .. 2^480, or an
                // exact zero numerator over a finite normal divisor (exact_div.cuh)
                const double dg = M[(size_t)ic * ld + ic];
                const unsigned ex = ((unsigned)__double2hiint(xp) >> 20) & 0x7ffu, ed = ((unsigned)__double2hiint(dg) >> 20) & 0x7ffu;
                const bool ok = (xp == 0.0) ? (ed - 1u < 0x7feu) : ((ex - 543u) <= 960u && (ed - 543u) <= 960u);
                done = __all_sync(0xffffffffu, ok);
                if (!done) x = x_in;
            }
            if (!done) {
#pragma unroll 4
                for (int k = ktop; k >= k0; --k) {
                    const double c = M[(size_t)k * ld + ic];  // lane == k - k0 reads u_kk
                    double xv = x;
                    if (i == k) xv = rdiag ? div_by_recip(x, c, rdiag[k]) : x / c;
                    const double bk = __shfl_sync(0xffffffffu, xv, k - k0);
                    const double t = x - c * bk;
                    x = (i == k) ? bk : ((i < k) ? t : x);
                }
            }
            if (i < n) b[i] = x;
        }
        __syncthreads();
        for (int r = w; r < R0; r += nw) {
            const int i = 32 * r + lane;  // always < n
            double acc = b[i];
#pragma unroll 8
            for (int k = ktop; k >= k0; --k) acc -= M[(size_t)k * ld + i] * b[k];
            b[i] = acc;
        }
        __syncthreads();
    }
}

}  // namespace b200ida
