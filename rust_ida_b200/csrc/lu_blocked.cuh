// lu_blocked.cuh -- blocked, register-tiled partial-pivot LU of ONE column-major matrix, executed by a whole CTA.
// lub_factor / lub_factor_1w: the matrix sits in shared memory (the CTA-per-system IDA kernel, ida_team.cuh, for the
// iteration matrix of the heat problem, n = 100 / the Level-1 batched getrf for 33 <= n <= 114, linalg_batched.cu).
// lub_factor_global (end of file): the matrix stays in global memory / L2 and only the panels go through shared
// memory (Level-1 getrf for 115 <= n <= 256).
//
// Semantics are bit-for-bit dense_get_rf (crates/linear/src/dense.rs:86-158): pivot = first row with strictly
// largest |a_ik| (a NaN never wins; a NaN in row k keeps l = k); whole-row swap; l_ik = a_ik * (1/a_kk);
// a_ij -= a_kj * l_ik with column j skipped when a_kj == 0; zero pivot -> info = k+1 and stop.
//
// Why blocking is bit-exact.  The final value of a_ij is  (..((a_ij - l_i0 u_0j) - l_i1 u_1j) - ..), one un-fused
// multiply and one subtract per k, in increasing k.  The unblocked reference interleaves the k-steps of different
// (i, j); any schedule that keeps, for each (i, j), the k-order and the operands is bit-identical.  Here NB = 8
// consecutive k-steps are applied to a register tile before it goes back to shared memory, which divides the
// shared-memory traffic of the trailing update by NB (the unblocked lu_team.cuh kernel is bound by it).
// Built with -fmad=false.
//
// One block step (columns kb .. kb+nb), three CTA barriers (the panel of the NEXT step runs inside [U], see lub_factor):
//   [P] panel: warp 0 eliminates the nb columns of the (n-kb) x nb panel in shared memory (lane = row, 32-row
//       slots), arg-max by warp integer reductions, never moving rows: each lane tracks the LOGICAL position the
//       reference's row swaps would have given its rows.  It permutes the panel to the logical positions at the end
//       and publishes the list of row moves (src -> dst, at most 2 nb) of this panel.
//   [S] every other column (one thread per column): apply the row moves; right of the panel also the nb-step
//       unit-lower solve that produces the U rows kb..kb+nb (kept in registers, chain of nb), copied to Ub.
//   [U] trailing update: warps own (64-row x 4-column) register tiles; per tile 8 loads + 8 stores of the matrix
//       against 2*2*4*8 FP64 operations, multipliers cached in registers across the tiles of a row block,
//       U values read as 16-byte broadcasts from Ub.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200ida {

#ifdef LUB_PROFILE  // debug build only: per-phase clock accumulation (tools/prof_lub.py)
__device__ unsigned long long lub_prof[32];
#define LUB_ACC2(i, a, b) lub_acc[i] += (unsigned long long)((b) - (a))
#define LUB_ACC2_DECL unsigned long long lub_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define LUB_ACC2_FLUSH do { if (threadIdx.x == 0 && blockIdx.x == 0) { for (int i_ = 0; i_ < 8; ++i_) lub_prof[16 + i_] += lub_acc[i_]; } } while (0)
#define LUB_T(var) const long long var = clock64()
#define LUB_ACC(i, a, b) do { if ((threadIdx.x == 0 || threadIdx.x == 32) && blockIdx.x == 0) atomicAdd(&lub_prof[(i) + (threadIdx.x ? 8 : 0)], (unsigned long long)((b) - (a))); } while (0)
#else
#define LUB_T(var)
#define LUB_ACC(i, a, b)
#define LUB_ACC2(i, a, b)
#define LUB_ACC2_DECL
#define LUB_ACC2_FLUSH
#endif

constexpr int LUB_NB = 8;                   // panel width = k-steps fused in registers
constexpr int LUB_CT = 4;                   // columns per register tile in [U]
constexpr int LUB_MAXMOVES = 2 * LUB_NB;    // a panel of nb pivots moves at most 2 nb rows

constexpr int LUB_MAXPW = 8;                // at most this many warps share one panel
struct LubCtl {  // per-CTA control block in shared memory
    int info;    // 0 or the 1-based column of the zero pivot
    int nmoves;  // row moves of the current panel
    int ndone;   // panel columns actually eliminated (== nb unless a zero pivot stopped the panel)
    int qhead;   // next chunk of the trailing-update work queue (lub_update_q)
    int src[LUB_MAXMOVES];
    int dst[LUB_MAXMOVES];
    int4 xs[2][LUB_MAXPW];    // per-warp arg-max candidates of a panel column (fast key), double-buffered by column parity
    int xs2[LUB_MAXPW][8];    // the same for the full-key path: hi, lo, pp, pk, nank
};
// barrier among the warps of one panel (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void lub_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

__host__ __device__ inline int lub_ldu(int n) { return (n + 7) & ~3; }  // row pitch of Ub (multiple of 4, >= n + 3)
// scratch after the matrix: Ub (LUB_NB x ldu doubles) followed by the per-column zero flags (ldu bytes)
__host__ __device__ inline size_t lub_ub_bytes(int n) { return (size_t)LUB_NB * lub_ldu(n) * sizeof(double) + (size_t)lub_ldu(n); }

// ---- [P] panel factorisation by one warp, panel resident in shared memory ------------------------------------------------
// Lane = row (32-row slots, R of them hold rows).  Rows are NOT moved while the panel is eliminated: pos[r] is the
// LOGICAL position the reference's row swaps would have given register row r; the panel is permuted once at the end.
// Per column: arg-max by three warp integer reductions (redux.sync) on the IEEE bit pattern of |a_ik| (for finite
// non-negative doubles the bit pattern orders like the value; NaN maps below every candidate so it never wins; ties
// go to the lowest logical position, which also carries the physical row in its low bits), exact reciprocal of the
// pivot started from the reduced |pivot| bits while its sign is fetched, right-looking update of the remaining panel
// columns straight in shared memory.
// shared-memory accesses by 32-bit shared-window address: no generic-address arithmetic in the serial chain
__device__ __forceinline__ double lub_lds(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void lub_sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// step-c update of the ND panel columns right of column c: v -= a_d * x for the rows still active; a zero pivot-row
// entry a_d skips column d like the reference (select on the result, so the block stays straight-line code)
template <int R, int ND>
__device__ __forceinline__ void lub_panel_update(const unsigned (&ra)[R], unsigned pa, unsigned ldb, const double (&x)[R],
                                                 const bool (&act)[R]) {
    double a[ND], v[ND][R];
#pragma unroll
    for (int d = 0; d < ND; ++d) a[d] = lub_lds(pa + (unsigned)(d + 1) * ldb);
#pragma unroll
    for (int d = 0; d < ND; ++d)
#pragma unroll
        for (int r = 0; r < R; ++r) v[d][r] = lub_lds(ra[r] + (unsigned)(d + 1) * ldb);
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const bool nz = a[d] != 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double t = v[d][r] - a[d] * x[r];
            if (act[r] && nz) lub_sts(ra[r] + (unsigned)(d + 1) * ldb, t);
        }
    }
}

// ---- the ONE-WARP panel (Level-1 getrf kernels): lane = row, R 32-row slots per lane.  With many CTAs per SM and the
// look-ahead update running beside it, one warp on a scheduler of its own is the cheapest way through a panel (measured:
// the shared panel below is 20-25 % slower there); the CTA-per-system IDA kernel, whose SM is mostly idle, uses the shared
// panel (lub_panel_mw).
template <int R>
__device__ __noinline__ void lub_panel_1w(double* M, int ld, int n, int kb, int nb, LubCtl* ctl, int* pivs) {
    // One warp, in-order issue: the cost of a column step is (instructions issued) x (4..8 cycles), so the loop is
    // written for instruction count -- 32-bit shared addresses advanced per column, a "pivoted" flag per row
    // instead of position compares, unconditional loads (rows beyond the matrix alias row n-1; their results are
    // never stored) and predicated lone stores, so the body is straight-line code.
    const int lane = threadIdx.x & 31;
    LUB_ACC2_DECL;
    const unsigned mbase = (unsigned)__cvta_generic_to_shared(M);
    const unsigned ldb = (unsigned)ld * 8u;
    unsigned ra[R];   // shared address of (row r, current column)
    unsigned pp[R];   // (logical position << 16) | physical row
    bool act[R];      // row exists and has not been chosen as a pivot yet
    int rowc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = kb + lane + 32 * r;
        act[r] = row < n;
        rowc[r] = act[r] ? row : n - 1;
        ra[r] = mbase + (unsigned)(kb * ld + rowc[r]) * 8u;
        pp[r] = ((unsigned)rowc[r] << 16) | (unsigned)rowc[r];
    }
    unsigned colb = mbase + (unsigned)(kb * ld) * 8u;  // shared address of (row 0, current column)
    int ndone = nb, info = 0;
#pragma unroll 1
    for (int c = 0; c < nb; ++c) {
        const int k = kb + c;
        LUB_T(p0);
        double x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = lub_lds(ra[r]);
        // Fast path (the common case): compare only the high words of |x| -- sign stripped, so for finite values
        // they order like the values -- and accept the winner when it is unique: exactly one lane holds the maximum
        // high word, that lane saw it once, and it is finite.  A unique, strictly largest high word means a strictly
        // largest |x|, so the choice equals the reference's; everything else (equal high words, i.e. possible ties,
        // and any inf / NaN candidate) takes the full lexicographic path below.
        int bkey = -1;
        unsigned bpp = 0xffffffffu;
        bool tie = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int key = act[r] ? (__double2hiint(x[r]) & 0x7fffffff) : -1;
            const bool gt = key > bkey;
            tie = gt ? false : (tie || (key == bkey && key >= 0));
            bkey = gt ? key : bkey;
            bpp = gt ? pp[r] : bpp;
        }
        const int mkey = __reduce_max_sync(0xffffffffu, bkey);
        const unsigned wmask = __ballot_sync(0xffffffffu, bkey == mkey);
        const bool unique = __popc(wmask) == 1 && !__any_sync(0xffffffffu, tie && bkey == mkey) && mkey < 0x7ff00000;
        unsigned mpp;
        bool forced = false;
        double mult, pv;
        unsigned pvoff;
        LUB_T(p1);
        LUB_ACC2(0, p0, p1);
        if (unique) {  // warp-uniform
            mpp = __shfl_sync(0xffffffffu, bpp, __ffs(wmask) - 1);
            pvoff = (mpp & 0xffffu) * 8u;
            pv = lub_lds(colb + pvoff);
            mult = __drcp_rn(pv);  // == 1.0 / pv, both correctly rounded
        } else {
            // full key = bits of |x| (NaN -> -1, not a candidate -> -2), ties -> lowest logical position
            int bhi = -2;
            unsigned blo = 0u;
            bpp = 0xffffffffu;
            bool nank = false;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool isnan_ = x[r] != x[r];
                const int hi = isnan_ ? -1 : (__double2hiint(x[r]) & 0x7fffffff);
                const unsigned lo = isnan_ ? 0xffffffffu : (unsigned)__double2loint(x[r]);
                const bool gt = hi > bhi || (hi == bhi && (lo > blo || (lo == blo && pp[r] < bpp)));
                const bool take = act[r] && gt;
                bhi = take ? hi : bhi;
                blo = take ? lo : blo;
                bpp = take ? pp[r] : bpp;
                nank = nank || (isnan_ && (pp[r] >> 16) == (unsigned)k);
            }
            const int mhi = __reduce_max_sync(0xffffffffu, bhi);
            const unsigned mlo = __reduce_max_sync(0xffffffffu, (bhi == mhi) ? blo : 0u);
            mpp = __reduce_min_sync(0xffffffffu, (bhi == mhi && blo == mlo) ? bpp : 0xffffffffu);
            forced = mhi < 0 || __any_sync(0xffffffffu, nank);  // nothing comparable, or NaN in row k: l = k
            if (forced) {
                unsigned pk = 0xffffffffu;
#pragma unroll
                for (int r = 0; r < R; ++r) pk = (act[r] && (pp[r] >> 16) == (unsigned)k) ? pp[r] : pk;
                mpp = __reduce_min_sync(0xffffffffu, pk);
            }
            pvoff = (mpp & 0xffffu) * 8u;
            pv = lub_lds(colb + pvoff);
            mult = 1.0 / pv;
        }
        const int l = (int)(mpp >> 16);
        LUB_T(p2);
        LUB_ACC2(1, p1, p2);
        if (lane == 0) pivs[k] = l;
        if (pv == 0.0) {
            info = k + 1;
            ndone = c;
            break;
        }
        LUB_T(p3);
        LUB_ACC2(2, p2, p3);
        // logical positions: the pivot row goes to k, the row that sat at k goes to l; multipliers below the pivot
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool isl = pp[r] == mpp;
            const bool isk = (pp[r] >> 16) == (unsigned)k;
            const unsigned low = pp[r] & 0xffffu;
            pp[r] = isl ? (((unsigned)k << 16) | low) : (isk ? (((unsigned)l << 16) | low) : pp[r]);
            act[r] = act[r] && !isl;
            x[r] = x[r] * mult;
            if (act[r]) lub_sts(ra[r], x[r]);
        }
        LUB_T(p4);
        LUB_ACC2(3, p3, p4);
        // right-looking update of the remaining nd panel columns: all loads, then the arithmetic, then the stores
        // (specialised on nd so that nothing is issued for columns that do not exist)
        switch (nb - 1 - c) {
        case 7: lub_panel_update<R, 7>(ra, colb + pvoff, ldb, x, act); break;
        case 6: lub_panel_update<R, 6>(ra, colb + pvoff, ldb, x, act); break;
        case 5: lub_panel_update<R, 5>(ra, colb + pvoff, ldb, x, act); break;
        case 4: lub_panel_update<R, 4>(ra, colb + pvoff, ldb, x, act); break;
        case 3: lub_panel_update<R, 3>(ra, colb + pvoff, ldb, x, act); break;
        case 2: lub_panel_update<R, 2>(ra, colb + pvoff, ldb, x, act); break;
        case 1: lub_panel_update<R, 1>(ra, colb + pvoff, ldb, x, act); break;
        default: break;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) ra[r] += ldb;
        colb += ldb;
        LUB_T(p5);
        LUB_ACC2(4, p4, p5);
        __syncwarp();
        LUB_T(p6);
        LUB_ACC2(5, p5, p6);
    }
    __syncwarp();
    LUB_T(q0);
    // permute the panel to the logical positions (two columns per trip), publish the row moves for the other columns
    bool moved[R];
    bool anymoved = false;
    int posc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int pos = (int)(pp[r] >> 16);
        moved[r] = kb + lane + 32 * r < n && pos != rowc[r];
        anymoved = anymoved || moved[r];
        posc[r] = moved[r] ? pos : rowc[r];
    }
    if (__any_sync(0xffffffffu, anymoved)) {
#pragma unroll 1
        for (int c = 0; c < nb; c += 2) {
            double* Mc = M + (kb + c) * ld;
            const int o1 = (c + 1 < nb) ? ld : 0;  // odd width: the second column of the trip is the first again
            double t0[R], t1[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                t0[r] = Mc[rowc[r]];
                t1[r] = Mc[o1 + rowc[r]];
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (moved[r]) Mc[posc[r]] = t0[r];
            }
            if (o1 != 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (moved[r]) Mc[o1 + posc[r]] = t1[r];
                }
            }
        }
    }
    int base = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned m = __ballot_sync(0xffffffffu, moved[r]);
        const int idx = base + __popc(m & ((1u << lane) - 1u));
        if (moved[r]) ctl->src[idx] = rowc[r];
        if (moved[r]) ctl->dst[idx] = posc[r];
        base += __popc(m);
    }
    if (lane >= base && lane < LUB_MAXMOVES) {  // unused slots: a valid row, so [S] can load unconditionally
        ctl->src[lane] = kb;
        ctl->dst[lane] = kb;
    }
    if (lane == 0) {
        ctl->nmoves = base;
        ctl->ndone = ndone;
        if (info != 0) ctl->info = info;
    }
    __syncwarp();
    LUB_T(q1);
    LUB_ACC2(6, q0, q1);
    LUB_ACC2_FLUSH;
}

// dispatch on the number of 32-row slots that still hold rows, so later (shorter) panels issue no dead instructions
template <int RP>
__device__ __forceinline__ void lub_panel_any_1w(double* M, int ld, int n, int kb, int nb, LubCtl* ctl, int* pivs) {
    const int nr = (n - kb + 31) >> 5;
    if (RP >= 8 && nr >= 7) lub_panel_1w<8>(M, ld, n, kb, nb, ctl, pivs);
    else if (RP >= 6 && nr == 6) lub_panel_1w<6>(M, ld, n, kb, nb, ctl, pivs);
    else if (RP >= 5 && nr >= 5) lub_panel_1w<5>(M, ld, n, kb, nb, ctl, pivs);
    else if (RP >= 4 && nr == 4) lub_panel_1w<4>(M, ld, n, kb, nb, ctl, pivs);
    else if (RP >= 3 && nr == 3) lub_panel_1w<3>(M, ld, n, kb, nb, ctl, pivs);
    else if (RP >= 2 && nr == 2) lub_panel_1w<2>(M, ld, n, kb, nb, ctl, pivs);
    else lub_panel_1w<1>(M, ld, n, kb, nb, ctl, pivs);
}


// ---- the SHARED panel: P warps (P = 1 degenerates to one warp).  Warp pw owns the 32-row slots pw, pw + P, .. (RW of them
// per lane), so with P = ceil(rows / 32) every lane carries ONE row and a column step issues a quarter of the
// instructions of the one-warp panel at n = 100 -- the panel is bound by instructions issued in order by its warp(s).
// Per column the warps exchange their best candidate through shared memory (one 16-byte record per warp, double-buffered
// by column parity) with ONE named barrier; that barrier also orders the previous step's panel stores (each thread only
// ever writes its own rows) before this step's reads of the pivot row.  The full-key path (ties / inf / NaN among the
// leaders) exchanges a second record behind a second barrier.  Every panel warp takes the same path: the decision is
// made from the exchanged records.
template <int RW>
__device__ __noinline__ void lub_panel_mw(double* M, int ld, int n, int kb, int nb, LubCtl* ctl, int* pivs, int pw, int P) {
    // In-order issue: the cost of a column step is (instructions issued) x (4..8 cycles), so the loop is written for
    // instruction count -- 32-bit shared addresses advanced per column, a "pivoted" flag per row instead of position
    // compares, unconditional loads (rows beyond the matrix alias row n-1; their results are never stored) and
    // predicated lone stores, so the body is straight-line code.
    const int lane = threadIdx.x & 31;
    const int nth = 32 * P;
    LUB_ACC2_DECL;
    const unsigned mbase = (unsigned)__cvta_generic_to_shared(M);
    const unsigned ldb = (unsigned)ld * 8u;
    unsigned ra[RW];   // shared address of (row r, current column)
    unsigned pp[RW];   // (logical position << 16) | physical row
    bool act[RW];      // row exists and has not been chosen as a pivot yet
    bool mine[RW];     // row exists
    int rowc[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        const int row = kb + lane + 32 * (pw + P * r);
        mine[r] = row < n;
        act[r] = mine[r];
        rowc[r] = act[r] ? row : n - 1;
        ra[r] = mbase + (unsigned)(kb * ld + rowc[r]) * 8u;
        pp[r] = ((unsigned)rowc[r] << 16) | (unsigned)rowc[r];
    }
    if (pw == 0) {  // ordered before every other warp's use by the first column's barrier (same warp when P == 1)
        if (lane < LUB_MAXMOVES) {  // unused move slots: a valid row, so [S] can load unconditionally
            ctl->src[lane] = kb;
            ctl->dst[lane] = kb;
        }
        if (lane == 0) ctl->nmoves = 0;
    }
    unsigned colb = mbase + (unsigned)(kb * ld) * 8u;  // shared address of (row 0, current column)
    int ndone = nb, info = 0;
#pragma unroll 1
    for (int c = 0; c < nb; ++c) {
        const int k = kb + c;
        LUB_T(p0);
        double x[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) x[r] = lub_lds(ra[r]);
        // Fast path (the common case): compare only the high words of |x| -- sign stripped, so for finite values
        // they order like the values -- and accept the winner when it is unique: exactly one lane of one warp holds the
        // maximum high word, that lane saw it once, and it is finite.  A unique, strictly largest high word means a
        // strictly largest |x|, so the choice equals the reference's; everything else (equal high words, i.e. possible
        // ties, and any inf / NaN candidate) takes the full lexicographic path below.
        int bkey = -1;
        unsigned bpp = 0xffffffffu;
        bool tie = false;
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const int key = act[r] ? (__double2hiint(x[r]) & 0x7fffffff) : -1;
            const bool gt = key > bkey;
            tie = gt ? false : (tie || (key == bkey && key >= 0));
            bkey = gt ? key : bkey;
            bpp = gt ? pp[r] : bpp;
        }
        int mkey = __reduce_max_sync(0xffffffffu, bkey);
        const unsigned wmask = __ballot_sync(0xffffffffu, bkey == mkey);
        bool unique = __popc(wmask) == 1 && !__any_sync(0xffffffffu, tie && bkey == mkey) && mkey < 0x7ff00000;
        unsigned mpp = __shfl_sync(0xffffffffu, bpp, __ffs(wmask) - 1);
        if (P > 1) {  // uniform
            if (lane == 0) ctl->xs[c & 1][pw] = make_int4(mkey, (int)mpp, unique ? 1 : 0, 0);
            lub_bar(nth);
            int gk = -2, cnt = 0, gu = 0;
            unsigned gpp = 0xffffffffu;
            for (int j = 0; j < P; ++j) {
                const int4 e = ctl->xs[c & 1][j];
                const bool gt = e.x > gk;
                cnt = gt ? 1 : (cnt + (e.x == gk ? 1 : 0));
                gu = gt ? e.z : gu;
                gpp = gt ? (unsigned)e.y : gpp;
                gk = gt ? e.x : gk;
            }
            unique = cnt == 1 && gu != 0;
            mpp = gpp;
        }
        bool forced = false;
        double mult, pv;
        unsigned pvoff;
        LUB_T(p1);
        LUB_ACC2(0, p0, p1);
        if (unique) {  // uniform over the panel's warps
            pvoff = (mpp & 0xffffu) * 8u;
            pv = lub_lds(colb + pvoff);
            mult = __drcp_rn(pv);  // == 1.0 / pv, both correctly rounded
        } else {
            // full key = bits of |x| (NaN -> -1, not a candidate -> -2), ties -> lowest logical position
            int bhi = -2;
            unsigned blo = 0u;
            bpp = 0xffffffffu;
            bool nank = false;
            unsigned pk = 0xffffffffu;  // the row at logical position k (forced choice)
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                const bool isnan_ = x[r] != x[r];
                const int hi = isnan_ ? -1 : (__double2hiint(x[r]) & 0x7fffffff);
                const unsigned lo = isnan_ ? 0xffffffffu : (unsigned)__double2loint(x[r]);
                const bool gt = hi > bhi || (hi == bhi && (lo > blo || (lo == blo && pp[r] < bpp)));
                const bool take = act[r] && gt;
                bhi = take ? hi : bhi;
                blo = take ? lo : blo;
                bpp = take ? pp[r] : bpp;
                const bool atk = act[r] && (pp[r] >> 16) == (unsigned)k;
                nank = nank || (isnan_ && atk);
                pk = atk ? pp[r] : pk;
            }
            int mhi = __reduce_max_sync(0xffffffffu, bhi);
            unsigned mlo = __reduce_max_sync(0xffffffffu, (bhi == mhi) ? blo : 0u);
            mpp = __reduce_min_sync(0xffffffffu, (bhi == mhi && blo == mlo) ? bpp : 0xffffffffu);
            int anynank = __any_sync(0xffffffffu, nank) ? 1 : 0;
            pk = __reduce_min_sync(0xffffffffu, pk);
            if (P > 1) {
                if (lane == 0) {
                    int* e = ctl->xs2[pw];
                    e[0] = mhi;
                    e[1] = (int)mlo;
                    e[2] = (int)mpp;
                    e[3] = (int)pk;
                    e[4] = anynank;
                }
                lub_bar(nth);
                mhi = -2;
                mlo = 0u;
                mpp = 0xffffffffu;
                pk = 0xffffffffu;
                anynank = 0;
                for (int j = 0; j < P; ++j) {
                    const int* e = ctl->xs2[j];
                    const int hi = e[0];
                    const unsigned lo = (unsigned)e[1], qq = (unsigned)e[2];
                    const bool gt = hi > mhi || (hi == mhi && (lo > mlo || (lo == mlo && qq < mpp)));
                    mhi = gt ? hi : mhi;
                    mlo = gt ? lo : mlo;
                    mpp = gt ? qq : mpp;
                    pk = min(pk, (unsigned)e[3]);
                    anynank |= e[4];
                }
            }
            forced = mhi < 0 || anynank != 0;  // nothing comparable, or NaN in row k: l = k
            if (forced) mpp = pk;
            pvoff = (mpp & 0xffffu) * 8u;
            pv = lub_lds(colb + pvoff);
            mult = 1.0 / pv;
        }
        const int l = (int)(mpp >> 16);
        LUB_T(p2);
        LUB_ACC2(1, p1, p2);
        if (pw == 0 && lane == 0) pivs[k] = l;
        if (pv == 0.0) {
            info = k + 1;
            ndone = c;
            break;
        }
        LUB_T(p3);
        LUB_ACC2(2, p2, p3);
        // logical positions: the pivot row goes to k, the row that sat at k goes to l; multipliers below the pivot
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const bool isl = pp[r] == mpp;
            const bool isk = (pp[r] >> 16) == (unsigned)k;
            const unsigned low = pp[r] & 0xffffu;
            pp[r] = isl ? (((unsigned)k << 16) | low) : (isk ? (((unsigned)l << 16) | low) : pp[r]);
            act[r] = act[r] && !isl;
            x[r] = x[r] * mult;
            if (act[r]) lub_sts(ra[r], x[r]);
        }
        LUB_T(p4);
        LUB_ACC2(3, p3, p4);
        // right-looking update of the remaining nd panel columns: all loads, then the arithmetic, then the stores
        // (specialised on nd so that nothing is issued for columns that do not exist)
        switch (nb - 1 - c) {
        case 7: lub_panel_update<RW, 7>(ra, colb + pvoff, ldb, x, act); break;
        case 6: lub_panel_update<RW, 6>(ra, colb + pvoff, ldb, x, act); break;
        case 5: lub_panel_update<RW, 5>(ra, colb + pvoff, ldb, x, act); break;
        case 4: lub_panel_update<RW, 4>(ra, colb + pvoff, ldb, x, act); break;
        case 3: lub_panel_update<RW, 3>(ra, colb + pvoff, ldb, x, act); break;
        case 2: lub_panel_update<RW, 2>(ra, colb + pvoff, ldb, x, act); break;
        case 1: lub_panel_update<RW, 1>(ra, colb + pvoff, ldb, x, act); break;
        default: break;
        }
#pragma unroll
        for (int r = 0; r < RW; ++r) ra[r] += ldb;
        colb += ldb;
        LUB_T(p5);
        LUB_ACC2(4, p4, p5);
        if (P == 1) __syncwarp();
        LUB_T(p6);
        LUB_ACC2(5, p5, p6);
    }
    LUB_T(q0);
    // permute the panel to the logical positions: every thread picks up its rows' nb values, then (behind a barrier,
    // the destination may be another warp's row) drops them at the logical position; publish the row moves
    bool moved[RW];
    int posc[RW];
    double t[LUB_NB][RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        const int pos = (int)(pp[r] >> 16);
        moved[r] = mine[r] && pos != rowc[r];
        posc[r] = moved[r] ? pos : rowc[r];
    }
    if (P > 1) lub_bar(nth); else __syncwarp();  // the last column's stores
    double* Mp = M + kb * ld;
#pragma unroll
    for (int c = 0; c < LUB_NB; ++c) {
        const int co = (c < nb) ? c * ld : 0;
#pragma unroll
        for (int r = 0; r < RW; ++r) t[c][r] = Mp[co + rowc[r]];
    }
    if (P > 1) lub_bar(nth); else __syncwarp();
#pragma unroll
    for (int c = 0; c < LUB_NB; ++c) {
#pragma unroll
        for (int r = 0; r < RW; ++r)
            if (moved[r] && c < nb) Mp[c * ld + posc[r]] = t[c][r];
    }
    int cntw = 0;
    unsigned mm[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        mm[r] = __ballot_sync(0xffffffffu, moved[r]);
        cntw += __popc(mm[r]);
    }
    int base = 0;
    if (cntw > 0) {  // warp-uniform
        if (lane == 0) base = atomicAdd(&ctl->nmoves, cntw);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const int idx = base + __popc(mm[r] & ((1u << lane) - 1u));
            if (moved[r]) ctl->src[idx] = rowc[r];
            if (moved[r]) ctl->dst[idx] = posc[r];
            base += __popc(mm[r]);
        }
    }
    if (pw == 0 && lane == 0) {
        ctl->ndone = ndone;
        if (info != 0) ctl->info = info;
    }
    __syncwarp();
    LUB_T(q1);
    LUB_ACC2(6, q0, q1);
    if (pw == 0) { LUB_ACC2_FLUSH; }
}

// Number of warps that share the panel of rows kb..n: one 32-row slot per lane while the CTA has the warps for it (at
// most half of an 8+-warp CTA, so that the look-ahead update keeps warps of its own), else two.
__device__ __forceinline__ int lub_panel_warps(int n, int kb, int nw, int pol) {
    const int nslots = (n - kb + 31) >> 5;
    int pmax = nw >= 8 ? (nw >> 1 < LUB_MAXPW ? nw >> 1 : LUB_MAXPW) : (nw < LUB_MAXPW ? nw : LUB_MAXPW);
    if (pol & 15) pmax = (pol & 15) < nw ? (pol & 15) : nw;
    if (nslots <= pmax) return nslots;
    return (nslots + 1) >> 1;  // two slots per lane (needs nw >= 4 for n <= 256)
}
// warp w of the CTA takes part in the panel if w < P (returns P); RP = ceil(n / 32) bounds the slots per lane
template <int RP>
__device__ __forceinline__ int lub_panel_any(double* M, int ld, int n, int kb, int nb, LubCtl* ctl, int* pivs, int w, int nw,
                                             int pol = 0) {
    const int nslots = (n - kb + 31) >> 5;
    const int P = lub_panel_warps(n, kb, nw, pol);
    if (w < P) {
        if (nslots <= P) lub_panel_mw<1>(M, ld, n, kb, nb, ctl, pivs, w, P);
        else if (RP > 1 && nslots <= 2 * P) lub_panel_mw<2>(M, ld, n, kb, nb, ctl, pivs, w, P);
        else __trap();
    }
    return P;
}

// ---- [S] row moves for the columns outside the panel + the U rows of the columns right of it --------------------------
// zcol[j] (right columns): 0 = no zero among the nbd U values of column j, 2 = all zero, 1 = mixed
// (Lp, ldl): the current panel's columns -- column c at Lp + c * ldl, rows by absolute index.  For the matrix resident in
// shared memory that is M + kb * ld; for the global-memory variant the panel sits in its own shared-memory buffer.
static __device__ __noinline__ void lub_swap_solve(double* M, int ld, int n, int kb, int nbp, int nbd, const LubCtl* ctl,
                                            const double* Lp, int ldl, double* Ub, unsigned char* zcol, int ldu, int tid,
                                            int nt) {
    // loads unconditional (unused move slots point at a valid row), stores predicated: see lub_panel
    const int nm = ctl->nmoves;
    int src[LUB_MAXMOVES], dst[LUB_MAXMOVES];
#pragma unroll
    for (int m = 0; m < LUB_MAXMOVES; ++m) {
        src[m] = ctl->src[m];
        dst[m] = ctl->dst[m];
    }
    for (int j = tid; j < n; j += nt) {
        if (j >= kb && j < kb + nbp) continue;
        double* col = M + j * ld;
        if (nm > 0) {  // CTA-uniform
            double tmp[LUB_MAXMOVES];
#pragma unroll
            for (int m = 0; m < LUB_MAXMOVES; ++m) tmp[m] = col[src[m]];
#pragma unroll
            for (int m = 0; m < LUB_MAXMOVES; ++m)
                if (m < nm) col[dst[m]] = tmp[m];
        }
        if (j >= kb + nbp) {  // right of the panel: kb + LUB_NB <= n here, so rows kb..kb+7 exist
            double u[LUB_NB];
#pragma unroll
            for (int c = 0; c < LUB_NB; ++c) u[c] = col[kb + c];
            int nz = 0;
#pragma unroll
            for (int c = 0; c < LUB_NB; ++c) {
                const double uc = u[c];
                const bool on = c < nbd && uc != 0.0;
                nz += on ? 1 : 0;
                const double* lc = Lp + c * ldl + kb;
#pragma unroll
                for (int i = c + 1; i < LUB_NB; ++i) {
                    const double t = u[i] - uc * lc[i];
                    u[i] = (on && i < nbd) ? t : u[i];
                }
            }
#pragma unroll
            for (int c = 0; c < LUB_NB; ++c)
                if (c < nbd) {
                    col[kb + c] = u[c];
                    Ub[c * ldu + j] = u[c];
                }
            zcol[j] = (nz == nbd) ? 0 : (nz == 0 ? 2 : 1);
        }
    }
}

// ---- [U] trailing update with K = nbd fused k-steps per register tile -----------------------------------------------------
// EXACT = false: every U value of the tile is non-zero, plain multiply / subtract.  EXACT = true: a zero u_kj skips
// the update of column j at step k like the reference does (it matters for signed zeros and non-finite multipliers).
template <bool TWO, bool EXACT, bool FULL>
__device__ __forceinline__ void lub_tile(double* M, int ld, int n, int K, int i0, int j0, const double (&l0)[LUB_NB],
                                         const double (&l1)[LUB_NB], const double* Ub, int ldu) {
    // FULL: K == LUB_NB and the whole (32|64) x 4 tile is inside the matrix -> no predicates at all.
    // Otherwise loads use clamped (valid) addresses and only the stores are predicated (see lub_panel).
    const bool v0 = FULL || i0 < n, v1 = TWO && (FULL || i0 + 32 < n);
    const int i0c = v0 ? i0 : n - 1, i1c = v1 ? i0 + 32 : n - 1;
    double* c0 = M + j0 * ld;
    double a0[LUB_CT], a1[LUB_CT];
    int jo[LUB_CT];
#pragma unroll
    for (int cc = 0; cc < LUB_CT; ++cc) {
        jo[cc] = (FULL || j0 + cc < n) ? cc * ld : 0;
        a0[cc] = c0[jo[cc] + i0c];
        if (TWO) a1[cc] = c0[jo[cc] + i1c];
    }
    const double* ur = Ub + j0;
#pragma unroll
    for (int k = 0; k < LUB_NB; ++k) {
        if (FULL || k < K) {
            const double2 ua = *reinterpret_cast<const double2*>(ur + k * ldu);
            const double2 ub = *reinterpret_cast<const double2*>(ur + k * ldu + 2);
            const double u[LUB_CT] = {ua.x, ua.y, ub.x, ub.y};
#pragma unroll
            for (int cc = 0; cc < LUB_CT; ++cc) {
                if (!EXACT || u[cc] != 0.0) {
                    a0[cc] -= u[cc] * l0[k];
                    if (TWO) a1[cc] -= u[cc] * l1[k];
                }
            }
        }
    }
#pragma unroll
    for (int cc = 0; cc < LUB_CT; ++cc) {
        const bool ok = FULL || j0 + cc < n;
        if (ok && v0) c0[jo[cc] + i0c] = a0[cc];
        if (TWO && ok && v1) c0[jo[cc] + i1c] = a1[cc];
    }
}

// one (row block rb, column tile ct) item of the trailing update: rows r0 + rb * RB .., columns c0 + 4 ct ..
template <int RB>
__device__ __forceinline__ void lub_update_item(double* M, int ld, int n, int r0, int K, int c0, int c1, const double* Lp, int ldl,
                                                const double* Ub, const unsigned char* zcol, int ldu, int lane, int rb, int ct,
                                                int& cur_rb, double (&l0)[LUB_NB], double (&l1)[LUB_NB]) {
    const int j0 = c0 + ct * LUB_CT;
    const unsigned zf = *reinterpret_cast<const unsigned*>(zcol + j0);  // flags of the tile's 4 columns
    if (zf == 0x02020202u) return;  // every u_kj is zero: nothing to do
    const int ib = r0 + rb * RB;
    const int i0 = ib + lane;
    const bool two = RB == 64 && ib + 32 < n;  // warp-uniform
    if (rb != cur_rb) {
        cur_rb = rb;
#pragma unroll
        for (int k = 0; k < LUB_NB; ++k) {
            const int kc = (k < K) ? k : 0;
            const double* ckc = Lp + kc * ldl;
            l0[k] = ckc[i0 < n ? i0 : n - 1];
            if (RB == 64) l1[k] = ckc[i0 + 32 < n ? i0 + 32 : n - 1];
        }
    }
    const bool full = K == LUB_NB && j0 + LUB_CT <= c1 && ib + (two ? 64 : 32) <= n;
    if (zf == 0u) {
        if (two) {
            if (full) lub_tile<true, false, true>(M, ld, n, K, i0, j0, l0, l1, Ub, ldu);
            else lub_tile<true, false, false>(M, ld, n, K, i0, j0, l0, l1, Ub, ldu);
        } else {
            if (full) lub_tile<false, false, true>(M, ld, n, K, i0, j0, l0, l1, Ub, ldu);
            else lub_tile<false, false, false>(M, ld, n, K, i0, j0, l0, l1, Ub, ldu);
        }
    } else {
        if (two) lub_tile<true, true, false>(M, ld, n, K, i0, j0, l0, l1, Ub, ldu);
        else lub_tile<false, true, false>(M, ld, n, K, i0, j0, l0, l1, Ub, ldu);
    }
}

// rows >= kb + K, columns [c0, c1); warp w of nw takes a contiguous chunk of (row block, column tile) items.
// RB = rows per block: 64 (two register rows per lane) for the bulk, 32 to spread a narrow strip over more warps.
template <int RB>
__device__ __noinline__ void lub_update(double* M, int ld, int n, int kb, int K, int c0, int c1, const double* Lp, int ldl,
                                        const double* Ub, const unsigned char* zcol, int ldu, int w, int nw, int lane) {
    const int r0 = kb + K;
    const int nrows = n - r0, ncols = c1 - c0;
    if (nrows <= 0 || ncols <= 0 || K <= 0) return;
    const int nrb = (nrows + RB - 1) / RB, nct = (ncols + LUB_CT - 1) / LUB_CT;
    const int items = nrb * nct, chunk = (items + nw - 1) / nw;
    int it = w * chunk;
    if (it >= items) return;
    const int itend = (it + chunk < items) ? it + chunk : items;
    int rb = it / nct, ct = it - rb * nct;
    int cur_rb = -1;
    double l0[LUB_NB], l1[LUB_NB];
    for (; it < itend; ++it) {
        lub_update_item<RB>(M, ld, n, r0, K, c0, c1, Lp, ldl, Ub, zcol, ldu, lane, rb, ct, cur_rb, l0, l1);
        if (++ct == nct) {
            ct = 0;
            ++rb;
        }
    }
}

// The same update drawn from a work queue in shared memory (*qhead, zeroed before the barrier that precedes the phase):
// the warps that are not in the panel start at once, the panel's warps join when the panel is done, and whoever is
// free takes the next chunk of LUB_QG column tiles of one row block.  Which warp updates which tile has no influence on
// the values (each element gets its K steps from one warp, in k order).
constexpr int LUB_QG = 3;
template <int RB>
__device__ __noinline__ void lub_update_q(double* M, int ld, int n, int kb, int K, int c0, int c1, const double* Lp, int ldl,
                                          const double* Ub, const unsigned char* zcol, int ldu, int lane, int* qhead) {
    const int r0 = kb + K;
    const int nrows = n - r0, ncols = c1 - c0;
    if (nrows <= 0 || ncols <= 0 || K <= 0) return;
    const int nrb = (nrows + RB - 1) / RB, nct = (ncols + LUB_CT - 1) / LUB_CT;
    const int ng = (nct + LUB_QG - 1) / LUB_QG, nchunks = nrb * ng;
    int cur_rb = -1;
    double l0[LUB_NB], l1[LUB_NB];
    for (;;) {
        int ch = 0;
        if (lane == 0) ch = atomicAdd(qhead, 1);
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch >= nchunks) break;
        const int rb = ch / ng, g = ch - rb * ng;
        const int cte = ((g + 1) * LUB_QG < nct) ? (g + 1) * LUB_QG : nct;
        for (int ct = g * LUB_QG; ct < cte; ++ct)
            lub_update_item<RB>(M, ld, n, r0, K, c0, c1, Lp, ldl, Ub, zcol, ldu, lane, rb, ct, cur_rb, l0, l1);
    }
}

// ---- [U] on the FP64 tensor cores: the TOLERANCE mode of the trailing update (B200IDA_LU_DMMA) ---------------------------
// A block step's update  A22 -= L21 * U12  (K = 8) really is a dense contraction, and for n >= 128 it dominates the
// flops.  mma.sync.m8n8k4.f64 (SASS: DMMA) does 8 x 8 x 4 fused multiply-adds per warp instruction, i.e. the work of 16
// un-fused DMUL + DADD warp instructions of lub_tile.  It is NOT the reference's arithmetic -- products are fused into
// the accumulation and the four k of one instruction are summed in hardware order -- so this path is an explicit mode:
// pivots still compare equal to the oracle's on well-conditioned inputs, factors agree to ~1e-15 relative (the tests
// bound them at 1e-12), and the default stays the bit-exact lub_update.
//
// The product is formed TRANSPOSED so that a thread's two accumulator elements are two consecutive ROWS of one column
// of the column-major matrix (one 16-byte load / store):  D^T[j][i] = sum_k U[k][j] * (-L[i][k]) + C^T[j][i], with the mma
// "m" dimension = 8 matrix columns, "n" = 8 matrix rows:
//   A fragment (8 x 4, row):  a  = U[k = lane % 4 (+4)][j0 + lane / 4]                     from Ub
//   B fragment (4 x 8, col):  b  = -L[i0 + lane / 4][k = lane % 4 (+4)]                    from the panel Lp
//   C / D fragment (8 x 8):   c0, c1 = M[j0 + lane / 4][i0 + 2 (lane % 4) + {0, 1}]        one double2
// A warp item is a block of 64 rows (its 16 negated L fragments stay in registers) times the column tiles of its chunk.
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
// rows >= kb + 8, columns [c0, c1) (c0 a multiple of 8), K == 8 fused k-steps; ld even and M 16-byte aligned.
static __device__ __noinline__ void lub_update_dmma(double* M, int ld, int n, int kb, int c0, int c1, const double* Lp, int ldl,
                                                    const double* Ub, int ldu, int w, int nw, int lane) {
    constexpr int RB = 64;
    const int r0 = kb + LUB_NB;
    const int nrows = n - r0, ncols = c1 - c0;
    if (nrows <= 0 || ncols <= 0) return;
    const int nrb = (nrows + RB - 1) / RB, nct = (ncols + 7) / 8;
    const int items = nrb * nct, chunk = (items + nw - 1) / nw;
    int it = w * chunk;
    if (it >= items) return;
    const int itend = (it + chunk < items) ? it + chunk : items;
    int rb = it / nct, ct = it - rb * nct;
    int cur_rb = -1;
    const int g = lane >> 2, q = lane & 3;
    double bl[8][2];  // -L fragments of the 8 row tiles of this row block, k halves 0..3 / 4..7
    for (; it < itend; ++it) {
        const int ib = r0 + rb * RB;
        if (rb != cur_rb) {
            cur_rb = rb;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int i = ib + 8 * t + g;
                const bool ok = i < n;
                const int ic = ok ? i : n - 1;
                const double v0 = Lp[q * ldl + ic], v1 = Lp[(q + 4) * ldl + ic];
                bl[t][0] = ok ? -v0 : 0.0;
                bl[t][1] = ok ? -v1 : 0.0;
            }
        }
        const int j = c0 + ct * 8 + g;  // this thread's matrix column
        const bool jok = j < c1;
        const int jc = jok ? j : c1 - 1;
        const double a0 = jok ? Ub[q * ldu + jc] : 0.0, a1 = jok ? Ub[(q + 4) * ldu + jc] : 0.0;
        double* col = M + (size_t)jc * ld;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = ib + 8 * t + 2 * q;  // first of this thread's two rows (even: r0 and the tile offsets are even)
            if (ib + 8 * t < n) {              // warp-uniform: the tile has at least one row
                const bool ok0 = jok && i < n, ok1 = jok && i + 1 < n;
                double d0 = 0.0, d1 = 0.0;
                if (ok1) {
                    const double2 c = *reinterpret_cast<const double2*>(col + i);
                    d0 = c.x;
                    d1 = c.y;
                } else if (ok0) {
                    d0 = col[i];
                }
                dmma_m8n8k4(d0, d1, a0, bl[t][0]);
                dmma_m8n8k4(d0, d1, a1, bl[t][1]);
                if (ok1) *reinterpret_cast<double2*>(col + i) = make_double2(d0, d1);
                else if (ok0) col[i] = d0;
            }
        }
        if (++ct == nct) {
            ct = 0;
            ++rb;
        }
    }
}

// In-place LU of the n x n column-major matrix M (leading dimension ld) in shared memory.  Every thread of the CTA
// calls it (blockDim.x a multiple of 32, at least 64).  RP >= ceil(n / 32).  Ub: lub_ub_bytes(n) of shared scratch,
// 16-byte aligned.  On return (after a barrier) ctl->info holds 0 or the 1-based singular column, pivs[0..n) the pivot
// rows; after a zero pivot in column k the matrix is the reference's partially factored state (steps < k applied).
//
// The TEAM flavour (CTA-per-system IDA kernel; B200IDA_LU_TEAM_PANEL in the Level-1 entry point): blockDim.x >= 128.
// In block step kb all warps first bring the NEXT panel (columns kb+nb .. kb+2nb) up to date; then its first
// P = lub_panel_warps() warps eliminate it together (lub_panel_mw) while the rest of the trailing update is a work queue
// that the other warps start on at once and the panel's warps join; three CTA barriers per block step.
// pol (experiments, tools/build_variants.sh -DIDA_TEAM_LUB_POL): low 4 bits = cap on P (0: the default rule), bit 4 =
// warps that share a scheduler with a panel warp (w % 4 < P) stay out of the update.
template <int RP>
__device__ __forceinline__ void lub_factor(double* M, int ld, int n, LubCtl* ctl, int* pivs, double* Ub, int pol = 0) {
    const int tid = threadIdx.x, nt = blockDim.x, w = tid >> 5, nw = nt >> 5, lane = tid & 31;
    const int ldu = lub_ldu(n);
    unsigned char* zcol = reinterpret_cast<unsigned char*>(Ub + LUB_NB * ldu);
    if (tid == 0) ctl->info = 0;
    for (int j = n + tid; j < ldu; j += nt) zcol[j] = 0;
    __syncthreads();
    LUB_T(t0);
    lub_panel_any<RP>(M, ld, n, 0, n < LUB_NB ? n : LUB_NB, ctl, pivs, w, nw, pol);
    LUB_T(t1);
    __syncthreads();
    LUB_ACC(0, t0, t1);
    for (int kb = 0; kb < n; kb += LUB_NB) {
        const int nb = (n - kb < LUB_NB) ? n - kb : LUB_NB;
        const int nbd = ctl->ndone;
        const int info = ctl->info;
        LUB_T(ta);
        lub_swap_solve(M, ld, n, kb, nb, nbd, ctl, M + kb * ld, ld, Ub, zcol, ldu, tid, nt);
        LUB_T(tb);
        __syncthreads();
        LUB_T(tc);
        const int c0 = kb + nb;                                  // first column right of this panel
        const int nb2 = (n - c0 < LUB_NB) ? n - c0 : LUB_NB;     // width of the next panel (<= 0: none)
        if (nb2 > 0) {
            // the strip that becomes the next panel first, spread over all warps in 32-row blocks
            lub_update<32>(M, ld, n, kb, nbd, c0, c0 + nb2, M + kb * ld, ld, Ub, zcol, ldu, w, nw, lane);
            if (tid == 0) ctl->qhead = 0;
            LUB_T(td);
            __syncthreads();
            LUB_T(te);
            // the first warps eliminate the next panel (one row per lane where the CTA has the warps for it); the rest of
            // the trailing update is a work queue that the other warps start on at once and the panel's warps join
            int P = 0;
            if (info == 0) P = lub_panel_any<RP>(M, ld, n, c0, nb2, ctl, pivs, w, nw, pol);
            if (!((pol & 16) && w >= P && (w & 3) < P))
                lub_update_q<64>(M, ld, n, kb, nbd, c0 + nb2, n, M + kb * ld, ld, Ub, zcol, ldu, lane, &ctl->qhead);
            LUB_T(tf);
            LUB_ACC(3, tc, td);
            LUB_ACC(4, td, te);
            LUB_ACC(5, te, tf);
            __syncthreads();
            LUB_T(tg);
            LUB_ACC(6, tf, tg);
        } else {
            __syncthreads();
        }
        LUB_ACC(1, ta, tb);
        LUB_ACC(2, tb, tc);
        if (info != 0) break;  // uniform: read before warp 0 could overwrite it
    }
}

// The Level-1 flavour: one-warp panel, static split of the look-ahead update (see lub_panel_1w).
template <int RP>
__device__ __forceinline__ void lub_factor_1w(double* M, int ld, int n, LubCtl* ctl, int* pivs, double* Ub) {
    const int tid = threadIdx.x, nt = blockDim.x, w = tid >> 5, nw = nt >> 5, lane = tid & 31;
    const int ldu = lub_ldu(n);
    unsigned char* zcol = reinterpret_cast<unsigned char*>(Ub + LUB_NB * ldu);
    if (tid == 0) ctl->info = 0;
    for (int j = n + tid; j < ldu; j += nt) zcol[j] = 0;
    __syncthreads();
    LUB_T(t0);
    if (w == 0) lub_panel_any_1w<RP>(M, ld, n, 0, n < LUB_NB ? n : LUB_NB, ctl, pivs);
    LUB_T(t1);
    __syncthreads();
    LUB_ACC(0, t0, t1);
    for (int kb = 0; kb < n; kb += LUB_NB) {
        const int nb = (n - kb < LUB_NB) ? n - kb : LUB_NB;
        const int nbd = ctl->ndone;
        const int info = ctl->info;
        LUB_T(ta);
        lub_swap_solve(M, ld, n, kb, nb, nbd, ctl, M + kb * ld, ld, Ub, zcol, ldu, tid, nt);
        LUB_T(tb);
        __syncthreads();
        LUB_T(tc);
        const int c0 = kb + nb;                                  // first column right of this panel
        const int nb2 = (n - c0 < LUB_NB) ? n - c0 : LUB_NB;     // width of the next panel (<= 0: none)
        if (nb2 > 0) {
            // the strip that becomes the next panel first, spread over all warps in 32-row blocks
            lub_update<32>(M, ld, n, kb, nbd, c0, c0 + nb2, M + kb * ld, ld, Ub, zcol, ldu, w, nw, lane);
            LUB_T(td);
            __syncthreads();
            LUB_T(te);
            // warp w runs on scheduler (sub-partition) w % 4: the panel warp's serial chain must not queue behind
            // FP64-saturating update warps, so with 8 warps warp 4 (same scheduler as warp 0) sits the update out
            if (w == 0) {
                if (info == 0) lub_panel_any_1w<RP>(M, ld, n, c0, nb2, ctl, pivs);
            } else if (nw <= 4) {
                lub_update<64>(M, ld, n, kb, nbd, c0 + nb2, n, M + kb * ld, ld, Ub, zcol, ldu, w - 1, nw - 1, lane);
            } else if ((w & 3) != 0) {
                lub_update<64>(M, ld, n, kb, nbd, c0 + nb2, n, M + kb * ld, ld, Ub, zcol, ldu, w - 1 - (w >> 2), nw - (nw >> 2), lane);
            }
            LUB_T(tf);
            LUB_ACC(3, tc, td);
            LUB_ACC(4, td, te);
            LUB_ACC(5, te, tf);
            __syncthreads();
            LUB_T(tg);
            LUB_ACC(6, tf, tg);
        } else {
            __syncthreads();
        }
        LUB_ACC(1, ta, tb);
        LUB_ACC(2, tb, tc);
        if (info != 0) break;  // uniform: read before warp 0 could overwrite it
    }
}

// ---- the same factorisation for a matrix that does not fit in shared memory (161 <= n <= 256) -----------------------------
// The matrix stays in global memory (column-major, ld = n: it lives in L2 while its CTA works on it); only the panels
// go through shared memory.  Two panel buffers of LUB_NB full-height columns: Pb[cur] holds the factored panel of the
// current block step -- the multipliers and L11 the [S] and [U] phases read come from there, not from global -- while
// warp 0 eliminates the next panel in Pb[nxt].  A factored panel is written back to global right after the barrier that
// ends its elimination (later panels apply their row moves to it there).  Trailing tiles are read and written once per
// block step, i.e. the L2 traffic of the update is 1/LUB_NB of the unblocked kernel's.
__device__ __forceinline__ void lub_panel_copy(double* dst, int ldd, const double* src, int lds, int r0, int n, int ncols,
                                               int w, int nw, int lane) {
    for (int c = w; c < ncols; c += nw)
        for (int i = r0 + lane; i < n; i += 32) dst[c * ldd + i] = src[c * lds + i];
}

template <int RP>
__device__ __forceinline__ void lub_factor_global(double* Mg, int n, LubCtl* ctl, int* pivs, double* Ub, double* Pb0,
                                                  double* Pb1, bool dmma = false) {
    const int tid = threadIdx.x, nt = blockDim.x, w = tid >> 5, nw = nt >> 5, lane = tid & 31;
    const int ld = n, ldp = n | 1, ldu = lub_ldu(n);
    unsigned char* zcol = reinterpret_cast<unsigned char*>(Ub + LUB_NB * ldu);
    double* Pcur = Pb0;
    double* Pnxt = Pb1;
    if (tid == 0) ctl->info = 0;
    for (int j = n + tid; j < ldu; j += nt) zcol[j] = 0;
    const int nb0 = n < LUB_NB ? n : LUB_NB;
    lub_panel_copy(Pcur, ldp, Mg, ld, 0, n, nb0, w, nw, lane);
    __syncthreads();
    if (w == 0) lub_panel_any_1w<RP>(Pcur, ldp, n, 0, nb0, ctl, pivs);  // panel 0: column c of the panel is column c
    __syncthreads();
    lub_panel_copy(Mg, ld, Pcur, ldp, 0, n, nb0, w, nw, lane);
    for (int kb = 0; kb < n; kb += LUB_NB) {
        const int nb = (n - kb < LUB_NB) ? n - kb : LUB_NB;
        const int nbd = ctl->ndone;
        const int info = ctl->info;
        lub_swap_solve(Mg, ld, n, kb, nb, nbd, ctl, Pcur, ldp, Ub, zcol, ldu, tid, nt);
        __syncthreads();
        const int c0 = kb + nb;
        const int nb2 = (n - c0 < LUB_NB) ? n - c0 : LUB_NB;
        if (nb2 > 0) {
            lub_update<32>(Mg, ld, n, kb, nbd, c0, c0 + nb2, Pcur, ldp, Ub, zcol, ldu, w, nw, lane);
            __syncthreads();
            lub_panel_copy(Pnxt, ldp, Mg + (size_t)c0 * ld, ld, c0, n, nb2, w, nw, lane);  // rows >= c0 of the strip
            __syncthreads();
            if (w == 0) {
                // lub_panel addresses column kb' + c at M + (kb' + c) * ld: shift the base so that it lands in the buffer
                if (info == 0) lub_panel_any_1w<RP>(Pnxt - (size_t)c0 * ldp, ldp, n, c0, nb2, ctl, pivs);
            } else if ((w & 3) != 0 || nw <= 4) {
                const int uw = nw <= 4 ? w - 1 : w - 1 - (w >> 2), unw = nw <= 4 ? nw - 1 : nw - (nw >> 2);
                if (dmma && nbd == LUB_NB) lub_update_dmma(Mg, ld, n, kb, c0 + nb2, n, Pcur, ldp, Ub, ldu, uw, unw, lane);
                else lub_update<64>(Mg, ld, n, kb, nbd, c0 + nb2, n, Pcur, ldp, Ub, zcol, ldu, uw, unw, lane);
            }
            __syncthreads();
            if (info == 0) lub_panel_copy(Mg + (size_t)c0 * ld, ld, Pnxt, ldp, c0, n, nb2, w, nw, lane);
            double* t = Pcur;
            Pcur = Pnxt;
            Pnxt = t;
        }
        __syncthreads();
        if (info != 0) break;
    }
}

}  // namespace b200ida
