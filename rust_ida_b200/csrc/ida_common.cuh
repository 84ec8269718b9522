// ida_common.cuh -- constants, state layout and small helpers shared by the ensemble-IDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200ida.h"

namespace b200ida {

// ---- algorithm constants: src/constants.rs:4-31 ; RATEMAX src/ida_nls.rs:15 ---------------------------
constexpr int MXORDP1 = 6;
constexpr int MAXORD = 5;
constexpr int MXSTEP = 500;
constexpr int MXNCF = 10;
constexpr int MXNEF = 10;
constexpr int MAXNLSIT = 4;
constexpr double EPCON = 0.33;
constexpr double XRATE = 0.25;
constexpr double RATEMAX = 0.9;
constexpr double HMAX_INV = 0.0;
constexpr double DBL_EPS = 2.220446049250313e-16;

constexpr int MAX_ATOL = 256;   // largest n any kernel supports
constexpr int MAX_TOUTS = 64;   // ladder length per launch

// ---- per-instance persistent state in HBM, structure-of-arrays over the ensemble -----------------------
// doubles: dstate[field][batch], ints: istate[field][batch].  Field offsets for system size n:
struct Layout {
    int n;
    __host__ __device__ constexpr explicit Layout(int n_) : n(n_) {}
    // doubles
    __host__ __device__ constexpr int phi(int j, int i) const { return j * n + i; }  // [6][n]
    __host__ __device__ constexpr int psi(int j) const { return 6 * n + j; }
    __host__ __device__ constexpr int alpha(int j) const { return 6 * n + 6 + j; }
    __host__ __device__ constexpr int beta(int j) const { return 6 * n + 12 + j; }
    __host__ __device__ constexpr int sigma(int j) const { return 6 * n + 18 + j; }
    __host__ __device__ constexpr int gamma(int j) const { return 6 * n + 24 + j; }
    __host__ __device__ constexpr int ee(int i) const { return 6 * n + 30 + i; }
    __host__ __device__ constexpr int ewt(int i) const { return 7 * n + 30 + i; }
    __host__ __device__ constexpr int yy(int i) const { return 8 * n + 30 + i; }
    __host__ __device__ constexpr int yp(int i) const { return 9 * n + 30 + i; }
    __host__ __device__ constexpr int scal(int s) const { return 10 * n + 30 + s; }
    __host__ __device__ constexpr int J(int i, int j) const { return 10 * n + 30 + NSCAL + j * n + i; }  // col-major
    __host__ __device__ constexpr int ndouble() const { return 10 * n + 30 + NSCAL + n * n; }
    // ints
    __host__ __device__ constexpr int piv(int i) const { return NINT + i; }
    __host__ __device__ constexpr int nint() const { return NINT + n; }

    enum Scal { HH, HUSED, RR, CJ, CJOLD, CJLAST, CJRATIO, SS, TOLDEL, EPS_NEWT, TN, TRETLAST, H0U, TOLSF, TRET, NSCAL };
    enum Int { KK, KUSED, KNEW, PHASE, NS, SETUP_DONE, JCUR, NST, NRE, NJE, NNI, NETF, NCFN, NSETUPS, NLSCF, NATT, NINT };
};

struct TolParams {   // TolControlSV (src/tol_control.rs:51-83), passed by value to kernels
    double rtol;
    double atol[MAX_ATOL];
};

struct LadderParams {
    int ntout;
    double touts[MAX_TOUTS];
};

// per-attempt state trace of ONE instance (the device side of the reference's `data_trace` feature, which serialises
// the whole solver at the top of every step attempt: src/lib.rs:297-301, 635-639).  Row c of d / i is the instance's
// complete persistent state (the Layout fields, as doubles / ints) when its c-th attempt begins.
struct TraceParams {
    double* d;       // [max][Layout::ndouble()]
    int* i;          // [max][Layout::nint()]
    int* count;      // [1] attempts seen so far (rows beyond max are counted, not stored)
    int max;
    long long id;    // caller-visible instance id
};

// rootfinding (src/impl_r_check.rs): per-instance root state persisted next to the solver state, and the event log
// a ladder launch fills -- the device plays the caller of examples/roberts.rs:117-140, calling solve() again after
// every Root return and recording (t, get_root_info) of each.
struct RootParams {
    double* rd;         // [4 NR + 5][batch]: glo, ghi, grout, iroots [NR each], tlo, thi, trout, ttol, toutc
    int* ri;            // [NR + 2][batch]: gactive [NR], irfnd, nge
    double* ev_t;       // [max_ev][batch] root locations
    double* ev_iroots;  // [max_ev][NR][batch] get_root_info at each event (a Scalar array in the reference)
    int* ev_count;      // [batch] events seen (beyond max_ev: counted, not stored)
    int* nge_out;       // [batch] or null: get_num_g_evals after the launch
    int max_ev;
};

// deferred output interpolation (production thread kernels): a lane that passes an output time only SNAPSHOTS what
// get_solution (src/lib.rs:1274-1343) reads -- phi[0..kord], psi[0..kord-1], kused, tn, hh, hused -- and a second,
// lane-coherent kernel (ida_interp_kernel) evaluates y(tout), y'(tout) for every (instance, output time) afterwards:
// same operations on the same operands, so the values are bit-identical, but the ~400-instruction interpolation is no
// longer issued once per lane per output time by a warp whose lanes reach their output times on different attempts.
// snap: [batch][ntout][snap_doubles(n)] by caller-visible instance id (slots are re-assigned when the ensemble is
// re-grouped between segments; ids are not), AoS so that the lone lane writes whole sectors and the second pass streams; fields 0..6n+5 are phi / psi at their Layout offsets, then kused, tn, hh, hused.  valid: [batch][ntout] bytes,
// written for every call of every instance (1 = snapshot taken, 0 = the call returned some other way and interpolated
// in line).
struct DeferParams {
    double* snap;
    unsigned char* valid;
};
__host__ __device__ constexpr int snap_doubles(int n) { return 6 * n + 6 + 4; }

// segmented launches (re-grouping of the ensemble between segments, ida_engine.cu): a launch runs at most max_trips
// trips of the flattened solve loop, then parks the loop's own state -- which output time the lane is on, whether a
// step is in progress, the per-call / per-step counters -- next to the solver state, and the next launch resumes it.
// li: [2][batch] ints (word 0: it | need_prologue << 8 | in_step << 9 | ncf << 10 | nef << 18 ; word 1: nstloc),
// ld: [1][batch] doubles (saved_t).  max_trips < 0: one unsegmented launch (nothing is loaded or stored).
struct SegParams {
    int max_trips = -1;
    int resume = 0;      // 0: the launch starts the API call (fresh loop state), 1: continue from li / ld
    const int* li = nullptr;   // read when resume != 0: words at li[slot], li[stride + slot]
    const double* ld = nullptr;
    int* li_out = nullptr;     // written at the end of the launch: a DIFFERENT buffer, so that the exact pass of the same
    double* ld_out = nullptr;  // segment still finds the loop state its flagged lanes started the segment with
    long long stride = 0;      // distance between the two loop-state words (batch for [2][batch] arrays, 1 inside a row)
};

// How a launch of the thread-per-instance kernel finds and leaves the per-instance state.  Default: the engine's
// structure-of-arrays buffers, slot = thread.  Between the segments of a segmented solve the state travels through an
// EXCHANGE buffer of one contiguous row per instance -- [ndouble state doubles | saved_t | params] and
// [nint state ints | loop word | nstloc | instance id] -- which a lane writes at its slot and the next segment's lane
// reads at order[slot]: the re-grouping of the slots costs no pass over the state at all, it is the load of the next
// segment (each lane reads its own contiguous row; whole sectors either way).
struct IoParams {
    const double* xd_in = nullptr;  // exchange rows in (null: SoA in)
    const int* xi_in = nullptr;
    const int* order = nullptr;     // row of slot s = order[s] (null: s)
    double* xd_out = nullptr;       // exchange rows out, row = slot (null: SoA out)
    int* xi_out = nullptr;
    double* params_out = nullptr;   // SoA out only: the instance's parameters and id follow it to its final slot
    int* perm_out = nullptr;
};
__host__ __device__ constexpr int xrow_doubles(int n, int np) { return (10 * n + 45 + n * n + 1 + np + 1) & ~1; }
__host__ __device__ constexpr int xrow_ints(int n) { return (16 + n + 3 + 3) & ~3; }

// select a[k] from a 6-entry register array without dynamic indexing
__device__ __forceinline__ double pick6(const double (&a)[MXORDP1], int k) {
    double v = a[0];
#pragma unroll
    for (int j = 1; j < MXORDP1; ++j) v = (k == j) ? a[j] : v;
    return v;
}

}  // namespace b200ida
