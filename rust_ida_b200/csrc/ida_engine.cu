// ida_engine.cu -- C ABI (include/b200ida.h, "Level 2") of the ensemble IDA engine and its kernels.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false  (no FMA contraction:
// the reference's Rust arithmetic is un-fused; see DESIGN.md "Arithmetic policy").
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>

#include <cub/cub.cuh>

#include "host_util.h"
#include "ida_common.cuh"
#include "ida_team.cuh"
#include "ida_thread_launch.cuh"
#include "problems.cuh"

namespace b200ida {

long long& launch_counter() {
    static long long n = 0;
    return n;
}

static const UserProblemHooks* g_user_hooks = nullptr;
void register_user_problem(const UserProblemHooks* h) { g_user_hooks = h; }
const UserProblemHooks* user_problem_hooks() { return g_user_hooks; }

char* last_error_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

// ---- kernels -----------------------------------------------------------------------------------------

// Team-per-instance solve (ida_team.cuh): persistent teams (warps or whole CTAs) stride over the ensemble.
template <class Team>
__global__ void __launch_bounds__(Team::MAXT, Team::MINB) ida_solve_team_kernel(double* __restrict__ ds, int* __restrict__ is, const double* __restrict__ params,
                                      int64_t B, int n, int kind, int nparams, double rtol, const double* __restrict__ atol,
                                      LadderParams lad, double* __restrict__ yy_traj, int32_t* __restrict__ status,
                                      unsigned arena_bytes, const int* __restrict__ perm, TraceParams tr,
                                      const double* __restrict__ idmask) {
    extern __shared__ __align__(16) unsigned char team_smem[];
    const int tpc = Team::teams_per_cta();
    IdaTeam<Team> s;
    s.bind(team_smem + (size_t)Team::team_in_cta() * arena_bytes, n);
    s.rtol = rtol;
    s.atol = atol;
    s.idmask = idmask;  // null unless ida_suppressalg is set
    s.tr = tr;
    s.prob.kind = kind;
    s.prob.n = n;
    int M = 1;
    while (M * M < n) ++M;
    s.prob.M = M;
    for (int64_t b = (int64_t)blockIdx.x * tpc + Team::team_in_cta(); b < B; b += (int64_t)gridDim.x * tpc) {
        for (int k = 0; k < 4; ++k) s.prob.p[k] = k < nparams ? params[(int64_t)k * B + b] : 0.0;
        s.load(ds, is, B, b);
        s.solve_ladder(lad, yy_traj, status, B, perm ? (int64_t)perm[b] : b);
        s.store(ds, is, B, b);
    }
}

// Ida::new (src/lib.rs:278-405): phi[0] = yy0, phi[1] = yp0, yy = yy0, yp = yp0 (src/ida_nls.rs:82-83),
// tolsf = 1, everything else 0 (the state buffers are zeroed before this kernel).
__global__ void ida_init_kernel(double* __restrict__ ds, const double* __restrict__ yy0, const double* __restrict__ yp0,
                                int n, int64_t B, const int* __restrict__ perm) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t ob = perm ? (int64_t)perm[b] : b;
    const Layout lay(n);
    for (int i = 0; i < n; ++i) {
        const double y = yy0[(int64_t)i * B + ob], yp = yp0[(int64_t)i * B + ob];
        ds[(int64_t)lay.phi(0, i) * B + b] = y;
        ds[(int64_t)lay.phi(1, i) * B + b] = yp;
        ds[(int64_t)lay.yy(i) * B + b] = y;
        ds[(int64_t)lay.yp(i) * B + b] = yp;
    }
    ds[(int64_t)lay.scal(Layout::TOLSF) * B + b] = 1.0;
}

// slot order -> caller order: dst[r][perm[s]] = (TD)src[r][s]  (perm == nullptr: identity)
template <class TS, class TD>
__global__ void unpermute_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int rows, int64_t B,
                                 const int* __restrict__ perm) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    const int64_t o = perm ? (int64_t)perm[s] : s;
    for (int r = 0; r < rows; ++r) dst[(int64_t)r * B + o] = (TD)src[(int64_t)r * B + s];
}
// caller order -> slot order: dst[r][s] = src[r][perm[s]]
__global__ void permute_rows_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int64_t B,
                                    const int* __restrict__ perm) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    const int64_t o = perm ? (int64_t)perm[s] : s;
    for (int r = 0; r < rows; ++r) dst[(int64_t)r * B + s] = src[(int64_t)r * B + o];
}
// Ida::get_dky (src/lib.rs:424-529) for every instance of the ensemble at one (t, k): k-th derivative of the
// interpolating polynomial of the last step, from the persistent phi / psi / tn / hh / hused / kused.  One thread per
// instance over the SoA state (coalesced rows), written straight to caller order.  HBM-bound: reads (kused-k+1) n + ~10
// doubles, writes n.  The recurrences are followed as the reference has them (inner loops to kused-k+1 for every i, so
// for k >= 2 the higher coefficients stay 0; psij_1 carried across passes); k = 0 with kused = 5 would index past the
// reference's 6-entry array (it panics), here the recurrence stops at j = 5, which the sum does not use anyway.
// Per-instance status: 0, B200IDA_ST_BAD_K, B200IDA_ST_BAD_T (dky of that instance is left untouched).
__global__ void __launch_bounds__(256) ida_dky_kernel(const double* __restrict__ ds, const int* __restrict__ is, int n, int64_t B,
                                                      double t, int k, double* __restrict__ dky, int* __restrict__ status,
                                                      const int* __restrict__ perm) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    const int64_t o = perm ? (int64_t)perm[s] : s;
    const Layout lay(n);
    const int kused = is[(int64_t)Layout::KUSED * B + s];
    if (k < 0 || k > kused) {
        if (status) status[o] = B200IDA_ST_BAD_K;
        return;
    }
    const double tn = ds[(int64_t)lay.scal(Layout::TN) * B + s], hh = ds[(int64_t)lay.scal(Layout::HH) * B + s],
                 hused = ds[(int64_t)lay.scal(Layout::HUSED) * B + s];
    const double sgn = (hh > 0.0 || (hh == 0.0 && !signbit(hh))) ? 1.0 : -1.0;
    const double tfuzz = 100.0 * DBL_EPS * (fabs(tn) + fabs(hh)) * sgn;
    const double tp = tn - hused - tfuzz;
    if ((t - tp) * hh < 0.0) {
        if (status) status[o] = B200IDA_ST_BAD_T;
        return;
    }
    double psi[MXORDP1], cjk[MXORDP1], cjk_1[MXORDP1];
#pragma unroll
    for (int j = 0; j < MXORDP1; ++j) {
        psi[j] = ds[(int64_t)lay.psi(j) * B + s];
        cjk[j] = 0.0;
        cjk_1[j] = 0.0;
    }
    const double delt = t - tn;
    double psij_1 = 0.0;
    const int jtop = min(kused - k + 1, MXORDP1 - 1);
    // fully unrolled with predicates so that psi / cjk stay in registers
#pragma unroll
    for (int i = 0; i < MXORDP1; ++i) {
        if (i <= k) {
            const double si = (double)i;
            if (i == 0) {
                cjk[0] = 1.0;
            } else {
                cjk[i] = cjk[i - 1] * si / psi[i - 1];
                psij_1 = psi[i - 1];
            }
#pragma unroll
            for (int j = i + 1; j < MXORDP1; ++j) {
                if (j <= jtop) {
                    cjk[j] = (si * cjk_1[j - 1] + cjk[j - 1] * (delt + psij_1)) / psi[j - 1];
                    psij_1 = psi[j - 1];
                }
            }
#pragma unroll
            for (int j = i + 1; j < MXORDP1; ++j)
                if (j <= jtop) cjk_1[j] = cjk[j];
        }
    }
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < MXORDP1; ++j)
            if (j >= k && j <= kused) acc = acc + ds[(int64_t)lay.phi(j, i) * B + s] * cjk[j];
        dky[(int64_t)i * B + o] = acc;
    }
    if (status) status[o] = 0;
}
// Morton key of the (min/max-normalised) parameter vector of each instance: instances that are close in
// parameter space get adjacent slots, so the lanes of a warp take similar step/Newton/retry decisions.
__global__ void order_key_kernel(const double* __restrict__ params, const double* __restrict__ lo, const double* __restrict__ hi,
                                 int nparams, int64_t B, unsigned* __restrict__ keys, int* __restrict__ ids) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int bits = 30 / nparams > 10 ? 10 : 30 / nparams;
    unsigned key = 0;
    for (int j = 0; j < nparams; ++j) {
        const double range = hi[j] - lo[j];
        double u = range > 0.0 ? (params[(int64_t)j * B + b] - lo[j]) / range : 0.0;
        u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        unsigned q = (unsigned)(u * (double)((1u << bits) - 1u));
        for (int k = 0; k < bits; ++k) key |= ((q >> k) & 1u) << (k * nparams + j);
    }
    keys[b] = key;
    ids[b] = (int)b;
}

// ---- re-grouping of the ensemble between the segments of a solve ------------------------------------
// key: finished instances last; otherwise (order kk, phase, ns, quantised log2 |h|) -- lanes that agree on these walk the
// same loop trip counts and take the same branches for a while.  The radix sort is stable, so equal keys keep their
// previous (parameter-Morton) neighbourhood.
__global__ void regroup_key_kernel(const double* __restrict__ xd, const int* __restrict__ xi, int n, int XD, int XI, int64_t B, int ntout,
                                   unsigned* __restrict__ keys, int* __restrict__ ids) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    const Layout lay(n);
    const int* ir = xi + s * XI;
    const unsigned done = ((ir[lay.nint()] & 0xff) >= ntout) ? 1u : 0u;  // loop word 0: output-time index
    const unsigned kk = (unsigned)ir[Layout::KK] & 7u;
    const unsigned ph = (unsigned)ir[Layout::PHASE] & 1u;
    const unsigned ns = (unsigned)min(ir[Layout::NS], 7) & 7u;
    const unsigned hq = ((unsigned)__double2hiint(fabs(xd[s * XD + lay.scal(Layout::HH)])) >> 16) & 0x7fffu;  // exponent + 4 mantissa bits
    keys[s] = done << 23 | kk << 20 | ph << 19 | ns << 16 | hq;
    ids[s] = (int)s;
}
__global__ void invert_perm_kernel(const int* __restrict__ perm, int* __restrict__ inv, int64_t B) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < B) inv[perm[s]] = (int)s;
}

__global__ void detpow_kernel(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out,
                              int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = detpow::det_pow(x[i], y[i]);
}

// q = a / b and r = 1 / b through fastdiv.cuh: three quotients per thread in one batch, like the solver's call sites
__global__ void fastdiv_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ q,
                               double* __restrict__ r, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = 3 * t;
    if (i0 + 2 < n) {
        const double aa[3] = {a[i0], a[i0 + 1], a[i0 + 2]}, bb[3] = {b[i0], b[i0 + 1], b[i0 + 2]};
        double qq[3], rr[3];
        div_batch<3>(aa, bb, qq);
        rcp_batch<3>(bb, rr);
        bool okz;
        const double qz = div_fast_z(aa[0], bb[0], okz);  // element 0 of every triple: the zero-numerator variant
        if (okz) qq[0] = qz;
        for (int k = 0; k < 3; ++k) {
            q[i0 + k] = qq[k];
            r[i0 + k] = rr[k];
        }
    } else {
        for (int64_t i = i0; i < n; ++i) {
            bool ok;
            const double qz = div_fast_z(a[i], b[i], ok);  // the zero-numerator variant, with the cold fallback of div1
            q[i] = ok ? qz : div_slow(a[i], b[i]);
            r[i] = rcp1(b[i]);
        }
    }
}

// FP64 issue-rate microbenchmark: 8 independent dependency chains per thread, register resident.
template <int KIND>
__global__ void __launch_bounds__(256) fp64_rate_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + (double)(threadIdx.x + i) * 1e-9;
    const double m = 1.0000000001, c = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0) a[i] = __fma_rn(a[i], m, c);
            else if (KIND == 1) a[i] = __dadd_rn(a[i], c);
            else a[i] = __dmul_rn(a[i], m);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456) out[0] = s;  // keep the loop alive
}

}  // namespace b200ida

using namespace b200ida;

// ---- the opaque handle ---------------------------------------------------------------------------------
struct b200ida_ensemble {
    int problem = 0, n = 0, nparams = 0, device = 0;
    int nr = 0;  // Root::num_roots of the problem (Roberts 2; a user problem: its functor's NR)
    int64_t B = 0;
    double rtol = 0.0;
    double atol[MAX_ATOL];
    cudaStream_t stream = nullptr;
    double* dstate = nullptr;  // [ndouble][B]
    int* istate = nullptr;     // [nint][B]
    double* params = nullptr;  // [nparams][B]
    double* d_atol = nullptr;  // [n]
    int kernel = B200IDA_KERNEL_AUTO;
    int reorder = B200IDA_REORDER_AUTO;  // captured from the process default at create (b200ida_set_default_reorder)
    int* perm = nullptr;  // [B] slot -> caller-visible instance id, or null (identity)
    TraceParams trace = {nullptr, nullptr, nullptr, 0, -1};  // b200ida_set_trace; id < 0: off
    RootParams roots = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};  // b200ida_set_roots; max_ev == 0: off
    DevBuf root_d, root_i;  // the per-instance root state (engine-owned)
    // ida_id / ida_suppressalg (src/lib.rs:102,122; b200ida_set_id): host copy as 1.0 / 0.0 and a device copy for the
    // team kernels
    double idmask[MAX_ATOL];
    double* d_idmask = nullptr;
    int suppressalg = 0;
    // production thread path: second state buffers (the FAST pass writes there; swapped with dstate / istate after the
    // launch) and the per-slot flags of the instances it handed to the exact pass
    double* dstate2 = nullptr;
    int* istate2 = nullptr;
    // re-grouping between segments (thread-per-instance production path, large re-ordered ensembles)
    int regroup_trips = 96, regroup_segments = 4;  // b200ida_set_regroup; segments <= 1: off
    DevBuf params2, perm2, inv_perm, perm0, params0, rg_keys, rg_keys2, rg_ids, rg_order, rg_tmp;
    DevBuf xd[2], xi[2];  // the exchange buffers the state travels through between segments (IoParams, ida_common.cuh)
    size_t rg_tmp_bytes = 0;
    bool regrouped = false;  // perm / params differ from the parameter-Morton order built at (re)init
    DevBuf poisoned, todo;  // todo: [batch + 1] ints, the compacted slot list of the exact pass and its length
    bool fast_ran = false;  // the last solve went through the FAST pass (b200ida_get_fallback_count)
    DevBuf snap, snap_valid;  // deferred output interpolation (DeferParams), grown to the longest ladder seen
    size_t snap_cap = 0, valid_cap = 0;
    // persistent device scratch for the host-buffer (end-to-end) entry points and the order builder: a few blocks
    // grown on demand, sub-allocated per API call (scratch_reset() rewinds), freed only by destroy -- so there is no
    // cudaMalloc / cudaFree (and no implicit device synchronisation) on the per-step path once warmed up
    static constexpr int MAX_BLOCKS = 16;
    void* blk[MAX_BLOCKS] = {nullptr};
    size_t blk_cap[MAX_BLOCKS] = {0}, blk_used[MAX_BLOCKS] = {0};
    int nblk = 0;
    void scratch_reset() {
        for (int i = 0; i < nblk; ++i) blk_used[i] = 0;
    }
    cudaError_t scratch_take(size_t bytes, void** out) {
        bytes = (bytes + 255) / 256 * 256;
        if (bytes == 0) bytes = 256;
        for (int i = 0; i < nblk; ++i) {
            if (blk_cap[i] - blk_used[i] >= bytes) {
                *out = static_cast<unsigned char*>(blk[i]) + blk_used[i];
                blk_used[i] += bytes;
                return cudaSuccess;
            }
        }
        if (nblk == MAX_BLOCKS) {
            // every block is rewound between API calls (scratch_reset after a stream sync): drop the idle ones and
            // replace them by one block that holds everything this call has asked for so far, twice over
            size_t total = bytes, keep = 0;
            for (int i = 0; i < nblk; ++i) total += blk_used[i];
            for (int i = 0; i < nblk; ++i) {
                if (blk_used[i] == 0) {
                    cudaFree(blk[i]);
                } else {
                    blk[keep] = blk[i];
                    blk_cap[keep] = blk_cap[i];
                    blk_used[keep] = blk_used[i];
                    ++keep;
                }
            }
            nblk = keep;
            if (nblk == MAX_BLOCKS) return cudaErrorMemoryAllocation;
            cudaError_t err2 = cudaMalloc(&blk[nblk], 2 * total);
            if (err2 != cudaSuccess) return err2;
            blk_cap[nblk] = 2 * total;
            blk_used[nblk] = bytes;
            *out = blk[nblk];
            ++nblk;
            return cudaSuccess;
        }
        const size_t cap = bytes > ((size_t)32 << 20) ? bytes : ((size_t)32 << 20);
        cudaError_t err = cudaMalloc(&blk[nblk], cap);
        if (err != cudaSuccess) return err;
        blk_cap[nblk] = cap;
        blk_used[nblk] = bytes;
        *out = blk[nblk];
        ++nblk;
        return cudaSuccess;
    }
    void scratch_free() {
        for (int i = 0; i < nblk; ++i) cudaFree(blk[i]);
        nblk = 0;
    }
};

// process-wide DEFAULT for handles created afterwards (each handle keeps its own copy: changing the default never
// re-orders a live ensemble); an atomic so that concurrent creators see a whole value
static std::atomic<int> g_reorder_mode{B200IDA_REORDER_AUTO};

// (re)build e->perm from parameters that are already on the device in CALLER order
static int build_order(b200ida_ensemble* e, const double* d_params_caller_order) {
    const bool want = e->reorder == B200IDA_REORDER_ON ||
                      (e->reorder == B200IDA_REORDER_AUTO && e->problem != B200IDA_PROBLEM_HEAT2D && e->B >= 4096);
    if (!want || e->nparams < 1 || e->nparams > 8 || e->B >= (int64_t)1 << 31) {
        if (e->perm) cudaFree(e->perm);
        e->perm = nullptr;
        return B200IDA_OK;
    }
    const int64_t B = e->B;
    const int P = e->nparams;
    cudaStream_t st = e->stream;
    if (!e->perm) B200_CUDA_TRY(cudaMalloc(&e->perm, (size_t)B * sizeof(int)));
    struct PermGuard {  // an error return below must not leave an unfilled permutation attached to the handle
        b200ida_ensemble* e;
        bool done = false;
        ~PermGuard() {
            if (!done && e->perm) {
                cudaFree(e->perm);
                e->perm = nullptr;
            }
        }
    } guard{e};
    size_t tb1 = 0, tb2 = 0, tb3 = 0;
    cub::DeviceReduce::Min(nullptr, tb1, d_params_caller_order, (double*)nullptr, (int)B, st);
    cub::DeviceReduce::Max(nullptr, tb2, d_params_caller_order, (double*)nullptr, (int)B, st);
    cub::DeviceRadixSort::SortPairs(nullptr, tb3, (unsigned*)nullptr, (unsigned*)nullptr, (int*)nullptr, (int*)nullptr, (int)B, 0, 32, st);
    size_t tbytes = tb1 > tb2 ? tb1 : tb2;
    tbytes = tbytes > tb3 ? tbytes : tb3;
    void *p_lohi, *p_keys, *p_keys2, *p_ids, *p_tmp;
    B200_CUDA_TRY(e->scratch_take(2 * (size_t)P * sizeof(double), &p_lohi));
    B200_CUDA_TRY(e->scratch_take((size_t)B * sizeof(unsigned), &p_keys));
    B200_CUDA_TRY(e->scratch_take((size_t)B * sizeof(unsigned), &p_keys2));
    B200_CUDA_TRY(e->scratch_take((size_t)B * sizeof(int), &p_ids));
    B200_CUDA_TRY(e->scratch_take(tbytes, &p_tmp));
    double* lo = static_cast<double*>(p_lohi);
    double* hi = lo + P;
    unsigned *keys = static_cast<unsigned*>(p_keys), *keys2 = static_cast<unsigned*>(p_keys2);
    int* ids = static_cast<int*>(p_ids);
    for (int j = 0; j < P; ++j) {
        B200_CUDA_TRY(cub::DeviceReduce::Min(p_tmp, tbytes, d_params_caller_order + (size_t)j * B, lo + j, (int)B, st));
        B200_CUDA_TRY(cub::DeviceReduce::Max(p_tmp, tbytes, d_params_caller_order + (size_t)j * B, hi + j, (int)B, st));
    }
    const int tb = 256;
    order_key_kernel<<<(unsigned)((B + tb - 1) / tb), tb, 0, st>>>(d_params_caller_order, lo, hi, P, B, keys, ids);
    count_launch();
    B200_CUDA_TRY(cudaGetLastError());
    B200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(p_tmp, tbytes, keys, keys2, ids, e->perm, (int)B, 0, 32, st));
    guard.done = true;
    return B200IDA_OK;  // scratch is persistent and stream-ordered: no synchronisation needed
}

// e->perm changed: refresh its inverse (instance id -> slot), used by the interpolation pass
static int refresh_inverse_perm(b200ida_ensemble* e) {
    if (!e->perm) return B200IDA_OK;
    if (!e->inv_perm.p) B200_CUDA_TRY(e->inv_perm.alloc((size_t)e->B * sizeof(int)));
    invert_perm_kernel<<<(unsigned)((e->B + 255) / 256), 256, 0, e->stream>>>(e->perm, e->inv_perm.as<int>(), e->B);
    count_launch();
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}

static int check_mem(int mem) { return (mem == B200IDA_MEM_DEVICE || mem == B200IDA_MEM_HOST) ? 0 : -1; }

// slot-ordered state rows -> caller-ordered output (device or host), converting TS -> TD
template <class TS, class TD>
static int rows_out(b200ida_ensemble* e, TD* dst, const TS* src, int rows, int mem) {
    if (!dst) return B200IDA_OK;
    const int64_t B = e->B;
    const int tb = 256;
    const unsigned grid = (unsigned)((B + tb - 1) / tb);
    if (mem == B200IDA_MEM_DEVICE) {
        unpermute_kernel<TS, TD><<<grid, tb, 0, e->stream>>>(src, dst, rows, B, e->perm);
        count_launch();
        B200_CUDA_TRY(cudaGetLastError());
        return B200IDA_OK;
    }
    const size_t bytes = (size_t)rows * B * sizeof(TD);
    if (!e->perm && sizeof(TS) == sizeof(TD)) {  // identity order, same type: straight copy
        B200_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
        return B200IDA_OK;
    }
    void* tmp;
    e->scratch_reset();
    B200_CUDA_TRY(e->scratch_take(bytes, &tmp));
    unpermute_kernel<TS, TD><<<grid, tb, 0, e->stream>>>(src, static_cast<TD*>(tmp), rows, B, e->perm);
    count_launch();
    B200_CUDA_TRY(cudaGetLastError());
    B200_CUDA_TRY(cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, e->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200IDA_OK;
}


extern "C" {

#ifdef ROOT_PROF
int b200ida_debug_root_prof(unsigned long long* out8, int reset) {
    if (out8) cudaMemcpyFromSymbol(out8, b200ida::root_prof, 8 * sizeof(unsigned long long));
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(b200ida::root_prof, z, sizeof(z));
    }
    return 0;
}
#endif

const char* b200ida_last_error(void) { return last_error_buf(); }
int b200ida_version(void) { return B200IDA_VERSION; }
long long b200ida_kernel_launches(void) { return launch_counter(); }
int b200ida_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return c;
}

int b200ida_reinit(b200ida_ensemble* e, const double* params, const double* yy0, const double* yp0, int mem) {
    if (!e || !yy0 || !yp0 || check_mem(mem)) return set_error(B200IDA_ERR_ARG, "b200ida_reinit: bad argument");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    const Layout lay(e->n);
    const int64_t batch = e->B;
    const int n = e->n;
    const size_t nd = (size_t)lay.ndouble() * batch, ni = (size_t)lay.nint() * batch;
    cudaStream_t st = e->stream;
    const cudaMemcpyKind kind = (mem == B200IDA_MEM_HOST) ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    e->scratch_reset();
    const double *dy0 = yy0, *dyp0 = yp0;
    const size_t vbytes = (size_t)n * batch * sizeof(double);
    B200_CUDA_TRY(cudaMemsetAsync(e->dstate, 0, nd * sizeof(double), st));
    B200_CUDA_TRY(cudaMemsetAsync(e->istate, 0, ni * sizeof(int), st));
    if (e->roots.max_ev > 0) {  // a re-initialised ensemble starts rootfinding afresh (Ida::new, src/lib.rs:360-375)
        B200_CUDA_TRY(cudaMemsetAsync(e->root_d.p, 0, (size_t)(4 * e->nr + 5) * batch * sizeof(double), st));
        B200_CUDA_TRY(cudaMemsetAsync(e->root_i.p, 0, (size_t)(e->nr + 2) * batch * sizeof(int), st));
        B200_CUDA_TRY(cudaMemsetAsync(e->roots.ev_count, 0, (size_t)batch * sizeof(int), st));
    }
    if (params) {
        // new parameters: bring them to the device in caller order, rebuild the slot order, store them by slot
        const size_t pbytes = (size_t)e->nparams * batch * sizeof(double);
        const double* dp = params;
        if (mem == B200IDA_MEM_HOST) {
            void* pb;
            B200_CUDA_TRY(e->scratch_take(pbytes, &pb));
            B200_CUDA_TRY(cudaMemcpyAsync(pb, params, pbytes, kind, st));
            dp = static_cast<double*>(pb);
        }
        int rc = build_order(e, dp);
        if (rc != B200IDA_OK) return rc;
        permute_rows_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, st>>>(dp, e->params, e->nparams, batch, e->perm);
        count_launch();
        B200_CUDA_TRY(cudaGetLastError());
        e->regrouped = false;
        if (e->perm) {
            // keep the parameter-Morton order: a solve may re-group the slots, the next re-initialisation starts from here
            if (!e->perm0.p) B200_CUDA_TRY(e->perm0.alloc((size_t)batch * sizeof(int)));
            if (!e->params0.p) B200_CUDA_TRY(e->params0.alloc(pbytes));
            B200_CUDA_TRY(cudaMemcpyAsync(e->perm0.p, e->perm, (size_t)batch * sizeof(int), cudaMemcpyDeviceToDevice, st));
            B200_CUDA_TRY(cudaMemcpyAsync(e->params0.p, e->params, pbytes, cudaMemcpyDeviceToDevice, st));
            if ((rc = refresh_inverse_perm(e))) return rc;
        }
    } else if (e->regrouped && e->perm && e->perm0.p) {
        B200_CUDA_TRY(cudaMemcpyAsync(e->perm, e->perm0.p, (size_t)batch * sizeof(int), cudaMemcpyDeviceToDevice, st));
        B200_CUDA_TRY(cudaMemcpyAsync(e->params, e->params0.p, (size_t)e->nparams * batch * sizeof(double), cudaMemcpyDeviceToDevice, st));
        e->regrouped = false;
        int rc2 = refresh_inverse_perm(e);
        if (rc2) return rc2;
    }
    if (mem == B200IDA_MEM_HOST) {
        void *y0b, *yp0b;
        B200_CUDA_TRY(e->scratch_take(vbytes, &y0b));
        B200_CUDA_TRY(e->scratch_take(vbytes, &yp0b));
        B200_CUDA_TRY(cudaMemcpyAsync(y0b, yy0, vbytes, kind, st));
        B200_CUDA_TRY(cudaMemcpyAsync(yp0b, yp0, vbytes, kind, st));
        dy0 = static_cast<double*>(y0b);
        dyp0 = static_cast<double*>(yp0b);
    }
    const int tb = 256;
    ida_init_kernel<<<(unsigned)((batch + tb - 1) / tb), tb, 0, st>>>(e->dstate, dy0, dyp0, n, batch, e->perm);
    count_launch();
    B200_CUDA_TRY(cudaGetLastError());
    // host buffers: the caller may reuse them as soon as we return, so wait for the H2D copies; the device scratch
    // itself is persistent and stream-ordered
    if (mem == B200IDA_MEM_HOST) B200_CUDA_TRY(cudaStreamSynchronize(st));
    return B200IDA_OK;
}

int b200ida_create(b200ida_ensemble** out, int problem, int n, int64_t batch, int nparams, const double* params,
                   const double* yy0, const double* yp0, double rtol, const double* atol, int mem, int device,
                   void* stream) {
    if (!out || !params || !yy0 || !yp0 || !atol || batch <= 0 || check_mem(mem))
        return set_error(B200IDA_ERR_ARG, "b200ida_create: null pointer, batch <= 0 or bad mem");
    int need_n = 0, need_p = 0;
    switch (problem) {
    case B200IDA_PROBLEM_ROBERTS: need_n = Roberts::N; need_p = Roberts::NP; break;
    case B200IDA_PROBLEM_LORENZ63: need_n = Lorenz63::N; need_p = Lorenz63::NP; break;
    case B200IDA_PROBLEM_HEAT2D: {
        int M = 1;
        while (M * M < n) ++M;
        if (M * M != n || M < 3 || n > 256)
            return set_error(B200IDA_ERR_ARG, "b200ida_create: heat2d needs n = M*M with 3 <= M <= 16 (got n=%d)", n);
        need_n = n;
        need_p = 1;
        break;
    }
    case B200IDA_PROBLEM_USER: {
        const UserProblemHooks* hk = user_problem_hooks();
        if (!hk)
            return set_error(B200IDA_ERR_UNSUPPORTED, "b200ida_create: this library was built without a user problem "
                                                      "(make -C rust_ida_b200/csrc user PROBLEM_HDR=...)");
        need_n = hk->n;
        need_p = hk->nparams;
        break;
    }
    default: return set_error(B200IDA_ERR_UNSUPPORTED, "b200ida_create: problem %d has no kernel", problem);
    }
    if (n != need_n || nparams != need_p)
        return set_error(B200IDA_ERR_ARG, "b200ida_create: problem %d needs n=%d nparams=%d (got %d, %d)", problem,
                         need_n, need_p, n, nparams);
    if (b200ida_device_count() <= device || device < 0)
        return set_error(B200IDA_ERR_NO_DEVICE, "b200ida_create: CUDA device %d not available", device);
    B200_CUDA_TRY(cudaSetDevice(device));

    b200ida_ensemble* e = new (std::nothrow) b200ida_ensemble();
    if (!e) return set_error(B200IDA_ERR_ARG, "out of host memory");
    e->problem = problem;
    e->nr = problem == B200IDA_PROBLEM_ROBERTS ? Roberts::NR : (problem == B200IDA_PROBLEM_USER ? user_problem_hooks()->nroots : 0);
    e->reorder = g_reorder_mode.load();
    e->n = n;
    e->nparams = nparams;
    e->device = device;
    e->B = batch;
    e->rtol = rtol;
    memset(e->atol, 0, sizeof(e->atol));
    memcpy(e->atol, atol, sizeof(double) * n);
    for (int i = 0; i < MAX_ATOL; ++i) e->idmask[i] = 0.0;  // ida_id is all false in Ida::new (src/lib.rs:321)
    e->stream = static_cast<cudaStream_t>(stream);
    const Layout lay(n);
    const size_t nd = (size_t)lay.ndouble() * batch, ni = (size_t)lay.nint() * batch;
    cudaError_t err;
    if ((err = cudaMalloc(&e->dstate, nd * sizeof(double))) != cudaSuccess ||
        (err = cudaMalloc(&e->istate, ni * sizeof(int))) != cudaSuccess ||
        (err = cudaMalloc(&e->params, (size_t)nparams * batch * sizeof(double))) != cudaSuccess ||
        (err = cudaMalloc(&e->d_atol, (size_t)n * sizeof(double))) != cudaSuccess ||
        (err = cudaMemcpy(e->d_atol, atol, (size_t)n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
        b200ida_destroy(e);
        return set_error(B200IDA_ERR_CUDA, "b200ida_create: cudaMalloc -> %s", cudaGetErrorString(err));
    }
    int rc = b200ida_reinit(e, params, yy0, yp0, mem);
    if (rc != B200IDA_OK) {
        b200ida_destroy(e);
        return rc;
    }
    *out = e;
    return B200IDA_OK;
}

int b200ida_destroy(b200ida_ensemble* e) {
    if (!e) return B200IDA_OK;
    cudaSetDevice(e->device);
    if (e->dstate) cudaFree(e->dstate);
    if (e->istate) cudaFree(e->istate);
    if (e->dstate2) cudaFree(e->dstate2);
    if (e->istate2) cudaFree(e->istate2);
    if (e->params) cudaFree(e->params);
    if (e->d_atol) cudaFree(e->d_atol);
    if (e->d_idmask) cudaFree(e->d_idmask);
    if (e->perm) cudaFree(e->perm);
    e->scratch_free();
    delete e;
    return B200IDA_OK;
}

int b200ida_set_id(b200ida_ensemble* e, const unsigned char* id, int suppressalg) {
    if (!e) return set_error(B200IDA_ERR_ARG, "b200ida_set_id: null handle");
    if (suppressalg && !id) return set_error(B200IDA_ERR_ARG, "b200ida_set_id: suppressalg needs an id vector");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    for (int i = 0; i < e->n; ++i) e->idmask[i] = (id && id[i]) ? 1.0 : 0.0;
    e->suppressalg = suppressalg ? 1 : 0;
    if (!e->d_idmask) B200_CUDA_TRY(cudaMalloc(&e->d_idmask, (size_t)e->n * sizeof(double)));
    B200_CUDA_TRY(cudaMemcpyAsync(e->d_idmask, e->idmask, (size_t)e->n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(e->stream));  // e->idmask may change again before the copy would have run
    return B200IDA_OK;
}

int b200ida_set_trace(b200ida_ensemble* e, int64_t instance, double* dtrace, int32_t* itrace, int32_t* count,
                      int max_attempts) {
    if (!e) return set_error(B200IDA_ERR_ARG, "b200ida_set_trace: null handle");
    if (instance < 0) {
        e->trace = {nullptr, nullptr, nullptr, 0, -1};
        return B200IDA_OK;
    }
    if (instance >= e->B || !dtrace || !itrace || !count || max_attempts < 1)
        return set_error(B200IDA_ERR_ARG, "b200ida_set_trace: bad argument (0 <= instance < batch, non-null device buffers)");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    B200_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), e->stream));
    e->trace = {dtrace, reinterpret_cast<int*>(itrace), reinterpret_cast<int*>(count), max_attempts, (long long)instance};
    return B200IDA_OK;
}

int b200ida_set_roots(b200ida_ensemble* e, int max_events, double* ev_t, double* ev_iroots, int32_t* ev_count, int32_t* nge) {
    if (!e) return set_error(B200IDA_ERR_ARG, "b200ida_set_roots: null handle");
    if (max_events <= 0) {
        e->roots = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
        return B200IDA_OK;
    }
    if (e->nr <= 0)
        return set_error(B200IDA_ERR_UNSUPPORTED, "b200ida_set_roots: this problem has no root functions (Root::num_roots == 0)");
    if (!ev_t || !ev_iroots || !ev_count) return set_error(B200IDA_ERR_ARG, "b200ida_set_roots: null event buffer");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    const int NR = e->nr;
    const size_t db = (size_t)(4 * NR + 5) * e->B * sizeof(double), ib = (size_t)(NR + 2) * e->B * sizeof(int);
    if (!e->root_d.p) B200_CUDA_TRY(e->root_d.alloc(db));  // the batch size of a handle is fixed
    if (!e->root_i.p) B200_CUDA_TRY(e->root_i.alloc(ib));
    B200_CUDA_TRY(cudaMemsetAsync(e->root_d.p, 0, db, e->stream));
    B200_CUDA_TRY(cudaMemsetAsync(e->root_i.p, 0, ib, e->stream));  // gactive = false, irfnd = false, nge = 0 (src/lib.rs:360-375)
    B200_CUDA_TRY(cudaMemsetAsync(ev_count, 0, (size_t)e->B * sizeof(int32_t), e->stream));
    e->roots = {e->root_d.as<double>(), e->root_i.as<int>(), ev_t, ev_iroots, reinterpret_cast<int*>(ev_count),
                reinterpret_cast<int*>(nge), max_events};
    return B200IDA_OK;
}

int b200ida_solve_ladder(b200ida_ensemble* e, const double* touts, int ntout, int itask, double* yy_traj,
                         int32_t* status, int mem) {
    if (!e || !touts || ntout < 1 || ntout > MAX_TOUTS || check_mem(mem))
        return set_error(B200IDA_ERR_ARG, "b200ida_solve_ladder: bad argument (1 <= ntout <= %d)", MAX_TOUTS);
    if (itask != B200IDA_TASK_NORMAL)
        return set_error(B200IDA_ERR_UNSUPPORTED, "b200ida_solve: only IdaTask::Normal is on the device path");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    LadderParams lad;
    lad.ntout = ntout;
    for (int i = 0; i < MAX_TOUTS; ++i) lad.touts[i] = i < ntout ? touts[i] : 0.0;

    double* d_traj = yy_traj;
    int32_t* d_stat = status;
    const size_t traj_bytes = (size_t)ntout * e->n * e->B * sizeof(double);
    const size_t stat_bytes = (size_t)ntout * e->B * sizeof(int32_t);
    if (mem == B200IDA_MEM_HOST) {
        e->scratch_reset();
        void* sp;
        if (yy_traj) {
            B200_CUDA_TRY(e->scratch_take(traj_bytes, &sp));
            d_traj = static_cast<double*>(sp);
        }
        if (status) {
            B200_CUDA_TRY(e->scratch_take(stat_bytes, &sp));
            d_stat = static_cast<int32_t*>(sp);
        }
    }
    int kernel = e->kernel;
    if (kernel == B200IDA_KERNEL_AUTO)
        kernel = (e->n == 3 && e->problem != B200IDA_PROBLEM_HEAT2D) ? B200IDA_KERNEL_THREAD
                 : (e->n <= 32 ? B200IDA_KERNEL_WARP : B200IDA_KERNEL_CTA);
    if (e->roots.max_ev > 0 && kernel != B200IDA_KERNEL_THREAD)
        return set_error(B200IDA_ERR_UNSUPPORTED, "rootfinding is implemented by the thread-per-instance kernel only "
                                                  "(b200ida_set_kernel selected another one): no events would be reported");
    if (e->roots.max_ev > 0 && e->trace.id >= 0)
        return set_error(B200IDA_ERR_UNSUPPORTED, "rootfinding and the per-attempt trace cannot be active in the same launch");
    if (kernel == B200IDA_KERNEL_THREAD) {
        const bool user = e->problem == B200IDA_PROBLEM_USER;
        if (!user && (e->n != 3 || e->problem == B200IDA_PROBLEM_HEAT2D))
            return set_error(B200IDA_ERR_UNSUPPORTED, "thread-per-instance kernels exist for the n = 3 problems and for user problems only");
        ThreadLaunch a;
        a.ds = e->dstate;
        a.is = e->istate;
        a.params = e->params;
        a.B = e->B;
        a.rtol = e->rtol;
        a.atol = e->atol;
        a.idmask = e->idmask;
        a.suppressalg = e->suppressalg;
        a.lad = lad;
        a.traj = d_traj;
        a.stat = d_stat;
        a.perm = e->perm;
        a.trace = e->trace;
        a.roots = e->roots;
        a.stream = e->stream;
        a.df = DeferParams{nullptr, nullptr};
        if (!e->suppressalg && e->roots.max_ev <= 0 && e->trace.id < 0) {
            // production path: snapshot buffers of the deferred output interpolation, grown to the longest ladder seen
            const size_t need_snap = (size_t)e->B * ntout * snap_doubles(e->n) * sizeof(double), need_valid = (size_t)e->B * ntout;
            if (need_snap > e->snap_cap) {
                e->snap.release();
                e->snap_cap = 0;
                B200_CUDA_TRY(e->snap.alloc(need_snap));
                e->snap_cap = need_snap;
            }
            if (need_valid > e->valid_cap) {
                e->snap_valid.release();
                e->valid_cap = 0;
                B200_CUDA_TRY(e->snap_valid.alloc(need_valid));
                e->valid_cap = need_valid;
            }
            a.df = DeferParams{e->snap.as<double>(), e->snap_valid.as<unsigned char>()};
        }
        if (!e->suppressalg && e->trace.id < 0) {  // plain and rootfinding solves: FAST pass + exact pass
            if (IDA_FAST) {
                const Layout lay2(e->n);
                if (!e->dstate2) B200_CUDA_TRY(cudaMalloc(&e->dstate2, (size_t)lay2.ndouble() * e->B * sizeof(double)));
                if (!e->istate2) B200_CUDA_TRY(cudaMalloc(&e->istate2, (size_t)lay2.nint() * e->B * sizeof(int)));
                if (!e->poisoned.p) B200_CUDA_TRY(e->poisoned.alloc((size_t)e->B));
                if (!e->todo.p) B200_CUDA_TRY(e->todo.alloc(((size_t)e->B + 1) * sizeof(int)));
                a.todo = e->todo.as<int>() + 1;
                a.todo_count = e->todo.as<int>();
                a.ds2 = e->dstate2;
                a.is2 = e->istate2;
                a.poisoned = e->poisoned.as<unsigned char>();
            }
        }
        auto launch = [&](const ThreadLaunch& la) -> int {
            if (user) {
                const UserProblemHooks* hk = user_problem_hooks();
                if (!hk) return set_error(B200IDA_ERR_UNSUPPORTED, "no user problem is linked into this library");
                return hk->launch(la);
            }
            if (e->problem == B200IDA_PROBLEM_ROBERTS) return launch_thread_family<Roberts, CAP_TRACE | CAP_ROOTS | CAP_MASK>(la);
            return launch_thread_family<Lorenz63, CAP_TRACE | CAP_MASK>(la);
        };
        a.slot_of_id = e->perm ? e->inv_perm.as<int>() : nullptr;
        // Segmented solve with re-grouping: production path, internally re-ordered ensembles that are
        // large enough for a sort + gather of the state (~1 ms per 2^20 instances) to pay
        const bool segmented = a.ds2 && a.df.snap && e->roots.max_ev <= 0 && e->perm && e->regroup_segments > 1 && e->regroup_trips > 0 &&
                               e->B >= ((int64_t)1 << 16);
        int rc;
        if (!segmented) {
            rc = launch(a);
            if (rc != B200IDA_OK) return rc;
        } else {
            const int64_t B = e->B;
            const int np = e->nparams;
            const size_t XD = (size_t)xrow_doubles(e->n, np), XI = (size_t)xrow_ints(e->n);
            if (!e->xd[0].p) {
                for (int k = 0; k < 2; ++k) {
                    B200_CUDA_TRY(e->xd[k].alloc(XD * B * sizeof(double)));
                    B200_CUDA_TRY(e->xi[k].alloc(XI * B * sizeof(int)));
                }
                B200_CUDA_TRY(e->params2.alloc((size_t)np * B * sizeof(double)));
                B200_CUDA_TRY(e->perm2.alloc((size_t)B * sizeof(int)));
                B200_CUDA_TRY(e->rg_keys.alloc((size_t)B * sizeof(unsigned)));
                B200_CUDA_TRY(e->rg_keys2.alloc((size_t)B * sizeof(unsigned)));
                B200_CUDA_TRY(e->rg_ids.alloc((size_t)B * sizeof(int)));
                B200_CUDA_TRY(e->rg_order.alloc((size_t)B * sizeof(int)));
                size_t tb = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, tb, (unsigned*)nullptr, (unsigned*)nullptr, (int*)nullptr, (int*)nullptr, (int)B, 0, 24, e->stream);
                B200_CUDA_TRY(e->rg_tmp.alloc(tb));
                e->rg_tmp_bytes = tb;
            }
            const int nseg = e->regroup_segments;
            const unsigned g256 = (unsigned)((B + 255) / 256);
            int cur = 0;  // exchange buffer the NEXT segment reads
            for (int seg = 0; seg < nseg; ++seg) {
                const bool last = seg == nseg - 1;
                ThreadLaunch la = a;
                la.run_interp = 0;
                la.seg.max_trips = last ? 0x7fffffff : e->regroup_trips;
                la.seg.resume = seg > 0;
                if (seg > 0) {  // state comes from the rows the previous segment wrote, in the re-grouped order
                    la.io.xd_in = e->xd[cur].as<double>();
                    la.io.xi_in = e->xi[cur].as<int>();
                    la.io.order = e->rg_order.as<int>();
                }
                if (!last) {
                    la.io.xd_out = e->xd[cur ^ 1].as<double>();
                    la.io.xi_out = e->xi[cur ^ 1].as<int>();
                } else {  // back to the structure-of-arrays buffers (dstate2 / istate2), with parameters and ids in the final slot order
                    la.io.params_out = e->params2.as<double>();
                    la.io.perm_out = e->perm2.as<int>();
                }
                rc = launch(la);  // FAST pass + exact pass over its flagged slots; inputs are never written
                if (rc != B200IDA_OK) return rc;
                if (last) break;
                cur ^= 1;
                // re-group: sort the slots by behaviour key; the next segment's lane s loads row order[s]
                regroup_key_kernel<<<g256, 256, 0, e->stream>>>(e->xd[cur].as<double>(), e->xi[cur].as<int>(), e->n, (int)XD, (int)XI, B, ntout,
                                                                e->rg_keys.as<unsigned>(), e->rg_ids.as<int>());
                count_launch();
                size_t tb = e->rg_tmp_bytes;
                B200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(e->rg_tmp.p, tb, e->rg_keys.as<unsigned>(), e->rg_keys2.as<unsigned>(),
                                                              e->rg_ids.as<int>(), e->rg_order.as<int>(), (int)B, 0, 24, e->stream));
                B200_CUDA_TRY(cudaGetLastError());
            }
            { void* t = e->params; e->params = e->params2.as<double>(); e->params2.p = t; }
            { void* t = e->perm; e->perm = e->perm2.as<int>(); e->perm2.p = t; }
            e->regrouped = true;
            if ((rc = refresh_inverse_perm(e))) return rc;
        }
        e->fast_ran = a.ds2 != nullptr;
        if (a.ds2) {  // the FAST / exact passes wrote the new state to the second buffers: they are the state now
            double* td = e->dstate;
            e->dstate = e->dstate2;
            e->dstate2 = td;
            int* ti = e->istate;
            e->istate = e->istate2;
            e->istate2 = ti;
        }
        if (segmented) {  // the interpolation pass over all segments' snapshots, on the final state and slot order
            ThreadLaunch la = a;
            la.run_solve = 0;
            la.run_interp = 1;
            la.ds = e->dstate;
            la.ds2 = nullptr;
            la.slot_of_id = e->inv_perm.as<int>();
            rc = launch(la);
            if (rc != B200IDA_OK) return rc;
        }
    } else {
        const bool warp = kernel == B200IDA_KERNEL_WARP;
        if (warp && e->n > 32) return set_error(B200IDA_ERR_UNSUPPORTED, "warp-per-system needs n <= 32 (n = %d)", e->n);
        const unsigned arena = (unsigned)team_arena_bytes(e->n);
        const int tb = warp ? 128 : (e->n <= 128 ? 128 : 256);
        const int tpc = warp ? tb / 32 : 1;
        const size_t smem = (size_t)arena * tpc;
        if (smem > 227 * 1024) return set_error(B200IDA_ERR_UNSUPPORTED, "n = %d does not fit the shared-memory arena", e->n);
        int sms = 148, occ = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
        if (warp) {
            B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_team_kernel<WarpTeam>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ida_solve_team_kernel<WarpTeam>, tb, smem));
        } else {
            B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_team_kernel<CtaTeam>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ida_solve_team_kernel<CtaTeam>, tb, smem));
        }
        if (occ < 1) occ = 1;
        const int64_t want = (e->B + tpc - 1) / tpc;
        const unsigned grid = (unsigned)(want < (int64_t)sms * occ ? want : (int64_t)sms * occ);  // persistent teams
        if (warp)
            ida_solve_team_kernel<WarpTeam><<<grid, tb, smem, e->stream>>>(e->dstate, e->istate, e->params, e->B, e->n, e->problem,
                                                                           e->nparams, e->rtol, e->d_atol, lad, d_traj, d_stat, arena, e->perm, e->trace,
                                                                           e->suppressalg ? e->d_idmask : nullptr);
        else
            ida_solve_team_kernel<CtaTeam><<<grid, tb, smem, e->stream>>>(e->dstate, e->istate, e->params, e->B, e->n, e->problem,
                                                                          e->nparams, e->rtol, e->d_atol, lad, d_traj, d_stat, arena, e->perm, e->trace,
                                                                          e->suppressalg ? e->d_idmask : nullptr);
        count_launch();
    }
    B200_CUDA_TRY(cudaGetLastError());
    if (mem == B200IDA_MEM_HOST) {
        if (yy_traj) B200_CUDA_TRY(cudaMemcpyAsync(yy_traj, d_traj, traj_bytes, cudaMemcpyDeviceToHost, e->stream));
        if (status) B200_CUDA_TRY(cudaMemcpyAsync(status, d_stat, stat_bytes, cudaMemcpyDeviceToHost, e->stream));
        B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return B200IDA_OK;
}

int b200ida_get_fallback_count(b200ida_ensemble* e, int64_t* count, int64_t* hist16) {
    if (!e || !count) return set_error(B200IDA_ERR_ARG, "b200ida_get_fallback_count: bad argument");
    *count = 0;
    if (hist16)
        for (int i = 0; i < 16; ++i) hist16[i] = 0;
    if (!e->poisoned.p || !e->fast_ran) return B200IDA_OK;  // the last solve did not take the FAST path
    B200_CUDA_TRY(cudaSetDevice(e->device));
    unsigned char* h = static_cast<unsigned char*>(malloc((size_t)e->B));
    if (!h) return set_error(B200IDA_ERR_ARG, "out of host memory");
    cudaError_t err = cudaMemcpyAsync(h, e->poisoned.p, (size_t)e->B, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    int64_t c = 0;
    if (err == cudaSuccess)
        for (int64_t i = 0; i < e->B; ++i) {
            c += h[i] != 0;
            if (hist16) hist16[h[i] & 15] += 1;
        }
    free(h);
    B200_CUDA_TRY(err);
    *count = c;
    return B200IDA_OK;
}

int b200ida_set_regroup(b200ida_ensemble* e, int trips_per_segment, int max_segments) {
    if (!e || trips_per_segment < 0 || max_segments < 0 || max_segments > 64)
        return set_error(B200IDA_ERR_ARG, "b200ida_set_regroup: bad argument");
    e->regroup_trips = trips_per_segment;
    e->regroup_segments = max_segments;
    return B200IDA_OK;
}

int b200ida_set_kernel(b200ida_ensemble* e, int kernel) {
    if (!e || kernel < B200IDA_KERNEL_AUTO || kernel > B200IDA_KERNEL_CTA)
        return set_error(B200IDA_ERR_ARG, "b200ida_set_kernel: bad argument");
    e->kernel = kernel;
    return B200IDA_OK;
}

int b200ida_solve(b200ida_ensemble* e, double tout, int itask, int32_t* status, int mem) {
    return b200ida_solve_ladder(e, &tout, 1, itask, nullptr, status, mem);
}

int b200ida_get_state(b200ida_ensemble* e, double* yy, double* yp, double* tret, int mem) {
    if (!e || check_mem(mem)) return set_error(B200IDA_ERR_ARG, "b200ida_get_state: bad argument");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    const Layout lay(e->n);
    int rc;
    if ((rc = rows_out<double, double>(e, yy, e->dstate + (size_t)lay.yy(0) * e->B, e->n, mem))) return rc;
    if ((rc = rows_out<double, double>(e, yp, e->dstate + (size_t)lay.yp(0) * e->B, e->n, mem))) return rc;
    if ((rc = rows_out<double, double>(e, tret, e->dstate + (size_t)lay.scal(Layout::TRET) * e->B, 1, mem))) return rc;
    return B200IDA_OK;
}

int b200ida_get_counters(b200ida_ensemble* e, int64_t* nst, int64_t* nre, int64_t* nje, int64_t* nni, int64_t* netf,
                         int64_t* ncfn, int64_t* nsetups, int64_t* nattempts, int mem) {
    if (!e || check_mem(mem)) return set_error(B200IDA_ERR_ARG, "b200ida_get_counters: bad argument");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    const size_t B = (size_t)e->B;
    int rc;
    if ((rc = rows_out<int, int64_t>(e, nst, e->istate + (size_t)Layout::NST * B, 1, mem))) return rc;
    if ((rc = rows_out<int, int64_t>(e, nre, e->istate + (size_t)Layout::NRE * B, 1, mem))) return rc;
    if ((rc = rows_out<int, int64_t>(e, nje, e->istate + (size_t)Layout::NJE * B, 1, mem))) return rc;
    if ((rc = rows_out<int, int64_t>(e, nni, e->istate + (size_t)Layout::NNI * B, 1, mem))) return rc;
    if ((rc = rows_out<int, int64_t>(e, netf, e->istate + (size_t)Layout::NETF * B, 1, mem))) return rc;
    if ((rc = rows_out<int, int64_t>(e, ncfn, e->istate + (size_t)Layout::NCFN * B, 1, mem))) return rc;
    if ((rc = rows_out<int, int64_t>(e, nsetups, e->istate + (size_t)Layout::NSETUPS * B, 1, mem))) return rc;
    if ((rc = rows_out<int, int64_t>(e, nattempts, e->istate + (size_t)Layout::NATT * B, 1, mem))) return rc;
    return B200IDA_OK;
}

int b200ida_get_step_info(b200ida_ensemble* e, int32_t* kused, int32_t* kcur, double* hused, double* hcur,
                          double* tcur, double* h0u, int mem) {
    if (!e || check_mem(mem)) return set_error(B200IDA_ERR_ARG, "b200ida_get_step_info: bad argument");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    const Layout lay(e->n);
    const size_t B = (size_t)e->B;
    int rc;
    if ((rc = rows_out<int, int32_t>(e, kused, e->istate + (size_t)Layout::KUSED * B, 1, mem))) return rc;
    if ((rc = rows_out<int, int32_t>(e, kcur, e->istate + (size_t)Layout::KK * B, 1, mem))) return rc;
    if ((rc = rows_out<double, double>(e, hused, e->dstate + (size_t)lay.scal(Layout::HUSED) * B, 1, mem))) return rc;
    if ((rc = rows_out<double, double>(e, hcur, e->dstate + (size_t)lay.scal(Layout::HH) * B, 1, mem))) return rc;
    if ((rc = rows_out<double, double>(e, tcur, e->dstate + (size_t)lay.scal(Layout::TN) * B, 1, mem))) return rc;
    if ((rc = rows_out<double, double>(e, h0u, e->dstate + (size_t)lay.scal(Layout::H0U) * B, 1, mem))) return rc;
    return B200IDA_OK;
}

int b200ida_get_dky(b200ida_ensemble* e, double t, int k, double* dky, int32_t* status, int mem) {
    if (!e || !dky || check_mem(mem)) return set_error(B200IDA_ERR_ARG, "b200ida_get_dky: bad argument");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    const int64_t B = e->B;
    const unsigned grid = (unsigned)((B + 255) / 256);
    if (mem == B200IDA_MEM_DEVICE) {
        ida_dky_kernel<<<grid, 256, 0, e->stream>>>(e->dstate, e->istate, e->n, B, t, k, dky, status, e->perm);
        count_launch();
        B200_CUDA_TRY(cudaGetLastError());
        return B200IDA_OK;
    }
    // host buffers: compute into scratch (pre-filled from the caller's dky so that failed instances stay untouched)
    const size_t dbytes = (size_t)e->n * B * sizeof(double), sbytes = (size_t)B * sizeof(int32_t);
    void *d_dky, *d_st;
    e->scratch_reset();
    B200_CUDA_TRY(e->scratch_take(dbytes, &d_dky));
    B200_CUDA_TRY(e->scratch_take(sbytes, &d_st));
    B200_CUDA_TRY(cudaMemcpyAsync(d_dky, dky, dbytes, cudaMemcpyHostToDevice, e->stream));
    ida_dky_kernel<<<grid, 256, 0, e->stream>>>(e->dstate, e->istate, e->n, B, t, k, static_cast<double*>(d_dky),
                                                static_cast<int*>(d_st), e->perm);
    count_launch();
    B200_CUDA_TRY(cudaGetLastError());
    B200_CUDA_TRY(cudaMemcpyAsync(dky, d_dky, dbytes, cudaMemcpyDeviceToHost, e->stream));
    if (status) B200_CUDA_TRY(cudaMemcpyAsync(status, d_st, sbytes, cudaMemcpyDeviceToHost, e->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200IDA_OK;
}

int b200ida_set_default_reorder(int mode) {
    if (mode < B200IDA_REORDER_OFF || mode > B200IDA_REORDER_AUTO) return set_error(B200IDA_ERR_ARG, "b200ida_set_default_reorder: bad mode");
    g_reorder_mode.store(mode);
    return B200IDA_OK;
}

int b200ida_sync(b200ida_ensemble* e) {
    if (!e) return set_error(B200IDA_ERR_ARG, "b200ida_sync: null handle");
    B200_CUDA_TRY(cudaSetDevice(e->device));
    B200_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return B200IDA_OK;
}

int b200ida_detpow_batched(const double* x, const double* y, double* out, int64_t n, int mem, void* stream) {
    if (!x || !y || !out || n < 1 || check_mem(mem)) return set_error(B200IDA_ERR_ARG, "b200ida_detpow_batched: bad argument");
    if (b200ida_device_count() < 1) return set_error(B200IDA_ERR_NO_DEVICE, "b200ida_detpow_batched: no CUDA device");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (mem == B200IDA_MEM_DEVICE) {
        detpow_kernel<<<grid, 256, 0, st>>>(x, y, out, n);
        B200_CUDA_TRY(cudaGetLastError());
        return B200IDA_OK;
    }
    DevBuf dx, dy, dout;
    const size_t bytes = (size_t)n * sizeof(double);
    B200_CUDA_TRY(dx.alloc(bytes));
    B200_CUDA_TRY(dy.alloc(bytes));
    B200_CUDA_TRY(dout.alloc(bytes));
    B200_CUDA_TRY(cudaMemcpyAsync(dx.p, x, bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(dy.p, y, bytes, cudaMemcpyHostToDevice, st));
    detpow_kernel<<<grid, 256, 0, st>>>(dx.as<double>(), dy.as<double>(), dout.as<double>(), n);
    B200_CUDA_TRY(cudaGetLastError());
    B200_CUDA_TRY(cudaMemcpyAsync(out, dout.p, bytes, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    return B200IDA_OK;
}

int b200ida_fastdiv_batched(const double* a, const double* b, double* q, double* r, int64_t n) {
    if (!a || !b || !q || !r || n < 1) return set_error(B200IDA_ERR_ARG, "b200ida_fastdiv_batched: bad argument");
    if (b200ida_device_count() < 1) return set_error(B200IDA_ERR_NO_DEVICE, "b200ida_fastdiv_batched: no CUDA device");
    DevBuf da, db, dq, dr;
    const size_t bytes = (size_t)n * sizeof(double);
    B200_CUDA_TRY(da.alloc(bytes));
    B200_CUDA_TRY(db.alloc(bytes));
    B200_CUDA_TRY(dq.alloc(bytes));
    B200_CUDA_TRY(dr.alloc(bytes));
    B200_CUDA_TRY(cudaMemcpy(da.p, a, bytes, cudaMemcpyHostToDevice));
    B200_CUDA_TRY(cudaMemcpy(db.p, b, bytes, cudaMemcpyHostToDevice));
    const int64_t threads = (n + 2) / 3;
    fastdiv_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(da.as<double>(), db.as<double>(), dq.as<double>(), dr.as<double>(), n);
    B200_CUDA_TRY(cudaGetLastError());
    B200_CUDA_TRY(cudaMemcpy(q, dq.p, bytes, cudaMemcpyDeviceToHost));
    B200_CUDA_TRY(cudaMemcpy(r, dr.p, bytes, cudaMemcpyDeviceToHost));
    return B200IDA_OK;
}

int b200ida_fp64_issue_rate(int kind, int device, double* ginstr_per_s) {
    if (!ginstr_per_s || kind < 0 || kind > 2) return set_error(B200IDA_ERR_ARG, "b200ida_fp64_issue_rate: bad argument");
    if (b200ida_device_count() <= device) return set_error(B200IDA_ERR_NO_DEVICE, "no CUDA device %d", device);
    B200_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, tb = 256, iters = 20000;
    DevBuf out;
    B200_CUDA_TRY(out.alloc(8));
    cudaEvent_t e0, e1;
    B200_CUDA_TRY(cudaEventCreate(&e0));
    B200_CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        B200_CUDA_TRY(cudaEventRecord(e0));
        if (kind == 0) fp64_rate_kernel<0><<<blocks, tb>>>(out.as<double>(), iters, 1.0);
        else if (kind == 1) fp64_rate_kernel<1><<<blocks, tb>>>(out.as<double>(), iters, 1.0);
        else fp64_rate_kernel<2><<<blocks, tb>>>(out.as<double>(), iters, 1.0);
        B200_CUDA_TRY(cudaEventRecord(e1));
        B200_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        B200_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double ops = (double)blocks * tb * (double)iters * 8.0;
    *ginstr_per_s = ops / ((double)best * 1e-3) * 1e-9;
    return B200IDA_OK;
}

}  // extern "C"
