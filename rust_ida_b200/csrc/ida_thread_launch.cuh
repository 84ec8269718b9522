// ida_thread_launch.cuh -- the thread-per-instance kernels and the one place that launches them, for ANY problem functor.
//
// Included by ida_engine.cu (built-in problems: Roberts, Lorenz63) and by user_problem.cu (a user-supplied functor,
// `make user PROBLEM_HDR=...`, INTEGRATION.md "Adding a problem"): both get the same kernels, instantiated for their
// problem type.  A problem functor P provides (see problems.cuh):
//     static constexpr int N, NP, NR;  static constexpr bool ROOT_USES_YP;
//     void load(const double* params, int64_t B, int64_t b);            // per-instance parameters, SoA [NP][B]
//     void res(double t, const double (&yy)[N], const double (&yp)[N], double (&r)[N]) const;         // Residual::res
//     void jac(double t, double cj, const double (&yy)[N], const double (&yp)[N], const double (&r)[N], double (&J)[N][N]) const;
//     void root(double t, const double (&y)[N], const double (&yp)[N], double (&g)[NR]) const;        // Root::root (NR > 0)
// i.e. the reference's Residual / Jacobian / Root traits (src/traits.rs:28-89) as device code.
#pragma once
#include "host_util.h"
#include "ida_thread.cuh"

namespace b200ida {

// Thread-per-instance solve over a ladder of output times (ida_thread.cuh).
#ifndef IDA_MIN_BLOCKS
#define IDA_MIN_BLOCKS 6  // x 64 threads: 168 regs, 12 warps/SM; 128 regs (16 warps) is slower: the spills thrash the small L1
#endif
#ifndef IDA_FAST
#define IDA_FAST 1  // production path = FAST pass + exact pass over the flagged instances; 0: the exact kernel alone
#endif
#ifndef IDA_DEFER
#define IDA_DEFER 1  // deferred output interpolation (DeferParams, ida_common.cuh); 0: interpolate in the solve kernel
#endif
// FAST (production): fast paths only, poisoned lanes flagged in xp.poisoned (IdaThread, ida_thread.cuh).  The state is
// read from (ds, is) and written to (xp.ds_out, xp.is_out) -- a different buffer, so that the exact instantiation can
// re-integrate the flagged instances from the untouched input: launched right after with xp.select = the flags.
struct FastParams {
    double* ds_out = nullptr;        // null: in place
    int* is_out = nullptr;
    unsigned char* poisoned = nullptr;  // [batch] by slot, written by the FAST kernel (diagnostics)
    int* list = nullptr;                // FAST kernel: appends the slots it hands to the exact pass ...
    int* count = nullptr;               // ... counting them here (zeroed before the launch)
    const int* todo = nullptr;          // exact kernel: thread t integrates slot todo[t], t < *todo_count (null: slot = t)
    const int* todo_count = nullptr;
};
template <class P, bool TRACE = false, bool ROOTS = false, bool MASK = false, bool FAST = false>
__global__ void __launch_bounds__(THREAD_BLOCK, IDA_MIN_BLOCKS)
ida_solve_thread_kernel(double* __restrict__ ds, int* __restrict__ is, const double* __restrict__ params, int64_t B,
                        TolN<P::N> tol, LadderParams lad, double* __restrict__ yy_traj, int32_t* __restrict__ status,
                        const int* __restrict__ perm, TraceParams tr = TraceParams(), RootParams rp = RootParams(),
                        DeferParams df = DeferParams(), FastParams xp = FastParams(), SegParams sg = SegParams(),
                        IoParams io = IoParams()) {
    extern __shared__ double smem[];
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // slot
    if (b >= B) return;
    if (xp.todo) {  // exact pass over the compacted list of flagged slots: full warps however the flags are scattered
        if (b >= (int64_t)*xp.todo_count) return;
        b = xp.todo[b];
    }
    constexpr Layout xlay(P::N);
    constexpr int ND = xlay.ndouble(), NI = xlay.nint(), XD = xrow_doubles(P::N, P::NP), XI = xrow_ints(P::N);
    // where this lane's state comes from: its SoA column, or its row of the exchange buffer (IoParams)
    const bool row_in = io.xd_in != nullptr;
    const int64_t r_in = row_in ? (io.order ? (int64_t)io.order[b] : b) : 0;
    const double* din = row_in ? io.xd_in + r_in * XD : ds;
    const int* iin = row_in ? io.xi_in + r_in * XI : is;
    const int64_t Bin = row_in ? 1 : B, bin = row_in ? 0 : b;
    const int64_t ob = row_in ? (int64_t)iin[NI + 2] : (perm ? (int64_t)perm[b] : b);  // caller-visible instance id
    IdaThread<P, MASK, FAST> s;
    s.sm = smem + threadIdx.x;
    if constexpr (MASK) {
#pragma unroll
        for (int i = 0; i < P::N; ++i) s.idm[i] = tol.idmask[i];
    }
    double pr[P::NP];  // the instance's parameters travel with it
#pragma unroll
    for (int k = 0; k < P::NP; ++k) pr[k] = row_in ? din[ND + 1 + k] : params[(int64_t)k * B + b];
    s.prob.load(pr, 1, 0);
    s.load(din, iin, Bin, bin);
    if (sg.max_trips >= 0) {  // loop state of a segmented solve lives in the exchange rows
        sg.li = iin + NI;
        sg.ld = din + ND;
        sg.stride = 1;
        if (io.xd_out) {
            sg.li_out = io.xi_out + b * XI + NI;
            sg.ld_out = io.xd_out + b * XD + ND;
        } else {  // last segment: nothing resumes after it; park into the (discarded) tail of this lane's own scratch
            sg.li_out = nullptr;
            sg.ld_out = nullptr;
        }
    }
    if constexpr (ROOTS) {
        s.load_roots(rp, B, b);
        s.load_event_count(rp, ob);
    }
    // the production instantiation defers the output interpolation to ida_interp_kernel; the TRACE and ROOTS
    // instantiations read yy / yp right after get_solution (trace frames, g(t, y)) and keep it in line
    s.template solve_ladder<TRACE, ROOTS, !TRACE && !ROOTS && !MASK && IDA_DEFER>(lad, tol, yy_traj, status, B, ob, tr, rp, df, 0, sg);
    if constexpr (FAST) {
#if IDA_POISON_REASON
        xp.poisoned[b] = (unsigned char)__ffs((int)s.poison);  // 1 + index of the first failed test (ida_thread.cuh)
#else
        xp.poisoned[b] = s.poison ? 1 : 0;
#endif
        if (s.poison) {  // its outputs are rewritten by the exact pass
            xp.list[atomicAdd(xp.count, 1)] = (int)b;
            return;
        }
    }
    if (io.xd_out) {  // a segment that is not the last: the whole instance goes to row `slot` of the exchange buffer
        double* dout = io.xd_out + b * XD;
        int* iout = io.xi_out + b * XI;
        s.store(dout, iout, 1, 0);
#pragma unroll
        for (int k = 0; k < P::NP; ++k) dout[ND + 1 + k] = pr[k];
        iout[NI + 2] = (int)ob;
    } else {
        s.store(xp.ds_out ? xp.ds_out : ds, xp.is_out ? xp.is_out : is, B, b);
        if (io.params_out) {
#pragma unroll
            for (int k = 0; k < P::NP; ++k) io.params_out[(int64_t)k * B + b] = pr[k];
            io.perm_out[b] = (int)ob;
        }
    }
    if constexpr (ROOTS) {
        s.store_roots(rp, B, b);
        s.store_event_count(rp, ob);
        if (rp.nge_out) rp.nge_out[ob] = s.nge;
    }
}

// Second pass of the deferred output interpolation (DeferParams, ida_common.cuh): get_solution (src/lib.rs:1274-1343)
// for every (instance, output time) whose call returned through stop_test2, from the snapshot the solve kernel took.
// One thread per (slot, output time); the snapshot rows of consecutive threads are consecutive in HBM, each thread
// pulls its 224-byte row with 16-byte loads issued back to back and interpolates from registers -- an HBM-streaming
// kernel (algorithmic bytes: one row read, n doubles written).  y(tout) goes to yy_traj (caller order) when one was
// asked for; y, y' of the LAST call go to the persistent state, where the reference's solve() leaves them.  The
// arithmetic is get_solution_fn's (ida_thread.cuh) operation for operation, with the order loop unrolled and guarded.
template <int N>
__global__ void __launch_bounds__(256)
ida_interp_kernel(double* __restrict__ ds, int64_t B, LadderParams lad, DeferParams df, double* __restrict__ yy_traj,
                  const int* __restrict__ slot_of_id) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // instance id * ntout + it
    if (g >= B * lad.ntout) return;
    if (!df.valid[g]) return;
    const int64_t ob = g / lad.ntout;  // caller-visible instance id
    const int it = (int)(g - ob * lad.ntout);
    constexpr Layout lay(N);
    constexpr int SD = snap_doubles(N);
    static_assert(SD % 2 == 0, "rows are read as 16-byte pairs");
    double v[SD];
    const double2* row = reinterpret_cast<const double2*>(df.snap + (size_t)g * SD);
#pragma unroll
    for (int k = 0; k < SD / 2; ++k) {
        const double2 t = row[k];
        v[2 * k] = t.x;
        v[2 * k + 1] = t.y;
    }
    const int kused = (int)v[6 * N + 6];
    const double tn = v[6 * N + 7], hh = v[6 * N + 8], hused = v[6 * N + 9], t = lad.touts[it];
    const double sgn = (hh > 0.0 || (hh == 0.0 && !signbit(hh))) ? 1.0 : -1.0;
    const double tfuzz = 100.0 * DBL_EPS * (fabs(tn) + fabs(hh)) * sgn;
    const double tp = tn - hused - tfuzz;
    if ((t - tp) * hh < 0.0) return;  // unreachable after stop_test2 (tout lies inside the last step)
    const int kord = (kused == 0) ? 1 : kused;
    const double delt = t - tn;
    double c = 1.0, d = 0.0, num = delt;
    double y[N], yp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        y[i] = 0.0 + c * v[i];
        yp[i] = 0.0;
    }
#pragma unroll
    for (int j = 1; j <= MAXORD; ++j) {
        if (j <= kord) {
            const double psi_jm1 = v[6 * N + j - 1];
            const double gam = num / psi_jm1;
            d = d * gam + c / psi_jm1;
            c = c * gam;
            num = delt + psi_jm1;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double p = v[j * N + i];
                y[i] += c * p;
                yp[i] += d * p;
            }
        }
    }
    if (yy_traj) {
#pragma unroll
        for (int i = 0; i < N; ++i) yy_traj[((int64_t)it * N + i) * B + ob] = y[i];
    }
    if (it == lad.ntout - 1) {
        const int64_t b = slot_of_id ? (int64_t)slot_of_id[ob] : ob;  // where this instance's state lives now
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ds[(int64_t)lay.yy(i) * B + b] = y[i];
            ds[(int64_t)lay.yp(i) * B + b] = yp[i];
        }
    }
}


// everything one launch of the thread-per-instance family needs (filled by b200ida_solve_ladder)
struct ThreadLaunch {
    double* ds;
    int* is;
    const double* params;
    int64_t B;
    double rtol;
    const double* atol;    // host, [n]
    const double* idmask;  // host, [n]: ida_id as 1.0 / 0.0
    int suppressalg;
    LadderParams lad;
    double* traj;
    int32_t* stat;
    const int* perm;
    TraceParams trace;  // id < 0: off
    RootParams roots;   // max_ev == 0: off
    DeferParams df;     // snapshot buffers sized for snap_doubles(n) (production instantiation)
    double* ds2 = nullptr;  // production path: the output state buffers of the FAST pass (the caller swaps them in
    int* is2 = nullptr;     // afterwards) and the per-slot poison flags
    unsigned char* poisoned = nullptr;
    int* todo = nullptr;        // [batch] compacted slots of the flagged instances
    int* todo_count = nullptr;  // [1]
    SegParams seg;              // segmented launch (max_trips < 0: the whole ladder in one launch)
    IoParams io;                // where the state comes from / goes to (exchange rows between segments)
    int run_solve = 1;          // 0: only the interpolation pass (after the last segment of a segmented solve)
    int run_interp = 1;         // 0: this is not the last segment, the interpolation pass comes later
    const int* slot_of_id = nullptr;  // inverse of perm for the interpolation pass's final-state write (null: identity)
    cudaStream_t stream;
};

// which optional instantiations a problem is compiled with (each one is a full copy of the solver: ~1 minute of nvcc)
enum { CAP_TRACE = 1, CAP_ROOTS = 2, CAP_MASK = 4 };

template <class P, int CAPS>
int launch_thread_family(const ThreadLaunch& a) {
    constexpr int N = P::N;
    const int tb = THREAD_BLOCK;
    const unsigned grid = (unsigned)((a.B + tb - 1) / tb);
    const size_t smem = (size_t)IdaThread<P>::SMEM_DOUBLES * tb * sizeof(double);
    if (smem > 75 * 1024) return set_error(B200IDA_ERR_UNSUPPORTED, "thread-per-instance: n = %d needs %zu bytes of shared memory per CTA", N, smem);
    TolN<N> tol;
    tol.rtol = a.rtol;
    for (int i = 0; i < N; ++i) tol.atol[i] = a.atol[i];
    for (int i = 0; i < N; ++i) tol.idmask[i] = a.idmask ? a.idmask[i] : 0.0;
    const int ntout = a.lad.ntout;
    if (a.suppressalg && (a.roots.max_ev > 0 || a.trace.id >= 0))
        return set_error(B200IDA_ERR_UNSUPPORTED, "suppressalg cannot be combined with rootfinding or the attempt trace");
    if (a.suppressalg) {
        // masked local-error norms: the MASK instantiation (interpolates in line; it is not the benchmarked path)
        if constexpr ((CAPS & CAP_MASK) != 0) {
            B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_thread_kernel<P, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ida_solve_thread_kernel<P, false, false, true><<<grid, tb, smem, a.stream>>>(a.ds, a.is, a.params, a.B, tol, a.lad, a.traj, a.stat, a.perm);
            count_launch();
        } else {
            return set_error(B200IDA_ERR_UNSUPPORTED, "this problem was built without the suppressalg instantiation");
        }
    } else if (a.roots.max_ev > 0) {
        if constexpr ((CAPS & CAP_ROOTS) != 0 && P::NR > 0) {
            const size_t rsmem = (size_t)(IdaThread<P>::SMEM_DOUBLES + IdaThread<P>::ROOT_DOUBLES) * tb * sizeof(double);
            if (IDA_FAST && a.ds2 && a.is2 && a.poisoned) {
                // FAST pass (root checks included; the rare Illinois iteration and the interpolation are exact out-of-line
                // functions either way) + exact pass over the flagged instances; the root state is written at the very
                // end of a lane's launch only, so an abandoned lane leaves it untouched
                FastParams xf;
                xf.ds_out = a.ds2;
                xf.is_out = a.is2;
                xf.poisoned = a.poisoned;
                xf.list = a.todo;
                xf.count = a.todo_count;
                B200_CUDA_TRY(cudaMemsetAsync(a.todo_count, 0, sizeof(int), a.stream));
                B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_thread_kernel<P, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
                ida_solve_thread_kernel<P, false, true, false, true><<<grid, tb, rsmem, a.stream>>>(a.ds, a.is, a.params, a.B, tol, a.lad, a.traj, a.stat,
                                                                                                   a.perm, TraceParams(), a.roots, DeferParams(), xf);
                count_launch();
                B200_CUDA_TRY(cudaGetLastError());
                FastParams xe;
                xe.ds_out = a.ds2;
                xe.is_out = a.is2;
                xe.todo = a.todo;
                xe.todo_count = a.todo_count;
                B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_thread_kernel<P, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
                ida_solve_thread_kernel<P, false, true><<<grid, tb, rsmem, a.stream>>>(a.ds, a.is, a.params, a.B, tol, a.lad, a.traj, a.stat, a.perm,
                                                                                      TraceParams(), a.roots, DeferParams(), xe);
                count_launch();
            } else {
                B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_thread_kernel<P, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
                ida_solve_thread_kernel<P, false, true><<<grid, tb, rsmem, a.stream>>>(a.ds, a.is, a.params, a.B, tol, a.lad, a.traj, a.stat, a.perm,
                                                                                      TraceParams(), a.roots);
                count_launch();
            }
        } else {
            return set_error(B200IDA_ERR_UNSUPPORTED, "rootfinding: this problem has no root functions (Root::num_roots == 0)");
        }
    } else if (a.trace.id >= 0) {
        // traced run: the TRACE instantiation of the same kernel (the production instantiation carries no trace code)
        if constexpr ((CAPS & CAP_TRACE) != 0) {
            B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_thread_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ida_solve_thread_kernel<P, true><<<grid, tb, smem, a.stream>>>(a.ds, a.is, a.params, a.B, tol, a.lad, a.traj, a.stat, a.perm, a.trace);
            count_launch();
        } else {
            return set_error(B200IDA_ERR_UNSUPPORTED, "this problem was built without the trace instantiation");
        }
    } else {
        // production path: FAST pass (state a.ds -> a.ds2, snapshots at the output times), the exact pass over the
        // instances the FAST pass flagged (normally none: its CTAs exit at once), the coherent interpolation pass
        if (!a.run_solve) {
        } else if (IDA_FAST && a.ds2 && a.is2 && a.poisoned) {
            FastParams xf;
            xf.ds_out = a.ds2;
            xf.is_out = a.is2;
            xf.poisoned = a.poisoned;
            xf.list = a.todo;
            xf.count = a.todo_count;
            B200_CUDA_TRY(cudaMemsetAsync(a.todo_count, 0, sizeof(int), a.stream));
            B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_thread_kernel<P, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ida_solve_thread_kernel<P, false, false, false, true><<<grid, tb, smem, a.stream>>>(a.ds, a.is, a.params, a.B, tol, a.lad, a.traj, a.stat,
                                                                                               a.perm, TraceParams(), RootParams(), a.df, xf, a.seg, a.io);
            count_launch();
            B200_CUDA_TRY(cudaGetLastError());
            FastParams xe;
            xe.ds_out = a.ds2;
            xe.is_out = a.is2;
            xe.todo = a.todo;
            xe.todo_count = a.todo_count;
            B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_thread_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ida_solve_thread_kernel<P><<<grid, tb, smem, a.stream>>>(a.ds, a.is, a.params, a.B, tol, a.lad, a.traj, a.stat, a.perm, TraceParams(),
                                                                     RootParams(), a.df, xe, a.seg, a.io);
            count_launch();
        } else {
            B200_CUDA_TRY(cudaFuncSetAttribute(ida_solve_thread_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ida_solve_thread_kernel<P><<<grid, tb, smem, a.stream>>>(a.ds, a.is, a.params, a.B, tol, a.lad, a.traj, a.stat, a.perm, TraceParams(),
                                                                     RootParams(), a.df);
            count_launch();
        }
        B200_CUDA_TRY(cudaGetLastError());
        if (IDA_DEFER && a.run_interp) {
            ida_interp_kernel<N><<<(unsigned)((a.B * ntout + 255) / 256), 256, 0, a.stream>>>((IDA_FAST && a.ds2 && a.is2 && a.poisoned) ? a.ds2 : a.ds,
                                                                                              a.B, a.lad, a.df, a.traj, a.slot_of_id);
            count_launch();
        }
    }
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}

// ---- user-supplied problems (user_problem.cu registers itself through this at load time) -----------------------------
struct UserProblemHooks {
    int n, nparams, nroots;
    int (*launch)(const ThreadLaunch&);
};
void register_user_problem(const UserProblemHooks* h);  // ida_engine.cu
const UserProblemHooks* user_problem_hooks();

}  // namespace b200ida
