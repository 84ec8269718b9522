/* b200ida.h -- C ABI of the B200-native batched IDA engine (libb200ida.so).
 *
 * Drop-in boundary for ONE hot path of jondo2010/rust-ida: the per-step BDF/Newton inner loop run over
 * large ensembles of independent DAE instances.  Every entry point names the reference interface it
 * replaces (file:line under the reference tree).  Plain pointers and sizes only; no C++/torch types.
 * There is no CPU fallback and no backend dispatch: every compute entry point launches hand-written
 * sm_100a CUDA kernels and fails with B200IDA_ERR_CUDA / B200IDA_ERR_NO_DEVICE when it cannot.
 *
 * Conventions
 *   - "SoA" arrays are component-major: element (i, b) of an [n][batch] array is at  i*batch + b ,
 *     so that consecutive instances are consecutive in memory (coalesced, vectorisable FP64 access).
 *   - "AoS" matrices are one column-major n x n matrix per instance (nalgebra's owned storage, the
 *     layout `LSolver::setup` receives, crates/linear/src/traits.rs:52-54): element (i, j) of
 *     instance b is at  b*n*n + j*n + i ; AoS vectors are  b*n + i .
 *   - `mem` says where caller buffers live: B200IDA_MEM_DEVICE (device pointers, no copies, work is
 *     enqueued on `stream` and the call returns without synchronising) or B200IDA_MEM_HOST (host
 *     pointers; the call copies in, runs, copies out and synchronises -- the end-to-end path).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - Return value: 0 on success, negative B200IDA_ERR_* otherwise.  Per-instance outcomes are
 *     reported in status / info arrays, never through the return value.
 *   - Arithmetic policy: IEEE-754 binary64, round-to-nearest, NO fused multiply-add contraction
 *     (kernels are built with -fmad=false) so results are bit-identical to the reference's
 *     un-fused Rust arithmetic; pow is glibc-2.39-compatible (csrc/detpow.h).
 */
#ifndef B200IDA_H
#define B200IDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200IDA_VERSION 100

/* ---- return codes -------------------------------------------------------------------------------- */
#define B200IDA_OK 0
#define B200IDA_ERR_ARG (-1)       /* bad argument (null pointer, n out of range, unknown enum ...) */
#define B200IDA_ERR_CUDA (-2)      /* a CUDA runtime call or kernel launch failed; see b200ida_last_error */
#define B200IDA_ERR_NO_DEVICE (-3) /* no sm_100-class CUDA device is available */
#define B200IDA_ERR_UNSUPPORTED (-4) /* (problem, n) combination has no kernel */

/* ---- memory space / layout ------------------------------------------------------------------------ */
#define B200IDA_MEM_DEVICE 0
#define B200IDA_MEM_HOST 1
#define B200IDA_LAYOUT_SOA 0
#define B200IDA_LAYOUT_AOS 1
/* OR-ed into `layout` of b200ida_getrf_batched: TOLERANCE mode -- for AoS matrices with even n >= 115 the trailing-matrix
 * update of the blocked LU runs on the FP64 tensor cores (mma.sync m8n8k4, SASS DMMA) instead of the bit-exact un-fused
 * multiply / subtract.  Factors then agree with the reference's dense_get_rf to rounding (tests: rel 1e-12), pivot
 * sequences are equal unless two pivot candidates tie to within rounding.  Ignored for other sizes / layouts. */
#define B200IDA_LU_DMMA 0x100
/* OR into `layout` of b200ida_getrf_batched: factor 33 <= n <= 160 (AoS) with the panel / update schedule of the
 * CTA-per-system IDA kernel -- the panel shared by one warp per 32-row slot, the trailing update drawn from a work queue
 * (lu_blocked.cuh, lub_factor) -- instead of the one-warp panel.  Results are bit-identical; at the benchmark batches it is
 * 20-25 % slower than the default.  Exposed so that the team flavour can be tested and measured through
 * the Level-1 entry point.  Ignored for other sizes / layouts. */
#define B200IDA_LU_TEAM_PANEL 0x200

const char* b200ida_last_error(void); /* thread-local text of the last failure */
int b200ida_version(void);
int b200ida_device_count(void);

/* =====================================================================================================
 * Level 1 -- batched dense linear algebra and norms.
 * Replaces, for `batch` independent systems at once:
 *   LSolver::setup   crates/linear/src/traits.rs:52-54  -> Dense::setup / dense_get_rf  crates/linear/src/dense.rs:38-44, 86-158
 *   LSolver::solve   crates/linear/src/traits.rs:70-80  -> Dense::solve / dense_get_rs  crates/linear/src/dense.rs:46-63, 165-206
 *   NormWRMS::norm_wrms  crates/nonlinear/src/norm_wrms.rs:24-33 ; NormRms::norm_wrms  src/norm_rms.rs:31-38 ;
 *   NormRmsMasked::norm_wrms_masked  src/norm_rms.rs:49-57
 * ===================================================================================================== */

/* In-place LU with partial (row) pivoting of `batch` n x n matrices.
 *   A     [in,out]  matrices, `layout` SoA ([n][n][batch]: element (i, j) of system b at (i*n+j)*batch + b)
 *                   or AoS (one column-major matrix per system, see Conventions)
 *   piv   [out]     0-based pivot rows, int64 like the reference's `usize` pivots; SoA [n][batch] or AoS [batch][n]
 *   info  [out]     [batch]; 0, or the 1-based column of the first zero pivot exactly as
 *                   linear::Error::LUFactFail{col} (crates/linear/src/lib.rs:9-13, dense.rs:120-122);
 *                   the factorisation of that matrix stops there, like the reference's early return.
 * Semantics are bit-exact dense_get_rf: pivot = first row with strictly largest |a_ik|; whole-row swap;
 * multipliers a_ik * (1/a_kk); column-oriented update skipped when a_kj == 0.  1 <= n <= 256. */
int b200ida_getrf_batched(double* A, int64_t* piv, int32_t* info, int n, int64_t batch, int layout, int mem,
                          void* stream);

/* Solve A x = b for `batch` systems from the factors and pivots of b200ida_getrf_batched.
 *   bx [in,out]  right-hand sides on entry, solutions on return (the reference copies b into x first,
 *                dense.rs:58-61); SoA [n][batch] or AoS [batch][n]. */
int b200ida_getrs_batched(const double* LU, const int64_t* piv, double* bx, int n, int64_t batch, int layout,
                          int mem, void* stream);

/* out[b] = sqrt( (sum_i (x_i w_i)^2) / n ), summed in index order from 0 (bit-exact with the fold in
 * src/norm_rms.rs:32-37).  x, w: SoA [n][batch] or AoS [batch][n]; out: [batch]. */
int b200ida_wrms_batched(const double* x, const double* w, double* out, int n, int64_t batch, int layout,
                         int mem, void* stream);

/* Masked variant, NormRmsMasked::norm_wrms_masked (src/norm_rms.rs:49-57; what Ida::wrms_norm switches to when
 * `suppressalg` is set, src/lib.rs:1353-1370):  out[b] = sqrt( (sum_i ((x_i w_i) m_i)^2) / n ),  m_i = id[i] ? 1 : 0,
 * same fold order; the divisor stays n like the reference.  id: [n] bytes (non-zero = differential component), one
 * mask for the whole batch (it is a property of the problem, `ida_id` src/lib.rs:122), in the same memory space
 * as x. */
int b200ida_wrms_masked_batched(const double* x, const double* w, const unsigned char* id, double* out, int n,
                                int64_t batch, int layout, int mem, void* stream);

/* =====================================================================================================
 * Level 1.5 -- batched nonlinear solver.  Mirrors `NLSolver<T, D>` (crates/nonlinear/src/traits.rs:129-209) as implemented
 * by `Newton<T, D>` (crates/nonlinear/src/newton.rs:14-167): one `Newton::new(maxiters)` + `solve(problem, y0, y, w, tol,
 * call_lsetup)` per system of the batch.  The `NLProblem` callbacks (traits.rs:5-127: sys / setup / solve / ctest) are
 * device functors selected by `nlp_kind` (a kernel cannot call back into host generics once per Newton iteration):
 *   B200IDA_NLP_QUAD3  the reference's known-answer problem, newton.rs:192-343 (n = 3, analytic Jacobian, Dense LU,
 *                      ctest = wrms(del, w) <= tol).
 * y0, w, y: [n][batch] SoA.  Per-system results (each pointer may be NULL):
 *   status      0 = Ok(()), 1 = Err(Error::ConvergenceRecover), -col = Err(Error::LinearSetupFailed{LUFactFail{col}})
 *               (crates/nonlinear/src/lib.rs:10-35, crates/linear/src/dense.rs:43,121; col is 1-based)
 *   niters      NLSolver::get_num_iters     (newton.rs:169-171)
 *   nconvfails  NLSolver::get_num_conv_fails (newton.rs:177-179)
 * Deviation shared with the oracle: "ConvergenceRecover while the Jacobian is current" returns the recoverable error
 * (SUNDIALS semantics) where the reference's loop does not terminate (newton.rs:146-153).
 * ===================================================================================================== */
#define B200IDA_NLP_QUAD3 0
int b200ida_newton_batched(int nlp_kind, int n, int64_t batch, const double* y0, const double* w, double tol, int maxiters,
                           int call_lsetup, double* y, int32_t* status, int32_t* niters, int32_t* nconvfails, int mem,
                           void* stream);

/* =====================================================================================================
 * Level 2 -- ensemble IDA.  One handle = `batch` independent instances of
 *   Ida<P, Dense, Newton, TolControlSV>   (src/lib.rs:89-95)
 * whose whole solve() loop -- ewt_set, set_coeffs, predict, Newton with res/jac/getrf/getrs, ctest,
 * test_error, restore/handle_n_flag, complete_step, stop tests, get_solution -- runs on the device
 * with per-instance step size, order and retry state; the host never sees a step.
 * ===================================================================================================== */

/* Problems = device-side implementations of the Residual/Jacobian traits (src/traits.rs:28-38, 57-69). */
#define B200IDA_PROBLEM_ROBERTS 0  /* src/sample_problems/roberts.rs, params (k1,k2,k3) = (0.04,1e4,3e7), n=3 */
#define B200IDA_PROBLEM_LORENZ63 1 /* tests/lorenz63.rs posed as F = y' - f(y), params (p,r,b), n=3 */
#define B200IDA_PROBLEM_HEAT2D 2   /* SUNDIALS idaHeat2D-style M x M grid, params (coeff), n=M*M */
#define B200IDA_PROBLEM_USER 3     /* a user-supplied Residual / Jacobian (/ Root) functor (src/traits.rs:28-89) compiled in
                                      with `make -C rust_ida_b200/csrc user PROBLEM_HDR=my_problem.cuh` -- INTEGRATION.md,
                                      "Adding a problem"; n, nparams and the number of root functions come from the functor */

/* IdaTask (src/lib.rs:51-55).  Only Normal is on the hot path. */
#define B200IDA_TASK_NORMAL 0

/* Per-instance status of one solve call.  0 = IdaSolveStatus::Success (src/lib.rs:57-63); negative =
 * the IdaError the reference would have returned (src/error.rs:18-126). */
#define B200IDA_ST_SUCCESS 0
#define B200IDA_ST_TOO_MUCH_WORK (-1) /* mxstep (500) steps in this call, src/impl_solve.rs:249-260 */
#define B200IDA_ST_TOO_MUCH_ACC (-2)  /* IdaError::TooMuchAccuracy, src/impl_solve.rs:296-308 */
#define B200IDA_ST_ERR_FAIL (-3)      /* IDA_ERR_FAIL: MXNEF error-test failures, src/lib.rs:1194-1198 */
#define B200IDA_ST_CONV_FAIL (-4)     /* IdaError::ConvergenceFail: MXNCF failures, src/lib.rs:1232-1234 */
#define B200IDA_ST_LSETUP_FAIL (-5)   /* singular Jacobian; the reference panics (src/ida_ls.rs:287) */
#define B200IDA_ST_BAD_EWT (-6)       /* an ewt component <= 0, src/impl_solve.rs:272-285 */
#define B200IDA_ST_ILL_INPUT (-7)     /* tout too close to t0, src/impl_solve.rs:97-109 */
#define B200IDA_ST_BAD_T (-8)         /* IdaError::BadTimeValue from get_solution, src/lib.rs:1286-1292 */
#define B200IDA_ST_CLOSE_ROOTS (-9)   /* IdaError::CloseRoots, src/impl_r_check.rs:199-203 */
#define B200IDA_ST_BAD_K (-10)       /* IdaError::BadK from get_dky, src/lib.rs:443-447 */
#define B200IDA_ST_ROOT 2             /* IdaSolveStatus::Root (src/lib.rs:57-63): internal to a ladder launch, see b200ida_set_roots */

typedef struct b200ida_ensemble b200ida_ensemble; /* opaque */

/* Ida::new(problem, yy0, yp0, TolControlSV::new(rtol, atol))  -- src/lib.rs:278-283, src/tol_control.rs:59-66.
 *   problem  B200IDA_PROBLEM_*            n       system size (3 for Roberts/Lorenz, M*M for heat)
 *   batch    number of instances          nparams rows of `params`
 *   params   SoA [nparams][batch]         yy0,yp0 SoA [n][batch]       atol host array [n]
 *   mem      where params/yy0/yp0 live    device  CUDA device ordinal
 * tn starts at 0 as in the reference (src/ida_nls.rs:90); there is no t0 argument. */
int b200ida_create(b200ida_ensemble** out, int problem, int n, int64_t batch, int nparams, const double* params,
                   const double* yy0, const double* yp0, double rtol, const double* atol, int mem, int device,
                   void* stream);
int b200ida_destroy(b200ida_ensemble* e);

/* Re-initialise every instance to t = 0 with new initial conditions (and new parameters when `params`
 * is non-NULL), keeping the allocation: the ensemble analogue of constructing a fresh Ida::new
 * (src/lib.rs:278-405) -- all counters, phi, psi.. are reset. */
int b200ida_reinit(b200ida_ensemble* e, const double* params, const double* yy0, const double* yp0, int mem);

/* Instance reordering (process-wide default, read by b200ida_create / b200ida_reinit-with-params).
 * Lanes of a warp advance in lockstep, so the engine may store instances internally in an order that puts
 * instances with similar parameters in the same warp (Morton order of the min/max-normalised parameter
 * vector): they take similar step / Newton / retry decisions and diverge less.  Purely internal: every
 * input and output of this ABI stays in the caller's instance order, and per-instance results are
 * bit-identical with or without it (instances are independent).
 *   OFF never, ON always, AUTO (default) for the thread-per-instance kernels when batch >= 4096. */
#define B200IDA_REORDER_OFF 0
#define B200IDA_REORDER_ON 1
#define B200IDA_REORDER_AUTO 2
int b200ida_set_default_reorder(int mode);

/* Which kernel family integrates the ensemble (all give bit-identical results; tests check that):
 *   AUTO    thread-per-instance for the n = 3 problems, warp-per-system for n <= 32, CTA-per-system above
 *   THREAD  one thread per instance, state in registers + a shared-memory column (csrc/ida_thread.cuh)
 *   WARP    one warp per instance, state in a per-warp shared-memory arena (csrc/ida_team.cuh), n <= 32
 *   CTA     one CTA per instance, LU factors resident in shared memory (csrc/ida_team.cuh), n <= 256 */
#define B200IDA_KERNEL_AUTO 0
#define B200IDA_KERNEL_THREAD 1
#define B200IDA_KERNEL_WARP 2
#define B200IDA_KERNEL_CTA 3
int b200ida_set_kernel(b200ida_ensemble* e, int kernel);

/* Ida::solve(tout, &mut tret, IdaTask::Normal) for every instance -- src/impl_solve.rs:69-376.
 * One kernel launch; returns after enqueueing (device outputs) or after copying `status` back (host).
 *   status [out, may be NULL]  [batch] B200IDA_ST_*, in memory space `mem`. */
int b200ida_solve(b200ida_ensemble* e, double tout, int itask, int32_t* status, int mem);

/* The same, for a whole ladder of output times in ONE launch: equivalent to calling b200ida_solve for
 * touts[0], touts[1], ... in order (h0 depends on the first tout, mxstep is per call), but each instance
 * moves on to its next tout as soon as it reaches the previous one, which removes the per-call
 * reconvergence of slow and fast instances.
 *   yy_traj [out, may be NULL]  SoA [ntout][n][batch] interpolated y(tout_k)  (get_solution, src/lib.rs:1274-1343)
 *   status  [out, may be NULL]  [ntout][batch] */
int b200ida_solve_ladder(b200ida_ensemble* e, const double* touts /* host */, int ntout, int itask, double* yy_traj,
                         int32_t* status, int mem);

/* Engine tuning, results are bit-identical for every setting: large internally re-ordered ensembles on the
 * thread-per-instance production path are integrated in segments of `trips_per_segment` step attempts per instance;
 * between segments the engine re-sorts its slots by a behaviour key (finished, order, phase, ns, log2 |h|) so that the
 * lanes of a warp take the same branches again.  max_segments <= 1 or trips_per_segment == 0:
 * one launch for the whole ladder.  Default: 96 attempts, 4 segments. */
int b200ida_set_regroup(b200ida_ensemble* e, int trips_per_segment, int max_segments);

/* ida_id and ida_suppressalg (src/lib.rs:102,122): id[i] != 0 marks component i as differential, 0 as algebraic.  With
 * suppressalg != 0 every local-error norm -- wrms_norm(.., mask = suppressalg), src/lib.rs:1353-1370: the first step-size
 * guess (src/impl_solve.rs:122-126), the tolsf check (:289-293), test_error (src/lib.rs:983-1004) and complete_step
 * (src/impl_complete_step.rs:76) -- becomes norm_wrms_masked (src/norm_rms.rs:49-57); the Newton convergence norm stays
 * unmasked.  The reference keeps both at their defaults (no setter exists, src/lib.rs:318-321); id == NULL restores
 * them.  Applies to the solves that follow, in every kernel family; not combinable with rootfinding or the trace. */
int b200ida_set_id(b200ida_ensemble* e, const unsigned char* id /* [n], host */, int suppressalg);

/* Per-attempt state trace of ONE instance -- the device side of the reference's `data_trace` feature, which
 * serialises the whole solver at the top of every step attempt (src/lib.rs:297-301, 635-639; scripts/data_trace.ipynb).
 * Until switched off (instance < 0), every b200ida_solve / b200ida_solve_ladder of this handle records (thread
 * kernels: a separate traced instantiation; team kernels: a team-uniform test per attempt): when attempt c of
 * `instance` begins, its complete persistent state is written to row c of
 *   dtrace [max_attempts][10n + 45 + n^2] doubles: phi[6][n], psi, alpha, beta, sigma, gamma [6 each], ee, ewt, yy, yp [n],
 *          hh, hused, rr, cj, cjold, cjlast, cjratio, ss, toldel, eps_newt, tn, tretlast, h0u, tolsf, tret, J[n][n] (col-major)
 *   itrace [max_attempts][16 + n] int32:  kk, kused, knew, phase, ns, setup_done, jcur, nst, nre, nje, nni, netf, ncfn,
 *          nsetups, nlscf, natt, pivots[n]
 * and count[0] (reset to 0 here) counts the attempts seen.  All three are DEVICE buffers owned by the caller. */
int b200ida_set_trace(b200ida_ensemble* e, int64_t instance, double* dtrace, int32_t* itrace, int32_t* count,
                      int max_attempts);

/* Rootfinding -- Root::{num_roots, root} (src/traits.rs:72-89), IDARcheck1/2/3 + IDARootfind (src/impl_r_check.rs:32-576),
 * the Root returns of Ida::solve (src/impl_solve.rs:161-164, 186-227, 335-344) and get_root_info / get_num_g_evals
 * (src/ida_io.rs:107-117).  Available for the Roberts problem (g0 = y1 - 1e-4, g1 = y3 - 0.01,
 * src/sample_problems/roberts.rs:99-113) through a separate instantiation of the thread-per-instance kernel.
 * After b200ida_set_roots(e, max_events > 0, ..) every b200ida_solve_ladder launch plays the caller of
 * examples/roberts.rs:117-140 on the device: when solve() returns Root for an instance, the event is logged and solve()
 * is called again for the same tout, so the per-tout status stays Success and
 *   ev_t      [max_events][batch]      root locations tret, in the order they were found
 *   ev_iroots [max_events][2][batch]   get_root_info at each event (a Scalar array in the reference: signum(g(tlo)) for
 *                                      the g_i that crossed, 0 otherwise -- the reference's sign convention, which is the
 *                                      opposite of SUNDIALS')
 *   ev_count  [batch]                  events seen (beyond max_events: counted, not stored); reset to 0 here
 *   nge       [batch], may be NULL     get_num_g_evals after the launch
 * all DEVICE buffers owned by the caller, indexed by the caller-visible instance id.  Rootfinding never alters the
 * step sequence (same nst / attempts / states as without).  The reference starts with every g_i INACTIVE
 * (`ida_gactive` false, src/lib.rs:373; SUNDIALS: active), so roots in the first step are not seen -- kept as is.
 * max_events <= 0 switches rootfinding off. */
int b200ida_set_roots(b200ida_ensemble* e, int max_events, double* ev_t, double* ev_iroots, int32_t* ev_count,
                      int32_t* nge);

/* get_yy / get_yp (src/ida_io.rs:11-19) and the tret of the last solve; SoA [n][batch] / [batch]; NULL skips. */
int b200ida_get_state(b200ida_ensemble* e, double* yy, double* yp, double* tret, int mem);

/* Ida::get_dky (src/lib.rs:424-529) on every instance: dky[i][b] = k-th derivative at time t of instance b's
 * interpolating polynomial over its last step (k = 0: the solution itself, 1: y', ... up to the order last used).
 * dky [out] [n][batch] SoA; status [out, may be NULL] [batch]: 0, B200IDA_ST_BAD_K (k > order last used by that
 * instance, IdaError::BadK :443-447) or B200IDA_ST_BAD_T (t before the instance's last step, IdaError::BadTimeValue
 * :456-466); the dky column of a failed instance is left as it was.  Reads the persistent state only -- no stepping. */
int b200ida_get_dky(b200ida_ensemble* e, double t, int k, double* dky, int32_t* status, int mem);

/* Statistics, [batch] each, NULL skips.  Names follow src/ida_io.rs:
 *   nst get_num_steps :58   nre get_num_res_evals :64   nje get_num_jac_evals :79
 *   nni get_num_nonlin_solv_iters :84   netf get_num_err_test_fails :74
 *   ncfn get_num_nonlin_solv_conv_fails :104   nsetups get_num_lin_solv_setups :69
 *   nattempts = step attempts (one data_trace frame each, src/lib.rs:635-639). */
int b200ida_get_counters(b200ida_ensemble* e, int64_t* nst, int64_t* nre, int64_t* nje, int64_t* nni, int64_t* netf,
                         int64_t* ncfn, int64_t* nsetups, int64_t* nattempts, int mem);

/* get_last_order :22, get_current_order :27, get_last_step :38, get_current_step :43,
 * get_current_time :48, get_actual_init_step :32 (src/ida_io.rs); [batch] each, NULL skips. */
int b200ida_get_step_info(b200ida_ensemble* e, int32_t* kused, int32_t* kcur, double* hused, double* hcur,
                          double* tcur, double* h0u, int mem);

/* wait for everything enqueued on the ensemble's stream */
int b200ida_sync(b200ida_ensemble* e);

/* =====================================================================================================
 * Measurement helpers (used by bench.py for the FP64 roofline denominator; not part of the drop-in).
 * ===================================================================================================== */
/* Runs a register-resident FP64 loop (kind 0 = DFMA, 1 = DADD, 2 = DMUL) and returns its rate in
 * G instructions/s per GPU (thread-level operations: one DFMA counts 1). */
int b200ida_fp64_issue_rate(int kind, int device, double* ginstr_per_s);

/* Number of the engine's own CUDA kernels launched by Level-2 calls (create / reinit / solve / getters) of this process
 * so far (library kernels such as CUB's sort are not counted).  bench.py reports the difference over its timed region
 * as `gpu_launches`. */
long long b200ida_kernel_launches(void);

/* Diagnostic: how many instances the last b200ida_solve / solve_ladder of this handle handed from the production
 * (fast-path-only) pass to the exact pass -- a fast division / square root / pow left its range, or a rare control
 * path was reached (csrc/ida_thread.cuh, "FAST").  Results do not depend on it; throughput does.  Synchronises.
 * hist16 (may be NULL): histogram of the per-instance flag bytes -- 0 / 1 in production builds; in -DIDA_POISON_REASON=1
 * diagnostic builds 1 + the index of the first range test that failed. */
int b200ida_get_fallback_count(b200ida_ensemble* e, int64_t* count, int64_t* hist16);

/* out[i] = det_pow(x[i], y[i]) on the device: the glibc-2.39-compatible pow the engine uses wherever the
 * reference calls f64::powf (src/impl_complete_step.rs:128-132, src/lib.rs:1163-1169, src/ida_nls.rs:249-253).
 * Exposed so tests can prove device pow == host libm pow bit-for-bit. */
int b200ida_detpow_batched(const double* x, const double* y, double* out, int64_t n, int mem, void* stream);

/* q[i] = a[i] / b[i] and r[i] = 1.0 / b[i] through the straight-line division fast paths of csrc/fastdiv.cuh (what the
 * thread-per-instance kernel uses for ewt_set src/tol_control.rs:80, set_coeffs src/lib.rs:744-750 and dense_get_rs
 * crates/linear/src/dense.rs:196-205).  Exposed so tests can prove them bit-identical to IEEE division.  HOST buffers. */
int b200ida_fastdiv_batched(const double* a, const double* b, double* q, double* r, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* B200IDA_H */
