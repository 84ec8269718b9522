//! Rust binding of the B200 batched IDA engine (include/b200ida.h).
//!
//! UNCOMPILED SOURCE: there is no Rust toolchain in the build image, so this file has never been through
//! rustc.  It shows exactly what a rust-ida maintainer binds: the `extern "C"` block mirrors include/b200ida.h
//! one to one; `GpuDense` implements the reference's own `linear::LSolver<f64, D>` trait
//! (crates/linear/src/traits.rs:27-91) for a batch of one, and `EnsembleIda` is the ensemble analogue of
//! `Ida::new` / `Ida::solve` / the getters (src/lib.rs:278-283, src/impl_solve.rs:69-74, src/ida_io.rs).

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub const MEM_DEVICE: c_int = 0;
pub const MEM_HOST: c_int = 1;
pub const LAYOUT_SOA: c_int = 0;
pub const LAYOUT_AOS: c_int = 1;
pub const PROBLEM_ROBERTS: c_int = 0;
pub const PROBLEM_LORENZ63: c_int = 1;
pub const PROBLEM_HEAT2D: c_int = 2;
pub const TASK_NORMAL: c_int = 0;

#[repr(C)]
pub struct RawEnsemble {
    _private: [u8; 0],
}

extern "C" {
    pub fn b200ida_last_error() -> *const c_char;
    pub fn b200ida_version() -> c_int;
    pub fn b200ida_device_count() -> c_int;
    pub fn b200ida_getrf_batched(a: *mut f64, piv: *mut i64, info: *mut i32, n: c_int, batch: i64, layout: c_int,
                                 mem: c_int, stream: *mut c_void) -> c_int;
    pub fn b200ida_getrs_batched(lu: *const f64, piv: *const i64, bx: *mut f64, n: c_int, batch: i64, layout: c_int,
                                 mem: c_int, stream: *mut c_void) -> c_int;
    pub fn b200ida_wrms_batched(x: *const f64, w: *const f64, out: *mut f64, n: c_int, batch: i64, layout: c_int,
                                mem: c_int, stream: *mut c_void) -> c_int;
    pub fn b200ida_wrms_masked_batched(x: *const f64, w: *const f64, id: *const u8, out: *mut f64, n: c_int, batch: i64,
                                       layout: c_int, mem: c_int, stream: *mut c_void) -> c_int;
    /// NLSolver level (crates/nonlinear/src/traits.rs:129-209): Newton::solve over a batch of device NLProblems
    pub fn b200ida_newton_batched(nlp_kind: c_int, n: c_int, batch: i64, y0: *const f64, w: *const f64, tol: f64,
                                  maxiters: c_int, call_lsetup: c_int, y: *mut f64, status: *mut i32, niters: *mut i32,
                                  nconvfails: *mut i32, mem: c_int, stream: *mut c_void) -> c_int;
    /// ida_id / ida_suppressalg (src/lib.rs:102,122)
    pub fn b200ida_set_id(e: *mut RawEnsemble, id: *const u8, suppressalg: c_int) -> c_int;
    pub fn b200ida_create(out: *mut *mut RawEnsemble, problem: c_int, n: c_int, batch: i64, nparams: c_int,
                          params: *const f64, yy0: *const f64, yp0: *const f64, rtol: f64, atol: *const f64,
                          mem: c_int, device: c_int, stream: *mut c_void) -> c_int;
    pub fn b200ida_destroy(e: *mut RawEnsemble) -> c_int;
    /// rootfinding (src/impl_r_check.rs): event log per instance, device buffers; max_events <= 0 switches it off
    pub fn b200ida_set_roots(e: *mut RawEnsemble, max_events: c_int, ev_t: *mut f64, ev_iroots: *mut f64, ev_count: *mut i32,
                             nge: *mut i32) -> c_int;
    /// per-attempt state trace of one instance (the `data_trace` feature, src/lib.rs:297-301, 635-639); device buffers
    pub fn b200ida_set_trace(e: *mut RawEnsemble, instance: i64, dtrace: *mut f64, itrace: *mut i32, count: *mut i32,
                             max_attempts: c_int) -> c_int;
    pub fn b200ida_reinit(e: *mut RawEnsemble, params: *const f64, yy0: *const f64, yp0: *const f64, mem: c_int) -> c_int;
    pub fn b200ida_set_kernel(e: *mut RawEnsemble, kernel: c_int) -> c_int;
    pub fn b200ida_solve(e: *mut RawEnsemble, tout: f64, itask: c_int, status: *mut i32, mem: c_int) -> c_int;
    pub fn b200ida_solve_ladder(e: *mut RawEnsemble, touts: *const f64, ntout: c_int, itask: c_int, yy_traj: *mut f64,
                                status: *mut i32, mem: c_int) -> c_int;
    pub fn b200ida_get_state(e: *mut RawEnsemble, yy: *mut f64, yp: *mut f64, tret: *mut f64, mem: c_int) -> c_int;
    pub fn b200ida_get_counters(e: *mut RawEnsemble, nst: *mut i64, nre: *mut i64, nje: *mut i64, nni: *mut i64,
                                netf: *mut i64, ncfn: *mut i64, nsetups: *mut i64, nattempts: *mut i64, mem: c_int) -> c_int;
    pub fn b200ida_get_step_info(e: *mut RawEnsemble, kused: *mut i32, kcur: *mut i32, hused: *mut f64, hcur: *mut f64,
                                 tcur: *mut f64, h0u: *mut f64, mem: c_int) -> c_int;
    pub fn b200ida_get_dky(e: *mut RawEnsemble, t: f64, k: c_int, dky: *mut f64, status: *mut i32, mem: c_int) -> c_int;
    pub fn b200ida_sync(e: *mut RawEnsemble) -> c_int;
}

#[derive(Debug, thiserror::Error)]
#[error("b200ida error {code}: {msg}")]
pub struct EngineError {
    pub code: i32,
    pub msg: String,
}

fn check(rc: c_int) -> Result<(), EngineError> {
    if rc == 0 {
        Ok(())
    } else {
        let msg = unsafe { CStr::from_ptr(b200ida_last_error()) }.to_string_lossy().into_owned();
        Err(EngineError { code: rc, msg })
    }
}

// ---------------------------------------------------------------------------------------------------------------
// `impl LSolver<f64, D> for GpuDense<D>`: drop-in for `linear::Dense<D>` (crates/linear/src/dense.rs:22-63).
// nalgebra's owned storage is column-major contiguous = the ABI's AoS layout with batch = 1.
// ---------------------------------------------------------------------------------------------------------------
use linear::{Error as LError, LSolver, LSolverType};
use nalgebra::{allocator::Allocator, DefaultAllocator, DimName, Matrix, OVector, StorageMut, U1};

pub struct GpuDense<D: DimName>
where
    DefaultAllocator: Allocator<i64, D>,
{
    pivots: OVector<i64, D>,
}

impl<D> LSolver<f64, D> for GpuDense<D>
where
    D: DimName,
    DefaultAllocator: Allocator<f64, D> + Allocator<i64, D>,
{
    fn new() -> Self {
        GpuDense { pivots: OVector::<i64, D>::zeros() }
    }

    fn get_type(&self) -> LSolverType {
        LSolverType::Direct
    }

    fn setup<S>(&mut self, mat_a: &mut Matrix<f64, D, D, S>) -> Result<(), LError>
    where
        S: StorageMut<f64, D, D>,
    {
        let mut info: i32 = 0;
        let rc = unsafe {
            b200ida_getrf_batched(mat_a.as_mut_ptr(), self.pivots.as_mut_ptr(), &mut info, D::dim() as c_int, 1,
                                  LAYOUT_AOS, MEM_HOST, std::ptr::null_mut())
        };
        assert_eq!(rc, 0, "b200ida_getrf_batched failed");
        if info != 0 {
            // same error, same 1-based column as dense.rs:43,121
            return Err(LError::LUFactFail { col: info as usize });
        }
        Ok(())
    }

    fn solve<SA, SB, SC>(&self, mat_a: &Matrix<f64, D, D, SA>, x: &mut Matrix<f64, D, U1, SB>,
                         b: &Matrix<f64, D, U1, SC>, _tol: f64) -> Result<(), LError>
    where
        SA: StorageMut<f64, D, D>,
        SB: StorageMut<f64, D>,
        SC: StorageMut<f64, D>,
    {
        x.set_column(0, b); // dense.rs:58-59
        let rc = unsafe {
            b200ida_getrs_batched(mat_a.as_ptr(), self.pivots.as_ptr(), x.as_mut_ptr(), D::dim() as c_int, 1, LAYOUT_AOS,
                                  MEM_HOST, std::ptr::null_mut())
        };
        assert_eq!(rc, 0, "b200ida_getrs_batched failed");
        Ok(())
    }
}

// ---------------------------------------------------------------------------------------------------------------
// EnsembleIda: Ida::new / Ida::solve / getters over `batch` instances (host buffers, SoA [n][batch]).
// ---------------------------------------------------------------------------------------------------------------
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum IdaTask {
    Normal,
}

pub struct EnsembleIda {
    h: *mut RawEnsemble,
    n: usize,
    batch: usize,
}

impl EnsembleIda {
    /// `Ida::new(problem, yy0, yp0, TolControlSV::new(rtol, atol))`
    pub fn new(problem: c_int, n: usize, batch: usize, params: &[f64], yy0: &[f64], yp0: &[f64], rtol: f64,
               atol: &[f64]) -> Result<Self, EngineError> {
        assert_eq!(yy0.len(), n * batch); // src/lib.rs:284
        assert_eq!(yp0.len(), n * batch);
        assert_eq!(atol.len(), n);
        let nparams = params.len() / batch;
        let mut h: *mut RawEnsemble = std::ptr::null_mut();
        check(unsafe {
            b200ida_create(&mut h, problem, n as c_int, batch as i64, nparams as c_int, params.as_ptr(), yy0.as_ptr(),
                           yp0.as_ptr(), rtol, atol.as_ptr(), MEM_HOST, 0, std::ptr::null_mut())
        })?;
        Ok(EnsembleIda { h, n, batch })
    }

    /// `Ida::solve(tout, &mut tret, itask)`: per-instance status (0 = IdaSolveStatus::Success)
    pub fn solve(&mut self, tout: f64, _itask: IdaTask) -> Result<Vec<i32>, EngineError> {
        let mut st = vec![0i32; self.batch];
        check(unsafe { b200ida_solve(self.h, tout, TASK_NORMAL, st.as_mut_ptr(), MEM_HOST) })?;
        Ok(st)
    }

    pub fn get_yy(&self) -> Result<Vec<f64>, EngineError> {
        let mut y = vec![0f64; self.n * self.batch];
        check(unsafe { b200ida_get_state(self.h, y.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut(), MEM_HOST) })?;
        Ok(y)
    }

    /// every output time of `touts` in one launch; `y_traj`, if wanted, is `[ntout][n][batch]`
    pub fn solve_ladder(&mut self, touts: &[f64], want_traj: bool) -> Result<(Vec<i32>, Option<Vec<f64>>), EngineError> {
        let mut st = vec![0i32; touts.len() * self.batch];
        let mut traj = if want_traj { Some(vec![0f64; touts.len() * self.n * self.batch]) } else { None };
        let tp = traj.as_mut().map_or(std::ptr::null_mut(), |v| v.as_mut_ptr());
        check(unsafe { b200ida_solve_ladder(self.h, touts.as_ptr(), touts.len() as c_int, TASK_NORMAL, tp, st.as_mut_ptr(), MEM_HOST) })?;
        Ok((st, traj))
    }

    /// `Ida::get_dky(t, k, &mut dky)` (src/lib.rs:424-529) on every instance: `dky` is `[n][batch]`, the returned
    /// vector holds each instance's outcome (0, BadK = -10, BadTimeValue = -8); failed columns keep their contents
    pub fn get_dky(&self, t: f64, k: usize, dky: &mut [f64]) -> Result<Vec<i32>, EngineError> {
        assert_eq!(dky.len(), self.n * self.batch);
        let mut st = vec![0i32; self.batch];
        check(unsafe { b200ida_get_dky(self.h, t, k as c_int, dky.as_mut_ptr(), st.as_mut_ptr(), MEM_HOST) })?;
        Ok(st)
    }

    pub fn get_yp(&self) -> Result<Vec<f64>, EngineError> {
        let mut y = vec![0f64; self.n * self.batch];
        check(unsafe { b200ida_get_state(self.h, std::ptr::null_mut(), y.as_mut_ptr(), std::ptr::null_mut(), MEM_HOST) })?;
        Ok(y)
    }

    /// `get_last_order` / `get_last_step` (src/ida_io.rs:22, 38)
    pub fn get_last_order(&self) -> Result<Vec<i32>, EngineError> {
        let mut v = vec![0i32; self.batch];
        let (zi, zd) = (std::ptr::null_mut::<i32>(), std::ptr::null_mut::<f64>());
        check(unsafe { b200ida_get_step_info(self.h, v.as_mut_ptr(), zi, zd, zd, zd, zd, MEM_HOST) })?;
        Ok(v)
    }
    pub fn get_last_step(&self) -> Result<Vec<f64>, EngineError> {
        let mut v = vec![0f64; self.batch];
        let (zi, zd) = (std::ptr::null_mut::<i32>(), std::ptr::null_mut::<f64>());
        check(unsafe { b200ida_get_step_info(self.h, zi, zi, v.as_mut_ptr(), zd, zd, zd, MEM_HOST) })?;
        Ok(v)
    }

    pub fn get_num_steps(&self) -> Result<Vec<i64>, EngineError> {
        let mut c = vec![0i64; self.batch];
        let z = std::ptr::null_mut();
        check(unsafe { b200ida_get_counters(self.h, c.as_mut_ptr(), z, z, z, z, z, z, z, MEM_HOST) })?;
        Ok(c)
    }
}

impl Drop for EnsembleIda {
    fn drop(&mut self) {
        unsafe { b200ida_destroy(self.h) };
    }
}


// ---- NLSolver level ------------------------------------------------------------------------------------------------
/// Problem kinds with a device functor (include/b200ida.h B200IDA_NLP_*)
pub const NLP_QUAD3: c_int = 0;

/// `impl NLSolver<f64, D>`-shaped wrapper over a batch: `GpuNewton::new(maxiters)` then `solve(..)`, with the reference's
/// getters.  `nonlinear::NLSolver::solve` takes the problem as `&mut NLP where NLP: NLProblem<T, D>`; a GPU kernel cannot
/// call those closures per iteration, so the batched form takes the problem KIND instead and returns one
/// `Result`-equivalent status per system: 0 = Ok(()), 1 = Err(Error::ConvergenceRecover {}), -col =
/// Err(Error::LinearSetupFailed { source: LUFactFail { col } }) (crates/nonlinear/src/lib.rs:10-35).
pub struct GpuNewton {
    maxiters: usize,
    niters: Vec<usize>,
    nconvfails: Vec<usize>,
    curiter: Vec<usize>,
}

impl GpuNewton {
    /// NLSolver::new (traits.rs:136)
    pub fn new(maxiters: usize) -> Self {
        GpuNewton { maxiters, niters: Vec::new(), nconvfails: Vec::new(), curiter: Vec::new() }
    }

    /// NLSolver::solve (traits.rs:176-193) for `batch` systems; y0, y, w are SoA [n][batch]
    pub fn solve(&mut self, nlp_kind: c_int, n: usize, batch: usize, y0: &[f64], y: &mut [f64], w: &[f64], tol: f64,
                 call_lsetup: bool) -> Result<Vec<i32>, EngineError> {
        assert_eq!(y0.len(), n * batch);
        assert_eq!(y.len(), n * batch);
        assert_eq!(w.len(), n * batch);
        let mut st = vec![0i32; batch];
        let mut ni = vec![0i32; batch];
        let mut nc = vec![0i32; batch];
        check(unsafe {
            b200ida_newton_batched(nlp_kind, n as c_int, batch as i64, y0.as_ptr(), w.as_ptr(), tol, self.maxiters as c_int,
                                   call_lsetup as c_int, y.as_mut_ptr(), st.as_mut_ptr(), ni.as_mut_ptr(), nc.as_mut_ptr(),
                                   MEM_HOST, std::ptr::null_mut())
        })?;
        if self.niters.len() != batch {
            self.niters = vec![0; batch];
            self.nconvfails = vec![0; batch];
        }
        for b in 0..batch {
            self.niters[b] += ni[b] as usize;
            self.nconvfails[b] += nc[b] as usize;
        }
        self.curiter = ni.iter().map(|&v| v as usize).collect();
        Ok(st)
    }

    /// NLSolver::get_num_iters / get_cur_iter / get_num_conv_fails (traits.rs:196-208), per system
    pub fn get_num_iters(&self) -> &[usize] { &self.niters }
    pub fn get_cur_iter(&self) -> &[usize] { &self.curiter }
    pub fn get_num_conv_fails(&self) -> &[usize] { &self.nconvfails }
}
