//! Safe wrapper over `liblbfgsb200.so` with the reference crate's public names.
//!
//! ```ignore
//! use lbfgs_b200::{lbfgs, DeviceBuffer, Rosenbrock};
//! let mut x = DeviceBuffer::from_host(0, &x0)?;          // x lives in HBM
//! let report = lbfgs()
//!     .with_max_iterations(5)
//!     .with_orthantwise(1.0, 0, 99)
//!     .minimize(&mut x, Rosenbrock::new(0)?, |prgr| { println!("{} {}", prgr.niter, prgr.fx); false })?;
//! ```
//! mirrors `lbfgs().with_max_iterations(5).with_orthantwise(1.0, 0, 99).minimize(&mut x, evaluate, progress)`
//! of the reference (src/lib.rs:38-50).  `Progress` / `Report` keep their meaning (src/core.rs:221-299); the
//! new piece is `DeviceEvaluate`, which hands the objective raw device pointers.
use anyhow::{bail, Result};
use lbfgs_b200_sys as sys;
use std::ffi::CStr;
use std::os::raw::{c_int, c_void};

/// f64 device memory owned by Rust.
pub struct DeviceBuffer { ptr: *mut f64, len: usize, device: i32 }
impl DeviceBuffer {
    pub fn new(device: i32, len: usize) -> Result<Self> {
        let mut p: *mut c_void = std::ptr::null_mut();
        let rc = unsafe { sys::lbfgsb200_device_alloc(device, (len * 8) as i64, &mut p) };
        if rc != 0 { bail!("device allocation failed (status {rc}); there is no CPU fallback"); }
        Ok(Self { ptr: p as *mut f64, len, device })
    }
    pub fn from_host(device: i32, src: &[f64]) -> Result<Self> {
        let b = Self::new(device, src.len())?;
        let rc = unsafe { sys::lbfgsb200_copy_h2d(b.ptr as *mut c_void, src.as_ptr() as *const c_void, (src.len() * 8) as i64, std::ptr::null_mut()) };
        if rc != 0 { bail!("host -> device copy failed") }
        Ok(b)
    }
    pub fn to_host(&self, dst: &mut [f64]) -> Result<()> {
        assert_eq!(dst.len(), self.len);
        let rc = unsafe { sys::lbfgsb200_copy_d2h(dst.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, (self.len * 8) as i64, std::ptr::null_mut()) };
        if rc != 0 { bail!("device -> host copy failed") }
        Ok(())
    }
    pub fn as_ptr(&self) -> *const f64 { self.ptr }
    pub fn as_mut_ptr(&mut self) -> *mut f64 { self.ptr }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    pub fn device(&self) -> i32 { self.device }
}
impl Drop for DeviceBuffer { fn drop(&mut self) { unsafe { sys::lbfgsb200_device_free(self.ptr as *mut c_void); } } }

/// Replaces `E: FnMut(&[f64], &mut [f64]) -> Result<f64>` (src/core.rs:10-13): the objective receives raw
/// device pointers and the solver's stream, enqueues its kernels there, writes the gradient to `g_dev` and this
/// rank's partial value to `*fx_dev`.  It must not synchronise.  `Err` has the reference's meaning.
pub trait DeviceEvaluate {
    /// # Safety
    /// `x_dev`/`g_dev` point to `n` f64 in device memory; `fx_dev` to one f64 in device memory.
    unsafe fn evaluate(&mut self, x_dev: *const f64, g_dev: *mut f64, n: usize, stream: *mut c_void, fx_dev: *mut f64) -> Result<()>;
    /// Optional fused line-search trial (x = xp + step*d, gradient, f, g.d, g.g, x.x in one pass).
    fn trial(&mut self) -> Option<(sys::lbfgsb200_trial_eval_fn, *mut c_void)> { None }
    /// Built-in objectives bypass the trampoline.
    fn raw(&mut self) -> Option<(sys::lbfgsb200_eval_fn, *mut c_void)> { None }
}

macro_rules! builtin {
    ($name:ident, $doc:expr, $ctor:expr) => {
        #[doc = $doc]
        pub struct $name { h: *mut sys::lbfgsb200_objective_t }
        impl Drop for $name { fn drop(&mut self) { unsafe { sys::lbfgsb200_objective_destroy(self.h) } } }
        impl DeviceEvaluate for $name {
            unsafe fn evaluate(&mut self, x: *const f64, g: *mut f64, n: usize, s: *mut c_void, fx: *mut f64) -> Result<()> {
                if sys::lbfgsb200_objective_eval(self.h as *mut c_void, x, g, n as i64, s, fx) != 0 { bail!("evaluate failed") }
                Ok(())
            }
            fn raw(&mut self) -> Option<(sys::lbfgsb200_eval_fn, *mut c_void)> { Some((Some(sys::lbfgsb200_objective_eval), self.h as *mut c_void)) }
            fn trial(&mut self) -> Option<(sys::lbfgsb200_trial_eval_fn, *mut c_void)> {
                if unsafe { sys::lbfgsb200_objective_has_trial_eval(self.h) } == 1 {
                    Some((Some(sys::lbfgsb200_objective_trial_eval), self.h as *mut c_void))
                } else { None }
            }
        }
    };
}
builtin!(Rosenbrock, "`default_evaluate()` (src/lib.rs:79-94) on the device.", sys::lbfgsb200_objective_rosenbrock);
builtin!(Booth, "tests/simple.rs:65-74 on the device.", sys::lbfgsb200_objective_booth);
builtin!(LennardJones, "examples/lj.rs:20-64 on the device.", sys::lbfgsb200_objective_lennard_jones);
impl Rosenbrock { pub fn new(device: i32) -> Result<Self> { let mut h = std::ptr::null_mut(); if unsafe { sys::lbfgsb200_objective_rosenbrock(device, &mut h) } != 0 { bail!("no CUDA device") } Ok(Self { h }) } }
impl Booth { pub fn new(device: i32) -> Result<Self> { let mut h = std::ptr::null_mut(); if unsafe { sys::lbfgsb200_objective_booth(device, &mut h) } != 0 { bail!("no CUDA device") } Ok(Self { h }) } }
impl LennardJones { pub fn new(device: i32, epsilon: f64, sigma: f64) -> Result<Self> { let mut h = std::ptr::null_mut(); if unsafe { sys::lbfgsb200_objective_lennard_jones(device, epsilon, sigma, &mut h) } != 0 { bail!("no CUDA device") } Ok(Self { h }) } }

/// src/core.rs:221-250; `x` / `gx` are device pointers.
#[derive(Debug, Clone)]
pub struct Progress { pub x: *const f64, pub gx: *const f64, pub n: usize, pub fx: f64, pub xnorm: f64, pub gnorm: f64,
                      pub step: f64, pub niter: usize, pub neval: usize, pub ncall: usize }
/// src/core.rs:271-285
#[derive(Debug, Clone, Default)]
pub struct Report { pub fx: f64, pub xnorm: f64, pub gnorm: f64, pub neval: usize }

fn progress_from(p: &sys::lbfgsb200_progress_t) -> Progress {
    Progress { x: p.x_dev, gx: p.gx_dev, n: p.n_local as usize, fx: p.fx, xnorm: p.xnorm, gnorm: p.gnorm, step: p.step,
               niter: p.niter as usize, neval: p.neval as usize, ncall: p.ncall as usize }
}

unsafe extern "C" fn eval_tramp<E: DeviceEvaluate>(user: *mut c_void, x: *const f64, g: *mut f64, n: i64, stream: *mut c_void, fx: *mut f64) -> c_int {
    let e = &mut *(user as *mut E);
    match e.evaluate(x, g, n as usize, stream, fx) { Ok(()) => 0, Err(_) => 1 }
}
unsafe extern "C" fn progress_tramp<G: FnMut(&Progress) -> bool>(user: *mut c_void, p: *const sys::lbfgsb200_progress_t) -> c_int {
    let g = &mut *(user as *mut G);
    if g(&progress_from(&*p)) { 1 } else { 0 }
}

/// The builder (src/lbfgs.rs:179-384); `param` is private like the reference's.
#[derive(Clone, Debug)]
pub struct Lbfgs { param: sys::lbfgsb200_param_t, fused_trial: bool }
impl Default for Lbfgs {
    fn default() -> Self {
        let mut p = std::mem::MaybeUninit::<sys::lbfgsb200_param_t>::zeroed();
        unsafe { sys::lbfgsb200_param_default(p.as_mut_ptr()); Self { param: p.assume_init(), fused_trial: true } }
    }
}
/// Create a default LBFGS optimizer (src/lib.rs:74-76).
pub fn lbfgs() -> Lbfgs { Lbfgs::default() }

impl Lbfgs {
    pub fn with_epsilon(mut self, epsilon: f64) -> Self { assert!(epsilon.is_sign_positive(), "Invalid parameter epsilon specified."); self.param.epsilon = epsilon; self }
    pub fn with_initial_step_size(mut self, b: f64) -> Self { assert!(b.is_sign_positive(), "Invalid beta parameter for scaling the initial step size."); self.param.initial_inverse_hessian = b; self }
    pub fn with_max_step_size(mut self, s: f64) -> Self { assert!(s.is_sign_positive(), "Invalid max_step_size parameter."); self.param.max_step_size = s; self }
    pub fn with_damping(mut self, damped: bool) -> Self { self.param.damping = damped as i64; self }
    pub fn with_orthantwise(mut self, c: f64, start: usize, end: impl Into<Option<usize>>) -> Self {
        assert!(c.is_sign_positive(), "Invalid parameter orthantwise c parameter specified.");
        self.param.orthantwise = 1; self.param.owl_c = c; self.param.owl_start = start as i64;
        self.param.owl_end = end.into().map(|e| e as i64).unwrap_or(-1); self
    }
    pub fn with_linesearch_ftol(mut self, ftol: f64) -> Self { assert!(ftol >= 0.0, "Invalid parameter ftol specified."); self.param.ls_ftol = ftol; self }
    pub fn with_linesearch_gtol(mut self, gtol: f64) -> Self {
        assert!(gtol >= 0.0 && gtol < 1.0 && gtol > self.param.ls_ftol, "Invalid parameter gtol specified."); self.param.ls_gtol = gtol; self
    }
    pub fn with_gradient_only(mut self) -> Self {
        self.param.ls_gradient_only = 1; self.param.damping = 1; self.param.ls_algorithm = sys::LBFGSB200_LS_BACKTRACKING_STRONG_WOLFE; self
    }
    pub fn with_max_linesearch(mut self, n: usize) -> Self { self.param.ls_max_linesearch = n as i64; self }
    pub fn with_linesearch_xtol(mut self, xtol: f64) -> Self { assert!(xtol >= 0.0, "Invalid parameter xtol specified."); self.param.ls_xtol = xtol; self }
    pub fn with_linesearch_min_step(mut self, min_step: f64) -> Self { assert!(min_step >= 0.0, "Invalid parameter min_step specified."); self.param.ls_min_step = min_step; self }
    pub fn with_max_iterations(mut self, niter: usize) -> Self { self.param.max_iterations = niter as i64; self }
    pub fn with_max_evaluations(mut self, neval: usize) -> Self { self.param.max_evaluations = neval as i64; self }
    pub fn with_fx_delta(mut self, delta: f64, past: usize) -> Self { assert!(delta >= 0.0, "Invalid parameter delta specified."); self.param.delta = delta; self.param.past = past as i64; self }
    pub fn with_linesearch_algorithm(mut self, algo: &str) -> Self {
        self.param.ls_algorithm = match algo {
            "MoreThuente" => sys::LBFGSB200_LS_MORETHUENTE,
            "BacktrackingArmijo" => sys::LBFGSB200_LS_BACKTRACKING_ARMIJO,
            "BacktrackingStrongWolfe" => sys::LBFGSB200_LS_BACKTRACKING_STRONG_WOLFE,
            "BacktrackingWolfe" | "Backtracking" => sys::LBFGSB200_LS_BACKTRACKING_WOLFE,
            _ => unimplemented!(),
        };
        self
    }
    // extensions
    pub fn with_m(mut self, m: usize) -> Self { assert!(m >= 1); self.param.m = m as i64; self }
    pub fn with_sequential_reduction(mut self, on: bool) -> Self { self.param.reduction = on as i64; self }
    pub fn with_fused_trial(mut self, on: bool) -> Self { self.fused_trial = on; self }

    /// `minimize(&mut x, eval_fn, prgr_fn)` (src/lbfgs.rs:399-421) with x in device memory.
    pub fn minimize<E, G>(self, x: &mut DeviceBuffer, mut eval_fn: E, mut prgr_fn: G) -> Result<Report>
    where E: DeviceEvaluate, G: FnMut(&Progress) -> bool {
        let xp = x.as_mut_ptr();
        let state = self.create(x.len(), x.device(), &mut eval_fn)?;
        let mut rep = sys::lbfgsb200_report_t::default();
        let st = unsafe { sys::lbfgsb200_minimize(state.solver, xp, state.eval.0, state.eval.1, Some(progress_tramp::<G>),
                                                  &mut prgr_fn as *mut G as *mut c_void, &mut rep) };
        if st < 0 { bail!("{}", state.last_error()) }
        Ok(Report { fx: rep.fx, xnorm: rep.xnorm, gnorm: rep.gnorm, neval: rep.neval as usize })
    }

    /// `build` (src/lbfgs.rs:443-481): the iterative API.  `x` and `eval_fn` stay borrowed while the state lives.
    pub fn build<'a, E: DeviceEvaluate>(self, x: &'a mut DeviceBuffer, eval_fn: &'a mut E) -> Result<LbfgsState<'a>> {
        let state = self.create(x.len(), x.device(), eval_fn)?;
        let rc = unsafe { sys::lbfgsb200_build(state.solver, x.as_mut_ptr(), state.eval.0, state.eval.1) };
        if rc != 0 { bail!("{}", state.last_error()) }
        Ok(state)
    }

    fn create<'a, E: DeviceEvaluate>(self, n: usize, device: i32, eval_fn: &'a mut E) -> Result<LbfgsState<'a>> {
        let mut solver = std::ptr::null_mut();
        let rc = unsafe { sys::lbfgsb200_create(&self.param, n as i64, n as i64, 0, device, std::ptr::null_mut(), std::ptr::null_mut(), &mut solver) };
        if rc != 0 { bail!("lbfgsb200_create failed with status {rc} (no CUDA device? there is no CPU fallback)") }
        let eval = eval_fn.raw().unwrap_or((Some(eval_tramp::<E>), eval_fn as *mut E as *mut c_void));
        if self.fused_trial { if let Some((tf, tu)) = eval_fn.trial() { unsafe { sys::lbfgsb200_set_trial_evaluate(solver, tf, tu); } } }
        Ok(LbfgsState { solver, eval, _x: std::marker::PhantomData })
    }
}

/// src/lbfgs.rs:425-566
pub struct LbfgsState<'a> { solver: *mut sys::lbfgsb200_solver_t, eval: (sys::lbfgsb200_eval_fn, *mut c_void), _x: std::marker::PhantomData<&'a mut ()> }
impl<'a> LbfgsState<'a> {
    pub fn is_converged(&mut self) -> bool { let mut st = 0; unsafe { sys::lbfgsb200_is_converged(self.solver, &mut st) == 1 } }
    pub fn propagate(&mut self) -> Result<Progress> {
        let mut p = std::mem::MaybeUninit::<sys::lbfgsb200_progress_t>::zeroed();
        let rc = unsafe { sys::lbfgsb200_propagate(self.solver, p.as_mut_ptr()) };
        if rc != 0 { bail!("{}", self.last_error()) }
        Ok(progress_from(unsafe { &p.assume_init() }))
    }
    pub fn report(&mut self) -> Report {
        let mut r = sys::lbfgsb200_report_t::default();
        unsafe { sys::lbfgsb200_report(self.solver, &mut r); }
        Report { fx: r.fx, xnorm: r.xnorm, gnorm: r.gnorm, neval: r.neval as usize }
    }
    /// x and xp ping-pong between two buffers; this makes the caller's buffer hold the current point.
    pub fn finish(&mut self) -> Result<()> { if unsafe { sys::lbfgsb200_finish(self.solver) } != 0 { bail!("{}", self.last_error()) } Ok(()) }
    fn last_error(&self) -> String { unsafe { CStr::from_ptr(sys::lbfgsb200_last_error(self.solver)).to_string_lossy().into_owned() } }
}
impl<'a> Drop for LbfgsState<'a> { fn drop(&mut self) { unsafe { sys::lbfgsb200_destroy(self.solver) } } }
