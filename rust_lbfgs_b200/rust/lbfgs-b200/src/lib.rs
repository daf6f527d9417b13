//! Safe wrapper over `liblbfgsb200.so` with the reference crate's public names.
//!
//! ```ignore
//! use lbfgs_b200::{default_progress, lbfgs, DeviceBuffer, Rosenbrock};
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
    /// Optional fused line-search entries (`lbfgsb200_fused_ops_t`): write-free probes + one commit per iteration,
    /// and/or the one-pass trial.  `None`: trials run as K1 + evaluate + K2.
    fn fused_ops(&mut self) -> Option<sys::lbfgsb200_fused_ops_t> { None }
    /// Multi-GPU: hand the objective the solve's communicator (see `lbfgsb200_objective_set_shard`); called by
    /// `Lbfgs::with_shard` solves before the objective's fused entries are queried.
    fn attach_comm(&mut self, _comm: *mut sys::lbfgsb200_comm_t) -> Result<()> { Ok(()) }
    /// Built-in objectives bypass the trampoline.
    fn raw(&mut self) -> Option<(sys::lbfgsb200_eval_fn, *mut c_void)> { None }
}

macro_rules! builtin {
    ($name:ident, $doc:expr, $ctor:expr) => {
        #[doc = $doc]
        pub struct $name { h: *mut sys::lbfgsb200_objective_t, offsets: Vec<i64> }
        impl $name { fn offsets_ptr(&self) -> *const i64 { if self.offsets.is_empty() { std::ptr::null() } else { self.offsets.as_ptr() } } }
        impl Drop for $name { fn drop(&mut self) { unsafe { sys::lbfgsb200_objective_destroy(self.h) } } }
        impl DeviceEvaluate for $name {
            unsafe fn evaluate(&mut self, x: *const f64, g: *mut f64, n: usize, s: *mut c_void, fx: *mut f64) -> Result<()> {
                if sys::lbfgsb200_objective_eval(self.h as *mut c_void, x, g, n as i64, s, fx) != 0 { bail!("evaluate failed") }
                Ok(())
            }
            fn raw(&mut self) -> Option<(sys::lbfgsb200_eval_fn, *mut c_void)> { Some((Some(sys::lbfgsb200_objective_eval), self.h as *mut c_void)) }
            fn fused_ops(&mut self) -> Option<sys::lbfgsb200_fused_ops_t> {
                let mut ops = sys::lbfgsb200_fused_ops_t::default();
                if unsafe { sys::lbfgsb200_objective_fused_ops(self.h, &mut ops) } != 0 { return None }
                if ops.trial.is_none() && ops.probe.is_none() { None } else { Some(ops) }
            }
            fn attach_comm(&mut self, comm: *mut sys::lbfgsb200_comm_t) -> Result<()> {
                // Lennard-Jones needs its shard offsets first: `LennardJones::shard(offsets)`
                if unsafe { sys::lbfgsb200_objective_set_shard(self.h, comm, self.offsets_ptr()) } != 0 { bail!("lbfgsb200_objective_set_shard failed") }
                Ok(())
            }
        }
        impl $name {
            /// the raw objective handle (`lbfgsb200_objective_t*`)
            pub fn handle(&self) -> *mut sys::lbfgsb200_objective_t { self.h }
        }
    };
}
builtin!(Rosenbrock, "`default_evaluate()` (src/lib.rs:79-94) on the device.", sys::lbfgsb200_objective_rosenbrock);
builtin!(Booth, "tests/simple.rs:65-74 on the device.", sys::lbfgsb200_objective_booth);
builtin!(LennardJones, "examples/lj.rs:20-64 on the device.", sys::lbfgsb200_objective_lennard_jones);
builtin!(Glm, "Dense GLM (tests/owlqn.rs:22-43 Poisson, or logistic): X row-major nrow x ncol in device memory.", sys::lbfgsb200_objective_glm);
impl Rosenbrock { pub fn new(device: i32) -> Result<Self> { let mut h = std::ptr::null_mut(); if unsafe { sys::lbfgsb200_objective_rosenbrock(device, &mut h) } != 0 { bail!("no CUDA device") } Ok(Self { h, offsets: vec![] }) } }
impl Booth { pub fn new(device: i32) -> Result<Self> { let mut h = std::ptr::null_mut(); if unsafe { sys::lbfgsb200_objective_booth(device, &mut h) } != 0 { bail!("no CUDA device") } Ok(Self { h, offsets: vec![] }) } }
impl LennardJones {
    pub fn new(device: i32, epsilon: f64, sigma: f64) -> Result<Self> { let mut h = std::ptr::null_mut(); if unsafe { sys::lbfgsb200_objective_lennard_jones(device, epsilon, sigma, &mut h) } != 0 { bail!("no CUDA device") } Ok(Self { h, offsets: vec![] }) }
    /// The 1/r^2 molecular-dynamics arithmetic with fused multiply-adds instead of the reference's per-pair arithmetic.
    pub fn fast(self, on: bool) -> Self { unsafe { sys::lbfgsb200_objective_set_lj_fast(self.h, on as c_int); } self }
    /// Atoms sharded over the ranks: `offsets[r]..offsets[r + 1]` are rank r's elements (multiples of 3).
    pub fn shard(mut self, offsets: &[i64]) -> Self { self.offsets = offsets.to_vec(); self }
}
/// GLM kind: `tests/owlqn.rs:22-43` is Poisson.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum GlmKind { Poisson = 0, Logistic = 1 }
impl Glm {
    /// `x` (nrow x ncol, row-major) and `y` (nrow) stay borrowed by the objective: keep the buffers alive.
    pub fn new(device: i32, kind: GlmKind, x: &DeviceBuffer, y: &DeviceBuffer, nrow: usize, ncol: usize) -> Result<Self> {
        assert_eq!(x.len(), nrow * ncol); assert_eq!(y.len(), nrow);
        let mut h = std::ptr::null_mut();
        if unsafe { sys::lbfgsb200_objective_glm(device, kind as c_int, x.as_ptr(), y.as_ptr(), nrow as i64, ncol as i64, &mut h) } != 0 { bail!("creating the GLM objective failed") }
        Ok(Self { h, offsets: vec![] })
    }
    /// Which kernels the last evaluation ran (`LBFGSB200_GLM_PATH_*`).
    pub fn last_path(&self) -> i32 { unsafe { sys::lbfgsb200_objective_last_path(self.h) } }
}

/// The fixture loader of the reference's OWL-QN test (tests/owlqn.rs:66-83): a headerless CSV of f64, row-major.
pub fn read_csv(path: &std::path::Path) -> Result<(Vec<f64>, usize, usize)> {
    let text = std::fs::read_to_string(path)?;
    let mut data = Vec::new();
    let (mut nrow, mut ncol) = (0usize, 0usize);
    for line in text.lines().filter(|l| !l.trim().is_empty()) {
        let row: Vec<f64> = line.split(',').map(|t| t.trim().parse::<f64>()).collect::<std::result::Result<_, _>>()?;
        if nrow == 0 { ncol = row.len() } else if row.len() != ncol { bail!("ragged CSV: row {} has {} fields, expected {}", nrow, row.len(), ncol) }
        data.extend(row);
        nrow += 1;
    }
    Ok((data, nrow, ncol))
}

/// One rank of a multi-GPU solve (`lbfgsb200_comm_t`): one process per GPU, the unique id comes from rank 0.
pub struct Comm { handle: *mut sys::lbfgsb200_comm_t, pub rank: i32, pub nranks: i32 }
impl Comm {
    /// Rank 0 calls this and ships the 128 bytes to the other ranks (MPI, a file, a socket ...).
    pub fn unique_id() -> Result<[u8; 128]> {
        let mut id = [0u8; 128];
        if unsafe { sys::lbfgsb200_comm_unique_id(id.as_mut_ptr() as *mut std::os::raw::c_char) } != 0 { bail!("ncclGetUniqueId failed (libnccl.so.2 not loadable?)") }
        Ok(id)
    }
    pub fn new(id: &[u8; 128], rank: i32, nranks: i32, device: i32) -> Result<Self> {
        let mut handle = std::ptr::null_mut();
        let rc = unsafe { sys::lbfgsb200_comm_create(id.as_ptr() as *const std::os::raw::c_char, rank, nranks, device, &mut handle) };
        if rc != 0 { bail!("lbfgsb200_comm_create failed with status {rc}") }
        Ok(Self { handle, rank, nranks })
    }
    /// true: the scalar exchange is fused into the reducing kernels (peer mailboxes over NVLink); false: ncclAllReduce.
    pub fn peer_mailboxes(&self) -> bool { unsafe { sys::lbfgsb200_comm_transport(self.handle) == 1 } }
}
impl Drop for Comm { fn drop(&mut self) { unsafe { sys::lbfgsb200_comm_destroy(self.handle) } } }

/// src/core.rs:221-250; `x` / `gx` are device pointers.
#[derive(Debug, Clone)]
pub struct Progress { pub x: *const f64, pub gx: *const f64, pub n: usize, pub fx: f64, pub xnorm: f64, pub gnorm: f64,
                      pub step: f64, pub niter: usize, pub neval: usize, pub ncall: usize }
/// src/core.rs:271-285
#[derive(Debug, Clone, Default)]
pub struct Report { pub fx: f64, pub xnorm: f64, pub gnorm: f64, pub neval: usize }

/// `default_progress()` (src/lib.rs:102-112): prints the iteration line of the reference and never cancels.
pub fn default_progress() -> impl FnMut(&Progress) -> bool {
    move |prgr| {
        println!("Iteration {}, Evaluation: {}", prgr.niter, prgr.neval);
        println!(" fx = {:-12.6} xnorm = {:-12.6}, gnorm = {:-12.6}, ls = {}, step = {}", prgr.fx, prgr.xnorm, prgr.gnorm, prgr.ncall, prgr.step);
        false
    }
}

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
pub struct Lbfgs { param: sys::lbfgsb200_param_t, fused_trial: bool, compact: Option<bool>, shard: Option<(*mut sys::lbfgsb200_comm_t, i64, i64)> }
impl Default for Lbfgs {
    fn default() -> Self {
        let mut p = std::mem::MaybeUninit::<sys::lbfgsb200_param_t>::zeroed();
        unsafe { sys::lbfgsb200_param_default(p.as_mut_ptr()); Self { param: p.assume_init(), fused_trial: true, compact: None, shard: None } }
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
    /// false: line-search trials as K1 + evaluate + K2 even if the objective offers probe + commit / a fused trial.
    pub fn with_fused_trial(mut self, on: bool) -> Self { self.fused_trial = on; self }
    /// true: the search direction from two passes over the ring (alpha_j / beta_j derived from inner products of the
    /// unmodified ring vectors, include/lbfgsb200.h: LBFGSB200_DIRECTION_COMPACT) instead of the reference's 2 * min(m, k)
    /// dependent trips; same element-wise operations, scalars equal up to rounding.  m <= 32.
    pub fn with_compact_direction(mut self, on: bool) -> Self { self.compact = Some(on); self }
    /// This rank's `x` is elements `[global_offset, global_offset + x.len())` of an `n_global` vector (one process per GPU).
    pub fn with_shard(mut self, comm: &Comm, n_global: usize, global_offset: usize) -> Self {
        self.shard = Some((comm.handle, n_global as i64, global_offset as i64)); self
    }

    /// The reference's exact call shape: `x` is a HOST slice (`minimize(&mut x, ..)`, src/lbfgs.rs:399).  One C-ABI call
    /// copies it to the device, solves there and copies the result back.
    pub fn minimize_host<E, G>(self, x: &mut [f64], device: i32, mut eval_fn: E, mut prgr_fn: G) -> Result<Report>
    where E: DeviceEvaluate, G: FnMut(&Progress) -> bool {
        let eval = eval_fn.raw().unwrap_or((Some(eval_tramp::<E>), &mut eval_fn as *mut E as *mut c_void));
        let (comm, n_global, goff) = self.shard.unwrap_or((std::ptr::null_mut(), x.len() as i64, 0));
        if !comm.is_null() { eval_fn.attach_comm(comm)?; }
        let ops = if self.fused_trial { eval_fn.fused_ops() } else { None };
        let mut rep = sys::lbfgsb200_report_t::default();
        // the solver is created inside the call: it takes the process-wide default
        if let Some(on) = self.compact {
            assert!(!on || self.param.m <= 32, "the compact direction supports m <= 32");
            unsafe { sys::lbfgsb200_set_default_direction(on as c_int); }
        }
        let st = unsafe { sys::lbfgsb200_minimize_host_ex(&self.param, x.as_mut_ptr(), x.len() as i64, n_global, goff, device, comm,
                                                          eval.0, eval.1, ops.as_ref().map_or(std::ptr::null(), |o| o as *const _),
                                                          Some(progress_tramp::<G>), &mut prgr_fn as *mut G as *mut c_void, &mut rep) };
        if self.compact.is_some() { unsafe { sys::lbfgsb200_set_default_direction(-1); } }
        if st < 0 { bail!("minimize failed with status {st}") }
        Ok(Report { fx: rep.fx, xnorm: rep.xnorm, gnorm: rep.gnorm, neval: rep.neval as usize })
    }

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
        let (comm, n_global, goff) = self.shard.unwrap_or((std::ptr::null_mut(), n as i64, 0));
        let rc = unsafe { sys::lbfgsb200_create(&self.param, n as i64, n_global, goff, device, std::ptr::null_mut(), comm, &mut solver) };
        if rc != 0 { bail!("lbfgsb200_create failed with status {rc} (no CUDA device? there is no CPU fallback)") }
        if let Some(on) = self.compact {
            let rc = unsafe { sys::lbfgsb200_set_direction(solver, on as c_int) };
            if rc != 0 { unsafe { sys::lbfgsb200_destroy(solver) }; bail!("lbfgsb200_set_direction failed with status {rc} (the compact direction needs m <= 32)") }
        }
        let eval = eval_fn.raw().unwrap_or((Some(eval_tramp::<E>), eval_fn as *mut E as *mut c_void));
        if !comm.is_null() { eval_fn.attach_comm(comm)?; }
        if self.fused_trial {
            if let Some(ops) = eval_fn.fused_ops() { unsafe { sys::lbfgsb200_set_fused_ops(solver, &ops); } }
        }
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
