//! Raw bindings to `include/lbfgsb200.h` (ABI version 3).  One item per C declaration, same order.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const LBFGSB200_ABI_VERSION: c_int = 3;

pub const LBFGSB200_OK_CONVERGED: c_int = 0;
pub const LBFGSB200_OK_MAX_ITERATIONS: c_int = 1;
pub const LBFGSB200_OK_MAX_EVALUATIONS: c_int = 2;
pub const LBFGSB200_OK_CANCELLED: c_int = 3;
pub const LBFGSB200_ERR_EVALUATE: c_int = -1;
pub const LBFGSB200_ERR_X_NOT_CHANGED: c_int = -2;
pub const LBFGSB200_ERR_G_NOT_CHANGED: c_int = -3;
pub const LBFGSB200_ERR_LINESEARCH: c_int = -4;
pub const LBFGSB200_ERR_INVALID_PARAM: c_int = -5;
pub const LBFGSB200_ERR_OWLQN_ZERO_DIRECTION: c_int = -6;
pub const LBFGSB200_ERR_INVALID_DNORM: c_int = -7;
pub const LBFGSB200_ERR_CUDA: c_int = -20;
pub const LBFGSB200_ERR_NCCL: c_int = -21;
pub const LBFGSB200_ERR_STATE: c_int = -22;
pub const LBFGSB200_ERR_UNSUPPORTED: c_int = -23;

pub const LBFGSB200_LS_MORETHUENTE: i64 = 0;
pub const LBFGSB200_LS_BACKTRACKING_ARMIJO: i64 = 1;
pub const LBFGSB200_LS_BACKTRACKING_WOLFE: i64 = 2;
pub const LBFGSB200_LS_BACKTRACKING_STRONG_WOLFE: i64 = 3;

pub const LBFGSB200_REDUCE_TREE: i64 = 0;
pub const LBFGSB200_REDUCE_SEQUENTIAL: i64 = 1;
pub const LBFGSB200_DIRECTION_TWO_LOOP: c_int = 0;
pub const LBFGSB200_DIRECTION_COMPACT: c_int = 1;

/// `lbfgsb200_param_t`: LbfgsParam + LineSearch + Orthantwise flattened (8-byte fields only).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct lbfgsb200_param_t {
    pub struct_size: i64,
    pub m: i64,
    pub epsilon: f64,
    pub past: i64,
    pub delta: f64,
    pub max_iterations: i64,
    pub max_evaluations: i64,
    pub ls_algorithm: i64,
    pub ls_ftol: f64,
    pub ls_gtol: f64,
    pub ls_xtol: f64,
    pub ls_min_step: f64,
    pub ls_max_step: f64,
    pub ls_max_linesearch: i64,
    pub ls_gradient_only: i64,
    pub orthantwise: i64,
    pub owl_c: f64,
    pub owl_start: i64,
    pub owl_end: i64,
    pub initial_inverse_hessian: f64,
    pub max_step_size: f64,
    pub damping: i64,
    pub constrain_step_size: i64,
    pub reduction: i64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct lbfgsb200_progress_t {
    pub x_dev: *const f64,
    pub gx_dev: *const f64,
    pub n_local: i64,
    pub n_global: i64,
    pub fx: f64,
    pub xnorm: f64,
    pub gnorm: f64,
    pub step: f64,
    pub niter: i64,
    pub neval: i64,
    pub ncall: i64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct lbfgsb200_report_t {
    pub fx: f64,
    pub xnorm: f64,
    pub gnorm: f64,
    pub neval: i64,
    pub niter: i64,
    pub last_ls_error: i64,
    pub status: i64,
}

pub const LBFGSB200_K_COUNT: usize = 15;
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct lbfgsb200_profile_t {
    pub launches: [i64; LBFGSB200_K_COUNT],
    pub bytes: [f64; LBFGSB200_K_COUNT],
    pub ms: [f64; LBFGSB200_K_COUNT],
    pub host_syncs: i64,
    pub allreduces: i64,
}

#[repr(C)] pub struct lbfgsb200_solver_t { _p: [u8; 0] }
#[repr(C)] pub struct lbfgsb200_comm_t { _p: [u8; 0] }
#[repr(C)] pub struct lbfgsb200_objective_t { _p: [u8; 0] }
#[repr(C)] pub struct lbfgsb200_linesearch_t { _p: [u8; 0] }

pub type lbfgsb200_eval_fn = Option<unsafe extern "C" fn(user: *mut c_void, x_dev: *const f64, g_dev: *mut f64,
    n_local: i64, stream: *mut c_void, fx_dev: *mut f64) -> c_int>;
pub type lbfgsb200_trial_eval_fn = Option<unsafe extern "C" fn(user: *mut c_void, xp_dev: *const f64, d_dev: *const f64,
    step: f64, x_dev: *mut f64, g_dev: *mut f64, n_local: i64, stream: *mut c_void, out_dev: *mut f64) -> c_int>;
pub type lbfgsb200_probe_fn = Option<unsafe extern "C" fn(user: *mut c_void, xp_dev: *const f64, d_dev: *const f64, step: f64,
    step_dev: *const f64, n_local: i64, stream: *mut c_void, out_dev: *mut f64) -> c_int>;
pub type lbfgsb200_commit_fn = Option<unsafe extern "C" fn(user: *mut c_void, xp_dev: *const f64, d_dev: *const f64,
    gp_dev: *const f64, step: f64, bs_scale: f64, x_dev: *mut f64, g_dev: *mut f64, s_dev: *mut f64, y_dev: *mut f64,
    n_local: i64, stream: *mut c_void, out_dev: *mut f64) -> c_int>;

/// The commit fused with pass A of the compact search direction (include/lbfgsb200.h: lbfgsb200_commit_gram_fn).
pub type lbfgsb200_commit_gram_fn = Option<unsafe extern "C" fn(user: *mut c_void, xp_dev: *const f64, d_dev: *const f64,
    gp_dev: *const f64, step: f64, bs_scale: f64, x_dev: *mut f64, g_dev: *mut f64, s_dev: *mut f64, y_dev: *mut f64,
    s_old_dev: *const *const f64, y_old_dev: *const *const f64, n_old: c_int, n_local: i64, stream: *mut c_void,
    out_dev: *mut f64, gram_out_dev: *mut f64, newdot_out_dev: *mut f64) -> c_int>;

/// Several write-free trials in one pass (include/lbfgsb200.h: lbfgsb200_probe_multi_fn).
pub type lbfgsb200_probe_multi_fn = Option<unsafe extern "C" fn(user: *mut c_void, xp_dev: *const f64, d_dev: *const f64,
    steps: *const f64, step0_dev: *const f64, k: c_int, n_local: i64, stream: *mut c_void, out_dev: *mut f64) -> c_int>;

pub const LBFGSB200_FUSED_OPS_SIZE_V1: i64 = 48;
pub const LBFGSB200_FUSED_OPS_SIZE_V2: i64 = 56;
pub const LBFGSB200_FUSED_SUMS_OVER_RANKS: i64 = 1;
pub const LBFGSB200_FUSED_COMMIT_SKIPS_GP: i64 = 2;
/// `lbfgsb200_fused_ops_t`: what an objective offers beyond evaluate.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct lbfgsb200_fused_ops_t {
    pub struct_size: i64,
    pub trial: lbfgsb200_trial_eval_fn,
    pub probe: lbfgsb200_probe_fn,
    pub commit: lbfgsb200_commit_fn,
    pub user: *mut c_void,
    pub flags: i64,
    pub commit_gram: lbfgsb200_commit_gram_fn,
    pub probe_multi: lbfgsb200_probe_multi_fn,
}
impl Default for lbfgsb200_fused_ops_t {
    fn default() -> Self {
        Self { struct_size: std::mem::size_of::<Self>() as i64, trial: None, probe: None, commit: None, user: std::ptr::null_mut(), flags: 0,
               commit_gram: None, probe_multi: None }
    }
}

pub const LBFGSB200_GLM_PATH_TWO_PASS: c_int = 1;
pub const LBFGSB200_GLM_PATH_FUSED: c_int = 2;
pub const LBFGSB200_GLM_PATH_FUSED_ODD: c_int = 3;
pub const LBFGSB200_GLM_PATH_FUSED_CLUSTER: c_int = 4;

pub type lbfgsb200_progress_fn = Option<unsafe extern "C" fn(user: *mut c_void, progress: *const lbfgsb200_progress_t) -> c_int>;

extern "C" {
    pub fn lbfgsb200_abi_version() -> c_int;
    pub fn lbfgsb200_param_default(param: *mut lbfgsb200_param_t);

    pub fn lbfgsb200_comm_unique_id(id: *mut c_char) -> c_int;
    pub fn lbfgsb200_comm_create(id: *const c_char, rank: c_int, nranks: c_int, device: c_int, out: *mut *mut lbfgsb200_comm_t) -> c_int;
    pub fn lbfgsb200_comm_destroy(comm: *mut lbfgsb200_comm_t);
    pub fn lbfgsb200_comm_transport(comm: *const lbfgsb200_comm_t) -> c_int;
    pub fn lbfgsb200_comm_allreduce_sum(comm: *mut lbfgsb200_comm_t, buf_dev: *mut f64, count: c_int, stream: *mut c_void) -> c_int;

    pub fn lbfgsb200_create(param: *const lbfgsb200_param_t, n_local: i64, n_global: i64, global_offset: i64, device: c_int,
                            stream: *mut c_void, comm: *mut lbfgsb200_comm_t, out: *mut *mut lbfgsb200_solver_t) -> c_int;
    pub fn lbfgsb200_destroy(solver: *mut lbfgsb200_solver_t);
    pub fn lbfgsb200_last_error(solver: *const lbfgsb200_solver_t) -> *const c_char;
    pub fn lbfgsb200_minimize(solver: *mut lbfgsb200_solver_t, x_dev: *mut f64, eval: lbfgsb200_eval_fn, eval_user: *mut c_void,
                              progress: lbfgsb200_progress_fn, progress_user: *mut c_void, report: *mut lbfgsb200_report_t) -> c_int;
    pub fn lbfgsb200_set_trial_evaluate(solver: *mut lbfgsb200_solver_t, f: lbfgsb200_trial_eval_fn, user: *mut c_void) -> c_int;
    pub fn lbfgsb200_set_fused_ops(solver: *mut lbfgsb200_solver_t, ops: *const lbfgsb200_fused_ops_t) -> c_int;
    pub fn lbfgsb200_set_direction(solver: *mut lbfgsb200_solver_t, mode: c_int) -> c_int;
    pub fn lbfgsb200_get_direction(solver: *const lbfgsb200_solver_t) -> c_int;
    pub fn lbfgsb200_set_default_direction(mode: c_int) -> c_int;
    pub fn lbfgsb200_build(solver: *mut lbfgsb200_solver_t, x_dev: *mut f64, eval: lbfgsb200_eval_fn, eval_user: *mut c_void) -> c_int;
    pub fn lbfgsb200_is_converged(solver: *mut lbfgsb200_solver_t, stop_status: *mut c_int) -> c_int;
    pub fn lbfgsb200_propagate(solver: *mut lbfgsb200_solver_t, progress_out: *mut lbfgsb200_progress_t) -> c_int;
    pub fn lbfgsb200_report(solver: *mut lbfgsb200_solver_t, report_out: *mut lbfgsb200_report_t) -> c_int;
    pub fn lbfgsb200_finish(solver: *mut lbfgsb200_solver_t) -> c_int;
    pub fn lbfgsb200_x(solver: *const lbfgsb200_solver_t) -> *const f64;
    pub fn lbfgsb200_gx(solver: *const lbfgsb200_solver_t) -> *const f64;
    pub fn lbfgsb200_direction(solver: *const lbfgsb200_solver_t) -> *const f64;
    pub fn lbfgsb200_minimize_host(param: *const lbfgsb200_param_t, x_host: *mut f64, n: i64, device: c_int, eval: lbfgsb200_eval_fn,
                                   eval_user: *mut c_void, progress: lbfgsb200_progress_fn, progress_user: *mut c_void,
                                   report: *mut lbfgsb200_report_t) -> c_int;

    pub fn lbfgsb200_minimize_host_ex(param: *const lbfgsb200_param_t, x_host: *mut f64, n_local: i64, n_global: i64, global_offset: i64,
                                      device: c_int, comm: *mut lbfgsb200_comm_t, eval: lbfgsb200_eval_fn, eval_user: *mut c_void,
                                      fused: *const lbfgsb200_fused_ops_t, progress: lbfgsb200_progress_fn,
                                      progress_user: *mut c_void, report: *mut lbfgsb200_report_t) -> c_int;
    pub fn lbfgsb200_profile_enable(solver: *mut lbfgsb200_solver_t, timing: c_int) -> c_int;
    pub fn lbfgsb200_profile_get(solver: *mut lbfgsb200_solver_t, out: *mut lbfgsb200_profile_t) -> c_int;
    pub fn lbfgsb200_profile_reset(solver: *mut lbfgsb200_solver_t) -> c_int;

    // LbfgsMath on device pointers, src/math.rs:31-82
    pub fn lbfgsb200_vecadd(y_dev: *mut f64, x_dev: *const f64, c: f64, n: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_vecdot(x_dev: *const f64, y_dev: *const f64, n: i64, stream: *mut c_void, out_host: *mut f64) -> c_int;
    pub fn lbfgsb200_vecscale(y_dev: *mut f64, c: f64, n: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_veccpy(y_dev: *mut f64, x_dev: *const f64, n: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_vecncpy(y_dev: *mut f64, x_dev: *const f64, n: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_vecdiff(z_dev: *mut f64, x_dev: *const f64, y_dev: *const f64, n: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_vec2norm(x_dev: *const f64, n: i64, stream: *mut c_void, out_host: *mut f64) -> c_int;
    pub fn lbfgsb200_vec2norminv(x_dev: *const f64, n: i64, stream: *mut c_void, out_host: *mut f64) -> c_int;

    pub fn lbfgsb200_dots3(g_dev: *const f64, d_dev: *const f64, x_dev: *const f64, n: i64, stream: *mut c_void, out_host: *mut f64) -> c_int;
    pub fn lbfgsb200_trial_step(x_dev: *mut f64, xp_dev: *const f64, d_dev: *const f64, step: f64, n: i64, wp_dev: *const i8,
                                start: i64, end: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_owl_pseudo_gradient(pg_dev: *mut f64, x_dev: *const f64, g_dev: *const f64, n: i64, c: f64, start: i64, end: i64,
                                         stream: *mut c_void, out_host: *mut f64) -> c_int;
    pub fn lbfgsb200_owl_orthant(wp_dev: *mut i8, xp_dev: *const f64, pg_dev: *const f64, n: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_owl_constrain_direction(d_dev: *mut f64, pg_dev: *const f64, n: i64, start: i64, end: i64, stream: *mut c_void,
                                             out_host: *mut f64) -> c_int;

    // the update chain's kernels, one call each (scalars by value)
    pub fn lbfgsb200_init_direction(d_dev: *mut f64, g_dev: *const f64, n: i64, stream: *mut c_void, out_host: *mut f64) -> c_int;
    pub fn lbfgsb200_history_update(s_dev: *mut f64, y_dev: *mut f64, x_dev: *const f64, xp_dev: *const f64, g_dev: *const f64,
                                    gp_dev: *const f64, pg_dev: *const f64, n: i64, step: f64, damping: c_int, stream: *mut c_void,
                                    out_host: *mut f64) -> c_int;
    pub fn lbfgsb200_damp_y(y_dev: *mut f64, gp_dev: *const f64, n: i64, step: f64, ys: f64, sbs: f64, stream: *mut c_void,
                            applied_host: *mut c_int) -> c_int;
    pub fn lbfgsb200_two_loop_backward_step(q_dev: *mut f64, g_first_dev: *const f64, y_j_dev: *const f64, s_next_dev: *const f64,
                                            n: i64, sq: f64, ys_j: f64, gamma: f64, stream: *mut c_void, out_host: *mut f64) -> c_int;
    pub fn lbfgsb200_two_loop_forward_step(r_dev: *mut f64, s_j_dev: *const f64, y_next_dev: *const f64, g_last_dev: *const f64,
                                           n: i64, yr: f64, ys_j: f64, alpha_j: f64, owl: c_int, owl_start: i64, owl_end: i64,
                                           stream: *mut c_void, out_host: *mut f64) -> c_int;

    pub fn lbfgsb200_objective_rosenbrock(device: c_int, out: *mut *mut lbfgsb200_objective_t) -> c_int;
    pub fn lbfgsb200_objective_booth(device: c_int, out: *mut *mut lbfgsb200_objective_t) -> c_int;
    pub fn lbfgsb200_objective_glm(device: c_int, kind: c_int, x_dev: *const f64, y_dev: *const f64, nrow: i64, ncol: i64,
                                   out: *mut *mut lbfgsb200_objective_t) -> c_int;
    pub fn lbfgsb200_objective_lennard_jones(device: c_int, epsilon: f64, sigma: f64, out: *mut *mut lbfgsb200_objective_t) -> c_int;
    pub fn lbfgsb200_objective_set_reduction(objective: *mut lbfgsb200_objective_t, reduction: c_int) -> c_int;
    pub fn lbfgsb200_objective_set_shard(objective: *mut lbfgsb200_objective_t, comm: *mut lbfgsb200_comm_t, shard_offsets: *const i64) -> c_int;
    pub fn lbfgsb200_objective_destroy(objective: *mut lbfgsb200_objective_t);
    pub fn lbfgsb200_objective_eval(objective: *mut c_void, x_dev: *const f64, g_dev: *mut f64, n_local: i64, stream: *mut c_void,
                                    fx_dev: *mut f64) -> c_int;
    pub fn lbfgsb200_objective_trial_eval(objective: *mut c_void, xp_dev: *const f64, d_dev: *const f64, step: f64, x_dev: *mut f64,
                                          g_dev: *mut f64, n_local: i64, stream: *mut c_void, out_dev: *mut f64) -> c_int;
    pub fn lbfgsb200_objective_has_trial_eval(objective: *const lbfgsb200_objective_t) -> c_int;
    pub fn lbfgsb200_objective_probe(objective: *mut c_void, xp_dev: *const f64, d_dev: *const f64, step: f64, step_dev: *const f64,
                                     n_local: i64, stream: *mut c_void, out_dev: *mut f64) -> c_int;
    pub fn lbfgsb200_objective_commit(objective: *mut c_void, xp_dev: *const f64, d_dev: *const f64, gp_dev: *const f64, step: f64,
                                      bs_scale: f64, x_dev: *mut f64, g_dev: *mut f64, s_dev: *mut f64, y_dev: *mut f64,
                                      n_local: i64, stream: *mut c_void, out_dev: *mut f64) -> c_int;
    pub fn lbfgsb200_objective_probe_multi(objective: *mut c_void, xp_dev: *const f64, d_dev: *const f64, steps: *const f64,
                                           step0_dev: *const f64, k: c_int, n_local: i64, stream: *mut c_void,
                                           out_dev: *mut f64) -> c_int;
    pub fn lbfgsb200_objective_commit_gram(objective: *mut c_void, xp_dev: *const f64, d_dev: *const f64, gp_dev: *const f64, step: f64,
                                           bs_scale: f64, x_dev: *mut f64, g_dev: *mut f64, s_dev: *mut f64, y_dev: *mut f64,
                                           s_old_dev: *const *const f64, y_old_dev: *const *const f64, n_old: c_int, n_local: i64,
                                           stream: *mut c_void, out_dev: *mut f64, gram_out_dev: *mut f64,
                                           newdot_out_dev: *mut f64) -> c_int;
    pub fn lbfgsb200_objective_fused_ops(objective: *mut lbfgsb200_objective_t, out: *mut lbfgsb200_fused_ops_t) -> c_int;
    pub fn lbfgsb200_objective_last_path(objective: *const lbfgsb200_objective_t) -> c_int;
    pub fn lbfgsb200_objective_set_lj_fast(objective: *mut lbfgsb200_objective_t, fast: c_int) -> c_int;

    pub fn lbfgsb200_linesearch_begin(param: *const lbfgsb200_param_t, orthantwise: c_int, finit: f64, dginit: f64, step: f64)
        -> *mut lbfgsb200_linesearch_t;
    pub fn lbfgsb200_linesearch_next(ls: *mut lbfgsb200_linesearch_t, step_out: *mut f64) -> c_int;
    pub fn lbfgsb200_linesearch_predict(ls: *const lbfgsb200_linesearch_t, steps_out: *mut f64, kmax: c_int) -> c_int;
    pub fn lbfgsb200_linesearch_feed(ls: *mut lbfgsb200_linesearch_t, eval_ok: c_int, f: f64, dg: f64);
    pub fn lbfgsb200_linesearch_result(ls: *mut lbfgsb200_linesearch_t, ncall: *mut i64, step: *mut f64) -> c_int;
    pub fn lbfgsb200_linesearch_end(ls: *mut lbfgsb200_linesearch_t);

    pub fn lbfgsb200_device_count() -> c_int;
    pub fn lbfgsb200_device_alloc(device: c_int, bytes: i64, out_dev: *mut *mut c_void) -> c_int;
    pub fn lbfgsb200_device_free(dev: *mut c_void) -> c_int;
    pub fn lbfgsb200_copy_h2d(dst_dev: *mut c_void, src_host: *const c_void, bytes: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_copy_d2h(dst_host: *mut c_void, src_dev: *const c_void, bytes: i64, stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_stream_synchronize(stream: *mut c_void) -> c_int;
    pub fn lbfgsb200_trim_pool(device: c_int) -> c_int;
}
