/* lbfgsb200.h — C ABI of the B200-native L-BFGS / OWL-QN solver (liblbfgsb200.so).
 *
 * This is the drop-in boundary for the hot path of ybyygu/rust-lbfgs (`liblbfgs` 0.2.0).  The
 * reference has no FFI of its own; its seams are `trait LbfgsMath<f64>` (src/math.rs:4-29), the
 * builder `Lbfgs::with_*` / `minimize` / `build` / `propagate` (src/lbfgs.rs:185-566) and the two
 * user closures `E: FnMut(&[f64], &mut [f64]) -> Result<f64>` (src/core.rs:10-13) and
 * `G: FnMut(&Progress) -> bool` (src/lbfgs.rs:402).  Each entry point below names the reference
 * item it replaces.  A Rust `-sys` crate (rust_lbfgs_b200/rust/), the C++ builder
 * (rust_lbfgs_b200/cxx/lbfgsb200.hpp) and the Python ctypes mirror (rust_lbfgs_b200/api.py) all
 * bind exactly these symbols; see INTEGRATION.md.
 *
 * Conventions
 *   - plain C: pointers, sizes, PODs with 8-byte fields only; no CUDA or torch types.  A
 *     `stream` argument is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - every vector (x, g, d, xp, gp, pg, the m-deep s/y ring) lives in HBM for the whole solve;
 *     `x_dev` is caller-owned device memory, 16-byte aligned, length n_local; everything else is
 *     owned by the solver handle.
 *   - all arithmetic is IEEE f64; element-wise results are bit-identical to the reference's
 *     (kernels are built with -fmad=false), reductions are deterministic two-level tree sums.
 *   - one handle per GPU / rank; a handle is not thread-safe (same as the reference).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     LBFGSB200_ERR_CUDA.
 */
#ifndef LBFGSB200_H
#define LBFGSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBFGSB200_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------------------------- */
enum {
    LBFGSB200_OK_CONVERGED = 0,             /* gnorm / max(1, xnorm) <= epsilon   src/lbfgs.rs:714-722 */
    LBFGSB200_OK_MAX_ITERATIONS = 1,        /* src/lbfgs.rs:726-735 */
    LBFGSB200_OK_MAX_EVALUATIONS = 2,       /* src/lbfgs.rs:739-748 */
    LBFGSB200_OK_CANCELLED = 3,             /* progress callback returned non-zero  src/lbfgs.rs:412-416 */
    LBFGSB200_OK = 0,
    LBFGSB200_ERR_EVALUATE = -1,            /* Err from evaluate at the initial point  src/lbfgs.rs:454 */
    LBFGSB200_ERR_X_NOT_CHANGED = -2,       /* "x not changed with step ..."  src/lbfgs.rs:645-646 */
    LBFGSB200_ERR_G_NOT_CHANGED = -3,       /* "gx not changed"  src/lbfgs.rs:655 */
    LBFGSB200_ERR_LINESEARCH = -4,          /* Err out of LineSearch::find itself  src/line.rs:198-201,208 */
    LBFGSB200_ERR_INVALID_PARAM = -5,       /* the reference's assert!/panic on parameters */
    LBFGSB200_ERR_OWLQN_ZERO_DIRECTION = -6,/* assert_ne!(d.vec2norm(), 0.0)  src/orthantwise.rs:160 */
    LBFGSB200_ERR_INVALID_DNORM = -7,       /* ensure!(dnorm.is_sign_positive())  src/lbfgs.rs:544 */
    LBFGSB200_ERR_CUDA = -20,               /* CUDA runtime failure, or no CUDA device */
    LBFGSB200_ERR_NCCL = -21,               /* NCCL failure, or libnccl.so.2 not loadable */
    LBFGSB200_ERR_STATE = -22,              /* call order violated (e.g. propagate before build) */
    LBFGSB200_ERR_UNSUPPORTED = -23         /* the objective has no implementation of the requested entry */
};

/* line-search algorithms  src/line.rs:39-80 */
enum {
    LBFGSB200_LS_MORETHUENTE = 0,
    LBFGSB200_LS_BACKTRACKING_ARMIJO = 1,
    LBFGSB200_LS_BACKTRACKING_WOLFE = 2,
    LBFGSB200_LS_BACKTRACKING_STRONG_WOLFE = 3
};

/* reduction order of every dot product / norm / objective sum (param.reduction) */
enum {
    LBFGSB200_REDUCE_TREE = 0,        /* production: deterministic two-level tree (warp shuffle, CTA, last-CTA) */
    LBFGSB200_REDUCE_SEQUENTIAL = 1   /* validation: one thread, the reference's left-to-right fold
                                         (`iter().sum()`, src/math.rs:40-42): with it a whole solve is
                                         bit-identical to the reference's CPU arithmetic.  Slow by design. */
};

/* swallowed line-search errors (src/line.rs:213-220 prints and reverts), report.last_ls_error */
enum {
    LBFGSB200_LS_ERR_NONE = 0,
    LBFGSB200_LS_ERR_EVALUATE = 1,
    LBFGSB200_LS_ERR_ROUNDING = 2,          /* src/line.rs:292-298 */
    LBFGSB200_LS_ERR_XTOL = 3,              /* src/line.rs:300-302 */
    LBFGSB200_LS_ERR_MAX_STEP = 4,          /* src/line.rs:305-308,171-174 */
    LBFGSB200_LS_ERR_MIN_STEP = 5,          /* src/line.rs:310-313,167-170 */
    LBFGSB200_LS_ERR_OUT_OF_INTERVAL = 6,   /* src/line.rs:474-476 */
    LBFGSB200_LS_ERR_INCREASE_GRADIENT = 7, /* src/line.rs:477-479 */
    LBFGSB200_LS_ERR_INCORRECT_TMINMAX = 8  /* src/line.rs:480-483 */
};

/* ---- parameters --------------------------------------------------------------------------- */
/* LbfgsParam (src/lbfgs.rs:72-154) + LineSearch (src/line.rs:91-148) + Orthantwise
 * (src/orthantwise.rs:19-45), flattened.  `struct_size` must be sizeof(lbfgsb200_param_t). */
typedef struct lbfgsb200_param {
    int64_t struct_size;
    int64_t m;                      /* history depth; default 6 (no setter in the reference: src/lbfgs.rs:163,182) */
    double  epsilon;                /* 1e-5 */
    int64_t past;                   /* 0; stored, never consumed (dead code src/lbfgs.rs:766-787) */
    double  delta;                  /* 1e-5; stored, never consumed */
    int64_t max_iterations;         /* 0 = unlimited */
    int64_t max_evaluations;        /* 0 = unlimited */
    int64_t ls_algorithm;           /* LBFGSB200_LS_* */
    double  ls_ftol;                /* 1e-4 */
    double  ls_gtol;                /* 0.9 */
    double  ls_xtol;                /* f64::EPSILON */
    double  ls_min_step;            /* 1e-20 */
    double  ls_max_step;            /* 1e+20 */
    int64_t ls_max_linesearch;      /* 20 */
    int64_t ls_gradient_only;       /* 0 */
    int64_t orthantwise;            /* 0 = None */
    double  owl_c;                  /* 1.0 */
    int64_t owl_start;              /* 0   (global index) */
    int64_t owl_end;                /* < 0 = None => n_global */
    double  initial_inverse_hessian;/* 1.0 */
    double  max_step_size;          /* 1.0 */
    int64_t damping;                /* 0 */
    int64_t constrain_step_size;    /* 1 (no setter in the reference: src/lbfgs.rs:153,174) */
    int64_t reduction;              /* LBFGSB200_REDUCE_*; 0 (extension, not in the reference) */
} lbfgsb200_param_t;

/* Lbfgs::default()  src/lbfgs.rs:156-177, src/line.rs:150-163, src/orthantwise.rs:47-55 */
void lbfgsb200_param_default(lbfgsb200_param_t *param);

/* ---- callbacks ---------------------------------------------------------------------------- */
/* Device-resident evaluate: replaces `E: FnMut(&[f64], &mut [f64]) -> Result<f64>`
 * (src/core.rs:10-13,120).  Must enqueue, on `stream`, work that writes the gradient of this
 * rank's shard to g_dev[0..n_local) and this rank's PARTIAL objective value to *fx_dev (device
 * memory; the solver sums the partials over ranks).  Must not synchronise.  Non-zero = Err. */
typedef int (*lbfgsb200_eval_fn)(void *user, const double *x_dev, double *g_dev, int64_t n_local,
                                 void *stream, double *fx_dev);

/* Optional FUSED line-search trial (north_star: "line-search trial points x+alpha*d come from fused
 * multi-reductions").  One call replaces, for one trial of src/line.rs:283-288 / :741-742,
 *   take_line_step (src/core.rs:155-164)  x = xp + step*d
 *   evaluate       (src/core.rs:119-132)  g = grad f(x), f
 *   dg_unchecked   (src/core.rs:114-116)  g.d          and gnorm / xnorm (src/core.rs:183-194)
 * by ONE pass that reads xp and d and writes x and g (2R 2W instead of 5R 3W over three kernels):
 *   out_dev[0] = this rank's partial f(x), out_dev[1] = g.d, out_dev[2] = g.g, out_dev[3] = x.x  (partials)
 * Element-wise arithmetic must be that of the unfused path (x = xp + step*d without FMA).  Same rules as
 * lbfgsb200_eval_fn: enqueue on `stream`, do not synchronise, non-zero = Err.  Not used for OWL-QN. */
typedef int (*lbfgsb200_trial_eval_fn)(void *user, const double *xp_dev, const double *d_dev, double step,
                                       double *x_dev, double *g_dev, int64_t n_local, void *stream,
                                       double *out_dev);

/* PROBE + COMMIT: the fused line search without wasted writes.  A line search evaluates t trial points and
 * keeps one; only {f, g.d} of the rejected ones are ever used (src/line.rs:283-320, :741-760).  So a trial is a
 * PROBE that reads xp and d and writes nothing (2R instead of the fused trial's 2R 2W):
 *   out_dev[0..3] = partial f(xp + step*d), g.d, g.g, x.x          (same sums as lbfgsb200_trial_eval_fn)
 * step_dev non-NULL: the trial step is *step_dev — device memory written by an earlier kernel on the stream — and
 * `step` is ignored.  Because a probe writes nothing it can be launched SPECULATIVELY: the solver enqueues the next
 * iteration's first trial right behind the two-loop recursion (its step, min(max_step_size, |d|) / |d|, is formed
 * on the device), so that trial's result arrives with the update's own scalars in one host synchronisation.
 * and, once the search has accepted a step, ONE COMMIT pass materialises the point and performs
 * IterationData::update's vector work (src/lbfgs.rs:640-656, :670-673) on the way:
 *   x = xp + step*d;  g = grad f(x);  s = x - xp;  y = g - gp
 *   out_dev[0..4] = partial s.s, y.s, y.y, s.(-g), s.(gp*bs_scale)      (bs_scale = -step_returned, for damping)
 * 3R 4W; per iteration 2t + 7 passes replace the fused trial's 4t + 6.  gp_dev is, by contract, the gradient at
 * xp_dev as THIS objective computed it (its evaluate, fused trial or previous commit wrote it); an objective whose
 * gradient is element-local may therefore recompute it from xp instead of reading it (the built-in Rosenbrock
 * does: 2R 4W, 2t + 6 passes per iteration).  Element-wise arithmetic must be that of
 * the unfused kernels (no FMA), so that all three paths produce the same bits.  Same rules as lbfgsb200_eval_fn:
 * enqueue on `stream`, do not synchronise, non-zero = Err.  Not used for OWL-QN. */
typedef int (*lbfgsb200_probe_fn)(void *user, const double *xp_dev, const double *d_dev, double step,
                                  const double *step_dev, int64_t n_local, void *stream, double *out_dev);
typedef int (*lbfgsb200_commit_fn)(void *user, const double *xp_dev, const double *d_dev, const double *gp_dev,
                                   double step, double bs_scale, double *x_dev, double *g_dev, double *s_dev,
                                   double *y_dev, int64_t n_local, void *stream, double *out_dev);

/* Optional: the commit fused with pass A of the compact search direction (lbfgsb200_set_direction).  Everything the
 * commit does, plus — from the registers that hold the new pair s = x - xp, y = g - gp and the new gradient g — the
 * inner products with n_old <= 5 older ring pairs (HOST arrays of device pointers s_old_dev[k], y_old_dev[k]):
 *   gram_out_dev[5 k + 0..4] = { s_k.d0, y_k.d0, s.y_k, s_k.y, y.y_k },  newdot_out_dev[0..1] = { y.d0, y.y },  d0 = -g
 * (ALWAYS partials of this rank, also in out_dev and also when the other entries carry LBFGSB200_FUSED_SUMS_OVER_RANKS:
 * the solver sums them over the ranks with one all-reduce.  LBFGSB200_ERR_UNSUPPORTED makes the solver run commit +
 * pass A separately, as it does with Powell damping, which rewrites y after the commit, and for OWL-QN). */
typedef int (*lbfgsb200_commit_gram_fn)(void *user, const double *xp_dev, const double *d_dev, const double *gp_dev,
                                        double step, double bs_scale, double *x_dev, double *g_dev, double *s_dev,
                                        double *y_dev, const double *const *s_old_dev, const double *const *y_old_dev,
                                        int n_old, int64_t n_local, void *stream, double *out_dev, double *gram_out_dev,
                                        double *newdot_out_dev);

/* Optional: k <= 6 write-free trials in ONE pass over xp and d.  steps[0 .. k) are host values; with step0_dev != NULL
 * they are instead the More-Thuente extrapolation chain s_0 = *step0_dev, s_{j+1} = s_j + 4 (s_j - s_{j-1}), s_{-1} = 0
 * (src/line.rs:266), formed on the device with exactly that expression.  out_dev[4 j + 0..3] = { f, g.d, g.g, x.x } at
 * xp + s_j d, each with the bits lbfgsb200_probe_fn would give for that step; out_dev[4 k + j] = s_j.  The solver asks
 * for the steps it EXPECTS the search to take and uses a result only if the search then asks for exactly that step. */
typedef int (*lbfgsb200_probe_multi_fn)(void *user, const double *xp_dev, const double *d_dev, const double *steps,
                                        const double *step0_dev, int k, int64_t n_local, void *stream, double *out_dev);

/* What an objective offers beyond lbfgsb200_eval_fn.  Unused entries are NULL.  probe needs commit. */
#define LBFGSB200_FUSED_SUMS_OVER_RANKS 1  /* the callbacks leave sums over ALL ranks in out_dev (the built-in
                                              objectives do, in their kernels' epilogue, once
                                              lbfgsb200_objective_set_shard gave them the communicator); such a
                                              callback must fail on every rank or on none */
#define LBFGSB200_FUSED_COMMIT_SKIPS_GP 2   /* commit recomputes gp from xp (accounting: 2R 4W instead of 3R 4W) */
typedef struct lbfgsb200_fused_ops {
    int64_t struct_size;                /* sizeof(lbfgsb200_fused_ops_t) */
    lbfgsb200_trial_eval_fn trial;      /* one-pass trial that writes x and g */
    lbfgsb200_probe_fn probe;           /* write-free trial */
    lbfgsb200_commit_fn commit;         /* accepted point + history update */
    void *user;
    int64_t flags;                      /* LBFGSB200_FUSED_* */
    lbfgsb200_commit_gram_fn commit_gram;  /* commit + pass A of the compact direction (struct_size tells whether the
                                              caller's struct has this field: LBFGSB200_FUSED_OPS_SIZE_V1 = without) */
    lbfgsb200_probe_multi_fn probe_multi;  /* several trials per pass (LBFGSB200_FUSED_OPS_SIZE_V2 = without) */
} lbfgsb200_fused_ops_t;
#define LBFGSB200_FUSED_OPS_SIZE_V1 48
#define LBFGSB200_FUSED_OPS_SIZE_V2 56

/* Progress  src/core.rs:221-250; x/gx are device pointers to this rank's shard */
typedef struct lbfgsb200_progress {
    const double *x_dev;
    const double *gx_dev;
    int64_t n_local;
    int64_t n_global;
    double  fx;
    double  xnorm;
    double  gnorm;
    double  step;
    int64_t niter;
    int64_t neval;
    int64_t ncall;
} lbfgsb200_progress_t;

/* replaces `G: FnMut(&Progress) -> bool`; non-zero cancels  src/lbfgs.rs:402,412-416 */
typedef int (*lbfgsb200_progress_fn)(void *user, const lbfgsb200_progress_t *progress);

/* Report  src/core.rs:271-285 (+ diagnostics) */
typedef struct lbfgsb200_report {
    double  fx;
    double  xnorm;
    double  gnorm;
    int64_t neval;
    int64_t niter;                  /* number of propagate() calls made */
    int64_t last_ls_error;          /* LBFGSB200_LS_ERR_* of the last swallowed line-search failure */
    int64_t status;                 /* status of the last minimize() */
} lbfgsb200_report_t;

/* ---- multi-GPU communicator (one NCCL rank per process/GPU) ------------------------------- */
typedef struct lbfgsb200_comm lbfgsb200_comm_t;
#define LBFGSB200_UNIQUE_ID_BYTES 128
/* rank 0 creates the id and ships it to the other ranks (torch.distributed / MPI / a file) */
int  lbfgsb200_comm_unique_id(char id[LBFGSB200_UNIQUE_ID_BYTES]);
int  lbfgsb200_comm_create(const char id[LBFGSB200_UNIQUE_ID_BYTES], int rank, int nranks, int device,
                           lbfgsb200_comm_t **out);
void lbfgsb200_comm_destroy(lbfgsb200_comm_t *comm);
/* How the solver sums its scalars over the ranks: 1 = peer mailboxes (CUDA IPC over NVLink; the exchange is fused
 * into the reducing kernels' epilogue, no collective call), 0 = ncclAllReduce per step (fallback, or
 * LBFGSB200_PEER_REDUCE=0). */
int  lbfgsb200_comm_transport(const lbfgsb200_comm_t *comm);
/* in-place sum of `count` doubles in device memory over all ranks (ncclAllReduce, ncclDouble, ncclSum) */
int  lbfgsb200_comm_allreduce_sum(lbfgsb200_comm_t *comm, double *buf_dev, int count, void *stream);

/* ---- solver lifecycle --------------------------------------------------------------------- */
typedef struct lbfgsb200_solver lbfgsb200_solver_t;

/* Problem::new (src/core.rs:59-75) + the ring allocation of build (src/lbfgs.rs:449): allocates
 * g/gp/x'/d (+pg, wp for OWL-QN) and the 2m history vectors in HBM.
 *   n_local        this rank's shard length;  n_global  total length (== n_local on one GPU)
 *   global_offset  global index of local element 0 (even, so Rosenbrock pairs stay together)
 *   comm           NULL on one GPU */
int  lbfgsb200_create(const lbfgsb200_param_t *param, int64_t n_local, int64_t n_global,
                      int64_t global_offset, int device, void *stream, lbfgsb200_comm_t *comm,
                      lbfgsb200_solver_t **out);
void lbfgsb200_destroy(lbfgsb200_solver_t *solver);
const char *lbfgsb200_last_error(const lbfgsb200_solver_t *solver);

/* Lbfgs::minimize  src/lbfgs.rs:399-421.  Returns an LBFGSB200_OK_* / ERR_* status; x_dev holds
 * the final point on return. */
int lbfgsb200_minimize(lbfgsb200_solver_t *solver, double *x_dev, lbfgsb200_eval_fn eval, void *eval_user,
                       lbfgsb200_progress_fn progress, void *progress_user, lbfgsb200_report_t *report);

/* Registers (fn != NULL) or clears the fused trial evaluate for the following build()/minimize() calls. */
int lbfgsb200_set_trial_evaluate(lbfgsb200_solver_t *solver, lbfgsb200_trial_eval_fn fn, void *user);
/* Registers everything an objective offers (ops != NULL) or clears it (NULL) for the following build()/minimize()
 * calls.  With probe + commit the line search runs write-free probes and one commit per iteration; with only
 * `trial` it runs the one-pass trial; otherwise K1 + evaluate + K2.  All three give the same bits. */
int lbfgsb200_set_fused_ops(lbfgsb200_solver_t *solver, const lbfgsb200_fused_ops_t *ops);

/* How the search direction H.(-g) of src/lbfgs.rs:569-604 is formed (an extension; the reference has one way).
 *   TWO_LOOP (default)  the reference's recursion, one fused pass per trip: 2 * min(m, k) dependent passes over the
 *                       vector, (8 b - 1) V of memory traffic, 2 b reductions (and 2 b cross-GPU exchanges).
 *   COMPACT             the same recursion with the same element-wise operations in the same order, but its 2 b
 *                       scalars alpha_j / beta_j come from inner products of the UNMODIFIED ring vectors (S^T Y and
 *                       Y^T Y are kept on the device across iterations): two passes over the ring, (4 b + 4) V, 2
 *                       reductions, one all-reduce.  Given equal scalars d is bit-identical; the scalars differ from
 *                       the reference's by rounding only — less than the reference's own sensitivity to the order
 *                       in which its dot products are summed.  m <= 32.
 * Call before build() / minimize().  The environment variable LBFGSB200_DIRECTION=compact makes COMPACT the default
 * of every solver created afterwards (that is how lbfgsb200_minimize_host* pick it up). */
enum {
    LBFGSB200_DIRECTION_TWO_LOOP = 0,
    LBFGSB200_DIRECTION_COMPACT = 1
};
int lbfgsb200_set_direction(lbfgsb200_solver_t *solver, int mode);
int lbfgsb200_get_direction(const lbfgsb200_solver_t *solver);
/* Process-wide default for solvers created afterwards, including the ones lbfgsb200_minimize_host* create
 * internally; -1 = back to the environment's choice. */
int lbfgsb200_set_default_direction(int mode);

/* The iterative API  src/lbfgs.rs:443-566 */
int lbfgsb200_build(lbfgsb200_solver_t *solver, double *x_dev, lbfgsb200_eval_fn eval, void *eval_user);
/* is_converged (src/lbfgs.rs:489-494): 1 = stop, 0 = continue; *stop_status gets the OK_* reason */
int lbfgsb200_is_converged(lbfgsb200_solver_t *solver, int *stop_status);
int lbfgsb200_propagate(lbfgsb200_solver_t *solver, lbfgsb200_progress_t *progress_out);
int lbfgsb200_report(lbfgsb200_solver_t *solver, lbfgsb200_report_t *report_out);
/* copies the current point into the caller's x_dev if it lives in the solver's own buffer
 * (x and xp ping-pong instead of save_state's copies, src/core.rs:207-210) */
int lbfgsb200_finish(lbfgsb200_solver_t *solver);
/* device pointers of the current point / gradient / search direction (this rank's shard) */
const double *lbfgsb200_x(const lbfgsb200_solver_t *solver);
const double *lbfgsb200_gx(const lbfgsb200_solver_t *solver);
const double *lbfgsb200_direction(const lbfgsb200_solver_t *solver);

/* Reference-shaped convenience: x is a HOST slice as in `minimize(&mut x, ..)` (src/lbfgs.rs:399).
 * Copies x to the device, solves on one GPU, copies the result back (both copies inside the call). */
int lbfgsb200_minimize_host(const lbfgsb200_param_t *param, double *x_host, int64_t n, int device,
                            lbfgsb200_eval_fn eval, void *eval_user, lbfgsb200_progress_fn progress,
                            void *progress_user, lbfgsb200_report_t *report);

/* The same with everything the device API offers: this rank's shard of a sharded vector (comm != NULL) and the
 * objective's fused line-search entries (fused may be NULL).  x_host should be pinned memory for full PCIe
 * speed. */
int lbfgsb200_minimize_host_ex(const lbfgsb200_param_t *param, double *x_host, int64_t n_local, int64_t n_global,
                               int64_t global_offset, int device, lbfgsb200_comm_t *comm, lbfgsb200_eval_fn eval,
                               void *eval_user, const lbfgsb200_fused_ops_t *fused,
                               lbfgsb200_progress_fn progress, void *progress_user, lbfgsb200_report_t *report);

/* ---- instrumentation ---------------------------------------------------------------------- */
enum {
    LBFGSB200_K_DOTS = 0,       /* {g.d, g.g, x.x} in one read                       src/core.rs:114-116,183-194 */
    LBFGSB200_K_OWL_PG = 1,     /* l1 norm + pseudo-gradient + norms                 src/orthantwise.rs:70-112 */
    LBFGSB200_K_INIT_DIR = 2,   /* d = -g | -pg, d.d, g.d                            src/core.rs:95-101 */
    LBFGSB200_K_TRIAL = 3,      /* x = xp + step*d (+ orthant projection)            src/core.rs:155-164 */
    LBFGSB200_K_ORTHANT = 4,    /* wp                                                src/core.rs:167-180 */
    LBFGSB200_K_HISTORY = 5,    /* s, y, s.s, y.s, y.y, s.d, s.Bs                    src/lbfgs.rs:640-656 */
    LBFGSB200_K_DAMP = 6,       /* Powell damping of y                               src/lbfgs.rs:666-689 */
    LBFGSB200_K_BACKWARD = 7,   /* two-loop backward step                            src/lbfgs.rs:582-591 */
    LBFGSB200_K_FORWARD = 8,    /* two-loop forward step                             src/lbfgs.rs:594-601 */
    LBFGSB200_K_EVALUATE = 9,   /* the user's device evaluate */
    LBFGSB200_K_PRIMITIVE = 10, /* unfused LbfgsMath primitives                      src/math.rs:31-82 */
    LBFGSB200_K_TRIAL_EVAL = 11,/* fused trial step + evaluate + dots (lbfgsb200_trial_eval_fn) */
    LBFGSB200_K_PROBE = 12,     /* write-free trial (lbfgsb200_probe_fn) */
    LBFGSB200_K_COMMIT = 13,    /* accepted point + history update (lbfgsb200_commit_fn) */
    LBFGSB200_K_UPDATE_SMALL = 14, /* launch-bound regime: history + whole two-loop in ONE cooperative kernel */
    LBFGSB200_K_COUNT = 15
};
typedef struct lbfgsb200_profile {
    int64_t launches[LBFGSB200_K_COUNT];        /* kernels launched (evaluate: callback invocations) */
    double  bytes[LBFGSB200_K_COUNT];           /* algorithmic bytes moved (passes * 8 * n_local) */
    double  ms[LBFGSB200_K_COUNT];              /* CUDA-event time on the solver's stream (timing on) */
    int64_t host_syncs;                         /* stream synchronisations made by the solver */
    int64_t allreduces;                         /* scalar all-reduces issued */
} lbfgsb200_profile_t;
/* timing = 1 brackets every launch with CUDA events on the solver's stream (costs ~2 % at n = 1e8);
 * timing = 2 << LBFGSB200_K_<kind> (OR-able) times only those kinds; 0 = counters only */
int lbfgsb200_profile_enable(lbfgsb200_solver_t *solver, int timing);
int lbfgsb200_profile_get(lbfgsb200_solver_t *solver, lbfgsb200_profile_t *out);
int lbfgsb200_profile_reset(lbfgsb200_solver_t *solver);

/* ---- LbfgsMath primitives on device pointers  src/math.rs:31-82 ---------------------------- */
int lbfgsb200_vecadd(double *y_dev, const double *x_dev, double c, int64_t n, void *stream);     /* y += c*x */
int lbfgsb200_vecdot(const double *x_dev, const double *y_dev, int64_t n, void *stream, double *out_host);
int lbfgsb200_vecscale(double *y_dev, double c, int64_t n, void *stream);                        /* y *= c */
int lbfgsb200_veccpy(double *y_dev, const double *x_dev, int64_t n, void *stream);               /* y = x */
int lbfgsb200_vecncpy(double *y_dev, const double *x_dev, int64_t n, void *stream);              /* y = -x */
int lbfgsb200_vecdiff(double *z_dev, const double *x_dev, const double *y_dev, int64_t n, void *stream); /* z = x - y */
int lbfgsb200_vec2norm(const double *x_dev, int64_t n, void *stream, double *out_host);
int lbfgsb200_vec2norminv(const double *x_dev, int64_t n, void *stream, double *out_host);

/* ---- fused hot-path steps, exposed for unit parity ----------------------------------------- */
/* {g.d, g.g, x.x}; d_dev may be NULL (then out[0] = 0) */
int lbfgsb200_dots3(const double *g_dev, const double *d_dev, const double *x_dev, int64_t n, void *stream,
                    double out_host[3]);
/* x = xp + step*d; wp_dev (int8 signs) non-NULL applies the orthant projection on [start,end) */
int lbfgsb200_trial_step(double *x_dev, const double *xp_dev, const double *d_dev, double step, int64_t n,
                         const signed char *wp_dev, int64_t start, int64_t end, void *stream);
/* OWL-QN pseudo-gradient (src/orthantwise.rs:70-112): out = {sum c|x| on [start,end), pg.pg, x.x} */
int lbfgsb200_owl_pseudo_gradient(double *pg_dev, const double *x_dev, const double *g_dev, int64_t n,
                                  double c, int64_t start, int64_t end, void *stream, double out_host[3]);
/* wp = xp == 0 ? signum(-pg) : signum(xp)  (src/core.rs:167-180), stored as int8 */
int lbfgsb200_owl_orthant(signed char *wp_dev, const double *xp_dev, const double *pg_dev, int64_t n,
                          void *stream);
/* d = 0 where signum(d) != signum(-pg) on [start,end) (src/orthantwise.rs:140-161); out = {d.d after} */
int lbfgsb200_owl_constrain_direction(double *d_dev, const double *pg_dev, int64_t n, int64_t start,
                                      int64_t end, void *stream, double out_host[1]);

/* The update chain's own kernels, one call each.  Scalars that the solver keeps on the device (alpha, beta,
 * y.s, gamma) are passed BY VALUE here and uploaded by the wrapper, so every production kernel can be driven
 * with arbitrary vectors.  Results are read back with one stream synchronisation per call. */
/* d = -g; out = {d.d, g.d}                                             src/core.rs:95-101, src/lbfgs.rs:457-461 */
int lbfgsb200_init_direction(double *d_dev, const double *g_dev, int64_t n, void *stream, double out_host[2]);
/* IterationData::update's vector work (src/lbfgs.rs:640-656, :670-673): s = x - xp, y = g - gp,
 *   out = {s.s, y.s, y.y, s.(-g), s.(gp * -step)}   (the last one only with damping != 0, else 0)
 * pg_dev non-NULL (OWL-QN): out[3] = s.(-pg), the first alpha's numerator with d = -pg (src/core.rs:96-97) */
int lbfgsb200_history_update(double *s_dev, double *y_dev, const double *x_dev, const double *xp_dev,
                             const double *g_dev, const double *gp_dev, const double *pg_dev, int64_t n,
                             double step, int damping, void *stream, double out_host[5]);
/* Powell damping (src/lbfgs.rs:664-689), decided by the kernel from y.s and s.Bs: case 1 (y.s < 0.4 s.Bs)
 * rewrites y = ((gp * -step) * (1 - theta)) + theta * y, theta = 0.6 s.Bs / (s.Bs - y.s); otherwise y is left
 * alone (case 2 computes and discards).  *applied_host = 1 when y was rewritten. */
int lbfgsb200_damp_y(double *y_dev, const double *gp_dev, int64_t n, double step, double ys, double sbs,
                     void *stream, int *applied_host);
/* One trip of the backward loop (src/lbfgs.rs:582-591): alpha = sq / ys_j; q = q - alpha*y_j, where the first
 * trip takes q = -g from g_first_dev (non-NULL) instead of reading q.  s_next_dev non-NULL: out = {alpha,
 * s_next.q} (the next trip's numerator); NULL (last trip): q *= gamma (:591) and out = {alpha, y_j.q} */
int lbfgsb200_two_loop_backward_step(double *q_dev, const double *g_first_dev, const double *y_j_dev,
                                     const double *s_next_dev, int64_t n, double sq, double ys_j, double gamma,
                                     void *stream, double out_host[2]);
/* One trip of the forward loop (src/lbfgs.rs:594-601): beta = yr / ys_j; r += (alpha_j - beta)*s_j.
 * y_next_dev non-NULL: out = {beta, y_next.r}.  NULL (last trip): g_last_dev = g, out = {beta, r.r, g.r}
 * (src/lbfgs.rs:543, src/core.rs:78-92); with owl != 0 g_last_dev = pg and the direction is projected on
 * [owl_start, owl_end) (src/orthantwise.rs:140-161): out = {beta, r.r before, pg.d after, d.d after} */
int lbfgsb200_two_loop_forward_step(double *r_dev, const double *s_j_dev, const double *y_next_dev,
                                    const double *g_last_dev, int64_t n, double yr, double ys_j, double alpha_j,
                                    int owl, int64_t owl_start, int64_t owl_end, void *stream, double out_host[4]);

/* ---- built-in device objectives (lbfgsb200_eval_fn-compatible) ------------------------------ */
typedef struct lbfgsb200_objective lbfgsb200_objective_t;
/* default_evaluate (Rosenbrock)  src/lib.rs:79-94; n_local must be even */
int  lbfgsb200_objective_rosenbrock(int device, lbfgsb200_objective_t **out);
/* Booth function  tests/simple.rs:65-74 (n = 2) */
int  lbfgsb200_objective_booth(int device, lbfgsb200_objective_t **out);
/* dense GLM, X row-major nrow x ncol in device memory (not copied), y nrow:
 *   kind 0 = Poisson log-linear  tests/owlqn.rs:22-43;  kind 1 = logistic (BASELINE.json configs[2]) */
int  lbfgsb200_objective_glm(int device, int kind, const double *X_dev, const double *y_dev, int64_t nrow,
                             int64_t ncol, lbfgsb200_objective_t **out);
/* all-pairs Lennard-Jones  examples/lj.rs:20-64,114-117; n = 3 * atoms */
int  lbfgsb200_objective_lennard_jones(int device, double epsilon, double sigma, lbfgsb200_objective_t **out);
/* Which kernels the last GLM evaluation ran (diagnostic; 0 before the first one): the one-pass kernel covers even
 * ncol up to 10 240 per CTA, odd ncol up to 6 143 (rows only 8-byte aligned), and even ncol up to 163 840 with the
 * columns split over a thread-block cluster; anything else takes the two-pass kernels (X read twice). */
enum {
    LBFGSB200_GLM_PATH_TWO_PASS = 1,
    LBFGSB200_GLM_PATH_FUSED = 2,
    LBFGSB200_GLM_PATH_FUSED_ODD = 3,
    LBFGSB200_GLM_PATH_FUSED_CLUSTER = 4
};
int  lbfgsb200_objective_last_path(const lbfgsb200_objective_t *objective);
/* Lennard-Jones per-pair arithmetic.  0 (default): the reference's — sqrt, sigma/r, powi, g*dr/r with IEEE
 * divisions (examples/lj.rs:23-32,50-57), so every pair term has the reference's bits.  1: the molecular-dynamics
 * form — only 1/r^2 (reciprocal seed + two Newton steps), fused multiply-adds; ~2.5x fewer FP64 instructions,
 * every pair term within a few ulp of the reference's.  Ignored (always 0) with LBFGSB200_REDUCE_SEQUENTIAL. */
int  lbfgsb200_objective_set_lj_fast(lbfgsb200_objective_t *objective, int fast);
void lbfgsb200_objective_destroy(lbfgsb200_objective_t *objective);
/* LBFGSB200_REDUCE_* for the objective's own sum (f); SEQUENTIAL is implemented for Rosenbrock, Booth and
 * Lennard-Jones (exp/log in the GLMs are not bit-reproducible against a CPU libm anyway) */
int  lbfgsb200_objective_set_reduction(lbfgsb200_objective_t *objective, int reduction);
/* Multi-GPU objectives.  Rosenbrock is shard-local: nothing to exchange for evaluate; its fused
 * trial / probe / commit kernels use the communicator to sum their scalars over the ranks in their own epilogue
 * (LBFGSB200_FUSED_SUMS_OVER_RANKS), so a trial costs no extra launch on N GPUs.
 *   GLM: this rank's X / y hold a block of ROWS; w is replicated on every rank and the solver runs unsharded
 *        (comm = NULL in lbfgsb200_create): eval all-reduces f and the ncol-vector gradient, so every rank sees the
 *        same bits.  shard_offsets is ignored.
 *   Lennard-Jones: the solver shards the 3N coordinates (lbfgsb200_create with comm); shard_offsets[0..nranks] are
 *        the element offsets of every rank's shard (multiples of 3): eval gathers all positions over NVLink and
 *        computes this rank's forces against all atoms (same bits as on one GPU) and its partial energy.
 * comm = NULL resets to single-GPU behaviour. */
int  lbfgsb200_objective_set_shard(lbfgsb200_objective_t *objective, lbfgsb200_comm_t *comm, const int64_t *shard_offsets);
/* the lbfgsb200_eval_fn for every built-in objective: pass the objective handle as `user` */
int  lbfgsb200_objective_eval(void *objective, const double *x_dev, double *g_dev, int64_t n_local,
                              void *stream, double *fx_dev);

/* the lbfgsb200_trial_eval_fn of the built-in objectives (`user` = the objective handle); implemented for
 * Rosenbrock, returns LBFGSB200_ERR_UNSUPPORTED for the others.  _has_trial_eval tells without calling. */
int  lbfgsb200_objective_trial_eval(void *objective, const double *xp_dev, const double *d_dev, double step,
                                    double *x_dev, double *g_dev, int64_t n_local, void *stream, double *out_dev);
int  lbfgsb200_objective_has_trial_eval(const lbfgsb200_objective_t *objective);
/* the lbfgsb200_probe_fn / lbfgsb200_commit_fn of the built-in objectives (Rosenbrock; the others return
 * LBFGSB200_ERR_UNSUPPORTED), and everything an objective offers in one struct (entries it lacks are NULL;
 * flags has LBFGSB200_FUSED_SUMS_OVER_RANKS once lbfgsb200_objective_set_shard attached a communicator whose
 * peer mailboxes the kernels can use). */
int  lbfgsb200_objective_probe(void *objective, const double *xp_dev, const double *d_dev, double step,
                               const double *step_dev, int64_t n_local, void *stream, double *out_dev);
int  lbfgsb200_objective_commit(void *objective, const double *xp_dev, const double *d_dev, const double *gp_dev,
                                double step, double bs_scale, double *x_dev, double *g_dev, double *s_dev,
                                double *y_dev, int64_t n_local, void *stream, double *out_dev);
/* the lbfgsb200_probe_multi_fn of the built-in objectives (Rosenbrock; LBFGSB200_ERR_UNSUPPORTED otherwise) */
int  lbfgsb200_objective_probe_multi(void *objective, const double *xp_dev, const double *d_dev, const double *steps,
                                     const double *step0_dev, int k, int64_t n_local, void *stream, double *out_dev);
/* the lbfgsb200_commit_gram_fn of the built-in objectives (Rosenbrock on one GPU; LBFGSB200_ERR_UNSUPPORTED otherwise) */
int  lbfgsb200_objective_commit_gram(void *objective, const double *xp_dev, const double *d_dev, const double *gp_dev,
                                     double step, double bs_scale, double *x_dev, double *g_dev, double *s_dev,
                                     double *y_dev, const double *const *s_old_dev, const double *const *y_old_dev,
                                     int n_old, int64_t n_local, void *stream, double *out_dev, double *gram_out_dev,
                                     double *newdot_out_dev);
int  lbfgsb200_objective_fused_ops(lbfgsb200_objective_t *objective, lbfgsb200_fused_ops_t *out);

/* ---- line-search state machines (pure host code; exposed so the scalar logic can be checked
 *      without a GPU)  src/line.rs:226-399, 446-709, 716-784 ----------------------------------- */
typedef struct lbfgsb200_linesearch lbfgsb200_linesearch_t;
lbfgsb200_linesearch_t *lbfgsb200_linesearch_begin(const lbfgsb200_param_t *param, int orthantwise,
                                                   double finit, double dginit, double step);
/* returns 1 and *step_out = next trial step; 0 when finished (see _result) */
int  lbfgsb200_linesearch_next(lbfgsb200_linesearch_t *ls, double *step_out);
/* between _next and _feed: the steps the search will ask for next IF the pending trial and each one after it
 * extrapolates (More-Thuente, interval not bracketed: stp + 4 (stp - stx), src/line.rs:266); returns how many were
 * written (0: not predictable).  Does not change the state: the driver evaluates them ahead of time in the same pass
 * (lbfgsb200_probe_multi_fn) and uses a result only when the search then asks for exactly that step. */
int  lbfgsb200_linesearch_predict(const lbfgsb200_linesearch_t *ls, double *steps_out, int kmax);
void lbfgsb200_linesearch_feed(lbfgsb200_linesearch_t *ls, int eval_ok, double f, double dg);
/* after _next returned 0: *ncall, final *step; returns LBFGSB200_LS_ERR_* (0 = success) */
int  lbfgsb200_linesearch_result(lbfgsb200_linesearch_t *ls, int64_t *ncall, double *step);
void lbfgsb200_linesearch_end(lbfgsb200_linesearch_t *ls);

/* ---- device-memory helpers for hosts without a CUDA binding (Rust, ctypes) ----------------- */
int  lbfgsb200_device_count(void);
int  lbfgsb200_device_alloc(int device, int64_t bytes, void **out_dev);
int  lbfgsb200_device_free(void *dev);
int  lbfgsb200_copy_h2d(void *dst_dev, const void *src_host, int64_t bytes, void *stream);
int  lbfgsb200_copy_d2h(void *dst_host, const void *src_dev, int64_t bytes, void *stream);
int  lbfgsb200_stream_synchronize(void *stream);
/* Solver arenas come from a private CUDA memory pool per device (the device's default pool is left alone) and
 * stay cached there after lbfgsb200_destroy, so repeated solves do not pay the driver's map/unmap of ~(2m+5)
 * n-vectors each time.  This hands the cached pages back to the driver (e.g. before another library needs the
 * HBM).  LBFGSB200_POOL=0 disables the pool. */
int  lbfgsb200_trim_pool(int device);
/* 3.  Everything added since ABI 3 was first published is additive: new functions (lbfgsb200_set_direction,
 * _get_direction, _set_default_direction, _linesearch_predict, _objective_probe_multi, _objective_commit_gram) and two
 * optional entries appended to lbfgsb200_fused_ops_t, whose struct_size tells the library which layout the caller
 * was built against (LBFGSB200_FUSED_OPS_SIZE_V1 / _V2 / sizeof). */
int  lbfgsb200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LBFGSB200_H */
