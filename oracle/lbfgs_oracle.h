/* lbfgs_oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a single-threaded CPU restatement of the
 * reference solver (ybyygu/rust-lbfgs, `liblbfgs` 0.2.0).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (rust_lbfgs_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  The reference cannot be compiled here (no rustc/cargo
 * in the image), so the oracle is pinned against every known-answer the
 * reference's own tests assert (src/math.rs:84-122, tests/simple.rs:37-40,
 * :52-54, :81-82, tests/owlqn.rs:60) — see tests/test_oracle_pins.py.
 */
#ifndef LBFGS_ORACLE_H
#define LBFGS_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* line-search algorithms, src/line.rs:39-80 */
enum {
    ORACLE_LS_MORETHUENTE = 0,
    ORACLE_LS_BACKTRACKING_ARMIJO = 1,
    ORACLE_LS_BACKTRACKING_WOLFE = 2,
    ORACLE_LS_BACKTRACKING_STRONG_WOLFE = 3
};

/* termination status of oracle_minimize */
enum {
    ORACLE_OK_CONVERGED = 0,        /* gnorm/max(1,xnorm) <= epsilon, src/lbfgs.rs:714-722 */
    ORACLE_OK_MAX_ITERATIONS = 1,   /* src/lbfgs.rs:726-735 */
    ORACLE_OK_MAX_EVALUATIONS = 2,  /* src/lbfgs.rs:739-748 */
    ORACLE_OK_CANCELLED = 3,        /* progress callback returned true, src/lbfgs.rs:412-416 */
    ORACLE_ERR_EVALUATE = -1,       /* evaluate failed at the initial point, src/lbfgs.rs:454 */
    ORACLE_ERR_X_NOT_CHANGED = -2,  /* src/lbfgs.rs:645-646 */
    ORACLE_ERR_G_NOT_CHANGED = -3,  /* src/lbfgs.rs:655 */
    ORACLE_ERR_LINESEARCH = -4,     /* Err out of LineSearch::find itself, src/line.rs:198-201,208 */
    ORACLE_ERR_INVALID_PARAM = -5,  /* the reference's assert!/panic on parameters */
    ORACLE_ERR_OWLQN_ZERO_DIRECTION = -6, /* src/orthantwise.rs:160 */
    ORACLE_ERR_INVALID_DNORM = -7   /* src/lbfgs.rs:544 */
};

/* LbfgsParam + LineSearch + Orthantwise flattened, src/lbfgs.rs:72-154,
 * src/line.rs:91-148, src/orthantwise.rs:19-45.  Only 8-byte fields. */
typedef struct oracle_param {
    int64_t m;
    double  epsilon;
    int64_t past;
    double  delta;
    int64_t max_iterations;
    int64_t max_evaluations;
    int64_t ls_algorithm;
    double  ls_ftol;
    double  ls_gtol;
    double  ls_xtol;
    double  ls_min_step;
    double  ls_max_step;
    int64_t ls_max_linesearch;
    int64_t ls_gradient_only;
    int64_t orthantwise;        /* 0 = None */
    double  owl_c;
    int64_t owl_start;
    int64_t owl_end;            /* < 0 = None (=> n) */
    double  initial_inverse_hessian;
    double  max_step_size;
    int64_t damping;
    int64_t constrain_step_size;
    int64_t reduction_mode;     /* 0 = faithful sequential sums; 1 = compensated (Neumaier) sums */
} oracle_param_t;

/* Progress, src/core.rs:221-250 */
typedef struct oracle_progress {
    const double *x;
    const double *gx;
    int64_t n;
    double  fx;
    double  xnorm;
    double  gnorm;
    double  step;
    int64_t niter;
    int64_t neval;
    int64_t ncall;
} oracle_progress_t;

/* Report, src/core.rs:271-285 */
typedef struct oracle_report {
    double  fx;
    double  xnorm;
    double  gnorm;
    int64_t neval;
    int64_t niter;              /* extra: number of propagate() calls made */
    int64_t last_ls_error;      /* extra: code of the last swallowed line-search error (0 = none) */
} oracle_report_t;

/* E: FnMut(&[f64], &mut [f64]) -> Result<f64>, src/core.rs:10-13; *err != 0 is Err */
typedef double (*oracle_eval_fn)(void *user, const double *x, double *g, int64_t n, int *err);
/* G: FnMut(&Progress) -> bool, true cancels, src/lbfgs.rs:402 */
typedef int (*oracle_progress_fn)(void *user, const oracle_progress_t *prgr);

void oracle_param_default(oracle_param_t *p);

int oracle_minimize(const oracle_param_t *param, double *x, int64_t n,
                    oracle_eval_fn eval, void *eval_user,
                    oracle_progress_fn progress, void *progress_user,
                    oracle_report_t *report, char *errbuf, size_t errbuf_len);

/* src/line.rs:15-31: one LineSearch::find from x along a caller-chosen direction d */
int oracle_line_search(const oracle_param_t *param, double *x, int64_t n, const double *d, double *step,
                       oracle_eval_fn eval, void *eval_user, int64_t *ncall, int64_t *ls_error, double *fx_out);

/* LbfgsMath for [f64], src/math.rs:31-82 */
void   oracle_vecadd(double *y, const double *x, double c, int64_t n);
double oracle_vecdot(const double *x, const double *y, int64_t n);
void   oracle_vecscale(double *y, double c, int64_t n);
void   oracle_veccpy(double *y, const double *x, int64_t n);
void   oracle_vecncpy(double *y, const double *x, int64_t n);
void   oracle_vecdiff(double *z, const double *x, const double *y, int64_t n);
double oracle_vec2norm(const double *x, int64_t n);
double oracle_vec2norminv(const double *x, int64_t n);

/* Orthantwise pieces, src/orthantwise.rs:70-180, src/core.rs:167-180 */
double oracle_owl_x1norm(const double *x, int64_t n, double c, int64_t start, int64_t end);
void   oracle_owl_pseudo_gradient(double *pg, const double *x, const double *g, int64_t n,
                                  double c, int64_t start, int64_t end);
void   oracle_owl_project(double *x, const double *sign_of, int64_t n, int64_t start, int64_t end,
                          int negate_sign);
void   oracle_owl_orthant(double *wp, const double *xp, const double *pg, int64_t n);

/* Built-in objectives with the oracle_eval_fn signature. */
/* src/lib.rs:79-94; user = NULL, or an int64_t* holding reduction_mode */
double oracle_eval_rosenbrock(void *user, const double *x, double *g, int64_t n, int *err);
/* tests/simple.rs:65-74 */
double oracle_eval_booth(void *user, const double *x, double *g, int64_t n, int *err);

/* dense GLM objectives; X is row-major nrow x ncol */
typedef struct oracle_glm {
    const double *X;
    const double *y;
    int64_t nrow;
    int64_t ncol;
    int64_t reduction_mode;
} oracle_glm_t;
/* tests/owlqn.rs:22-43: fx = -(sum(y*Xb - exp(Xb))), g = -X^T (y - exp(Xb)) */
double oracle_eval_poisson(void *user, const double *x, double *g, int64_t n, int *err);
/* logistic analogue (BASELINE.json configs[2]): fx = sum(log(1+exp(z)) - y z), g = X^T (sigmoid(z) - y) */
double oracle_eval_logistic(void *user, const double *x, double *g, int64_t n, int *err);

/* examples/lj.rs:20-64,114-117: all-pairs Lennard-Jones, x = 3*natoms; user = double[2]{epsilon, sigma} or NULL */
double oracle_eval_lj(void *user, const double *x, double *g, int64_t n, int *err);

#ifdef __cplusplus
}
#endif
#endif
