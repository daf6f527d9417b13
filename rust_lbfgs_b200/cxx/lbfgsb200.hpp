// lbfgsb200.hpp — header-only C++17 host side over the C ABI (include/lbfgsb200.h).
//
// The reference is a Rust crate; this image has no Rust toolchain, so the compiled-language host mirror is
// C++ (the Rust `-sys` + safe wrapper sources are under rust_lbfgs_b200/rust/).  It
// reproduces the reference's public surface name for name:
//
//     let report = lbfgs().with_max_iterations(5)                      // src/lib.rs:38-50
//                         .with_orthantwise(1.0, 0, 99)
//                         .minimize(&mut x, evaluate, progress)?;
//
//     auto report = lbfgsb200::lbfgs().with_max_iterations(5)
//                                     .with_orthantwise(1.0, 0, 99)
//                                     .minimize(x_dev, n, evaluate, progress);
//
// with Progress / Report keeping their meaning (src/core.rs:221-299).  What is new is what north_star asks
// for: `x` is DEVICE memory and `evaluate` is a device-resident objective (DeviceEvaluate below: it gets raw
// device pointers and a stream), so x, g, d and the s/y history never leave HBM.  A reference-style HOST
// closure `FnMut(&[f64], &mut [f64]) -> Result<f64>` still works through host_evaluate(), at the price of
// one PCIe round trip per evaluation — the solver's own vector algebra stays on the GPU either way.
//
// Error behaviour follows the reference: `assert!` on a parameter -> std::invalid_argument at the call
// that the reference panics in; `Err(..)` from minimize -> lbfgsb200::Error carrying the status code and
// the reference's message text.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/lbfgsb200.h"

namespace lbfgsb200 {

// Progress, src/core.rs:221-250.  x / gx are device pointers (n elements of this rank's shard).
struct Progress {
    const double *x;
    const double *gx;
    int64_t n;
    double fx, xnorm, gnorm, step;
    int64_t niter, neval, ncall;
};

// Report, src/core.rs:271-285 (+ status / diagnostics)
struct Report {
    double fx = 0, xnorm = 0, gnorm = 0;
    int64_t neval = 0, niter = 0, last_ls_error = 0;
    int status = 0;
};

// The Err arm of the reference's Result<Report>.
class Error : public std::runtime_error {
  public:
    Error(int status, const std::string &msg, Report rep = {}) : std::runtime_error(msg), status(status), report(rep) {}
    int status;
    Report report;
};

// ---- device-resident evaluate -----------------------------------------------------------------------
// Replaces `E: FnMut(&[f64], &mut [f64]) -> Result<f64>` (src/core.rs:10-13).  The callable enqueues, on
// `stream`, work that writes the gradient to g_dev[0..n) and this rank's partial objective value to *fx_dev
// (device memory).  It must not synchronise.  Return 0 for Ok, non-zero for Err.
using DeviceEvaluate = std::function<int(const double *x_dev, double *g_dev, int64_t n, void *stream, double *fx_dev)>;
// `G: FnMut(&Progress) -> bool`; true cancels (src/lbfgs.rs:402,412-416)
using ProgressFn = std::function<bool(const Progress &)>;

// Adapter for a reference-style host closure: copies x to the host, calls f(x, gx) -> (ok, fx), copies gx and
// fx back.  One PCIe round trip per evaluation; meant for porting, not for speed.
inline DeviceEvaluate host_evaluate(std::function<bool(const std::vector<double> &x, std::vector<double> &gx, double &fx)> f) {
    auto xs = std::make_shared<std::vector<double>>();
    auto gs = std::make_shared<std::vector<double>>();
    return [f, xs, gs](const double *x_dev, double *g_dev, int64_t n, void *stream, double *fx_dev) -> int {
        xs->resize((size_t)n);
        gs->assign((size_t)n, 0.0);
        if (lbfgsb200_copy_d2h(xs->data(), x_dev, n * (int64_t)sizeof(double), stream) != 0) return 1;
        double fx = 0.0;
        if (!f(*xs, *gs, fx)) return 1;
        if (lbfgsb200_copy_h2d(g_dev, gs->data(), n * (int64_t)sizeof(double), stream) != 0) return 1;
        return lbfgsb200_copy_h2d(fx_dev, &fx, sizeof(double), stream) != 0;
    };
}

// Built-in device objectives (csrc/objectives.cu); usable wherever a DeviceEvaluate is expected.
class Objective {
  public:
    Objective() = default;
    Objective(const Objective &) = delete;
    Objective &operator=(const Objective &) = delete;
    Objective(Objective &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ~Objective() { if (h_) lbfgsb200_objective_destroy(h_); }
    // default_evaluate(), src/lib.rs:79-94
    static Objective rosenbrock(int device = 0) { Objective o; check(lbfgsb200_objective_rosenbrock(device, &o.h_)); return o; }
    // tests/simple.rs:65-74
    static Objective booth(int device = 0) { Objective o; check(lbfgsb200_objective_booth(device, &o.h_)); return o; }
    // examples/lj.rs:20-64
    // fast: the 1/r^2 form with fused multiply-adds instead of the reference's per-pair arithmetic (a few ulp apart)
    static Objective lennard_jones(double epsilon = 1.0, double sigma = 1.0, int device = 0, bool fast = false) {
        Objective o; check(lbfgsb200_objective_lennard_jones(device, epsilon, sigma, &o.h_));
        if (fast) check(lbfgsb200_objective_set_lj_fast(o.h_, 1));
        return o;
    }
    // kind 0 = Poisson (tests/owlqn.rs:22-43), 1 = logistic; X row-major nrow x ncol in device memory
    static Objective glm(int kind, const double *X_dev, const double *y_dev, int64_t nrow, int64_t ncol, int device = 0) {
        Objective o; check(lbfgsb200_objective_glm(device, kind, X_dev, y_dev, nrow, ncol, &o.h_)); return o;
    }
    lbfgsb200_objective_t *handle() const { return h_; }
    // multi-GPU: see lbfgsb200_objective_set_shard (offsets: Lennard-Jones only)
    void shard(lbfgsb200_comm_t *comm, const int64_t *shard_offsets = nullptr) const {
        check(lbfgsb200_objective_set_shard(h_, comm, shard_offsets));
    }
    // what the objective offers beyond evaluate (probe + commit, fused trial); fused = false: nothing
    lbfgsb200_fused_ops_t fused_ops(bool fused) const {
        lbfgsb200_fused_ops_t ops{};
        ops.struct_size = (int64_t)sizeof(ops);
        if (fused) check(lbfgsb200_objective_fused_ops(h_, &ops));
        return ops;
    }

  private:
    static void check(int rc) { if (rc != 0) throw Error(rc, "creating a device objective failed (no CUDA device?)"); }
    lbfgsb200_objective_t *h_ = nullptr;
};

// default_progress(), src/lib.rs:102-112
inline ProgressFn default_progress() {
    return [](const Progress &p) {
        std::printf("Iteration %lld, Evaluation %lld:\n", (long long)p.niter, (long long)p.neval);
        std::printf(" fx = %-12.6f xnorm = %-12.6f, gnorm = %-12.6f, ls = %lld, step = %g\n", p.fx, p.xnorm, p.gnorm,
                    (long long)p.ncall, p.step);
        return false;
    };
}

class LbfgsState;

// The builder, src/lbfgs.rs:179-384.
class Lbfgs {
  public:
    Lbfgs() { lbfgsb200_param_default(&p_); }

    Lbfgs &with_epsilon(double epsilon) {                       // :194-199
        require(!std::signbit(epsilon), "Invalid parameter epsilon specified.");
        p_.epsilon = epsilon; return *this;
    }
    Lbfgs &with_initial_step_size(double b) {                   // :203-211
        require(!std::signbit(b), "Invalid beta parameter for scaling the initial step size.");
        p_.initial_inverse_hessian = b; return *this;
    }
    Lbfgs &with_max_step_size(double s) {                       // :215-220
        require(!std::signbit(s), "Invalid max_step_size parameter.");
        p_.max_step_size = s; return *this;
    }
    Lbfgs &with_damping(bool damped) { p_.damping = damped; return *this; }   // :224-227
    // :231-245; end < 0 means None (all of x from `start`)
    Lbfgs &with_orthantwise(double c, int64_t start, int64_t end = -1) {
        require(!std::signbit(c), "Invalid parameter orthantwise c parameter specified.");
        p_.orthantwise = 1; p_.owl_c = c; p_.owl_start = start; p_.owl_end = end; return *this;
    }
    Lbfgs &with_linesearch_ftol(double ftol) {                  // :253-258
        require(ftol >= 0.0, "Invalid parameter ftol specified.");
        p_.ls_ftol = ftol; return *this;
    }
    Lbfgs &with_linesearch_gtol(double gtol) {                  // :268-276
        require(gtol >= 0.0 && gtol < 1.0 && gtol > p_.ls_ftol, "Invalid parameter gtol specified.");
        p_.ls_gtol = gtol; return *this;
    }
    Lbfgs &with_gradient_only() {                               // :283-289
        p_.ls_gradient_only = 1; p_.damping = 1; p_.ls_algorithm = LBFGSB200_LS_BACKTRACKING_STRONG_WOLFE; return *this;
    }
    Lbfgs &with_max_linesearch(int64_t n) { p_.ls_max_linesearch = n; return *this; }   // :292-296
    Lbfgs &with_linesearch_xtol(double xtol) {                  // :307-312
        require(xtol >= 0.0, "Invalid parameter xtol specified.");
        p_.ls_xtol = xtol; return *this;
    }
    Lbfgs &with_linesearch_min_step(double min_step) {          // :320-325
        require(min_step >= 0.0, "Invalid parameter min_step specified.");
        p_.ls_min_step = min_step; return *this;
    }
    Lbfgs &with_max_iterations(int64_t niter) { p_.max_iterations = niter; return *this; }     // :334-337
    Lbfgs &with_max_evaluations(int64_t neval) { p_.max_evaluations = neval; return *this; }   // :345-348
    Lbfgs &with_fx_delta(double delta, int64_t past) {          // :360-366 (stored, never consumed — as the reference)
        require(delta >= 0.0, "Invalid parameter delta specified.");
        p_.delta = delta; p_.past = past; return *this;
    }
    Lbfgs &with_linesearch_algorithm(const std::string &algo) { // :371-383
        if (algo == "MoreThuente") p_.ls_algorithm = LBFGSB200_LS_MORETHUENTE;
        else if (algo == "BacktrackingArmijo") p_.ls_algorithm = LBFGSB200_LS_BACKTRACKING_ARMIJO;
        else if (algo == "BacktrackingStrongWolfe") p_.ls_algorithm = LBFGSB200_LS_BACKTRACKING_STRONG_WOLFE;
        else if (algo == "BacktrackingWolfe" || algo == "Backtracking") p_.ls_algorithm = LBFGSB200_LS_BACKTRACKING_WOLFE;
        else throw std::logic_error("not implemented: " + algo);   // unimplemented!(), :379
        return *this;
    }
    // ---- extensions (not in the reference) ----
    Lbfgs &with_m(int64_t m) { require(m >= 1, "Invalid parameter m specified."); p_.m = m; return *this; }
    Lbfgs &with_reduction(int mode) { p_.reduction = mode; return *this; }                 // LBFGSB200_REDUCE_*
    Lbfgs &with_device(int device, void *stream = nullptr) { device_ = device; stream_ = stream; return *this; }
    Lbfgs &with_shard(lbfgsb200_comm_t *comm, int64_t n_global, int64_t global_offset) {
        comm_ = comm; n_global_ = n_global; goff_ = global_offset; return *this;
    }
    // false: line-search trials as K1 + evaluate + K2 even if the objective offers probe + commit / a fused trial
    Lbfgs &with_fused_trial(bool fused) { fused_ = fused; return *this; }
    // LBFGSB200_DIRECTION_COMPACT: the direction from two passes over the ring instead of 2 * min(m, k) dependent trips
    // (same element-wise operations; alpha_j / beta_j from inner products of the unmodified ring vectors).  m <= 32.
    Lbfgs &with_direction(int mode) { direction_ = mode; return *this; }

    // minimize, src/lbfgs.rs:399-421.  x_dev: n doubles of device memory, updated in place.
    Report minimize(double *x_dev, int64_t n, DeviceEvaluate evaluate, ProgressFn progress = nullptr) const {
        Handle h(*this, n);
        Report r = run(h.s, x_dev, tramp_eval, &evaluate, progress);
        return r;
    }
    Report minimize(double *x_dev, int64_t n, const Objective &objective, ProgressFn progress = nullptr) const {
        Handle h(*this, n);
        lbfgsb200_objective_set_reduction(objective.handle(), (int)p_.reduction);
        if (comm_) objective.shard(comm_);
        const lbfgsb200_fused_ops_t ops = objective.fused_ops(fused_);
        lbfgsb200_set_fused_ops(h.s, &ops);
        return run(h.s, x_dev, lbfgsb200_objective_eval, objective.handle(), progress);
    }
    // The reference's exact shape: x is a HOST slice (`minimize(&mut x, ..)`); copied to the device, solved on
    // one GPU, copied back.
    Report minimize(std::vector<double> &x, const Objective &objective, ProgressFn progress = nullptr) const {
        lbfgsb200_report_t rep{};
        const DefaultDirection dd(*this);
        lbfgsb200_objective_set_reduction(objective.handle(), (int)p_.reduction);
        if (comm_) objective.shard(comm_);
        const lbfgsb200_fused_ops_t ops = objective.fused_ops(fused_);
        const int64_t n = (int64_t)x.size();
        int st = lbfgsb200_minimize_host_ex(&p_, x.data(), n, n_global_ > 0 ? n_global_ : n, goff_, device_, comm_,
                                            lbfgsb200_objective_eval, objective.handle(), &ops,
                                            progress ? tramp_progress : nullptr, progress ? &progress : nullptr, &rep);
        Report r = convert(rep, st);
        if (st < 0) throw Error(st, "minimize failed (status " + std::to_string(st) + ")", r);
        return r;
    }
    Report minimize(std::vector<double> &x, DeviceEvaluate evaluate, ProgressFn progress = nullptr) const {
        lbfgsb200_report_t rep{};
        const DefaultDirection dd(*this);
        int st = lbfgsb200_minimize_host(&p_, x.data(), (int64_t)x.size(), device_, tramp_eval, &evaluate,
                                         progress ? tramp_progress : nullptr, progress ? &progress : nullptr, &rep);
        Report r = convert(rep, st);
        if (st < 0) throw Error(st, "minimize failed (status " + std::to_string(st) + ")", r);
        return r;
    }

    // build, src/lbfgs.rs:443-481 (the iterative API)
    inline LbfgsState build(double *x_dev, int64_t n, const Objective &objective) const;

    const lbfgsb200_param_t &param() const { return p_; }

  private:
    friend class LbfgsState;
    // The host-buffer entry points create their solver themselves: it takes the process-wide default direction.
    struct DefaultDirection {
        bool set = false;
        explicit DefaultDirection(const Lbfgs &b) {
            if (b.direction_ < 0) return;
            require(b.direction_ != LBFGSB200_DIRECTION_COMPACT || b.p_.m <= 32, "the compact direction supports m <= 32");
            set = lbfgsb200_set_default_direction(b.direction_) == 0;
        }
        DefaultDirection(const DefaultDirection &) = delete;
        DefaultDirection &operator=(const DefaultDirection &) = delete;
        ~DefaultDirection() { if (set) lbfgsb200_set_default_direction(-1); }
    };
    struct Handle {
        lbfgsb200_solver_t *s = nullptr;
        Handle(const Lbfgs &b, int64_t n) {
            const int64_t ng = b.n_global_ > 0 ? b.n_global_ : n;
            int rc = lbfgsb200_create(&b.p_, n, ng, b.goff_, b.device_, b.stream_, b.comm_, &s);
            if (rc == LBFGSB200_ERR_INVALID_PARAM) throw std::invalid_argument("invalid L-BFGS parameter");
            if (rc != 0) throw Error(rc, "lbfgsb200_create failed (no CUDA device? there is no CPU fallback)");
            if (b.direction_ >= 0 && (rc = lbfgsb200_set_direction(s, b.direction_)) != 0) {
                const std::string msg = lbfgsb200_last_error(s);
                lbfgsb200_destroy(s);
                s = nullptr;
                if (rc == LBFGSB200_ERR_INVALID_PARAM) throw std::invalid_argument(msg);
                throw Error(rc, msg);
            }
        }
        Handle(const Handle &) = delete;
        Handle &operator=(const Handle &) = delete;
        ~Handle() { if (s) lbfgsb200_destroy(s); }
    };
    static void require(bool ok, const char *msg) { if (!ok) throw std::invalid_argument(msg); }   // assert!(.., msg)
    static int tramp_eval(void *user, const double *x, double *g, int64_t n, void *stream, double *fx) {
        try { return (*static_cast<DeviceEvaluate *>(user))(x, g, n, stream, fx); } catch (...) { return 1; }
    }
    static int tramp_progress(void *user, const lbfgsb200_progress_t *p) {
        Progress q{p->x_dev, p->gx_dev, p->n_local, p->fx, p->xnorm, p->gnorm, p->step, p->niter, p->neval, p->ncall};
        return (*static_cast<ProgressFn *>(user))(q) ? 1 : 0;
    }
    static Report convert(const lbfgsb200_report_t &r, int st) {
        Report o; o.fx = r.fx; o.xnorm = r.xnorm; o.gnorm = r.gnorm; o.neval = r.neval; o.niter = r.niter;
        o.last_ls_error = r.last_ls_error; o.status = st; return o;
    }
    static Report run(lbfgsb200_solver_t *s, double *x_dev, lbfgsb200_eval_fn fn, void *user, ProgressFn &progress) {
        lbfgsb200_report_t rep{};
        int st = lbfgsb200_minimize(s, x_dev, fn, user, progress ? tramp_progress : nullptr, progress ? &progress : nullptr, &rep);
        Report r = convert(rep, st);
        if (st < 0) throw Error(st, lbfgsb200_last_error(s), r);
        return r;
    }

    lbfgsb200_param_t p_{};
    int device_ = 0;
    void *stream_ = nullptr;
    lbfgsb200_comm_t *comm_ = nullptr;
    int64_t n_global_ = 0, goff_ = 0;
    bool fused_ = true;
    int direction_ = -1;   // < 0: the library's default
};

// LbfgsState, src/lbfgs.rs:425-566
class LbfgsState {
  public:
    LbfgsState(const Lbfgs &b, double *x_dev, int64_t n, const Objective &objective) : h_(b, n) {
        lbfgsb200_objective_set_reduction(objective.handle(), (int)b.p_.reduction);
        if (b.comm_) objective.shard(b.comm_);
        const lbfgsb200_fused_ops_t ops = objective.fused_ops(b.fused_);
        lbfgsb200_set_fused_ops(h_.s, &ops);
        int st = lbfgsb200_build(h_.s, x_dev, lbfgsb200_objective_eval, objective.handle());
        if (st != 0) throw Error(st, lbfgsb200_last_error(h_.s));
    }
    bool is_converged() { int st = 0; return lbfgsb200_is_converged(h_.s, &st) == 1; }      // :489-494
    Progress propagate() {                                                                   // :503-560
        lbfgsb200_progress_t p{};
        int st = lbfgsb200_propagate(h_.s, &p);
        if (st != 0) throw Error(st, lbfgsb200_last_error(h_.s));
        return Progress{p.x_dev, p.gx_dev, p.n_local, p.fx, p.xnorm, p.gnorm, p.step, p.niter, p.neval, p.ncall};
    }
    Report report() {                                                                        // :497-499
        lbfgsb200_report_t r{};
        lbfgsb200_report(h_.s, &r);
        Report o; o.fx = r.fx; o.xnorm = r.xnorm; o.gnorm = r.gnorm; o.neval = r.neval; o.niter = r.niter;
        o.last_ls_error = r.last_ls_error; o.status = (int)r.status; return o;
    }
    void finish() { int st = lbfgsb200_finish(h_.s); if (st != 0) throw Error(st, lbfgsb200_last_error(h_.s)); }
    int direction() const { return lbfgsb200_get_direction(h_.s); }                          // LBFGSB200_DIRECTION_*

  private:
    Lbfgs::Handle h_;
};

inline LbfgsState Lbfgs::build(double *x_dev, int64_t n, const Objective &objective) const {
    return LbfgsState(*this, x_dev, n, objective);
}

// lbfgs(), src/lib.rs:74-76
inline Lbfgs lbfgs() { return Lbfgs(); }

}  // namespace lbfgsb200
