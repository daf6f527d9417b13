// lbfgs_oracle.cpp — CPU oracle for the L-BFGS / OWL-QN hot path.
//
// TEST INFRASTRUCTURE ONLY (see lbfgs_oracle.h).  A single-threaded restatement of the
// reference algorithm with strictly sequential left-to-right sums and no FMA contraction
// (build with -O2 -ffp-contract=off, never -ffast-math), so that on IEEE f64 it produces
// what rustc produces for the reference.  Every function cites the reference lines it
// follows (paths relative to the reference checkout).
//
// Parity status: PINNED against the reference's own asserted known answers
// (tests/test_oracle_pins.py): src/math.rs:84-122, tests/simple.rs:37-40,52-54,81-82,
// tests/owlqn.rs:60.  The Lennard-Jones objective (examples/lj.rs) has no reference
// test: its parity is UNPINNED (depends on vecfx 0.1 `vecdist`, un-vendored).
//
// reduction_mode 1 swaps every sum for a Neumaier-compensated sum (nothing else); it is
// the "accurate" comparison arm for n >= 1e7 where a sequential f64 sum is itself off by
// ~1e-9 relative (SURVEY.md §7 "Hard parts").

#include "lbfgs_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

typedef std::vector<double> Vec;

int g_mode = 0;  // reduction mode of the solve in flight (the oracle is single-threaded)
int g_ls_variant = 0;  // experiment switch, see line_search_morethuente
int g_dir_variant = 0; // 0 = the reference's two-loop recursion; 1 = "compact" direction, see two_loop_compact

// ----- sums ---------------------------------------------------------------------------
struct Acc {
    double s = 0.0, c = 0.0;
    int mode;
    explicit Acc(int m) : mode(m) {}
    inline void add(double v) {
        if (mode == 0) {
            s += v;  // Iterator::sum() is a sequential fold
        } else {
            double t = s + v;
            if (std::fabs(s) >= std::fabs(v)) c += (s - t) + v; else c += (v - t) + s;
            s = t;
        }
    }
    inline double value() const { return mode == 0 ? s : s + c; }
};

// ----- src/math.rs:31-82 -----------------------------------------------------------------
void vecadd(double *y, const double *x, double c, int64_t n) {  // :33-37
    for (int64_t i = 0; i < n; ++i) y[i] += c * x[i];
}
double vecdot(const double *x, const double *y, int64_t n) {  // :40-42
    Acc a(g_mode);
    for (int64_t i = 0; i < n; ++i) a.add(x[i] * y[i]);
    return a.value();
}
void vecscale(double *y, double c, int64_t n) {  // :45-49
    for (int64_t i = 0; i < n; ++i) y[i] *= c;
}
void veccpy(double *y, const double *x, int64_t n) {  // :52-56
    for (int64_t i = 0; i < n; ++i) y[i] = x[i];
}
void vecncpy(double *y, const double *x, int64_t n) {  // :59-63
    for (int64_t i = 0; i < n; ++i) y[i] = -x[i];
}
void vecdiff(double *z, const double *x, const double *y, int64_t n) {  // :66-70
    for (int64_t i = 0; i < n; ++i) z[i] = x[i] - y[i];
}
double vec2norm(const double *x, int64_t n) { return std::sqrt(vecdot(x, x, n)); }  // :73-76
double vec2norminv(const double *x, int64_t n) { return 1.0 / vec2norm(x, n); }     // :79-81

// ----- src/orthantwise.rs ----------------------------------------------------------------
double rust_signum(double x) {  // f64::signum
    if (std::isnan(x)) return x;
    return std::signbit(x) ? -1.0 : 1.0;
}
double signum(double x) {  // :174-180
    if (std::isnan(x) || x == 0.0) return 0.0;
    return rust_signum(x);
}

struct Owl {
    bool on = false;
    double c = 1.0;
    int64_t start = 0;
    int64_t end = -1;

    // :59-67; returns false where the reference's assert! panics
    bool start_end(int64_t n, int64_t &s, int64_t &e) const {
        s = start;
        e = (end < 0 ? n : end);
        if (e > n) e = n;
        return s < e;
    }
    double x1norm(const double *x, int64_t n) const {  // :70-79
        int64_t s, e;
        start_end(n, s, e);
        Acc a(g_mode);
        for (int64_t i = s; i < e; ++i) a.add(c * std::fabs(x[i]));
        return a.value();
    }
    void pseudo_gradient(double *pg, const double *x, const double *g, int64_t n) const {  // :82-112
        int64_t s, e;
        start_end(n, s, e);
        for (int64_t i = 0; i < s; ++i) pg[i] = g[i];
        for (int64_t i = s; i < e; ++i) {
            if (x[i] != 0.0) {
                pg[i] = g[i] + rust_signum(x[i]) * c;
            } else {
                double right_partial = g[i] + c;
                double left_partial = g[i] - c;
                if (right_partial < 0.0) pg[i] = right_partial;
                else if (left_partial > 0.0) pg[i] = left_partial;
                else pg[i] = 0.0;
            }
        }
        for (int64_t i = e; i < n; ++i) pg[i] = g[i];
    }
};

// :165-171 with the two call shapes of :118-133 (y = wp) and :140-161 (y = -pg)
void project(double *x, const double *y, int64_t s, int64_t e, bool negate) {
    for (int64_t i = s; i < e; ++i) {
        double yi = negate ? -y[i] : y[i];
        if (signum(x[i]) != signum(yi)) x[i] = 0.0;
    }
}

// ----- error plumbing (anyhow::Result restated as codes) -----------------------------------
enum LsErr {
    LS_OK = 0,
    LS_ERR_EVALUATE = 1,
    LS_ERR_ROUNDING = 2,         // src/line.rs:292-298
    LS_ERR_XTOL = 3,             // :300-302
    LS_ERR_MAX_STEP = 4,         // :305-308, :171-174
    LS_ERR_MIN_STEP = 5,         // :310-313, :167-170
    LS_ERR_OUT_OF_INTERVAL = 6,  // :474-476
    LS_ERR_INCREASE_GRADIENT = 7,// :477-479
    LS_ERR_INCORRECT_TMINMAX = 8 // :480-483
};

// ----- src/core.rs Problem -----------------------------------------------------------------
struct Problem {
    double *x;
    int64_t n;
    double fx = 0.0;
    Vec gx, xp, gp, pg, wp, d;
    oracle_eval_fn eval;
    void *eval_user;
    Owl owl;
    int64_t neval = 0;

    Problem(double *x_, int64_t n_, oracle_eval_fn e, void *u, const Owl &o)  // :59-75
        : x(x_), n(n_), gx(n_, 0.0), xp(n_, 0.0), gp(n_, 0.0), pg(n_, 0.0), wp(n_, 0.0), d(n_, 0.0),
          eval(e), eval_user(u), owl(o) {}

    double dginit() const {  // :78-92 (dginit > 0 only warns)
        return owl.on ? vecdot(pg.data(), d.data(), n) : vecdot(gx.data(), d.data(), n);
    }
    void update_search_direction() {  // :95-101
        vecncpy(d.data(), owl.on ? pg.data() : gx.data(), n);
    }
    double dg_unchecked() const { return vecdot(gx.data(), d.data(), n); }  // :114-116
    bool evaluate() {  // :119-132; false = Err
        int err = 0;
        double f = eval(eval_user, x, gx.data(), n, &err);
        if (err) return false;
        fx = f;
        if (owl.on) {
            fx += owl.x1norm(x, n);
            owl.pseudo_gradient(pg.data(), x, gx.data(), n);
        }
        neval += 1;
        return true;
    }
    void take_line_step(double step) {  // :155-164
        veccpy(x, xp.data(), n);
        vecadd(x, d.data(), step, n);
        if (owl.on) {
            int64_t s, e;
            owl.start_end(n, s, e);
            project(x, wp.data(), s, e, false);  // constraint_line_search, orthantwise.rs:118-133
        }
    }
    void update_orthant_new_point() {  // :167-180
        for (int64_t i = 0; i < n; ++i)
            wp[i] = (xp[i] == 0.0) ? signum(-pg[i]) : signum(xp[i]);
    }
    double gnorm() const { return owl.on ? vec2norm(pg.data(), n) : vec2norm(gx.data(), n); }  // :183-189
    double xnorm() const { return vec2norm(x, n); }                                          // :192-194
    void revert() {  // :201-204 (fx is NOT restored)
        veccpy(x, xp.data(), n);
        veccpy(gx.data(), gp.data(), n);
    }
    void save_state() {  // :207-210
        veccpy(xp.data(), x, n);
        veccpy(gp.data(), gx.data(), n);
    }
    // :213-217 + orthantwise.rs:140-161; false where assert_ne! panics
    bool constrain_search_direction() {
        if (!owl.on) return true;
        int64_t s, e;
        owl.start_end(n, s, e);
        project(d.data(), pg.data(), s, e, true);
        return vec2norm(d.data(), n) != 0.0;
    }
};

// ----- src/line.rs -------------------------------------------------------------------------
struct LineSearch {  // :91-163
    int algorithm;
    double ftol, gtol, xtol, min_step, max_step;
    int64_t max_linesearch;
    bool gradient_only;

    LsErr validate_step(double step) const {  // :166-177
        if (step < min_step) return LS_ERR_MIN_STEP;
        if (step > max_step) return LS_ERR_MAX_STEP;
        return LS_OK;
    }
};

// :620-637
void cubic_minimizer(double &cm, double u, double fu, double du, double v, double fv, double dv) {
    double d = v - u;
    double theta = (fu - fv) * 3.0 / d + du + dv;
    double p = std::fabs(theta);
    double q = std::fabs(du);
    double r = std::fabs(dv);
    double s = std::fmax(std::fmax(p, q), r);
    double a = theta / s;
    double gamma = s * std::sqrt(a * a - du / s * (dv / s));
    if (v < u) gamma = -gamma;
    p = gamma - du + theta;
    q = gamma - du + gamma + dv;
    r = p / q;
    cm = u + r * d;
}

// :652-680
void cubic_minimizer2(double &cm, double u, double fu, double du, double v, double fv, double dv,
                      double xmin, double xmax) {
    double d = v - u;
    double theta = (fu - fv) * 3.0 / d + du + dv;
    double p = std::fabs(theta);
    double q = std::fabs(du);
    double r = std::fabs(dv);
    double s = std::fmax(std::fmax(p, q), r);
    double a = theta / s;
    double gamma = s * std::sqrt(std::fmax(0.0, a * a - du / s * (dv / s)));
    if (u < v) gamma = -gamma;
    p = gamma - dv + theta;
    q = gamma - dv + gamma + du;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) cm = v - r * d;
    else if (v > u) cm = xmax;
    else cm = xmin;
}

// :692-695
void quard_minimizer(double &qm, double u, double fu, double du, double v, double fv) {
    double a = v - u;
    qm = u + du / ((fu - fv) / a + du) / 2.0 * a;
}
// :706-709
void quard_minimizer2(double &qm, double u, double du, double v, double dv) {
    double a = u - v;
    qm = v + dv / (dv - du) * a;
}

// mcstep::update_trial_interval, :446-606 (always returns uinfo = 0 on success)
LsErr update_trial_interval(double &x, double &fx, double &dx, double &y, double &fy, double &dy,
                            double &t, double ft, double dt, double tmin, double tmax, bool &brackt) {
    bool dsign = dt * (dx / std::fabs(dx)) < 0.0;  // :461
    double mc = 0.0, mq = 0.0, newt = 0.0;

    if (brackt) {  // :470-484
        if (t <= std::fmin(x, y) || std::fmax(x, y) <= t) return LS_ERR_OUT_OF_INTERVAL;
        else if (0.0 <= dx * (t - x)) return LS_ERR_INCREASE_GRADIENT;
        else if (tmax < tmin) return LS_ERR_INCORRECT_TMINMAX;
    }

    int bound;
    if (fx < ft) {  // case 1, :487-501
        brackt = true;
        cubic_minimizer(mc, x, fx, dx, t, ft, dt);
        quard_minimizer(mq, x, fx, dx, t, ft);
        if (std::fabs(mc - x) < std::fabs(mq - x)) newt = mc;
        else newt = mc + 0.5 * (mq - mc);
        bound = 1;
    } else if (dsign) {  // case 2, :502-516
        brackt = true;
        cubic_minimizer(mc, x, fx, dx, t, ft, dt);
        quard_minimizer2(mq, x, dx, t, dt);
        if (std::fabs(mc - t) > std::fabs(mq - t)) newt = mc;
        else newt = mq;
        bound = 0;
    } else if (std::fabs(dt) < std::fabs(dx)) {  // case 3, :517-542
        cubic_minimizer2(mc, x, fx, dx, t, ft, dt, tmin, tmax);
        quard_minimizer2(mq, x, dx, t, dt);
        if (brackt) {
            if (std::fabs(t - mc) < std::fabs(t - mq)) newt = mc;
            else newt = mq;
        } else if (std::fabs(t - mc) > std::fabs(t - mq)) newt = mc;
        else newt = mq;
        bound = 1;
    } else {  // case 4, :543-557
        if (brackt) cubic_minimizer(newt, t, ft, dt, y, fy, dy);
        else if (x < t) newt = tmax;
        else newt = tmin;
        bound = 0;
    }

    if (fx < ft) {  // :567-583
        y = t; fy = ft; dy = dt;
    } else {
        if (dsign) { y = x; fy = fx; dy = dx; }
        x = t; fx = ft; dx = dt;
    }

    if (tmax < newt) newt = tmax;  // :586-591
    if (newt < tmin) newt = tmin;

    if (brackt && bound != 0) {  // :595-604
        mq = x + 0.66 * (y - x);
        if (x < y) { if (mq < newt) newt = mq; }
        else if (newt < mq) newt = mq;
    }
    t = newt;
    return LS_OK;
}

// :226-399; returns the evaluation count through `ncall`
LsErr line_search_morethuente(Problem &prb, double &stp, const LineSearch &param, int64_t &ncall) {
    double dginit = prb.dginit();
    bool brackt = false;
    int stage1 = 1;
    int uinfo = 0;

    double finit = prb.fx;
    double dgtest = param.ftol * dginit;
    double width = param.max_step - param.min_step;
    double prev_width = 2.0 * width;

    double stx = 0.0, sty = 0.0;
    double fx = finit, fy = finit;
    double dgy = dginit, dgx = dgy;

    for (int64_t count = 1; count < param.max_linesearch; ++count) {  // :258
        double stmin, stmax;
        if (brackt) {
            stmin = (stx <= sty) ? stx : sty;
            stmax = (stx >= sty) ? stx : sty;
        } else {
            stmin = stx;
            stmax = stp + 4.0 * (stp - stx);
        }
        if (stp < param.min_step) stp = param.min_step;  // :269-274
        if (param.max_step < stp) stp = param.max_step;

        if ((brackt && (stp <= stmin || stmax <= stp || param.max_linesearch <= count + 1 || uinfo != 0)) ||
            (brackt && stmax - stmin <= param.xtol * stmax)) {  // :278-282
            stp = stx;
        }

        prb.take_line_step(stp);
        if (!prb.evaluate()) return LS_ERR_EVALUATE;  // :286
        double f = prb.fx;
        double dg = prb.dg_unchecked();
        double ftest1 = finit + stp * dgtest;

        if (brackt && (stp <= stmin || stmax <= stp || uinfo != 0)) return LS_ERR_ROUNDING;  // :292-298
        if (brackt && stmax - stmin <= param.xtol * stmax) return LS_ERR_XTOL;               // :300-302
        if (stp == param.max_step && f <= ftest1 && dg <= dgtest) return LS_ERR_MAX_STEP;    // :305-308
        if (stp == param.min_step && (ftest1 < f || dgtest <= dg)) return LS_ERR_MIN_STEP;   // :310-313

        // ORACLE_LS_VARIANT=1 (an experiment, never the default): skip the curvature-only exit so that only the
        // sufficient-decrease + curvature branch below can end the search — the C liblbfgs condition that :315-317
        // shadows.  Used by scripts/explain_iteration_counts.py to explain tests/simple.rs:33-35,48-50.
        if (g_ls_variant != 1 && std::fabs(dg) <= param.gtol * -dginit) {  // :315-317 (curvature alone)
            ncall = count;
            return LS_OK;
        } else if (f <= ftest1 && std::fabs(dg) <= param.gtol * -dginit) {  // :318-320 (unreachable)
            ncall = count;
            return LS_OK;
        } else {
            if (stage1 != 0 && f <= ftest1 && std::fmin(param.ftol, param.gtol) * dginit <= dg) stage1 = 0;  // :324-326

            LsErr e;
            if (stage1 != 0 && ftest1 < f && f <= fx) {  // :333-361
                double fm = f - stp * dgtest;
                double fxm = fx - stx * dgtest;
                double fym = fy - sty * dgtest;
                double dgm = dg - dgtest;
                double dgxm = dgx - dgtest;
                double dgym = dgy - dgtest;
                e = update_trial_interval(stx, fxm, dgxm, sty, fym, dgym, stp, fm, dgm, stmin, stmax, brackt);
                if (e != LS_OK) return e;
                uinfo = 0;
                fx = fxm + stx * dgtest;
                fy = fym + sty * dgtest;
                dgx = dgxm + dgtest;
                dgy = dgym + dgtest;
            } else {  // :362-377
                e = update_trial_interval(stx, fx, dgx, sty, fy, dgy, stp, f, dg, stmin, stmax, brackt);
                if (e != LS_OK) return e;
                uinfo = 0;
            }

            if (!brackt) continue;  // :381-383
            if (0.66 * prev_width <= std::fabs(sty - stx)) stp = stx + 0.5 * (sty - stx);  // :385-387
            prev_width = width;
            width = std::fabs(sty - stx);
        }
    }
    ncall = param.max_linesearch;  // :396-398
    return LS_OK;
}

// :716-784
LsErr line_search_backtracking(Problem &prb, double &stp, const LineSearch &param, int64_t &ncall) {
    double dginit = prb.dginit();
    const double dec = 0.5, inc = 2.1;
    double finit = prb.fx;
    double dgtest = param.ftol * dginit;

    bool orthantwise = prb.owl.on;
    if (orthantwise) prb.update_orthant_new_point();  // :734-736

    double width;
    for (int64_t count = 1; count < param.max_linesearch; ++count) {  // :739
        prb.take_line_step(stp);
        if (!prb.evaluate()) return LS_ERR_EVALUATE;

        if (prb.fx > finit + stp * dgtest) {  // :745
            width = dec;
        } else if (param.algorithm == ORACLE_LS_BACKTRACKING_ARMIJO || orthantwise) {  // :747-750
            ncall = count;
            return LS_OK;
        } else {
            double dg = prb.dg_unchecked();
            if (dg < param.gtol * dginit) {  // :754
                width = inc;
            } else if (param.algorithm == ORACLE_LS_BACKTRACKING_WOLFE) {  // :756-758
                ncall = count;
                return LS_OK;
            } else if (dg > -param.gtol * dginit) {  // :759
                width = dec;
            } else {
                ncall = count;
                return LS_OK;
            }
        }

        if (param.gradient_only) {  // :768-774
            double dg = prb.dg_unchecked();
            if (std::fabs(dg) <= -param.gtol * std::fabs(dginit)) {
                ncall = count;
                return LS_OK;
            }
        }

        LsErr e = param.validate_step(stp);  // :776
        if (e != LS_OK) return e;
        stp *= width;
    }
    ncall = param.max_linesearch;  // :783
    return LS_OK;
}

// LineSearch::find, :193-223.  Returns false for an Err out of find itself.
bool linesearch_find(const LineSearch &ls, Problem &prb, double &step, int64_t &ncall, int64_t &swallowed) {
    if (std::signbit(step)) return false;  // :198-201 ensure!(step.is_sign_positive())
    LsErr e;
    int64_t count = 0;
    if (ls.algorithm == ORACLE_LS_MORETHUENTE && !prb.owl.on) {  // :204
        if (ls.gradient_only) return false;                      // :208 bail!
        e = line_search_morethuente(prb, step, ls, count);
    } else {
        e = line_search_backtracking(prb, step, ls, count);
    }
    if (e != LS_OK) {  // :213-220
        prb.revert();
        swallowed = e;
        count = 0;
    }
    ncall = count;
    return true;
}

// ----- src/lbfgs.rs ------------------------------------------------------------------------
struct IterationData {  // :607-627
    double alpha = 0.0;
    Vec s, y;
    double ys = 0.0;
    explicit IterationData(int64_t n) : s(n, 0.0), y(n, 0.0) {}

    // :640-692; returns 0, or the error status
    int update(const double *x, const double *xp, const double *gx, const double *gp, int64_t n,
               double step, bool damping, double &gamma) {
        vecdiff(s.data(), x, xp, n);
        double d = vec2norm(s.data(), n);
        if (!(d != 0.0)) return ORACLE_ERR_X_NOT_CHANGED;  // ensure!(d != 0.0); NaN != 0 is true
        vecdiff(y.data(), gx, gp, n);

        double ys_ = vecdot(y.data(), s.data(), n);
        double yy = vecdot(y.data(), y.data(), n);
        if (!(yy != 0.0)) return ORACLE_ERR_G_NOT_CHANGED;
        ys = ys_;

        const double sigma2 = 0.6, sigma3 = 3.0;  // :664-665
        if (damping) {
            Vec bs(gp, gp + n);  // :670
            vecscale(bs.data(), -step, n);
            double sbs = vecdot(s.data(), bs.data(), n);
            if (ys_ < (1.0 - sigma2) * sbs) {  // case 1, :675-680
                double theta = sigma2 * sbs / (sbs - ys_);
                vecscale(bs.data(), 1.0 - theta, n);
                vecadd(bs.data(), y.data(), theta, n);
                veccpy(y.data(), bs.data(), n);
            } else if (ys_ > (1.0 + sigma3) * sbs) {  // case 2, :681-685: bs computed, y untouched
                double theta = sigma3 * sbs / (ys_ - sbs);
                vecscale(bs.data(), 1.0 - theta, n);
                vecadd(bs.data(), y.data(), theta, n);
            }
        }
        gamma = ys_ / yy;  // :691
        return 0;
    }
};

// NOT reference behaviour — the checker of the product's opt-in LBFGSB200_DIRECTION_COMPACT mode.  The same recursion
// as :569-604 with the same element-wise operations in the same order, but the 2 * bound scalars alpha_j / beta_j are
// derived from inner products of the UNMODIFIED vectors (s_i.y_j, y_i.y_j, s_i.d0, y_i.d0) instead of from the
// vector as it is rewritten: s_j.q = s_j.d0 - sum_{i newer than j} alpha_i (s_j.y_i), and so on.  In exact arithmetic
// both give the same numbers; in floating point the scalars differ by rounding only.
int64_t two_loop_compact(std::vector<IterationData> &lm, double *d, int64_t n, double gamma, int64_t m, int64_t k,
                         int64_t end) {
    end = (end + 1) % m;
    const int64_t bound = (m < k) ? m : k;
    std::vector<int64_t> slot(bound);
    for (int64_t t = 0, j = end; t < bound; ++t) { j = (j + m - 1) % m; slot[t] = j; }   // newest ... oldest
    std::vector<double> sg(bound), yg(bound), SY(bound * bound), YY(bound * bound), alpha(bound), coef(bound);
    for (int64_t a = 0; a < bound; ++a) {
        sg[a] = vecdot(lm[slot[a]].s.data(), d, n);
        yg[a] = vecdot(lm[slot[a]].y.data(), d, n);
        for (int64_t b = 0; b < bound; ++b) {
            SY[a * bound + b] = vecdot(lm[slot[a]].s.data(), lm[slot[b]].y.data(), n);
            YY[a * bound + b] = vecdot(lm[slot[a]].y.data(), lm[slot[b]].y.data(), n);
        }
    }
    for (int64_t a = 0; a < bound; ++a) {               // backward: newest -> oldest
        double acc = sg[a];
        for (int64_t i = 0; i < a; ++i) acc += -alpha[i] * SY[a * bound + i];
        alpha[a] = acc / lm[slot[a]].ys;
        lm[slot[a]].alpha = alpha[a];
    }
    for (int64_t a = bound - 1; a >= 0; --a) {          // forward: oldest -> newest
        double acc = yg[a];
        for (int64_t i = 0; i < bound; ++i) acc += -alpha[i] * YY[a * bound + i];
        acc = acc * gamma;
        for (int64_t i = bound - 1; i > a; --i) acc += coef[i] * SY[i * bound + a];
        const double beta = acc / lm[slot[a]].ys;
        coef[a] = alpha[a] - beta;
    }
    for (int64_t a = 0; a < bound; ++a) vecadd(d, lm[slot[a]].y.data(), -alpha[a], n);   // the element-wise passes of :589-599
    vecscale(d, gamma, n);
    for (int64_t a = bound - 1; a >= 0; --a) vecadd(d, lm[slot[a]].s.data(), coef[a], n);
    return end;
}

// :569-604
int64_t two_loop_recursion(std::vector<IterationData> &lm, double *d, int64_t n, double gamma,
                           int64_t m, int64_t k, int64_t end) {
    if (g_dir_variant == 1) return two_loop_compact(lm, d, n, gamma, m, k, end);
    end = (end + 1) % m;
    int64_t j = end;
    int64_t bound = (m < k) ? m : k;
    for (int64_t t = 0; t < bound; ++t) {
        j = (j + m - 1) % m;
        IterationData &it = lm[j];
        it.alpha = vecdot(it.s.data(), d, n) / it.ys;
        vecadd(d, it.y.data(), -it.alpha, n);
    }
    vecscale(d, gamma, n);
    for (int64_t t = 0; t < bound; ++t) {
        IterationData &it = lm[j];
        double beta = vecdot(it.y.data(), d, n) / it.ys;
        vecadd(d, it.s.data(), it.alpha - beta, n);
        j = (j + 1) % m;
    }
    return end;
}

struct State {  // LbfgsState, :425-439
    oracle_param_t vars;
    LineSearch ls;
    Problem prb;
    int64_t end = 0;
    double step = 0.0;
    int64_t k = 0;
    std::vector<IterationData> lm;
    int64_t ncall = 0;
    int64_t swallowed = 0;

    State(const oracle_param_t &p, const LineSearch &l, double *x, int64_t n, oracle_eval_fn e, void *u,
          const Owl &o)
        : vars(p), ls(l), prb(x, n, e, u, o) {
        for (int64_t i = 0; i < p.m; ++i) lm.emplace_back(n);  // :449
    }

    void fill_progress(oracle_progress_t &pr, double step_value) const {  // core.rs:253-268
        pr.x = prb.x;
        pr.gx = prb.gx.data();
        pr.n = prb.n;
        pr.fx = prb.fx;
        pr.xnorm = prb.xnorm();
        pr.gnorm = prb.gnorm();
        pr.neval = prb.neval;
        pr.ncall = ncall;
        pr.step = step_value;
        pr.niter = k;
    }

    // satisfying_stop_conditions, :697-748; returns -100 when no condition holds
    int stop_status() const {
        oracle_progress_t pr;
        fill_progress(pr, step);
        if (vars.max_iterations != 0 && pr.niter >= vars.max_iterations) return ORACLE_OK_MAX_ITERATIONS;
        if (vars.max_evaluations != 0 && pr.neval >= vars.max_evaluations) return ORACLE_OK_MAX_EVALUATIONS;
        if (pr.gnorm / std::fmax(pr.xnorm, 1.0) <= vars.epsilon) return ORACLE_OK_CONVERGED;
        return -100;
    }

    // :503-560; returns 0 or an error status
    int propagate(oracle_progress_t &pr) {
        k += 1;
        if (k == 1) {  // :507-510
            fill_progress(pr, step);
            return 0;
        }
        prb.save_state();
        if (!linesearch_find(ls, prb, step, ncall, swallowed)) return ORACLE_ERR_LINESEARCH;
        double step_ls = step;

        double gamma = 0.0;
        int rc = lm[end].update(prb.x, prb.xp.data(), prb.gx.data(), prb.gp.data(), prb.n, step,
                                vars.damping != 0, gamma);
        if (rc != 0) return rc;

        prb.update_search_direction();
        end = two_loop_recursion(lm, prb.d.data(), prb.n, gamma, vars.m, k - 1, end);

        double dnorm = vec2norm(prb.d.data(), prb.n);
        if (std::signbit(dnorm)) return ORACLE_ERR_INVALID_DNORM;  // :544

        if (vars.constrain_step_size) step = std::fmin(vars.max_step_size, dnorm) / dnorm;  // :547-551
        else step = 1.0;

        if (!prb.constrain_search_direction()) return ORACLE_ERR_OWLQN_ZERO_DIRECTION;  // :554

        fill_progress(pr, step_ls);  // :556-557
        return 0;
    }
};

const char *status_text(int st) {
    switch (st) {
        case ORACLE_ERR_EVALUATE: return "evaluate failed at the initial point";
        case ORACLE_ERR_X_NOT_CHANGED: return "x not changed";
        case ORACLE_ERR_G_NOT_CHANGED: return "gx not changed";
        case ORACLE_ERR_LINESEARCH: return "Failure during line search";
        case ORACLE_ERR_INVALID_PARAM: return "invalid parameter";
        case ORACLE_ERR_OWLQN_ZERO_DIRECTION: return "invalid direction vector after constraints";
        case ORACLE_ERR_INVALID_DNORM: return "invalid norm value";
        default: return "";
    }
}

}  // namespace

// ----- C ABI -------------------------------------------------------------------------------
extern "C" {

void oracle_param_default(oracle_param_t *p) {  // src/lbfgs.rs:156-177, src/line.rs:150-163, orthantwise.rs:47-55
    p->m = 6;
    p->epsilon = 1e-5;
    p->past = 0;
    p->delta = 1e-5;
    p->max_iterations = 0;
    p->max_evaluations = 0;
    p->ls_algorithm = ORACLE_LS_MORETHUENTE;
    p->ls_ftol = 1e-4;
    p->ls_gtol = 0.9;
    p->ls_xtol = 2.220446049250313e-16;  // f64::EPSILON
    p->ls_min_step = 1e-20;
    p->ls_max_step = 1e+20;
    p->ls_max_linesearch = 20;
    p->ls_gradient_only = 0;
    p->orthantwise = 0;
    p->owl_c = 1.0;
    p->owl_start = 0;
    p->owl_end = -1;
    p->initial_inverse_hessian = 1.0;
    p->max_step_size = 1.0;
    p->damping = 0;
    p->constrain_step_size = 1;
    p->reduction_mode = 0;
}

int oracle_minimize(const oracle_param_t *param, double *x, int64_t n, oracle_eval_fn eval, void *eval_user,
                    oracle_progress_fn progress, void *progress_user, oracle_report_t *report, char *errbuf,
                    size_t errbuf_len) {
    if (errbuf && errbuf_len) errbuf[0] = 0;
    auto fail = [&](int st) {
        if (errbuf && errbuf_len) snprintf(errbuf, errbuf_len, "%s", status_text(st));
        return st;
    };
    g_mode = (int)param->reduction_mode;
    {
        const char *v = getenv("ORACLE_LS_VARIANT");
        g_ls_variant = (v && v[0] == '1') ? 1 : 0;
        const char *w = getenv("ORACLE_DIRECTION_VARIANT");
        g_dir_variant = (w && w[0] == '1') ? 1 : 0;
    }

    Owl owl;
    owl.on = param->orthantwise != 0;
    owl.c = param->owl_c;
    owl.start = param->owl_start;
    owl.end = param->owl_end;
    if (param->m < 1 || n < 1) return fail(ORACLE_ERR_INVALID_PARAM);
    if (owl.on) {
        int64_t s, e;
        if (!owl.start_end(n, s, e)) return fail(ORACLE_ERR_INVALID_PARAM);   // orthantwise.rs:64
        if (std::signbit(owl.c)) return fail(ORACLE_ERR_INVALID_PARAM);       // orthantwise.rs:91
    }
    LineSearch ls;
    ls.algorithm = (int)param->ls_algorithm;
    ls.ftol = param->ls_ftol;
    ls.gtol = param->ls_gtol;
    ls.xtol = param->ls_xtol;
    ls.min_step = param->ls_min_step;
    ls.max_step = param->ls_max_step;
    ls.max_linesearch = param->ls_max_linesearch;
    ls.gradient_only = param->ls_gradient_only != 0;

    // build, src/lbfgs.rs:443-481
    State st(*param, ls, x, n, eval, eval_user, owl);
    auto fill_report = [&]() {
        if (!report) return;
        report->fx = st.prb.fx;  // core.rs:288-298
        report->xnorm = st.prb.xnorm();
        report->gnorm = st.prb.gnorm();
        report->neval = st.prb.neval;
        report->niter = st.k;
        report->last_ls_error = st.swallowed;
    };
    if (!st.prb.evaluate()) { fill_report(); return fail(ORACLE_ERR_EVALUATE); }
    st.prb.update_search_direction();
    st.step = vec2norminv(st.prb.d.data(), n) * param->initial_inverse_hessian;

    // minimize, src/lbfgs.rs:399-421
    int status;
    for (;;) {
        int s = st.stop_status();
        if (s != -100) { status = s; break; }
        oracle_progress_t pr;
        int rc = st.propagate(pr);
        if (rc != 0) { fill_report(); return fail(rc); }
        if (progress && progress(progress_user, &pr)) { status = ORACLE_OK_CANCELLED; break; }
    }
    fill_report();
    return status;
}

// The public low-level entry of src/line.rs:15-31: Problem::new, evaluate, a caller-chosen search
// direction, LineSearch::find.  Returns 0 or ORACLE_ERR_*; *ls_error gets the swallowed
// line-search error (0 = the search itself succeeded), *ncall / *step the results of find.
int oracle_line_search(const oracle_param_t *param, double *x, int64_t n, const double *d, double *step,
                       oracle_eval_fn eval, void *eval_user, int64_t *ncall, int64_t *ls_error, double *fx_out) {
    g_mode = (int)param->reduction_mode;
    Owl owl;
    owl.on = param->orthantwise != 0;
    owl.c = param->owl_c;
    owl.start = param->owl_start;
    owl.end = param->owl_end;
    LineSearch ls;
    ls.algorithm = (int)param->ls_algorithm;
    ls.ftol = param->ls_ftol;
    ls.gtol = param->ls_gtol;
    ls.xtol = param->ls_xtol;
    ls.min_step = param->ls_min_step;
    ls.max_step = param->ls_max_step;
    ls.max_linesearch = param->ls_max_linesearch;
    ls.gradient_only = param->ls_gradient_only != 0;
    Problem prb(x, n, eval, eval_user, owl);
    if (!prb.evaluate()) return ORACLE_ERR_EVALUATE;
    veccpy(prb.d.data(), d, n);
    prb.save_state();
    int64_t swallowed = 0, count = 0;
    if (!linesearch_find(ls, prb, *step, count, swallowed)) return ORACLE_ERR_LINESEARCH;
    *ncall = count;
    *ls_error = swallowed;
    *fx_out = prb.fx;
    return 0;
}

void   oracle_vecadd(double *y, const double *x, double c, int64_t n) { vecadd(y, x, c, n); }
double oracle_vecdot(const double *x, const double *y, int64_t n) { g_mode = 0; return vecdot(x, y, n); }
void   oracle_vecscale(double *y, double c, int64_t n) { vecscale(y, c, n); }
void   oracle_veccpy(double *y, const double *x, int64_t n) { veccpy(y, x, n); }
void   oracle_vecncpy(double *y, const double *x, int64_t n) { vecncpy(y, x, n); }
void   oracle_vecdiff(double *z, const double *x, const double *y, int64_t n) { vecdiff(z, x, y, n); }
double oracle_vec2norm(const double *x, int64_t n) { g_mode = 0; return vec2norm(x, n); }
double oracle_vec2norminv(const double *x, int64_t n) { g_mode = 0; return vec2norminv(x, n); }

double oracle_owl_x1norm(const double *x, int64_t n, double c, int64_t start, int64_t end) {
    Owl o; o.on = true; o.c = c; o.start = start; o.end = end;
    g_mode = 0;
    return o.x1norm(x, n);
}
void oracle_owl_pseudo_gradient(double *pg, const double *x, const double *g, int64_t n, double c,
                                int64_t start, int64_t end) {
    Owl o; o.on = true; o.c = c; o.start = start; o.end = end;
    o.pseudo_gradient(pg, x, g, n);
}
void oracle_owl_project(double *x, const double *sign_of, int64_t n, int64_t start, int64_t end,
                        int negate_sign) {
    Owl o; o.on = true; o.start = start; o.end = end;
    int64_t s, e;
    o.start_end(n, s, e);
    project(x, sign_of, s, e, negate_sign != 0);
}
void oracle_owl_orthant(double *wp, const double *xp, const double *pg, int64_t n) {  // core.rs:167-180
    for (int64_t i = 0; i < n; ++i) wp[i] = (xp[i] == 0.0) ? signum(-pg[i]) : signum(xp[i]);
}

// src/lib.rs:79-94 (n is assumed even, as in the reference, which would index out of bounds otherwise)
double oracle_eval_rosenbrock(void *user, const double *x, double *g, int64_t n, int *err) {
    (void)err;
    int mode = user ? (int)*(const int64_t *)user : 0;
    Acc fx(mode);
    for (int64_t i = 0; i + 1 < n; i += 2) {
        double t1 = 1.0 - x[i];
        double t2 = 10.0 * (x[i + 1] - x[i] * x[i]);
        g[i + 1] = 20.0 * t2;
        g[i] = -2.0 * (x[i] * g[i + 1] + t1);
        fx.add(t1 * t1 + t2 * t2);
    }
    return fx.value();
}

// tests/simple.rs:65-74 (powi(2) = v*v)
double oracle_eval_booth(void *user, const double *x, double *g, int64_t n, int *err) {
    (void)user; (void)n; (void)err;
    double x1 = x[0], x2 = x[1];
    double a = x1 + 2.0 * x2 - 7.0;
    double b = 2.0 * x1 + x2 - 5.0;
    double fx = a * a + b * b;
    g[0] = 10.0 * x1 + 8.0 * x2 - 34.0;
    g[1] = 8.0 * x1 + 10.0 * x2 - 38.0;
    return fx;
}

// tests/owlqn.rs:22-43.  The reference evaluates this with nalgebra (via vecfx 0.1, version
// unpinned, un-vendored); nalgebra's gemv accumulates column by column (axpy over columns), and
// `.sum()` is a sequential fold.  prec = 0.0 as in tests/owlqn.rs:21.
double oracle_eval_poisson(void *user, const double *par, double *g, int64_t n, int *err) {
    (void)err;
    const oracle_glm_t *p = (const oracle_glm_t *)user;
    const int64_t nr = p->nrow, nc = p->ncol;
    const double prec = 0.0;
    (void)n;
    Vec xbeta(nr), t(nr);
    for (int64_t r = 0; r < nr; ++r) {  // xbeta = X * par, column-axpy order
        double z = par[0] * p->X[r * nc + 0];
        for (int64_t c = 1; c < nc; ++c) z += par[c] * p->X[r * nc + c];
        xbeta[r] = z;
    }
    Acc s1((int)p->reduction_mode), s2((int)p->reduction_mode);
    for (int64_t r = 0; r < nr; ++r) {
        double e = std::exp(xbeta[r]);
        s1.add(p->y[r] * xbeta[r] - e);
        t[r] = p->y[r] - e;
    }
    for (int64_t c = 0; c < nc; ++c) s2.add(prec * (par[c] * par[c]));
    double fx = -1.0 * s1.value() + 0.5 * s2.value();
    for (int64_t c = 0; c < nc; ++c) {  // g = (-X^T) * t + par * prec
        Acc a((int)p->reduction_mode);
        if (p->reduction_mode == 0) {
            double v = t[0] * (-p->X[0 * nc + c]);
            for (int64_t r = 1; r < nr; ++r) v += t[r] * (-p->X[r * nc + c]);
            g[c] = v + par[c] * prec;
        } else {
            for (int64_t r = 0; r < nr; ++r) a.add(t[r] * (-p->X[r * nc + c]));
            g[c] = a.value() + par[c] * prec;
        }
    }
    return fx;
}

// Logistic analogue of the fixture for BASELINE.json configs[2] (no reference code; the formulas
// below are the definition both the oracle and the device objective follow):
//   z = X w;  fx = sum_r softplus(z_r) - y_r z_r;  g = X^T (sigmoid(z) - y)
//   softplus(z) = max(z,0) + log1p(exp(-|z|));  sigmoid(z) = z>=0 ? 1/(1+exp(-z)) : exp(z)/(1+exp(z))
double oracle_eval_logistic(void *user, const double *w, double *g, int64_t n, int *err) {
    (void)err; (void)n;
    const oracle_glm_t *p = (const oracle_glm_t *)user;
    const int64_t nr = p->nrow, nc = p->ncol;
    Vec r_(nr);
    Acc f((int)p->reduction_mode);
    for (int64_t r = 0; r < nr; ++r) {
        Acc z((int)p->reduction_mode);
        for (int64_t c = 0; c < nc; ++c) z.add(w[c] * p->X[r * nc + c]);
        double zv = z.value();
        double sp = std::fmax(zv, 0.0) + std::log1p(std::exp(-std::fabs(zv)));
        double mu;
        if (zv >= 0.0) mu = 1.0 / (1.0 + std::exp(-zv));
        else { double e = std::exp(zv); mu = e / (1.0 + e); }
        f.add(sp - p->y[r] * zv);
        r_[r] = mu - p->y[r];
    }
    for (int64_t c = 0; c < nc; ++c) {
        Acc a((int)p->reduction_mode);
        for (int64_t r = 0; r < nr; ++r) a.add(r_[r] * p->X[r * nc + c]);
        g[c] = a.value();
    }
    return f.value();
}

// examples/lj.rs:20-64 + the closure at :114-117 (forces negated into a gradient).
// powi(v, 6) is the square-and-multiply chain (v*v)*((v*v)*(v*v)); vecdist is taken as
// sqrt(dx*dx + dy*dy + dz*dz) summed left to right (vecfx is un-vendored: UNPINNED).
double oracle_eval_lj(void *user, const double *x, double *g, int64_t n, int *err) {
    (void)err;
    double eps = 1.0, sigma = 1.0;
    if (user) { eps = ((const double *)user)[0]; sigma = ((const double *)user)[1]; }
    int64_t na = n / 3;
    for (int64_t i = 0; i < n; ++i) g[i] = 0.0;
    double energy = 0.0;
    for (int64_t i = 0; i < na; ++i) {
        for (int64_t j = 0; j < i; ++j) {
            double d0 = x[3 * i] - x[3 * j], d1 = x[3 * i + 1] - x[3 * j + 1], d2 = x[3 * i + 2] - x[3 * j + 2];
            double r = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            double q = sigma / r;
            double q2 = q * q;
            double s6 = q2 * (q2 * q2);
            energy += 4.0 * eps * (s6 * s6 - s6);
            double gr = 24.0 * eps * (s6 - 2.0 * (s6 * s6)) / r;
            for (int k = 0; k < 3; ++k) {
                double dr = x[3 * j + k] - x[3 * i + k];
                g[3 * i + k] += 1.0 * gr * dr / r;
                g[3 * j + k] += -1.0 * gr * dr / r;
            }
        }
    }
    for (int64_t i = 0; i < n; ++i) g[i] *= -1.0;
    return energy;
}

}  // extern "C"
