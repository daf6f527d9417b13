// linesearch.cpp — see linesearch.h.  Host-only scalar code; keep -ffp-contract=off.
#include "linesearch.h"

#include <cmath>

namespace lb {
namespace {

// A sample of the 1-D function phi(t) = f(xp + t d): position, value, slope.
struct Sample {
    double t, f, d;
};

// Shared head of the two cubic fits (src/line.rs:621-628, :653-662): returns theta and s.
inline void cubic_head(const Sample &u, const Sample &v, double &theta, double &s) {
    const double span = v.t - u.t;
    theta = (u.f - v.f) * 3.0 / span + u.d + v.d;
    s = std::fmax(std::fmax(std::fabs(theta), std::fabs(u.d)), std::fabs(v.d));
}

// Minimizer of the cubic through u and v (cubic_minimizer, src/line.rs:620-637).
inline double cubic_min(const Sample &u, const Sample &v) {
    double theta, s;
    cubic_head(u, v, theta, s);
    const double a = theta / s;
    double gamma = s * std::sqrt(a * a - u.d / s * (v.d / s));
    if (v.t < u.t) gamma = -gamma;
    const double p = gamma - u.d + theta;
    const double q = gamma - u.d + gamma + v.d;
    return u.t + p / q * (v.t - u.t);
}

// cubic_minimizer2 (src/line.rs:652-680): falls back to lo/hi when the cubic has no usable minimum.
inline double cubic_min_bounded(const Sample &u, const Sample &v, double lo, double hi) {
    double theta, s;
    cubic_head(u, v, theta, s);
    const double a = theta / s;
    double gamma = s * std::sqrt(std::fmax(0.0, a * a - u.d / s * (v.d / s)));
    if (u.t < v.t) gamma = -gamma;
    const double p = gamma - v.d + theta;
    const double q = gamma - v.d + gamma + u.d;
    const double r = p / q;
    if (r < 0.0 && gamma != 0.0) return v.t - r * (v.t - u.t);
    return (v.t > u.t) ? hi : lo;
}

// quard_minimizer (src/line.rs:692-695): quadratic through f(u), f'(u), f(v).
inline double quad_min_fdf(const Sample &u, const Sample &v) {
    const double a = v.t - u.t;
    return u.t + u.d / ((u.f - v.f) / a + u.d) / 2.0 * a;
}
// quard_minimizer2 (src/line.rs:706-709): secant through f'(u), f'(v).
inline double quad_min_secant(const Sample &u, const Sample &v) {
    const double a = u.t - v.t;
    return v.t + v.d / (v.d - u.d) * a;
}

// mcstep::update_trial_interval (src/line.rs:446-606).  best/other are the interval endpoints,
// `t` the trial (in: current, out: next).  Returns an LBFGSB200_LS_ERR_* code.
int update_interval(Sample &best, Sample &other, double &t, double ft, double dt, double tmin, double tmax,
                    bool &brackt) {
    const Sample cur{t, ft, dt};
    const bool opposite = dt * (best.d / std::fabs(best.d)) < 0.0;  // :461

    if (brackt) {  // :470-484
        if (t <= std::fmin(best.t, other.t) || std::fmax(best.t, other.t) <= t) return LBFGSB200_LS_ERR_OUT_OF_INTERVAL;
        if (0.0 <= best.d * (t - best.t)) return LBFGSB200_LS_ERR_INCREASE_GRADIENT;
        if (tmax < tmin) return LBFGSB200_LS_ERR_INCORRECT_TMINMAX;
    }

    double next;
    bool clamp_to_two_thirds;
    if (best.f < ft) {  // case 1: higher value => bracketed (:487-501)
        brackt = true;
        const double mc = cubic_min(best, cur);
        const double mq = quad_min_fdf(best, cur);
        next = (std::fabs(mc - best.t) < std::fabs(mq - best.t)) ? mc : mc + 0.5 * (mq - mc);
        clamp_to_two_thirds = true;
    } else if (opposite) {  // case 2: lower value, slopes of opposite sign => bracketed (:502-516)
        brackt = true;
        const double mc = cubic_min(best, cur);
        const double mq = quad_min_secant(best, cur);
        next = (std::fabs(mc - t) > std::fabs(mq - t)) ? mc : mq;
        clamp_to_two_thirds = false;
    } else if (std::fabs(dt) < std::fabs(best.d)) {  // case 3: slope magnitude decreases (:517-542)
        const double mc = cubic_min_bounded(best, cur, tmin, tmax);
        const double mq = quad_min_secant(best, cur);
        if (brackt) next = (std::fabs(t - mc) < std::fabs(t - mq)) ? mc : mq;
        else next = (std::fabs(t - mc) > std::fabs(t - mq)) ? mc : mq;
        clamp_to_two_thirds = true;
    } else {  // case 4 (:543-557)
        if (brackt) next = cubic_min(cur, other);
        else next = (best.t < t) ? tmax : tmin;
        clamp_to_two_thirds = false;
    }

    if (best.f < ft) {  // :567-583
        other = cur;
    } else {
        if (opposite) other = best;
        best = cur;
    }

    if (tmax < next) next = tmax;  // :586-591
    if (next < tmin) next = tmin;

    if (brackt && clamp_to_two_thirds) {  // :595-604
        const double mq = best.t + 0.66 * (other.t - best.t);
        if (best.t < other.t) {
            if (mq < next) next = mq;
        } else if (next < mq) {
            next = mq;
        }
    }
    t = next;
    return LBFGSB200_LS_ERR_NONE;
}

}  // namespace

int LineSearchMachine::begin(const LsConfig &cfg, bool orthantwise, double finit, double dginit, double step) {
    cfg_ = cfg;
    owl_ = orthantwise;
    err_ = 0;
    ncall_ = 0;
    count_ = 1;
    trials_ = 0;
    awaiting_ = false;
    done_ = true;
    stp_ = step;
    if (std::signbit(step)) return LBFGSB200_ERR_LINESEARCH;  // src/line.rs:198-201
    mt_ = (cfg.algorithm == LBFGSB200_LS_MORETHUENTE) && !orthantwise;  // :204
    if (mt_ && cfg.gradient_only) return LBFGSB200_ERR_LINESEARCH;    // :208
    done_ = false;
    finit_ = finit;
    dginit_ = dginit;
    dgtest_ = cfg.ftol * dginit;  // :243 / :730
    if (mt_) {  // :234-255
        brackt_ = false;
        stage1_ = 1;
        uinfo_ = 0;
        width_ = cfg.max_step - cfg.min_step;
        prev_width_ = 2.0 * width_;
        stx_ = sty_ = 0.0;
        fx_ = fy_ = finit;
        dgx_ = dgy_ = dginit;
    }
    return 0;
}

bool LineSearchMachine::next_trial(double *step_out) {
    if (done_) return false;
    if (!(count_ < cfg_.max_linesearch)) {  // loop `1..max_linesearch` exhausted: Ok(max_linesearch), :396-398,:783
        finish(cfg_.max_linesearch);
        return false;
    }
    if (mt_) {
        if (brackt_) {  // :261-265
            stmin_ = (stx_ <= sty_) ? stx_ : sty_;
            stmax_ = (stx_ >= sty_) ? stx_ : sty_;
        } else {
            stmin_ = stx_;
            stmax_ = stp_ + 4.0 * (stp_ - stx_);
        }
        if (stp_ < cfg_.min_step) stp_ = cfg_.min_step;  // :269-274
        if (cfg_.max_step < stp_) stp_ = cfg_.max_step;
        const bool unusual = (brackt_ && (stp_ <= stmin_ || stmax_ <= stp_ || cfg_.max_linesearch <= count_ + 1 || uinfo_ != 0)) ||
                             (brackt_ && stmax_ - stmin_ <= cfg_.xtol * stmax_);
        if (unusual) stp_ = stx_;  // :278-282
    }
    *step_out = stp_;
    awaiting_ = true;
    ++trials_;
    return true;
}

int LineSearchMachine::predict(double *steps, int kmax) const {
    if (done_ || !awaiting_ || !mt_ || brackt_) return 0;
    double stx = stx_, stp = stp_;
    int n = 0;
    int64_t count = count_;
    while (n < kmax && count + 1 < cfg_.max_linesearch) {
        const double next = stp + 4.0 * (stp - stx);        // stmax of the NEXT round, :266; the clamp of :583-588 lands on it
        if (!(next < cfg_.max_step) || !(next > stp)) break;
        steps[n++] = next;
        stx = stp;
        stp = next;
        ++count;
    }
    return n;
}

void LineSearchMachine::feed(bool eval_ok, double f, double dg) {
    if (done_ || !awaiting_) return;
    awaiting_ = false;
    if (!eval_ok) {
        fail(LBFGSB200_LS_ERR_EVALUATE);
        return;
    }
    if (mt_) feed_morethuente(f, dg);
    else feed_backtracking(f, dg);
}

void LineSearchMachine::feed_morethuente(double f, double dg) {
    const double ftest1 = finit_ + stp_ * dgtest_;  // :289

    // :292-313
    if (brackt_ && (stp_ <= stmin_ || stmax_ <= stp_ || uinfo_ != 0)) return fail(LBFGSB200_LS_ERR_ROUNDING);
    if (brackt_ && stmax_ - stmin_ <= cfg_.xtol * stmax_) return fail(LBFGSB200_LS_ERR_XTOL);
    if (stp_ == cfg_.max_step && f <= ftest1 && dg <= dgtest_) return fail(LBFGSB200_LS_ERR_MAX_STEP);
    if (stp_ == cfg_.min_step && (ftest1 < f || dgtest_ <= dg)) return fail(LBFGSB200_LS_ERR_MIN_STEP);

    // :315-320 — the curvature test alone decides; the sufficient-decrease arm behind it is unreachable
    if (std::fabs(dg) <= cfg_.gtol * -dginit_) return finish(count_);

    if (stage1_ != 0 && f <= ftest1 && std::fmin(cfg_.ftol, cfg_.gtol) * dginit_ <= dg) stage1_ = 0;  // :324-326

    int rc;
    if (stage1_ != 0 && ftest1 < f && f <= fx_) {  // modified function, :333-361
        Sample best{stx_, fx_ - stx_ * dgtest_, dgx_ - dgtest_};
        Sample other{sty_, fy_ - sty_ * dgtest_, dgy_ - dgtest_};
        const double fm = f - stp_ * dgtest_;
        const double dgm = dg - dgtest_;
        rc = update_interval(best, other, stp_, fm, dgm, stmin_, stmax_, brackt_);
        if (rc == 0) {
            stx_ = best.t;
            sty_ = other.t;
            fx_ = best.f + stx_ * dgtest_;
            fy_ = other.f + sty_ * dgtest_;
            dgx_ = best.d + dgtest_;
            dgy_ = other.d + dgtest_;
        }
    } else {  // :362-377
        Sample best{stx_, fx_, dgx_};
        Sample other{sty_, fy_, dgy_};
        rc = update_interval(best, other, stp_, f, dg, stmin_, stmax_, brackt_);
        if (rc == 0) {
            stx_ = best.t; fx_ = best.f; dgx_ = best.d;
            sty_ = other.t; fy_ = other.f; dgy_ = other.d;
        }
    }
    if (rc != 0) return fail(rc);
    uinfo_ = 0;  // update_trial_interval only ever returns Ok(0), :605

    if (brackt_) {  // :381-391
        if (0.66 * prev_width_ <= std::fabs(sty_ - stx_)) stp_ = stx_ + 0.5 * (sty_ - stx_);
        prev_width_ = width_;
        width_ = std::fabs(sty_ - stx_);
    }
    ++count_;
}

void LineSearchMachine::feed_backtracking(double f, double dg) {
    const double dec = 0.5, inc = 2.1;  // :725-726
    double width;
    if (f > finit_ + stp_ * dgtest_) {  // :745
        width = dec;
    } else if (cfg_.algorithm == LBFGSB200_LS_BACKTRACKING_ARMIJO || owl_) {  // :747-750
        return finish(count_);
    } else if (dg < cfg_.gtol * dginit_) {  // :754
        width = inc;
    } else if (cfg_.algorithm == LBFGSB200_LS_BACKTRACKING_WOLFE) {  // :756-758
        return finish(count_);
    } else if (dg > -cfg_.gtol * dginit_) {  // :759
        width = dec;
    } else {
        return finish(count_);
    }

    if (cfg_.gradient_only) {  // :768-774 (RHS <= 0: fires only when dg == 0 == dginit)
        if (std::fabs(dg) <= -cfg_.gtol * std::fabs(dginit_)) return finish(count_);
    }

    if (stp_ < cfg_.min_step) return fail(LBFGSB200_LS_ERR_MIN_STEP);  // validate_step, :166-177
    if (stp_ > cfg_.max_step) return fail(LBFGSB200_LS_ERR_MAX_STEP);
    stp_ *= width;
    ++count_;
}

}  // namespace lb
