// linesearch.h — re-entrant scalar state machines for the line searches (pure host code).
//
// The reference's searches (src/line.rs:226-399 MoreThuente, :716-784 backtracking) are loops
// that call `take_line_step` / `evaluate` / `dg_unchecked` on host slices.  Here the vectors live
// in HBM, so the loop is turned inside out: the machine says "evaluate at this step"
// (next_trial), the driver launches trial_step + evaluate + dots on the GPU, and feeds the two
// scalars (f, dg) back (feed).  The scalar arithmetic — interval bookkeeping, the safeguarded
// cubic/quadratic interpolation of mcstep (src/line.rs:446-709) — follows the reference's
// operation order exactly (this file is compiled with -ffp-contract=off), because every rank of
// a multi-GPU solve replays it on bit-identical all-reduced scalars.
#pragma once

#include <stdint.h>

#include "../../include/lbfgsb200.h"

namespace lb {

struct LsConfig {                    // LineSearch, src/line.rs:91-148
    int algorithm = LBFGSB200_LS_MORETHUENTE;
    double ftol = 1e-4, gtol = 0.9, xtol = 2.220446049250313e-16;
    double min_step = 1e-20, max_step = 1e+20;
    int64_t max_linesearch = 20;
    bool gradient_only = false;
};

class LineSearchMachine {
  public:
    // LineSearch::find (src/line.rs:193-223).  Returns 0, or LBFGSB200_ERR_LINESEARCH for the two
    // Err paths of find itself (negative step :198-201; gradient-only + MoreThuente :208).
    int begin(const LsConfig &cfg, bool orthantwise, double finit, double dginit, double step);
    // true: evaluate at *step_out and call feed(); false: the search is over.
    bool next_trial(double *step_out);
    // (f, dg) at the trial point; eval_ok == false is an Err from evaluate (src/line.rs:286,743).
    void feed(bool eval_ok, double f, double dg);

    // The steps the search WILL ask for next if the trial just handed out (and each one after it) extrapolates — More-
    // Thuente with the interval not yet bracketed clamps the new trial to stmax = stp + 4 (stp - stx), src/line.rs:266 —
    // computed with the search's own expression, so a driver may evaluate them ahead of time and compare bit for bit.
    // Returns how many were written (0: not predictable).  Pure: does not change the state.
    int predict(double *steps, int kmax) const;

    int error() const { return err_; }          // LBFGSB200_LS_ERR_*; non-zero => caller reverts
    int64_t ncall() const { return ncall_; }    // the Ok(count) of the reference
    double step() const { return stp_; }        // the in/out `stp`
    int64_t trials() const { return trials_; }  // evaluations requested so far
    bool uses_morethuente() const { return mt_; }

  private:
    void finish(int64_t ncall) { ncall_ = ncall; done_ = true; }
    void fail(int code) { err_ = code; ncall_ = 0; done_ = true; }
    void feed_morethuente(double f, double dg);
    void feed_backtracking(double f, double dg);

    LsConfig cfg_;
    bool owl_ = false, mt_ = true, done_ = true, awaiting_ = false;
    int err_ = 0;
    int64_t ncall_ = 0, count_ = 1, trials_ = 0;
    double stp_ = 0.0, finit_ = 0.0, dginit_ = 0.0, dgtest_ = 0.0;
    // MoreThuente interval state (src/line.rs:234-255)
    bool brackt_ = false;
    int stage1_ = 1, uinfo_ = 0;
    double width_ = 0.0, prev_width_ = 0.0;
    double stx_ = 0.0, sty_ = 0.0, fx_ = 0.0, fy_ = 0.0, dgx_ = 0.0, dgy_ = 0.0;
    double stmin_ = 0.0, stmax_ = 0.0;
};

}  // namespace lb
