// kernels.h — host-side launchers of the sm_100a hot-path kernels (kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "types.h"

namespace lb {

// Everything a launch needs besides its operands.
struct Launch {
    cudaStream_t stream = nullptr;
    int max_grid = 148 * 8;   // CTAs: a multiple of the SM count (set from the device at create)
    int max_grid_trial = 148 * 2; // the same for K2, which shares its grid with the objectives' trial-family kernels
    bool streaming = true;    // evict-first loads/stores (vectors much larger than L2)
    bool sequential = false;  // reference-order reductions: every kernel runs as <<<1, 1>>> (validation only)
    ReduceWs ws{};            // level-2 reduction workspace (+ the peer mailboxes when sharded over GPUs)
    unsigned long long *peer_seq = nullptr;   // the communicator's exchange counter (advanced per reducing launch)
    int64_t *launch_counter = nullptr;
};

// K2  {g.d, g.g, x.x} in one read; d may be null.           src/core.rs:114-116,183-194
void launch_dots(const Launch &L, const double *g, const double *d, const double *x, int64_t n, double *out);
// K3  l1 = sum c|x| on [start,end), pseudo-gradient, pg.pg, x.x (, g.d when d != null)
//                                                            src/orthantwise.rs:70-112
void launch_owl_pg(const Launch &L, double *pg, const double *x, const double *g, const double *d, int64_t n,
                   double c, int64_t start, int64_t end, int64_t goff, double *out);
// K0  d = -src; out = {d.d, src.d}                           src/core.rs:95-101, src/lbfgs.rs:457-461
void launch_init_dir(const Launch &L, double *d, const double *src, int64_t n, double *out);
// K1  x = xp + step*d; wp != null: x = 0 where signum(x) != wp on [start,end)
//                                                            src/core.rs:155-164, src/orthantwise.rs:118-133
void launch_trial(const Launch &L, double *x, const double *xp, const double *d, double step, int64_t n,
                  const signed char *wp, int64_t start, int64_t end, int64_t goff);
// K4  wp = xp == 0 ? signum(-pg) : signum(xp)                src/core.rs:167-180
void launch_orthant(const Launch &L, signed char *wp, const double *xp, const double *pg, int64_t n);
// K5  s = x - xp; y = g - gp; out = {s.s, y.s, y.y, s.(-g), s.(gp*nstep)}
//                                                            src/lbfgs.rs:640-656,670-673, first trip of :587
//     (pg != null: OWL-QN, the first alpha uses d = -pg)
void launch_history(const Launch &L, const double *x, const double *xp, const double *g, const double *gp,
                    const double *pg, double *s, double *y, int64_t n, double nstep, bool damping, double *out);
// K6  Powell damping, decided on the device from hist = K5's sums: if y.s < 0.4 s.Bs (case 1)
//     y = ((gp*nstep)*(1-theta)) + theta*y with theta = 0.6 s.Bs / (s.Bs - y.s); otherwise the kernel exits at once
//                                                            src/lbfgs.rs:664-689
void launch_damp(const Launch &L, double *y, const double *gp, int64_t n, double nstep, const double *hist);
// K7  alpha = *red_in / *ys_in; q = (first ? -g : q) - alpha*y_j; !last: out = {s_next.q}
//     last: d = q*gamma with gamma = hist[1] / hist[2] (y.s / y.y of the newest pair), out = {y_j.d}
//     ys_in, hist, red_in are DEVICE scalars (ys_store != null: *ys_store = *ys_in)   src/lbfgs.rs:582-591,597,691
void launch_backward(const Launch &L, bool first, bool last, double *q, const double *g, const double *y,
                     const double *s_next, int64_t n, const double *red_in, const double *ys_in, double *ys_store,
                     const double *hist, double *alpha_out, double *out);
// K8  beta = *red_in / *ys_j (device scalar); r += (alpha_j - beta) s_j; !last: out = {y_next.r}
//     last: out = {r.r, g.r}; last+owl: out = {r.r (before projection), pg.d, d.d (after)}
//                                                            src/lbfgs.rs:594-601,543, src/orthantwise.rs:140-161
void launch_forward(const Launch &L, bool last, bool owl, double *r, const double *s, const double *y_next,
                    const double *g_or_pg, int64_t n, const double *red_in, const double *ys_j, const double *alpha_in,
                    int64_t start, int64_t end, int64_t goff, double *out);
// OWL-QN direction projection alone (K8's epilogue as a stand-alone op): out = {d.d after}
void launch_owl_constrain(const Launch &L, double *d, const double *pg, int64_t n, int64_t start, int64_t end,
                          int64_t goff, double *out);

// K11  LbfgsMath primitives                                   src/math.rs:31-82
void launch_vecadd(const Launch &L, double *y, const double *x, double c, int64_t n);
void launch_vecscale(const Launch &L, double *y, double c, int64_t n);
void launch_veccpy(const Launch &L, double *y, const double *x, int64_t n, bool negate);
void launch_vecdiff(const Launch &L, double *z, const double *x, const double *y, int64_t n);
void launch_vecdot(const Launch &L, const double *x, const double *y, int64_t n, double *out);
// multi-GPU: in-place sum of buf[0..count) (count <= kMailVals) over all ranks through the peer mailboxes, as a
// stand-alone 1-warp kernel (for results that were not produced by one of the reducing kernels above)
void launch_peer_allreduce(const Launch &L, double *buf, int count);

// *step_out = constrain ? min(max_step, |d|) / |d| : 1 with |d| = sqrt(dots[0])  (src/lbfgs.rs:543-551), formed on the
// device so that the next iteration's first (speculative, write-free) trial can be enqueued without a host round trip
void launch_next_step(const Launch &L, const double *dots, double max_step, bool constrain, double *step_out);

// The launch-bound regime (small.cu): the whole two-loop recursion (2 * bound trips) in ONE thread-block-cluster
// kernel; q stays in shared memory, the dependent dot products are reduced over DSMEM.  `ring` = S_0 (the ring
// vectors S_0, Y_0, S_1, ... lie `stride` doubles apart); out = {d.d before projection, g.d | pg.d, d.d after}.
int64_t two_loop_small_max_n();
cudaError_t launch_two_loop_small(const Launch &L, int device, int64_t n, int m, int bound, int slot_new, double *d,
                                  const double *dsrc, double *ring, int64_t stride, double *ys_dev, const double *hist,
                                  double *out, bool owl, int64_t start, int64_t end, int64_t goff, double max_step,
                                  bool constrain, double *step_out);


// the compact direction (below) for n <= two_loop_small_max_n(): pass A + scalar recursions + pass B in ONE cluster launch
cudaError_t launch_compact_small(const Launch &L, int device, int64_t n, int m, int bound, int slot_new, double *d,
                                 const double *dsrc, const double *ring, int64_t stride, double *ys_dev, double *SY, double *YY,
                                 const double *hist, double *out, bool owl, int64_t start, int64_t end, int64_t goff,
                                 double max_step, bool constrain, double *step_out);

// The opt-in compact search direction (compact.cu; src/lbfgs.rs:569-604 with the 2 * bound scalars derived from inner
// products of the unmodified ring vectors): pass A for the t-th newest slots given by s[] / y[] (cnt <=
// kCompactGroupMax older slots per launch, 5 sums each; `newdot` adds {y_new.d0, y_new.y_new}), the scalar recursions,
// and pass B (d written once; out = {d.d before projection, g.d | pg.d, d.d after}).
void launch_gram(const Launch &L, const double *s_new, const double *y_new, const double *src, const double *const *s,
                 const double *const *y, int cnt, bool newdot, int64_t n, double *wide_partials, double *out);
void launch_compact_solve(const Launch &L, int m, int bound, int slot_new, const double *sums, const double *hist,
                          double *SY, double *YY, double *ys_dev, double *coefs);
void launch_direction(const Launch &L, double *d, const double *src, const double *ring, int64_t stride, int64_t n, int m,
                      int bound, int slot_new, const double *coefs, bool owl, int64_t start, int64_t end, int64_t goff,
                      double *out);

}  // namespace lb
