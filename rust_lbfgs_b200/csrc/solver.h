// solver.h — the device-resident L-BFGS / OWL-QN driver (host side).
//
// `Solver` is what `Problem` (src/core.rs:10-75) + `LbfgsState` (src/lbfgs.rs:425-439) become when
// every n-vector lives in HBM: it owns g/gp (ping-pong), the second x buffer, d, pg, wp and the
// 2m-vector s/y ring, launches the fused kernels of kernels.cu on one stream, and keeps only
// scalars on the host.  The host synchronises where the scalar logic needs a value: once per
// line-search trial (f, dg) and once per iteration after the two-loop (||d|| for the next step, together
// with the history sums for the `x not changed` / `gx not changed` checks); y.s, gamma and the Powell
// damping branch are consumed on the device.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/lbfgsb200.h"
#include "kernels.h"
#include "linesearch.h"

namespace lb {

struct Comm;
int comm_rank(const Comm *c);
int comm_size(const Comm *c);
int comm_allreduce_sum(Comm *c, double *buf_dev, int count, cudaStream_t stream);
int comm_allgatherv(Comm *c, const double *send, double *recv_all, const int64_t *offsets, cudaStream_t stream);
// peer-mailbox transport (nullptr: not available, the solver calls comm_allreduce_sum instead)
const PeerCtx *comm_peer(const Comm *c);
unsigned long long *comm_peer_seq(Comm *c);
int comm_peer_fault(const Comm *c);

// Launch configuration for a device (SM count, L2 size, env overrides); shared by the solver,
// the stand-alone primitives and the built-in objectives.
struct DeviceInfo {
    int device = 0;
    int sm_count = 148;
    int64_t l2_bytes = 126ll << 20;
    int blocks_per_sm = 1;   // CTAs per SM in a full grid (LBFGSB200_BLOCKS_PER_SM); tuned, see types.h
    int blocks_per_sm_trial = 2;   // the same for the line-search trial family: K2, evaluate, fused trial, probe
                                   // (LBFGSB200_TRIAL_BLOCKS_PER_SM)
};
int query_device(int device, DeviceInfo *out);
// process-wide default direction mode of solvers created afterwards (-1: back to the environment's choice)
int set_default_direction(int mode);
// returns the pages cached by the solver arenas' memory pool to the driver
int trim_pool(int device);
// allocates the level-2 reduction workspace for `info`
int alloc_reduce_ws(const DeviceInfo &info, ReduceWs *ws);
void free_reduce_ws(ReduceWs *ws);

class Solver {
  public:
    Solver() = default;
    ~Solver();
    int init(const lbfgsb200_param_t &p, int64_t n_local, int64_t n_global, int64_t goff, int device,
             cudaStream_t stream, Comm *comm);

    int build(double *x_dev, lbfgsb200_eval_fn eval, void *user);          // src/lbfgs.rs:443-481
    bool is_converged(int *stop_status);                                   // :489-494, :697-748
    int propagate(lbfgsb200_progress_t *out);                              // :503-560
    void report(lbfgsb200_report_t *out) const;                            // src/core.rs:288-298
    int finish();
    int minimize(double *x_dev, lbfgsb200_eval_fn eval, void *user, lbfgsb200_progress_fn prog, void *prog_user,
                 lbfgsb200_report_t *rep);                                 // src/lbfgs.rs:399-421

    const double *x() const { return xbuf_[cur_x_]; }
    const double *gx() const { return gbuf_[cur_g_]; }
    const double *direction() const { return d_; }
    double *spare_x() const { return x_spare_; }   // an n-vector of the arena for the host-buffer entry points
    const std::string &error() const { return err_; }

    // optional fused line-search entries of the objective (lbfgsb200_fused_ops_t); ignored for OWL-QN
    void set_trial_evaluate(lbfgsb200_trial_eval_fn fn, void *user) {
        fused_ = lbfgsb200_fused_ops_t{};
        fused_.trial = fn;
        fused_.user = user;
        spec_.count = 0;
    }
    void set_fused_ops(const lbfgsb200_fused_ops_t *ops) { fused_ = ops ? *ops : lbfgsb200_fused_ops_t{}; spec_.count = 0; }

    // LBFGSB200_DIRECTION_*: how the search direction is formed (before build(); compact needs m <= kCompactMaxM)
    int set_direction(int mode);
    int direction_mode() const { return compact_ ? LBFGSB200_DIRECTION_COMPACT : LBFGSB200_DIRECTION_TWO_LOOP; }

    // timing: 0 = off, 1 = every kernel kind, otherwise a mask: bit (1 + kind) times LBFGSB200_K_<kind> only
    void profile_enable(int timing) { timing_ = timing != 0; timing_mask_ = (timing == 1) ? ~0u : ((unsigned)timing >> 1); }
    void profile_get(lbfgsb200_profile_t *out);
    void profile_reset();

  private:
    struct Speculation;
    // scalar slots in device memory (kMaxAcc doubles each)
    // SLOT_STEP[0]: the next search's first step as formed on the device (speculative first trial)
    // SLOT_SPEC .. SLOT_SPEC + 3: the results of a multi-step probe, kSpecMax x {f, g.d, g.g, x.x} then the steps used
    enum Slot { SLOT_EVAL = 0, SLOT_HIST = 1, SLOT_LOOP_A = 2, SLOT_LOOP_B = 3, SLOT_INIT = 4, SLOT_STEP = 5, SLOT_SPEC = 6, SLOT_COUNT = 10 };
    static constexpr int kSpecMax = 6;          // trial points per multi-step probe (5 * kSpecMax doubles <= 4 slots)
    static constexpr int kHostWords = 128;      // pinned mirror: every slot (80 doubles) + the evaluate-flag staging word
    static constexpr int kFlagWord = 120;
    double *slot(int s) const { return scal_dev_ + (size_t)s * kMaxAcc; }

    int fail(int status, const char *msg);
    int cuda_fail(cudaError_t e, const char *what);
    bool evaluate_point(const double *d_or_null, double *dg_out);  // evaluate + K2/K3 + allreduce + sync
    bool trial_point(const double *xp, double stp, double *dg_out); // K1 + evaluate_point, or the fused callbacks
    // one pass for the trial at steps[0] AND the k - 1 predicted ones after it; fills `sp` with all k results
    bool trial_multi(const double *xp, const double *steps, int k, Speculation *sp);
    int multi_probe_cap() const;  // how many trial points a pass may carry here (1: the multi-step probe is unavailable)
    // allreduce (unless `exchanged`: the producer already summed over the ranks) + D2H + sync of SLOT_EVAL
    bool finish_eval(bool fused, bool exchanged, double *dg_out);
    bool use_probe() const { return fused_.probe && fused_.commit && !owl_; }
    bool use_trial() const { return fused_.trial && !owl_; }
    bool fused_exchanges() const { return (fused_.flags & LBFGSB200_FUSED_SUMS_OVER_RANKS) != 0; }
    void post_eval_flag(int erc);
    int fetch(int s, int count, double *host, bool ours = true);   // allreduce + D2H + sync of a slot
    int fetch2(int s1, int c1, double *h1, int s2, int c2, double *h2);  // two reduced slots, one sync
    int fetch_all(double *hall);                                   // every slot with one copy and one sync
    int check_peers(const double *h, int count);                   // NaN sums + the communicator's fault word => ERR_NCCL
    // ours = the slot was produced by one of our reducing kernels (already exchanged inside it with peer mailboxes)
    int reduce_across_ranks(int s, int count, bool ours = true);
    void fill_progress(lbfgsb200_progress_t *out, double step_value) const;
    Launch launch_cfg();
    // history (or the objective's commit) of one iteration, enqueued on L.stream
    int enqueue_history(const Launch &L, const double *xp, const double *gp, double step_eval);
    // (+ damping) + two-loop of one iteration, enqueued on L.stream; *so_last = slot of the final dots
    int enqueue_two_loop(const Launch &L, const double *gp, int64_t bound, int *so_last);
    // the compact direction (compact.cu): pass A + scalar recursions + pass B instead of the 2 * bound trips
    int compact_direction(const Launch &L, int64_t bound, int *so_last);
    int compact_small(const Launch &L, int64_t bound, int *so_last);   // the same in one cluster launch (n <= 2^18)
    bool small_eligible() const;   // the cluster-persistent two-loop kernel (small.cu) applies
    int two_loop_small(const Launch &L, int64_t bound, int *so_last);
    bool graph_eligible(int64_t bound) const;
    int two_loop_graphed(const Launch &L, const double *gp, int64_t bound, int *so_last);
    void drop_graphs();

    // timing instrumentation
    struct Pending { int kind; cudaEvent_t a, b; };
    void prof_begin(int kind);
    void prof_end(int kind, double bytes);
    void prof_resolve();

    lbfgsb200_param_t p_{};
    LsConfig ls_{};
    bool owl_ = false;
    int64_t owl_start_ = 0, owl_end_ = 0;
    int64_t n_ = 0, n_global_ = 0, goff_ = 0;
    int64_t m_ = 6;
    DeviceInfo dev_{};
    cudaStream_t stream_ = nullptr;
    Comm *comm_ = nullptr;
    bool streaming_ = true;
    bool sequential_ = false;     // param.reduction == LBFGSB200_REDUCE_SEQUENTIAL

    // HBM
    void *arena_ = nullptr;
    bool arena_pooled_ = false;             // from this library's private per-device memory pool
    double *xbuf_[2] = {nullptr, nullptr};  // [0] = caller's x, [1] = ours
    double *gbuf_[2] = {nullptr, nullptr};
    double *d_ = nullptr, *pg_ = nullptr, *x_spare_ = nullptr;
    signed char *wp_ = nullptr;
    std::vector<double *> S_, Y_;
    double *scal_dev_ = nullptr;    // SLOT_COUNT * kMaxAcc doubles, then alpha[m]
    double *alpha_dev_ = nullptr;
    double *ys_dev_ = nullptr;      // y.s of every ring slot (src/lbfgs.rs:613 `ys`), device-only
    double *scal_host_ = nullptr;   // pinned mirror of one slot
    size_t scal_count_ = 0;         // doubles behind scal_dev_ (the buffers are recycled, see solver.cpp)
    ReduceWs ws_{};
    int cur_x_ = 0, cur_g_ = 0;

    // host scalars
    lbfgsb200_eval_fn eval_ = nullptr;
    void *eval_user_ = nullptr;
    lbfgsb200_fused_ops_t fused_{};
    // The next iteration's first trial, probed speculatively behind the two-loop recursion (write-free, so harmless
    // if it is never used): its step and {f, g.d, g.g, x.x}.  Consumed by the next propagate() if the line search
    // asks for exactly that step.
    // With the objective's multi-step probe the trials the search is EXPECTED to take next (More-Thuente's extrapolation
    // chain, LineSearchMachine::predict) ride in the same pass: up to kSpecMax entries.
    struct Speculation { int count = 0; double step[kSpecMax] = {}; double h[kSpecMax][4] = {}; } spec_;
    int spec_k_ = 1;              // trial points to evaluate per pass: what the previous search needed (adaptive)
    int multi_probe_max_ = kSpecMax;   // LBFGSB200_MULTI_PROBE_MAX (1 disables)
    bool speculate_ = true;       // LBFGSB200_SPECULATE=0 disables
    bool built_ = false;
    double fx_ = 0.0, xx_ = 0.0, gg_ = 0.0;   // f(x), x.x, g.g (pg.pg for OWL-QN) at the current point
    double dginit_ = 0.0;                      // g.d (pg.d) for the next line search
    double step_ = 0.0;
    int64_t k_ = 0, end_ = 0, ncall_ = 0, neval_ = 0;
    int64_t last_ls_error_ = 0;
    int last_status_ = 0;
    std::string err_;

    // CUDA graphs of the update chain (launch-bound regime: n <= kGraphMaxN), one per (ring position, parity)
    static constexpr int64_t kGraphMaxN = 1 << 22;
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        int so_last = 0;
        int64_t kernels = 0;
        int64_t launches[LBFGSB200_K_COUNT] = {};
        double bytes[LBFGSB200_K_COUNT] = {};
    };
    std::vector<GraphEntry> graphs_;
    cudaStream_t cap_stream_ = nullptr;
    bool graphs_enabled_ = true;     // LBFGSB200_GRAPHS=0 disables
    bool small_enabled_ = true;      // LBFGSB200_SMALL=0 disables; cleared if the cluster launch is refused
    int64_t ring_stride_ = 0;        // doubles between consecutive ring vectors (S_0, Y_0, S_1, ...)
    int64_t graph_replays_ = 0;

    // compact direction: S^T Y, Y^T Y (m x m each), this iteration's sums, the coefficients, wide reduction partials
    bool compact_ = false;
    bool commit_gram_enabled_ = true;   // LBFGSB200_COMMIT_GRAM=0 disables; cleared when the objective declines
    int gram_fused_ = 0;                // older ring pairs whose pass-A sums this iteration's commit produced (-1: none, but
                                        // the newest pair's own two sums)
    double *cmp_block_ = nullptr;
    double *cmp_sy_ = nullptr, *cmp_yy_ = nullptr, *cmp_sums_ = nullptr, *cmp_coefs_ = nullptr, *cmp_partials_ = nullptr;

    // profile
    bool timing_ = false;
    unsigned timing_mask_ = ~0u;
    lbfgsb200_profile_t prof_{};
    std::vector<Pending> pending_;
    std::vector<cudaEvent_t> event_pool_;
    int64_t launch_counter_ = 0;
};

}  // namespace lb
