// objectives.cu — device-resident objectives with the lbfgsb200_eval_fn signature.
//
// They exist so a solve (and the benchmark) never moves x or g over PCIe: each one reads x from
// HBM, writes g to HBM and leaves this rank's partial f in *fx_dev, all on the caller's stream.
//   Rosenbrock     default_evaluate, src/lib.rs:79-94            (streaming, 1R 1W)
//   Booth          tests/simple.rs:65-74                          (n = 2)
//   GLM            tests/owlqn.rs:22-43 (Poisson) + logistic       (dense X in HBM)
//   Lennard-Jones  examples/lj.rs:20-64,114-117                    (all pairs, FP64)
// Element-wise arithmetic keeps the reference's operation order (-fmad=false).
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>

#include "../../include/lbfgsb200.h"
#include "reduce.cuh"
#include "solver.h"

namespace lb {
namespace {

enum Kind { OBJ_ROSENBROCK = 0, OBJ_BOOTH = 1, OBJ_GLM = 2, OBJ_LJ = 3 };

struct Objective {
    int kind = 0;
    bool sequential = false;     // LBFGSB200_REDUCE_SEQUENTIAL: f summed in the reference's order by one thread
    bool recompute_gp = true;    // Rosenbrock commit: gp = grad f(xp) recomputed instead of read (LBFGSB200_COMMIT_RECOMPUTE_GP=0: read)
    DeviceInfo dev{};
    ReduceWs ws{};
    // GLM
    int glm_kind = 0;
    const double *X = nullptr, *y = nullptr;
    int64_t nrow = 0, ncol = 0;
    double *t = nullptr;         // per-row residual
    double *gpart = nullptr;     // [row_chunks][ncol] partial gradients
    int row_chunks = 0;
    double *gfused = nullptr;    // [sm_count][ncol] per-CTA partial gradients of the fused one-pass kernel
    bool fused = true;           // LBFGSB200_GLM_FUSED=0 forces the two-pass kernels
    int last_path = 0;           // which kernels the last GLM evaluation used (lbfgsb200_objective_last_path)
    // LJ
    double eps = 1.0, sigma = 1.0;
    bool lj_fast = false;        // lj_pair_fast instead of the reference's per-pair arithmetic (opt-in)
    // multi-GPU (lbfgsb200_objective_set_shard)
    Comm *comm = nullptr;            // GLM: rows of X sharded, f and g summed over ranks; LJ: atoms sharded
    std::vector<int64_t> offsets;    // LJ: element offsets of every rank's shard (nranks + 1)
    double *xall = nullptr;          // LJ: all positions, gathered before every evaluation
    double *wide_partials = nullptr; // Rosenbrock commit fused with pass A of the compact direction: 32 sums per CTA
};

// ---- Rosenbrock ------------------------------------------------------------------------------
// One (x0, x1) pair of default_evaluate: the gradient pair and the pair's term of f.  Every Rosenbrock kernel below
// goes through this one function, so evaluate, the fused trial, the probe and the commit see the same arithmetic.
__device__ __forceinline__ double rosen_pair(double x0, double x1, double2 &o) {
    const double t1 = 1.0 - x0;                             // lib.rs:85
    const double t2 = 10.0 * (x1 - x0 * x0);                // :86
    o.y = 20.0 * t2;                                        // :87
    o.x = -2.0 * (x0 * o.y + t1);                           // :88
    return t1 * t1 + t2 * t2;                               // :89
}
// x = xp + step*d for one pair (veccpy + vecadd, core.rs:156-157), then rosen_pair and the trial's four sums
// {f, g.d, g.g, x.x} (core.rs:114-116,183-194) in the order the unfused kernels add them.
__device__ __forceinline__ void rosen_trial_pair(double2 xp, double2 d, double step, double2 &xo, double2 &o,
                                                 double (&acc)[4]) {
    xo.x = xp.x + step * d.x;
    xo.y = xp.y + step * d.y;
    acc[0] += rosen_pair(xo.x, xo.y, o);
    acc[1] += o.x * d.x;
    acc[2] += o.x * o.x;
    acc[3] += xo.x * xo.x;
    acc[1] += o.y * d.y;
    acc[2] += o.y * o.y;
    acc[3] += xo.y * xo.y;
}

template <bool S>
struct RosenOp {
    const double *x;
    double *g;
    struct Regs { double2 x; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const { r.x = ld2<S>(x, i); }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[1]) const {
        double2 o;
        acc[0] += rosen_pair(r.x.x, r.x.y, o);
        st2<S>(g, i, o);
    }
    __device__ __forceinline__ void tail(int64_t, double (&)[1]) const {}
};
template <bool S>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_rosenbrock(RosenOp<S> op, int64_t n, ReduceWs ws, double *fx) {
    double acc[1] = {0.0};
    stream_pairs<1, kUt>(n, op, acc);
    grid_reduce<1>(acc, ws, fx);
}

// Fused line-search trial for Rosenbrock (lbfgsb200_trial_eval_fn): x = xp + step*d, g = grad f(x) and
// {f, g.d, g.g, x.x} in ONE pass, 2R 2W instead of K1 (2R 1W) + k_rosenbrock (1R 1W) + K2 (3R).  Same
// per-element arithmetic, same tile shape (kUt) and grid as those three kernels, so every sum sees the same
// terms in the same order: the fused and unfused paths give bit-identical scalars.
template <bool S>
struct RosenTrialOp {
    const double *xp, *d;
    double *x, *g;
    double step;
    struct Regs { double2 xp, d; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.xp = ld2<S>(xp, i);
        r.d = ld2<S>(d, i);
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[4]) const {
        double2 xo, o;
        rosen_trial_pair(r.xp, r.d, step, xo, o, acc);
        st2<S>(x, i, xo);
        st2<S>(g, i, o);
    }
    __device__ __forceinline__ void tail(int64_t, double (&)[4]) const {}
};
template <bool S>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_rosenbrock_trial(RosenTrialOp<S> op, int64_t n, ReduceWs ws, double *out) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    stream_pairs<4, kUt>(n, op, acc);
    grid_reduce<4>(acc, ws, out);
}

// Write-free trial (lbfgsb200_probe_fn): the line search only ever looks at {f, g.d} of a trial point (and the
// driver at g.g, x.x of the accepted one), so a probe reads xp and d and stores nothing — 2R instead of 2R 2W.
// Same arithmetic, tile shape and grid as the fused trial: same bits.
template <bool S>
struct RosenProbeOp {
    const double *xp, *d;
    double step;
    struct Regs { double2 xp, d; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.xp = ld2<S>(xp, i);
        r.d = ld2<S>(d, i);
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t, double (&acc)[4]) const {
        double2 xo, o;
        rosen_trial_pair(r.xp, r.d, step, xo, o, acc);
    }
    __device__ __forceinline__ void tail(int64_t, double (&)[4]) const {}
};
template <bool S>
__global__ void __launch_bounds__(kThreads, kProbePrefetch ? 1 : kMinBlocks)
k_rosenbrock_probe(RosenProbeOp<S> op, const double *step_dev, int64_t n, ReduceWs ws, double *out) {
    if (step_dev) op.step = __ldcg(step_dev);   // a speculative trial: the step was formed on the device
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (kProbePrefetch) stream_pairs_prefetch<4, kUt>(n, op, acc);
    else stream_pairs<4, kUt>(n, op, acc);
    grid_reduce<4>(acc, ws, out);
}

// Several write-free trials in ONE pass (lbfgsb200_probe_multi_fn).  A More-Thuente search whose trial is still far
// short of the minimum extrapolates: the next step is stp + 4 (stp - stx) for as long as the interval is not bracketed
// (src/line.rs:266,  update_trial_interval's clamp to tmax) — a sequence the driver can write down BEFORE the first
// result is back (with the reference's step cap of |step * d| <= 1 the first trial is short by construction once
// |x* - x| >> 1: the headline workload takes the steps s, 5 s, 21 s, 85 s in every iteration).  K trial points cost one
// read of xp and d instead of K: the pass becomes FP64-bound (25 operations per 32 bytes and step) instead of
// HBM-bound.  Per step the arithmetic, the tile shape and the order of every sum are those of k_rosenbrock_probe, so
// each of the K results has the bits the single probe would have produced; the line search consumes them only if it
// asks for exactly those steps.  out = K x {f, g.d, g.g, x.x}, then the K steps used.
template <bool S, int K>
struct RosenProbeMultiOp {
    const double *xp, *d;
    double step[K];
    struct Regs { double2 xp, d; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.xp = ld2<S>(xp, i);
        r.d = ld2<S>(d, i);
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t, double (&acc)[4 * K]) const {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double2 xo, o;
            rosen_trial_pair(r.xp, r.d, step[k], xo, o, reinterpret_cast<double (&)[4]>(acc[4 * k]));
        }
    }
    __device__ __forceinline__ void tail(int64_t, double (&)[4 * K]) const {}
};
#ifndef LB_PM_BLOCKS
#define LB_PM_BLOCKS kMinBlocks   // resident CTAs per SM of the multi-step probe (FP64-bound: tuned separately from the HBM-bound kernels)
#endif
template <bool S, int K>
__global__ void __launch_bounds__(kThreads, LB_PM_BLOCKS)
k_rosenbrock_probe_multi(RosenProbeMultiOp<S, K> op, const double *step_dev, int64_t n, ReduceWs ws, double *out) {
    if (step_dev) {   // the chain is formed here from the device's first step, with the line search's own expression
        double stx = 0.0, stp = __ldcg(step_dev);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            op.step[k] = stp;
            const double next = stp + 4.0 * (stp - stx);    // src/line.rs:266 (stmax; the clamp of update_trial_interval)
            stx = stp;
            stp = next;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) out[4 * K + k] = op.step[k];
    }
    double acc[4 * K];
#pragma unroll
    for (int a = 0; a < 4 * K; ++a) acc[a] = 0.0;
    stream_pairs<4 * K, kUt>(n, op, acc);
    grid_reduce<4 * K>(acc, ws, out);
}

// Accepted point + history update in one pass (lbfgsb200_commit_fn): x = xp + step*d and g = grad f(x) are
// recomputed with the probe's arithmetic (same bits as the accepted probe saw), written once, and
// s = x - xp, y = g - gp and IterationData::update's sums come out of the same registers — replacing the last
// trial's 2W and k_history's 4R 2W.  Tile shape and accumulation order are k_history's (kUh, history_elem).
// RECOMPUTE_GP: gp is by contract this objective's gradient at xp, and Rosenbrock's gradient is element-local, so
// rosen_pair(xp) reproduces the stored gp bit for bit from a vector the kernel reads anyway: 2R 4W instead of 3R 4W.
template <bool S, bool RECOMPUTE_GP>
struct RosenCommitOp {
    const double *xp, *d, *gp;
    double *x, *g, *s, *y;
    double step, nstep;
    struct Regs { double2 xp, d, gp; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.xp = ld2<S>(xp, i);
        r.d = ld2<S>(d, i);
        if (!RECOMPUTE_GP) r.gp = ld2<S>(gp, i);
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[5]) const {
        double2 xo, o, so, yo;
        if (RECOMPUTE_GP) rosen_pair(r.xp.x, r.xp.y, r.gp);
        xo.x = r.xp.x + step * r.d.x;                       // core.rs:156-157
        xo.y = r.xp.y + step * r.d.y;
        rosen_pair(xo.x, xo.y, o);
        history_elem<true, false>(xo.x, r.xp.x, o.x, r.gp.x, 0.0, nstep, so.x, yo.x, acc);
        history_elem<true, false>(xo.y, r.xp.y, o.y, r.gp.y, 0.0, nstep, so.y, yo.y, acc);
        st2<S>(x, i, xo);
        st2<S>(g, i, o);
        st2<S>(s, i, so);
        st2<S>(y, i, yo);
    }
    __device__ __forceinline__ void tail(int64_t, double (&)[5]) const {}
};
template <bool S, bool RECOMPUTE_GP>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_rosenbrock_commit(RosenCommitOp<S, RECOMPUTE_GP> op, int64_t n, ReduceWs ws, double *out) {
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    stream_pairs<5, kUh>(n, op, acc);
    grid_reduce<5>(acc, ws, out);
}

// The commit fused with pass A of the compact search direction (lbfgsb200_commit_gram_fn; csrc/compact.cu): the new
// pair (s, y) and the new gradient are in registers here, so their inner products with G older ring pairs — the new
// row and column of S^T Y, Y^T Y and S^T d0, Y^T d0 with d0 = -g — are formed before they are stored: 2R 4W + 2G R in
// one pass instead of 2R 4W followed by (3 + 2G) R.  Sums: [0, 5) the history sums of the commit, then per older pair
// {s_k.d0, y_k.d0, s.y_k, s_k.y, y.y_k}, then {y.d0, y.y}.  Same element-wise arithmetic as the commit and as k_gram.
template <bool S, int G>
struct RosenCommitGramOp {
    const double *xp, *d;
    double *x, *g, *s, *y;
    double step, nstep;
    const double *so[G > 0 ? G : 1], *yo[G > 0 ? G : 1];
    static constexpr int kAcc = 5 + 5 * G + 2;
    struct Regs { double2 xp, d; double2 so[G > 0 ? G : 1], yo[G > 0 ? G : 1]; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.xp = ld2<S>(xp, i);
        r.d = ld2<S>(d, i);
#pragma unroll
        for (int k = 0; k < G; ++k) {
            r.so[k] = ld2<S>(so[k], i);
            r.yo[k] = ld2<S>(yo[k], i);
        }
    }
    __device__ __forceinline__ void gram(double sn, double yn, double gn, const double *sk, const double *yk, double (&acc)[kAcc]) const {
        const double ng = -gn;                              // d0 = -g, core.rs:95-101
#pragma unroll
        for (int k = 0; k < G; ++k) {
            acc[5 + 5 * k + 0] += sk[k] * ng;
            acc[5 + 5 * k + 1] += yk[k] * ng;
            acc[5 + 5 * k + 2] += sn * yk[k];
            acc[5 + 5 * k + 3] += sk[k] * yn;
            acc[5 + 5 * k + 4] += yn * yk[k];
        }
        acc[5 + 5 * G + 0] += yn * ng;
        acc[5 + 5 * G + 1] += yn * yn;
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[kAcc]) const {
        double2 xo, o, sn, yn, gp;
        rosen_pair(r.xp.x, r.xp.y, gp);                     // gp recomputed from xp (see RosenCommitOp)
        xo.x = r.xp.x + step * r.d.x;                       // core.rs:156-157
        xo.y = r.xp.y + step * r.d.y;
        rosen_pair(xo.x, xo.y, o);
        double (&h)[5] = reinterpret_cast<double (&)[5]>(acc);
        history_elem<true, false>(xo.x, r.xp.x, o.x, gp.x, 0.0, nstep, sn.x, yn.x, h);
        history_elem<true, false>(xo.y, r.xp.y, o.y, gp.y, 0.0, nstep, sn.y, yn.y, h);
        double sx[G > 0 ? G : 1], yx[G > 0 ? G : 1], sy[G > 0 ? G : 1], yy[G > 0 ? G : 1];
#pragma unroll
        for (int k = 0; k < G; ++k) { sx[k] = r.so[k].x; yx[k] = r.yo[k].x; sy[k] = r.so[k].y; yy[k] = r.yo[k].y; }
        gram(sn.x, yn.x, o.x, sx, yx, acc);
        gram(sn.y, yn.y, o.y, sy, yy, acc);
        st2<S>(x, i, xo);
        st2<S>(g, i, o);
        st2<S>(s, i, sn);
        st2<S>(y, i, yn);
    }
    __device__ __forceinline__ void tail(int64_t, double (&)[kAcc]) const {}
};
#ifndef LB_CG_U
#define LB_CG_U 3        // pairs per thread per tile ...
#endif
#ifndef LB_CG_BLOCKS
#define LB_CG_BLOCKS 1   // ... and resident CTAs per SM of the fused commit + pass A kernel (tuned: profiles/r02_tuning.md)
#endif
template <bool S, int G>
__global__ void __launch_bounds__(kThreads, LB_CG_BLOCKS)
k_rosenbrock_commit_gram(RosenCommitGramOp<S, G> op, int64_t n, ReduceWs ws, double *hist, double *gram_out, double *newdot_out) {
    constexpr int kAcc = RosenCommitGramOp<S, G>::kAcc;
    double acc[kAcc];
#pragma unroll
    for (int a = 0; a < kAcc; ++a) acc[a] = 0.0;
    stream_pairs<kAcc, LB_CG_U>(n, op, acc);
    grid_reduce_split<kAcc>(acc, ws, hist, 5, gram_out, 5 * G, newdot_out);
}

// ---- Booth -----------------------------------------------------------------------------------
__global__ void k_booth(const double *x, double *g, double *fx) {
    const double x1 = x[0], x2 = x[1];
    const double a = x1 + 2.0 * x2 - 7.0;
    const double b = 2.0 * x1 + x2 - 5.0;
    *fx = a * a + b * b;                                    // powi(2), tests/simple.rs:68
    g[0] = 10.0 * x1 + 8.0 * x2 - 34.0;                     // :69
    g[1] = 8.0 * x1 + 10.0 * x2 - 38.0;                     // :70
}

// ---- GLM, pass 1: z = X w per row (one warp per row), f terms, residual t ---------------------
// kind 0 Poisson:  f += -(y z - exp z) ... accumulated as (y z - exp z) and negated once, t = y - exp z
// kind 1 logistic: f += softplus(z) - y z,                                                t = sigmoid(z) - y
__global__ void __launch_bounds__(kThreads) k_glm_rows(const double *__restrict__ X, const double *__restrict__ y,
                                                       const double *__restrict__ w, double *__restrict__ t,
                                                       int64_t nrow, int64_t ncol, int kind, ReduceWs ws, double *fx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    double acc[1] = {0.0};
    for (int64_t r = warp0; r < nrow; r += nwarps) {
        const double *row = X + r * ncol;
        double z = 0.0;
        for (int64_t c = lane; c < ncol; c += 32) z += w[c] * row[c];
        z = warp_sum(z);
        z = __shfl_sync(0xffffffffu, z, 0);
        if (lane == 0) {
            const double yr = y[r];
            if (kind == 0) {
                const double e = exp(z);
                acc[0] += yr * z - e;
                t[r] = yr - e;
            } else {
                const double sp = fmax(z, 0.0) + log1p(exp(-fabs(z)));
                double mu;
                if (z >= 0.0) mu = 1.0 / (1.0 + exp(-z));
                else { const double e = exp(z); mu = e / (1.0 + e); }
                acc[0] += sp - yr * z;
                t[r] = mu - yr;
            }
        }
    }
    if (kind == 0) acc[0] = -1.0 * acc[0];
    grid_reduce<1>(acc, ws, fx);
}

// ---- GLM, pass 2: partial gradients over row chunks (thread = column; coalesced rows) ---------
__global__ void __launch_bounds__(kThreads) k_glm_grad_partial(const double *__restrict__ X, const double *__restrict__ t,
                                                               double *__restrict__ gpart, int64_t nrow, int64_t ncol,
                                                               int rows_per_chunk, int kind) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    int64_t r1 = r0 + rows_per_chunk;
    if (r1 > nrow) r1 = nrow;
    if (c >= ncol) return;
    double a = 0.0;
    if (kind == 0) for (int64_t r = r0; r < r1; ++r) a += t[r] * (-X[r * ncol + c]);   // (-X^T) t, owlqn.rs:40
    else for (int64_t r = r0; r < r1; ++r) a += t[r] * X[r * ncol + c];
    gpart[(int64_t)blockIdx.y * ncol + c] = a;
}
__global__ void __launch_bounds__(kThreads) k_glm_grad_final(const double *__restrict__ gpart, double *__restrict__ g,
                                                             int64_t ncol, int chunks) {
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= ncol) return;
    double a = 0.0;
    for (int k = 0; k < chunks; ++k) a += gpart[(int64_t)k * ncol + c];
    g[c] = a;
}


// ---- GLM, fused: ONE pass over X ---------------------------------------------------------------------
// z_r = X[r,:].w, the per-row loss term and residual t_r, and the gradient update g += t_r X[r,:] all while
// row r is on chip, so X (80 GB at 1e6 x 1e4) crosses HBM once per evaluation instead of twice.
//   * one 256-thread CTA per SM owns row groups of R consecutive rows (R = 1 for 80 KB rows, up to 8 for short
//     ones: ~64 KB per group, so the two block barriers per group are amortised);
//   * groups are staged in shared memory by 1-D TMA bulk copies (cp.async.bulk -> mbarrier complete_tx), a ring
//     of `stages` buffers so the next groups are in flight while this one is consumed (2 stages of 80 KB at
//     ncol = 1e4, more for shorter rows);
//   * every thread owns KP fixed column pairs: its slice of w and its slice of the gradient accumulator live in
//     registers for the whole kernel; the row is read from shared memory twice (dot, then axpy);
//   * per-CTA partial gradients go to gpart[cta][ncol] and are summed in CTA order by k_glm_grad_final;
//     f goes through the usual deterministic two-level reduction.
// ODD (ncol odd, e.g. the reference's 500 x 21 fixture): rows are only 8-byte aligned, so R is even (every group
// starts 16-byte aligned), elements are read from shared memory as scalars and the odd 8 bytes a short last group
// may leave are copied by hand.
// CLUSTER (ncol > 10 240: a row no longer fits one CTA's registers / shared memory): the C CTAs of a thread-block
// cluster split the COLUMNS of the same rows; each stages and keeps its slice, the partial dot products z_r meet
// through distributed shared memory (one hardware cluster barrier per row, summed in rank order by every CTA),
// and each CTA updates its own slice of the gradient.  Still one pass over X.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kGlmMaxStages = 8;
constexpr uint32_t kTmaChunk = 32768;  // bytes per bulk copy

// KP column pairs per thread, R consecutive rows per pipeline stage.  Short rows (ncol of a few thousand) cannot
// pay two block barriers each — a 16 KB row lasts 0.36 us at this SM's share of the HBM bandwidth — so a stage
// holds R rows (they are contiguous in X: one bulk copy), the R dot products are reduced together and the R
// rank-one updates applied together.
template <int KP, int R, bool ODD, bool CLUSTER>
__global__ void __launch_bounds__(kThreads, 1)
k_glm_fused(const double *__restrict__ X, const double *__restrict__ y, const double *__restrict__ w,
            double *__restrict__ gpart, int64_t nrow, int64_t ncol, int kind, int stages, uint32_t stage_stride,
            ReduceWs ws, double *fx) {
    static_assert(!(ODD && CLUSTER), "the column-split kernel takes even ncol only");
    static_assert(!CLUSTER || R == 1, "column split: one row per stage");
    extern __shared__ __align__(128) unsigned char glm_smem[];
    __shared__ __align__(8) uint64_t full_bar[kGlmMaxStages];
    __shared__ double red[2][R][kWarps];
    __shared__ double zcta[2][R];                    // CLUSTER: this CTA's partial dot products, read by its peers
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    namespace cgx = cooperative_groups;
    cgx::cluster_group cluster = cgx::this_cluster();
    const int csize = CLUSTER ? (int)cluster.num_blocks() : 1;
    const int crank = CLUSTER ? (int)cluster.block_rank() : 0;
    const int64_t cid = blockIdx.x / csize;          // which row-group stream this CTA (cluster) walks
    const int64_t ncl = gridDim.x / csize;
    const int64_t npairs_all = (ncol + 1) >> 1;      // ODD: the last pair has one element
    // this CTA's columns: pairs [p0, p1)
    const int64_t chunk = (npairs_all + csize - 1) / csize;
    const int64_t p0 = (int64_t)crank * chunk;
    const int64_t p1 = (p0 + chunk < npairs_all) ? (p0 + chunk) : npairs_all;
    const uint32_t row_bytes = (uint32_t)(ncol * 8);                 // a full row of X
    const uint32_t slice_bytes = CLUSTER ? (uint32_t)((p1 > p0 ? p1 - p0 : 0) * 16) : row_bytes;   // what this CTA stages per row
    const int64_t ngroups = (nrow + R - 1) / R;      // groups of R rows; the last one may be short
    const int64_t my_groups = (ngroups > cid) ? (ngroups - cid + ncl - 1) / ncl : 0;

    double2 wv[KP], gv[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int64_t p = p0 + tid + (int64_t)k * kThreads;
        wv[k] = make_double2(0.0, 0.0);
        if (p < p1) {
            if (ODD) {
                wv[k].x = w[2 * p];
                if (2 * p + 1 < ncol) wv[k].y = w[2 * p + 1];
            } else {
                wv[k] = reinterpret_cast<const double2 *>(w)[p];
            }
        }
        gv[k] = make_double2(0.0, 0.0);
    }
    auto rows_in = [&](int64_t i) {  // rows of this CTA's i-th group
        const int64_t r0 = (cid + i * ncl) * R;
        return (int)((nrow - r0 < R) ? (nrow - r0) : R);
    };
    auto issue = [&](int64_t i) {  // thread 0: stage group i of this CTA
        const int s = (int)(i % stages);
        const int64_t r0 = (cid + i * ncl) * R;
        const uint32_t bytes = CLUSTER ? slice_bytes : row_bytes * (uint32_t)rows_in(i);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(X + r0 * ncol) + (CLUSTER ? p0 * 16 : 0);
        unsigned char *dst = glm_smem + (size_t)s * stage_stride;
        const uint32_t bulk = ODD ? (bytes & ~15u) : bytes;           // ODD: a short last group may end on 8 bytes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this buffer
        mbar_expect_tx(&full_bar[s], bulk);
        for (uint32_t off = 0; off < bulk; off += kTmaChunk) {
            const uint32_t nb = (bulk - off < kTmaChunk) ? (bulk - off) : kTmaChunk;
            tma_load_1d(dst + off, src + off, nb, &full_bar[s]);
        }
        if (ODD && bulk != bytes)   // consumers see it through the block barriers between here and their reads
            *reinterpret_cast<double *>(dst + bulk) = *reinterpret_cast<const double *>(src + bulk);
    };
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int64_t i = 0; i < stages && i < my_groups; ++i) issue(i);
    if (ODD) __syncthreads();   // the hand-copied tail of a prologue group

    double facc = 0.0;   // lanes 0..R-1 of warp 0 accumulate the loss terms of "their" row of every group
    auto load_y = [&](int64_t i) {
        const int64_t r = (cid + i * ncl) * R + lane;
        return (lane < R && i < my_groups && r < nrow) ? y[r] : 0.0;
    };
    auto row_pair = [&](const unsigned char *stage, int r, int64_t p) {   // pair p of row r of a staged group
        if (ODD) {
            const double *row = reinterpret_cast<const double *>(stage + (size_t)r * row_bytes);
            return make_double2(row[2 * p], (2 * p + 1 < ncol) ? row[2 * p + 1] : 0.0);
        }
        return reinterpret_cast<const double2 *>(stage + (size_t)r * (CLUSTER ? slice_bytes : row_bytes))[p - p0];
    };
    double y_next = load_y(0);
    for (int64_t i = 0; i < my_groups; ++i) {
        const int s = (int)(i % stages);
        const uint32_t parity = (uint32_t)((i / stages) & 1);
        const int nr = (R == 1) ? 1 : rows_in(i);   // R == 1: every group is one full row (compile-time)
        const double yr = y_next;
        y_next = load_y(i + 1);
        mbar_wait(&full_bar[s], parity);
        const unsigned char *stage = glm_smem + (size_t)s * stage_stride;
        double part[R];
#pragma unroll
        for (int r = 0; r < R; ++r) part[r] = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const int64_t p = p0 + tid + (int64_t)k * kThreads;
            if (p < p1) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r < nr) {
                        const double2 xv = row_pair(stage, r, p);
                        part[r] += wv[k].x * xv.x;
                        part[r] += wv[k].y * xv.y;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double v = warp_sum(part[r]);
            if (lane == 0) red[i & 1][r][warp] = v;
        }
        __syncthreads();
        if (CLUSTER) {   // this CTA's partial z, then the peers' through DSMEM
            if (tid < nr) {
                double zc = 0.0;
#pragma unroll
                for (int q = 0; q < kWarps; ++q) zc += red[i & 1][tid][q];
                zcta[i & 1][tid] = zc;
            }
            cluster.sync();
        }
        // lane r of every warp finishes row r: z, the link function, t; warp 0 also books the loss term
        double t_mine = 0.0;
        if (lane < nr) {
            double z = 0.0;
            if (CLUSTER) {
                for (int rk = 0; rk < csize; ++rk) {   // fixed rank order: the same z in every CTA of the cluster
                    const double(*peer)[R] = cluster.map_shared_rank(zcta, rk);
                    z += peer[i & 1][lane];
                }
            } else {
#pragma unroll
                for (int q = 0; q < kWarps; ++q) z += red[i & 1][lane][q];
            }
            double term;
            if (kind == 0) {
                const double e = exp(z);
                term = yr * z - e;
                t_mine = -(yr - e);                       // g = (-X^T) t, tests/owlqn.rs:40
            } else {
                const double sp = fmax(z, 0.0) + log1p(exp(-fabs(z)));
                double mu;
                if (z >= 0.0) mu = 1.0 / (1.0 + exp(-z));
                else { const double e = exp(z); mu = e / (1.0 + e); }
                term = sp - yr * z;
                t_mine = mu - yr;
            }
            if (warp == 0 && crank == 0) facc += term;    // one CTA of a cluster books the row's loss term
        }
        double t[R];
#pragma unroll
        for (int r = 0; r < R; ++r) t[r] = __shfl_sync(0xffffffffu, t_mine, r);
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const int64_t p = p0 + tid + (int64_t)k * kThreads;
            if (p < p1) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r < nr) {
                        const double2 xv = row_pair(stage, r, p);
                        gv[k].x += t[r] * xv.x;
                        gv[k].y += t[r] * xv.y;
                    }
                }
            }
        }
        __syncthreads();  // every thread is done with stage s: it may be refilled
        if (tid == 0 && i + stages < my_groups) issue(i + stages);
    }
    double *gp = gpart + cid * ncol;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int64_t p = p0 + tid + (int64_t)k * kThreads;
        if (p < p1) {
            if (ODD) {
                gp[2 * p] = gv[k].x;
                if (2 * p + 1 < ncol) gp[2 * p + 1] = gv[k].y;
            } else {
                reinterpret_cast<double2 *>(gp)[p] = gv[k];
            }
        }
    }
    double acc[1] = {(kind == 0) ? -1.0 * facc : facc};
    grid_reduce<1>(acc, ws, fx);
    if (CLUSTER) cluster.sync();   // no CTA exits while a peer may still read its zcta
}

// function attributes are per device: one opt-in per (instantiation, device)
template <int KP, int R, bool ODD, bool CLUSTER>
bool glm_fused_attrs(int device, bool nonportable) {
    static bool attr_set[64] = {};
    if (device >= 0 && device < 64 && attr_set[device]) return true;
    if (cudaFuncSetAttribute(k_glm_fused<KP, R, ODD, CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        (nonportable && cudaFuncSetAttribute(k_glm_fused<KP, R, ODD, CLUSTER>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
        cudaGetLastError();
        return false;
    }
    if (device >= 0 && device < 64) attr_set[device] = true;
    return true;
}

template <int KP, int R, bool ODD>
int launch_glm_fused(Objective *o, const double *w, double *g, cudaStream_t stream, double *fx) {
    const uint32_t row_bytes = (uint32_t)(o->ncol * 8);
    const uint32_t stage_stride = (row_bytes * (uint32_t)R + 127u) & ~127u;
    int stages = (int)((200u * 1024u) / stage_stride);
    if (stages > kGlmMaxStages) stages = kGlmMaxStages;
    if (stages < 2) return LBFGSB200_ERR_UNSUPPORTED;
    const size_t smem = (size_t)stages * stage_stride;
    if (!glm_fused_attrs<KP, R, ODD, false>(o->dev.device, false)) return LBFGSB200_ERR_UNSUPPORTED;   // two-pass kernels instead
    const int64_t ngroups = (o->nrow + R - 1) / R;
    int grid = o->dev.sm_count;
    if ((int64_t)grid > ngroups) grid = (int)ngroups;
    k_glm_fused<KP, R, ODD, false><<<grid, kThreads, smem, stream>>>(o->X, o->y, w, o->gfused, o->nrow, o->ncol, o->glm_kind,
                                                                      stages, stage_stride, o->ws, fx);
    if (cudaGetLastError() != cudaSuccess) return LBFGSB200_ERR_UNSUPPORTED;   // launch refused: two-pass kernels instead
    const unsigned cb = (unsigned)((o->ncol + kThreads - 1) / kThreads);
    k_glm_grad_final<<<cb, kThreads, 0, stream>>>(o->gfused, g, o->ncol, grid);
    return 0;
}

// ncol > 10 240 (even): C CTAs of a cluster split the columns, KP = 20 pairs per thread each.
int launch_glm_fused_cluster(Objective *o, const double *w, double *g, cudaStream_t stream, double *fx) {
    const int64_t npairs = o->ncol / 2;
    int c = 2;
    while (c < 16 && (npairs + c - 1) / c > 20 * kThreads) c *= 2;
    if ((npairs + c - 1) / c > 20 * kThreads) return LBFGSB200_ERR_UNSUPPORTED;
    const uint32_t slice = (uint32_t)(((npairs + c - 1) / c) * 16);
    const uint32_t stage_stride = (slice + 127u) & ~127u;
    int stages = (int)((200u * 1024u) / stage_stride);
    if (stages > kGlmMaxStages) stages = kGlmMaxStages;
    if (stages < 2) return LBFGSB200_ERR_UNSUPPORTED;
    if (!glm_fused_attrs<20, 1, false, true>(o->dev.device, c > 8)) return LBFGSB200_ERR_UNSUPPORTED;
    int nclusters = o->dev.sm_count / c;
    if ((int64_t)nclusters > o->nrow) nclusters = (int)o->nrow;
    if (nclusters < 1) return LBFGSB200_ERR_UNSUPPORTED;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(nclusters * c));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)stages * stage_stride;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_glm_fused<20, 1, false, true>, o->X, o->y, w, o->gfused, o->nrow, o->ncol,
                                             o->glm_kind, stages, stage_stride, o->ws, fx);
    if (e != cudaSuccess) { cudaGetLastError(); return LBFGSB200_ERR_UNSUPPORTED; }
    const unsigned cb = (unsigned)((o->ncol + kThreads - 1) / kThreads);
    k_glm_grad_final<<<cb, kThreads, 0, stream>>>(o->gfused, g, o->ncol, nclusters);
    return 0;
}

// rows per stage: enough for ~32-64 KB per stage (R * 16 B-aligned rows are contiguous in X)
template <int KP, bool ODD>
int launch_glm_fused_r(Objective *o, const double *w, double *g, cudaStream_t stream, double *fx) {
    const int64_t row_bytes = o->ncol * 8;
    if (row_bytes * 8 <= 65536) return launch_glm_fused<KP, 8, ODD>(o, w, g, stream, fx);
    if (row_bytes * 4 <= 65536) return launch_glm_fused<KP, 4, ODD>(o, w, g, stream, fx);
    if (row_bytes * 2 <= 65536 || ODD) return launch_glm_fused<KP, 2, ODD>(o, w, g, stream, fx);   // ODD: R stays even
    return launch_glm_fused<KP, 1, false>(o, w, g, stream, fx);
}

// 0 = launched; LBFGSB200_ERR_UNSUPPORTED = shape not covered (caller falls back to the two-pass kernels)
int glm_fused(Objective *o, const double *w, double *g, cudaStream_t stream, double *fx) {
    if (!o->gfused || (((uintptr_t)o->X | (uintptr_t)w) & 15u)) return LBFGSB200_ERR_UNSUPPORTED;
    const int64_t need = ((o->ncol + 1) / 2 + kThreads - 1) / kThreads;  // column pairs per thread
    if (o->ncol & 1) {   // odd ncol: scalar shared-memory reads, even R
        if (need <= 1) return launch_glm_fused_r<1, true>(o, w, g, stream, fx);
        if (need <= 2) return launch_glm_fused_r<2, true>(o, w, g, stream, fx);
        if (need <= 4) return launch_glm_fused_r<4, true>(o, w, g, stream, fx);
        if (need <= 8) return launch_glm_fused_r<8, true>(o, w, g, stream, fx);
        if (need <= 12) return launch_glm_fused<12, 2, true>(o, w, g, stream, fx);   // up to ncol = 6 143
        return LBFGSB200_ERR_UNSUPPORTED;
    }
    if (need <= 1) return launch_glm_fused_r<1, false>(o, w, g, stream, fx);
    if (need <= 2) return launch_glm_fused_r<2, false>(o, w, g, stream, fx);
    if (need <= 4) return launch_glm_fused_r<4, false>(o, w, g, stream, fx);
    if (need <= 8) return launch_glm_fused_r<8, false>(o, w, g, stream, fx);
    if (need <= 12) return launch_glm_fused<12, 1, false>(o, w, g, stream, fx);
    if (need <= 16) return launch_glm_fused<16, 1, false>(o, w, g, stream, fx);
    if (need <= 20) return launch_glm_fused<20, 1, false>(o, w, g, stream, fx);
    return launch_glm_fused_cluster(o, w, g, stream, fx);   // ncol > 10 240: columns split over a CTA cluster
}

// ---- Lennard-Jones: all pairs, FP64-pipe bound -------------------------------------------------------------
// For atom i the reference adds its pair forces in ascending partner order (pairs (i, j<i) during
// row i, then pairs (i', i) for i' > i), and (p_i - p_j) == -(p_j - p_i) exactly, so a sequential
// ascending loop over all partners reproduces forces[i] bit for bit (k_lj: one thread per atom, used for
// reference-order validation).  The production kernel (k_lj_lanes) spreads the partners of an atom over P lanes
// and combines them with a fixed butterfly: same per-pair arithmetic, different association.
constexpr int kLjTile = 256;

// a / b from the correctly rounded reciprocal y = RN(1 / b): q = RN(a y); rem = a - b q (exact, one FMA);
// RN(q + rem y) is the correctly rounded quotient (Markstein's theorem) — bit-identical to the IEEE division the
// reference performs, at 3 instructions instead of ~20.  A pair needs five divisions by the same r, so the one
// true division that yields y is shared.  (Checked against a / b on 4e8 random and adversarial operands.)
__device__ __forceinline__ double div_by(double a, double b, double y) {
    const double q = a * y;
    const double rem = __fma_rn(-b, q, a);
    return __fma_rn(rem, y, q);
}

// One ordered pair in the REFERENCE's arithmetic (examples/lj.rs:23-32,50-57): d = p_i - p_j, `other` = (j != i).
// pe = pair energy; (c0, c1, c2) = the pair's contribution to forces[i].
__device__ __forceinline__ void lj_pair_ref(double d0, double d1, double d2, bool other, double eps, double sigma,
                                            double &pe, double &c0, double &c1, double &c2) {
    const double r2 = d0 * d0 + d1 * d1 + d2 * d2;
    const double r = sqrt(other ? r2 : 1.0);                  // vecdist, lj.rs:50 (self pair: masked by the caller)
    const double y = 1.0 / r;                                 // the one true division of the pair
    const double qq = div_by(sigma, r, y);
    const double q2 = qq * qq;
    const double s6 = q2 * (q2 * q2);                         // powi(sigma/r, 6), lj.rs:23,30
    pe = 4.0 * eps * (s6 * s6 - s6);                          // pair_energy, lj.rs:51
    const double gr = div_by(24.0 * eps * (s6 - 2.0 * (s6 * s6)), r, y);  // pair_gradient, lj.rs:32
    // forces[i][k] += g*dr/r with dr = p_j - p_i = -d  (lj.rs:55-57)
    c0 = div_by(gr * (-d0), r, y);
    c1 = div_by(gr * (-d1), r, y);
    c2 = div_by(gr * (-d2), r, y);
}

// The same pair in the arithmetic a molecular-dynamics code would use (opt-in: lbfgsb200_objective_set_lj_fast):
// only 1/r^2 is needed — (sigma/r)^6 = (sigma^2/r^2)^3 and g*dr/r = 24 eps (s6 - 2 s12) dr / r^2 — so the square
// root and five of the six divisions disappear; 1/r^2 comes from the SFU's reciprocal seed and two Newton steps
// (full double precision, not correctly rounded), and multiply-adds are fused.  ~23 FP64 instructions per pair
// instead of ~60; every pair term agrees with the reference's to a few ulp.
__device__ __forceinline__ double rcp_newton(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));     // ~20 good bits (MUFU.RCP64H)
    double e = __fma_rn(-a, y, 1.0);
    y = __fma_rn(y, e, y);                                    // ~40 bits
    e = __fma_rn(-a, y, 1.0);
    return __fma_rn(y, e, y);                                 // full precision
}
__device__ __forceinline__ void lj_pair_fast(double d0, double d1, double d2, bool other, double eps4, double eps24,
                                             double sigma2, double &pe, double &c0, double &c1, double &c2) {
    const double r2 = __fma_rn(d2, d2, __fma_rn(d1, d1, d0 * d0));
    const double inv = rcp_newton(other ? r2 : 1.0);
    const double s2 = sigma2 * inv;
    const double s6 = s2 * s2 * s2;
    pe = eps4 * __fma_rn(s6, s6, -s6);
    const double w = eps24 * __fma_rn(-2.0 * s6, s6, s6) * inv;
    c0 = w * (-d0);
    c1 = w * (-d1);
    c2 = w * (-d2);
}

// Thread = atom, partners in ascending order: forces bit-identical to the reference's (validation / SEQUENTIAL).
// Sharded over GPUs: x holds ALL natoms positions (gathered), this rank owns atoms [a0, a0 + nloc) and writes
// their gradient to g[0 .. 3 nloc).
__global__ void __launch_bounds__(kLjTile) k_lj(const double *__restrict__ x, double *__restrict__ g, int64_t natoms,
                                                int64_t a0, int64_t nloc, double eps, double sigma, ReduceWs ws, double *fx) {
    __shared__ double sp[kLjTile * 3];
    const int64_t il = (int64_t)blockIdx.x * kLjTile + threadIdx.x;
    const int64_t i = a0 + il;
    const bool active = il < nloc;
    double pi0 = 0.0, pi1 = 0.0, pi2 = 0.0;
    if (active) { pi0 = x[3 * i]; pi1 = x[3 * i + 1]; pi2 = x[3 * i + 2]; }
    double f0 = 0.0, f1 = 0.0, f2 = 0.0, e = 0.0;
    for (int64_t base = 0; base < natoms; base += kLjTile) {
        const int64_t cnt = (natoms - base < kLjTile) ? (natoms - base) : kLjTile;
        __syncthreads();
        for (int64_t q = threadIdx.x; q < cnt * 3; q += kLjTile) sp[q] = x[3 * base + q];
        __syncthreads();
        if (!active) continue;
        // Branch-free body, unrolled: the sqrt / division chains of four partners are independent and overlap in
        // the FP64 pipe, while the accumulations stay in ascending partner order (bit-identical forces).
#pragma unroll 4
        for (int jj = 0; jj < (int)cnt; ++jj) {
            const int64_t j = base + jj;
            const bool other = (j != i);
            double pe, c0, c1, c2;
            lj_pair_ref(pi0 - sp[3 * jj], pi1 - sp[3 * jj + 1], pi2 - sp[3 * jj + 2], other, eps, sigma, pe, c0, c1, c2);
            if (j < i) e += pe;                                       // counted once
            if (other) { f0 += c0; f1 += c1; f2 += c2; }
        }
    }
    if (active) {  // gx = -forces, lj.rs:116
        g[3 * il] = -f0;
        g[3 * il + 1] = -f1;
        g[3 * il + 2] = -f2;
    }
    double acc[1] = {e};
    grid_reduce<1>(acc, ws, fx);
}

// The production kernel: P lanes of a warp share an atom.  Lane `sub` visits partners sub, sub + P, ... of every
// tile (ascending), and the P partial forces are combined with a fixed xor-butterfly — deterministic, but the
// association differs from the one-thread order, so forces agree with k_lj to rounding (1e-16 relative), not bit
// for bit.  P is chosen so that the grid has >= 16 CTAs per SM: one thread per atom gives 1e5 atoms only 2.7 CTAs
// per SM (30 % of the warps a B200 can hold, and a 3-vs-2 CTA imbalance between SMs); with 8 lanes per atom the
// FP64 pipe — the roofline of this kernel — stays busy.  FAST selects lj_pair_fast.
// One tile of partners for one lane.  SELF: the tile may contain atom i itself (and partners on both sides of it), so
// every pair checks j != i and j < i; otherwise those tests are hoisted out of the loop: all partners of a tile
// BELOW atom i's tile have j < i (energy counted), all of a tile ABOVE have j > i (no energy term is even computed).
// For 1e5 atoms 99.7 % of the tiles take the test-free paths.
template <int P, bool FAST, bool SELF, bool BELOW>
__device__ __forceinline__ void lj_tile(const double *__restrict__ sp, int cnt, int sub, int base, int i, double pi0,
                                        double pi1, double pi2, double eps, double sigma, double eps4, double eps24,
                                        double sigma2, double &f0, double &f1, double &f2, double &e) {
#pragma unroll 4
    for (int jj = sub; jj < cnt; jj += P) {
        const bool other = SELF ? (base + jj != i) : true;
        const double d0 = pi0 - sp[3 * jj], d1 = pi1 - sp[3 * jj + 1], d2 = pi2 - sp[3 * jj + 2];
        double pe, c0, c1, c2;
        if (FAST) lj_pair_fast(d0, d1, d2, other, eps4, eps24, sigma2, pe, c0, c1, c2);
        else lj_pair_ref(d0, d1, d2, other, eps, sigma, pe, c0, c1, c2);
        if (SELF) {
            if (base + jj < i) e += pe;
            if (other) { f0 += c0; f1 += c1; f2 += c2; }
        } else {
            if (BELOW) e += pe;
            f0 += c0;
            f1 += c1;
            f2 += c2;
        }
    }
}

template <int P, bool FAST>
__global__ void __launch_bounds__(kLjTile) k_lj_lanes(const double *__restrict__ x, double *__restrict__ g, int64_t natoms,
                                                      int64_t a0, int64_t nloc, double eps, double sigma, ReduceWs ws,
                                                      double *fx) {
    __shared__ double sp[kLjTile * 3];
    constexpr int kAtoms = kLjTile / P;
    const int sub = threadIdx.x % P;
    const int64_t il = (int64_t)blockIdx.x * kAtoms + threadIdx.x / P;
    const int i = (int)(a0 + il);                 // atoms are indexed with 32 bits (3 * natoms * 8 B would be 51 GB at 2^31)
    const bool active = il < nloc;
    double pi0 = 0.0, pi1 = 0.0, pi2 = 0.0;
    if (active) { pi0 = x[3 * (int64_t)i]; pi1 = x[3 * (int64_t)i + 1]; pi2 = x[3 * (int64_t)i + 2]; }
    const double eps4 = 4.0 * eps, eps24 = 24.0 * eps, sigma2 = sigma * sigma;
    // the tiles that hold this CTA's own atoms (CTA-uniform): only they need the per-pair self / order tests
    const int first_atom = (int)(a0 + (int64_t)blockIdx.x * kAtoms);
    const int tile_lo = first_atom / kLjTile, tile_hi = (first_atom + kAtoms - 1) / kLjTile;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0, e = 0.0;
    const int ntiles = (int)((natoms + kLjTile - 1) / kLjTile);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int base = tile * kLjTile;
        const int cnt = (natoms - base < kLjTile) ? (int)(natoms - base) : kLjTile;
        __syncthreads();
        for (int q = threadIdx.x; q < cnt * 3; q += kLjTile) sp[q] = x[3 * (int64_t)base + q];
        __syncthreads();
        if (!active) continue;
        if (tile < tile_lo) lj_tile<P, FAST, false, true>(sp, cnt, sub, base, i, pi0, pi1, pi2, eps, sigma, eps4, eps24, sigma2, f0, f1, f2, e);
        else if (tile > tile_hi) lj_tile<P, FAST, false, false>(sp, cnt, sub, base, i, pi0, pi1, pi2, eps, sigma, eps4, eps24, sigma2, f0, f1, f2, e);
        else lj_tile<P, FAST, true, false>(sp, cnt, sub, base, i, pi0, pi1, pi2, eps, sigma, eps4, eps24, sigma2, f0, f1, f2, e);
    }
#pragma unroll
    for (int off = P / 2; off > 0; off >>= 1) {
        f0 += __shfl_xor_sync(0xffffffffu, f0, off);
        f1 += __shfl_xor_sync(0xffffffffu, f1, off);
        f2 += __shfl_xor_sync(0xffffffffu, f2, off);
    }
    if (active && sub == 0) {
        g[3 * il] = -f0;
        g[3 * il + 1] = -f1;
        g[3 * il + 2] = -f2;
    }
    double acc[1] = {e};
    grid_reduce<1>(acc, ws, fx);
}

template <int P>
void launch_lj_lanes(Objective *o, const double *x, double *g, int64_t natoms, int64_t a0, int64_t nloc, cudaStream_t stream,
                     double *fx) {
    const int64_t blocks = (nloc + kLjTile / P - 1) / (kLjTile / P);
    if (o->lj_fast) k_lj_lanes<P, true><<<(int)blocks, kLjTile, 0, stream>>>(x, g, natoms, a0, nloc, o->eps, o->sigma, o->ws, fx);
    else k_lj_lanes<P, false><<<(int)blocks, kLjTile, 0, stream>>>(x, g, natoms, a0, nloc, o->eps, o->sigma, o->ws, fx);
}

// Lanes per atom: the smallest power of two that gives the grid >= 16 CTAs per SM (at most a warp per atom).
inline int lj_lanes(const Objective *o, int64_t nloc) {
    const int64_t want = (int64_t)o->dev.sm_count * 16 * kLjTile;
    int P = 1;
    while (P < 32 && nloc * P < want) P *= 2;
    return P;
}
// The level-2 reduction workspace must hold one partial per CTA: grow it for large systems.
inline int lj_ensure_ws(Objective *o, int64_t blocks) {
    if (blocks <= o->ws.stride) return 0;
    ReduceWs bigger{};
    bigger.stride = (int)blocks;
    if (cudaMalloc((void **)&bigger.partials, sizeof(double) * kMaxAcc * (size_t)bigger.stride) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    cudaFree(o->ws.partials);       // synchronises: nothing of ours is still running on the old workspace
    o->ws.partials = bigger.partials;
    o->ws.stride = bigger.stride;
    return 0;
}
int launch_lj(Objective *o, const double *x, double *g, int64_t natoms, int64_t a0, int64_t nloc, cudaStream_t stream, double *fx) {
    const int P = o->sequential ? 1 : lj_lanes(o, nloc);
    const int64_t blocks = (nloc * P + kLjTile - 1) / kLjTile;
    const int rc = lj_ensure_ws(o, blocks);
    if (rc != 0) return rc;
    switch (P) {
        case 1: k_lj<<<(int)blocks, kLjTile, 0, stream>>>(x, g, natoms, a0, nloc, o->eps, o->sigma, o->ws, fx); break;
        case 2: launch_lj_lanes<2>(o, x, g, natoms, a0, nloc, stream, fx); break;
        case 4: launch_lj_lanes<4>(o, x, g, natoms, a0, nloc, stream, fx); break;
        case 8: launch_lj_lanes<8>(o, x, g, natoms, a0, nloc, stream, fx); break;
        case 16: launch_lj_lanes<16>(o, x, g, natoms, a0, nloc, stream, fx); break;
        default: launch_lj_lanes<32>(o, x, g, natoms, a0, nloc, stream, fx); break;
    }
    return 0;
}

// Reference-order energy (LBFGSB200_REDUCE_SEQUENTIAL): one thread walks the pairs (i, j < i) exactly as
// examples/lj.rs:48-52 does, so the energy is the same left-to-right fold (forces above already are).
__global__ void k_lj_energy_seq(const double *__restrict__ x, int64_t natoms, double eps, double sigma, double *fx) {
    double e = 0.0;
    for (int64_t i = 0; i < natoms; ++i) {
        for (int64_t j = 0; j < i; ++j) {
            const double d0 = x[3 * i] - x[3 * j], d1 = x[3 * i + 1] - x[3 * j + 1], d2 = x[3 * i + 2] - x[3 * j + 2];
            const double r = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            const double qq = sigma / r;
            const double q2 = qq * qq;
            const double s6 = q2 * (q2 * q2);
            e += 4.0 * eps * (s6 * s6 - s6);
        }
    }
    *fx = e;
}

inline int stream_grid(const Objective *o, int64_t n, int U = kU) {
    if (o->sequential) return 1;
    const int64_t tile = (int64_t)kThreads * U;
    int64_t tiles = ((n >> 1) + tile - 1) / tile;
    // the kUt family (evaluate, fused trial, probe — and the solver's K2) shares one grid so that all paths sum alike
    const int64_t cap = (int64_t)o->dev.sm_count * (U == kUt ? o->dev.blocks_per_sm_trial : o->dev.blocks_per_sm);
    if (tiles > cap) tiles = cap;
    if (tiles < 1) tiles = 1;
    return (int)tiles;
}

// The reduction workspace of one fused line-search launch.  Sharded over GPUs (set_shard gave us the communicator)
// the kernel sums its scalars over the ranks in its own epilogue through the peer mailboxes — the same exchange,
// and the same sequence counter, as the solver's own reducing kernels (every rank issues the same launches in the
// same order) — so a trial on N GPUs costs no extra launch.
inline bool sums_over_ranks(const Objective *o) {
    return o->kind == OBJ_ROSENBROCK && o->comm && comm_size(o->comm) > 1 && comm_peer(o->comm) != nullptr;
}
inline ReduceWs fused_ws(Objective *o) {
    ReduceWs ws = o->ws;
    ws.peer = PeerCtx{};
    if (sums_over_ranks(o)) {
        ws.peer = *comm_peer(o->comm);
        ws.peer.seq = ++*comm_peer_seq(o->comm);
        ws.peer.extra[0] = ws.peer.extra[1] = nullptr;
    }
    return ws;
}
// 4 vectors in flight per trial; the same L2 rule as the solver's (working set vs 0.75 L2)
inline bool rosen_streaming(const Objective *o, int64_t n) { return (double)n * 16.0 > 0.75 * (double)o->dev.l2_bytes; }

int trial_impl(Objective *o, const double *xp, const double *d, double step, double *x, double *g, int64_t n,
               cudaStream_t stream, double *out) {
    if (o->kind != OBJ_ROSENBROCK) return LBFGSB200_ERR_UNSUPPORTED;
    if (n & 1) return LBFGSB200_ERR_INVALID_PARAM;
    if (cudaSetDevice(o->dev.device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    const int grid = stream_grid(o, n, kUt);
    const int threads = o->sequential ? 1 : kThreads;
    if (rosen_streaming(o, n)) k_rosenbrock_trial<true><<<grid, threads, 0, stream>>>({xp, d, x, g, step}, n, fused_ws(o), out);
    else k_rosenbrock_trial<false><<<grid, threads, 0, stream>>>({xp, d, x, g, step}, n, fused_ws(o), out);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

int probe_impl(Objective *o, const double *xp, const double *d, double step, const double *step_dev, int64_t n,
               cudaStream_t stream, double *out) {
    if (o->kind != OBJ_ROSENBROCK) return LBFGSB200_ERR_UNSUPPORTED;
    if (n & 1) return LBFGSB200_ERR_INVALID_PARAM;
    if (cudaSetDevice(o->dev.device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    const int grid = stream_grid(o, n, kUt);
    const int threads = o->sequential ? 1 : kThreads;
    if (rosen_streaming(o, n)) k_rosenbrock_probe<true><<<grid, threads, 0, stream>>>({xp, d, step}, step_dev, n, fused_ws(o), out);
    else k_rosenbrock_probe<false><<<grid, threads, 0, stream>>>({xp, d, step}, step_dev, n, fused_ws(o), out);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

template <int K>
int probe_multi_launch(Objective *o, const double *xp, const double *d, const double *steps, const double *step_dev, int64_t n,
                       cudaStream_t stream, double *out) {
    const int grid = stream_grid(o, n, kUt);
    const int threads = o->sequential ? 1 : kThreads;
    ReduceWs ws = fused_ws(o);
    ws.partials = o->wide_partials;   // 4 K rows
    if (rosen_streaming(o, n)) {
        RosenProbeMultiOp<true, K> op{xp, d, {}};
        for (int k = 0; k < K; ++k) op.step[k] = steps ? steps[k] : 0.0;
        k_rosenbrock_probe_multi<true, K><<<grid, threads, 0, stream>>>(op, step_dev, n, ws, out);
    } else {
        RosenProbeMultiOp<false, K> op{xp, d, {}};
        for (int k = 0; k < K; ++k) op.step[k] = steps ? steps[k] : 0.0;
        k_rosenbrock_probe_multi<false, K><<<grid, threads, 0, stream>>>(op, step_dev, n, ws, out);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

inline int ensure_wide_partials(Objective *o) {
    if (o->wide_partials) return 0;
    if (cudaMalloc((void **)&o->wide_partials, sizeof(double) * (size_t)(5 * kCompactGroupMax + 7) * (size_t)o->ws.stride) != cudaSuccess) {
        o->wide_partials = nullptr;
        cudaGetLastError();
        return LBFGSB200_ERR_UNSUPPORTED;
    }
    return 0;
}

// k trial points in one pass: steps[0 .. k) from the host, or (step_dev != null) the extrapolation chain formed on the
// device from *step_dev.  k <= 6; with the peer exchange in the epilogue the 4 k sums must fit one mailbox entry.
int probe_multi_impl(Objective *o, const double *xp, const double *d, const double *steps, const double *step_dev, int k,
                     int64_t n, cudaStream_t stream, double *out) {
    if (o->kind != OBJ_ROSENBROCK) return LBFGSB200_ERR_UNSUPPORTED;
    if (k < 1 || k > 6 || (!steps && !step_dev) || (sums_over_ranks(o) && 4 * k > kMailVals)) return LBFGSB200_ERR_INVALID_PARAM;
    if (n & 1) return LBFGSB200_ERR_INVALID_PARAM;
    if (cudaSetDevice(o->dev.device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    if (ensure_wide_partials(o) != 0) return LBFGSB200_ERR_UNSUPPORTED;
    switch (k) {
        case 1: return probe_multi_launch<1>(o, xp, d, steps, step_dev, n, stream, out);
        case 2: return probe_multi_launch<2>(o, xp, d, steps, step_dev, n, stream, out);
        case 3: return probe_multi_launch<3>(o, xp, d, steps, step_dev, n, stream, out);
        case 4: return probe_multi_launch<4>(o, xp, d, steps, step_dev, n, stream, out);
        case 5: return probe_multi_launch<5>(o, xp, d, steps, step_dev, n, stream, out);
        default: return probe_multi_launch<6>(o, xp, d, steps, step_dev, n, stream, out);
    }
}

int commit_impl(Objective *o, const double *xp, const double *d, const double *gp, double step, double bs_scale,
                double *x, double *g, double *s, double *y, int64_t n, cudaStream_t stream, double *out) {
    if (o->kind != OBJ_ROSENBROCK) return LBFGSB200_ERR_UNSUPPORTED;
    if (n & 1) return LBFGSB200_ERR_INVALID_PARAM;
    if (cudaSetDevice(o->dev.device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    const int grid = stream_grid(o, n, kUh);
    const int threads = o->sequential ? 1 : kThreads;
    const bool st = rosen_streaming(o, n);
    if (o->recompute_gp) {
        if (st) k_rosenbrock_commit<true, true><<<grid, threads, 0, stream>>>({xp, d, gp, x, g, s, y, step, bs_scale}, n, fused_ws(o), out);
        else k_rosenbrock_commit<false, true><<<grid, threads, 0, stream>>>({xp, d, gp, x, g, s, y, step, bs_scale}, n, fused_ws(o), out);
    } else {
        if (st) k_rosenbrock_commit<true, false><<<grid, threads, 0, stream>>>({xp, d, gp, x, g, s, y, step, bs_scale}, n, fused_ws(o), out);
        else k_rosenbrock_commit<false, false><<<grid, threads, 0, stream>>>({xp, d, gp, x, g, s, y, step, bs_scale}, n, fused_ws(o), out);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

template <int G>
int commit_gram_launch(Objective *o, const double *xp, const double *d, double step, double bs_scale, double *x, double *g,
                       double *s, double *y, const double *const *s_old, const double *const *y_old, int64_t n,
                       cudaStream_t stream, double *hist, double *gram_out, double *newdot_out) {
    int grid = 1;
    if (!o->sequential) {
        const int64_t tile = (int64_t)kThreads * LB_CG_U, nv = n >> 1;
        int64_t tiles = (nv + tile - 1) / tile;
        if (tiles < 1) tiles = 1;
        if (tiles > (int64_t)o->dev.sm_count * LB_CG_BLOCKS) tiles = (int64_t)o->dev.sm_count * LB_CG_BLOCKS;
        grid = (int)tiles;
    }
    const int threads = o->sequential ? 1 : kThreads;
    ReduceWs ws = o->ws;
    ws.partials = o->wide_partials;
    ws.peer.nranks = 0;
    if (rosen_streaming(o, n)) {
        RosenCommitGramOp<true, G> op{xp, d, x, g, s, y, step, bs_scale, {}, {}};
        for (int k = 0; k < G; ++k) { op.so[k] = s_old[k]; op.yo[k] = y_old[k]; }
        k_rosenbrock_commit_gram<true, G><<<grid, threads, 0, stream>>>(op, n, ws, hist, gram_out, newdot_out);
    } else {
        RosenCommitGramOp<false, G> op{xp, d, x, g, s, y, step, bs_scale, {}, {}};
        for (int k = 0; k < G; ++k) { op.so[k] = s_old[k]; op.yo[k] = y_old[k]; }
        k_rosenbrock_commit_gram<false, G><<<grid, threads, 0, stream>>>(op, n, ws, hist, gram_out, newdot_out);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

int commit_gram_impl(Objective *o, const double *xp, const double *d, double step, double bs_scale, double *x, double *g,
                     double *s, double *y, const double *const *s_old, const double *const *y_old, int n_old, int64_t n,
                     cudaStream_t stream, double *hist, double *gram_out, double *newdot_out) {
    if (o->kind != OBJ_ROSENBROCK || !o->recompute_gp || n_old < 0 || n_old > kCompactGroupMax) return LBFGSB200_ERR_UNSUPPORTED;
    // always rank-local partials (5 n_old + 7 sums do not fit one mailbox entry): the solver sums them over the ranks
    if (n & 1) return LBFGSB200_ERR_INVALID_PARAM;
    if (cudaSetDevice(o->dev.device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    if (ensure_wide_partials(o) != 0) return LBFGSB200_ERR_UNSUPPORTED;   // the caller runs commit + pass A separately
#define LB_CG(G) commit_gram_launch<G>(o, xp, d, step, bs_scale, x, g, s, y, s_old, y_old, n, stream, hist, gram_out, newdot_out)
    switch (n_old) {
        case 0: return LB_CG(0);
        case 1: return LB_CG(1);
        case 2: return LB_CG(2);
        case 3: return LB_CG(3);
        case 4: return LB_CG(4);
        default: return LB_CG(5);
    }
#undef LB_CG
}

int eval_impl(Objective *o, const double *x, double *g, int64_t n, cudaStream_t stream, double *fx) {
    if (cudaSetDevice(o->dev.device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    switch (o->kind) {
        case OBJ_ROSENBROCK: {
            if (n & 1) return LBFGSB200_ERR_INVALID_PARAM;  // the reference indexes x[i+1] (lib.rs:86)
            const int grid = stream_grid(o, n, kUt);
            const int threads = o->sequential ? 1 : kThreads;
            if (rosen_streaming(o, n)) k_rosenbrock<true><<<grid, threads, 0, stream>>>({x, g}, n, o->ws, fx);
            else k_rosenbrock<false><<<grid, threads, 0, stream>>>({x, g}, n, o->ws, fx);
            break;
        }
        case OBJ_BOOTH:
            if (n != 2) return LBFGSB200_ERR_INVALID_PARAM;
            k_booth<<<1, 1, 0, stream>>>(x, g, fx);
            break;
        case OBJ_GLM: {
            if (n != o->ncol) return LBFGSB200_ERR_INVALID_PARAM;
            if (o->fused) {
                const int frc = glm_fused(o, x, g, stream, fx);
                if (frc == 0) {
                    o->last_path = (o->ncol & 1) ? LBFGSB200_GLM_PATH_FUSED_ODD
                                                 : (o->ncol > 20 * 2 * kThreads ? LBFGSB200_GLM_PATH_FUSED_CLUSTER : LBFGSB200_GLM_PATH_FUSED);
                    break;
                }
                if (frc != LBFGSB200_ERR_UNSUPPORTED) return frc;
            }
            o->last_path = LBFGSB200_GLM_PATH_TWO_PASS;
            int64_t blocks = (o->nrow + kWarps - 1) / kWarps;
            const int64_t cap = (int64_t)o->dev.sm_count * 8;
            if (blocks > cap) blocks = cap;
            k_glm_rows<<<(int)blocks, kThreads, 0, stream>>>(o->X, o->y, x, o->t, o->nrow, o->ncol, o->glm_kind, o->ws, fx);
            const int rows_per_chunk = (int)((o->nrow + o->row_chunks - 1) / o->row_chunks);
            dim3 grid((unsigned)((o->ncol + kThreads - 1) / kThreads), (unsigned)o->row_chunks);
            k_glm_grad_partial<<<grid, kThreads, 0, stream>>>(o->X, o->t, o->gpart, o->nrow, o->ncol, rows_per_chunk, o->glm_kind);
            k_glm_grad_final<<<grid.x, kThreads, 0, stream>>>(o->gpart, g, o->ncol, o->row_chunks);
            break;
        }
        case OBJ_LJ: {
            if (n % 3 != 0 || n < 3) return LBFGSB200_ERR_INVALID_PARAM;
            const int64_t nloc = n / 3;
            if (o->comm && comm_size(o->comm) > 1) {  // atoms sharded: gather all positions, forces stay local
                const int r = comm_rank(o->comm);
                if (o->offsets[r + 1] - o->offsets[r] != n) return LBFGSB200_ERR_INVALID_PARAM;
                const int grc = comm_allgatherv(o->comm, x, o->xall, o->offsets.data(), stream);
                if (grc != 0) return grc;
                const int lrc = launch_lj(o, o->xall, g, o->offsets.back() / 3, o->offsets[r] / 3, nloc, stream, fx);
                if (lrc != 0) return lrc;
                break;
            }
            const int lrc = launch_lj(o, x, g, nloc, 0, nloc, stream, fx);
            if (lrc != 0) return lrc;
            if (o->sequential) k_lj_energy_seq<<<1, 1, 0, stream>>>(x, nloc, o->eps, o->sigma, fx);
            break;
        }
        default:
            return LBFGSB200_ERR_INVALID_PARAM;
    }
    if (o->kind == OBJ_GLM && o->comm && comm_size(o->comm) > 1) {
        // rows of X are sharded over the ranks, w is replicated: every rank needs the full f and gradient
        // (ncclAllReduce returns the same bits on every rank, so the replicated solvers stay in lockstep)
        int arc = comm_allreduce_sum(o->comm, g, (int)o->ncol, stream);
        if (arc == 0) arc = comm_allreduce_sum(o->comm, fx, 1, stream);
        if (arc != 0) return arc;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

int make(int device, int kind, Objective **out) {
    if (!out) return LBFGSB200_ERR_INVALID_PARAM;
    *out = nullptr;
    Objective *o = new (std::nothrow) Objective();
    if (!o) return LBFGSB200_ERR_CUDA;
    o->kind = kind;
    int rc = query_device(device, &o->dev);
    if (rc != 0) { delete o; return rc; }
    if (cudaSetDevice(device) != cudaSuccess) { delete o; return LBFGSB200_ERR_CUDA; }
    rc = alloc_reduce_ws(o->dev, &o->ws);
    if (rc != 0) { delete o; return rc; }
    const char *rg = getenv("LBFGSB200_COMMIT_RECOMPUTE_GP");
    o->recompute_gp = !(rg && rg[0] == '0');
    *out = o;
    return 0;
}

}  // namespace
}  // namespace lb

extern "C" {

int lbfgsb200_objective_rosenbrock(int device, lbfgsb200_objective_t **out) {
    return lb::make(device, lb::OBJ_ROSENBROCK, reinterpret_cast<lb::Objective **>(out));
}
int lbfgsb200_objective_booth(int device, lbfgsb200_objective_t **out) {
    return lb::make(device, lb::OBJ_BOOTH, reinterpret_cast<lb::Objective **>(out));
}
int lbfgsb200_objective_glm(int device, int kind, const double *X_dev, const double *y_dev, int64_t nrow, int64_t ncol,
                            lbfgsb200_objective_t **out) {
    if (!X_dev || !y_dev || nrow < 1 || ncol < 1 || kind < 0 || kind > 1) return LBFGSB200_ERR_INVALID_PARAM;
    lb::Objective *o = nullptr;
    int rc = lb::make(device, lb::OBJ_GLM, &o);
    if (rc != 0) return rc;
    o->glm_kind = kind;
    o->X = X_dev;
    o->y = y_dev;
    o->nrow = nrow;
    o->ncol = ncol;
    int chunks = (int)((nrow + 255) / 256);
    const int col_blocks = (int)((ncol + lb::kThreads - 1) / lb::kThreads);
    const int want = (o->dev.sm_count * 8 + col_blocks - 1) / col_blocks;
    if (chunks > want) chunks = want;
    if (chunks < 1) chunks = 1;
    o->row_chunks = chunks;
    const char *fenv = getenv("LBFGSB200_GLM_FUSED");
    o->fused = !(fenv && fenv[0] == '0');
    if (cudaMalloc((void **)&o->t, sizeof(double) * (size_t)nrow) != cudaSuccess ||
        cudaMalloc((void **)&o->gpart, sizeof(double) * (size_t)chunks * (size_t)ncol) != cudaSuccess ||
        cudaMalloc((void **)&o->gfused, sizeof(double) * (size_t)o->dev.sm_count * (size_t)ncol) != cudaSuccess) {
        lbfgsb200_objective_destroy(reinterpret_cast<lbfgsb200_objective_t *>(o));
        return LBFGSB200_ERR_CUDA;
    }
    *out = reinterpret_cast<lbfgsb200_objective_t *>(o);
    return 0;
}
int lbfgsb200_objective_lennard_jones(int device, double epsilon, double sigma, lbfgsb200_objective_t **out) {
    lb::Objective *o = nullptr;
    int rc = lb::make(device, lb::OBJ_LJ, &o);
    if (rc != 0) return rc;
    o->eps = epsilon;
    o->sigma = sigma;
    *out = reinterpret_cast<lbfgsb200_objective_t *>(o);
    return 0;
}
int lbfgsb200_objective_set_reduction(lbfgsb200_objective_t *objective, int reduction) {
    lb::Objective *o = reinterpret_cast<lb::Objective *>(objective);
    if (!o || (reduction != LBFGSB200_REDUCE_TREE && reduction != LBFGSB200_REDUCE_SEQUENTIAL)) return LBFGSB200_ERR_INVALID_PARAM;
    if (reduction == LBFGSB200_REDUCE_SEQUENTIAL && o->kind == lb::OBJ_GLM) return LBFGSB200_ERR_UNSUPPORTED;
    o->sequential = reduction == LBFGSB200_REDUCE_SEQUENTIAL;
    return 0;
}
int lbfgsb200_objective_last_path(const lbfgsb200_objective_t *objective) {
    const lb::Objective *o = reinterpret_cast<const lb::Objective *>(objective);
    return o ? o->last_path : 0;
}
int lbfgsb200_objective_set_lj_fast(lbfgsb200_objective_t *objective, int fast) {
    lb::Objective *o = reinterpret_cast<lb::Objective *>(objective);
    if (!o || o->kind != lb::OBJ_LJ) return LBFGSB200_ERR_INVALID_PARAM;
    o->lj_fast = fast != 0;
    return 0;
}
int lbfgsb200_objective_set_shard(lbfgsb200_objective_t *objective, lbfgsb200_comm_t *comm, const int64_t *shard_offsets) {
    lb::Objective *o = reinterpret_cast<lb::Objective *>(objective);
    if (!o) return LBFGSB200_ERR_INVALID_PARAM;
    lb::Comm *c = reinterpret_cast<lb::Comm *>(comm);
    if (o->xall) { cudaFree(o->xall); o->xall = nullptr; }
    o->offsets.clear();
    o->comm = nullptr;
    if (!c || lb::comm_size(c) == 1) return 0;
    if (o->kind == lb::OBJ_LJ) {
        if (!shard_offsets) return LBFGSB200_ERR_INVALID_PARAM;
        const int nr = lb::comm_size(c);
        o->offsets.assign(shard_offsets, shard_offsets + nr + 1);
        for (int r = 0; r <= nr; ++r)
            if (o->offsets[r] % 3 != 0 || (r > 0 && o->offsets[r] < o->offsets[r - 1])) return LBFGSB200_ERR_INVALID_PARAM;
        if (cudaSetDevice(o->dev.device) != cudaSuccess ||
            cudaMalloc((void **)&o->xall, sizeof(double) * (size_t)o->offsets.back()) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    }
    o->comm = c;   // Rosenbrock / Booth are shard-local: nothing to exchange
    return 0;
}
int lbfgsb200_objective_has_trial_eval(const lbfgsb200_objective_t *objective) {
    const lb::Objective *o = reinterpret_cast<const lb::Objective *>(objective);
    return (o && o->kind == lb::OBJ_ROSENBROCK) ? 1 : 0;
}
int lbfgsb200_objective_trial_eval(void *objective, const double *xp_dev, const double *d_dev, double step,
                                   double *x_dev, double *g_dev, int64_t n_local, void *stream, double *out_dev) {
    if (!objective || !xp_dev || !d_dev || !x_dev || !g_dev || !out_dev) return LBFGSB200_ERR_INVALID_PARAM;
    return lb::trial_impl(reinterpret_cast<lb::Objective *>(objective), xp_dev, d_dev, step, x_dev, g_dev, n_local,
                          (cudaStream_t)stream, out_dev);
}
int lbfgsb200_objective_probe(void *objective, const double *xp_dev, const double *d_dev, double step, const double *step_dev,
                              int64_t n_local, void *stream, double *out_dev) {
    if (!objective || !xp_dev || !d_dev || !out_dev) return LBFGSB200_ERR_INVALID_PARAM;
    return lb::probe_impl(reinterpret_cast<lb::Objective *>(objective), xp_dev, d_dev, step, step_dev, n_local,
                          (cudaStream_t)stream, out_dev);
}
int lbfgsb200_objective_commit(void *objective, const double *xp_dev, const double *d_dev, const double *gp_dev, double step,
                               double bs_scale, double *x_dev, double *g_dev, double *s_dev, double *y_dev, int64_t n_local,
                               void *stream, double *out_dev) {
    if (!objective || !xp_dev || !d_dev || !gp_dev || !x_dev || !g_dev || !s_dev || !y_dev || !out_dev) return LBFGSB200_ERR_INVALID_PARAM;
    return lb::commit_impl(reinterpret_cast<lb::Objective *>(objective), xp_dev, d_dev, gp_dev, step, bs_scale, x_dev, g_dev,
                           s_dev, y_dev, n_local, (cudaStream_t)stream, out_dev);
}
int lbfgsb200_objective_probe_multi(void *objective, const double *xp_dev, const double *d_dev, const double *steps,
                                    const double *step0_dev, int k, int64_t n_local, void *stream, double *out_dev) {
    if (!objective || !xp_dev || !d_dev || !out_dev) return LBFGSB200_ERR_INVALID_PARAM;
    return lb::probe_multi_impl(reinterpret_cast<lb::Objective *>(objective), xp_dev, d_dev, steps, step0_dev, k, n_local,
                                (cudaStream_t)stream, out_dev);
}
int lbfgsb200_objective_commit_gram(void *objective, const double *xp_dev, const double *d_dev, const double *gp_dev, double step,
                                    double bs_scale, double *x_dev, double *g_dev, double *s_dev, double *y_dev,
                                    const double *const *s_old_dev, const double *const *y_old_dev, int n_old, int64_t n_local,
                                    void *stream, double *out_dev, double *gram_out_dev, double *newdot_out_dev) {
    (void)gp_dev;   // recomputed from xp
    if (!objective || !xp_dev || !d_dev || !x_dev || !g_dev || !s_dev || !y_dev || !out_dev || !gram_out_dev || !newdot_out_dev ||
        (n_old > 0 && (!s_old_dev || !y_old_dev)))
        return LBFGSB200_ERR_INVALID_PARAM;
    return lb::commit_gram_impl(reinterpret_cast<lb::Objective *>(objective), xp_dev, d_dev, step, bs_scale, x_dev, g_dev, s_dev,
                                y_dev, s_old_dev, y_old_dev, n_old, n_local, (cudaStream_t)stream, out_dev, gram_out_dev,
                                newdot_out_dev);
}
int lbfgsb200_objective_fused_ops(lbfgsb200_objective_t *objective, lbfgsb200_fused_ops_t *out) {
    lb::Objective *o = reinterpret_cast<lb::Objective *>(objective);
    if (!o || !out) return LBFGSB200_ERR_INVALID_PARAM;
    out->struct_size = (int64_t)sizeof(lbfgsb200_fused_ops_t);
    out->trial = nullptr;
    out->probe = nullptr;
    out->commit = nullptr;
    out->commit_gram = nullptr;
    out->probe_multi = nullptr;
    out->user = o;
    out->flags = 0;
    if (o->kind == lb::OBJ_ROSENBROCK) {
        out->trial = lbfgsb200_objective_trial_eval;
        out->probe = lbfgsb200_objective_probe;
        out->commit = lbfgsb200_objective_commit;
        if (o->recompute_gp) out->commit_gram = lbfgsb200_objective_commit_gram;
        out->probe_multi = lbfgsb200_objective_probe_multi;
        if (lb::sums_over_ranks(o)) out->flags |= LBFGSB200_FUSED_SUMS_OVER_RANKS;
        if (o->recompute_gp) out->flags |= LBFGSB200_FUSED_COMMIT_SKIPS_GP;
    }
    return 0;
}
void lbfgsb200_objective_destroy(lbfgsb200_objective_t *objective) {
    lb::Objective *o = reinterpret_cast<lb::Objective *>(objective);
    if (!o) return;
    if (o->t) cudaFree(o->t);
    if (o->gpart) cudaFree(o->gpart);
    if (o->gfused) cudaFree(o->gfused);
    if (o->xall) cudaFree(o->xall);
    if (o->wide_partials) cudaFree(o->wide_partials);
    lb::free_reduce_ws(&o->ws);
    delete o;
}
int lbfgsb200_objective_eval(void *objective, const double *x_dev, double *g_dev, int64_t n_local, void *stream,
                             double *fx_dev) {
    if (!objective || !x_dev || !g_dev || !fx_dev) return LBFGSB200_ERR_INVALID_PARAM;
    return lb::eval_impl(reinterpret_cast<lb::Objective *>(objective), x_dev, g_dev, n_local, (cudaStream_t)stream, fx_dev);
}

}  // extern "C"
