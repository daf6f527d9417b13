// capi.cpp — the extern "C" surface declared in include/lbfgsb200.h.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>

#include "../../include/lbfgsb200.h"
#include "kernels.h"
#include "linesearch.h"
#include "solver.h"

namespace {

using lb::Solver;

inline Solver *S(lbfgsb200_solver_t *s) { return reinterpret_cast<Solver *>(s); }
inline const Solver *S(const lbfgsb200_solver_t *s) { return reinterpret_cast<const Solver *>(s); }

// Per-device context for the stand-alone primitives (reduction workspace + one result slot).
struct PrimCtx {
    lb::DeviceInfo dev{};
    lb::ReduceWs ws{};
    double *out_dev = nullptr;      // kMaxAcc result doubles, then kPrimIn uploaded scalars (the exported fused steps)
    std::mutex mu;                  // one launch + read-back at a time per device: ticket, partials and out_dev are shared
    bool ok = false;
};
constexpr int kPrimIn = 16;
std::mutex g_prim_mu;
std::map<int, PrimCtx> g_prim;

int prim_ctx(PrimCtx **out) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    std::lock_guard<std::mutex> lock(g_prim_mu);
    PrimCtx &c = g_prim[device];
    if (!c.ok) {
        int rc = lb::query_device(device, &c.dev);
        if (rc != 0) return rc;
        rc = lb::alloc_reduce_ws(c.dev, &c.ws);
        if (rc != 0) return rc;
        if (cudaMalloc((void **)&c.out_dev, sizeof(double) * (lb::kMaxAcc + kPrimIn)) != cudaSuccess) return LBFGSB200_ERR_CUDA;
        c.ok = true;
    }
    *out = &c;
    return 0;
}

lb::Launch prim_launch(const PrimCtx &c, void *stream, int64_t n, int nvec) {
    lb::Launch L;
    L.stream = (cudaStream_t)stream;
    L.max_grid = c.dev.sm_count * c.dev.blocks_per_sm;
    L.max_grid_trial = c.dev.sm_count * c.dev.blocks_per_sm_trial;
    L.streaming = (double)n * 8.0 * nvec > 0.75 * (double)c.dev.l2_bytes;
    L.ws = c.ws;
    return L;
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

int read_back(const PrimCtx &c, void *stream, int count, double *out_host) {
    if (cudaMemcpyAsync(out_host, c.out_dev, sizeof(double) * count, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess)
        return LBFGSB200_ERR_CUDA;
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

}  // namespace

extern "C" {

int lbfgsb200_abi_version(void) { return LBFGSB200_ABI_VERSION; }

void lbfgsb200_param_default(lbfgsb200_param_t *p) {
    if (!p) return;
    p->struct_size = (int64_t)sizeof(lbfgsb200_param_t);
    p->m = 6;                               // src/lbfgs.rs:163
    p->epsilon = 1e-5;                      // :164
    p->past = 0;                            // :165
    p->delta = 1e-5;                        // :166
    p->max_iterations = 0;                  // :167
    p->max_evaluations = 0;                 // :168
    p->ls_algorithm = LBFGSB200_LS_MORETHUENTE;  // src/line.rs:82-88
    p->ls_ftol = 1e-4;                      // src/line.rs:153
    p->ls_gtol = 0.9;                       // :154
    p->ls_xtol = 2.220446049250313e-16;     // :155 f64::EPSILON
    p->ls_min_step = 1e-20;                 // :156
    p->ls_max_step = 1e+20;                 // :157
    p->ls_max_linesearch = 20;              // :158
    p->ls_gradient_only = 0;                // :159
    p->orthantwise = 0;                     // src/lbfgs.rs:169
    p->owl_c = 1.0;                         // src/orthantwise.rs:50
    p->owl_start = 0;                       // :51
    p->owl_end = -1;                        // :52 None
    p->initial_inverse_hessian = 1.0;       // src/lbfgs.rs:171
    p->max_step_size = 1.0;                 // :172
    p->damping = 0;                         // :173
    p->constrain_step_size = 1;             // :174
    p->reduction = LBFGSB200_REDUCE_TREE;   // extension
}

// ---- solver ----------------------------------------------------------------------------------
int lbfgsb200_create(const lbfgsb200_param_t *param, int64_t n_local, int64_t n_global, int64_t global_offset,
                     int device, void *stream, lbfgsb200_comm_t *comm, lbfgsb200_solver_t **out) {
    if (!param || !out) return LBFGSB200_ERR_INVALID_PARAM;
    *out = nullptr;
    Solver *s = new (std::nothrow) Solver();
    if (!s) return LBFGSB200_ERR_CUDA;
    int rc = s->init(*param, n_local, n_global, global_offset, device, (cudaStream_t)stream,
                     reinterpret_cast<lb::Comm *>(comm));
    if (rc != 0) {
        delete s;
        return rc;
    }
    *out = reinterpret_cast<lbfgsb200_solver_t *>(s);
    return 0;
}
void lbfgsb200_destroy(lbfgsb200_solver_t *solver) { delete S(solver); }
const char *lbfgsb200_last_error(const lbfgsb200_solver_t *solver) { return solver ? S(solver)->error().c_str() : ""; }

int lbfgsb200_minimize(lbfgsb200_solver_t *solver, double *x_dev, lbfgsb200_eval_fn eval, void *eval_user,
                       lbfgsb200_progress_fn progress, void *progress_user, lbfgsb200_report_t *report) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    return S(solver)->minimize(x_dev, eval, eval_user, progress, progress_user, report);
}
int lbfgsb200_set_trial_evaluate(lbfgsb200_solver_t *solver, lbfgsb200_trial_eval_fn fn, void *user) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    S(solver)->set_trial_evaluate(fn, user);
    return 0;
}
int lbfgsb200_set_fused_ops(lbfgsb200_solver_t *solver, const lbfgsb200_fused_ops_t *ops) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    if (!ops) { S(solver)->set_fused_ops(nullptr); return 0; }
    // a caller built against the header before `commit_gram` was appended passes the shorter struct
    if ((ops->struct_size != (int64_t)sizeof(lbfgsb200_fused_ops_t) && ops->struct_size != LBFGSB200_FUSED_OPS_SIZE_V1 &&
         ops->struct_size != LBFGSB200_FUSED_OPS_SIZE_V2) ||
        (ops->probe && !ops->commit))
        return LBFGSB200_ERR_INVALID_PARAM;
    lbfgsb200_fused_ops_t full{};
    memcpy(&full, ops, (size_t)ops->struct_size);
    full.struct_size = (int64_t)sizeof(lbfgsb200_fused_ops_t);
    S(solver)->set_fused_ops(&full);
    return 0;
}
int lbfgsb200_set_direction(lbfgsb200_solver_t *solver, int mode) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    return S(solver)->set_direction(mode);
}
int lbfgsb200_set_default_direction(int mode) { return lb::set_default_direction(mode); }
int lbfgsb200_get_direction(const lbfgsb200_solver_t *solver) {
    return solver ? S(solver)->direction_mode() : LBFGSB200_ERR_INVALID_PARAM;
}
int lbfgsb200_build(lbfgsb200_solver_t *solver, double *x_dev, lbfgsb200_eval_fn eval, void *eval_user) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    return S(solver)->build(x_dev, eval, eval_user);
}
int lbfgsb200_is_converged(lbfgsb200_solver_t *solver, int *stop_status) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    return S(solver)->is_converged(stop_status) ? 1 : 0;
}
int lbfgsb200_propagate(lbfgsb200_solver_t *solver, lbfgsb200_progress_t *progress_out) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    return S(solver)->propagate(progress_out);
}
int lbfgsb200_report(lbfgsb200_solver_t *solver, lbfgsb200_report_t *report_out) {
    if (!solver || !report_out) return LBFGSB200_ERR_INVALID_PARAM;
    S(solver)->report(report_out);
    return 0;
}
int lbfgsb200_finish(lbfgsb200_solver_t *solver) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    return S(solver)->finish();
}
const double *lbfgsb200_x(const lbfgsb200_solver_t *solver) { return solver ? S(solver)->x() : nullptr; }
const double *lbfgsb200_gx(const lbfgsb200_solver_t *solver) { return solver ? S(solver)->gx() : nullptr; }
const double *lbfgsb200_direction(const lbfgsb200_solver_t *solver) { return solver ? S(solver)->direction() : nullptr; }

int lbfgsb200_minimize_host_ex(const lbfgsb200_param_t *param, double *x_host, int64_t n_local, int64_t n_global,
                               int64_t global_offset, int device, lbfgsb200_comm_t *comm, lbfgsb200_eval_fn eval,
                               void *eval_user, const lbfgsb200_fused_ops_t *fused,
                               lbfgsb200_progress_fn progress, void *progress_user, lbfgsb200_report_t *report) {
    if (!param || !x_host || n_local < 1 || !eval) return LBFGSB200_ERR_INVALID_PARAM;
    if (cudaSetDevice(device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    const char *dbg_env = getenv("LBFGSB200_DEBUG_TIMING");
    const bool dbg = dbg_env && dbg_env[0] != '0';
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!dbg) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[lbfgsb200] minimize_host: %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    cudaStream_t stream = nullptr;
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    lap("stream create");
    double *x_dev = nullptr;
    int status = LBFGSB200_ERR_CUDA;
    lbfgsb200_solver_t *solver = nullptr;
    const size_t bytes = sizeof(double) * (size_t)n_local;
    do {
        // the arena first: it is the big block a previous solve left in the pool; x then takes a small one
        status = lbfgsb200_create(param, n_local, n_global, global_offset, device, stream, comm, &solver);
        if (status != 0) break;
        lap("solver create");
        status = LBFGSB200_ERR_CUDA;
        x_dev = S(solver)->spare_x();   // the device copy of x lives in the solver's (pooled) arena
        if (cudaMemcpyAsync(x_dev, x_host, bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) break;
        lap("H2D enqueue");
        if (fused && (status = lbfgsb200_set_fused_ops(solver, fused)) != 0) break;
        status = lbfgsb200_minimize(solver, x_dev, eval, eval_user, progress, progress_user, report);
        lap("minimize");
        if (cudaMemcpyAsync(x_host, x_dev, bytes, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess) {
            if (status >= 0) status = LBFGSB200_ERR_CUDA;
        }
        lap("D2H + sync");
    } while (0);
    x_dev = nullptr;
    if (solver) lbfgsb200_destroy(solver);
    lap("solver destroy");
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    lap("stream destroy");
    return status;
}

int lbfgsb200_minimize_host(const lbfgsb200_param_t *param, double *x_host, int64_t n, int device,
                            lbfgsb200_eval_fn eval, void *eval_user, lbfgsb200_progress_fn progress,
                            void *progress_user, lbfgsb200_report_t *report) {
    return lbfgsb200_minimize_host_ex(param, x_host, n, n, 0, device, nullptr, eval, eval_user, nullptr, progress,
                                      progress_user, report);
}

int lbfgsb200_profile_enable(lbfgsb200_solver_t *solver, int timing) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    S(solver)->profile_enable(timing);
    return 0;
}
int lbfgsb200_profile_get(lbfgsb200_solver_t *solver, lbfgsb200_profile_t *out) {
    if (!solver || !out) return LBFGSB200_ERR_INVALID_PARAM;
    S(solver)->profile_get(out);
    return 0;
}
int lbfgsb200_profile_reset(lbfgsb200_solver_t *solver) {
    if (!solver) return LBFGSB200_ERR_INVALID_PARAM;
    S(solver)->profile_reset();
    return 0;
}

// ---- LbfgsMath primitives ----------------------------------------------------------------------
#define LB_PRIM_PROLOGUE(NVEC)                                       \
    if (n < 1) return LBFGSB200_ERR_INVALID_PARAM;                   \
    PrimCtx *c = nullptr;                                            \
    int rc = prim_ctx(&c);                                           \
    if (rc != 0) return rc;                                          \
    std::lock_guard<std::mutex> prim_lock(c->mu);                    \
    lb::Launch L = prim_launch(*c, stream, n, NVEC);

int lbfgsb200_vecadd(double *y, const double *x, double cc, int64_t n, void *stream) {
    if (!aligned16(y) || !aligned16(x)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(2)
    lb::launch_vecadd(L, y, x, cc, n);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_vecdot(const double *x, const double *y, int64_t n, void *stream, double *out_host) {
    if (!aligned16(y) || !aligned16(x) || !out_host) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(2)
    lb::launch_vecdot(L, x, y, n, c->out_dev);
    return read_back(*c, stream, 1, out_host);
}
int lbfgsb200_vecscale(double *y, double cc, int64_t n, void *stream) {
    if (!aligned16(y)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(1)
    lb::launch_vecscale(L, y, cc, n);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_veccpy(double *y, const double *x, int64_t n, void *stream) {
    if (!aligned16(y) || !aligned16(x)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(2)
    lb::launch_veccpy(L, y, x, n, false);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_vecncpy(double *y, const double *x, int64_t n, void *stream) {
    if (!aligned16(y) || !aligned16(x)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(2)
    lb::launch_veccpy(L, y, x, n, true);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_vecdiff(double *z, const double *x, const double *y, int64_t n, void *stream) {
    if (!aligned16(z) || !aligned16(y) || !aligned16(x)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(3)
    lb::launch_vecdiff(L, z, x, y, n);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_vec2norm(const double *x, int64_t n, void *stream, double *out_host) {
    double v = 0.0;
    int rc = lbfgsb200_vecdot(x, x, n, stream, &v);
    if (rc != 0) return rc;
    *out_host = std::sqrt(v);  // math.rs:73-76
    return 0;
}
int lbfgsb200_vec2norminv(const double *x, int64_t n, void *stream, double *out_host) {
    double v = 0.0;
    int rc = lbfgsb200_vec2norm(x, n, stream, &v);
    if (rc != 0) return rc;
    *out_host = 1.0 / v;  // math.rs:79-81
    return 0;
}

// ---- fused steps, stand-alone -------------------------------------------------------------------
int lbfgsb200_dots3(const double *g, const double *d, const double *x, int64_t n, void *stream, double out_host[3]) {
    if (!aligned16(g) || !aligned16(x) || (d && !aligned16(d))) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(3)
    if (cudaMemsetAsync(c->out_dev, 0, sizeof(double) * 3, (cudaStream_t)stream) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    lb::launch_dots(L, g, d, x, n, c->out_dev);
    return read_back(*c, stream, 3, out_host);
}
int lbfgsb200_trial_step(double *x, const double *xp, const double *d, double step, int64_t n,
                         const signed char *wp, int64_t start, int64_t end, void *stream) {
    if (!aligned16(x) || !aligned16(xp) || !aligned16(d)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(3)
    if (wp) { if (end < 0 || end > n) end = n; }
    lb::launch_trial(L, x, xp, d, step, n, wp, start, end, 0);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_owl_pseudo_gradient(double *pg, const double *x, const double *g, int64_t n, double cc,
                                  int64_t start, int64_t end, void *stream, double out_host[3]) {
    if (!aligned16(pg) || !aligned16(x) || !aligned16(g)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(3)
    if (end < 0 || end > n) end = n;
    if (!(start < end)) return LBFGSB200_ERR_INVALID_PARAM;
    lb::launch_owl_pg(L, pg, x, g, nullptr, n, cc, start, end, 0, c->out_dev);
    return read_back(*c, stream, 3, out_host);
}
int lbfgsb200_owl_orthant(signed char *wp, const double *xp, const double *pg, int64_t n, void *stream) {
    if (!aligned16(xp) || !aligned16(pg) || ((uintptr_t)wp & 1u)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(2)
    lb::launch_orthant(L, wp, xp, pg, n);
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_owl_constrain_direction(double *d, const double *pg, int64_t n, int64_t start, int64_t end,
                                      void *stream, double out_host[1]) {
    if (!aligned16(d) || !aligned16(pg)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(2)
    if (end < 0 || end > n) end = n;
    if (!(start < end)) return LBFGSB200_ERR_INVALID_PARAM;
    lb::launch_owl_constrain(L, d, pg, n, start, end, 0, c->out_dev);
    return read_back(*c, stream, 1, out_host);
}

// ---- the update chain's kernels, one call each ------------------------------------------------------
namespace {
// uploads `count` scalars behind the result slot; returns their device address
const double *upload(PrimCtx *c, void *stream, const double *vals, int count) {
    double *dst = c->out_dev + lb::kMaxAcc;
    if (cudaMemcpyAsync(dst, vals, sizeof(double) * count, cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) return nullptr;
    return dst;   // `vals` is pageable host memory: the copy has been staged when cudaMemcpyAsync returns
}
}  // namespace

int lbfgsb200_init_direction(double *d, const double *g, int64_t n, void *stream, double out_host[2]) {
    if (!aligned16(d) || !aligned16(g) || !out_host) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(2)
    lb::launch_init_dir(L, d, g, n, c->out_dev);
    return read_back(*c, stream, 2, out_host);
}
int lbfgsb200_history_update(double *s, double *y, const double *x, const double *xp, const double *g, const double *gp,
                             const double *pg, int64_t n, double step, int damping, void *stream, double out_host[5]) {
    if (!aligned16(s) || !aligned16(y) || !aligned16(x) || !aligned16(xp) || !aligned16(g) || !aligned16(gp) ||
        (pg && !aligned16(pg)) || !out_host)
        return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(6)
    if (cudaMemsetAsync(c->out_dev, 0, sizeof(double) * 5, (cudaStream_t)stream) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    lb::launch_history(L, x, xp, g, gp, pg, s, y, n, -step, damping != 0, c->out_dev);
    rc = read_back(*c, stream, 5, out_host);
    if (rc == 0 && !damping) out_host[4] = 0.0;
    return rc;
}
int lbfgsb200_damp_y(double *y, const double *gp, int64_t n, double step, double ys, double sbs, void *stream,
                     int *applied_host) {
    if (!aligned16(y) || !aligned16(gp)) return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(2)
    const double hist[5] = {0.0, ys, 0.0, 0.0, sbs};
    const double *hist_dev = upload(c, stream, hist, 5);
    if (!hist_dev) return LBFGSB200_ERR_CUDA;
    lb::launch_damp(L, y, gp, n, -step, hist_dev);
    if (applied_host) *applied_host = (ys < (1.0 - 0.6) * sbs) ? 1 : 0;   // the kernel's own test, src/lbfgs.rs:664-675
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_two_loop_backward_step(double *q, const double *g_first, const double *y_j, const double *s_next, int64_t n,
                                     double sq, double ys_j, double gamma, void *stream, double out_host[2]) {
    if (!aligned16(q) || !aligned16(y_j) || (g_first && !aligned16(g_first)) || (s_next && !aligned16(s_next)) || !out_host)
        return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(4)
    // in: [0] = s_j.q, [1] = y_j.s_j, hist = in + 2: hist[1] / hist[2] = gamma / 1; out: [0] = alpha, [1] = the dot
    const double in[5] = {sq, ys_j, 0.0, gamma, 1.0};
    const double *in_dev = upload(c, stream, in, 5);
    if (!in_dev) return LBFGSB200_ERR_CUDA;
    lb::launch_backward(L, g_first != nullptr, s_next == nullptr, q, g_first, y_j, s_next, n, in_dev + 0, in_dev + 1, nullptr,
                        in_dev + 2, c->out_dev + 0, c->out_dev + 1);
    return read_back(*c, stream, 2, out_host);
}
int lbfgsb200_two_loop_forward_step(double *r, const double *s_j, const double *y_next, const double *g_last, int64_t n,
                                    double yr, double ys_j, double alpha_j, int owl, int64_t owl_start, int64_t owl_end,
                                    void *stream, double out_host[4]) {
    const bool last = y_next == nullptr;
    if (!aligned16(r) || !aligned16(s_j) || (y_next && !aligned16(y_next)) || (last && (!g_last || !aligned16(g_last))) || !out_host)
        return LBFGSB200_ERR_INVALID_PARAM;
    LB_PRIM_PROLOGUE(4)
    if (owl_end < 0 || owl_end > n) owl_end = n;
    const double in[3] = {yr, ys_j, alpha_j};
    const double *in_dev = upload(c, stream, in, 3);
    if (!in_dev) return LBFGSB200_ERR_CUDA;
    if (cudaMemsetAsync(c->out_dev, 0, sizeof(double) * 4, (cudaStream_t)stream) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    lb::launch_forward(L, last, last && owl != 0, r, s_j, y_next, g_last, n, in_dev + 0, in_dev + 1, in_dev + 2, owl_start,
                       owl_end, 0, c->out_dev + 1);
    rc = read_back(*c, stream, 4, out_host);
    if (rc == 0) out_host[0] = yr / ys_j;   // beta, the kernel's own quotient (src/lbfgs.rs:597)
    return rc;
}

// ---- line-search state machine ------------------------------------------------------------------
lbfgsb200_linesearch_t *lbfgsb200_linesearch_begin(const lbfgsb200_param_t *p, int orthantwise, double finit,
                                                   double dginit, double step) {
    if (!p) return nullptr;
    lb::LsConfig cfg;
    cfg.algorithm = (int)p->ls_algorithm;
    cfg.ftol = p->ls_ftol;
    cfg.gtol = p->ls_gtol;
    cfg.xtol = p->ls_xtol;
    cfg.min_step = p->ls_min_step;
    cfg.max_step = p->ls_max_step;
    cfg.max_linesearch = p->ls_max_linesearch;
    cfg.gradient_only = p->ls_gradient_only != 0;
    lb::LineSearchMachine *m = new (std::nothrow) lb::LineSearchMachine();
    if (!m) return nullptr;
    if (m->begin(cfg, orthantwise != 0, finit, dginit, step) != 0) {
        delete m;
        return nullptr;
    }
    return reinterpret_cast<lbfgsb200_linesearch_t *>(m);
}
int lbfgsb200_linesearch_next(lbfgsb200_linesearch_t *ls, double *step_out) {
    return reinterpret_cast<lb::LineSearchMachine *>(ls)->next_trial(step_out) ? 1 : 0;
}
int lbfgsb200_linesearch_predict(const lbfgsb200_linesearch_t *ls, double *steps_out, int kmax) {
    if (!ls || !steps_out || kmax < 1) return 0;
    return reinterpret_cast<const lb::LineSearchMachine *>(ls)->predict(steps_out, kmax);
}
void lbfgsb200_linesearch_feed(lbfgsb200_linesearch_t *ls, int eval_ok, double f, double dg) {
    reinterpret_cast<lb::LineSearchMachine *>(ls)->feed(eval_ok != 0, f, dg);
}
int lbfgsb200_linesearch_result(lbfgsb200_linesearch_t *ls, int64_t *ncall, double *step) {
    lb::LineSearchMachine *m = reinterpret_cast<lb::LineSearchMachine *>(ls);
    if (ncall) *ncall = m->ncall();
    if (step) *step = m->step();
    return m->error();
}
void lbfgsb200_linesearch_end(lbfgsb200_linesearch_t *ls) { delete reinterpret_cast<lb::LineSearchMachine *>(ls); }

// ---- device helpers -----------------------------------------------------------------------------
int lbfgsb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int lbfgsb200_device_alloc(int device, int64_t bytes, void **out_dev) {
    if (!out_dev || bytes < 0) return LBFGSB200_ERR_INVALID_PARAM;
    if (cudaSetDevice(device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    return cudaMalloc(out_dev, (size_t)bytes) == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_device_free(void *dev) { return cudaFree(dev) == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA; }
int lbfgsb200_copy_h2d(void *dst_dev, const void *src_host, int64_t bytes, void *stream) {
    if (cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    return cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_copy_d2h(void *dst_host, const void *src_dev, int64_t bytes, void *stream) {
    if (cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    return cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}
int lbfgsb200_trim_pool(int device) {
    if (cudaSetDevice(device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    return lb::trim_pool(device);
}
int lbfgsb200_stream_synchronize(void *stream) {
    return cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

}  // extern "C"
