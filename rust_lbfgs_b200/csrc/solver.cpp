// solver.cpp — see solver.h.  Compiled by nvcc as host code with -ffp-contract=off.
#include "solver.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace lb {

namespace {
constexpr int64_t kAlign = 256;  // bytes; every owned vector starts on a 256-byte boundary
inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}
struct DbgTimer {  // LBFGSB200_DEBUG_TIMING=1: wall time of the host-side phases to stderr
    bool on = env_int("LBFGSB200_DEBUG_TIMING", 0) != 0;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[lbfgsb200] %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};
}  // namespace

namespace {
std::atomic<int> g_default_direction{-1};   // -1: not set, the environment decides
}

int query_device(int device, DeviceInfo *out) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return LBFGSB200_ERR_CUDA;
    int sm = 0, l2 = 0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    if (cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    out->device = device;
    out->sm_count = sm;
    out->l2_bytes = l2;
    out->blocks_per_sm = env_int("LBFGSB200_BLOCKS_PER_SM", 1);  // tuned: see types.h
    if (out->blocks_per_sm < 1) out->blocks_per_sm = 1;
    if (out->blocks_per_sm > 16) out->blocks_per_sm = 16;
    // the write-free probe carries ~25 flops per 32 bytes: with ONE resident CTA all 8 warps load, then all
    // compute, and the memory pipe idles meanwhile; two co-resident CTAs de-phase (measured at n = 1e8: 6.07 ->
    // 6.80 TB/s; a register-prefetching single CTA gives the same 6.81, profiles/r02_tuning.md)
    out->blocks_per_sm_trial = env_int("LBFGSB200_TRIAL_BLOCKS_PER_SM", 2);
    if (out->blocks_per_sm_trial < 1) out->blocks_per_sm_trial = 1;
    if (out->blocks_per_sm_trial > 16) out->blocks_per_sm_trial = 16;
    return 0;
}

int alloc_reduce_ws(const DeviceInfo &info, ReduceWs *ws) {
    ws->stride = info.sm_count * 16;
    if (cudaMalloc((void **)&ws->partials, sizeof(double) * kMaxAcc * ws->stride) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    if (cudaMalloc((void **)&ws->ticket, 256) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    if (cudaMemset(ws->ticket, 0, 256) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    return 0;
}
void free_reduce_ws(ReduceWs *ws) {
    if (ws->partials) cudaFree(ws->partials);
    if (ws->ticket) cudaFree(ws->ticket);
    ws->partials = nullptr;
    ws->ticket = nullptr;
}

// Small per-solver resources (device scalar slots, their pinned host mirror, the reduction workspace) are
// recycled through a process-wide cache: on the B200 boxes a cudaFree / cudaFreeHost at solver destruction
// sporadically blocked for 0.4 - 0.8 s (and cudaMallocHost for 80 ms), i.e. the cost of ~70 L-BFGS iterations at
// n = 1e8.  lbfgsb200_trim_pool() releases the cache.
namespace {
struct Scratch {
    int device = 0;
    size_t scal_count = 0;
    double *scal_dev = nullptr;
    double *scal_host = nullptr;
    ReduceWs ws{};
};
std::mutex g_scratch_mu;
std::vector<Scratch> g_scratch;

void scratch_free(Scratch &s) {
    if (s.scal_dev) cudaFree(s.scal_dev);
    if (s.scal_host) cudaFreeHost(s.scal_host);
    free_reduce_ws(&s.ws);
    s = Scratch();
}

int scratch_acquire(const DeviceInfo &dev, size_t scal_count, Scratch *out) {
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        for (size_t i = 0; i < g_scratch.size(); ++i) {
            if (g_scratch[i].device == dev.device && g_scratch[i].scal_count >= scal_count &&
                g_scratch[i].ws.stride >= dev.sm_count * 16) {
                *out = g_scratch[i];
                g_scratch.erase(g_scratch.begin() + i);
                return 0;
            }
        }
    }
    Scratch s;
    s.device = dev.device;
    s.scal_count = scal_count < 256 ? 256 : scal_count;   // room for alpha[m] up to m ~ 200 without a re-allocation
    if (cudaMalloc((void **)&s.scal_dev, sizeof(double) * s.scal_count) != cudaSuccess ||
        cudaMallocHost((void **)&s.scal_host, sizeof(double) * 128) != cudaSuccess ||   // Solver::kHostWords
        alloc_reduce_ws(dev, &s.ws) != 0) {
        scratch_free(s);
        return LBFGSB200_ERR_CUDA;
    }
    *out = s;
    return 0;
}

void scratch_release(const Scratch &s) {
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    g_scratch.push_back(s);
}
}  // namespace

// The arena (2m + 4..5 n-vectors: 12.8 GB at n = 1e8, m = 6) comes from a PRIVATE CUDA memory pool per device whose
// release threshold is lifted, so a process that solves repeatedly pays the driver's map/unmap of those pages
// once (measured on B200: cudaMalloc + cudaFree of the arena cost 30 ms .. 1.3 s per solve, against ~8 ms per
// L-BFGS iteration).  The device's default pool — which other libraries in the process may use — is left alone.
// lbfgsb200_trim_pool() hands the cached pages back.  LBFGSB200_POOL=0 falls back to cudaMalloc / cudaFree.
namespace {
std::mutex g_pool_mu;
cudaMemPool_t g_pool[64];
int g_pool_state[64];  // 0 = unknown, 1 = pool, 2 = plain
}  // namespace
static cudaMemPool_t arena_pool(int device) {
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (g_pool_state[device] == 0) {
        int ok = 0;
        cudaMemPool_t pool = nullptr;
        if (env_int("LBFGSB200_POOL", 1) != 0 &&
            cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, device) == cudaSuccess && ok) {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = device;
            uint64_t keep = UINT64_MAX;
            ok = cudaMemPoolCreate(&pool, &props) == cudaSuccess &&
                 cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
        } else {
            ok = 0;
        }
        cudaGetLastError();
        g_pool[device] = ok ? pool : nullptr;
        g_pool_state[device] = ok ? 1 : 2;
    }
    return g_pool_state[device] == 1 ? g_pool[device] : nullptr;
}

int trim_pool(int device) {
    if (cudaDeviceSynchronize() != cudaSuccess) return LBFGSB200_ERR_CUDA;
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        for (size_t i = 0; i < g_scratch.size();) {
            if (g_scratch[i].device == device) {
                scratch_free(g_scratch[i]);
                g_scratch.erase(g_scratch.begin() + i);
            } else {
                ++i;
            }
        }
    }
    cudaMemPool_t pool = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if (device >= 0 && device < 64 && g_pool_state[device] == 1) pool = g_pool[device];
    }
    if (!pool) return 0;   // nothing was ever pooled on this device
    return cudaMemPoolTrimTo(pool, 0) == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
}

Solver::~Solver() {
    DbgTimer tm;
    drop_graphs();
    for (auto &p : pending_) { event_pool_.push_back(p.a); event_pool_.push_back(p.b); }
    for (auto e : event_pool_) cudaEventDestroy(e);
    tm.lap("destroy: events");
    if (cmp_block_) { cudaStreamSynchronize(stream_); cudaFree(cmp_block_); }
    if (arena_) {
        if (arena_pooled_) cudaFreeAsync(arena_, stream_);
        else cudaFree(arena_);
    }
    tm.lap("destroy: arena");
    if (scal_dev_) {  // back to the cache; make sure nothing of ours is still running on these buffers
        cudaStreamSynchronize(stream_);
        Scratch sc;
        sc.device = dev_.device;
        sc.scal_count = scal_count_;
        sc.scal_dev = scal_dev_;
        sc.scal_host = scal_host_;
        sc.ws = ws_;
        scratch_release(sc);
    }
    tm.lap("destroy: scalars + workspace");
}

int Solver::fail(int status, const char *msg) {
    err_ = msg;
    last_status_ = status;
    return status;
}
int Solver::cuda_fail(cudaError_t e, const char *what) {
    err_ = std::string(what) + ": " + cudaGetErrorString(e);
    last_status_ = LBFGSB200_ERR_CUDA;
    return LBFGSB200_ERR_CUDA;
}

int Solver::init(const lbfgsb200_param_t &p, int64_t n_local, int64_t n_global, int64_t goff, int device,
                 cudaStream_t stream, Comm *comm) {
    if (p.struct_size != (int64_t)sizeof(lbfgsb200_param_t)) return fail(LBFGSB200_ERR_INVALID_PARAM, "param.struct_size mismatch");
    if (n_local < 1 || n_global < n_local || goff < 0 || goff + n_local > n_global) return fail(LBFGSB200_ERR_INVALID_PARAM, "invalid shard geometry");
    if (p.m < 1) return fail(LBFGSB200_ERR_INVALID_PARAM, "m must be >= 1");
    // the builder's assert!s, src/lbfgs.rs:195,204,216,232,254,269,308,321,361
    if (std::signbit(p.epsilon)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid parameter epsilon specified.");
    if (std::signbit(p.initial_inverse_hessian)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid beta parameter for scaling the initial step size.");
    if (std::signbit(p.max_step_size)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid max_step_size parameter.");
    if (!(p.ls_ftol >= 0.0)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid parameter ftol specified.");
    if (!(p.ls_gtol >= 0.0 && p.ls_gtol < 1.0)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid parameter gtol specified.");
    if (!(p.ls_xtol >= 0.0)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid parameter xtol specified.");
    if (!(p.ls_min_step >= 0.0)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid parameter min_step specified.");
    if (!(p.delta >= 0.0)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid parameter delta specified.");
    if (p.ls_algorithm < 0 || p.ls_algorithm > 3) return fail(LBFGSB200_ERR_INVALID_PARAM, "unknown line search algorithm");
    if (p.reduction != LBFGSB200_REDUCE_TREE && p.reduction != LBFGSB200_REDUCE_SEQUENTIAL)
        return fail(LBFGSB200_ERR_INVALID_PARAM, "unknown reduction mode");
    if (p.reduction == LBFGSB200_REDUCE_SEQUENTIAL && comm && comm_size(comm) > 1)
        return fail(LBFGSB200_ERR_INVALID_PARAM, "sequential (reference-order) reductions are single-GPU only");

    p_ = p;
    m_ = p.m;
    n_ = n_local;
    n_global_ = n_global;
    goff_ = goff;
    comm_ = comm;
    stream_ = stream;
    owl_ = p.orthantwise != 0;
    sequential_ = p.reduction == LBFGSB200_REDUCE_SEQUENTIAL;
    if (owl_) {
        if (std::signbit(p.owl_c)) return fail(LBFGSB200_ERR_INVALID_PARAM, "Invalid parameter orthantwise c parameter specified.");
        owl_start_ = p.owl_start;                                       // Orthantwise::start_end, orthantwise.rs:59-67
        owl_end_ = (p.owl_end < 0) ? n_global : (p.owl_end < n_global ? p.owl_end : n_global);
        if (!(owl_start_ < owl_end_)) return fail(LBFGSB200_ERR_INVALID_PARAM, "invalid start for orthantwise");
    }
    ls_.algorithm = (int)p.ls_algorithm;
    ls_.ftol = p.ls_ftol;
    ls_.gtol = p.ls_gtol;
    ls_.xtol = p.ls_xtol;
    ls_.min_step = p.ls_min_step;
    ls_.max_step = p.ls_max_step;
    ls_.max_linesearch = p.ls_max_linesearch;
    ls_.gradient_only = p.ls_gradient_only != 0;

    DbgTimer tm;
    int rc = query_device(device, &dev_);
    if (rc != 0) return fail(rc, "no usable CUDA device (this library has no CPU fallback)");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");

    // One arena for every owned vector: x', g, gp, d, [pg], S[m], Y[m], one spare vector, [wp].  The spare is the
    // device copy of x for the host-buffer entry points (lbfgsb200_minimize_host*): keeping it inside the arena
    // makes every solver of the same (n, m) ask the memory pool for the same block size, so host-API and
    // device-API solves recycle each other's arena instead of going to the driver (whose physical allocations
    // of even 0.8 GB were measured at 5 .. 640 ms on these boxes).
    const int64_t vec_bytes = round_up(n_ * (int64_t)sizeof(double), kAlign);
    const int64_t nvec = 4 + (owl_ ? 1 : 0) + 2 * m_ + 1;
    const int64_t wp_bytes = owl_ ? round_up(n_, kAlign) : 0;
    cudaMemPool_t pool = arena_pool(device);
    arena_pooled_ = pool != nullptr;
    if (arena_pooled_) e = cudaMallocFromPoolAsync(&arena_, (size_t)(nvec * vec_bytes + wp_bytes), pool, stream_);
    else e = cudaMalloc(&arena_, (size_t)(nvec * vec_bytes + wp_bytes));
    if (e != cudaSuccess) { arena_ = nullptr; return cuda_fail(e, "cudaMalloc(arena)"); }
    tm.lap("create: arena");
    char *base = (char *)arena_;
    auto take = [&]() { double *r = (double *)base; base += vec_bytes; return r; };
    xbuf_[1] = take();
    gbuf_[0] = take();
    gbuf_[1] = take();
    d_ = take();
    if (owl_) pg_ = take();
    S_.resize(m_);
    Y_.resize(m_);
    for (int64_t i = 0; i < m_; ++i) { S_[i] = take(); Y_[i] = take(); }
    x_spare_ = take();
    if (owl_) wp_ = (signed char *)base;

    const size_t scal_need = SLOT_COUNT * kMaxAcc + 2 * (size_t)m_;   // slots, alpha[m], ys[m]
    Scratch sc;
    rc = scratch_acquire(dev_, scal_need, &sc);
    if (rc != 0) return fail(rc, "cudaMalloc(scalars / pinned mirror / reduce workspace)");
    scal_dev_ = sc.scal_dev;
    scal_host_ = sc.scal_host;
    scal_count_ = sc.scal_count;
    ws_ = sc.ws;
    e = cudaMemsetAsync(scal_dev_, 0, sizeof(double) * scal_need, stream_);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(scalars)");
    e = cudaMemsetAsync(ws_.ticket, 0, 256, stream_);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(ticket)");
    alpha_dev_ = scal_dev_ + SLOT_COUNT * kMaxAcc;
    ys_dev_ = alpha_dev_ + m_;
    tm.lap("create: scalars + pinned + workspace");

    // evict-first accesses once the working set cannot live in L2
    const double working_set = (double)nvec * (double)vec_bytes;
    streaming_ = working_set > 0.75 * (double)dev_.l2_bytes;
    const int force = env_int("LBFGSB200_STREAMING", -1);
    if (force == 0) streaming_ = false;
    if (force == 1) streaming_ = true;
    memset(&prof_, 0, sizeof(prof_));
    graphs_enabled_ = env_int("LBFGSB200_GRAPHS", 1) != 0;
    small_enabled_ = env_int("LBFGSB200_SMALL", 1) != 0;
    speculate_ = env_int("LBFGSB200_SPECULATE", 1) != 0;
    commit_gram_enabled_ = env_int("LBFGSB200_COMMIT_GRAM", 1) != 0;
    multi_probe_max_ = env_int("LBFGSB200_MULTI_PROBE_MAX", kSpecMax);
    if (multi_probe_max_ < 1) multi_probe_max_ = 1;
    if (multi_probe_max_ > kSpecMax) multi_probe_max_ = kSpecMax;
    ring_stride_ = vec_bytes / (int64_t)sizeof(double);
    // process-wide default (lbfgsb200_set_default_direction, else LBFGSB200_DIRECTION=compact); lbfgsb200_set_direction wins
    int dflt = g_default_direction.load();
    if (dflt < 0) {
        const char *dm = getenv("LBFGSB200_DIRECTION");
        dflt = (dm && strcmp(dm, "compact") == 0) ? LBFGSB200_DIRECTION_COMPACT : LBFGSB200_DIRECTION_TWO_LOOP;
    }
    if (dflt == LBFGSB200_DIRECTION_COMPACT && m_ <= kCompactMaxM) return set_direction(LBFGSB200_DIRECTION_COMPACT);
    return 0;
}

int set_default_direction(int mode) {
    if (mode != -1 && mode != LBFGSB200_DIRECTION_TWO_LOOP && mode != LBFGSB200_DIRECTION_COMPACT) return LBFGSB200_ERR_INVALID_PARAM;
    g_default_direction.store(mode);
    return 0;
}

int Solver::set_direction(int mode) {
    if (mode != LBFGSB200_DIRECTION_TWO_LOOP && mode != LBFGSB200_DIRECTION_COMPACT)
        return fail(LBFGSB200_ERR_INVALID_PARAM, "unknown direction mode");
    if (!arena_) return fail(LBFGSB200_ERR_STATE, "solver not initialised");
    if (built_ && k_ > 1) return fail(LBFGSB200_ERR_STATE, "the direction mode cannot change in the middle of a solve");
    if (mode == LBFGSB200_DIRECTION_TWO_LOOP) { compact_ = false; return 0; }
    if (m_ > kCompactMaxM) return fail(LBFGSB200_ERR_INVALID_PARAM, "the compact direction supports m <= 32");
    if (!cmp_block_) {
        cudaError_t e = cudaSetDevice(dev_.device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
        const size_t mm = (size_t)round_up(m_ * m_, 32);
        const size_t nsums = (size_t)round_up(5 * m_ + 8, 32);
        const size_t ncoef = (size_t)round_up(1 + 2 * kCompactMaxM, 32);
        const size_t npart = (size_t)(5 * kCompactGroupMax + 2) * (size_t)ws_.stride;
        const size_t total = 2 * mm + 32 + nsums + ncoef + npart;
        e = cudaMalloc((void **)&cmp_block_, total * sizeof(double));
        if (e != cudaSuccess) { cmp_block_ = nullptr; return cuda_fail(e, "cudaMalloc(compact direction state)"); }
        e = cudaMemsetAsync(cmp_block_, 0, total * sizeof(double), stream_);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(compact direction state)");
        cmp_sy_ = cmp_block_;
        cmp_yy_ = cmp_sy_ + mm;
        cmp_sums_ = cmp_yy_ + mm + 32;   // the 32 doubles in front: the commit's history sums ride along with the all-reduce
        cmp_coefs_ = cmp_sums_ + nsums;
        cmp_partials_ = cmp_coefs_ + ncoef;
    }
    compact_ = true;
    return 0;
}

Launch Solver::launch_cfg() {
    Launch L;
    L.stream = stream_;
    L.max_grid = dev_.sm_count * dev_.blocks_per_sm;
    L.max_grid_trial = dev_.sm_count * dev_.blocks_per_sm_trial;
    L.streaming = streaming_;
    L.sequential = sequential_;
    L.ws = ws_;
    if (const PeerCtx *pc = comm_peer(comm_)) {
        L.ws.peer = *pc;
        L.peer_seq = comm_peer_seq(comm_);
    }
    L.launch_counter = &launch_counter_;
    return L;
}

// ---- instrumentation -----------------------------------------------------------------------
void Solver::prof_begin(int kind) {
    if (!timing_ || !((timing_mask_ >> kind) & 1u)) return;
    Pending p;
    p.kind = kind;
    for (cudaEvent_t *ev : {&p.a, &p.b}) {
        if (!event_pool_.empty()) { *ev = event_pool_.back(); event_pool_.pop_back(); }
        else cudaEventCreate(ev);
    }
    cudaEventRecord(p.a, stream_);
    pending_.push_back(p);
}
void Solver::prof_end(int kind, double bytes) {
    prof_.launches[kind] += 1;
    prof_.bytes[kind] += bytes;
    if (!timing_ || !((timing_mask_ >> kind) & 1u)) return;
    cudaEventRecord(pending_.back().b, stream_);
}
void Solver::prof_resolve() {
    for (auto &p : pending_) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) prof_.ms[p.kind] += ms;
        event_pool_.push_back(p.a);
        event_pool_.push_back(p.b);
    }
    pending_.clear();
}
void Solver::profile_get(lbfgsb200_profile_t *out) {
    if (timing_) { cudaStreamSynchronize(stream_); prof_resolve(); }
    *out = prof_;
}
void Solver::profile_reset() {
    if (timing_) { cudaStreamSynchronize(stream_); prof_resolve(); }
    memset(&prof_, 0, sizeof(prof_));
}

// ---- scalar plumbing ------------------------------------------------------------------------
int Solver::reduce_across_ranks(int s, int count, bool ours) {
    if (!comm_ || comm_size(comm_) == 1) return 0;
    prof_.allreduces += 1;
    if (comm_peer(comm_)) {
        if (ours) return 0;                       // summed over the ranks inside the producing kernel's epilogue
        Launch L = launch_cfg();
        launch_peer_allreduce(L, slot(s), count);  // e.g. a user objective's fused trial: 1-warp exchange kernel
        return cudaGetLastError() == cudaSuccess ? 0 : LBFGSB200_ERR_CUDA;
    }
    return comm_allreduce_sum(comm_, slot(s), count, stream_);
}

int Solver::fetch(int s, int count, double *host, bool ours) {
    int rc = reduce_across_ranks(s, count, ours);
    if (rc != 0) return fail(rc, "ncclAllReduce failed");
    cudaError_t e = cudaMemcpyAsync(scal_host_, slot(s), sizeof(double) * count, cudaMemcpyDeviceToHost, stream_);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(D2H scalars)");
    e = cudaStreamSynchronize(stream_);
    prof_.host_syncs += 1;
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    if (timing_) prof_resolve();
    for (int i = 0; i < count; ++i) host[i] = scal_host_[i];
    return check_peers(host, count);
}

// A reducing kernel that waited in vain for a peer's mailbox entry returns NaN sums and raises the communicator's
// fault word; NaN alone is a legitimate value (the objective may produce it), so the word decides.  Only looked at
// when something is NaN: no cost on the normal path.
int Solver::check_peers(const double *h, int count) {
    if (!comm_ || comm_size(comm_) == 1 || !comm_peer(comm_)) return 0;
    bool nan = false;
    for (int i = 0; i < count; ++i) nan = nan || std::isnan(h[i]);
    if (!nan || !comm_peer_fault(comm_)) return 0;
    return fail(LBFGSB200_ERR_NCCL, "a peer rank did not post its partial sums in time (dead or out of step)");
}

// Two slots (already summed over the ranks) with one synchronisation.
int Solver::fetch2(int s1, int c1, double *h1, int s2, int c2, double *h2) {
    cudaError_t e = cudaMemcpyAsync(scal_host_, slot(s1), sizeof(double) * c1, cudaMemcpyDeviceToHost, stream_);
    if (e == cudaSuccess) e = cudaMemcpyAsync(scal_host_ + kMaxAcc, slot(s2), sizeof(double) * c2, cudaMemcpyDeviceToHost, stream_);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(D2H scalars)");
    e = cudaStreamSynchronize(stream_);
    prof_.host_syncs += 1;
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    if (timing_) prof_resolve();
    for (int i = 0; i < c1; ++i) h1[i] = scal_host_[i];
    for (int i = 0; i < c2; ++i) h2[i] = scal_host_[kMaxAcc + i];
    const int rc = check_peers(h1, c1);
    return rc != 0 ? rc : check_peers(h2, c2);
}

// Every slot (already summed over the ranks) with one copy and one synchronisation: hall[slot * kMaxAcc + i].
int Solver::fetch_all(double *hall) {
    constexpr int kAll = SLOT_COUNT * kMaxAcc;
    cudaError_t e = cudaMemcpyAsync(scal_host_, scal_dev_, sizeof(double) * kAll, cudaMemcpyDeviceToHost, stream_);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(D2H scalars)");
    e = cudaStreamSynchronize(stream_);
    prof_.host_syncs += 1;
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    if (timing_) prof_resolve();
    for (int i = 0; i < kAll; ++i) hall[i] = scal_host_[i];
    return check_peers(hall + SLOT_HIST * kMaxAcc, 5);
}

// Problem::evaluate (src/core.rs:119-132) followed by the reductions the driver needs at this
// point: g.d for the line search, g.g (pg.pg) and x.x for Progress / the stop test.
bool Solver::evaluate_point(const double *d_or_null, double *dg_out) {
    const double *x = xbuf_[cur_x_];
    double *g = gbuf_[cur_g_];
    const double vbytes = 8.0 * (double)n_;
    Launch L = launch_cfg();
    double *sl = slot(SLOT_EVAL);

    prof_begin(LBFGSB200_K_EVALUATE);
    int erc = eval_(eval_user_, x, g, n_, (void *)stream_, sl + 0);
    prof_end(LBFGSB200_K_EVALUATE, 0.0);
    const bool multi = comm_ && comm_size(comm_) > 1;
    if (erc != 0 && !multi) return false;
    if (multi) {
        post_eval_flag(erc);
        // the objective's partial f (sl[0]) and the Err flag (sl[7]) ride along with the exchange that the
        // following reducing kernel performs in its epilogue (peer mailboxes); with NCCL they are part of the slot
        L.ws.peer.extra[0] = sl + 0;
        L.ws.peer.extra[1] = sl + 7;
    }

    // OWL-QN with gradient_only is the one combination that needs g.d next to the pseudo-gradient
    const bool want_gd = d_or_null != nullptr && (!owl_ || ls_.gradient_only);
    if (owl_) {
        prof_begin(LBFGSB200_K_OWL_PG);
        launch_owl_pg(L, pg_, x, g, want_gd ? d_or_null : nullptr, n_, p_.owl_c, owl_start_, owl_end_, goff_, sl + 1);
        prof_end(LBFGSB200_K_OWL_PG, (want_gd ? 4.0 : 3.0) * vbytes);
    } else {
        prof_begin(LBFGSB200_K_DOTS);
        launch_dots(L, g, want_gd ? d_or_null : nullptr, x, n_, sl + 1);
        prof_end(LBFGSB200_K_DOTS, (want_gd ? 3.0 : 2.0) * vbytes);
    }
    return finish_eval(false, /*exchanged=*/true, dg_out);   // K2 / K3 summed over the ranks in their epilogue
}

// An Err from evaluate on any rank must be seen by every rank (replicated control flow): the flag is summed
// over the ranks with the dot products.
void Solver::post_eval_flag(int erc) {
    scal_host_[kFlagWord] = erc != 0 ? 1.0 : 0.0;   // every trial ends with a stream sync, so the word is free again
    const cudaError_t e = cudaMemcpyAsync(slot(SLOT_EVAL) + 7, scal_host_ + kFlagWord, sizeof(double),
                                          cudaMemcpyHostToDevice, stream_);
    if (e != cudaSuccess) cuda_fail(e, "cudaMemcpyAsync(H2D evaluate flag)");
}

// The scalar half of an evaluation: (optional) cross-rank sum, D2H of the slot, the one host sync per trial.
// `fused`: the slot was written by the objective's fused trial (not by one of our reducing kernels).
bool Solver::finish_eval(bool fused, bool exchanged, double *dg_out) {
    const bool multi = comm_ && comm_size(comm_) > 1;
    double h[kMaxAcc];
    if (last_status_ <= LBFGSB200_ERR_CUDA) return false;   // post_eval_flag failed
    if (fetch(SLOT_EVAL, kMaxAcc, h, /*ours=*/exchanged) != 0) return false;
    if (multi && !(fused && exchanged) && h[7] != 0.0) return false;

    neval_ += 1;
    if (owl_ && !fused) {
        fx_ = h[0];
        fx_ += h[1];        // fx += x1norm, core.rs:123-124
        gg_ = h[2];
        xx_ = h[3];
        if (dg_out) *dg_out = h[4];
    } else {
        fx_ = h[0];
        if (dg_out) *dg_out = h[1];
        gg_ = h[2];
        xx_ = h[3];
    }
    return true;
}

// One line-search trial: take_line_step (core.rs:155-164) + evaluate (:119-132) + dg (:114-116).  With the
// objective's probe (+ commit) registered this is one WRITE-FREE pass over xp and d; with only its fused trial,
// one pass that also writes x and g; otherwise K1 + evaluate + K2.  Never fused for OWL-QN.
bool Solver::trial_point(const double *xp, double stp, double *dg_out) {
    const double vbytes = 8.0 * (double)n_;
    double *x = xbuf_[cur_x_];
    if (use_probe() || use_trial()) {
        const bool probe = use_probe();
        const int kind = probe ? LBFGSB200_K_PROBE : LBFGSB200_K_TRIAL_EVAL;
        prof_begin(kind);
        const int erc = probe ? fused_.probe(fused_.user, xp, d_, stp, nullptr, n_, (void *)stream_, slot(SLOT_EVAL))
                              : fused_.trial(fused_.user, xp, d_, stp, x, gbuf_[cur_g_], n_, (void *)stream_, slot(SLOT_EVAL));
        prof_end(kind, (probe ? 2.0 : 4.0) * vbytes);
        launch_counter_ += 1;
        const bool multi = comm_ && comm_size(comm_) > 1;
        const bool exchanged = !multi || fused_exchanges();
        if (erc != 0 && exchanged) return false;      // single GPU, or "fails on every rank or on none"
        if (multi && !exchanged) post_eval_flag(erc);
        return finish_eval(true, exchanged, dg_out);
    }
    Launch L = launch_cfg();
    prof_begin(LBFGSB200_K_TRIAL);
    launch_trial(L, x, xp, d_, stp, n_, owl_ ? wp_ : nullptr, owl_start_, owl_end_, goff_);
    prof_end(LBFGSB200_K_TRIAL, 3.0 * vbytes + (owl_ ? (double)n_ : 0.0));
    return evaluate_point(d_, dg_out);
}

int Solver::multi_probe_cap() const {
    if (!use_probe() || !fused_.probe_multi || multi_probe_max_ < 2) return 1;
    const bool multi = comm_ && comm_size(comm_) > 1;
    if (!multi) return multi_probe_max_;
    if (!fused_exchanges()) return 1;                      // rank-local partials would need an exchange of 4 k values
    const int fit = kMailVals / 4;                          // 4 sums per trial point in one mailbox entry
    return multi_probe_max_ < fit ? multi_probe_max_ : fit;
}

// steps[0] is the trial the search asked for; steps[1 .. k) are the ones it will ask for if it keeps extrapolating.
bool Solver::trial_multi(const double *xp, const double *steps, int k, Speculation *sp) {
    const double vbytes = 8.0 * (double)n_;
    prof_begin(LBFGSB200_K_PROBE);
    const int erc = fused_.probe_multi(fused_.user, xp, d_, steps, nullptr, k, n_, (void *)stream_, slot(SLOT_SPEC));
    prof_end(LBFGSB200_K_PROBE, 2.0 * vbytes);
    launch_counter_ += 1;
    if (erc != 0) return false;                             // single GPU, or "fails on every rank or on none"
    double h[5 * kSpecMax];
    if (last_status_ <= LBFGSB200_ERR_CUDA) return false;
    if (fetch(SLOT_SPEC, 5 * k, h, /*ours=*/true) != 0) return false;
    sp->count = k;
    for (int i = 0; i < k; ++i) {
        sp->step[i] = h[4 * k + i];
        for (int j = 0; j < 4; ++j) sp->h[i][j] = h[4 * i + j];
    }
    return true;
}

void Solver::fill_progress(lbfgsb200_progress_t *out, double step_value) const {  // core.rs:253-268
    if (!out) return;
    out->x_dev = xbuf_[cur_x_];
    out->gx_dev = gbuf_[cur_g_];
    out->n_local = n_;
    out->n_global = n_global_;
    out->fx = fx_;
    out->xnorm = std::sqrt(xx_);
    out->gnorm = std::sqrt(gg_);
    out->step = step_value;
    out->niter = k_;
    out->neval = neval_;
    out->ncall = ncall_;
}

// ---- Lbfgs::build, src/lbfgs.rs:443-481 ------------------------------------------------------
int Solver::build(double *x_dev, lbfgsb200_eval_fn eval, void *user) {
    if (!arena_) return fail(LBFGSB200_ERR_STATE, "solver not initialised");
    if (!x_dev || !eval) return fail(LBFGSB200_ERR_INVALID_PARAM, "x_dev and eval must be non-null");
    if (((uintptr_t)x_dev & 15u) != 0) return fail(LBFGSB200_ERR_INVALID_PARAM, "x_dev must be 16-byte aligned");
    cudaError_t e = cudaSetDevice(dev_.device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    if (xbuf_[0] != x_dev) drop_graphs();   // captured chains hold the caller's x pointer
    xbuf_[0] = x_dev;
    cur_x_ = 0;
    cur_g_ = 0;
    eval_ = eval;
    eval_user_ = user;
    k_ = 0;
    end_ = 0;
    ncall_ = 0;
    neval_ = 0;
    last_ls_error_ = 0;
    last_status_ = 0;
    err_.clear();
    built_ = false;
    spec_.count = 0;
    spec_k_ = 1;

    if (!evaluate_point(nullptr, nullptr)) {  // :454
        if (last_status_ != 0) return last_status_;
        return fail(LBFGSB200_ERR_EVALUATE, "evaluate failed at the initial point");
    }
    // d = -g (or -pg), :457; step = 1/||d|| * h0, :460-461
    Launch L = launch_cfg();
    prof_begin(LBFGSB200_K_INIT_DIR);
    launch_init_dir(L, d_, owl_ ? pg_ : gbuf_[cur_g_], n_, slot(SLOT_INIT));
    prof_end(LBFGSB200_K_INIT_DIR, 2.0 * 8.0 * (double)n_);
    double h[2];
    int rc = fetch(SLOT_INIT, 2, h);
    if (rc != 0) return rc;
    dginit_ = h[1];
    step_ = (1.0 / std::sqrt(h[0])) * p_.initial_inverse_hessian;
    built_ = true;
    return 0;
}

// ---- satisfying_stop_conditions, src/lbfgs.rs:697-748 -----------------------------------------
bool Solver::is_converged(int *stop_status) {
    int st = -100;
    if (p_.max_iterations != 0 && k_ >= p_.max_iterations) st = LBFGSB200_OK_MAX_ITERATIONS;
    else if (p_.max_evaluations != 0 && neval_ >= p_.max_evaluations) st = LBFGSB200_OK_MAX_EVALUATIONS;
    else if (std::sqrt(gg_) / std::fmax(std::sqrt(xx_), 1.0) <= p_.epsilon) st = LBFGSB200_OK_CONVERGED;
    if (st == -100) return false;
    if (stop_status) *stop_status = st;
    return true;
}

// ---- the update chain of one iteration: history (+ damping) + two-loop ------------------------------------------
// Everything here is enqueued without a host round trip: y.s, y.y, gamma and the damping branch are consumed
// on the device, and the host checks the history sums (`x not changed` / `gx not changed`) when it reads the
// final dot products.
//
// IterationData::update (src/lbfgs.rs:640-692).  With the objective's probe + commit registered the accepted point
// has not been written yet: the commit materialises x and g (from xp, d and the accepted trial's step) and forms
// s, y and the history sums in the same pass; otherwise k_history reads x, xp, g, gp.
int Solver::enqueue_history(const Launch &L, const double *xp, const double *gp, double step_eval) {
    const int64_t slot_new = end_;
    const bool damping = p_.damping != 0;
    const double vbytes = 8.0 * (double)n_;
    double *hist = slot(SLOT_HIST);   // {s.s, y.s, y.y, s.(-g | -pg), s.Bs}
    int rc = 0;
    gram_fused_ = 0;
    if (use_probe()) {
        int erc = LBFGSB200_ERR_UNSUPPORTED;
        const bool multi = comm_ && comm_size(comm_) > 1;
        bool hist_done = false;   // the history sums are already summed over the ranks
        if (compact_ && commit_gram_enabled_ && fused_.commit_gram && !damping && !small_eligible()) {
            // compact direction: the commit also forms the new pair's inner products with the first group of older
            // ring pairs (pass A for them), from the registers that hold s, y and g — 3 V less than commit + k_gram
            const int64_t bnd = (m_ < k_ - 1) ? m_ : (k_ - 1);
            const int nold = (int)bnd - 1;
            const int groups = nold > 0 ? (nold + kCompactGroupMax - 1) / kCompactGroupMax : 1;
            const int cnt = nold > 0 ? (nold + groups - 1) / groups : 0;
            const double *sp[kCompactGroupMax], *yp[kCompactGroupMax];
            for (int c = 0; c < cnt; ++c) {
                const int j = (int)((slot_new + m_ - (1 + c)) % m_);
                sp[c] = S_[j];
                yp[c] = Y_[j];
            }
            // N > 1 GPUs: commit_gram leaves rank-local partials; its history sums go to the eight doubles in front of
            // cmp_sums_, so ONE all-reduce covers them and the first group (and the newest pair's two sums when that
            // is all there is), then they are copied to the history slot
            double *hist_out = multi ? cmp_sums_ - kMaxAcc : hist;
            prof_begin(LBFGSB200_K_COMMIT);
            erc = fused_.commit_gram(fused_.user, xp, d_, gp, step_eval, -step_, xbuf_[cur_x_], gbuf_[cur_g_], S_[slot_new],
                                     Y_[slot_new], sp, yp, cnt, n_, (void *)L.stream, hist_out, cmp_sums_, cmp_sums_ + 5 * nold);
            if (erc == 0) {
                prof_end(LBFGSB200_K_COMMIT, (6.0 + 2.0 * cnt) * vbytes);
                launch_counter_ += 1;
                gram_fused_ = cnt > 0 ? cnt : -1;   // -1: the newest pair's own two sums only
                if (multi) {
                    prof_.allreduces += 1;
                    const int count = kMaxAcc + 5 * cnt + (cnt == nold ? 2 : 0);
                    rc = comm_allreduce_sum(comm_, hist_out, count, stream_);
                    if (rc != 0) return fail(rc, "ncclAllReduce failed");
                    const cudaError_t ce = cudaMemcpyAsync(hist, hist_out, 5 * sizeof(double), cudaMemcpyDeviceToDevice, stream_);
                    if (ce != cudaSuccess) return cuda_fail(ce, "cudaMemcpyAsync(history sums)");
                    hist_done = true;
                }
            } else {   // not offered for this shape: undo the pending event pair and run the plain commit
                if (timing_ && ((timing_mask_ >> LBFGSB200_K_COMMIT) & 1u)) {
                    event_pool_.push_back(pending_.back().a);
                    event_pool_.push_back(pending_.back().b);
                    pending_.pop_back();
                }
                cudaGetLastError();
                if (erc == LBFGSB200_ERR_UNSUPPORTED) commit_gram_enabled_ = false;
            }
        }
        if (erc != 0) {
            prof_begin(LBFGSB200_K_COMMIT);
            erc = fused_.commit(fused_.user, xp, d_, gp, step_eval, -step_, xbuf_[cur_x_], gbuf_[cur_g_], S_[slot_new],
                                Y_[slot_new], n_, (void *)L.stream, hist);
            prof_end(LBFGSB200_K_COMMIT, ((fused_.flags & LBFGSB200_FUSED_COMMIT_SKIPS_GP) ? 6.0 : 7.0) * vbytes);
            launch_counter_ += 1;
        }
        if (erc != 0) return fail(erc <= LBFGSB200_ERR_CUDA ? erc : LBFGSB200_ERR_EVALUATE, "the objective's commit failed");
        if (!hist_done) rc = reduce_across_ranks(SLOT_HIST, 5, /*ours=*/fused_exchanges());
    } else {
        prof_begin(LBFGSB200_K_HISTORY);
        launch_history(L, xbuf_[cur_x_], xp, gbuf_[cur_g_], gp, owl_ ? pg_ : nullptr, S_[slot_new], Y_[slot_new], n_, -step_,
                       damping, hist);
        prof_end(LBFGSB200_K_HISTORY, (owl_ ? 7.0 : 6.0) * vbytes);
        rc = reduce_across_ranks(SLOT_HIST, 5);
    }
    if (rc != 0) return fail(rc, "ncclAllReduce failed");
    if (damping) {  // :664-689, decided inside the kernel (case 1 rewrites y, otherwise it exits at once)
        prof_begin(LBFGSB200_K_DAMP);
        launch_damp(L, Y_[slot_new], gp, n_, -step_, hist);
        prof_end(LBFGSB200_K_DAMP, 0.0);
    }
    return 0;
}

int Solver::enqueue_two_loop(const Launch &L, const double *gp, int64_t bound, int *so_last) {
    (void)gp;
    const double vbytes = 8.0 * (double)n_;
    const double *g = gbuf_[cur_g_];
    double *hist = slot(SLOT_HIST);
    int rc = 0;
    // lbfgs_two_loop_recursion, src/lbfgs.rs:569-604, one fused kernel per trip (end = (end + 1) % m, :575)
    int64_t j = (end_ + 1) % m_;
    const double *red_in = hist + 3;              // s_new . (-g), produced by the history kernel
    int pp = 0;
    const double *dsrc = owl_ ? pg_ : g;          // d = -g | -pg, core.rs:95-101
    for (int64_t t = 0; t < bound; ++t) {
        j = (j + m_ - 1) % m_;
        const bool first = (t == 0), last = (t == bound - 1);
        const int64_t jn = (j + m_ - 1) % m_;
        const int so = pp ? SLOT_LOOP_B : SLOT_LOOP_A;
        // it.ys of the newest pair still sits in the history slot; the first trip files it into ys_dev[slot_new]
        const double *ys_in = first ? hist + 1 : ys_dev_ + j;
        prof_begin(LBFGSB200_K_BACKWARD);
        launch_backward(L, first, last, d_, dsrc, Y_[j], last ? nullptr : S_[jn], n_, red_in, ys_in,
                        first ? ys_dev_ + j : nullptr, hist, alpha_dev_ + j, slot(so));
        prof_end(LBFGSB200_K_BACKWARD, (last ? 3.0 : 4.0) * vbytes);
        rc = reduce_across_ranks(so, 1);
        if (rc != 0) return fail(rc, "ncclAllReduce failed");
        red_in = slot(so);
        pp ^= 1;
    }
    for (int64_t t = 0; t < bound; ++t) {
        const bool last = (t == bound - 1);
        const int64_t jn = (j + 1) % m_;
        const int so = pp ? SLOT_LOOP_B : SLOT_LOOP_A;
        prof_begin(LBFGSB200_K_FORWARD);
        launch_forward(L, last, owl_, d_, S_[j], last ? nullptr : Y_[jn], dsrc, n_, red_in, ys_dev_ + j, alpha_dev_ + j,
                       owl_start_, owl_end_, goff_, slot(so));
        prof_end(LBFGSB200_K_FORWARD, 4.0 * vbytes);
        rc = reduce_across_ranks(so, last ? 3 : 1);
        if (rc != 0) return fail(rc, "ncclAllReduce failed");
        red_in = slot(so);
        *so_last = so;
        pp ^= 1;
        j = jn;
    }
    return 0;
}

// The compact direction (compact.cu): the recursion of src/lbfgs.rs:569-604 with its 2 * bound scalars derived from
// inner products of the unmodified ring vectors — two passes over the ring instead of 2 * bound dependent ones.
int Solver::compact_direction(const Launch &L, int64_t bound, int *so_last) {
    const double vbytes = 8.0 * (double)n_;
    const int b = (int)bound, e = (int)end_, m = (int)m_;
    const int nold = b - 1;
    const double *src = owl_ ? pg_ : gbuf_[cur_g_];   // d0 = -g | -pg, core.rs:95-101
    auto slot_of = [&](int t) { return (e + m - t) % m; };   // the t-th newest pair
    if (nold == 0) {
        if (gram_fused_ == 0) {
            prof_begin(LBFGSB200_K_BACKWARD);
            launch_gram(L, S_[e], Y_[e], src, nullptr, nullptr, 0, true, n_, cmp_partials_, cmp_sums_);
            prof_end(LBFGSB200_K_BACKWARD, 2.0 * vbytes);
        }
    } else {
        const int groups = (nold + kCompactGroupMax - 1) / kCompactGroupMax;
        const int per = (nold + groups - 1) / groups;
        // the objective's commit_gram already produced the first group (and the newest pair's own two sums)
        for (int t0 = 1 + (gram_fused_ > 0 ? gram_fused_ : 0); t0 <= nold; t0 += per) {
            const int cnt = (nold - t0 + 1 < per) ? (nold - t0 + 1) : per;
            const double *sp[kCompactGroupMax], *yp[kCompactGroupMax];
            for (int c = 0; c < cnt; ++c) { sp[c] = S_[slot_of(t0 + c)]; yp[c] = Y_[slot_of(t0 + c)]; }
            const bool newdot = gram_fused_ == 0 && t0 + cnt > nold;   // the last group also sums y_new.d0 and y_new.y_new
            prof_begin(LBFGSB200_K_BACKWARD);
            launch_gram(L, S_[e], Y_[e], src, sp, yp, cnt, newdot, n_, cmp_partials_, cmp_sums_ + 5 * (t0 - 1));
            prof_end(LBFGSB200_K_BACKWARD, (3.0 + 2.0 * cnt) * vbytes);
        }
    }
    if (comm_ && comm_size(comm_) > 1) {   // one all-reduce for every sum of the iteration (same bits on every rank)
        // (what the commit produced: the first group, and the newest pair's two sums when there was no other group,
        // was summed right behind it)
        const int done = gram_fused_ > 0 ? gram_fused_ : 0;
        const int count = (gram_fused_ != 0 && done == nold) ? 0 : 5 * (nold - done) + 2;
        if (count > 0) {
            prof_.allreduces += 1;
            const int rc = comm_allreduce_sum(comm_, cmp_sums_ + 5 * done, count, stream_);
            if (rc != 0) return fail(rc, "ncclAllReduce failed");
        }
    }
    launch_compact_solve(L, m, b, e, cmp_sums_, slot(SLOT_HIST), cmp_sy_, cmp_yy_, ys_dev_, cmp_coefs_);
    prof_begin(LBFGSB200_K_FORWARD);
    launch_direction(L, d_, src, S_[0], ring_stride_, n_, m, b, e, cmp_coefs_, owl_, owl_start_, owl_end_, goff_,
                     slot(SLOT_LOOP_A));
    prof_end(LBFGSB200_K_FORWARD, (2.0 * b + 2.0) * vbytes);
    const int rc = reduce_across_ranks(SLOT_LOOP_A, 3);
    if (rc != 0) return fail(rc, "ncclAllReduce failed");
    *so_last = SLOT_LOOP_A;
    return 0;
}

// The compact direction in the launch-bound regime: pass A + scalar recursions + pass B in ONE cluster launch (small.cu).
int Solver::compact_small(const Launch &L, int64_t bound, int *so_last) {
    const double vbytes = 8.0 * (double)n_;
    prof_begin(LBFGSB200_K_UPDATE_SMALL);
    const cudaError_t e = launch_compact_small(L, dev_.device, n_, (int)m_, (int)bound, (int)end_, d_, owl_ ? pg_ : gbuf_[cur_g_],
                                               S_[0], ring_stride_, ys_dev_, cmp_sy_, cmp_yy_, slot(SLOT_HIST), slot(SLOT_LOOP_A),
                                               owl_, owl_start_, owl_end_, goff_, p_.max_step_size,
                                               p_.constrain_step_size != 0, slot(SLOT_STEP));
    if (e != cudaSuccess) {   // refused (cluster shape): never try again, use the three-launch form
        cudaGetLastError();
        if (timing_ && ((timing_mask_ >> LBFGSB200_K_UPDATE_SMALL) & 1u)) {
            event_pool_.push_back(pending_.back().a);
            event_pool_.push_back(pending_.back().b);
            pending_.pop_back();
        }
        small_enabled_ = false;
        return compact_direction(L, bound, so_last);
    }
    prof_end(LBFGSB200_K_UPDATE_SMALL, (4.0 * (double)bound + 2.0) * vbytes);
    *so_last = SLOT_LOOP_A;
    return 0;
}

// The launch-bound regime: all 2 * bound trips in one cluster-persistent kernel (small.cu).  One GPU, tree
// reductions; anything it cannot take (or a refused cluster launch) goes to the multi-kernel chain.
bool Solver::small_eligible() const {
    if (!small_enabled_ || sequential_ || m_ > 64) return false;
    if (comm_ && comm_size(comm_) > 1) return false;
    return n_ <= two_loop_small_max_n();
}

int Solver::two_loop_small(const Launch &L, int64_t bound, int *so_last) {
    const double vbytes = 8.0 * (double)n_;
    prof_begin(LBFGSB200_K_UPDATE_SMALL);
    const cudaError_t e = launch_two_loop_small(L, dev_.device, n_, (int)m_, (int)bound, (int)end_, d_, owl_ ? pg_ : gbuf_[cur_g_],
                                                S_[0], ring_stride_, ys_dev_, slot(SLOT_HIST), slot(SLOT_LOOP_A), owl_,
                                                owl_start_, owl_end_, goff_, p_.max_step_size, p_.constrain_step_size != 0,
                                                slot(SLOT_STEP));
    if (e != cudaSuccess) {   // refused (cluster shape / shared memory): never try again, use the chain
        cudaGetLastError();
        if (timing_ && ((timing_mask_ >> LBFGSB200_K_UPDATE_SMALL) & 1u)) {   // undo prof_begin's pending record
            event_pool_.push_back(pending_.back().a);
            event_pool_.push_back(pending_.back().b);
            pending_.pop_back();
        }
        small_enabled_ = false;
        return enqueue_two_loop(L, nullptr, bound, so_last);
    }
    prof_end(LBFGSB200_K_UPDATE_SMALL, (8.0 * (double)bound - 1.0) * vbytes);
    *so_last = SLOT_LOOP_A;
    return 0;
}

// CUDA-graph replay of enqueue_two_loop for the launch-bound regime.  The kernel arguments of the chain depend only
// on the ring position (which s/y slots), the x/g buffer parity and — fixed for a solver — the option set, so
// after the history ring is full there are at most 2m distinct chains; each is captured once and then replayed
// with a single cudaGraphLaunch (1 + 2m kernels cost ~4.5 us of CPU launch time each otherwise).
bool Solver::graph_eligible(int64_t bound) const {
    if (!graphs_enabled_ || timing_ || sequential_) return false;
    if (comm_ && comm_size(comm_) > 1) return false;                                   // exchange sequence numbers are by value
    if (bound != m_) return false;                                                     // ring still filling
    // capturing + instantiating one chain costs about as much as 6 iterations of direct launches and saves
    // ~20 us per replay (measured, profiles/README.md): only long solves get graphs
    if (k_ <= 10 * m_ + 4) return false;
    return n_ <= kGraphMaxN;
}

int Solver::two_loop_graphed(const Launch &L, const double *gp, int64_t bound, int *so_last) {
    if (!cap_stream_) {
        if (cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            graphs_enabled_ = false;
            return enqueue_two_loop(L, gp, bound, so_last);
        }
        graphs_.assign((size_t)(4 * m_), GraphEntry{});
    }
    // a chain bakes in the ring position and BOTH buffer parities (x/xp and g/gp flip together in propagate, but
    // finish() re-homes x alone)
    GraphEntry &ge = graphs_[(size_t)(end_ * 4 + cur_x_ * 2 + cur_g_)];
    if (!ge.exec) {
        const lbfgsb200_profile_t before = prof_;
        const int64_t launches_before = launch_counter_;
        Launch Lc = L;
        Lc.stream = cap_stream_;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        int rc = ok ? enqueue_two_loop(Lc, gp, bound, &ge.so_last) : 0;
        if (ok) ok = cudaStreamEndCapture(cap_stream_, &graph) == cudaSuccess && rc == 0 && graph != nullptr;
        if (ok) ok = cudaGraphInstantiate(&ge.exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        // nothing ran during the capture: take the counters back, they are added per replay below
        for (int k = 0; k < LBFGSB200_K_COUNT; ++k) {
            ge.launches[k] = prof_.launches[k] - before.launches[k];
            ge.bytes[k] = prof_.bytes[k] - before.bytes[k];
        }
        ge.kernels = launch_counter_ - launches_before;
        prof_ = before;
        launch_counter_ = launches_before;
        if (!ok) {
            cudaGetLastError();
            ge.exec = nullptr;
            graphs_enabled_ = false;
            return enqueue_two_loop(L, gp, bound, so_last);
        }
    }
    if (cudaGraphLaunch(ge.exec, stream_) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaGraphLaunch");
    for (int k = 0; k < LBFGSB200_K_COUNT; ++k) {
        prof_.launches[k] += ge.launches[k];
        prof_.bytes[k] += ge.bytes[k];
    }
    launch_counter_ += ge.kernels;
    graph_replays_ += 1;
    *so_last = ge.so_last;
    return 0;
}

void Solver::drop_graphs() {
    for (auto &ge : graphs_)
        if (ge.exec) cudaGraphExecDestroy(ge.exec);
    graphs_.clear();
    if (cap_stream_) { cudaStreamDestroy(cap_stream_); cap_stream_ = nullptr; }
}

// ---- LbfgsState::propagate, src/lbfgs.rs:503-560 ----------------------------------------------
int Solver::propagate(lbfgsb200_progress_t *out) {
    if (!built_) return fail(LBFGSB200_ERR_STATE, "propagate() before build()");
    cudaError_t ce = cudaSetDevice(dev_.device);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaSetDevice");
    k_ += 1;
    if (k_ == 1) {  // :507-510
        fill_progress(out, step_);
        return 0;
    }
    Launch L = launch_cfg();
    const double vbytes = 8.0 * (double)n_;

    // save_state (core.rs:207-210) without moving a byte: the current buffers become xp / gp
    const double *xp = xbuf_[cur_x_];
    const double *gp = gbuf_[cur_g_];
    const double xx_prev = xx_, gg_prev = gg_;
    cur_x_ ^= 1;
    cur_g_ ^= 1;
    auto revert = [&]() {  // core.rs:201-204: x, gx restored; fx and pg are not
        cur_x_ ^= 1;
        cur_g_ ^= 1;
        xx_ = xx_prev;
        if (!owl_) gg_ = gg_prev;
    };

    // LineSearch::find, src/line.rs:193-223
    LineSearchMachine ls;
    if (ls.begin(ls_, owl_, fx_, dginit_, step_) != 0) {
        revert();
        return fail(LBFGSB200_ERR_LINESEARCH, "Failure during line search");
    }
    if (owl_) {  // update_orthant_new_point, line.rs:734-736
        prof_begin(LBFGSB200_K_ORTHANT);
        launch_orthant(L, wp_, xp, pg_, n_);
        prof_end(LBFGSB200_K_ORTHANT, 2.0 * vbytes + (double)n_);
    }
    double stp = 0.0, stp_eval = 0.0;   // stp_eval: the step of the last EVALUATED trial, i.e. where x is (line.rs:396-398
                                        // leaves `stp` one update ahead when the search runs out of trials)
    Speculation spec = spec_;
    spec_.count = 0;
    const int cap = multi_probe_cap();
    while (ls.next_trial(&stp)) {
        double dg = 0.0;
        stp_eval = stp;
        bool ok;
        int hit = -1;
        if (use_probe())
            for (int i = 0; i < spec.count; ++i)
                if (spec.step[i] == stp) hit = i;
        if (hit < 0 && cap > 1 && spec_k_ > (int)ls.trials()) {   // (a search that outlives the guess goes on one trial at a time)
            // one pass for this trial and the ones the search is expected to ask for next (as many as the previous search
            // needed): they share the read of xp and d
            double steps[kSpecMax] = {};
            steps[0] = stp;
            int want = spec_k_ - ((int)ls.trials() - 1);     // trials() counts the one just handed out
            if (want > cap) want = cap;
            const int k = 1 + ls.predict(steps + 1, want - 1);
            spec.count = 0;
            if (k > 1) {
                if (trial_multi(xp, steps, k, &spec)) hit = 0;
                else if (last_status_ <= LBFGSB200_ERR_CUDA) return last_status_;
                else spec.count = 0;                        // the objective declined: probe this trial alone below
            }
        }
        if (hit >= 0) {
            // this trial was probed ahead — behind the previous update, or in the same pass as an earlier trial — and came
            // back with its scalars: no launch, no round trip
            neval_ += 1;
            fx_ = spec.h[hit][0];
            dg = spec.h[hit][1];
            gg_ = spec.h[hit][2];
            xx_ = spec.h[hit][3];
            ok = true;
        } else {
            spec.count = 0;
            ok = trial_point(xp, stp, &dg);
        }
        if (!ok && last_status_ <= LBFGSB200_ERR_CUDA) return last_status_;  // CUDA / NCCL failure is fatal
        ls.feed(ok, fx_, dg);
    }
    step_ = ls.step();
    const double step_ls = step_;
    if (ls.error() != 0 || ls.trials() == 0) {
        // line.rs:213-220: revert and report Ok(0); x == xp then fails `ensure!(d != 0.0)` (lbfgs.rs:645).
        // With max_linesearch <= 1 the loop body never runs and x was never moved: same outcome.
        revert();
        ncall_ = ls.error() != 0 ? 0 : ls.ncall();
        if (ls.error() != 0) last_ls_error_ = ls.error();
        char msg[96];
        snprintf(msg, sizeof(msg), "x not changed with step %g", step_);
        return fail(LBFGSB200_ERR_X_NOT_CHANGED, msg);
    }
    ncall_ = ls.ncall();
    spec_k_ = ncall_ < 1 ? 1 : (ncall_ > kSpecMax ? kSpecMax : (int)ncall_);   // what the next search will probably need

    // IterationData::update (src/lbfgs.rs:640-692) + lbfgs_two_loop_recursion (:569-604): 1 + 2*bound kernels with no
    // host round trip in between.  In the launch-bound regime (small n, steady state) the chain is captured once
    // per (ring position, buffer parity) into a CUDA graph and replayed with one launch.
    const bool damping = p_.damping != 0;
    const int64_t bound = (m_ < k_ - 1) ? m_ : (k_ - 1);
    int so_last = SLOT_LOOP_A;
    int rc = 0;
    rc = enqueue_history(L, xp, gp, stp_eval);
    if (rc != 0) return rc;
    bool small_ran = false;
    if (compact_ && small_eligible()) { rc = compact_small(L, bound, &so_last); small_ran = small_enabled_; }
    else if (compact_) rc = compact_direction(L, bound, &so_last);
    else if (small_eligible()) { rc = two_loop_small(L, bound, &so_last); small_ran = small_enabled_; }
    else if (graph_eligible(bound)) rc = two_loop_graphed(L, gp, bound, &so_last);
    else rc = enqueue_two_loop(L, gp, bound, &so_last);
    if (rc != 0) return rc;
    end_ = (end_ + 1) % m_;
    // The next search's first trial, speculatively: a probe writes nothing, so it can run right behind the two-loop
    // recursion with its step formed on the device (min(max_step_size, |d|) / |d|, src/lbfgs.rs:547-551) and return
    // with the update's own scalars — an iteration whose search accepts its first trial then costs ONE host
    // round trip instead of two.  Unused (the solve stops, the search wants another first step) it is just dropped.
    const bool multi = comm_ && comm_size(comm_) > 1;
    bool speculated = false;
    int spec_multi = 0;   // trial points of the speculative pass when it went through the multi-step probe
    if (speculate_ && use_probe() && (!multi || fused_exchanges())) {
        if (!small_ran)   // (the cluster kernel forms the step in its own epilogue)
            launch_next_step(L, slot(so_last), p_.max_step_size, p_.constrain_step_size != 0, slot(SLOT_STEP));
        int kq = spec_k_ < cap ? spec_k_ : cap;
        if (!ls.uses_morethuente()) kq = 1;                 // the chain is More-Thuente's extrapolation
        prof_begin(LBFGSB200_K_PROBE);
        int erc;
        if (kq > 1) {
            erc = fused_.probe_multi(fused_.user, xbuf_[cur_x_], d_, nullptr, slot(SLOT_STEP), kq, n_, (void *)stream_, slot(SLOT_SPEC));
            spec_multi = erc == 0 ? kq : 0;
        } else {
            erc = fused_.probe(fused_.user, xbuf_[cur_x_], d_, 0.0, slot(SLOT_STEP), n_, (void *)stream_, slot(SLOT_EVAL));
        }
        prof_end(LBFGSB200_K_PROBE, 2.0 * vbytes);
        launch_counter_ += 1;
        speculated = erc == 0;
    }
    // the one host round trip of the update: the history sums and the final dot products together
    double h[5], hd[3];
    Speculation next_spec;
    if (speculated) {
        double hall[SLOT_COUNT * kMaxAcc];
        rc = fetch_all(hall);
        if (rc != 0) return rc;
        for (int i = 0; i < 5; ++i) h[i] = hall[SLOT_HIST * kMaxAcc + i];
        for (int i = 0; i < 3; ++i) hd[i] = hall[so_last * kMaxAcc + i];
        if (spec_multi > 0) {
            const double *q = hall + SLOT_SPEC * kMaxAcc;
            next_spec.count = spec_multi;
            for (int i = 0; i < spec_multi; ++i) {
                next_spec.step[i] = q[4 * spec_multi + i];
                for (int j = 0; j < 4; ++j) next_spec.h[i][j] = q[4 * i + j];
            }
        } else {
            next_spec.count = 1;
            next_spec.step[0] = hall[SLOT_STEP * kMaxAcc];
            for (int i = 0; i < 4; ++i) next_spec.h[0][i] = hall[SLOT_EVAL * kMaxAcc + i];
        }
    } else {
        rc = fetch2(SLOT_HIST, 5, h, so_last, 3, hd);
        if (rc != 0) return rc;
    }
    const double ss = h[0], ys = h[1], yy = h[2], sbs = h[4];
    if (!(std::sqrt(ss) != 0.0)) {  // :645-646
        char msg[96];
        snprintf(msg, sizeof(msg), "x not changed with step %g", step_);
        return fail(LBFGSB200_ERR_X_NOT_CHANGED, msg);
    }
    if (!(yy != 0.0)) return fail(LBFGSB200_ERR_G_NOT_CHANGED, "gx not changed");  // :655
    if (damping && ys < (1.0 - 0.6) * sbs) prof_.bytes[LBFGSB200_K_DAMP] += 3.0 * vbytes;  // case 1 ran (accounting only)
    const double dnorm = std::sqrt(hd[0]);  // :543 (before the orthant projection)
    dginit_ = hd[1];
    if (std::signbit(dnorm)) return fail(LBFGSB200_ERR_INVALID_DNORM, "invalid norm value");  // :544
    if (p_.constrain_step_size) step_ = std::fmin(p_.max_step_size, dnorm) / dnorm;        // :547-551
    else step_ = 1.0;
    if (owl_ && !(std::sqrt(hd[2]) != 0.0))                                                  // orthantwise.rs:160
        return fail(LBFGSB200_ERR_OWLQN_ZERO_DIRECTION, "invalid direction vector after constraints");

    spec_ = next_spec;            // the update was sound: the speculative trial (if any) belongs to the next search
    fill_progress(out, step_ls);  // :556-557
    return 0;
}

void Solver::report(lbfgsb200_report_t *out) const {
    if (!out) return;
    out->fx = fx_;
    out->xnorm = std::sqrt(xx_);
    out->gnorm = std::sqrt(gg_);
    out->neval = neval_;
    out->niter = k_;
    out->last_ls_error = last_ls_error_;
    out->status = last_status_;
}

int Solver::finish() {
    if (cur_x_ != 0 && xbuf_[0]) {
        cudaError_t e = cudaMemcpyAsync(xbuf_[0], xbuf_[1], sizeof(double) * (size_t)n_, cudaMemcpyDeviceToDevice, stream_);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(D2D x)");
        // keep the roles consistent: the caller's buffer is current again, ours holds the copy
        cur_x_ = 0;
    }
    cudaError_t e = cudaStreamSynchronize(stream_);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    if (timing_) prof_resolve();
    return 0;
}

// ---- Lbfgs::minimize, src/lbfgs.rs:399-421 -----------------------------------------------------
int Solver::minimize(double *x_dev, lbfgsb200_eval_fn eval, void *user, lbfgsb200_progress_fn prog, void *prog_user,
                     lbfgsb200_report_t *rep) {
    const bool dbg = env_int("LBFGSB200_DEBUG_TIMING", 0) != 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    auto t0 = now();
    int rc = build(x_dev, eval, user);
    if (dbg) fprintf(stderr, "[lbfgsb200] build %.3f ms\n", ms(t0, now()));
    if (rc != 0) {
        report(rep);
        return rc;
    }
    int status = 0;
    for (;;) {
        int st = 0;
        if (is_converged(&st)) { status = st; break; }
        lbfgsb200_progress_t pr;
        auto t1 = now();
        rc = propagate(&pr);
        if (dbg) fprintf(stderr, "[lbfgsb200] propagate k=%lld ncall=%lld %.3f ms\n", (long long)k_, (long long)ncall_, ms(t1, now()));
        if (rc != 0) { status = rc; break; }
        if (prog && prog(prog_user, &pr) != 0) { status = LBFGSB200_OK_CANCELLED; break; }
    }
    auto t2 = now();
    int frc = finish();
    if (dbg) fprintf(stderr, "[lbfgsb200] finish %.3f ms\n", ms(t2, now()));
    if (frc != 0 && status >= 0) status = frc;
    last_status_ = status;
    report(rep);
    return status;
}

}  // namespace lb
