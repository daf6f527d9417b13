// compact.cu — the opt-in "compact" search direction (lbfgsb200_set_direction(LBFGSB200_DIRECTION_COMPACT)).
//
// The reference's two-loop recursion (src/lbfgs.rs:569-604) is a chain of 2 * bound DEPENDENT passes: every alpha_j /
// beta_j is a dot product with the vector as the previous trip left it, so the vector is re-read and re-written
// 2 * bound times ((8 * bound - 1) V of HBM traffic, 2 * bound grid-wide reductions, 2 * bound cross-GPU exchanges).
// The ring vectors themselves do not change between iterations, only one pair enters, so every one of those dot
// products is a linear combination of inner products of UNMODIFIED vectors:
//     s_j . q_j   = s_j . d0 - sum_{i newer than j} alpha_i (s_j . y_i)
//     y_j . r_j   = gamma (y_j . d0 - sum_i alpha_i (y_j . y_i)) + sum_{i older than j} (alpha_i - beta_i) (y_j . s_i)
// with d0 = -g (-pg).  This file keeps S^T Y and Y^T Y (bound x bound, device memory, indexed by ring slot) across
// iterations and per iteration runs
//   pass A  k_gram        one read of g and the ring: the new pair's row / column of S^T Y, Y^T Y and S^T d0, Y^T d0
//                         (5 sums per older slot, 2 for the newest), one deterministic grid reduction
//   solve   k_compact_solve   one warp: files the sums, then the two recursions on SCALARS -> -alpha_j, alpha_j - beta_j
//   pass B  k_direction   one read of g and the ring, one write of d: the element-wise operations of :589-599 in the
//                         reference's order (d = -g; d += -alpha_j y_j ...; d *= gamma; d += (alpha_j - beta_j) s_j ...),
//                         fused with d.d, g.d and the OWL-QN projection like the last forward trip of kernels.cu
// = (4 * bound + 4) V instead of (8 * bound - 1) V, 2 grid reductions instead of 2 * bound, and on N GPUs one
// all-reduce of 5 * bound - 3 doubles plus one exchange instead of 2 * bound exchanges.  Given the same scalars the
// element-wise result is bit-identical to the trips'; the scalars differ from the reference's by rounding only
// (measured: less than the drift between two summation orders of the reference's own dot products).
// The test suite's CPU checker restates this file's arithmetic; in reference-order mode (<<<1, 1>>>) the two agree bit
// for bit (tests/test_gpu_compact.py).
#include <cstdlib>

#include "kernels.h"
#include "reduce.cuh"

namespace lb {
namespace {

__device__ __forceinline__ double sgn_c(double v) { return (double)((v > 0.0) - (v < 0.0)); }

inline int threads_for(const Launch &L) { return L.sequential ? 1 : kThreads; }
inline int grid_for_u(const Launch &L, int64_t n, int U) {
    if (L.sequential) return 1;
    const int64_t nv = n >> 1;
    const int64_t tile = (int64_t)kThreads * U;
    int64_t tiles = (nv + tile - 1) / tile;
    if (tiles < 1) tiles = 1;
    if (tiles > L.max_grid) tiles = L.max_grid;
    return (int)tiles;
}

// ---------------------------------------------------------------------------------------------
// pass A: inner products of the newest pair and of d0 = -src with G older ring slots
template <int G>
struct GramPtrs {
    const double *s[G > 0 ? G : 1];
    const double *y[G > 0 ? G : 1];
};
template <bool S, int G, bool NEWDOT>
struct GramOp {
    const double *sn, *yn, *src;
    GramPtrs<G> p;
    static constexpr int kAcc = 5 * G + (NEWDOT ? 2 : 0);
    struct Regs { double2 sn, yn, g; double2 s[G > 0 ? G : 1], y[G > 0 ? G : 1]; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        if (G > 0) r.sn = ld2<S>(sn, i);
        r.yn = ld2<S>(yn, i);
        r.g = ld2<S>(src, i);
#pragma unroll
        for (int k = 0; k < G; ++k) {
            r.s[k] = ld2<S>(p.s[k], i);
            r.y[k] = ld2<S>(p.y[k], i);
        }
    }
    template <class A>
    __device__ __forceinline__ void elem(double sni, double yni, double gi, const double *sk, const double *yk, A &acc) const {
        const double ng = -gi;                              // d0 = -g | -pg, core.rs:95-101
#pragma unroll
        for (int k = 0; k < G; ++k) {
            acc[5 * k + 0] += sk[k] * ng;                   // s_k . d0
            acc[5 * k + 1] += yk[k] * ng;                   // y_k . d0
            acc[5 * k + 2] += sni * yk[k];                  // s_new . y_k
            acc[5 * k + 3] += sk[k] * yni;                  // s_k . y_new
            acc[5 * k + 4] += yni * yk[k];                  // y_new . y_k
        }
        if (NEWDOT) {
            acc[5 * G + 0] += yni * ng;                     // y_new . d0
            acc[5 * G + 1] += yni * yni;                    // y_new . y_new (of the vector as stored: after damping)
        }
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t, double (&acc)[kAcc > 0 ? kAcc : 1]) const {
        double sx[G > 0 ? G : 1], yx[G > 0 ? G : 1], sy[G > 0 ? G : 1], yy[G > 0 ? G : 1];
#pragma unroll
        for (int k = 0; k < G; ++k) { sx[k] = r.s[k].x; yx[k] = r.y[k].x; sy[k] = r.s[k].y; yy[k] = r.y[k].y; }
        elem(G > 0 ? r.sn.x : 0.0, r.yn.x, r.g.x, sx, yx, acc);
        elem(G > 0 ? r.sn.y : 0.0, r.yn.y, r.g.y, sy, yy, acc);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[kAcc > 0 ? kAcc : 1]) const {
        double sk[G > 0 ? G : 1], yk[G > 0 ? G : 1];
#pragma unroll
        for (int k = 0; k < G; ++k) { sk[k] = p.s[k][e]; yk[k] = p.y[k][e]; }
        elem(G > 0 ? sn[e] : 0.0, yn[e], src[e], sk, yk, acc);
    }
};
#ifndef LB_GRAM_U
#define LB_GRAM_U 2   // pairs per thread per tile of pass A with 3 .. 5 older ring pairs per launch (tuned: profiles/r02_tuning.md)
#endif
template <int G> struct GramU { static constexpr int value = G == 0 ? 8 : (G == 1 ? 4 : (G == 2 ? 3 : LB_GRAM_U)); };

template <bool S, int G, bool NEWDOT>
__global__ void __launch_bounds__(kThreads, 1) k_gram(GramOp<S, G, NEWDOT> op, int64_t n, ReduceWs ws, double *out) {
    constexpr int kAcc = GramOp<S, G, NEWDOT>::kAcc;
    double acc[kAcc];
#pragma unroll
    for (int a = 0; a < kAcc; ++a) acc[a] = 0.0;
    stream_pairs<kAcc, GramU<G>::value>(n, op, acc);
    grid_reduce<kAcc>(acc, ws, out);
}

// ---------------------------------------------------------------------------------------------
// the scalar recursions.  One warp; lanes stage the two matrices, lane 0 runs the (inherently serial) recurrences in
// exactly this operation order (the CPU checker of the tests restates it).
struct SolveArgs {
    int m, bound, slot_new;
    const double *sums;     // [5 * (t - 1) + c] for the t-th newest slot (t = 1 .. bound - 1), then {y_new.d0, y_new.y_new}
    const double *hist;     // {s.s, y.s, y.y, s.d0, s.Bs} of the newest pair (k_history / the objective's commit)
    double *SY, *YY;        // m x m, ring-slot indexed: SY[a * m + b] = s_a . y_b
    double *ys_dev;         // it.ys of every slot (src/lbfgs.rs:613,653)
    double *coefs;          // out: [0] = gamma, [1 + t] = -alpha_t, [1 + kCompactMaxM + t] = alpha_t - beta_t  (t-th newest)
};
__global__ void __launch_bounds__(32, 1) k_compact_solve(SolveArgs a) {
    __shared__ double sy[kCompactMaxM * kCompactMaxM], yy[kCompactMaxM * kCompactMaxM];
    __shared__ double alpha[kCompactMaxM], coef[kCompactMaxM], sg[kCompactMaxM], yg[kCompactMaxM], ysr[kCompactMaxM];
    __shared__ int slot[kCompactMaxM];
    const int m = a.m, b = a.bound, e = a.slot_new, lane = threadIdx.x;
    for (int i = lane; i < m * m; i += 32) { sy[i] = a.SY[i]; yy[i] = a.YY[i]; }
    for (int t = lane; t < b; t += 32) slot[t] = (e + m - t) % m;           // newest ... oldest
    __syncwarp();
    const int nold = b - 1;
    // file the new pair's row / column (all lanes; disjoint entries)
    for (int t = 1 + lane; t < b; t += 32) {
        const double *q = a.sums + 5 * (t - 1);
        const int j = slot[t];
        sg[t] = q[0];
        yg[t] = q[1];
        sy[e * m + j] = q[2];
        sy[j * m + e] = q[3];
        yy[e * m + j] = q[4];
        yy[j * m + e] = q[4];
        a.SY[e * m + j] = q[2];
        a.SY[j * m + e] = q[3];
        a.YY[e * m + j] = q[4];
        a.YY[j * m + e] = q[4];
        ysr[t] = a.ys_dev[j];
    }
    if (lane == 0) {
        sg[0] = a.hist[3];                                  // s_new . d0 (the first trip's numerator, :587)
        yg[0] = a.sums[5 * nold + 0];
        const double ynyn = a.sums[5 * nold + 1];
        yy[e * m + e] = ynyn;
        a.YY[e * m + e] = ynyn;
        ysr[0] = a.hist[1];                                 // it.ys of the newest pair (:653; before damping, as in the reference)
        a.ys_dev[e] = ysr[0];
    }
    __syncwarp();
    if (lane != 0) return;
    const double gamma = a.hist[1] / a.hist[2];             // :691
    for (int t = 0; t < b; ++t) {                           // backward, :582-591
        double acc = sg[t];
        const int j = slot[t];
        for (int i = 0; i < t; ++i) acc += -alpha[i] * sy[j * m + slot[i]];
        alpha[t] = acc / ysr[t];                            // :587
    }
    for (int t = b - 1; t >= 0; --t) {                      // forward, :594-601
        double acc = yg[t];
        const int j = slot[t];
        for (int i = 0; i < b; ++i) acc += -alpha[i] * yy[j * m + slot[i]];
        acc = acc * gamma;                                  // :591
        for (int i = b - 1; i > t; --i) acc += coef[i] * sy[slot[i] * m + j];
        const double beta = acc / ysr[t];                   // :597
        coef[t] = alpha[t] - beta;                          // :599
    }
    a.coefs[0] = gamma;
    for (int t = 0; t < b; ++t) {
        a.coefs[1 + t] = -alpha[t];
        a.coefs[1 + kCompactMaxM + t] = coef[t];
    }
}

// ---------------------------------------------------------------------------------------------
// pass B: d from g and the ring, element-wise in the reference's order
struct DirArgs {
    double *d;
    const double *src;          // g | pg
    const double *ring;         // S_0; S_j = ring + 2 j stride, Y_j = ring + (2 j + 1) stride
    int64_t stride;
    int64_t n;
    int m, bound, slot_new;
    const double *coefs;
    int64_t start, end, goff;
    int q_lo, q_hi;             // k_direction_gen: the steps [q_lo, q_hi) of the 2 * bound (y newest -> oldest, s oldest -> newest)
};
// bound == BT at compile time (1 .. 8): the loops over the ring are fully unrolled, every load of a tile in flight at once.
template <bool S, bool OWL, int BT, int U, int C>
__global__ void __launch_bounds__(kThreads, 1) k_direction(DirArgs a, ReduceWs ws, double *out) {
    __shared__ double nal[kCompactMaxM], cf[kCompactMaxM];
    __shared__ const double *yp[kCompactMaxM], *sp[kCompactMaxM];
    const int b = BT;
    for (int t = threadIdx.x; t < b; t += blockDim.x) {
        const int j = (a.slot_new + a.m - t) % a.m;
        nal[t] = __ldcg(a.coefs + 1 + t);
        cf[t] = __ldcg(a.coefs + 1 + kCompactMaxM + t);
        sp[t] = a.ring + (int64_t)(2 * j) * a.stride;
        yp[t] = a.ring + (int64_t)(2 * j + 1) * a.stride;
    }
    const double gamma = __ldcg(a.coefs);
    __syncthreads();
    double acc[3] = {0.0, 0.0, 0.0};
    auto finish = [&](int64_t e, double v, double gi) -> double {
        acc[0] += v * v;                                    // dnorm^2 before projection, lbfgs.rs:543
        if (OWL) {
            const int64_t gidx = a.goff + e;
            if (gidx >= a.start && gidx < a.end && sgn_c(v) != sgn_c(-gi)) v = 0.0;   // orthantwise.rs:140-147
            acc[2] += v * v;                                // ||d|| after projection, :160
        }
        acc[1] += gi * v;                                   // next dginit: g.d | pg.d, core.rs:78-92
        return v;
    };
    const int64_t nv = a.n >> 1;
    const int64_t T = blockDim.x;                           // 1 in reference-order mode: pairs in index order
    const int64_t tile = T * U;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < nv; base += (int64_t)gridDim.x * tile) {
        double2 g[U], v[U];
        bool in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * T + threadIdx.x;
            in[u] = i < nv;
            if (in[u]) g[u] = ld2<S>(a.src, i);
            else g[u] = make_double2(0.0, 0.0);
            v[u].x = -g[u].x;                               // vecncpy, core.rs:99
            v[u].y = -g[u].y;
        }
        // backward element-wise passes, newest -> oldest (:589)
#pragma unroll
        for (int t0 = 0; t0 < BT; t0 += C) {
            double2 w[C][U];
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (t0 + c < b && in[u]) w[c][u] = ld2<S>(yp[t0 + c], base + u * T + threadIdx.x);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (t0 + c < b) {
                    const double na = nal[t0 + c];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (in[u]) {
                            v[u].x = v[u].x + na * w[c][u].x;
                            v[u].y = v[u].y + na * w[c][u].y;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {                       // vecscale(gamma), :591
            v[u].x = v[u].x * gamma;
            v[u].y = v[u].y * gamma;
        }
        // forward element-wise passes, oldest -> newest (:599)
#pragma unroll
        for (int t0 = 0; t0 < BT; t0 += C) {
            double2 w[C][U];
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (t0 + c < b && in[u]) w[c][u] = ld2<S>(sp[b - 1 - (t0 + c)], base + u * T + threadIdx.x);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (t0 + c < b) {
                    const double co = cf[b - 1 - (t0 + c)];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (in[u]) {
                            v[u].x = v[u].x + co * w[c][u].x;
                            v[u].y = v[u].y + co * w[c][u].y;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (in[u]) {
                const int64_t i = base + u * T + threadIdx.x;
                double2 o;
                o.x = finish(2 * i, v[u].x, g[u].x);
                o.y = finish(2 * i + 1, v[u].y, g[u].y);
                st2<S>(a.d, i, o);
            }
        }
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // odd trailing element
        const int64_t e = a.n - 1;
        const double gi = a.src[e];
        double v = -gi;
        for (int t = 0; t < b; ++t) v = v + nal[t] * yp[t][e];
        v = v * gamma;
        for (int t = b - 1; t >= 0; --t) v = v + cf[t] * sp[t][e];
        a.d[e] = finish(e, v, gi);
    }
    grid_reduce<3>(acc, ws, out);
}

// Any bound (9 .. kCompactMaxM): the 2 * bound ring vectors are walked as ONE sequence of steps (y newest -> oldest,
// the scale by gamma, s oldest -> newest), C vectors per chunk, the NEXT chunk's loads issued before the current chunk
// is consumed (two register sets, the loop unrolled by two: no register moves), two co-resident CTAs per SM.
// A launch covers the steps [q_lo, q_hi): a long recursion CAN run as a few sub-passes that hand the partly built
// vector on through d (LBFGSB200_COMPACT_SPLIT = ring vectors per sub-pass).  Measured on B200 the single pass wins at
// every depth — 41 vectors open per CTA at m = 20 still stream at 6.96 TB/s, a split costs its extra read + write of d
// (profiles/r02_tuning.md) — so the default is one pass and the knob stays for other parts.
template <bool S, bool OWL, int U, int C>
__global__ void __launch_bounds__(kThreads, 2) k_direction_gen(DirArgs a, ReduceWs ws, double *out) {
    __shared__ double co[2 * kCompactMaxM];
    __shared__ const double *vp[2 * kCompactMaxM];
    const int b = a.bound, nq = 2 * a.bound, q_lo = a.q_lo, q_hi = a.q_hi;
    const bool first = q_lo == 0, final = q_hi == nq;
    for (int q = q_lo + threadIdx.x; q < q_hi; q += blockDim.x) {
        const int t = q < b ? q : nq - 1 - q;               // the t-th newest pair
        const int j = (a.slot_new + a.m - t) % a.m;
        co[q] = q < b ? __ldcg(a.coefs + 1 + t) : __ldcg(a.coefs + 1 + kCompactMaxM + t);
        vp[q] = a.ring + (int64_t)(2 * j + (q < b ? 1 : 0)) * a.stride;
    }
    const double gamma = __ldcg(a.coefs);
    __syncthreads();
    double acc[3] = {0.0, 0.0, 0.0};
    auto finish = [&](int64_t e, double v, double gi) -> double {
        acc[0] += v * v;                                    // dnorm^2 before projection, lbfgs.rs:543
        if (OWL) {
            const int64_t gidx = a.goff + e;
            if (gidx >= a.start && gidx < a.end && sgn_c(v) != sgn_c(-gi)) v = 0.0;   // orthantwise.rs:140-147
            acc[2] += v * v;                                // ||d|| after projection, :160
        }
        acc[1] += gi * v;                                   // next dginit: g.d | pg.d, core.rs:78-92
        return v;
    };
    const int64_t nv = a.n >> 1;
    const int64_t T = blockDim.x;
    const int64_t tile = T * U;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < nv; base += (int64_t)gridDim.x * tile) {
        double2 g[U], v[U], wa[C][U], wb[C][U];
        bool in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * T + threadIdx.x;
            in[u] = i < nv;
            g[u] = make_double2(0.0, 0.0);
            v[u] = make_double2(0.0, 0.0);
            if (in[u] && (first || final)) g[u] = ld2<S>(a.src, i);
            if (in[u] && !first) v[u] = ld2<S>(a.d, i);
        }
        auto load = [&](double2 (&w)[C][U], int q0) {
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (q0 + c < q_hi && in[u]) w[c][u] = ld2<S>(vp[q0 + c], base + u * T + threadIdx.x);
        };
        auto consume = [&](double2 (&w)[C][U], int q0) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int q = q0 + c;
                if (q < q_hi) {
                    if (q == b) {                           // vecscale(gamma) between the two loops, :591
#pragma unroll
                        for (int u = 0; u < U; ++u) { v[u].x = v[u].x * gamma; v[u].y = v[u].y * gamma; }
                    }
                    const double cq = co[q];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        v[u].x = v[u].x + cq * w[c][u].x;   // vecadd, :589 / :599
                        v[u].y = v[u].y + cq * w[c][u].y;
                    }
                }
            }
        };
        load(wa, q_lo);
        if (first) {
#pragma unroll
            for (int u = 0; u < U; ++u) { v[u].x = -g[u].x; v[u].y = -g[u].y; }   // vecncpy, core.rs:99
        }
        for (int q0 = q_lo; q0 < q_hi; q0 += 2 * C) {
            load(wb, q0 + C);
            consume(wa, q0);
            load(wa, q0 + 2 * C);
            consume(wb, q0 + C);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (in[u]) {
                const int64_t i = base + u * T + threadIdx.x;
                double2 o = v[u];
                if (final) {
                    o.x = finish(2 * i, v[u].x, g[u].x);
                    o.y = finish(2 * i + 1, v[u].y, g[u].y);
                }
                st2<S>(a.d, i, o);
            }
        }
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // odd trailing element
        const int64_t e = a.n - 1;
        const double gi = a.src[e];
        double v = first ? -gi : a.d[e];
        for (int q = q_lo; q < q_hi; ++q) {
            if (q == b) v = v * gamma;
            v = v + co[q] * vp[q][e];
        }
        a.d[e] = final ? finish(e, v, gi) : v;
    }
    if (final) grid_reduce<3>(acc, ws, out);
}

inline ReduceWs ws_peer(const Launch &L) {   // a reducing launch that exchanges over the ranks in its epilogue
    ReduceWs ws = L.ws;
    if (ws.peer.nranks > 1 && L.peer_seq) ws.peer.seq = ++*L.peer_seq;
    else ws.peer.nranks = 0;
    return ws;
}

template <bool S, int G, bool NEWDOT>
void gram_launch(const Launch &L, const double *sn, const double *yn, const double *src, const GramPtrs<G> &p, int64_t n,
                 const ReduceWs &ws, double *out) {
    GramOp<S, G, NEWDOT> op{sn, yn, src, p};
    k_gram<S, G, NEWDOT><<<grid_for_u(L, n, GramU<G>::value), threads_for(L), 0, L.stream>>>(op, n, ws, out);
}
template <int G>
void gram_group(const Launch &L, bool newdot, const double *sn, const double *yn, const double *src, const double *const *s,
                const double *const *y, int64_t n, const ReduceWs &ws, double *out) {
    GramPtrs<G> p{};
    for (int k = 0; k < G; ++k) { p.s[k] = s[k]; p.y[k] = y[k]; }
    if constexpr (G == 0) {   // the newest pair alone: always with its own two sums
        if (L.streaming) gram_launch<true, 0, true>(L, sn, yn, src, p, n, ws, out);
        else gram_launch<false, 0, true>(L, sn, yn, src, p, n, ws, out);
    } else if (L.streaming) {
        if (newdot) gram_launch<true, G, true>(L, sn, yn, src, p, n, ws, out);
        else gram_launch<true, G, false>(L, sn, yn, src, p, n, ws, out);
    } else {
        if (newdot) gram_launch<false, G, true>(L, sn, yn, src, p, n, ws, out);
        else gram_launch<false, G, false>(L, sn, yn, src, p, n, ws, out);
    }
}

// Ring vectors per sub-pass of k_direction_gen (LBFGSB200_COMPACT_SPLIT overrides; 0 = never split).
static int compact_split() {
    static const int v = [] { const char *e = getenv("LBFGSB200_COMPACT_SPLIT"); return (e && *e) ? atoi(e) : kCompactSplitDefault; }();
    return v;
}

template <bool S, bool OWL>
void direction_launch(const Launch &L, const DirArgs &a, const ReduceWs &ws, double *out) {
    const int th = threads_for(L);
#define LB_DIR(BT, U, C) k_direction<S, OWL, BT, U, C><<<grid_for_u(L, a.n, U), th, 0, L.stream>>>(a, ws, out)
    switch (a.bound) {
        case 1: LB_DIR(1, 4, 1); break;
        case 2: LB_DIR(2, 4, 2); break;
        case 3: LB_DIR(3, 3, 3); break;
        case 4: LB_DIR(4, 2, 4); break;
        case 5: LB_DIR(5, 2, 5); break;
        case 6: LB_DIR(6, 2, 6); break;
        case 7: LB_DIR(7, 2, 7); break;
        case 8: LB_DIR(8, 2, 8); break;
        default: {   // two co-resident CTAs per SM: twice the grid of the one-CTA kernels
            int grid = L.sequential ? 1 : 2 * grid_for_u(L, a.n, 2);
            const int64_t tiles = ((a.n >> 1) + (int64_t)kThreads * 2 - 1) / ((int64_t)kThreads * 2);
            if (!L.sequential && grid > tiles) grid = (int)(tiles < 1 ? 1 : tiles);
            const int nq = 2 * a.bound, split = compact_split();
            int parts = (split > 0 && !L.sequential) ? (nq + split - 1) / split : 1;
            if (parts < 1) parts = 1;
            const int per = (nq + parts - 1) / parts;
            for (int lo = 0; lo < nq; lo += per) {
                DirArgs sub = a;
                sub.q_lo = lo;
                sub.q_hi = lo + per < nq ? lo + per : nq;
                k_direction_gen<S, OWL, 2, 4><<<grid, th, 0, L.stream>>>(sub, ws, out);
            }
            break;
        }
    }
#undef LB_DIR
}

}  // namespace

// Pass A for the t-th newest slots [t0, t0 + cnt) (1 <= t0; cnt <= kCompactGroupMax, may be 0 when the ring holds the
// newest pair only).  `newdot`: this launch also produces {y_new.d0, y_new.y_new} at out[5 * cnt ..].
// wide_partials: [5 * kCompactGroupMax + 2][L.ws.stride] doubles of device memory.
void launch_gram(const Launch &L, const double *s_new, const double *y_new, const double *src, const double *const *s,
                 const double *const *y, int cnt, bool newdot, int64_t n, double *wide_partials, double *out) {
    ReduceWs ws = L.ws;
    ws.partials = wide_partials;
    ws.peer.nranks = 0;          // the caller all-reduces the whole sums array at once
    if (L.launch_counter) ++*L.launch_counter;
    switch (cnt) {
        case 0: gram_group<0>(L, true, s_new, y_new, src, s, y, n, ws, out); break;
        case 1: gram_group<1>(L, newdot, s_new, y_new, src, s, y, n, ws, out); break;
        case 2: gram_group<2>(L, newdot, s_new, y_new, src, s, y, n, ws, out); break;
        case 3: gram_group<3>(L, newdot, s_new, y_new, src, s, y, n, ws, out); break;
        case 4: gram_group<4>(L, newdot, s_new, y_new, src, s, y, n, ws, out); break;
        default: gram_group<5>(L, newdot, s_new, y_new, src, s, y, n, ws, out); break;
    }
}

void launch_compact_solve(const Launch &L, int m, int bound, int slot_new, const double *sums, const double *hist,
                          double *SY, double *YY, double *ys_dev, double *coefs) {
    if (L.launch_counter) ++*L.launch_counter;
    SolveArgs a{m, bound, slot_new, sums, hist, SY, YY, ys_dev, coefs};
    k_compact_solve<<<1, 32, 0, L.stream>>>(a);
}

// Pass B.  out = {d.d before projection, g.d | pg.d, d.d after projection} (summed over the ranks in the epilogue).
void launch_direction(const Launch &L, double *d, const double *src, const double *ring, int64_t stride, int64_t n, int m,
                      int bound, int slot_new, const double *coefs, bool owl, int64_t start, int64_t end, int64_t goff,
                      double *out) {
    if (L.launch_counter) ++*L.launch_counter;
    DirArgs a{d, src, ring, stride, n, m, bound, slot_new, coefs, start, end, goff, 0, 2 * bound};
    const ReduceWs ws = ws_peer(L);
    if (L.streaming) {
        if (owl) direction_launch<true, true>(L, a, ws, out);
        else direction_launch<true, false>(L, a, ws, out);
    } else {
        if (owl) direction_launch<false, true>(L, a, ws, out);
        else direction_launch<false, false>(L, a, ws, out);
    }
}

}  // namespace lb
