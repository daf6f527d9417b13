// kernels.cu — hand-written f64 sm_100a kernels for the L-BFGS / OWL-QN hot path.
//
// Every kernel is a single streaming pass (HBM-bound, <= 0.25 flop/byte — no tensor cores): 128-bit
// coalesced loads, U independent loads per vector in flight per thread, fused element-wise update
// + up to 5 dot products, deterministic two-level reduction (reduce.cuh).  Built with -fmad=false
// so each element-wise result is bit-identical to the reference's scalar Rust (rustc never
// contracts a*b+c); the arithmetic below keeps the reference's operation order.
//
// Scalars that only the device knows (the two-loop's alpha_j / beta_j) never visit the host: the
// producing kernel leaves its dot product in a scalar slot, and the consuming kernel's prologue
// reads it (after the optional cross-rank all-reduce) and derives alpha / beta itself.
#include "kernels.h"
#include "reduce.cuh"

namespace lb {
namespace {

__device__ __forceinline__ double sgn(double v) {  // orthantwise.rs:174-180: 0 for 0 and NaN
    return (double)((v > 0.0) - (v < 0.0));
}

inline int grid_for(const Launch &L, int64_t n, int U, bool trial_family = false) {
    if (L.sequential) return 1;
    const int64_t nv = n >> 1;
    const int64_t tile = (int64_t)kThreads * U;
    int64_t tiles = (nv + tile - 1) / tile;
    const int64_t cap = trial_family ? L.max_grid_trial : L.max_grid;
    if (tiles < 1) tiles = 1;
    if (tiles > cap) tiles = cap;
    return (int)tiles;
}
// The reduction workspace of one reducing launch: with peers, this launch gets the next exchange sequence number
// (every rank issues the same launches in the same order, so the numbers agree).
inline ReduceWs ws_for(const Launch &L) {
    ReduceWs ws = L.ws;
    if (ws.peer.nranks > 1 && L.peer_seq) ws.peer.seq = ++*L.peer_seq;
    else ws.peer.nranks = 0;
    return ws;
}
inline int threads_for(const Launch &L) { return L.sequential ? 1 : kThreads; }
inline void count(const Launch &L) {
    if (L.launch_counter) ++*L.launch_counter;
}

// ---------------------------------------------------------------------------------------------
// K2 dots: {g.d, g.g, x.x}
template <bool S, bool HAS_D>
struct DotsOp {
    const double *g, *d, *x;
    struct Regs { double2 g, d, x; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.g = ld2<S>(g, i);
        if (HAS_D) r.d = ld2<S>(d, i);
        r.x = ld2<S>(x, i);
    }
    __device__ __forceinline__ void elem(double gi, double di, double xi, double (&acc)[3]) const {
        if (HAS_D) acc[0] += gi * di;
        acc[1] += gi * gi;
        acc[2] += xi * xi;
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t, double (&acc)[3]) const {
        elem(r.g.x, HAS_D ? r.d.x : 0.0, r.x.x, acc);
        elem(r.g.y, HAS_D ? r.d.y : 0.0, r.x.y, acc);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[3]) const {
        elem(g[e], HAS_D ? d[e] : 0.0, x[e], acc);
    }
};
template <bool S, bool HAS_D>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_dots(DotsOp<S, HAS_D> op, int64_t n, ReduceWs ws, double *out) {
    double acc[3] = {0.0, 0.0, 0.0};
    stream_pairs<3, kUt>(n, op, acc);   // kUt: same element -> thread map as the fused trial kernel
    grid_reduce<3>(acc, ws, out);
}

// ---------------------------------------------------------------------------------------------
// K3 OWL-QN pseudo-gradient + l1 norm + norms
template <bool S, bool HAS_D>
struct OwlPgOp {
    double *pg;
    const double *x, *g, *d;
    double c;
    int64_t start, end, goff;
    struct Regs { double2 x, g, d; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.x = ld2<S>(x, i);
        r.g = ld2<S>(g, i);
        if (HAS_D) r.d = ld2<S>(d, i);
    }
    __device__ __forceinline__ double elem(int64_t e, double xi, double gi, double di, double (&acc)[4]) const {
        const int64_t gidx = goff + e;
        double p;
        if (gidx >= start && gidx < end) {
            acc[0] += c * fabs(xi);                         // x1norm, orthantwise.rs:74-76
            if (xi != 0.0) {                                // :95-97
                const double sg = (xi > 0.0) ? 1.0 : ((xi < 0.0) ? -1.0 : xi);
                p = gi + sg * c;
            } else {                                        // :98-108
                const double right_partial = gi + c;
                const double left_partial = gi - c;
                p = (right_partial < 0.0) ? right_partial : ((left_partial > 0.0) ? left_partial : 0.0);
            }
        } else {
            p = gi;                                         // :85-87,110-112
        }
        acc[1] += p * p;
        acc[2] += xi * xi;
        if (HAS_D) acc[3] += gi * di;
        return p;
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[4]) const {
        double2 p;
        p.x = elem(2 * i, r.x.x, r.g.x, HAS_D ? r.d.x : 0.0, acc);
        p.y = elem(2 * i + 1, r.x.y, r.g.y, HAS_D ? r.d.y : 0.0, acc);
        st2<S>(pg, i, p);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[4]) const {
        pg[e] = elem(e, x[e], g[e], HAS_D ? d[e] : 0.0, acc);
    }
};
template <bool S, bool HAS_D>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_owl_pg(OwlPgOp<S, HAS_D> op, int64_t n, ReduceWs ws, double *out) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    stream_pairs<4, HAS_D ? kUh : kU>(n, op, acc);   // 2R 1W (3R 1W with d): the tile shapes of the same traffic classes
    grid_reduce<4>(acc, ws, out);
}

// ---------------------------------------------------------------------------------------------
// K0 init direction: d = -src; {d.d, src.d}
template <bool S>
struct InitDirOp {
    double *d;
    const double *src;
    struct Regs { double2 v; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const { r.v = ld2<S>(src, i); }
    __device__ __forceinline__ double elem(double v, double (&acc)[2]) const {
        const double di = -v;
        acc[0] += di * di;
        acc[1] += v * di;
        return di;
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[2]) const {
        double2 o;
        o.x = elem(r.v.x, acc);
        o.y = elem(r.v.y, acc);
        st2<S>(d, i, o);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[2]) const { d[e] = elem(src[e], acc); }
};
template <bool S>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_init_dir(InitDirOp<S> op, int64_t n, ReduceWs ws, double *out) {
    double acc[2] = {0.0, 0.0};
    stream_pairs<2, kU>(n, op, acc);
    grid_reduce<2>(acc, ws, out);
}

// ---------------------------------------------------------------------------------------------
// K1 trial step: x = xp + step*d (+ orthant projection)
template <bool S, bool OWL>
struct TrialOp {
    double *x;
    const double *xp, *d;
    const signed char *wp;
    double step;
    int64_t start, end, goff;
    struct Regs { double2 xp, d; char2 w; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.xp = ld2<S>(xp, i);
        r.d = ld2<S>(d, i);
        if (OWL) r.w = reinterpret_cast<const char2 *>(wp)[i];
    }
    __device__ __forceinline__ double elem(int64_t e, double xpi, double di, signed char w) const {
        double v = xpi + step * di;                         // veccpy + vecadd, core.rs:156-157
        if (OWL) {
            const int64_t gidx = goff + e;
            if (gidx >= start && gidx < end && sgn(v) != (double)w) v = 0.0;  // orthantwise.rs:165-171
        }
        return v;
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&)[1]) const {
        double2 o;
        o.x = elem(2 * i, r.xp.x, r.d.x, OWL ? r.w.x : (signed char)0);
        o.y = elem(2 * i + 1, r.xp.y, r.d.y, OWL ? r.w.y : (signed char)0);
        st2<S>(x, i, o);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&)[1]) const {
        x[e] = elem(e, xp[e], d[e], OWL ? wp[e] : (signed char)0);
    }
};
template <bool S, bool OWL>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_trial(TrialOp<S, OWL> op, int64_t n) {
    double acc[1] = {0.0};
    stream_pairs<1, kU>(n, op, acc);
}

// ---------------------------------------------------------------------------------------------
// K4 orthant: wp = xp == 0 ? signum(-pg) : signum(xp)   (int8)
template <bool S>
struct OrthantOp {
    signed char *wp;
    const double *xp, *pg;
    struct Regs { double2 xp, pg; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.xp = ld2<S>(xp, i);
        r.pg = ld2<S>(pg, i);
    }
    __device__ __forceinline__ signed char elem(double xpi, double pgi) const {
        return (signed char)((xpi == 0.0) ? sgn(-pgi) : sgn(xpi));
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&)[1]) const {
        char2 o;
        o.x = elem(r.xp.x, r.pg.x);
        o.y = elem(r.xp.y, r.pg.y);
        reinterpret_cast<char2 *>(wp)[i] = o;
    }
    __device__ __forceinline__ void tail(int64_t e, double (&)[1]) const { wp[e] = elem(xp[e], pg[e]); }
};
template <bool S>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_orthant(OrthantOp<S> op, int64_t n) {
    double acc[1] = {0.0};
    stream_pairs<1, kU>(n, op, acc);
}

// ---------------------------------------------------------------------------------------------
// K5 history update: s = x - xp; y = g - gp; {s.s, y.s, y.y, s.(-g | -pg), s.(gp*nstep)}
template <bool S, bool DAMP, bool OWL>
struct HistoryOp {
    const double *x, *xp, *g, *gp, *pg;
    double *s, *y;
    double nstep;
    struct Regs { double2 x, xp, g, gp, pg; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.x = ld2<S>(x, i);
        r.xp = ld2<S>(xp, i);
        r.g = ld2<S>(g, i);
        r.gp = ld2<S>(gp, i);
        if (OWL) r.pg = ld2<S>(pg, i);
    }
    __device__ __forceinline__ void elem(double xi, double xpi, double gi, double gpi, double pgi, double &si,
                                         double &yi, double (&acc)[5]) const {
        history_elem<DAMP, OWL>(xi, xpi, gi, gpi, pgi, nstep, si, yi, acc);
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[5]) const {
        double2 so, yo;
        elem(r.x.x, r.xp.x, r.g.x, r.gp.x, OWL ? r.pg.x : 0.0, so.x, yo.x, acc);
        elem(r.x.y, r.xp.y, r.g.y, r.gp.y, OWL ? r.pg.y : 0.0, so.y, yo.y, acc);
        st2<S>(s, i, so);
        st2<S>(y, i, yo);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[5]) const {
        double so, yo;
        elem(x[e], xp[e], g[e], gp[e], OWL ? pg[e] : 0.0, so, yo, acc);
        s[e] = so;
        y[e] = yo;
    }
};
template <bool S, bool DAMP, bool OWL>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_history(HistoryOp<S, DAMP, OWL> op, int64_t n, ReduceWs ws, double *out) {
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    stream_pairs<5, kUh>(n, op, acc);
    grid_reduce<5>(acc, ws, out);
}

// ---------------------------------------------------------------------------------------------
// K6 Powell damping: y = ((gp*nstep)*omt) + theta*y
template <bool S>
struct DampOp {
    double *y;
    const double *gp;
    double nstep, omt, theta;
    struct Regs { double2 y, gp; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.y = ld2<S>(y, i);
        r.gp = ld2<S>(gp, i);
    }
    __device__ __forceinline__ double elem(double yi, double gpi) const {
        double bs = gpi * nstep;                            // lbfgs.rs:670-671
        bs = bs * omt;                                      // :678
        return bs + theta * yi;                             // :679-680
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&)[1]) const {
        double2 o;
        o.x = elem(r.y.x, r.gp.x);
        o.y = elem(r.y.y, r.gp.y);
        st2<S>(y, i, o);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&)[1]) const { y[e] = elem(y[e], gp[e]); }
};
// The damping branch (src/lbfgs.rs:664-689) is taken on the device from the history kernel's sums, so the host
// does not have to synchronise between the history update and the two-loop: hist = {s.s, y.s, y.y, s.(-g), s.Bs}.
template <bool S>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_damp(DampOp<S> op, int64_t n, const double *hist) {
    const double sigma2 = 0.6;                              // :664 (sigma3 = 3.0 selects case 2, which leaves y alone)
    const double ys = __ldcg(hist + 1), sbs = __ldcg(hist + 4);
    if (!(ys < (1.0 - sigma2) * sbs)) return;               // not case 1: y stays as it is (:681-689)
    const double theta = sigma2 * sbs / (sbs - ys);         // :676
    op.omt = 1.0 - theta;
    op.theta = theta;
    double acc[1] = {0.0};
    stream_pairs<1, kU>(n, op, acc);
}

// ---------------------------------------------------------------------------------------------
// K7 two-loop backward step
template <bool S, bool FIRST, bool LAST>
struct BackwardOp {
    double *q;              // in/out (out only when FIRST)
    const double *g;        // FIRST: q_in = -g
    const double *y;        // y_j
    const double *snext;    // s_{j-1} (!LAST)
    double nalpha, gamma;
    struct Regs { double2 q, y, s; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.q = ld2<S>(FIRST ? g : q, i);
        r.y = ld2<S>(y, i);
        if (!LAST) r.s = ld2<S>(snext, i);
    }
    __device__ __forceinline__ double elem(double qi, double yi, double si, double (&acc)[1]) const {
        if (FIRST) qi = -qi;                                // vecncpy, core.rs:99
        double v = qi + nalpha * yi;                        // vecadd(y, -alpha), lbfgs.rs:589
        if (LAST) {
            v = v * gamma;                                  // vecscale(gamma), :591
            acc[0] += yi * v;                               // y_j . d for the first beta, :597
        } else {
            acc[0] += si * v;                               // s_{j-1} . q for the next alpha, :587
        }
        return v;
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[1]) const {
        double2 o;
        o.x = elem(r.q.x, r.y.x, LAST ? 0.0 : r.s.x, acc);
        o.y = elem(r.q.y, r.y.y, LAST ? 0.0 : r.s.y, acc);
        st2<S>(q, i, o);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[1]) const {
        q[e] = elem(FIRST ? g[e] : q[e], y[e], LAST ? 0.0 : snext[e], acc);
    }
};
template <bool S, bool FIRST, bool LAST>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_backward(BackwardOp<S, FIRST, LAST> op, int64_t n, const double *red_in, const double *ys_in, double *ys_store,
           const double *hist, double *alpha_out, ReduceWs ws, double *out) {
    const double ys_j = __ldcg(ys_in);                      // it.ys of slot j: a device scalar, never on the host path
    const double alpha = __ldcg(red_in) / ys_j;             // lbfgs.rs:587
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *alpha_out = alpha;
        if (ys_store) *ys_store = ys_j;                     // the newest pair: keep y.s for the next m iterations (:653)
    }
    op.nalpha = -alpha;
    if (LAST) op.gamma = __ldcg(hist + 1) / __ldcg(hist + 2);  // gamma = ys / yy of the newest pair (:691)
    double acc[1] = {0.0};
    stream_pairs<1, kU>(n, op, acc);
    grid_reduce<1>(acc, ws, out);
}

// ---------------------------------------------------------------------------------------------
// K8 two-loop forward step
template <bool S, bool LAST, bool OWL>
struct ForwardOp {
    double *r;              // in/out
    const double *s;        // s_j
    const double *aux;      // !LAST: y_{j+1};  LAST: g (or pg when OWL)
    double coef;
    int64_t start, end, goff;
    struct Regs { double2 r, s, a; };
    __device__ __forceinline__ void load(Regs &v, int64_t i) const {
        v.r = ld2<S>(r, i);
        v.s = ld2<S>(s, i);
        v.a = ld2<S>(aux, i);
    }
    __device__ __forceinline__ double elem(int64_t e, double ri, double si, double ai, double (&acc)[3]) const {
        double v = ri + coef * si;                          // vecadd(s, alpha - beta), lbfgs.rs:599
        if (!LAST) {
            acc[0] += ai * v;                               // y_{j+1} . r for the next beta, :597
        } else {
            acc[0] += v * v;                                // dnorm^2 before projection, :543
            if (OWL) {
                const int64_t gidx = goff + e;
                if (gidx >= start && gidx < end && sgn(v) != sgn(-ai)) v = 0.0;  // orthantwise.rs:140-147
                acc[2] += v * v;                            // ||d|| after projection, :160
            }
            acc[1] += ai * v;                               // next dginit: g.d or pg.d, core.rs:78-92
        }
        return v;
    }
    __device__ __forceinline__ void apply(Regs &v, int64_t i, double (&acc)[3]) const {
        double2 o;
        o.x = elem(2 * i, v.r.x, v.s.x, v.a.x, acc);
        o.y = elem(2 * i + 1, v.r.y, v.s.y, v.a.y, acc);
        st2<S>(r, i, o);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[3]) const { r[e] = elem(e, r[e], s[e], aux[e], acc); }
};
template <bool S, bool LAST, bool OWL>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_forward(ForwardOp<S, LAST, OWL> op, int64_t n, const double *red_in, const double *ys_in, const double *alpha_in,
          ReduceWs ws, double *out) {
    const double beta = __ldcg(red_in) / __ldcg(ys_in);     // lbfgs.rs:597
    op.coef = __ldcg(alpha_in) - beta;                      // :599
    double acc[3] = {0.0, 0.0, 0.0};
    stream_pairs<3, kU>(n, op, acc);
    grid_reduce<3>(acc, ws, out);
}

// ---------------------------------------------------------------------------------------------
// stand-alone OWL-QN direction projection
template <bool S>
struct OwlConstrainOp {
    double *d;
    const double *pg;
    int64_t start, end, goff;
    struct Regs { double2 d, pg; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.d = ld2<S>(d, i);
        r.pg = ld2<S>(pg, i);
    }
    __device__ __forceinline__ double elem(int64_t e, double di, double pgi, double (&acc)[1]) const {
        const int64_t gidx = goff + e;
        if (gidx >= start && gidx < end && sgn(di) != sgn(-pgi)) di = 0.0;
        acc[0] += di * di;
        return di;
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&acc)[1]) const {
        double2 o;
        o.x = elem(2 * i, r.d.x, r.pg.x, acc);
        o.y = elem(2 * i + 1, r.d.y, r.pg.y, acc);
        st2<S>(d, i, o);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[1]) const { d[e] = elem(e, d[e], pg[e], acc); }
};
template <bool S>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_owl_constrain(OwlConstrainOp<S> op, int64_t n, ReduceWs ws, double *out) {
    double acc[1] = {0.0};
    stream_pairs<1, kU>(n, op, acc);
    grid_reduce<1>(acc, ws, out);
}

// ---------------------------------------------------------------------------------------------
// K11 LbfgsMath primitives (src/math.rs:31-82), unfused
enum { P_ADD = 0, P_SCALE = 1, P_CPY = 2, P_NCPY = 3, P_DIFF = 4 };
template <bool S, int KIND>
struct PrimOp {
    double *out;            // y (or z)
    const double *a, *b;    // x (, y)
    double c;
    struct Regs { double2 a, b; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        if (KIND != P_SCALE) r.a = ld2<S>(a, i);
        if (KIND == P_ADD || KIND == P_SCALE) r.b = ld2<S>(out, i);
        if (KIND == P_DIFF) r.b = ld2<S>(b, i);
    }
    __device__ __forceinline__ double elem(double ai, double bi) const {
        if (KIND == P_ADD) return bi + c * ai;              // math.rs:33-37
        if (KIND == P_SCALE) return bi * c;                 // :45-49
        if (KIND == P_CPY) return ai;                       // :52-56
        if (KIND == P_NCPY) return -ai;                     // :59-63
        return ai - bi;                                     // :66-70
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t i, double (&)[1]) const {
        double2 o;
        o.x = elem(r.a.x, r.b.x);
        o.y = elem(r.a.y, r.b.y);
        st2<S>(out, i, o);
    }
    __device__ __forceinline__ void tail(int64_t e, double (&)[1]) const {
        const double ai = (KIND != P_SCALE) ? a[e] : 0.0;
        const double bi = (KIND == P_ADD || KIND == P_SCALE) ? out[e] : ((KIND == P_DIFF) ? b[e] : 0.0);
        out[e] = elem(ai, bi);
    }
};
template <bool S, int KIND>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_prim(PrimOp<S, KIND> op, int64_t n) {
    double acc[1] = {0.0};
    stream_pairs<1, kU>(n, op, acc);
}

template <bool S>
struct DotOp {
    const double *x, *y;
    struct Regs { double2 x, y; };
    __device__ __forceinline__ void load(Regs &r, int64_t i) const {
        r.x = ld2<S>(x, i);
        r.y = ld2<S>(y, i);
    }
    __device__ __forceinline__ void apply(Regs &r, int64_t, double (&acc)[1]) const {
        acc[0] += r.x.x * r.y.x;
        acc[0] += r.x.y * r.y.y;
    }
    __device__ __forceinline__ void tail(int64_t e, double (&acc)[1]) const { acc[0] += x[e] * y[e]; }
};
template <bool S>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_dot(DotOp<S> op, int64_t n, ReduceWs ws, double *out) {
    double acc[1] = {0.0};
    stream_pairs<1, kU>(n, op, acc);
    grid_reduce<1>(acc, ws, out);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
#define LB_DISPATCH_S(L, EXPR_TRUE, EXPR_FALSE) \
    do {                                        \
        if ((L).streaming) { EXPR_TRUE; } else { EXPR_FALSE; } \
    } while (0)

void launch_dots(const Launch &L, const double *g, const double *d, const double *x, int64_t n, double *out) {
    const int grid = grid_for(L, n, kUt, /*trial_family=*/true);
    count(L);
    if (d) {
        LB_DISPATCH_S(L, (k_dots<true, true><<<grid, threads_for(L), 0, L.stream>>>({g, d, x}, n, ws_for(L), out)),
                      (k_dots<false, true><<<grid, threads_for(L), 0, L.stream>>>({g, d, x}, n, ws_for(L), out)));
    } else {
        LB_DISPATCH_S(L, (k_dots<true, false><<<grid, threads_for(L), 0, L.stream>>>({g, d, x}, n, ws_for(L), out)),
                      (k_dots<false, false><<<grid, threads_for(L), 0, L.stream>>>({g, d, x}, n, ws_for(L), out)));
    }
}

void launch_owl_pg(const Launch &L, double *pg, const double *x, const double *g, const double *d, int64_t n,
                   double c, int64_t start, int64_t end, int64_t goff, double *out) {
    const int grid = grid_for(L, n, d ? kUh : kU);
    count(L);
    if (d) {
        LB_DISPATCH_S(L,
                      (k_owl_pg<true, true><<<grid, threads_for(L), 0, L.stream>>>({pg, x, g, d, c, start, end, goff}, n, ws_for(L), out)),
                      (k_owl_pg<false, true><<<grid, threads_for(L), 0, L.stream>>>({pg, x, g, d, c, start, end, goff}, n, ws_for(L), out)));
    } else {
        LB_DISPATCH_S(L,
                      (k_owl_pg<true, false><<<grid, threads_for(L), 0, L.stream>>>({pg, x, g, d, c, start, end, goff}, n, ws_for(L), out)),
                      (k_owl_pg<false, false><<<grid, threads_for(L), 0, L.stream>>>({pg, x, g, d, c, start, end, goff}, n, ws_for(L), out)));
    }
}

void launch_init_dir(const Launch &L, double *d, const double *src, int64_t n, double *out) {
    const int grid = grid_for(L, n, kU);
    count(L);
    LB_DISPATCH_S(L, (k_init_dir<true><<<grid, threads_for(L), 0, L.stream>>>({d, src}, n, ws_for(L), out)),
                  (k_init_dir<false><<<grid, threads_for(L), 0, L.stream>>>({d, src}, n, ws_for(L), out)));
}

void launch_trial(const Launch &L, double *x, const double *xp, const double *d, double step, int64_t n,
                  const signed char *wp, int64_t start, int64_t end, int64_t goff) {
    const int grid = grid_for(L, n, kU);
    count(L);
    if (wp) {
        LB_DISPATCH_S(L, (k_trial<true, true><<<grid, threads_for(L), 0, L.stream>>>({x, xp, d, wp, step, start, end, goff}, n)),
                      (k_trial<false, true><<<grid, threads_for(L), 0, L.stream>>>({x, xp, d, wp, step, start, end, goff}, n)));
    } else {
        LB_DISPATCH_S(L, (k_trial<true, false><<<grid, threads_for(L), 0, L.stream>>>({x, xp, d, wp, step, start, end, goff}, n)),
                      (k_trial<false, false><<<grid, threads_for(L), 0, L.stream>>>({x, xp, d, wp, step, start, end, goff}, n)));
    }
}

void launch_orthant(const Launch &L, signed char *wp, const double *xp, const double *pg, int64_t n) {
    const int grid = grid_for(L, n, kU);
    count(L);
    LB_DISPATCH_S(L, (k_orthant<true><<<grid, threads_for(L), 0, L.stream>>>({wp, xp, pg}, n)),
                  (k_orthant<false><<<grid, threads_for(L), 0, L.stream>>>({wp, xp, pg}, n)));
}

template <bool S>
static void history_impl(const Launch &L, const double *x, const double *xp, const double *g, const double *gp,
                         const double *pg, double *s, double *y, int64_t n, double nstep, bool damping, double *out,
                         int grid) {
#define LB_HIST(D, O) \
    k_history<S, D, O><<<grid, threads_for(L), 0, L.stream>>>({x, xp, g, gp, pg, s, y, nstep}, n, ws_for(L), out)
    if (damping && pg) LB_HIST(true, true);
    else if (damping) LB_HIST(true, false);
    else if (pg) LB_HIST(false, true);
    else LB_HIST(false, false);
#undef LB_HIST
}

void launch_history(const Launch &L, const double *x, const double *xp, const double *g, const double *gp,
                    const double *pg, double *s, double *y, int64_t n, double nstep, bool damping, double *out) {
    const int grid = grid_for(L, n, kUh);
    count(L);
    if (L.streaming) history_impl<true>(L, x, xp, g, gp, pg, s, y, n, nstep, damping, out, grid);
    else history_impl<false>(L, x, xp, g, gp, pg, s, y, n, nstep, damping, out, grid);
}

void launch_damp(const Launch &L, double *y, const double *gp, int64_t n, double nstep, const double *hist) {
    const int grid = grid_for(L, n, kU);
    count(L);
    LB_DISPATCH_S(L, (k_damp<true><<<grid, threads_for(L), 0, L.stream>>>({y, gp, nstep, 0.0, 0.0}, n, hist)),
                  (k_damp<false><<<grid, threads_for(L), 0, L.stream>>>({y, gp, nstep, 0.0, 0.0}, n, hist)));
}

template <bool S>
static void backward_impl(const Launch &L, bool first, bool last, double *q, const double *g, const double *y,
                          const double *s_next, int64_t n, const double *red_in, const double *ys_in, double *ys_store,
                          const double *hist, double *alpha_out, double *out, int grid) {
#define LB_BWD(F, LA) \
    k_backward<S, F, LA><<<grid, threads_for(L), 0, L.stream>>>({q, g, y, s_next, 0.0, 0.0}, n, red_in, ys_in, ys_store, hist, alpha_out, ws_for(L), out)
    if (first && last) LB_BWD(true, true);
    else if (first) LB_BWD(true, false);
    else if (last) LB_BWD(false, true);
    else LB_BWD(false, false);
#undef LB_BWD
}

void launch_backward(const Launch &L, bool first, bool last, double *q, const double *g, const double *y,
                     const double *s_next, int64_t n, const double *red_in, const double *ys_in, double *ys_store,
                     const double *hist, double *alpha_out, double *out) {
    const int grid = grid_for(L, n, kU);
    count(L);
    if (L.streaming) backward_impl<true>(L, first, last, q, g, y, s_next, n, red_in, ys_in, ys_store, hist, alpha_out, out, grid);
    else backward_impl<false>(L, first, last, q, g, y, s_next, n, red_in, ys_in, ys_store, hist, alpha_out, out, grid);
}

template <bool S>
static void forward_impl(const Launch &L, bool last, bool owl, double *r, const double *s, const double *aux,
                         int64_t n, const double *red_in, const double *ys_j, const double *alpha_in, int64_t start,
                         int64_t end, int64_t goff, double *out, int grid) {
#define LB_FWD(LA, OW) \
    k_forward<S, LA, OW><<<grid, threads_for(L), 0, L.stream>>>({r, s, aux, 0.0, start, end, goff}, n, red_in, ys_j, alpha_in, ws_for(L), out)
    if (last && owl) LB_FWD(true, true);
    else if (last) LB_FWD(true, false);
    else LB_FWD(false, false);
#undef LB_FWD
}

void launch_forward(const Launch &L, bool last, bool owl, double *r, const double *s, const double *y_next,
                    const double *g_or_pg, int64_t n, const double *red_in, const double *ys_j, const double *alpha_in,
                    int64_t start, int64_t end, int64_t goff, double *out) {
    const int grid = grid_for(L, n, kU);
    count(L);
    const double *aux = last ? g_or_pg : y_next;
    if (L.streaming) forward_impl<true>(L, last, owl, r, s, aux, n, red_in, ys_j, alpha_in, start, end, goff, out, grid);
    else forward_impl<false>(L, last, owl, r, s, aux, n, red_in, ys_j, alpha_in, start, end, goff, out, grid);
}

void launch_owl_constrain(const Launch &L, double *d, const double *pg, int64_t n, int64_t start, int64_t end,
                          int64_t goff, double *out) {
    const int grid = grid_for(L, n, kU);
    count(L);
    LB_DISPATCH_S(L, (k_owl_constrain<true><<<grid, threads_for(L), 0, L.stream>>>({d, pg, start, end, goff}, n, ws_for(L), out)),
                  (k_owl_constrain<false><<<grid, threads_for(L), 0, L.stream>>>({d, pg, start, end, goff}, n, ws_for(L), out)));
}

template <int KIND>
static void prim_impl(const Launch &L, double *out, const double *a, const double *b, double c, int64_t n) {
    const int grid = grid_for(L, n, kU);
    count(L);
    LB_DISPATCH_S(L, (k_prim<true, KIND><<<grid, threads_for(L), 0, L.stream>>>({out, a, b, c}, n)),
                  (k_prim<false, KIND><<<grid, threads_for(L), 0, L.stream>>>({out, a, b, c}, n)));
}
void launch_vecadd(const Launch &L, double *y, const double *x, double c, int64_t n) { prim_impl<P_ADD>(L, y, x, nullptr, c, n); }
void launch_vecscale(const Launch &L, double *y, double c, int64_t n) { prim_impl<P_SCALE>(L, y, nullptr, nullptr, c, n); }
void launch_veccpy(const Launch &L, double *y, const double *x, int64_t n, bool negate) {
    if (negate) prim_impl<P_NCPY>(L, y, x, nullptr, 0.0, n);
    else prim_impl<P_CPY>(L, y, x, nullptr, 0.0, n);
}
void launch_vecdiff(const Launch &L, double *z, const double *x, const double *y, int64_t n) { prim_impl<P_DIFF>(L, z, x, y, 0.0, n); }
void launch_vecdot(const Launch &L, const double *x, const double *y, int64_t n, double *out) {
    const int grid = grid_for(L, n, kU);
    count(L);
    LB_DISPATCH_S(L, (k_dot<true><<<grid, threads_for(L), 0, L.stream>>>({x, y}, n, ws_for(L), out)),
                  (k_dot<false><<<grid, threads_for(L), 0, L.stream>>>({x, y}, n, ws_for(L), out)));
}

__global__ void k_next_step(const double *dots, double max_step, int constrain, double *step_out) {
    const double dnorm = sqrt(__ldcg(dots));                                  // lbfgs.rs:543
    *step_out = constrain ? fmin(max_step, dnorm) / dnorm : 1.0;              // :547-551, the host's own expression
}
void launch_next_step(const Launch &L, const double *dots, double max_step, bool constrain, double *step_out) {
    count(L);
    k_next_step<<<1, 1, 0, L.stream>>>(dots, max_step, constrain ? 1 : 0, step_out);
}

__global__ void k_peer_allreduce(PeerCtx pc, double *buf, int count) {
    __shared__ double tab[kMaxPeers + 1][kMailVals];
    if (threadIdx.x < count) tab[kMaxPeers][threadIdx.x] = buf[threadIdx.x];
    peer_allreduce(pc, count, tab);
    if (threadIdx.x < count) buf[threadIdx.x] = tab[kMaxPeers][threadIdx.x];
}
void launch_peer_allreduce(const Launch &L, double *buf, int count) {
    ReduceWs ws = L.ws;
    if (ws.peer.nranks <= 1 || !L.peer_seq) return;
    ws.peer.seq = ++*L.peer_seq;
    if (L.launch_counter) ++*L.launch_counter;
    k_peer_allreduce<<<1, 32, 0, L.stream>>>(ws.peer, buf, count);
}

}  // namespace lb
