// small.cu — the launch-bound regime (n up to ~2e5): the WHOLE two-loop recursion in one kernel.
//
// At n <= ~1e5 a vector is a few hundred KB: every kernel of the update chain runs for about as long as it takes to
// launch it, and the 2m dependent launches of lbfgs_two_loop_recursion (src/lbfgs.rs:569-604) cost 2m x (launch
// gap + drain) regardless of n.  Here ONE thread-block cluster runs all 2m trips:
//   * q (the vector the recursion rewrites 2m times) never leaves shared memory: each CTA of the cluster keeps its
//     contiguous slice of q in its own SMEM for the whole kernel; s_j / y_j stream in from L2 once per trip;
//   * the 2m dependent dot products are reduced thread -> warp -> CTA -> cluster with a fixed-shape tree: each CTA
//     publishes its partial in its own SMEM, one hardware cluster barrier (~0.2 us), then EVERY CTA sums the
//     partials of all ranks in rank order through DSMEM — all CTAs hold the same bits, no second barrier
//     (the partial slots are double-buffered across trips);
//   * alpha_j, beta_j, y_j.s_j and gamma stay on chip; the device ring ys[] is updated for later iterations.
// Element-wise arithmetic is that of k_backward / k_forward (-fmad=false, the reference's operation order), so
// the result differs from the multi-kernel chain by the summation tree only; the tree depends on (n, cluster
// size) alone: deterministic run to run.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "kernels.h"
#include "reduce.cuh"

namespace cg = cooperative_groups;

namespace lb {
namespace {

constexpr int kSmallThreads = 512;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int kSmallMaxM = 64;
constexpr int kSmallMaxCluster = 16;
constexpr int kChunk = 8;   // elements per thread whose loads are issued together

__device__ __forceinline__ double sgn_small(double v) { return (double)((v > 0.0) - (v < 0.0)); }

// CTA-level then cluster-level sum of NACC accumulators; every thread of every CTA returns the same totals.
template <int NACC>
__device__ __forceinline__ void cluster_sum(double (&acc)[NACC], double (*warp_part)[kSmallWarps], double (*cta_part)[4],
                                            int parity, cg::cluster_group &cluster, int nctas) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        const double v = warp_sum(acc[a]);
        if (lane == 0) warp_part[a][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            double v = (lane < kSmallWarps) ? warp_part[a][lane] : 0.0;
#pragma unroll
            for (int off = kSmallWarps / 2; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0) cta_part[parity][a] = v;
        }
    }
    if (nctas > 1) cluster.sync();   // partials of this trip are visible cluster-wide (also a CTA barrier)
    else __syncthreads();
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double v = 0.0;
        for (int r = 0; r < nctas; ++r) {   // fixed rank order: identical bits in every CTA
            const double(*remote)[4] = (nctas > 1) ? cluster.map_shared_rank(cta_part, r) : cta_part;
            v += remote[parity][a];
        }
        acc[a] = v;
    }
}

struct SmallArgs {
    int64_t n;
    int m, bound, slot_new;
    double *d;               // out: the new search direction
    const double *dsrc;      // g, or pg for OWL-QN: the recursion starts from -dsrc (src/core.rs:95-101)
    double *ring;            // S_0, Y_0, S_1, Y_1, ... each `stride` doubles apart
    int64_t stride;
    double *ys_dev;          // y_j.s_j of every ring slot (device ring, kept for later iterations)
    const double *hist;      // {s.s, y.s, y.y, s.(-g | -pg), s.Bs} of the newest pair (k_history / the commit)
    double *out;             // {d.d before projection, g.d | pg.d, d.d after projection}
    int owl;
    int64_t start, end, goff;
    double max_step;         // the next search's first step, formed here: constrain ? min(max_step, |d|) / |d| : 1
    int constrain;
    double *step_out;
};

__global__ void __launch_bounds__(kSmallThreads, 1) k_two_loop_small(SmallArgs a) {
    extern __shared__ __align__(16) double q[];             // this CTA's slice of q
    __shared__ double warp_part[3][kSmallWarps];
    __shared__ double cta_part[2][4];
    __shared__ double alpha_s[kSmallMaxM], ys_s[kSmallMaxM];
    cg::cluster_group cluster = cg::this_cluster();
    const int nctas = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    // contiguous slices of even length (OWL-QN and the tails need no special casing: scalar accesses)
    int64_t per = (a.n + nctas - 1) / nctas;
    per = (per + 1) & ~(int64_t)1;
    const int64_t lo = (int64_t)rank * per;
    const int64_t cnt = (lo >= a.n) ? 0 : ((a.n - lo < per) ? (a.n - lo) : per);
    const int m = a.m;
    if (tid < m) ys_s[tid] = a.ys_dev[tid];
    __syncthreads();
    const double ys_new = a.hist[1];
    const double gamma = ys_new / a.hist[2];                // ys / yy of the newest pair (src/lbfgs.rs:691)
    int parity = 0;
    double red = a.hist[3];                                 // s_new . (-g): the first alpha's numerator

    // ---- backward loop (src/lbfgs.rs:582-591) ----
    int j = (a.slot_new + 1) % m;
    for (int t = 0; t < a.bound; ++t) {
        j = (j + m - 1) % m;
        const bool first = (t == 0), last = (t == a.bound - 1);
        const int jn = (j + m - 1) % m;
        const double ys_j = first ? ys_new : ys_s[j];
        const double alpha = red / ys_j;                    // :587
        if (tid == 0) {
            alpha_s[j] = alpha;
            if (first) ys_s[j] = ys_j;
            if (first && rank == 0) a.ys_dev[j] = ys_j;     // keep y.s of the newest pair for the next m iterations (:653)
        }
        const double nalpha = -alpha;
        // read-only for the whole kernel (the ring and g were written by earlier launches): ld.global.nc, so the
        // loads of an unrolled body are issued together instead of waiting behind the shared-memory stores of q
        const double *__restrict__ y = a.ring + (int64_t)(2 * j + 1) * a.stride + lo;
        const double *__restrict__ sn = a.ring + (int64_t)(2 * jn) * a.stride + lo;
        const double *__restrict__ src = a.dsrc + lo;
        double acc[1] = {0.0};
        // kChunk elements per thread at a time, ALL their loads issued (predicated) before the first use: a thread's
        // few elements cost one L2 latency per chunk instead of one per element (a partially unrolled loop runs
        // its remainder iterations — here usually all of them — one load-use at a time)
        for (int64_t i0 = tid; i0 < cnt; i0 += (int64_t)kChunk * kSmallThreads) {
            double yv[kChunk], sv[kChunk], qv[kChunk];
#pragma unroll
            for (int c = 0; c < kChunk; ++c) {
                const int64_t i = i0 + (int64_t)c * kSmallThreads;
                const bool in = i < cnt;
                yv[c] = in ? __ldg(y + i) : 0.0;
                sv[c] = (in && !last) ? __ldg(sn + i) : 0.0;
                qv[c] = in ? (first ? -__ldg(src + i) : q[i]) : 0.0;   // vecncpy, core.rs:99
            }
#pragma unroll
            for (int c = 0; c < kChunk; ++c) {
                const int64_t i = i0 + (int64_t)c * kSmallThreads;
                if (i < cnt) {
                    double v = qv[c] + nalpha * yv[c];      // vecadd(y, -alpha), :589
                    if (last) {
                        v = v * gamma;                      // vecscale(gamma), :591
                        acc[0] += yv[c] * v;                // y_j . d for the first beta, :597
                    } else {
                        acc[0] += sv[c] * v;                // s_{j-1} . q for the next alpha, :587
                    }
                    q[i] = v;
                }
            }
        }
        cluster_sum<1>(acc, warp_part, cta_part, parity, cluster, nctas);
        parity ^= 1;
        red = acc[0];
    }
    // ---- forward loop (src/lbfgs.rs:594-601) ----
    for (int t = 0; t < a.bound; ++t) {
        const bool last = (t == a.bound - 1);
        const int jn = (j + 1) % m;
        const double beta = red / ys_s[j];                  // :597
        const double coef = alpha_s[j] - beta;              // :599
        const double *__restrict__ s = a.ring + (int64_t)(2 * j) * a.stride + lo;
        const double *__restrict__ aux = last ? a.dsrc + lo : a.ring + (int64_t)(2 * jn + 1) * a.stride + lo;
        double acc[3] = {0.0, 0.0, 0.0};
        double *__restrict__ dout = a.d + lo;
        for (int64_t i0 = tid; i0 < cnt; i0 += (int64_t)kChunk * kSmallThreads) {
            double sv[kChunk], av[kChunk];
#pragma unroll
            for (int c = 0; c < kChunk; ++c) {
                const int64_t i = i0 + (int64_t)c * kSmallThreads;
                const bool in = i < cnt;
                sv[c] = in ? __ldg(s + i) : 0.0;
                av[c] = in ? __ldg(aux + i) : 0.0;
            }
#pragma unroll
            for (int c = 0; c < kChunk; ++c) {
                const int64_t i = i0 + (int64_t)c * kSmallThreads;
                if (i < cnt) {
                    double v = q[i] + coef * sv[c];         // vecadd(s, alpha - beta), :599
                    if (!last) {
                        acc[0] += av[c] * v;                // y_{j+1} . r for the next beta, :597
                        q[i] = v;
                    } else {
                        acc[0] += v * v;                    // dnorm^2 before projection, :543
                        if (a.owl) {
                            const int64_t gidx = a.goff + lo + i;
                            if (gidx >= a.start && gidx < a.end && sgn_small(v) != sgn_small(-av[c])) v = 0.0;  // orthantwise.rs:140-147
                            acc[2] += v * v;                // ||d|| after projection, :160
                        }
                        acc[1] += av[c] * v;                // next dginit: g.d or pg.d, core.rs:78-92
                        dout[i] = v;
                    }
                }
            }
        }
        if (last) cluster_sum<3>(acc, warp_part, cta_part, parity, cluster, nctas);
        else cluster_sum<1>(reinterpret_cast<double(&)[1]>(acc), warp_part, cta_part, parity, cluster, nctas);
        parity ^= 1;
        red = acc[0];
        if (last && rank == 0 && tid == 0) {
            a.out[0] = acc[0];
            a.out[1] = acc[1];
            a.out[2] = acc[2];
            const double dnorm = sqrt(acc[0]);                                // lbfgs.rs:543
            *a.step_out = a.constrain ? fmin(a.max_step, dnorm) / dnorm : 1.0;   // :547-551
        }
        j = jn;
    }
    if (nctas > 1) cluster.sync();   // no CTA may exit while a peer still reads its partials through DSMEM
}

// ---- the compact search direction in the launch-bound regime -----------------------------------------------------
// csrc/compact.cu's pass A + scalar recursions + pass B in ONE cluster kernel: the recursion's 2 * bound dependent
// reductions become TWO cluster-wide reductions (the 5 bound - 3 inner products of pass A, then d.d / g.d), every CTA
// runs the scalar recursions redundantly on the same totals (same bits everywhere, no broadcast), and the ring is read
// from L2 twice with all loads of an element independent.  Same element-wise arithmetic as k_gram / k_direction.
constexpr int kCsMaxSums = 5 * (kCompactMaxM - 1) + 2;
constexpr int kCsThreads = 256;   // 255 registers per thread: 27 running sums + two pairs of elements of 13 vectors in flight
constexpr int kCsWarps = kCsThreads / 32;

struct CompactSmallArgs {
    int64_t n;
    int m, bound, slot_new;
    double *d;
    const double *dsrc;      // g | pg
    const double *ring;      // S_0; S_j = ring + 2 j stride, Y_j = ring + (2 j + 1) stride
    int64_t stride;
    double *ys_dev, *SY, *YY;   // device state kept across iterations (csrc/compact.cu)
    const double *hist;      // {s.s, y.s, y.y, s.d0, s.Bs} of the newest pair
    double *out;             // {d.d before projection, g.d | pg.d, d.d after projection}
    int owl;
    int64_t start, end, goff;
    double max_step;
    int constrain;
    double *step_out;
};

// sums of one group of G older pairs over this CTA's slice -> acc (registers), block-reduced into part[off ..]
template <int G, bool NEWDOT>
__device__ __forceinline__ void cs_gram_group(const CompactSmallArgs &a, int64_t lo, int64_t cnt, const int *slot, int t0,
                                              double (*red)[kCsWarps], double *part, int off, int newdot_off) {
    constexpr int kAcc = 5 * G + (NEWDOT ? 2 : 0);
    double acc[kAcc > 0 ? kAcc : 1];
#pragma unroll
    for (int q = 0; q < kAcc; ++q) acc[q] = 0.0;
    const int e = a.slot_new;
    const double *__restrict__ sn = a.ring + (int64_t)(2 * e) * a.stride + lo;
    const double *__restrict__ yn = a.ring + (int64_t)(2 * e + 1) * a.stride + lo;
    const double *__restrict__ src = a.dsrc + lo;
    const double *sp[G > 0 ? G : 1], *yp[G > 0 ? G : 1];
#pragma unroll
    for (int k = 0; k < G; ++k) {
        sp[k] = a.ring + (int64_t)(2 * slot[t0 + k]) * a.stride + lo;
        yp[k] = a.ring + (int64_t)(2 * slot[t0 + k] + 1) * a.stride + lo;
    }
    // slices start on even elements of 256-byte aligned vectors: 128-bit loads; two pairs of elements per trip, all
    // their loads (2 x (3 + 2 G)) issued before the first use
    struct Regs { double2 s, y, g, sk[G > 0 ? G : 1], yk[G > 0 ? G : 1]; };
    auto load = [&](Regs &r, int64_t p) {
        r.s = G > 0 ? __ldg(reinterpret_cast<const double2 *>(sn) + p) : make_double2(0.0, 0.0);
        r.y = __ldg(reinterpret_cast<const double2 *>(yn) + p);
        r.g = __ldg(reinterpret_cast<const double2 *>(src) + p);
#pragma unroll
        for (int k = 0; k < G; ++k) {
            r.sk[k] = __ldg(reinterpret_cast<const double2 *>(sp[k]) + p);
            r.yk[k] = __ldg(reinterpret_cast<const double2 *>(yp[k]) + p);
        }
    };
    auto apply = [&](const Regs &r) {
        const double ngx = -r.g.x, ngy = -r.g.y;            // d0 = -g | -pg, core.rs:95-101
#pragma unroll
        for (int k = 0; k < G; ++k) {
            acc[5 * k + 0] += r.sk[k].x * ngx;
            acc[5 * k + 1] += r.yk[k].x * ngx;
            acc[5 * k + 2] += r.s.x * r.yk[k].x;
            acc[5 * k + 3] += r.sk[k].x * r.y.x;
            acc[5 * k + 4] += r.y.x * r.yk[k].x;
        }
        if (NEWDOT) {
            acc[5 * G + 0] += r.y.x * ngx;
            acc[5 * G + 1] += r.y.x * r.y.x;
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
            acc[5 * k + 0] += r.sk[k].y * ngy;
            acc[5 * k + 1] += r.yk[k].y * ngy;
            acc[5 * k + 2] += r.s.y * r.yk[k].y;
            acc[5 * k + 3] += r.sk[k].y * r.y.y;
            acc[5 * k + 4] += r.y.y * r.yk[k].y;
        }
        if (NEWDOT) {
            acc[5 * G + 0] += r.y.y * ngy;
            acc[5 * G + 1] += r.y.y * r.y.y;
        }
    };
    const int64_t npair = cnt >> 1;
    int64_t p = threadIdx.x;
    for (; p + kCsThreads < npair; p += 2 * kCsThreads) {
        Regs ra, rb;
        load(ra, p);
        load(rb, p + kCsThreads);
        apply(ra);
        apply(rb);
    }
    if (p < npair) {
        Regs ra;
        load(ra, p);
        apply(ra);
    }
    if ((cnt & 1) && threadIdx.x == 0) {                    // odd n: the last slice ends with a single element
        const int64_t i = cnt - 1;
        const double ng = -src[i], yni = yn[i], sni = G > 0 ? sn[i] : 0.0;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const double ski = sp[k][i], yki = yp[k][i];
            acc[5 * k + 0] += ski * ng;
            acc[5 * k + 1] += yki * ng;
            acc[5 * k + 2] += sni * yki;
            acc[5 * k + 3] += ski * yni;
            acc[5 * k + 4] += yni * yki;
        }
        if (NEWDOT) {
            acc[5 * G + 0] += yni * ng;
            acc[5 * G + 1] += yni * yni;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                                        // red[] free again
#pragma unroll
    for (int q = 0; q < kAcc; ++q) {
        const double v = warp_sum(acc[q]);
        if (lane == 0) red[q][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < kAcc; ++q) {
            double v = (lane < kCsWarps) ? red[q][lane] : 0.0;
#pragma unroll
            for (int o = kCsWarps / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) part[(q < 5 * G) ? off + q : newdot_off + (q - 5 * G)] = v;
        }
    }
}

// d over this CTA's slice.  BT > 0: bound == BT, all 2 BT + 1 loads of a pair of elements are issued before the first
// use; BT == 0: any bound, four ring vectors at a time.
template <int BT>
__device__ __forceinline__ void cs_direction(const CompactSmallArgs &a, int64_t lo, int64_t cnt, const int *slot, const double *nal,
                                             const double *coef, double gamma, double (&acc)[3]) {
    const int b = BT > 0 ? BT : a.bound;
    const double *__restrict__ src = a.dsrc + lo;
    double *__restrict__ dout = a.d + lo;
    auto finish = [&](int64_t i, double v, double gi) -> double {
        acc[0] += v * v;                                    // dnorm^2 before projection, lbfgs.rs:543
        if (a.owl) {
            const int64_t gidx = a.goff + lo + i;
            if (gidx >= a.start && gidx < a.end && sgn_small(v) != sgn_small(-gi)) v = 0.0;   // orthantwise.rs:140-147
            acc[2] += v * v;                                // ||d|| after projection, :160
        }
        acc[1] += gi * v;                                   // next dginit: g.d | pg.d, core.rs:78-92
        return v;
    };
    auto yptr = [&](int t) { return a.ring + (int64_t)(2 * slot[t] + 1) * a.stride + lo; };
    auto sptr = [&](int t) { return a.ring + (int64_t)(2 * slot[t]) * a.stride + lo; };
    const int64_t npair = cnt >> 1;
    for (int64_t p = threadIdx.x; p < npair; p += kCsThreads) {
        const double2 g2 = __ldg(reinterpret_cast<const double2 *>(src) + p);
        double2 v = make_double2(-g2.x, -g2.y);             // vecncpy, core.rs:99
        if (BT > 0) {
            double2 yv[BT > 0 ? BT : 1], sv[BT > 0 ? BT : 1];
#pragma unroll
            for (int t = 0; t < BT; ++t) {
                yv[t] = __ldg(reinterpret_cast<const double2 *>(yptr(t)) + p);
                sv[t] = __ldg(reinterpret_cast<const double2 *>(sptr(t)) + p);
            }
#pragma unroll
            for (int t = 0; t < BT; ++t) { v.x = v.x + nal[t] * yv[t].x; v.y = v.y + nal[t] * yv[t].y; }   // :589
            v.x = v.x * gamma;                              // :591
            v.y = v.y * gamma;
#pragma unroll
            for (int t = BT - 1; t >= 0; --t) { v.x = v.x + coef[t] * sv[t].x; v.y = v.y + coef[t] * sv[t].y; }   // :599
        } else {
            for (int t0 = 0; t0 < b; t0 += 4) {
                double2 w[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (t0 + c < b) w[c] = __ldg(reinterpret_cast<const double2 *>(yptr(t0 + c)) + p);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (t0 + c < b) { v.x = v.x + nal[t0 + c] * w[c].x; v.y = v.y + nal[t0 + c] * w[c].y; }
            }
            v.x = v.x * gamma;
            v.y = v.y * gamma;
            for (int t0 = 0; t0 < b; t0 += 4) {
                double2 w[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (t0 + c < b) w[c] = __ldg(reinterpret_cast<const double2 *>(sptr(b - 1 - (t0 + c))) + p);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (t0 + c < b) { v.x = v.x + coef[b - 1 - (t0 + c)] * w[c].x; v.y = v.y + coef[b - 1 - (t0 + c)] * w[c].y; }
            }
        }
        double2 o;
        o.x = finish(2 * p, v.x, g2.x);
        o.y = finish(2 * p + 1, v.y, g2.y);
        reinterpret_cast<double2 *>(dout)[p] = o;
    }
    if ((cnt & 1) && threadIdx.x == 0) {
        const int64_t i = cnt - 1;
        const double gi = src[i];
        double v = -gi;
        for (int t = 0; t < b; ++t) v = v + nal[t] * yptr(t)[i];
        v = v * gamma;
        for (int t = b - 1; t >= 0; --t) v = v + coef[t] * sptr(t)[i];
        dout[i] = finish(i, v, gi);
    }
}

__global__ void __launch_bounds__(kCsThreads, 1) k_compact_small(CompactSmallArgs a) {
    __shared__ double sy[kCompactMaxM * kCompactMaxM], yy[kCompactMaxM * kCompactMaxM];
    __shared__ double part[2][kCsMaxSums + 1], tot[kCsMaxSums + 1];
    __shared__ double red[5 * kCompactGroupMax + 2][kCsWarps];
    __shared__ double alpha[kCompactMaxM], coef[kCompactMaxM], nal[kCompactMaxM], sg[kCompactMaxM], yg[kCompactMaxM], ysr[kCompactMaxM];
    __shared__ double gamma_s;
    __shared__ int slot[kCompactMaxM];
    cg::cluster_group cluster = cg::this_cluster();
    const int nctas = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const int m = a.m, b = a.bound, e = a.slot_new, nold = a.bound - 1;
    int64_t per = (a.n + nctas - 1) / nctas;
    per = (per + 1) & ~(int64_t)1;
    const int64_t lo = (int64_t)rank * per;
    const int64_t cnt = (lo >= a.n) ? 0 : ((a.n - lo < per) ? (a.n - lo) : per);
    for (int i = tid; i < m * m; i += kCsThreads) { sy[i] = a.SY[i]; yy[i] = a.YY[i]; }
    if (tid < b) slot[tid] = (e + m - tid) % m;             // newest ... oldest
    __syncthreads();

    // ---- pass A over this CTA's slice: 5 sums per older pair, 2 for the newest ----
    const int nsums = 5 * nold + 2;
    if (nold == 0) {
        cs_gram_group<0, true>(a, lo, cnt, slot, 1, red, part[0], 0, 0);
    } else {
        const int groups = (nold + kCompactGroupMax - 1) / kCompactGroupMax;
        const int pg = (nold + groups - 1) / groups;
        for (int t0 = 1; t0 <= nold; t0 += pg) {
            const int g = (nold - t0 + 1 < pg) ? (nold - t0 + 1) : pg;
            const bool nd = t0 == 1;                        // the first group also sums y_new.d0 and y_new.y_new
            const int off = 5 * (t0 - 1);
#define LB_CS(G) (nd ? cs_gram_group<G, true>(a, lo, cnt, slot, t0, red, part[0], off, 5 * nold) \
                     : cs_gram_group<G, false>(a, lo, cnt, slot, t0, red, part[0], off, 5 * nold))
            switch (g) {
                case 1: LB_CS(1); break;
                case 2: LB_CS(2); break;
                case 3: LB_CS(3); break;
                case 4: LB_CS(4); break;
                default: LB_CS(5); break;
            }
#undef LB_CS
        }
    }
    // ---- totals over the cluster, the same bits in every CTA (fixed rank order) ----
    if (nctas > 1) cluster.sync();
    else __syncthreads();
    for (int q = tid; q < nsums; q += kCsThreads) {
        double v = 0.0;
        for (int r = 0; r < nctas; ++r) {
            const double *remote = (nctas > 1) ? cluster.map_shared_rank(&part[0][0], r) : &part[0][0];
            v += remote[q];
        }
        tot[q] = v;
    }
    __syncthreads();

    // ---- the scalar recursions (compact.cu: k_compact_solve), redundantly in every CTA ----
    for (int t = 1 + tid; t < b; t += kCsThreads) {
        const double *q = tot + 5 * (t - 1);
        const int j = slot[t];
        sg[t] = q[0];
        yg[t] = q[1];
        sy[e * m + j] = q[2];
        sy[j * m + e] = q[3];
        yy[e * m + j] = q[4];
        yy[j * m + e] = q[4];
        ysr[t] = a.ys_dev[j];
        if (rank == 0) {
            a.SY[e * m + j] = q[2];
            a.SY[j * m + e] = q[3];
            a.YY[e * m + j] = q[4];
            a.YY[j * m + e] = q[4];
        }
    }
    if (tid == 0) {
        sg[0] = a.hist[3];
        yg[0] = tot[5 * nold];
        yy[e * m + e] = tot[5 * nold + 1];
        ysr[0] = a.hist[1];
        if (rank == 0) {
            a.YY[e * m + e] = tot[5 * nold + 1];
            a.ys_dev[e] = a.hist[1];
        }
    }
    __syncthreads();
    if (tid == 0) {
        const double gamma = a.hist[1] / a.hist[2];         // lbfgs.rs:691
        for (int t = 0; t < b; ++t) {                       // backward, :582-591
            double acc = sg[t];
            const int j = slot[t];
            for (int i = 0; i < t; ++i) acc += -alpha[i] * sy[j * m + slot[i]];
            alpha[t] = acc / ysr[t];
        }
        for (int t = b - 1; t >= 0; --t) {                  // forward, :594-601
            double acc = yg[t];
            const int j = slot[t];
            for (int i = 0; i < b; ++i) acc += -alpha[i] * yy[j * m + slot[i]];
            acc = acc * gamma;
            for (int i = b - 1; i > t; --i) acc += coef[i] * sy[slot[i] * m + j];
            coef[t] = alpha[t] - acc / ysr[t];
        }
        for (int t = 0; t < b; ++t) nal[t] = -alpha[t];
        gamma_s = gamma;
    }
    __syncthreads();

    // ---- pass B over the slice: d, element-wise in the reference's order ----
    const double gamma = gamma_s;
    double acc3[3] = {0.0, 0.0, 0.0};
    switch (b) {
        case 1: cs_direction<1>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
        case 2: cs_direction<2>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
        case 3: cs_direction<3>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
        case 4: cs_direction<4>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
        case 5: cs_direction<5>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
        case 6: cs_direction<6>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
        case 7: cs_direction<7>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
        case 8: cs_direction<8>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
        default: cs_direction<0>(a, lo, cnt, slot, nal, coef, gamma, acc3); break;
    }
    {
        const int lane = tid & 31, warp = tid >> 5;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double v = warp_sum(acc3[q]);
            if (lane == 0) red[q][warp] = v;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                double v = (lane < kCsWarps) ? red[q][lane] : 0.0;
#pragma unroll
                for (int o = kCsWarps / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0) part[1][q] = v;
            }
        }
        if (nctas > 1) cluster.sync();
        else __syncthreads();
        if (rank == 0 && tid == 0) {
            double o3[3] = {0.0, 0.0, 0.0};
            for (int r = 0; r < nctas; ++r) {
                const double *remote = (nctas > 1) ? cluster.map_shared_rank(&part[1][0], r) : &part[1][0];
                for (int q = 0; q < 3; ++q) o3[q] += remote[q];
            }
            a.out[0] = o3[0];
            a.out[1] = o3[1];
            a.out[2] = o3[2];
            const double dnorm = sqrt(o3[0]);                                        // lbfgs.rs:543
            *a.step_out = a.constrain ? fmin(a.max_step, dnorm) / dnorm : 1.0;       // :547-551
        }
    }
    if (nctas > 1) cluster.sync();   // no CTA may exit while a peer still reads its partials through DSMEM
}

}  // namespace

// Largest n the cluster kernel takes: beyond ~2^18 elements a CTA's slice streams from L2 for longer per trip than
// a full-grid kernel of the multi-kernel chain takes to launch and run.
constexpr int64_t kSmallSliceMax = 24576;   // doubles of q per CTA (192 KB of shared memory)
int64_t two_loop_small_max_n() { return (int64_t)1 << 18; }

// Returns cudaSuccess, or the launch error (the caller falls back to the multi-kernel chain).
cudaError_t launch_two_loop_small(const Launch &L, int device, int64_t n, int m, int bound, int slot_new, double *d,
                                  const double *dsrc, double *ring, int64_t stride, double *ys_dev, const double *hist,
                                  double *out, bool owl, int64_t start, int64_t end, int64_t goff, double max_step,
                                  bool constrain, double *step_out) {
    if (m > kSmallMaxM || bound < 1 || n > two_loop_small_max_n()) return cudaErrorInvalidValue;
    // cluster size: ~2048 elements per CTA at least (4 per thread), as few CTAs as that allows, a power of two
    int c = 1;
    while (c < kSmallMaxCluster && (n > (int64_t)c * 2048 || (n + c - 1) / c + 2 > kSmallSliceMax)) c *= 2;
    if ((n + c - 1) / c + 2 > kSmallSliceMax) return cudaErrorInvalidValue;
    static bool attr_set[64] = {};   // function attributes are per device
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (!attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(k_two_loop_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmallSliceMax * 8));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_two_loop_small, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
        attr_set[device] = true;
    }
    int64_t per = (n + c - 1) / c;
    per = (per + 1) & ~(int64_t)1;
    SmallArgs a{n, m, bound, slot_new, d, dsrc, ring, stride, ys_dev, hist, out, owl ? 1 : 0, start, end, goff,
                max_step, constrain ? 1 : 0, step_out};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)c);
    cfg.blockDim = dim3(kSmallThreads);
    cfg.dynamicSmemBytes = (size_t)per * sizeof(double);
    cfg.stream = L.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (L.launch_counter) ++*L.launch_counter;
    return cudaLaunchKernelEx(&cfg, k_two_loop_small, a);
}

// The compact direction for n <= two_loop_small_max_n(): pass A + scalar recursions + pass B in one cluster launch.
cudaError_t launch_compact_small(const Launch &L, int device, int64_t n, int m, int bound, int slot_new, double *d,
                                 const double *dsrc, const double *ring, int64_t stride, double *ys_dev, double *SY, double *YY,
                                 const double *hist, double *out, bool owl, int64_t start, int64_t end, int64_t goff,
                                 double max_step, bool constrain, double *step_out) {
    if (m > kCompactMaxM || bound < 1 || n > two_loop_small_max_n()) return cudaErrorInvalidValue;
    int c = 1;
    while (c < kSmallMaxCluster && n > (int64_t)c * 1024) c *= 2;   // >= ~2 pairs per thread per CTA before another CTA joins
    static bool attr_set[64] = {};   // function attributes are per device
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (!attr_set[device]) {
        const cudaError_t e = cudaFuncSetAttribute(k_compact_small, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
        attr_set[device] = true;
    }
    CompactSmallArgs a{n, m, bound, slot_new, d, dsrc, ring, stride, ys_dev, SY, YY, hist, out, owl ? 1 : 0, start, end, goff,
                       max_step, constrain ? 1 : 0, step_out};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)c);
    cfg.blockDim = dim3(kCsThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = L.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (L.launch_counter) ++*L.launch_counter;
    return cudaLaunchKernelEx(&cfg, k_compact_small, a);
}

}  // namespace lb
