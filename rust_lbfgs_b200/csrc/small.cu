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

// Registers holding one trip's first kChunk elements per thread of its two streamed vectors.  They are loaded for
// trip k + 1 BEFORE trip k's cluster reduction, so the L2 latency of the next trip's operands hides behind the
// barrier of this one: a trip then costs about one barrier, not barrier + load latency.
constexpr int kChunk = 16;
struct Prefetch {
    double a[kChunk], b[kChunk];
};
__device__ __forceinline__ void prefetch(Prefetch &p, const double *__restrict__ va, const double *__restrict__ vb,
                                         int64_t cnt, int tid) {
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
        const int64_t i = tid + (int64_t)k * kSmallThreads;
        p.a[k] = (i < cnt) ? __ldg(va + i) : 0.0;             // read-only for the whole kernel: ld.global.nc
        p.b[k] = (vb != nullptr && i < cnt) ? __ldg(vb + i) : 0.0;
    }
}

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
    const int m = a.m, bound = a.bound, trips = 2 * a.bound;
    if (tid < m) ys_s[tid] = a.ys_dev[tid];
    __syncthreads();
    const double ys_new = a.hist[1];
    const double gamma = ys_new / a.hist[2];                // ys / yy of the newest pair (src/lbfgs.rs:691)
    int parity = 0;
    double red = a.hist[3];                                 // s_new . (-g): the first alpha's numerator
    const double *__restrict__ src = a.dsrc + lo;           // g | pg: the recursion starts from -src, the last trip dots with it

    // The operands of trip k (src/lbfgs.rs:582-601): backward trips walk j = slot_new, slot_new - 1, ... and stream
    // (y_j, s_{j-1}); forward trips walk back up and stream (s_j, y_{j+1}); the last forward trip dots with g | pg.
    auto slot_of = [&](int k) {   // ring slot of trip k
        const int back = (k < bound) ? k : (trips - 1 - k);
        return ((a.slot_new - back) % m + m) % m;
    };
    auto operands = [&](int k, const double *&va, const double *&vb) {
        const int j = slot_of(k);
        if (k < bound) {
            va = a.ring + (int64_t)(2 * j + 1) * a.stride + lo;                                  // y_j
            vb = (k == bound - 1) ? nullptr : a.ring + (int64_t)(2 * ((j + m - 1) % m)) * a.stride + lo;   // s_{j-1}
        } else {
            va = a.ring + (int64_t)(2 * j) * a.stride + lo;                                      // s_j
            vb = (k == trips - 1) ? src : a.ring + (int64_t)(2 * ((j + 1) % m) + 1) * a.stride + lo;        // y_{j+1} | g
        }
    };
    Prefetch pf;
    {
        const double *va, *vb;
        operands(0, va, vb);
        prefetch(pf, va, vb, cnt, tid);
    }
    for (int k = 0; k < trips; ++k) {
        const bool backward = k < bound;
        const bool first = (k == 0), last_b = (k == bound - 1), last_f = (k == trips - 1);
        const int j = slot_of(k);
        const double *va, *vb;
        operands(k, va, vb);
        double coef;
        if (backward) {
            const double ys_j = first ? ys_new : ys_s[j];
            const double alpha = red / ys_j;                // :587
            if (tid == 0) {
                alpha_s[j] = alpha;
                if (first) ys_s[j] = ys_j;
                if (first && rank == 0) a.ys_dev[j] = ys_j;  // keep y.s of the newest pair for the next m iterations (:653)
            }
            coef = -alpha;
        } else {
            const double beta = red / ys_s[j];              // :597
            coef = alpha_s[j] - beta;                       // :599
        }
        double acc[3] = {0.0, 0.0, 0.0};
        for (int64_t i0 = tid; i0 < cnt; i0 += (int64_t)kChunk * kSmallThreads) {
            if (i0 != tid) prefetch(pf, va + (i0 - tid), vb ? vb + (i0 - tid) : nullptr, cnt - (i0 - tid), tid);   // slices beyond 8 192 elements
#pragma unroll
            for (int c = 0; c < kChunk; ++c) {
                const int64_t i = i0 + (int64_t)c * kSmallThreads;
                if (i < cnt) {
                    if (backward) {
                        const double qi = first ? -__ldg(src + i) : q[i];   // vecncpy, core.rs:99
                        double v = qi + coef * pf.a[c];         // vecadd(y, -alpha), :589
                        if (last_b) {
                            v = v * gamma;                      // vecscale(gamma), :591
                            acc[0] += pf.a[c] * v;              // y_j . d for the first beta, :597
                        } else {
                            acc[0] += pf.b[c] * v;              // s_{j-1} . q for the next alpha, :587
                        }
                        q[i] = v;
                    } else {
                        double v = q[i] + coef * pf.a[c];       // vecadd(s, alpha - beta), :599
                        if (!last_f) {
                            acc[0] += pf.b[c] * v;              // y_{j+1} . r for the next beta, :597
                            q[i] = v;
                        } else {
                            acc[0] += v * v;                    // dnorm^2 before projection, :543
                            if (a.owl) {
                                const int64_t gidx = a.goff + lo + i;
                                if (gidx >= a.start && gidx < a.end && sgn_small(v) != sgn_small(-pf.b[c])) v = 0.0;  // orthantwise.rs:140-147
                                acc[2] += v * v;                // ||d|| after projection, :160
                            }
                            acc[1] += pf.b[c] * v;              // next dginit: g.d or pg.d, core.rs:78-92
                            a.d[lo + i] = v;
                        }
                    }
                }
            }
        }
        if (k + 1 < trips) {   // the next trip's operands do not depend on this trip's dot product: fetch them now
            const double *na, *nb;
            operands(k + 1, na, nb);
            prefetch(pf, na, nb, cnt, tid);
        }
        if (last_f) cluster_sum<3>(acc, warp_part, cta_part, parity, cluster, nctas);
        else cluster_sum<1>(reinterpret_cast<double(&)[1]>(acc), warp_part, cta_part, parity, cluster, nctas);
        parity ^= 1;
        red = acc[0];
        if (last_f && rank == 0 && tid == 0) {
            a.out[0] = acc[0];
            a.out[1] = acc[1];
            a.out[2] = acc[2];
            const double dnorm = sqrt(acc[0]);                                // lbfgs.rs:543
            *a.step_out = a.constrain ? fmin(a.max_step, dnorm) / dnorm : 1.0;   // :547-551
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

}  // namespace lb
