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
