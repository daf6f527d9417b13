// reduce.cuh — streaming map + deterministic two-level reduction skeleton (sm_100a, f64).
//
// Every hot-path kernel is "stream a few n-vectors once, write at most a couple, and emit up to
// kMaxAcc dot-product-like sums".  Level 1: each thread keeps its accumulators over a fixed set of
// elements, a fixed-shape warp-shuffle tree and a fixed-shape cross-warp tree give one partial per
// CTA.  Level 2: the LAST CTA to finish (atomic ticket) sums the per-CTA partials in index order
// with the same fixed trees and writes the results to device memory.  Which CTA runs level 2 does
// not matter: the summation order depends only on (n, grid), so results are bit-reproducible run
// to run.  This replaces the sequential fold of `vecdot` (src/math.rs:40-42) — and fuses what the
// survey calls K9 (`finalize_reduce`) into the producing kernel, so no extra launch is needed.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "types.h"

namespace lb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// Sum acc[] over the CTA; result valid in thread 0.
template <int NACC>
__device__ __forceinline__ void block_sum(double (&acc)[NACC], double (*sm)[kWarps]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double v = warp_sum(acc[a]);
        if (lane == 0) sm[a][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            double v = (lane < kWarps) ? sm[a][lane] : 0.0;
#pragma unroll
            for (int off = kWarps / 2; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            acc[a] = v;
        }
    }
}

// Sum nv values over all ranks through the peer mailboxes; every thread of warp 0 must call it.  See PeerCtx in
// types.h.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifndef LB_PEER_TIMEOUT_MS
#define LB_PEER_TIMEOUT_MS 20000
#endif
constexpr unsigned long long kPeerTimeoutNs = (unsigned long long)LB_PEER_TIMEOUT_MS * 1000000ull;
// In/out: tab[kMaxPeers][0..nv) in shared memory (written by lane 0 before the call, read by it afterwards).
__device__ __forceinline__ void peer_allreduce(const PeerCtx &pc, int nv, double (*tab)[kMailVals]) {
    const int lane = threadIdx.x & 31;
    const unsigned long long seq = pc.seq;
    const size_t entry = (size_t)(seq & (kMailRing - 1)) * pc.nranks;
    double *mail_lane = (lane < pc.nranks) ? pc.mail_table[lane] : nullptr;   // the peers' mailboxes, from device memory
    double *mail_self = pc.mail_self;
    __syncwarp();
    if (lane < pc.nranks) {  // lane t publishes this rank's totals into rank t's mailbox (a peer store over NVLink)
        volatile double *dst = mail_lane + (entry + pc.rank) * kMailStride;
        for (int a = 0; a < nv; ++a) dst[1 + a] = tab[kMaxPeers][a];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(dst) = seq;
    }
    if (lane < pc.nranks) {  // lane t collects sender t's entry from the local mailbox
        volatile double *src = mail_self + (entry + lane) * kMailStride;
        const unsigned long long t0 = globaltimer_ns();
        bool ok = true;
        while (*reinterpret_cast<volatile unsigned long long *>(src) != seq) {
            if (globaltimer_ns() - t0 > kPeerTimeoutNs) { ok = false; break; }  // a peer died (or never launched)
        }
        if (!ok && pc.fault) atomicExch(pc.fault, 1u);   // the host turns this into LBFGSB200_ERR_NCCL (Solver::fetch)
        __threadfence_system();
        for (int a = 0; a < nv; ++a) tab[lane][a] = ok ? src[1 + a] : __longlong_as_double(0x7ff8000000000000ll);
    }
    __syncwarp();
    double v = 0.0;
    if (lane < nv)
        for (int r = 0; r < pc.nranks; ++r) v += tab[r][lane];  // fixed rank order
    __syncwarp();
    if (lane < nv) tab[kMaxPeers][lane] = v;
    __syncwarp();
}

// Level 1 + level 2.  `out[a]` receives the grid-wide sum of accumulator a.
template <int NACC>
__device__ __forceinline__ void grid_reduce(double (&acc)[NACC], const ReduceWs &ws, double *__restrict__ out) {
    __shared__ double sm[NACC][kWarps];
    __shared__ bool is_last;
    if (blockDim.x == 1) {  // reference-order mode (<<<1, 1>>>): acc[] already is the sequential fold
#pragma unroll
        for (int a = 0; a < NACC; ++a) out[a] = acc[a];
        return;
    }
    block_sum<NACC>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) __stcg(&ws.partials[(size_t)a * ws.stride + blockIdx.x], acc[a]);
        __threadfence();
        unsigned int t = atomicAdd(ws.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double v = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads)
            v += __ldcg(&ws.partials[(size_t)a * ws.stride + i]);
        acc[a] = v;
    }
    __syncthreads();  // sm reuse
    block_sum<NACC>(acc, sm);
    if (ws.peer.nranks > 1) {  // level 3: the same totals from every GPU, summed in rank order (warp 0)
        __shared__ double tab[kMaxPeers + 1][kMailVals];
        if (threadIdx.x < 32) {
            int nv = NACC;
            if (threadIdx.x == 0) {
#pragma unroll
                for (int a = 0; a < NACC; ++a) tab[kMaxPeers][a] = acc[a];
                if (ws.peer.extra[0]) tab[kMaxPeers][nv++] = *ws.peer.extra[0];
                if (ws.peer.extra[1]) tab[kMaxPeers][nv++] = *ws.peer.extra[1];
            }
            nv = __shfl_sync(0xffffffffu, nv, 0);
            peer_allreduce(ws.peer, nv, tab);
            if (threadIdx.x == 0) {
#pragma unroll
                for (int a = 0; a < NACC; ++a) acc[a] = tab[kMaxPeers][a];
                int k = NACC;
                if (ws.peer.extra[0]) *ws.peer.extra[0] = tab[kMaxPeers][k++];
                if (ws.peer.extra[1]) *ws.peer.extra[1] = tab[kMaxPeers][k++];
            }
        }
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) out[a] = acc[a];
        *ws.ticket = 0u;
    }
}

// The same two levels for a kernel whose sums belong to three different consumers: accumulators [0, n0) go to out0,
// [n0, n0 + n1) to out1, the rest to out2.  No cross-GPU level (single-GPU fusions only); ws.partials must hold
// NACC * ws.stride doubles.
template <int NACC>
__device__ __forceinline__ void grid_reduce_split(double (&acc)[NACC], const ReduceWs &ws, double *__restrict__ out0, int n0,
                                                  double *__restrict__ out1, int n1, double *__restrict__ out2) {
    __shared__ double sm[NACC][kWarps];
    __shared__ bool is_last;
    auto put = [&](int a, double v) {
        if (a < n0) out0[a] = v;
        else if (a < n0 + n1) out1[a - n0] = v;
        else out2[a - n0 - n1] = v;
    };
    if (blockDim.x == 1) {  // reference-order mode
#pragma unroll
        for (int a = 0; a < NACC; ++a) put(a, acc[a]);
        return;
    }
    block_sum<NACC>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) __stcg(&ws.partials[(size_t)a * ws.stride + blockIdx.x], acc[a]);
        __threadfence();
        unsigned int t = atomicAdd(ws.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double v = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads)
            v += __ldcg(&ws.partials[(size_t)a * ws.stride + i]);
        acc[a] = v;
    }
    __syncthreads();  // sm reuse
    block_sum<NACC>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) put(a, acc[a]);
        *ws.ticket = 0u;
    }
}

// 128-bit streaming loads/stores.  kStream marks the instantiations used when the vectors are far larger than L2
// (n = 1e8: 0.8 GB per vector vs 126 MB of L2); which cache hints they use is a build-time policy (below).
// Cache hints for the streaming (working set >> L2) instantiations, measured on B200 at n = 1e8
// (profiles/r01_tuning.md): evict-first loads (ld.global.cs) cost 4 % on most boxes (6.70 vs 7.00 TB/s for the
// 3R 1W kernels), no-allocate loads are no better, and evict-first / .cg stores only help together with plain
// loads.  Plain loads and stores are the robust choice, so the hints are off by default.
#ifndef LB_LD_POLICY
#define LB_LD_POLICY 1   // 0 = ld.global.cs (evict-first), 1 = default, 2 = ld.global.nc.L1::no_allocate
#endif
#ifndef LB_ST_POLICY
#define LB_ST_POLICY 1   // 0 = st.global.cs (evict-first), 1 = default, 2 = st.global.cg
#endif
template <bool kStream>
__device__ __forceinline__ double2 ld2(const double *__restrict__ p, int64_t i) {
    const double2 *q = reinterpret_cast<const double2 *>(p) + i;
    if (kStream) {
#if LB_LD_POLICY == 0
        return __ldcs(q);
#elif LB_LD_POLICY == 2
        double2 v;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(q));
        return v;
#else
        return *q;
#endif
    }
    return *q;
}
template <bool kStream>
__device__ __forceinline__ void st2(double *__restrict__ p, int64_t i, double2 v) {
    double2 *q = reinterpret_cast<double2 *>(p) + i;
    if (kStream) {
#if LB_ST_POLICY == 0
        __stcs(q, v);
#elif LB_ST_POLICY == 2
        __stcg(q, v);
#else
        *q = v;
#endif
    } else {
        *q = v;
    }
}

// Streams n elements through `op` as double2 pairs: U independent 128-bit loads per input vector
// are issued per thread before any use (memory-level parallelism), tiles are contiguous per CTA
// (coalesced 4 KB * U per vector), grid-stride over tiles.  An odd trailing element is handled by
// one thread through op.tail().
//   Op::Regs                      registers holding one pair per input vector
//   op.load(Regs&, int64 pair)    issue the loads
//   op.apply(Regs&, int64 pair, double (&acc)[NACC])   compute, store, accumulate
//   op.tail(int64 elem, acc)      scalar path for element n-1 when n is odd
//
// Reference-order mode: launched as <<<1, 1>>> the single thread walks the elements in index order, so
// every accumulator is exactly the sequential left-to-right fold of `vecdot` (src/math.rs:40-42) and a
// whole solve is bit-identical to the reference's CPU arithmetic (validation only: one thread).
template <int NACC, int U, class Op>
__device__ __forceinline__ void stream_pairs(int64_t n, Op &op, double (&acc)[NACC]) {
    const int64_t nv = n >> 1;
    if (blockDim.x == 1) {
        for (int64_t i = 0; i < nv; ++i) {
            typename Op::Regs r;
            op.load(r, i);
            op.apply(r, i, acc);
        }
        if (n & 1) op.tail(n - 1, acc);
        return;
    }
    constexpr int64_t kTile = (int64_t)kThreads * U;
    for (int64_t base = (int64_t)blockIdx.x * kTile; base < nv; base += (int64_t)gridDim.x * kTile) {
        typename Op::Regs r[U];
        if (base + kTile <= nv) {
#pragma unroll
            for (int u = 0; u < U; ++u) op.load(r[u], base + u * kThreads + threadIdx.x);
#pragma unroll
            for (int u = 0; u < U; ++u) op.apply(r[u], base + u * kThreads + threadIdx.x, acc);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + u * kThreads + threadIdx.x;
                if (i < nv) op.load(r[u], i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + u * kThreads + threadIdx.x;
                if (i < nv) op.apply(r[u], i, acc);
            }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) op.tail(n - 1, acc);
}

// The same walk with the NEXT tile's loads issued before the current tile is consumed (two register sets, the loop
// unrolled by two so no register moves): for kernels with enough arithmetic per byte that "load a tile, then
// compute on it" leaves the memory pipe idle during the compute phase (the Rosenbrock probe: ~25 flops per 32
// bytes, all 8 warps of the one resident CTA in the same phase).  Same element -> thread map and the same
// accumulation order as stream_pairs, hence the same bits.
template <int NACC, int U, class Op>
__device__ __forceinline__ void stream_pairs_prefetch(int64_t n, Op &op, double (&acc)[NACC]) {
    const int64_t nv = n >> 1;
    if (blockDim.x == 1) {
        stream_pairs<NACC, U>(n, op, acc);
        return;
    }
    constexpr int64_t kTile = (int64_t)kThreads * U;
    const int64_t stride = (int64_t)gridDim.x * kTile;
    int64_t base = (int64_t)blockIdx.x * kTile;
    typename Op::Regs ra[U], rb[U];
    bool full = base + kTile <= nv;
    if (full) {
#pragma unroll
        for (int u = 0; u < U; ++u) op.load(ra[u], base + u * kThreads + threadIdx.x);
    }
    while (full) {
        int64_t next = base + stride;
        bool next_full = next + kTile <= nv;
        if (next_full) {
#pragma unroll
            for (int u = 0; u < U; ++u) op.load(rb[u], next + u * kThreads + threadIdx.x);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) op.apply(ra[u], base + u * kThreads + threadIdx.x, acc);
        base = next;
        full = next_full;
        if (!full) break;
        next = base + stride;
        next_full = next + kTile <= nv;
        if (next_full) {
#pragma unroll
            for (int u = 0; u < U; ++u) op.load(ra[u], next + u * kThreads + threadIdx.x);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) op.apply(rb[u], base + u * kThreads + threadIdx.x, acc);
        base = next;
        full = next_full;
    }
    if (base < nv) {   // this CTA's last, partial tile
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kThreads + threadIdx.x;
            if (i < nv) op.load(ra[u], i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kThreads + threadIdx.x;
            if (i < nv) op.apply(ra[u], i, acc);
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) op.tail(n - 1, acc);
}

// IterationData::update for one element (src/lbfgs.rs:644-656, :670-673).  Shared by k_history and by the objectives'
// commit kernels (lbfgsb200_commit_fn), so the fused and the unfused paths accumulate the same terms.
template <bool DAMP, bool OWL>
__device__ __forceinline__ void history_elem(double xi, double xpi, double gi, double gpi, double pgi, double nstep,
                                             double &si, double &yi, double (&acc)[5]) {
    si = xi - xpi;                                      // lbfgs.rs:644
    yi = gi - gpi;                                      // :647 (raw gradients, also for OWL-QN: :529-530)
    acc[0] += si * si;                                  // :645
    acc[1] += yi * si;                                  // :653
    acc[2] += yi * yi;                                  // :654
    acc[3] += si * (-(OWL ? pgi : gi));                 // first trip of :587 with d = -g | -pg (core.rs:95-101)
    if (DAMP) acc[4] += si * (gpi * nstep);             // :670-673, nstep = -step
}

}  // namespace lb
