// types.h — constants and PODs shared by host code and kernels.
#pragma once

#include <stdint.h>

namespace lb {

#ifndef LB_THREADS
#define LB_THREADS 256
#endif
constexpr int kThreads = LB_THREADS;   // threads per CTA for every streaming kernel
constexpr int kWarps = kThreads / 32;
#ifndef LB_MINBLOCKS
#define LB_MINBLOCKS 2
#endif
#ifndef LB_U
#define LB_U 8
#endif
#ifndef LB_UH
#define LB_UH 4
#endif
#ifndef LB_UT
#define LB_UT 6
#endif
constexpr int kUt = LB_UT;             // the same for the fused trial kernel (2R 2W) and, so that the fused and the
                                       // unfused paths sum the same terms in the same order, K2 and the objective
#ifndef LB_PROBE_PREFETCH
#define LB_PROBE_PREFETCH 0
#endif
constexpr bool kProbePrefetch = LB_PROBE_PREFETCH != 0;   // software-pipelined tile walk for the write-free probe
constexpr int kUh = LB_UH;             // the same for the history kernel (5 input vectors)
// Tile shapes, tuned on B200 at n = 1e8 (profiles/r01_tuning.md): one 256-thread CTA per SM with 8 independent
// 128-bit loads per input vector in flight per thread (2048 per SM per vector) streams 3R 1W at ~7.0 TB/s; more
// resident CTAs or deeper unrolling lose 5-7 %.  The 2R 2W fused trial is best with 24 KB (non power of two)
// tiles, the 4R 2W history kernel with 16 KB tiles.
constexpr int kMinBlocks = LB_MINBLOCKS;  // resident CTAs per SM every kernel is compiled for (2 => <= 128 registers)
constexpr int kU = LB_U;               // independent 128-bit loads per input vector per thread (tile = kU * 4 KB)
constexpr int kMaxAcc = 8;             // accumulators per kernel (and doubles per scalar slot)

// Cross-GPU exchange fused into the reducing kernels (one process per GPU, peers mapped with CUDA IPC over
// NVLink / NVSwitch).  Every rank owns a mailbox ring in its HBM: entry (seq % kMailRing, sender) holds a sequence
// word and up to kMailVals doubles.  The last CTA of a reducing kernel stores its totals into EVERY rank's
// mailbox (peer stores), waits until all senders' entries for this seq have landed in its own mailbox, and sums
// them in rank order — deterministic and bit-identical on every rank, which the replicated scalar control
// logic needs.  No NCCL call, no extra launch.
constexpr int kMaxPeers = 8;
constexpr int kMailRing = 4;
constexpr int kMailStride = 32;        // doubles per entry (256 B): [0] = seq word, [1..] = values
constexpr int kMailVals = 26;          // 24 = six line-search trial points x {f, g.d, g.g, x.x} in one exchange, + 2 riders
struct PeerCtx {
    int nranks;                    // 0 / 1 = no exchange
    int rank;
    unsigned long long seq;        // this launch's sequence number (same on every rank)
    double *mail_self;             // this rank's mailbox
    double *const *mail_table;     // device array: mail_table[r] = rank r's mailbox as mapped into this process
    double *extra[2];              // optional device scalars that ride along with the exchange (summed in place)
    unsigned int *fault;           // communicator-owned device word, set when a peer did not answer in time
};

// The opt-in compact search direction (compact.cu): history depth it supports and older ring slots per pass-A launch
// (5 running sums each, kept in registers).
constexpr int kCompactMaxM = 32;
constexpr int kCompactGroupMax = 5;
constexpr int kCompactSplitDefault = 0;   // ring vectors per sub-pass of pass B beyond bound 8 (0 = one pass); tuned on B200

// Per-solver reduction workspace in HBM.
struct ReduceWs {
    double *partials;      // [kMaxAcc][stride]
    unsigned int *ticket;  // zero between kernels
    int stride;            // >= max grid size
    PeerCtx peer;          // multi-GPU: fused all-reduce of the results (nranks <= 1: none)
};

}  // namespace lb
