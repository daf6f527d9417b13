// types.h — constants and PODs shared by host code and kernels.
#pragma once

#include <stdint.h>

namespace lb {

constexpr int kThreads = 256;          // threads per CTA for every streaming kernel
constexpr int kWarps = kThreads / 32;
constexpr int kMinBlocks = 4;         // resident CTAs per SM every kernel is compiled for (<= 64 registers)
constexpr int kMaxAcc = 8;             // accumulators per kernel (and doubles per scalar slot)

// Per-solver reduction workspace in HBM.
struct ReduceWs {
    double *partials;      // [kMaxAcc][stride]
    unsigned int *ticket;  // zero between kernels
    int stride;            // >= max grid size
};

}  // namespace lb
