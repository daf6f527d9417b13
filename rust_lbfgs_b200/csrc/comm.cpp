// comm.cpp — one NCCL rank per process/GPU, used for exactly one thing: summing a handful of f64
// partial dot products over the ranks after each fused reduction step.
//
// libnccl.so.2 is resolved with dlopen at first use, so the single-GPU path has no NCCL
// dependency at all and a process that already loaded torch's bundled NCCL shares that copy.
// ncclAllReduce delivers bit-identical results on every rank, which the replicated scalar
// line-search logic relies on.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/lbfgsb200.h"
#include "solver.h"

namespace lb {
namespace {

// Minimal mirror of the NCCL ABI we use (nccl.h 2.x): stable since 2.0.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };  // ncclDataType_t: ncclDouble
enum { ncclSum = 0 };      // ncclRedOp_t

struct Nccl {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Nccl &nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle) return;
        n.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(n.handle, "ncclGetUniqueId");
        n.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(n.handle, "ncclCommInitRank");
        n.CommDestroy = (int (*)(ncclComm_t))dlsym(n.handle, "ncclCommDestroy");
        n.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(n.handle, "ncclAllReduce");
        n.GetErrorString = (const char *(*)(int))dlsym(n.handle, "ncclGetErrorString");
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce;
    });
    return n;
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
};

int comm_rank(const Comm *c) { return c ? c->rank : 0; }
int comm_size(const Comm *c) { return c ? c->nranks : 1; }

int comm_allreduce_sum(Comm *c, double *buf_dev, int count, cudaStream_t stream) {
    if (!c || c->nranks == 1) return 0;
    Nccl &n = nccl();
    if (!n.ok) return LBFGSB200_ERR_NCCL;
    int rc = n.AllReduce(buf_dev, buf_dev, (size_t)count, ncclFloat64, ncclSum, c->comm, stream);
    return rc == ncclSuccess ? 0 : LBFGSB200_ERR_NCCL;
}

}  // namespace lb

extern "C" {

int lbfgsb200_comm_unique_id(char id[LBFGSB200_UNIQUE_ID_BYTES]) {
    lb::Nccl &n = lb::nccl();
    if (!n.ok) return LBFGSB200_ERR_NCCL;
    lb::ncclUniqueId uid;
    if (n.GetUniqueId(&uid) != lb::ncclSuccess) return LBFGSB200_ERR_NCCL;
    static_assert(sizeof(uid) == LBFGSB200_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id, &uid, sizeof(uid));
    return 0;
}

int lbfgsb200_comm_create(const char id[LBFGSB200_UNIQUE_ID_BYTES], int rank, int nranks, int device,
                          lbfgsb200_comm_t **out) {
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return LBFGSB200_ERR_INVALID_PARAM;
    *out = nullptr;
    lb::Nccl &n = lb::nccl();
    if (!n.ok) return LBFGSB200_ERR_NCCL;
    if (cudaSetDevice(device) != cudaSuccess) return LBFGSB200_ERR_CUDA;
    lb::ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    lb::Comm *c = new lb::Comm();
    c->rank = rank;
    c->nranks = nranks;
    c->device = device;
    if (n.CommInitRank(&c->comm, nranks, uid, rank) != lb::ncclSuccess) {
        delete c;
        return LBFGSB200_ERR_NCCL;
    }
    *out = reinterpret_cast<lbfgsb200_comm_t *>(c);
    return 0;
}

void lbfgsb200_comm_destroy(lbfgsb200_comm_t *comm) {
    lb::Comm *c = reinterpret_cast<lb::Comm *>(comm);
    if (!c) return;
    lb::Nccl &n = lb::nccl();
    if (n.ok && c->comm) n.CommDestroy(c->comm);
    delete c;
}

int lbfgsb200_comm_allreduce_sum(lbfgsb200_comm_t *comm, double *buf_dev, int count, void *stream) {
    return lb::comm_allreduce_sum(reinterpret_cast<lb::Comm *>(comm), buf_dev, count, (cudaStream_t)stream);
}

}  // extern "C"
