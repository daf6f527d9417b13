// comm.cpp — one rank per process/GPU, used for exactly one thing: summing a handful of f64 partial dot
// products over the ranks after each fused reduction step.
//
// Two transports.  (1) The default on an NVLink/NVSwitch box: PEER MAILBOXES — every rank allocates a small
// mailbox ring in its HBM, the ranks exchange CUDA IPC handles once (over NCCL, at communicator creation) and
// map each other's mailboxes; from then on the all-reduce happens INSIDE the reducing kernels (reduce.cuh:
// peer_allreduce — peer stores over NVLink, fixed rank-order sum), with no collective call and no extra launch.
// (2) ncclAllReduce(ncclDouble, ncclSum) per step: the fallback when IPC mapping is unavailable, or with
// LBFGSB200_PEER_REDUCE=0.  Both give bit-identical results on every rank, which the replicated scalar
// line-search logic relies on.
//
// libnccl.so.2 is resolved with dlopen at first use, so the single-GPU path has no NCCL
// dependency at all and a process that already loaded torch's bundled NCCL shares that copy.
// ncclAllReduce delivers bit-identical results on every rank, which the replicated scalar
// line-search logic relies on.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

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
    int (*CommAbort)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
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
        n.CommAbort = (int (*)(ncclComm_t))dlsym(n.handle, "ncclCommAbort");
        n.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(n.handle, "ncclAllReduce");
        n.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(n.handle, "ncclAllGather");
        n.Broadcast = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(n.handle, "ncclBroadcast");
        n.GroupStart = (int (*)())dlsym(n.handle, "ncclGroupStart");
        n.GroupEnd = (int (*)())dlsym(n.handle, "ncclGroupEnd");
        n.GetErrorString = (const char *(*)(int))dlsym(n.handle, "ncclGetErrorString");
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce;
    });
    return n;
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    // peer mailboxes (transport 1)
    PeerCtx peer{};                    // nranks == 0: not available, use NCCL
    unsigned long long seq = 0;        // exchange counter, advanced identically on every rank
    double *mailbox = nullptr;         // this rank's ring: kMailRing x nranks entries of kMailStride doubles, then the fault word
    bool aborted = false;              // ncclCommAbort was called (setup could not even stage its handshake)
    double **mail_table = nullptr;     // device array of every rank's mailbox pointer (as mapped here)
    void *opened[kMaxPeers] = {};
};

int comm_rank(const Comm *c) { return c ? c->rank : 0; }
int comm_size(const Comm *c) { return c ? c->nranks : 1; }
const PeerCtx *comm_peer(const Comm *c) { return (c && c->peer.nranks > 1) ? &c->peer : nullptr; }
unsigned long long *comm_peer_seq(Comm *c) { return c ? &c->seq : nullptr; }
// 1 if a reducing kernel gave up waiting for a peer's mailbox entry (the sums it returned are NaN)
int comm_peer_fault(const Comm *c) {
    if (!c || !c->peer.fault) return 0;
    unsigned int v = 0;
    if (cudaMemcpy(&v, c->peer.fault, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return v != 0;
}

namespace {
// Maps every rank's mailbox into this process.  Collective; returns false (on every rank) if any rank failed.
bool setup_peer_mailboxes(Comm *c) {
    Nccl &n = nccl();
    const char *env = getenv("LBFGSB200_PEER_REDUCE");
    bool ok = !(env && env[0] == '0') && n.AllGather && c->nranks <= kMaxPeers;
    const size_t ring_bytes = sizeof(double) * kMailRing * (size_t)c->nranks * kMailStride;
    const size_t bytes = ring_bytes + 128;   // + the fault word
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    // The staging buffer of the handshake comes first: every later failure is reported THROUGH the collectives
    // below (status byte 0), so no rank is left waiting in them.  If not even these few hundred bytes can be
    // allocated the communicator is aborted, which makes the peers' collectives fail instead of hang.
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 16;
    unsigned char *dev = nullptr;
    std::vector<unsigned char> host(rec * c->nranks, 0);
    if (cudaMalloc((void **)&dev, rec * c->nranks) != cudaSuccess) {
        cudaGetLastError();
        if (n.CommAbort) { n.CommAbort(c->comm); c->comm = nullptr; c->aborted = true; }
        return false;
    }
    if (ok) ok = cudaMalloc((void **)&c->mailbox, bytes) == cudaSuccess;
    if (ok) ok = cudaMemset(c->mailbox, 0, bytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    if (ok) ok = cudaIpcGetMemHandle(&mine, c->mailbox) == cudaSuccess;
    if (!ok) cudaGetLastError();
    // all-gather the handles (+ one status byte per rank) through NCCL
    memcpy(&host[rec * c->rank], &mine, sizeof(mine));
    host[rec * c->rank + sizeof(mine)] = ok ? 1 : 0;
    cudaMemcpy(dev, host.data(), rec * c->nranks, cudaMemcpyHostToDevice);
    bool talk = n.AllGather && n.AllGather(dev + rec * c->rank, dev, rec, /*ncclChar*/ 0, c->comm, nullptr) == ncclSuccess &&
                cudaDeviceSynchronize() == cudaSuccess;
    if (talk) cudaMemcpy(host.data(), dev, rec * c->nranks, cudaMemcpyDeviceToHost);
    bool all_ok = talk;
    for (int r = 0; all_ok && r < c->nranks; ++r) all_ok = host[rec * r + sizeof(mine)] == 1;
    if (all_ok) {
        double *table[kMaxPeers] = {};
        for (int r = 0; r < c->nranks; ++r) {
            if (r == c->rank) { table[r] = c->mailbox; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, &host[rec * r], sizeof(h));
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all_ok = false; cudaGetLastError(); break; }
            c->opened[r] = p;
            table[r] = (double *)p;
        }
        if (all_ok) all_ok = cudaMalloc((void **)&c->mail_table, sizeof(table)) == cudaSuccess &&
                             cudaMemcpy(c->mail_table, table, sizeof(table), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    // second round: did every rank manage to map every peer?
    double *flag = (double *)dev;
    double v = all_ok ? 0.0 : 1.0;
    cudaMemcpy(flag, &v, sizeof(double), cudaMemcpyHostToDevice);
    if (talk && n.AllReduce(flag, flag, 1, ncclFloat64, ncclSum, c->comm, nullptr) == ncclSuccess && cudaDeviceSynchronize() == cudaSuccess) {
        cudaMemcpy(&v, flag, sizeof(double), cudaMemcpyDeviceToHost);
        all_ok = all_ok && v == 0.0;
    } else {
        all_ok = false;
    }
    cudaFree(dev);
    if (!all_ok) {
        for (int r = 0; r < kMaxPeers; ++r) if (c->opened[r]) { cudaIpcCloseMemHandle(c->opened[r]); c->opened[r] = nullptr; }
        if (c->mailbox) { cudaFree(c->mailbox); c->mailbox = nullptr; }
        if (c->mail_table) { cudaFree(c->mail_table); c->mail_table = nullptr; }
        c->peer = PeerCtx{};
        return false;
    }
    c->peer.nranks = c->nranks;
    c->peer.rank = c->rank;
    c->peer.mail_self = c->mailbox;
    c->peer.mail_table = c->mail_table;
    c->peer.seq = 0;
    c->peer.extra[0] = c->peer.extra[1] = nullptr;
    c->peer.fault = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(c->mailbox) + ring_bytes);
    return true;
}
}  // namespace

int comm_allreduce_sum(Comm *c, double *buf_dev, int count, cudaStream_t stream) {
    if (!c || c->nranks == 1) return 0;
    Nccl &n = nccl();
    if (!n.ok || !c->comm) return LBFGSB200_ERR_NCCL;
    int rc = n.AllReduce(buf_dev, buf_dev, (size_t)count, ncclFloat64, ncclSum, c->comm, stream);
    return rc == ncclSuccess ? 0 : LBFGSB200_ERR_NCCL;
}

// All-gather of contiguous shards of unequal length: rank r's `send` (offsets[r+1] - offsets[r] doubles) lands at
// recv_all + offsets[r] on every rank (one grouped ncclBroadcast per rank over NVLink).
int comm_allgatherv(Comm *c, const double *send, double *recv_all, const int64_t *offsets, cudaStream_t stream) {
    if (!c || c->nranks == 1) return 0;
    Nccl &n = nccl();
    if (!n.ok || !c->comm || !n.Broadcast || !n.GroupStart || !n.GroupEnd) return LBFGSB200_ERR_NCCL;
    if (n.GroupStart() != ncclSuccess) return LBFGSB200_ERR_NCCL;
    int bad = 0;
    for (int r = 0; r < c->nranks; ++r) {
        const size_t cnt = (size_t)(offsets[r + 1] - offsets[r]);
        double *dst = recv_all + offsets[r];
        const void *src = (r == c->rank) ? (const void *)send : (const void *)dst;
        if (n.Broadcast(src, dst, cnt, ncclFloat64, r, c->comm, stream) != ncclSuccess) bad = 1;
    }
    if (n.GroupEnd() != ncclSuccess || bad) return LBFGSB200_ERR_NCCL;
    return 0;
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
    if (nranks > 1) {
        const bool peers = lb::setup_peer_mailboxes(c);
        if (getenv("LBFGSB200_DEBUG_TIMING") && rank == 0)
            fprintf(stderr, "[lbfgsb200] scalar all-reduce transport: %s\n", peers ? "peer mailboxes (fused into the kernels)" : "ncclAllReduce");
    }
    *out = reinterpret_cast<lbfgsb200_comm_t *>(c);
    return 0;
}

int lbfgsb200_comm_transport(const lbfgsb200_comm_t *comm) {
    return lb::comm_peer(reinterpret_cast<const lb::Comm *>(comm)) ? 1 : 0;
}

void lbfgsb200_comm_destroy(lbfgsb200_comm_t *comm) {
    lb::Comm *c = reinterpret_cast<lb::Comm *>(comm);
    if (!c) return;
    lb::Nccl &n = lb::nccl();
    for (int r = 0; r < lb::kMaxPeers; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->mail_table) cudaFree(c->mail_table);
    if (n.ok && c->comm) n.CommDestroy(c->comm);
    delete c;
}

int lbfgsb200_comm_allreduce_sum(lbfgsb200_comm_t *comm, double *buf_dev, int count, void *stream) {
    return lb::comm_allreduce_sum(reinterpret_cast<lb::Comm *>(comm), buf_dev, count, (cudaStream_t)stream);
}

}  // extern "C"
