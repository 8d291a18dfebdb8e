// comm.cpp -- packet exchange between the GPUs of one box, one process per GPU.
// The only data that ever crosses GPUs is the p(p+1)/2 + p + 1 double packet (80 B at p = 3) per
// pass, so the transport is NCCL's all-reduce over NVLink; libnccl is resolved lazily so that the
// library loads (and single-GPU fits run) on machines without it.
#include "comm.hpp"

#include <dlfcn.h>

#include <cstring>
#include <string>

#include <cstdlib>
#include <vector>

#include "../../include/gslnls_b200.h"
#include "nls_abi.h"

namespace gslnls {
void set_error(const std::string &s);

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t *);
typedef int (*fn_CommInitRank)(void **, int, ncclUniqueId_t, int);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_AllGather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*fn_CommDestroy)(void *);
typedef int (*fn_CommInitAll)(void **, int, const int *);
typedef const char *(*fn_GetErrorString)(int);

struct Nccl {
    void *h = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_CommInitAll CommInitAll = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
    bool ok = false;
};

Nccl &nccl()
{
    static Nccl N;
    if (N.h)
        return N;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        N.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (N.h)
            break;
    }
    if (!N.h)
        return N;
    N.GetUniqueId = (fn_GetUniqueId)dlsym(N.h, "ncclGetUniqueId");
    N.CommInitRank = (fn_CommInitRank)dlsym(N.h, "ncclCommInitRank");
    N.AllReduce = (fn_AllReduce)dlsym(N.h, "ncclAllReduce");
    N.AllGather = (fn_AllGather)dlsym(N.h, "ncclAllGather");
    N.CommDestroy = (fn_CommDestroy)dlsym(N.h, "ncclCommDestroy");
    N.CommInitAll = (fn_CommInitAll)dlsym(N.h, "ncclCommInitAll");
    N.GetErrorString = (fn_GetErrorString)dlsym(N.h, "ncclGetErrorString");
    N.ok = N.GetUniqueId && N.CommInitRank && N.AllReduce && N.AllGather && N.CommDestroy;
    return N;
}
} // namespace

int comm_allreduce_sum(gslnls_comm *c, double *dev_buf, size_t count, cudaStream_t stream)
{
    if (!c || c->nranks <= 1)
        return GSLNLS_SUCCESS;
    Nccl &N = nccl();
    const int ncclFloat64 = 8, ncclSum = 0;
    const int rc = N.AllReduce(dev_buf, dev_buf, count, ncclFloat64, ncclSum, c->nccl, stream);
    if (rc != 0) {
        set_error(std::string("ncclAllReduce: ") + (N.GetErrorString ? N.GetErrorString(rc) : "error"));
        return GSLNLS_ECOMM;
    }
    return GSLNLS_SUCCESS;
}

// every rank receives all ranks' `count` doubles, rank-major; the caller sums them in rank order so
// that all ranks hold bitwise the same packet whatever NCCL's internal reduction order would be
int comm_allgather(gslnls_comm *c, const double *dev_send, double *dev_recv, size_t count, cudaStream_t stream)
{
    if (!c || c->nranks <= 1)
        return GSLNLS_SUCCESS;
    Nccl &N = nccl();
    const int ncclFloat64 = 8;
    const int rc = N.AllGather(dev_send, dev_recv, count, ncclFloat64, c->nccl, stream);
    if (rc != 0) {
        set_error(std::string("ncclAllGather: ") + (N.GetErrorString ? N.GetErrorString(rc) : "error"));
        return GSLNLS_ECOMM;
    }
    return GSLNLS_SUCCESS;
}
} // namespace gslnls

using namespace gslnls;

// Map every rank's channel block into this process (cudaIpc over NVLink peer access).  The 64-byte
// handles travel through one NCCL all-gather; all ranks then agree (a second all-gather of an "ok"
// word) on whether the peer path is usable, so nobody waits on a mailbox nobody writes.
static void setup_peer_channel(gslnls_comm *c)
{
    Nccl &N = nccl();
    const int R = c->nranks;
    cudaStream_t st = nullptr;
    double *d_send = nullptr, *d_recv = nullptr;
    const size_t words = sizeof(cudaIpcMemHandle_t) / sizeof(double) + 1; // handle + ok word
    std::vector<double> h_send(words, 0.0), h_recv(words * R, 0.0);
    bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&c->channel, NLS_CH_BYTES) == cudaSuccess;
    ok = ok && cudaMemset(c->channel, 0, NLS_CH_BYTES) == cudaSuccess;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    ok = ok && cudaIpcGetMemHandle(&mine, c->channel) == cudaSuccess;
    std::memcpy(h_send.data(), &mine, sizeof(mine));
    h_send[words - 1] = ok ? 1.0 : 0.0;
    bool xfer = cudaMalloc(&d_send, sizeof(double) * words) == cudaSuccess &&
                cudaMalloc(&d_recv, sizeof(double) * words * R) == cudaSuccess && st != nullptr;
    // the collectives below must be entered by every rank, whatever happened locally
    const int ncclFloat64 = 8;
    if (xfer) {
        cudaMemcpyAsync(d_send, h_send.data(), sizeof(double) * words, cudaMemcpyHostToDevice, st);
        xfer = N.AllGather(d_send, d_recv, words, ncclFloat64, c->nccl, st) == 0;
        cudaMemcpyAsync(h_recv.data(), d_recv, sizeof(double) * words * R, cudaMemcpyDeviceToHost, st);
        xfer = cudaStreamSynchronize(st) == cudaSuccess && xfer;
    }
    bool all_ok = xfer;
    for (int r = 0; r < R && all_ok; ++r)
        all_ok = h_recv[(size_t)r * words + words - 1] == 1.0;
    if (all_ok) {
        for (int r = 0; r < R; ++r) {
            if (r == c->rank) {
                c->peer_channel[r] = c->channel;
                continue;
            }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, &h_recv[(size_t)r * words], sizeof(h));
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                all_ok = false;
                break;
            }
            c->peer_channel[r] = (char *)ptr;
        }
    }
    // second round: did every rank manage to open every handle?
    double flag = all_ok ? 1.0 : 0.0;
    if (xfer) {
        cudaMemcpyAsync(d_send, &flag, sizeof(double), cudaMemcpyHostToDevice, st);
        bool g = N.AllGather(d_send, d_recv, 1, ncclFloat64, c->nccl, st) == 0;
        cudaMemcpyAsync(h_recv.data(), d_recv, sizeof(double) * R, cudaMemcpyDeviceToHost, st);
        g = cudaStreamSynchronize(st) == cudaSuccess && g;
        for (int r = 0; r < R && g; ++r)
            g = h_recv[r] == 1.0;
        all_ok = all_ok && g;
    }
    c->p2p = all_ok;
    cudaFree(d_send);
    cudaFree(d_recv);
    if (st)
        cudaStreamDestroy(st);
    cudaGetLastError();
}

extern "C" {

GSLNLS_API void gslnls_comm_free(gslnls_comm *c);

GSLNLS_API int gslnls_comm_get_unique_id(void *id_bytes)
{
    if (!id_bytes)
        return GSLNLS_EINVAL;
    Nccl &N = nccl();
    if (!N.ok) {
        set_error("libnccl.so.2 not found");
        return GSLNLS_ECOMM;
    }
    ncclUniqueId_t id;
    if (N.GetUniqueId(&id) != 0) {
        set_error("ncclGetUniqueId failed");
        return GSLNLS_ECOMM;
    }
    static_assert(sizeof(id) == GSLNLS_COMM_ID_BYTES, "id size");
    std::memcpy(id_bytes, &id, sizeof(id));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_comm_create(const void *id_bytes, int rank, int nranks, int device, gslnls_comm **out)
{
    if (!out || nranks < 1 || rank < 0 || rank >= nranks)
        return GSLNLS_EINVAL;
    *out = nullptr;
    gslnls_comm *c = new gslnls_comm();
    c->rank = rank;
    c->nranks = nranks;
    c->device = device;
    if (nranks > 1) {
        Nccl &N = nccl();
        if (!N.ok || !id_bytes) {
            delete c;
            set_error("libnccl.so.2 not found or no unique id");
            return GSLNLS_ECOMM;
        }
        if (cudaSetDevice(device) != cudaSuccess) {
            delete c;
            set_error("cudaSetDevice failed");
            return GSLNLS_ENODEVICE;
        }
        ncclUniqueId_t id;
        std::memcpy(&id, id_bytes, sizeof(id));
        const int rc = N.CommInitRank(&c->nccl, nranks, id, rank);
        if (rc != 0) {
            set_error(std::string("ncclCommInitRank: ") + (N.GetErrorString ? N.GetErrorString(rc) : "error"));
            delete c;
            return GSLNLS_ECOMM;
        }
    }
    if (nranks > 1 && nranks <= NLS_MAX_RANKS) {
        const char *env = std::getenv("GSLNLS_P2P");
        if (!env || std::atoi(env) != 0)
            setup_peer_channel(c); // on failure the NCCL all-gather path remains
    }
    *out = c;
    return GSLNLS_SUCCESS;
}

// All ranks of one process (one host thread per GPU drives its own rank): the channel blocks are
// reached through peer access in the unified address space, no cudaIpc, no rendezvous.  NCCL is
// initialised too when available (ncclCommInitAll) for the paths that do not use the mailboxes.
GSLNLS_API int gslnls_comm_create_local(int ndev, const int *devices, gslnls_comm **out)
{
    if (!out || !devices || ndev < 1 || ndev > NLS_MAX_RANKS)
        return GSLNLS_EINVAL;
    for (int r = 0; r < ndev; ++r)
        out[r] = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess)
        have = 0;
    for (int r = 0; r < ndev; ++r) {
        if (devices[r] < 0 || devices[r] >= have) {
            set_error("no such CUDA device");
            return GSLNLS_ENODEVICE;
        }
        for (int q = 0; q < r; ++q)
            if (devices[q] == devices[r]) {
                set_error("a device may appear only once");
                return GSLNLS_EINVAL;
            }
    }
    int prev = 0;
    cudaGetDevice(&prev);
    bool ok = true;
    for (int r = 0; r < ndev && ok; ++r) {
        gslnls_comm *c = new gslnls_comm();
        c->rank = r;
        c->nranks = ndev;
        c->device = devices[r];
        c->local = true;
        out[r] = c;
        if (ndev > 1) {
            ok = cudaSetDevice(devices[r]) == cudaSuccess && cudaMalloc(&c->channel, NLS_CH_BYTES) == cudaSuccess &&
                 cudaMemset(c->channel, 0, NLS_CH_BYTES) == cudaSuccess;
        }
    }
    bool p2p = ok && ndev > 1;
    for (int r = 0; r < ndev && p2p; ++r) {
        cudaSetDevice(devices[r]);
        for (int q = 0; q < ndev && p2p; ++q) {
            if (q == r)
                continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[r], devices[q]);
            if (!can) {
                p2p = false;
                break;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                p2p = false;
            cudaGetLastError();
        }
    }
    const char *env = std::getenv("GSLNLS_P2P");
    if (env && std::atoi(env) == 0)
        p2p = false;
    for (int r = 0; r < ndev && ok; ++r) {
        out[r]->p2p = p2p;
        for (int q = 0; q < ndev; ++q)
            out[r]->peer_channel[q] = out[q]->channel;
    }
    if (ok && ndev > 1) {
        Nccl &N = nccl();
        if (N.ok && N.CommInitAll) {
            void *comms[NLS_MAX_RANKS] = {nullptr};
            if (N.CommInitAll(comms, ndev, devices) == 0)
                for (int r = 0; r < ndev; ++r)
                    out[r]->nccl = comms[r];
        }
        if (!p2p && !out[0]->nccl) {
            set_error("the devices have neither peer access nor a usable libnccl.so.2");
            ok = false;
        }
    }
    cudaSetDevice(prev);
    if (!ok) {
        set_error("could not set up the device group (allocation, peer access or NCCL initialisation failed)");
        for (int r = 0; r < ndev; ++r) {
            gslnls_comm_free(out[r]);
            out[r] = nullptr;
        }
        return GSLNLS_ECOMM;
    }
    return GSLNLS_SUCCESS;
}

GSLNLS_API void gslnls_comm_free(gslnls_comm *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    for (int r = 0; r < c->nranks && r < NLS_MAX_RANKS && !c->local; ++r)
        if (r != c->rank && c->peer_channel[r])
            cudaIpcCloseMemHandle(c->peer_channel[r]);
    if (c->channel)
        cudaFree(c->channel);
    if (c->nccl)
        nccl().CommDestroy(c->nccl);
    delete c;
}

GSLNLS_API int gslnls_comm_has_peer_memory(const gslnls_comm *c) { return c && c->p2p ? 1 : 0; }

GSLNLS_API int gslnls_comm_rank(const gslnls_comm *c) { return c ? c->rank : 0; }
GSLNLS_API int gslnls_comm_size(const gslnls_comm *c) { return c ? c->nranks : 1; }

} // extern "C"
