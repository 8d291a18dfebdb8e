// comm.cpp -- packet exchange between the GPUs of one box, one process per GPU.
// The only data that ever crosses GPUs is the p(p+1)/2 + p + 1 double packet (80 B at p = 3) per
// pass, so the transport is NCCL's all-reduce over NVLink; libnccl is resolved lazily so that the
// library loads (and single-GPU fits run) on machines without it.
#include "comm.hpp"

#include <dlfcn.h>

#include <cstring>
#include <string>

#include "../../include/gslnls_b200.h"

namespace gslnls {
void set_error(const std::string &s);

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t *);
typedef int (*fn_CommInitRank)(void **, int, ncclUniqueId_t, int);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_AllGather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*fn_CommDestroy)(void *);
typedef const char *(*fn_GetErrorString)(int);

struct Nccl {
    void *h = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
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

extern "C" {

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
    *out = c;
    return GSLNLS_SUCCESS;
}

GSLNLS_API void gslnls_comm_free(gslnls_comm *c)
{
    if (!c)
        return;
    if (c->nccl)
        nccl().CommDestroy(c->nccl);
    delete c;
}

GSLNLS_API int gslnls_comm_rank(const gslnls_comm *c) { return c ? c->rank : 0; }
GSLNLS_API int gslnls_comm_size(const gslnls_comm *c) { return c ? c->nranks : 1; }

} // extern "C"
