// comm.hpp -- multi-GPU exchange of the normal-equation packet (one process per GPU).
#pragma once
#include <cuda_runtime.h>

struct gslnls_comm {
    void *nccl = nullptr; // ncclComm_t (bootstrap, and the exchange when peer memory is unavailable)
    int rank = 0, nranks = 1, device = 0;
    // peer-memory channel (nls_abi.h NLS_CH_*): this rank's block and every rank's block as mapped here
    char *channel = nullptr;
    char *peer_channel[8] = {nullptr};
    bool p2p = false;
    bool local = false;         // created by gslnls_comm_create_local: peers are mapped by peer access, not cudaIpc
    long long n_total_hint = -1; // local groups: the caller knows the global row count (no size exchange needed)
};

namespace gslnls {
// sum `count` doubles in place across ranks on `stream`; every rank receives bitwise the same result
int comm_allreduce_sum(gslnls_comm *c, double *dev_buf, size_t count, cudaStream_t stream);
int comm_allgather(gslnls_comm *c, const double *dev_send, double *dev_recv, size_t count, cudaStream_t stream);
}
