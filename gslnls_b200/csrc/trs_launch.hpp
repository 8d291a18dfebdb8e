// trs_launch.hpp -- host-callable launchers of the K3 kernels (trs_kernel.cu)
#pragma once
#include <cuda_runtime.h>

#include "trs_core.h"

namespace gslnls {
// what the resident server needs to sum the partial packets of the persistent pass kernel itself
struct TrsLinks {
    const double *partials = nullptr; // [nctas][pk_stride] per-CTA partial packets; nullptr: packets arrive reduced
    int nctas = 0, pk_stride = 0, rank = 0, pad_ = 0;
    char *peer[8] = {nullptr}; // every rank's channel block as mapped on this GPU
};
int trs_max_p();
cudaError_t trs_launch_step(const trs::Params &P, double *state, const double *packet, double *req,
                            double *partrace, double *ssrtrace, double *condtrace, int *ndone,
                            cudaStream_t stream);
cudaError_t trs_launch_step_batch(const trs::Params &P, double *states, int state_stride, const double *packets,
                                  int pk_stride, double *reqs, int req_stride, int ncand, int *ndone,
                                  cudaStream_t stream);
cudaError_t trs_launch_reset(double *states, int state_stride, double *reqs, int req_stride, const double *starts,
                             int p, int ncand, int *ndone, cudaStream_t stream);
cudaError_t trs_launch_set_request(double *req, int mode, const double *theta, const double *v, int p,
                                   cudaStream_t stream);
// resident-server mode (one fit, one candidate): channel bookkeeping + the server kernel itself
int trs_server_max_p();
cudaError_t trs_launch_channel_begin(char *channel, cudaStream_t stream);
// reset + channel_begin in one launch, start values by value (p <= 64)
cudaError_t trs_launch_fit_begin(double *state, double *req, const double *start_host, int p, int *ndone,
                                 char *channel, cudaStream_t stream);
cudaError_t trs_launch_server(const trs::Params &P, char *channel, int nranks, int pk_count, double *state,
                              double *packet, double *req, double *partrace, double *ssrtrace, double *condtrace,
                              int *ndone, int *host_flags_dev, double *host_state_dev, unsigned long long watchdog_ns,
                              unsigned long long handshake_ns, const TrsLinks &links, cudaStream_t stream);
cudaError_t launch_sum_rank_packets(const double *gathered, double *packet, int count, int nranks,
                                    cudaStream_t stream);
} // namespace gslnls
