// trs_kernel.cu -- K3: device instantiations of the trust-region state machine (trs_core.h).
//   trs_step_warp  : one warp per fit, p x p matrices in shared memory, cooperative Cholesky
//   trs_step_batch : one thread per multi-start candidate (src/nls_mstart.c:75-91 inner loops)
// Compiled offline for sm_100a into libgslnls_b200.so.
#include <cuda_runtime.h>

#include "trs_core.h"
#include "trs_launch.hpp"

namespace gslnls {

template <int PMAX>
__global__ void __launch_bounds__(32) trs_step_warp(const trs::Params P, double *state, const double *packet,
                                                    double *req, double *partrace, double *ssrtrace,
                                                    double *condtrace, int *ndone)
{
    extern __shared__ double trs_smem[];
    const int before = (int)state[trs::S_PHASE];
    if (before == trs::PH_DONE)
        return;
    trs::Solver<PMAX, trs::WarpLanes> S(P, trs::WarpLanes(), trs_smem, trs_smem + P.p * P.p);
    S.advance(state, packet, req, partrace, ssrtrace, condtrace);
    __syncwarp();
    if ((threadIdx.x & 31) == 0 && S.phase == trs::PH_DONE)
        atomicAdd(ndone, 1);
}

template <int PMAX>
__global__ void __launch_bounds__(128) trs_step_batch(const trs::Params P, double *states, int state_stride,
                                                      const double *packets, int pk_stride, double *reqs,
                                                      int req_stride, int ncand, int *ndone)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand)
        return;
    double *state = states + (size_t)c * state_stride;
    if ((int)state[trs::S_PHASE] == trs::PH_DONE)
        return;
    double jtj[PMAX * PMAX], work[PMAX * PMAX];
    trs::Solver<PMAX, trs::SingleLane> S(P, trs::SingleLane(), jtj, work);
    S.advance(state, packets + (size_t)c * pk_stride, reqs + (size_t)c * req_stride, nullptr, nullptr, nullptr);
    if (S.phase == trs::PH_DONE)
        atomicAdd(ndone, 1);
}

// reset kernels: write the initial state / request records on the device
__global__ void trs_reset(double *states, int state_stride, double *reqs, int req_stride, const double *starts,
                          int p, int ncand, int *ndone)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0)
        *ndone = 0;
    if (c >= ncand)
        return;
    trs::state_reset(states + (size_t)c * state_stride, reqs + (size_t)c * req_stride, starts + (size_t)c * p, p);
}

__global__ void trs_set_request(double *req, int mode, const double *theta, const double *v, int p)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        req[0] = (double)mode;
        for (int i = 0; i < p; ++i) {
            req[1 + i] = theta[i];
            req[1 + p + i] = v ? v[i] : 0.0;
        }
    }
}

// packet[e] = sum over ranks, in rank order, of gathered[r][e]
__global__ void sum_rank_packets(const double *gathered, double *packet, int count, int nranks)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count)
        return;
    double s = 0.0;
    for (int r = 0; r < nranks; ++r)
        s += gathered[(size_t)r * count + e];
    packet[e] = s;
}

cudaError_t launch_sum_rank_packets(const double *gathered, double *packet, int count, int nranks,
                                    cudaStream_t stream)
{
    sum_rank_packets<<<(count + 127) / 128, 128, 0, stream>>>(gathered, packet, count, nranks);
    return cudaGetLastError();
}

int trs_max_p() { return 100; }

static size_t warp_smem_bytes(int p) { return sizeof(double) * 2 * (size_t)p * p; }

cudaError_t trs_launch_step(const trs::Params &P, double *state, const double *packet, double *req,
                            double *partrace, double *ssrtrace, double *condtrace, int *ndone,
                            cudaStream_t stream)
{
    const size_t smem = warp_smem_bytes(P.p);
    if (P.p <= 8) {
        trs_step_warp<8><<<1, 32, smem, stream>>>(P, state, packet, req, partrace, ssrtrace, condtrace, ndone);
    } else if (P.p <= 32) {
        trs_step_warp<32><<<1, 32, smem, stream>>>(P, state, packet, req, partrace, ssrtrace, condtrace, ndone);
    } else if (P.p <= 100) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(trs_step_warp<100>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)warp_smem_bytes(100));
            if (e != cudaSuccess)
                return e;
            attr_set = true;
        }
        trs_step_warp<100><<<1, 32, smem, stream>>>(P, state, packet, req, partrace, ssrtrace, condtrace, ndone);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t trs_launch_step_batch(const trs::Params &P, double *states, int state_stride, const double *packets,
                                  int pk_stride, double *reqs, int req_stride, int ncand, int *ndone,
                                  cudaStream_t stream)
{
    if (P.p > 8)
        return cudaErrorInvalidValue;
    const int threads = 128, blocks = (ncand + threads - 1) / threads;
    trs_step_batch<8><<<blocks, threads, 0, stream>>>(P, states, state_stride, packets, pk_stride, reqs,
                                                     req_stride, ncand, ndone);
    return cudaGetLastError();
}

cudaError_t trs_launch_reset(double *states, int state_stride, double *reqs, int req_stride, const double *starts,
                             int p, int ncand, int *ndone, cudaStream_t stream)
{
    const int threads = 128, blocks = (ncand + threads - 1) / threads;
    trs_reset<<<blocks, threads, 0, stream>>>(states, state_stride, reqs, req_stride, starts, p, ncand, ndone);
    return cudaGetLastError();
}

cudaError_t trs_launch_set_request(double *req, int mode, const double *theta, const double *v, int p,
                                   cudaStream_t stream)
{
    trs_set_request<<<1, 32, 0, stream>>>(req, mode, theta, v, p);
    return cudaGetLastError();
}

} // namespace gslnls
