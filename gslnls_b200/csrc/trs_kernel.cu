// trs_kernel.cu -- K3: device instantiations of the trust-region state machine (trs_core.h).
//   trs_step_warp  : one warp per fit, p x p matrices in shared memory, cooperative Cholesky
//   trs_step_batch : one thread per multi-start candidate (src/nls_mstart.c:75-91 inner loops)
// Compiled offline for sm_100a into libgslnls_b200.so.
#include <cuda_runtime.h>

#include "nls_abi.h"
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

// ---------------------------------------------------------------------------------------------
// trs_server: the trust-region warp stays resident for a whole fit.  It waits until the packet of
// pass k has been deposited in this GPU's mailbox by every rank's pass kernel (local stores, or
// NVLink peer stores from the other GPUs), adds the rank packets in rank order -- so all GPUs hold
// bitwise the same normal equations and take bitwise the same step without any broadcast --,
// advances the state machine and publishes request k+1, on which the already-dispatched next pass
// kernel is spinning.  Exchange + step are one resident kernel: no collective call, no launch and
// no cold instruction cache between two passes.
static __device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
static __device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

static __device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <int PMAX, bool FIXED>
__global__ void __launch_bounds__(32) trs_server(const trs::Params P, char *channel, int nranks, int pk_count,
                                                 double *state, double *packet, double *req, double *partrace,
                                                 double *ssrtrace, double *condtrace, int *ndone,
                                                 volatile int *host_flags, double *host_state,
                                                 unsigned long long watchdog_ns, unsigned long long handshake_ns,
                                                 const TrsLinks links)
{
    extern __shared__ double trs_smem[];
    const int lane = threadIdx.x & 31;
    unsigned long long *req_seq = (unsigned long long *)(channel + NLS_CH_REQ_SEQ);
    const unsigned long long *flags = (const unsigned long long *)(channel + NLS_CH_FLAGS);
    const unsigned long long *abort_w = (const unsigned long long *)(channel + NLS_CH_ABORT);
    const double *mbox = (const double *)(channel + NLS_CH_DATA);
    const unsigned long long *pass_seen = (const unsigned long long *)(channel + NLS_CH_PASS_SEEN);
    const unsigned long long *cta_count = (const unsigned long long *)(channel + NLS_CH_CTA_COUNT);
    // server_reduce: the (persistent) pass kernel's CTAs deposit partial packets and count themselves in; this
    // warp sums them in CTA order -- and, with several GPUs, forwards the rank packet to every peer's mailbox
    const bool reduce_here = links.partials != nullptr;
    const unsigned long long k0 = __ldcg((const unsigned long long *)(channel + NLS_CH_FIT_SEQ0));
    unsigned long long k = k0;
    // start-of-fit handshake: until the first pass kernel of this fit has been seen RUNNING next to this
    // kernel, the wait below is bounded by handshake_ns, not by the watchdog.  Kernels are not guaranteed to
    // run concurrently (ncu, compute-sanitizer, cuda-gdb and CUDA_LAUNCH_BLOCKING=1 serialise them): the
    // server then leaves with host flag 3 and the host steps the fit launch-ordered (K1 -> K3 -> K1 ...).
    bool partner_seen = false;
    trs::Solver<PMAX, trs::WarpLanes, FIXED> S(P, trs::WarpLanes(), trs_smem, trs_smem + P.p * P.p);
    S.keep_state = true; // the record is read once and written back when the fit ends or the server leaves
    if ((int)__ldcg(state + trs::S_PHASE) == trs::PH_DONE)
        return;
    for (;; ++k) {
        // ---- wait for pass k: from the local CTAs (server_reduce), then from every rank (lane r watches rank r) ----
        unsigned long long t0 = 0ull, spins = 0ull;
        int why = 0; // 0 packet, 1 abort, 2 watchdog, 3 no concurrent pass kernel (handshake)
        bool local_done = !reduce_here;
        for (;;) {
            if (!local_done) {
                if (ld_acquire_gpu(cta_count) >= (k - k0 + 1ull) * (unsigned long long)links.nctas) {
                    // every CTA's partial is in: sum them in CTA order, lanes striding over the CTAs
                    const unsigned long long t_in = globaltimer_ns();
                    // all loads of this lane go out before the first is used: one L2 round trip for the whole
                    // 147 x 11 block instead of one per packet entry (measured: 7 us -> ~1 us)
                    constexpr int kMaxPk = 16, kChunks = 5; // persistent mode: p <= 4, at most 160 CTAs
                    double v[kMaxPk][kChunks];
#pragma unroll
                    for (int e = 0; e < kMaxPk; ++e)
#pragma unroll
                        for (int j = 0; j < kChunks; ++j) {
                            const int b = lane + 32 * j;
                            v[e][j] = (e < pk_count && b < links.nctas) ? __ldcg(links.partials + (size_t)b * links.pk_stride + e) : 0.0;
                        }
#pragma unroll
                    for (int e = 0; e < kMaxPk; ++e) {
                        if (e >= pk_count)
                            break;
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < kChunks; ++j)
                            if (lane + 32 * j < links.nctas)
                                s += v[e][j];
                        s += __shfl_down_sync(0xffffffffu, s, 16);
                        s += __shfl_down_sync(0xffffffffu, s, 8);
                        s += __shfl_down_sync(0xffffffffu, s, 4);
                        s += __shfl_down_sync(0xffffffffu, s, 2);
                        s += __shfl_down_sync(0xffffffffu, s, 1);
                        s = __shfl_sync(0xffffffffu, s, 0);
                        if (nranks > 1) {
                            // rank packet -> slot [parity][rank] of every rank's mailbox (NVLink peer stores)
                            const size_t slot = ((size_t)(k & 1ull) * NLS_MAX_RANKS + (size_t)links.rank) * NLS_CH_MAXPK;
                            if (lane < nranks)
                                ((double *)(links.peer[lane] + NLS_CH_DATA))[slot + e] = s;
                        } else if (lane == 0) {
                            packet[e] = s;
                        }
                    }
                    if (lane == 0) {
                        unsigned long long *tm = (unsigned long long *)(channel + NLS_CH_TIMER);
                        tm[1] += t_in - __ldcg(tm); // request seen by the pass kernel -> all partials in
                        tm[2] += 1ull;
                        *(unsigned long long *)(channel + NLS_CH_PASS_CTR) = k;
                    }
                    if (nranks > 1) {
                        __threadfence_system();
                        if (lane < nranks)
                            st_release_sys((unsigned long long *)(links.peer[lane] + NLS_CH_FLAGS + 128 * links.rank), k);
                    }
                    local_done = true;
                    partner_seen = true;
                    if (nranks == 1)
                        break;
                    continue;
                }
            } else {
                const bool have = lane >= nranks || ld_acquire_sys(flags + 16 * lane) >= k;
                if (__all_sync(0xffffffffu, have))
                    break;
            }
            if ((++spins & 255ull) == 0ull) {
                unsigned long long a = 0ull, t = 0ull, seen = 0ull;
                if (lane == 0) {
                    a = ld_acquire_sys(abort_w);
                    t = globaltimer_ns();
                    if (!partner_seen)
                        seen = ld_acquire_sys(pass_seen);
                }
                a = __shfl_sync(0xffffffffu, a, 0);
                t = __shfl_sync(0xffffffffu, t, 0);
                seen = __shfl_sync(0xffffffffu, seen, 0);
                if (a) {
                    why = 1;
                    break;
                }
                if (!partner_seen && seen >= k)
                    partner_seen = true;
                if (t0 == 0ull)
                    t0 = t;
                else if (!partner_seen && t - t0 > handshake_ns) {
                    why = 3;
                    break;
                } else if (t - t0 > watchdog_ns) {
                    why = 2;
                    break;
                }
            }
        }
        if (why == 1) {
            // fit_end before completion: request k stays published; the state record is brought up to date
            S.flush(state);
            __threadfence();
            return;
        }
        if (why == 3) {
            // nothing has been consumed: the state record is still the start of the fit.  Queued pass
            // launches fall through as idle no-ops (request mode IDLE, request sequence far ahead).
            if (lane == 0) {
                req[0] = (double)trs::MODE_IDLE;
                __threadfence();
                st_release_gpu(req_seq, k + NLS_CH_GONE_BUMP);
                __threadfence_system();
                host_flags[0] = 3;
            }
            return;
        }
        partner_seen = true; // a packet arrived: the kernels do run side by side
        if (why == 2) {
            // a peer never delivered: fail the fit instead of hanging the GPU
            S.flush(state);
            __syncwarp();
            if (lane == 0) {
                state[trs::S_STATUS] = (double)trs::E_FAILURE;
                state[trs::S_PHASE] = (double)trs::PH_DONE;
                req[0] = (double)trs::MODE_IDLE;
                __threadfence();
                st_release_gpu(req_seq, k + 1ull);
                atomicAdd(ndone, 1);
                __threadfence_system();
                host_flags[0] = 2;
            }
            return;
        }
        const unsigned long long t_in = globaltimer_ns();
        if (!reduce_here || nranks > 1) {
            // ---- rank-ordered sum of the deposited packets ----
            const double *slot = mbox + (size_t)(k & 1ull) * NLS_MAX_RANKS * NLS_CH_MAXPK;
            for (int e = lane; e < pk_count; e += 32) {
                double s = __ldcg(slot + e);
                for (int r = 1; r < nranks; ++r)
                    s += __ldcg(slot + (size_t)r * NLS_CH_MAXPK + e);
                packet[e] = s;
            }
        }
        __syncwarp();
        S.advance(state, packet, req, partrace, ssrtrace, condtrace);
        // the request record (and the trace rows) are written by lane 0 alone: its release store orders them; a
        // finished fit also stored the state record from every lane, which the fence below the loop covers
        if (S.phase == trs::PH_DONE)
            __threadfence();
        __syncwarp();
        if (lane == 0) {
            st_release_gpu(req_seq, k + 1ull);
            unsigned long long *tm = (unsigned long long *)(channel + NLS_CH_TIMER);
            tm[3] += globaltimer_ns() - t_in; // packet complete -> next request published, summed
            tm[4] += 1ull;
        }
        if (S.phase == trs::PH_DONE) {
            // the final state record goes straight to mapped host memory, then the done word: the host
            // needs no copy and no stream synchronisation to return the result
            if (host_state) {
                const int ns = trs::state_doubles(P.p);
                for (int e = lane; e < ns; e += 32)
                    host_state[e] = __ldcg(state + e);
            }
            __threadfence_system();
            __syncwarp();
            if (lane == 0) {
                atomicAdd(ndone, 1);
                __threadfence_system();
                host_flags[0] = 1;
            }
            return;
        }
    }
}

// start-of-fit bookkeeping of the channel (stream-ordered after every earlier pass kernel)
__global__ void trs_channel_begin(char *channel)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned long long done = *(unsigned long long *)(channel + NLS_CH_PASS_CTR);
        *(unsigned long long *)(channel + NLS_CH_FIT_SEQ0) = done + 1ull;
        *(unsigned long long *)(channel + NLS_CH_ABORT) = 0ull;
        *(unsigned long long *)(channel + NLS_CH_PASS_SEEN) = 0ull; // idle launches behind the last fit wrote it
        __threadfence();
        *(unsigned long long *)(channel + NLS_CH_REQ_SEQ) = done + 1ull; // request 1 of this fit = trs_reset's
    }
}

// start of a fit in resident-server mode, one launch: state and request records from start values passed BY
// VALUE (no host-to-device copy in front of the fit) plus the channel bookkeeping of trs_channel_begin
struct TrsStart {
    double v[64];
};
__global__ void trs_fit_begin(double *state, double *req, const TrsStart st, int p, int *ndone, char *channel)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        *ndone = 0;
        trs::state_reset(state, req, st.v, p);
        const unsigned long long done = *(unsigned long long *)(channel + NLS_CH_PASS_CTR);
        *(unsigned long long *)(channel + NLS_CH_FIT_SEQ0) = done + 1ull;
        *(unsigned long long *)(channel + NLS_CH_ABORT) = 0ull;
        *(unsigned long long *)(channel + NLS_CH_PASS_SEEN) = 0ull;
        *(unsigned long long *)(channel + NLS_CH_CTA_COUNT) = 0ull;
        __threadfence();
        *(unsigned long long *)(channel + NLS_CH_REQ_SEQ) = done + 1ull;
    }
}

cudaError_t trs_launch_fit_begin(double *state, double *req, const double *start_host, int p, int *ndone,
                                 char *channel, cudaStream_t stream)
{
    if (p > 64)
        return cudaErrorInvalidValue;
    TrsStart st;
    for (int i = 0; i < 64; ++i)
        st.v[i] = i < p ? start_host[i] : 0.0;
    trs_fit_begin<<<1, 32, 0, stream>>>(state, req, st, p, ndone, channel);
    return cudaGetLastError();
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
int trs_server_max_p() { return 64; }

cudaError_t trs_launch_channel_begin(char *channel, cudaStream_t stream)
{
    trs_channel_begin<<<1, 32, 0, stream>>>(channel);
    return cudaGetLastError();
}

cudaError_t trs_launch_server(const trs::Params &P, char *channel, int nranks, int pk_count, double *state,
                              double *packet, double *req, double *partrace, double *ssrtrace, double *condtrace,
                              int *ndone, int *host_flags_dev, double *host_state_dev, unsigned long long watchdog_ns,
                              unsigned long long handshake_ns, const TrsLinks &links, cudaStream_t stream)
{
    const size_t smem = sizeof(double) * 2 * (size_t)P.p * P.p;
#define TRS_SERVER_LAUNCH(PM, FX)                                                                                  \
    trs_server<PM, FX><<<1, 32, smem, stream>>>(P, channel, nranks, pk_count, state, packet, req, partrace, ssrtrace, \
                                                condtrace, ndone, host_flags_dev, host_state_dev, watchdog_ns, \
                                                handshake_ns, links)
    if (P.p == 2)
        TRS_SERVER_LAUNCH(2, true);
    else if (P.p == 3)
        TRS_SERVER_LAUNCH(3, true);
    else if (P.p == 4)
        TRS_SERVER_LAUNCH(4, true);
    else if (P.p <= 8)
        TRS_SERVER_LAUNCH(8, false);
    else if (P.p <= 32)
        TRS_SERVER_LAUNCH(32, false);
    else if (P.p <= 64) {
        // 2 p^2 doubles of shared memory: 64 KB at p = 64, above the 48 KB default limit
        cudaError_t e = cudaFuncSetAttribute(trs_server<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(double) * 2 * 64 * 64));
        if (e != cudaSuccess)
            return e;
        TRS_SERVER_LAUNCH(64, false);
    }
#undef TRS_SERVER_LAUNCH
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

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
        // per device, and cheap: set it on every launch rather than remembering which devices have it
        cudaError_t e = cudaFuncSetAttribute(trs_step_warp<100>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)warp_smem_bytes(100));
        if (e != cudaSuccess)
            return e;
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
