// problem.cu -- host orchestration behind the C ABI (include/gslnls_b200.h).
//
// Plays the role of C_nls_large_internal (src/nls_large.c:77-424) and of the driver loop
// gsl_multilarge_nlinear_driver2 (src/nls_fit.c:153-224), except that the loop body is two kernels
// per trial step -- K1 fused pass (NVRTC, nls_pass_kernel.cuh) and K3 trust-region step
// (trs_kernel.cu) -- enqueued in chunks on one stream; the host only reads back a completion
// counter between chunks.  The p x p state never leaves the device during a fit.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gslnls_b200.h"
#include "comm.hpp"
#include "model.hpp"
#include "mstart.hpp"
#include "nls_abi.h"
#include "trs_launch.hpp"
#include "upload.hpp"

namespace gslnls {
thread_local std::string g_last_error;
void set_error(const std::string &s) { g_last_error = s; }
} // namespace gslnls
using namespace gslnls;

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                            \
            return GSLNLS_ECUDA;                                                                       \
        }                                                                                              \
    } while (0)

struct gslnls_problem {
    const gslnls_model *model = nullptr;
    int p = 0, nvar = 0, has_w = 0, device = 0;
    int64_t n = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    // optional per-pass timing (bench.py roofline): event pairs around K1 launches
    bool profile = false;
    int prof_stride = 1;          // time every prof_stride-th pass launch (event records between launches cost a few us)
    int64_t prof_seen = 0;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    int *d_prof_flags = nullptr; // one word per timed launch: did it stream, or was it an idle no-op
    size_t prof_flags_cap = 0;
    // data
    std::vector<double *> owned; // buffers we allocated
    int64_t owned_cap = 0;       // rows each owned buffer can hold
    const double *dvars[NLS_MAX_VARS] = {nullptr};
    const double *dy = nullptr, *dw = nullptr;
    bool bound = false;
    unsigned long long *d_trace = nullptr; // developer hook: per-CTA phase stamps of the last pass
    int trace_cap = 0;
    double keep_mb = -1.0; // L2-resident head of the shard in MB (GSLNLS_L2_KEEP_MB; < 0: automatic)
    int wgsl = 0; // weights mode: 0 rows of J weighted too (default), 1 GSL multilarge's (f, fvv only)
    int upload_sharing = 1; // uploads running side by side in this process (one per GPU of a multi-GPU call)
    // kernels
    Variant *var = nullptr;
    VariantKey vkey{};
    int num_sms = 0, grid_x = 0, occ = 0;
    // workspace (sized for cap candidates)
    int cap = 0, cap_grid = 0, cap_trace = 0;
    int req_stride = 0, pk_stride = 0, state_stride = 0;
    double *d_gparts = nullptr;
    unsigned *d_gticket = nullptr;
    double *d_req = nullptr, *d_partials = nullptr, *d_packet = nullptr, *d_state = nullptr, *d_starts = nullptr;
    double *d_partrace = nullptr, *d_ssrtrace = nullptr, *d_condtrace = nullptr, *d_theta = nullptr;
    unsigned *d_ticket = nullptr;
    int *d_ndone = nullptr;
    int *h_ndone = nullptr; // pinned
    gslnls_comm *comm = nullptr;
    double *d_gather = nullptr;
    size_t gather_cap = 0;
    int64_t n_total = 0; // observations over all ranks
    // active fit
    trs::Params P{};
    bool active = false;
    int ncand = 1;
    std::vector<double> start;
    double h_df = 0, h_fvv = 0;
    int64_t launches = 0, passes = 0;
    int chunk = 4;
    // resident-server mode: the trust-region warp lives on its own stream for the whole fit and the
    // pass launches are sequenced on the device through the channel (nls_abi.h NLS_CH_*)
    bool allow_server = true, server_on = false;
    bool server_reduce = true; // persistent mode: the server sums the CTA partials (GSLNLS_SERVER_REDUCE=0: last CTA does)
    bool allow_persistent = true, persistent_on = false; // one pass-kernel launch per fit (GSLNLS_PERSISTENT=0 disables)
    bool server_pending = false; // fit_begin chose the server; it is launched with the first pass (fit_run)
    unsigned long long handshake_ns = 100000000ull; // start-of-fit handshake period (GSLNLS_HANDSHAKE_MS)
    char *own_channel = nullptr; // single-GPU problems own their channel; sharded ones use the comm's
    cudaStream_t srv_stream = nullptr, ctl_stream = nullptr;
    cudaEvent_t ev_reset = nullptr, ev_chunk[2] = {nullptr, nullptr};
    int *h_flags = nullptr, *d_flags = nullptr; // mapped pinned: [0] done (1) / watchdog (2)
    double *h_state = nullptr, *d_hstate = nullptr; // mapped pinned: the server's final state record
    int h_state_cap = 0;
    unsigned long long watchdog_ns = 60000000000ull;
};

static char *channel_of(const gslnls_problem *pb)
{
    return (pb->comm && pb->comm->nranks > 1) ? pb->comm->channel : pb->own_channel;
}

static void free_workspace(gslnls_problem *pb)
{
    cudaFree(pb->d_req); cudaFree(pb->d_partials); cudaFree(pb->d_packet); cudaFree(pb->d_state);
    cudaFree(pb->d_starts); cudaFree(pb->d_ticket); cudaFree(pb->d_ndone);
    cudaFree(pb->d_gparts); cudaFree(pb->d_gticket);
    pb->d_gparts = nullptr; pb->d_gticket = nullptr;
    pb->d_req = pb->d_partials = pb->d_packet = pb->d_state = pb->d_starts = nullptr;
    pb->d_ticket = nullptr; pb->d_ndone = nullptr;
    pb->cap = 0;
}

static int ensure_kernels(gslnls_problem *pb, bool batch, bool persistent = false)
{
    int vec = 2;
    for (int k = 0; k < pb->nvar; ++k)
        if (reinterpret_cast<uintptr_t>(pb->dvars[k]) & 15u)
            vec = 1;
    if ((reinterpret_cast<uintptr_t>(pb->dy) & 15u) || (pb->dw && (reinterpret_cast<uintptr_t>(pb->dw) & 15u)))
        vec = 1;
    KernelTune t = default_tune(pb->p, batch ? 0.0 : 8.0 * (pb->nvar + 1 + pb->has_w) * (double)pb->n,
                                persistent && !batch && vec == 2);
    if (t.tiled == 2 && (vec != 2 || batch)) { // bulk copies need 16-byte aligned columns; fall back to LDG
        t = default_tune(pb->p, 0.0);
        if (t.tiled == 2) { // forced through GSLNLS_TUNE
            t.tiled = 0;
            t.block = std::max(32, t.block - 32);
        }
    }
    if (t.tiled == 2) {
        // the ring must fit one SM's shared memory whatever the number of columns: fewer stages first,
        // then shorter tiles, and the LDG path when even two stages of the shortest tile do not fit
        const int narr = pb->nvar + 1 + pb->has_w;
        const size_t budget = 200 * 1024;
        while (tma_smem_bytes(narr, t.block, t.unroll, t.stages) > budget && t.stages > 3)
            --t.stages;
        while (tma_smem_bytes(narr, t.block, t.unroll, t.stages) > budget && t.unroll > 1)
            --t.unroll;
        while (tma_smem_bytes(narr, t.block, t.unroll, t.stages) > budget && t.stages > 2)
            --t.stages;
        if (tma_smem_bytes(narr, t.block, t.unroll, t.stages) > budget) {
            t = default_tune(pb->p, 0.0);
            if (t.tiled == 2) {
                t.tiled = 0;
                t.block = std::max(32, t.block - 32);
            }
        }
    }
    VariantKey key{pb->has_w, vec, batch ? 0 : 1, t.block, t.unroll, t.minb, batch ? 0 : t.tiled, t.stages, batch ? 0 : t.prefetch, t.fexp,
                   pb->has_w ? pb->wgsl : 0};
    if (pb->var && !(key < pb->vkey) && !(pb->vkey < key))
        return GSLNLS_SUCCESS;
    try {
        pb->var = &const_cast<gslnls_model *>(pb->model)->load(key);
    } catch (const std::exception &e) {
        set_error(e.what());
        return GSLNLS_ECOMPILE;
    }
    pb->vkey = key;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)pb->var->pass, key.block, pb->var->pass_smem));
    pb->occ = std::max(occ, 1);
    return GSLNLS_SUCCESS;
}

// grid size along x for one candidate
static int pick_grid(const gslnls_problem *pb, int ncand, bool reserve_slot = false)
{
    int64_t need;
    if (pb->vkey.tiled == 2) { // tiles of (consumer threads x 2 x unroll) observations go to CTAs round-robin
        const int64_t tile = (int64_t)(pb->vkey.block - 32) * 2 * pb->vkey.unroll;
        need = std::max<int64_t>(1, pb->n / tile);
    } else if (pb->vkey.tiled) { // a producer warp takes slabs of 32 observations
        const int64_t nslab = (pb->n + 31) / 32, nw = pb->vkey.block / 32;
        const int64_t np = nw / (1 + pb->vkey.unroll) * pb->vkey.unroll;
        need = std::max<int64_t>(1, (nslab + np - 1) / np);
    } else {
        const int64_t nv = pb->vkey.vec == 2 ? pb->n / 2 : pb->n;
        need = std::max<int64_t>(1, (nv + pb->vkey.block - 1) / pb->vkey.block);
    }
    int64_t full = (int64_t)pb->num_sms * pb->occ;
    if (ncand > 1) // candidates fill the machine together
        full = std::max<int64_t>(1, full / std::min<int64_t>(ncand, full));
    if (reserve_slot && full > 1)
        --full; // one CTA slot stays free for the resident trust-region warp (trs_server)
    return (int)std::min<int64_t>(need, full);
}

static int ensure_workspace(gslnls_problem *pb, int ncand, int grid_x, int ntrace)
{
    const int p = pb->p;
    pb->req_stride = trs::request_doubles(p);
    pb->pk_stride = trs::packet_doubles(p) + 1;
    pb->state_stride = trs::state_doubles(p);
    if (ncand > pb->cap || grid_x > pb->cap_grid) {
        free_workspace(pb);
        const size_t c = (size_t)ncand;
        CK(cudaMalloc(&pb->d_req, sizeof(double) * c * pb->req_stride));
        CK(cudaMalloc(&pb->d_partials, sizeof(double) * c * grid_x * pb->pk_stride));
        CK(cudaMalloc(&pb->d_packet, sizeof(double) * c * pb->pk_stride));
        CK(cudaMalloc(&pb->d_state, sizeof(double) * c * pb->state_stride));
        CK(cudaMalloc(&pb->d_starts, sizeof(double) * c * p));
        CK(cudaMalloc(&pb->d_ticket, sizeof(unsigned) * c));
        CK(cudaMalloc(&pb->d_ndone, sizeof(int)));
        const size_t ngrp = ((size_t)grid_x + NLS_RED_GROUP - 1) / NLS_RED_GROUP;
        CK(cudaMalloc(&pb->d_gparts, sizeof(double) * ngrp * pb->pk_stride));
        CK(cudaMalloc(&pb->d_gticket, sizeof(unsigned) * ngrp));
        CK(cudaMemsetAsync(pb->d_gticket, 0, sizeof(unsigned) * ngrp, pb->stream));
        CK(cudaMemsetAsync(pb->d_ticket, 0, sizeof(unsigned) * c, pb->stream));
        CK(cudaMemsetAsync(pb->d_packet, 0, sizeof(double) * c * pb->pk_stride, pb->stream));
        CK(cudaMemsetAsync(pb->d_req, 0, sizeof(double) * c * pb->req_stride, pb->stream));
        pb->cap = ncand;
        pb->cap_grid = grid_x;
    }
    if (ntrace > pb->cap_trace) {
        cudaFree(pb->d_partrace); cudaFree(pb->d_ssrtrace); cudaFree(pb->d_condtrace);
        CK(cudaMalloc(&pb->d_partrace, sizeof(double) * (size_t)ntrace * p));
        CK(cudaMalloc(&pb->d_ssrtrace, sizeof(double) * ntrace));
        CK(cudaMalloc(&pb->d_condtrace, sizeof(double) * ntrace));
        pb->cap_trace = ntrace;
    }
    if (!pb->d_theta)
        CK(cudaMalloc(&pb->d_theta, sizeof(double) * 2 * p));
    return GSLNLS_SUCCESS;
}

// how much of a shard to hold in L2 across passes: nothing is gained once the pass is many times the cache
static double default_keep_mb(double pass_bytes)
{
    (void)pass_bytes;
    return 0.0; // set from measurements, see DESIGN.md
}

static void fill_pass_params(gslnls_problem *pb, int ncand, int force_mode, NlsPassParams &prm);

// persistent mode: does the resident server sum the CTA partials itself?  (its lanes hold a 16 x 5 block each)
static bool use_server_reduce(const gslnls_problem *pb)
{
    return pb->persistent_on && pb->server_reduce && pb->grid_x <= 160 && pb->pk_stride - 1 <= 16;
}

static int launch_pass(gslnls_problem *pb, int ncand, int force_mode)
{
    NlsPassParams prm;
    fill_pass_params(pb, ncand, force_mode, prm);
    const bool timed = pb->profile && pb->prof_used + 2 <= pb->prof_ev.size() &&
                       (pb->prof_seen++ % std::max(pb->prof_stride, 1)) == 0;
    if (timed)
        prm.prof_flag = pb->d_prof_flags + pb->prof_used / 2;
    void *args[] = {&prm};
    if (timed)
        CK(cudaEventRecord(pb->prof_ev[pb->prof_used], pb->stream));
    CK(cudaLaunchKernel((const void *)pb->var->pass, dim3(pb->grid_x, ncand, 1), dim3(pb->vkey.block, 1, 1), args,
                        pb->var->pass_smem, pb->stream));
    if (timed) {
        CK(cudaEventRecord(pb->prof_ev[pb->prof_used + 1], pb->stream));
        pb->prof_used += 2;
    }
    ++pb->launches;
    ++pb->passes;
    return GSLNLS_SUCCESS;
}

// one launch for (up to) max_passes passes of the current fit: the persistent pass kernel (resident-server
// mode, TMA-ring variant).  Profiling: one event pair around the launch, the kernel counts its real passes.
static int launch_persistent(gslnls_problem *pb, int max_passes)
{
    NlsPassParams prm;
    fill_pass_params(pb, 1, 0, prm);
    prm.max_passes = max_passes;
    prm.server_reduce = use_server_reduce(pb) ? 1 : 0;
    static const int l2_ahead = std::getenv("GSLNLS_L2_AHEAD") ? std::atoi(std::getenv("GSLNLS_L2_AHEAD")) : 8;
    prm.l2_ahead = std::max(0, l2_ahead);
    const bool timed = pb->profile && pb->prof_used + 2 <= pb->prof_ev.size();
    if (timed)
        prm.prof_flag = pb->d_prof_flags + pb->prof_used / 2;
    void *args[] = {&prm};
    if (timed)
        CK(cudaEventRecord(pb->prof_ev[pb->prof_used], pb->stream));
    // Every CTA of this grid has to be resident at once (the last CTA to arrive reduces, all wait for the next
    // request): the grid is sized to one CTA per SM, which the occupancy query of prepare() guarantees to fit.
    // A cooperative launch would assert the same, but the driver does not run a cooperative grid next to the
    // resident trust-region kernel (measured on B200: it waits for the server to leave -- the start-of-fit
    // handshake then times out and the fit falls back to launch-ordered stepping).
    CK(cudaLaunchKernel((const void *)pb->var->persistent, dim3(pb->grid_x, 1, 1), dim3(pb->vkey.block, 1, 1), args,
                        pb->var->pass_smem, pb->stream));
    if (timed) {
        CK(cudaEventRecord(pb->prof_ev[pb->prof_used + 1], pb->stream));
        pb->prof_used += 2;
    }
    ++pb->launches;
    pb->passes += max_passes;
    return GSLNLS_SUCCESS;
}

static void fill_pass_params(gslnls_problem *pb, int ncand, int force_mode, NlsPassParams &prm)
{
    std::memset(&prm, 0, sizeof(prm));
    for (int k = 0; k < pb->nvar; ++k)
        prm.vars[k] = pb->dvars[k];
    prm.y = pb->dy;
    prm.w = pb->dw;
    prm.n = pb->n;
    prm.req = pb->d_req;
    prm.partials = pb->d_partials;
    prm.packet = pb->d_packet;
    prm.ticket = pb->d_ticket;
    prm.h_df = pb->h_df;
    prm.h_fvv = pb->h_fvv;
    prm.req_stride = pb->req_stride;
    prm.pk_stride = pb->pk_stride;
    prm.force_mode = force_mode;
    prm.watchdog_ns = pb->watchdog_ns;
    prm.trace = (pb->d_trace && pb->grid_x <= pb->trace_cap) ? pb->d_trace : nullptr;
    {
        // L2 residency: the first keep_mb of the pass's bytes are loaded evict_last (nls_pass_kernel.cuh)
        const double row_bytes = 8.0 * (pb->nvar + 1 + pb->has_w);
        const double keep = pb->keep_mb >= 0.0 ? pb->keep_mb : default_keep_mb(row_bytes * (double)pb->n);
        prm.keep_rows = ncand == 1 ? (long long)std::min((double)pb->n, keep * 1048576.0 / row_bytes) : 0;
    }
    static const bool flat_red = std::getenv("GSLNLS_FLAT_RED") != nullptr; // developer aid
    // two-level grid reduction for long packets (p > 8); a short packet is summed faster by one CTA
    if (ncand == 1 && pb->grid_x > NLS_RED_GROUP && pb->pk_stride >= 48 && !flat_red) {
        prm.group_partials = pb->d_gparts;
        prm.group_ticket = pb->d_gticket;
    }
    prm.nranks = 1;
    if (pb->server_on) {
        prm.channel = channel_of(pb);
        prm.peer_channel[0] = prm.channel;
        if (pb->comm && pb->comm->nranks > 1) {
            prm.nranks = pb->comm->nranks;
            prm.rank = pb->comm->rank;
            for (int r = 0; r < prm.nranks; ++r)
                prm.peer_channel[r] = pb->comm->peer_channel[r];
        }
    }
}

static int exchange_packet(gslnls_problem *pb, size_t count)
{
    if (!pb->comm || pb->comm->nranks <= 1)
        return GSLNLS_SUCCESS;
    const int R = pb->comm->nranks;
    if (pb->gather_cap < count * R) {
        cudaFree(pb->d_gather);
        CK(cudaMalloc(&pb->d_gather, sizeof(double) * count * R));
        pb->gather_cap = count * R;
    }
    int rc = comm_allgather(pb->comm, pb->d_packet, pb->d_gather, count, pb->stream);
    if (rc)
        return rc;
    CK(launch_sum_rank_packets(pb->d_gather, pb->d_packet, (int)count, R, pb->stream));
    ++pb->launches;
    return GSLNLS_SUCCESS;
}

// make the resident trust-region warp leave (no-op when it already has) and wait for it
static void stop_server(gslnls_problem *pb)
{
    if (!pb->server_on)
        return;
    if (pb->server_pending) { // chosen by fit_begin but never launched (no fit_run in between)
        pb->server_pending = false;
        pb->server_on = false;
        return;
    }
    char *ch = channel_of(pb);
    if (pb->h_flags[0] == 0 && ch) {
        static const unsigned long long one = 1ull;
        cudaMemcpyAsync(ch + NLS_CH_ABORT, &one, sizeof(one), cudaMemcpyHostToDevice, pb->ctl_stream);
        cudaStreamSynchronize(pb->ctl_stream);
    }
    cudaStreamSynchronize(pb->srv_stream);
    pb->server_on = false;
}

// block until the server has published the request that follows the last completed pass (or has
// finished the fit); the pass launches themselves must already have drained from the main stream
static int wait_server_caught_up(gslnls_problem *pb)
{
    unsigned long long *w = reinterpret_cast<unsigned long long *>(pb->h_ndone); // pinned, 16 bytes
    char *ch = channel_of(pb);
    if (pb->server_pending)
        return GSLNLS_SUCCESS; // never launched: nothing to catch up with
    const auto t0 = std::chrono::steady_clock::now();
    const double limit_s = 1e-9 * (double)pb->watchdog_ns + 1.0;
    const bool counted = use_server_reduce(pb); // the server, not the pass kernel, advances PASS_CTR
    while (pb->h_flags[0] == 0) {
        CK(cudaMemcpyAsync(&w[0], ch + NLS_CH_REQ_SEQ, sizeof(w[0]), cudaMemcpyDeviceToHost, pb->ctl_stream));
        CK(cudaMemcpyAsync(&w[1], ch + (counted ? NLS_CH_CTA_COUNT : NLS_CH_PASS_CTR), sizeof(w[1]),
                           cudaMemcpyDeviceToHost, pb->ctl_stream));
        if (counted)
            CK(cudaMemcpyAsync(&w[2], ch + NLS_CH_FIT_SEQ0, sizeof(w[2]), cudaMemcpyDeviceToHost, pb->ctl_stream));
        CK(cudaStreamSynchronize(pb->ctl_stream));
        // passes whose partials are all in: the request that follows the last of them must be out
        const unsigned long long need = counted ? w[2] + w[1] / (unsigned long long)std::max(pb->grid_x, 1) : w[1] + 1ull;
        if (w[0] >= need)
            return GSLNLS_SUCCESS;
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) {
            set_error("the resident trust-region server did not digest the last pass within the watchdog period");
            return GSLNLS_ECOMM;
        }
    }
    return GSLNLS_SUCCESS;
}

// ---- one-shot call cache -------------------------------------------------------------------
// gslnls_fit_large() is called once per fit from the host language, so everything that does not depend on
// the data values -- the n-sized device buffers, the solver workspace, streams, events, pinned words, the
// loaded kernels -- is kept from one call to the next (one problem per device).  Measured on B200 at
// n = 1e8: cudaMalloc/cudaFree of the 1.6 GB of columns plus the pinned allocations cost 5-90 ms per
// call against 30 ms of PCIe copy and 4 ms of fit.  gslnls_cache_clear() returns the memory.
namespace {
std::mutex g_cache_mu;
std::vector<gslnls_problem *> g_cache;
bool cache_enabled()
{
    const char *c = std::getenv("GSLNLS_CACHE");
    return !(c && std::atoi(c) == 0);
}
gslnls_problem *cache_take(const gslnls_model *m, int has_w, int device)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = 0; i < g_cache.size(); ++i) {
        gslnls_problem *pb = g_cache[i];
        if (pb->model == m && pb->has_w == has_w && pb->device == device) {
            g_cache.erase(g_cache.begin() + (long)i);
            return pb;
        }
    }
    return nullptr;
}
} // namespace

extern "C" void gslnls_problem_free(gslnls_problem *pb);

static void cache_put(gslnls_problem *pb)
{
    std::vector<gslnls_problem *> evict;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_cache.size();) {
            if (g_cache[i]->device == pb->device) {
                evict.push_back(g_cache[i]);
                g_cache.erase(g_cache.begin() + (long)i);
            } else {
                ++i;
            }
        }
        g_cache.push_back(pb);
    }
    for (gslnls_problem *e : evict)
        gslnls_problem_free(e);
}

static void multi_cache_drop(const gslnls_model *m);

namespace gslnls {
// a model is going away (gslnls_model_free) or the user asked for the memory back (m == nullptr: all)
void cache_drop(const gslnls_model *m)
{
    multi_cache_drop(m);
    std::vector<gslnls_problem *> evict;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_cache.size();) {
            if (!m || g_cache[i]->model == m) {
                evict.push_back(g_cache[i]);
                g_cache.erase(g_cache.begin() + (long)i);
            } else {
                ++i;
            }
        }
    }
    for (gslnls_problem *e : evict)
        gslnls_problem_free(e);
}
} // namespace gslnls

// The resident trust-region server needs the pass kernel to run NEXT TO it.  CUDA does not promise
// concurrent kernels, and some environments rule them out: CUDA_LAUNCH_BLOCKING=1, and the injection
// libraries of ncu / compute-sanitizer / cuda-gdb serialise launches.  Those are detected up front; anything
// else is caught by the start-of-fit handshake (trs_server), after which the process stops using the server.
namespace {
std::atomic<bool> g_server_unsafe{false};
std::atomic<int> g_handshake_failures{0}; // consecutive start-of-fit handshakes that timed out
bool env_set(const char *name)
{
    const char *c = std::getenv(name);
    return c && *c && !(c[0] == '0' && c[1] == '\0');
}
bool serialising_environment()
{
    return env_set("CUDA_LAUNCH_BLOCKING") || env_set("CUDA_INJECTION64_PATH") || env_set("CUDA_INJECTION32_PATH") ||
           env_set("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || env_set("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") ||
           env_set("NV_SANITIZER_INJECTION_PORT_BASE");
}
} // namespace

// weights mode of problems created from now on: gslnls_set_weights_mode(), else GSLNLS_WEIGHTS_MODE=gsl|consistent
namespace {
std::atomic<int> g_weights_mode{-1};
int default_weights_mode()
{
    const int m = g_weights_mode.load();
    if (m >= 0)
        return m;
    const char *c = std::getenv("GSLNLS_WEIGHTS_MODE");
    return (c && (std::strcmp(c, "gsl") == 0 || std::strcmp(c, "1") == 0)) ? GSLNLS_WEIGHTS_GSL : GSLNLS_WEIGHTS_CONSISTENT;
}
} // namespace

// ------------------------------------------------------------------------------------ C ABI

extern "C" {

GSLNLS_API int gslnls_set_weights_mode(int mode)
{
    if (mode != GSLNLS_WEIGHTS_CONSISTENT && mode != GSLNLS_WEIGHTS_GSL)
        return GSLNLS_EINVAL;
    g_weights_mode.store(mode);
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_set_weights_mode(gslnls_problem *pb, int mode)
{
    if (!pb || (mode != GSLNLS_WEIGHTS_CONSISTENT && mode != GSLNLS_WEIGHTS_GSL))
        return GSLNLS_EINVAL;
    if (pb->wgsl != mode) {
        pb->wgsl = mode;
        pb->var = nullptr; // another kernel variant
    }
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

GSLNLS_API int gslnls_problem_create(const gslnls_model *m, int64_t n_local, int has_weights, int device,
                                     gslnls_problem **out)
{
    if (!m || !out || n_local < 0)
        return GSLNLS_EINVAL;
    *out = nullptr;
    if (gslnls_device_count() <= device) {
        set_error("no usable CUDA device (this library has no CPU path)");
        return GSLNLS_ENODEVICE;
    }
    if (m->nvar > NLS_MAX_VARS) {
        set_error("too many predictor columns");
        return GSLNLS_EINVAL;
    }
    CK(cudaSetDevice(device));
    gslnls_problem *pb = new gslnls_problem();
    pb->model = m;
    pb->p = m->p;
    pb->nvar = m->nvar;
    pb->n = n_local;
    pb->has_w = has_weights ? 1 : 0;
    pb->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    pb->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&pb->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&pb->ev0));
    CK(cudaEventCreate(&pb->ev1));
    CK(cudaEventCreate(&pb->ev2));
    CK(cudaEventCreate(&pb->ev3));
    CK(cudaMallocHost(&pb->h_ndone, sizeof(int) * 8));
    if (const char *c = std::getenv("GSLNLS_CHUNK"))
        pb->chunk = std::max(1, std::atoi(c));
    if (const char *c = std::getenv("GSLNLS_PROF_STRIDE"))
        pb->prof_stride = std::max(1, std::atoi(c));
    if (const char *c = std::getenv("GSLNLS_SERVER"))
        pb->allow_server = std::atoi(c) != 0;
    else if (serialising_environment())
        pb->allow_server = false;
    if (const char *c = std::getenv("GSLNLS_SERVER_REDUCE"))
        pb->server_reduce = std::atoi(c) != 0;
    if (const char *c = std::getenv("GSLNLS_PERSISTENT"))
        pb->allow_persistent = std::atoi(c) != 0;
    if (const char *c = std::getenv("GSLNLS_WATCHDOG_S"))
        pb->watchdog_ns = (unsigned long long)std::max(1, std::atoi(c)) * 1000000000ull;
    pb->wgsl = default_weights_mode();
    if (const char *c = std::getenv("GSLNLS_L2_KEEP_MB"))
        pb->keep_mb = std::atof(c);
    if (const char *c = std::getenv("GSLNLS_L2_PERSIST_MB")) { // developer aid: L2 set-aside for persisting (evict_last) lines
        const size_t want = (size_t)(std::atof(c) * 1048576.0);
        const size_t cap = (size_t)prop.persistingL2CacheMaxSize;
        cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(want, cap));
        if (e != cudaSuccess)
            cudaGetLastError();
        static bool told = false;
        if (!told) {
            std::fprintf(stderr, "gslnls: persisting L2 set-aside %.0f MB requested, device maximum %.0f MB (%s)\n",
                         want / 1048576.0, cap / 1048576.0, cudaGetErrorString(e));
            told = true;
        }
    }
    if (const char *c = std::getenv("GSLNLS_HANDSHAKE_MS"))
        pb->handshake_ns = (unsigned long long)std::max(1, std::atoi(c)) * 1000000ull;
    CK(cudaStreamCreateWithFlags(&pb->srv_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&pb->ctl_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&pb->ev_reset, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&pb->ev_chunk[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&pb->ev_chunk[1], cudaEventDisableTiming));
    CK(cudaHostAlloc(&pb->h_flags, sizeof(int) * 4, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&pb->d_flags, pb->h_flags, 0));
    pb->h_flags[0] = 0;
    *out = pb;
    return GSLNLS_SUCCESS;
}

GSLNLS_API void gslnls_problem_free(gslnls_problem *pb)
{
    if (!pb)
        return;
    cudaSetDevice(pb->device);
    stop_server(pb);
    cudaStreamSynchronize(pb->stream);
    cudaFree(pb->own_channel);
    cudaFree(pb->d_trace);
    cudaFree(pb->d_prof_flags);
    cudaFreeHost(pb->h_flags);
    cudaFreeHost(pb->h_state);
    cudaEventDestroy(pb->ev_reset);
    cudaEventDestroy(pb->ev_chunk[0]);
    cudaEventDestroy(pb->ev_chunk[1]);
    cudaStreamDestroy(pb->srv_stream);
    cudaStreamDestroy(pb->ctl_stream);
    for (double *b : pb->owned)
        cudaFree(b);
    free_workspace(pb);
    cudaFree(pb->d_partrace); cudaFree(pb->d_ssrtrace); cudaFree(pb->d_condtrace); cudaFree(pb->d_theta);
    cudaFree(pb->d_gather);
    cudaFreeHost(pb->h_ndone);
    cudaEventDestroy(pb->ev0);
    cudaEventDestroy(pb->ev1);
    cudaEventDestroy(pb->ev2);
    cudaEventDestroy(pb->ev3);
    for (cudaEvent_t e : pb->prof_ev)
        cudaEventDestroy(e);
    cudaStreamDestroy(pb->stream);
    delete pb;
}

GSLNLS_API int gslnls_problem_upload(gslnls_problem *pb, const double *const *vars, const double *y,
                                     const double *weights)
{
    if (!pb || !y || (pb->nvar > 0 && !vars) || (pb->has_w && !weights))
        return GSLNLS_EINVAL;
    CK(cudaSetDevice(pb->device));
    const int nbuf = pb->nvar + 1 + pb->has_w;
    if (pb->bound || (int)pb->owned.size() != nbuf || pb->owned_cap < pb->n) {
        for (double *b : pb->owned)
            cudaFree(b);
        pb->owned.clear();
        pb->owned_cap = 0;
        const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(pb->n, 1);
        for (int k = 0; k < nbuf; ++k) {
            double *d = nullptr;
            CK(cudaMalloc(&d, bytes));
            pb->owned.push_back(d);
        }
        pb->owned_cap = std::max<int64_t>(pb->n, 1);
        pb->bound = false;
    }
    // pageable host memory (what R hands over) is pinned by the library: multi-threaded staging ring, upload.cpp
    const void *src[NLS_MAX_VARS + 2];
    void *dst[NLS_MAX_VARS + 2];
    for (int k = 0; k < pb->nvar; ++k) {
        src[k] = vars[k];
        dst[k] = pb->owned[k];
        pb->dvars[k] = pb->owned[k];
    }
    src[pb->nvar] = y;
    dst[pb->nvar] = pb->owned[pb->nvar];
    pb->dy = pb->owned[pb->nvar];
    pb->dw = nullptr;
    if (pb->has_w) {
        src[pb->nvar + 1] = weights;
        dst[pb->nvar + 1] = pb->owned[pb->nvar + 1];
        pb->dw = pb->owned[pb->nvar + 1];
    }
    if (pb->n > 0) {
        int rc = staged_upload(pb->device, src, dst, nbuf, sizeof(double) * (size_t)pb->n, pb->stream,
                               upload_threads_default(pb->upload_sharing));
        if (rc)
            return rc;
    }
    pb->var = nullptr;
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_bind_device(gslnls_problem *pb, const double *const *dev_vars, const double *dev_y,
                                          const double *dev_weights)
{
    if (!pb || !dev_y || (pb->nvar > 0 && !dev_vars) || (pb->has_w && !dev_weights))
        return GSLNLS_EINVAL;
    for (int k = 0; k < pb->nvar; ++k)
        pb->dvars[k] = dev_vars[k];
    pb->dy = dev_y;
    pb->dw = pb->has_w ? dev_weights : nullptr;
    pb->bound = true;
    pb->var = nullptr;
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_set_comm(gslnls_problem *pb, gslnls_comm *comm)
{
    if (!pb)
        return GSLNLS_EINVAL;
    pb->comm = comm;
    pb->n_total = pb->n;
    if (comm && comm->nranks > 1 && comm->n_total_hint >= 0) {
        pb->n_total = comm->n_total_hint; // single-process group: the caller split the rows itself
    } else if (comm && comm->nranks > 1) {
        // global number of observations: gather the shard sizes once
        CK(cudaSetDevice(pb->device));
        const int R = comm->nranks;
        double *d = nullptr;
        CK(cudaMalloc(&d, sizeof(double) * (R + 1)));
        const double mine = (double)pb->n;
        CK(cudaMemcpyAsync(d + R, &mine, sizeof(double), cudaMemcpyHostToDevice, pb->stream));
        int rc = comm_allgather(comm, d + R, d, 1, pb->stream);
        if (rc) {
            cudaFree(d);
            return rc;
        }
        std::vector<double> h(R);
        CK(cudaMemcpyAsync(h.data(), d, sizeof(double) * R, cudaMemcpyDeviceToHost, pb->stream));
        CK(cudaStreamSynchronize(pb->stream));
        cudaFree(d);
        pb->n_total = 0;
        for (double v : h)
            pb->n_total += (int64_t)v;
    }
    return GSLNLS_SUCCESS;
}

static int prepare(gslnls_problem *pb, int ncand, int ntrace, bool batch, bool reserve_slot = false,
                   bool persistent = false)
{
    if (!pb->dy) {
        set_error("no data: call gslnls_problem_upload or gslnls_problem_bind_device first");
        return GSLNLS_EINVAL;
    }
    CK(cudaSetDevice(pb->device));
    int rc = ensure_kernels(pb, batch, persistent);
    if (rc)
        return rc;
    pb->persistent_on = persistent && pb->vkey.tiled == 2 && pb->var->persistent != nullptr;
    // The resident trust-region warp gets an SM slot of its own.  Registers and shared memory of a TMA-ring CTA
    // would leave room for it on the same SM, but an SM that is running the server (tiny shared-memory carve-out)
    // cannot be re-configured for a 150 KB CTA until it drains, which the server never does: a 148-CTA
    // persistent grid then waits forever for its last CTA (measured: watchdog).  The pass is HBM-bound; 147
    // streaming SMs lose nothing.
    pb->grid_x = pick_grid(pb, ncand, reserve_slot);
    rc = ensure_workspace(pb, ncand, pb->grid_x, ntrace);
    return rc;
}

GSLNLS_API int gslnls_problem_eval_packet(gslnls_problem *pb, const double *theta, double *packet)
{
    if (!pb || !theta || !packet)
        return GSLNLS_EINVAL;
    int rc = prepare(pb, 1, 0, false);
    if (rc)
        return rc;
    const int p = pb->p;
    pb->h_df = pb->h_df > 0 ? pb->h_df : 1.4901161193847656e-08;
    CK(cudaMemcpyAsync(pb->d_theta, theta, sizeof(double) * p, cudaMemcpyHostToDevice, pb->stream));
    CK(trs_launch_set_request(pb->d_req, trs::MODE_FJ, pb->d_theta, nullptr, p, pb->stream));
    rc = launch_pass(pb, 1, trs::MODE_FJ);
    if (rc)
        return rc;
    rc = exchange_packet(pb, (size_t)pb->pk_stride);
    if (rc)
        return rc;
    CK(cudaMemcpyAsync(packet, pb->d_packet, sizeof(double) * trs::packet_doubles(p), cudaMemcpyDeviceToHost,
                       pb->stream));
    CK(cudaStreamSynchronize(pb->stream));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_eval_jtfvv(gslnls_problem *pb, const double *theta, const double *v, double *out)
{
    if (!pb || !theta || !v || !out)
        return GSLNLS_EINVAL;
    int rc = prepare(pb, 1, 0, false);
    if (rc)
        return rc;
    const int p = pb->p;
    pb->h_df = pb->h_df > 0 ? pb->h_df : 1.4901161193847656e-08;
    pb->h_fvv = pb->h_fvv > 0 ? pb->h_fvv : 0.02;
    CK(cudaMemcpyAsync(pb->d_theta, theta, sizeof(double) * p, cudaMemcpyHostToDevice, pb->stream));
    CK(cudaMemcpyAsync(pb->d_theta + p, v, sizeof(double) * p, cudaMemcpyHostToDevice, pb->stream));
    CK(trs_launch_set_request(pb->d_req, trs::MODE_FVV, pb->d_theta, pb->d_theta + p, p, pb->stream));
    rc = launch_pass(pb, 1, trs::MODE_FVV);
    if (rc)
        return rc;
    rc = exchange_packet(pb, p + 1);
    if (rc)
        return rc;
    CK(cudaMemcpyAsync(out, pb->d_packet, sizeof(double) * p, cudaMemcpyDeviceToHost, pb->stream));
    CK(cudaStreamSynchronize(pb->stream));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_time_passes(gslnls_problem *pb, const double *theta, int npass, float *ms_per_pass)
{
    if (!pb || !theta || npass < 1 || !ms_per_pass)
        return GSLNLS_EINVAL;
    int rc = prepare(pb, 1, 0, false);
    if (rc)
        return rc;
    const int p = pb->p;
    pb->h_df = pb->h_df > 0 ? pb->h_df : 1.4901161193847656e-08;
    CK(cudaMemcpyAsync(pb->d_theta, theta, sizeof(double) * p, cudaMemcpyHostToDevice, pb->stream));
    CK(trs_launch_set_request(pb->d_req, trs::MODE_FJ, pb->d_theta, nullptr, p, pb->stream));
    CK(cudaEventRecord(pb->ev0, pb->stream));
    for (int i = 0; i < npass; ++i) {
        rc = launch_pass(pb, 1, trs::MODE_FJ);
        if (rc)
            return rc;
    }
    CK(cudaEventRecord(pb->ev1, pb->stream));
    CK(cudaEventSynchronize(pb->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, pb->ev0, pb->ev1));
    *ms_per_pass = ms / (float)npass;
    return GSLNLS_SUCCESS;
}

static int problem_residuals_ld(gslnls_problem *pb, const double *theta, double *resid, double *grad, int64_t ld);

GSLNLS_API int gslnls_problem_residuals(gslnls_problem *pb, const double *theta, double *resid, double *grad)
{
    return problem_residuals_ld(pb, theta, resid, grad, pb ? pb->n : 0);
}

// K4 with the host Jacobian's leading dimension given: a shard of a multi-GPU session writes its rows into
// the caller's n x p column-major array directly (column j of this shard starts at grad + j * ld)
static int problem_residuals_ld(gslnls_problem *pb, const double *theta, double *resid, double *grad, int64_t ld)
{
    if (!pb || !theta)
        return GSLNLS_EINVAL;
    int rc = prepare(pb, 1, 0, false);
    if (rc)
        return rc;
    const int p = pb->p;
    double *d_resid = nullptr, *d_grad = nullptr;
    struct Guard { // the scratch arrays go back on every exit path
        double *&a, *&b;
        ~Guard() { cudaFree(a); cudaFree(b); }
    } guard{d_resid, d_grad};
    const size_t n1 = (size_t)std::max<int64_t>(pb->n, 1);
    if (resid)
        CK(cudaMalloc(&d_resid, sizeof(double) * n1));
    if (grad)
        CK(cudaMalloc(&d_grad, sizeof(double) * n1 * p));
    CK(cudaMemcpyAsync(pb->d_theta, theta, sizeof(double) * p, cudaMemcpyHostToDevice, pb->stream));
    NlsMaterialiseParams prm;
    std::memset(&prm, 0, sizeof(prm));
    for (int k = 0; k < pb->nvar; ++k)
        prm.vars[k] = pb->dvars[k];
    prm.y = pb->dy;
    prm.w = pb->dw;
    prm.n = pb->n;
    prm.theta = pb->d_theta;
    prm.resid = d_resid;
    prm.grad = d_grad;
    prm.h_df = pb->h_df > 0 ? pb->h_df : 1.4901161193847656e-08;
    void *args[] = {&prm};
    const int blocks = (int)std::min<int64_t>((pb->n + 255) / 256 > 0 ? (pb->n + 255) / 256 : 1, (int64_t)pb->num_sms * 8);
    CK(cudaLaunchKernel((const void *)pb->var->materialise, dim3(blocks), dim3(256), args, 0, pb->stream));
    ++pb->launches;
    if (resid)
        CK(cudaMemcpyAsync(resid, d_resid, sizeof(double) * pb->n, cudaMemcpyDeviceToHost, pb->stream));
    if (grad) {
        if (ld == pb->n) {
            CK(cudaMemcpyAsync(grad, d_grad, sizeof(double) * pb->n * p, cudaMemcpyDeviceToHost, pb->stream));
        } else {
            for (int j = 0; j < p; ++j)
                CK(cudaMemcpyAsync(grad + (size_t)j * ld, d_grad + (size_t)j * pb->n, sizeof(double) * pb->n,
                                   cudaMemcpyDeviceToHost, pb->stream));
        }
    }
    CK(cudaStreamSynchronize(pb->stream));
    return GSLNLS_SUCCESS;
}

static int fill_params(gslnls_problem *pb, const int *ci, const double *cd, int batch_iters)
{
    trs::Params &P = pb->P;
    P.p = pb->p;
    P.maxiter = ci[0];
    P.trace = ci[1] ? 1 : 0;
    P.trs = (ci[2] >= 1 && ci[2] <= 5) ? ci[2] : 0;       // src/nls_large.c:97-116
    P.scale = (ci[3] == 1 || ci[3] == 2) ? ci[3] : 0;     // :119-129
    P.batch_iters = batch_iters;
    const int64_t ntot = (pb->comm && pb->comm->nranks > 1) ? pb->n_total : pb->n;
    P.cg_maxit = std::max<int64_t>(ntot, 1);             // GSL default max_iter = 0 -> n
    P.factor_up = cd[0]; P.factor_down = cd[1]; P.avmax = cd[2]; P.h_df = cd[3]; P.h_fvv = cd[4]; // :135-139
    P.xtol = cd[5]; P.ftol = cd[6]; P.gtol = cd[7];
    P.cg_tol = 1.0e-6;                                     // GSL default tol
    pb->h_df = cd[3];
    pb->h_fvv = cd[4];
    if (P.maxiter < 1) {
        set_error("maxiter must be >= 1");
        return GSLNLS_EINVAL;
    }
    if (P.p > trs_max_p()) {
        set_error("p exceeds the dense trust-region kernel limit (100)");
        return GSLNLS_EINVAL;
    }
    if (P.trs == trs::TRS_LMACCEL && pb->model->spec.fvv_mode == GSLNLS_FVV_NONE) {
        // R/nls_large.R:354-356
        set_error("analytic second derivative function 'fvv' is required, but none is available");
        return GSLNLS_EINVAL;
    }
    return GSLNLS_SUCCESS;
}

// device-side start of a fit with the parameters already in pb->P / pb->start: records, channel, traces.
// Called by fit_begin, and again by fit_run when the resident server had to be abandoned.
static int fit_setup(gslnls_problem *pb)
{
    const int ntrace = pb->P.trace ? pb->P.maxiter + 1 : 0;
    stop_server(pb);
    const bool sharded = pb->comm && pb->comm->nranks > 1;
    const bool use_server = pb->allow_server && !g_server_unsafe.load() && pb->p <= trs_server_max_p() &&
                            trs::packet_doubles(pb->p) + 1 <= NLS_CH_MAXPK && (!sharded || pb->comm->p2p);
    const bool want_persistent = use_server && pb->allow_persistent && pb->p <= 4;
    int rc = prepare(pb, 1, ntrace, false, use_server, want_persistent);
    if (rc)
        return rc;
    if (use_server && pb->h_state_cap < pb->state_stride) {
        cudaFreeHost(pb->h_state);
        pb->h_state = nullptr;
        CK(cudaHostAlloc(&pb->h_state, sizeof(double) * pb->state_stride, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(&pb->d_hstate, pb->h_state, 0));
        pb->h_state_cap = pb->state_stride;
    }
    if (use_server && !sharded && !pb->own_channel) {
        CK(cudaMalloc(&pb->own_channel, NLS_CH_BYTES));
        CK(cudaMemsetAsync(pb->own_channel, 0, NLS_CH_BYTES, pb->stream));
    }
    const int p = pb->p;
    pb->ncand = 1;
    if (ntrace) {
        CK(cudaMemsetAsync(pb->d_partrace, 0, sizeof(double) * (size_t)ntrace * p, pb->stream));
        CK(cudaMemsetAsync(pb->d_ssrtrace, 0, sizeof(double) * ntrace, pb->stream));
        CK(cudaMemsetAsync(pb->d_condtrace, 0, sizeof(double) * ntrace, pb->stream));
    }
    if (use_server) {
        // one launch writes the records (start values travel as kernel arguments) and opens the channel; the
        // pass launches that follow on the main stream find their request through the channel, not through
        // stream order; the server kernel itself goes out with the first pass (fit_run)
        CK(trs_launch_fit_begin(pb->d_state, pb->d_req, pb->start.data(), p, pb->d_ndone, channel_of(pb), pb->stream));
        CK(cudaEventRecord(pb->ev_reset, pb->stream));
        ++pb->launches;
        pb->h_flags[0] = 0;
        pb->server_on = true;
        pb->server_pending = true;
    } else {
        CK(cudaMemcpyAsync(pb->d_starts, pb->start.data(), sizeof(double) * p, cudaMemcpyHostToDevice, pb->stream));
        CK(trs_launch_reset(pb->d_state, pb->state_stride, pb->d_req, pb->req_stride, pb->d_starts, p, 1, pb->d_ndone,
                            pb->stream));
        ++pb->launches;
    }
    pb->active = true;
    pb->passes = 0;
    return GSLNLS_SUCCESS;
}

// launch the resident server chosen by fit_setup, right before the first pass of the fit is enqueued: its
// start-of-fit handshake expects a running pass kernel within handshake_ns
static int launch_pending_server(gslnls_problem *pb)
{
    if (!pb->server_pending)
        return GSLNLS_SUCCESS;
    const bool sharded = pb->comm && pb->comm->nranks > 1;
    const bool tr = pb->P.trace != 0;
    CK(cudaStreamWaitEvent(pb->srv_stream, pb->ev_reset, 0));
    TrsLinks links;
    if (use_server_reduce(pb)) {
        links.partials = pb->d_partials;
        links.nctas = pb->grid_x;
        links.pk_stride = pb->pk_stride;
        if (sharded) {
            links.rank = pb->comm->rank;
            for (int r = 0; r < pb->comm->nranks; ++r)
                links.peer[r] = pb->comm->peer_channel[r];
        }
    }
    CK(trs_launch_server(pb->P, channel_of(pb), sharded ? pb->comm->nranks : 1, pb->pk_stride, pb->d_state,
                         pb->d_packet, pb->d_req, tr ? pb->d_partrace : nullptr, tr ? pb->d_ssrtrace : nullptr,
                         tr ? pb->d_condtrace : nullptr, pb->d_ndone, pb->d_flags, pb->d_hstate, pb->watchdog_ns,
                         pb->handshake_ns, links, pb->srv_stream));
    ++pb->launches;
    pb->server_pending = false;
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_fit_begin(gslnls_problem *pb, const double *start, const int *control_int,
                                        const double *control_dbl)
{
    if (!pb || !start || !control_int || !control_dbl)
        return GSLNLS_EINVAL;
    int rc = fill_params(pb, control_int, control_dbl, 0);
    if (rc)
        return rc;
    pb->start.assign(start, start + pb->p);
    return fit_setup(pb);
}

GSLNLS_API int gslnls_problem_fit_run(gslnls_problem *pb, int max_passes, int *done, int64_t *passes_run,
                                      float *device_ms)
{
    if (!pb || !pb->active)
        return GSLNLS_EINVAL;
    CK(cudaSetDevice(pb->device));
    const bool tr = pb->P.trace != 0;
    int64_t run = 0;
    int fin = 0;
    CK(cudaEventRecord(pb->ev0, pb->stream));
    if (pb->server_on) {
        int rc0 = launch_pending_server(pb);
        if (rc0)
            return rc0;
        // keep two chunks of pass launches in flight; the only host work per chunk is waiting for the
        // older chunk's event and looking at the done word the server writes into mapped host memory
        int slot = 0, inflight = 0;
        fin = pb->h_flags[0] != 0;
        if (pb->persistent_on && !fin) {
            // one launch for the rest of the fit (or for max_passes of it): the grid stays resident, every CTA
            // loops over the passes; the host only watches the done word and the end of the kernel
            const int64_t cap = (int64_t)pb->P.maxiter * 34 + 8;
            const int budget = (int)(max_passes > 0 ? std::min<int64_t>(max_passes, cap) : cap);
            int rc = launch_persistent(pb, budget);
            if (rc)
                return rc;
            run = budget;
            CK(cudaEventRecord(pb->ev_chunk[0], pb->stream));
            cudaError_t q;
            while ((q = cudaEventQuery(pb->ev_chunk[0])) == cudaErrorNotReady && pb->h_flags[0] == 0) {
            }
            if (q != cudaSuccess && q != cudaErrorNotReady)
                CK(q);
            fin = pb->h_flags[0] != 0;
        }
        while (!pb->persistent_on && !fin && (max_passes <= 0 || run < max_passes)) {
            int todo = pb->chunk;
            if (max_passes > 0)
                todo = (int)std::min<int64_t>(todo, max_passes - run);
            for (int i = 0; i < todo; ++i) {
                int rc = launch_pass(pb, 1, 0);
                if (rc)
                    return rc;
            }
            run += todo;
            CK(cudaEventRecord(pb->ev_chunk[slot], pb->stream));
                    slot ^= 1;
            if (++inflight == 2) {
                // wait for the older chunk -- or for the done word, whichever comes first
                cudaError_t q;
                while ((q = cudaEventQuery(pb->ev_chunk[slot])) == cudaErrorNotReady && pb->h_flags[0] == 0) {
                }
                if (q != cudaSuccess && q != cudaErrorNotReady)
                    CK(q);
                --inflight;
            }
            fin = pb->h_flags[0] != 0;
        }
        if (pb->h_flags[0] == 3) {
            // The server never saw a pass kernel running beside it (start-of-fit handshake): kernels are being
            // serialised by a tool or by the platform.  Nothing has been consumed; what is queued drains as idle
            // launches.  Step this fit -- and every later one in this process -- launch-ordered instead.
            CK(cudaStreamSynchronize(pb->stream));
            CK(cudaStreamSynchronize(pb->srv_stream));
            pb->server_on = false;
            // one timeout can be a hiccup (the host thread descheduled between the two launches); two in a row
            // mean kernels are being serialised: stop trying in this process
            const bool give_up = g_handshake_failures.fetch_add(1) + 1 >= 2;
            if (give_up)
                g_server_unsafe.store(true);
            const bool saved_allow = pb->allow_server;
            pb->allow_server = false; // this fit restarts launch-ordered either way
            struct Restore {
                gslnls_problem *pb;
                bool v;
                ~Restore() { pb->allow_server = v; }
            } restore{pb, saved_allow};
            if (pb->comm && pb->comm->nranks > 1) {
                set_error("the resident trust-region server cannot run next to the pass kernel on this rank "
                          "(kernels are serialised); rerun every rank with GSLNLS_SERVER=0");
                return GSLNLS_ECOMM;
            }
            int rc = fit_setup(pb);
            if (rc)
                return rc;
            run = 0;
            fin = 0;
            goto launch_ordered;
        }
        if (!fin || device_ms) {
            // a finished fit needs no synchronisation: its state record is already in host memory and the
            // launches still queued are no-ops that the next fit's launches simply follow
            CK(cudaEventRecord(pb->ev1, pb->stream));
            CK(cudaEventSynchronize(pb->ev1));
        }
        if (!fin) {
            // the last pass has left the stream; its step may still be running in the server
            int rc = wait_server_caught_up(pb);
            if (rc)
                return rc;
            fin = pb->h_flags[0] != 0;
        }
        if (pb->h_flags[0] == 1)
            g_handshake_failures.store(0);
        if (pb->h_flags[0] == 2) {
            if (pb->comm && pb->comm->nranks > 1)
                set_error("trust-region server watchdog: a rank never delivered its packet");
            else
                set_error("trust-region server watchdog: the pass kernel never delivered its packet "
                          "(GSLNLS_WATCHDOG_S seconds; GSLNLS_SERVER=0 selects launch-ordered stepping)");
            return GSLNLS_ECOMM;
        }
        if (device_ms)
            CK(cudaEventElapsedTime(device_ms, pb->ev0, pb->ev1));
        if (done)
            *done = fin;
        if (passes_run)
            *passes_run = run;
        return GSLNLS_SUCCESS;
    }
launch_ordered:
    while (!fin && (max_passes <= 0 || run < max_passes)) {
        int todo = pb->chunk;
        if (max_passes > 0)
            todo = (int)std::min<int64_t>(todo, max_passes - run);
        for (int i = 0; i < todo; ++i) {
            int rc = launch_pass(pb, 1, 0);
            if (rc)
                return rc;
            rc = exchange_packet(pb, (size_t)pb->pk_stride);
            if (rc)
                return rc;
            CK(trs_launch_step(pb->P, pb->d_state, pb->d_packet, pb->d_req, tr ? pb->d_partrace : nullptr,
                               tr ? pb->d_ssrtrace : nullptr, tr ? pb->d_condtrace : nullptr, pb->d_ndone,
                               pb->stream));
            ++pb->launches;
        }
        run += todo;
        CK(cudaMemcpyAsync(pb->h_ndone, pb->d_ndone, sizeof(int), cudaMemcpyDeviceToHost, pb->stream));
        CK(cudaStreamSynchronize(pb->stream));
        fin = pb->h_ndone[0] >= 1;
    }
    CK(cudaEventRecord(pb->ev1, pb->stream));
    CK(cudaEventSynchronize(pb->ev1));
    if (device_ms)
        CK(cudaEventElapsedTime(device_ms, pb->ev0, pb->ev1));
    if (done)
        *done = fin;
    if (passes_run)
        *passes_run = run;
    return GSLNLS_SUCCESS;
}

static double *dup(const double *src, size_t n)
{
    double *d = (double *)std::malloc(sizeof(double) * (n ? n : 1));
    if (src)
        std::memcpy(d, src, sizeof(double) * n);
    return d;
}

GSLNLS_API int gslnls_problem_fit_end(gslnls_problem *pb, int want_resid_grad, gslnls_result *out)
{
    if (!pb || !out || !pb->active)
        return GSLNLS_EINVAL;
    CK(cudaSetDevice(pb->device));
    const int p = pb->p;
    std::memset(out, 0, sizeof(*out));
    std::vector<double> S(pb->state_stride);
    if (pb->server_on && pb->h_flags[0] == 1) {
        // finished fit: the server left its final state record in mapped host memory before raising the
        // done word; nothing to copy, nothing to wait for
        std::memcpy(S.data(), pb->h_state, sizeof(double) * pb->state_stride);
        pb->server_on = false;
    } else {
        if (pb->server_on) {
            // wait until the server has digested every pass that was launched (request seq = passes + 1),
            // then let it go; an unfinished fit leaves its state record current
            CK(cudaStreamSynchronize(pb->stream));
            int rc = wait_server_caught_up(pb);
            if (rc)
                return rc;
            stop_server(pb);
        }
        CK(cudaMemcpyAsync(S.data(), pb->d_state, sizeof(double) * pb->state_stride, cudaMemcpyDeviceToHost, pb->stream));
        CK(cudaStreamSynchronize(pb->stream));
    }
    pb->active = false;

    int status = (int)S[trs::S_STATUS];
    const bool finished = (int)S[trs::S_PHASE] == trs::PH_DONE;
    if (!finished)
        status = GSLNLS_CONTINUE; // fit_end before completion (benchmark use)
    const bool ok = status == GSLNLS_SUCCESS || status == GSLNLS_EMAXITER || !finished;
    out->n = (pb->comm && pb->comm->nranks > 1) ? pb->n_total : pb->n;
    out->n_local = pb->n;
    out->p = p;
    const double *v = S.data() + trs::S_COUNT;
    out->par = dup(ok ? v : pb->start.data(), p);          // src/nls_large.c:293-302
    out->covar = dup(v + 6 * p + p * p, (size_t)p * p);    // :311-326
    if (!ok || !finished)
        for (int i = 0; i < p * p; ++i)
            out->covar[i] = NAN;
    out->jtj = dup(v + 6 * p, (size_t)p * p);
    for (int i = 0; i < p; ++i) // mirror the stored lower triangle
        for (int j = i + 1; j < p; ++j)
            out->jtj[i * p + j] = out->jtj[j * p + i];
    out->grad_vec = dup(v + 2 * p, p);
    out->x_final = dup(v, p);
    out->ssr = S[trs::S_CHISQ1];
    out->ssrtol = S[trs::S_CHISQ0] - S[trs::S_CHISQ1];
    out->chisq_init = S[trs::S_CHISQ_INIT];
    out->niter = (int)S[trs::S_NITER];
    out->conv = status;
    out->info = (int)S[trs::S_INFO];
    out->status = gslnls_strerror(status);
    out->algorithm = gslnls_trs_name(pb->P.trs);
    out->neval[0] = (int64_t)S[trs::S_NEVAL_F];
    out->neval[1] = (int64_t)S[trs::S_NEVAL_DFU];
    out->neval[2] = (int64_t)S[trs::S_NEVAL_DF2];
    out->neval[3] = (int64_t)S[trs::S_NEVAL_FVV];
    out->npass = (int64_t)S[trs::S_NPASS];
    if (pb->P.trace) {
        const int nt = pb->P.maxiter + 1;
        out->ntrace = nt;
        out->partrace = dup(nullptr, (size_t)nt * p);
        out->ssrtrace = dup(nullptr, nt);
        out->condtrace = dup(nullptr, nt);
        CK(cudaMemcpy(out->partrace, pb->d_partrace, sizeof(double) * (size_t)nt * p, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out->ssrtrace, pb->d_ssrtrace, sizeof(double) * nt, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out->condtrace, pb->d_condtrace, sizeof(double) * nt, cudaMemcpyDeviceToHost));
    }
    if (want_resid_grad) {
        out->resid = dup(nullptr, (size_t)pb->n);
        out->grad = dup(nullptr, (size_t)pb->n * p);
        if (ok && finished) {
            int rc = gslnls_problem_residuals(pb, out->par, out->resid, out->grad); // :339-385
            if (rc)
                return rc;
        } else {
            for (int64_t i = 0; i < pb->n; ++i)
                out->resid[i] = NAN;
            for (int64_t i = 0; i < pb->n * p; ++i)
                out->grad[i] = NAN;
        }
    }
    return status;
}

GSLNLS_API int gslnls_problem_fit(gslnls_problem *pb, const double *start, const int *control_int,
                                  const double *control_dbl, int want_resid_grad, gslnls_result *out)
{
    int rc = gslnls_problem_fit_begin(pb, start, control_int, control_dbl);
    if (rc)
        return rc;
    int done = 0;
    // every trial step is one pass (two with geodesic acceleration); maxiter * 17 trials * 2 bounds it
    const int64_t hard_cap = (int64_t)pb->P.maxiter * 34 + 8;
    int64_t total = 0;
    while (!done && total < hard_cap) {
        int64_t run = 0;
        // launch-ordered mode returns to the host every few chunks; with the resident server one call
        // runs the whole fit
        const int64_t budget = pb->server_on ? std::min<int64_t>(hard_cap - total, 1 << 30) : pb->chunk * 4;
        rc = gslnls_problem_fit_run(pb, (int)budget, &done, &run, nullptr);
        if (rc)
            return rc;
        total += run;
    }
    return gslnls_problem_fit_end(pb, want_resid_grad, out);
}

/* developer hook: record globaltimer stamps of every CTA's phases in the pass kernel (0 entry, 1 request seen,
 * 2 thread 0 done streaming, 3 CTA done streaming, 4 partial written, 5 packet published by the last CTA);
 * read returns the stamps of the LAST pass launched, [ctas][8] nanoseconds, and the number of CTAs */
GSLNLS_API int gslnls_problem_trace(gslnls_problem *pb, int enable, unsigned long long *out, int cap_ctas, int *nctas)
{
    if (!pb)
        return GSLNLS_EINVAL;
    CK(cudaSetDevice(pb->device));
    if (enable && !pb->d_trace) {
        pb->trace_cap = 1024;
        CK(cudaMalloc(&pb->d_trace, sizeof(unsigned long long) * 32 * pb->trace_cap));
        CK(cudaMemset(pb->d_trace, 0, sizeof(unsigned long long) * 32 * pb->trace_cap));
    }
    if (out && pb->d_trace) {
        CK(cudaStreamSynchronize(pb->stream));
        const int n = std::min(std::min(cap_ctas, pb->trace_cap), pb->grid_x);
        CK(cudaMemcpy(out, pb->d_trace, sizeof(unsigned long long) * 32 * n, cudaMemcpyDeviceToHost));
        if (nctas)
            *nctas = n;
    }
    if (!enable && pb->d_trace) {
        cudaFree(pb->d_trace);
        pb->d_trace = nullptr;
    }
    return GSLNLS_SUCCESS;
}

GSLNLS_API int64_t gslnls_problem_launch_count(const gslnls_problem *pb) { return pb ? pb->launches : 0; }

GSLNLS_API int gslnls_problem_timer_start(gslnls_problem *pb)
{
    if (!pb)
        return GSLNLS_EINVAL;
    CK(cudaSetDevice(pb->device));
    CK(cudaStreamSynchronize(pb->stream));
    CK(cudaEventRecord(pb->ev2, pb->stream));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_timer_stop(gslnls_problem *pb, float *ms)
{
    if (!pb || !ms)
        return GSLNLS_EINVAL;
    CK(cudaEventRecord(pb->ev3, pb->stream));
    CK(cudaEventSynchronize(pb->ev3));
    CK(cudaEventElapsedTime(ms, pb->ev2, pb->ev3));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_set_profile(gslnls_problem *pb, int max_passes)
{
    if (!pb)
        return GSLNLS_EINVAL;
    CK(cudaSetDevice(pb->device));
    pb->profile = max_passes > 0;
    pb->prof_used = 0;
    pb->prof_seen = 0;
    while (pb->prof_ev.size() < (size_t)2 * (size_t)std::max(max_passes, 0)) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        pb->prof_ev.push_back(e);
    }
    if (pb->prof_flags_cap < pb->prof_ev.size() / 2) {
        cudaFree(pb->d_prof_flags);
        pb->d_prof_flags = nullptr;
        CK(cudaMalloc(&pb->d_prof_flags, sizeof(int) * (pb->prof_ev.size() / 2)));
        pb->prof_flags_cap = pb->prof_ev.size() / 2;
    }
    if (pb->d_prof_flags)
        CK(cudaMemsetAsync(pb->d_prof_flags, 0, sizeof(int) * pb->prof_flags_cap, pb->stream));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_profile(gslnls_problem *pb, float *avg_pass_ms, int64_t *npasses_timed)
{
    if (!pb || !avg_pass_ms || !npasses_timed)
        return GSLNLS_EINVAL;
    CK(cudaStreamSynchronize(pb->stream));
    double tot = 0.0;
    std::vector<int> real(pb->prof_used / 2 + 1, 0);
    if (pb->prof_used)
        CK(cudaMemcpy(real.data(), pb->d_prof_flags, sizeof(int) * (pb->prof_used / 2), cudaMemcpyDeviceToHost));
    int64_t cnt = 0;
    for (size_t i = 0; i + 1 < pb->prof_used; i += 2) {
        if (!real[i / 2])
            continue; // an idle launch behind a finished fit: not a pass
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, pb->prof_ev[i], pb->prof_ev[i + 1]));
        tot += ms;
        cnt += real[i / 2]; // 1 for a per-pass launch; the persistent kernel counts the passes it ran
    }
    *npasses_timed = cnt;
    *avg_pass_ms = *npasses_timed ? (float)(tot / (double)*npasses_timed) : 0.f;
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_channel_stats(gslnls_problem *pb, int reset, double *avg_stream_us, double *avg_step_us,
                                            int64_t *npasses)
{
    if (!pb)
        return GSLNLS_EINVAL;
    char *ch = channel_of(pb);
    if (avg_stream_us) *avg_stream_us = 0.0;
    if (avg_step_us) *avg_step_us = 0.0;
    if (npasses) *npasses = 0;
    if (!ch)
        return GSLNLS_SUCCESS;
    CK(cudaSetDevice(pb->device));
    CK(cudaStreamSynchronize(pb->stream));
    unsigned long long tm[5] = {0, 0, 0, 0, 0};
    CK(cudaMemcpy(tm, ch + NLS_CH_TIMER, sizeof(tm), cudaMemcpyDeviceToHost));
    if (avg_stream_us && tm[2]) *avg_stream_us = 1e-3 * (double)tm[1] / (double)tm[2];
    if (avg_step_us && tm[4]) *avg_step_us = 1e-3 * (double)tm[3] / (double)tm[4];
    if (npasses) *npasses = (int64_t)tm[2];
    if (reset)
        CK(cudaMemset(ch + NLS_CH_TIMER, 0, sizeof(tm)));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_fit_large_sharded(const gslnls_model *m, const double *const *vars, const double *y,
                                        const double *weights, int64_t n_local, const double *start,
                                        const int *control_int, const double *control_dbl, int device,
                                        gslnls_comm *comm, int want_resid_grad, gslnls_result *out)
{
    if (!m || !y || !start || !control_int || !control_dbl || !out)
        return GSLNLS_EINVAL;
    std::memset(out, 0, sizeof(*out));
    const bool sharded = comm && comm->nranks > 1;
    if (!sharded && n_local < m->p) {
        // R/nls_large.R:286-288
        set_error("negative residual degrees of freedom, cannot fit a model with less observations than parameters");
        return GSLNLS_EINVAL;
    }
    static const bool trace = std::getenv("GSLNLS_TRACE_E2E") != nullptr; // developer aid: phase times on stderr
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    const bool use_cache = cache_enabled();
    gslnls_problem *pb = use_cache ? cache_take(m, weights != nullptr, device) : nullptr;
    int rc = GSLNLS_SUCCESS;
    if (pb) {
        pb->n = n_local; // the buffers grow on demand in upload(); grid and workspace follow in prepare()
        pb->n_total = n_local;
        pb->comm = nullptr;
        gslnls_problem_set_weights_mode(pb, default_weights_mode());
    } else {
        rc = gslnls_problem_create(m, n_local, weights != nullptr, device, &pb);
        if (rc)
            return rc;
    }
    pb->upload_sharing = (comm && comm->local) ? comm->nranks : 1; // local group: all shards upload at once
    const double t1 = now();
    rc = gslnls_problem_upload(pb, vars, y, weights);
    if (rc == GSLNLS_SUCCESS && sharded) {
        rc = gslnls_problem_set_comm(pb, comm);
        if (rc == GSLNLS_SUCCESS && pb->n_total < m->p) {
            set_error("negative residual degrees of freedom, cannot fit a model with less observations than parameters");
            rc = GSLNLS_EINVAL;
        }
    }
    double t2 = t1, t3 = t1;
    if (rc == GSLNLS_SUCCESS) {
        if (trace) {
            cudaStreamSynchronize(pb->stream);
            t2 = now();
        }
        rc = gslnls_problem_fit(pb, start, control_int, control_dbl, want_resid_grad, out);
        t3 = now();
    }
    if (use_cache && rc < 1000) {
        pb->comm = nullptr; // the caller owns the comm and may free it before the next call
        cache_put(pb);
    } else {
        gslnls_problem_free(pb);
    }
    if (trace)
        std::fprintf(stderr, "gslnls_fit_large: create %.2f ms, upload %.2f ms, fit %.2f ms, free %.2f ms\n", t1 - t0,
                     t2 - t1, t3 - t2, now() - t3);
    return rc;
}

GSLNLS_API int gslnls_fit_large(const gslnls_model *m, const double *const *vars, const double *y,
                                const double *weights, int64_t n, const double *start, const int *control_int,
                                const double *control_dbl, int device, int want_resid_grad, gslnls_result *out)
{
    return gslnls_fit_large_sharded(m, vars, y, weights, n, start, control_int, control_dbl, device, nullptr,
                                    want_resid_grad, out);
}

} // extern "C"

// ---- sessions: one handle for "the data of a fit, resident on 1..R GPUs of this process" ------------
// What the R shim keeps behind an external pointer: the rows are split into contiguous, 2-aligned shards,
// one gslnls_problem per GPU, each driven by its own host thread (upload over that GPU's PCIe link, sharded
// fit with packets crossing NVLink peer memory).  resid / grad of src/nls_large.c:339-385 are produced on
// demand from the resident shards (lazy post-fit accessors), not by the fit.
struct gslnls_session {
    const gslnls_model *model = nullptr;
    int64_t n = 0, per = 0;
    int has_w = 0;
    std::vector<int> dev;
    std::vector<gslnls_problem *> pb;
    std::vector<gslnls_comm *> comm; // empty for one GPU
    int R() const { return (int)pb.size(); }
    int64_t lo(int r) const { return std::min<int64_t>(n, (int64_t)r * per); }
    int64_t hi(int r) const { return std::min<int64_t>(n, lo(r) + per); }
};

namespace {
template <class F>
int on_every_rank(gslnls_session *s, F f)
{
    const int R = s->R();
    if (R == 1)
        return f(0);
    std::vector<int> rcs(R, GSLNLS_SUCCESS);
    std::vector<std::string> errs(R);
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r)
        th.emplace_back([&, r] {
            rcs[r] = f(r);
            errs[r] = g_last_error;
        });
    for (std::thread &t : th)
        t.join();
    for (int r = 0; r < R; ++r)
        if (rcs[r] >= 1000 || rcs[r] == GSLNLS_EINVAL) {
            set_error(errs[r]);
            return rcs[r];
        }
    return rcs[0];
}

// the session of the last gslnls_fit_large_multi call, kept for the next one (peer-access setup, the n-sized
// device buffers, streams and loaded kernels are all slow to create)
std::mutex g_multi_mu;
gslnls_session *g_multi = nullptr;
} // namespace

extern "C" GSLNLS_API void gslnls_session_free(gslnls_session *s);
static void multi_cache_drop(const gslnls_model *m)
{
    std::lock_guard<std::mutex> lk(g_multi_mu);
    if (g_multi && (!m || g_multi->model == m)) {
        gslnls_session_free(g_multi);
        g_multi = nullptr;
    }
}

extern "C" {

GSLNLS_API void gslnls_session_free(gslnls_session *s)
{
    if (!s)
        return;
    for (gslnls_problem *pb : s->pb)
        gslnls_problem_free(pb);
    for (gslnls_comm *c : s->comm)
        gslnls_comm_free(c);
    delete s;
}

GSLNLS_API int gslnls_session_create(const gslnls_model *m, int64_t n, int has_weights, int ngpu, const int *devices,
                                     gslnls_session **out)
{
    if (!m || !out || n < 0 || ngpu < 1 || ngpu > NLS_MAX_RANKS)
        return GSLNLS_EINVAL;
    *out = nullptr;
    // contiguous, 2-aligned row ranges (keeps every shard's columns 16-byte aligned); no empty shards
    int R = (int)std::min<int64_t>(ngpu, std::max<int64_t>(1, n / 2));
    auto rows_per = [&](int r) {
        int64_t q = (n + r - 1) / r;
        return q + (q & 1);
    };
    while (R > 1 && (int64_t)(R - 1) * rows_per(R) >= n) // rounding the shard length up to even can empty the last shard
        --R;
    gslnls_session *s = new gslnls_session();
    s->model = m;
    s->n = n;
    s->has_w = has_weights ? 1 : 0;
    s->per = R > 1 ? rows_per(R) : n;
    for (int r = 0; r < R; ++r)
        s->dev.push_back(devices ? devices[r] : r);
    if (R > 1) {
        s->comm.assign(R, nullptr);
        int rc = gslnls_comm_create_local(R, s->dev.data(), s->comm.data());
        if (rc) {
            s->comm.clear();
            gslnls_session_free(s);
            return rc;
        }
    }
    for (int r = 0; r < R; ++r) {
        gslnls_problem *pb = nullptr;
        int rc = gslnls_problem_create(m, s->hi(r) - s->lo(r), s->has_w, s->dev[r], &pb);
        if (rc) {
            gslnls_session_free(s);
            return rc;
        }
        pb->upload_sharing = R;
        s->pb.push_back(pb);
    }
    *out = s;
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_session_ngpu(const gslnls_session *s) { return s ? s->R() : 0; }

GSLNLS_API int gslnls_session_set_weights_mode(gslnls_session *s, int mode)
{
    if (!s)
        return GSLNLS_EINVAL;
    for (gslnls_problem *pb : s->pb) {
        int rc = gslnls_problem_set_weights_mode(pb, mode);
        if (rc)
            return rc;
    }
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_session_upload(gslnls_session *s, const double *const *vars, const double *y, const double *weights)
{
    if (!s || !y || (s->model->nvar > 0 && !vars) || (s->has_w && !weights))
        return GSLNLS_EINVAL;
    const int nvar = s->model->nvar;
    return on_every_rank(s, [&](int r) {
        const int64_t lo = s->lo(r);
        std::vector<const double *> v(std::max(nvar, 1), nullptr);
        for (int k = 0; k < nvar; ++k)
            v[k] = vars[k] + lo;
        int rc = gslnls_problem_upload(s->pb[r], v.data(), y + lo, weights ? weights + lo : nullptr);
        if (rc == GSLNLS_SUCCESS && s->R() > 1) {
            s->comm[r]->n_total_hint = s->n;
            rc = gslnls_problem_set_comm(s->pb[r], s->comm[r]);
        }
        return rc;
    });
}

GSLNLS_API int gslnls_session_residuals(gslnls_session *s, const double *theta, double *resid, double *grad)
{
    if (!s || !theta)
        return GSLNLS_EINVAL;
    // every rank writes its rows straight into the caller's n-row arrays (grad is n x p column-major)
    return on_every_rank(s, [&](int r) {
        const int64_t lo = s->lo(r);
        return problem_residuals_ld(s->pb[r], theta, resid ? resid + lo : nullptr, grad ? grad + lo : nullptr, s->n);
    });
}

GSLNLS_API int gslnls_session_fit(gslnls_session *s, const double *start, const int *control_int,
                                  const double *control_dbl, int want_resid_grad, gslnls_result *out)
{
    if (!s || !start || !control_int || !control_dbl || !out)
        return GSLNLS_EINVAL;
    std::memset(out, 0, sizeof(*out));
    if (s->n < s->model->p) {
        // R/nls_large.R:286-288
        set_error("negative residual degrees of freedom, cannot fit a model with less observations than parameters");
        return GSLNLS_EINVAL;
    }
    const int R = s->R();
    std::vector<gslnls_result> res(R);
    for (gslnls_result &q : res)
        std::memset(&q, 0, sizeof(q));
    int rc = on_every_rank(s, [&](int r) {
        return gslnls_problem_fit(s->pb[r], start, control_int, control_dbl, 0, &res[r]);
    });
    if (rc >= 1000 || rc == GSLNLS_EINVAL) {
        for (gslnls_result &q : res)
            gslnls_result_free(&q);
        return rc;
    }
    *out = res[0]; // every rank holds bitwise the same result
    out->n = s->n;
    out->n_local = s->n;
    for (int r = 1; r < R; ++r)
        gslnls_result_free(&res[r]);
    if (want_resid_grad) {
        // src/nls_large.c:339-385; NaN-filled on failure (:345-349, :371-376)
        const int p = s->model->p;
        out->resid = (double *)std::malloc(sizeof(double) * (size_t)std::max<int64_t>(s->n, 1));
        out->grad = (double *)std::malloc(sizeof(double) * (size_t)std::max<int64_t>(s->n, 1) * p);
        if (rc == GSLNLS_SUCCESS || rc == GSLNLS_EMAXITER) {
            int r2 = gslnls_session_residuals(s, out->par, out->resid, out->grad);
            if (r2) {
                gslnls_result_free(out);
                return r2;
            }
        } else {
            for (int64_t i = 0; i < s->n; ++i)
                out->resid[i] = NAN;
            for (int64_t i = 0; i < s->n * p; ++i)
                out->grad[i] = NAN;
        }
    }
    return rc;
}

// ---- one call, several GPUs, one process: the `int ngpu, const int *devices` form of src/nls_large.c:66 ----
GSLNLS_API int gslnls_fit_large_multi(const gslnls_model *m, const double *const *vars, const double *y,
                                      const double *weights, int64_t n, const double *start,
                                      const int *control_int, const double *control_dbl, int ngpu,
                                      const int *devices, int want_resid_grad, gslnls_result *out)
{
    if (!m || !y || !start || !control_int || !control_dbl || !out || ngpu < 1 || ngpu > NLS_MAX_RANKS)
        return GSLNLS_EINVAL;
    std::memset(out, 0, sizeof(*out));
    if (n < m->p) {
        set_error("negative residual degrees of freedom, cannot fit a model with less observations than parameters");
        return GSLNLS_EINVAL;
    }
    if (ngpu == 1)
        return gslnls_fit_large_sharded(m, vars, y, weights, n, start, control_int, control_dbl,
                                        devices ? devices[0] : 0, nullptr, want_resid_grad, out);
    std::lock_guard<std::mutex> lk(g_multi_mu); // one multi-GPU one-shot fit at a time per process
    std::vector<int> dev(ngpu);
    for (int r = 0; r < ngpu; ++r)
        dev[r] = devices ? devices[r] : r;
    gslnls_session *s = g_multi;
    bool reuse = s && s->model == m && s->has_w == (weights ? 1 : 0) && (int)s->dev.size() <= ngpu &&
                 std::equal(s->dev.begin(), s->dev.end(), dev.begin());
    if (reuse && s->n != n) {
        // same devices, another row count: keep the peer-access group and the streams, re-split the rows (the
        // column buffers grow on demand in upload())
        int R = (int)std::min<int64_t>(ngpu, std::max<int64_t>(1, n / 2));
        auto rows_per = [&](int r) {
            int64_t q = (n + r - 1) / r;
            return q + (q & 1);
        };
        while (R > 1 && (int64_t)(R - 1) * rows_per(R) >= n)
            --R;
        if (R != s->R()) {
            reuse = false;
        } else {
            s->n = n;
            s->per = R > 1 ? rows_per(R) : n;
            for (int r = 0; r < R; ++r) {
                s->pb[r]->n = s->hi(r) - s->lo(r);
                s->pb[r]->n_total = n;
            }
        }
    }
    if (!reuse) {
        gslnls_session_free(g_multi);
        g_multi = nullptr;
        int rc = gslnls_session_create(m, n, weights != nullptr, ngpu, dev.data(), &s);
        if (rc)
            return rc;
        if (cache_enabled())
            g_multi = s;
    }
    gslnls_session_set_weights_mode(s, default_weights_mode());
    int rc = gslnls_session_upload(s, vars, y, weights);
    if (rc == GSLNLS_SUCCESS)
        rc = gslnls_session_fit(s, start, control_int, control_dbl, want_resid_grad, out);
    if (s != g_multi)
        gslnls_session_free(s);
    else if (rc >= 1000) { // do not keep a session that failed at the library level
        gslnls_session_free(g_multi);
        g_multi = nullptr;
    }
    return rc;
}

GSLNLS_API void gslnls_cache_clear(void)
{
    cache_drop(nullptr);
    upload_pools_release();
}

GSLNLS_API void gslnls_result_free(gslnls_result *r)
{
    if (!r)
        return;
    std::free(r->par); std::free(r->covar); std::free(r->partrace); std::free(r->ssrtrace);
    std::free(r->condtrace); std::free(r->resid); std::free(r->grad); std::free(r->jtj); std::free(r->grad_vec);
    std::free(r->x_final);
    std::memset(r, 0, sizeof(*r));
}

// ---- batched multi-start inner kernels ------------------------------------------------------
// S candidates side by side: candidates ride blockIdx.y of the pass kernel and one thread each of the batched
// trust-region step; `iters` outer iterations per candidate.  Leaves the S state records in `states`.
static int batch_run(gslnls_problem *pb, const double *starts, int S, const int *control_int,
                     const double *control_dbl, std::vector<double> &states)
{
    if (pb->p > 8) {
        set_error("batched multi-start supports p <= 8");
        return GSLNLS_EINVAL;
    }
    if (S > 65535) {
        set_error("at most 65535 candidates per batch");
        return GSLNLS_EINVAL;
    }
    int rc = fill_params(pb, control_int, control_dbl, control_int[0]);
    if (rc)
        return rc;
    pb->P.trace = 0;
    stop_server(pb);
    rc = prepare(pb, S, 0, true);
    if (rc)
        return rc;
    const int p = pb->p;
    CK(cudaMemcpyAsync(pb->d_starts, starts, sizeof(double) * (size_t)S * p, cudaMemcpyHostToDevice, pb->stream));
    CK(trs_launch_reset(pb->d_state, pb->state_stride, pb->d_req, pb->req_stride, pb->d_starts, p, S, pb->d_ndone,
                        pb->stream));
    ++pb->launches;
    const int64_t hard_cap = (int64_t)pb->P.maxiter * 34 + 8;
    int64_t total = 0;
    int fin = 0;
    while (!fin && total < hard_cap) {
        for (int i = 0; i < pb->chunk; ++i) {
            rc = launch_pass(pb, S, 0);
            if (rc)
                return rc;
            rc = exchange_packet(pb, (size_t)S * pb->pk_stride);
            if (rc)
                return rc;
            CK(trs_launch_step_batch(pb->P, pb->d_state, pb->state_stride, pb->d_packet, pb->pk_stride, pb->d_req,
                                     pb->req_stride, S, pb->d_ndone, pb->stream));
            ++pb->launches;
        }
        total += pb->chunk;
        CK(cudaMemcpyAsync(pb->h_ndone, pb->d_ndone, sizeof(int), cudaMemcpyDeviceToHost, pb->stream));
        CK(cudaStreamSynchronize(pb->stream));
        fin = pb->h_ndone[0] >= S;
    }
    states.resize((size_t)S * pb->state_stride);
    CK(cudaMemcpy(states.data(), pb->d_state, sizeof(double) * states.size(), cudaMemcpyDeviceToHost));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_fit_batch(gslnls_problem *pb, const double *starts, int S, const int *control_int,
                                        const double *control_dbl, double *par_out, double *ssr_out,
                                        double *logdet_out, int *conv_out, int *niter_out)
{
    if (!pb || !starts || S < 1 || !control_int || !control_dbl)
        return GSLNLS_EINVAL;
    std::vector<double> St;
    int rc = batch_run(pb, starts, S, control_int, control_dbl, St);
    if (rc)
        return rc;
    const int p = pb->p;
    for (int c = 0; c < S; ++c) {
        const double *s = St.data() + (size_t)c * pb->state_stride;
        if (par_out)
            std::memcpy(par_out + (size_t)c * p, s + trs::S_COUNT, sizeof(double) * p);
        if (ssr_out)
            ssr_out[c] = s[trs::S_CHISQ1];
        if (logdet_out)
            logdet_out[c] = s[trs::S_LOGDET0];
        if (conv_out)
            conv_out[c] = (int)s[trs::S_PHASE] == trs::PH_DONE ? (int)s[trs::S_STATUS] : GSLNLS_CONTINUE;
        if (niter_out)
            niter_out[c] = (int)s[trs::S_NITER];
    }
    return GSLNLS_SUCCESS;
}

} // extern "C"

// ---- multi-start global search (control logic: mstart.hpp; local searches: the batched kernels) ----
namespace {
struct GpuBatchEvaluator {
    gslnls_problem *pb;
    const int *ci;
    const double *cd;
    int rc = GSLNLS_SUCCESS;
    void operator()(const std::vector<double> &starts, int S, int iters, std::vector<mstart::BatchResult> &out)
    {
        out.assign(S, mstart::BatchResult());
        const int p = pb->p;
        int c_int[7];
        double c_dbl[8];
        std::memcpy(c_int, ci, sizeof(c_int));
        std::memcpy(c_dbl, cd, sizeof(c_dbl));
        c_int[0] = iters;
        c_int[1] = 0;
        c_dbl[7] = 1.0e-3; // gtol of the inner searches, src/nls_mstart.c:90, :250
        std::vector<double> St;
        if (rc == GSLNLS_SUCCESS)
            rc = batch_run(pb, starts.data(), S, c_int, c_dbl, St);
        for (int c = 0; c < S; ++c) {
            mstart::BatchResult &b = out[c];
            b.par.assign(starts.begin() + (size_t)c * p, starts.begin() + (size_t)(c + 1) * p);
            b.diag.assign(p, 1.0);
            b.ssr = b.ssr_prev = b.ssr_start = std::numeric_limits<double>::infinity();
            b.logdet_start = b.logdet_end = -std::numeric_limits<double>::infinity();
            b.status = GSLNLS_FAILURE;
            if (rc != GSLNLS_SUCCESS)
                continue;
            const double *s = St.data() + (size_t)c * pb->state_stride;
            const double *v = s + trs::S_COUNT;
            b.par.assign(v, v + p);
            b.diag.assign(v + 3 * p, v + 4 * p);
            b.ssr = s[trs::S_CHISQ1];
            b.ssr_prev = s[trs::S_CHISQ0];
            b.ssr_start = s[trs::S_CHISQ_INIT];
            b.logdet_start = s[trs::S_LOGDET0];
            b.logdet_end = s[trs::S_LOGDET1];
            b.status = (int)s[trs::S_STATUS];
        }
    }
};
} // namespace

extern "C" {

GSLNLS_API int gslnls_problem_multistart(gslnls_problem *pb, const double *range, const int *has_range,
                                         const int *control_int, const double *control_dbl, const int *mstart_int,
                                         const double *mstart_dbl, gslnls_mstart_result *out)
{
    if (!pb || !range || !has_range || !control_int || !control_dbl || !mstart_int || !mstart_dbl || !out)
        return GSLNLS_EINVAL;
    std::memset(out, 0, sizeof(*out));
    mstart::Control c;
    c.n = mstart_int[0]; c.p = mstart_int[1]; c.q = mstart_int[2]; c.s = mstart_int[3];
    c.niter = mstart_int[4]; c.max = mstart_int[5]; c.minsp = mstart_int[6];
    c.r = mstart_dbl[0]; c.tol = mstart_dbl[1];
    if (c.n < 1 || c.p < 1 || c.q < 1 || c.s < 1 || c.niter < 1 || c.max < 1 || c.minsp < 1 || !(c.r > 1.0) || !(c.tol > 0.0)) {
        set_error("invalid multi-start control values"); // R/nls.R:681-689
        return GSLNLS_EINVAL;
    }
    GpuBatchEvaluator ev{pb, control_int, control_dbl};
    mstart::Driver<GpuBatchEvaluator> drv(pb->p, c, range, has_range, control_dbl[5], control_dbl[6], ev);
    const mstart::Outcome o = drv.run();
    if (ev.rc)
        return ev.rc;
    const int p = pb->p;
    out->p = p;
    out->par = dup(o.par.data(), p);
    out->range = dup(o.range.data(), 2 * (size_t)p);
    out->ssr = o.ssr;
    out->ssrconv = o.ssrconv;
    out->nsp = o.nsp;
    out->nwsp = o.nwsp;
    out->mstarts = o.mstarts;
    out->status = o.status;
    out->searches = o.searches;
    return GSLNLS_SUCCESS;
}

GSLNLS_API void gslnls_mstart_result_free(gslnls_mstart_result *r)
{
    if (!r)
        return;
    std::free(r->par);
    std::free(r->range);
    std::memset(r, 0, sizeof(*r));
}

/* test hook: the first `count` points of the quasi-random generator the multi-start sampler uses for `dim`
 * parameters (Sobol below 41 dimensions, Halton above), row-major count x dim */
GSLNLS_API int gslnls_qrng_points(int dim, int count, double *out)
{
    if (dim < 1 || count < 0 || !out)
        return GSLNLS_EINVAL;
    if (dim < 41) {
        mstart::Sobol g(dim);
        for (int i = 0; i < count; ++i)
            if (!g.next(out + (size_t)i * dim))
                return GSLNLS_FAILURE;
    } else {
        mstart::Halton g(dim);
        for (int i = 0; i < count; ++i)
            g.next(out + (size_t)i * dim);
    }
    return GSLNLS_SUCCESS;
}

} // extern "C"

// ---- IRLS: robust losses on the large path (src/nls_irls.c:412-546) -------------------------------------
namespace {
struct IrlsScratch {
    unsigned long long *d_hist = nullptr; // [256] + cnt_min [2]
    double *d_partial = nullptr;
    double *d_userw = nullptr;
    double *d_theta = nullptr;
    ~IrlsScratch()
    {
        cudaFree(d_hist); cudaFree(d_partial); cudaFree(d_userw); cudaFree(d_theta);
    }
};

void irls_base_params(const gslnls_problem *pb, const double *d_theta, NlsIrlsParams &prm)
{
    std::memset(&prm, 0, sizeof(prm));
    for (int k = 0; k < pb->nvar; ++k)
        prm.vars[k] = pb->dvars[k];
    prm.y = pb->dy;
    prm.n = pb->n;
    prm.theta = d_theta;
    prm.h_df = pb->h_df;
}

// median of |fn(theta) - y| as gsl_median computes it (src/nls_utils.c:162-189): middle element, or the mean
// of the two middle elements; radix select over the bit patterns, 8 bits per streaming pass
int irls_median(gslnls_problem *pb, const double *d_theta, IrlsScratch &sc, double *median)
{
    const int64_t n = pb->n;
    if (n == 0) {
        *median = 0.0;
        return GSLNLS_SUCCESS;
    }
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)pb->num_sms * 8);
    NlsIrlsParams prm;
    irls_base_params(pb, d_theta, prm);
    prm.hist = sc.d_hist;
    prm.cnt_min = sc.d_hist + 256;
    unsigned long long rank = (unsigned long long)((n - 1) / 2), prefix = 0ull, hist[256];
    for (int shift = 56; shift >= 0; shift -= 8) {
        CK(cudaMemsetAsync(sc.d_hist, 0, sizeof(unsigned long long) * 256, pb->stream));
        prm.prefix = prefix;
        prm.shift = shift;
        void *args[] = {&prm};
        CK(cudaLaunchKernel((const void *)pb->var->irls_hist, dim3(blocks), dim3(256), args, 0, pb->stream));
        ++pb->launches;
        CK(cudaMemcpyAsync(hist, sc.d_hist, sizeof(hist), cudaMemcpyDeviceToHost, pb->stream));
        CK(cudaStreamSynchronize(pb->stream));
        int b = 0;
        for (; b < 255 && rank >= hist[b]; ++b)
            rank -= hist[b];
        prefix = (prefix << 8) | (unsigned long long)b;
    }
    double lo;
    std::memcpy(&lo, &prefix, sizeof(lo));
    *median = lo;
    if ((n - 1) / 2 != n / 2) { // even length: the next order statistic as well
        unsigned long long init[2] = {0ull, ~0ull}, got[2];
        CK(cudaMemcpyAsync(sc.d_hist + 256, init, sizeof(init), cudaMemcpyHostToDevice, pb->stream));
        prm.pivot = prefix;
        void *args[] = {&prm};
        CK(cudaLaunchKernel((const void *)pb->var->irls_above, dim3(blocks), dim3(256), args, 0, pb->stream));
        ++pb->launches;
        CK(cudaMemcpyAsync(got, sc.d_hist + 256, sizeof(got), cudaMemcpyDeviceToHost, pb->stream));
        CK(cudaStreamSynchronize(pb->stream));
        double hi = lo;
        if (got[0] <= (unsigned long long)(n / 2)) // fewer than n/2 + 1 elements <= lo: the upper middle is larger
            std::memcpy(&hi, &got[1], sizeof(hi));
        *median = (lo + hi) / 2.0;
    }
    return GSLNLS_SUCCESS;
}

int irls_scratch_init(gslnls_problem *pb, IrlsScratch &sc, bool want_userw)
{
    CK(cudaMalloc(&sc.d_hist, sizeof(unsigned long long) * 258));
    CK(cudaMalloc(&sc.d_partial, sizeof(double) * (size_t)pb->num_sms * 8));
    CK(cudaMalloc(&sc.d_theta, sizeof(double) * pb->p));
    if (want_userw) {
        CK(cudaMalloc(&sc.d_userw, sizeof(double) * (size_t)std::max<int64_t>(pb->n, 1)));
        CK(cudaMemcpyAsync(sc.d_userw, pb->dw, sizeof(double) * (size_t)pb->n, cudaMemcpyDeviceToDevice, pb->stream));
    }
    return GSLNLS_SUCCESS;
}
} // namespace

extern "C" {

GSLNLS_API int gslnls_problem_median_abs_resid(gslnls_problem *pb, const double *theta, double *median)
{
    if (!pb || !theta || !median)
        return GSLNLS_EINVAL;
    int rc = prepare(pb, 1, 0, false);
    if (rc)
        return rc;
    IrlsScratch sc;
    rc = irls_scratch_init(pb, sc, false);
    if (rc)
        return rc;
    CK(cudaMemcpyAsync(sc.d_theta, theta, sizeof(double) * pb->p, cudaMemcpyHostToDevice, pb->stream));
    return irls_median(pb, sc.d_theta, sc, median);
}

GSLNLS_API int gslnls_problem_get_weights(gslnls_problem *pb, double *weights)
{
    if (!pb || !weights || !pb->dw)
        return GSLNLS_EINVAL;
    CK(cudaSetDevice(pb->device));
    CK(cudaMemcpyAsync(weights, pb->dw, sizeof(double) * (size_t)pb->n, cudaMemcpyDeviceToHost, pb->stream));
    CK(cudaStreamSynchronize(pb->stream));
    return GSLNLS_SUCCESS;
}

GSLNLS_API int gslnls_problem_fit_irls(gslnls_problem *pb, const double *start, const int *control_int,
                                       const double *control_dbl, int loss, const double *cc, int irls_maxiter,
                                       double irls_xtol, gslnls_result *out, gslnls_irls_info *info)
{
    if (!pb || !start || !control_int || !control_dbl || !cc || !out || !info || loss < 1 || loss > 8 || irls_maxiter < 1)
        return GSLNLS_EINVAL;
    if (!pb->has_w || !pb->dw || pb->bound) {
        set_error("IRLS needs a problem created with has_weights = 1 and library-owned columns (the weights "
                  "column is the working vector)");
        return GSLNLS_EINVAL;
    }
    if (pb->comm && pb->comm->nranks > 1) {
        set_error("IRLS on a sharded problem is not built (the median needs a cross-rank histogram)");
        return GSLNLS_EINVAL;
    }
    std::memset(out, 0, sizeof(*out));
    info->sigma = 1.0; info->delta = 0.0; info->niter = 0; info->status = GSLNLS_FAILURE;
    const int p = pb->p;
    int rc = prepare(pb, 1, 0, false);
    if (rc)
        return rc;
    IrlsScratch sc;
    rc = irls_scratch_init(pb, sc, true);
    if (rc)
        return rc;
    double *d_work = const_cast<double *>(pb->dw); // library-owned (checked above)
    std::vector<double> prev(start, start + p), cur(p);
    const int blocks = (int)std::min<int64_t>(std::max<int64_t>((pb->n + 255) / 256, 1), (int64_t)pb->num_sms * 8);
    int status = GSLNLS_CONTINUE;
    for (;;) {
        info->niter += 1;
        gslnls_result_free(out);
        // weighted fit with the current weights; every IRLS iteration restarts from the given start values
        // (gsl_multifit_nlinear_winit(pars->mpopt, ...), src/nls_irls.c:452-457)
        status = gslnls_problem_fit(pb, start, control_int, control_dbl, 0, out);
        if (status >= 1000 || status == GSLNLS_EINVAL)
            return status;
        if (status == GSLNLS_EBADFUNC || (status == GSLNLS_ENOPROG && info->niter == 1))
            return status; // :479-484
        std::memcpy(cur.data(), out->x_final, sizeof(double) * p);
        rc = prepare(pb, 1, 0, false); // kernels of the launch-ordered variant (the fit may have switched variant)
        if (rc)
            return rc;
        // sigma = 1.4826 median |unweighted residual| (:494)
        CK(cudaMemcpyAsync(sc.d_theta, cur.data(), sizeof(double) * p, cudaMemcpyHostToDevice, pb->stream));
        double med = 0.0;
        rc = irls_median(pb, sc.d_theta, sc, &med);
        if (rc)
            return rc;
        info->sigma = 1.482602218505602 * med;
        // w_i = max(psi(r_i / sigma) / (r_i / sigma), eps), normalised to sum n, times the user's weights (:496-515)
        NlsIrlsParams prm;
        irls_base_params(pb, sc.d_theta, prm);
        prm.sigma = info->sigma;
        prm.loss = loss;
        prm.cc[0] = cc[0]; prm.cc[1] = cc[1]; prm.cc[2] = cc[2];
        prm.wout = d_work;
        prm.partial = sc.d_partial;
        void *args[] = {&prm};
        CK(cudaLaunchKernel((const void *)pb->var->irls_weights, dim3(blocks), dim3(256), args, 0, pb->stream));
        std::vector<double> part(blocks);
        CK(cudaMemcpyAsync(part.data(), sc.d_partial, sizeof(double) * blocks, cudaMemcpyDeviceToHost, pb->stream));
        CK(cudaStreamSynchronize(pb->stream));
        double sum_wts = 0.0;
        for (double v : part)
            sum_wts += v; // CTA order: deterministic
        prm.scale = (double)pb->n / sum_wts;
        prm.userw = sc.d_userw;
        CK(cudaLaunchKernel((const void *)pb->var->irls_scale, dim3(blocks), dim3(256), args, 0, pb->stream));
        pb->launches += 2;
        // convergence of the parameters (test_delta_irls, :365-384)
        bool conv = true;
        info->delta = 0.0;
        for (int i = 0; i < p; ++i) {
            const double dxi = std::fabs(prev[i] - cur[i]), rel = dxi / std::fabs(cur[i]);
            info->delta = std::max(info->delta, dxi);
            if (!((rel < dxi ? rel : dxi) < irls_xtol))
                conv = false;
        }
        if (conv) {
            info->status = GSLNLS_SUCCESS;
            CK(cudaStreamSynchronize(pb->stream));
            return status;
        }
        prev = cur;
        if (info->niter >= irls_maxiter)
            break;
    }
    info->status = GSLNLS_EMAXITER; // :535-541
    out->conv = GSLNLS_EMAXITER;
    out->status = gslnls_strerror(GSLNLS_EMAXITER);
    CK(cudaStreamSynchronize(pb->stream));
    return GSLNLS_EMAXITER;
}

} // extern "C"
