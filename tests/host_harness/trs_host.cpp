// tests/host_harness/trs_host.cpp -- TEST INFRASTRUCTURE: compiles the device trust-region state
// machine (gslnls_b200/csrc/trs_core.h) for the host with the SingleLane policy so that the CPU
// test-suite can step it against the oracle without a GPU.  Never linked into libgslnls_b200.so.
#include <barrier>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "../../gslnls_b200/csrc/trs_core.h"

extern "C" {

typedef int (*packet_cb)(void *ctx, int mode, const double *theta, const double *v, double *packet);

struct trs_host_params {
    int p, maxiter, trs, scale, trace, batch_iters;
    long long cg_maxit;
    double factor_up, factor_down, avmax, h_df, h_fvv, xtol, ftol, gtol, cg_tol;
};

int trs_host_state_doubles(int p) { return trs::state_doubles(p); }

// returns number of packets consumed; state (state_doubles), traces ((maxiter+1)*p, maxiter+1, maxiter+1)
long trs_host_fit(const trs_host_params *hp, const double *start, packet_cb cb, void *ctx, double *state,
                  double *partrace, double *ssrtrace, double *condtrace, long max_packets)
{
    trs::Params P;
    P.p = hp->p; P.maxiter = hp->maxiter; P.trs = hp->trs; P.scale = hp->scale; P.trace = hp->trace;
    P.batch_iters = hp->batch_iters; P.cg_maxit = hp->cg_maxit; P.factor_up = hp->factor_up;
    P.factor_down = hp->factor_down; P.avmax = hp->avmax; P.h_df = hp->h_df; P.h_fvv = hp->h_fvv;
    P.xtol = hp->xtol; P.ftol = hp->ftol; P.gtol = hp->gtol; P.cg_tol = hp->cg_tol;
    const int p = P.p;
    std::vector<double> req(trs::request_doubles(p)), pk(trs::packet_doubles(p) + 2), jtj(p * p), work(p * p);
    trs::state_reset(state, req.data(), start, p);
    long n = 0;
    typedef trs::Solver<128, trs::SingleLane> S;
    S *solver = new S(P, trs::SingleLane(), jtj.data(), work.data());
    while ((int)state[trs::S_PHASE] != trs::PH_DONE && n < max_packets) {
        const int mode = (int)req[0];
        if (mode == trs::MODE_IDLE)
            break;
        if (cb(ctx, mode, req.data() + 1, req.data() + 1 + p, pk.data()))
            break;
        solver->advance(state, pk.data(), req.data(), partrace, ssrtrace, condtrace);
        ++n;
    }
    delete solver;
    return n;
}

// ---- the warp, emulated: one host thread per lane --------------------------------------------
// The device runs the state machine with 32 lanes that share the p x p matrices and the O(p^2) / O(p^3)
// loops (trs::WarpLanes: __syncwarp, shuffles, votes).  Here every lane is a thread with its own Solver
// (private vectors), the matrices are shared, sync() is a barrier, bcast() / all() go through a shared
// exchange array.  Unlike a warp the threads are scheduled independently, so a missing sync() in
// trs_core.h shows up here as a wrong answer (or under -fsanitize=thread) instead of going unnoticed.
}  // extern "C"

namespace {
struct LaneShared {
    explicit LaneShared(int n) : bar(n), xch(n), flag(n), n(n) {}
    std::barrier<> bar;
    std::vector<double> xch;
    std::vector<int> flag;
    int n;
};
struct ThreadLanes {
    LaneShared *sh;
    int id;
    int lane() const { return id; }
    int nlanes() const { return sh->n; }
    void sync() const { sh->bar.arrive_and_wait(); }
    double bcast(double v, int src) const
    {
        sh->xch[id] = v;
        sync();
        const double r = sh->xch[src];
        sync();
        return r;
    }
    bool all(bool b) const
    {
        sh->flag[id] = b ? 1 : 0;
        sync();
        bool r = true;
        for (int i = 0; i < sh->n; ++i)
            r = r && sh->flag[i] != 0;
        sync();
        return r;
    }
};
}  // namespace

extern "C" {

long trs_host_fit_lanes(const trs_host_params *hp, const double *start, packet_cb cb, void *ctx, double *state,
                        double *partrace, double *ssrtrace, double *condtrace, long max_packets, int nlanes)
{
    trs::Params P;
    P.p = hp->p; P.maxiter = hp->maxiter; P.trs = hp->trs; P.scale = hp->scale; P.trace = hp->trace;
    P.batch_iters = hp->batch_iters; P.cg_maxit = hp->cg_maxit; P.factor_up = hp->factor_up;
    P.factor_down = hp->factor_down; P.avmax = hp->avmax; P.h_df = hp->h_df; P.h_fvv = hp->h_fvv;
    P.xtol = hp->xtol; P.ftol = hp->ftol; P.gtol = hp->gtol; P.cg_tol = hp->cg_tol;
    const int p = P.p;
    if (nlanes != 32)
        return -1; // the lane-slot arithmetic of trs_core.h assumes a 32-wide warp
    std::vector<double> req(trs::request_doubles(p)), pk(trs::packet_doubles(p) + 2), jtj(p * p), work(p * p);
    trs::state_reset(state, req.data(), start, p);
    typedef trs::Solver<128, ThreadLanes> S;
    LaneShared shared(nlanes);
    std::vector<std::unique_ptr<S>> solver;
    for (int l = 0; l < nlanes; ++l)
        solver.emplace_back(new S(P, ThreadLanes{&shared, l}, jtj.data(), work.data()));
    long n = 0;
    while ((int)state[trs::S_PHASE] != trs::PH_DONE && n < max_packets) {
        const int mode = (int)req[0];
        if (mode == trs::MODE_IDLE)
            break;
        if (cb(ctx, mode, req.data() + 1, req.data() + 1 + p, pk.data()))
            break;
        std::vector<std::thread> th;
        for (int l = 0; l < nlanes; ++l)
            th.emplace_back([&, l] { solver[l]->advance(state, pk.data(), req.data(), partrace, ssrtrace, condtrace); });
        for (auto &t : th)
            t.join();
        ++n;
    }
    return n;
}
}

// ---- multi-start control logic (gslnls_b200/csrc/mstart.hpp) on the host ------------------------------
// The evaluator runs every candidate of a batch through the one-lane host build of the trust-region core
// (the arithmetic of the device's trs_step_batch), packets supplied by the caller's callback.
#include "../../gslnls_b200/csrc/mstart.hpp"

namespace {
struct HostBatchEvaluator {
    trs::Params P;
    packet_cb cb;
    void *ctx;
    void operator()(const std::vector<double> &starts, int S, int iters,
                    std::vector<gslnls::mstart::BatchResult> &out)
    {
        const int p = P.p;
        trs::Params Q = P;
        Q.maxiter = iters;
        Q.batch_iters = iters;
        Q.trace = 0;
        Q.gtol = 1.0e-3;
        out.assign(S, gslnls::mstart::BatchResult());
        std::vector<double> state(trs::state_doubles(p)), req(trs::request_doubles(p)), pk(trs::packet_doubles(p) + 2);
        std::vector<double> jtj(p * p), work(p * p);
        typedef trs::Solver<128, trs::SingleLane> S1;
        std::unique_ptr<S1> solver(new S1(Q, trs::SingleLane(), jtj.data(), work.data()));
        for (int c = 0; c < S; ++c) {
            trs::state_reset(state.data(), req.data(), starts.data() + (size_t)c * p, p);
            long guard = 0;
            while ((int)state[trs::S_PHASE] != trs::PH_DONE && guard++ < 100000) {
                const int mode = (int)req[0];
                if (mode == trs::MODE_IDLE || cb(ctx, mode, req.data() + 1, req.data() + 1 + p, pk.data()))
                    break;
                solver->advance(state.data(), pk.data(), req.data(), nullptr, nullptr, nullptr);
            }
            gslnls::mstart::BatchResult &b = out[c];
            const double *v = state.data() + trs::S_COUNT;
            b.par.assign(v, v + p);
            b.diag.assign(v + 3 * p, v + 4 * p);
            b.ssr = state[trs::S_CHISQ1];
            b.ssr_prev = state[trs::S_CHISQ0];
            b.ssr_start = state[trs::S_CHISQ_INIT];
            b.logdet_start = state[trs::S_LOGDET0];
            b.logdet_end = state[trs::S_LOGDET1];
            b.status = (int)state[trs::S_STATUS];
        }
    }
};
} // namespace

extern "C" {
// out: [p par | 2p range | ssr ssrconv nsp nwsp mstarts status searches]
int trs_host_multistart(const trs_host_params *hp, const double *range, const int *has_range, const int *mstart_int,
                        const double *mstart_dbl, packet_cb cb, void *ctx, double *out)
{
    HostBatchEvaluator ev;
    trs::Params &P = ev.P;
    P.p = hp->p; P.maxiter = hp->maxiter; P.trs = hp->trs; P.scale = hp->scale; P.trace = 0;
    P.batch_iters = hp->batch_iters; P.cg_maxit = hp->cg_maxit; P.factor_up = hp->factor_up;
    P.factor_down = hp->factor_down; P.avmax = hp->avmax; P.h_df = hp->h_df; P.h_fvv = hp->h_fvv;
    P.xtol = hp->xtol; P.ftol = hp->ftol; P.gtol = hp->gtol; P.cg_tol = hp->cg_tol;
    ev.cb = cb;
    ev.ctx = ctx;
    gslnls::mstart::Control c;
    c.n = mstart_int[0]; c.p = mstart_int[1]; c.q = mstart_int[2]; c.s = mstart_int[3];
    c.niter = mstart_int[4]; c.max = mstart_int[5]; c.minsp = mstart_int[6];
    c.r = mstart_dbl[0]; c.tol = mstart_dbl[1];
    gslnls::mstart::Driver<HostBatchEvaluator> drv(P.p, c, range, has_range, P.xtol, P.ftol, ev);
    const gslnls::mstart::Outcome o = drv.run();
    const int p = P.p;
    for (int k = 0; k < p; ++k)
        out[k] = o.par[k];
    for (int k = 0; k < 2 * p; ++k)
        out[p + k] = o.range[k];
    double *s = out + 3 * p;
    s[0] = o.ssr; s[1] = o.ssrconv; s[2] = o.nsp; s[3] = o.nwsp; s[4] = o.mstarts; s[5] = o.status; s[6] = (double)o.searches;
    return 0;
}

int trs_host_qrng(int dim, int count, double *out)
{
    if (dim < 41) {
        gslnls::mstart::Sobol g(dim);
        for (int i = 0; i < count; ++i)
            g.next(out + (size_t)i * dim);
    } else {
        gslnls::mstart::Halton g(dim);
        for (int i = 0; i < count; ++i)
            g.next(out + (size_t)i * dim);
    }
    return 0;
}
}
