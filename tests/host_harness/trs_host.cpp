// tests/host_harness/trs_host.cpp -- TEST INFRASTRUCTURE: compiles the device trust-region state
// machine (gslnls_b200/csrc/trs_core.h) for the host with the SingleLane policy so that the CPU
// test-suite can step it against the oracle without a GPU.  Never linked into libgslnls_b200.so.
#include <cstdlib>
#include <cstring>
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
}
