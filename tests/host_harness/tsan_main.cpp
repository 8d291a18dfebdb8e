// tests/host_harness/tsan_main.cpp -- TEST INFRASTRUCTURE: the 32-lane host emulation of the trust-region
// warp (trs_host.cpp) under -fsanitize=thread.  A mildly nonlinear least-squares problem at p = 3, 12, 40, all
// six methods, one-lane and 32-lane runs side by side: prints one line per case; ThreadSanitizer reports any
// pair of lane accesses to the shared matrices that is not ordered by a sync() -- on the GPU that would be a
// missing __syncwarp().  Built and run by tests/test_trs_core_cpu.py.
#include "trs_host.cpp"
#include <cstdio>
#include <random>
struct Ctx { int n, p; std::vector<double> J, y; };
static int cb(void *c, int mode, const double *th, const double *v, double *pk)
{
    Ctx &C = *(Ctx *)c;
    const int n = C.n, p = C.p, npk = p * (p + 1) / 2;
    for (int e = 0; e < npk + p + 2; ++e) pk[e] = 0.0;
    if (mode == 1) {
        for (int i = 0; i < n; ++i) {
            double f = -C.y[i];
            for (int j = 0; j < p; ++j) f += C.J[i * p + j] * th[j] + 0.01 * th[j] * th[j] * C.J[i * p + j];
            int e = 0;
            for (int a = 0; a < p; ++a) {
                const double ja = C.J[i * p + a] * (1.0 + 0.02 * th[a]);
                for (int b = 0; b <= a; ++b, ++e) pk[e] += ja * C.J[i * p + b] * (1.0 + 0.02 * th[b]);
                pk[npk + a] += ja * f;
            }
            pk[npk + p] += f * f;
        }
    } else {
        for (int i = 0; i < n; ++i) {
            double h = 0.0;
            for (int j = 0; j < p; ++j) h += 0.02 * C.J[i * p + j] * v[j] * v[j];
            for (int a = 0; a < p; ++a) pk[a] += C.J[i * p + a] * (1.0 + 0.02 * th[a]) * h;
            pk[p] += h * h;
        }
    }
    return 0;
}
int main()
{
    for (int p : {3, 12, 40}) {
        Ctx C; C.n = 200; C.p = p; C.J.resize(C.n * p); C.y.resize(C.n);
        std::mt19937 g(p);
        std::normal_distribution<double> N(0, 1);
        for (auto &v : C.J) v = N(g);
        for (auto &v : C.y) v = N(g);
        for (int trs = 0; trs < 6; ++trs) {
            trs_host_params hp = {p, 20, trs, 0, 1, 0, 200, 2.0, 3.0, 0.75, 1.5e-8, 0.02, 1.5e-8, 1.5e-8, 1.5e-8, 1e-6};
            std::vector<double> start(p, 0.1), st1(trs::state_doubles(p)), st2(st1.size());
            std::vector<double> pt((21) * p), ss(21), cc(21), pt2(pt.size()), ss2(21), cc2(21);
            long a = trs_host_fit(&hp, start.data(), cb, &C, st1.data(), pt.data(), ss.data(), cc.data(), 1000);
            long b = trs_host_fit_lanes(&hp, start.data(), cb, &C, st2.data(), pt2.data(), ss2.data(), cc2.data(), 1000, 32);
            int same = a == b && !memcmp(st1.data(), st2.data(), st1.size() * 8);
            printf("p %d trs %d packets %ld %ld status %d niter %d bitwise %s\n", p, trs, a, b, (int)st1[trs::S_STATUS], (int)st1[trs::S_NITER], same ? "same" : "DIFFERENT");
        }
    }
}
