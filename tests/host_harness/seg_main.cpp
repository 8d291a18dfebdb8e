// TEST INFRASTRUCTURE: host check of gslnls_b200/csrc/seg_build.hpp -- the threaded stable grouping must reproduce
// the serial counting sort entry for entry, and the item classes must follow their definitions.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../gslnls_b200/csrc/seg_build.hpp"
using namespace gslnls;

static int check(long long n, long long nseg, int nthreads, unsigned seed, int pattern)
{
    std::mt19937_64 rng(seed);
    std::vector<int> keys((size_t)n);
    for (long long i = 0; i < n; ++i)
        keys[(size_t)i] = pattern == 0 ? (int)(rng() % (unsigned long long)nseg)   // random
                          : pattern == 1 ? (int)(i * nseg / n)                      // sorted runs
                                         : (int)(i % nseg);                         // interleaved
    std::vector<long long> p1, p2;
    std::vector<int> o1((size_t)n), o2((size_t)n);
    seg_group(keys.data(), n, nseg, p1, o1.data(), 1);
    seg_group(keys.data(), n, nseg, p2, o2.data(), nthreads);
    if (p1 != p2 || o1 != o2)
        return 1;
    for (long long s = 0; s < nseg; ++s)
        for (long long e = p1[(size_t)s]; e < p1[(size_t)s + 1]; ++e) {
            if (keys[(size_t)o1[(size_t)e]] != s)
                return 2;
            if (e > p1[(size_t)s] && o1[(size_t)e] <= o1[(size_t)e - 1])
                return 3; // stable: original order inside a segment
        }
    SegLists B;
    B.ent_a = o1;
    seg_items(B, p1, 2048, 64, 8, nthreads);
    const size_t ni = B.item_begin.size() - 1;
    size_t wide = 0;
    for (size_t it = 0; it < ni; ++it) {
        const long long a = B.item_begin[it], b = B.item_begin[it + 1];
        if (b <= a || b - a > 2048)
            return 4;
        bool run = true;
        for (long long e = a + 1; e < b; ++e)
            run = run && B.ent_a[(size_t)e] == B.ent_a[(size_t)e - 1] + 1;
        if ((B.item_a0[it] >= 0) != run || (run && B.item_a0[it] != B.ent_a[(size_t)a]))
            return 5;
        wide += (b - a > 8);
    }
    if (wide != B.wide_item.size() || (size_t)B.nshort != ni - wide)
        return 6;
    for (int s : B.long_seg)
        if (B.seg_itemptr[(size_t)s + 1] - B.seg_itemptr[(size_t)s] <= 64)
            return 7;
    return 0;
}

int main()
{
    int rc = 0;
    const long long cases[][2] = {{1, 1}, {10, 3}, {1000, 1000}, {100000, 7}, {300000, 40000}, {2000000, 3}, {500000, 1}};
    for (auto &c : cases)
        for (int pattern = 0; pattern < 3; ++pattern)
            for (int nt : {2, 5, 16}) {
                const int r = check(c[0], c[1], nt, 7u + (unsigned)pattern, pattern);
                if (r) {
                    std::printf("FAIL n=%lld nseg=%lld threads=%d pattern=%d code=%d\n", c[0], c[1], nt, pattern, r);
                    rc = 1;
                }
            }
    std::printf(rc ? "seg_build: FAILED\n" : "seg_build: ok\n");
    return rc;
}
