// seg_build.hpp -- host side of the sparse-row path: the gather lists of the segmented sums (csrc/sparse.cu).
// Entries (terms, or stored nonzeros) are grouped by segment (row, or column) with a STABLE counting sort, each
// segment is cut into items of <= item_cap entries, and every item is classified: consecutive run (streamed
// without index lists), short (one thread), long segment (a CTA adds its item sums).  Plain C++ with std::thread
// so that tests/host_harness can build and check it without a GPU: the threaded sort must reproduce the serial one.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

namespace gslnls {

inline int seg_threads(long long work)
{
    const unsigned hw = std::thread::hardware_concurrency();
    const long long t = std::min<long long>(std::min<long long>(hw ? hw : 1, 16), work / (1 << 18) + 1);
    return (int)std::max<long long>(1, t);
}

// fn(begin, end, thread) over [0, n) in nthreads contiguous chunks
inline void seg_parallel(long long n, int nthreads, const std::function<void(long long, long long, int)> &fn)
{
    if (nthreads <= 1 || n < 2) {
        fn(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    for (int c = 0; c < nthreads; ++c)
        th.emplace_back([&, c] { fn(n * c / nthreads, n * (c + 1) / nthreads, c); });
    for (auto &t : th)
        t.join();
}

// order[] = indices 0..n-1 sorted by keys[] (stable), ptr[s] = first position of segment s.  keys in [0, nseg).
inline void seg_group(const int *keys, long long n, long long nseg, std::vector<long long> &ptr, int *order, int nthreads)
{
    // per-thread histograms cost nseg counters each: fewer threads when there are very many segments
    nthreads = (int)std::max<long long>(1, std::min<long long>(nthreads, (64ll << 20) / std::max<long long>(nseg, 1)));
    if (n < 2 * (long long)nthreads)
        nthreads = 1;
    std::vector<std::vector<long long>> off((size_t)nthreads);
    seg_parallel(n, nthreads, [&](long long a, long long b, int c) {
        off[(size_t)c].assign((size_t)nseg, 0);
        long long *h = off[(size_t)c].data();
        for (long long i = a; i < b; ++i)
            ++h[keys[i]];
    });
    ptr.assign((size_t)nseg + 1, 0);
    long long run = 0;
    for (long long s = 0; s < nseg; ++s) {
        ptr[(size_t)s] = run;
        for (int c = 0; c < nthreads; ++c) {
            const long long cnt = off[(size_t)c][(size_t)s];
            off[(size_t)c][(size_t)s] = run; // where thread c writes its first entry of segment s
            run += cnt;
        }
    }
    ptr[(size_t)nseg] = run;
    seg_parallel(n, nthreads, [&](long long a, long long b, int c) {
        long long *o = off[(size_t)c].data();
        for (long long i = a; i < b; ++i)
            order[o[keys[i]]++] = (int)i;
    });
}

struct SegLists {
    std::vector<int> ent_a, ent_b;      // per sorted entry: gather index / row (columns only)
    std::vector<long long> item_begin;  // [nitems + 1]
    std::vector<int> seg_itemptr;       // [nseg + 1]
    std::vector<int> item_a0, item_b0;  // first ent_a / ent_b of a consecutive item, else -1
    std::vector<int> long_seg;          // segments of more than long_items items
    std::vector<int> wide_item;         // items of more than short_entries entries
    int nshort = 0;
};

// items and their classes from the segment pointers and the sorted entry lists
inline void seg_items(SegLists &B, const std::vector<long long> &ptr, int item_cap, int long_items, int short_entries,
                      int nthreads)
{
    const size_t nseg = ptr.size() - 1;
    B.seg_itemptr.assign(nseg + 1, 0);
    size_t nitems = 0;
    for (size_t s = 0; s < nseg; ++s) {
        B.seg_itemptr[s] = (int)nitems;
        nitems += (size_t)((ptr[s + 1] - ptr[s] + item_cap - 1) / item_cap);
    }
    B.seg_itemptr[nseg] = (int)nitems;
    B.item_begin.resize(nitems + 1);
    B.long_seg.clear();
    for (size_t s = 0; s < nseg; ++s) {
        size_t it = (size_t)B.seg_itemptr[s];
        for (long long a = ptr[s]; a < ptr[s + 1]; a += item_cap)
            B.item_begin[it++] = a;
        if (B.seg_itemptr[s + 1] - B.seg_itemptr[s] > long_items)
            B.long_seg.push_back((int)s);
    }
    B.item_begin[nitems] = ptr[nseg];
    B.item_a0.assign(nitems, -1);
    B.item_b0.assign(nitems, -1);
    const bool has_b = !B.ent_b.empty();
    seg_parallel((long long)nitems, nthreads, [&](long long i0, long long i1, int) {
        for (long long it = i0; it < i1; ++it) {
            const long long a = B.item_begin[(size_t)it], b = B.item_begin[(size_t)it + 1];
            bool run = b > a;
            for (long long e = a + 1; e < b && run; ++e)
                run = B.ent_a[(size_t)e] == B.ent_a[(size_t)e - 1] + 1 &&
                      (!has_b || B.ent_b[(size_t)e] == B.ent_b[(size_t)e - 1] + 1);
            if (run) {
                B.item_a0[(size_t)it] = B.ent_a[(size_t)a];
                B.item_b0[(size_t)it] = has_b ? B.ent_b[(size_t)a] : 0;
            }
        }
    });
    B.wide_item.clear();
    for (size_t it = 0; it < nitems; ++it)
        if (B.item_begin[it + 1] - B.item_begin[it] > short_entries)
            B.wide_item.push_back((int)it);
    B.nshort = (int)(nitems - B.wide_item.size());
}

} // namespace gslnls
