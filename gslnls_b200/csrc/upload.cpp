// upload.cpp -- host columns -> HBM at PCIe rate, from PAGEABLE memory.
//
// The reference boundary hands the library plain R vectors (REAL(y), src/nls_large.c:66-75): pageable host
// memory.  cudaMemcpyAsync from pageable memory is staged by the driver through one thread and one small
// bounce buffer (10-25 GB/s measured), and cudaHostRegister of gigabytes costs more than the copy it would
// speed up.  So the library pins for the caller: a per-device pool of worker threads, each with its own pair
// of pinned staging buffers and its own stream, copies slices of the columns into pinned memory and enqueues
// the DMA; the memcpy of slice k+1 overlaps the DMA of slice k, and the threads together out-run the PCIe
// link (Gen5 x16: ~55 GB/s).  Already pinned / registered / managed source memory skips the staging.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gslnls_b200.h"
#include "upload.hpp"

namespace gslnls {
void set_error(const std::string &s);

namespace {
constexpr size_t kSlice = (size_t)4 << 20; // bytes per staged slice
constexpr size_t kDirectBelow = (size_t)1 << 20; // total bytes below which the driver's own staging is as good

struct Lane { // one worker thread's resources on one device
    void *pin[2] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
};
struct DevicePool {
    int device = -1;
    std::vector<Lane> lanes;
    std::mutex mu; // one staged upload at a time per device
};
std::mutex g_pools_mu;
std::vector<DevicePool *> g_pools;

DevicePool *pool_for(int device)
{
    std::lock_guard<std::mutex> lk(g_pools_mu);
    for (DevicePool *p : g_pools)
        if (p->device == device)
            return p;
    DevicePool *p = new DevicePool();
    p->device = device;
    g_pools.push_back(p);
    return p;
}

bool is_pageable(const void *ptr)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
} // namespace

int upload_threads_default(int sharing)
{
    if (const char *c = std::getenv("GSLNLS_UPLOAD_THREADS"))
        return std::max(1, std::atoi(c));
    // `sharing` uploads run side by side (one per GPU of a multi-GPU call) and share the host cores; so do the
    // sibling ranks of a one-process-per-GPU job (torchrun / mpirun export their count)
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    int siblings = 1;
    for (const char *name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS"})
        if (const char *c = std::getenv(name)) {
            siblings = std::max(1, std::atoi(c));
            break;
        }
    return std::max(1, std::min(8, hw / std::max(1, sharing * siblings)));
}

// copy `ncol` host columns of `bytes` bytes each to their device buffers; returns a GSLNLS code.
// On success every copy has been enqueued AND `done` has been made to wait for them (the caller's stream).
int staged_upload(int device, const void *const *src, void *const *dst, int ncol, size_t bytes, cudaStream_t done,
                  int nthreads)
{
    if (ncol <= 0 || bytes == 0)
        return GSLNLS_SUCCESS;
    auto fail = [](const char *what, cudaError_t e) {
        set_error(std::string(what) + ": " + cudaGetErrorString(e));
        return (int)GSLNLS_ECUDA;
    };
    // columns that are already DMA-able (pinned, registered, managed, device) go straight to the copy engine
    std::vector<int> staged;
    for (int c = 0; c < ncol; ++c) {
        if ((size_t)ncol * bytes >= kDirectBelow && is_pageable(src[c])) {
            staged.push_back(c);
        } else {
            cudaError_t e = cudaMemcpyAsync(dst[c], src[c], bytes, cudaMemcpyDefault, done);
            if (e != cudaSuccess)
                return fail("cudaMemcpyAsync(upload)", e);
        }
    }
    if (staged.empty())
        return GSLNLS_SUCCESS;

    DevicePool *pool = pool_for(device);
    std::lock_guard<std::mutex> lk(pool->mu);
    const size_t nslice_col = (bytes + kSlice - 1) / kSlice;
    const size_t nslice = nslice_col * staged.size();
    const int T = (int)std::min<size_t>((size_t)std::max(1, nthreads), nslice);
    while ((int)pool->lanes.size() < T) {
        Lane ln;
        cudaError_t e = cudaSetDevice(device);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
            e = cudaHostAlloc(&ln.pin[b], kSlice, cudaHostAllocDefault);
            if (e == cudaSuccess)
                e = cudaEventCreateWithFlags(&ln.ev[b], cudaEventDisableTiming);
        }
        if (e == cudaSuccess)
            e = cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess)
            return fail("staging buffers", e);
        pool->lanes.push_back(ln);
    }
    // the staged DMAs overwrite the device columns: order them after whatever the caller's stream still has
    // queued on those buffers (lane 0's first event doubles as the fork marker)
    {
        cudaError_t e = cudaEventRecord(pool->lanes[0].ev[0], done);
        for (int t = 0; t < T && e == cudaSuccess; ++t)
            e = cudaStreamWaitEvent(pool->lanes[t].stream, pool->lanes[0].ev[0], 0);
        if (e != cudaSuccess)
            return fail("staged upload (fork)", e);
    }
    std::atomic<size_t> next{0};
    std::atomic<int> err{(int)cudaSuccess};
    auto work = [&](int t) {
        Lane &ln = pool->lanes[t];
        if (cudaSetDevice(device) != cudaSuccess)
            return;
        int b = 0;
        bool used[2] = {false, false};
        for (;;) {
            const size_t s = next.fetch_add(1); // slices are handed out in address order: sequential host reads
            if (s >= nslice || err.load() != (int)cudaSuccess)
                break;
            const int c = staged[s / nslice_col];
            const size_t off = (s % nslice_col) * kSlice, len = std::min(kSlice, bytes - off);
            cudaError_t e = used[b] ? cudaEventSynchronize(ln.ev[b]) : cudaSuccess; // DMA out of this buffer finished?
            if (e == cudaSuccess) {
                std::memcpy(ln.pin[b], (const char *)src[c] + off, len);
                e = cudaMemcpyAsync((char *)dst[c] + off, ln.pin[b], len, cudaMemcpyHostToDevice, ln.stream);
            }
            if (e == cudaSuccess)
                e = cudaEventRecord(ln.ev[b], ln.stream);
            if (e != cudaSuccess) {
                err.store((int)e);
                break;
            }
            used[b] = true;
            b ^= 1;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t)
        th.emplace_back(work, t);
    work(0);
    for (std::thread &t : th)
        t.join();
    if (err.load() != (int)cudaSuccess)
        return fail("staged upload", (cudaError_t)err.load());
    // the caller's stream continues once every lane's last DMA has landed
    for (int t = 0; t < T; ++t) {
        Lane &ln = pool->lanes[t];
        cudaError_t e = cudaEventRecord(ln.ev[0], ln.stream);
        if (e == cudaSuccess)
            e = cudaStreamWaitEvent(done, ln.ev[0], 0);
        if (e != cudaSuccess)
            return fail("staged upload (join)", e);
    }
    return GSLNLS_SUCCESS;
}

void upload_pools_release()
{
    std::lock_guard<std::mutex> lk(g_pools_mu);
    for (DevicePool *p : g_pools) {
        std::lock_guard<std::mutex> l2(p->mu);
        cudaSetDevice(p->device);
        for (Lane &ln : p->lanes) {
            cudaStreamSynchronize(ln.stream);
            cudaFreeHost(ln.pin[0]);
            cudaFreeHost(ln.pin[1]);
            cudaEventDestroy(ln.ev[0]);
            cudaEventDestroy(ln.ev[1]);
            cudaStreamDestroy(ln.stream);
        }
        p->lanes.clear();
    }
}

} // namespace gslnls
