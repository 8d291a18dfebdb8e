// upload.hpp -- host columns -> HBM from pageable memory through a pinned, multi-threaded staging ring
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace gslnls {
// worker threads of one staged upload when `sharing` uploads run side by side (GSLNLS_UPLOAD_THREADS overrides)
int upload_threads_default(int sharing);
// enqueue the copy of ncol host columns (bytes each) to their device buffers; `done` waits for all of them
int staged_upload(int device, const void *const *src, void *const *dst, int ncol, size_t bytes, cudaStream_t done,
                  int nthreads);
void upload_pools_release(); // pinned staging memory back to the system (gslnls_cache_clear)
} // namespace gslnls
