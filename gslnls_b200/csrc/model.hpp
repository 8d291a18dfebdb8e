// model.hpp -- compiled model object: generated source + NVRTC-built kernel variants.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "expr.hpp"

namespace gslnls {

struct KernelTune {
    int block = 256, unroll = 4, minb = 2;
    int tiled = 0; // 1: shared-memory J tiles + FP64 DMMA SYRK (nls_pass_tiled.cuh); 2: TMA-staged register kernel
    int stages = 4; // tiled == 2: tiles in flight per CTA
    int prefetch = 0; // register kernel: software-pipelined loads (next trip's loads before this trip's arithmetic)
    int fexp = 0;     // exp() of the model: 0 CUDA library, 1 branch-free table variant, 2 branch-free polynomial
};

struct VariantKey {
    int has_w, vec, stream, block, unroll, minb, tiled, stages, prefetch, fexp;
    int wgsl = 0; // weights as GSL's multilarge applies them: sqrt(w) on f and fvv only, J^T J unweighted
    bool operator<(const VariantKey &o) const
    {
        return std::tie(has_w, vec, stream, block, unroll, minb, tiled, stages, prefetch, fexp, wgsl) <
               std::tie(o.has_w, o.vec, o.stream, o.block, o.unroll, o.minb, o.tiled, o.stages, o.prefetch, o.fexp, o.wgsl);
    }
};

struct Variant {
    std::vector<char> cubin;
    std::string log;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t pass = nullptr, materialise = nullptr;
    cudaKernel_t irls_hist = nullptr, irls_above = nullptr, irls_weights = nullptr, irls_scale = nullptr;
    cudaKernel_t persistent = nullptr; // nls_pass_persistent (TMA-ring variants only): one launch per fit
    cudaKernel_t sparse_eval = nullptr; // nls_sparse_eval (symbolic Jacobian, p <= 16): term values + nonzeros of J
    bool loaded = false;
    size_t pass_smem = 0; // dynamic shared memory of one nls_pass CTA (tiled variant)
    std::vector<int> smem_devices; // devices on which the dynamic shared-memory limit has been raised
};

} // namespace gslnls

struct gslnls_model {
    gslnls::ModelSpec spec;
    std::string source; // generated device functions
    int p = 0, nvar = 0;
    std::map<gslnls::VariantKey, gslnls::Variant> variants;
    std::mutex mu, load_mu;

    // NVRTC-compile (if needed) the kernel variant; no device required. Throws std::runtime_error.
    gslnls::Variant &compile(const gslnls::VariantKey &key);
    // compile + load on the current device and resolve kernel handles
    gslnls::Variant &load(const gslnls::VariantKey &key);
    ~gslnls_model();
};

namespace gslnls {
// shard_bytes: bytes one pass reads on this GPU (0 = unknown / small); picks the load path of the p <= 4 kernel
// persistent: the fit will run the one-launch-per-fit kernel, which exists for the TMA-ring variant only
KernelTune default_tune(int p, double shard_bytes = 0.0, bool persistent = false);
size_t tiled_smem_bytes(int p, int block, int nprod, int nconst); // dynamic shared memory of the tiled pass kernel
int tiled_num_buffers(int p, int block, int nprod, int nconst);   // tile buffers of its CTA-wide ring (NT_NBUF)
size_t tma_smem_bytes(int narr, int block, int unroll, int stages); // dynamic shared memory of the TMA-staged pass kernel
std::string nvrtc_arch_for_device(int device); // "sm_100a" on B200; used as --gpu-architecture
void cache_drop(const gslnls_model *m); // problem.cu: release one-shot problems cached for m (nullptr: all)
} // namespace gslnls
