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
};

struct VariantKey {
    int has_w, vec, stream, block, unroll, minb;
    bool operator<(const VariantKey &o) const
    {
        return std::tie(has_w, vec, stream, block, unroll, minb) <
               std::tie(o.has_w, o.vec, o.stream, o.block, o.unroll, o.minb);
    }
};

struct Variant {
    std::vector<char> cubin;
    std::string log;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t pass = nullptr, materialise = nullptr;
    bool loaded = false;
};

} // namespace gslnls

struct gslnls_model {
    gslnls::ModelSpec spec;
    std::string source; // generated device functions
    int p = 0, nvar = 0;
    std::map<gslnls::VariantKey, gslnls::Variant> variants;
    std::mutex mu;

    // NVRTC-compile (if needed) the kernel variant; no device required. Throws std::runtime_error.
    gslnls::Variant &compile(const gslnls::VariantKey &key);
    // compile + load on the current device and resolve kernel handles
    gslnls::Variant &load(const gslnls::VariantKey &key);
    ~gslnls_model();
};

namespace gslnls {
KernelTune default_tune(int p);
std::string nvrtc_arch_for_device(int device); // "sm_100a" on B200; used as --gpu-architecture
} // namespace gslnls
