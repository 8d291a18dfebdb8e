// capi.cpp -- model compilation entry points and small utilities of the C ABI.
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/gslnls_b200.h"
#include "model.hpp"

namespace gslnls {
extern thread_local std::string g_last_error;
void set_error(const std::string &s);
} // namespace gslnls
using namespace gslnls;

extern "C" {

GSLNLS_API int gslnls_model_compile(const char *rhs_expr, const char *const *param_names, int p,
                                    const char *const *var_names, int nvar, int jac_mode, int fvv_mode,
                                    gslnls_model **out, char *errbuf, size_t errlen)
{
    auto fail = [&](int code, const std::string &msg) {
        set_error(msg);
        if (errbuf && errlen) {
            std::snprintf(errbuf, errlen, "%s", msg.c_str());
        }
        return code;
    };
    if (!rhs_expr || !out || p < 1 || (p > 0 && !param_names) || (nvar > 0 && !var_names))
        return fail(GSLNLS_EINVAL, "invalid argument");
    if (jac_mode < 0 || jac_mode > 2 || fvv_mode < 0 || fvv_mode > 2)
        return fail(GSLNLS_EINVAL, "invalid jac_mode / fvv_mode");
    *out = nullptr;
    gslnls_model *m = new gslnls_model();
    m->spec.rhs = rhs_expr;
    for (int j = 0; j < p; ++j)
        m->spec.params.emplace_back(param_names[j]);
    for (int k = 0; k < nvar; ++k)
        m->spec.vars.emplace_back(var_names[k]);
    m->spec.jac_mode = jac_mode;
    m->spec.fvv_mode = fvv_mode;
    m->p = p;
    m->nvar = nvar;
    try {
        m->source = generate_model_source(m->spec); // R/nls_large.R:297,334: deriv() equivalents
    } catch (const std::exception &e) {
        delete m;
        return fail(GSLNLS_EPARSE, e.what());
    }
    try {
        // compile the default variant now so that NVRTC errors surface at model-build time,
        // like stop("failed to symbolically derive 'jac'") does at R/nls_large.R:298-299
        const KernelTune t = default_tune(p);
        m->compile(VariantKey{0, 2, 1, t.block, t.unroll, t.minb, t.tiled, t.stages, t.prefetch, t.fexp});
    } catch (const std::exception &e) {
        delete m;
        return fail(GSLNLS_ECOMPILE, e.what());
    }
    *out = m;
    return GSLNLS_SUCCESS;
}

GSLNLS_API void gslnls_model_free(gslnls_model *m)
{
    if (m)
        gslnls::cache_drop(m); // cached one-shot problems hold kernels of this model
    delete m;
}
GSLNLS_API int gslnls_model_p(const gslnls_model *m) { return m ? m->p : 0; }
GSLNLS_API int gslnls_model_nvar(const gslnls_model *m) { return m ? m->nvar : 0; }
GSLNLS_API const char *gslnls_model_source(const gslnls_model *m) { return m ? m->source.c_str() : ""; }

GSLNLS_API const char *gslnls_strerror(int code)
{
    switch (code) { // gsl_strerror() texts for the GSL codes
    case GSLNLS_SUCCESS: return "success";
    case GSLNLS_FAILURE: return "failure";
    case GSLNLS_CONTINUE: return "the iteration has not converged yet";
    case GSLNLS_EDOM: return "input domain error";
    case GSLNLS_EINVAL: return "invalid argument supplied by user";
    case GSLNLS_ENOMEM: return "malloc failed";
    case GSLNLS_EBADFUNC: return "problem with user-supplied function";
    case GSLNLS_EMAXITER: return "exceeded max number of iterations";
    case GSLNLS_EBADLEN: return "matrix, vector lengths are not conformant";
    case GSLNLS_ENOPROG: return "iteration is not making progress towards solution";
    case GSLNLS_ETOLF: return "cannot reach the specified tolerance in F";
    case GSLNLS_ETOLX: return "cannot reach the specified tolerance in X";
    case GSLNLS_ETOLG: return "cannot reach the specified tolerance in gradient";
    case GSLNLS_EPARSE: return "model formula cannot be translated to device code";
    case GSLNLS_ECOMPILE: return "NVRTC compilation of the model kernel failed";
    case GSLNLS_ECUDA: return "CUDA runtime error";
    case GSLNLS_ENODEVICE: return "no usable CUDA device (no CPU fallback exists)";
    case GSLNLS_ECOMM: return "multi-GPU exchange failed";
    default: return "unknown error code";
    }
}

GSLNLS_API const char *gslnls_trs_name(int algorithm)
{
    switch (algorithm) { // gsl_multilarge_nlinear_trs_name(); README.md:595,649 confirm the first two
    case 1: return "levenberg-marquardt+accel";
    case 2: return "dogleg";
    case 3: return "double-dogleg";
    case 4: return "2D-subspace";
    case 5: return "steihaug-toint";
    default: return "levenberg-marquardt";
    }
}

GSLNLS_API const char *gslnls_last_error(void) { return g_last_error.c_str(); }
GSLNLS_API const char *gslnls_version(void) { return "gslnls_b200 0.1.0 (sm_100a)"; }

} // extern "C"
