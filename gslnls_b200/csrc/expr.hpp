// expr.hpp -- model-formula compiler front half: R arithmetic parser, hash-consed expression
// DAG, symbolic differentiation (the job stats::deriv does at R/nls_large.R:297 and :334),
// and straight-line code generation for the device functions.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace gslnls {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

enum class Op : uint8_t { Const, Param, Var, Vel, Add, Sub, Mul, Div, Neg, Pow, Func };

enum Fn : int {
    F_EXP, F_LOG, F_LOG2, F_LOG10, F_LOG1P, F_EXPM1, F_SQRT, F_SIN, F_COS, F_TAN, F_ASIN, F_ACOS,
    F_ATAN, F_SINH, F_COSH, F_TANH, F_ABS, F_SIGN, F_PNORM, F_DNORM, F_SINPI, F_COSPI, F_COUNT
};

struct Node {
    Op op;
    int a = -1, b = -1; // children (or index for Param/Var/Vel, function id for Func in b)
    double c = 0.0;     // Const value
};

class Graph {
public:
    int cst(double v);
    int param(int j);
    int var(int k);
    int vel(int j);
    int add(int a, int b);
    int sub(int a, int b);
    int mul(int a, int b);
    int div(int a, int b);
    int neg(int a);
    int pow(int a, int b);
    int func(int fid, int a);

    int diff(int e, int wrt);  // d e / d theta_wrt
    int ddir(int e);           // sum_j vel_j d e / d theta_j

    const Node &at(int i) const { return nodes_[i]; }
    size_t size() const { return nodes_.size(); }
    bool is_const(int e, double v) const { return nodes_[e].op == Op::Const && nodes_[e].c == v; }
    bool is_const(int e) const { return nodes_[e].op == Op::Const; }
    int nparams = 0;

private:
    int intern(const Node &n);
    int dgeneric(int e, int wrt); // wrt >= 0: partial; wrt == -1: directional
    int dleaf(int j, int wrt) { return wrt >= 0 ? cst(j == wrt ? 1.0 : 0.0) : vel(j); }
    std::vector<Node> nodes_;
    std::map<std::tuple<int, int, int, uint64_t>, int> memo_;
    std::map<std::pair<int, int>, int> dmemo_;
};

// Parse an R arithmetic expression. Names found in `params` become Param nodes, names in `vars`
// Var nodes; anything else is an error.
int parse_rhs(Graph &g, const std::string &text, const std::vector<std::string> &params,
              const std::vector<std::string> &vars);

// Emit `const double t<i> = ...;` lines computing all `roots`; returns the C expression naming
// each root. `prefix` distinguishes temporaries of different functions.
std::string emit_block(const Graph &g, const std::vector<int> &roots, std::vector<std::string> &root_names,
                       const std::string &indent);

// Two-stage form of the same block: statements that do not depend on a predictor go to `prep`
// (run once per launch, results in c[0..nconst)), the per-observation rest to `body`.
struct SplitCode {
    std::string prep, body;
    int nconst = 0;
    std::vector<std::string> root_names;
};
SplitCode emit_split(Graph &g, const std::vector<int> &roots, const std::string &indent, bool sink_rows);

struct ModelSpec {
    std::string rhs;
    std::vector<std::string> params, vars;
    int jac_mode = 0, fvv_mode = 0;
};

// Full model source: device functions nls_model_f / nls_model_fj / nls_model_jfvv plus the
// GSLNLS_P / GSLNLS_NVAR / mode defines consumed by the kernel template.
std::string generate_model_source(const ModelSpec &spec);

} // namespace gslnls
