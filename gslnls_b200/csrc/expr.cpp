// expr.cpp -- see expr.hpp.  Host-side C++ only (no CUDA); part of libgslnls_b200.so.
#include "expr.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <sstream>

namespace gslnls {

// ------------------------------------------------------------------------------------------
// hash-consed DAG with local simplification
// ------------------------------------------------------------------------------------------

static uint64_t dbits(double v)
{
    uint64_t u;
    std::memcpy(&u, &v, 8);
    return u;
}

int Graph::intern(const Node &n)
{
    auto key = std::make_tuple(static_cast<int>(n.op), n.a, n.b, dbits(n.c));
    auto it = memo_.find(key);
    if (it != memo_.end())
        return it->second;
    nodes_.push_back(n);
    const int id = static_cast<int>(nodes_.size()) - 1;
    memo_.emplace(key, id);
    return id;
}

int Graph::cst(double v)
{
    if (v == 0.0)
        v = 0.0; // fold -0.0
    Node n{Op::Const};
    n.c = v;
    return intern(n);
}
int Graph::param(int j) { Node n{Op::Param}; n.a = j; return intern(n); }
int Graph::var(int k) { Node n{Op::Var}; n.a = k; return intern(n); }
int Graph::vel(int j) { Node n{Op::Vel}; n.a = j; return intern(n); }

int Graph::neg(int a)
{
    const Node &A = nodes_[a];
    if (A.op == Op::Const)
        return cst(-A.c);
    if (A.op == Op::Neg)
        return A.a;
    Node n{Op::Neg};
    n.a = a;
    return intern(n);
}

int Graph::add(int a, int b)
{
    const Node A = nodes_[a], B = nodes_[b];
    if (A.op == Op::Const && B.op == Op::Const)
        return cst(A.c + B.c);
    if (is_const(a, 0.0))
        return b;
    if (is_const(b, 0.0))
        return a;
    if (B.op == Op::Neg)
        return sub(a, B.a);
    if (A.op == Op::Neg)
        return sub(b, A.a);
    Node n{Op::Add};
    n.a = a < b ? a : b;
    n.b = a < b ? b : a;
    return intern(n);
}

int Graph::sub(int a, int b)
{
    const Node A = nodes_[a], B = nodes_[b];
    if (A.op == Op::Const && B.op == Op::Const)
        return cst(A.c - B.c);
    if (is_const(b, 0.0))
        return a;
    if (is_const(a, 0.0))
        return neg(b);
    if (a == b)
        return cst(0.0);
    if (B.op == Op::Neg)
        return add(a, B.a);
    Node n{Op::Sub};
    n.a = a;
    n.b = b;
    return intern(n);
}

int Graph::mul(int a, int b)
{
    const Node A = nodes_[a], B = nodes_[b];
    if (A.op == Op::Const && B.op == Op::Const)
        return cst(A.c * B.c);
    if (is_const(a, 0.0) || is_const(b, 0.0))
        return cst(0.0);
    if (is_const(a, 1.0))
        return b;
    if (is_const(b, 1.0))
        return a;
    if (is_const(a, -1.0))
        return neg(b);
    if (is_const(b, -1.0))
        return neg(a);
    if (A.op == Op::Neg && B.op == Op::Neg)
        return mul(A.a, B.a);
    if (A.op == Op::Neg)
        return neg(mul(A.a, b));
    if (B.op == Op::Neg)
        return neg(mul(a, B.a));
    Node n{Op::Mul};
    n.a = a < b ? a : b;
    n.b = a < b ? b : a;
    return intern(n);
}

int Graph::div(int a, int b)
{
    const Node A = nodes_[a], B = nodes_[b];
    if (A.op == Op::Const && B.op == Op::Const)
        return cst(A.c / B.c);
    if (is_const(a, 0.0))
        return cst(0.0);
    if (is_const(b, 1.0))
        return a;
    if (is_const(b, -1.0))
        return neg(a);
    if (A.op == Op::Neg && B.op == Op::Neg)
        return div(A.a, B.a);
    if (A.op == Op::Neg)
        return neg(div(A.a, b));
    if (B.op == Op::Neg)
        return neg(div(a, B.a));
    Node n{Op::Div};
    n.a = a;
    n.b = b;
    return intern(n);
}

int Graph::pow(int a, int b)
{
    const Node A = nodes_[a], B = nodes_[b];
    if (A.op == Op::Const && B.op == Op::Const)
        return cst(std::pow(A.c, B.c));
    if (is_const(b, 0.0))
        return cst(1.0);
    if (is_const(b, 1.0))
        return a;
    Node n{Op::Pow};
    n.a = a;
    n.b = b;
    return intern(n);
}

static double host_fn(int fid, double x)
{
    switch (fid) {
    case F_EXP: return std::exp(x);
    case F_LOG: return std::log(x);
    case F_LOG2: return std::log2(x);
    case F_LOG10: return std::log10(x);
    case F_LOG1P: return std::log1p(x);
    case F_EXPM1: return std::expm1(x);
    case F_SQRT: return std::sqrt(x);
    case F_SIN: return std::sin(x);
    case F_COS: return std::cos(x);
    case F_TAN: return std::tan(x);
    case F_ASIN: return std::asin(x);
    case F_ACOS: return std::acos(x);
    case F_ATAN: return std::atan(x);
    case F_SINH: return std::sinh(x);
    case F_COSH: return std::cosh(x);
    case F_TANH: return std::tanh(x);
    case F_ABS: return std::fabs(x);
    case F_SIGN: return (x > 0.0) - (x < 0.0);
    case F_PNORM: return 0.5 * std::erfc(-x * 0.70710678118654752440);
    case F_DNORM: return std::exp(-0.5 * x * x) * 0.39894228040143267794;
    case F_SINPI: return std::sin(3.14159265358979323846 * x);
    case F_COSPI: return std::cos(3.14159265358979323846 * x);
    default: return NAN;
    }
}

int Graph::func(int fid, int a)
{
    if (nodes_[a].op == Op::Const)
        return cst(host_fn(fid, nodes_[a].c));
    Node n{Op::Func};
    n.a = a;
    n.b = fid;
    return intern(n);
}

// ------------------------------------------------------------------------------------------
// differentiation (partial wrt one parameter, or directional along v)
// ------------------------------------------------------------------------------------------

int Graph::diff(int e, int wrt) { return dgeneric(e, wrt); }
int Graph::ddir(int e) { return dgeneric(e, -1); }

int Graph::dgeneric(int e, int wrt)
{
    auto key = std::make_pair(e, wrt);
    auto it = dmemo_.find(key);
    if (it != dmemo_.end())
        return it->second;
    const Node N = nodes_[e];
    int r = -1;
    switch (N.op) {
    case Op::Const:
    case Op::Var:
    case Op::Vel: r = cst(0.0); break;
    case Op::Param: r = dleaf(N.a, wrt); break;
    case Op::Add: r = add(dgeneric(N.a, wrt), dgeneric(N.b, wrt)); break;
    case Op::Sub: r = sub(dgeneric(N.a, wrt), dgeneric(N.b, wrt)); break;
    case Op::Neg: r = neg(dgeneric(N.a, wrt)); break;
    case Op::Mul: {
        const int da = dgeneric(N.a, wrt), db = dgeneric(N.b, wrt);
        r = add(mul(da, N.b), mul(N.a, db));
        break;
    }
    case Op::Div: {
        const int da = dgeneric(N.a, wrt), db = dgeneric(N.b, wrt);
        // da/b - a*db/b^2  (the form deriv() produces)
        const int t1 = div(da, N.b);
        const int t2 = is_const(db, 0.0) ? cst(0.0) : div(mul(N.a, db), mul(N.b, N.b));
        r = sub(t1, t2);
        break;
    }
    case Op::Pow: {
        const int da = dgeneric(N.a, wrt), db = dgeneric(N.b, wrt);
        int t1 = cst(0.0), t2 = cst(0.0);
        if (!is_const(da, 0.0)) {
            if (is_const(N.b)) {
                const double c = nodes_[N.b].c;
                if (c == 2.0)
                    t1 = mul(mul(cst(2.0), N.a), da);
                else
                    t1 = mul(mul(cst(c), pow(N.a, cst(c - 1.0))), da);
            } else {
                t1 = mul(mul(pow(N.a, sub(N.b, cst(1.0))), N.b), da);
            }
        }
        if (!is_const(db, 0.0))
            t2 = mul(mul(e, func(F_LOG, N.a)), db);
        r = add(t1, t2);
        break;
    }
    case Op::Func: {
        const int da = dgeneric(N.a, wrt);
        if (is_const(da, 0.0)) {
            r = cst(0.0);
            break;
        }
        const int a = N.a;
        switch (N.b) {
        case F_EXP: r = mul(e, da); break;
        case F_LOG: r = div(da, a); break;
        case F_LOG2: r = div(da, mul(a, cst(0.69314718055994530942))); break;
        case F_LOG10: r = div(da, mul(a, cst(2.30258509299404568402))); break;
        case F_LOG1P: r = div(da, add(cst(1.0), a)); break;
        case F_EXPM1: r = mul(func(F_EXP, a), da); break;
        case F_SQRT: r = div(da, mul(cst(2.0), e)); break;
        case F_SIN: r = mul(func(F_COS, a), da); break;
        case F_COS: r = neg(mul(func(F_SIN, a), da)); break;
        case F_TAN: { const int c = func(F_COS, a); r = div(da, mul(c, c)); break; }
        case F_ASIN: r = div(da, func(F_SQRT, sub(cst(1.0), mul(a, a)))); break;
        case F_ACOS: r = neg(div(da, func(F_SQRT, sub(cst(1.0), mul(a, a))))); break;
        case F_ATAN: r = div(da, add(cst(1.0), mul(a, a))); break;
        case F_SINH: r = mul(func(F_COSH, a), da); break;
        case F_COSH: r = mul(func(F_SINH, a), da); break;
        case F_TANH: r = mul(sub(cst(1.0), mul(e, e)), da); break;
        case F_ABS: r = mul(func(F_SIGN, a), da); break;
        case F_SIGN: r = cst(0.0); break;
        case F_PNORM: r = mul(func(F_DNORM, a), da); break;
        case F_DNORM: r = neg(mul(mul(a, e), da)); break;
        case F_SINPI: r = mul(mul(cst(3.14159265358979323846), func(F_COSPI, a)), da); break;
        case F_COSPI: r = neg(mul(mul(cst(3.14159265358979323846), func(F_SINPI, a)), da)); break;
        default: throw ParseError("internal: derivative of unknown function");
        }
        break;
    }
    }
    dmemo_.emplace(key, r);
    return r;
}

// ------------------------------------------------------------------------------------------
// parser for R arithmetic
// ------------------------------------------------------------------------------------------

namespace {

struct Parser {
    Graph &g;
    const std::string &s;
    const std::vector<std::string> &params, &vars;
    size_t i = 0;

    [[noreturn]] void fail(const std::string &msg) const
    {
        std::ostringstream os;
        os << msg << " at position " << i << " of '" << s << "'";
        throw ParseError(os.str());
    }
    void ws()
    {
        while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i])))
            ++i;
    }
    bool eat(char c)
    {
        ws();
        if (i < s.size() && s[i] == c) {
            ++i;
            return true;
        }
        return false;
    }
    char peek()
    {
        ws();
        return i < s.size() ? s[i] : '\0';
    }

    int expr()
    {
        int l = term();
        for (;;) {
            if (eat('+'))
                l = g.add(l, term());
            else if (eat('-'))
                l = g.sub(l, term());
            else
                return l;
        }
    }
    int term()
    {
        int l = unary();
        for (;;) {
            ws();
            if (i + 1 < s.size() && s[i] == '*' && s[i + 1] == '*')
                fail("unexpected '**'");
            if (eat('*'))
                l = g.mul(l, unary());
            else if (eat('/'))
                l = g.div(l, unary());
            else
                return l;
        }
    }
    int unary()
    {
        if (eat('-'))
            return g.neg(unary());
        if (eat('+'))
            return unary();
        return power();
    }
    int power()
    {
        int b = primary();
        ws();
        if (i < s.size() && s[i] == '^') {
            ++i;
            return g.pow(b, unary());
        }
        if (i + 1 < s.size() && s[i] == '*' && s[i + 1] == '*') {
            i += 2;
            return g.pow(b, unary());
        }
        return b;
    }
    std::vector<int> args()
    {
        std::vector<int> a;
        if (!eat('('))
            fail("expected '('");
        if (eat(')'))
            return a;
        for (;;) {
            a.push_back(expr());
            if (eat(','))
                continue;
            if (eat(')'))
                return a;
            fail("expected ',' or ')'");
        }
    }
    int primary()
    {
        ws();
        if (i >= s.size())
            fail("unexpected end of expression");
        const char c = s[i];
        if (c == '(') {
            ++i;
            const int e = expr();
            if (!eat(')'))
                fail("expected ')'");
            return e;
        }
        if (std::isdigit(static_cast<unsigned char>(c)) ||
            (c == '.' && i + 1 < s.size() && std::isdigit(static_cast<unsigned char>(s[i + 1])))) {
            size_t j = i;
            while (j < s.size() && (std::isdigit(static_cast<unsigned char>(s[j])) || s[j] == '.'))
                ++j;
            if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
                size_t k = j + 1;
                if (k < s.size() && (s[k] == '+' || s[k] == '-'))
                    ++k;
                if (k < s.size() && std::isdigit(static_cast<unsigned char>(s[k]))) {
                    while (k < s.size() && std::isdigit(static_cast<unsigned char>(s[k])))
                        ++k;
                    j = k;
                }
            }
            const double v = std::strtod(s.substr(i, j - i).c_str(), nullptr);
            i = j;
            if (i < s.size() && s[i] == 'L')
                ++i; // R integer literal
            return g.cst(v);
        }
        if (std::isalpha(static_cast<unsigned char>(c)) || c == '.' || c == '_') {
            size_t j = i;
            while (j < s.size() && (std::isalnum(static_cast<unsigned char>(s[j])) || s[j] == '.' || s[j] == '_'))
                ++j;
            const std::string name = s.substr(i, j - i);
            i = j;
            if (peek() == '(')
                return call(name);
            for (size_t k = 0; k < params.size(); ++k)
                if (params[k] == name)
                    return g.param(static_cast<int>(k));
            for (size_t k = 0; k < vars.size(); ++k)
                if (vars[k] == name)
                    return g.var(static_cast<int>(k));
            if (name == "pi")
                return g.cst(3.14159265358979323846);
            fail("unknown symbol '" + name + "' (neither a parameter with a starting value nor a data variable)");
        }
        if (c == '`') {
            const size_t j = s.find('`', i + 1);
            if (j == std::string::npos)
                fail("unterminated backtick name");
            const std::string name = s.substr(i + 1, j - i - 1);
            i = j + 1;
            for (size_t k = 0; k < params.size(); ++k)
                if (params[k] == name)
                    return g.param(static_cast<int>(k));
            for (size_t k = 0; k < vars.size(); ++k)
                if (vars[k] == name)
                    return g.var(static_cast<int>(k));
            fail("unknown symbol '" + name + "'");
        }
        fail(std::string("unexpected character '") + c + "'");
    }
    int call(const std::string &name)
    {
        static const std::map<std::string, int> f1 = {
            {"exp", F_EXP}, {"log", F_LOG}, {"log2", F_LOG2}, {"log10", F_LOG10}, {"log1p", F_LOG1P},
            {"expm1", F_EXPM1}, {"sqrt", F_SQRT}, {"sin", F_SIN}, {"cos", F_COS}, {"tan", F_TAN},
            {"asin", F_ASIN}, {"acos", F_ACOS}, {"atan", F_ATAN}, {"sinh", F_SINH}, {"cosh", F_COSH},
            {"tanh", F_TANH}, {"abs", F_ABS}, {"sign", F_SIGN}, {"pnorm", F_PNORM}, {"dnorm", F_DNORM},
            {"sinpi", F_SINPI}, {"cospi", F_COSPI}};
        std::vector<int> a = args();
        auto need = [&](size_t k) {
            if (a.size() != k)
                fail("function '" + name + "' expects " + std::to_string(k) + " argument(s)");
        };
        auto it = f1.find(name);
        if (it != f1.end()) {
            if (name == "log" && a.size() == 2)
                return g.div(g.func(F_LOG, a[0]), g.func(F_LOG, a[1]));
            need(1);
            return g.func(it->second, a[0]);
        }
        if (name == "I") { need(1); return a[0]; }
        // stats:: self-start model shapes (the formulas only; start values must be supplied)
        auto eexp = [&](int e) { return g.func(F_EXP, e); };
        if (name == "SSasymp") { need(4); return g.add(a[1], g.mul(g.sub(a[2], a[1]), eexp(g.neg(g.mul(eexp(a[3]), a[0]))))); }
        if (name == "SSasympOff") { need(4); return g.mul(a[1], g.sub(g.cst(1.0), eexp(g.neg(g.mul(eexp(a[2]), g.sub(a[0], a[3])))))); }
        if (name == "SSasympOrig") { need(3); return g.mul(a[1], g.sub(g.cst(1.0), eexp(g.neg(g.mul(eexp(a[2]), a[0]))))); }
        if (name == "SSbiexp") { need(5); return g.add(g.mul(a[1], eexp(g.neg(g.mul(eexp(a[2]), a[0])))), g.mul(a[3], eexp(g.neg(g.mul(eexp(a[4]), a[0]))))); }
        if (name == "SSfpl") { need(5); return g.add(a[1], g.div(g.sub(a[2], a[1]), g.add(g.cst(1.0), eexp(g.div(g.sub(a[3], a[0]), a[4]))))); }
        if (name == "SSgompertz") { need(4); return g.mul(a[1], eexp(g.neg(g.mul(a[2], g.pow(a[3], a[0]))))); }
        if (name == "SSlogis") { need(4); return g.div(a[1], g.add(g.cst(1.0), eexp(g.div(g.sub(a[2], a[0]), a[3])))); }
        if (name == "SSmicmen") { need(3); return g.div(g.mul(a[1], a[0]), g.add(a[2], a[0])); }
        if (name == "SSweibull") { need(5); return g.sub(a[1], g.mul(a[2], eexp(g.neg(g.mul(eexp(a[3]), g.pow(a[0], a[4])))))); }
        fail("function '" + name + "' cannot be translated to device code");
    }
};

} // namespace

int parse_rhs(Graph &g, const std::string &text, const std::vector<std::string> &params,
              const std::vector<std::string> &vars)
{
    Parser P{g, text, params, vars};
    g.nparams = static_cast<int>(params.size());
    const int e = P.expr();
    P.ws();
    if (P.i != text.size())
        P.fail("trailing characters");
    return e;
}

// ------------------------------------------------------------------------------------------
// code generation
// ------------------------------------------------------------------------------------------

static std::string lit(double v)
{
    if (std::isnan(v))
        return "NLS_NAN";
    if (std::isinf(v))
        return v > 0 ? "NLS_INF" : "(-NLS_INF)";
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".eEn") == std::string::npos)
        s += ".0";
    if (v < 0)
        s = "(" + s + ")";
    return s;
}

static const char *fn_name(int fid)
{
    switch (fid) {
    case F_EXP: return "NLS_EXP";
    case F_LOG: return "log";
    case F_LOG2: return "log2";
    case F_LOG10: return "log10";
    case F_LOG1P: return "log1p";
    case F_EXPM1: return "expm1";
    case F_SQRT: return "sqrt";
    case F_SIN: return "sin";
    case F_COS: return "cos";
    case F_TAN: return "tan";
    case F_ASIN: return "asin";
    case F_ACOS: return "acos";
    case F_ATAN: return "atan";
    case F_SINH: return "sinh";
    case F_COSH: return "cosh";
    case F_TANH: return "tanh";
    case F_ABS: return "fabs";
    case F_SIGN: return "nls_sign";
    case F_PNORM: return "nls_pnorm";
    case F_DNORM: return "nls_dnorm";
    case F_SINPI: return "nls_sinpi";
    case F_COSPI: return "nls_cospi";
    default: return "nls_unknown";
    }
}

// one statement computing node e, operands named by `name`
static std::string node_rhs(const Graph &g, int e, const std::function<std::string(int)> &name)
{
    const Node &N = g.at(e);
    switch (N.op) {
    case Op::Add: return name(N.a) + " + " + name(N.b);
    case Op::Sub: return name(N.a) + " - " + name(N.b);
    case Op::Mul: return name(N.a) + " * " + name(N.b);
    case Op::Div: return name(N.a) + " / " + name(N.b);
    case Op::Neg: return "-" + name(N.a);
    case Op::Func: return std::string(fn_name(N.b)) + "(" + name(N.a) + ")";
    case Op::Pow: {
        const Node &B = g.at(N.b);
        if (B.op == Op::Const) {
            const double c = B.c;
            if (c == 2.0)
                return name(N.a) + " * " + name(N.a);
            if (c == 0.5)
                return "sqrt(" + name(N.a) + ")";
            if (c == -0.5)
                return "1.0 / sqrt(" + name(N.a) + ")";
            if (c == std::floor(c) && std::fabs(c) <= 64.0)
                return "nls_powi(" + name(N.a) + ", " + std::to_string(static_cast<int>(c)) + ")";
            return "pow(" + name(N.a) + ", " + lit(c) + ")";
        }
        return "pow(" + name(N.a) + ", " + name(N.b) + ")";
    }
    default: return "";
    }
}

static void mark_live(const Graph &g, const std::vector<int> &roots, std::vector<char> &live)
{
    live.assign(g.size(), 0);
    std::function<void(int)> mark = [&](int e) {
        if (live[e])
            return;
        live[e] = 1;
        const Node &N = g.at(e);
        switch (N.op) {
        case Op::Add: case Op::Sub: case Op::Mul: case Op::Div: case Op::Pow: mark(N.a); mark(N.b); break;
        case Op::Neg: case Op::Func: mark(N.a); break;
        default: break;
        }
    };
    for (int r : roots)
        mark(r);
}

static bool is_leaf(const Node &N)
{
    return N.op == Op::Const || N.op == Op::Param || N.op == Op::Var || N.op == Op::Vel;
}

std::string emit_block(const Graph &g, const std::vector<int> &roots, std::vector<std::string> &root_names,
                       const std::string &indent)
{
    std::vector<char> live;
    mark_live(g, roots, live);
    std::function<std::string(int)> name = [&](int e) -> std::string {
        const Node &N = g.at(e);
        switch (N.op) {
        case Op::Const: return lit(N.c);
        case Op::Param: return "th[" + std::to_string(N.a) + "]";
        case Op::Var: return "x[" + std::to_string(N.a) + "]";
        case Op::Vel: return "v[" + std::to_string(N.a) + "]";
        default: return "t" + std::to_string(e);
        }
    };
    std::ostringstream os;
    for (int e = 0; e < static_cast<int>(g.size()); ++e) {
        if (!live[e] || is_leaf(g.at(e)))
            continue;
        os << indent << "const double t" << e << " = " << node_rhs(g, e, name) << ";\n";
    }
    root_names.clear();
    for (int r : roots)
        root_names.push_back(name(r));
    return os.str();
}

// Two-stage emission for the tiled kernel (p > 8), where neither the parameters nor anything
// derived from them should occupy registers across the observation loop.  Every subexpression
// that does not depend on a predictor ("invariant": built from constants, parameters and the
// velocity only) is computed once per launch by `prep` into c[]; a division of a row quantity by
// an invariant becomes a multiplication by the invariant's reciprocal.  `body` then evaluates one
// observation from th[], c[] (both in shared memory) and x[].
SplitCode emit_split(Graph &g, const std::vector<int> &roots_in, const std::string &indent, bool sink_rows)
{
    // ---- pass 1: rewrite a / b (b invariant, a not) -> a * (1 / b) ----
    std::vector<char> live;
    mark_live(g, roots_in, live);
    const int n0 = static_cast<int>(g.size());
    std::vector<int> re(n0, -1);
    std::vector<char> inv0(n0, 0);
    const int one = g.cst(1.0);
    for (int e = 0; e < n0; ++e) {
        if (!live[e])
            continue;
        const Node N = g.at(e); // by value: g grows below
        switch (N.op) {
        case Op::Const: case Op::Param: case Op::Vel: inv0[e] = 1; re[e] = e; break;
        case Op::Var: inv0[e] = 0; re[e] = e; break;
        case Op::Neg: inv0[e] = inv0[N.a]; re[e] = re[N.a] == N.a ? e : g.neg(re[N.a]); break;
        case Op::Func: inv0[e] = inv0[N.a]; re[e] = re[N.a] == N.a ? e : g.func(N.b, re[N.a]); break;
        default: {
            inv0[e] = inv0[N.a] && inv0[N.b];
            const int a = re[N.a], b = re[N.b];
            if (N.op == Op::Div && inv0[N.b] && !inv0[N.a] && !g.is_const(b)) {
                re[e] = g.mul(a, g.div(one, b));
            } else if (a == N.a && b == N.b) {
                re[e] = e;
            } else {
                switch (N.op) {
                case Op::Add: re[e] = g.add(a, b); break;
                case Op::Sub: re[e] = g.sub(a, b); break;
                case Op::Mul: re[e] = g.mul(a, b); break;
                case Op::Div: re[e] = g.div(a, b); break;
                default: re[e] = g.pow(a, b); break;
                }
            }
            break;
        }
        }
    }
    std::vector<int> roots;
    for (int r : roots_in)
        roots.push_back(re[r]);

    // ---- pass 2: classify on the rewritten graph ----
    mark_live(g, roots, live);
    const int n1 = static_cast<int>(g.size());
    std::vector<char> inv(n1, 0), used_by_row(n1, 0);
    for (int e = 0; e < n1; ++e) {
        if (!live[e])
            continue;
        const Node &N = g.at(e);
        switch (N.op) {
        case Op::Const: case Op::Param: case Op::Vel: inv[e] = 1; break;
        case Op::Var: inv[e] = 0; break;
        case Op::Neg: case Op::Func: inv[e] = inv[N.a]; break;
        default: inv[e] = inv[N.a] && inv[N.b]; break;
        }
        if (!inv[e] && !is_leaf(N)) {
            used_by_row[N.a] = 1;
            if (N.op != Op::Neg && N.op != Op::Func)
                used_by_row[N.b] = 1;
        }
    }
    for (int r : roots)
        used_by_row[r] = 1;
    std::vector<int> slot(n1, -1);
    int nc = 0;
    for (int e = 0; e < n1; ++e)
        if (live[e] && inv[e] && !is_leaf(g.at(e)) && used_by_row[e])
            slot[e] = nc++;

    auto leaf_name = [&](const Node &N) -> std::string {
        switch (N.op) {
        case Op::Const: return lit(N.c);
        case Op::Param: return "th[" + std::to_string(N.a) + "]";
        case Op::Var: return "x[" + std::to_string(N.a) + "]";
        default: return "v[" + std::to_string(N.a) + "]";
        }
    };
    std::function<std::string(int)> prep_name = [&](int e) -> std::string {
        const Node &N = g.at(e);
        return is_leaf(N) ? leaf_name(N) : "t" + std::to_string(e);
    };
    std::function<std::string(int)> body_name = [&](int e) -> std::string {
        const Node &N = g.at(e);
        if (is_leaf(N))
            return leaf_name(N);
        if (slot[e] >= 0)
            return "c[" + std::to_string(slot[e]) + "]";
        return "t" + std::to_string(e);
    };
    SplitCode out;
    std::ostringstream prep, body;
    for (int e = 0; e < n1; ++e) {
        if (!live[e] || is_leaf(g.at(e)))
            continue;
        if (inv[e]) {
            prep << indent << "const double t" << e << " = " << node_rhs(g, e, prep_name) << ";\n";
            if (slot[e] >= 0)
                prep << indent << "c[" << slot[e] << "] = t" << e << ";\n";
        } else {
            body << indent << "const double t" << e << " = " << node_rhs(g, e, body_name) << ";\n";
            // hand a Jacobian entry to the caller as soon as it exists (roots[0] is f, roots[1+j] is J_j),
            // so that a long row never has to sit in registers
            if (sink_rows)
                for (size_t r = 1; r < roots.size(); ++r)
                    if (roots[r] == e)
                        body << indent << "J(" << (r - 1) << ", t" << e << ");\n";
        }
    }
    if (sink_rows)
        for (size_t r = 1; r < roots.size(); ++r)
            if (is_leaf(g.at(roots[r])) || inv[roots[r]])
                body << indent << "J(" << (r - 1) << ", " << body_name(roots[r]) << ");\n";
    out.prep = prep.str();
    out.body = body.str();
    out.nconst = nc;
    for (int r : roots)
        out.root_names.push_back(body_name(r));
    return out;
}

std::string generate_model_source(const ModelSpec &spec)
{
    Graph g;
    const int p = static_cast<int>(spec.params.size());
    const int nvar = static_cast<int>(spec.vars.size());
    const int f = parse_rhs(g, spec.rhs, spec.params, spec.vars);

    std::ostringstream os;
    os << "// generated by gslnls_b200 from: " << spec.rhs << "\n";
    os << "#define GSLNLS_P " << p << "\n";
    os << "#define GSLNLS_NVAR " << (nvar > 0 ? nvar : 0) << "\n";
    os << "#define GSLNLS_JAC_MODE " << spec.jac_mode << "\n";
    os << "#define GSLNLS_FVV_MODE " << spec.fvv_mode << "\n";
    os << "#include \"nls_model_prelude.h\"\n\n";

    std::vector<std::string> names;
    {
        const std::string body = emit_block(g, {f}, names, "    ");
        os << "NLS_FN double nls_model_f(const double *th, const double *x)\n{\n"
           << "    (void)th; (void)x;\n"
           << body << "    return " << names[0] << ";\n}\n\n";
    }
    if (spec.jac_mode == 0) {
        std::vector<int> roots{f};
        for (int j = 0; j < p; ++j)
            roots.push_back(g.diff(f, j));
        const std::string body = emit_block(g, roots, names, "    ");
        os << "NLS_FN void nls_model_fj(const double *th, const double *x, double &f, double *J)\n{\n"
           << "    (void)th; (void)x;\n"
           << body << "    f = " << names[0] << ";\n";
        for (int j = 0; j < p; ++j)
            os << "    J[" << j << "] = " << names[j + 1] << ";\n";
        os << "}\n\n";
        // partial derivatives that are literal constants (an additive or linear parameter): the sparse-row path
        // stores and streams no column for them (csrc/sparse.cu)
        if (p <= 30) {
            unsigned mask = 0;
            for (int j = 0; j < p; ++j)
                if (g.is_const(roots[(size_t)j + 1]))
                    mask |= 1u << j;
            os << "#define GSLNLS_JCONST_MASK " << mask << "\n";
            for (int j = 0; j < p; ++j)
                if (mask >> j & 1u) {
                    char buf[64];
                    std::snprintf(buf, sizeof buf, "%a", g.at(roots[(size_t)j + 1]).c);
                    os << "#define GSLNLS_JCONST_" << j << " " << buf << "\n";
                }
        }
        // two-stage form for the tiled kernel: invariants once per launch, rows from th[], c[], x[]
        const SplitCode sc = emit_split(g, roots, "    ", true);
        os << "#define GSLNLS_NC_FJ " << sc.nconst << "\n";
        os << "NLS_FN void nls_model_prep_fj(const double *th, double *c)\n{\n"
           << "    (void)th; (void)c;\n" << sc.prep << "}\n";
        os << "template <class RowSink>\n"
           << "NLS_FN void nls_model_fj_c(const double *th, const double *c, const double *x, double &f, RowSink &J)\n{\n"
           << "    (void)th; (void)x; (void)c;\n"
           << sc.body << "    f = " << sc.root_names[0] << ";\n";
        os << "}\n\n";
    }
    if (spec.fvv_mode == 1) {
        const int d1 = g.ddir(f);
        const int d2 = g.ddir(d1);
        const std::string body = emit_block(g, {d2}, names, "    ");
        os << "NLS_FN double nls_model_fvv(const double *th, const double *v, const double *x)\n{\n"
           << "    (void)th; (void)x; (void)v;\n"
           << body << "    return " << names[0] << ";\n}\n";
        const SplitCode sc = emit_split(g, {d2}, "    ", false);
        os << "#define GSLNLS_NC_FVV " << sc.nconst << "\n";
        os << "NLS_FN void nls_model_prep_fvv(const double *th, const double *v, double *c)\n{\n"
           << "    (void)th; (void)v; (void)c;\n" << sc.prep << "}\n";
        os << "NLS_FN double nls_model_fvv_c(const double *th, const double *v, const double *c, const double *x)\n{\n"
           << "    (void)th; (void)x; (void)v; (void)c;\n"
           << sc.body << "    return " << sc.root_names[0] << ";\n}\n";
    }
    return os.str();
}

} // namespace gslnls
